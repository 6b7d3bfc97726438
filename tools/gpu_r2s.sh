#!/bin/bash
# round 2, run S: native trainer handle parity, smem im2col, unrolled bn_bwd_reduce
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/r2s_train_tests.txt
timeout 300 python -m pytest tests/test_retrieval_gpu.py -q -x -s -k market 2>&1 | tail -8 > gpurun_out/r2s_market.txt
timeout 300 python tools/bench_train.py 256 > gpurun_out/r2s_bench_train.txt 2>&1
timeout 300 python bench.py --workload train --steps 30 --warmup 5 > gpurun_out/r2s_bench_train_step.json 2> gpurun_out/r2s_bench_train_step.err
cat gpurun_out/r2s_train_tests.txt gpurun_out/r2s_market.txt gpurun_out/r2s_bench_train.txt; tail -c 1500 gpurun_out/r2s_bench_train_step.json
