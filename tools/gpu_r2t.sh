#!/bin/bash
# round 2, run T: cheap-tile retrieval passes (approx / skip / pid-sorted planes), native trainer IBN parity
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2t_retrieval_tests.txt
timeout 300 python -m pytest tests/test_train_gpu.py -q -x -k "native" 2>&1 | tail -8 > gpurun_out/r2t_native_trainer.txt
timeout 300 python -m pytest tests/test_modules_gpu.py tests/test_losses_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/r2t_modules.txt
timeout 400 python bench.py --workload retrieval --steps 20 --warmup 5 > gpurun_out/r2t_bench_retrieval.json 2> gpurun_out/r2t_bench_retrieval.err
cat gpurun_out/r2t_retrieval_tests.txt gpurun_out/r2t_native_trainer.txt gpurun_out/r2t_modules.txt; tail -c 2500 gpurun_out/r2t_bench_retrieval.json; tail -5 gpurun_out/r2t_bench_retrieval.err
