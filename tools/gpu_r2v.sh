#!/bin/bash
# round 2, run V: tile-list passes (positives + threshold subset), index map through the column metadata
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2v_retrieval_tests.txt
timeout 300 python tools/prof_retrieval2.py > gpurun_out/r2v_prof_retrieval.txt 2>&1
timeout 400 python bench.py --workload retrieval --steps 20 --warmup 5 > gpurun_out/r2v_bench_retrieval.json 2> gpurun_out/r2v_bench_retrieval.err
cat gpurun_out/r2v_retrieval_tests.txt gpurun_out/r2v_prof_retrieval.txt; tail -c 2200 gpurun_out/r2v_bench_retrieval.json; tail -c 600 gpurun_out/r2v_bench_retrieval.err
