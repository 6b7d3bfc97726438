#!/bin/bash
# round 2, run Y: count pass in units of 4 gallery tiles per query tile, bucket counters in shared memory
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval_gpu.py -q -x 2>&1 | tail -25 > gpurun_out/r2y_retrieval_tests.txt
timeout 300 python tools/prof_retrieval3.py > gpurun_out/r2y_prof_pass2.txt 2>&1
timeout 300 python tools/prof_retrieval2.py 2>&1 | head -12 > gpurun_out/r2y_prof_retrieval.txt
cat gpurun_out/r2y_retrieval_tests.txt gpurun_out/r2y_prof_pass2.txt gpurun_out/r2y_prof_retrieval.txt
