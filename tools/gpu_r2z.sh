#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/r2z_retrieval_tests.txt
{
  timeout 300 python tools/prof_retrieval3.py 2>&1 | cut -c1-330
  timeout 300 python tools/prof_retrieval2.py 2>&1 | head -7
} > gpurun_out/r2z_prof.txt 2>&1
cat gpurun_out/r2z_retrieval_tests.txt gpurun_out/r2z_prof.txt
