#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval_gpu.py -q -x 2>&1 | tail -3 > gpurun_out/r2aa.txt
timeout 300 python tools/prof_retrieval4.py >> gpurun_out/r2aa.txt 2>&1
timeout 300 python tools/prof_retrieval2.py 2>&1 | head -7 >> gpurun_out/r2aa.txt
cat gpurun_out/r2aa.txt
