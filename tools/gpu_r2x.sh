#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_retrieval3.py > gpurun_out/r2x_prof_pass2.txt 2>&1
cat gpurun_out/r2x_prof_pass2.txt
