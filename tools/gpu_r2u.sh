#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_retrieval2.py > gpurun_out/r2u_prof_retrieval.txt 2>&1
cat gpurun_out/r2u_prof_retrieval.txt
