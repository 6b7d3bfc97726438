#!/usr/bin/env bash
# round-2 GPU check C: whole GPU suite (no -x), default bench, train workload
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$" > gpurun_out/pytest_gpu_r2c.txt; grep -E "engine vs|train features|train-mode features|worst gradient|fp16-sim|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r2c.txt | cut -c1-300 | tail -40
timeout 900 python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 400 gpurun_out/bench_r2c.err; tail -c 2500 gpurun_out/bench_r2c.json
