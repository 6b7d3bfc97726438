#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$" > gpurun_out/pytest_gpu_r2f.txt; grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r2f.txt | cut -c1-300 | tail -30
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-420
CTL_DYNAMIC_LOSS_SCALE=0 timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-420
timeout 600 python bench.py --workload retrieval --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-300
