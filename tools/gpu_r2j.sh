#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -q 2>&1 | tail -6
timeout 300 python tools/bench_train.py 256 2>&1 | tail -2
CTL_WGRAD_PAIR=0 timeout 300 python tools/bench_train.py 256 2>&1 | tail -2
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-420
CTL_WGRAD_PAIR=0 timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-420
timeout 300 python tools/bench_train.py 256 2>&1 | tail -2
