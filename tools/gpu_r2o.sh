#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do timeout 50 python -u tools/bench_train.py 256 2>&1 | grep "train trunk" | cut -c1-120; echo "rc=${PIPESTATUS[0]}"; done
for i in 1 2 3; do timeout 100 python -u bench.py --workload train --steps 20 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c150-260; echo "rc=${PIPESTATUS[0]}"; done
timeout 700 python -m pytest tests -m gpu -q 2>&1 | tail -4
