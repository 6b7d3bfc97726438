#!/usr/bin/env bash
set -u
timeout 400 python -m pytest tests/test_train_gpu.py tests/test_reference_autocast_gpu.py -m gpu -q 2>&1 | tail -5
for i in 1 2 3; do timeout 50 python -u tools/bench_train.py 256 2>&1 | grep "train trunk" | cut -c1-120; echo "rc=${PIPESTATUS[0]}"; done
CTL_FUSE_BN_STATS=0 timeout 50 python -u tools/bench_train.py 256 2>&1 | grep "train trunk" | cut -c1-120
for i in 1 2; do timeout 100 python -u bench.py --workload train --steps 20 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c150-260; echo "rc=${PIPESTATUS[0]}"; done
