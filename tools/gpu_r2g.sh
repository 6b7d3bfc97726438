#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_modules_gpu.py tests/test_solver_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/bench_train.py 256 2>&1 | tail -2
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-420
CTL_DYNAMIC_LOSS_SCALE=0 timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-420
CTL_TRAIN_GRAPHS=0 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/train_launches_r2g.csv python tools/bench_train.py 256 > /dev/null 2>&1
python tools/ncu_sum.py gpurun_out/train_launches_r2g.csv > gpurun_out/train_launches_r2g_sum_all.txt; head -32 gpurun_out/train_launches_r2g_sum_all.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('embed', d['value'], 'e2e', d['e2e']['value'], d['e2e']['fp32_input']['value'], 'frac', d['roofline']['frac'])"
