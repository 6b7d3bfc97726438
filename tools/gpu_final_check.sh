#!/bin/bash
# last check of the round: whole GPU suite, smoke, the training workload line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --workload train --steps 30 --warmup 5 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('train', d['ms_per_step'], d['value'], d['roofline']['frac'])"
timeout 200 python tools/bench_train.py 256 2>&1 | tail -2
