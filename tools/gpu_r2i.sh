#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; tail -c 300 gpurun_out/bench_r2i.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/bench_r2i.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('embed', d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['fp32_input']['value'], d['clocks'])
print('retrieval', d['retrieval'].get('ms_per_step'), d['retrieval'].get('value'), d['retrieval'].get('error'))
print('train', d['train_step'].get('ms_per_step'), d['train_step'].get('error'))
print('cpu', d['cpu_baseline'], d['retrieval'].get('cpu_baseline'), d['train_step'].get('cpu_baseline'))
PY
timeout 300 python tools/bench_train.py 256 2>&1 | tail -2
