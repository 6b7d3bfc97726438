#!/usr/bin/env bash
# final state: GPU suite, smoke, default bench line, the 1000-step (sustained, >= 2.5 s) line
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 200 gpurun_out/bench_final.err
timeout 300 python bench.py --steps 1000 --warmup 5 --no-secondary > gpurun_out/bench_final_steps1000.json 2>/dev/null
python - <<'PY'
import json
for f in ('bench_final','bench_final_steps1000'):
    lines=[l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')]
    d=json.loads(lines[-1])
    print(f, 'embed', round(d['value']), d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', round(d['e2e']['value']), round(d['e2e']['fp32_input']['value']), d['clocks'])
    if 'retrieval' in d: print('  retrieval', d['retrieval'].get('ms_per_step'), d['retrieval'].get('error'), 'train', d['train_step'].get('ms_per_step'), d['train_step'].get('error'))
PY
