#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_trunk_gpu.py -q -x -k "u8 or native_trunk" 2>&1 | tail -4 > gpurun_out/r2ae.txt
timeout 400 python bench.py --no-secondary > gpurun_out/r2ae_bench.json 2> gpurun_out/r2ae_err.txt
cat gpurun_out/r2ae.txt; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ae_bench.json') if l.startswith('{')][-1])
print('embed', round(d['value']), d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', round(d['e2e']['value']), round(d['e2e']['fp32_input']['value']), d['clocks'])
PY
tail -2 gpurun_out/r2ae_err.txt | cut -c1-300
