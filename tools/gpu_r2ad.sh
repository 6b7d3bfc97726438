#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_retrieval_gpu.py -q -x 2>&1 | tail -5 > gpurun_out/r2ad.txt
timeout 400 python bench.py --workload retrieval --steps 20 --warmup 5 > gpurun_out/r2ad_bench_retrieval.json 2> gpurun_out/r2ad_err.txt
cat gpurun_out/r2ad.txt; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2ad_bench_retrieval.json') if l.startswith('{')][-1])
print('retrieval', d['ms_per_step'], d['value'], 'e2e', d['e2e']['value'], 'mAP', d['mAP'], d['roofline'].get('pass_ms'), d['roofline'].get('pass2_ms'))
PY
tail -3 gpurun_out/r2ad_err.txt | cut -c1-300
