#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/r2ac_bench_n2.json 2> gpurun_out/r2ac_err.txt
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r2ac_bench_n2.json') if l.startswith('{')]
d=json.loads(lines[-1])
print('embed', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['clocks']['reasons'])
r=d.get('retrieval',{}); t=d.get('train_step',{})
print('retrieval', r.get('ms_per_step'), r.get('sharded_equals_single_gpu'), r.get('error'))
print('train', t.get('ms_per_step'), t.get('value'), t.get('error'))
PY
tail -3 gpurun_out/r2ac_err.txt | cut -c1-300
