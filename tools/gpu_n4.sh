#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n4.json') if l.startswith('{')][-1])
print('embed', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
r=d.get('retrieval',{}); t=d.get('train_step',{})
print('retrieval', r.get('ms_per_step'), r.get('sharded_equals_single_gpu'), r.get('error'))
print('train', t.get('ms_per_step'), t.get('value'), t.get('error'))
PY
tail -2 gpurun_out/bench_r2_n4.err | cut -c1-200
