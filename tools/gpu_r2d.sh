#!/usr/bin/env bash
# round-2 GPU check D: GPU suite, bench (train leg), launch list of one retrieval step
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$" > gpurun_out/pytest_gpu_r2d.txt; grep -E "unresolved|train features|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r2d.txt | cut -c1-400 | tail -30
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-900
timeout 600 python bench.py --workload retrieval --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-700
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/ret_launches.csv python tools/ncu_retrieval.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[l for l in open('gpurun_out/ret_launches.csv') if not l.startswith('==')]
r=list(csv.DictReader(rows))
n=len(r)
per=n//3 if n>=3 else n
last=r[-per:]
tot=0
for x in last:
    v=float(x['Metric Value'].replace(',','')); u=x['Metric Unit']
    us=v/1e3 if u in('nsecond','ns') else v
    tot+=us
    print(f"{x['Kernel Name'][:70]:70s} {us:9.1f} us")
print('total', tot)
PY
