#!/bin/bash
mkdir -p gpurun_out
for v in caller sorted_all sorted_list; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_launches_$v.csv python tools/ncu_retrieval2.py $v > /dev/null 2>&1
  echo "== $v"; python tools/ncu_sum.py gpurun_out/r2w_launches_$v.csv | head -24
done > gpurun_out/r2w_summary.txt 2>&1
cat gpurun_out/r2w_summary.txt
