#!/usr/bin/env bash
# round-2 GPU check B: the whole GPU suite, smoke, the default bench line, dist_gemm DRAM traffic
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -vE "^\s*$" | tail -40 > gpurun_out/pytest_gpu_r2b.txt; tail -25 gpurun_out/pytest_gpu_r2b.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; tail -c 600 gpurun_out/bench_r2b.err; head -c 3000 gpurun_out/bench_r2b.json
timeout 300 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:dist_gemm \
   --csv --log-file gpurun_out/dist_launches.csv python tools/ncu_retrieval.py > /dev/null 2>&1
python tools/dist_traffic.py gpurun_out/dist_launches.csv gpurun_out/dist_traffic.json
