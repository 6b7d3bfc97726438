#!/usr/bin/env bash
# round-2 GPU check A: trunk parity tests, A/B of the fused shortcut, launch list with DRAM bytes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trunk_gpu.py -x -q 2>&1 | tail -15
timeout 300 python tools/bench_trunk.py 256 2>&1 | tail -2
CTL_FUSE_SHORTCUT=0 timeout 300 python tools/bench_trunk.py 256 2>&1 | tail -2
CTL_GRAPH=0 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    -k regex:'conv|stem|maxpool|gap_bn|instnorm' -s 153 -c 51 --csv --log-file gpurun_out/launches_r2a.csv \
    python tools/bench_trunk.py 256 > /dev/null 2>&1
python tools/launchlist.py gpurun_out/launches_r2a.csv gpurun_out/conv_traffic_r2a.json | tee gpurun_out/launches_r2a.txt | tail -60
