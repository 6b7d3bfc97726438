#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
CTL_WGRAD_PAIR=0 CTL_TRAIN_GRAPHS=0 timeout 100 python tools/bench_train.py 256 2>&1 | tail -3
echo "--- eager pair on"
CTL_TRAIN_GRAPHS=0 timeout 100 python tools/bench_train.py 256 2>&1 | tail -3
echo "--- ncu pair off"
CTL_WGRAD_PAIR=0 CTL_TRAIN_GRAPHS=0 timeout 240 ncu --clock-control none --metrics gpu__time_duration.sum -c 1100 --csv --log-file gpurun_out/train_launches_k.csv python tools/bench_train.py 256 > /dev/null 2>&1
python tools/ncu_sum.py gpurun_out/train_launches_k.csv | head -14
echo "--- graphs pair off"
CTL_WGRAD_PAIR=0 timeout 100 python tools/bench_train.py 256 2>&1 | tail -3
