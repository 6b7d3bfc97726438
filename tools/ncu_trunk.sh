#!/usr/bin/env bash
# GPU box: launch list of one steady-state trunk forward (our kernels only) + full captures.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 3 warm-up forwards x 56 launches = 168 of our kernels skipped, then one forward (56 launches)
$NCU --metrics gpu__time_duration.sum -k regex:'conv_gemm|stem_conv|maxpool|gap_bn|instnorm' -s 168 -c 56 --csv \
    --log-file gpurun_out/launches_trunk.csv python tools/bench_trunk.py 256 > gpurun_out/ncu_trunk_run.log 2>&1
echo "launch list exit $?"
# full captures: a 3x3 conv of layer1 (launch 4 of a forward = id 170 overall), one of layer4
$NCU --set full --import-source on -k regex:conv_gemm -s 156 -c 52 -o gpurun_out/prof_conv python tools/bench_trunk.py 256 > gpurun_out/ncu_conv_full.log 2>&1
echo "conv full exit $?"
