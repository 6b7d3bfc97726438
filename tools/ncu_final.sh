#!/usr/bin/env bash
# GPU box: launch list of one steady-state trunk forward (with DRAM bytes) + `--set full` of the CTA-pair conv kernel
# (layer4 3x3) and of the weight-gradient kernel; reports exported to text, .ncu-rep files deleted.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
METRICS='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|launch__registers_per_thread|gpu__dram_throughput|lts__t_bytes.sum|sm__throughput|l1tex__data_pipe|smsp__cycles_active.avg|sm__cycles_elapsed.max|lts__throughput|dram__throughput'
CTL_GRAPH=0 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:'conv|stem|maxpool|gap_bn|instnorm' \
    -s 165 -c 55 --csv --log-file gpurun_out/launches_trunk.csv python tools/bench_trunk.py 256 > /dev/null 2>&1
echo "launch list exit $?"
CTL_GRAPH=0 $NCU --set full --import-source on -k regex:conv_gemm_pair -s 186 -c 1 -f -o gpurun_out/prof_pair python tools/bench_trunk.py 256 > /dev/null 2>&1
ncu -i gpurun_out/prof_pair.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py "$METRICS" > gpurun_out/prof_pair.txt
ncu -i gpurun_out/prof_pair.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Grid Size|Tensor|Achieved Occupancy|L2 Hit|DRAM|Shared Memory" | head -40 >> gpurun_out/prof_pair.txt
rm -f gpurun_out/prof_pair.ncu-rep
CTL_TRAIN_GRAPHS=0 $NCU --set full --import-source on -k regex:conv_wgrad -s 140 -c 1 -f -o gpurun_out/prof_wgrad python tools/bench_train.py 256 > /dev/null 2>&1
ncu -i gpurun_out/prof_wgrad.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py "$METRICS" > gpurun_out/prof_wgrad.txt
ncu -i gpurun_out/prof_wgrad.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Grid Size|Tensor|Achieved Occupancy|L2 Hit|DRAM|Shared Memory" | head -40 >> gpurun_out/prof_wgrad.txt
rm -f gpurun_out/prof_wgrad.ncu-rep
ls -la gpurun_out
