#!/usr/bin/env bash
# GPU box, round 2: launch list of one steady-state trunk forward (with DRAM bytes), `--set full` captures of the kernels
# round 2 changed (fused conv3+shortcut launch, residual-in-slab variant, radix-select tau) and of dist_gemm; reports
# exported to text, .ncu-rep files deleted.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
METRICS='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|launch__registers_per_thread|gpu__dram_throughput|lts__t_bytes.sum|sm__throughput|l1tex__data_pipe|smsp__cycles_active.avg|sm__cycles_elapsed.max|lts__throughput|dram__throughput'
CTL_GRAPH=0 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:'conv|stem|maxpool|gap_bn|instnorm' \
    -s 153 -c 51 --csv --log-file gpurun_out/launches_trunk_r2.csv python tools/bench_trunk.py 256 > /dev/null 2>&1
python tools/launchlist.py gpurun_out/launches_trunk_r2.csv gpurun_out/conv_traffic_r2.json > gpurun_out/launches_trunk_r2.txt; tail -3 gpurun_out/launches_trunk_r2.txt
full() {  # name, kernel regex, skip, script args...
  local name=$1 rx=$2 skip=$3; shift 3
  CTL_GRAPH=0 CTL_TRAIN_GRAPHS=0 $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > /dev/null 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py "$METRICS" > gpurun_out/$name.txt
  ncu -i gpurun_out/$name.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Grid Size|Tensor|Achieved Occupancy|L2 Hit|DRAM|Shared Memory" | head -40 >> gpurun_out/$name.txt
  rm -f gpurun_out/$name.ncu-rep
  head -3 gpurun_out/$name.txt
}
full r2_conv_dual_L4b0_ncu_full conv_gemm_pair 161 python tools/bench_trunk.py 256
full r2_conv_residual_L4b1_ncu_full conv_gemm_pair 164 python tools/bench_trunk.py 256
full r2_conv_residual_L3b1_ncu_full conv_gemm_pair 146 python tools/bench_trunk.py 256
full r2_select_tau_ncu_full select_tau 2 python tools/ncu_retrieval.py
full r2_dist_gemm_ncu_full dist_gemm 4 python tools/ncu_retrieval.py
full r2_wgrad_reduce_nchw_ncu_full wgrad_reduce_nchw 200 python tools/bench_train.py 256
ls -la gpurun_out | tail -12
