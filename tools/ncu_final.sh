#!/usr/bin/env bash
# GPU box, end of round 2: `--set full` captures of the kernels changed late in the round (both passes of dist_gemm with the
# new count epilogue, the shared-memory im2col, the two-row bn_bwd_reduce) and the launch list of one training trunk step.
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
METRICS='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|launch__registers_per_thread|gpu__dram_throughput|lts__t_bytes.sum|sm__throughput|l1tex__data_pipe|smsp__cycles_active.avg|sm__cycles_elapsed.max|lts__throughput|dram__throughput'
full() {  # name, kernel regex, skip, script args...
  local name=$1 rx=$2 skip=$3; shift 3
  CTL_GRAPH=0 CTL_TRAIN_GRAPHS=0 timeout 400 $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$name "$@" > /dev/null 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py "$METRICS" > gpurun_out/$name.txt
  ncu -i gpurun_out/$name.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Grid Size|Tensor|Achieved Occupancy|L2 Hit|DRAM|Shared Memory" | head -40 >> gpurun_out/$name.txt
  rm -f gpurun_out/$name.ncu-rep
  head -3 gpurun_out/$name.txt
}
full r2_dist_gemm_pass1_ncu_full dist_gemm 4 python tools/ncu_retrieval.py
full r2_dist_gemm_pass2_ncu_full dist_gemm 5 python tools/ncu_retrieval.py
full r2_stem_im2col_ncu_full stem_im2col 3 python tools/bench_train.py 256
full r2_bn_bwd_reduce_ncu_full bn_bwd_reduce 230 python tools/bench_train.py 256
CTL_TRAIN_GRAPHS=0 timeout 600 $NCU --metrics gpu__time_duration.sum -c 4400 --csv --log-file gpurun_out/r2_train_launches.csv python tools/bench_train.py 256 > /dev/null 2>&1
python tools/ncu_sum.py gpurun_out/r2_train_launches.csv > gpurun_out/r2_train_launches_sum8.txt; head -30 gpurun_out/r2_train_launches_sum8.txt
rm -f gpurun_out/r2_train_launches.csv
