#!/usr/bin/env bash
# GPU box: (1) launch list of one steady-state trunk forward (our kernels only);
# (2) `--set full` captures of three representative conv launches and one dist_gemm launch, exported
#     to CSV/text on the box (the .ncu-rep files stay small enough to travel; gpurun_out <= 64 MiB).
set -u
mkdir -p gpurun_out
NCU="ncu --clock-control none"
KERN='conv|stem|maxpool|gap_bn|instnorm'
CTL_GRAPH=0 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:"$KERN" -s 165 -c 55 --csv \
    --log-file gpurun_out/launches_trunk.csv python tools/bench_trunk.py 256 > gpurun_out/ncu_trunk_run.log 2>&1
echo "launch list exit $?"
METRICS='dram__bytes_read.sum|dram__bytes_write.sum|gpu__time_duration.sum|sm__pipe_tensor_cycles_active|sm__inst_executed_pipe_tensor|sm__warps_active.avg.pct|launch__registers_per_thread|gpu__dram_throughput|lts__t_bytes.sum|lts__t_sectors_srcunit_tex|sm__throughput|l1tex__data_pipe|smsp__cycles_active.avg|sm__cycles_elapsed.max|lts__throughput|dram__throughput'
i=0
for skip in 157 159 199; do
  i=$((i+1))
  CTL_GRAPH=0 $NCU --set full --import-source on -k regex:conv -s $skip -c 1 -f -o gpurun_out/prof_conv_$i python tools/bench_trunk.py 256 > /dev/null 2>&1
  ncu -i gpurun_out/prof_conv_$i.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py "$METRICS" > gpurun_out/prof_conv_$i.txt
  ncu -i gpurun_out/prof_conv_$i.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Grid Size|Tensor|Achieved Occupancy|L2 Hit|DRAM" | head -40 >> gpurun_out/prof_conv_$i.txt
  rm -f gpurun_out/prof_conv_$i.ncu-rep
done
CTL_GRAPH=0 $NCU --set full --import-source on -k regex:stem_pool -s 3 -c 1 -f -o gpurun_out/prof_stem python tools/bench_trunk.py 256 > /dev/null 2>&1
ncu -i gpurun_out/prof_stem.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Grid Size|Achieved Occupancy|L2 Hit|DRAM|Pipe|Issue" | head -40 > gpurun_out/prof_stem.txt
rm -f gpurun_out/prof_stem.ncu-rep
$NCU --set full --import-source on -k regex:dist_gemm -s 2 -c 1 -f -o gpurun_out/prof_dist python tools/ncu_retrieval.py > /dev/null 2>&1
ncu -i gpurun_out/prof_dist.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py "$METRICS" > gpurun_out/prof_dist.txt
ncu -i gpurun_out/prof_dist.ncu-rep --page details 2>/dev/null | grep -E "Duration|Throughput|Registers|Grid Size|Tensor|Achieved Occupancy|L2 Hit|DRAM" | head -40 >> gpurun_out/prof_dist.txt
ls -la gpurun_out
