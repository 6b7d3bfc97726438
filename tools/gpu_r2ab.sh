#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload retrieval --steps 5 --warmup 3 > gpurun_out/r2ab_bench_retrieval_n2.json 2> gpurun_out/r2ab_err.txt
tail -c 1800 gpurun_out/r2ab_bench_retrieval_n2.json; tail -5 gpurun_out/r2ab_err.txt | cut -c1-300
