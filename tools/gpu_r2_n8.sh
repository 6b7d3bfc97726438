#!/usr/bin/env bash
# 8 x B200: the default bench line under torchrun (embedding scaling + nested config-5 sharded retrieval + config-4 training)
set -u
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err
tail -c 1200 gpurun_out/bench_r2_n8.err; tail -c 6000 gpurun_out/bench_r2_n8.json
