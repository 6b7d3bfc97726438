#!/usr/bin/env bash
# round-2 GPU check E (2 GPUs): GPU suite on one device, then the default bench under torchrun at N=2
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s 2>&1 | grep -vE "^\s*$" > gpurun_out/pytest_gpu_r2e.txt; grep -E "unresolved|train features|passed|failed|FAILED|Error" gpurun_out/pytest_gpu_r2e.txt | cut -c1-400 | tail -30
timeout 600 python bench.py --workload train --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-420
timeout 600 python bench.py --workload retrieval --steps 10 --warmup 3 --no-secondary 2>/dev/null | tail -1 | cut -c1-300
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2e_n2.json 2> gpurun_out/bench_r2e_n2.err; tail -c 1500 gpurun_out/bench_r2e_n2.err; tail -c 5000 gpurun_out/bench_r2e_n2.json
