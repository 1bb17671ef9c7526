#!/bin/bash
# 8 GPUs: the weak-scaling bench line (2x2x2 blocks: i, j and k interfaces over NCCL); nothing else, to keep the 8x charge short
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29658 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n8.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/bench_n8.log | cut -c1-300
