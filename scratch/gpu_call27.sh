#!/bin/bash
# 4 GPUs: multi-rank NCCL parity (2x2x1 blocks: i and j interfaces) and the weak-scaling bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nccl" > gpurun_out/pytest_nccl4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29657 bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n4.log 2>&1
tail -3 gpurun_out/pytest_nccl4.log; tail -1 gpurun_out/bench_n4.log | cut -c1-260
