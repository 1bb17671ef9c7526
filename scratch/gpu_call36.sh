#!/bin/bash
# 2 GPUs, final kernel set: multi-rank NCCL parity and the weak-scaling bench line
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nccl" > gpurun_out/pytest_nccl2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl2.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29659 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n2_final.log 2>&1
tail -3 gpurun_out/pytest_nccl2.log; tail -1 gpurun_out/bench_n2_final.log | cut -c1-260
