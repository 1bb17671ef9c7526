#!/bin/bash
# 2 GPUs: multi-rank parity with the overlapped halo swap, then A/B of the weak-scaling bench line (overlap on / off)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nccl" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n2_overlap.log 2>&1
F3D_OVERLAP=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29656 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_n2_nooverlap.log 2>&1
tail -4 gpurun_out/pytest_nccl.log; tail -1 gpurun_out/bench_n2_overlap.log | cut -c1-260; tail -1 gpurun_out/bench_n2_nooverlap.log | cut -c1-260
