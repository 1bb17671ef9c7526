#!/bin/bash
# 2-GPU check of generation 3: multi-rank NCCL parity test and the weak-scaling bench line at N=2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nccl" > gpurun_out/pytest_nccl.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2.log 2>&1
tail -4 gpurun_out/pytest_nccl.log; tail -2 gpurun_out/bench_n2.log | cut -c1-700
