#!/bin/bash
# multi-GPU call: NCCL parity (1 and 2 blocks per rank), weak and strong scaling bench lines at N = number of visible GPUs
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests -m gpu -x -q -k "nccl or multiblock or interface or periodic or four_blocks or checkpoint or smoothbump_two" > gpurun_out/m${N}_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/m${N}_pytest.txt; tail -5 gpurun_out/m${N}_pytest.txt
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu ${@:3} > gpurun_out/m${N}_$2.json 2> gpurun_out/m${N}_$2.err; cut -c1-260 gpurun_out/m${N}_$2.json; tail -2 gpurun_out/m${N}_$2.err | cut -c1-300; }
run 29701 weak
run 29702 strong --scaling strong
if [ "$N" = "2" ]; then NCCL_DEBUG=INFO run 29703 weak_ncclinfo; grep -m3 -E "NVLS|via P2P|Channel 00" gpurun_out/m${N}_weak_ncclinfo.err | cut -c1-200; fi
