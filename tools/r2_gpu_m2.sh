#!/bin/bash
# Round-2 multi-GPU session (2 GPUs): fused peer reduce with folded barriers vs NCCL, config 4 / config 5 legs, scaling line.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/m${N}_dp_check.log 2>&1
DP_P=1000000 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/m${N}_dp_check_1M.log 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo finished
