#!/bin/bash
# Round-2 multi-GPU session (N GPUs): fused peer reduce vs NCCL with per-kernel times, then the full bench line (view-parallel
# headline, shared-model step, BASELINE configs 4 and 5 at size).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-8}
DP_P=1000000 DP_W=1920 DP_H=1080 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/m${N}_dp_check_1M.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo finished
