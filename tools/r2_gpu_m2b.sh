#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-2}
DP_P=1000000 DP_W=1920 DP_H=1080 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/dp_check.py > gpurun_out/m${N}b_dp_check_1M.log 2>&1
echo finished
