#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="--steps 30 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs"
GS_PROJECT_CULL=1 timeout 600 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -25 > gpurun_out/v4_pytest_cull.log
GS_PROJECT_CULL=1 timeout 300 python bench.py $B > gpurun_out/v4_cull.json 2> gpurun_out/v4_cull.err
timeout 300 python bench.py $B > gpurun_out/v4_v1.json 2> gpurun_out/v4_v1.err
echo finished
