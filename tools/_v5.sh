#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="--steps 30 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs"
GS_PROJECT_CULL=1 timeout 600 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -25 > gpurun_out/v5_pytest_cull.log
GS_PROJECT_CULL=1 timeout 300 python bench.py $B > gpurun_out/v5_cull.json 2> gpurun_out/v5_cull.err
GS_PROJECT_CULL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_project --launch-skip 8 -c 1 -o gpurun_out/v5_kproject python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-next-rows --no-configs --no-graph --no-shared-model > gpurun_out/v5_ncu.log 2>&1
echo finished
