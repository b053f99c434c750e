#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for combo in "12 10" "12 13" "12 12" "14 13" "16 13"; do
set -- $combo
GS_BLEND_OCC_FWD=$1 GS_BLEND_OCC_BWD=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs --no-graph > gpurun_out/occ_$1_$2.json 2> gpurun_out/occ_$1_$2.err
done
echo finished
