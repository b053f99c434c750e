#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for c in 296 592 888 1184; do
GS_FILL_CTAS=$c timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs --no-graph > gpurun_out/s12_fill_$c.json 2> gpurun_out/s12_fill_$c.err
done
echo finished
