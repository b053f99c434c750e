#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
GS_NO_PREFILL=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs --no-graph > gpurun_out/s10_bench_noprefill.json 2> gpurun_out/s10_bench_noprefill.err
timeout 300 python bench.py --scene stress --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model > gpurun_out/s10_stress.json 2> gpurun_out/s10_stress.err
echo finished
