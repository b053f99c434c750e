#!/bin/bash
# Round-2 GPU session 4: TMA pipelines (k_grad_dense, k_shade_emit), small-rect binning, merge-path sort, blend fixes.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -120 > gpurun_out/s4_pytest.log
timeout 600 python tools/graph_check.py > gpurun_out/s4_graph_check.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
for P in 1000000 2000000; do
timeout 300 python bench.py --scene frustum --P $P --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model > gpurun_out/s4_frustum_${P}.json 2> gpurun_out/s4_frustum_${P}.err
done
GS_NO_TMA=1 timeout 300 python bench.py --scene frustum --P 1000000 --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s4_frustum_1000000_notma.json 2> gpurun_out/s4_frustum_notma.err
timeout 300 python bench.py --scene stress --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s4_stress.json 2> gpurun_out/s4_stress.err
timeout 900 ncu --set full --clock-control none --import-source on -c 14 -f -o gpurun_out/s4_full_frustum \
    python tools/profile_step.py --scene frustum --P 1000000 --W 512 --H 512 --seed 2001 --steps 1 > gpurun_out/s4_ncu_frustum.log 2>&1
echo finished
