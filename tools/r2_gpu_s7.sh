#!/bin/bash
# Round-2 GPU session 7: clear kernel instead of memset nodes (graph e2e), walk fix, wider sort CTAs, TMA zero-fill in k_grad_write.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x 2>&1 | tail -40 > gpurun_out/s7_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs > gpurun_out/s7_bench.json 2> gpurun_out/s7_bench.err
for P in 1000000 2000000; do
timeout 300 python bench.py --scene frustum --P $P --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model > gpurun_out/s7_frustum_${P}.json 2> gpurun_out/s7_frustum_${P}.err
done
timeout 300 python bench.py --scene stress --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s7_stress.json 2> gpurun_out/s7_stress.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_grad_write|k_shade_emit|k_count_tiles" -c 3 -f -o gpurun_out/s7_cfg3_misc \
    python tools/profile_step.py --steps 1 > gpurun_out/s7_ncu_misc.log 2>&1
echo finished
