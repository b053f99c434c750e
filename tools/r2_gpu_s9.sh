#!/bin/bash
# Round-2 GPU session 9: zero-fill beside the tile pass (gs_backward_prefill).
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x 2>&1 | tail -40 > gpurun_out/s9_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
timeout 300 python bench.py --scene frustum --P 1000000 --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model > gpurun_out/s9_frustum_1000000.json 2> gpurun_out/s9_frustum_1000000.err
timeout 300 python bench.py --scene stress --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s9_stress.json 2> gpurun_out/s9_stress.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_blend_bwd|k_fill_zero|k_grad_vis" -c 3 -f -o gpurun_out/s9_grad_write \
    python tools/profile_step.py --steps 1 > gpurun_out/s9_ncu.log 2>&1
echo finished
