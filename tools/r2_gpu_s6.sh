#!/bin/bash
# Round-2 GPU session 6: staged single H2D per step (e2e), warp-aggregated binning atomics, raster-ordered frustum scene,
# ncu of the binning + mid-sort kernels with source.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x 2>&1 | tail -40 > gpurun_out/s6_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs > gpurun_out/s6_bench.json 2> gpurun_out/s6_bench.err
for P in 1000000 2000000; do
timeout 300 python bench.py --scene frustum --P $P --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s6_frustum_${P}.json 2> gpurun_out/s6_frustum_${P}.err
timeout 300 python bench.py --impl reference --scene frustum --P $P --W 512 --H 512 --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > gpurun_out/s6_frustum_${P}_ref.json 2> gpurun_out/s6_frustum_${P}_ref.err
done
timeout 300 python bench.py --scene frustum --shuffled --P 1000000 --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s6_frustum_1000000_shuffled.json 2> gpurun_out/s6_frustum_shuffled.err
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_count_tiles|k_shade_emit|k_tile_sort_mid|k_tile_sort_big|k_project" -c 5 -f -o gpurun_out/s6_binning_frustum \
    python tools/profile_step.py --scene frustum --P 1000000 --W 512 --H 512 --seed 2001 --steps 1 > gpurun_out/s6_ncu_binning.log 2>&1
echo finished
