#!/bin/bash
# Round-2 GPU session 3: new blend kernels (strip geometry, a_eff fast path), k_grad_dense (TMA), static capacity + graphs.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -120 > gpurun_out/s3_pytest.log
timeout 600 python tools/graph_check.py > gpurun_out/s3_graph_check.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
for P in 1000000 2000000; do
timeout 300 python bench.py --scene frustum --P $P --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s3_frustum_${P}.json 2> gpurun_out/s3_frustum_${P}.err
done
GS_BLEND_VARIANT=2 timeout 300 python bench.py --scene frustum --P 1000000 --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s3_frustum_1000000_blocks.json 2> gpurun_out/s3_frustum_blocks.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_blend|k_grad_dense" -c 6 -f -o gpurun_out/s3_blend_cfg3 \
    python tools/profile_step.py --steps 2 > gpurun_out/s3_ncu_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_blend|k_grad_dense" -c 6 -f -o gpurun_out/s3_blend_frustum \
    python tools/profile_step.py --scene frustum --P 1000000 --W 512 --H 512 --seed 2001 --steps 2 > gpurun_out/s3_ncu_frustum.log 2>&1
echo finished
