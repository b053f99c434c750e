#!/bin/bash
# Round-2 GPU session 1: parity of the new blend kernels, A/B against the round-1 kernels, launch list, ncu of the blends.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -60 > gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2a_bench_new.json 2> gpurun_out/r2a_bench_new.err
GS_BLEND_VARIANT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/r2a_bench_r1.json 2> gpurun_out/r2a_bench_r1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:k_blend -c 4 -f -o gpurun_out/r2a_blend \
    python tools/profile_step.py > gpurun_out/r2a_ncu.log 2>&1
echo finished
