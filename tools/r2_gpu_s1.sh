#!/bin/bash
# Round-2 GPU session (re-entry): parity suite, A/B of blend variants, reference arm, frustum scene, launch list + ncu full.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/s1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/s1_bench_new.json 2> gpurun_out/s1_bench_new.err
GS_BLEND_VARIANT=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model > gpurun_out/s1_bench_r1.json 2> gpurun_out/s1_bench_r1.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > gpurun_out/s1_bench_ref.json 2> gpurun_out/s1_bench_ref.err
for P in 1000000 2000000; do
timeout 300 python bench.py --scene frustum --P $P --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model > gpurun_out/s1_frustum_${P}.json 2> gpurun_out/s1_frustum_${P}.err
timeout 300 python bench.py --impl reference --scene frustum --P $P --W 512 --H 512 --steps 10 --warmup 3 --no-cpu-baseline --no-next-rows > gpurun_out/s1_frustum_${P}_ref.json 2> gpurun_out/s1_frustum_${P}_ref.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/s1_launches_ours.csv python tools/profile_step.py --steps 4 > gpurun_out/s1_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -c 13 -f -o gpurun_out/s1_full_cfg3 \
    python tools/profile_step.py --steps 1 > gpurun_out/s1_ncu_cfg3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -c 13 -f -o gpurun_out/s1_full_frustum \
    python tools/profile_step.py --scene frustum --P 1000000 --W 512 --H 512 --seed 2001 --steps 1 > gpurun_out/s1_ncu_frustum.log 2>&1
echo finished
