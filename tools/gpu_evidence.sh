#!/bin/bash
# Evidence session (1 GPU), the source of profiles/r02_*: GPU test suite, sanitizers, launch lists (ours + reference), ncu --set full of one
# config-3 step and one LucidDreamer-shaped step, the default bench line and the reference arm.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,driver_version --format=csv > gpurun_out/f_smi.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -15 > gpurun_out/f_pytest.log
SEL='test_cuda_matches_cpu_oracle and (micro_1k_64 or stress_4k_96_x6 or opaque_40k or frustum_20k)'
for tool in memcheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_parity_gpu.py -q -x -k "$SEL" > gpurun_out/f_sanitizer_$tool.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches_ours.csv python tools/profile_step.py --steps 4 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/f_launches_ref.csv python tools/profile_step.py --steps 4 --ref > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -c 16 -f -o gpurun_out/f_full_cfg3 python tools/profile_step.py --steps 1 > gpurun_out/f_ncu_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -c 16 -f -o gpurun_out/f_full_frustum python tools/profile_step.py --scene frustum --P 1000000 --W 512 --H 512 --seed 2001 --steps 1 > gpurun_out/f_ncu_frustum.log 2>&1
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
echo finished
