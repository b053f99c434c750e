#!/bin/bash
# Round-2 GPU session 2: full GPU test suite + compute-sanitizer (memcheck / racecheck / initcheck) on small cases.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -80 > gpurun_out/r2b_pytest.log
SEL='test_cuda_matches_cpu_oracle and (micro_1k_64 or stress_4k_96_x6 or opaque_40k or needles)'
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_parity_gpu.py -q -x -k "$SEL" > gpurun_out/r2b_sanitizer_$tool.log 2>&1
done
timeout 900 ncu --set full --clock-control none --import-source on -c 26 -f -o gpurun_out/r2c_frustum \
    python tools/profile_step.py --scene frustum --P 1000000 --W 512 --H 512 --seed 2001 --steps 2 > gpurun_out/r2c_ncu.log 2>&1
echo finished
