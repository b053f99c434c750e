#!/bin/bash
# Round-2 GPU session 5: binning rework (cooperative walk + hit bitmap), padded merge-path sort, e2e probes, sanitizers.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --tb=short -x 2>&1 | tail -60 > gpurun_out/s5_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs > gpurun_out/s5_bench.json 2> gpurun_out/s5_bench.err
for sk in u c l ucl; do
GS_E2E_SKIP=$sk timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs --no-shared-model > gpurun_out/s5_bench_skip_$sk.json 2> gpurun_out/s5_bench_skip_$sk.err
done
for P in 1000000 2000000; do
timeout 300 python bench.py --scene frustum --P $P --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s5_frustum_${P}.json 2> gpurun_out/s5_frustum_${P}.err
done
timeout 300 python bench.py --scene stress --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/s5_stress.json 2> gpurun_out/s5_stress.err
SEL='test_cuda_matches_cpu_oracle and (micro_1k_64 or stress_4k_96_x6 or opaque_40k or frustum_20k)'
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_parity_gpu.py -q -x -k "$SEL" > gpurun_out/s5_sanitizer_$tool.log 2>&1
done
echo finished
