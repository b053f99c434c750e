#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="--steps 30 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs"
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -25 > gpurun_out/v3_pytest.log
if ! grep -q " passed" gpurun_out/v3_pytest.log || grep -q "failed" gpurun_out/v3_pytest.log; then
  GS_PROJECT_V1=1 GS_NO_EARLY_PREFILL=1 GS_SCAN_KERNEL=1 timeout 600 python -m pytest tests -m gpu -q --timeout 600 --tb=line 2>&1 | tail -8 > gpurun_out/v3_pytest_fallback.log
  GS_NO_EARLY_PREFILL=1 GS_SCAN_KERNEL=1 timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 --tb=line 2>&1 | tail -8 > gpurun_out/v3_pytest_projonly.log
fi
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v3_smoke.log 2>&1
timeout 300 python bench.py $B > gpurun_out/v3_new.json 2> gpurun_out/v3_new.err
GS_PROJECT_V1=1 timeout 300 python bench.py $B > gpurun_out/v3_projv1.json 2> gpurun_out/v3_projv1.err
GS_NO_EARLY_PREFILL=1 timeout 300 python bench.py $B > gpurun_out/v3_noearly.json 2> gpurun_out/v3_noearly.err
GS_SCAN_KERNEL=1 timeout 300 python bench.py $B > gpurun_out/v3_scank.json 2> gpurun_out/v3_scank.err
timeout 300 python bench.py --scene frustum --P 1000000 --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/v3_frustum.json 2> gpurun_out/v3_frustum.err
echo finished
