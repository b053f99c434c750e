#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
B="--steps 30 --warmup 5 --no-cpu-baseline --no-next-rows --no-configs"
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --tb=short 2>&1 | tail -25 > gpurun_out/v6_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v6_smoke.log 2>&1
timeout 300 python bench.py $B > gpurun_out/v6_bench.json 2> gpurun_out/v6_bench.err
GS_PROJECT_V1=1 timeout 300 python bench.py $B > gpurun_out/v6_v1.json 2> gpurun_out/v6_v1.err
timeout 300 python bench.py --scene frustum --P 1000000 --W 512 --H 512 --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows --no-shared-model --no-graph > gpurun_out/v6_frustum.json 2> gpurun_out/v6_frustum.err
echo finished
