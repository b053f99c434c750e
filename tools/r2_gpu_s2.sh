#!/bin/bash
# Round-2 GPU session 2: failing tests in detail, sync-free forward + CUDA-graph step check, headline bench with the graphed e2e.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -k "needles or adam or dropin or reference_" 2>&1 | tail -150 > gpurun_out/s2_pytest_detail.log
timeout 600 python tools/graph_check.py > gpurun_out/s2_graph_check.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
echo finished
