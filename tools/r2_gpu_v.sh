#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --tb=short -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/v_pytest.log
timeout 600 python - > gpurun_out/v_shaped.json 2> gpurun_out/v_shaped.err <<'PY'
import json, sys, torch
sys.argv = ["bench.py"]
import bench
dev = torch.device("cuda:0")
print(json.dumps(bench.shaped_legs(bench.Ours, dev, only=("config2_100k_512", "frustum_1M_512_randcam"))))
PY
echo finished
