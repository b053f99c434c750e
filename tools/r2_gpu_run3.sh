#!/bin/bash
# Round-2 GPU session 3: ncu --set full of every kernel of one LucidDreamer-shaped step (1 M Gaussians, 512x512, all visible)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -c 26 -f -o gpurun_out/r2c_frustum \
    python tools/profile_step.py --scene frustum --P 1000000 --W 512 --H 512 --seed 2001 --steps 2 > gpurun_out/r2c_ncu.log 2>&1
echo finished
