#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply3d_v4 -s 80 -c 2 -o gpurun_out/prof_v4 -f python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-newton --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
