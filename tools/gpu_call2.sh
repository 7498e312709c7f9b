#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for r in 4 3; do
  timeout 300 python tools/newton_bench.py --refine $r --steps 2 --quiet > gpurun_out/nb_r${r}_n1.json 2>&1
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 tools/newton_bench.py --refine $r --steps 2 --quiet > gpurun_out/nb_r${r}_n2.json 2>&1
done
tail -n 2 gpurun_out/nb_r*.json
