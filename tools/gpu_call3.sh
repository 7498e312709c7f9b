#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
N=$(nvidia-smi -L | wc -l)
export PF_PY_TRACE=1
for n in 2 4 8; do
  [ $n -le $N ] || continue
  echo "== $n GPUs"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29617 tools/newton_bench.py --refine 4 --steps 2 --quiet 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | cut -c1-500
done
