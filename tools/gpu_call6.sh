#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -x -q 2>&1 | tail -3
for n in 1 2 4; do
  [ $n -le $N ] || continue
  for g in 1 0; do
  echo "== $n GPUs, PF_MG_GRAPH=$g"
  if [ $n -eq 1 ]; then
    PF_MG_GRAPH=$g timeout 300 python tools/newton_bench.py --refine 4 --steps 2 --quiet 2>&1 | cut -c1-330
  else
    PF_MG_GRAPH=$g timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29617 tools/newton_bench.py --refine 4 --steps 2 --quiet 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | cut -c1-330
  fi
  done
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_apply3d -s 3 -c 2 -o gpurun_out/prof_v4_full -f python tools/profile_apply.py --refine 4 --applies 6 > gpurun_out/ncu_full2.log 2>&1; echo "ncu rc=$?"
