#!/bin/bash
# What to run on a GPU box (gpurun -- 'bash tools/gpu_round_checks.sh [step ...]'), cheapest first.
# Every step is wrapped in `timeout`; a multi-rank hang must never eat the budget again
# (round 1 lost 53 GPU-minutes to two 300 s tear-down hangs on a 4-GPU box).
# Suggested first call of a round (1 GPU, about 15 minutes of box time):
#   bash tools/gpu_round_checks.sh tests bench variants feed block fp32 chunks miehe2d ncu
# then make the fastest apply variant / V-cycle precision / chunk count the default and re-run `tests bench`;
# `bench_multi` and the 2-rank half of `fp32` need a multi-GPU box (gpurun --gpus 2|4|8).
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
steps=${*:-"tests bench"}
for s in $steps; do
  case $s in
    tests)        # the gating suite
      timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ;;
    forest)       # forest / hanging-node device path (part of `tests` as well)
      timeout 600 python -m pytest tests/test_gpu_forest.py -x -q 2>&1 | tail -30 ;;
    bench)
      timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
      cat gpurun_out/bench_n1.json ;;
    bench_multi)  # only on a multi-GPU box; 120 s cap per rank count
      for n in 2 4 8; do
        [ "$n" -le "$N" ] || continue
        timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
          --master-port 29617 bench.py --gpus $n --steps 50 --warmup 10 2> gpurun_out/bench_n$n.err | tail -1 | tee gpurun_out/bench_n$n.json
      done ;;
    variants)     # needs a tuning build (make -C cracks_b200/csrc TUNING=1).  Measured in round 2 (profiles/r2_variants_ab.json):
                  # every variant below is slower than 16.  A/B of apply-kernel variants: 16 = default, 23 = v4,
                  # 17 / 18 = v4 with __launch_bounds__ (64, 6 / 5): 168 registers (12 / 10 warps/SM, 164 / 256 B spills),
                  # 20 / 21 = v4 with __maxnreg__ (200 / 184): 10 warps/SM, 56 / 176 B spills,
                  # 19 = v5 (y-collapse staged in shared memory: 178 registers, no spills, 54 KB per CTA -> 8 warps/SM)
      for v in 16 20 17 18 21 19; do
        echo "variant $v"; timeout 300 python bench.py --variant $v --steps 50 --warmup 10 --no-cpu-baseline --no-newton --e2e-steps 1 | cut -c1-330
      done ;;
    chunks)       # host-buffer apply (the e2e number): chunks of the H2D / apply / D2H pipeline, 8 was measured (3.75 ms)
      for c in 8 16 32; do
        echo "PF_E2E_CHUNKS=$c"; PF_E2E_CHUNKS=$c timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-newton | cut -c1-400
      done ;;
    miehe2d)      # BASELINE config 2 / 4 at ~2e5 DoF through the C++ command line with the 2-D multigrid (kind 3),
                  # written in round 1 without a GPU; prints wall time and iteration counts
      make -C cracks_b200/host -s && timeout 300 python tools/miehe_scale.py --refine 7 --steps 3 ;;
    fp32)         # A/B of the multigrid V-cycle precision (pf_mg_lowp.cuh, written in round 1 without a GPU):
                  # Newton-its/s and #LinIts with the FP64 and the FP32 V-cycle; same energies expected
      for f in 0 1; do
        echo "PF_MG_FP32=$f"; PF_MG_FP32=$f timeout 180 python tools/newton_bench.py --refine 4 --steps 2 --quiet | cut -c1-600
      done
      if [ "$N" -ge 2 ]; then
        for f in 0 1; do
          echo "PF_MG_FP32=$f, 2 ranks"
          PF_MG_FP32=$f timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
            --master-port 29618 tools/newton_bench.py --refine 4 --steps 2 --quiet | cut -c1-600
        done
      fi ;;
    graph)        # PF_MG_GRAPH=1: V-cycle as a CUDA graph; hung at tear-down with NCCL nodes (2 ranks) in round 1
      PF_MG_GRAPH=1 timeout 120 python tools/newton_bench.py --refine 4 --steps 2 --quiet | cut -c1-300 ;;
    feed)         # needs a tuning build: coefficient feed / CTAs per SM of k_apply3d_v6 (profiles/r2_v6_feed_modes.json)
      for m in 0 1 2 3; do
        echo "PF_V6_MODE=$m"; PF_V6_MODE=$m timeout 200 python bench.py --steps 30 --warmup 5 --no-newton --no-cpu-baseline | cut -c1-330
      done ;;
    block)        # linear solves as u stage + phi stage (default) against one GMRES on the whole system, stage trace
      for b in 1 0; do
        echo "block solve $b"; PF_BLOCK_TRACE=1 timeout 120 python tools/newton_bench.py --refine 4 --steps 2 --quiet --block-solve $b 2>&1 | tail -30 | cut -c1-400
      done
      echo "4-component Krylov basis in the phi stage"; PF_BLOCK_COMPACT=0 timeout 120 python tools/newton_bench.py --refine 4 --steps 2 --quiet | cut -c1-400 ;;
    ncu)          # full capture of the default apply kernel (full grid, deterministic launch index)
      timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply3d -s 3 -c 2 \
        -o gpurun_out/prof_apply -f python tools/profile_apply.py --refine 4 --applies 6 > gpurun_out/ncu.log 2>&1
      ls -la gpurun_out/prof_apply.ncu-rep ;;
    trace)        # per-entry-point host time of the Newton loop (PF_PY_TRACE) and NCCL segment timers (PF_TRACE)
      PF_PY_TRACE=1 timeout 120 python tools/newton_bench.py --refine 4 --steps 2 --quiet | cut -c1-1200 ;;
    *) echo "unknown step $s" ;;
  esac
done
