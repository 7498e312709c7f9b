#!/usr/bin/env python
"""Minimal driver for ncu: the operator-benchmark state of bench.py (Sneddon-3D, refine 4) and
N full-grid applies of the exact operator, nothing else (Jacobi preconditioner, so the only
k_apply3d launches are the ones of interest).

  ncu --set full --clock-control none --import-source on -k regex:k_apply3d -s 3 -c 2 \
      -o gpurun_out/prof python tools/profile_apply.py --refine 4 --applies 6
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refine", type=int, default=4)
    ap.add_argument("--applies", type=int, default=6)
    ap.add_argument("--bits", type=int, default=64, help="precision of the Krylov operator (pf_set_jacobian_precision)")
    ap.add_argument("--variant", type=int, default=0, help="pf_debug_set_variant (23 = v4)")
    args = ap.parse_args()
    import cracks_b200 as pf
    from bench import sneddon_state, SEED
    mesh = pf.sneddon_mesh(3, args.refine)
    ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh))
    ctx.set_preconditioner(0, 2, 20.0)
    if args.variant:
        ctx._check(ctx.lib.pf_debug_set_variant(ctx.h, args.variant))
    if args.bits != 64:
        ctx.set_jacobian_precision(args.bits)
    sol, active = sneddon_state(mesh.n[0], mesh.h[0])
    ctx.set_state(sol, sol, sol, 1.0, 1.0, False, 1e-3)
    ctx.set_dirichlet_all_faces()
    ctx.set_constraints(None, active)
    ctx.setup_jacobian()
    x = np.random.default_rng(SEED).standard_normal(ctx.n_dofs)
    x_dev, y_dev = ctx.device_vector(), ctx.device_vector()
    ctx.upload(x, x_dev)
    for _ in range(args.applies):
        ctx.vmult_dev(y_dev, x_dev)
    ctx.synchronize()
    print("applies done:", args.applies, "launches:", ctx.launch_count)
    ctx.close()


if __name__ == "__main__":
    main()
