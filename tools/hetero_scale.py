#!/usr/bin/env python
"""BASELINE config 5 (`multiple het`, parameters_hetero_3d.prm) at scale: the forest operator apply on N GPUs.

  python tools/hetero_scale.py --global-refine 7 --local-refine 1                      # one GPU
  torchrun --nproc-per-node 8 ... tools/hetero_scale.py --global-refine 8      # 6.8e7 DoF on 8 GPUs

The single-tree cube [0,10]^3 is refined `--global-refine` times, then `--local` times where the interpolated initial
cracks lie (ref strategy = phase field, cracks.cc:3971-3995), Lame coefficients per cell from a heterogeneous E-modulus
field (synthetic smooth field in [E, 10 E]: the reference's test.pgm is not shipped), u = 0 on the six faces.  Every rank
builds the same host forest; cells are cut into N contiguous ranges (pf_create_forest_distributed).  Timed: y = J(U) x,
one residual evaluation, one diagonal.  |y| is printed as a check that does not depend on N."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-refine", type=int, default=6)
    ap.add_argument("--local-refine", type=int, default=1)
    ap.add_argument("--applies", type=int, default=10)
    args = ap.parse_args()
    import torch
    import cracks_b200 as pf
    from cracks_b200 import api
    from cracks_b200.forest import ForestContext, HostForest, initial_multiple_het_3d

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as td
        td.init_process_group("nccl", device_id=torch.device("cuda", local))

        def fresh_id():
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.frombuffer(bytearray(pf.PhaseFieldContext.nccl_unique_id()), dtype=torch.uint8).cuda()
            td.broadcast(idt, 0)
            return idt.cpu().numpy().tobytes()
        dist = (rank, world, fresh_id)

    t0 = time.perf_counter()
    f = HostForest(3, (1, 1, 1), (0.0,) * 3, (10.0,) * 3)
    f.refine_global(args.global_refine)
    cap = args.global_refine + args.local_refine
    for _ in range(args.local_refine):
        t = f.tables()
        phi = initial_multiple_het_3d(t["coords"], f.min_cell_diameter)
        f.refine((t["level"] < cap) & (phi[t["conn"]] < 0.4).any(axis=1))
    t = f.tables()
    t_forest = time.perf_counter() - t0
    h = f.min_cell_diameter
    centres = t["coords"][t["conn"]].mean(axis=1)
    E = 1.0 + 4.5 * (1.0 + np.sin(0.7 * centres[:, 0]) * np.cos(0.9 * centres[:, 1]) * np.sin(1.1 * centres[:, 2]))
    nu = 0.2

    def lame(Ev):
        mu = Ev / (2.0 * (1 + nu))
        return np.stack([(2 * nu * mu) / (1.0 - 2 * nu), mu], axis=1)

    params = api.Params(1.0, 1.0, 1.0, 0.0, 2.0 * h, 0.0)          # lambda, mu per cell; Eps reg = 2 h, K reg = 0
    ctx = ForestContext(f, params, device=local, cell_lame=lame(E + 1.0), cell_lame_energy=lame(E), dist=dist)
    xyz = t["coords"]
    nn = ctx.n_nodes
    on_b = ((xyz == 0.0) | (xyz == 10.0)).any(axis=1)
    m = np.zeros((nn, 4), dtype=np.uint8)
    m[on_b, :3] = 1
    ctx.set_constraints(ctx.to_block(m.reshape(-1)).astype(np.uint8), np.zeros(ctx.n_dofs, dtype=np.uint8))
    sol = np.zeros((nn, 4))
    s, c = (lambda a: np.sin(np.pi * a / 10.0)), (lambda a: np.sin(np.pi * a / 10.0))
    for d in range(3):
        sol[:, d] = 1e-3 * s(xyz[:, 0]) * s(xyz[:, 1]) * s(xyz[:, 2]) * (d + 1)
    sol[:, 3] = initial_multiple_het_3d(xyz, h)
    blk = ctx.to_block(sol.reshape(-1))
    ctx.set_state(blk, blk, blk, 0.01, 0.01, False, 10.0)
    del sol, blk
    ev = lambda: torch.cuda.Event(enable_timing=True)
    stream = torch.cuda.ExternalStream(ctx.stream)

    def timed(fn, reps):
        fn()
        ctx.synchronize()
        if world > 1:
            td.barrier()
        e0, e1 = ev(), ev()
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(reps):
            fn()
        with torch.cuda.stream(stream):
            e1.record()
        ctx.synchronize()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
            td.all_reduce(tt, op=td.ReduceOp.MAX)
            ms = float(tt[0])
        return ms

    ms_setup = timed(ctx.setup_jacobian, 2)
    x = np.random.default_rng(20240229).standard_normal(ctx.n_dofs)
    x[ctx.to_block(m.reshape(-1)) == 1] = 0.0
    x_dev, y_dev = ctx.device_vector(), ctx.device_vector()
    ctx.upload(x, x_dev)
    ms_apply = timed(lambda: ctx.vmult_dev(y_dev, x_dev), args.applies)
    ms_res = timed(lambda: ctx.residual(want_vectors=False), 3)
    y = ctx.download(y_dev)
    if rank == 0:
        print(json.dumps({"what": "forest operator apply, multiple het 3-D (BASELINE config 5)", "n_gpus": world,
                          "global_refine": args.global_refine, "local_refine": args.local_refine, "n_cells": int(f.n_cells),
                          "n_nodes": int(nn), "n_dofs": int(ctx.n_dofs), "n_hanging": int(f.n_hanging),
                          "host_forest_s": round(t_forest, 1), "ms_per_apply": ms_apply,
                          "MDoF_per_s": ctx.n_dofs / (ms_apply * 1e-3) / 1e6, "ms_per_residual": ms_res,
                          "ms_setup_jacobian": ms_setup, "norm_y": float(np.linalg.norm(y)),
                          "parallelism": "cells cut into %d contiguous ranges, nodal vectors replicated, one all-reduce per apply" % world}))
    ctx.device_vector_free(x_dev)
    ctx.device_vector_free(y_dev)
    ctx.close()
    if world > 1:
        td.destroy_process_group()


if __name__ == "__main__":
    main()
