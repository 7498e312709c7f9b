#!/usr/bin/env python
"""Newton-its/s of the device-resident active-set Newton loop (cracks.cc:2780-2994)
on Sneddon-3D (parameters_sneddon_3d.prm values) at a given global refinement.

  python tools/newton_bench.py --refine 3 --steps 2 [--precond 1 --degree 2 --ratio 6]

Prints the reference-style Newton table and one JSON line with Newton
iterations per second, total linear iterations and the energies.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--refine", type=int, default=3)
    ap.add_argument("--steps", type=int, default=2, help="time steps to run")
    ap.add_argument("--precond", type=int, default=1)
    ap.add_argument("--degree", type=int, default=2)
    ap.add_argument("--ratio", type=float, default=6.0)
    ap.add_argument("--gmres-max-it", type=int, default=200)
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--jacobian-bits", type=int, default=64, help="32: inexact Newton, Jacobian apply in FP32")
    ap.add_argument("--uncoupled", action="store_true", help="block-diagonal smoother operator (pf_set_multigrid_coupling 0)")
    ap.add_argument("--block-solve", type=int, default=-1, help="1 / 0: pf_set_block_solve (default: the library's setting)")
    ap.add_argument("--mg-bits", type=int, default=0, help="32 / 64: precision of the V-cycle (0 = library default)")
    args = ap.parse_args()
    import cracks_b200 as pf
    from cracks_b200.api import mesh_diameter

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    nccl_id = None
    if world > 1:                                                        # torchrun: one rank per GPU
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pf.PhaseFieldContext.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = idt.cpu().numpy().tobytes()
        if rank != 0:
            args.quiet = True
    mesh = pf.sneddon_mesh(3, args.refine)
    ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh), device=local, rank=rank, nranks=world,
                               nccl_id=nccl_id)                          # K reg = 1e-8*h, Eps reg = 2h
    ctx.set_preconditioner(args.precond, args.degree, args.ratio)
    if args.jacobian_bits != 64:
        ctx.set_jacobian_precision(args.jacobian_bits)
    if args.mg_bits:
        ctx.set_multigrid_precision(args.mg_bits)
    if args.uncoupled:
        ctx.set_multigrid_coupling(False)
    if args.block_solve >= 0:
        ctx.set_block_solve(bool(args.block_solve))
    log = (lambda s: None) if args.quiet else (lambda s: print(s, flush=True))
    # parameters_sneddon_3d.prm: Newton lower bound 1e-7, max 50 steps, line search 10 x 0.5
    drv = pf.SneddonDriver(ctx, pressure=lambda t: 1e-3, max_no_timesteps=args.steps - 1, newton_lower_bound=1e-7,
                           max_newton=50, max_line_search=10, gmres_max_it=args.gmres_max_it, log=log)
    ctx.synchronize()
    t0 = time.perf_counter()
    try:
        stats = drv.run(mesh_diameter(mesh))
        err = None
    except pf.PFError as e:
        stats, err = drv.statistics, str(e)
    ctx.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0 and hasattr(ctx.lib, "calls"):
        print(json.dumps({k: [v[0], round(v[1], 4)] for k, v in sorted(ctx.lib.calls.items(), key=lambda kv: -kv[1][1])}))
    if rank == 0:
      print(json.dumps({"refine": args.refine, "n_gpus": world, "phase_s": drv.phase_s, "n_dofs": ctx.n_dofs, "time_steps": len(stats),
                        "newton_its": drv.newton_its, "linear_its": drv.lin_its, "wall_s": dt,
                        "newton_its_per_s": drv.newton_its / dt if dt > 0 else None,
                        "precond": args.precond, "cheb_degree": args.degree, "cheb_ratio": args.ratio,
                        "statistics": stats, "error": err}))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
