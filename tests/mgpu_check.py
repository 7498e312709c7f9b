"""Multi-GPU parity check, launched by torchrun (one rank per GPU, NCCL):
the slab-decomposed CUDA path must reproduce the CPU oracle (single domain)
on apply / residual / diagonal / energies, and the KAT-1 golden end to end.
Prints MGPU_OK on rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def small_mesh_parity(pf, orc, rank, world, local, nccl_id, allsum, nz=None):
    """apply / residual / diagonal / functionals of the slab-decomposed CUDA path on `world` ranks against the
    single-domain CPU oracle on a small anisotropic mesh (12 x 9 x max(10, 3*world) cells); relative errors, rank 0."""
    nz = nz or max(10, 3 * world)        # >= 16 layers per rank: pf_apply_jacobian takes its chunk-pipelined path
    n, h = (12, 9, nz), (0.5, 0.4, 0.3)
    lo = tuple(-0.5 * n[d] * h[d] for d in range(3))
    hi = tuple(0.5 * n[d] * h[d] for d in range(3))
    prob = orc.Problem(3, n, lo, hi, kappa_of_h=lambda hh: 1e-3, pressure=1e-3)
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 0.5
    rng = np.random.default_rng(7)
    nn, nc = prob.n_nodes, 4
    sol = np.zeros((nn, nc)); sol[:, :3] = 1e-2 * rng.standard_normal((nn, 3)); sol[:, 3] = rng.random(nn)
    old = sol.copy(); old[:, 3] = rng.random(nn)
    oo = old.copy(); oo[:, 3] += 0.5 * (rng.random(nn) - 0.5)
    sol, old, oo = sol.reshape(-1), old.reshape(-1), oo.reshape(-1)
    con = prob.dirichlet_mask().reshape(nn, nc); con[rng.random(nn) < 0.2, 3] = 1
    con = np.ascontiguousarray(con.reshape(-1))
    x = rng.standard_normal(nn * nc)

    mesh = pf.Mesh(); mesh.dim = 3
    for d in range(3):
        mesh.n[d], mesh.h[d], mesh.origin[d] = n[d], h[d], lo[d]
    params = pf.Params(prob.prm.lam, prob.prm.mu, prob.prm.G_c, prob.prm.kappa, prob.prm.eps, 0.0)
    ctx = pf.PhaseFieldContext(mesh, params, device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    ctx.set_state(ctx.to_block(sol), ctx.to_block(old), ctx.to_block(oo), 1.0, 0.5, False, prob.pressure)
    cb = ctx.to_block(con).astype(np.uint8)
    ctx.set_constraints(cb, cb)
    r_pde, r_tot, nrm = ctx.residual()
    r_pde, r_tot = allsum(r_pde), allsum(r_tot)           # every rank filled only its owned planes
    ctx.setup_jacobian()
    y = np.zeros(nn * nc)
    ctx.vmult(y, ctx.to_block(x))
    y = allsum(y)
    diag = allsum(ctx.jacobian_diagonal())
    bulk, crack = ctx.energy()
    tcv = ctx.tcv()
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    errs = {}
    if rank == 0:
        r_pde_ref, r_tot_ref = prob.residual(sol, old, oo, con)
        errs["apply"] = rel(ctx.to_nodal(y), prob.apply_jacobian(sol, old, oo, con, x))
        errs["r_total"] = rel(ctx.to_nodal(r_tot), r_tot_ref)
        errs["r_pde"] = rel(ctx.to_nodal(r_pde), r_pde_ref)
        errs["norm"] = abs(nrm - np.linalg.norm(r_pde_ref)) / np.linalg.norm(r_pde_ref)
        errs["diag"] = rel(ctx.to_nodal(diag), prob.jacobian(sol, old, oo, None).diagonal())
        b_ref, c_ref = prob.energy(sol)
        errs["bulk"], errs["crack"] = abs(bulk - b_ref) / b_ref, abs(crack - c_ref) / c_ref
        errs["tcv"] = abs(tcv - prob.tcv(sol)) / abs(prob.tcv(sol))
    ctx.close()

    return errs


def main():
    import torch
    import torch.distributed as dist
    import cracks_b200 as pf
    import newton_oracle as orc
    from cracks_b200.api import mesh_diameter

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(pf.PhaseFieldContext.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    nccl_id = idt.cpu().numpy().tobytes()

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
        dist.all_reduce(t)
        return t.cpu().numpy()

    errs = small_mesh_parity(pf, orc, rank, world, local, nccl_id, allsum)
    if rank == 0:
        print("errors vs single-domain oracle:", errs, flush=True)
        assert all(v <= 1e-11 for v in errs.values()), errs
    # slabs of 18 cell layers: the host-buffer apply runs as the H2D / apply / D2H pipeline on every rank
    errs = small_mesh_parity(pf, orc, rank, world, local, _fresh_id(pf, dist, torch, rank), allsum, nz=18 * world)
    if rank == 0:
        print("errors vs single-domain oracle, pipelined host apply:", errs, flush=True)
        assert all(v <= 1e-11 for v in errs.values()), errs

    # KAT-1 end to end on `world` GPUs
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "sneddon_3d_1.json")))
    prm = golden["prm"]
    mesh = pf.sneddon_mesh(3, 0)
    ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh, kappa_of_h=lambda hh: 0.0), device=local, rank=rank,
                               nranks=world, nccl_id=None if world == 1 else _fresh_id(pf, dist, torch, rank))
    drv = pf.SneddonDriver(ctx, pressure=lambda t: prm["pressure"], max_no_timesteps=prm["max_no_timesteps"],
                           newton_lower_bound=prm["newton_lower_bound"], max_newton=prm["newton_max_steps"],
                           max_line_search=prm["line_search_max_steps"], gmres_max_it=3000)
    stats = drv.run(mesh_diameter(mesh))
    if rank == 0:
        for got, ref in zip(stats, golden["statistics"]):
            assert abs(got["crack"] - ref["crack"]) <= 1e-8 * ref["crack"], (got, ref)
            assert abs(got["bulk"] - ref["bulk"]) <= 1e-6 * ref["bulk"], (got, ref)
        assert abs(drv.tcv - golden["tcv"]) <= 1e-5 * golden["tcv"]
        print("KAT-1 on %d GPUs:" % world, [(s["bulk"], s["crack"]) for s in stats], "tcv", drv.tcv, flush=True)
    ctx.close()

    # multigrid across ranks (z-slab levels 40 -> 20 -> 10, replicated 5^3 below): one time
    # step at 2 refinements must converge like the single-GPU hierarchy and give its energies
    mesh = pf.sneddon_mesh(3, 2)
    ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh), device=local, rank=rank, nranks=world,
                               nccl_id=None if world == 1 else _fresh_id(pf, dist, torch, rank))
    drv = pf.SneddonDriver(ctx, pressure=lambda t: 1e-3, max_no_timesteps=0, newton_lower_bound=1e-7, gmres_max_it=60)
    st = drv.run(mesh_diameter(mesh))[-1]
    ctx.close()
    if rank == 0:
        ctx1 = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh), device=local)
        drv1 = pf.SneddonDriver(ctx1, pressure=lambda t: 1e-3, max_no_timesteps=0, newton_lower_bound=1e-7, gmres_max_it=60)
        st1 = drv1.run(mesh_diameter(mesh))[-1]
        ctx1.close()
        print("multigrid on %d ranks: newton %d linear %d | single GPU: newton %d linear %d | crack %.10e vs %.10e"
              % (world, drv.newton_its, drv.lin_its, drv1.newton_its, drv1.lin_its, st["crack"], st1["crack"]), flush=True)
        assert abs(st["crack"] - st1["crack"]) <= 1e-8 * st1["crack"], (st, st1)
        assert abs(st["bulk"] - st1["bulk"]) <= 1e-6 * st1["bulk"], (st, st1)
        assert drv.lin_its <= 1.5 * drv1.lin_its + 10, (drv.lin_its, drv1.lin_its)
        print("MGPU_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def _fresh_id(pf, dist, torch, rank):
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.frombuffer(bytearray(pf.PhaseFieldContext.nccl_unique_id()), dtype=torch.uint8).cuda()
    dist.broadcast(idt, 0)
    return idt.cpu().numpy().tobytes()


if __name__ == "__main__":
    main()
