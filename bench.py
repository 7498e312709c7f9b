#!/usr/bin/env python
"""Benchmark of the hot path: matrix-free operator apply y = J(U) x of the
(u,phi) phase-field system on the Sneddon-3D geometry (BASELINE.json config 3).

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path

A "step" is one application of the Jacobian to a device-resident vector (the
vmult inside GMRES, cracks.cc:2770).  One JSON line is printed by rank 0.
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "MDoF/s operator-apply, Sneddon-3D"
UNIT = "MDoF/s"
SEED = 20240229


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def sneddon_state(n, h):
    """Operator-benchmark state of SURVEY.md 8d in block layout [u | phi]:
    phi = InitialValuesSneddon, u_d = 1e-3 sin(pi x_d/10) prod_{e!=d} cos(pi x_e/20)."""
    ax = -10.0 + h * np.arange(n + 1)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")         # [x][y][z]
    X, Y, Z = (np.ascontiguousarray(a.transpose(2, 1, 0)).reshape(-1) for a in (X, Y, Z))  # x fastest
    hd = h * np.sqrt(3.0)
    phi = np.where((X * X + Z * Z <= 1.0) & (np.abs(2.0 * Y) <= 2.0 * hd), 0.0, 1.0)
    s = lambda a: np.sin(np.pi * a / 10.0)
    c = lambda a: np.cos(np.pi * a / 20.0)
    u = np.stack([1e-3 * s(X) * c(Y) * c(Z), 1e-3 * c(X) * s(Y) * c(Z), 1e-3 * c(X) * c(Y) * s(Z)], axis=1)
    nn = phi.shape[0]
    sol = np.concatenate([u.reshape(-1), phi])
    active = np.zeros(4 * nn, dtype=np.uint8)
    active[3 * nn:] = (phi == 0.0)
    return sol, active


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
            except ValueError:
                continue
            for name, v in zip(names, r[4:8]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def pinned_array(lib, n, dtype=np.float64):
    p = ctypes.c_void_p()
    nbytes = n * np.dtype(dtype).itemsize
    if lib.pf_host_alloc(ctypes.byref(p), nbytes) != 0:
        raise RuntimeError("pf_host_alloc failed")
    buf = (ctypes.c_char * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=n), p


def cpu_reference_sample(steps, warmup, refine=2):
    """The reference's CPU operator apply: assemble the Jacobian into CSR
    (cracks.cc:2200-2468) once, then time its vmult = CSR SpMV (cracks.cc:2770)
    with all host threads.  Bounded sample: Sneddon-3D at `refine` global
    refinements.  Uses the CPU oracle (the reference itself cannot be built
    in this image: no deal.II/Trilinos/p4est/MPI)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import newton_oracle as orc
    prob = orc.sneddon_3d(refine, kappa_of_h=lambda h: 1e-8 * h)
    n = 10 * 2 ** refine
    sol_b, act_b = sneddon_state(n, 20.0 / n)
    nn = prob.n_nodes
    sol = np.empty((nn, 4))
    sol[:, :3] = sol_b[: 3 * nn].reshape(nn, 3)
    sol[:, 3] = sol_b[3 * nn:]
    sol = sol.reshape(-1)
    con = prob.dirichlet_mask().reshape(nn, 4)
    con[:, 3] = act_b[3 * nn:]
    con = np.ascontiguousarray(con.reshape(-1))
    rowptr, col = prob.csr_pattern()
    val = np.empty(col.shape[0])
    t0 = time.perf_counter()
    getattr(orc.lib(), "pfo_assemble_jacobian_3d")(ctypes.byref(prob.mesh), ctypes.byref(prob.prm), sol, sol, sol,
                                                   con.ctypes.data_as(ctypes.c_void_p), rowptr, col, val)
    t_asm = time.perf_counter() - t0
    x = np.random.default_rng(SEED).standard_normal(prob.n_dofs)
    y = np.empty(prob.n_dofs)
    spmv = getattr(orc.lib(), "pfo_spmv_3d")
    for _ in range(warmup):
        spmv(prob.n_dofs, rowptr, col, val, x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        spmv(prob.n_dofs, rowptr, col, val, x, y)
    t = (time.perf_counter() - t0) / steps
    cores = orc.lib().pfo_num_threads()
    # the reference assembles the Jacobian once per Newton step (cracks.cc:2917) and then needs one
    # vmult per GMRES iteration (about 12 per step with a multigrid-quality preconditioner): an upper
    # bound on its Newton-its/s that ignores the AMG set-up, the V-cycles and the residual assemblies
    newton_bound = 1.0 / (t_asm + 12.0 * t)
    return {
        "value": prob.n_dofs / t / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": (f"CSR SpMV (the reference's vmult) on Sneddon-3D at {refine} global refinements: "
                   f"{prob.n_dofs} DoF, {col.shape[0]} nnz, {steps} applies after {warmup} warm-ups; "
                   f"Jacobian assembly (needed once per Newton step by the reference) took {t_asm:.2f} s "
                   f"= {prob.n_dofs / t_asm / 1e6:.3f} MDoF/s; assembly + 12 vmults bound the CPU path to "
                   f"{newton_bound:.2f} Newton-its/s at this size"),
        "newton_its_per_s_upper_bound": newton_bound,
        "ms_per_apply": t * 1e3, "assembly_s": t_asm, "n_dofs": prob.n_dofs,
    }


def run_reference(args, rank):
    if rank != 0:
        return
    cb = cpu_reference_sample(args.steps, args.warmup, refine=args.cpu_refine)
    n = 10 * 2 ** args.refine
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_apply"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Sneddon-3D operator apply (parameters_sneddon_3d.prm), {n}^3 cells Q1; CPU arm "
                               f"timed on a bounded sample at {args.cpu_refine} refinements ({cb['n_dofs']} DoF)",
                   "timing": "host wall clock around OpenMP CSR SpMV"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--refine", type=int, default=4, help="global pre-refinement steps (4 -> 16.7M DoF)")
    ap.add_argument("--cpu-refine", type=int, default=3)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--variant", type=int, default=0, help="debug: apply-kernel variant (0 = library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import cracks_b200 as pf
    from cracks_b200.api import mesh_diameter

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    def fresh_nccl_id():
        if world == 1:
            return None
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(pf.PhaseFieldContext.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        return bytes(idt.cpu().numpy().tobytes())

    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nccl_id = fresh_nccl_id()

    mesh = pf.sneddon_mesh(3, args.refine)
    params = pf.sneddon_params(mesh)
    n = mesh.n[0]
    ctx = pf.PhaseFieldContext(mesh, params, device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id)
    lib = ctx.lib
    if args.variant:
        lib.pf_debug_set_variant(args.variant)
    nd, nn = ctx.n_dofs, ctx.n_nodes

    sol, active = sneddon_state(n, mesh.h[0])
    ctx.set_state(sol, sol, sol, 1.0, 1.0, False, 1e-3)
    ctx.set_dirichlet_all_faces()
    ctx.set_constraints(None, active)
    ctx.setup_jacobian()
    x_host, _xp = pinned_array(lib, nd)
    y_host, _yp = pinned_array(lib, nd)
    x_host[:] = np.random.default_rng(SEED).standard_normal(nd)
    x_dev, y_dev = ctx.device_vector(), ctx.device_vector()
    ctx.upload(x_host, x_dev)
    del sol

    stream = torch.cuda.ExternalStream(ctx.stream)

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # clocks are sampled from here on: W warm-up steps, then an untimed pre-roll of
    # >= 0.7 s under the same load (nvidia-smi needs ~100 ms to start reporting
    # and the timed region is only tens of milliseconds), then the K timed steps
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        ctx.vmult_dev(y_dev, x_dev)
    barrier()
    t_pre = time.perf_counter()
    while True:
        for _ in range(20):
            ctx.vmult_dev(y_dev, x_dev)
        ctx.synchronize()
        done = torch.tensor([1.0 if time.perf_counter() - t_pre > 0.7 else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(done, op=dist.ReduceOp.MAX)
        if float(done[0]) > 0:
            break
    barrier()
    ctx.profile_enable(True)
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.steps):
        ctx.vmult_dev(y_dev, x_dev)
    with torch.cuda.stream(stream):
        e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count - l0
    kern_ms, kern_cnt = ctx.profile_read()
    ctx.profile_enable(False)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total, kern_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, kern_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps

    # ---- end to end through the host-buffer ABI call (H2D + apply + D2H per step)
    ctx.vmult(y_host, x_host)
    barrier()
    with torch.cuda.stream(stream):
        e0.record()
    for _ in range(args.e2e_steps):
        ctx.vmult(y_host, x_host)
    with torch.cuda.stream(stream):
        e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1) / args.e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
    lay = ctx.layout
    local_nodes = (lay.plane_end - lay.plane_begin) * lay.n_nodes_plane
    owned_nodes = (lay.owned_end - lay.owned_begin) * lay.n_nodes_plane

    newton = None
    if not args.no_newton:
        # second half of BASELINE.json's metric: Newton-its/s of the device-resident
        # active-set Newton loop (cracks.cc:2780-2994) on the same mesh (all ranks, same
        # z-slab decomposition), real time steps 0..1 of parameters_sneddon_3d.prm from the
        # interpolated initial condition
        nctx = pf.PhaseFieldContext(mesh, params, device=local_rank, rank=rank, nranks=world, nccl_id=fresh_nccl_id())
        drv = pf.SneddonDriver(nctx, pressure=lambda t: 1e-3, max_no_timesteps=1, newton_lower_bound=1e-7,
                               max_newton=50, max_line_search=10, gmres_max_it=200)
        barrier()
        t0 = time.perf_counter()
        try:
            # every decision in the loop is taken on all-reduced values, so a failure is raised on all ranks alike
            nstats, nerr = drv.run(mesh_diameter(mesh)), None
        except pf.PFError as exc:
            nstats, nerr = drv.statistics, str(exc)
        nctx.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        newton = {"newton_its_per_s": drv.newton_its / dt, "newton_its": drv.newton_its,
                  "linear_its": drv.lin_its, "time_steps": len(nstats), "wall_s": dt,
                  "crack_energy": nstats[-1]["crack"] if nstats else None,
                  "bulk_energy": nstats[-1]["bulk"] if nstats else None,
                  "preconditioner": "matrix-free geometric multigrid V-cycle (z-slab levels, replicated below), "
                                    "Chebyshev-Jacobi smoothing"}
        if nerr:
            newton["error"] = nerr
        nctx.close()

    if rank == 0:
        peak, peak_src = load_peaks()
        # algorithmic bytes of one launch of the dominant kernel on this rank
        # (SURVEY.md 8d): 24 B/DoF (x, y, U) + 9 B/node (phi~, mask)
        b_alg = 24 * 4 * local_nodes + 9 * local_nodes
        k_ms = kern_ms / max(kern_cnt, 1)
        achieved = b_alg / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        line = {
            "metric": METRIC, "value": nd / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Sneddon-3D operator apply y=J(U)x (parameters_sneddon_3d.prm, Global "
                                   f"pre-refinement steps = {args.refine}): {n}^3 cells Q1, {nn} nodes, {nd} DoF",
                       "parallelism": f"z-slab x{world}" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: x, y, U are 3 x %.0f MB per GPU vs 126 MB L2" % (8 * 4 * local_nodes / 1e6),
                       "timing": "CUDA events on the library stream, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": nd / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(8 * 4 * local_nodes), "d2h_bytes_per_step": int(8 * 4 * owned_nodes),
                    "note": "pf_apply_jacobian with pinned host buffers in the reference's block layout"
                            + ("; H2D / apply / D2H pipelined over %s chunks of cell layers" % os.environ.get("PF_E2E_CHUNKS", "16")
                               if world == 1 else "")},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_apply3d_v4<16,4,1>" if str(args.variant or os.environ.get("PF_APPLY_VARIANT", "16")) == "16" else "variant %s" % (args.variant or os.environ["PF_APPLY_VARIANT"]), "kernel_ms": k_ms,
                         "kernel_share_of_step": kern_ms / ms_total, "algorithmic_bytes": b_alg, "peak_source": peak_src,
                         "note": "FP64-pipe bound: exact 27-point FP64 quadrature, no f64 tensor path (DESIGN.md)"},
        }
        if newton is not None:
            line["newton"] = newton
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference_sample(10, 2, refine=args.cpu_refine)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    ctx.device_vector_free(x_dev)
    ctx.device_vector_free(y_dev)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
