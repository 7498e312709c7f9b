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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_refine_for_this_host(requested):
    """--cpu-refine auto: the BASELINE size (4 refinements: 1.8e9 nnz = 21.6 GB of CSR plus 14 GB of
    assembly scratch) when the host has the memory for it, else 3 (2.8 GB); the rate is size-normalised."""
    if requested != "auto":
        return int(requested)
    try:
        avail_kb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
    except (OSError, StopIteration):
        return 3
    return 4 if avail_kb >= 96 * 1024 * 1024 else 3


def cpu_reference_sample(steps, warmup, refine=2, newton=False):
    """The reference's CPU operator apply: assemble the Jacobian into CSR
    (cracks.cc:2200-2468) once, then time its vmult = CSR SpMV (cracks.cc:2770)
    with all host threads.  Bounded sample: Sneddon-3D at `refine` global
    refinements.  Uses the CPU oracle (the reference itself cannot be built
    in this image: no deal.II/Trilinos/p4est/MPI)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import newton_oracle as orc
    # torchrun exports OMP_NUM_THREADS=1 to its ranks: ask for every core this process may run on
    orc.lib().pfo_set_num_threads(host_cores())
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
    nnz = int(col.shape[0])
    del val, col, rowptr
    measured = None
    if newton:
        # a MEASURED active-set Newton step of the assembled-matrix path (SURVEY.md 8d iii): residual + active
        # set + assembly + GMRES(1e-8) + one line-search residual, at 3 refinements (about 20 s of CPU work)
        import cpu_newton
        measured = cpu_newton.time_one_newton_step(refine=min(refine, 3), threads=host_cores())
    return {
        "newton_its_per_s": measured["newton_its_per_s"] if measured else None,
        "newton_sample": ("one active-set Newton step (cracks.cc:2780-2994) of the assembled-matrix CPU path at %d DoF: "
                          "%.1f s = residual + active set + CSR assembly + %d GMRES iterations (%s) to 1e-8 + one "
                          "line-search residual" % (measured["n_dofs"], measured["wall_s"], measured["linear_its"],
                                                    measured["preconditioner"])) if measured else None,
        "value": prob.n_dofs / t / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": (f"CSR SpMV (the reference's vmult) on Sneddon-3D at {refine} global refinements: "
                   f"{prob.n_dofs} DoF, {nnz} nnz, {steps} applies after {warmup} warm-ups; "
                   f"Jacobian assembly (needed once per Newton step by the reference) took {t_asm:.2f} s "
                   f"= {prob.n_dofs / t_asm / 1e6:.3f} MDoF/s; assembly + 12 vmults bound the CPU path to "
                   f"{newton_bound:.2f} Newton-its/s at this size"),
        "newton_its_per_s_upper_bound": newton_bound,
        "ms_per_apply": t * 1e3, "assembly_s": t_asm, "n_dofs": prob.n_dofs,
    }


def kernel_info(variant):
    """name, FP64 instructions per cell and captured DRAM traffic of the apply kernel variant (profiles/kernels.json,
    written from the SASS / ncu captures named there)"""
    p = os.path.join(ROOT, "profiles", "kernels.json")
    info = json.load(open(p)) if os.path.exists(p) else {}
    k = dict(info.get("kernels", {}).get(str(variant), {"name": "apply variant %s" % variant}))
    k.setdefault("fp64_peak_tinst_per_s", info.get("fp64_peak_tinst_per_s"))
    if not k.get("fp64_peak_tinst_per_s"):
        k.pop("fp64_inst_per_cell", None)
    return k


def workload_config(refine, world):
    """`config` of both arms (identical by construction: the CPU arm times a bounded sample of this workload)"""
    n = 10 * 2 ** refine
    nn = (n + 1) ** 3
    return {"workload": f"Sneddon-3D operator apply y=J(U)x (parameters_sneddon_3d.prm, Global pre-refinement "
                        f"steps = {refine}): {n}^3 cells Q1, {nn} nodes, {4 * nn} DoF",
            "parallelism": f"z-slab x{world}" if world > 1 else "single GPU"}


def run_reference(args, rank):
    if rank != 0:
        return
    refine = cpu_refine_for_this_host(args.cpu_refine)
    cb = cpu_reference_sample(args.steps, args.warmup, refine=refine)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_apply"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.refine, args.gpus),
        "timing": f"host wall clock around the OpenMP CSR SpMV, {cb['cores']} threads; bounded sample of the workload at "
                  f"{refine} refinements ({cb['n_dofs']} DoF), the rate is per DoF",
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
    ap.add_argument("--cpu-refine", default="auto", help="refinements of the CPU arm's sample: 3, 4 or auto (4 if the host has >= 96 GB free)")
    ap.add_argument("--no-cpu-newton", action="store_true", help="skip the measured CPU Newton step (about 20 s)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-newton", action="store_true")
    ap.add_argument("--no-fp32", action="store_true", help="skip the timing of the FP32 (inexact-Newton) Jacobian")
    ap.add_argument("--no-block-solve", action="store_true", help="skip the A/B Newton runs with the other linear-solve structure")
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

    parity = None
    if world > 1:
        # multi-GPU parity record (outside every timed region; the oracle is the checker): apply / residual /
        # diagonal / functionals of the slab-decomposed path on `world` ranks against the single-domain CPU oracle
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import newton_oracle as orc
        from mgpu_check import small_mesh_parity

        def allsum(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
            dist.all_reduce(t)
            return t.cpu().numpy()

        errs = small_mesh_parity(pf, orc, rank, world, local_rank, nccl_id, allsum)
        if rank == 0:
            parity = {"n_ranks": world, "jv_relerr": errs["apply"], "residual_relerr": max(errs["r_pde"], errs["r_total"]),
                      "diag_relerr": errs["diag"], "crack_energy_relerr": errs["crack"], "bulk_energy_relerr": errs["bulk"],
                      "against": "single-domain CPU oracle, 12 x 9 x %d cells" % max(10, 3 * world),
                      "ok": bool(all(v <= 1e-11 for v in errs.values()))}
        nccl_id = fresh_nccl_id()

    mesh = pf.sneddon_mesh(3, args.refine)
    params = pf.sneddon_params(mesh)
    n = mesh.n[0]
    ctx = pf.PhaseFieldContext(mesh, params, device=local_rank, rank=rank, nranks=world, nccl_id=nccl_id)
    lib = ctx.lib
    if args.variant:
        lib.pf_debug_set_variant(ctx.h, args.variant)
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

    # ---- the inexact-Newton operator: the same 27-point evaluation in FP32 on FP64 vectors (pf_set_jacobian_precision).
    # Reported beside the headline, never instead of it: `value` above is the exact FP64 operator.
    inexact = None
    if not args.no_fp32 and not args.variant:
        ctx.vmult_dev(y_dev, x_dev)
        y64 = ctx.download(y_dev)
        ctx.set_jacobian_precision(32)
        ctx.setup_jacobian()
        for _ in range(args.warmup):
            ctx.vmult_dev(y_dev, x_dev)
        barrier()
        ctx.profile_enable(True)
        with torch.cuda.stream(stream):
            e0.record()
        for _ in range(args.steps):
            ctx.vmult_dev(y_dev, x_dev)
        with torch.cuda.stream(stream):
            e1.record()
        barrier()
        ms32 = e0.elapsed_time(e1) / args.steps
        k32_ms, k32_cnt = ctx.profile_read()
        ctx.profile_enable(False)
        y32 = ctx.download(y_dev)
        if world > 1:
            t = torch.tensor([ms32, k32_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms32, k32_ms = float(t[0]), float(t[1])
        inexact = {"dtype": "f32 arithmetic on f64 vectors", "ms_per_step": ms32, "value": nd / (ms32 * 1e-3) / 1e6, "unit": UNIT,
                   "kernel_ms": k32_ms / max(k32_cnt, 1),
                   "relerr_vs_f64": float(np.max(np.abs(y32 - y64)) / np.max(np.abs(y64)))}
        del y32, y64
        ctx.set_jacobian_precision(64)
        ctx.setup_jacobian()

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
    newton_inexact = None
    if not args.no_newton:
        # second half of BASELINE.json's metric: Newton-its/s of the device-resident
        # active-set Newton loop (cracks.cc:2780-2994) on the same mesh (all ranks, same
        # z-slab decomposition), real time steps 0..1 of parameters_sneddon_3d.prm from the
        # interpolated initial condition.  Run twice: with the exact FP64 Jacobian (the library default) and
        # as inexact Newton with the FP32 Jacobian (pf_set_jacobian_precision); residuals are FP64 in both.
        def newton_run(jacobian_bits, block_solve=None):
            nctx = pf.PhaseFieldContext(mesh, params, device=local_rank, rank=rank, nranks=world, nccl_id=fresh_nccl_id())
            if block_solve is not None:
                nctx.set_block_solve(block_solve)
            stages = nctx.block_solve()
            if jacobian_bits != 64:
                nctx.set_jacobian_precision(jacobian_bits)
                nctx.set_multigrid_precision(jacobian_bits)
            # one untimed time step first: the context allocates its Krylov basis, multigrid levels and coefficient
            # records on first use (the reference allocates in setup_system(), outside newton_active_set() too)
            pf.SneddonDriver(nctx, pressure=lambda t: 1e-3, max_no_timesteps=0, newton_lower_bound=1e-7, max_newton=50,
                             max_line_search=10, gmres_max_it=200).run(mesh_diameter(mesh))
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
            out = {"newton_its_per_s": drv.newton_its / dt, "newton_its": drv.newton_its,
                   "linear_its": drv.lin_its, "time_steps": len(nstats), "wall_s": dt,
                   "crack_energy": nstats[-1]["crack"] if nstats else None,
                   "bulk_energy": nstats[-1]["bulk"] if nstats else None,
                   "jacobian": "exact FP64 27-point apply" if jacobian_bits == 64 else "FP32 27-point apply on FP64 vectors (inexact Newton)",
                   "preconditioner": "matrix-free geometric multigrid V-cycle in FP%d (z-slab levels, replicated below), "
                                     "Chebyshev-Jacobi smoothing" % jacobian_bits,
                   "linear_solve": ("u stage + phi stage (pf_set_block_solve: block (u,phi) of the Jacobian is zero, "
                                    "cracks.cc:2333-2337)" if stages else "one GMRES on the whole system")
                                   + (" [library default]" if block_solve is None else "")}
            if nerr:
                out["error"] = nerr
            if stages:
                try:    # counters since the context was created (the untimed first time step included)
                    out["stages"] = nctx.block_solve_stats()
                except Exception as exc:    # a diagnostic must not cost the bench line
                    out["stages"] = str(exc)
            nctx.close()
            return out

        newton = newton_run(64)
        newton_inexact = newton_run(32)
        newton_stages = None
        if not args.no_block_solve:
            # the same two runs with the other linear-solve structure than the library default (A/B record)
            other = not ctx.block_solve()
            newton_stages = {"block_solve": other, "exact": newton_run(64, other), "inexact": newton_run(32, other)}

    if rank == 0:
        peak, peak_src = load_peaks()
        # algorithmic bytes of one launch of the dominant kernel on this rank
        # (SURVEY.md 8d): 24 B/DoF (x, y, U) + 9 B/node (phi~, mask)
        b_alg = 24 * 4 * local_nodes + 9 * local_nodes
        k_ms = kern_ms / max(kern_cnt, 1)
        achieved = b_alg / (k_ms * 1e-3) / 1e9
        kinfo = kernel_info(args.variant or int(os.environ.get("PF_APPLY_VARIANT", "0")) or "default")
        # DRAM bytes of one launch from the committed ncu capture of the same kernel at this size; only quoted
        # for the launch it was captured on (one GPU, the whole mesh)
        traffic = kinfo.get("dram_bytes_per_launch") if (world == 1 and args.refine == 4) else None
        local_cells = (lay.plane_end - lay.plane_begin - 1) * n * n      # incl. the redundant layer of a slab
        fp64 = None
        if kinfo.get("fp64_inst_per_cell"):
            # the resource that binds an exact FP64 evaluation: thread-level FP64 instructions (DFMA / DADD / DMUL share one
            # pipe) per second against the DFMA issue rate measured on this pool (tools/fp64_peak.cu)
            rate = kinfo["fp64_inst_per_cell"] * local_cells / (k_ms * 1e-3) / 1e12
            fp64 = {"bound": "fp64", "achieved": rate, "peak": kinfo["fp64_peak_tinst_per_s"], "unit": "T FP64 inst/s",
                    "frac": rate / kinfo["fp64_peak_tinst_per_s"], "fp64_inst_per_cell": kinfo["fp64_inst_per_cell"],
                    "source": kinfo.get("source")}
        line = {
            "metric": METRIC, "value": nd / (ms_step * 1e-3) / 1e6, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.refine, world),
            "timing": "CUDA events on the library stream, max over ranks; inputs larger than L2: x, y, U are 3 x %.0f MB "
                      "per GPU vs 126 MB L2" % (8 * 4 * local_nodes / 1e6),
            "clocks": clocks,
            "e2e": {"value": nd / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(8 * 4 * local_nodes), "d2h_bytes_per_step": int(8 * 4 * owned_nodes),
                    "note": "pf_apply_jacobian with pinned host buffers in the reference's block layout; H2D / apply / D2H "
                            "pipelined over chunks of cell layers"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": kinfo["name"], "kernel_ms": k_ms,
                         "kernel_share_of_step": kern_ms / ms_total, "algorithmic_bytes": b_alg, "peak_source": peak_src,
                         "binding_resource": "fp64 pipe (see roofline_fp64): exact 27-point FP64 quadrature, no f64 tensor "
                                             "path on sm_100a (DESIGN.md 5.1)"},
        }
        if fp64:
            line["roofline_fp64"] = fp64
        if inexact:
            inexact["roofline_frac_hbm"] = b_alg / (inexact["kernel_ms"] * 1e-3) / 1e9 / peak
            line["inexact_newton_operator"] = inexact
        if parity is not None:
            line["parity"] = parity
        if newton is not None:
            line["newton"] = newton
            line["newton_inexact"] = newton_inexact
            if newton_stages is not None:
                line["newton_other_linear_solve"] = newton_stages
        if not args.no_cpu_baseline:
            cb = cpu_reference_sample(10, 2, refine=3, newton=not args.no_cpu_newton)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample", "newton_its_per_s",
                                                       "newton_sample")}
        print(json.dumps(line))
    ctx.device_vector_free(x_dev)
    ctx.device_vector_free(y_dev)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
