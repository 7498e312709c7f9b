"""The WHOLE product library on the CPU: tests/emu/build_emulated_library.py compiles cracks_b200/csrc's own
sources (pf_api.cu and every kernel header) with g++ against a shim of the CUDA runtime; "device" memory is
host memory and kernel launches run their blocks on the CPU (one OS thread per CUDA thread where a kernel
synchronises).  The Python mirror (cracks_b200/api.py, forest.py) is pointed at that build for the duration of
a test, so the same calls the GPU suite makes -- C ABI, launch sequences, Newton / GMRES / multigrid glue,
drivers -- run end to end and are held against the oracle and the reference's goldens without a GPU.
This is a checker (test infrastructure), not a fallback: the product loads only libcracks_b200.so and fails
without a CUDA device."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu_so():
    sys.path.insert(0, os.path.join(HERE, "emu"))
    import build_emulated_library as b
    so = os.path.join(HERE, "emu", "libcracks_b200_emu.so")
    srcs = [os.path.join(b.SRC, f) for f in os.listdir(b.SRC) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(HERE, "emu", f) for f in ("emu_runtime.cc", "build_emulated_library.py",
                                                   os.path.join("cuda_shim_full", "cuda_runtime.h"))]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        b.build()
    return so


@pytest.fixture()
def epf(emu_so, monkeypatch):
    """cracks_b200 with its ctypes mirror bound to the emulated build"""
    import cracks_b200 as pf
    from cracks_b200 import api
    monkeypatch.setattr(api, "library_path", lambda: emu_so)
    monkeypatch.setattr(api, "_LIB", None)
    yield pf
    api._LIB = None


def _relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


def test_apply_residual_diag_3d(oracle, epf):
    pf = epf
    n, h = (18, 6, 3), (0.5, 0.5, 0.5)
    rng = np.random.default_rng(12)
    lo = tuple(-0.5 * n[d] * h[d] for d in range(3)); hi = tuple(0.5 * n[d] * h[d] for d in range(3))
    prob = oracle.Problem(3, n, lo, hi, kappa_of_h=lambda hh: 1e-3, eps_of_h=lambda hh: 2.0 * hh, pressure=1e-3)
    nn = prob.n_nodes
    sol = np.zeros((nn, 4)); sol[:, :3] = 1e-2 * rng.standard_normal((nn, 3)); sol[:, 3] = rng.random(nn)
    old = sol.copy(); old[:, 3] = rng.random(nn)
    oo = old.copy(); oo[:, 3] += 0.5 * (rng.random(nn) - 0.5)
    sol, old, oo = sol.reshape(-1), old.reshape(-1), oo.reshape(-1)
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 0.5
    con = prob.dirichlet_mask().reshape(nn, 4); con[rng.random(nn) < 0.2, 3] = 1
    con = np.ascontiguousarray(con.reshape(-1))
    mesh = pf.Mesh(); mesh.dim = 3
    for d in range(3):
        mesh.n[d], mesh.h[d], mesh.origin[d] = n[d], h[d], lo[d]
    ctx = pf.PhaseFieldContext(mesh, pf.Params(prob.prm.lam, prob.prm.mu, prob.prm.G_c, prob.prm.kappa, prob.prm.eps, 0.0))
    ctx.set_state(ctx.to_block(sol), ctx.to_block(old), ctx.to_block(oo), 1.0, 0.5, False, prob.pressure)
    cb = ctx.to_block(con).astype(np.uint8)
    ctx.set_constraints(cb, cb)
    r_pde, r_tot, nrm = ctx.residual()
    r_pde_ref, r_tot_ref = prob.residual(sol, old, oo, con)
    assert _relerr(ctx.to_nodal(r_tot), r_tot_ref) <= 1e-12 and _relerr(ctx.to_nodal(r_pde), r_pde_ref) <= 1e-12
    assert nrm == pytest.approx(np.linalg.norm(r_pde_ref), rel=1e-12)
    ctx.set_preconditioner(0, 2, 20.0)
    ctx.setup_jacobian()
    x = rng.standard_normal(prob.n_dofs)
    y = np.zeros(prob.n_dofs)
    ctx.vmult(y, ctx.to_block(x))                                  # n[2] < 16: upload / apply_dev / download path
    assert _relerr(ctx.to_nodal(y), prob.apply_jacobian(sol, old, oo, con, x)) <= 1e-12
    assert _relerr(ctx.to_nodal(ctx.jacobian_diagonal()), prob.jacobian(sol, old, oo, None).diagonal()) <= 1e-12
    b_ref, c_ref = prob.energy(sol)
    bulk, crack = ctx.energy()
    assert bulk == pytest.approx(b_ref, rel=1e-12) and crack == pytest.approx(c_ref, rel=1e-12)
    # pf_set_deterministic: the same operator / residual / diagonal launched colour by colour (8 launches whose tiles
    # share no node); blocks run concurrently in the emulation too, so the bits must repeat
    ctx.set_deterministic(True)
    ctx.setup_jacobian()
    outs = []
    for _ in range(2):
        yd = np.zeros(prob.n_dofs)
        ctx.vmult(yd, ctx.to_block(x))
        rd = ctx.residual()[1]
        outs.append((yd, rd, ctx.jacobian_diagonal()))
        ctx.setup_jacobian()
    assert all(np.array_equal(a, b) for a, b in zip(*outs))
    assert _relerr(ctx.to_nodal(outs[0][0]), prob.apply_jacobian(sol, old, oo, con, x)) <= 1e-12
    assert _relerr(ctx.to_nodal(outs[0][1]), r_tot_ref) <= 1e-12
    assert _relerr(ctx.to_nodal(outs[0][2]), prob.jacobian(sol, old, oo, None).diagonal()) <= 1e-12
    ctx.close()


@pytest.mark.parametrize("nz", [16, 45])
def test_host_buffer_apply_pipeline(oracle, epf, nz):
    """pf_apply_jacobian with n[2] >= 16 takes the three-stream pipeline over chunks of cell layers (upload,
    permute + operator on a layer range, download of the finished planes): 8 chunks of 2 layers / 16 ragged
    chunks of 2-3 layers, against the oracle."""
    pf = epf
    n, h = (5, 4, nz), (0.5, 0.5, 0.5)
    rng = np.random.default_rng(5)
    lo = tuple(-0.5 * n[d] * h[d] for d in range(3)); hi = tuple(0.5 * n[d] * h[d] for d in range(3))
    prob = oracle.Problem(3, n, lo, hi, kappa_of_h=lambda hh: 1e-3, eps_of_h=lambda hh: 2.0 * hh, pressure=1e-3)
    nn = prob.n_nodes
    sol = np.zeros((nn, 4)); sol[:, :3] = 1e-2 * rng.standard_normal((nn, 3)); sol[:, 3] = rng.random(nn)
    sol = sol.reshape(-1)
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 1.0
    con = prob.dirichlet_mask().reshape(nn, 4); con[rng.random(nn) < 0.2, 3] = 1
    con = np.ascontiguousarray(con.reshape(-1))
    mesh = pf.Mesh(); mesh.dim = 3
    for d in range(3):
        mesh.n[d], mesh.h[d], mesh.origin[d] = n[d], h[d], lo[d]
    ctx = pf.PhaseFieldContext(mesh, pf.Params(prob.prm.lam, prob.prm.mu, prob.prm.G_c, prob.prm.kappa, prob.prm.eps, 0.0))
    blk = ctx.to_block(sol)
    ctx.set_state(blk, blk, blk, 1.0, 1.0, False, prob.pressure)
    cb = ctx.to_block(con).astype(np.uint8)
    ctx.set_constraints(cb, cb)
    ctx.set_preconditioner(0, 2, 20.0)
    ctx.setup_jacobian()
    for _ in range(2):                                              # twice: the staging buffers are reused
        x = rng.standard_normal(prob.n_dofs)
        y = np.zeros(prob.n_dofs)
        ctx.vmult(y, ctx.to_block(x))
        assert _relerr(ctx.to_nodal(y), prob.apply_jacobian(sol, sol, sol, con, x)) <= 1e-12
    ctx.close()


def test_kat1_first_time_steps_with_multigrid(epf):
    """sneddon_3d_1 golden through SneddonDriver: active-set Newton, GMRES, the multigrid V-cycle (10^3 -> 5^3)"""
    pf = epf
    from cracks_b200.api import mesh_diameter
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_3d_1.json")))
    mesh = pf.sneddon_mesh(3, 0)
    ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh, kappa_of_h=lambda hh: 0.0))
    drv = pf.SneddonDriver(ctx, pressure=lambda t: g["prm"]["pressure"], max_no_timesteps=0,
                           newton_lower_bound=g["prm"]["newton_lower_bound"], max_newton=g["prm"]["newton_max_steps"],
                           max_line_search=g["prm"]["line_search_max_steps"], gmres_max_it=300)
    stats = drv.run(mesh_diameter(mesh))
    assert stats[0]["crack"] == pytest.approx(g["statistics"][0]["crack"], rel=1e-8)
    assert stats[0]["bulk"] == pytest.approx(g["statistics"][0]["bulk"], rel=1e-7)
    assert stats[0]["diff"] == pytest.approx(g["timestep_difference_linfty"][0], rel=2e-6)
    assert drv.lin_its / drv.newton_its < 40                        # the V-cycle preconditions (Jacobi needs hundreds)
    ctx.close()


@pytest.mark.parametrize("h", [(0.5, 0.5, 0.5), (0.5, 0.4, 0.3)])
def test_block_restricted_operator(oracle, epf, h):
    """pf_set_block_solve: the operator restricted to the u block / the phi block (pf_debug_set_block) equals the
    oracle's Jacobian with the other block's dofs constrained -- on cubic cells the phi block is the dedicated
    kernel of pf_apply3d_phi.cuh (27-point records) -- and block (u,phi) of the Jacobian is zero
    (cracks.cc:2333-2337)."""
    pf = epf
    n = (18, 6, 3)
    rng = np.random.default_rng(21)
    lo = tuple(-0.5 * n[d] * h[d] for d in range(3)); hi = tuple(0.5 * n[d] * h[d] for d in range(3))
    prob = oracle.Problem(3, n, lo, hi, kappa_of_h=lambda hh: 1e-3, eps_of_h=lambda hh: 2.0 * hh, pressure=1e-3)
    nn = prob.n_nodes
    sol = np.zeros((nn, 4)); sol[:, :3] = 1e-2 * rng.standard_normal((nn, 3)); sol[:, 3] = rng.random(nn)
    old = sol.copy(); old[:, 3] = rng.random(nn)
    sol, old = sol.reshape(-1), old.reshape(-1)
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 1.0
    con = prob.dirichlet_mask().reshape(nn, 4); con[rng.random(nn) < 0.2, 3] = 1
    con_u = con.copy(); con_u[:, 3] = 1                             # u stage: every phi dof constrained
    con_p = con.copy(); con_p[:, :3] = 1                            # phi stage: every u dof constrained
    con, con_u, con_p = (np.ascontiguousarray(c.reshape(-1)) for c in (con, con_u, con_p))
    J = prob.jacobian(sol, old, old, None).tocsr()
    iu = np.arange(4 * nn).reshape(nn, 4)[:, :3].ravel(); ip = np.arange(4 * nn).reshape(nn, 4)[:, 3]
    assert J[iu][:, ip].nnz == 0 or abs(J[iu][:, ip]).max() == 0.0
    mesh = pf.Mesh(); mesh.dim = 3
    for d in range(3):
        mesh.n[d], mesh.h[d], mesh.origin[d] = n[d], h[d], lo[d]
    ctx = pf.PhaseFieldContext(mesh, pf.Params(prob.prm.lam, prob.prm.mu, prob.prm.G_c, prob.prm.kappa, prob.prm.eps, 0.0))
    ctx.set_state(ctx.to_block(sol), ctx.to_block(old), ctx.to_block(old), 1.0, 1.0, False, prob.pressure)
    cb = ctx.to_block(con).astype(np.uint8)
    ctx.set_constraints(cb, cb)
    ctx.set_preconditioner(0, 2, 20.0)
    ctx.set_block_solve(True)
    ctx.setup_jacobian()
    x = rng.standard_normal(prob.n_dofs)
    for block, c in ((1, con_u), (2, con_p), (0, con)):
        ctx._check(ctx.lib.pf_debug_set_block(ctx.h, block))
        y = np.zeros(prob.n_dofs)
        ctx.vmult(y, ctx.to_block(x))
        assert _relerr(ctx.to_nodal(y), prob.apply_jacobian(sol, old, old, c, x)) <= 1e-12, block
    ctx.close()


@pytest.mark.parametrize("mg_bits,jac_bits,steps", [(64, 64, 1), (32, 32, 0)])
def test_block_solve_kat1(epf, mg_bits, jac_bits, steps):
    """pf_set_block_solve on the sneddon_3d_1 golden: pf_solve as u stage + phi stage gives the golden energies and the
    Newton history of the monolithic solve; once u has been solved the later Newton steps of a time step skip the
    u stage (their GMRES iterations are phi iterations: fewer per Newton step than the monolithic solve needs)."""
    pf = epf
    from cracks_b200.api import mesh_diameter
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_3d_1.json")))
    out = {}
    for block in (False, True):
        mesh = pf.sneddon_mesh(3, 0)
        ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh, kappa_of_h=lambda hh: 0.0))
        ctx.set_multigrid_precision(mg_bits)
        if jac_bits != 64:
            ctx.set_jacobian_precision(jac_bits)
        ctx.set_block_solve(block)
        drv = pf.SneddonDriver(ctx, pressure=lambda t: g["prm"]["pressure"], max_no_timesteps=steps,
                               newton_lower_bound=g["prm"]["newton_lower_bound"], max_newton=g["prm"]["newton_max_steps"],
                               max_line_search=g["prm"]["line_search_max_steps"], gmres_max_it=300)
        stats = drv.run(mesh_diameter(mesh))
        out[block] = (stats, drv.newton_its, drv.lin_its, [[(r.n_active, r.line_search) for r in rows] for rows in drv.history])
        st = ctx.block_solve_stats()
        if block:
            # every solve ran as stages; the u stage ran in the first Newton step of a time step and was skipped in
            # others (|b_u| is what the previous u solve left); the iteration counts add up to the driver's
            assert st["solves"] == drv.newton_its and 1 <= st["with_u_stage"] < st["solves"], st
            assert st["u_iterations"] + st["phi_iterations"] == drv.lin_its and st["phi_iterations"] > 0, st
        else:
            assert st["solves"] == 0, st
        ctx.close()
    for block in (False, True):
        for got, ref in zip(out[block][0], g["statistics"]):
            assert got["crack"] == pytest.approx(ref["crack"], rel=1e-8)
            assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-6)
    # The intermediate active sets are decided by the last digits of the linear solves (cracks.cc:2863 tests `> 0` at
    # tolerance 0; SURVEY.md 8c: not a parity quantity) and the stages solve u more accurately than the monolithic
    # GMRES: the histories may differ by a few nodes on the way, the converged sets and the step counts agree.
    print("Newton histories (active, line search) monolithic:", out[False][3], "block stages:", out[True][3])
    print("GMRES iterations monolithic / block stages:", out[False][2], out[True][2])
    assert abs(out[True][1] - out[False][1]) <= 2
    assert [rows[-1][0] for rows in out[True][3]] == [rows[-1][0] for rows in out[False][3]]


def test_fp32_vcycle_against_fp64(epf):
    """pf_set_multigrid_precision(32): the V-cycle in float (pf_mg_lowp.cuh: smoother operator, Chebyshev steps,
    transfers) returns the FP64 V-cycle's vector to single-precision accuracy on a two-level problem, and the
    KAT-1 time step converges to the same golden energies with (nearly) the same number of GMRES iterations."""
    pf = epf
    from cracks_b200.api import mesh_diameter
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_3d_1.json")))
    out = {}
    for bits in (64, 32):
        mesh = pf.sneddon_mesh(3, 0)
        ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh, kappa_of_h=lambda hh: 0.0))
        ctx.set_multigrid_precision(bits)
        drv = pf.SneddonDriver(ctx, pressure=lambda t: g["prm"]["pressure"], max_no_timesteps=0,
                               newton_lower_bound=g["prm"]["newton_lower_bound"], max_newton=g["prm"]["newton_max_steps"],
                               max_line_search=g["prm"]["line_search_max_steps"], gmres_max_it=300)
        stats = drv.run(mesh_diameter(mesh))
        # one more V-cycle on a fixed vector, linearised at the converged state
        ctx.setup_jacobian()
        v = np.random.default_rng(7).standard_normal(ctx.n_dofs)
        out[bits] = (stats, drv.lin_its, drv.newton_its, ctx.apply_preconditioner(v))
        ctx.close()
    s64, s32 = out[64][0], out[32][0]
    assert s32[0]["crack"] == pytest.approx(g["statistics"][0]["crack"], rel=1e-8)
    assert s32[0]["bulk"] == pytest.approx(g["statistics"][0]["bulk"], rel=1e-7)
    assert s32[0]["bulk"] == pytest.approx(s64[0]["bulk"], rel=1e-7)
    assert out[32][2] == out[64][2]                                  # same Newton history
    # the preconditioner is as good; what the FP32 cycle adds are the iterations of the refinement cycle pf_solve runs
    # when the true residual b - J x shows the rounding of x = M^-1 V y (about 2 per solve)
    assert out[32][1] - out[64][1] <= max(3, out[64][1] // 20) + 3 * out[64][2]
    z64, z32 = out[64][3], out[32][3]
    print("GMRES iterations fp64 / fp32:", out[64][1], out[32][1], " |z32 - z64| / |z64| =",
          np.linalg.norm(z32 - z64) / np.linalg.norm(z64))
    assert np.linalg.norm(z32 - z64) <= 2e-4 * np.linalg.norm(z64)
    assert np.linalg.norm(z32 - z64) > 0                            # and really is another code path


def test_miehe_shear_small(oracle, epf):
    """Miehe shear with the stress split on the 4 x 4 slit mesh, three time steps, against the oracle's run of the
    same mesh (the golden-sized run is part of the GPU suite; here the point is the library's plumbing)"""
    pf = epf
    lam, mu = 121.15e3, 80.77e3
    ref = oracle.MieheRun("miehe shear", 1, 5e-4, lam, mu, 1e3, kappa_of_h=lambda h: 1e-10 * h, d_rhs=1.0, d_mat=1.0,
                          max_no_timesteps=2).run()
    hf = pf.miehe_final_h(1, 0)
    ctx = pf.PhaseFieldContext(pf.miehe_mesh(1), pf.Params(lam, mu, 2.7, 1e-10 * hf, 2 * hf, 0.0))
    ctx.set_krylov_dim(100)
    drv = pf.MieheDriver(ctx, "miehe shear", E=1e3, timestep=5e-4, max_no_timesteps=2, d_rhs=1.0, d_mat=1.0,
                         newton_lower_bound=1e-6, max_newton=100, max_line_search=10, line_search_damping=0.6,
                         gmres_max_it=3000)
    got = drv.run()
    assert len(got) == len(ref) == 3
    for a, b in zip(got, ref):
        for k in ("bulk", "crack", "load"):
            assert a[k] == pytest.approx(b[k], rel=1e-6), (a["step"], k)
    ctx.close()


def test_multigrid_2d_on_the_slit_mesh(epf):
    """pf_set_preconditioner kind 3: geometric multigrid on the unit square with the slit (16 x 16 -> ... -> 2 x 2,
    doubled slit nodes on every level, stress split from the second step on).  Same converged steps as with
    Jacobi-GMRES, a small fraction of the iterations."""
    pf = epf
    lam, mu = 121.15e3, 80.77e3
    out = {}
    for kind in (1, 3):
        hf = pf.miehe_final_h(3, 0)
        ctx = pf.PhaseFieldContext(pf.miehe_mesh(3), pf.Params(lam, mu, 2.7, 1e-10 * hf, 2 * hf, 0.0))
        ctx.set_krylov_dim(300 if kind == 1 else 40)
        ctx.set_preconditioner(kind, 2, 8.0)
        drv = pf.MieheDriver(ctx, "miehe shear", E=1e3, timestep=1e-3, max_no_timesteps=2, d_rhs=1.0, d_mat=1.0,
                             newton_lower_bound=1e-6, max_newton=100, max_line_search=10, line_search_damping=0.6,
                             gmres_max_it=3000)
        out[kind] = (drv.run(), drv.newton_its, drv.lin_its)
        ctx.close()
    for a, b in zip(out[3][0], out[1][0]):
        for k in ("bulk", "crack", "load"):
            assert a[k] == pytest.approx(b[k], rel=1e-9), (a["step"], k)
    assert out[3][1] == out[1][1]
    assert out[3][2] * 4 < out[1][2] and out[3][2] / out[3][1] < 20
    # with kind 1 a 2-D context keeps using Jacobi (the path the golden GPU runs were verified with)


def test_forest_context_kat2_end_to_end(oracle, epf):
    """pf_create_forest and the hanging-node plumbing of pf_api.cu (apply_forest_dev, the fold / distribute hooks)
    through ForestSneddonDriver: the sneddon_2d_1 golden"""
    pf = epf
    from cracks_b200.forest import ForestContext, ForestSneddonDriver
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_2d_1.json")))
    f = ForestSneddonDriver.prerefined_forest()
    h = f.min_cell_diameter
    mu = 1.0 / (2.0 * 1.2)
    ctx = ForestContext(f, pf.Params(0.4 * mu / 0.6, mu, 1.0, 1e-8 * h, 2.0 * h, 0.0))
    ctx.set_krylov_dim(300)
    drv = ForestSneddonDriver(ctx, pressure=lambda t: 1e-3, max_no_timesteps=3, newton_lower_bound=1e-7, max_newton=50,
                              max_line_search=10, gmres_max_it=3000)
    stats = drv.run_on_forest()
    assert ctx.n_dofs == 453 and len(stats) == 4
    for got, ref in zip(stats, g["statistics"]):
        assert got["crack"] == pytest.approx(ref["crack"], rel=1e-8)
        assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-6)
    assert drv.tcv == pytest.approx(g["tcv"], rel=1e-5)
    ctx.close()


def test_forest_context_entry_points(oracle, epf):
    """pf_create_forest on a refined slit forest: residual / apply / diagonal / mass / energy through the C ABI
    against the hanging-node oracle, plus the Miehe-on-forest entry points (pf_set_dirichlet_values, pf_load_cells,
    pf_get_state)"""
    import scipy.sparse as sp
    pf = epf
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import adaptive_oracle as ao
    from cracks_b200.forest import ForestContext, HostForest
    lam, mu = 121.15e3, 80.77e3
    run = ao.AdaptiveMieheRun("miehe shear", 2, 1e-3, lam, mu, 1e3, kappa_of_h=lambda h: 1e-6, cycles=1)
    flagged = [c for c in run.forest.order if run.forest.cell_box(c)[0] >= 0.5]
    run.forest = run.forest.copy(); run.forest.refine(flagged); run._setup_system()
    p, prm = run.p, run.prm
    f = HostForest(2, (2, 2), (0.0, 0.0), (1.0, 1.0), slit=True)
    f.refine_global(2)
    t = f.tables()
    centre_x = t["coords"][t["conn"]][:, :, 0].mean(axis=1)
    f.refine(centre_x - 0.5 * t["level_h"][t["level"], 0] >= 0.5 - 1e-12)
    assert np.array_equal(f.tables()["conn"], p.cells) and f.n_hanging == len(p.hanging) > 0
    ctx = ForestContext(f, pf.Params(lam, mu, prm.G_c, prm.kappa, prm.eps, 0.0))
    rng = np.random.default_rng(21)
    nn = p.n_nodes
    sol = np.zeros((nn, 3)); sol[:, :2] = 1e-3 * rng.standard_normal((nn, 2)); sol[:, 2] = rng.random(nn)
    old = sol.copy(); old[:, 2] = rng.random(nn)
    oo = old.copy(); oo[:, 2] += 0.5 * (rng.random(nn) - 0.5)
    sol, old, oo = (p.distribute_hanging(v.reshape(-1)) for v in (sol, old, oo))
    prm.dt_old, prm.dt_oldold = 1.0, 0.5
    con = p.dirichlet.reshape(nn, 3).copy(); con[:, 2] |= (rng.random(nn) < 0.2) & ~p.is_hanging_node
    con = con.reshape(-1)
    ctx.set_state(ctx.to_block(sol), ctx.to_block(old), ctx.to_block(oo), 1.0, 0.5, False, 0.0)
    cb = ctx.to_block(con.astype(np.uint8)).astype(np.uint8)
    ctx.set_constraints(cb, cb)
    raw = p.raw_residual(sol, old, oo)
    r_total_ref = p.H.T @ raw
    r_pde, r_tot, nrm = ctx.residual()
    assert _relerr(ctx.to_nodal(r_tot), r_total_ref) <= 1e-12
    assert _relerr(ctx.to_nodal(r_pde), np.where(con, 0.0, r_total_ref)) <= 1e-12
    ctx.setup_jacobian()
    free = ~(con | p.is_hanging_dof)
    Cm = p.H @ sp.diags(free.astype(float))
    x = np.where(free, rng.standard_normal(p.n_dofs), 0.0)
    y = np.zeros(p.n_dofs)
    ctx.vmult(y, ctx.to_block(x))
    ref = Cm.T @ (p.raw_jacobian(sol, old, oo) @ (Cm @ x))
    assert _relerr(ctx.to_nodal(y)[free], ref[free]) <= 1e-12
    assert _relerr(ctx.lumped_mass(), p.lumped_mass()) <= 1e-14
    b_ref, c_ref, _ = p.functionals(sol)
    bulk, crack = ctx.energy()
    assert bulk == pytest.approx(b_ref, rel=1e-12) and crack == pytest.approx(c_ref, rel=1e-12)
    # Miehe-on-forest entry points
    top_cells = np.where(p.xy[p.cells[:, 2], 1] == 1.0)[0].astype(np.int64)
    lx, ly = ctx.load_cells(top_cells)
    lx_ref, ly_ref = run.load(sol)
    assert lx == pytest.approx(lx_ref, rel=1e-12) and ly == pytest.approx(ly_ref, rel=1e-12)
    run._time = 0.0125
    expect = sol.copy()
    run.set_initial_bc(expect)
    expect = p.distribute_hanging(expect)
    vals = np.zeros((nn, 3)); vals[run._top, 0] = -0.0125
    ctx.set_constraints(ctx.to_block(p.dirichlet.astype(np.uint8)).astype(np.uint8), None)
    ctx.set_dirichlet_values(ctx.to_block(vals.reshape(-1)))
    assert np.allclose(ctx.to_nodal(ctx.get_state(0)), expect, rtol=0, atol=1e-15)
    assert np.array_equal(ctx.to_nodal(ctx.get_state(1)), old) and np.array_equal(ctx.to_nodal(ctx.get_state(2)), oo)
    ctx.close()


def test_cpp_host_driver_on_the_forest_path(emu_so, tmp_path):
    """cracks_b200/host (C++: .prm surface, FracturePhaseFieldProblem, host forest) linked against the emulated
    build: tests/sneddon_2d_1.prm -- local pre-refinement, hanging nodes, the refinement cycle at the end --
    through the command line, like tests/test_gpu_forest.py::test_kat2_through_the_cli on a GPU."""
    host = os.path.join(ROOT, "cracks_b200", "host")
    exe = os.path.join(HERE, "emu", "cracks_b200_run_emu")
    srcs = [os.path.join(host, f) for f in ("main.cc", "fracture_problem.cc", "parameter_handler.cc", "function_parser.cc",
                                            "forest.cc", "bitmap_function.cc", "vtu_writer.cc")]
    deps = srcs + [os.path.join(host, f) for f in os.listdir(host) if f.endswith(".h")] + [emu_so]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, *srcs, "-L", os.path.join(HERE, "emu"),
                               "-lcracks_b200_emu", "-Wl,-rpath," + os.path.join(HERE, "emu"), "-pthread"])
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_2d_1.json")))
    sections = {"Global parameters": ["Global pre-refinement steps", "Local pre-refinement steps", "Adaptive refinement cycles",
                                      "Max No of timesteps", "Timestep size", "outer solver", "test case", "ref strategy",
                                      "value phase field for refinement"],
                "Problem dependent parameters": ["K reg", "Eps reg", "Gamma penalization", "Pressure", "Fracture toughness G_c",
                                                 "Poisson ratio nu", "E modulus"],
                "Solver parameters": ["Use Direct Inner Solver", "Newton lower bound", "Newton maximum steps",
                                      "Decompose stress in rhs", "Decompose stress in matrix", "Line search maximum steps"]}
    lines = []
    for sec, keys in sections.items():
        lines.append("subsection " + sec)
        if sec == "Global parameters":
            lines += ["  set Dimension = 2", "  set Output directory = " + str(tmp_path / "out")]
        lines += ["  set %s = %s" % (k, g["prm"][k]) for k in keys if k in g["prm"]]
        lines.append("end")
    (tmp_path / "t.prm").write_text("\n".join(lines) + "\n")
    r = subprocess.run([exe, str(tmp_path / "t.prm")], capture_output=True, text=True, timeout=900)
    print(r.stdout[-2500:], r.stderr[-800:])
    assert r.returncode == 0, r.stderr
    assert "Prerefinement step with h= 2.82843" in r.stdout
    assert "DoFs: 302 solid + 151 phase = 453" in r.stdout and "DoFs: 518 solid + 259 phase = 777" in r.stdout
    assert "0\t\t\t1.491639e+01" in r.stdout                          # tests/sneddon_2d_1.output
    assert "Refinement cycle 0" in r.stdout and "TCV: value= 0.0418879" in r.stdout
    # output_results(): the initial condition and the 4 time steps, like the golden's "Write solution 0..4"; the
    # pieces are valid VTK XML with the forest's cells and the nodal fields
    from vtu_reader import read_vtu
    written = [l.strip() for l in r.stdout.splitlines() if l.startswith("Write solution")]
    assert written[:5] == ["Write solution %d" % i for i in range(5)]
    assert "\tas " + str(tmp_path / "out" / "solution_00003.visit") in r.stdout
    v = read_vtu(tmp_path / "out" / "solution_00004.0000.vtu")
    assert (v["n_points"], v["n_cells"]) == (151, 124) and v["point_data"] == ["displacement", "phasefield", "active_set"]
    assert v["connectivity"].shape == (124 * 4,) and set(v["types"]) == {9} and v["offsets"][-1] == 124 * 4
    assert v["phasefield"].min() == 0.0 and 0.9 < v["phasefield"].max() <= 1.0 and 0 < v["active_set"].sum() < 151
    assert np.all(v["displacement"][:, 2] == 0.0) and np.abs(v["displacement"]).max() > 0
    quad = v["points"][v["connectivity"].reshape(-1, 4)]                 # counter-clockwise quads: positive area
    e1, e2 = quad[:, 1] - quad[:, 0], quad[:, 3] - quad[:, 0]
    assert np.all(e1[:, 0] * e2[:, 1] - e1[:, 1] * e2[:, 0] > 0)
    pvd = open(tmp_path / "out" / "solution.pvd").read()
    assert pvd.count("<DataSet") == len(written) and 'file="solution_00004.pvtu"' in pvd
    assert open(tmp_path / "out" / "solution.visit").read().splitlines()[:2] == ["!NBLOCKS 1", "solution_00000.0000.vtu"]
    rows = [l.split() for l in open(tmp_path / "out" / "statistics") if not l.startswith("#")]
    assert len(rows) == 4
    for row, ref in zip(rows, g["statistics"]):
        assert int(row[2]) == 453 and float(row[3]) == pytest.approx(ref["h"], rel=1e-8)
        assert float(row[5]) == pytest.approx(ref["crack"], rel=1e-8)
        assert float(row[4]) == pytest.approx(ref["bulk"], rel=1e-6)


def test_cpp_host_driver_miehe_small(oracle, emu_so, tmp_path):
    """The C++ Miehe branch (slit mesh, split from the second step on, Load x column) on a 4 x 4 mesh against the
    oracle's run; the golden-sized runs are in the GPU suite (tests/test_gpu_host_driver.py)."""
    exe = os.path.join(HERE, "emu", "cracks_b200_run_emu")
    if not os.path.exists(exe):
        pytest.skip("built by test_cpp_host_driver_on_the_forest_path")
    lam, mu = 121.15e3, 80.77e3
    ref = oracle.MieheRun("miehe shear", 1, 5e-4, lam, mu, 1e3, kappa_of_h=lambda h: 1e-10 * h, d_rhs=1.0, d_mat=1.0,
                          max_no_timesteps=2).run()
    (tmp_path / "m.prm").write_text(f"""subsection Global parameters
  set Dimension = 2
  set Global pre-refinement steps = 1
  set Max No of timesteps = 2
  set Timestep size = 5.0e-4
  set test case = miehe shear
  set ref strategy = phase field
  set value phase field for refinement = 0.8
  set Output directory = {tmp_path / 'out'}
end
subsection Problem dependent parameters
  set K reg = 1.0e-10*h
  set Eps reg = 2*h
  set Fracture toughness G_c = 2.7
  set E modulus = 1e+3
  set Poisson ratio nu = 0.2
  set Lame mu = 80.77e+3
  set Lame lambda = 121.15e+3
end
subsection Solver parameters
  set Use Direct Inner Solver = true
  set Newton lower bound = 1.0e-6
  set Newton maximum steps = 100
  set Line search maximum steps = 10
  set Line search damping = 0.6
  set Decompose stress in rhs = 1.0
  set Decompose stress in matrix = 1.0
end
""")
    r = subprocess.run([exe, str(tmp_path / "m.prm")], capture_output=True, text=True, timeout=900)
    print(r.stdout[-1500:], r.stderr[-800:])
    assert r.returncode == 0, r.stderr
    text = open(tmp_path / "out" / "statistics").read()
    assert "# 7: Load x" in text
    # the piece of the last step: 4 x 4 cells with the slit = 25 + 2 doubled nodes, the cell row above the slit
    # is connected to the upper copies
    from vtu_reader import read_vtu
    v = read_vtu(tmp_path / "out" / "solution_00003.0000.vtu")
    assert (v["n_points"], v["n_cells"]) == (27, 16)
    cells = v["connectivity"].reshape(-1, 4)
    assert {25, 26} <= set(cells[8:12].reshape(-1)) and not ({25, 26} & set(cells[:8].reshape(-1)))
    assert np.allclose(v["points"][25:27, :2], [[0.75, 0.5], [1.0, 0.5]])
    rows = [l.split() for l in text.splitlines() if not l.startswith("#")]
    assert len(rows) == len(ref) == 3
    for row, b in zip(rows, ref):
        assert int(row[2]) == b["dofs"]
        for col, k in ((4, "bulk"), (5, "crack"), (6, "load")):
            assert float(row[col]) == pytest.approx(b[k], rel=1e-6), (row[0], k)


def test_cpp_host_driver_miehe_adaptive(emu_so, tmp_path):
    """tests/miehe_shear_1.prm through the C++ command line with --adaptive: predictor-corrector refinement on the
    slit forest (phase-field flags, 2:1 balance across the slit, SolutionTransfer of the three vectors, redo of
    the step) against the reference's golden statistics.  Steps 0-6 by default (the mesh changes in step 6:
    891 -> 918 DoFs); PF_SLOW_TESTS=1 runs all 11 rows."""
    from prm_from_golden import write_prm, read_statistics
    exe = os.path.join(HERE, "emu", "cracks_b200_run_emu")
    if not os.path.exists(exe):
        pytest.skip("built by test_cpp_host_driver_on_the_forest_path")
    g = json.load(open(os.path.join(HERE, "golden", "miehe_shear_1.json")))
    last = 10 if os.environ.get("PF_SLOW_TESTS") == "1" else 6
    write_prm(tmp_path / "a.prm", g["prm"], 2, tmp_path / "out", Max_No_of_timesteps=last)
    r = subprocess.run([exe, str(tmp_path / "a.prm"), "--adaptive"], capture_output=True, text=True, timeout=3000)
    print(r.stdout[-1500:], r.stderr[-800:])
    assert r.returncode == 0, r.stderr
    assert "MESH CHANGED!" in r.stdout
    rows = read_statistics(tmp_path / "out" / "statistics")
    assert len(rows) == last + 1
    for row, ref in zip(rows, g["statistics"]):
        assert int(row[2]) == ref["dofs"] and float(row[3]) == pytest.approx(ref["h"], rel=1e-8)
        for col, k in ((4, "bulk"), (5, "crack"), (6, "load")):
            assert float(row[col]) == pytest.approx(ref[k], rel=2e-7), (row[0], k)


def test_cpp_host_driver_miehe_multigrid_from_64_cells(epf, emu_so, tmp_path):
    """From 64 x 64 cells on the C++ driver preconditions the Miehe runs with the 2-D multigrid (kind 3): two time
    steps of miehe shear at 5 global refinements through the command line = the Python driver with the same
    preconditioner, a dozen GMRES iterations per Newton step."""
    from prm_from_golden import write_prm, read_statistics
    import re
    pf = epf
    exe = os.path.join(HERE, "emu", "cracks_b200_run_emu")
    if not os.path.exists(exe):
        pytest.skip("built by test_cpp_host_driver_on_the_forest_path")
    g = json.load(open(os.path.join(HERE, "golden", "miehe_shear_2.json")))
    write_prm(tmp_path / "m.prm", g["prm"], 2, tmp_path / "out", Max_No_of_timesteps=1,
              exact={"Global pre-refinement steps": 5})
    r = subprocess.run([exe, str(tmp_path / "m.prm"), "--no-output"], capture_output=True, text=True, timeout=900)
    print(r.stdout[-1500:], r.stderr[-800:])
    assert r.returncode == 0, r.stderr
    its = [tuple(map(int, m)) for m in re.findall(r"Newton iterations: (\d+) total linear iterations: (\d+)", r.stdout)]
    assert len(its) == 2 and all(lin <= 20 * newton for newton, lin in its)
    p = g["prm"]
    lam, mu = float(p["Lame lambda"]), float(p["Lame mu"])
    hf = pf.miehe_final_h(5, 0)
    fh = lambda expr: eval(expr, {"h": hf, "pow": pow})
    ctx = pf.PhaseFieldContext(pf.miehe_mesh(5), pf.Params(lam, mu, float(p["Fracture toughness G_c"]), fh(p["K reg"]),
                                                           fh(p["Eps reg"]), 0.0))
    ctx.set_krylov_dim(300)
    ctx.set_preconditioner(3, 2, 8.0)
    drv = pf.MieheDriver(ctx, p["test case"], E=float(p["E modulus"]), timestep=float(p["Timestep size"]), max_no_timesteps=1,
                         d_rhs=float(p["Decompose stress in rhs"]), d_mat=float(p["Decompose stress in matrix"]),
                         newton_lower_bound=float(p["Newton lower bound"]), max_newton=int(p["Newton maximum steps"]),
                         max_line_search=int(p["Line search maximum steps"]), line_search_damping=float(p["Line search damping"]),
                         gmres_max_it=3000)
    ref = drv.run()
    ctx.close()
    rows = read_statistics(tmp_path / "out" / "statistics")
    assert len(rows) == len(ref) == 2 and not os.path.exists(tmp_path / "out" / "solution_00000.0000.vtu")
    for row, b in zip(rows, ref):
        assert int(row[2]) == 12771
        for col, k in ((4, "bulk"), (5, "crack"), (6, "load")):
            # the crack energy of these first steps is 1e-3 of the bulk energy and moves with the Newton stopping point
            # (|r| < 1e-6): an absolute floor on the bulk energy's scale
            assert float(row[col]) == pytest.approx(b[k], rel=1e-6, abs=1e-6 * abs(b["bulk"])), (row[0], k)


def test_forest_hetero_3d_kat5_end_to_end(epf):
    """BASELINE config 5 in small through the library (emulated): octree with edge / face hanging nodes from the
    phase-field pre-refinement, per-cell Lame coefficients (both sets), pressure(time): tests/hetero_3d_1 golden."""
    pf = epf
    from cracks_b200.forest import ForestHeteroDriver
    g = json.load(open(os.path.join(HERE, "golden", "hetero_3d_1.json")))
    field = {tuple(k): e for k, e in zip(g["cell_keys"], g["e_modulus"])}

    def e_of(centres):
        # the fixture is keyed by (level, i, j, k); recover the key from the centre and the cell size
        out = np.empty(centres.shape[0])
        for n, c in enumerate(centres):
            for level in (3, 4):
                h = 10.0 / (1 << level)
                idx = tuple(int(round(v / h - 0.5)) for v in c)
                if abs((idx[0] + 0.5) * h - c[0]) < 1e-9 and (level,) + idx in field:
                    out[n] = field[(level,) + idx]
                    break
            else:
                raise KeyError(c)
        return out

    drv = ForestHeteroDriver(e_of, newton_lower_bound=1e-6, max_newton=20, max_line_search=8, gmres_max_it=3000)
    assert drv.prerefinement[0][1] == g["dofs_before_prerefinement"]
    assert drv.ctx.n_dofs == g["statistics"][0]["dofs"] == 5288
    for got, ref in zip(drv.run(), g["statistics"]):
        assert got["crack"] == pytest.approx(ref["crack"], rel=1e-7)
        assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-6)
    drv.ctx.close()


def test_cpp_host_driver_hetero_3d(emu_so, tmp_path):
    """`test case = multiple het` (BASELINE config 5 in small) through the C++ command line on the emulated build:
    single-tree cube, phase-field pre-refinement, E-modulus field read from the reference's test.pgm by the C++
    BitmapFunction, hanging nodes in 3-D -- tests/hetero_3d_1.mpirun-4.statistics.  Needs the reference tree for
    the bitmap (skipped elsewhere)."""
    if not os.path.exists("/root/reference/test.pgm"):
        pytest.skip("the reference's test.pgm is only present in the build container")
    exe = os.path.join(HERE, "emu", "cracks_b200_run_emu")
    host = os.path.join(ROOT, "cracks_b200", "host")
    srcs = [os.path.join(host, f) for f in ("main.cc", "fracture_problem.cc", "parameter_handler.cc", "function_parser.cc",
                                            "forest.cc", "bitmap_function.cc", "vtu_writer.cc")]
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, *srcs, "-L", os.path.join(HERE, "emu"),
                           "-lcracks_b200_emu", "-Wl,-rpath," + os.path.join(HERE, "emu"), "-pthread"])
    g = json.load(open(os.path.join(HERE, "golden", "hetero_3d_1.json")))
    (tmp_path / "h.prm").write_text(f"""subsection Global parameters
  set Dimension = 3
  set Global pre-refinement steps = 3
  set Local pre-refinement steps = 1
  set Adaptive refinement cycles = 0
  set Max No of timesteps = 1
  set Timestep size = 0.01
  set outer solver = active set
  set test case = multiple het
  set ref strategy = phase field
  set value phase field for refinement = 0.4
  set Output directory = {tmp_path / 'out'}
end
subsection Problem dependent parameters
  set K reg = 0
  set Eps reg = 1.5
  set Pressure = 0 + time *1e3
  set Fracture toughness G_c = 1.0
  set Poisson ratio nu = 0.2
  set E modulus = 1e4
end
subsection Solver parameters
  set Newton lower bound = 1.0e-6
  set Newton maximum steps = 20
  set Line search maximum steps = 8
  set Line search damping = 0.5
end
""")
    r = subprocess.run([exe, str(tmp_path / "h.prm"), "--source-dir", "/root/reference"], capture_output=True, text=True,
                       timeout=900)
    print(r.stdout[-2500:], r.stderr[-800:])
    assert r.returncode == 0, r.stderr
    assert "Cells:\t512" in r.stdout and "DoFs: 2187 solid + 729 phase = 2916" in r.stdout   # tests/hetero_3d_1 output
    assert "Prerefinement step with h= 2.16506" in r.stdout and "DoFs: 3966 solid + 1322 phase = 5288" in r.stdout
    assert "0\t\t\t2.772590e+01" in r.stdout and "0\t\t\t9.870742e+01" in r.stdout
    rows = [l.split() for l in open(tmp_path / "out" / "statistics") if not l.startswith("#")]
    assert len(rows) == 2
    for row, ref in zip(rows, g["statistics"]):
        assert int(row[2]) == 5288 and float(row[3]) == pytest.approx(ref["h"], rel=1e-8)
        assert float(row[5]) == pytest.approx(ref["crack"], rel=1e-7)
        assert float(row[4]) == pytest.approx(ref["bulk"], rel=1e-6)
    # output_results() of `multiple het`: hexahedra of the octree, the emodulus cell field (1 + E, cracks.cc:3179)
    from vtu_reader import read_vtu
    v = read_vtu(tmp_path / "out" / "solution_00002.0000.vtu")
    assert v["n_points"] == 1322 and set(v["types"]) == {12} and v["cell_data"] == ["emodulus", "subdomain"]
    assert v["emodulus"].min() >= 1.0 and v["emodulus"].max() > 2.0 * v["emodulus"].min()
    hexa = v["points"][v["connectivity"].reshape(-1, 8)]
    vol = np.einsum("ci,ci->c", np.cross(hexa[:, 1] - hexa[:, 0], hexa[:, 3] - hexa[:, 0]), hexa[:, 4] - hexa[:, 0])
    assert np.all(vol > 0) and vol.sum() == pytest.approx(1000.0, rel=1e-12)      # [0,10]^3, right-handed cells
