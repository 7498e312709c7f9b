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
