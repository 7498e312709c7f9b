"""The CUDA kernel SOURCES for locally refined meshes, executed on the CPU.

tests/emu/emu_kernels.cc compiles cracks_b200/csrc/{pf_generic,pf_forest,pf_vector,pf_split2d}.cuh with g++
against a shim of <cuda_runtime.h> and runs the thread-per-item kernels one emulated thread at a time, in
the order pf_api.cu launches them on a forest mesh (mark hanging, lumped mass, residual + fold, diagonal +
fold, distribute -> init -> cell kernel -> fold).  That holds the kernel logic of the hanging-node device
path against the oracle -- which reproduces the reference's sneddon_2d_1 and hetero_3d_1 goldens -- without
a GPU.  It is test infrastructure, not a CPU fallback: nothing under cracks_b200/ links it, and the launch
plumbing of pf_api.cu itself still needs its first GPU run (tests/test_gpu_forest.py)."""
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
def emu():
    d = os.path.join(HERE, "emu")
    so = os.path.join(d, "libemu_kernels.so")
    srcs = [os.path.join(d, "emu_kernels.cc"), os.path.join(d, "cuda_shim", "cuda_runtime.h")] + \
           [os.path.join(ROOT, "cracks_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "cracks_b200", "csrc"))
            if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(d, "cuda_shim"),
                               "-o", so, os.path.join(d, "emu_kernels.cc")])
    return C.CDLL(so)


@pytest.fixture(scope="module")
def ao(oracle):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import adaptive_oracle
    return adaptive_oracle


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _run_emulation(emu, p, prm, sol, old, oo, con, x, cell_lame=None):
    """p: oracle AdaptiveProblem (mesh tables + reference results), returns the emulated kernel outputs"""
    nc, nn = p.nc, p.n_nodes
    level_h, inv = np.unique(np.round(p.cell_h, 12), axis=0, return_inverse=True)
    level = np.ascontiguousarray(inv.reshape(-1).astype(np.uint8))
    hang = np.full((len(p.hanging), 5), -1, dtype=np.int64)
    for k, (h, parents) in enumerate(sorted(p.hanging.items())):
        hang[k, 0] = h
        hang[k, 1:1 + len(parents)] = parents
    ct = (prm.dt_old + prm.dt_oldold) / prm.dt_oldold
    pt = np.ascontiguousarray(oo.reshape(nn, nc)[:, -1] + ct * (old.reshape(nn, nc)[:, -1] - oo.reshape(nn, nc)[:, -1]))
    mask = np.zeros(nn, dtype=np.uint8)
    for c in range(nc):
        mask |= (con.reshape(nn, nc)[:, c].astype(np.uint8) << c)
    phys = np.array([prm.lam, prm.mu, prm.G_c, prm.kappa, prm.eps, (prm.alpha_biot - 1.0) * prm.pressure, 1.0])
    out = {k: np.zeros(nn * nc) for k in ("r_total", "r_pde", "diag", "y")}
    out["mass"] = np.zeros(nn)
    cells = np.ascontiguousarray(p.cells)
    emu.emu_forest(C.c_int(p.dim), C.c_longlong(p.n_cells), C.c_longlong(nn), _ptr(cells), _ptr(level),
                   C.c_int(level_h.shape[0]), _ptr(np.ascontiguousarray(level_h)), C.c_longlong(hang.shape[0]), _ptr(hang),
                   _ptr(cell_lame), _ptr(phys), _ptr(sol), _ptr(pt), _ptr(mask), _ptr(x), _ptr(out["r_total"]),
                   _ptr(out["r_pde"]), _ptr(out["diag"]), _ptr(out["y"]), _ptr(out["mass"]))
    return out


def _check(p, prm, out, sol, old, oo, con, x):
    import scipy.sparse as sp
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    raw = p.raw_residual(sol, old, oo)
    r_total_ref = p.H.T @ raw
    assert rel(out["r_total"], r_total_ref) <= 1e-12
    assert rel(out["r_pde"], np.where(con, 0.0, r_total_ref)) <= 1e-12
    assert rel(out["mass"], p.lumped_mass()) <= 1e-14
    free = ~(con | p.is_hanging_dof)
    Cm = p.H @ sp.diags(free.astype(float))
    J = p.raw_jacobian(sol, old, oo)
    ref = Cm.T @ (J @ (Cm @ x))
    assert rel(out["y"][free], ref[free]) <= 1e-12
    assert np.all(out["y"][p.is_hanging_dof] == 0.0)                       # decoupled rows times x_h = 0
    assert np.allclose(out["y"][con & ~p.is_hanging_dof], (out["diag"] * x)[con & ~p.is_hanging_dof])
    assert np.all(out["diag"] > 0)
    # Jacobi diagonal: the raw diagonal on regular dofs plus the folded share of the hanging ones
    d_raw = np.abs(J.diagonal())
    reg = ~p.is_hanging_dof
    assert np.all(out["diag"][reg] >= d_raw[reg] * (1 - 1e-12))


def test_forest_kernels_2d_on_the_kat2_mesh(emu, ao):
    run = ao.AdaptiveSneddonRun()
    p, prm = run.p, run.prm
    rng = np.random.default_rng(2)
    nn = p.n_nodes
    sol = np.zeros((nn, 3)); sol[:, :2] = 1e-2 * rng.standard_normal((nn, 2)); sol[:, 2] = rng.random(nn)
    old = sol.copy(); old[:, 2] = rng.random(nn)
    oo = old.copy(); oo[:, 2] += 0.5 * (rng.random(nn) - 0.5)        # extrapolation leaves [0, 1]: the clamp is exercised
    sol, old, oo = (np.ascontiguousarray(p.distribute_hanging(v.reshape(-1))) for v in (sol, old, oo))
    prm.dt_old, prm.dt_oldold = 1.0, 0.5
    con = p.dirichlet.reshape(nn, 3).copy()
    con[:, 2] |= (rng.random(nn) < 0.25) & ~p.is_hanging_node
    con = con.reshape(-1)
    free = ~(con | p.is_hanging_dof)
    x = np.where(free, rng.standard_normal(p.n_dofs), 0.0)
    x[con] = rng.standard_normal(int(con.sum()))                     # constrained entries must not leak into free rows
    x[p.is_hanging_dof] = 0.0
    out = _run_emulation(emu, p, prm, sol, old, oo, con, x)
    _check(p, prm, out, sol, old, oo, con, x)                         # free rows only see free columns


def test_forest_kernels_3d_hetero_on_the_kat5_mesh(emu, ao):
    g = json.load(open(os.path.join(HERE, "golden", "hetero_3d_1.json")))
    field = {tuple(k): e for k, e in zip(g["cell_keys"], g["e_modulus"])}
    run = ao.HeteroRun3D(lambda cell, centre: field[cell])
    p, prm = run.p, run.prm
    prm.pressure = 10.0
    rng = np.random.default_rng(3)
    nn = p.n_nodes
    sol = np.zeros((nn, 4)); sol[:, :3] = 1e-4 * rng.standard_normal((nn, 3)); sol[:, 3] = rng.random(nn)
    sol = np.ascontiguousarray(p.distribute_hanging(sol.reshape(-1)))
    prm.dt_old = prm.dt_oldold = 0.01
    con = p.dirichlet.reshape(nn, 4).copy()
    con[:, 3] |= (rng.random(nn) < 0.25) & ~p.is_hanging_node
    con = con.reshape(-1)
    free = ~(con | p.is_hanging_dof)
    x = np.where(free, rng.standard_normal(p.n_dofs), 0.0)
    out = _run_emulation(emu, p, prm, sol, sol, sol, con, x, cell_lame=p._lame_a)
    _check(p, prm, out, sol, sol, sol, con, x)
    assert {len(v) for v in p.hanging.values()} == {2, 4}             # edge and face hanging nodes were exercised


def test_active_set_kernel_skips_hanging_nodes(emu, ao):
    run = ao.AdaptiveSneddonRun()
    p = run.p
    rng = np.random.default_rng(4)
    nn = p.n_nodes
    r_total = rng.standard_normal(nn * 3)
    mass = p.lumped_mass()
    old = rng.random(nn * 3)
    sol = old + 0.1 * rng.standard_normal(nn * 3)
    sol0 = sol.copy()
    cycle = np.zeros(nn, dtype=np.int32)
    mask = np.where(p.is_hanging_node, 0x80, 0).astype(np.uint8)
    counts = np.zeros(3, dtype=np.uint64)
    emu.emu_active_set(C.c_int(2), C.c_longlong(nn), C.c_double(10.0), _ptr(r_total), _ptr(mass), _ptr(old), _ptr(sol),
                       _ptr(cycle), _ptr(mask), _ptr(counts))
    crit = r_total.reshape(nn, 3)[:, 2] / mass + 10.0 * (sol0.reshape(nn, 3)[:, 2] - old.reshape(nn, 3)[:, 2])
    active_ref = (~p.is_hanging_node) & (crit > 0.0)                       # cracks.cc:2855-2881
    assert np.array_equal((mask & 4) != 0, active_ref)
    assert np.all((mask & 0x80) == np.where(p.is_hanging_node, 0x80, 0))   # hanging bit untouched
    assert int(counts[0]) == int(active_ref.sum())
    phi, phi_old = sol.reshape(nn, 3)[:, 2], old.reshape(nn, 3)[:, 2]
    assert np.array_equal(phi[active_ref], phi_old[active_ref]) and np.array_equal(phi[~active_ref], sol0.reshape(nn, 3)[~active_ref, 2])


# ---- the tiled 3-D hot kernels, one OS thread per CUDA thread -----------------------------------------

@pytest.fixture(scope="module")
def emu_tiled():
    d = os.path.join(HERE, "emu")
    so = os.path.join(d, "libemu_tiled.so")
    srcs = [os.path.join(d, "emu_tiled.cc"), os.path.join(d, "cuda_shim_block", "cuda_runtime.h")] + \
           [os.path.join(ROOT, "cracks_b200", "csrc", f) for f in os.listdir(os.path.join(ROOT, "cracks_b200", "csrc"))
            if f.endswith(".cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++20", "-fPIC", "-shared", "-pthread", "-I",
                               os.path.join(d, "cuda_shim_block"), "-o", so, os.path.join(d, "emu_tiled.cc")])
    return C.CDLL(so)


def _box_case(oracle, n, h, seed):
    rng = np.random.default_rng(seed)
    lo = tuple(-0.5 * n[d] * h[d] for d in range(3))
    hi = tuple(0.5 * n[d] * h[d] for d in range(3))
    prob = oracle.Problem(3, tuple(n), lo, hi, kappa_of_h=lambda hh: 1e-3, eps_of_h=lambda hh: 2.0 * hh, pressure=1e-3)
    nn = prob.n_nodes
    sol = np.zeros((nn, 4)); sol[:, :3] = 1e-2 * rng.standard_normal((nn, 3)); sol[:, 3] = rng.random(nn)
    old = sol.copy(); old[:, 3] = rng.random(nn)
    oo = old.copy(); oo[:, 3] = old[:, 3] + 0.5 * (rng.random(nn) - 0.5)
    sol, old, oo = sol.reshape(-1), old.reshape(-1), oo.reshape(-1)
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 0.5
    con = prob.dirichlet_mask().reshape(nn, 4); con[rng.random(nn) < 0.2, 3] = 1
    con = np.ascontiguousarray(con.reshape(-1))
    x = rng.standard_normal(nn * 4)
    pt = np.ascontiguousarray(oo.reshape(nn, 4)[:, 3] + 3.0 * (old.reshape(nn, 4)[:, 3] - oo.reshape(nn, 4)[:, 3]))
    mask = np.zeros(nn, dtype=np.uint8)
    for c in range(4):
        mask |= (con.reshape(nn, 4)[:, c].astype(np.uint8) << c)
    phys = np.array([prob.prm.lam, prob.prm.mu, prob.prm.G_c, prob.prm.kappa, prob.prm.eps, -prob.prm.pressure, 1.0])
    return prob, sol, old, oo, con, x, pt, mask, phys


@pytest.mark.parametrize("n,h", [((17, 5, 2), (0.5, 0.5, 0.5)),      # cubic cells: the ISO specialisation, ragged tiles
                                 ((36, 3, 1), (0.25, 0.25, 0.25)),   # two packed-FP32 tiles in x, a partial tile row
                                 ((7, 6, 3), (0.3, 0.45, 0.7))])     # anisotropic cells: the general variant
def test_tiled_apply_kernel_sources_match_the_oracle(oracle, emu_tiled, n, h):
    prob, sol, old, oo, con, x, pt, mask, phys = _box_case(oracle, n, h, seed=sum(n))
    y_ref = prob.apply_jacobian(sol, old, oo, con, x)
    free = con == 0
    diag = np.ones(prob.n_dofs)
    nv, hv = (C.c_int * 3)(*n), (C.c_double * 3)(*h)
    for variant in (16, 3, 19):                                       # v4 (default), v2 and v5 (y-collapse staged)
        y = np.zeros(prob.n_dofs)
        emu_tiled.emu_apply3d(C.c_int(variant), C.c_int(3), nv, hv, _ptr(phys), _ptr(x), _ptr(sol), _ptr(pt), _ptr(mask),
                              _ptr(diag), _ptr(y))
        assert np.max(np.abs(y[free] - y_ref[free])) / np.max(np.abs(y_ref[free])) <= 1e-12, variant
        assert np.array_equal(y[~free], x[~free])                     # constrained rows: diag * x, untouched by the tiles
    if h[0] == h[1] == h[2]:
        # v6 (cubic cells, cached state coefficients): the default Krylov operator, and its FP32 twin
        for variant, tol in ((26, 1e-12), (27, 2e-5)):
            y = np.zeros(prob.n_dofs)
            emu_tiled.emu_apply3d(C.c_int(variant), C.c_int(3), nv, hv, _ptr(phys), _ptr(x), _ptr(sol), _ptr(pt), _ptr(mask),
                                  _ptr(diag), _ptr(y))
            assert np.max(np.abs(y[free] - y_ref[free])) / np.max(np.abs(y_ref[free])) <= tol, variant
            assert np.array_equal(y[~free], x[~free])
    # the 2-point-rule operator of the multigrid smoother: v4 and v2 must agree with each other
    y2 = [np.zeros(prob.n_dofs), np.zeros(prob.n_dofs)]
    for out, variant in zip(y2, (16, 3)):
        emu_tiled.emu_apply3d(C.c_int(variant), C.c_int(2), nv, hv, _ptr(phys), _ptr(x), _ptr(sol), _ptr(pt), _ptr(mask),
                              _ptr(diag), _ptr(out))
    assert np.max(np.abs(y2[0] - y2[1])) / np.max(np.abs(y2[1])) <= 1e-12
    if h[0] == h[1] == h[2]:
        for variant, tol in ((26, 1e-12), (27, 2e-5)):                # v6 with the 2-point rule (FP64; packed FP32): the same operator
            y26 = np.zeros(prob.n_dofs)
            emu_tiled.emu_apply3d(C.c_int(variant), C.c_int(2), nv, hv, _ptr(phys), _ptr(x), _ptr(sol), _ptr(pt), _ptr(mask),
                                  _ptr(diag), _ptr(y26))
            assert np.max(np.abs(y26 - y2[1])) / np.max(np.abs(y2[1])) <= tol, variant
        # the block-diagonal smoother operator (no (phi,u) block): the u rows are those of the coupled operator, the phi
        # rows differ by exactly the (phi,u) block, i.e. agree for a direction without displacement part
        y28 = np.zeros(prob.n_dofs)
        emu_tiled.emu_apply3d(C.c_int(28), C.c_int(2), nv, hv, _ptr(phys), _ptr(x), _ptr(sol), _ptr(pt), _ptr(mask),
                              _ptr(diag), _ptr(y28))
        urow = np.arange(prob.n_dofs) % 4 != 3
        assert np.max(np.abs(y28[urow] - y2[1][urow])) / np.max(np.abs(y2[1][urow])) <= 1e-12
        if (con.reshape(-1, 4)[:, :3] == 0).any():                    # (a one-layer mesh has every displacement dof on a face)
            assert np.max(np.abs(y28[~urow] - y2[1][~urow])) > 1e-8 * np.max(np.abs(y2[1]))
        xp = np.where(urow, 0.0, x)
        ya, yb = np.zeros(prob.n_dofs), np.zeros(prob.n_dofs)
        for out, variant in ((ya, 28), (yb, 26)):
            emu_tiled.emu_apply3d(C.c_int(variant), C.c_int(2), nv, hv, _ptr(phys), _ptr(xp), _ptr(sol), _ptr(pt), _ptr(mask),
                                  _ptr(diag), _ptr(out))
        assert np.max(np.abs(ya - yb)) <= 1e-12 * np.max(np.abs(yb))
    # and it is a different (under-integrated) operator, close to the exact one
    assert 1e-6 < np.max(np.abs(y2[0][free] - y_ref[free])) / np.max(np.abs(y_ref[free])) < 0.5


def test_tiled_residual_kernel_source_matches_the_oracle(oracle, emu_tiled):
    n, h = (18, 5, 2), (0.5, 0.4, 0.3)
    prob, sol, old, oo, con, x, pt, mask, phys = _box_case(oracle, n, h, seed=9)
    _, r_tot_ref = prob.residual(sol, old, oo, con)
    r = np.zeros(prob.n_dofs)
    emu_tiled.emu_residual3d((C.c_int * 3)(*n), (C.c_double * 3)(*h), _ptr(phys), _ptr(sol), _ptr(pt), _ptr(r))
    assert np.max(np.abs(r - r_tot_ref)) / np.max(np.abs(r_tot_ref)) <= 1e-12


@pytest.mark.parametrize("split", [False, True])
def test_generic_2d_kernel_sources_on_the_slit_mesh(oracle, emu, split):
    """The 2-D path the GPU suite covers (tests/test_gpu_miehe.py) once more in the CPU suite: slit
    connectivity, Miehe split and its linearisation, diagonal, lumped mass -- kernel sources vs oracle."""
    lam, mu = 121.15e3, 80.77e3
    n = 8
    prob = oracle.Problem(2, (n, n), (0.0, 0.0), (1.0, 1.0), G_c=2.7, pressure=0.0, kappa_of_h=lambda h: 1e-6,
                          eps_of_h=lambda h: 2.0 * h, slit=True, lame=(lam, mu))
    prob.prm.split, prob.prm.d_rhs, prob.prm.d_mat = int(split), 1.0, 1.0
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 0.5
    rng = np.random.default_rng(8)
    nn = prob.n_nodes
    sol = np.zeros((nn, 3)); sol[:, :2] = 1e-3 * rng.standard_normal((nn, 2)); sol[:, 2] = rng.random(nn)
    old = sol.copy(); old[:, 2] = rng.random(nn)
    oo = old.copy(); oo[:, 2] += 0.5 * (rng.random(nn) - 0.5)
    sol, old, oo = sol.reshape(-1), old.reshape(-1), oo.reshape(-1)
    run = oracle.MieheRun("miehe shear", 2, 5e-4, lam, mu, 1e3)
    con = run.dirichlet.reshape(nn, 3).copy(); con[rng.random(nn) < 0.2, 2] = 1
    con = np.ascontiguousarray(con.reshape(-1))
    x = rng.standard_normal(nn * 3)
    pt = np.ascontiguousarray(oo.reshape(nn, 3)[:, 2] + 3.0 * (old.reshape(nn, 3)[:, 2] - oo.reshape(nn, 3)[:, 2]))
    mask = np.zeros(nn, dtype=np.uint8)
    for c in range(3):
        mask |= (con.reshape(nn, 3)[:, c].astype(np.uint8) << c)
    phys = np.array([lam, mu, 2.7, prob.prm.kappa, prob.prm.eps, 0.0, 1.0, float(split), 1.0, 1.0])
    out = {k: np.zeros(nn * 3) for k in ("r", "diag", "y")}
    mass = np.zeros(nn)
    emu.emu_box2d((C.c_int * 2)(n, n), (C.c_double * 2)(1.0 / n, 1.0 / n), C.c_int(1), _ptr(phys), _ptr(sol), _ptr(pt),
                  _ptr(mask), _ptr(x), _ptr(out["r"]), _ptr(out["diag"]), _ptr(out["y"]), _ptr(mass))
    rel = lambda a, b: np.max(np.abs(a - b)) / np.max(np.abs(b))
    _, r_tot_ref = prob.residual(sol, old, oo, con)
    assert rel(out["r"], r_tot_ref) <= (1e-12 if not split else 1e-11)
    y_ref = prob.apply_jacobian(sol, old, oo, con, x)
    assert rel(out["y"], y_ref) <= (1e-12 if not split else 1e-10)    # constrained rows: the same deal.II-style diagonal
    assert rel(out["diag"], np.abs(prob.jacobian(sol, old, oo, None).diagonal())) <= (1e-12 if not split else 1e-10)
    assert rel(mass, prob.lumped_mass()) <= 1e-14


def test_forest_load_and_dirichlet_value_kernels(emu, ao):
    """k_load_top_forest / k_set_dirichlet_values (Miehe tests on a refined slit forest) vs the adaptive oracle"""
    lam, mu = 121.15e3, 80.77e3
    run = ao.AdaptiveMieheRun("miehe shear", 2, 1e-3, lam, mu, 1e3, cycles=1)
    # refine the cells right of the crack tip once so that the top edge and the mesh carry several levels
    flagged = [c for c in run.forest.order if run.forest.cell_box(c)[0] >= 0.5]
    run.forest = run.forest.copy(); run.forest.refine(flagged); run._setup_system()
    p = run.p
    rng = np.random.default_rng(6)
    sol = np.ascontiguousarray(p.distribute_hanging(rng.standard_normal(p.n_dofs)))
    level_h, inv = np.unique(np.round(p.cell_h, 12), axis=0, return_inverse=True)
    level = np.ascontiguousarray(inv.reshape(-1).astype(np.uint8))
    top = np.ascontiguousarray(np.where(p.xy[p.cells[:, 2], 1] == 1.0)[0].astype(np.int64))
    assert len(set(p.cell_h[top, 0])) == 2                                   # two cell sizes along the top edge
    out = np.zeros(2)
    cells = np.ascontiguousarray(p.cells)
    emu.emu_load_forest(C.c_longlong(p.n_cells), C.c_longlong(p.n_nodes), _ptr(cells), _ptr(level),
                        _ptr(np.ascontiguousarray(level_h)), C.c_double(lam), C.c_double(mu), C.c_longlong(top.shape[0]),
                        _ptr(top), _ptr(sol), _ptr(out))
    lx_ref, ly_ref = run.load(sol)
    assert -out[0] == pytest.approx(lx_ref, rel=1e-12) and out[1] == pytest.approx(ly_ref, rel=1e-12)
    # set_initial_bc(time)
    run._time = 0.0125
    ref = sol.copy()
    run.set_initial_bc(ref)
    vals = np.zeros((p.n_nodes, 3)); vals[run._top, 0] = -0.0125
    mask = np.zeros(p.n_nodes, dtype=np.uint8)
    for c in range(3):
        mask |= (p.dirichlet.reshape(-1, 3)[:, c].astype(np.uint8) << c)
    got = sol.copy()
    emu.emu_set_dirichlet_values(C.c_int(2), C.c_longlong(p.n_nodes), _ptr(mask), _ptr(np.ascontiguousarray(vals.reshape(-1))),
                                 _ptr(got))
    assert np.array_equal(got, ref)
