"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs.  Bit-level agreement is not expected for
FP64 sums in a different order; the tolerance is 1e-12 relative l-infinity
(SURVEY.md 8c)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-12


def _mesh(pf, dim, n, h, origin):
    m = pf.Mesh()
    m.dim = dim
    for d in range(3):
        m.n[d] = n[d] if d < dim else 1
        m.h[d] = h[d] if d < dim else 1.0
        m.origin[d] = origin[d] if d < dim else 0.0
    return m


def _case(oracle, pf, dim, n, h, seed, kappa=1e-3, clamp_active=True):
    """Random but physical state on an n[0] x n[1] (x n[2]) box; returns the oracle
    problem, a GPU context with the same state, and the state in nodal layout."""
    rng = np.random.default_rng(seed)
    lo = tuple(-0.5 * n[d] * h[d] for d in range(dim))
    hi = tuple(0.5 * n[d] * h[d] for d in range(dim))
    prob = oracle.Problem(dim, tuple(n), lo, hi, kappa_of_h=lambda hh: kappa, eps_of_h=lambda hh: 2.0 * hh,
                          pressure=1e-3)
    nc = dim + 1
    sol = np.zeros((prob.n_nodes, nc))
    sol[:, :dim] = 1e-2 * rng.standard_normal((prob.n_nodes, dim))
    sol[:, dim] = rng.random(prob.n_nodes)
    old = sol.copy()
    oo = sol.copy()
    old[:, dim] = rng.random(prob.n_nodes)
    # extrapolation leaves [0,1] on a good part of the nodes when clamp_active
    oo[:, dim] = old[:, dim] + (0.5 if clamp_active else 0.05) * (rng.random(prob.n_nodes) - 0.5)
    sol, old, oo = sol.reshape(-1), old.reshape(-1), oo.reshape(-1)
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 0.5
    con = prob.dirichlet_mask().reshape(-1, nc)
    con[rng.random(prob.n_nodes) < 0.2, dim] = 1
    con = np.ascontiguousarray(con.reshape(-1))
    mesh = _mesh(pf, dim, n, h, lo)
    params = pf.Params(prob.prm.lam, prob.prm.mu, prob.prm.G_c, prob.prm.kappa, prob.prm.eps, 0.0)
    ctx = pf.PhaseFieldContext(mesh, params)
    ctx.set_state(ctx.to_block(sol), ctx.to_block(old), ctx.to_block(oo), dt_old=1.0, dt_oldold=0.5,
                  use_old_timestep_pf=False, pressure=prob.pressure)
    cb = ctx.to_block(con).astype(np.uint8)
    ctx.set_constraints(cb, cb)
    return prob, ctx, sol, old, oo, con, rng


def _relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


CASES_3D = [((10, 10, 10), (2.0, 2.0, 2.0)),      # the KAT-1 mesh
            ((16, 4, 2), (0.5, 0.5, 0.5)),        # exactly one tile
            ((19, 7, 5), (0.3, 0.45, 0.7)),       # ragged tiles, anisotropic cells
            ((1, 1, 1), (1.0, 1.0, 1.0)),         # a single cell
            ((33, 9, 3), (0.25, 0.25, 0.25)),
            ((12, 6, 21), (0.4, 0.5, 0.3))]       # >= 16 layers: the host call takes the chunk-pipelined path


@pytest.mark.parametrize("n,h", CASES_3D)
def test_apply_jacobian_3d(oracle, pf, n, h):
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 3, n, h, seed=sum(n))
    ctx.setup_jacobian()
    x = rng.standard_normal(prob.n_dofs)
    y_ref = prob.apply_jacobian(sol, old, oo, con, x)
    y = np.zeros(prob.n_dofs)
    ctx.vmult(y, ctx.to_block(x))
    assert _relerr(ctx.to_nodal(y), y_ref) <= TOL
    # the dimension-generic kernel is an independent second implementation
    ctx.lib.pf_debug_force_generic(ctx.h, 1)
    try:
        y2 = np.zeros(prob.n_dofs)
        ctx.vmult(y2, ctx.to_block(x))
    finally:
        ctx.lib.pf_debug_force_generic(ctx.h, 0)
    assert _relerr(ctx.to_nodal(y2), y_ref) <= TOL
    # the first-generation tiled kernel stays available for A/B measurements
    # cubic cells use a specialisation with the gradient scales folded into constants
    ctx.lib.pf_debug_disable_iso(ctx.h, 1)
    try:
        y4 = np.zeros(prob.n_dofs)
        ctx.vmult(y4, ctx.to_block(x))
    finally:
        ctx.lib.pf_debug_disable_iso(ctx.h, 0)
    assert _relerr(ctx.to_nodal(y4), y_ref) <= TOL
    # earlier generations / tuning experiments of the tiled kernel exist only in a `make TUNING=1` library
    # (1 = first generation, 2..7 = tile shapes of v2, 12..15 = persistent TMA-fed v3, 17..21 = v4 under register
    # caps and v5); the product build answers PF_UNSUPPORTED and ships variant 16 alone
    for variant in (1, 2, 3, 4, 5, 6, 7, 12, 13, 14, 15, 17, 18, 19, 20, 21):
        if ctx.lib.pf_debug_set_variant(ctx.h, variant) != 0:
            continue
        try:
            y3 = np.zeros(prob.n_dofs)
            ctx.vmult(y3, ctx.to_block(x))
        finally:
            ctx.lib.pf_debug_set_variant(ctx.h, 16)
        assert _relerr(ctx.to_nodal(y3), y_ref) <= TOL, variant
    ctx.close()


def test_apply_jacobian_3d_use_old_timestep_pf(oracle, pf):
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 3, (9, 6, 4), (0.5, 0.5, 0.5), seed=11)
    prob.prm.use_old_timestep_pf = 1
    ctx.set_time_parameters(1.0, 0.5, True, prob.pressure)
    ctx.setup_jacobian()
    x = rng.standard_normal(prob.n_dofs)
    y_ref = prob.apply_jacobian(sol, old, oo, con, x)
    y = np.zeros(prob.n_dofs)
    ctx.vmult(y, ctx.to_block(x))
    assert _relerr(ctx.to_nodal(y), y_ref) <= TOL
    ctx.close()


def test_apply_jacobian_unit_vectors_match_csr_columns(oracle, pf):
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 3, (4, 3, 3), (1.0, 0.8, 1.2), seed=2)
    ctx.setup_jacobian()
    J = prob.jacobian(sol, old, oo, con).tocsc()
    scale = abs(J).max()
    for j in rng.choice(prob.n_dofs, 24, replace=False):
        e = np.zeros(prob.n_dofs)
        e[j] = 1.0
        y = np.zeros(prob.n_dofs)
        ctx.vmult(y, ctx.to_block(e))
        col = np.asarray(J[:, j].todense()).reshape(-1)
        assert np.max(np.abs(ctx.to_nodal(y) - col)) <= TOL * scale
    ctx.close()


@pytest.mark.parametrize("n,h", [((10, 10), (2.0, 2.0)), ((13, 7), (0.3, 0.5)), ((1, 1), (1.0, 1.0))])
def test_apply_jacobian_2d(oracle, pf, n, h):
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 2, n, h, seed=sum(n))
    ctx.setup_jacobian()
    x = rng.standard_normal(prob.n_dofs)
    y_ref = prob.apply_jacobian(sol, old, oo, con, x)
    y = np.zeros(prob.n_dofs)
    ctx.vmult(y, ctx.to_block(x))
    assert _relerr(ctx.to_nodal(y), y_ref) <= TOL
    ctx.close()


@pytest.mark.parametrize("dim,n,h", [(3, (10, 10, 10), (2.0, 2.0, 2.0)), (3, (7, 5, 3), (0.3, 0.45, 0.7)),
                                     (2, (13, 7), (0.3, 0.5))])
def test_residual_diag_energy(oracle, pf, dim, n, h):
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, dim, n, h, seed=7 + sum(n))
    r_pde_ref, r_tot_ref = prob.residual(sol, old, oo, con)
    r_pde, r_tot, nrm = ctx.residual()
    assert _relerr(ctx.to_nodal(r_tot), r_tot_ref) <= TOL
    assert _relerr(ctx.to_nodal(r_pde), r_pde_ref) <= TOL
    assert nrm == pytest.approx(np.linalg.norm(r_pde_ref), rel=1e-12)
    # 3-D has a tiled residual kernel; the generic one is the independent second implementation
    ctx.lib.pf_debug_force_generic(ctx.h, 1)
    try:
        r_pde2, r_tot2, nrm2 = ctx.residual()
    finally:
        ctx.lib.pf_debug_force_generic(ctx.h, 0)
    assert _relerr(ctx.to_nodal(r_tot2), r_tot_ref) <= TOL and nrm2 == pytest.approx(nrm, rel=1e-12)
    ctx.setup_jacobian()
    d_ref = prob.jacobian(sol, old, oo, None).diagonal()
    assert np.all(d_ref > 0)
    assert _relerr(ctx.to_nodal(ctx.jacobian_diagonal()), d_ref) <= TOL
    b_ref, c_ref = prob.energy(sol)
    b, c = ctx.energy()
    assert b == pytest.approx(b_ref, rel=1e-12) and c == pytest.approx(c_ref, rel=1e-12)
    assert ctx.tcv() == pytest.approx(prob.tcv(sol), rel=1e-11, abs=1e-15)
    assert np.allclose(ctx.lumped_mass(), prob.lumped_mass(), rtol=1e-15)
    # crack opening displacement on a mesh plane and off the mesh planes
    x_plane = prob.lo[0] + prob.mesh.h[0] * (n[0] // 2)
    v_ref, nf_ref = prob.cod(sol, x_plane)
    v, nf = ctx.cod(x_plane)
    assert nf == nf_ref and nf > 0 and v == pytest.approx(v_ref, rel=1e-11, abs=1e-16)
    assert ctx.cod(x_plane + 0.37 * prob.mesh.h[0])[1] == 0
    ctx.close()


def test_active_set_update(oracle, pf):
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 3, (6, 5, 4), (0.5, 0.5, 0.5), seed=21)
    _, r_tot, _ = ctx.residual()
    r_tot_nodal = ctx.to_nodal(r_tot)
    mass = prob.lumped_mass()
    cycle = np.zeros(prob.n_nodes, dtype=np.int32)
    sol_ref = sol.copy()
    act_ref, cnt_ref, _ = prob.active_set(10.0, r_tot_nodal, mass, old, sol_ref, cycle)
    ctx.active_set_reset()
    act, cnt, ncyc, changed = ctx.active_set_update(10.0)
    assert cnt == cnt_ref and ncyc == 0 and changed
    assert np.array_equal(act, act_ref)
    assert np.array_equal(ctx.to_nodal(ctx.get_solution()), sol_ref)
    # a second update with the same residual changes nothing
    ctx.residual(want_vectors=False)
    ctx.close()


def test_gmres_solves_the_newton_system(oracle, pf):
    import scipy.sparse.linalg as spla
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 3, (6, 6, 6), (1.0, 1.0, 1.0), seed=5, clamp_active=False)
    r_pde_ref, _ = prob.residual(sol, old, oo, con)
    J = prob.jacobian(sol, old, oo, con).tocsc()
    dx_ref = spla.spsolve(J, r_pde_ref)
    dx_ref[con == 1] = 0.0
    ctx.residual(want_vectors=False)
    ctx.setup_jacobian()
    dx, its = ctx.solve(1e-10, 2000, want_dx=True)
    assert 0 < its <= 2000
    assert _relerr(ctx.to_nodal(dx), dx_ref) <= 1e-7
    with pytest.raises(pf.NoConvergence):
        ctx.solve(1e-14, 3)
    ctx.close()


def test_errors_are_loud(pf):
    m = pf.sneddon_mesh(3, 0)
    ctx = pf.PhaseFieldContext(m, pf.sneddon_params(m))
    y = np.zeros(ctx.n_dofs)
    with pytest.raises(pf.PFError):
        ctx.vmult(y, y.copy())          # no pf_setup_jacobian yet
    with pytest.raises(pf.PFError):
        ctx.active_set_update(10.0)     # no residual yet
    ctx.close()


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_preconditioners_give_the_same_newton_update(oracle, pf, kind):
    """Jacobi, multigrid with the under-integrated smoother operator (default) and
    multigrid with the exact operator are only preconditioners: the solution of
    J dx = r must not depend on them (the reference's ML AMG is unpinned too)."""
    import scipy.sparse.linalg as spla
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 3, (16, 16, 16), (1.0, 1.0, 1.0), seed=9, clamp_active=False)
    r_pde_ref, _ = prob.residual(sol, old, oo, con)
    J = prob.jacobian(sol, old, oo, con).tocsc()
    dx_ref = spla.spsolve(J, r_pde_ref)
    dx_ref[con == 1] = 0.0
    ctx.set_preconditioner(kind)
    ctx.residual(want_vectors=False)
    ctx.setup_jacobian()
    dx, its = ctx.solve(1e-10, 2000, want_dx=True)
    assert _relerr(ctx.to_nodal(dx), dx_ref) <= 1e-7
    if kind:
        assert its <= 40          # multigrid: a few dozen at most on this random state
    with pytest.raises(pf.PFError):
        ctx.set_preconditioner(4)        # 3 = multigrid also on 2-D meshes
    with pytest.raises(pf.PFError):
        ctx.set_preconditioner(1, cheb_degree=0)
    ctx.close()


def test_deterministic_mode_is_bit_identical_run_to_run(oracle, pf):
    """pf_set_deterministic: operator, residual and diagonal launched colour by colour.  Two Newton runs (time step 0
    of the Sneddon test at 1 refinement, multigrid V-cycle, about 10 Newton steps with a round-off-determined active-set
    history, SURVEY.md preamble) must repeat each other bit for bit -- residual history, active-set sizes, GMRES
    iteration counts, energies -- and the apply must still match the oracle to 1e-12."""
    from cracks_b200.api import mesh_diameter
    prob, ctx, sol, old, oo, con, rng = _case(oracle, pf, 3, (33, 9, 5), (0.25, 0.25, 0.25), seed=3)
    ctx.set_deterministic(True)
    ctx.setup_jacobian()
    x = rng.standard_normal(prob.n_dofs)
    ys = []
    for _ in range(3):
        y = np.zeros(prob.n_dofs)
        ctx.vmult(y, ctx.to_block(x))
        ys.append(y)
    assert np.array_equal(ys[0], ys[1]) and np.array_equal(ys[0], ys[2])
    assert _relerr(ctx.to_nodal(ys[0]), prob.apply_jacobian(sol, old, oo, con, x)) <= TOL
    ctx.close()
    histories = []
    for _ in range(2):
        mesh = pf.sneddon_mesh(3, 1)
        c = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh))
        c.set_deterministic(True)
        drv = pf.SneddonDriver(c, pressure=lambda t: 1e-3, max_no_timesteps=0, newton_lower_bound=1e-7, max_newton=50,
                               max_line_search=10, gmres_max_it=200)
        st = drv.run(mesh_diameter(mesh))
        histories.append(([(r.n_active, r.residual, r.line_search, r.lin_its) for r in drv.history[0]], st[0]["bulk"], st[0]["crack"]))
        c.close()
    assert histories[0][0] == histories[1][0]
    # the energy functionals add their block sums with atomics (reporting only, never fed back): equal to round-off
    assert histories[0][1] == pytest.approx(histories[1][1], rel=1e-13) and histories[0][2] == pytest.approx(histories[1][2], rel=1e-13)
