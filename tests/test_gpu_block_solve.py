"""pf_set_block_solve: the linear solves of the active-set Newton method as a u stage followed by a phi stage.
Block (u,phi) of the reference's Jacobian is identically zero (cracks.cc:2333-2337: the linearised stresses are zeroed
for phi trial functions), so the two stages solve the same system J dx = b as the monolithic GMRES
(cracks.cc:2762-2771); the phi stage evaluates the (phi,phi) block alone (cracks_b200/csrc/pf_apply3d_phi.cuh)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _relerr(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("h", [(0.5, 0.5, 0.5), (0.5, 0.4, 0.3)])
def test_block_restricted_operator_vs_oracle(oracle, pf, h):
    """the operator of each stage (pf_debug_set_block) against the oracle's Jacobian with the other block's dofs
    constrained, <= 1e-12; cubic cells: the phi stage is k_apply3d_phi on the 27-point coefficient records"""
    n = (37, 10, 5)
    rng = np.random.default_rng(21)
    lo = tuple(-0.5 * n[d] * h[d] for d in range(3)); hi = tuple(0.5 * n[d] * h[d] for d in range(3))
    prob = oracle.Problem(3, n, lo, hi, kappa_of_h=lambda hh: 1e-3, eps_of_h=lambda hh: 2.0 * hh, pressure=1e-3)
    nn = prob.n_nodes
    sol = np.zeros((nn, 4)); sol[:, :3] = 1e-2 * rng.standard_normal((nn, 3)); sol[:, 3] = rng.random(nn)
    old = sol.copy(); old[:, 3] = rng.random(nn)
    sol, old = sol.reshape(-1), old.reshape(-1)
    prob.prm.dt_old, prob.prm.dt_oldold = 1.0, 1.0
    con = prob.dirichlet_mask().reshape(nn, 4); con[rng.random(nn) < 0.2, 3] = 1
    con_u = con.copy(); con_u[:, 3] = 1
    con_p = con.copy(); con_p[:, :3] = 1
    con, con_u, con_p = (np.ascontiguousarray(c.reshape(-1)) for c in (con, con_u, con_p))
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


def _run(pf, refine, block, steps, mg_bits=64, jacobian_bits=64):
    from cracks_b200.api import mesh_diameter
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_3d_1.json")))
    mesh = pf.sneddon_mesh(3, refine)
    ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh, kappa_of_h=(lambda hh: 0.0) if refine == 0 else (lambda hh: 1e-8 * hh)))
    ctx.set_multigrid_precision(mg_bits)
    if jacobian_bits != 64:
        ctx.set_jacobian_precision(jacobian_bits)
    ctx.set_block_solve(block)
    drv = pf.SneddonDriver(ctx, pressure=lambda t: g["prm"]["pressure"], max_no_timesteps=steps,
                           newton_lower_bound=g["prm"]["newton_lower_bound"], max_newton=g["prm"]["newton_max_steps"],
                           max_line_search=g["prm"]["line_search_max_steps"], gmres_max_it=300)
    stats = drv.run(mesh_diameter(mesh))
    ctx.close()
    return g, stats, drv


@pytest.mark.parametrize("mg_bits,jacobian_bits", [(64, 64), (32, 32)])
def test_kat1_golden_with_the_block_solve(pf, mg_bits, jacobian_bits):
    """tests/sneddon_3d_1.mpirun=4.statistics, all four time steps, with exact and with inexact (FP32) operators"""
    g, stats, drv = _run(pf, 0, True, 3, mg_bits, jacobian_bits)
    assert len(stats) == 4
    for got, ref in zip(stats, g["statistics"]):
        assert got["crack"] == pytest.approx(ref["crack"], rel=1e-8)
        assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-7 if got["step"] == 0 else 1e-6)
    assert drv.tcv == pytest.approx(g["tcv"], rel=1e-5)


def test_block_solve_against_the_monolithic_solve_refine2(pf):
    """262 144 DoF, kappa = 1e-8 h, two time steps: same energies and Newton step counts as the monolithic GMRES,
    and the phi stages carry the iteration count (the u stage runs only while |b_u| is above its tolerance)"""
    _, s0, d0 = _run(pf, 2, False, 1)
    _, s1, d1 = _run(pf, 2, True, 1)
    print("Newton", d0.newton_its, d1.newton_its, "GMRES", d0.lin_its, d1.lin_its)
    for a, b in zip(s0, s1):
        assert b["crack"] == pytest.approx(a["crack"], rel=1e-9)
        assert b["bulk"] == pytest.approx(a["bulk"], rel=1e-6)
    assert abs(d1.newton_its - d0.newton_its) <= 2
