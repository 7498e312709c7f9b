"""Pins the CPU oracle's 2-D path -- unit_slit mesh, Miehe stress split and its
linearisation, time-dependent Dirichlet data, load functional -- against the
reference's own goldens:
  KAT-4  tests/miehe_shear_2.statistics (+ mpirun=2 variant), split + all 25 steps on a fixed mesh,
  KAT-3  tests/miehe_tension_adaptive_1.statistics rows 0-24 (fixed 891-dof mesh; row 25 refines),
  KAT-6  the six Catch cases of eigen_vectors_and_values (cracks.cc:1740-1919).
The reference's own np1/np2 goldens agree to 9 digits up to row 18 and differ by up
to 5.5e-4 afterwards (crack propagation), so later rows get the looser tolerance."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _load(name):
    return json.load(open(os.path.join(HERE, "golden", name)))


def miehe_run(oracle, g, **over):
    p = g["prm"]
    num = lambda k: float(p[k])
    fh = lambda expr: (lambda h: eval(expr.replace("pow", "pow"), {"h": h, "pow": pow}))
    kw = dict(timestep=num("Timestep size"), lam=num("Lame lambda"), mu=num("Lame mu"), E=num("E modulus"),
              G_c=num("Fracture toughness G_c"), kappa_of_h=fh(p["K reg"]), eps_of_h=fh(p["Eps reg"]),
              cycles=int(p["Adaptive refinement cycles"]), max_no_timesteps=int(p["Max No of timesteps"]),
              timestep_2=num("Timestep size to switch to"), switch_timestep=int(p["Switch timestep after steps"]),
              newton_lower_bound=num("Newton lower bound"), max_newton=int(p["Newton maximum steps"]),
              max_line_search=int(p["Line search maximum steps"]), line_search_damping=num("Line search damping"),
              d_rhs=float(p.get("Decompose stress in rhs", 0.0)), d_mat=float(p.get("Decompose stress in matrix", 0.0)),
              refine_threshold=num("value phase field for refinement"))
    kw.update(over)
    return oracle.MieheRun(p["test case"], int(p["Global pre-refinement steps"]), **kw)


def test_eigen_2x2_catch_cases(oracle):
    for case in _load("eigen_2x2.json")["cases"]:
        e1, e2, P = C.c_double(), C.c_double(), np.zeros(4)
        oracle.lib().pfo_eigen_2x2(np.array(case["m"], dtype=float), C.byref(e1), C.byref(e2), P)
        # Catch's Approx: relative epsilon 1.2e-5 (scaled by the value) plus a small margin
        assert e1.value == pytest.approx(case["e1"], rel=1.2e-5, abs=1e-12), case["name"]
        assert e2.value == pytest.approx(case["e2"], rel=1.2e-5, abs=1e-12), case["name"]
        assert [P[0], P[2]] == pytest.approx(case["v1"], rel=1.2e-5, abs=1e-12), case["name"]
        assert [P[1], P[3]] == pytest.approx(case["v2"], rel=1.2e-5, abs=1e-12), case["name"]


def test_split_is_consistent(oracle):
    """sigma+ + sigma- = sigma, and the derivative branch is the directional derivative of sigma+."""
    rng = np.random.default_rng(3)
    lam, mu = 121.15e3, 80.77e3
    for _ in range(20):
        A = rng.standard_normal((2, 2)); E = 0.5 * (A + A.T)
        B = rng.standard_normal((2, 2)); L = 0.5 * (B + B.T)
        sp, sm = np.zeros(4), np.zeros(4)
        oracle.lib().pfo_decompose_stress_2d(E.reshape(-1).copy(), np.zeros(4), lam, mu, 0, sp, sm)
        sig = lam * np.trace(E) * np.eye(2) + 2 * mu * E
        assert np.allclose((sp + sm).reshape(2, 2), sig, rtol=1e-12, atol=1e-9)
        dp, dm = np.zeros(4), np.zeros(4)
        oracle.lib().pfo_decompose_stress_2d(E.reshape(-1).copy(), L.reshape(-1).copy(), lam, mu, 1, dp, dm)
        t = 1e-6
        sp1, sp0, tmp = np.zeros(4), np.zeros(4), np.zeros(4)
        oracle.lib().pfo_decompose_stress_2d((E + t * L).reshape(-1).copy(), np.zeros(4), lam, mu, 0, sp1, tmp)
        oracle.lib().pfo_decompose_stress_2d((E - t * L).reshape(-1).copy(), np.zeros(4), lam, mu, 0, sp0, tmp)
        assert np.allclose(dp, (sp1 - sp0) / (2 * t), rtol=1e-5, atol=1e-3)


def test_slit_mesh_matches_unit_slit_inp(oracle):
    g = _load("miehe_shear_2.json")
    run = miehe_run(oracle, g)
    row = g["statistics"][0]
    assert run.p.n_dofs == row["dofs"] == 891                       # 594 solid + 297 phase
    assert run.h_final == pytest.approx(row["h"], rel=1e-8)
    cells = run.p.cells()
    x, y = run.p.node_coords()
    # the doubled nodes sit on the slit y = 1/2, x in (1/2, 1]; only the cells above use them
    dup = np.arange(17 * 17, run.p.n_nodes)
    assert np.all(y[dup] == 0.5) and np.all(x[dup] > 0.5)
    users = np.unique(np.where(np.isin(cells, dup))[0])
    assert np.all(users // 16 == 8)


@pytest.fixture(scope="module")
def shear(oracle):
    g = _load("miehe_shear_2.json")
    run = miehe_run(oracle, g)
    run.run()
    return run, g


def test_kat4_miehe_shear_statistics(shear):
    run, g = shear
    assert len(run.statistics) == len(g["statistics"]) == 25
    for got, ref, ref2 in zip(run.statistics, g["statistics"], g["statistics_np2"]):
        # up to row 18 the reference's 1- and 2-rank goldens are identical: full 8-digit parity;
        # afterwards allow the spread the reference shows between its own partitions
        spread = max(abs(ref[k] - ref2[k]) / abs(ref[k]) for k in ("bulk", "crack", "load"))
        tol = 2e-8 if got["step"] <= 18 else max(10 * spread, 1e-4)
        for k in ("bulk", "crack", "load"):
            assert got[k] == pytest.approx(ref[k], rel=tol), (got["step"], k)
        assert got["time"] == pytest.approx(ref["time"], abs=5.1e-5)


def test_kat4_initial_newton_residuals(shear):
    run, g = shear
    for lg, ref in zip(run.logs[:19], g["initial_newton_residual"]):
        assert lg.initial_residual == pytest.approx(ref, rel=2e-6)


def test_kat3_miehe_tension_fixed_mesh_rows(oracle):
    g = _load("miehe_tension_adaptive_1.json")
    run = miehe_run(oracle, g)
    with pytest.raises(oracle.MeshWouldRefine) as exc:
        run.run()
    # the golden's DoF column changes at row 25: that is where refine_mesh() first fires
    first_refined = next(r["step"] for r in g["statistics"] if r["dofs"] != 891)
    assert exc.value.args[0] == first_refined == 25
    assert len(run.statistics) == 25
    for got, ref in zip(run.statistics, g["statistics"]):
        tol = 2e-8 if got["step"] <= 21 else 1e-4
        for k in ("bulk", "crack", "load"):
            assert got[k] == pytest.approx(ref[k], rel=tol), (got["step"], k)
