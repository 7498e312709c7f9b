"""Host logic of the forest drivers (cracks_b200/forest.py) without a GPU: ForestMieheDriver and
ForestSneddonDriver are run with a test double of the device context that is backed by the CPU oracle
(tests/mock_forest_context.py), on the tables of the C++ host forest.  What is exercised is exactly what the
device cannot check for them: refinement flags and level cap, 2:1 balance, solution transfer, the redo of a
time step, Dirichlet rows / values incl. the doubled slit nodes, the list of top-edge cells -- against the
reference's goldens for the adaptive cases."""
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


@pytest.fixture()
def mocked(oracle, pf, monkeypatch):
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import adaptive_oracle as ao
    import cracks_b200.forest as forest_mod
    from mock_forest_context import make_mock
    monkeypatch.setattr(forest_mod, "ForestContext", make_mock(ao, oracle))
    return forest_mod


def _driver(pf, forest_mod, g, max_steps=None):
    p = g["prm"]
    num = lambda k: float(p[k])
    fh = lambda expr: (lambda h: eval(expr, {"h": h, "pow": pow}))
    params_of_h = lambda h: pf.Params(num("Lame lambda"), num("Lame mu"), num("Fracture toughness G_c"), fh(p["K reg"])(h),
                                      fh(p["Eps reg"])(h), 0.0)
    return forest_mod.ForestMieheDriver(
        p["test case"], int(p["Global pre-refinement steps"]), params_of_h, E=num("E modulus"), timestep=num("Timestep size"),
        max_no_timesteps=int(p["Max No of timesteps"]) if max_steps is None else max_steps,
        cycles=int(p["Adaptive refinement cycles"]), timestep_2=num("Timestep size to switch to"),
        switch_timestep=int(p["Switch timestep after steps"]), d_rhs=float(p.get("Decompose stress in rhs", 0.0)),
        d_mat=float(p.get("Decompose stress in matrix", 0.0)), refine_threshold=num("value phase field for refinement"),
        newton_lower_bound=num("Newton lower bound"), max_newton=int(p["Newton maximum steps"]),
        max_line_search=int(p["Line search maximum steps"]), line_search_damping=num("Line search damping"))


def test_forest_miehe_driver_follows_miehe_shear_1(pf, mocked):
    g = json.load(open(os.path.join(HERE, "golden", "miehe_shear_1.json")))
    drv = _driver(pf, mocked, g)
    stats = drv.run()
    assert [r["dofs"] for r in stats] == [r["dofs"] for r in g["statistics"]] == [891] * 6 + [918, 984, 1068, 1173, 1506]
    assert sorted(set(drv.redone)) == [6, 7, 8, 9, 10]
    for got, ref in zip(stats, g["statistics"]):
        tol = 2e-8 if got["step"] <= 9 else 1e-5
        for k in ("bulk", "crack", "load"):
            assert got[k] == pytest.approx(ref[k], rel=tol), (got["step"], k)


def test_forest_miehe_driver_follows_the_adaptive_tension_test(pf, mocked):
    g = json.load(open(os.path.join(HERE, "golden", "miehe_tension_adaptive_1.json")))
    drv = _driver(pf, mocked, g, max_steps=27)
    stats = drv.run()
    assert [r["dofs"] for r in stats] == [r["dofs"] for r in g["statistics"][:28]]
    for got, ref in zip(stats, g["statistics"]):
        k = got["step"]
        tol = 2e-8 if k <= 21 else 1e-4 if k <= 26 else 5e-3
        for key in ("bulk", "crack", "load"):
            assert got[key] == pytest.approx(ref[key], rel=tol), (k, key)


def test_forest_sneddon_driver_follows_sneddon_2d_1(pf, mocked):
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_2d_1.json")))
    f = mocked.ForestSneddonDriver.prerefined_forest()
    h = f.min_cell_diameter
    mu = 1.0 / (2.0 * 1.2)
    ctx = mocked.ForestContext(f, pf.Params(0.4 * mu / 0.6, mu, 1.0, 1e-8 * h, 2.0 * h, 0.0))
    drv = mocked.ForestSneddonDriver(ctx, pressure=lambda t: 1e-3, max_no_timesteps=3, newton_lower_bound=1e-7,
                                     max_newton=50, max_line_search=10)
    stats = drv.run_on_forest()
    assert ctx.n_dofs == 453 and len(stats) == 4
    for got, ref, diff in zip(stats, g["statistics"], g["timestep_difference_linfty"]):
        assert got["crack"] == pytest.approx(ref["crack"], rel=1e-8)
        assert got["bulk"] == pytest.approx(ref["bulk"], rel=5e-8)
    assert [s["diff"] for s in stats][:3] == pytest.approx(g["timestep_difference_linfty"][:3], rel=2e-6)
    assert drv.tcv == pytest.approx(g["tcv"], rel=2e-6)
