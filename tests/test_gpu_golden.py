"""End-to-end KAT-1 on the GPU: the device-resident active-set Newton loop must
reproduce tests/sneddon_3d_1.mpirun=4.statistics of the reference."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sneddon_3d_golden(pf):
    from cracks_b200.api import mesh_diameter
    golden = json.load(open(os.path.join(HERE, "golden", "sneddon_3d_1.json")))
    prm = golden["prm"]
    mesh = pf.sneddon_mesh(3, prm["global_refine"])
    params = pf.sneddon_params(mesh, E=prm["E"], nu=prm["nu"], G_c=prm["G_c"], kappa_of_h=lambda h: 0.0)
    ctx = pf.PhaseFieldContext(mesh, params)
    log = []
    drv = pf.SneddonDriver(ctx, E=prm["E"], pressure=lambda t: prm["pressure"], timestep=prm["timestep"],
                           max_no_timesteps=prm["max_no_timesteps"], newton_lower_bound=prm["newton_lower_bound"],
                           max_newton=prm["newton_max_steps"], max_line_search=prm["line_search_max_steps"],
                           line_search_damping=prm["line_search_damping"], gmres_max_it=3000, gmres_tol=1e-8,
                           log=log.append)
    stats = drv.run(mesh_diameter(mesh))
    print("\n".join(log))
    assert len(stats) == len(golden["statistics"])
    assert drv.history[0][0].residual < 1e-5 * golden["initial_newton_residual"][0]
    for got, ref in zip(stats, golden["statistics"]):
        assert got["crack"] == pytest.approx(ref["crack"], rel=1e-8)
    assert stats[0]["bulk"] == pytest.approx(golden["statistics"][0]["bulk"], rel=1e-7)
    for got, ref in zip(stats[1:], golden["statistics"][1:]):
        assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-6)
    assert stats[0]["diff"] == pytest.approx(golden["timestep_difference_linfty"][0], rel=1e-5)
    assert drv.tcv == pytest.approx(golden["tcv"], rel=1e-5)
    assert len(drv.cod) == 1 and drv.cod[0][0] == 0.0
    assert drv.cod[0][1] == pytest.approx(golden["cod"][0][1], rel=1e-5)
    ctx.close()


def test_properties_at_benchmark_size(pf):
    """Size-independent properties on the BASELINE config-3 mesh (refine 3 here to
    keep the test short; bench.py runs refine 4): linearity, rigid translations
    in the kernel of the (u,u) block, symmetry of the (u,u) block."""
    from cracks_b200.api import mesh_diameter
    mesh = pf.sneddon_mesh(3, 3)
    params = pf.sneddon_params(mesh)
    ctx = pf.PhaseFieldContext(mesh, params)
    ctx.interpolate_sneddon(mesh_diameter(mesh))
    ctx.set_time_parameters(1.0, 1.0, False, 1e-3)
    ctx.setup_jacobian()                        # no constraints: mask is empty
    rng = np.random.default_rng(1)
    n, nd, dim = ctx.n_nodes, ctx.n_dofs, 3
    x1, x2 = rng.standard_normal(nd), rng.standard_normal(nd)
    y1, y2, y3 = np.zeros(nd), np.zeros(nd), np.zeros(nd)
    ctx.vmult(y1, x1)
    ctx.vmult(y2, x2)
    ctx.vmult(y3, 2.0 * x1 - 3.0 * x2)
    assert np.max(np.abs(y3 - (2.0 * y1 - 3.0 * y2))) <= 1e-11 * np.max(np.abs(y3))
    # rigid translation: J (c, 0) = 0 without Dirichlet rows
    t = np.zeros(nd)
    t[: n * dim] = np.tile([1.0, -2.0, 0.5], n)
    yt = np.zeros(nd)
    ctx.vmult(yt, t)
    assert np.max(np.abs(yt)) <= 1e-12 * np.max(np.abs(y1))
    # (u,u) block symmetric: <A x1u, x2u> == <x1u, A x2u>
    xu1, xu2 = x1.copy(), x2.copy()
    xu1[n * dim:] = 0
    xu2[n * dim:] = 0
    ctx.vmult(y1, xu1)
    ctx.vmult(y2, xu2)
    a, b = np.dot(y1[: n * dim], xu2[: n * dim]), np.dot(xu1[: n * dim], y2[: n * dim])
    assert a == pytest.approx(b, rel=1e-10)
    ctx.close()
