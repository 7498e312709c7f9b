"""The multigrid V-cycle in FP32 (pf_set_multigrid_precision(32), cracks_b200/csrc/pf_mg_lowp.cuh): the
preconditioner's arithmetic is unpinned, the Krylov operator and every residual stay FP64, so
the KAT-1 golden must be reproduced.  Gating since round 2."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(pf, refine, bits, steps=0, jacobian_bits=64):
    from cracks_b200.api import mesh_diameter
    g = json.load(open(os.path.join(HERE, "golden", "sneddon_3d_1.json")))
    mesh = pf.sneddon_mesh(3, refine)
    ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh, kappa_of_h=(lambda hh: 0.0) if refine == 0 else (lambda hh: 1e-8 * hh)))
    ctx.set_multigrid_precision(bits)
    if jacobian_bits != 64:
        ctx.set_jacobian_precision(jacobian_bits)
    drv = pf.SneddonDriver(ctx, pressure=lambda t: g["prm"]["pressure"], max_no_timesteps=steps,
                           newton_lower_bound=g["prm"]["newton_lower_bound"], max_newton=g["prm"]["newton_max_steps"],
                           max_line_search=g["prm"]["line_search_max_steps"], gmres_max_it=300)
    stats = drv.run(mesh_diameter(mesh))
    ctx.setup_jacobian()
    z = ctx.apply_preconditioner(np.random.default_rng(7).standard_normal(ctx.n_dofs))
    ctx.close()
    return g, stats, drv.newton_its, drv.lin_its, z


def test_kat1_golden_with_the_fp32_vcycle(pf):
    g, stats, _, _, _ = _run(pf, 0, 32, steps=3)
    for got, ref in zip(stats, g["statistics"]):
        assert got["crack"] == pytest.approx(ref["crack"], rel=1e-8)
        # the bulk energy of the later steps moves by 1e-7 with the Newton stopping point (FP64 cycle: 2e-8,
        # FP32 cycle: 2e-7 on the emulated library); tests/test_gpu_host_driver.py uses the same 1e-6
        assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-7 if got["step"] == 0 else 1e-6)


@pytest.mark.parametrize("refine", [1, 2])
def test_fp32_vcycle_is_as_good_a_preconditioner(pf, refine):
    """three / four levels, kappa = 1e-8 h like parameters_sneddon_3d.prm"""
    _, s64, n64, l64, z64 = _run(pf, refine, 64)
    _, s32, n32, l32, z32 = _run(pf, refine, 32)
    print("refine", refine, "Newton", n64, n32, "GMRES", l64, l32, "|dz|/|z|", np.linalg.norm(z32 - z64) / np.linalg.norm(z64))
    assert s32[0]["crack"] == pytest.approx(s64[0]["crack"], rel=1e-9)
    assert s32[0]["bulk"] == pytest.approx(s64[0]["bulk"], rel=1e-6)
    assert n32 == n64
    assert l32 <= 1.15 * l64 + 3
    assert np.linalg.norm(z32 - z64) <= 1e-3 * np.linalg.norm(z64)


def test_kat1_golden_with_the_fp32_jacobian(pf):
    """Inexact Newton (pf_set_jacobian_precision 32 + FP32 V-cycle): the Krylov operator is the packed-FP32 27-point
    apply, the residual that defines the Newton fixed point stays FP64 -- the reference's golden energies of all four
    time steps must come out as with the exact operator (tests/sneddon_3d_1.mpirun=4.statistics)."""
    g, stats, _, _, _ = _run(pf, 0, 32, steps=3, jacobian_bits=32)
    assert len(stats) == 4
    for got, ref in zip(stats, g["statistics"]):
        assert got["crack"] == pytest.approx(ref["crack"], rel=1e-8)
        assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-7 if got["step"] == 0 else 1e-6)


def test_inexact_newton_at_refine2_matches_the_exact_run(pf):
    """275 684 DoF, four multigrid levels: same Newton history length and energies with the FP32 Jacobian, and the
    CPU stand-in's energies (tests/golden/sneddon_3d_refine2_cpu.json) to 1e-6"""
    _, s64, n64, l64, _ = _run(pf, 2, 64)
    _, s32, n32, l32, _ = _run(pf, 2, 32, jacobian_bits=32)
    cpu = json.load(open(os.path.join(HERE, "golden", "sneddon_3d_refine2_cpu.json")))["statistics"][0]
    print("Newton", n64, n32, "GMRES", l64, l32)
    assert n32 <= n64 + 1
    assert s32[0]["crack"] == pytest.approx(s64[0]["crack"], rel=1e-9)
    assert s32[0]["crack"] == pytest.approx(cpu["crack"], rel=1e-6)
    assert s32[0]["bulk"] == pytest.approx(cpu["bulk"], rel=1e-6)
