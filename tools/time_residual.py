#!/usr/bin/env python
"""CUDA-event timing of pf_residual (device-resident) on Sneddon-3D: tiled vs generic kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cracks_b200 as pf
from cracks_b200.api import mesh_diameter
refine = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mesh = pf.sneddon_mesh(3, refine)
ctx = pf.PhaseFieldContext(mesh, pf.sneddon_params(mesh))
ctx.set_dirichlet_all_faces(); ctx.interpolate_sneddon(mesh_diameter(mesh)); ctx.set_time_parameters(1.0, 1.0, False, 1e-3)
st = torch.cuda.ExternalStream(ctx.stream)
for generic in (0, 1):
    ctx.lib.pf_debug_force_generic(ctx.h, generic)
    for _ in range(3):
        ctx.residual(want_vectors=False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st): e0.record()
    for _ in range(20):
        ctx.residual(want_vectors=False)
    with torch.cuda.stream(st): e1.record()
    ctx.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    nb = 16 * ctx.n_dofs + 17 * ctx.n_nodes
    print(f"{'generic' if generic else 'tiled  '} residual incl. memset/finish/norm: {ms:.3f} ms  ({ctx.n_dofs/ms/1e3:.0f} MDoF/s, {nb/ms/1e6:.0f} GB/s algorithmic)")
ctx.lib.pf_debug_force_generic(ctx.h, 0)
ctx.close()
