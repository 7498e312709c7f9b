"""The forest drivers on the reference's adaptive goldens, shared by tests/test_gpu_forest.py (one GPU),
tests/mgpu_forest_check.py (torchrun, several GPUs) and tests/emu/multirank_forest_worker.py (emulated ranks)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def miehe_forest_driver(pf, g, max_steps=None, dist=None, device=0):
    from cracks_b200.forest import ForestMieheDriver
    p = g["prm"]
    num = lambda k: float(p[k])
    fh = lambda expr: (lambda h: eval(expr, {"h": h, "pow": pow}))
    params_of_h = lambda h: pf.Params(num("Lame lambda"), num("Lame mu"), num("Fracture toughness G_c"), fh(p["K reg"])(h),
                                      fh(p["Eps reg"])(h), 0.0)
    return ForestMieheDriver(p["test case"], int(p["Global pre-refinement steps"]), params_of_h, E=num("E modulus"),
                             timestep=num("Timestep size"),
                             max_no_timesteps=int(p["Max No of timesteps"]) if max_steps is None else max_steps,
                             cycles=int(p["Adaptive refinement cycles"]), timestep_2=num("Timestep size to switch to"),
                             switch_timestep=int(p["Switch timestep after steps"]),
                             d_rhs=float(p.get("Decompose stress in rhs", 0.0)), d_mat=float(p.get("Decompose stress in matrix", 0.0)),
                             refine_threshold=num("value phase field for refinement"),
                             newton_lower_bound=num("Newton lower bound"), max_newton=int(p["Newton maximum steps"]),
                             max_line_search=int(p["Line search maximum steps"]), line_search_damping=num("Line search damping"),
                             gmres_max_it=3000, dist=dist, device=device)


def hetero_driver(pf, dist=None, device=0):
    """BASELINE config 5 in small: tests/hetero_3d_1.mpirun-4.statistics (KAT-5) on the 3-D forest path"""
    from cracks_b200.forest import ForestHeteroDriver
    g = json.load(open(os.path.join(HERE, "golden", "hetero_3d_1.json")))
    field = {tuple(k): e for k, e in zip(g["cell_keys"], g["e_modulus"])}

    def e_of(centres):
        out = np.empty(centres.shape[0])
        for n, c in enumerate(centres):
            for level in (3, 4):
                h = 10.0 / (1 << level)
                idx = tuple(int(round(v / h - 0.5)) for v in c)
                if abs((idx[0] + 0.5) * h - c[0]) < 1e-9 and (level,) + idx in field:
                    out[n] = field[(level,) + idx]
                    break
        return out

    drv = ForestHeteroDriver(e_of, newton_lower_bound=1e-6, max_newton=20, max_line_search=8, gmres_max_it=3000, dist=dist,
                             device=device)
    return drv, g
