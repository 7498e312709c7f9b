"""Generates tests/golden/sneddon_3d_refine{N}_cpu.json: time step 0 of parameters_sneddon_3d.prm at N global
refinements solved by the CPU stand-in of the reference's assembled-matrix Newton path
(oracle/cpu_newton.py: the oracle's loop, CSR Jacobian, Jacobi-GMRES to 1e-8).  N = 3 is the CPU-feasible
BASELINE size (2 125 764 DoF); the GPU path is held against these energies in
tests/test_gpu_golden.py::test_time_step_0_at_refine3_vs_cpu_stand_in (crack energy <= 1e-6).

  python tests/golden/make_sneddon_refine_cpu.py 3        # about 10 minutes on 8 cores, 6 GB
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import cpu_newton
import newton_oracle as orc

refine = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prob = orc.sneddon_3d(refine, kappa_of_h=lambda h: 1e-8 * h)
run = cpu_newton.KrylovSneddonRun(prob, newton_lower_bound=1e-7, max_newton=50, max_line_search=10, max_no_timesteps=0)
t0 = time.perf_counter()
stats = run.run()
dt = time.perf_counter() - t0
log = run.logs[0]
out = {"refine": refine, "n_dofs": prob.n_dofs, "h": prob.hdiam, "kappa": prob.prm.kappa, "eps": prob.prm.eps,
       "pressure": prob.pressure, "newton_lower_bound": 1e-7,
       "initial_newton_residual": log.initial_residual,
       "newton_rows": [list(r) for r in log.rows],
       "statistics": stats, "newton_its": run.newton_its, "linear_its": run.lin_its, "wall_s": dt,
       "cores": orc.lib().pfo_num_threads(), "solver": "Jacobi-GMRES(60), rel. tol 1e-8, on the assembled CSR Jacobian"}
json.dump(out, open(os.path.join(HERE, "sneddon_3d_refine%d_cpu.json" % refine), "w"), indent=1)
print(json.dumps(out)[:600])
