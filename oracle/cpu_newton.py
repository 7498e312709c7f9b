"""CPU stand-in for the reference's Newton path at benchmark sizes (TEST / BASELINE INFRASTRUCTURE, like
everything under oracle/: the product never imports it).

`newton_oracle.SneddonRun` solves its linear systems with a sparse direct solver, which a 3-D mesh of
10^6 DoF does not allow.  `KrylovSneddonRun` keeps the oracle's loop (cracks.cc:2780-2994) and the
assembled CSR Jacobian (cracks.cc:2200-2468) and replaces only the solve by what the reference does
(cracks.cc:2744-2777): GMRES on the assembled matrix, relative tolerance 1e-8 -- with a Jacobi
preconditioner, because Trilinos ML is not available here (iteration counts are therefore far above the
reference's; said wherever a number from this file is printed).  The matrix-vector product is the oracle's
OpenMP CSR SpMV, i.e. the reference's `vmult`.

Used by
  * bench.py (`cpu_baseline.newton_its_per_s`): one timed Newton step at refine 3;
  * tests/golden/make_sneddon_refine3_cpu.py: the converged energies of time step 0 at refine 3
    (2.1e6 DoF), the "CPU stand-in at benchmark size" the GPU path is held against.
"""
import time

import numpy as np

import newton_oracle as orc


def jacobi_gmres(spmv, diag, b, tol_rel=1e-8, restart=60, max_it=4000):
    """right-preconditioned restarted GMRES (classical Gram-Schmidt, twice), returns (x, iterations)"""
    n = b.shape[0]
    idiag = 1.0 / diag
    x = np.zeros(n)
    bnorm = float(np.linalg.norm(b))
    if bnorm == 0.0:
        return x, 0
    V = np.empty((restart + 1, n))
    w = np.empty(n)
    total = 0
    r = b.copy()
    while True:
        beta = float(np.linalg.norm(r))
        if beta <= tol_rel * bnorm or total >= max_it:
            return x, total
        V[0] = r / beta
        H = np.zeros((restart + 1, restart))
        g = np.zeros(restart + 1)
        g[0] = beta
        cs, sn = np.zeros(restart), np.zeros(restart)
        k_used = 0
        for k in range(restart):
            spmv(idiag * V[k], w)
            h = V[: k + 1] @ w
            w -= h @ V[: k + 1]
            h2 = V[: k + 1] @ w
            w -= h2 @ V[: k + 1]
            h += h2
            hn = float(np.linalg.norm(w))
            H[: k + 1, k] = h
            H[k + 1, k] = hn
            for i in range(k):
                t = cs[i] * H[i, k] + sn[i] * H[i + 1, k]
                H[i + 1, k] = -sn[i] * H[i, k] + cs[i] * H[i + 1, k]
                H[i, k] = t
            d = np.hypot(H[k, k], H[k + 1, k])
            cs[k], sn[k] = H[k, k] / d, H[k + 1, k] / d
            H[k, k] = d
            H[k + 1, k] = 0.0
            g[k + 1] = -sn[k] * g[k]
            g[k] = cs[k] * g[k]
            total += 1
            k_used = k + 1
            if abs(g[k + 1]) <= tol_rel * bnorm or hn == 0.0 or total >= max_it:
                break
            V[k + 1] = w / hn
        yk = np.linalg.solve(np.triu(H[:k_used, :k_used]), g[:k_used])
        x += idiag * (yk @ V[:k_used])
        spmv(x, w)
        r = b - w


class KrylovSneddonRun(orc.SneddonRun):
    """SneddonRun with the reference's iterative solve (Jacobi instead of ML AMG), timed per phase."""

    def __init__(self, prob, max_newton_steps_total=None, **kw):
        super().__init__(prob, **kw)
        self.lin_its = 0
        self.newton_its = 0
        self.t_assembly = self.t_solve = 0.0
        self.max_total = max_newton_steps_total
        self._pattern_done = False

    def linear_solve(self, J, rhs):
        rowptr, col = self.p.csr_pattern()
        val = J.data
        n = self.p.n_dofs
        spmv = getattr(orc.lib(), f"pfo_spmv_{self.p.sfx}")

        def mv(x, y):
            spmv(n, rowptr, col, val, np.ascontiguousarray(x), y)

        t0 = time.perf_counter()
        dx, its = jacobi_gmres(mv, J.diagonal(), rhs)
        self.t_solve += time.perf_counter() - t0
        self.lin_its += its
        self.newton_its += 1
        if self.max_total is not None and self.newton_its >= self.max_total:
            raise StopAfter(dx)
        return dx


class StopAfter(Exception):
    """raised by KrylovSneddonRun after `max_newton_steps_total` linear solves (bounded timing sample)"""


def time_one_newton_step(refine=3, threads=None):
    """Wall time of ONE active-set Newton step of the assembled-matrix CPU path at `refine`:
    residual assembly + active set + Jacobian assembly + Jacobi-GMRES(1e-8) + one line-search residual."""
    if threads:
        orc.lib().pfo_set_num_threads(int(threads))
    prob = orc.sneddon_3d(refine, kappa_of_h=lambda h: 1e-8 * h)
    prob.csr_pattern()                                    # sparsity pattern: set-up, not part of a Newton step
    run = KrylovSneddonRun(prob, max_newton_steps_total=1, newton_lower_bound=1e-7, max_newton=50, max_line_search=10,
                           max_no_timesteps=0)
    t0 = time.perf_counter()
    try:
        run.run()
    except StopAfter:
        pass
    # the step is cut after the solve: add the one residual evaluation the line search needs at least
    sol = run.solution
    t1 = time.perf_counter()
    prob.residual(sol, sol, sol, run.constrained)
    t_res = time.perf_counter() - t1
    dt = time.perf_counter() - t0
    return {"wall_s": dt, "newton_its_per_s": 1.0 / dt, "linear_its": run.lin_its, "solve_s": run.t_solve,
            "residual_s": t_res, "n_dofs": prob.n_dofs, "cores": orc.lib().pfo_num_threads(),
            "preconditioner": "Jacobi (Trilinos ML is not available in this image)"}


if __name__ == "__main__":
    import json
    import sys
    print(json.dumps(time_one_newton_step(int(sys.argv[1]) if len(sys.argv) > 1 else 2)))
