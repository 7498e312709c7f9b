"""TEST DOUBLE for cracks_b200.forest.ForestContext, backed by the CPU oracle.

The forest drivers (cracks_b200/forest.py) only sequence calls on a context.  Their control flow --
flagging, refinement, solution transfer, step redo, boundary data, load cells -- is host logic that can be
exercised without a GPU by handing them this stand-in, which implements the context's methods with
oracle/adaptive_oracle.py on the very tables the C++ host forest produces.  Used by
tests/test_forest_driver_logic_cpu.py only; the product never imports it."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


class _TableForest:
    def __init__(self, tables, lo, hi):
        self.t, self.lo, self.hi = tables, lo, hi

    def build(self):
        t = self.t
        hanging = {int(r[0]): tuple(int(v) for v in r[1:] if v >= 0) for r in t["hanging"]}
        return t["conn"].copy(), t["level_h"][t["level"]].copy(), t["coords"].copy(), hanging


class _Lib:
    def __init__(self, ctx):
        self.ctx = ctx

    def pf_residual(self, h, a, b, c):
        self.ctx.residual(want_vectors=False)
        return 0

    def pf_restore_old_solution(self, h):
        self.ctx.sol = self.ctx.old.copy()
        return 0


def make_mock(ao, orc):
    """-> a class with ForestContext's constructor signature"""

    class MockForestContext:
        def __init__(self, forest, params, device=0, cell_lame=None, cell_lame_energy=None, dist=None):
            self.forest, self.params = forest, params
            self.tables = forest.tables()
            self.dim, self.nc = forest.dim, forest.dim + 1
            self.prm = orc.Params(params.lam, params.mu, params.G_c, params.kappa, params.eps, 0.0, params.alpha_biot,
                                  1.0, 1.0, 0, 0, 0.0, 0.0)
            lo = tuple(self.tables["coords"].min(axis=0))
            hi = tuple(self.tables["coords"].max(axis=0))
            self.p = ao.AdaptiveProblem(_TableForest(self.tables, lo, hi), self.prm)
            self.n_nodes, self.n_dofs = self.p.n_nodes, self.p.n_dofs
            self.h, self.lib = object(), _Lib(self)
            self.mass = self.p.lumped_mass()
            self.dirichlet = np.zeros(self.n_dofs, dtype=bool)
            self.active = np.zeros(self.n_nodes, dtype=bool)
            self.cycle = np.zeros(self.n_nodes, dtype=np.int32)
            z = np.zeros(self.n_dofs)
            self.sol, self.old, self.oldold = z.copy(), z.copy(), z.copy()

        # layout helpers (same as PhaseFieldContext)
        def to_block(self, nodal):
            a = np.asarray(nodal).reshape(self.n_nodes, self.nc)
            return np.concatenate([a[:, : self.dim].reshape(-1), a[:, self.dim]])

        def to_nodal(self, block):
            b = np.asarray(block)
            out = np.empty((self.n_nodes, self.nc), dtype=b.dtype)
            out[:, : self.dim] = b[: self.n_nodes * self.dim].reshape(self.n_nodes, self.dim)
            out[:, self.dim] = b[self.n_nodes * self.dim:]
            return out.reshape(-1)

        def _check(self, rc):
            assert rc == 0

        def close(self):
            pass

        def set_krylov_dim(self, m):
            pass

        # state
        def set_state(self, sol, old=None, oldold=None, dt_old=1.0, dt_oldold=1.0, use_old=False, pressure=0.0):
            self.sol = self.to_nodal(sol).copy()
            if old is not None:
                self.old = self.to_nodal(old).copy()
            if oldold is not None:
                self.oldold = self.to_nodal(oldold).copy()
            self.set_time_parameters(dt_old, dt_oldold, use_old, pressure)

        def get_state(self, which):
            return self.to_block((self.sol, self.old, self.oldold)[which])

        def set_time_parameters(self, dt_old, dt_oldold, use_old, pressure):
            self.prm.dt_old, self.prm.dt_oldold = dt_old, dt_oldold
            self.prm.use_old_timestep_pf, self.prm.pressure = int(use_old), pressure

        def set_stress_split(self, active, d_rhs, d_mat):
            self.prm.split, self.prm.d_rhs, self.prm.d_mat = int(active), d_rhs, d_mat

        def advance_timestep(self):
            self.oldold, self.old = self.old, self.sol.copy()

        def set_constraints(self, dirichlet_block, active_block):
            if dirichlet_block is not None:
                d = self.to_nodal(dirichlet_block).reshape(-1, self.nc).astype(bool)
                d[:, self.dim] = False
                self.dirichlet = d.reshape(-1)
            if active_block is not None:
                self.active = self.to_nodal(active_block).reshape(-1, self.nc)[:, self.dim].astype(bool)

        def set_dirichlet_values(self, values_block):
            v = self.to_nodal(values_block)
            self.sol[self.dirichlet] = v[self.dirichlet]
            self.sol = self.p.distribute_hanging(self.sol)

        def _constrained(self):
            c = self.dirichlet.reshape(-1, self.nc).copy()
            c[:, self.dim] |= self.active
            return c.reshape(-1)

        # hot path
        def residual(self, want_vectors=True):
            raw = self.p.raw_residual(self.sol, self.old, self.oldold)
            self.r_total = self.p.H.T @ raw
            self.r_pde = np.where(self._constrained(), 0.0, self.r_total)
            return None, None, float(np.linalg.norm(self.r_pde))

        def active_set_reset(self):
            self.active[:] = False
            self.cycle[:] = 0

        def active_set_update(self, c, want_mask=True):
            nc, dim = self.nc, self.dim
            phi, phi_old = self.sol.reshape(-1, nc)[:, dim], self.old.reshape(-1, nc)[:, dim]
            crit = self.r_total.reshape(-1, nc)[:, dim] / self.mass + c * (phi - phi_old)
            new = (~self.p.is_hanging_node) & (~((crit <= 0.0) & (self.cycle < 5)))
            n_cyc = int(np.sum(new & (self.cycle >= 5)))
            phi[new] = phi_old[new]
            self.sol = self.p.distribute_hanging(self.sol)
            self.cycle[self.active & ~new] += 1
            changed = bool(np.any(new != self.active))
            self.active = new
            return None, int(new.sum()), n_cyc, changed

        def setup_jacobian(self):
            self.J = self.p.raw_jacobian(self.sol, self.old, self.oldold)

        def solve(self, tol_rel=1e-8, max_it=200, want_dx=False):
            free = ~(self._constrained() | self.p.is_hanging_dof)
            Cm = self.p.H @ sp.diags(free.astype(float))
            A = (Cm.T @ self.J @ Cm + sp.diags((~free).astype(float))).tocsc()
            self.dx = Cm @ spla.spsolve(A, self.r_pde)
            return None, 1

        def save_solution(self):
            self.saved = self.sol.copy()

        def restore_saved_solution(self):
            self.sol = self.saved.copy()

        def update_solution(self, alpha=1.0):
            self.sol = self.sol + alpha * self.dx

        def scale_update(self, f):
            self.dx = self.dx * f

        def project_phase_field(self):
            phi = self.sol.reshape(-1, self.nc)[:, self.dim]
            np.clip(phi, 0.0, 1.0, out=phi)
            self.sol = self.p.distribute_hanging(self.sol)

        # functionals
        def energy(self):
            b, c, _ = self.p.functionals(self.sol)
            return b, c

        def tcv(self):
            return self.p.functionals(self.sol)[2]

        def timestep_difference(self):
            return float(np.max(np.abs(self.old - self.sol)))

        def phase_field_min(self):
            return float(self.sol.reshape(-1, self.nc)[:, self.dim].min())

        def load_cells(self, cells):
            # compute_load on the listed cells (top edge), same quadrature as the oracle's AdaptiveMieheRun.load
            gq = 0.5 * np.sqrt(3.0 / 5.0)
            xi, w = (0.5 - gq, 0.5, 0.5 + gq), (5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0)
            s = self.sol.reshape(-1, 3)
            lx = ly = 0.0
            for c in cells:
                nodes, (hx, hy) = self.p.cells[c], self.p.cell_h[c]
                for q in range(3):
                    gu = np.zeros((2, 2))
                    for v in range(4):
                        bx, by = v & 1, (v >> 1) & 1
                        Nx, Ny = (xi[q] if bx else 1.0 - xi[q]), (1.0 if by else 0.0)
                        g = np.array([(1.0 if bx else -1.0) / hx * Ny, Nx * (1.0 if by else -1.0) / hy])
                        gu += np.outer(s[nodes[v], :2], g)
                    tr = gu[0, 0] + gu[1, 1]
                    lx += self.prm.mu * (gu[0, 1] + gu[1, 0]) * hx * w[q]
                    ly += (self.prm.lam * tr + 2 * self.prm.mu * gu[1, 1]) * hx * w[q]
            return -lx, ly

    return MockForestContext
