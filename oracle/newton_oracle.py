"""CPU oracle driver (TEST INFRASTRUCTURE ONLY -- never imported by cracks_b200/).

ctypes binding of oracle/libpf_oracle.so plus a numpy/scipy restatement of the
reference's time loop (cracks.cc:4166-4581, Sneddon branch, uniform meshes) and
of its primal-dual active-set Newton solver (cracks.cc:2780-2994).  The linear
solve uses scipy's sparse direct solver on the oracle's assembled CSR Jacobian:
the reference's GMRES+ML arithmetic is unpinned and only the
converged quantities are compared against the goldens.

Vectors are node-major interleaved, dof = node*(dim+1)+component.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class Mesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int * 3), ("h", C.c_double * 3), ("origin", C.c_double * 3),
                ("slit", C.c_int)]


class Params(C.Structure):
    _fields_ = [
        ("lam", C.c_double), ("mu", C.c_double), ("G_c", C.c_double), ("kappa", C.c_double),
        ("eps", C.c_double), ("pressure", C.c_double), ("alpha_biot", C.c_double),
        ("dt_old", C.c_double), ("dt_oldold", C.c_double), ("use_old_timestep_pf", C.c_int),
        ("split", C.c_int), ("d_rhs", C.c_double), ("d_mat", C.c_double),
    ]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpf_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("pf_oracle.c", "pf_oracle_impl.h", "pf_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        dp = np.ctypeslib.ndpointer(np.float64, flags="C")
        u8 = np.ctypeslib.ndpointer(np.uint8, flags="C")
        lp = np.ctypeslib.ndpointer(np.int64, flags="C")
        ip = np.ctypeslib.ndpointer(np.int32, flags="C")
        MP, PP = C.POINTER(Mesh), C.POINTER(Params)
        for d in ("2d", "3d"):
            f = getattr(_LIB, f"pfo_residual_{d}"); f.restype = None
            f.argtypes = [MP, PP, dp, dp, dp, C.c_void_p, C.c_void_p, dp]
            f = getattr(_LIB, f"pfo_csr_nnz_{d}"); f.restype = C.c_long; f.argtypes = [MP]
            f = getattr(_LIB, f"pfo_csr_pattern_{d}"); f.restype = None; f.argtypes = [MP, lp, ip]
            f = getattr(_LIB, f"pfo_assemble_jacobian_{d}"); f.restype = None
            f.argtypes = [MP, PP, dp, dp, dp, C.c_void_p, lp, ip, dp]
            f = getattr(_LIB, f"pfo_apply_jacobian_{d}"); f.restype = None
            f.argtypes = [MP, PP, dp, dp, dp, C.c_void_p, dp, dp]
            f = getattr(_LIB, f"pfo_lumped_mass_{d}"); f.restype = None; f.argtypes = [MP, dp]
            f = getattr(_LIB, f"pfo_energy_{d}"); f.restype = None
            f.argtypes = [MP, PP, dp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
            f = getattr(_LIB, f"pfo_tcv_{d}"); f.restype = C.c_double; f.argtypes = [MP, dp]
            f = getattr(_LIB, f"pfo_cod_{d}"); f.restype = C.c_double
            f.argtypes = [MP, dp, C.c_double, C.POINTER(C.c_long)]
            f = getattr(_LIB, f"pfo_active_set_{d}"); f.restype = C.c_long
            f.argtypes = [MP, C.c_double, dp, dp, dp, dp, C.c_void_p, u8, C.POINTER(C.c_long)]
            f = getattr(_LIB, f"pfo_spmv_{d}"); f.restype = None
            f.argtypes = [C.c_long, lp, ip, dp, dp, dp]
            f = getattr(_LIB, f"pfo_n_nodes_{d}"); f.restype = C.c_long; f.argtypes = [MP]
            f = getattr(_LIB, f"pfo_cell_nodes_{d}"); f.restype = None; f.argtypes = [MP, lp]
            f = getattr(_LIB, f"pfo_node_coords_{d}"); f.restype = None; f.argtypes = [MP, dp]
        _LIB.pfo_eigen_2x2.restype = None
        _LIB.pfo_eigen_2x2.argtypes = [dp, C.POINTER(C.c_double), C.POINTER(C.c_double), dp]
        _LIB.pfo_decompose_stress_2d.restype = None
        _LIB.pfo_decompose_stress_2d.argtypes = [dp, dp, C.c_double, C.c_double, C.c_int, dp, dp]
        _LIB.pfo_load_2d.restype = None
        _LIB.pfo_load_2d.argtypes = [MP, PP, dp, dp]
        _LIB.pfo_num_threads.restype = C.c_int
        _LIB.pfo_set_num_threads.restype = None
        _LIB.pfo_set_num_threads.argtypes = [C.c_int]
    return _LIB


def _u8ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


@dataclass
class Problem:
    """Uniform box mesh + material/regularisation parameters (Sneddon set-up,
    cracks.cc:1207-1254, 1500-1511, 3820-3892)."""
    dim: int
    n: tuple
    lo: tuple
    hi: tuple
    E: float = 1.0
    nu: float = 0.2
    G_c: float = 1.0
    pressure: float = 1e-3
    kappa_of_h: object = lambda h: 0.0
    eps_of_h: object = lambda h: 2.0 * h
    slit: bool = False              # unit_slit.inp topology (2-D, Miehe tests)
    lame: object = None             # (lambda, mu) given directly (Miehe prms, cracks.cc:1512-1521)
    h_param: object = None          # h used for K reg / Eps reg if not the cell diameter (cracks.cc:3839-3854)
    mesh: Mesh = field(init=False)
    prm: Params = field(init=False)

    def __post_init__(self):
        self.mesh = Mesh()
        self.mesh.dim = self.dim
        for d in range(3):
            self.mesh.n[d] = self.n[d] if d < self.dim else 1
            self.mesh.h[d] = (self.hi[d] - self.lo[d]) / self.n[d] if d < self.dim else 1.0
            self.mesh.origin[d] = self.lo[d] if d < self.dim else 0.0
        self.hdiam = math.sqrt(sum(self.mesh.h[d] ** 2 for d in range(self.dim)))  # cell->diameter()
        mu = self.E / (2.0 * (1 + self.nu))
        lam = (2 * self.nu * mu) / (1.0 - 2 * self.nu)
        if self.lame is not None:
            lam, mu = self.lame
        self.mesh.slit = 1 if self.slit else 0
        hp = self.hdiam if self.h_param is None else self.h_param
        self.prm = Params(lam, mu, self.G_c, self.kappa_of_h(hp), self.eps_of_h(hp),
                          self.pressure, 0.0, 1.0, 1.0, 0, 0, 0.0, 0.0)
        self.nc = self.dim + 1
        self.nnode_dir = tuple(self.n[d] + 1 for d in range(self.dim))
        self.sfx = f"{self.dim}d"
        self.n_nodes = int(getattr(lib(), f"pfo_n_nodes_{self.sfx}")(C.byref(self.mesh)))
        self.n_dofs = self.n_nodes * self.nc
        self.sfx = f"{self.dim}d"

    # -- geometry ----------------------------------------------------------
    def cells(self):
        nv = 1 << self.dim
        out = np.empty((int(np.prod(self.n[: self.dim])), nv), dtype=np.int64)
        getattr(lib(), f"pfo_cell_nodes_{self.sfx}")(C.byref(self.mesh), out)
        return out

    def node_coords(self):
        if self.slit:
            xyz = np.empty((self.n_nodes, self.dim))
            getattr(lib(), f"pfo_node_coords_{self.sfx}")(C.byref(self.mesh), xyz)
            return [np.ascontiguousarray(xyz[:, d]) for d in range(self.dim)]
        axes = [self.lo[d] + self.mesh.h[d] * np.arange(self.nnode_dir[d]) for d in range(self.dim)]
        grids = np.meshgrid(*axes[::-1], indexing="ij")[::-1]   # x fastest
        return [g.reshape(-1) for g in grids]

    def dirichlet_mask(self):
        """u = 0 on all faces (cracks.cc:2575-2583 for 2D, 2686-2694 for 3D); phi free."""
        idx = np.meshgrid(*[np.arange(k) for k in self.nnode_dir[::-1]], indexing="ij")[::-1]
        on_b = np.zeros(self.n_nodes, dtype=bool)
        for d in range(self.dim):
            i = idx[d].reshape(-1)
            on_b |= (i == 0) | (i == self.nnode_dir[d] - 1)
        m = np.zeros((self.n_nodes, self.nc), dtype=np.uint8)
        m[on_b, : self.dim] = 1
        return m.reshape(-1)

    def initial_sneddon(self):
        """InitialValuesSneddon, cracks.cc:381-406."""
        xs = self.node_coords()
        r2 = xs[0] ** 2 if self.dim == 2 else xs[0] ** 2 + xs[2] ** 2
        broken = (r2 <= 1.0) & (np.abs(2.0 * xs[1]) <= 2.0 * self.hdiam)
        sol = np.zeros((self.n_nodes, self.nc))
        sol[:, self.dim] = np.where(broken, 0.0, 1.0)
        return sol.reshape(-1)

    # -- oracle calls ------------------------------------------------------
    def residual(self, sol, old, oldold, constrained):
        r_pde = np.empty(self.n_dofs); r_tot = np.empty(self.n_dofs)
        getattr(lib(), f"pfo_residual_{self.sfx}")(C.byref(self.mesh), C.byref(self.prm), sol, old, oldold,
                                                   _u8ptr(constrained), r_pde.ctypes.data_as(C.c_void_p), r_tot)
        return r_pde, r_tot

    def csr_pattern(self):
        if not hasattr(self, "_pattern") and self.slit:
            # node graph from the cell connectivity (the structured stencil does not know the slit)
            import scipy.sparse as sp
            cells, nv, nc = self.cells(), 1 << self.dim, self.nc
            rows = np.repeat(cells, nv, axis=1).reshape(-1)
            cols = np.tile(cells, (1, nv)).reshape(-1)
            g = sp.coo_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(self.n_nodes,) * 2).tocsr()
            g.sort_indices()
            blk = sp.kron(g, np.ones((nc, nc)), format="csr")
            blk.sort_indices()
            self._pattern = (blk.indptr.astype(np.int64), blk.indices.astype(np.int32))
        if not hasattr(self, "_pattern"):
            nnz = getattr(lib(), f"pfo_csr_nnz_{self.sfx}")(C.byref(self.mesh))
            rowptr = np.empty(self.n_dofs + 1, dtype=np.int64); col = np.empty(nnz, dtype=np.int32)
            getattr(lib(), f"pfo_csr_pattern_{self.sfx}")(C.byref(self.mesh), rowptr, col)
            self._pattern = (rowptr, col)
        return self._pattern

    def jacobian(self, sol, old, oldold, constrained):
        import scipy.sparse as sp
        rowptr, col = self.csr_pattern()
        val = np.empty(col.shape[0])
        getattr(lib(), f"pfo_assemble_jacobian_{self.sfx}")(C.byref(self.mesh), C.byref(self.prm), sol, old, oldold,
                                                            _u8ptr(constrained), rowptr, col, val)
        return sp.csr_matrix((val, col, rowptr), shape=(self.n_dofs, self.n_dofs))

    def apply_jacobian(self, sol, old, oldold, constrained, x):
        y = np.empty(self.n_dofs)
        getattr(lib(), f"pfo_apply_jacobian_{self.sfx}")(C.byref(self.mesh), C.byref(self.prm), sol, old, oldold,
                                                         _u8ptr(constrained), np.ascontiguousarray(x), y)
        return y

    def lumped_mass(self):
        m = np.empty(self.n_nodes)
        getattr(lib(), f"pfo_lumped_mass_{self.sfx}")(C.byref(self.mesh), m)
        return m

    def energy(self, sol):
        b, c = C.c_double(), C.c_double()
        getattr(lib(), f"pfo_energy_{self.sfx}")(C.byref(self.mesh), C.byref(self.prm), sol, C.byref(b), C.byref(c))
        return b.value, c.value

    def tcv(self, sol):
        return getattr(lib(), f"pfo_tcv_{self.sfx}")(C.byref(self.mesh), sol)

    def cod(self, sol, eval_line):
        """(value, n_faces) of compute_cod(eval_line), cracks.cc:3452-3549"""
        nf = C.c_long(0)
        v = getattr(lib(), f"pfo_cod_{self.sfx}")(C.byref(self.mesh), sol, eval_line, C.byref(nf))
        return v, nf.value

    def load(self, sol):
        """compute_load(), cracks.cc:3728-3816: (load_x, load_y) on boundary id 3"""
        out = np.zeros(2)
        lib().pfo_load_2d(C.byref(self.mesh), C.byref(self.prm), sol, out)
        return out

    def active_set(self, c_scale, r_total, mass, old, sol, cycle):
        act = np.zeros(self.n_nodes, dtype=np.uint8)
        ncyc = C.c_long(0)
        cnt = getattr(lib(), f"pfo_active_set_{self.sfx}")(C.byref(self.mesh), c_scale, r_total, mass, old, sol,
                                                           cycle.ctypes.data_as(C.c_void_p), act, C.byref(ncyc))
        return act, cnt, ncyc.value


class NoConvergence(Exception):
    pass


@dataclass
class NewtonLog:
    rows: list = field(default_factory=list)   # (it, n_active, n_cyc, residual, reduction, lsrch)
    initial_residual: float = 0.0


class SneddonRun:
    """Time loop of the reference for `test case = sneddon` on a uniform mesh."""

    def __init__(self, prob: Problem, newton_lower_bound=1e-7, max_newton=60, max_line_search=50,
                 line_search_damping=0.5, timestep=1.0, max_no_timesteps=5):
        self.p = prob
        self.lower = newton_lower_bound
        self.max_newton = max_newton
        self.max_ls = max_line_search
        self.damp = line_search_damping
        self.dt = timestep
        self.max_steps = max_no_timesteps
        self.mass = prob.lumped_mass()
        self.dirichlet = prob.dirichlet_mask()
        self.constrained = self.dirichlet.copy()     # constraints_update after setup_system (cracks.cc:1637-1641)
        self.solution = prob.initial_sneddon()
        self.statistics = []
        self.logs = []

    def linear_solve(self, J, rhs):
        """solve() of the reference (cracks.cc:2744-2777); the oracle's default is a sparse direct solve
        (oracle/cpu_newton.py overrides it with the Krylov solver a benchmark-sized mesh needs)."""
        import scipy.sparse.linalg as spla
        return spla.spsolve(J.tocsc(), rhs)

    def newton_active_set(self, old):
        p = self.p
        nc, dim = p.nc, p.dim
        log = NewtonLog()
        sol = self.solution
        oldold = self.oldold
        r_pde, r_total = p.residual(sol, old, oldold, self.constrained)
        newton_residual = float(np.linalg.norm(r_pde))
        log.initial_residual = newton_residual
        old_newton_residual = newton_residual
        active = np.zeros(p.n_nodes, dtype=np.uint8)
        cycle = np.zeros(p.n_nodes, dtype=np.int32)
        step = 0
        while True:
            active_old = active
            active, n_act, n_cyc = p.active_set(10.0 * p.E, r_total, self.mass, old, sol, cycle)
            cycle[(active_old == 1) & (active == 0)] += 1
            con = self.dirichlet.reshape(-1, nc).copy()
            con[:, dim] = active
            self.constrained = con.reshape(-1)
            changed = int(np.any(active != active_old))
            J = p.jacobian(sol, old, oldold, self.constrained)
            r_pde, _ = p.residual(sol, old, oldold, self.constrained)
            update = self.linear_solve(J, r_pde)
            update[self.constrained == 1] = 0.0
            saved = sol.copy()
            ls = 0
            new_res = 0.0
            while ls < self.max_ls:
                sol += update
                r_pde, r_total = p.residual(sol, old, oldold, self.constrained)
                new_res = float(np.linalg.norm(r_pde))
                if new_res < newton_residual:
                    break
                sol[:] = saved
                update *= self.damp
                ls += 1
            log.rows.append((step + 1, n_act, n_cyc, new_res, new_res / newton_residual, ls))
            old_newton_residual = newton_residual
            newton_residual = new_res
            step += 1
            if newton_residual < self.lower and changed == 0:
                break
            if step >= self.max_newton:
                raise NoConvergence()
        self.logs.append(log)
        return new_res / old_newton_residual

    def run(self):
        p = self.p
        sol = self.solution
        sol.reshape(-1, p.nc)[:, p.dim] = np.clip(sol.reshape(-1, p.nc)[:, p.dim], 0.0, 1.0)
        self.oldold = sol.copy()
        old = sol.copy()
        dt_old = dt_oldold = self.dt
        time = 0.0
        step_no = 0
        self.tcv = None
        while True:
            dt_oldold, dt_old = dt_old, self.dt
            self.oldold = old
            old = sol.copy()
            p.prm.dt_old, p.prm.dt_oldold = dt_old, dt_oldold
            p.prm.use_old_timestep_pf = 0
            time += self.dt
            self.newton_active_set(old)
            phi = sol.reshape(-1, p.nc)[:, p.dim]
            np.clip(phi, 0.0, 1.0, out=phi)
            bulk, crack = p.energy(sol)
            diff = float(np.max(np.abs(old - sol)))
            self.statistics.append(dict(step=step_no, time=time, dofs=p.n_dofs, h=p.hdiam, bulk=bulk, crack=crack,
                                        diff=diff, n_active=self.logs[-1].rows[-1][1]))
            step_no += 1
            if diff < 1.0e-5:
                self.tcv = p.tcv(sol)
                # compute_functional_values(): x = -1.5 + i/256, i = 0..768; lines without faces print nothing
                self.cod = [(x, v) for x in (-1.5 + i / 256.0 for i in range(3 * 256 + 1))
                            for v, nf in [p.cod(sol, x)] if nf > 0]
                break
            if step_no > self.max_steps:
                break
        return self.statistics


class MeshWouldRefine(Exception):
    """refine_mesh() of the reference would change the mesh (cracks.cc:3971-3995, 4108-4133):
    the uniform-mesh oracle stops here."""


class MieheRun(SneddonRun):
    """Time loop of the reference for `test case = miehe tension / miehe shear` on the
    uniformly refined unit_slit.inp mesh (cracks.cc:4166-4581 with the Dirichlet data of
    2584-2625, the load functional 3728-3816 and, for d_mat > 0, the stress split)."""

    def __init__(self, test: str, refine: int, timestep, lam, mu, E, G_c=2.7, kappa_of_h=lambda h: 0.0,
                 eps_of_h=lambda h: 2.0 * h, cycles=0, max_no_timesteps=24, timestep_2=None, switch_timestep=0,
                 newton_lower_bound=1e-6, max_newton=100, max_line_search=10, line_search_damping=0.6,
                 d_rhs=0.0, d_mat=0.0, refine_threshold=0.8):
        assert test in ("miehe tension", "miehe shear")
        n = 2 * 2 ** refine
        # determine_mesh_dependent_parameters: h of the FINAL level (cracks.cc:3839-3854)
        h_final = 0.5 * math.sqrt(2.0) * 2.0 ** (-(refine + cycles))
        prob = Problem(2, (n, n), (0.0, 0.0), (1.0, 1.0), E=E, G_c=G_c, pressure=0.0, kappa_of_h=kappa_of_h,
                       eps_of_h=eps_of_h, slit=True, lame=(lam, mu), h_param=h_final)
        self.test, self.cycles, self.threshold = test, cycles, refine_threshold
        self.h_final = h_final
        self.d_rhs, self.d_mat = d_rhs, d_mat
        self.dt2, self.switch = timestep_2, switch_timestep
        self.p = prob
        self.lower, self.max_newton = newton_lower_bound, max_newton
        self.max_ls, self.damp = max_line_search, line_search_damping
        self.dt, self.max_steps = timestep, max_no_timesteps
        self.mass = prob.lumped_mass()
        self.dirichlet, self._bc_nodes = self._dirichlet()
        self.constrained = self.dirichlet.copy()
        sol = np.zeros((prob.n_nodes, 3))
        sol[:, 2] = 1.0                          # InitialValuesTensionOrShear, cracks.cc:679-691
        self.solution = sol.reshape(-1)
        self.statistics, self.logs = [], []

    def _dirichlet(self):
        """Constrained displacement dofs (cracks.cc:2584-2625) and the top-edge nodes."""
        p = self.p
        x, y = p.node_coords()
        nreg = (p.n[0] + 1) * (p.n[1] + 1)
        m = np.zeros((p.n_nodes, 3), dtype=np.uint8)
        top, bottom = y == 1.0, y == 0.0
        left, right = x == 0.0, x == 1.0
        if self.test == "miehe tension":
            m[bottom, 1] = 1                     # id 2: u_y = 0
            m[top, 0] = m[top, 1] = 1            # id 3: u = (0, t)
        else:
            m[left, 1] = m[right, 1] = 1         # ids 0, 1: u_y = 0
            m[bottom, 0] = m[bottom, 1] = 1      # id 2: u = 0
            m[top, 0] = m[top, 1] = 1            # id 3: u = (-t, 0)
            # id 4 = lower face of the slit: the original (not the doubled) nodes on y = 1/2, x >= 1/2
            lower = (np.arange(p.n_nodes) < nreg) & (y == 0.5) & (x >= 0.5)
            m[lower, 1] = 1
        return m.reshape(-1), np.where(top)[0]

    def set_initial_bc(self, time):
        sol = self.solution.reshape(-1, 3)
        con = self.dirichlet.reshape(-1, 3)
        for c in range(2):
            sol[con[:, c] == 1, c] = 0.0
        if self.test == "miehe tension":
            sol[self._bc_nodes, 1] = time        # BoundaryTensionTest, cracks.cc:780-797
        else:
            sol[self._bc_nodes, 0] = -time       # BoundaryShearTest, cracks.cc:845-861

    def run(self):
        p = self.p
        sol = self.solution
        self.oldold = sol.copy()
        old = sol.copy()
        dt = self.dt
        dt_old = dt_oldold = dt
        time, step_no = 0.0, 0
        while step_no <= self.max_steps:
            if self.switch > 0 and step_no > self.switch:
                dt = self.dt2
            tmp_dt = dt
            dt_oldold, dt_old = dt_old, dt
            self.oldold = old
            old = sol.copy()
            p.prm.dt_old, p.prm.dt_oldold = dt_old, dt_oldold
            p.prm.use_old_timestep_pf = 0
            p.prm.split = 1 if (self.d_mat > 0 and step_no > 0) else 0
            p.prm.d_rhs, p.prm.d_mat = self.d_rhs, self.d_mat
            time += dt
            while True:
                try:
                    self.set_initial_bc(time)
                    self.newton_active_set(old)
                    break
                except NoConvergence:
                    sol[:] = old                 # time-step cut, cracks.cc:4333-4355
                    time -= dt
                    dt /= 10.0
                    time += dt
            phi = sol.reshape(-1, 3)[:, 2]
            np.clip(phi, 0.0, 1.0, out=phi)
            if self.cycles > 0 and phi.min() < self.threshold:
                raise MeshWouldRefine(step_no)
            dt = tmp_dt
            bulk, crack = p.energy(sol)
            load = p.load(sol)
            self.statistics.append(dict(step=step_no, time=time, dofs=p.n_dofs, h=self.h_final, bulk=bulk, crack=crack,
                                        load=float(load[1] if self.test == "miehe tension" else load[0])))
            step_no += 1
        return self.statistics


def sneddon_3d(refine: int = 0, kappa_of_h=lambda h: 0.0) -> Problem:
    n = 10 * 2 ** refine
    return Problem(3, (n, n, n), (-10.0,) * 3, (10.0,) * 3, kappa_of_h=kappa_of_h)


def sneddon_2d(refine: int = 0, kappa_of_h=lambda h: 0.0) -> Problem:
    n = 10 * 2 ** refine
    return Problem(2, (n, n), (-10.0,) * 2, (10.0,) * 2, kappa_of_h=kappa_of_h)


if __name__ == "__main__":
    run = SneddonRun(sneddon_3d(0))
    for row in run.run():
        print(row)
    for lg in run.logs:
        print("r0 = %.6e" % lg.initial_residual)
        for r in lg.rows:
            print("  %d\t%d\t%d\t%.6e\t%.6e\t%d" % r)
    print("TCV", run.tcv)
