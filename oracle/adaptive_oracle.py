"""CPU oracle for locally refined meshes with hanging nodes (TEST INFRASTRUCTURE ONLY --
never imported by cracks_b200/).

Restates, in numpy/scipy on top of oracle/libpf_oracle.so's raw cell sums, what the
reference does once `Local pre-refinement steps` / `Adaptive refinement cycles` are
non-zero (2-D):

  * refine_mesh(), strategy `fixed preref sneddon` (cracks.cc:3901-3923) on the forest of
    the box mesh, with p4est's 2:1 balance across faces and corners;
  * DoFTools::make_hanging_node_constraints (1630-1634): an edge midpoint that is a vertex of
    the refined neighbour but not of the coarse cell is constrained to the mean of the edge ends;
  * AffineConstraints::distribute_local_to_global / set_zero / distribute for the two
    constraint sets of the reference: hanging nodes only (system_total_residual, 2446-2456)
    and hanging nodes + Dirichlet + active set (constraints_update, 2909-2911).  In exact
    arithmetic these are the congruence C^T J C, C^T r with the interpolation matrix C;
  * the active-set loop skipping hanging nodes (2855-2857) and re-distributing them (2887-2890);
  * compute_cod on a line x = const of a non-uniform mesh (3452-3549).

Parity status: PINNED by tests/test_oracle_adaptive.py against the reference's golden
tests/sneddon_2d_1.{statistics,output} (KAT-2: 124 cells, 453 DoFs incl. 12 hanging nodes).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

import newton_oracle as orc


class GMesh(C.Structure):
    _fields_ = [("n_cells", C.c_long), ("n_nodes", C.c_long), ("cells", C.c_void_p), ("cell_h", C.c_void_p),
                ("cell_lame", C.c_void_p)]


def _lib():
    lib = orc.lib()
    if not hasattr(lib, "_g_ready"):
        dp = np.ctypeslib.ndpointer(np.float64, flags="C")
        GP, PP = C.POINTER(GMesh), C.POINTER(orc.Params)
        for d in ("2d", "3d"):
            f = getattr(lib, f"pfo_g_residual_{d}"); f.restype = None; f.argtypes = [GP, PP, dp, dp, dp, dp]
            f = getattr(lib, f"pfo_g_cell_matrices_{d}"); f.restype = None; f.argtypes = [GP, PP, dp, dp, dp, dp]
            f = getattr(lib, f"pfo_g_lumped_mass_{d}"); f.restype = None; f.argtypes = [GP, dp]
            f = getattr(lib, f"pfo_g_functionals_{d}"); f.restype = None; f.argtypes = [GP, PP, dp, dp]
        lib._g_ready = True
    return lib


class QuadForest:
    """Forest of quadtrees over an nx x ny box mesh; a cell is (level, i, j) with (i, j) counted in
    cells of its own level.  refine() splits flagged cells and restores the 2:1 balance across
    faces and corners the way p4est does for deal.II's parallel::distributed::Triangulation."""

    def __init__(self, nx, ny, lo, hi, slit=False):
        """slit: the topology of meshes/unit_slit.inp -- the line y = mid, x > mid is an internal boundary;
        the cells above it own doubled nodes, and only the tip (mid, mid) connects the two sides."""
        self.nx, self.ny, self.lo, self.hi, self.slit = nx, ny, lo, hi, slit
        assert not slit or (nx % 2 == 0 and ny % 2 == 0)
        self.cells = {(0, i, j) for j in range(ny) for i in range(nx)}

    def copy(self):
        f = QuadForest(self.nx, self.ny, self.lo, self.hi, self.slit)
        f.cells = set(self.cells)
        return f

    def _connected(self, L, i, j, di, dj):
        """is the level-L position (i + di, j + dj) a face / corner neighbour of (i, j) in the coarse-mesh
        connectivity?  Across the slit only through points with x <= mid (the tip is a shared vertex)."""
        if not self.slit or dj == 0:
            return True
        half_y, half_x = (self.ny << L) // 2, (self.nx << L) // 2
        j2 = j + dj
        if (j < half_y) == (j2 < half_y):
            return True                                    # same side of the slit line
        if di == 0:
            return i <= half_x                             # shared segment [i, i+1] reaches x <= mid
        return (i + 1 if di > 0 else i) <= half_x          # shared corner point

    # -- geometry ---------------------------------------------------------------
    def max_level(self):
        return max(c[0] for c in self.cells)

    def cell_box(self, c):
        L, i, j = c
        hx = (self.hi[0] - self.lo[0]) / (self.nx << L)
        hy = (self.hi[1] - self.lo[1]) / (self.ny << L)
        return self.lo[0] + i * hx, self.lo[1] + j * hy, hx, hy

    def vertices(self, c):
        x, y, hx, hy = self.cell_box(c)
        return [(x, y), (x + hx, y), (x, y + hy), (x + hx, y + hy)]

    # -- refinement -------------------------------------------------------------
    def _leaf_containing(self, L, i, j):
        """the active cell covering cell (L, i, j) or None if outside / finer cells cover it"""
        if i < 0 or j < 0 or i >= (self.nx << L) or j >= (self.ny << L):
            return None
        while L >= 0:
            if (L, i, j) in self.cells:
                return (L, i, j)
            L, i, j = L - 1, i >> 1, j >> 1
        return None

    def refine(self, flagged):
        for c in list(flagged):
            self._split(c)

    def _split(self, c):
        if c not in self.cells:
            return
        L, i, j = c
        # 2:1 balance: before c gets level L + 1 children, no face / corner neighbour may be coarser than L
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                if (di, dj) == (0, 0):
                    continue
                if not self._connected(L, i, j, di, dj):
                    continue
                nb = self._leaf_containing(L, i + di, j + dj)
                if nb is not None and nb[0] < L:
                    self._split(nb)
        self.cells.discard(c)
        for b in (0, 1):
            for a in (0, 1):
                self.cells.add((L + 1, 2 * i + a, 2 * j + b))

    # -- numbering ---------------------------------------------------------------
    def build(self):
        """-> (cells [n][4] node ids, cell_h [n][2], node coordinates [m][2], hanging {node: (a, b)})"""
        Lm = self.max_level()
        order = sorted(self.cells, key=lambda c: (c[2] << (Lm - c[0]), c[1] << (Lm - c[0]), c[0]))
        node_of, coords, cells, hs = {}, [], [], []
        hx0 = (self.hi[0] - self.lo[0]) / (self.nx << Lm)
        hy0 = (self.hi[1] - self.lo[1]) / (self.ny << Lm)

        xm, ym = (self.nx << Lm) // 2, (self.ny << Lm) // 2

        def key(p, upper):
            # nodes on the slit line right of the tip exist twice: flag 1 = the copy of the upper side
            return (p[0], p[1], 1 if (self.slit and upper and p[1] == ym and p[0] > xm) else 0)

        def node(k):
            if k not in node_of:
                node_of[k] = len(coords)
                coords.append((self.lo[0] + k[0] * hx0, self.lo[1] + k[1] * hy0))
            return node_of[k]

        lattice = []
        for (L, i, j) in order:
            s = 1 << (Lm - L)
            upper = j * s >= ym
            pts = [key(p, upper) for p in ((i * s, j * s), ((i + 1) * s, j * s), (i * s, (j + 1) * s),
                                           ((i + 1) * s, (j + 1) * s))]
            lattice.append((pts, upper))
            cells.append([node(k) for k in pts])
            hs.append((s * hx0, s * hy0))
        hanging = {}
        for pts, upper in lattice:
            for a, b in ((0, 1), (2, 3), (0, 2), (1, 3)):       # the four edges
                p, q = pts[a], pts[b]
                if (p[0] + q[0]) % 2 or (p[1] + q[1]) % 2:
                    continue
                mid = key(((p[0] + q[0]) // 2, (p[1] + q[1]) // 2), upper)
                if mid in node_of:
                    hanging[node_of[mid]] = (node_of[p], node_of[q])
        self.order, self.node_of, self.lattice_level = order, node_of, Lm
        return (np.array(cells, dtype=np.int64), np.array(hs, dtype=np.float64), np.array(coords, dtype=np.float64),
                hanging)


class AdaptiveProblem:
    """(u, phi) problem on a forest mesh (QuadForest / OctForest): raw cell sums from the C oracle,
    constraints here.  hanging = {node: (parents...)} with equal weights 1/len(parents)."""

    def __init__(self, forest, prm: orc.Params):
        import scipy.sparse as sp
        self.forest, self.prm = forest, prm
        self.cells, self.cell_h, self.xy, self.hanging = forest.build()
        self.dim = dim = self.cell_h.shape[1]
        self.nc = nc = dim + 1
        self.sfx = f"{dim}d"
        self.n_cells, self.n_nodes = self.cells.shape[0], self.xy.shape[0]
        self.n_dofs = self.n_nodes * nc
        self.cell_lame = None                                # set_cell_lame() for heterogeneous materials
        self.gm = GMesh(self.n_cells, self.n_nodes, self.cells.ctypes.data, self.cell_h.ctypes.data, None)
        self.h_min = float(np.min(np.sqrt((self.cell_h ** 2).sum(axis=1))))
        # H: hanging-node interpolation (identity on regular dofs), cracks.cc:1630-1634
        rows, cols, vals = [], [], []
        is_h = np.zeros(self.n_nodes, dtype=bool)
        for h in self.hanging:
            is_h[h] = True
        for n in range(self.n_nodes):
            for c in range(nc):
                if is_h[n]:
                    parents = self.hanging[n]
                    assert not any(is_h[q] for q in parents)
                    rows += [n * nc + c] * len(parents)
                    cols += [q * nc + c for q in parents]
                    vals += [1.0 / len(parents)] * len(parents)
                else:
                    rows.append(n * nc + c); cols.append(n * nc + c); vals.append(1.0)
        self.H = sp.csr_matrix((vals, (rows, cols)), shape=(self.n_dofs,) * 2)
        self.is_hanging_node = is_h
        self.is_hanging_dof = np.repeat(is_h, nc)
        on_b = np.zeros(self.n_nodes, dtype=bool)
        for d in range(dim):
            on_b |= (self.xy[:, d] == forest.lo[d]) | (self.xy[:, d] == forest.hi[d])
        m = np.zeros((self.n_nodes, nc), dtype=bool)
        m[on_b, :dim] = True                                # u = 0 on every face, cracks.cc:2575-2583, 2686-2694
        self.dirichlet = m.reshape(-1)
        # dof indices of the cell matrices
        ndpc = self.cells.shape[1] * nc
        self.ndpc = ndpc
        dofs = (self.cells[:, :, None] * nc + np.arange(nc)[None, None, :]).reshape(self.n_cells, ndpc)
        self._rows = np.repeat(dofs, ndpc, axis=1).reshape(-1)    # row = test function j
        self._cols = np.tile(dofs, (1, ndpc)).reshape(-1)

    def set_cell_lame(self, lame_assembly, lame_energy):
        """per-cell (lambda, mu): the values assemble_system uses and the ones compute_energy uses"""
        self._lame_a = np.ascontiguousarray(lame_assembly, dtype=np.float64)
        self._lame_e = np.ascontiguousarray(lame_energy, dtype=np.float64)
        self.gm.cell_lame = self._lame_a.ctypes.data

    def raw_residual(self, sol, old, oldold):
        r = np.empty(self.n_dofs)
        getattr(_lib(), f"pfo_g_residual_{self.sfx}")(C.byref(self.gm), C.byref(self.prm), sol, old, oldold, r)
        return r

    def raw_jacobian(self, sol, old, oldold):
        import scipy.sparse as sp
        mats = np.empty(self.n_cells * self.ndpc * self.ndpc)
        getattr(_lib(), f"pfo_g_cell_matrices_{self.sfx}")(C.byref(self.gm), C.byref(self.prm), sol, old, oldold, mats)
        return sp.coo_matrix((mats, (self._rows, self._cols)), shape=(self.n_dofs,) * 2).tocsr()

    def lumped_mass(self):
        m = np.empty(self.n_nodes)
        getattr(_lib(), f"pfo_g_lumped_mass_{self.sfx}")(C.byref(self.gm), m)
        return m

    def functionals(self, sol):
        out = np.zeros(3)
        if self.cell_lame is None and hasattr(self, "_lame_e"):
            self.gm.cell_lame = self._lame_e.ctypes.data      # compute_energy's coefficients (cracks.cc:3651)
        getattr(_lib(), f"pfo_g_functionals_{self.sfx}")(C.byref(self.gm), C.byref(self.prm), sol, out)
        if hasattr(self, "_lame_a"):
            self.gm.cell_lame = self._lame_a.ctypes.data
        return out                                            # bulk, crack, tcv

    def distribute_hanging(self, v):
        return self.H @ v

    def cod(self, sol, eval_line):
        """compute_cod(eval_line), cracks.cc:3452-3549, on a mesh whose cells differ in size"""
        gq = 0.5 * math.sqrt(3.0 / 5.0)
        xi, w = (0.5 - gq, 0.5, 0.5 + gq), (5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0)
        s = sol.reshape(-1, 3)
        total, faces = 0.0, 0
        for c in range(self.n_cells):
            nodes, (hx, hy) = self.cells[c], self.cell_h[c]
            x0 = self.xy[nodes[0], 0]
            for side in (0, 1):
                fx = x0 + side * hx
                if not (eval_line - 1e-8 < fx < eval_line + 1e-8):
                    continue
                faces += 1
                for q in range(3):
                    pt = (float(side), xi[q])
                    u, g = np.zeros(2), np.zeros(2)
                    for v in range(4):
                        bx, by = v & 1, (v >> 1) & 1
                        Nx = pt[0] if bx else 1.0 - pt[0]
                        Ny = pt[1] if by else 1.0 - pt[1]
                        u += Nx * Ny * s[nodes[v], :2]
                        g += np.array([(1.0 if bx else -1.0) / hx * Ny, Nx * (1.0 if by else -1.0) / hy]) * s[nodes[v], 2]
                    total += 0.5 * float(u @ g) * hy * w[q]
        return total / 2.0, faces


def flag_fixed_preref_sneddon(forest: QuadForest):
    """cells with a vertex in [-2.5, 2.5] x [-1.25, 1.25] (cracks.cc:3901-3923)"""
    return [c for c in forest.cells
            if any(-2.5 <= x <= 2.5 and -1.25 <= y <= 1.25 for x, y in forest.vertices(c))]


class AdaptiveSneddonRun:
    """run() of the reference for `test case = sneddon`, dim 2, `ref strategy = fixed preref sneddon`
    with local pre-refinement (cracks.cc:4166-4581), up to the first adaptive refinement cycle."""

    def __init__(self, global_refine=0, local_pre_refine=1, E=1.0, nu=0.2, G_c=1.0, pressure=1e-3,
                 kappa_of_h=lambda h: 1e-8 * h, eps_of_h=lambda h: 2.0 * h, newton_lower_bound=1e-7, max_newton=50,
                 max_line_search=10, line_search_damping=0.5, timestep=1.0, max_no_timesteps=3):
        n = 10 << global_refine
        self.forest = QuadForest(n, n, (-10.0, -10.0), (10.0, 10.0))
        self.prerefinement_h = []
        for _ in range(local_pre_refine):
            self.prerefinement_h.append(AdaptiveProblem(self.forest, orc.Params()).h_min)
            self.forest.refine(flag_fixed_preref_sneddon(self.forest))
        mu = E / (2.0 * (1 + nu))
        lam = (2 * nu * mu) / (1.0 - 2 * nu)
        self.prm = orc.Params(lam, mu, G_c, 0.0, 1.0, pressure, 0.0, 1.0, 1.0, 0, 0, 0.0, 0.0)
        self.p = AdaptiveProblem(self.forest, self.prm)
        self.prm.kappa, self.prm.eps = kappa_of_h(self.p.h_min), eps_of_h(self.p.h_min)
        self.E = E
        self.lower, self.max_newton = newton_lower_bound, max_newton
        self.max_ls, self.damp = max_line_search, line_search_damping
        self.dt, self.max_steps = timestep, max_no_timesteps
        self.statistics, self.logs, self.diffs = [], [], []

    def initial(self):
        p = self.p
        x, y = p.xy[:, 0], p.xy[:, 1]
        broken = (x * x <= 1.0) & (np.abs(2.0 * y) <= 2.0 * p.h_min)   # InitialValuesSneddon, cracks.cc:381-406
        sol = np.zeros((p.n_nodes, 3))
        sol[:, 2] = np.where(broken, 0.0, 1.0)
        return sol.reshape(-1)

    def set_initial_bc(self, sol):
        sol[self.p.dirichlet] = 0.0                            # u = 0 on the boundary, cracks.cc:2575-2583

    def newton_active_set(self, sol, old, oldold):
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        p = self.p
        log = orc.NewtonLog()
        self.set_initial_bc(sol)
        sol[:] = p.distribute_hanging(sol)
        # constraints_update still holds the active set of the previous solve (cracks.cc:2790-2794)
        constrained = self.constrained

        def residuals():
            raw = p.raw_residual(sol, old, oldold)
            r_total = p.H.T @ raw                              # hanging-node constraints only
            r_pde = r_total.copy()
            r_pde[constrained] = 0.0
            return r_pde, r_total

        r_pde, r_total = residuals()
        res = float(np.linalg.norm(r_pde))
        log.initial_residual = res
        old_res = res
        active = np.zeros(p.n_nodes, dtype=bool)
        cycle = np.zeros(p.n_nodes, dtype=np.int32)
        mass = self.mass
        step = 0
        while True:
            active_old = active
            nc, dim = p.nc, p.dim
            phi, phi_old = sol.reshape(-1, nc)[:, dim], old.reshape(-1, nc)[:, dim]
            crit = r_total.reshape(-1, nc)[:, dim] / mass + 10.0 * self.E * (phi - phi_old)
            active = (~p.is_hanging_node) & (~((crit <= 0.0) & (cycle < 5)))
            n_cyc = int(np.sum(active & (cycle >= 5)))
            phi[active] = phi_old[active]
            sol[:] = p.distribute_hanging(sol)
            cycle[active_old & ~active] += 1
            con = p.dirichlet.reshape(-1, nc).copy()
            con[:, dim] |= active
            constrained = self.constrained = con.reshape(-1)
            changed = bool(np.any(active != active_old))
            free = ~(constrained | p.is_hanging_dof)
            Cm = p.H @ sp.diags(free.astype(float))          # interpolation matrix of constraints_update
            J = p.raw_jacobian(sol, old, oldold)
            r_pde, _ = residuals()
            A = (Cm.T @ J @ Cm + sp.diags((~free).astype(float))).tocsc()
            update = Cm @ spla.spsolve(A, r_pde)
            saved = sol.copy()
            ls, new_res = 0, 0.0
            while ls < self.max_ls:
                sol += update
                r_pde, r_total = residuals()
                new_res = float(np.linalg.norm(r_pde))
                if new_res < res:
                    break
                sol[:] = saved
                update *= self.damp
                ls += 1
            log.rows.append((step + 1, int(active.sum()), n_cyc, new_res, new_res / res, ls))
            old_res, res = res, new_res
            step += 1
            if res < self.lower and not changed:
                break
            if step >= self.max_newton:
                raise orc.NoConvergence()
        self.logs.append(log)
        return res / old_res

    def run(self):
        p = self.p
        self.mass = p.lumped_mass()
        self.constrained = p.dirichlet.copy()                  # constraints_update after setup_system (1637-1641)
        sol = self.initial()
        phi = sol.reshape(-1, 3)[:, 2]
        np.clip(phi, 0.0, 1.0, out=phi)
        oldold, old = sol.copy(), sol.copy()
        self.prm.dt_old = self.prm.dt_oldold = self.dt
        time, step_no = 0.0, 0
        self.tcv, self.cod = None, None
        while step_no <= self.max_steps:
            oldold, old = old, sol.copy()
            time += self.dt
            self.newton_active_set(sol, old, oldold)
            np.clip(phi, 0.0, 1.0, out=phi)
            sol[:] = p.distribute_hanging(sol)                 # cracks.cc:4416-4417
            bulk, crack, tcv = p.functionals(sol)
            diff = float(np.max(np.abs(old - sol)))
            self.statistics.append(dict(step=step_no, time=time, dofs=p.n_dofs, h=p.h_min, bulk=bulk, crack=crack))
            self.diffs.append(diff)
            step_no += 1
            if diff < 1.0e-5:
                self.tcv = tcv
                # compute_functional_values(), cracks.cc:3704-3725: x = -1.5 + i/256; only lines that carry
                # mesh faces print a value (all vertex abscissae of these meshes are multiples of 1/256)
                xs_mesh = set(np.round(p.xy[:, 0] * 256.0).astype(int).tolist())
                self.cod = [(x, v) for i in range(3 * 256 + 1) for x in [-1.5 + i / 256.0]
                            if int(round(x * 256.0)) in xs_mesh
                            for v, nf in [p.cod(sol, x)] if nf > 0]
                break
        self.solution = sol
        return self.statistics

    def refined_once_more(self):
        """DoFs after `Refinement cycle 0` (cracks.cc:4525-4560): same strategy on the current forest"""
        f = QuadForest(self.forest.nx, self.forest.ny, self.forest.lo, self.forest.hi)
        f.cells = set(self.forest.cells)
        f.refine(flag_fixed_preref_sneddon(f))
        return AdaptiveProblem(f, self.prm)


def transfer(old_forest: QuadForest, old_prob: AdaptiveProblem, new_forest: QuadForest, new_prob: AdaptiveProblem,
             vectors):
    """SolutionTransfer::interpolate after refinement (cracks.cc:4137-4159): every new cell lies inside one
    old cell; its vertex values are the bilinear interpolant of that cell's four nodal values."""
    old_index = {c: k for k, c in enumerate(old_forest.order)}
    outs = [np.zeros(new_prob.n_dofs) for _ in vectors]
    done = np.zeros(new_prob.n_nodes, dtype=bool)
    for k, (L, i, j) in enumerate(new_forest.order):
        # the old leaf covering this cell
        l, a, b = L, i, j
        while (l, a, b) not in old_index:
            l, a, b = l - 1, a >> 1, b >> 1
            assert l >= 0, "coarsening is not part of the reference's refinement strategies used here"
        oc = old_index[(l, a, b)]
        s = 1 << (L - l)                                   # new cells per old cell edge
        onodes = old_prob.cells[oc]
        for v in range(4):
            n = new_prob.cells[k, v]
            if done[n]:
                continue
            xi = ((i - a * s) + (v & 1)) / s
            eta = ((j - b * s) + ((v >> 1) & 1)) / s
            w = ((1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta)
            for vec, out in zip(vectors, outs):
                vv = vec.reshape(-1, 3)
                out.reshape(-1, 3)[n] = sum(w[q] * vv[onodes[q]] for q in range(4))
            done[n] = True
    assert done.all()
    return outs


class AdaptiveMieheRun(AdaptiveSneddonRun):
    """run() of the reference for `test case = miehe tension / miehe shear` with the predictor-corrector
    refinement `ref strategy = phase field` (cracks.cc:4166-4581, 3971-3995, 4108-4159): after every
    converged step cells holding a phase-field dof below the threshold are refined (up to the level cap,
    2:1 balanced), the three solution vectors are interpolated to the new mesh and the step is redone."""

    def __init__(self, test, refine, timestep, lam, mu, E, G_c=2.7, kappa_of_h=lambda h: 0.0, eps_of_h=lambda h: 2.0 * h,
                 cycles=1, max_no_timesteps=32, timestep_2=None, switch_timestep=0, newton_lower_bound=1e-6,
                 max_newton=50, max_line_search=10, line_search_damping=0.6, d_rhs=0.0, d_mat=0.0,
                 refine_threshold=0.5):
        assert test in ("miehe tension", "miehe shear")
        self.test = test
        self.forest = QuadForest(2, 2, (0.0, 0.0), (1.0, 1.0), slit=True)
        for _ in range(refine):
            self.forest.refine(list(self.forest.cells))                       # refine_global
        self.level_cap = refine + cycles
        self.h_final = 0.5 * math.sqrt(2.0) * 2.0 ** (-(refine + cycles))    # cracks.cc:3839-3854
        self.prm = orc.Params(lam, mu, G_c, kappa_of_h(self.h_final), eps_of_h(self.h_final), 0.0, 0.0, 1.0, 1.0, 0,
                              0, d_rhs, d_mat)
        self.E, self.d_rhs, self.d_mat = E, d_rhs, d_mat
        self.lower, self.max_newton = newton_lower_bound, max_newton
        self.max_ls, self.damp = max_line_search, line_search_damping
        self.dt, self.max_steps = timestep, max_no_timesteps
        self.dt2, self.switch, self.threshold = timestep_2, switch_timestep, refine_threshold
        self.statistics, self.logs, self.diffs = [], [], []
        self._setup_system()

    def _setup_system(self):
        """setup_system() on the current forest (cracks.cc:1579-1680)"""
        self.p = p = AdaptiveProblem(self.forest, self.prm)
        x, y = p.xy[:, 0], p.xy[:, 1]
        m = np.zeros((p.n_nodes, 3), dtype=bool)
        top, bottom, left, right = y == 1.0, y == 0.0, x == 0.0, x == 1.0
        if self.test == "miehe tension":
            m[bottom, 1] = True
            m[top, 0] = m[top, 1] = True
        else:
            m[left, 1] = m[right, 1] = True
            m[bottom, 0] = m[bottom, 1] = True
            m[top, 0] = m[top, 1] = True
            upper_copy = np.zeros(p.n_nodes, dtype=bool)
            for k, n in self.forest.node_of.items():
                upper_copy[n] = k[2] == 1
            m[(y == 0.5) & (x >= 0.5) & ~upper_copy, 1] = True                # boundary id 4: lower slit face
        p.dirichlet = m.reshape(-1)
        self._top = np.where(top)[0]
        self.mass = p.lumped_mass()
        self.constrained = p.dirichlet.copy()

    def set_initial_bc(self, sol):
        time = self._time
        s = sol.reshape(-1, 3)
        con = self.p.dirichlet.reshape(-1, 3)
        for c in range(2):
            s[con[:, c], c] = 0.0
        if self.test == "miehe tension":
            s[self._top, 1] = time
        else:
            s[self._top, 0] = -time

    def load(self, sol):
        """compute_load() on the top edge (cracks.cc:3728-3816)"""
        p = self.p
        gq = 0.5 * math.sqrt(3.0 / 5.0)
        xi, w = (0.5 - gq, 0.5, 0.5 + gq), (5.0 / 18.0, 8.0 / 18.0, 5.0 / 18.0)
        s = sol.reshape(-1, 3)
        lx = ly = 0.0
        for c in range(p.n_cells):
            nodes, (hx, hy) = p.cells[c], p.cell_h[c]
            if p.xy[nodes[2], 1] != 1.0:
                continue
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

    def refine_mesh(self, vectors):
        """-> (changed, transferred vectors).  Strategy `phase field` + level cap + 2:1 balance."""
        p = self.p
        phi = vectors[0].reshape(-1, 3)[:, 2]
        flagged = [c for k, c in enumerate(self.forest.order)
                   if c[0] < self.level_cap and np.any(phi[p.cells[k]] < self.threshold)]
        if not flagged:
            return False, vectors
        old_forest, old_prob = self.forest, p
        self.forest = old_forest.copy()
        self.forest.refine(flagged)
        self._setup_system()
        return True, transfer(old_forest, old_prob, self.forest, self.p, vectors)

    def run(self):
        sol = np.zeros((self.p.n_nodes, 3))
        sol[:, 2] = 1.0
        sol = sol.reshape(-1)
        oldold, old = sol.copy(), sol.copy()
        dt = self.dt
        dt_old = dt_oldold = dt
        time, step_no = 0.0, 0
        self.redone = []
        while step_no <= self.max_steps:
            if self.switch > 0 and step_no > self.switch:
                dt = self.dt2
            tmp_dt = dt
            dt_oldold, dt_old = dt_old, dt
            oldold, old = old, sol.copy()
            while True:                                                       # redo_step
                self.prm.dt_old, self.prm.dt_oldold = dt_old, dt_oldold
                self.prm.split = 1 if (self.d_mat > 0 and step_no > 0) else 0
                time += dt
                while True:
                    try:
                        self._time = time
                        self.newton_active_set(sol, old, oldold)
                        break
                    except orc.NoConvergence:
                        sol[:] = old
                        time -= dt
                        dt /= 10.0
                        time += dt
                        if dt < 1e-6 * tmp_dt:
                            raise                                         # the reference would cut for ever
                phi = sol.reshape(-1, 3)[:, 2]
                np.clip(phi, 0.0, 1.0, out=phi)
                sol[:] = self.p.distribute_hanging(sol)
                changed, (sol, old, oldold) = self.refine_mesh([sol, old, oldold])
                if not changed:
                    break
                self.redone.append(step_no)                                   # "MESH CHANGED!", cracks.cc:4421-4431
                time -= dt
                sol = old.copy()
            dt = tmp_dt
            bulk, crack, _ = self.p.functionals(sol)
            lx, ly = self.load(sol)
            self.statistics.append(dict(step=step_no, time=time, dofs=self.p.n_dofs, h=self.h_final, bulk=bulk,
                                        crack=crack, load=ly if self.test == "miehe tension" else lx))
            step_no += 1
        self.solution = sol
        return self.statistics



class OctForest:
    """3-D counterpart of QuadForest (no slit): forest of octrees over an nx x ny x nz box mesh with
    p4est's full 2:1 balance (faces, edges, corners).  Hanging nodes: edge midpoints (two parents) and
    face centres (four parents) of cells whose neighbour is one level finer."""

    def __init__(self, n, lo, hi):
        self.n, self.lo, self.hi = tuple(n), tuple(lo), tuple(hi)
        self.cells = {(0, i, j, k) for k in range(n[2]) for j in range(n[1]) for i in range(n[0])}

    def copy(self):
        f = OctForest(self.n, self.lo, self.hi)
        f.cells = set(self.cells)
        return f

    def cell_size(self, L):
        return tuple((self.hi[d] - self.lo[d]) / (self.n[d] << L) for d in range(3))

    def centre(self, c):
        h = self.cell_size(c[0])
        return tuple(self.lo[d] + (c[1 + d] + 0.5) * h[d] for d in range(3))

    def vertices(self, c):
        h = self.cell_size(c[0])
        return [tuple(self.lo[d] + (c[1 + d] + ((v >> d) & 1)) * h[d] for d in range(3)) for v in range(8)]

    def _leaf_containing(self, L, idx):
        if any(idx[d] < 0 or idx[d] >= (self.n[d] << L) for d in range(3)):
            return None
        i, j, k = idx
        while L >= 0:
            if (L, i, j, k) in self.cells:
                return (L, i, j, k)
            L, i, j, k = L - 1, i >> 1, j >> 1, k >> 1
        return None

    def refine(self, flagged):
        for c in list(flagged):
            self._split(c)

    def _split(self, c):
        if c not in self.cells:
            return
        L, i, j, k = c
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                for dk in (-1, 0, 1):
                    if (di, dj, dk) == (0, 0, 0):
                        continue
                    nb = self._leaf_containing(L, (i + di, j + dj, k + dk))
                    if nb is not None and nb[0] < L:
                        self._split(nb)
        self.cells.discard(c)
        for v in range(8):
            self.cells.add((L + 1, 2 * i + (v & 1), 2 * j + ((v >> 1) & 1), 2 * k + ((v >> 2) & 1)))

    def build(self):
        Lm = max(c[0] for c in self.cells)
        order = sorted(self.cells, key=lambda c: (c[3] << (Lm - c[0]), c[2] << (Lm - c[0]), c[1] << (Lm - c[0]), c[0]))
        h0 = self.cell_size(Lm)
        node_of, coords, cells, hs, lattice = {}, [], [], [], []

        def node(pt):
            if pt not in node_of:
                node_of[pt] = len(coords)
                coords.append(tuple(self.lo[d] + pt[d] * h0[d] for d in range(3)))
            return node_of[pt]

        for c in order:
            s = 1 << (Lm - c[0])
            pts = [tuple((c[1 + d] + ((v >> d) & 1)) * s for d in range(3)) for v in range(8)]
            lattice.append(pts)
            cells.append([node(pt) for pt in pts])
            hs.append(tuple(s * h0[d] for d in range(3)))
        hanging = {}
        edges = [(a, a | (1 << d)) for d in range(3) for a in range(8) if not a & (1 << d)]
        faces = [[a for a in range(8) if ((a >> d) & 1) == side] for d in range(3) for side in (0, 1)]
        for pts in lattice:
            for group in edges + faces:
                ps = [pts[a] for a in group]
                ssum = [sum(q[d] for q in ps) for d in range(3)]
                if any(v % len(ps) for v in ssum):
                    continue
                mid = tuple(v // len(ps) for v in ssum)
                if mid in node_of:
                    hanging[node_of[mid]] = tuple(node_of[q] for q in ps)
        self.order, self.node_of, self.lattice_level = order, node_of, Lm
        return (np.array(cells, dtype=np.int64), np.array(hs, dtype=np.float64), np.array(coords, dtype=np.float64),
                hanging)


def initial_multiple_het_3d(xyz, width):
    """InitialValuesMultipleHet<3>, cracks.cc:595-610: two plate-shaped cracks, width = min cell diameter"""
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    w = width / 2.0
    c1 = (x >= 2.6 - w) & (x <= 2.6 + w) & (y >= 3.8 - w) & (y <= 5.5 + w) & (z >= 4.0 - w) & (z <= 4.0 + w)
    c2 = (x >= 5.5 - w) & (x <= 7.0 + w) & (y >= 4.0 - w) & (y <= 4.0 + w) & (z >= 6.0 - w) & (z <= 6.0 + w)
    return np.where(c1 | c2, 0.0, 1.0)


class HeteroRun3D(AdaptiveSneddonRun):
    """run() of the reference for `test case = multiple het`, dim 3 (cracks.cc:4166-4581): single-tree
    cube [0,10]^3, global refinement, local pre-refinement with `ref strategy = phase field` on the
    interpolated initial cracks, per-cell Lame coefficients from the E-modulus field, pressure(time)."""

    def __init__(self, e_modulus_of_cell, global_refine=3, local_pre_refine=1, nu=0.2, G_c=1.0,
                 pressure=lambda t: 1e3 * t, kappa_of_h=lambda h: 0.0, eps_of_h=lambda h: 1.5, E_active_set=1e4,
                 newton_lower_bound=1e-6, max_newton=20, max_line_search=8, line_search_damping=0.5, timestep=0.01,
                 max_no_timesteps=1, refine_threshold=0.4):
        self.forest = OctForest((1, 1, 1), (0.0,) * 3, (10.0,) * 3)
        for _ in range(global_refine):
            self.forest.refine(list(self.forest.cells))
        self.level_cap = global_refine + local_pre_refine
        self.prerefinement_h, self.prerefinement_dofs = [], []
        dummy = orc.Params(1.0, 1.0, G_c, 0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 0, 0, 0.0, 0.0)
        for _ in range(local_pre_refine):
            p = AdaptiveProblem(self.forest, dummy)
            self.prerefinement_h.append(p.h_min)
            self.prerefinement_dofs.append(p.n_dofs)
            phi = initial_multiple_het_3d(p.xy, p.h_min)
            flagged = [c for k, c in enumerate(self.forest.order)
                       if c[0] < self.level_cap and np.any(phi[p.cells[k]] < refine_threshold)]
            self.forest.refine(flagged)
        self.prm = orc.Params(1.0, 1.0, G_c, 0.0, 1.0, 0.0, 0.0, 1.0, 1.0, 0, 0, 0.0, 0.0)
        self.p = AdaptiveProblem(self.forest, self.prm)
        self.prm.kappa, self.prm.eps = kappa_of_h(self.p.h_min), eps_of_h(self.p.h_min)
        # E(cell centre); assembly adds 1.0 to it (cracks.cc:2209-2210), compute_energy does not (3651)
        E = np.array([e_modulus_of_cell(c, self.forest.centre(c)) for c in self.forest.order], dtype=np.float64)

        def lame(Ev):
            mu = Ev / (2.0 * (1 + nu))
            return np.stack([(2 * nu * mu) / (1.0 - 2 * nu), mu], axis=1)

        self.p.set_cell_lame(lame(E + 1.0), lame(E))
        self.E_cells = E
        self.E = E_active_set          # c = 10 E_modulus with whatever value the member holds (cracks.cc:2859)
        self.pressure = pressure
        self.lower, self.max_newton = newton_lower_bound, max_newton
        self.max_ls, self.damp = max_line_search, line_search_damping
        self.dt, self.max_steps = timestep, max_no_timesteps
        self.statistics, self.logs, self.diffs = [], [], []

    def initial(self):
        p = self.p
        sol = np.zeros((p.n_nodes, 4))
        sol[:, 3] = initial_multiple_het_3d(p.xy, p.h_min)
        return sol.reshape(-1)

    def run(self):
        p = self.p
        self.mass = p.lumped_mass()
        self.constrained = p.dirichlet.copy()
        sol = self.initial()
        phi = sol.reshape(-1, 4)[:, 3]
        np.clip(phi, 0.0, 1.0, out=phi)
        oldold, old = sol.copy(), sol.copy()
        self.prm.dt_old = self.prm.dt_oldold = self.dt
        time, step_no = 0.0, 0
        while step_no <= self.max_steps:
            oldold, old = old, sol.copy()
            time += self.dt
            self.prm.pressure = self.pressure(time)
            self.newton_active_set(sol, old, oldold)
            np.clip(phi, 0.0, 1.0, out=phi)
            sol[:] = p.distribute_hanging(sol)
            bulk, crack, _ = p.functionals(sol)
            self.statistics.append(dict(step=step_no, time=time, dofs=p.n_dofs, h=p.h_min, bulk=bulk, crack=crack))
            self.diffs.append(float(np.max(np.abs(old - sol))))
            step_no += 1
        self.solution = sol
        return self.statistics
