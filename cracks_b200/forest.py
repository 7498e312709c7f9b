"""ctypes mirror of the C++ host forest (cracks_b200/host/forest.h) and of pf_create_forest.

The host forest is the product-side stand-in for the reference's p4est triangulation +
make_hanging_node_constraints + SolutionTransfer (cracks.cc:3895-4163, 1630-1634); it is pure host
code and is checked against the CPU oracle's forests in tests/test_host_forest.py.

Device side: `ForestContext` (pf_create_forest / pf_create_forest_distributed) and the drivers below, held
against the hanging-node oracle and the reference's adaptive goldens in tests/test_gpu_forest.py (one GPU)
and tests/mgpu_forest_check.py (several ranks: replicated vectors, partitioned cells, one all-reduce per
operator application).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import api

_HERE = os.path.dirname(os.path.abspath(__file__))
_HOST = None


def host_library():
    global _HOST
    if _HOST is None:
        so = os.path.join(_HERE, "libcracks_b200_host.so")
        src = os.path.join(_HERE, "host")
        if not os.path.exists(so) or any(os.path.getmtime(os.path.join(src, f)) > os.path.getmtime(so)
                                         for f in ("forest.cc", "forest.h")):
            subprocess.check_call(["make", "-C", src, "-s", "../libcracks_b200_host.so"])
        lib = C.CDLL(so)
        vp, ll = C.c_void_p, C.c_longlong
        lib.pfh_forest_create.restype = vp
        lib.pfh_forest_create.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int]
        lib.pfh_forest_clone.restype = vp
        lib.pfh_forest_clone.argtypes = [vp]
        lib.pfh_forest_destroy.restype = None
        lib.pfh_forest_destroy.argtypes = [vp]
        lib.pfh_forest_refine_global.restype = None
        lib.pfh_forest_refine_global.argtypes = [vp, C.c_int]
        lib.pfh_forest_refine.restype = C.c_int
        lib.pfh_forest_refine.argtypes = [vp, vp, ll]
        for name in ("n_cells", "n_nodes", "n_hanging"):
            f = getattr(lib, "pfh_forest_" + name)
            f.restype, f.argtypes = ll, [vp]
        lib.pfh_forest_max_level.restype = C.c_int
        lib.pfh_forest_max_level.argtypes = [vp]
        lib.pfh_forest_min_cell_diameter.restype = C.c_double
        lib.pfh_forest_min_cell_diameter.argtypes = [vp]
        lib.pfh_forest_tables.restype = None
        lib.pfh_forest_tables.argtypes = [vp] * 7
        lib.pfh_forest_transfer.restype = C.c_int
        lib.pfh_forest_transfer.argtypes = [vp, vp, vp, vp, C.c_int]
        _HOST = lib
    return _HOST


class HostForest:
    """Locally refined box / slit mesh with 2:1 balance, node numbering and hanging-node table."""

    def __init__(self, dim, n, lo, hi, slit=False, _handle=None):
        self.lib = host_library()
        self.dim = dim
        if _handle is None:
            nn = (C.c_int * 3)(*(list(n) + [1] * (3 - len(n))))
            l = (C.c_double * 3)(*(list(lo) + [0.0] * (3 - len(lo))))
            h = (C.c_double * 3)(*(list(hi) + [1.0] * (3 - len(hi))))
            _handle = self.lib.pfh_forest_create(dim, nn, l, h, int(slit))
            if not _handle:
                raise ValueError("bad forest description")
        self.h = C.c_void_p(_handle)

    def __del__(self):
        try:
            if self.h:
                self.lib.pfh_forest_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def clone(self) -> "HostForest":
        return HostForest(self.dim, None, None, None, _handle=self.lib.pfh_forest_clone(self.h))

    def refine_global(self, times=1):
        self.lib.pfh_forest_refine_global(self.h, times)

    def refine(self, flags):
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        if self.lib.pfh_forest_refine(self.h, flags.ctypes.data, flags.shape[0]) != 0:
            raise ValueError("one flag per active cell")

    @property
    def n_cells(self):
        return self.lib.pfh_forest_n_cells(self.h)

    @property
    def n_nodes(self):
        return self.lib.pfh_forest_n_nodes(self.h)

    @property
    def n_hanging(self):
        return self.lib.pfh_forest_n_hanging(self.h)

    @property
    def max_level(self):
        return self.lib.pfh_forest_max_level(self.h)

    @property
    def min_cell_diameter(self):
        return self.lib.pfh_forest_min_cell_diameter(self.h)

    def tables(self) -> dict:
        nv, dim = 1 << self.dim, self.dim
        t = dict(conn=np.empty((self.n_cells, nv), dtype=np.int64), level=np.empty(self.n_cells, dtype=np.uint8),
                 coords=np.empty((self.n_nodes, dim)), hanging=np.empty((self.n_hanging, 5), dtype=np.int64),
                 level_h=np.empty((self.max_level + 1, dim)), upper_copy=np.empty(self.n_nodes, dtype=np.uint8))
        self.lib.pfh_forest_tables(self.h, *(t[k].ctypes.data for k in ("conn", "level", "coords", "hanging", "level_h",
                                                                         "upper_copy")))
        return t

    def transfer_from(self, coarser: "HostForest", vec: np.ndarray, ncomp: int) -> np.ndarray:
        """SolutionTransfer::interpolate of a nodal vector (ncomp interleaved) from a coarser state of this forest"""
        vec = np.ascontiguousarray(vec, dtype=np.float64)
        out = np.empty(self.n_nodes * ncomp)
        if self.lib.pfh_forest_transfer(self.h, coarser.h, vec.ctypes.data, out.ctypes.data, ncomp) != 0:
            raise ValueError("the target forest is not a refinement of the source")
        return out


class ForestMeshStruct(C.Structure):
    _fields_ = [("dim", C.c_int), ("n_cells", C.c_int64), ("n_nodes", C.c_int64), ("conn", C.c_void_p),
                ("cell_level", C.c_void_p), ("n_levels", C.c_int), ("level_h", C.c_void_p), ("n_hanging", C.c_int64),
                ("hanging", C.c_void_p), ("cell_lame", C.c_void_p), ("cell_lame_energy", C.c_void_p)]


class ForestContext(api.PhaseFieldContext):
    """pf_create_forest: the (u,phi) problem on a HostForest mesh.
    Vectors cross the ABI in the block layout [u | phi] over the forest's node numbering."""

    def __init__(self, forest: HostForest, params: api.Params, device: int = 0, cell_lame=None, cell_lame_energy=None,
                 dist=None):
        """dist = (rank, nranks, fresh_id) for several GPUs: fresh_id() returns an ncclUniqueId (128 bytes) that
        rank 0 made and every rank received -- a new one per context, the adaptive drivers build many"""
        self.lib = api.load_library()
        self.forest, self.params = forest, params
        self.dim, self.nc = forest.dim, forest.dim + 1
        self.mesh = None
        t = self.tables = forest.tables()
        self._keep = [t]
        fm = ForestMeshStruct(forest.dim, forest.n_cells, forest.n_nodes, t["conn"].ctypes.data, t["level"].ctypes.data,
                              forest.max_level + 1, t["level_h"].ctypes.data, forest.n_hanging,
                              t["hanging"].ctypes.data if forest.n_hanging else None, None, None)
        if cell_lame is not None:
            la = np.ascontiguousarray(cell_lame, dtype=np.float64)
            le = np.ascontiguousarray(cell_lame if cell_lame_energy is None else cell_lame_energy, dtype=np.float64)
            self._keep += [la, le]
            fm.cell_lame, fm.cell_lame_energy = la.ctypes.data, le.ctypes.data
        h = C.c_void_p()
        if dist is None or dist[1] == 1:
            rc = self.lib.pf_create_forest(C.cast(C.pointer(fm), C.c_void_p), C.byref(params), device, C.byref(h))
        else:
            rank, nranks, fresh_id = dist
            self._nccl_id = C.create_string_buffer(bytes(fresh_id()), 128)
            rc = self.lib.pf_create_forest_distributed(C.cast(C.pointer(fm), C.c_void_p), C.byref(params), device, rank,
                                                       nranks, self._nccl_id, C.byref(h))
        self.h = h
        self._check(rc)
        lay = api.Layout()
        self._check(self.lib.pf_get_layout(self.h, C.byref(lay)))
        self.layout = lay
        self.n_nodes = int(lay.n_nodes_global)
        self.n_dofs = self.n_nodes * self.nc
        self.n_local_dofs = self.n_dofs


class ForestSneddonDriver(api.SneddonDriver):
    """run() for `test case = sneddon`, dim 2, with `Local pre-refinement steps` (cracks.cc:4166-4581) on a
    HostForest: the mesh work (flagging, refinement, initial values, Dirichlet rows) is host logic, the
    Newton loop is the inherited device-resident one."""

    @staticmethod
    def prerefined_forest(global_refine=0, local_pre_refine=1):
        n = 10 << global_refine
        f = HostForest(2, (n, n), (-10.0, -10.0), (10.0, 10.0))
        for _ in range(local_pre_refine):
            t = f.tables()
            xy = t["coords"][t["conn"]]                                   # [cell][vertex][2]
            inside = (np.abs(xy[:, :, 0]) <= 2.5) & (np.abs(xy[:, :, 1]) <= 1.25)
            f.refine(inside.any(axis=1))                                  # fixed preref sneddon, cracks.cc:3901-3923
        return f

    def run_on_forest(self):
        c = self.ctx
        t = c.tables
        x, y = t["coords"][:, 0], t["coords"][:, 1]
        h = c.forest.min_cell_diameter
        on_b = (np.abs(x) == 10.0) | (np.abs(y) == 10.0)
        dmask = np.zeros((c.n_nodes, 3), dtype=np.uint8)
        dmask[on_b, :2] = 1
        zeros = np.zeros(c.n_nodes, dtype=np.uint8)
        c.set_constraints(c.to_block(dmask.reshape(-1)).astype(np.uint8), np.concatenate([np.zeros(2 * c.n_nodes, np.uint8), zeros]))
        sol = np.zeros((c.n_nodes, 3))
        sol[:, 2] = np.where((x * x <= 1.0) & (np.abs(2.0 * y) <= 2.0 * h), 0.0, 1.0)   # InitialValuesSneddon
        blk = c.to_block(sol.reshape(-1))
        c.set_state(blk, blk, blk, self.dt, self.dt, False, self.pressure(0.0))
        c.project_phase_field()
        dt_old = dt_oldold = self.dt
        time, step_no = 0.0, 0
        while True:
            dt_oldold, dt_old = dt_old, self.dt
            c.advance_timestep()
            time += self.dt
            c.set_time_parameters(dt_old, dt_oldold, False, self.pressure(time))
            self.newton_active_set()
            c.project_phase_field()
            bulk, crack = c.energy()
            diff = c.timestep_difference()
            self.statistics.append(dict(step=step_no, time=time, bulk=bulk, crack=crack, diff=diff))
            step_no += 1
            if diff < 1.0e-5:
                self.tcv = c.tcv()
                break
            if step_no > self.max_steps:
                break
        return self.statistics


class ForestMieheDriver(api.SneddonDriver):
    """run() for `test case = miehe tension / miehe shear` WITH the predictor-corrector refinement
    (`ref strategy = phase field`, cracks.cc:4166-4581, 3971-3995, 4108-4159, 4419-4431) on a HostForest:
    after every converged step the cells holding a phase-field dof below the threshold are refined (level
    cap, 2:1 balance), solution / old / old_old are interpolated to the new forest, a new device context
    is built and the step is redone.  """

    def __init__(self, test, refine, params_of_h, E, timestep, max_no_timesteps, cycles=1, timestep_2=None,
                 switch_timestep=0, d_rhs=0.0, d_mat=0.0, refine_threshold=0.8, device=0, krylov_dim=300, dist=None, **kw):
        self.kind = {"miehe tension": 1, "miehe shear": 2}[test]
        self.dist = dist
        self.forest = HostForest(2, (2, 2), (0.0, 0.0), (1.0, 1.0), slit=True)
        self.forest.refine_global(refine)
        self.level_cap = refine + cycles
        self.h_final = api.miehe_final_h(refine, cycles)
        self.params = params_of_h(self.h_final)                  # api.Params with K reg / Eps reg at the final h
        self.device, self.krylov_dim = device, krylov_dim
        self.dt2, self.switch = timestep_2, switch_timestep
        self.d_rhs, self.d_mat, self.threshold = d_rhs, d_mat, refine_threshold
        self.redone = []
        super().__init__(self._new_context(), E=E, timestep=timestep, max_no_timesteps=max_no_timesteps, **kw)

    def _new_context(self):
        """setup_system() on the current forest: tables -> device, Dirichlet rows of set_newton_bc (2584-2625)"""
        f = self.forest
        ctx = ForestContext(f, self.params, device=self.device, dist=self.dist)
        ctx.set_krylov_dim(self.krylov_dim)
        t = ctx.tables
        x, y = t["coords"][:, 0], t["coords"][:, 1]
        top, bottom, left, right = y == 1.0, y == 0.0, x == 0.0, x == 1.0
        m = np.zeros((ctx.n_nodes, 3), dtype=np.uint8)
        if self.kind == 1:
            m[bottom, 1] = 1
            m[top, 0] = m[top, 1] = 1
        else:
            m[left, 1] = m[right, 1] = 1
            m[bottom, 0] = m[bottom, 1] = 1
            m[top, 0] = m[top, 1] = 1
            m[(y == 0.5) & (x >= 0.5) & (t["upper_copy"] == 0), 1] = 1      # boundary id 4: lower face of the slit
        self._top_nodes = np.where(top)[0]
        self._top_cells = np.where(t["coords"][t["conn"][:, 2], 1] == 1.0)[0].astype(np.int64)
        none = np.zeros(ctx.n_dofs, dtype=np.uint8)
        ctx.set_constraints(ctx.to_block(m.reshape(-1)).astype(np.uint8), none)
        return ctx

    def _bc_values(self, time):
        v = np.zeros((self.ctx.n_nodes, 3))
        if self.kind == 1:
            v[self._top_nodes, 1] = time                              # BoundaryTensionTest
        else:
            v[self._top_nodes, 0] = -time                             # BoundaryShearTest
        return self.ctx.to_block(v.reshape(-1))

    def _refine_mesh(self):
        """-> True if the mesh changed; the three vectors are carried over (SolutionTransfer)"""
        c = self.ctx
        t = c.tables
        phi = c.to_nodal(c.get_state(0)).reshape(-1, 3)[:, 2]
        flags = (t["level"] < self.level_cap) & (phi[t["conn"]] < self.threshold).any(axis=1)
        if not flags.any():
            return False
        vecs = [c.to_nodal(c.get_state(w)) for w in (0, 1, 2)]
        coarse = self.forest
        self.forest = coarse.clone()
        self.forest.refine(flags)
        new = [self.forest.transfer_from(coarse, v, 3) for v in vecs]
        c.close()
        self.ctx = c = self._new_context()
        c.set_state(c.to_block(new[0]), c.to_block(new[1]), c.to_block(new[2]), self._dt_old, self._dt_oldold, False, 0.0)
        return True

    def run(self):
        c = self.ctx
        sol = np.zeros((c.n_nodes, 3))
        sol[:, 2] = 1.0                                               # InitialValuesTensionOrShear
        blk = c.to_block(sol.reshape(-1))
        c.set_state(blk, blk, blk, self.dt, self.dt, False, 0.0)
        dt = self.dt
        self._dt_old = self._dt_oldold = dt
        time, step_no = 0.0, 0
        while step_no <= self.max_steps:
            if self.switch > 0 and step_no > self.switch:
                dt = self.dt2
            tmp_dt = dt
            self._dt_oldold, self._dt_old = self._dt_old, dt
            self.ctx.advance_timestep()
            while True:                                               # redo_step
                c = self.ctx
                c.set_time_parameters(self._dt_old, self._dt_oldold, False, 0.0)
                c.set_stress_split(self.d_mat > 0 and step_no > 0, self.d_rhs, self.d_mat)
                time += dt
                while True:
                    try:
                        c.set_dirichlet_values(self._bc_values(time))   # set_initial_bc(time)
                        self.newton_active_set()
                        break
                    except api.NoConvergence:
                        c._check(c.lib.pf_restore_old_solution(c.h))
                        time -= dt
                        dt /= 10.0
                        time += dt
                        if dt < 1e-6 * tmp_dt:
                            raise
                c.project_phase_field()
                if not self._refine_mesh():
                    break
                self.redone.append(step_no)                           # "MESH CHANGED!": redo the step from old_solution
                time -= dt
                self.ctx._check(self.ctx.lib.pf_restore_old_solution(self.ctx.h))
            dt = tmp_dt
            c = self.ctx
            bulk, crack = c.energy()
            lx, ly = c.load_cells(self._top_cells)
            self.statistics.append(dict(step=step_no, time=time, dofs=c.n_dofs, bulk=bulk, crack=crack,
                                        load=ly if self.kind == 1 else lx))
            step_no += 1
        return self.statistics


def initial_multiple_het_3d(xyz: np.ndarray, width: float) -> np.ndarray:
    """InitialValuesMultipleHet<3> for the phase-field component (cracks.cc:595-610): two plate-shaped cracks"""
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    w = width / 2.0
    c1 = (x >= 2.6 - w) & (x <= 2.6 + w) & (y >= 3.8 - w) & (y <= 5.5 + w) & (z >= 4.0 - w) & (z <= 4.0 + w)
    c2 = (x >= 5.5 - w) & (x <= 7.0 + w) & (y >= 4.0 - w) & (y <= 4.0 + w) & (z >= 6.0 - w) & (z <= 6.0 + w)
    return np.where(c1 | c2, 0.0, 1.0)


class ForestHeteroDriver(api.SneddonDriver):
    """run() for `test case = multiple het`, dim 3 (BASELINE config 5; cracks.cc:4166-4581): the single-tree cube
    [0,10]^3 of meshes/unit_cube_10.inp, global refinement, local pre-refinement with `ref strategy = phase field`
    on the interpolated initial cracks, Lame coefficients per cell from an E-modulus field (the reference's
    BitmapFunction; `e_modulus_of_cells(centres) -> E`), `E + 1` in the assembly but not in compute_energy
    (2209-2210 vs 3651), pressure as a function of time.  """

    def __init__(self, e_modulus_of_cells, global_refine=3, local_pre_refine=1, nu=0.2, G_c=1.0, pressure=lambda t: 1e3 * t,
                 kappa_of_h=lambda h: 0.0, eps_of_h=lambda h: 1.5, E_active_set=1e4, refine_threshold=0.4, timestep=0.01,
                 max_no_timesteps=1, device=0, krylov_dim=300, dist=None, **kw):
        f = HostForest(3, (1, 1, 1), (0.0,) * 3, (10.0,) * 3)
        f.refine_global(global_refine)
        cap = global_refine + local_pre_refine
        self.prerefinement = []
        for _ in range(local_pre_refine):
            t = f.tables()
            self.prerefinement.append((f.min_cell_diameter, int(f.n_nodes) * 4))
            phi = initial_multiple_het_3d(t["coords"], f.min_cell_diameter)
            f.refine((t["level"] < cap) & (phi[t["conn"]] < refine_threshold).any(axis=1))
        self.forest = f
        t = f.tables()
        h = f.min_cell_diameter
        centres = t["coords"][t["conn"]].mean(axis=1)
        E = np.asarray(e_modulus_of_cells(centres), dtype=np.float64)

        def lame(Ev):
            mu = Ev / (2.0 * (1 + nu))
            return np.stack([(2 * nu * mu) / (1.0 - 2 * nu), mu], axis=1)

        params = api.Params(1.0, 1.0, G_c, kappa_of_h(h), eps_of_h(h), 0.0)     # lambda, mu come per cell
        ctx = ForestContext(f, params, device=device, cell_lame=lame(E + 1.0), cell_lame_energy=lame(E), dist=dist)
        ctx.set_krylov_dim(krylov_dim)
        xyz = t["coords"]
        on_b = ((xyz == 0.0) | (xyz == 10.0)).any(axis=1)
        m = np.zeros((ctx.n_nodes, 4), dtype=np.uint8)
        m[on_b, :3] = 1                                                        # u = 0 on the six faces, cracks.cc:2686-2694
        ctx.set_constraints(ctx.to_block(m.reshape(-1)).astype(np.uint8), np.zeros(ctx.n_dofs, dtype=np.uint8))
        super().__init__(ctx, E=E_active_set, pressure=pressure, timestep=timestep, max_no_timesteps=max_no_timesteps, **kw)

    def run(self):
        c = self.ctx
        sol = np.zeros((c.n_nodes, 4))
        sol[:, 3] = initial_multiple_het_3d(c.tables["coords"], c.forest.min_cell_diameter)
        blk = c.to_block(sol.reshape(-1))
        c.set_state(blk, blk, blk, self.dt, self.dt, False, self.pressure(0.0))
        c.project_phase_field()
        dt_old = dt_oldold = self.dt
        time, step_no = 0.0, 0
        while step_no <= self.max_steps:
            dt_oldold, dt_old = dt_old, self.dt
            c.advance_timestep()
            time += self.dt
            c.set_time_parameters(dt_old, dt_oldold, False, self.pressure(time))
            self.newton_active_set()
            c.project_phase_field()
            bulk, crack = c.energy()
            self.statistics.append(dict(step=step_no, time=time, dofs=c.n_dofs, bulk=bulk, crack=crack))
            step_no += 1
        return self.statistics
