"""ctypes mirror of the C++ host forest (cracks_b200/host/forest.h) and of pf_create_forest.

The host forest is the product-side stand-in for the reference's p4est triangulation +
make_hanging_node_constraints + SolutionTransfer (cracks.cc:3895-4163, 1630-1634); it is pure host
code and is checked against the CPU oracle's forests in tests/test_host_forest.py.

EXPERIMENTAL device side: `ForestContext` (pf_create_forest) and `ForestSneddonDriver` were written
against the hanging-node oracle but have not been run on a GPU yet; their tests are opt-in
(PF_EXPERIMENTAL=1).
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
    """pf_create_forest: the (u,phi) problem on a HostForest mesh (EXPERIMENTAL, see module docstring).
    Vectors cross the ABI in the block layout [u | phi] over the forest's node numbering."""

    def __init__(self, forest: HostForest, params: api.Params, device: int = 0, cell_lame=None, cell_lame_energy=None):
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
        rc = self.lib.pf_create_forest(C.cast(C.pointer(fm), C.c_void_p), C.byref(params), device, C.byref(h))
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
