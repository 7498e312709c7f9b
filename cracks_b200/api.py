"""ctypes mirror of include/cracks_b200.h.

Mirrors the reference's operator interface for the hot path: the object plays
the role of ``system_pde_matrix`` + ``assemble_system`` + ``solve`` in
``newton_active_set`` (cracks.cc:2780-2994); ``vmult`` has the argument order
of deal.II's ``A.vmult(dst, src)`` (cracks.cc:2729-2734, 2770).

There is no CPU fallback: if the CUDA library is missing or fails, this module
raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PF_OK, PF_BAD_ARG, PF_CUDA_ERROR, PF_NCCL_ERROR, PF_NO_CONVERGENCE, PF_NUMERIC, PF_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6


class PFError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cracks_b200 error {code}: {msg}")
        self.code = code


class NoConvergence(PFError):
    """SolverControl::NoConvergence of the reference (cracks.cc:2987, 2762-2771)."""


class Mesh(C.Structure):
    _fields_ = [("dim", C.c_int), ("n", C.c_int * 3), ("h", C.c_double * 3), ("origin", C.c_double * 3),
                ("slit", C.c_int)]


class Params(C.Structure):
    _fields_ = [("lam", C.c_double), ("mu", C.c_double), ("G_c", C.c_double), ("kappa", C.c_double),
                ("eps", C.c_double), ("alpha_biot", C.c_double)]


class Layout(C.Structure):
    _fields_ = [("n_nodes_global", C.c_int64), ("n_nodes_plane", C.c_int64), ("plane_begin", C.c_int),
                ("plane_end", C.c_int), ("owned_begin", C.c_int), ("owned_end", C.c_int), ("ncomp", C.c_int)]


class MgLevel(C.Structure):
    _fields_ = [("n", C.c_int * 3), ("replicated", C.c_int), ("mode_below", C.c_int), ("layout", Layout),
                ("inject_begin", C.c_int), ("inject_end", C.c_int), ("restrict_begin", C.c_int),
                ("restrict_end", C.c_int)]


def library_path() -> str:
    return os.path.join(_HERE, "libcracks_b200.so")


def build_library(force: bool = False) -> str:
    """Compile the CUDA library for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    so = library_path()
    deps = [os.path.join(src_dir, f) for f in os.listdir(src_dir) if f.endswith((".cu", ".cuh"))]
    deps.append(os.path.join(_HERE, "..", "include", "cracks_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["make", "-C", src_dir, "-s"])
    return so


_SIGS = {
    "pf_create": [C.POINTER(Mesh), C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)],
    "pf_create_forest": [C.c_void_p, C.POINTER(Params), C.c_int, C.POINTER(C.c_void_p)],   # see cracks_b200/forest.py
    "pf_create_forest_distributed": [C.c_void_p, C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.POINTER(C.c_void_p)],
    "pf_destroy": [C.c_void_p],
    "pf_nccl_unique_id": [C.c_void_p],
    "pf_get_layout": [C.c_void_p, C.POINTER(Layout)],
    "pf_slab_layout": [C.POINTER(Mesh), C.c_int, C.c_int, C.POINTER(Layout), C.POINTER(C.c_int), C.POINTER(C.c_int),
                       C.POINTER(C.c_int), C.POINTER(C.c_int)],
    "pf_mg_hierarchy": [C.POINTER(Mesh), C.c_int, C.c_int, C.POINTER(MgLevel), C.c_int, C.POINTER(C.c_int)],
    "pf_synchronize": [C.c_void_p],
    "pf_set_state": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double],
    "pf_get_solution": [C.c_void_p, C.c_void_p],
    "pf_get_state": [C.c_void_p, C.c_int, C.c_void_p],
    "pf_update_solution": [C.c_void_p, C.c_double],
    "pf_set_params": [C.c_void_p, C.POINTER(Params)],
    "pf_set_constraints": [C.c_void_p, C.c_void_p, C.c_void_p],
    "pf_set_dirichlet_all_faces": [C.c_void_p],
    "pf_residual": [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)],
    "pf_setup_jacobian": [C.c_void_p],
    "pf_set_preconditioner": [C.c_void_p, C.c_int, C.c_int, C.c_double],
    "pf_set_multigrid_precision": [C.c_void_p, C.c_int],
    "pf_set_jacobian_precision": [C.c_void_p, C.c_int],
    "pf_set_deterministic": [C.c_void_p, C.c_int],
    "pf_set_multigrid_coupling": [C.c_void_p, C.c_int],
    "pf_set_block_solve": [C.c_void_p, C.c_int],
    "pf_debug_set_block": [C.c_void_p, C.c_int],
    "pf_get_block_solve": [C.c_void_p],
    "pf_get_block_solve_stats": [C.c_void_p, C.POINTER(C.c_int64)],
    "pf_set_multigrid_graph": [C.c_void_p, C.c_int],
    "pf_apply_preconditioner": [C.c_void_p, C.c_void_p, C.c_void_p],
    "pf_set_krylov_dim": [C.c_void_p, C.c_int],
    "pf_apply_jacobian": [C.c_void_p, C.c_void_p, C.c_void_p],
    "pf_apply_jacobian_dev": [C.c_void_p, C.c_void_p, C.c_void_p],
    "pf_jacobian_diagonal": [C.c_void_p, C.c_void_p],
    "pf_lumped_mass": [C.c_void_p, C.c_void_p],
    "pf_active_set_update": [C.c_void_p, C.c_double, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                             C.POINTER(C.c_int)],
    "pf_get_active_set": [C.c_void_p, C.c_void_p],
    "pf_active_set_reset": [C.c_void_p],
    "pf_solve": [C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.POINTER(C.c_int)],
    "pf_energy": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "pf_tcv": [C.c_void_p, C.POINTER(C.c_double)],
    "pf_cod": [C.c_void_p, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_int64)],
    "pf_project_phase_field": [C.c_void_p],
    "pf_interpolate_sneddon": [C.c_void_p, C.c_double],
    "pf_set_stress_split": [C.c_void_p, C.c_int, C.c_double, C.c_double],
    "pf_dirichlet_miehe": [C.c_void_p, C.c_int, C.c_double, C.c_int],
    "pf_interpolate_unbroken": [C.c_void_p],
    "pf_load": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "pf_phase_field_min": [C.c_void_p, C.POINTER(C.c_double)],
    "pf_set_dirichlet_values": [C.c_void_p, C.c_void_p],
    "pf_load_cells": [C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)],
    "pf_advance_timestep": [C.c_void_p],
    "pf_set_time_parameters": [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_double],
    "pf_timestep_difference": [C.c_void_p, C.POINTER(C.c_double)],
    "pf_restore_old_solution": [C.c_void_p],
    "pf_save_solution": [C.c_void_p],
    "pf_restore_saved_solution": [C.c_void_p],
    "pf_scale_update": [C.c_void_p, C.c_double],
    "pf_device_vector": [C.c_void_p, C.POINTER(C.c_void_p)],
    "pf_device_vector_free": [C.c_void_p, C.c_void_p],
    "pf_upload": [C.c_void_p, C.c_void_p, C.c_void_p],
    "pf_download": [C.c_void_p, C.c_void_p, C.c_void_p],
    "pf_host_alloc": [C.POINTER(C.c_void_p), C.c_size_t],
    "pf_host_free": [C.c_void_p],
    "pf_debug_force_generic": [C.c_void_p, C.c_int],
    "pf_debug_disable_iso": [C.c_void_p, C.c_int],
    "pf_debug_set_variant": [C.c_void_p, C.c_int],
    "pf_profile_enable": [C.c_void_p, C.c_int],
    "pf_profile_read": [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)],
}


def load_library():
    """dlopen libcracks_b200.so and attach the signatures of every exported entry point."""
    global _LIB
    if _LIB is None:
        so = library_path()
        if not os.path.exists(so):
            raise PFError(PF_UNSUPPORTED, f"{so} is not built; run __graft_entry__.build()")
        lib = C.CDLL(so)
        for name, args in _SIGS.items():
            fn = getattr(lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
        lib.pf_last_error.argtypes = [C.c_void_p]
        lib.pf_last_error.restype = C.c_char_p
        lib.pf_n_dofs.argtypes = [C.c_void_p]
        lib.pf_n_dofs.restype = C.c_int64
        lib.pf_stream.argtypes = [C.c_void_p]
        lib.pf_stream.restype = C.c_void_p
        lib.pf_launch_count.argtypes = [C.c_void_p]
        lib.pf_launch_count.restype = C.c_int64
        _LIB = _TracedLib(lib) if os.environ.get("PF_PY_TRACE") else lib
    return _LIB


class _TracedLib:
    """Diagnostics (PF_PY_TRACE=1): host wall time and call count per C-ABI entry point."""

    def __init__(self, lib):
        self._lib, self.calls = lib, {}

    def __getattr__(self, name):
        import time
        fn = getattr(self._lib, name)

        def timed(*a):
            t0 = time.perf_counter()
            r = fn(*a)
            e = self.calls.setdefault(name, [0, 0.0])
            e[0] += 1
            e[1] += time.perf_counter() - t0
            return r
        return timed


def exported_symbols_in_header() -> list:
    """Names of the functions include/cracks_b200.h declares (for the ABI export test)."""
    import re
    hdr = open(os.path.join(_HERE, "..", "include", "cracks_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(pf_[a-z0-9_]+)\s*\(", hdr)))


def mg_hierarchy(mesh: Mesh, rank: int, nranks: int) -> list:
    """The multigrid levels (fine to coarse) `rank` of `nranks` builds; pure host code, no GPU needed."""
    levels = (MgLevel * 16)()
    n = C.c_int()
    rc = load_library().pf_mg_hierarchy(C.byref(mesh), rank, nranks, levels, 16, C.byref(n))
    if rc != PF_OK:
        raise PFError(rc, "pf_mg_hierarchy: bad arguments")
    out = []
    for L in levels[: n.value]:
        lay = L.layout
        out.append(dict(n=tuple(L.n), replicated=bool(L.replicated), mode_below=L.mode_below,
                        plane_begin=lay.plane_begin, plane_end=lay.plane_end, owned_begin=lay.owned_begin,
                        owned_end=lay.owned_end, inject=(L.inject_begin, L.inject_end),
                        restrict=(L.restrict_begin, L.restrict_end)))
    return out


def slab_layout(mesh: Mesh, rank: int, nranks: int) -> dict:
    """The library's z-slab decomposition for (rank, nranks); pure host code, no GPU needed."""
    lay = Layout()
    cb, ce, ocb, oce = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    rc = load_library().pf_slab_layout(C.byref(mesh), rank, nranks, C.byref(lay), C.byref(cb), C.byref(ce),
                                       C.byref(ocb), C.byref(oce))
    if rc != PF_OK:
        raise PFError(rc, "pf_slab_layout: bad arguments")
    return dict(plane_begin=lay.plane_begin, plane_end=lay.plane_end, owned_begin=lay.owned_begin,
                owned_end=lay.owned_end, cell_begin=cb.value, cell_end=ce.value, own_cell_begin=ocb.value,
                own_cell_end=oce.value, nodes_per_plane=lay.n_nodes_plane, n_nodes_global=lay.n_nodes_global)


def sneddon_mesh(dim: int, refine: int) -> Mesh:
    """subdivided_hyper_rectangle(10 per direction) on [-10,10]^dim + refine_global (cracks.cc:1207-1253, 1534)."""
    m = Mesh()
    m.dim = dim
    n = 10 * 2 ** refine
    for d in range(3):
        m.n[d] = n if d < dim else 1
        m.h[d] = 20.0 / n if d < dim else 1.0
        m.origin[d] = -10.0 if d < dim else 0.0
    return m


def mesh_diameter(m: Mesh) -> float:
    return math.sqrt(sum(m.h[d] ** 2 for d in range(m.dim)))


def sneddon_params(m: Mesh, E=1.0, nu=0.2, G_c=1.0, kappa_of_h=lambda h: 1e-8 * h, eps_of_h=lambda h: 2.0 * h) -> Params:
    """Material / regularisation parameters of parameters_sneddon_3d.prm (cracks.cc:1500-1511, 3876-3882)."""
    h = mesh_diameter(m)
    mu = E / (2.0 * (1 + nu))
    lam = (2 * nu * mu) / (1.0 - 2 * nu)
    return Params(lam, mu, G_c, kappa_of_h(h), eps_of_h(h), 0.0)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class PhaseFieldContext:
    """One rank's view of the matrix-free (u,phi) problem on one GPU."""

    def __init__(self, mesh: Mesh, params: Params, device: int = 0, rank: int = 0, nranks: int = 1, nccl_id: bytes | None = None):
        self.lib = load_library()
        self.mesh, self.params = mesh, params
        self.dim = mesh.dim
        self.nc = mesh.dim + 1
        h = C.c_void_p()
        idbuf = C.create_string_buffer(nccl_id, 128) if nccl_id is not None else None
        rc = self.lib.pf_create(C.byref(mesh), C.byref(params), device, rank, nranks, idbuf, C.byref(h))
        self.h = h
        self._check(rc)
        lay = Layout()
        self._check(self.lib.pf_get_layout(self.h, C.byref(lay)))
        self.layout = lay
        self.n_nodes = int(lay.n_nodes_global)      # includes the doubled nodes of a slit mesh
        self.n_dofs = self.n_nodes * self.nc
        self.n_local_dofs = (lay.plane_end - lay.plane_begin) * lay.n_nodes_plane * self.nc

    # -- plumbing ---------------------------------------------------------
    def _check(self, rc):
        if rc == PF_OK:
            return
        msg = self.lib.pf_last_error(self.h).decode() if self.h else ""
        if rc == PF_NO_CONVERGENCE:
            raise NoConvergence(rc, msg)
        raise PFError(rc, msg)

    def close(self):
        if self.h:
            self.lib.pf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        if load_library().pf_nccl_unique_id(buf) != PF_OK:
            raise PFError(PF_NCCL_ERROR, "ncclGetUniqueId failed")
        return buf.raw

    def synchronize(self):
        self._check(self.lib.pf_synchronize(self.h))

    @property
    def stream(self) -> int:
        return self.lib.pf_stream(self.h) or 0

    @property
    def launch_count(self) -> int:
        return self.lib.pf_launch_count(self.h)

    # -- layout helpers (block layout [u | phi] <-> (node, comp) arrays) ----
    def to_block(self, nodal: np.ndarray) -> np.ndarray:
        a = np.asarray(nodal).reshape(self.n_nodes, self.nc)
        return np.concatenate([a[:, : self.dim].reshape(-1), a[:, self.dim]])

    def to_nodal(self, block: np.ndarray) -> np.ndarray:
        b = np.asarray(block)
        out = np.empty((self.n_nodes, self.nc), dtype=b.dtype)
        out[:, : self.dim] = b[: self.n_nodes * self.dim].reshape(self.n_nodes, self.dim)
        out[:, self.dim] = b[self.n_nodes * self.dim:]
        return out.reshape(-1)

    # -- state ----------------------------------------------------------------
    def set_state(self, sol, old=None, oldold=None, dt_old=1.0, dt_oldold=1.0, use_old_timestep_pf=False, pressure=0.0):
        self._keep = (sol, old, oldold)
        self._check(self.lib.pf_set_state(self.h, _ptr(sol), _ptr(old), _ptr(oldold), dt_old, dt_oldold,
                                          int(use_old_timestep_pf), pressure))

    def set_time_parameters(self, dt_old, dt_oldold, use_old_timestep_pf, pressure):
        self._check(self.lib.pf_set_time_parameters(self.h, dt_old, dt_oldold, int(use_old_timestep_pf), pressure))

    def get_state(self, which: int) -> np.ndarray:
        """0: solution, 1: old_solution, 2: old_old_solution (block layout)"""
        out = np.zeros(self.n_dofs)
        self._check(self.lib.pf_get_state(self.h, which, _ptr(out)))
        return out

    def get_solution(self) -> np.ndarray:
        out = np.zeros(self.n_dofs)
        self._check(self.lib.pf_get_solution(self.h, _ptr(out)))
        return out

    def set_constraints(self, dirichlet_mask=None, active_mask=None):
        self._check(self.lib.pf_set_constraints(self.h, _ptr(dirichlet_mask), _ptr(active_mask)))

    def set_dirichlet_all_faces(self):
        self._check(self.lib.pf_set_dirichlet_all_faces(self.h))

    # -- hot path ---------------------------------------------------------------
    def residual(self, want_vectors=True):
        """assemble_nl_residual + set_zero + l2_norm -> (r_pde, r_total, norm)"""
        nrm = C.c_double()
        if want_vectors:
            r_pde, r_tot = np.zeros(self.n_dofs), np.zeros(self.n_dofs)
            self._check(self.lib.pf_residual(self.h, _ptr(r_pde), _ptr(r_tot), C.byref(nrm)))
            return r_pde, r_tot, nrm.value
        self._check(self.lib.pf_residual(self.h, None, None, C.byref(nrm)))
        return None, None, nrm.value

    def set_preconditioner(self, kind=1, cheb_degree=2, cheb_ratio=6.0):
        """0 = Jacobi, 1 = geometric multigrid (stand-in for the reference's ML AMG)"""
        self._check(self.lib.pf_set_preconditioner(self.h, kind, cheb_degree, cheb_ratio))

    def block_solve(self):
        """whether pf_solve runs as u stage + phi stage on this context"""
        return bool(self.lib.pf_get_block_solve(self.h))

    def block_solve_stats(self):
        """counters of the staged linear solves since the context was created"""
        out = (C.c_int64 * 4)()
        self._check(self.lib.pf_get_block_solve_stats(self.h, out))
        return {"solves": int(out[0]), "with_u_stage": int(out[1]), "u_iterations": int(out[2]), "phi_iterations": int(out[3])}

    def set_block_solve(self, on=True):
        """pf_solve as a u stage followed by a phi stage (the Jacobian has no (u,phi) block, cracks.cc:2333-2337)"""
        self._check(self.lib.pf_set_block_solve(self.h, int(on)))

    def set_multigrid_coupling(self, coupled=True):
        """False: block-diagonal smoother operator (no (phi,u) block), like the reference's BlockDiagonalPreconditioner"""
        self._check(self.lib.pf_set_multigrid_coupling(self.h, int(coupled)))

    def set_multigrid_graph(self, on=True):
        """the V-cycle as one CUDA graph launch (several GPUs: the cycle is launch-bound)"""
        self._check(self.lib.pf_set_multigrid_graph(self.h, int(on)))

    def set_deterministic(self, on=True):
        """scatter kernels colour by colour: bit-identical results run to run (3-D box meshes)"""
        self._check(self.lib.pf_set_deterministic(self.h, int(on)))

    def set_jacobian_precision(self, bits=64):
        """64 (default) or 32: inexact Newton with the Jacobian apply in FP32 (residuals and GMRES vectors stay FP64)"""
        self._check(self.lib.pf_set_jacobian_precision(self.h, bits))

    def set_multigrid_precision(self, bits=64):
        """64 (default) or 32: run the V-cycle of the multigrid preconditioner in FP32 (outer GMRES stays FP64)"""
        self._check(self.lib.pf_set_multigrid_precision(self.h, bits))

    def apply_preconditioner(self, v):
        """z = M^-1 v (one V-cycle, or Jacobi), block layout"""
        v = np.ascontiguousarray(v, dtype=np.float64)
        z = np.empty_like(v)
        self._check(self.lib.pf_apply_preconditioner(self.h, _ptr(v), _ptr(z)))
        return z

    def get_active_set(self):
        """one byte per node, 1 = in the active set (does not change it)"""
        m = np.zeros(self.n_nodes, dtype=np.uint8)
        self._check(self.lib.pf_get_active_set(self.h, _ptr(m)))
        return m

    def setup_jacobian(self):
        self._check(self.lib.pf_setup_jacobian(self.h))

    def vmult(self, dst: np.ndarray, src: np.ndarray):
        """dst = J(U) src, host buffers in block layout (deal.II argument order)."""
        self._check(self.lib.pf_apply_jacobian(self.h, _ptr(src), _ptr(dst)))

    def vmult_dev(self, dst_dev: int, src_dev: int):
        self._check(self.lib.pf_apply_jacobian_dev(self.h, C.c_void_p(src_dev), C.c_void_p(dst_dev)))

    def jacobian_diagonal(self) -> np.ndarray:
        out = np.zeros(self.n_dofs)
        self._check(self.lib.pf_jacobian_diagonal(self.h, _ptr(out)))
        return out

    def lumped_mass(self) -> np.ndarray:
        out = np.zeros(self.n_nodes)
        self._check(self.lib.pf_lumped_mass(self.h, _ptr(out)))
        return out

    def active_set_update(self, c: float, want_mask=True):
        na, ncy, ch = C.c_int64(), C.c_int64(), C.c_int()
        mask = np.zeros(self.n_nodes, dtype=np.uint8) if want_mask else None
        self._check(self.lib.pf_active_set_update(self.h, c, _ptr(mask), C.byref(na), C.byref(ncy), C.byref(ch)))
        return mask, na.value, ncy.value, bool(ch.value)

    def active_set_reset(self):
        self._check(self.lib.pf_active_set_reset(self.h))

    def solve(self, tol_rel=1e-8, max_it=200, want_dx=False):
        n_it = C.c_int()
        dx = np.zeros(self.n_dofs) if want_dx else None
        rc = self.lib.pf_solve(self.h, tol_rel, max_it, _ptr(dx), C.byref(n_it))
        self._check(rc)
        return dx, n_it.value

    def set_krylov_dim(self, m: int):
        self._check(self.lib.pf_set_krylov_dim(self.h, m))

    def set_stress_split(self, active: bool, d_rhs: float, d_mat: float):
        self._check(self.lib.pf_set_stress_split(self.h, int(active), d_rhs, d_mat))

    def dirichlet_miehe(self, kind: int, time: float, set_values: bool = True):
        self._check(self.lib.pf_dirichlet_miehe(self.h, kind, time, int(set_values)))

    def interpolate_unbroken(self):
        self._check(self.lib.pf_interpolate_unbroken(self.h))

    def load(self):
        lx, ly = C.c_double(), C.c_double()
        self._check(self.lib.pf_load(self.h, C.byref(lx), C.byref(ly)))
        return lx.value, ly.value

    def set_dirichlet_values(self, values_block: np.ndarray):
        self._check(self.lib.pf_set_dirichlet_values(self.h, _ptr(np.ascontiguousarray(values_block, dtype=np.float64))))

    def load_cells(self, cells: np.ndarray):
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        lx, ly = C.c_double(), C.c_double()
        self._check(self.lib.pf_load_cells(self.h, _ptr(cells), cells.shape[0], C.byref(lx), C.byref(ly)))
        return lx.value, ly.value

    def phase_field_min(self) -> float:
        v = C.c_double()
        self._check(self.lib.pf_phase_field_min(self.h, C.byref(v)))
        return v.value

    def update_solution(self, alpha=1.0):
        self._check(self.lib.pf_update_solution(self.h, alpha))

    def save_solution(self):
        self._check(self.lib.pf_save_solution(self.h))

    def restore_saved_solution(self):
        self._check(self.lib.pf_restore_saved_solution(self.h))

    def scale_update(self, f):
        self._check(self.lib.pf_scale_update(self.h, f))

    # -- functionals ---------------------------------------------------------------
    def energy(self):
        b, c = C.c_double(), C.c_double()
        self._check(self.lib.pf_energy(self.h, C.byref(b), C.byref(c)))
        return b.value, c.value

    def tcv(self) -> float:
        t = C.c_double()
        self._check(self.lib.pf_tcv(self.h, C.byref(t)))
        return t.value

    def cod(self, eval_line: float):
        """(value, n_faces) of compute_cod(eval_line), cracks.cc:3452-3549"""
        v, nf = C.c_double(), C.c_int64()
        self._check(self.lib.pf_cod(self.h, eval_line, C.byref(v), C.byref(nf)))
        return v.value, nf.value

    def project_phase_field(self):
        self._check(self.lib.pf_project_phase_field(self.h))

    def interpolate_sneddon(self, h_diam):
        self._check(self.lib.pf_interpolate_sneddon(self.h, h_diam))

    def advance_timestep(self):
        self._check(self.lib.pf_advance_timestep(self.h))

    def timestep_difference(self) -> float:
        d = C.c_double()
        self._check(self.lib.pf_timestep_difference(self.h, C.byref(d)))
        return d.value

    def restore_old_solution(self):
        self._check(self.lib.pf_restore_old_solution(self.h))

    def profile_enable(self, on=True):
        self._check(self.lib.pf_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms, cnt = C.c_double(), C.c_int64()
        self._check(self.lib.pf_profile_read(self.h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    # -- device vectors ----------------------------------------------------------------
    def device_vector(self) -> int:
        p = C.c_void_p()
        self._check(self.lib.pf_device_vector(self.h, C.byref(p)))
        return p.value

    def device_vector_free(self, p: int):
        self._check(self.lib.pf_device_vector_free(self.h, C.c_void_p(p)))

    def upload(self, host_block: np.ndarray, dev: int):
        self._check(self.lib.pf_upload(self.h, _ptr(host_block), C.c_void_p(dev)))
        self.synchronize()

    def download(self, dev: int) -> np.ndarray:
        out = np.zeros(self.n_dofs)
        self._check(self.lib.pf_download(self.h, C.c_void_p(dev), _ptr(out)))
        return out


def miehe_mesh(refine: int) -> Mesh:
    """meshes/unit_slit.inp (2 x 2 cells on the unit square, slit from the centre to the right
    edge) after `refine` global refinements (cracks.cc:1202-1205, 1534)."""
    m = Mesh()
    m.dim = 2
    n = 2 * 2 ** refine
    for d in range(2):
        m.n[d], m.h[d], m.origin[d] = n, 1.0 / n, 0.0
    m.n[2], m.h[2], m.origin[2] = 1, 1.0, 0.0
    m.slit = 1
    return m


def miehe_final_h(refine: int, cycles: int = 0) -> float:
    """determine_mesh_dependent_parameters() for the Miehe tests: the diameter the cells will
    have on the FINAL level (cracks.cc:3839-3854), coarse diameter sqrt(2)/2."""
    return 0.5 * math.sqrt(2.0) * 2.0 ** (-(refine + cycles))


@dataclass
class NewtonRow:
    it: int
    n_active: int
    n_cycling: int
    residual: float
    reduction: float
    line_search: int
    lin_its: int


class SneddonDriver:
    """Host-side restatement of run() / newton_active_set() for `test case = sneddon`
    (cracks.cc:4166-4581, 2780-2994), every vector resident on the GPU."""

    def __init__(self, ctx: PhaseFieldContext, E=1.0, pressure=lambda t: 1e-3, timestep=1.0, max_no_timesteps=5,
                 newton_lower_bound=1e-7, max_newton=50, max_line_search=10, line_search_damping=0.5,
                 gmres_max_it=200, gmres_tol=1e-8, log=None):
        self.ctx, self.E, self.pressure = ctx, E, pressure
        self.dt, self.max_steps = timestep, max_no_timesteps
        self.lower, self.max_newton = newton_lower_bound, max_newton
        self.max_ls, self.damp = max_line_search, line_search_damping
        self.gmres_max_it, self.gmres_tol = gmres_max_it, gmres_tol
        self.log = log or (lambda s: None)
        self.statistics, self.history = [], []
        self.newton_its = 0
        self.lin_its = 0
        self.tcv = None
        self.phase_s = {}          # host wall time per phase of the Newton loop (diagnostics)

    def _timed(self, name, fn, *a, **k):
        import time as _t
        t0 = _t.perf_counter()
        r = fn(*a, **k)
        self.phase_s[name] = self.phase_s.get(name, 0.0) + _t.perf_counter() - t0
        return r

    def newton_active_set(self):
        c = self.ctx
        _, _, res = c.residual(want_vectors=False)
        self.log("It.\t#A.Set\t#CycDoF\tResidual\tReduction\tLSrch\t#LinIts")
        self.log("0\t\t\t%.6e" % res)
        old_res = res
        # active_set.clear() + a fresh cycle_counter (cracks.cc:2802-2807); the
        # r_total of the residual above stays valid for the first update
        c.active_set_reset()
        rows = []
        step = 0
        while True:
            _, n_act, n_cyc, changed = self._timed("active_set", c.active_set_update, 10.0 * self.E, want_mask=False)
            self._timed("setup_jacobian", c.setup_jacobian)
            # rhs with the new constraints (cracks.cc:2917-2918)
            self._timed("residual", lambda: c._check(c.lib.pf_residual(c.h, None, None, None)))
            _, n_lin = self._timed("solve", c.solve, self.gmres_tol, self.gmres_max_it)
            c.save_solution()
            ls = 0
            new_res = 0.0
            while ls < self.max_ls:
                c.update_solution(1.0)
                _, _, new_res = self._timed("residual", c.residual, want_vectors=False)
                if new_res < res:
                    break
                c.restore_saved_solution()
                c.scale_update(self.damp)
                ls += 1
            rows.append(NewtonRow(step + 1, n_act, n_cyc, new_res, new_res / res, ls, n_lin))
            self.log("%d\t%d\t%d\t%.6e\t%.6e\t%d\t%d" % (step + 1, n_act, n_cyc, new_res, new_res / res, ls, n_lin))
            old_res, res = res, new_res
            step += 1
            self.lin_its += n_lin
            if res < self.lower and not changed:
                break
            if step >= self.max_newton:
                raise NoConvergence(PF_NO_CONVERGENCE, "Newton iteration did not converge in %d steps" % step)
            # NB: if every trial was rejected the solution is the restored one but
            # r_total stays that of the last rejected trial, like cracks.cc:2947
        self.newton_its += step
        self.history.append(rows)
        return res / old_res

    def run(self, h_diam):
        c = self.ctx
        c.set_dirichlet_all_faces()
        c.interpolate_sneddon(h_diam)
        c.project_phase_field()
        c.interpolate_sneddon(h_diam)       # old = oldold = solution (cracks.cc:4276-4277)
        dt_old = dt_oldold = self.dt
        time, step_no = 0.0, 0
        while True:
            dt_oldold, dt_old = dt_old, self.dt
            c.advance_timestep()
            time += self.dt
            c.set_time_parameters(dt_old, dt_oldold, False, self.pressure(time))
            self.log("Timestep %d: %g (%g)" % (step_no, time - self.dt, self.dt))
            self.newton_active_set()
            c.project_phase_field()
            bulk, crack = c.energy()
            diff = c.timestep_difference()
            self.statistics.append(dict(step=step_no, time=time, bulk=bulk, crack=crack, diff=diff))
            self.log("No %d time %g bulk energy: %.8e crack energy: %.8e  diff %.6g" % (step_no, time, bulk, crack, diff))
            step_no += 1
            if diff < 1.0e-5:
                self.tcv = c.tcv()
                # compute_functional_values(), cracks.cc:3704-3725: x = -1.5 + i/256, i = 0..768
                h0, x0 = c.mesh.h[0], c.mesh.origin[0]
                on_plane = lambda x: abs((x - x0) / h0 - round((x - x0) / h0)) * h0 <= 1e-8
                self.cod = [(x, v) for x in (-1.5 + i / 256.0 for i in range(3 * 256 + 1)) if on_plane(x)
                            for v, nf in [c.cod(x)] if nf > 0]
                break
            if step_no > self.max_steps:
                break
        return self.statistics


class MeshWouldRefine(PFError):
    """refine_mesh() of the reference would change the mesh here (cracks.cc:3971-3995, 4108-4133);
    predictor-corrector refinement is outside this library's scope."""

    def __init__(self, step):
        super().__init__(PF_UNSUPPORTED, "phase field below the refinement threshold at time step %d" % step)
        self.step = step


class MieheDriver(SneddonDriver):
    """Host-side restatement of run() for `test case = miehe tension / miehe shear` on the
    uniformly refined slit mesh (cracks.cc:4166-4581): time-dependent Dirichlet data, stress
    split from the second time step on, load functional, time-step cut on NoConvergence."""

    def __init__(self, ctx: PhaseFieldContext, test: str, E, timestep, max_no_timesteps, timestep_2=None,
                 switch_timestep=0, d_rhs=0.0, d_mat=0.0, cycles=0, refine_threshold=0.8, **kw):
        super().__init__(ctx, E=E, timestep=timestep, max_no_timesteps=max_no_timesteps, **kw)
        self.kind = {"miehe tension": 1, "miehe shear": 2}[test]
        self.dt2, self.switch = timestep_2, switch_timestep
        self.d_rhs, self.d_mat = d_rhs, d_mat
        self.cycles, self.threshold = cycles, refine_threshold

    def run(self):
        c = self.ctx
        c.interpolate_unbroken()
        c.project_phase_field()
        dt = self.dt
        dt_old = dt_oldold = dt
        time, step_no = 0.0, 0
        while step_no <= self.max_steps:
            if self.switch > 0 and step_no > self.switch:
                dt = self.dt2
            tmp_dt = dt
            dt_oldold, dt_old = dt_old, dt
            c.advance_timestep()
            c.set_time_parameters(dt_old, dt_oldold, False, 0.0)
            c.set_stress_split(self.d_mat > 0 and step_no > 0, self.d_rhs, self.d_mat)
            time += dt
            self.log("Timestep %d: %g (%g)" % (step_no, time - dt, dt))
            while True:
                try:
                    c.dirichlet_miehe(self.kind, time, True)      # set_initial_bc(time), cracks.cc:2787
                    self.newton_active_set()
                    break
                except NoConvergence:
                    self.log("Solver did not converge! Adjusting time step to %g" % (dt / 10))
                    c._check(c.lib.pf_restore_old_solution(c.h))  # cracks.cc:4333-4355
                    time -= dt
                    dt /= 10.0
                    time += dt
            c.project_phase_field()
            if self.cycles > 0 and c.phase_field_min() < self.threshold:
                raise MeshWouldRefine(step_no)
            dt = tmp_dt
            bulk, crack = c.energy()
            lx, ly = c.load()
            self.statistics.append(dict(step=step_no, time=time, bulk=bulk, crack=crack,
                                        load=ly if self.kind == 1 else lx))
            self.log("No %d time %g bulk energy: %.8e crack energy: %.8e  load %.8e"
                     % (step_no, time, bulk, crack, self.statistics[-1]["load"]))
            step_no += 1
        return self.statistics
