"""CPU-side checks of the C-ABI library: it builds for sm_100a, loads without a
GPU, and exports every symbol include/cracks_b200.h declares.  No compute calls."""
import ctypes
import os
import subprocess

import pytest


def test_library_builds_and_exports_header_symbols(pf):
    from cracks_b200 import api
    so = pf.library_path()
    assert os.path.exists(so)
    lib = ctypes.CDLL(so)
    declared = api.exported_symbols_in_header()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    # and the Python mirror binds every one of them
    bound = set(api._SIGS) | {"pf_last_error", "pf_n_dofs", "pf_stream", "pf_launch_count"}
    assert set(declared) <= bound, sorted(set(declared) - bound)


def test_library_is_sm100a_and_self_contained(pf):
    so = pf.library_path()
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    # no link-time dependency on NCCL, torch or the oracle
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    for lib in ("nccl", "torch", "pf_oracle"):
        assert lib not in needed


def test_bad_arguments_are_rejected_without_gpu(pf):
    from cracks_b200 import api
    lib = api.load_library()
    h = ctypes.c_void_p()
    m = pf.sneddon_mesh(3, 0)
    p = pf.sneddon_params(m)
    m.dim = 5
    assert lib.pf_create(ctypes.byref(m), ctypes.byref(p), 0, 0, 1, None, ctypes.byref(h)) == api.PF_BAD_ARG
    assert lib.pf_destroy(None) == api.PF_BAD_ARG
    assert lib.pf_last_error(None) == b"null context"


def test_product_does_not_import_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "cracks_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pf_oracle" not in text and "newton_oracle" not in text, os.path.join(dirpath, f)
