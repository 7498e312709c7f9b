"""cracks_b200 -- B200-native (u,phi) hot path of tjhei/cracks behind a C ABI.

The product is ``libcracks_b200.so`` (hand-written sm_100a CUDA, see
``cracks_b200/csrc`` and ``include/cracks_b200.h``).  This package is the thin
Python mirror of that ABI used by the tests and by ``bench.py``; it contains
no numerics of its own and never falls back to a CPU path.
"""
from .api import (  # noqa: F401
    PFError,
    NoConvergence,
    PhaseFieldContext,
    Mesh,
    Params,
    build_library,
    library_path,
    load_library,
    sneddon_mesh,
    sneddon_params,
    SneddonDriver,
    MieheDriver,
    MeshWouldRefine,
    miehe_mesh,
    miehe_final_h,
)
