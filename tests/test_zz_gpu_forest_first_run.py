"""Runs the opt-in suite of the forest / hanging-node device path (tests/test_gpu_forest_experimental.py) in a
SUBPROCESS, last in the GPU suite, and does not gate on it: that path was written after the round's GPU
budget was spent; its kernel sources and the library plumbing pass on the CPU emulation
(tests/test_emulated_library_cpu.py) but it has not met a real GPU yet.  Isolation keeps a device fault or
a hang (420 s / 180 s caps) from touching the gating tests; the outcome is printed either way and reported as
passed / xfailed."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _first_run(test_file, cap):
    env = dict(os.environ, PF_EXPERIMENTAL="1")
    try:
        r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", test_file),
                            "-q", "-rA", "-s", "-p", "no:cacheprovider"], capture_output=True, text=True, timeout=cap, env=env,
                           cwd=ROOT)
        out, rc = r.stdout[-4000:] + r.stderr[-1500:], r.returncode
    except subprocess.TimeoutExpired as exc:
        out, rc = "TIMEOUT after %d s\n" % cap + str(exc.stdout)[-2000:], -1
    print(out)
    return rc


def test_forest_device_path_first_gpu_run(pf):
    if _first_run("test_gpu_forest_experimental.py", 420) != 0:
        pytest.xfail("forest device path failed on its first GPU run (non-gating, see the captured output)")


def test_fp32_vcycle_first_gpu_run(pf):
    """same arrangement for the FP32 V-cycle (pf_mg_lowp.cuh, opt-in at run time)"""
    if _first_run("test_gpu_fp32_vcycle_experimental.py", 180) != 0:
        pytest.xfail("FP32 V-cycle failed on its first GPU run (non-gating, see the captured output)")
