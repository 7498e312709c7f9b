"""CPU tests of the C++ host side: the .prm surface (ParameterHandler subset +
expression evaluator) that replaces deal.II's for the hot-path driver."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "cracks_b200", "prm_check")


@pytest.fixture(scope="module")
def tool(pf):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cracks_b200", "host"), "-s"])
    return TOOL


def run(tool, path, h=1.0, t=1.0):
    return subprocess.run([tool, path, str(h), str(t)], capture_output=True, text=True)


def test_kat1_fixture_parses(tool):
    r = run(tool, os.path.join(ROOT, "tests", "golden", "kat1_sneddon_3d.prm"), h=3.4641016151377544)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert "K=0\n" in out
    assert "Eps=6.92820323027550" in out
    assert "Pressure=0.001" in out
    assert "set outer solver = active set\n" in out          # '#' glued to the value is a comment
    assert "set Max No of timesteps = 5\n" in out              # tabs / repeated blanks in the key
    assert "set Local pre-refinement steps = 0\n" in out
    assert "set Upper Newton rho = 0.999\n" in out             # default kept


@pytest.mark.parametrize("text,msg", [
    ("subsection Global parameters\n set No such key = 1\nend\n", "no entry with name"),
    ("subsection Global parameters\n set Dimension = three\nend\n", "not an integer"),
    ("subsection Global parameters\n set test case = sneddon2\nend\n", "not one of"),
    ("subsection Nope\nend\n", "no such subsection"),
    ("subsection Global parameters\n set Dimension = 3\n", "unbalanced"),
    ("subsection Solver parameters\n set Use Direct Inner Solver = maybe\nend\n", "'true' or 'false'"),
])
def test_bad_input_is_an_error(tool, tmp_path, text, msg):
    p = tmp_path / "bad.prm"
    p.write_text(text)
    r = run(tool, str(p))
    assert r.returncode == 1
    assert msg in r.stderr


def test_expressions(tool, tmp_path):
    p = tmp_path / "expr.prm"
    p.write_text("subsection Problem dependent parameters\n"
                 "  set K reg = 0.25 * pow(h,0.5)\n  set Eps reg = 2.0*h + sqrt(4)/2 - (1e-1)^2\n"
                 "  set Pressure = 1.0e+3*time\nend\n")
    r = run(tool, str(p), h=4.0, t=2.5)
    assert r.returncode == 0, r.stderr
    assert "K=0.5\n" in r.stdout
    assert "Eps=8.99" in r.stdout
    assert "Pressure=2500\n" in r.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_all_reference_prm_files_parse_unchanged(tool):
    files = sorted(glob.glob("/root/reference/*.prm") + glob.glob("/root/reference/tests/*.prm"))
    assert len(files) >= 15
    for f in files:
        r = run(tool, f, h=0.5, t=2.0)
        assert r.returncode == 0, (f, r.stderr)
