"""The C++ host driver (cracks_b200_run <file.prm>) on the GPU: KAT-1 through
the reference's own command-line / statistics-file surface."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_reproduces_golden_statistics(pf, tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cracks_b200", "host"), "-s"])
    golden = json.load(open(os.path.join(ROOT, "tests", "golden", "sneddon_3d_1.json")))
    prm = open(os.path.join(ROOT, "tests", "golden", "kat1_sneddon_3d.prm")).read()
    prm = prm.replace("output-kat1", str(tmp_path / "out"))
    (tmp_path / "kat1.prm").write_text(prm)
    r = subprocess.run([os.path.join(ROOT, "cracks_b200", "cracks_b200_run"), str(tmp_path / "kat1.prm"),
                        "--gmres-max-it", "3000"], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stderr
    assert "Problem dimension: 3" in r.stdout
    assert "DoFs: 3993 solid + 1331 phase = 5324" in r.stdout
    assert "0\t\t\t6.744161e+01" in r.stdout                      # tests/sneddon_3d_1.mpirun=4.output:29
    rows = [l.split() for l in open(tmp_path / "out" / "statistics") if not l.startswith("#")]
    assert len(rows) == len(golden["statistics"])
    for row, ref in zip(rows, golden["statistics"]):
        assert int(row[0]) == ref["step"] and float(row[1]) == ref["time"] and int(row[2]) == golden["dofs"]
        assert float(row[3]) == pytest.approx(golden["h_min"], rel=1e-8)
        assert float(row[5]) == pytest.approx(ref["crack"], rel=1e-8)
        assert float(row[4]) == pytest.approx(ref["bulk"], rel=1e-6)
    assert float(rows[0][4]) == pytest.approx(golden["statistics"][0]["bulk"], rel=1e-7)
    assert "TCV: value= 0.0399535" in r.stdout
    assert "\n0  0.00441323\n" in r.stdout                      # the COD line of the golden output


def test_cli_rejects_unsupported_cases(pf, tmp_path):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cracks_b200", "host"), "-s"])
    (tmp_path / "m.prm").write_text("subsection Global parameters\n set test case = miehe shear\nend\n")
    r = subprocess.run([os.path.join(ROOT, "cracks_b200", "cracks_b200_run"), str(tmp_path / "m.prm")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "miehe shear" in r.stderr
