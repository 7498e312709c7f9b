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
    (tmp_path / "m.prm").write_text("subsection Global parameters\n set test case = three point bending\nend\n")
    r = subprocess.run([os.path.join(ROOT, "cracks_b200", "cracks_b200_run"), str(tmp_path / "m.prm")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "three point bending" in r.stderr


SECTIONS = {
    "Global parameters": ["Global pre-refinement steps", "Local pre-refinement steps", "Adaptive refinement cycles",
                          "Max No of timesteps", "Timestep size", "Timestep size to switch to",
                          "Switch timestep after steps", "outer solver", "test case", "ref strategy",
                          "value phase field for refinement", "Output filename"],
    "Problem dependent parameters": ["K reg", "Eps reg", "Gamma penalization", "Pressure", "Fracture toughness G_c",
                                     "Poisson ratio nu", "E modulus", "Lame mu", "Lame lambda"],
    "Solver parameters": ["Use Direct Inner Solver", "Newton lower bound", "Newton maximum steps", "Upper Newton rho",
                          "Line search maximum steps", "Line search damping", "Decompose stress in rhs",
                          "Decompose stress in matrix"],
}


def _write_prm(path, values, outdir):
    """A .prm file with the values the golden fixture transcribed from the reference's test prm."""
    lines = []
    for sec, keys in SECTIONS.items():
        lines.append("subsection " + sec)
        if sec == "Global parameters":
            lines.append("  set Dimension = 2")
            lines.append("  set Output directory = " + str(outdir))
        for k in keys:
            if k in values:
                lines.append("  set %s = %s" % (k, values[k]))
        lines.append("end")
    path.write_text("\n".join(lines) + "\n")


@pytest.mark.parametrize("name,rows,stops", [("miehe_shear_2", 25, False), ("miehe_tension_adaptive_1", 25, True)])
def test_cli_miehe_goldens(pf, tmp_path, name, rows, stops):
    """KAT-4 / KAT-3 through `cracks_b200_run file.prm`: the statistics file has the reference's
    columns (incl. Load x / Load y) and values; the adaptive case stops where refine_mesh() would fire."""
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cracks_b200", "host"), "-s"])
    g = json.load(open(os.path.join(ROOT, "tests", "golden", name + ".json")))
    _write_prm(tmp_path / "t.prm", g["prm"], tmp_path / "out")
    r = subprocess.run([os.path.join(ROOT, "cracks_b200", "cracks_b200_run"), str(tmp_path / "t.prm")],
                       capture_output=True, text=True, timeout=900)
    print(r.stdout[-2000:], r.stderr[-1000:])
    assert "DoFs: 594 solid + 297 phase = 891" in r.stdout            # tests/miehe_shear_2.output
    if stops:
        assert r.returncode == 1 and "refine_mesh() would refine the mesh in time step 25" in r.stderr
    else:
        assert r.returncode == 0, r.stderr
    text = open(tmp_path / "out" / "statistics").read()
    assert ("# 7: Load x" if "shear" in name else "# 7: Load y") in text
    got = [l.split() for l in text.splitlines() if not l.startswith("#")]
    assert len(got) == rows
    for row, ref in zip(got, g["statistics"]):
        assert int(row[0]) == ref["step"] and int(row[2]) == 891
        assert float(row[3]) == pytest.approx(ref["h"], rel=1e-8)
        tol = 1e-6 if ref["step"] <= 18 else 1e-3
        for col, k in ((4, "bulk"), (5, "crack"), (6, "load")):
            assert float(row[col]) == pytest.approx(ref[k], rel=tol), (ref["step"], k)
