#!/usr/bin/env python
"""Transcribes the reference's own golden files for the Miehe tests into small
JSON fixtures (run in the build container, where /root/reference exists; the
fixtures travel, the reference does not):

  tests/miehe_shear_2.{prm,statistics,output}            -> miehe_shear_2.json            (KAT-4)
  tests/miehe_tension_adaptive_1.{prm,statistics}        -> miehe_tension_adaptive_1.json (KAT-3)
  tests/miehe_shear_1.{prm,statistics,output}            -> miehe_shear_1.json            (adaptive shear, split)
  the six Catch TEST_CASEs of cracks.cc:1740-1919        -> eigen_2x2.json                (KAT-6)
  tests/sneddon_2d_1.{prm,statistics,output}             -> sneddon_2d_1.json             (KAT-2, hanging nodes)
"""
import json
import math
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def prm_values(path):
    out = {}
    for line in open(path):
        line = line.split("#")[0].strip()
        m = re.match(r"set\s+(.*?)\s*=\s*(.*)$", line)
        if m:
            out[m.group(1).strip()] = m.group(2).strip()
    return out


def statistics(path):
    rows = []
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        f = line.split()
        rows.append(dict(step=int(f[0]), time=float(f[1]), dofs=int(f[2]), h=float(f[3]), bulk=float(f[4]),
                         crack=float(f[5]), load=float(f[6])))
    return rows


def sneddon_2d_1():
    out = open(f"{REF}/tests/sneddon_2d_1.output").read()
    rows = []
    for line in open(f"{REF}/tests/sneddon_2d_1.statistics"):
        if line.startswith("#") or not line.strip():
            continue
        f = line.split()
        rows.append(dict(step=int(f[0]), time=float(f[1]), dofs=int(f[2]), h=float(f[3]), bulk=float(f[4]),
                         crack=float(f[5])))
    tcv = re.search(r"TCV: value= (\S+) exact= (\S+)", out)
    cod = [(float(a), float(b)) for a, b in re.findall(r"^(-?\d+)  (\S+)$", out, flags=re.M)]
    d = dict(_source="tjhei/cracks tests/sneddon_2d_1.prm, .statistics, .output (transcribed by make_miehe_goldens.py)",
             prm=prm_values(f"{REF}/tests/sneddon_2d_1.prm"), statistics=rows,
             initial_newton_residual=initial_residuals(f"{REF}/tests/sneddon_2d_1.output"),
             timestep_difference_linfty=[float(x) for x in re.findall(r"Timestep difference linfty: (\S+)", out)],
             cells=int(re.search(r"Timestep 0: .*Cells: (\d+)", out).group(1)),
             prerefinement_h=float(re.search(r"Prerefinement step with h= (\S+)", out).group(1)),
             tcv=float(tcv.group(1)), cod=cod,
             dofs_after_refinement_cycle_0=int(re.search(r"Refinement cycle 0\s*\n-+\s*\n\s*\nDoFs: .* = (\d+)", out).group(1)))
    json.dump(d, open(os.path.join(HERE, "sneddon_2d_1.json"), "w"), indent=1)


def initial_residuals(path):
    return [float(m.group(1)) for m in re.finditer(r"^0\t\t\t(\S+)$", open(path).read(), flags=re.M)]


def main():
    sneddon_2d_1()
    for name in ("miehe_shear_2", "miehe_tension_adaptive_1", "miehe_shear_1"):
        d = dict(_source=f"tjhei/cracks tests/{name}.prm, .statistics, .output (transcribed by make_miehe_goldens.py)",
                 prm=prm_values(f"{REF}/tests/{name}.prm"), statistics=statistics(f"{REF}/tests/{name}.statistics"),
                 initial_newton_residual=initial_residuals(f"{REF}/tests/{name}.output"))
        if name == "miehe_shear_2":
            d["statistics_np2"] = statistics(f"{REF}/tests/miehe_shear_2.mpirun=2.statistics")
        json.dump(d, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
    s2, s17 = math.sqrt(2.0), math.sqrt(17.0)
    a, b = 3.5, s17 / 2.0
    v1, v2 = (-0.5 + b) / 2.0, (-0.5 - b) / 2.0
    l1, l2 = math.sqrt(v1 * v1 + 1.0), math.sqrt(v2 * v2 + 1.0)
    w1, w2 = 1.0 - s2, 1.0 + s2
    m1, m2 = math.sqrt(w1 * w1 + 1.0), math.sqrt(w2 * w2 + 1.0)
    cases = [  # matrix (row-major), eval1, evec1, eval2, evec2 -- cracks.cc:1740-1919
        dict(name="diagonal", m=[2, 0, 0, 3], e1=2.0, v1=[1, 0], e2=3.0, v2=[0, 1]),
        dict(name="(1,1)=0", m=[-2, 0, 0, 0], e1=-2.0, v1=[1, 0], e2=0.0, v2=[0, 1]),
        dict(name="(1,1)=0 test2", m=[5, 0, 0, 0], e1=5.0, v1=[1, 0], e2=0.0, v2=[0, 1]),
        dict(name="only offdiagonal", m=[0, -2, -2, 0], e1=2.0, v1=[1 / s2, -1 / s2], e2=-2.0, v2=[1 / s2, 1 / s2]),
        dict(name="full", m=[3, 2, 2, 4], e1=a + b, v1=[v1 / l1, 1 / l1], e2=a - b, v2=[-v2 / l2, -1 / l2]),
        dict(name="(0,0)=0", m=[0, -2, -2, 4], e1=2 + 2 * s2, v1=[-w1 / m1, -1 / m1], e2=2 - 2 * s2, v2=[w2 / m2, 1 / m2]),
    ]
    json.dump(dict(_source="tjhei/cracks cracks.cc:1740-1919 (Catch TEST_CASEs of eigen_vectors_and_values)", cases=cases),
              open(os.path.join(HERE, "eigen_2x2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
