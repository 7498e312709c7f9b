#!/usr/bin/env python
"""Builds tests/golden/hetero_3d_1.json (KAT-5) in the build container, where /root/reference exists:

  * the golden numbers of tests/hetero_3d_1.mpirun-4.{statistics,output}, transcribed;
  * the E-modulus field of `test case = multiple het` sampled at the cell centres of the KAT-5 mesh
    (932 cells), i.e. BitmapFunction<3>::value of cracks.cc:118-241 evaluated on the reference's
    test.pgm -- including its quirks: the "255" max-value token of the PGM header is read as the first
    pixel (150-155), and xi = eta = min(max(., 1), 0) = 0, so there is no interpolation (197-198).
    The 1.9 MB bitmap itself stays in the reference; only these 932 numbers travel.
"""
import json
import math
import os
import re
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))


class BitmapFile:
    def __init__(self, path):
        f = open(path)
        f.readline()
        second = f.readline()
        rest = f.read().split()
        if not second.startswith("#"):
            rest = second.split() + rest
        self.nx, self.ny = int(rest[0]), int(rest[1])
        self.data = np.array(rest[2:2 + self.nx * self.ny], dtype=np.float64) / 255.0
        self.hx, self.hy = 1.0 / (self.nx - 1), 1.0 / (self.ny - 1)

    def pixel(self, i, j):
        return self.data[self.nx * (self.ny - 1 - j) + i]

    def get_value(self, x, y):
        ix = min(max(int(x / self.hx), 0), self.nx - 2)
        iy = min(max(int(y / self.hy), 0), self.ny - 2)
        xi = min(max((x - ix * self.hx) / self.hx, 1.0), 0.0)
        eta = min(max((y - iy * self.hy) / self.hy, 1.0), 0.0)
        return ((1 - xi) * (1 - eta) * self.pixel(ix, iy) + xi * (1 - eta) * self.pixel(ix + 1, iy)
                + (1 - xi) * eta * self.pixel(ix, iy + 1) + xi * eta * self.pixel(ix + 1, iy + 1))


def main():
    import adaptive_oracle as ao
    bm = BitmapFile(f"{REF}/test.pgm")
    E_mod = 1e4                                   # `E modulus` of tests/hetero_3d_1.prm; field range [E, 10 E]

    def e_modulus(cell, centre):
        x, y, z = (centre[d] / 10.0 for d in range(3))
        v = (bm.get_value(x / 10.0, (y - z) / 10.0) + 0.5 * bm.get_value((x + y) / 2.0, (z + x) / 2.0)
             + 0.25 * bm.get_value(math.fmod(z + x - y, 10.0), math.fmod(y + x, 10.0)))
        return E_mod + v * (10.0 * E_mod - E_mod) / 2.25

    run = ao.HeteroRun3D(e_modulus)
    out = open(f"{REF}/tests/hetero_3d_1.mpirun-4.output").read()
    rows = []
    for line in open(f"{REF}/tests/hetero_3d_1.mpirun-4.statistics"):
        if line.startswith("#") or not line.strip():
            continue
        f = line.split()
        rows.append(dict(step=int(f[0]), time=float(f[1]), dofs=int(f[2]), h=float(f[3]), bulk=float(f[4]),
                         crack=float(f[5])))
    d = dict(_source="tjhei/cracks tests/hetero_3d_1.prm, .mpirun-4.statistics, .mpirun-4.output; E field from test.pgm "
                     "via make_hetero_golden.py",
             statistics=rows,
             initial_newton_residual=[float(m) for m in re.findall(r"^0\t\t\t(\S+)$", out, flags=re.M)],
             cells=int(re.search(r"Timestep 0: .*Cells: (\d+)", out).group(1)),
             dofs_before_prerefinement=int(re.search(r"DoFs: \d+ solid \+ \d+ phase = (\d+)\nPrerefinement", out).group(1)),
             prerefinement_h=float(re.search(r"Prerefinement step with h= (\S+)", out).group(1)),
             cell_keys=[list(c) for c in run.forest.order], e_modulus=[float(v) for v in run.E_cells])
    json.dump(d, open(os.path.join(HERE, "hetero_3d_1.json"), "w"))


if __name__ == "__main__":
    main()
