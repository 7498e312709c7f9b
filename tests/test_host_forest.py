"""The C++ host forest (cracks_b200/host/forest.{h,cc}: local refinement with 2:1 balance, node numbering,
hanging-node table, solution transfer -- the product-side stand-in for the reference's p4est forest and
make_hanging_node_constraints, cracks.cc:3895-4163, 1630-1634) against the CPU oracle's forests, which are
pinned to the reference's goldens (tests/test_oracle_adaptive.py): same cells, same connectivity, same
hanging nodes on the KAT-2 and KAT-5 meshes and on a refined slit mesh.  Pure host code, no GPU."""
import json
import math
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dump():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "cracks_b200", "host"), "-s", "../forest_dump"])
    exe = os.path.join(ROOT, "cracks_b200", "forest_dump")
    return lambda *args: json.loads(subprocess.check_output([exe, *args], text=True))


@pytest.fixture(scope="module")
def ao(oracle):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import adaptive_oracle
    return adaptive_oracle


def _compare(d, forest, dim):
    cells, cell_h, xy, hanging = forest.build()
    nv = 1 << dim
    assert d["n_cells"] == cells.shape[0] and d["n_nodes"] == xy.shape[0]
    keys = [tuple(c) + (0,) * (4 - len(c)) for c in forest.order]
    assert [tuple(c) for c in d["cells"]] == keys
    assert np.array_equal(np.array(d["conn"]).reshape(-1, nv), cells)
    assert np.allclose(np.array(d["coords"]).reshape(-1, dim), xy, rtol=0, atol=1e-14)
    got = {h[0]: tuple(sorted(h[1:])) for h in d["hanging"]}
    ref = {h: tuple(sorted(p)) for h, p in hanging.items()}
    assert got == ref
    assert d["h_min"] == pytest.approx(float(np.min(np.sqrt((cell_h ** 2).sum(axis=1)))), rel=1e-15)


def test_kat2_mesh(dump, ao):
    f = ao.QuadForest(10, 10, (-10.0, -10.0), (10.0, 10.0))
    f.refine(ao.flag_fixed_preref_sneddon(f))
    d = dump("kat2")
    _compare(d, f, 2)
    assert d["n_cells"] == 124 and d["n_nodes"] * 3 == 453 and len(d["hanging"]) == 12     # tests/sneddon_2d_1.output


def test_kat5_mesh(dump, ao):
    f = ao.OctForest((1, 1, 1), (0.0,) * 3, (10.0,) * 3)
    for _ in range(3):
        f.refine(list(f.cells))
    cells, cell_h, xyz, _ = f.build()
    h = float(np.min(np.sqrt((cell_h ** 2).sum(axis=1))))
    phi = ao.initial_multiple_het_3d(xyz, h)
    f.refine([c for k, c in enumerate(f.order) if np.any(phi[cells[k]] < 0.4)])
    d = dump("kat5")
    _compare(d, f, 3)
    assert d["n_cells"] == 932 and d["n_nodes"] * 4 == 5288                                 # tests/hetero_3d_1 output
    assert sorted({len(h) - 1 for h in d["hanging"]}) == [2, 4]


@pytest.mark.parametrize("refine", [1, 3])
def test_refined_slit_mesh(dump, ao, refine):
    f = ao.QuadForest(2, 2, (0.0, 0.0), (1.0, 1.0), slit=True)
    for _ in range(refine):
        f.refine(list(f.cells))
    for radius in (0.3, 0.12):
        f.build()
        flagged = []
        for c in f.order:
            x, y, hx, hy = f.cell_box(c)
            if math.hypot(x + 0.5 * hx - 0.7, y + 0.5 * hy - 0.45) < radius:
                flagged.append(c)
        f.refine(flagged)
    d = dump("slit", str(refine))
    _compare(d, f, 2)
    # the two sides of the slit are not connected: no hanging node has parents on both sides
    xy = np.array(d["coords"]).reshape(-1, 2)
    conn = np.array(d["conn"]).reshape(-1, 4)
    dup = [n for n in range(d["n_nodes"]) if np.sum(np.all(xy == xy[n], axis=1)) == 2]
    assert dup and all(xy[n, 1] == 0.5 and xy[n, 0] > 0.5 for n in dup)
    for n in dup:
        users = np.unique(np.where(conn == n)[0])
        ys = xy[conn[users]][:, :, 1]
        assert np.all(ys >= 0.5) or np.all(ys <= 0.5)


def test_solution_transfer_reproduces_linear_fields(dump):
    d = dump("transfer")
    assert d["fine_cells"] > d["coarse_cells"] and d["fine_hanging"] > 0
    assert d["max_error"] <= 1e-14


def test_ctypes_mirror_of_the_host_forest(dump, pf):
    """cracks_b200/forest.py (HostForest) hands out the same tables as the C++ class"""
    from cracks_b200.forest import ForestSneddonDriver
    f = ForestSneddonDriver.prerefined_forest()            # the KAT-2 recipe through the C binding
    d = dump("kat2")
    t = f.tables()
    assert (f.n_cells, f.n_nodes, f.n_hanging) == (d["n_cells"], d["n_nodes"], len(d["hanging"])) == (124, 151, 12)
    assert np.array_equal(t["conn"].reshape(-1), np.array(d["conn"]))
    assert np.allclose(t["coords"].reshape(-1), np.array(d["coords"]), rtol=0, atol=0)
    assert [[v for v in row if v >= 0] for row in t["hanging"].tolist()] == d["hanging"]
    assert np.allclose(t["level_h"], [[2.0, 2.0], [1.0, 1.0]])
    finer = f.clone()
    flags = np.zeros(finer.n_cells, dtype=np.uint8)
    flags[::7] = 1
    finer.refine(flags)
    xy0, xy1 = t["coords"], finer.tables()["coords"]
    lin = lambda xy: np.stack([1 + 2 * xy[:, 0] - xy[:, 1], 3 - xy[:, 0]], axis=1).reshape(-1)
    assert np.allclose(finer.transfer_from(f, lin(xy0), 2), lin(xy1), rtol=0, atol=1e-13)


def test_cpp_bitmap_function_reproduces_the_fixture_field(dump):
    """cracks_b200/host/bitmap_function.{h,cc} (BitmapFile / BitmapFunction<3> of cracks.cc:118-241 with its
    quirks) evaluated on the reference's test.pgm at the KAT-5 cell centres equals the field the hetero_3d_1
    fixture carries (which reproduces the golden).  Needs the reference tree: skipped where it is absent."""
    pgm = "/root/reference/test.pgm"
    if not os.path.exists(pgm):
        pytest.skip("the reference's test.pgm is only present in the build container")
    exe = os.path.join(ROOT, "cracks_b200", "forest_dump")
    got = np.array(json.loads(subprocess.check_output([exe, "bitmap", pgm], text=True)))
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "hetero_3d_1.json")))
    assert got.shape[0] == 932
    assert np.allclose(got, np.array(g["e_modulus"]), rtol=1e-14, atol=0)
