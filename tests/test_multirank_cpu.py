"""World-size-2 (and 3) tests of the multi-GPU path's host logic on CPU with gloo:
the slab decomposition exported by the library (pf_slab_layout) plus the halo
pattern of pf_api.cu::halo_exchange must reproduce the single-rank result on
every owned row.  The cell arithmetic is played by the CPU oracle on each
rank's sub-box; the exchange uses torch.distributed send/recv in the same
order as the NCCL group in the library."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, h, result_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import newton_oracle as orc
    import cracks_b200 as pf
    from cracks_b200.api import slab_layout

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dim = 3
    lo = (0.0, 0.0, 0.0)
    hi = tuple(n[d] * h[d] for d in range(3))
    glob = orc.Problem(dim, n, lo, hi, kappa_of_h=lambda hh: 1e-3, pressure=1e-3)
    rng = np.random.default_rng(42)                       # same global data on every rank
    nn, nc = glob.n_nodes, 4
    sol = np.zeros((nn, nc))
    sol[:, :3] = 1e-2 * rng.standard_normal((nn, 3))
    sol[:, 3] = rng.random(nn)
    sol = sol.reshape(-1)
    old = sol + 1e-2 * rng.standard_normal(nn * nc)
    con = glob.dirichlet_mask().reshape(nn, nc)
    con[rng.random(nn) < 0.2, 3] = 1
    con = np.ascontiguousarray(con.reshape(-1))
    x = rng.standard_normal(nn * nc)
    y_ref = glob.apply_jacobian(sol, old, old, con, x)
    e_ref = glob.energy(sol)

    mesh = pf.Mesh()
    mesh.dim = 3
    for d in range(3):
        mesh.n[d], mesh.h[d], mesh.origin[d] = n[d], h[d], lo[d]
    lay = slab_layout(mesh, rank, world)
    npp = lay["nodes_per_plane"]
    p0, p1 = lay["plane_begin"], lay["plane_end"]
    o0, o1 = lay["owned_begin"], lay["owned_end"]
    sl = slice(p0 * npp * nc, p1 * npp * nc)

    # local sub-box = the cell layers this rank evaluates
    nloc = (n[0], n[1], lay["cell_end"] - lay["cell_begin"])
    lo_l = (0.0, 0.0, lay["cell_begin"] * h[2])
    hi_l = (hi[0], hi[1], lay["cell_end"] * h[2])
    loc = orc.Problem(dim, nloc, lo_l, hi_l, kappa_of_h=lambda hh: 1e-3, pressure=1e-3)
    loc.prm = glob.prm
    assert loc.n_nodes == (p1 - p0) * npp

    # x is only valid on owned planes; ghosts arrive through the halo exchange
    xl = np.full((p1 - p0) * npp * nc, np.nan)
    xl[(o0 - p0) * npp * nc:(o1 - p0) * npp * nc] = x[o0 * npp * nc:o1 * npp * nc]
    cnt = npp * nc

    def plane(gp):
        return torch.from_numpy(xl[(gp - p0) * cnt:(gp - p0 + 1) * cnt])

    reqs = []
    if rank > 0:
        reqs.append(dist.irecv(plane(p0), src=rank - 1))
        reqs.append(dist.isend(plane(o0).clone(), dst=rank - 1))
    if rank < world - 1:
        reqs.append(dist.irecv(plane(p1 - 1), src=rank + 1))
        reqs.append(dist.isend(plane(o1 - 1).clone(), dst=rank + 1))
    for r in reqs:
        r.wait()
    assert not np.isnan(xl).any()
    assert np.array_equal(xl, x[sl])

    yl = loc.apply_jacobian(sol[sl].copy(), old[sl].copy(), old[sl].copy(), np.ascontiguousarray(con[sl]), xl)
    own = slice((o0 - p0) * cnt, (o1 - p0) * cnt)
    err = np.max(np.abs(yl[own] - y_ref[o0 * cnt:o1 * cnt])) / np.max(np.abs(y_ref))

    # functionals: owned cells only, summed over ranks
    own_cells = orc.Problem(dim, (n[0], n[1], lay["own_cell_end"] - lay["own_cell_begin"]),
                            (0.0, 0.0, lay["own_cell_begin"] * h[2]), (hi[0], hi[1], lay["own_cell_end"] * h[2]),
                            kappa_of_h=lambda hh: 1e-3)
    own_cells.prm = glob.prm
    s0 = lay["own_cell_begin"] * cnt
    s1 = (lay["own_cell_end"] + 1) * cnt
    e = torch.tensor(own_cells.energy(sol[s0:s1].copy()), dtype=torch.float64)
    dist.all_reduce(e)
    owned = torch.tensor([o1 - o0], dtype=torch.int64)
    dist.all_reduce(owned)
    np.save(os.path.join(result_dir, f"r{rank}.npy"),
            np.array([err, abs(e[0].item() - e_ref[0]) / abs(e_ref[0]), abs(e[1].item() - e_ref[1]) / abs(e_ref[1]),
                      owned.item()]))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, (5, 4, 6)), (3, (4, 3, 7)), (2, (3, 3, 2))])
def test_slab_decomposition_matches_single_rank(pf, oracle, tmp_path, world, n):
    import torch.multiprocessing as mp
    port = 29500 + os.getpid() % 2000 + world
    mp.spawn(_worker, args=(world, port, n, (0.5, 0.4, 0.3), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err, eb, ec, owned = np.load(tmp_path / f"r{r}.npy")
        assert err <= 1e-13
        assert eb <= 1e-12 and ec <= 1e-12
        assert owned == n[2] + 1            # every node plane is owned exactly once


def test_slab_layout_properties(pf):
    from cracks_b200.api import slab_layout
    m = pf.sneddon_mesh(3, 4)
    for world in (1, 2, 4, 8, 7):
        lays = [slab_layout(m, r, world) for r in range(world)]
        assert lays[0]["owned_begin"] == 0 and lays[-1]["owned_end"] == 161
        for a, b in zip(lays, lays[1:]):
            assert a["owned_end"] == b["owned_begin"]              # owned planes tile [0, 161)
            assert b["plane_begin"] == a["owned_end"] - 1          # lower ghost = neighbour's last owned plane
            assert a["plane_end"] - 1 == b["owned_begin"]          # upper ghost = neighbour's first owned plane
            assert a["cell_end"] == a["own_cell_end"] + 1           # one redundant layer
        assert sum(l["own_cell_end"] - l["own_cell_begin"] for l in lays) == 160
    with pytest.raises(pf.PFError):
        slab_layout(pf.sneddon_mesh(3, 0), 0, 11)                  # more ranks than cell layers
