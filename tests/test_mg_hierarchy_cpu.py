"""The multigrid hierarchy across ranks (pure host logic, no GPU): pf_mg_hierarchy reports the
levels and the inter-grid transfer ranges from the very functions the device code uses
(mg_inject_range / mg_restrict_range / mg_coarse_distributed_n in pf_api.cu).  Checked for every
rank count 1..8 on the Sneddon meshes: each coarse plane is produced exactly once, and no rank
reads a fine or coarse plane it does not hold."""
import pytest


def _mesh(pf, n):
    m = pf.Mesh()
    m.dim = 3
    for d in range(3):
        m.n[d], m.h[d], m.origin[d] = n[d], 20.0 / n[d], -10.0
    return m


CASES = [((160,) * 3, r) for r in range(1, 9)] + [((80,) * 3, r) for r in (1, 2, 3, 4, 8)] + \
        [((40, 40, 40), 8), ((20, 20, 20), 2), ((10, 10, 10), 2), ((10, 10, 10), 3), ((48, 32, 96), 6), ((16, 16, 30), 4)]


@pytest.mark.parametrize("n,nranks", CASES)
def test_hierarchy_is_consistent_across_ranks(pf, n, nranks):
    from cracks_b200.api import mg_hierarchy
    H = [mg_hierarchy(_mesh(pf, n), r, nranks) for r in range(nranks)]
    nlev = len(H[0])
    assert all(len(h) == nlev for h in H)
    for l in range(nlev):
        # same level shape and same decision on every rank (collective calls must line up)
        assert len({(h[l]["n"], h[l]["replicated"], h[l]["mode_below"]) for h in H}) == 1
    # the coarsest level cannot be halved any further under the library's rule
    assert H[0][-1]["mode_below"] == 0
    assert any(k % 2 or k // 2 < 4 for k in H[0][-1]["n"])
    for l in range(1, nlev):
        nn_f, nn_c = H[0][l - 1]["n"][2] + 1, H[0][l]["n"][2] + 1
        mode = H[0][l - 1]["mode_below"]
        assert mode in (1, 2)
        assert tuple(k * 2 for k in H[0][l]["n"]) == H[0][l - 1]["n"]
        ranks_f = range(nranks) if not H[0][l - 1]["replicated"] else [0]
        restricted, injected = [], []
        for r in ranks_f:
            F, Cl = H[r][l - 1], H[r][l]
            f_local = range(F["plane_begin"], F["plane_end"])
            f_owned = range(F["owned_begin"], F["owned_end"])
            c_local = range(Cl["plane_begin"], Cl["plane_end"])
            # restriction: coarse plane K reads the fine residual on planes 2K-1 .. 2K+1 (halo-exchanged)
            for K in range(*Cl["restrict"]):
                assert K in c_local
                for k in (2 * K - 1, 2 * K, 2 * K + 1):
                    if 0 <= k < nn_f:
                        assert k in f_local, (r, l, K, k)
                restricted.append(K)
            # injection reads fine plane 2K: any local plane (mode 1) / an owned plane (mode 2)
            for K in range(*Cl["inject"]):
                assert K in c_local and (2 * K in (f_owned if mode == 2 else f_local))
                injected.append(K)
            if mode == 1:
                # aligned decomposition: the coarse level owns half the cell layers of the fine one ...
                assert Cl["owned_end"] - 1 == (F["owned_end"] - 1) // 2 and Cl["replicated"] == F["replicated"]
                # ... injection covers everything but (possibly) the upper ghost plane, which the neighbour sends
                assert set(range(Cl["owned_begin"], Cl["owned_end"])) <= set(range(*Cl["inject"]))
                assert set(c_local) - set(range(*Cl["inject"])) <= {Cl["plane_end"] - 1}
                # prolongation: fine plane k reads coarse planes k//2 and (k+1)//2
                for k in f_local:
                    assert k // 2 in c_local and (k + 1) // 2 in c_local, (r, l, k)
            else:
                assert Cl["replicated"] and (Cl["plane_begin"], Cl["plane_end"]) == (0, nn_c)
        # every coarse plane is restricted into by exactly one rank; with a replicated coarse level the
        # all-reduce also relies on every plane being injected exactly once
        assert sorted(restricted) == list(range(nn_c)), (l, sorted(restricted))
        if mode == 2:
            assert sorted(injected) == list(range(nn_c))


def test_sneddon_refine4_on_8_ranks(pf):
    """The layout DESIGN.md quotes: 160, 80, 40 keep the z-slabs (20, 10, 5 layers per rank), 20, 10, 5 are replicated."""
    from cracks_b200.api import mg_hierarchy
    h = mg_hierarchy(_mesh(pf, (160,) * 3), 3, 8)
    assert [L["n"][0] for L in h] == [160, 80, 40, 20, 10, 5]
    assert [L["replicated"] for L in h] == [False, False, False, True, True, True]
    assert [L["mode_below"] for L in h] == [1, 1, 2, 1, 1, 0]
    assert [L["owned_end"] - L["owned_begin"] for L in h[:3]] == [20, 10, 5]
