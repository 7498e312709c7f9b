"""Several ranks of the library WITHOUT GPUs: every rank is a process running the CPU emulation of the library
(tests/emu/libcracks_b200_emu.so); "NCCL" is tests/emu/fake_nccl (Unix sockets), found by the library's own
dlopen("libnccl.so.2") through LD_LIBRARY_PATH.  What this exercises is the multi-rank host logic that only a
multi-GPU box runs otherwise: z-slab layout, halo exchanges, the multigrid hierarchy across ranks (levels that
keep the decomposition and levels replicated by all-reduce), all-reduced Krylov / active-set decisions.
The result must not depend on the number of ranks (up to FP reassociation)."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")


@pytest.fixture(scope="module")
def built():
    sys.path.insert(0, EMU)
    import build_emulated_library as b
    so = os.path.join(EMU, "libcracks_b200_emu.so")
    if not os.path.exists(so):
        b.build()
    fake = os.path.join(EMU, "fake_nccl", "libnccl.so.2")
    src = os.path.join(EMU, "fake_nccl", "fake_nccl.cc")
    if not os.path.exists(fake) or os.path.getmtime(src) > os.path.getmtime(fake):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", fake, src])
    return so


def _run(nranks, n, steps, tmp_path, timeout, **extra_env):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(EMU, "fake_nccl") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    env.update(extra_env)
    tag = "%d_%s_%s" % (nranks, "x".join(map(str, n)), "_".join(extra_env.values()))
    id_file, out_file = str(tmp_path / ("id_" + tag)), str(tmp_path / ("out_" + tag + ".json"))
    procs = [subprocess.Popen([sys.executable, os.path.join(EMU, "multirank_worker.py"), str(r), str(nranks), id_file, out_file,
                               *map(str, n), str(steps)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(nranks)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
    return json.load(open(out_file))


@pytest.fixture(scope="module")
def one_rank_8(built, tmp_path_factory):
    return _run(1, (8, 8, 8), 1, tmp_path_factory.mktemp("one"), 600)


def test_two_ranks_with_the_fp32_vcycle(built, one_rank_8, tmp_path):
    """PF_MG_FP32=1 on two ranks: float halo exchanges of the smoother input, the residual before the restriction
    and the coarse correction; the converged step is the FP64 single-rank one."""
    two = _run(2, (8, 8, 8), 1, tmp_path, 600, PF_MG_FP32="1")
    for a, b in zip(two["statistics"], one_rank_8["statistics"]):
        assert a["crack"] == pytest.approx(b["crack"], rel=1e-10)
        assert a["bulk"] == pytest.approx(b["bulk"], rel=1e-7)
    assert two["newton_its"] == one_rank_8["newton_its"]
    assert abs(two["linear_its"] - one_rank_8["linear_its"]) <= 0.2 * one_rank_8["linear_its"] + 3


@pytest.fixture(scope="module")
def one_rank_12(built, tmp_path_factory):
    return _run(1, (8, 8, 12), 1, tmp_path_factory.mktemp("one12"), 900)


@pytest.mark.parametrize("fp32", ["0", "1"])
def test_four_ranks_middle_slabs(built, one_rank_12, tmp_path, fp32):
    """12 cell layers on 4 ranks: ranks 1 and 2 own slabs with a neighbour on both sides, whose two boundary
    layers are evaluated by ONE strided launch after the halo exchange (Grid::layer_stride) while the interior
    layer overlaps it; the 4 x 4 x 6 level is replicated."""
    four = _run(4, (8, 8, 12), 1, tmp_path, 900, PF_MG_FP32=fp32)
    assert [(tuple(L[0]), L[1]) for L in four["levels"]] == [((8, 8, 12), False), ((4, 4, 6), True)]
    for a, b in zip(four["statistics"], one_rank_12["statistics"]):
        assert a["crack"] == pytest.approx(b["crack"], rel=1e-10)
        assert a["bulk"] == pytest.approx(b["bulk"], rel=1e-7)
    assert four["newton_its"] == one_rank_12["newton_its"]
    assert abs(four["linear_its"] - one_rank_12["linear_its"]) <= 0.2 * one_rank_12["linear_its"] + 3


def test_two_ranks_give_the_single_rank_result(built, one_rank_8, tmp_path):
    n = (8, 8, 8)
    one = one_rank_8
    two = _run(2, n, 1, tmp_path, 600)
    # 8 layers on 2 ranks: the 4 x 4 x 4 level keeps the z-slabs (mode 1, 2 layers per rank)
    assert [tuple(L[0]) for L in two["levels"]] == [(8, 8, 8), (4, 4, 4)] and two["levels"][0][2] == 1
    for a, b in zip(two["statistics"], one["statistics"]):
        assert a["crack"] == pytest.approx(b["crack"], rel=1e-10)
        assert a["bulk"] == pytest.approx(b["bulk"], rel=1e-7)
    assert two["newton_its"] == one["newton_its"]
    assert abs(two["linear_its"] - one["linear_its"]) <= 0.2 * one["linear_its"] + 3


@pytest.fixture(scope="module")
def one_rank_32(built, tmp_path_factory):
    return _run(1, (16, 16, 32), 1, tmp_path_factory.mktemp("one32"), 3000)


@pytest.mark.parametrize("fp32", ["0", "1"])
def test_eight_ranks_distributed_and_replicated_levels(built, one_rank_32, tmp_path, fp32):
    n = (16, 16, 32)
    one = one_rank_32
    eight = _run(8, n, 1, tmp_path, 3000, PF_MG_FP32=fp32, PF_EMU_THREADS="2")
    # 32 layers on 8 ranks: 8 x 8 x 16 keeps the slabs (2 layers per rank), 4 x 4 x 8 is replicated
    assert [(tuple(L[0]), L[1]) for L in eight["levels"]] == [((16, 16, 32), False), ((8, 8, 16), False), ((4, 4, 8), True)]
    for a, b in zip(eight["statistics"], one["statistics"]):
        assert a["crack"] == pytest.approx(b["crack"], rel=1e-10)
        assert a["bulk"] == pytest.approx(b["bulk"], rel=1e-7)
    assert eight["newton_its"] == one["newton_its"]


def _run_forest(nranks, case, tmp_path, timeout, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(EMU, "fake_nccl") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    id_prefix, out_file = str(tmp_path / ("fid_%s_%d" % (case, nranks))), str(tmp_path / ("fout_%s_%d.json" % (case, nranks)))
    procs = [subprocess.Popen([sys.executable, os.path.join(EMU, "multirank_forest_worker.py"), str(r), str(nranks), id_prefix,
                               out_file, case, *map(str, args)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(nranks)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=timeout)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
    return json.load(open(out_file))


def test_hetero_3d_golden_on_two_and_three_forest_ranks(built, tmp_path):
    """pf_create_forest_distributed (replicated vectors, cells cut into contiguous ranges, one all-reduce per operator
    application / residual / diagonal / functional) on KAT-5, tests/hetero_3d_1.mpirun-4.statistics: 932 cells with
    318 hanging nodes and per-cell Lame coefficients.  3 ranks: ranges of unequal length."""
    g = json.load(open(os.path.join(HERE, "golden", "hetero_3d_1.json")))
    for nranks in (2, 3):
        out = _run_forest(nranks, "hetero", tmp_path, 1500)
        for got, ref in zip(out["statistics"], g["statistics"]):
            assert got["dofs"] == 5288
            assert got["crack"] == pytest.approx(ref["crack"], rel=1e-7)
            assert got["bulk"] == pytest.approx(ref["bulk"], rel=1e-6)


def test_adaptive_miehe_shear_on_two_forest_ranks(built, tmp_path):
    """BASELINE config 4 in small on 2 ranks: stress split + predictor-corrector refinement (a new distributed
    context, i.e. a new communicator, after every mesh change); the first rows of tests/miehe_shear_1.statistics"""
    g = json.load(open(os.path.join(HERE, "golden", "miehe_shear_1.json")))
    out = _run_forest(2, "shear", tmp_path, 2400, 6)
    assert [r["dofs"] for r in out["statistics"]] == [r["dofs"] for r in g["statistics"][:7]]
    for got, ref in zip(out["statistics"], g["statistics"]):
        for k in ("bulk", "crack", "load"):
            assert got[k] == pytest.approx(ref[k], rel=1e-6), (got["step"], k)
