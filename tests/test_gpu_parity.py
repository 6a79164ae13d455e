"""Parity of the CUDA path (through the C-ABI, via the reference-shaped Python API)
against golden vectors of the unmodified reference and against the CPU oracle.

Bar: bin codes, selected gene sequences bit-exact; entropies / scores compared with
np.array_equal (bit-identical fp64) — the 1e-9 relative tolerance of the north star is
the fallback documented in DESIGN.md, not what these tests accept."""
import ctypes
import os

import numpy as np
import pytest
import scipy.sparse

import fixtures
from helpers import check_entropies, check_selectors, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pr():
    import picturedrocks_b200 as pr

    return pr


@pytest.fixture(scope="module")
def O():
    from oracle import oracle

    return oracle


def make_gpu(pr):
    return lambda kind, infoset: getattr(pr.markers, kind)(infoset)


def test_native_library_is_loaded(pr):
    from picturedrocks_b200 import _native

    lib = _native.load()
    assert lib.prb_abi_version() == 3
    with open("/proc/self/maps") as f:
        assert "libprb200.so" in f.read()


def test_sampleX(pr, golden):
    g = golden("sampleX")
    X = g["X"]
    check_entropies(pr.markers.SparseInformationSet(X), g, "sparse_")
    check_entropies(pr.markers.InformationSet(X, False), g, "dense_")
    s = pr.markers.SparseInformationSet(X)
    H = s.entropy_wrt(np.arange(0))
    assert np.issubdtype(H.dtype, np.floating)
    # python lists are accepted for cols (reference tests/test_entropy.py:173, 182)
    assert np.array_equal(s.entropy_wrt([2]), g["sparse_wrt_c_2"])
    assert s.entropy([4, 5]) == float(g["sparse_ent_c_4_5"])


@pytest.mark.parametrize("name", ["xor", "diminish_red"])
def test_selector_known_answers(pr, golden, name):
    g = golden(name)
    inf = pr.markers.SparseInformationSet(g["X"])
    inf.set_y(g["y"])
    check_entropies(inf, g)
    check_selectors(make_gpu(pr), inf, g, 2)
    for kind in ("CIFE", "JMI"):
        fs = getattr(pr.markers, kind)(inf)
        fs.autoselect(1)
        assert int(fs.S[0]) == int(g[kind + "_traj_first"])
        fs.remove(int(fs.S[0]))
        fs.add(0)
        assert np.array_equal(fs.score, g[kind + "_traj_score"])
        fs.autoselect(1)
        assert [int(s) for s in fs.S] == list(g[kind + "_traj_S"])


def test_reference_test_fs_expectations(pr):
    """Literal expectations of the reference's tests/test_fs.py:39-91."""
    X, y = fixtures.diminish_red_table()
    inf = pr.markers.SparseInformationSet(X)
    inf.set_y(y)
    cife, jmi, mim = pr.markers.CIFE(inf), pr.markers.JMI(inf), pr.markers.MIM(inf)
    cife.autoselect(2)
    jmi.autoselect(2)
    assert set(cife.S) == {0, 1} and set(jmi.S) == {0, 1}
    ninf = float("-inf")
    assert np.allclose(cife.score, [ninf, ninf, 1, 0])
    assert np.allclose(jmi.score, [ninf, ninf, 1, 0.5])
    assert np.allclose(mim.score, [2, 2, 1, 1])
    for objective in (pr.markers.CIFE, pr.markers.JMI, pr.markers.MIM, pr.markers.CIFEUnsup,
                      pr.markers.UniEntropy):
        fs = objective(inf)
        fs.autoselect(2)
        assert set(fs.S) == {0, 1}
        fs.remove(0)
        assert len(fs.S) == 1
        fs.autoselect(1)
        assert fs.S[-1] == 0
        fs.remove(1)
        assert len(fs.S) == 1
        fs.autoselect(1)
        assert fs.S[-1] == 1


def test_linearxy(pr, golden):
    g = golden("linearxy")
    X, y, _ = fixtures.linearxy()
    assert sha(X) == str(g["X_sha"])
    Xq = pr.markers.quantile_discretize(X)
    assert Xq.dtype == np.int64 and Xq.shape == X.shape
    assert np.array_equal(Xq.astype(np.uint8), g["codes"])
    inf = pr.markers.SparseInformationSet(Xq, y)
    check_entropies(inf, g, "sparse_")
    dense = pr.markers.InformationSet(np.concatenate((Xq, y[:, None]), axis=1), True)
    check_entropies(dense, g, "dense_")
    assert np.array_equal([inf.entropy(np.array([i])) for i in range(20)], g["ent_single"])
    assert np.array_equal([inf.entropy([-1, i]) for i in range(20)], g["ent_y_single"])
    # selectors on the dense set (label column = last column) take the fused step as well and
    # give the bits of the sparse set's (which the goldens pin)
    for kind in ("CIFE", "JMI", "MIM", "CIFEUnsup"):
        a, b = getattr(pr.markers, kind)(dense), getattr(pr.markers, kind)(inf)
        if hasattr(a, "_d"):
            assert a._d.fused
        a.autoselect(7)
        b.autoselect(7)
        if kind in ("CIFE", "JMI"):
            a.remove(a.S[2])
            b.remove(b.S[2])
            a.autoselect(2)
            b.autoselect(2)
        assert list(a.S) == list(b.S), kind
        assert np.array_equal(a.score, b.score), kind


def test_planted600(pr, golden):
    g = golden("planted600")
    X, y = g["X"], g["y"]
    Xq = pr.markers.quantile_discretize(X)
    assert np.array_equal(Xq.astype(np.uint8), g["codes"])
    inf = pr.markers.SparseInformationSet(Xq, y)
    check_entropies(inf, g)
    check_selectors(make_gpu(pr), inf, g, 12)
    for kind in ("CIFE", "JMI", "CIFEUnsup"):
        fs = getattr(pr.markers, kind)(inf)
        fs.autoselect(4)
        fs.remove(int(fs.S[1]))
        fs.add(17)
        fs.autoselect(3)
        assert [int(s) for s in fs.S] == list(g[kind + "_mix_S"])
        assert np.array_equal(fs.score, g[kind + "_mix_score"])
        assert np.array_equal(fs.penalty, g[kind + "_mix_penalty"])
    mim = pr.markers.MIM(inf)
    assert mim.base_score[7] == mim.base_score[3]
    assert mim.base_score[9] == 0.0


def test_makeinfoset_planted(pr, golden):
    """makeinfoset(adata-like) keeps codes on the GPU; same selections as the reference."""
    import types

    import pandas as pd

    g = golden("planted600")
    adata = types.SimpleNamespace(X=g["X"], obs=pd.DataFrame({"y": g["y"]}))
    inf = pr.markers.makeinfoset(adata, True)
    assert isinstance(inf, pr.markers.SparseInformationSet) and inf.has_y
    check_selectors(make_gpu(pr), inf, g, 12, kinds=("CIFE", "JMI"))
    assert np.array_equal(np.asarray(inf.X.todense()).astype(np.uint8), g["codes"])
    inf2 = pr.markers.makeinfoset(types.SimpleNamespace(X=scipy.sparse.csr_matrix(g["X"]),
                                                        obs=adata.obs), False)
    assert not inf2.has_y
    assert np.array_equal(inf2.entropy_wrt(np.arange(0)), g["wrt_c_empty"])


def test_discretize_dtypes(pr, golden):
    g = golden("discretize")
    Xc = fixtures.continuous_matrix(5)
    for dt in (np.float32, np.float64):
        for k in (2, 3, 5, 10, 20):
            got = pr.markers.quantile_discretize(Xc.astype(dt), k).astype(np.uint8)
            assert np.array_equal(got, g["codes_%s_k%d" % (np.dtype(dt).name, k)]), (dt, k)
    Xi = g["int_X"]
    for k in (3, 5, 8):
        assert np.array_equal(pr.markers.quantile_discretize(Xi, k).astype(np.uint8),
                              g["codes_int64_k%d" % k])
    sp = scipy.sparse.csr_matrix(Xc.astype(np.float32))
    assert np.array_equal(pr.markers.quantile_discretize(sp, 5).astype(np.uint8),
                          g["codes_float32_k5"])


def test_config1(pr, golden):
    """BASELINE.json configs[0]: 2730 x 3451, 19 clusters, makeinfoset then CIFE top-20."""
    import types

    import pandas as pd
    from picturedrocks_b200.synth import poisson_gamma

    g = golden("config1")
    X, y = poisson_gamma(2730, 3451, 19, 1)
    assert sha(X) == str(g["X_sha"])
    inf = pr.markers.makeinfoset(types.SimpleNamespace(X=X, obs=pd.DataFrame({"y": y})), True)
    codes = np.asarray(inf.X.todense()).astype(np.uint8)
    assert sha(codes) == str(g["codes_sha"])
    check_entropies(inf, g)
    check_selectors(make_gpu(pr), inf, g, 20, kinds=("CIFE", "JMI", "MIM"))


@pytest.mark.parametrize("name,kw,n,kinds", [
    ("synth_sparse", dict(N=3000, P=400, C=10, K=5, mean_density=0.05, seed=3), 15,
     ("CIFE", "JMI", "CIFEUnsup")),
    ("synth_sparse_k20", dict(N=2000, P=300, C=12, K=20, mean_density=0.08, seed=5), 8,
     ("CIFE", "JMI")),
])
def test_synth_device_generation_and_selection(pr, golden, name, kw, n, kinds):
    from picturedrocks_b200.synth import DiscretisedSpec

    spec = DiscretisedSpec(**kw)
    g = golden(name)
    eng, yd = spec.device_engine()
    y = spec.cpu_labels()
    assert np.array_equal(yd.cpu().numpy(), y)
    Xs = spec.cpu_csc(range(spec.P), y)
    Xg = eng.to_scipy_csc()
    assert np.array_equal(Xg.indptr, Xs.indptr)
    assert np.array_equal(Xg.indices, Xs.indices)
    assert np.array_equal(Xg.data, Xs.data)
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    check_entropies(inf, g)
    check_selectors(make_gpu(pr), inf, g, n, kinds=kinds)


def test_c_abi_host_entry_matches_oracle(O):
    """prb_sparse_entropy_wrt_host: the host-buffer drop-in for _sparse_entropy_wrt."""
    from picturedrocks_b200 import _native as nat
    from picturedrocks_b200.synth import DiscretisedSpec

    spec = DiscretisedSpec(N=5000, P=300, C=7, K=5, mean_density=0.06, seed=9)
    y = spec.cpu_labels()
    X = spec.cpu_csc(range(spec.P), y)
    inf = O.SparseInformationSet(X, y)
    for cols in ([], [-1], [4], [-1, 4], [4, 8], [-1, 4, 8]):
        has_y = bool(cols) and cols[0] == -1
        mc = cols[1:] if has_y else cols
        mi, md = O.make_master_col(inf.X, mc, inf._shift)
        out = np.empty(spec.P)
        ind = inf.X.indices.astype(np.int32)
        ptr = inf.X.indptr.astype(np.int64)
        dat = inf.X.data.astype(np.int64)
        cs = inf.classsizes if has_y else np.zeros(0)
        nat.call("prb_sparse_entropy_wrt_host", ind.ctypes.data, ptr.ctypes.data, dat.ctypes.data,
                 mi.ctypes.data, md.ctypes.data, mi.size, spec.N, spec.P, len(mc), inf._shift,
                 inf._ybits if has_y else 0, inf.y.ctypes.data if has_y else None,
                 cs.ctypes.data if has_y else None, cs.size, out.ctypes.data)
        assert np.array_equal(out, inf.entropy_wrt(np.array(cols, dtype=np.int64))), cols


def test_c_abi_host_scalar_entropy_matches_oracle(O):
    """prb_sparse_entropy_host: the host-buffer drop-in for _sparse_entropy (INFOSET:413-428),
    driven the way SparseInformationSet.entropy drives the reference kernel."""
    from picturedrocks_b200 import _native as nat
    from picturedrocks_b200.synth import DiscretisedSpec

    spec = DiscretisedSpec(N=7000, P=40, C=6, K=5, mean_density=0.2, seed=2)
    y = spec.cpu_labels()
    inf = O.SparseInformationSet(spec.cpu_csc(range(spec.P), y), y)
    for cols in ([3], [3, 9], [0, 5, 11]):
        mi, md = O.make_master_col(inf.X, cols, inf._shift)
        max_val = 2 ** (inf._shift * len(cols))
        out = np.empty(1)
        md = np.ascontiguousarray(md, dtype=np.int64)
        nat.call("prb_sparse_entropy_host", md.ctypes.data, md.size, spec.N, max_val,
                 out.ctypes.data)
        assert out[0] == O.sparse_entropy(mi, md, spec.N, max_val)
        assert out[0] == inf.entropy(np.array(cols, dtype=np.int64))
    out = np.empty(1)
    nat.call("prb_sparse_entropy_host", None, 0, 10, 7, out.ctypes.data)
    assert out[0] == 0.0  # a constant column
    with pytest.raises(nat.PrbError):
        bad = np.array([9], dtype=np.int64)
        nat.call("prb_sparse_entropy_host", bad.ctypes.data, 1, 10, 7, out.ctypes.data)


def test_medium_sparse_against_oracle(pr, O):
    """60 classes, skewed densities: GPU vs live oracle, entropies and sequences."""
    from picturedrocks_b200.synth import DiscretisedSpec

    spec = DiscretisedSpec(N=120000, P=1500, C=60, K=5, mean_density=0.07, seed=4)
    eng, yd = spec.device_engine()
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    X = inf.X
    nt = O.max_threads()
    ora = O.SparseInformationSet(X, yd.cpu().numpy().astype(np.int64), nthreads=nt)
    for cols in ([], [-1], [11], [-1, 11]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(inf.entropy_wrt(c), ora.entropy_wrt(c)), cols
    for kind in ("CIFE", "JMI"):
        a, b = getattr(pr.markers, kind)(inf), getattr(O, kind)(ora)
        a.autoselect(6)
        b.autoselect(6)
        assert a.S == [int(s) for s in b.S]
        assert np.array_equal(a.score, b.score)
    # planted structure: duplicated genes tie exactly, all-zero genes score exactly 0
    mim = pr.markers.MIM(inf)
    assert mim.base_score[4018] == mim.base_score[4017] if spec.P > 4018 else True


def test_edge_cases(pr):
    # single non-zero entry, empty genes, ragged columns
    X = np.zeros((37, 5), dtype=np.int64)
    X[3, 1] = 2
    X[:, 3] = np.arange(37) % 3
    y = (np.arange(37) % 4).astype(np.int64)
    from oracle import oracle as O

    a = pr.markers.SparseInformationSet(X, y)
    b = O.SparseInformationSet(X, y)
    for cols in ([], [-1], [1], [3], [-1, 3], [0], [-1, 0], [1, 3]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(a.entropy_wrt(c), b.entropy_wrt(c)), cols
        assert a.entropy(c) == b.entropy(c), cols
    with pytest.raises(AssertionError):
        pr.markers.SparseInformationSet(X.astype(np.float64))
    with pytest.raises(AssertionError):
        pr.markers.SparseInformationSet(X, y.astype(np.float64))
    with pytest.raises(AssertionError):
        pr.markers.CIFE(pr.markers.SparseInformationSet(X))
    # set_y re-targets labels in place (binary selection use, notebook cell 44)
    a.set_y((y == 1).astype(np.int64))
    b.set_y((y == 1).astype(np.int64))
    assert np.array_equal(a.entropy_wrt([-1]), b.entropy_wrt(np.array([-1])))


@pytest.mark.parametrize("tile_cells,vlayout", [(None, True), (4096, True), (256, True),
                                                 (None, False)])
def test_class_sorted_cells_and_v_layout(pr, O, monkeypatch, tile_cells, vlayout):
    """csrc/vstream.cu (K <= 5): cells relabelled by class inside the V layout, class segments
    split and packed into super-tiles of various sizes, tile-major padded runs, label
    re-targeting (layout rebuilt), export of X — all bit-identical to the oracle;
    vlayout=False runs the same checks through the generic shared-histogram kernel."""
    import picturedrocks_b200.engine as E
    from picturedrocks_b200.synth import DiscretisedSpec

    monkeypatch.setattr(E, "USE_VLAYOUT", vlayout)
    if tile_cells:
        monkeypatch.setenv("PRB_V_TILE_CELLS", str(tile_cells))
    spec = DiscretisedSpec(N=30011, P=257, C=7, K=5, mean_density=0.09, seed=11)
    eng, yd = spec.device_engine()
    assert eng.vmode == vlayout
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    y = spec.cpu_labels()
    Xs = spec.cpu_csc(range(spec.P), y)
    Xg = eng.to_scipy_csc()  # the exported matrix keeps the original cell ids
    assert (eng.newpos is not None) == vlayout  # cells are relabelled inside the V layout only
    assert np.array_equal(Xg.indptr, Xs.indptr) and np.array_equal(Xg.indices, Xs.indices)
    assert np.array_equal(Xg.data, Xs.data)
    ora = O.SparseInformationSet(Xs, y)
    rng = np.random.RandomState(5)
    y_skew = np.where(rng.rand(spec.N) < 0.8, 0, rng.randint(1, 13, spec.N)).astype(np.int64)
    for labels in (None, y_skew, (y == 3).astype(np.int64)):
        if labels is not None:
            inf.set_y(labels)
            ora.set_y(labels)
        for cols in ([], [-1], [5], [-1, 5], [200], [-1, 200], [5, 9], [-1, 5, 9]):
            c = np.array(cols, dtype=np.int64)
            assert np.array_equal(inf.entropy_wrt(c), ora.entropy_wrt(c)), cols
            assert inf.entropy(c) == ora.entropy(c), cols
        for kind in ("CIFE", "JMI", "CIFEUnsup"):
            a, b = getattr(pr.markers, kind)(inf), getattr(O, kind)(ora)
            a.autoselect(9)  # crosses GRAPH_MIN_STEPS: eager step, capture, replays
            b.autoselect(9)
            assert a.S == [int(s) for s in b.S], kind
            assert np.array_equal(a.score, b.score), kind
    Xg = eng.to_scipy_csc()
    assert np.array_equal(Xg.indices, Xs.indices) and np.array_equal(Xg.data, Xs.data)


@pytest.mark.parametrize("layout", ["ell", "ell-cap64", "ell-pool", "ell-int32hits", "ell1",
                                    "rows16", "rows32"])
@pytest.mark.parametrize("tile_cells", [None, 4096])
def test_v_layout_variants(pr, O, monkeypatch, tile_cells, layout):
    """The streamed layouts of the K <= 5 path: the streamed lane-runs of csrc/vell2.cu (the
    default; also with lane-runs cut at 64 entries — pieces of runs, added with REDs —, with the
    work pool forced on, and with the x_j-major int32 hit table), the lane-runs of csrc/vell.cu
    (global work queue) and warp-per-run rows with 16- or 32-bit entries (csrc/vstream.cu):
    entropies, scores and sequences bit-identical to the oracle."""
    from picturedrocks_b200.synth import DiscretisedSpec

    monkeypatch.setenv("PRB_V_KERNEL", "ell" if layout.startswith("ell") else "rows")
    monkeypatch.setenv("PRB_V16", "0" if layout == "rows32" else "1")
    if layout == "ell-cap64":
        monkeypatch.setenv("PRB_ELL_CAP", "64")
    if layout == "ell-pool":
        monkeypatch.setenv("PRB_ELL_POOL_MIN", "0")
        monkeypatch.setenv("PRB_ELL_POOL_UNIT", "60")
    if layout == "ell-int32hits":
        monkeypatch.setenv("PRB_ELL_VMAJOR", "0")
    if layout == "ell1":
        monkeypatch.setenv("PRB_ELL_V", "1")
    if tile_cells:
        monkeypatch.setenv("PRB_V_TILE_CELLS", str(tile_cells))
    spec = DiscretisedSpec(N=70001, P=301, C=7, K=5, mean_density=0.09, seed=12)  # several tiles
    eng, yd = spec.device_engine()
    assert eng.ell == layout.startswith("ell") and eng.v16 == (layout == "rows16")
    assert eng.ell2 == (layout.startswith("ell") and layout != "ell1")
    assert eng.vmajor == (eng.ell2 and layout != "ell-int32hits")
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    y = spec.cpu_labels()
    ora = O.SparseInformationSet(spec.cpu_csc(range(spec.P), y), y)
    for cols in ([], [-1], [5], [-1, 5], [200], [-1, 200]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(inf.entropy_wrt(c), ora.entropy_wrt(c)), cols
    for kind in ("CIFE", "JMI", "CIFEUnsup"):
        a, b = getattr(pr.markers, kind)(inf), getattr(O, kind)(ora)
        a.autoselect(9)
        b.autoselect(9)
        assert a.S == [int(s) for s in b.S], kind
        assert np.array_equal(a.score, b.score), kind
    if layout == "ell-pool":
        assert eng.v["pool1"] > eng.v["pool0"]  # the pool really was in use


@pytest.mark.parametrize("K,C", [(6, 3), (10, 20), (17, 5), (20, 20), (21, 2)])
def test_lane_run_layout_more_bins(pr, O, K, C):
    """K = 6..21 bins on the lane-run layout (2..5 counter registers per lane, csrc/vell2.cu):
    the K > 5 points of BASELINE config 5 no longer go through the shared-histogram kernel."""
    from picturedrocks_b200.synth import DiscretisedSpec

    spec = DiscretisedSpec(N=30011, P=211, C=C, K=K, mean_density=0.12, seed=K)
    eng, yd = spec.device_engine()
    assert eng.ell and eng.vmode
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    y = spec.cpu_labels()
    ora = O.SparseInformationSet(spec.cpu_csc(range(spec.P), y), y)
    for cols in ([], [-1], [3], [-1, 3], [3, 7]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(inf.entropy_wrt(c), ora.entropy_wrt(c)), cols
    for kind in ("JMI", "CIFEUnsup"):
        a, b = getattr(pr.markers, kind)(inf), getattr(O, kind)(ora)
        a.autoselect(9)
        b.autoselect(9)
        assert a.S == [int(s) for s in b.S], kind
        assert np.array_equal(a.score, b.score), kind


@pytest.mark.parametrize("case", ["fused", "class-split", "forced-generic"])
def test_more_bins_step_paths(pr, O, monkeypatch, case):
    """K > 5 selectors: the fused step with the term-list tail (csrc/finalize_wide.cu) needs
    every class in one segment; a class of more than 32 752 cells is cut into two and the step
    falls back to the unfused generic path, as it does with PRB_WIDE_FUSED=0 — same bits."""
    from picturedrocks_b200.synth import DiscretisedSpec

    if case == "forced-generic":
        monkeypatch.setenv("PRB_WIDE_FUSED", "0")
    N, C = (70001, 2) if case == "class-split" else (30011, 6)
    spec = DiscretisedSpec(N=N, P=157, C=C, K=9, mean_density=0.12, seed=3)
    eng, yd = spec.device_engine()
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    y = spec.cpu_labels()
    ora = O.SparseInformationSet(spec.cpu_csc(range(spec.P), y), y)
    for kind in ("CIFE", "JMI"):
        a, b = getattr(pr.markers, kind)(inf), getattr(O, kind)(ora)
        assert a._d.fused == (case == "fused")
        a.autoselect(7)
        a.remove(a.S[1])
        a.add(5)
        a.autoselect(2)
        b.autoselect(7)
        b.remove(b.S[1])
        b.add(5)
        b.autoselect(2)
        assert a.S == [int(s) for s in b.S], kind
        assert np.array_equal(a.score, b.score), kind
    c = np.array([-1, 3], dtype=np.int64)
    assert np.array_equal(inf.entropy_wrt(c), ora.entropy_wrt(c))  # the table is clean again


@pytest.mark.parametrize("K", [2, 3, 4])
def test_v_layout_fewer_bins(pr, O, K):
    from picturedrocks_b200.synth import DiscretisedSpec

    spec = DiscretisedSpec(N=9001, P=130, C=5, K=K, mean_density=0.2, seed=K)
    eng, yd = spec.device_engine()
    assert eng.vmode
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    ora = O.SparseInformationSet(inf.X, spec.cpu_labels())
    for cols in ([], [-1], [3], [-1, 3]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(inf.entropy_wrt(c), ora.entropy_wrt(c)), cols
    a, b = pr.markers.CIFE(inf), O.CIFE(ora)
    a.autoselect(5)
    b.autoselect(5)
    assert a.S == [int(s) for s in b.S] and np.array_equal(a.score, b.score)


@pytest.mark.timeout(900)
def test_full_size_config4_sampled_genes(pr, O):
    """BASELINE.json configs[3] at FULL size (1,306,127 x 27,998, 60 classes) on the GPU; the
    CPU oracle recomputes the same quantities for a sample of candidate genes (the synthetic
    matrix is a counter-based hash, so any column can be regenerated on the host):
    entropy_wrt([]) / ([-1]) / ([j]) / ([-1, j]) bit-identical on the sample, plus
    size-independent properties of the full vectors."""
    import torch
    from picturedrocks_b200.synth import DiscretisedSpec

    import gc

    gc.collect()
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~30 GB of free device memory")
    spec = DiscretisedSpec(N=1306127, P=27998, C=60, K=5, mean_density=0.07, seed=4)
    eng, yd = spec.device_engine()
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    y = spec.cpu_labels()
    assert np.array_equal(yd.cpu().numpy(), y)
    j = 24255  # the first gene CIFE selects on this matrix
    sample = [0, 1, 17, 18, 29, 4018, 11955, 12806, 13999, 20000, 27997, j]
    Xs = spec.cpu_csc(sample, y)
    ora = O.SparseInformationSet(Xs, y, nthreads=O.max_threads())
    jj = len(sample) - 1
    got = {c: inf.entropy_wrt(np.array(c, dtype=np.int64)) for c in ((), (-1,), (j,), (-1, j))}
    want = {(): ora.entropy_wrt(np.arange(0)), (-1,): ora.entropy_wrt(np.array([-1])),
            (j,): ora.entropy_wrt(np.array([jj])), (-1, j): ora.entropy_wrt(np.array([-1, jj]))}
    for c in got:
        assert got[c].shape == (spec.P,)
        assert np.array_equal(got[c][sample], want[c]), c
    Hx, Hyx, Hjx, Hyjx = got[()], got[(-1,)], got[(j,)], got[(-1, j)]
    Hy, Hj, Hyj = inf.entropy([-1]), inf.entropy([j]), inf.entropy([-1, j])
    assert Hy == ora.entropy(np.array([-1])) and Hj == ora.entropy(np.array([jj]))
    assert Hyj == ora.entropy(np.array([-1, jj]))
    eps = 1e-9
    # H(A,B) >= max(H(A), H(B)) and <= H(A) + H(B), for every candidate gene
    assert np.all(Hyx >= np.maximum(Hx, Hy) - eps) and np.all(Hyx <= Hx + Hy + eps)
    assert np.all(Hjx >= np.maximum(Hx, Hj) - eps) and np.all(Hjx <= Hx + Hj + eps)
    assert np.all(Hyjx >= np.maximum(Hyx, Hjx) - eps) and np.all(Hyjx <= Hyx + Hj + eps)
    # x_i = x_j: the joint with itself adds nothing; planted duplicate / all-zero genes
    assert Hjx[j] == Hj and Hyjx[j] == Hyj
    assert Hx[4018] == Hx[4017] and Hyjx[4018] == Hyjx[4017]  # planted duplicate
    assert Hx[29] == 0.0 and Hjx[29] == Hj and Hyjx[29] == Hyj  # planted all-zero gene
    # exported matrix (original cell order) equals the host regeneration on the sample
    Xg = eng.to_scipy_csc(30)
    Xh = spec.cpu_csc(range(30), y)
    assert np.array_equal(Xg.indptr, Xh.indptr) and np.array_equal(Xg.indices, Xh.indices)
    assert np.array_equal(Xg.data, Xh.data)
    fs = pr.markers.CIFE(inf)
    fs.autoselect(3)
    assert fs.S[0] == j and len(set(fs.S)) == 3


def test_makeinfoset_sparse_input_is_not_densified(pr, O):
    """Sparse raw counts (CSR as AnnData stores them, CSC, float32 / float64 / int64, explicit
    zeros, genes without zeros): the sparse-aware discretiser gives the codes the reference
    gets by densifying (infoset.py:66-67, oracle on X.toarray()); negative values fall back
    to the dense path."""
    import types

    import pandas as pd
    from picturedrocks_b200.markers.mutualinformation.infoset import discretize_sparse_to_engine
    from picturedrocks_b200.synth import poisson_gamma

    X, y = poisson_gamma(4000, 300, 7, 13)
    X[:, :6] += 2.0  # genes without zeros: edge_0 > 0, small stored values get code 0
    X[:, 6] = 0.0    # all-zero gene
    X[:, 7] = 5.0    # constant gene
    for dt in (np.float32, np.float64, np.int64):
        Xd = X.astype(dt)
        want = O.quantile_discretize(Xd, 5)
        for fmt in (scipy.sparse.csr_matrix, scipy.sparse.csc_matrix):
            eng = discretize_sparse_to_engine(fmt(Xd), 5, max_block_nnz=50000)
            assert eng is not None
            assert np.array_equal(np.asarray(eng.to_scipy_csc().todense()), want), (dt, fmt)
    for k in (2, 3, 8):
        eng = discretize_sparse_to_engine(scipy.sparse.csr_matrix(X), k)
        assert np.array_equal(np.asarray(eng.to_scipy_csc().todense()), O.quantile_discretize(X, k))
    # explicit zeros stored in the sparse matrix behave like implicit ones
    Xs = scipy.sparse.csr_matrix(X)
    Xs.data[::7] = 0.0
    eng = discretize_sparse_to_engine(Xs, 5)
    assert np.array_equal(np.asarray(eng.to_scipy_csc().todense()),
                          O.quantile_discretize(Xs.toarray(), 5))
    # the public entry point: same selection as the oracle on the densified matrix
    adata = types.SimpleNamespace(X=scipy.sparse.csr_matrix(X), obs=pd.DataFrame({"y": y}))
    inf = pr.markers.makeinfoset(adata, True)
    ora = O.makeinfoset(types.SimpleNamespace(X=X, obs=adata.obs), True)
    a, b = pr.markers.CIFE(inf), O.CIFE(ora)
    a.autoselect(6)
    b.autoselect(6)
    assert a.S == [int(s) for s in b.S] and np.array_equal(a.score, b.score)
    # negative values: the coded matrix is not sparse -> dense path, still the same codes
    Xn = X.copy()
    Xn[::3, 10:20] -= 1.5
    assert discretize_sparse_to_engine(scipy.sparse.csr_matrix(Xn), 5) is None
    inf = pr.markers.makeinfoset(types.SimpleNamespace(X=scipy.sparse.csr_matrix(Xn),
                                                       obs=adata.obs), False)
    assert np.array_equal(np.asarray(inf.X.todense()), O.quantile_discretize(Xn, 5))


def test_sparse_discretiser_pbmc_shaped_sample(pr, O):
    """PBMC-68k-shaped raw sparse counts (68,579 cells, 2 % stored, CSR): makeinfoset never
    densifies; the codes of a sample of genes equal the oracle's on the densified columns."""
    import types

    import pandas as pd

    rng = np.random.RandomState(3)
    N, P, n = 68579, 8192, 11_000_000
    rows, cols = rng.randint(0, N, n), (rng.beta(0.6, 3.0, n) * P).astype(np.int64)
    vals = (1 + rng.poisson(1.5, n)).astype(np.float32)
    X = scipy.sparse.csr_matrix((vals, (rows, cols)), shape=(N, P))
    y = rng.randint(0, 10, N).astype(np.int64)
    inf = pr.markers.makeinfoset(types.SimpleNamespace(X=X, obs=pd.DataFrame({"y": y})), True)
    assert inf.N == N and inf.P == P and inf.has_y
    got = inf.X
    sample = [0, 1, 2, 5, 17, 100, 999, 2048, 4095, 6000, 8191]
    want = O.quantile_discretize(np.asarray(X[:, sample].todense()), 5)
    assert np.array_equal(np.asarray(got[:, sample].todense()), want)
    fs = pr.markers.CIFE(inf)
    fs.autoselect(3)
    assert len(set(fs.S)) == 3


def test_ingestion_accepts_what_the_reference_accepts(pr, O, monkeypatch):
    """SparseInformationSet(X, y): canonical CSC of any index / data width goes to the device
    as it is (chunked upload, packed and checked there, csrc/ingest.cu); unsorted indices,
    duplicates and stored zeros take the reference's host canonicalisation (infoset.py:191-196);
    same entropies either way.  Codes above 255 are rejected."""
    from picturedrocks_b200.synth import DiscretisedSpec

    monkeypatch.setenv("PRB_INGEST_CHUNK", "50000")  # many chunks, both staging buffers reused
    spec = DiscretisedSpec(N=20000, P=300, C=12, K=5, mean_density=0.06, seed=7)
    y = spec.cpu_labels()
    X = spec.cpu_csc(range(spec.P), y)  # canonical, int64 data, int32 indices
    ora = O.SparseInformationSet(X, y)
    cols = [np.array(c, dtype=np.int64) for c in ([], [-1], [5], [-1, 5])]
    want = [ora.entropy_wrt(c) for c in cols]

    def check(Xin):
        inf = pr.markers.SparseInformationSet(Xin, y)
        for c, w in zip(cols, want):
            assert np.array_equal(inf.entropy_wrt(c), w)
        return inf

    inf = check(X)
    assert inf.X is X  # canonical CSC is not copied
    X8 = scipy.sparse.csc_matrix((X.data.astype(np.uint8), X.indices.astype(np.int64),
                                  X.indptr.astype(np.int64)), shape=X.shape)
    check(X8)
    check(scipy.sparse.csr_matrix(X))  # cells-major input: converted by scipy
    check(np.asarray(X.todense()))     # dense ndarray
    # unsorted rows inside the genes
    Xu = X.copy()
    for g in range(0, spec.P, 7):
        a, b = Xu.indptr[g], Xu.indptr[g + 1]
        Xu.indices[a:b] = Xu.indices[a:b][::-1].copy()
        Xu.data[a:b] = Xu.data[a:b][::-1].copy()
    Xu.has_sorted_indices = False
    inf = check(Xu)
    assert inf.X.has_canonical_format
    # duplicates (summed by the reference) and stored zeros (dropped)
    coo = X.tocoo()
    half = coo.data // 2
    rows = np.concatenate([coo.row, coo.row, coo.row[:50]])
    colsi = np.concatenate([coo.col, coo.col, coo.col[:50]])
    vals = np.concatenate([coo.data - half, half, np.zeros(50, dtype=coo.data.dtype)])
    Xd = scipy.sparse.csc_matrix(scipy.sparse.coo_matrix((vals, (rows, colsi)), shape=X.shape))
    Xd2 = scipy.sparse.csc_matrix((np.concatenate([X.data, [0]]), np.concatenate([X.indices, [3]]),
                                   np.concatenate([X.indptr[:-1], [X.indptr[-1] + 1]])),
                                  shape=X.shape)  # an explicit zero at the end of the last gene
    check(Xd)
    if Xd2.indices[-2] < 3:
        check(Xd2)
    Xbig = X.copy()
    Xbig.data[0] = 300
    with pytest.raises(ValueError):
        pr.markers.SparseInformationSet(Xbig, y)
    with pytest.raises(OverflowError):
        pr.markers.SparseInformationSet(scipy.sparse.csc_matrix((50, 4), dtype=np.int64))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,density,dtype", [((3001, 517), 0.03, np.float32),
                                                 ((257, 4099), 0.2, np.int64),
                                                 ((70000, 33), 0.01, np.float64),
                                                 ((5, 7), 0.0, np.float32)])
def test_csr_to_csc_counting_transpose(pr, shape, density, dtype):
    """csrc/ingest.cu prb_csr_count / prb_csr_fill: CSR (cells-major) -> by-gene CSC on the
    device equals scipy's tocsc() — offsets, row ids ascending inside every gene, values."""
    import scipy.sparse

    from picturedrocks_b200.markers.mutualinformation.infoset import _sparse_csc_on_device

    rng = np.random.RandomState(shape[0])
    X = scipy.sparse.random(shape[0], shape[1], density=density, format="csr", random_state=rng,
                            dtype=np.float64)
    X.data = (1 + np.floor(X.data * 50)).astype(dtype)
    X.indices = X.indices.astype(np.int64 if shape[0] == 257 else np.int32)
    ref = X.tocsc()
    ref.sort_indices()
    indptr, rows, vals, N, P = _sparse_csc_on_device(X)
    assert (N, P) == shape
    assert np.array_equal(indptr.cpu().numpy(), ref.indptr.astype(np.int64))
    assert np.array_equal(rows.cpu().numpy(), ref.indices.astype(np.int32))
    assert np.array_equal(vals.cpu().numpy(), ref.data.astype(vals.cpu().numpy().dtype))
