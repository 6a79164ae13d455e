"""Pins oracle/ (the CPU restatement) to golden vectors produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import scipy.sparse

import fixtures
from helpers import check_entropies, check_selectors, sha
from oracle import oracle as O
from picturedrocks_b200.synth import DiscretisedSpec, poisson_gamma


def make(kind, infoset):
    return getattr(O, kind)(infoset)


def test_sampleX_entropies(golden):
    g = golden("sampleX")
    X = g["X"]
    assert np.array_equal(X, fixtures.sampleX()[0])
    check_entropies(O.SparseInformationSet(X), g, "sparse_")
    check_entropies(O.InformationSet(X, False), g, "dense_")
    # analytic values of the reference's own test (tests/test_entropy.py:86-134)
    _, ent, mi = fixtures.sampleX()
    s = O.SparseInformationSet(X)
    assert np.allclose(s.entropy_wrt(np.arange(0)), ent)
    joint = np.array([s.entropy_wrt(np.array([i])) for i in range(6)])
    assert np.allclose(joint, ent[:, None] + ent[None, :] - mi)


@pytest.mark.parametrize("name", ["xor", "diminish_red"])
def test_selector_known_answers(golden, name):
    g = golden(name)
    inf = O.SparseInformationSet(g["X"], g["y"])
    check_entropies(inf, g)
    check_selectors(make, inf, g, 2)
    for kind in ("CIFE", "JMI"):
        fs = make(kind, inf)
        fs.autoselect(1)
        assert int(fs.S[0]) == int(g[kind + "_traj_first"])
        fs.remove(int(fs.S[0]))
        fs.add(0)
        assert np.array_equal(fs.score, g[kind + "_traj_score"])
        fs.autoselect(1)
        assert [int(s) for s in fs.S] == list(g[kind + "_traj_S"])


def test_reference_expected_values():
    """The literal expectations of the reference's tests/test_fs.py:39-91."""
    X, y = fixtures.diminish_red_table()
    inf = O.SparseInformationSet(X, y)
    cife, jmi, mim = O.CIFE(inf), O.JMI(inf), O.MIM(inf)
    cife.autoselect(2)
    jmi.autoselect(2)
    assert set(cife.S) == {0, 1} and set(jmi.S) == {0, 1}
    ninf = float("-inf")
    assert np.allclose(cife.score, [ninf, ninf, 1, 0])
    assert np.allclose(jmi.score, [ninf, ninf, 1, 0.5])
    assert np.allclose(mim.score, [2, 2, 1, 1])
    X, y = fixtures.xor_table()
    inf = O.SparseInformationSet(X, y)
    third = 1 + ((1 / 3) * np.log2(1 / 3) + (2 / 3) * np.log2(2 / 3))
    for mk in (O.CIFE, O.JMI):
        fs = mk(inf)
        assert np.allclose(fs.score, [0, 0, third])
        fs.autoselect(1)
        assert fs.S == [2]
        fs.remove(2)
        fs.add(0)
        assert np.allclose(fs.score, [ninf, 1, third])
        fs.autoselect(1)
        assert fs.S == [0, 1]


def test_linearxy(golden):
    g = golden("linearxy")
    X, y, _ = fixtures.linearxy()
    assert sha(X) == str(g["X_sha"]), "numpy RandomState stream changed"
    Xq = O.quantile_discretize(X)
    assert np.array_equal(Xq.astype(np.uint8), g["codes"])
    inf = O.SparseInformationSet(Xq, y)
    check_entropies(inf, g, "sparse_")
    dense = O.InformationSet(np.concatenate((Xq, y[:, None]), axis=1), True)
    check_entropies(dense, g, "dense_")
    assert np.array_equal([inf.entropy(np.array([i])) for i in range(20)], g["ent_single"])
    assert np.array_equal([inf.entropy(np.array([-1, i])) for i in range(20)], g["ent_y_single"])


def test_planted600(golden):
    g = golden("planted600")
    X, y = g["X"], g["y"]
    assert np.array_equal(X, fixtures.planted_counts(600, 250, 7, 3)[0])
    Xq = O.quantile_discretize(X)
    assert np.array_equal(Xq.astype(np.uint8), g["codes"])
    inf = O.SparseInformationSet(Xq, y)
    check_entropies(inf, g)
    check_selectors(make, inf, g, 12)
    for kind in ("CIFE", "JMI", "CIFEUnsup"):
        fs = make(kind, inf)
        fs.autoselect(4)
        fs.remove(int(fs.S[1]))
        fs.add(17)
        fs.autoselect(3)
        assert [int(s) for s in fs.S] == list(g[kind + "_mix_S"])
        assert np.array_equal(fs.score, g[kind + "_mix_score"])
        assert np.array_equal(fs.penalty, g[kind + "_mix_penalty"])
    # planted duplicate genes tie exactly; the all-zero gene scores exactly 0
    mim = O.MIM(inf)
    assert mim.base_score[7] == mim.base_score[3]
    assert mim.base_score[9] == 0.0


def test_discretize(golden):
    g = golden("discretize")
    Xc = fixtures.continuous_matrix(5)
    assert sha(Xc) == str(g["X_sha"])
    for dt in (np.float32, np.float64):
        for k in (2, 3, 5, 10, 20):
            got = O.quantile_discretize(Xc.astype(dt), k).astype(np.uint8)
            assert np.array_equal(got, g["codes_%s_k%d" % (np.dtype(dt).name, k)]), (dt, k)
    Xi = g["int_X"]
    for k in (3, 5, 8):
        assert np.array_equal(O.quantile_discretize(Xi, k).astype(np.uint8), g["codes_int64_k%d" % k])
    # sparse input is densified first (infoset.py:66-67)
    sp = scipy.sparse.csr_matrix(Xc.astype(np.float32))
    assert np.array_equal(O.quantile_discretize(sp, 5).astype(np.uint8), g["codes_float32_k5"])


def test_config1(golden):
    """BASELINE.json configs[0]: 2730 x 3451, 19 clusters, makeinfoset then CIFE top-20."""
    g = golden("config1")
    X, y = poisson_gamma(2730, 3451, 19, 1)
    assert sha(X) == str(g["X_sha"])
    Xq = O.quantile_discretize(X, nthreads=O.max_threads())
    assert sha(Xq.astype(np.uint8)) == str(g["codes_sha"])
    inf = O.SparseInformationSet(Xq, y, nthreads=O.max_threads())
    check_entropies(inf, g)
    check_selectors(make, inf, g, 20, kinds=("CIFE", "JMI", "MIM"))


@pytest.mark.parametrize("name,spec,n,kinds", [
    ("synth_sparse", DiscretisedSpec(N=3000, P=400, C=10, K=5, mean_density=0.05, seed=3), 15,
     ("CIFE", "JMI", "CIFEUnsup")),
    ("synth_sparse_k20", DiscretisedSpec(N=2000, P=300, C=12, K=20, mean_density=0.08, seed=5), 8,
     ("CIFE", "JMI")),
])
def test_synth_sparse(golden, name, spec, n, kinds):
    g = golden(name)
    y = spec.cpu_labels()
    Xs = spec.cpu_csc(range(spec.P), y)
    assert np.array_equal(Xs.indptr, g["indptr"])
    assert sha(Xs.data.astype(np.uint8)) == str(g["data_sha"])
    inf = O.SparseInformationSet(Xs, y, nthreads=O.max_threads())
    check_entropies(inf, g)
    check_selectors(make, inf, g, n, kinds=kinds)


def test_term_table_matches_libm():
    import math

    for N in (8, 2730, 68579):
        t = O.term_table(N)
        ref = [0.0] + [-(c / N) * math.log2(c / N) for c in range(1, N + 1)]
        assert np.array_equal(t, np.array(ref))
