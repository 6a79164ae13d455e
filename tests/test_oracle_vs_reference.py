"""Live check of the CPU oracle against the UNMODIFIED reference on random inputs (build
container only: skipped where /root/reference is absent, e.g. on the GPU box, where the
committed golden vectors of tests/golden/ pin the oracle instead).

The reference's hot-path modules are loaded by oracle/ref_shim.py in numba-JIT mode (its
intended execution mode).  Everything is compared bit-for-bit (fp64 np.array_equal)."""
import numpy as np
import pytest
import scipy.sparse

from oracle import oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(),
                                reason="reference tree not present (build container only)")


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def _random_codes(rng, N, P, K, density):
    X = rng.randint(1, K, size=(N, P)) * (rng.rand(N, P) < density * rng.rand(1, P) * 2)
    X[:, rng.randint(P)] = 0                      # an all-zero gene
    X[:, rng.randint(P)] = X[:, rng.randint(P)]   # a duplicated gene: exact score ties
    if X.max() < K - 1:
        X[0, 0] = K - 1
    return X.astype(np.int64)


CASES = [  # N, P, C, K, density, seed
    (300, 40, 2, 2, 0.5, 1), (500, 60, 7, 5, 0.2, 2), (800, 30, 19, 5, 0.05, 3),
    (400, 50, 60, 5, 0.3, 4), (350, 25, 4, 10, 0.4, 5), (600, 20, 9, 20, 0.6, 6),
    (64, 12, 3, 3, 1.0, 7), (1000, 16, 33, 8, 0.1, 8),
]


@pytest.mark.parametrize("N,P,C,K,density,seed", CASES)
def test_entropies_and_selectors_equal_the_reference(ref, N, P, C, K, density, seed):
    INFOSET, ITER = ref
    rng = np.random.RandomState(seed)
    X = _random_codes(rng, N, P, K, density)
    y = rng.randint(0, C, N).astype(np.int64)
    y[:C] = np.arange(C)
    r = INFOSET.SparseInformationSet(scipy.sparse.csc_matrix(X), y)
    o = O.SparseInformationSet(scipy.sparse.csc_matrix(X), y)
    assert (r._shift, r._ybits) == (o._shift, o._ybits)
    j, k = int(rng.randint(P)), int(rng.randint(P))
    for cols in ([], [-1], [j], [-1, j], [j, k], [-1, k, j]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(r.entropy_wrt(c), o.entropy_wrt(c)), cols
        if cols:
            assert r.entropy(c) == o.entropy(c), cols
    n = min(6, P - 2)
    for kind in ("CIFE", "JMI", "CIFEUnsup", "MIM", "UniEntropy"):
        a, b = getattr(ITER, kind)(r), getattr(O, kind)(o)
        assert np.array_equal(a.base_score, b.base_score), kind
        a.autoselect(n)
        b.autoselect(n)
        assert [int(s) for s in a.S] == [int(s) for s in b.S], kind
        assert np.array_equal(a.score, b.score), kind
        if kind in ("CIFE", "JMI", "CIFEUnsup"):
            g = int(a.S[1])
            a.remove(g)
            b.remove(g)
            assert np.array_equal(a.score, b.score) and np.array_equal(a.penalty, b.penalty), kind


@pytest.mark.parametrize("dtype,seed", [(np.float32, 1), (np.float64, 2), (np.int64, 3)])
@pytest.mark.parametrize("k", [2, 5, 8])
def test_discretiser_equals_the_reference(ref, dtype, seed, k):
    INFOSET, _ = ref
    rng = np.random.RandomState(seed)
    N, P = 700, 30
    X = rng.poisson(rng.gamma(0.5, 4.0, P), size=(N, P)).astype(np.float64)
    X[:, :6] *= rng.lognormal(size=(N, 6))        # continuous genes
    X[:, 7] = 0
    X[:, 8] = 4
    X[:, 9] = np.arange(N) % 10
    X[:, 10] = -X[:, 10]                          # negatives
    X = np.floor(X).astype(dtype) if dtype == np.int64 else X.astype(dtype)
    assert np.array_equal(INFOSET.quantile_discretize(X, k), O.quantile_discretize(X, k))
    Xs = scipy.sparse.csr_matrix(np.maximum(X, 0))
    assert np.array_equal(INFOSET.quantile_discretize(Xs, k), O.quantile_discretize(Xs, k))


def test_dense_information_set_equals_the_reference(ref):
    INFOSET, _ = ref
    rng = np.random.RandomState(11)
    X = _random_codes(rng, 400, 18, 5, 0.4)
    y = rng.randint(0, 6, 400).astype(np.int64)
    Xy = np.hstack([X, y[:, None]])
    r, o = INFOSET.InformationSet(Xy, True), O.InformationSet(Xy, True)
    for cols in ([], [-1], [3], [-1, 3], [2, 5]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(r.entropy_wrt(c), o.entropy_wrt(c)), cols
        if cols:
            assert r.entropy(c) == o.entropy(c), cols
