"""oracle/synth_cpu.c (C + OpenMP generator of the benchmark's synthetic matrices) against
the numpy statement of the same recipe in picturedrocks_b200/synth.py, and the flat-array
oracle set against the scipy-backed one."""
import numpy as np
import pytest

from oracle import oracle as O
from picturedrocks_b200.synth import DiscretisedSpec


@pytest.mark.parametrize("K,C,lo,hi", [(5, 13, 3990, 4100), (2, 3, 0, 40), (3, 60, 5000, 5040),
                                       (10, 20, 0, 30), (20, 20, 10, 40)])
def test_c_generator_equals_numpy_recipe(K, C, lo, hi):
    spec = DiscretisedSpec(N=30011, P=6000, C=C, K=K, mean_density=0.06, seed=9 + K)
    indptr, rows, codes, y = O.synth_csc(spec, lo, hi, nthreads=3)
    assert np.array_equal(y, spec.cpu_labels())
    X = spec.cpu_csc(range(lo, hi), y)
    assert np.array_equal(X.indptr, indptr)
    assert np.array_equal(X.indices, rows) and np.array_equal(X.data, codes)
    assert codes.min() >= 1 and codes.max() <= K - 1
    pick = [hi - 1, lo, lo + 7]
    ip2, r2, c2, _ = O.synth_csc(spec, genes=pick, nthreads=2)
    X2 = spec.cpu_csc(pick, y)
    assert np.array_equal(X2.indptr, ip2) and np.array_equal(X2.indices, r2)
    assert np.array_equal(X2.data, c2)


def test_flat_set_equals_scipy_set():
    spec = DiscretisedSpec(N=20000, P=300, C=12, K=5, mean_density=0.06, seed=7)
    indptr, rows, codes, y = O.synth_csc(spec)
    flat = O.FlatSparseInformationSet(indptr, rows, codes, spec.N, y, nthreads=4)
    ref = O.SparseInformationSet(spec.cpu_csc(range(spec.P), y), y)
    for cols in ([], [-1], [5], [-1, 5], [3, 7], [-1, 3, 7]):
        c = np.array(cols, dtype=np.int64)
        assert np.array_equal(flat.entropy_wrt(c), ref.entropy_wrt(c)), cols
        assert flat.entropy(c) == ref.entropy(c), cols
    for kind in ("CIFE", "JMI", "CIFEUnsup", "MIM"):
        a, b = getattr(O, kind)(flat), getattr(O, kind)(ref)
        a.autoselect(6)
        b.autoselect(6)
        assert [int(s) for s in a.S] == [int(s) for s in b.S], kind
        assert np.array_equal(a.score, b.score), kind
