"""Callers around the hot path: label coding, k-fold evaluation.
Host-side glue, checked against golden vectors produced by the unmodified reference
(tests/golden/make_golden_perf.py -> perf.npz); the GPU test runs the folds through the
device-resident selectors and compares with the CPU oracle on the same training cells."""
import os

import numpy as np
import pandas as pd
import pytest
import scipy.sparse

import fixtures
from picturedrocks_b200.data import ExpressionData
from picturedrocks_b200.performance import (FoldTester, NearestCentroidClassifier,
                                            PerformanceReport, kfoldindices, merge_markers,
                                            mi_selector, truncatemarkers)
from picturedrocks_b200.read import process_clusts


def _unpack(flat):
    out, i = [], 0
    while i < len(flat):
        n = int(flat[i])
        out.append(flat[i + 1:i + 1 + n])
        i += n + 1
    return out


@pytest.mark.parametrize("n,k", fixtures.FOLD_CASES)
def test_kfoldindices_match_reference(golden, n, k):
    g = golden("perf")
    ours = list(kfoldindices(n, k))
    ref = _unpack(g["folds_%d_%d" % (n, k)])
    assert len(ours) == len(ref) and all(np.array_equal(a, b) for a, b in zip(ours, ref))
    np.random.seed(n + k)
    ours = list(kfoldindices(n, k, random=True))
    ref = _unpack(g["rfolds_%d_%d" % (n, k)])
    assert len(ours) == len(ref) and all(np.array_equal(a, b) for a, b in zip(ours, ref))


def _labelled():
    X, names = fixtures.cluster_names()
    return process_clusts(ExpressionData(X, obs=pd.DataFrame({"louvain": names})), name="louvain")


def test_process_clusts_matches_reference(golden):
    g = golden("perf")
    ad = _labelled()
    assert np.array_equal(ad.obs["y"].to_numpy(), g["pc_y"])
    assert ad.obs["y"].to_numpy().dtype == g["pc_y"].dtype
    assert ad.uns["num_clusts"] == int(g["pc_num"])
    assert list(ad.obs["clust"].cat.categories) == list(g["pc_categories"])
    for k in range(ad.uns["num_clusts"]):
        assert np.array_equal(ad.uns["clusterindices"][k], g["pc_idx_%d" % k])
    # copy=True leaves the input alone
    X, names = fixtures.cluster_names()
    src = ExpressionData(X, obs=pd.DataFrame({"clust": names}))
    out = process_clusts(src, copy=True)
    assert "y" in out.obs and "y" not in src.obs and "num_clusts" not in src.uns


def test_process_clusts_warns_on_missing_labels():
    X, names = fixtures.cluster_names()
    names = names.astype(object)
    names[3] = None
    with pytest.warns(UserWarning):
        ad = process_clusts(ExpressionData(X, obs=pd.DataFrame({"clust": names})))
    assert ad.obs["y"].to_numpy()[3] == -1


def test_performance_report_matches_reference(golden):
    g = golden("perf")
    y, yhat = fixtures.predictions()
    rep = PerformanceReport(y, yhat)
    assert np.array_equal(rep.getconfusionmatrix(), g["conf"])
    assert rep.wrong() == float(g["wrong"])
    assert np.allclose(rep.getconfusionmatrix().sum(axis=1), 1.0)
    with pytest.raises(AssertionError):
        PerformanceReport(np.array([0, 2, 2]), np.array([0, 2, 2]))


def test_foldtester_protocol_matches_reference(golden, tmp_path):
    g = golden("perf")
    ft = FoldTester(_labelled())
    ft.makefolds(k=4)
    assert ft.validatefolds()
    ft.selectmarkers(fixtures.variance_selector(3))
    ft.classify(fixtures.MajorityClassifier)
    assert np.array_equal(ft.yhat, g["ft_yhat"])
    for i, m in enumerate(ft.markers):
        assert np.array_equal(np.asarray(m), g["ft_marker%d" % i])
    path = os.path.join(tmp_path, "folds.npz")
    ft.savefoldsandmarkers(path)
    assert sorted(np.load(path).files) == list(g["ft_saved_keys"])
    ft2 = FoldTester(_labelled())
    ft2.loadfoldsandmarkers(path)
    assert all(np.array_equal(a, b) for a, b in zip(ft.folds, ft2.folds))
    assert all(np.array_equal(a, b) for a, b in zip(ft.markers, ft2.markers))
    ft3 = FoldTester(_labelled())
    ft3.loadfolds(path)
    assert ft3.markers is None and len(ft3.folds) == 4
    assert np.array_equal(np.asarray(truncatemarkers(ft, 2).markers[0]), g["ft_trunc0"])
    # a file for other labels is refused; overlapping folds are detected
    other = _labelled()
    other.obs["y"] = (other.obs["y"].to_numpy() + 1) % 3
    with pytest.raises(AssertionError):
        FoldTester(other).loadfolds(path)
    ft.folds[0] = ft.folds[1]
    assert not ft.validatefolds()


def test_merge_markers():
    ft = FoldTester(_labelled())
    ft.makefolds(k=2)
    ft.markers = [[[1, 2, 3], [2, 5, 6]], [[0, 1, 2], [0, 7, 8]]]
    m = merge_markers(ft, 2)
    assert sorted(m.markers[0]) == [1, 2, 5] and sorted(m.markers[1]) == [0, 1, 7]


def test_expression_data_slicing():
    X, names = fixtures.cluster_names()
    for M in (X, scipy.sparse.csr_matrix(X)):
        ad = ExpressionData(M, obs=pd.DataFrame({"clust": names}))
        mask = np.zeros(ad.n_obs, dtype=bool)
        mask[::3] = True
        sub = ad[~mask, :][:, [4, 1]]
        assert sub.shape == ((~mask).sum(), 2) and sub.n_vars == 2
        dense = sub.X.toarray() if scipy.sparse.issparse(sub.X) else sub.X
        assert np.array_equal(dense, X[~mask][:, [4, 1]])
        assert list(sub.obs["clust"]) == list(names[~mask])
        sub.obs["clust"] = "x"          # a slice is a copy
        assert list(ad.obs["clust"]) == list(names)


def test_nearest_centroid_classifier_separates_planted_clusters():
    rng = np.random.RandomState(3)
    y = rng.randint(0, 3, 300)
    lam = np.array([[20, 1, 1, 5], [1, 20, 1, 5], [1, 1, 20, 5]], dtype=float)
    X = rng.poisson(lam[y]).astype(np.float32)
    ad = process_clusts(ExpressionData(X, obs=pd.DataFrame({"clust": y})))
    ft = FoldTester(ad)
    ft.makefolds(k=3, random=False)
    ft.markers = [[0, 1, 2]] * 3
    ft.classify(NearestCentroidClassifier)
    assert PerformanceReport(ad.obs["y"].to_numpy(), ft.yhat).wrong() <= 6
    c = NearestCentroidClassifier()
    c.train(ad)
    assert c.xkibar.shape == (3, 4)
    # an empty cell (no counts) is classified without a division by zero
    assert c.test(np.zeros((1, 4), dtype=np.float32)).shape == (1,)


@pytest.mark.gpu
def test_folds_through_the_gpu_selectors_match_the_oracle():
    """FoldTester.selectmarkers with the device-resident CIFE / MIM on raw counts: per fold the
    same genes as the CPU oracle run on the same training cells (discretiser included)."""
    from oracle import oracle as O
    from picturedrocks_b200.synth import poisson_gamma

    X, y = poisson_gamma(900, 160, 6, seed=13)
    ad = process_clusts(ExpressionData(X, obs=pd.DataFrame({"clust": y})))
    for kind, n in (("CIFE", 6), ("MIM", 8)):
        ft = FoldTester(ad)
        ft.makefolds(k=3)
        ft.selectmarkers(mi_selector(kind, n))
        for f, got in zip(ft.folds, ft.markers):
            mask = np.ones(ad.n_obs, dtype=bool)
            mask[f] = False
            codes = O.quantile_discretize(X[mask], 5)
            ref = getattr(O, kind)(O.SparseInformationSet(scipy.sparse.csc_matrix(codes),
                                                         ad.obs["y"].to_numpy()[mask].astype(np.int64)))
            ref.autoselect(n)
            assert [int(g) for g in ref.S] == list(got), kind
