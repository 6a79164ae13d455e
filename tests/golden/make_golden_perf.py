"""Golden vectors for the callers around the hot path, from the
UNMODIFIED reference (build container only):

    python tests/golden/make_golden_perf.py

picturedrocks/read.py and picturedrocks/performance.py are loaded from /root/reference with
stub modules for their absent third-party imports (anndata, scanpy, plotly — none of them is
touched by the functions exercised here) and driven with
`picturedrocks_b200.data.ExpressionData` in place of an AnnData.  Outputs: perf.npz.
"""
import importlib.util
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = os.environ.get("PRB_REFERENCE_ROOT", "/root/reference")

import fixtures  # noqa: E402
from picturedrocks_b200.data import ExpressionData  # noqa: E402


def load_reference():
    for name in ("anndata", "scanpy", "scanpy.preprocessing", "plotly", "plotly.graph_objs"):
        sys.modules.setdefault(name, mock.MagicMock())
    pkg = types.ModuleType("picturedrocks")
    pkg.__path__ = [os.path.join(REF, "picturedrocks")]
    sys.modules["picturedrocks"] = pkg
    mods = {}
    for name in ("read", "performance"):
        spec = importlib.util.spec_from_file_location(
            "picturedrocks." + name, os.path.join(REF, "picturedrocks", name + ".py"))
        mods[name] = importlib.util.module_from_spec(spec)
        sys.modules["picturedrocks." + name] = mods[name]
        spec.loader.exec_module(mods[name])
    return mods["read"], mods["performance"]


def main():
    read, perf = load_reference()
    out = {}
    # fold indices: ordered and shuffled (numpy's global generator, seeded)
    for n, k in fixtures.FOLD_CASES:
        out["folds_%d_%d" % (n, k)] = np.concatenate(
            [np.r_[len(f), f] for f in perf.kfoldindices(n, k)] or [np.zeros(0, int)])
        np.random.seed(n + k)
        out["rfolds_%d_%d" % (n, k)] = np.concatenate(
            [np.r_[len(f), f] for f in perf.kfoldindices(n, k, random=True)] or [np.zeros(0, int)])
    # label coding
    X, names = fixtures.cluster_names()
    ad = ExpressionData(X, obs=pd.DataFrame({"louvain": names}))
    ad = read.process_clusts(ad, name="louvain")
    out["pc_y"] = ad.obs["y"].to_numpy()
    out["pc_num"] = np.int64(ad.uns["num_clusts"])
    for k, v in ad.uns["clusterindices"].items():
        out["pc_idx_%d" % k] = v
    out["pc_categories"] = np.array(list(ad.obs["clust"].cat.categories))
    # confusion matrix / error count
    y, yhat = fixtures.predictions()
    rep = perf.PerformanceReport(y, yhat)
    out["conf"] = rep.getconfusionmatrix()
    out["wrong"] = np.float64(rep.wrong())
    # FoldTester protocol with a host-side select function and classifier
    ft = perf.FoldTester(ad)
    ft.makefolds(k=4)
    ft.selectmarkers(fixtures.variance_selector(3))
    ft.classify(fixtures.MajorityClassifier)
    out["ft_yhat"] = ft.yhat
    for i, m in enumerate(ft.markers):
        out["ft_marker%d" % i] = np.asarray(m)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "f.npz")
        ft.savefoldsandmarkers(path)
        saved = np.load(path)
        out["ft_saved_keys"] = np.array(sorted(saved.files))
    t = perf.truncatemarkers(ft, 2)
    out["ft_trunc0"] = np.asarray(t.markers[0])
    np.savez_compressed(os.path.join(HERE, "perf.npz"), **out)
    print("perf.npz: %d arrays" % len(out))


if __name__ == "__main__":
    main()
