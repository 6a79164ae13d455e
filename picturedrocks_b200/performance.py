"""K-fold evaluation of marker selection (reference: picturedrocks/performance.py:28-300,
361-390) — the one other in-repo caller of the selection path.

`FoldTester.selectmarkers(select_function)` runs a marker-selection function on the
complement of every fold; with this package the function is typically
`mi_selector("CIFE", n)`: discretisation and greedy selection on the GPU per fold.

`FoldTester`, `kfoldindices`, `PerformanceReport`, `truncatemarkers`, `merge_markers` and
`NearestCentroidClassifier` keep the reference's names, arguments, attributes and file
format (`np.savez` with keys k, y, fold<i>, marker<i>).  Host-side glue in numpy: none of
it is on the hot path.  The plotly confusion-matrix figure of the reference is not built.
"""
import numpy as np
import scipy.sparse

from .read import process_clusts


def kfoldindices(n, k, random=False):
    """Yield the index arrays of k folds of range(n): the first n % k folds hold
    n // k + 1 indices, the others n // k; `random` shuffles with `np.random.shuffle`
    (seed numpy's global generator for reproducible folds, as with the reference)."""
    order = np.arange(n)
    if random:
        np.random.shuffle(order)
    small, extra = divmod(n, k)
    start = 0
    for i in range(k):
        if start >= n:
            break
        size = small + 1 if i < extra else small
        yield order[start:start + size]
        start += size


class PerformanceReport:
    """Actual vs predicted cluster labels (performance.py:59-173).

    Args
    ----
    y: numpy.ndarray of actual labels 0..K-1 (every label present)
    yhat: numpy.ndarray of predicted labels
    """

    def __init__(self, y, yhat):
        self.y = y
        self.yhat = yhat
        self.N = y.shape[0]
        self.K = y.max() + 1
        assert np.array_equal(np.unique(self.y), np.arange(self.K)), \
            "Cluster labels should be 0, 1, 2, ..., K -1"
        flat = np.asarray(self.y).ravel()
        self.clusterindices = {k: np.flatnonzero(flat == k) for k in range(self.K)}
        self.nk = np.array([len(self.clusterindices[k]) for k in range(self.K)])

    def wrong(self):
        """Number of misclassified cells (a float, as in the reference)."""
        return np.sum((np.asarray(self.y).ravel() != self.yhat) * 1.0)

    def printscore(self):
        wrong = self.wrong()
        print("{} out of {} incorrect: {:.2f}%".format(wrong, self.N, 100 * wrong / self.N))

    def getconfusionmatrix(self):
        """(K, K) array; entry [i, j] = fraction of the cells of cluster i predicted as j."""
        K = self.K
        table = np.zeros([K, K])
        for i in range(K):
            pred = np.asarray(self.yhat)[self.clusterindices[i]]
            table[i] = np.bincount(pred, minlength=K)[:K] / self.nk[i]
        return table


class FoldTester:
    """K-fold cross validation for marker selection (performance.py:173-335).

    Args
    ----
    adata: AnnData-like (`picturedrocks_b200.data.ExpressionData` or `anndata.AnnData`)
        with `obs["y"]` (see `picturedrocks_b200.read.process_clusts`)
    """

    def __init__(self, adata):
        self.adata = adata
        self.folds = None
        self.yhat = None
        self.markers = None

    def makefolds(self, k=5, random=False):
        self.folds = list(kfoldindices(self.adata.n_obs, k, random))

    def _y(self):
        return np.asarray(self.adata.obs["y"])

    def savefolds(self, file):
        d = {"k": len(self.folds), "y": self._y()}
        for i, f in enumerate(self.folds):
            d["fold{}".format(i)] = f
        return np.savez(file, **d)

    def _load(self, file, with_markers):
        d = np.load(file)
        k = int(d["k"])
        self.folds = [d["fold{}".format(i)] for i in range(k)]
        if with_markers:
            self.markers = [d["marker{}".format(i)] for i in range(k)]
        assert np.array_equal(self._y(), d["y"].ravel()), "y vector does not match."
        assert self.validatefolds(), "folds are not partition of indices"

    def loadfolds(self, file):
        """Folds from a file written by `savefolds` or `savefoldsandmarkers`."""
        self._load(file, False)

    def validatefolds(self):
        """True if every observation lies in exactly one fold."""
        counts = np.zeros(self.adata.n_obs)
        for f in self.folds:
            np.add.at(counts, f, 1)
        return bool(np.all(counts == 1))

    def _train_mask(self, fold):
        mask = np.ones(self.adata.n_obs, dtype=bool)
        mask[fold] = False
        return mask

    def selectmarkers(self, select_function, comm=None):
        """`select_function(traindata) -> list of gene indices` on the complement of every
        fold; the results are kept in `self.markers` (performance.py:252-262).
        The folds are independent selections on row subsets: with `comm` (a
        `picturedrocks_b200.dist.Comm`, one process per GPU) rank r selects for the folds
        r, r + world, ... on its own GPU and every rank receives all the marker lists."""
        if comm is None or comm.world == 1:
            self.markers = [select_function(self.adata[self._train_mask(f), :])
                            for f in self.folds]
            return
        import torch.distributed as dist

        mine = {i: list(select_function(self.adata[self._train_mask(f), :]))
                for i, f in enumerate(self.folds) if i % comm.world == comm.rank}
        parts = [None] * comm.world
        dist.all_gather_object(parts, mine, group=comm.group)
        merged = {}
        for d in parts:
            merged.update(d)
        self.markers = [merged[i] for i in range(len(self.folds))]

    def classify(self, classifier):
        """For every fold: train `classifier()` on the other folds restricted to the fold's
        markers, predict the fold's cells; predictions end up in `self.yhat` (-1 = none)."""
        self.yhat = np.zeros(self.adata.n_obs, dtype=int) - 1
        for i, f in enumerate(self.folds):
            genes = np.asarray(self.markers[i])
            c = classifier()
            c.train(self.adata[self._train_mask(f), :][:, genes])
            self.yhat[f] = c.test(self.adata.X[f, :][:, genes])

    def savefoldsandmarkers(self, file):
        d = {"k": len(self.folds), "y": self._y()}
        for i, f in enumerate(self.folds):
            d["fold{}".format(i)] = f
        for i, m in enumerate(self.markers):
            d["marker{}".format(i)] = m
        return np.savez(file, **d)

    def loadfoldsandmarkers(self, file):
        self._load(file, True)


def truncatemarkers(ft, n_markers):
    """A FoldTester on the same data and folds with the first n_markers of every fold."""
    out = FoldTester(ft.adata)
    out.folds = ft.folds
    out.markers = [m[:n_markers] for m in ft.markers]
    return out


def merge_markers(ft, n_markers):
    """Per fold, the union of the first n_markers of several marker lists
    (`ft.markers[i]` = a list of equally long lists)."""
    out = FoldTester(ft.adata)
    out.folds = ft.folds
    out.markers = [list(set(np.array(m)[:, :n_markers].flatten())) for m in ft.markers]
    return out


def _toarray(a):
    return a.toarray() if scipy.sparse.issparse(a) else np.asarray(a)


def _lognorm(X, target=1000.0):
    """Counts per cell scaled to `target` (cells without counts are left at zero), then
    log(1 + x): what scanpy's `normalize_per_cell(adata, 1000, min_counts=0)` followed by
    `log1p` compute (performance.py:375-376, 385-386; scanpy is an un-pinned dependency,
    so parity with the reference's classifier is unpinned)."""
    X = _toarray(X).astype(np.float64)
    counts = X.sum(axis=1)
    counts = counts + (counts == 0)
    return np.log1p(X * (target / counts)[:, None])


class NearestCentroidClassifier:
    """Cluster centroids of the log-normalised training cells; a test cell gets the label
    of the nearest centroid (Euclidean; lowest label on ties) (performance.py:361-390)."""

    def __init__(self):
        self.xkibar = None

    def train(self, adata):
        adata = process_clusts(adata.copy())
        X = _lognorm(adata.X)
        idx = adata.uns["clusterindices"]
        self.xkibar = np.array([X[idx[k]].mean(axis=0) for k in range(adata.uns["num_clusts"])])

    def test(self, Xtest):
        X = _lognorm(Xtest)
        c = self.xkibar
        d2 = (X * X).sum(1)[:, None] - 2.0 * (X @ c.T) + (c * c).sum(1)[None, :]
        return d2.argmin(axis=1)


def mi_selector(kind, n_markers, k=5):
    """A `select_function` for `FoldTester.selectmarkers`: discretise the training cells on
    the GPU (`makeinfoset`), run the greedy selector `kind` ("CIFE", "JMI", "MIM", ...) for
    n_markers steps and return the selected gene indices."""
    from . import markers

    def select(traindata):
        fs = getattr(markers, kind)(markers.makeinfoset(traindata, True, k))
        fs.autoselect(n_markers)
        return [int(g) for g in fs.S]

    return select
