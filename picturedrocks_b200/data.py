"""`ExpressionData`: the few members of `anndata.AnnData` this path touches.

The reference takes `anndata.AnnData` objects (`adata.X`, `adata.obs["y"]`, `adata.uns`,
`adata.n_obs`, `adata[rows, :]`, `adata[:, genes]`, `adata.copy()`;
picturedrocks/markers/mutualinformation/infoset.py:41-46, performance.py:271-300,
read.py:100-111).  Every function of this package accepts a real AnnData just as well —
nothing here imports anndata — and this container is what the tests, the benchmark and
hosts without anndata pass instead.  Slicing copies (AnnData returns views; the reference
never writes through one on this path).
"""
import numpy as np
import pandas as pd
import scipy.sparse


class ExpressionData:
    """Cells x genes matrix `X` (ndarray or scipy sparse) with per-cell annotations `obs`
    (pandas DataFrame), per-gene annotations `var` and unstructured `uns` (dict)."""

    def __init__(self, X, obs=None, var=None, uns=None):
        if not scipy.sparse.issparse(X):
            X = np.asarray(X)
        assert X.ndim == 2, "X must be cells x genes"
        self.X = X
        n, p = X.shape
        self.obs = pd.DataFrame(index=pd.RangeIndex(n).astype(str)) if obs is None else obs
        self.var = pd.DataFrame(index=pd.RangeIndex(p).astype(str)) if var is None else var
        assert len(self.obs) == n and len(self.var) == p, "annotation lengths do not match X"
        self.uns = {} if uns is None else uns

    @property
    def n_obs(self):
        return self.X.shape[0]

    @property
    def n_vars(self):
        return self.X.shape[1]

    @property
    def shape(self):
        return self.X.shape

    @staticmethod
    def _index(ix, n):
        if isinstance(ix, slice):
            return np.arange(n)[ix]
        ix = np.asarray(ix)
        if ix.dtype == bool:
            assert ix.shape == (n,), "boolean mask of the wrong length"
            return np.flatnonzero(ix)
        return np.atleast_1d(ix).astype(np.int64)

    def __getitem__(self, key):
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        r, c = self._index(rows, self.n_obs), self._index(cols, self.n_vars)
        X = self.X
        if scipy.sparse.issparse(X):
            X = X.tocsr()[r][:, c]
        else:
            X = X[np.ix_(r, c)]
        return ExpressionData(X, self.obs.iloc[r].copy(), self.var.iloc[c].copy(), dict(self.uns))

    def copy(self):
        return ExpressionData(self.X.copy(), self.obs.copy(), self.var.copy(), dict(self.uns))
