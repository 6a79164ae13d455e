"""Label coding in front of the marker-selection path (reference: picturedrocks/read.py:74-111;
SURVEY §8(f) rank 3).

`process_clusts` turns a cluster column of `adata.obs` into what `makeinfoset(adata, True)`
and the cross-validation helpers read: `obs["clust"]` (categorical), `obs["y"]` (integer
codes 0..C-1 in category order), `uns["num_clusts"]`, `uns["clusterindices"]`.
Works on `anndata.AnnData` and on `picturedrocks_b200.data.ExpressionData` alike.
"""
from warnings import warn

import numpy as np


def process_clusts(adata, name="clust", copy=False):
    """Annotate `adata` with cluster codes and per-cluster cell indices.

    Args
    ----
    adata: AnnData-like
    name: str
        column of `adata.obs` holding the cluster labels
    copy: bool
        annotate and return a copy instead of `adata` itself

    The codes keep pandas' dtype (int8 for fewer than 128 clusters, as in the reference);
    the information sets of this package widen them before any arithmetic, so the
    reference's int8 overflow in `entropy([-1, j])` does not occur.
    """
    adata = adata.copy() if copy else adata
    adata.obs["clust"] = adata.obs[name].astype("category")
    adata.obs["y"] = adata.obs["clust"].cat.codes
    if bool(adata.obs["clust"].isnull().any()):
        warn("Some or all cells not assigned to cluster.")
    n = adata.obs["clust"].cat.categories.size
    adata.uns["num_clusts"] = n
    y = adata.obs["y"].to_numpy()
    adata.uns["clusterindices"] = {k: np.flatnonzero(y == k) for k in range(n)}
    return adata
