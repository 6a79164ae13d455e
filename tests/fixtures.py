"""Input fixtures shared by the golden generator and the tests.

The reference's own known-answer fixtures are restated here from their generating
rules (tests/test_entropy.py:15-80, tests/data/generate_data.py:15-49 of the
reference) rather than copied as files.
"""
import itertools

import numpy as np
from scipy import linalg


def sampleX():
    """8x6 matrix with analytic entropies [0,0,1,1,2,2] (reference tests/test_entropy.py:15-39)."""
    rows = []
    for a in range(2):
        for b in range(2):
            for c in range(2):
                # col2 = a, col3 = c, col4 = 2a+b, col5 = 2b+c
                rows.append([0, 1, a, c, 2 * a + b, 2 * b + c])
    X = np.array(rows)
    entropies = np.array([0, 0, 1, 1, 2, 2])
    mi = np.diag(entropies)
    mi[2, 4] = mi[3, 5] = mi[4, 5] = 1
    mi = np.maximum(mi, mi.T)
    return X, entropies, mi


def xor_table():
    """12x3 + y (reference tests/data/generate_data.py:31-42, columns x1,x2,x3,y)."""
    rows = []
    for b0, b1, b2 in itertools.product([0, 1], [0, 1], [0, 1, 2]):
        yv = b0 ^ b1
        rows.append([b0, b1, yv ^ int(b2 == 2), yv])
    a = np.array(rows)
    return a[:, :3], a[:, 3]


def diminish_red_table():
    """32x4 + y (reference tests/data/generate_data.py:15-28).  The CSV stores bit
    strings that pandas parses as decimal ints (e.g. "10" -> 10), so x1, x2 take
    values {0, 1, 10, 11}; y is re-labelled with np.unique(return_inverse)."""
    rows = []
    for bits in itertools.product([0, 1], repeat=5):
        s = [str(b) for b in bits]
        rows.append([int("".join(s[:2])), int("".join(s[2:4])), int(s[4]), int(s[0]),
                     int("".join(s))])
    a = np.array(rows)
    _, y = np.unique(a[:, 4], return_inverse=True)
    return a[:, :4], y


def linearxy():
    """Seeded Gaussian data with 20 correlated feature groups
    (reference tests/test_entropy.py:42-80)."""
    random = np.random.RandomState(0)
    n, grp, n_grp = 10000, 10, 20
    n_feats = grp * n_grp
    true_rel = random.choice(grp, n_grp)
    true_abs = grp * np.arange(n_grp) + true_rel
    blocks = []
    for _ in range(n_grp):
        q, _r = np.linalg.qr(random.randn(grp, grp))
        blocks.append(q @ np.diag(5.0 ** (-np.arange(grp))))
    covar = linalg.block_diag(*blocks)
    w = np.zeros((n_feats, 1))
    w[true_abs, 0] = random.randn(n_grp)
    X = random.randn(n, n_feats) @ covar
    y = ((np.sign(X @ w) + 1) // 2).astype(int).flatten()
    return X, y, true_abs


def planted_counts(N, P, C, seed, dtype=np.float32):
    """Poisson-gamma counts with a duplicated gene, an all-zero gene and a constant gene."""
    from picturedrocks_b200.synth import poisson_gamma

    X, y = poisson_gamma(N, P, C, seed, dtype)
    if P >= 12:
        X[:, 7] = X[:, 3]      # exact duplicate -> exact score ties
        X[:, 9] = 0            # all-zero gene
        X[:, 11] = 3           # constant gene
    return X, y


def continuous_matrix(seed, N=2000, P=24):
    rng = np.random.RandomState(seed)
    X = rng.lognormal(size=(N, P))
    X[rng.rand(N, P) < 0.3] = 0
    X[:, 3] = 0
    X[:, 4] = 2.5
    X[:, 5] = rng.rand(N) < 0.3
    X[:, 6] = -X[:, 6]
    X[:, 7] = np.arange(N) % 10
    return X


# ---- callers around the hot path (tests/test_performance.py, golden/make_golden_perf.py) ----
FOLD_CASES = [(10, 3), (11, 5), (100, 7), (3, 5), (8, 1)]


def cluster_names(n=57, p=9, seed=5):
    """Small count matrix + string cluster names (three named clusters, unsorted)."""
    rng = np.random.RandomState(seed)
    names = rng.choice(["T cell", "B cell", "NK"], size=n, p=[0.5, 0.3, 0.2])
    lam = 1.0 + 4.0 * rng.rand(3, p)
    code = {"T cell": 0, "B cell": 1, "NK": 2}
    X = rng.poisson(lam[[code[s] for s in names]]).astype(np.float32)
    return X, names


def predictions(n=200, K=5, seed=9):
    rng = np.random.RandomState(seed)
    y = rng.randint(0, K, n)
    y[:K] = np.arange(K)
    yhat = np.where(rng.rand(n) < 0.7, y, rng.randint(0, K, n))
    return y, yhat


def variance_selector(n):
    """select_function: the n genes of largest variance over the training cells."""
    def select(traindata):
        X = np.asarray(traindata.X, dtype=np.float64)
        return [int(g) for g in np.argsort(-X.var(axis=0), kind="stable")[:n]]
    return select


class MajorityClassifier:
    """classifier protocol of FoldTester.classify: predicts the most frequent training label."""

    def train(self, adata):
        self.label = int(np.bincount(np.asarray(adata.obs["y"])).argmax())

    def test(self, Xtest):
        return np.full(Xtest.shape[0], self.label)
