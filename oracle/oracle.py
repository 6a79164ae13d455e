"""CPU oracle for the PicturedRocks mutual-information hot path.

TEST INFRASTRUCTURE, NOT PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module.  ``picturedrocks_b200`` never does.

It is a restatement (numpy + the plain-C kernels of ``oracle.c``) of the
reference's algorithm; citations are ``file:line`` relative to the reference
tree, with INFOSET = picturedrocks/markers/mutualinformation/infoset.py,
ITER = picturedrocks/markers/mutualinformation/iterative.py and
BASE = picturedrocks/markers/_iterative.py.

Parity status: PINNED against the unmodified reference — see
``tests/golden/make_golden.py`` (generator, run in the build container where
/root/reference exists) and ``tests/test_oracle_golden.py``.
"""
import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force=False):
    """Compile oracle.c → liboracle.so (gcc).  Building the checker is not using it."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "synth_cpu.c")]
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(map(os.path.getmtime, srcs)):
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_sparse_entropy.restype = ctypes.c_double
        _LIB.orc_dense_entropy.restype = ctypes.c_double
    return _LIB


def max_threads():
    return int(lib().orc_max_threads())


def host_cores():
    """Cores this process may run on (ignores OMP_NUM_THREADS, which torchrun sets to 1)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def _p(a, t):
    return a.ctypes.data_as(t)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


# ----------------------------------------------------------------------------
# flat-array kernels (the reference's njit functions)
# ----------------------------------------------------------------------------
def sparse_entropy_wrt(dindices, dindptr, ddata, mindices, mdata, n_rows, n_feats, n_mcols,
                       shift, ybits, y, classsizes, nthreads=1, feat_range=None):
    """INFOSET:306-410 ``_sparse_entropy_wrt``; same argument order and meaning.

    ``nthreads`` > 1 shards the (independent) per-gene loop over OpenMP threads;
    ``feat_range=(b, e)`` restricts the computation to genes b..e-1.
    """
    L = lib()
    b, e = (0, int(n_feats)) if feat_range is None else map(int, feat_range)
    out = np.empty(e - b, dtype=np.float64)
    dindptr = _c(dindptr, np.int64)
    mindices = _c(mindices, np.int64)
    mdata = _c(mdata, np.int64)
    y = _c(y, np.int64)
    classsizes = _c(classsizes, np.float64)
    if ddata.dtype == np.uint8 and dindices.dtype == np.int32:
        fn, it, dt, ip, dp = L.orc_sparse_entropy_wrt_i32_u8, np.int32, np.uint8, _i32p, _u8p
    elif dindices.dtype == np.int32:
        fn, it, dt, ip, dp = L.orc_sparse_entropy_wrt_i32_i64, np.int32, np.int64, _i32p, _i64p
    else:
        fn, it, dt, ip, dp = L.orc_sparse_entropy_wrt_i64_i64, np.int64, np.int64, _i64p, _i64p
    dindices = _c(dindices, it)
    ddata = _c(ddata, dt)
    rc = fn(_p(dindices, ip), _p(dindptr, _i64p), _p(ddata, dp), _p(mindices, _i64p),
            _p(mdata, _i64p), ctypes.c_int64(mindices.size), ctypes.c_int64(n_rows),
            ctypes.c_int64(b), ctypes.c_int64(e), ctypes.c_int64(n_mcols),
            ctypes.c_int64(shift), ctypes.c_int64(ybits), _p(y, _i64p), _p(classsizes, _f64p),
            ctypes.c_int64(classsizes.size), _p(out, _f64p), ctypes.c_int(nthreads))
    assert rc == 0
    return out


def sparse_entropy(indices, data, n_rows, max_val):
    """INFOSET:413-428 ``_sparse_entropy``."""
    data = _c(data, np.int64)
    return float(lib().orc_sparse_entropy(_p(data, _i64p), ctypes.c_int64(data.size),
                                          ctypes.c_int64(n_rows), ctypes.c_int64(max_val)))


def make_master_col(X, cols, shift):
    """INFOSET:431-442 ``_sparse_make_master_col``: joint-state packing of the
    selected columns, first column most significant; returns (row ids ascending,
    packed values) of the non-zero entries."""
    N = X.shape[0]
    packed = np.zeros(N, dtype=np.int64)
    m = len(cols)
    for pos, c in enumerate(cols):
        col = X.getcol(int(c))
        packed[col.indices] += col.data.astype(np.int64) << (shift * (m - 1 - pos))
    idx = np.flatnonzero(packed)
    return idx.astype(np.int64), packed[idx]


def term_table(n_rows):
    out = np.empty(n_rows + 1, dtype=np.float64)
    lib().orc_term_table(ctypes.c_int64(n_rows), _p(out, _f64p))
    return out


# ----------------------------------------------------------------------------
# discretiser
# ----------------------------------------------------------------------------
def quantile_discretize_full(X, k=5, nthreads=1):
    """INFOSET:50-83.  Returns (codes uint8 column-major [P, N], edges [P, k-1],
    n_edges [P])."""
    if scipy.sparse.issparse(X):
        X = X.toarray()  # INFOSET:66-67
    X = np.asarray(X)
    N, P = X.shape
    if X.dtype == np.float32:
        fn, xt, et, xp = lib().orc_quantile_discretize_f32, np.float32, np.float32, ctypes.POINTER(ctypes.c_float)
    elif X.dtype == np.float64:
        fn, xt, et, xp = lib().orc_quantile_discretize_f64, np.float64, np.float64, _f64p
    elif np.issubdtype(X.dtype, np.integer):
        fn, xt, et, xp = lib().orc_quantile_discretize_i64, np.int64, np.float64, _i64p
    else:
        raise TypeError("unsupported dtype %s" % X.dtype)
    Xc = _c(X, xt)
    codes = np.zeros((P, N), dtype=np.uint8)
    edges = np.zeros((P, max(k - 1, 1)), dtype=et)
    n_edges = np.zeros(P, dtype=np.int32)
    rc = fn(_p(Xc, xp), ctypes.c_int64(N), ctypes.c_int64(P), ctypes.c_int64(k),
            _p(codes, _u8p), _p(edges, xp if et == xt else _f64p), _p(n_edges, _i32p),
            ctypes.c_int(nthreads))
    assert rc == 0
    return codes, edges, n_edges


def quantile_discretize(X, k=5, nthreads=1):
    """INFOSET:50-83 with the reference's return type: dense int64 [N, P]."""
    codes, _, _ = quantile_discretize_full(X, k, nthreads)
    return np.ascontiguousarray(codes.T).astype(np.int64)


# ----------------------------------------------------------------------------
# information sets
# ----------------------------------------------------------------------------
def _bit_length_like_ref(v):
    # INFOSET:112/199/216: int(np.log2(v) + 1)
    return int(np.log2(v) + 1)


class InformationSet:
    """INFOSET:86-173 (dense, numba jitclass in the reference)."""

    def __init__(self, X, has_y=False, nthreads=1):
        self.has_y = bool(has_y)
        self.X = _c(X, np.int64)
        self.N, self.P = self.X.shape
        self._shift = _bit_length_like_ref(self.X.max())
        self._nthreads = nthreads

    def entropy(self, cols):
        cols = _c(cols, np.int64)
        return float(lib().orc_dense_entropy(_p(self.X, _i64p), ctypes.c_int64(self.N),
                                             ctypes.c_int64(self.P), _p(cols, _i64p),
                                             ctypes.c_int64(cols.size),
                                             ctypes.c_int64(self._shift)))

    def entropy_wrt(self, cols):
        cols = _c(cols, np.int64)
        n_feats = self.P - 1 if self.has_y else self.P  # INFOSET:162
        out = np.empty(n_feats, dtype=np.float64)
        rc = lib().orc_dense_entropy_wrt(_p(self.X, _i64p), ctypes.c_int64(self.N),
                                         ctypes.c_int64(self.P), _p(cols, _i64p),
                                         ctypes.c_int64(cols.size), ctypes.c_int64(self._shift),
                                         ctypes.c_int64(n_feats), _p(out, _f64p),
                                         ctypes.c_int(self._nthreads))
        assert rc == 0
        return out


class SparseInformationSet:
    """INFOSET:176-303."""

    def __init__(self, X, y=None, nthreads=1):
        self.X = scipy.sparse.csc_matrix(X)
        assert np.issubdtype(self.X.dtype, np.integer), "X should be integer dtype"
        self.X.sum_duplicates()
        self.X.eliminate_zeros()
        assert self.X.has_canonical_format
        self.N, self.P = self.X.shape
        self._shift = _bit_length_like_ref(self.X.max())
        self._nthreads = nthreads
        self.set_y(y)

    def set_y(self, y):  # INFOSET:202-216
        self.has_y = y is not None
        self.y = y
        if self.has_y:
            if scipy.sparse.issparse(self.y):
                self.y = self.y.toarray().flatten()
            self.y = np.asarray(self.y)
            assert np.issubdtype(self.y.dtype, np.integer), "y should be integer dtype"
            assert self.y.shape == (self.N,)
            assert self.y.min() >= 0, "y should have non-negative integers only"
            self.y = self.y.astype(np.int64)
            self.classsizes = np.bincount(self.y, minlength=int(self.y.max()) + 1).astype(np.float64)
            self._ybits = _bit_length_like_ref(self.y.max())

    def entropy(self, cols):  # INFOSET:218-249
        cols = list(np.asarray(cols, dtype=np.int64))
        if len(cols) > 0 and cols[0] == -1:
            idx, vals = make_master_col(self.X, cols[1:], self._shift)
            packed = np.zeros(self.N, dtype=np.int64)
            packed[idx] = vals
            packed += self.y << (self._shift * len(cols))  # INFOSET:236 (len includes the -1)
            nz = np.flatnonzero(packed)
            return sparse_entropy(nz, packed[nz], self.N,
                                  2 ** (self._shift * len(cols) + self._ybits))
        idx, vals = make_master_col(self.X, cols, self._shift)
        return sparse_entropy(idx, vals, self.N, 2 ** (self._shift * len(cols)))

    def entropy_wrt(self, cols, feat_range=None):  # INFOSET:251-303
        cols = list(np.asarray(cols, dtype=np.int64))
        has_y = len(cols) > 0 and cols[0] == -1
        if has_y:
            cols = cols[1:]
            y, ybits, cs = self.y, self._ybits, self.classsizes
        else:
            y, ybits, cs = np.arange(0), 0, np.arange(0)
        mi, md = make_master_col(self.X, cols, self._shift)
        return sparse_entropy_wrt(self.X.indices, self.X.indptr, self.X.data, mi, md, self.N,
                                  self.P, len(cols), self._shift, ybits, y, cs,
                                  nthreads=self._nthreads, feat_range=feat_range)


class _FlatCSC:
    """The members of scipy's csc_matrix the oracle touches, over caller-owned flat arrays
    (no index widening, no int64 copy of the codes)."""

    def __init__(self, indptr, indices, data, shape):
        self.indptr, self.indices, self.data, self.shape = indptr, indices, data, tuple(shape)

    def max(self):
        return int(self.data.max())

    def getcol(self, j):
        b, e = int(self.indptr[j]), int(self.indptr[j + 1])
        return _FlatCSC(np.array([0, e - b]), self.indices[b:e], self.data[b:e], (self.shape[0], 1))


class FlatSparseInformationSet(SparseInformationSet):
    """INFOSET:176-303 over flat canonical CSC arrays (indptr int64[P+1], row ids int32[nnz]
    ascending inside a column, codes uint8[nnz], all non-zero).  For matrices with more than
    2^31 stored entries: scipy would widen the row ids to int64 and the reference keeps int64
    data (42 GB at BASELINE config 4); the algorithm and the C kernels are the same."""

    def __init__(self, indptr, indices, data, n_rows, y=None, nthreads=1):
        indptr = _c(indptr, np.int64)
        indices = _c(indices, np.int32)
        data = _c(data, np.uint8)
        assert indptr[0] == 0 and indptr[-1] == indices.size == data.size
        self.X = _FlatCSC(indptr, indices, data, (int(n_rows), indptr.size - 1))
        self.N, self.P = self.X.shape
        self._shift = _bit_length_like_ref(self.X.max())
        self._nthreads = nthreads
        self.set_y(y)


def synth_csc(spec, gene_lo=0, gene_hi=None, nthreads=None, genes=None):
    """Genes [gene_lo, gene_hi) — or the listed `genes` — of a
    picturedrocks_b200.synth.DiscretisedSpec through synth_cpu.c:
    (indptr int64, row ids int32, codes uint8, labels int64)."""
    L = lib()
    nthreads = host_cores() if nthreads is None else int(nthreads)
    gene_hi = spec.P if gene_hi is None else int(gene_hi)
    P = gene_hi - int(gene_lo)
    ids = None
    if genes is not None:
        ids = np.ascontiguousarray(genes, dtype=np.int64)
        P = ids.size
    y = np.empty(spec.N, dtype=np.int64)
    seed = ctypes.c_uint64(spec.seed & ((1 << 64) - 1))
    L.orc_synth_labels(_p(y, _i64p), ctypes.c_int64(spec.N), ctypes.c_int64(spec.C), seed)
    cum = np.ascontiguousarray(spec.cum_table(), dtype=np.uint32)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    args = (_p(y, _i64p), ctypes.c_int64(spec.N), ctypes.c_int64(gene_lo), ctypes.c_int64(P),
            ctypes.c_int64(spec.C), ctypes.c_int64(spec.K), ctypes.c_uint32(spec.mean_q24), seed,
            _p(cum, u32p))
    idp = _p(ids, _i64p) if ids is not None else None
    counts = np.zeros(P, dtype=np.int64)
    assert L.orc_synth_count(*args, _p(counts, _i64p), ctypes.c_int(nthreads), idp) == 0
    indptr = np.zeros(P + 1, dtype=np.int64)
    np.cumsum(counts, out=indptr[1:])
    rows = np.empty(int(indptr[-1]), dtype=np.int32)
    codes = np.empty(int(indptr[-1]), dtype=np.uint8)
    assert L.orc_synth_fill(*args, _p(indptr, _i64p), _p(rows, _i32p), _p(codes, _u8p),
                            ctypes.c_int(nthreads), idp) == 0
    return indptr, rows, codes, y


def makeinfoset(adata, include_y, k=5, nthreads=1):
    """INFOSET:25-47."""
    X = quantile_discretize(adata.X, k, nthreads)
    infoset = SparseInformationSet(X, None, nthreads)
    if include_y:
        infoset.set_y(adata.obs["y"].values)
    return infoset


# ----------------------------------------------------------------------------
# greedy selectors (ITER:24-134, BASE:22-83), restated as one table-driven class
# ----------------------------------------------------------------------------
class Selector:
    """kind ∈ {"CIFE", "JMI", "MIM", "CIFEUnsup", "UniEntropy"}.

    Score algebra is evaluated left-to-right exactly as numpy evaluates the
    reference's expressions (ITER:29-33, 50-56, 114-118):
      base  = (Hx + Hy) - Hyx                      supervised
      base  = Hx                                   unsupervised
      delta = (((base + Hj) - Hij) - Hyj) + Hyij   CIFE / JMI
      delta = (base + Hj) - Hij                    CIFEUnsup
    """

    SUPERVISED = {"CIFE": True, "JMI": True, "MIM": True, "CIFEUnsup": False, "UniEntropy": False}

    def __init__(self, kind, infoset):
        self.kind = kind
        self.infoset = infoset
        self.S = []
        e = np.arange(0)
        if self.SUPERVISED[kind]:
            assert infoset.has_y, "Information Set must have target labels"
            self.base_score = (infoset.entropy_wrt(e) + infoset.entropy(np.array([-1]))
                               - infoset.entropy_wrt(np.array([-1])))
        else:
            self.base_score = infoset.entropy_wrt(e)
        self.penalty = np.zeros(len(self.base_score))
        self.score = self.base_score.copy()

    def _delta(self, ind):
        I = self.infoset
        if self.kind == "CIFEUnsup":
            return self.base_score + I.entropy(np.array([ind])) - I.entropy_wrt(np.array([ind]))
        return (self.base_score + I.entropy(np.array([ind])) - I.entropy_wrt(np.array([ind]))
                - I.entropy(np.array([-1, ind])) + I.entropy_wrt(np.array([-1, ind])))

    def add(self, ind):
        ind = int(ind)
        self.S.append(ind)
        if self.kind in ("MIM", "UniEntropy"):  # BASE:70-72
            self.score[ind] = -np.inf
            return
        d = self._delta(ind)
        self.penalty += d
        if self.kind == "JMI":  # ITER:85-87
            self.score = self.base_score - (self.penalty / len(self.S))
            self.score[self.S] = -np.inf
        else:  # ITER:57-59, 119-121
            self.score -= d
            self.score[ind] = -np.inf

    def remove(self, ind):
        ind = int(ind)
        self.S.remove(ind)
        if self.kind in ("MIM", "UniEntropy"):  # BASE:74-76
            self.score[ind] = self.base_score[ind]
            return
        d = self._delta(ind)
        self.penalty -= d
        if self.kind == "JMI":  # ITER:98-100
            with np.errstate(divide="ignore", invalid="ignore"):
                self.score = self.base_score - (self.penalty / len(self.S))
        else:  # ITER:70-72, 131-133
            self.score = self.base_score - self.penalty
        self.score[self.S] = -np.inf

    def autoselect(self, n_feats):
        if self.kind in ("MIM", "UniEntropy"):  # BASE:78-83
            nbest = np.argpartition(self.score, -n_feats)[-n_feats:]
            nbest = nbest[np.argsort(self.score[nbest])[::-1]]
            self.S.extend(nbest)
            self.score[nbest] = -np.inf
            return nbest
        for _ in range(n_feats):  # BASE:64-66
            self.add(int(np.argmax(self.score)))


def CIFE(infoset):
    return Selector("CIFE", infoset)


def JMI(infoset):
    return Selector("JMI", infoset)


def MIM(infoset):
    return Selector("MIM", infoset)


def CIFEUnsup(infoset):
    return Selector("CIFEUnsup", infoset)


def UniEntropy(infoset):
    return Selector("UniEntropy", infoset)
