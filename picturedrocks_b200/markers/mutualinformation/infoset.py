"""Information sets — host-side mirror of the reference interface
(picturedrocks/markers/mutualinformation/infoset.py), computing on the GPU.

Same names, argument meaning and error behaviour as the reference:
  makeinfoset(adata, include_y, k=5)            infoset.py:25-47
  quantile_discretize(X, k=5)                   infoset.py:50-83
  InformationSet(X, has_y=False)                infoset.py:86-173
  SparseInformationSet(X, y=None)               infoset.py:176-303
Everything numeric runs in libprb200.so (see ../../engine.py); there is no CPU
fallback.
"""
import ctypes

import numpy as np
import scipy.sparse
import torch

from ... import _native as nat
from ...engine import Engine, NotCanonical

_DT = {np.dtype(np.float32): (0, torch.float32, torch.float32),
       np.dtype(np.float64): (1, torch.float64, torch.float64),
       np.dtype(np.int64): (2, torch.int64, torch.float64)}


def _device():
    nat.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def discretize_to_device(X, k=5, gene_block=None):
    """quantile_discretize on the GPU; returns column-major uint8 codes as a device
    tensor [P, N] (row g = gene g) — the layout the entropy kernels consume."""
    if not (2 <= int(k) <= 256):
        raise ValueError("k must be in [2, 256]")
    k = int(k)
    dev = _device()
    if scipy.sparse.issparse(X):
        X = X.toarray()  # same semantics as infoset.py:66-67 (zeros take part in the quantiles)
    if isinstance(X, torch.Tensor):
        Xh = None
        Xd_full = X.to(dev)
        if Xd_full.dtype.is_floating_point and Xd_full.dtype not in (torch.float32, torch.float64):
            Xd_full = Xd_full.to(torch.float32)  # half / bfloat16: never truncate to integers
        np_dt = {torch.float32: np.float32, torch.float64: np.float64}.get(Xd_full.dtype, np.int64)
        N, P = Xd_full.shape
    else:
        Xh = np.asarray(X)
        if Xh.ndim != 2:
            raise ValueError("X must be 2-d")
        if Xh.dtype not in (np.float32, np.float64):
            if not (np.issubdtype(Xh.dtype, np.integer) or Xh.dtype == np.bool_):
                raise TypeError("unsupported dtype %s" % Xh.dtype)
            Xh = Xh.astype(np.int64)
        np_dt = Xh.dtype.type
        N, P = Xh.shape
        Xd_full = None
    code, tdt, edt = _DT[np.dtype(np_dt)]
    esz = 4 if code == 0 else 8
    codes = torch.empty((P, N), dtype=torch.uint8, device=dev)
    if P == 0 or N == 0:
        return codes
    if gene_block is None:
        gene_block = max(1, min(P, 60000, ((1 << 31) - 1) // max(N, 1), (1 << 28) // max(N, 1) + 1))
    s = nat.stream()
    for g0 in range(0, P, gene_block):
        g1 = min(P, g0 + gene_block)
        pb = g1 - g0
        if Xd_full is not None:
            Xd = Xd_full[:, g0:g1].to(tdt).contiguous()
        else:
            Xd = torch.from_numpy(np.ascontiguousarray(Xh[:, g0:g1])).to(dev)
        Xt = torch.empty((pb, N), dtype=tdt, device=dev)
        nat.call("prb_transpose", nat.ptr(Xd), esz, N, pb, nat.ptr(Xt), s)
        Xs = torch.empty_like(Xt)
        ws_bytes = ctypes.c_int64()
        nat.call("prb_sort_columns_workspace", code, N, pb, ctypes.byref(ws_bytes))
        ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
        nat.call("prb_sort_columns", nat.ptr(Xt), nat.ptr(Xs), code, N, pb, nat.ptr(ws),
                 ws_bytes.value, s)
        edges = torch.zeros((pb, max(k - 1, 1)), dtype=edt, device=dev)
        n_edges = torch.zeros(pb, dtype=torch.int32, device=dev)
        nat.call("prb_qd_edges", nat.ptr(Xs), code, N, pb, k, nat.ptr(edges), nat.ptr(n_edges), s)
        nat.call("prb_qd_digitize", nat.ptr(Xt), code, N, pb, k, nat.ptr(edges), nat.ptr(n_edges),
                 nat.ptr(codes[g0:g1]), N, s)
    return codes


def _sparse_csc_on_device(X):
    """scipy sparse (cells x genes, any format) -> by-gene CSC of raw values on the device:
    (indptr int64[P+1], rows int32[nnz], vals float32 | float64 | int64 [nnz]), rows ascending
    inside a gene.  CSR input (how AnnData stores counts) is transposed on the GPU."""
    dev = _device()
    if not (scipy.sparse.isspmatrix_csr(X) or scipy.sparse.isspmatrix_csc(X)):
        X = scipy.sparse.csr_matrix(X)
    if not X.has_canonical_format:
        X = X.copy()
        X.sum_duplicates()  # also sorts the indices
    dt = X.dtype
    if dt not in (np.float32, np.float64):
        if not (np.issubdtype(dt, np.integer) or dt == np.bool_):
            raise TypeError("unsupported dtype %s" % dt)
        dt = np.dtype(np.int64)
    N, P = X.shape
    up = lambda a, t: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True).to(t)
    vals = torch.from_numpy(np.ascontiguousarray(X.data.astype(dt, copy=False))).to(dev)
    if scipy.sparse.isspmatrix_csc(X):
        return up(X.indptr, torch.int64), up(X.indices, torch.int32), vals, N, P
    # CSR -> CSC on the device: a counting transpose (csrc/ingest.cu), rows ascending inside
    # every gene without a sort
    col = torch.from_numpy(np.ascontiguousarray(X.indices)).to(dev, non_blocking=True)
    crow = up(X.indptr, torch.int64)
    nnz = int(col.numel())
    if nnz == 0:
        return (torch.zeros(P + 1, dtype=torch.int64, device=dev),
                torch.zeros(0, dtype=torch.int32, device=dev), vals[:0], N, P)
    B = int(max(1, min(1024, -(-N // 64), (256 << 20) // (4 * max(P, 1)))))
    T = torch.empty(B * P, dtype=torch.int32, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    s = nat.stream()
    nat.call("prb_csr_count", nat.ptr(crow), nat.ptr(col), col.element_size(), N, P, B,
             nat.ptr(T), nat.ptr(flags), s)
    indptr = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    torch.cumsum(T.view(B, P).sum(0, dtype=torch.int64), 0, out=indptr[1:])
    if int(flags.item()):
        raise ValueError("column index out of range in the CSR matrix")
    T64 = torch.empty(B * P, dtype=torch.int64, device=dev)
    rows = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    vout = torch.empty(max(nnz, 1), dtype=vals.dtype, device=dev)
    if nnz:
        nat.call("prb_csr_fill", nat.ptr(crow), nat.ptr(col), col.element_size(), nat.ptr(vals),
                 vals.element_size(), N, P, B, nat.ptr(T), nat.ptr(indptr), nat.ptr(T64),
                 nat.ptr(rows), nat.ptr(vout), s)
    return indptr, rows[:nnz], vout[:nnz], N, P


def discretize_sparse_to_engine(X, k=5, max_block_nnz=1 << 28, **engine_kw):
    """quantile_discretize of a NON-NEGATIVE scipy sparse matrix without densifying it
    (infoset.py:50-83 densifies, infoset.py:66-67): per gene the stored values are sorted,
    the N - nnz implicit zeros are accounted for by rank arithmetic, and only the stored
    values are digitised.  Returns the Engine holding the packed by-gene CSC of bin codes,
    or None when the coded matrix would not be sparse (an edge below zero)."""
    if not (2 <= int(k) <= 256):
        raise ValueError("k must be in [2, 256]")
    k = int(k)
    indptr, rows, vals, N, P = _sparse_csc_on_device(X)
    if N == 0 or P == 0:
        return None
    dev = indptr.device
    code, _, edt = _DT[np.dtype({torch.float32: np.float32, torch.float64: np.float64,
                                 torch.int64: np.int64}[vals.dtype])]
    if vals.numel() and bool((vals < 0).any().item()):
        return None
    s = nat.stream()
    edges = torch.zeros((P, max(k - 1, 1)), dtype=edt, device=dev)
    n_edges = torch.zeros(P, dtype=torch.int32, device=dev)
    zero_code = torch.zeros(P, dtype=torch.int32, device=dev)
    ip = indptr.cpu().numpy()
    g0 = 0
    while g0 < P:  # blocks of genes with a bounded number of stored values
        g1 = int(np.searchsorted(ip, ip[g0] + max_block_nnz, side="right")) - 1
        g1 = min(P, max(g1, g0 + 1))
        n = int(ip[g1] - ip[g0])
        if n >= (1 << 31):
            raise NotImplementedError("a single gene with 2^31 stored values")
        off = (indptr[g0: g1 + 1] - indptr[g0]).contiguous()
        vin = vals[int(ip[g0]): int(ip[g1])]
        vs = torch.empty_like(vin)
        if n:
            ws_bytes = ctypes.c_int64()
            nat.call("prb_sort_segments_workspace", code, n, g1 - g0, ctypes.byref(ws_bytes))
            ws = torch.empty(ws_bytes.value, dtype=torch.uint8, device=dev)
            nat.call("prb_sort_segments", nat.ptr(vin), nat.ptr(vs), code, n, g1 - g0,
                     nat.ptr(off), nat.ptr(ws), ws_bytes.value, s)
        nat.call("prb_qd_edges_sparse", nat.ptr(vs), nat.ptr(indptr[g0: g1 + 1].contiguous()),
                 code, N, g1 - g0, k, nat.ptr(edges[g0:g1]), nat.ptr(n_edges[g0:g1]),
                 nat.ptr(zero_code[g0:g1]), s)
        g0 = g1
    if bool((zero_code != 0).any().item()):
        return None
    counts = torch.zeros(P, dtype=torch.int64, device=dev)
    nat.call("prb_qd_count_sparse", nat.ptr(rows), nat.ptr(vals), nat.ptr(indptr), code, P, k,
             nat.ptr(edges), nat.ptr(n_edges), nat.ptr(counts), s)
    out_indptr = torch.zeros(P + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=out_indptr[1:])
    nnz = int(out_indptr[-1].item())
    from ...engine import PACKED_PAD

    packed = torch.zeros(nnz + PACKED_PAD, dtype=torch.int32, device=dev)
    nat.call("prb_qd_fill_sparse", nat.ptr(rows), nat.ptr(vals), nat.ptr(indptr), code, P, k,
             nat.ptr(edges), nat.ptr(n_edges), nat.ptr(out_indptr), nat.ptr(packed), s)
    K = int(n_edges.max().item()) + 1
    return Engine(out_indptr, packed, N, K, **engine_kw)


def quantile_discretize(X, k=5):
    """Discretize a data matrix with the recursive quantile transform (infoset.py:50-83).

    Returns the reference's type: dense int64 ndarray [n_obs, n_features]."""
    codes = discretize_to_device(X, k)
    return np.ascontiguousarray(codes.cpu().numpy().T).astype(np.int64)


def makeinfoset(adata, include_y, k=5):
    """Discretize `adata.X` and wrap it in a SparseInformationSet (infoset.py:25-47).
    The codes never leave the GPU: dense codes -> packed by-gene CSC on device; a
    non-negative scipy sparse `adata.X` (CSR counts, as AnnData stores them) is never
    densified (same codes as the reference, which densifies: infoset.py:66-67)."""
    eng = discretize_sparse_to_engine(adata.X, k) if scipy.sparse.issparse(adata.X) else None
    if eng is not None:
        infoset = SparseInformationSet._from_engine(eng)
    else:
        infoset = SparseInformationSet._from_device_codes(discretize_to_device(adata.X, k))
    if include_y:
        infoset.set_y(adata.obs["y"].values)
    return infoset


def _check_codes_u8(vals):
    if vals.size and (vals.min() < 0 or vals.max() > 255):
        raise ValueError("bin codes must lie in [0, 255]")


class _SelectorBackend:
    """What the greedy selectors need from an information set, as device tensors."""

    def __init__(self, engine, n_feats, ycol=None):
        self.engine = engine
        self.n_feats = n_feats  # local candidate features
        self.ycol = ycol  # dense sets: index of the label column inside the matrix

    @property
    def device(self):
        return self.engine.device

    def base_vectors(self, supervised):
        e = self.engine
        if self.ycol is None:
            Hx = e.entropy_wrt_device([])
            if not supervised:
                return Hx, None, None
            return Hx, e.entropy_scalar([-1]), e.entropy_wrt_device([-1])
        Hx = e.entropy_wrt_device([])[: self.n_feats].contiguous()
        if not supervised:
            return Hx, None, None
        return (Hx, e.entropy_scalar([self.ycol]),
                e.entropy_wrt_device([self.ycol])[: self.n_feats].contiguous())

    def step(self, supervised, host_ind=None):
        """(H(x_j, x_.), H(y, x_j, x_.)) for the gene held in engine.gene_ptr."""
        e = self.engine
        if self.ycol is None:
            full = torch.empty(e.P, dtype=torch.float64, device=e.device)
            if supervised:
                marg = torch.empty_like(full)
                e.direct_step(True, full, marg)
                return marg, full
            e.direct_step(False, full, None)
            return full, None
        j = int(e.gene_ptr.item()) if host_ind is None else int(host_ind)
        Hij = e.entropy_wrt_device([j])[: self.n_feats].contiguous()
        if not supervised:
            return Hij, None
        return Hij, e.entropy_wrt_device([self.ycol, j])[: self.n_feats].contiguous()


class SparseInformationSet:
    """Sparse discrete gene-expression matrix on the GPU (infoset.py:176-303).

    Args
    ----
    X: scipy.sparse matrix or ndarray, shape (num_obs, num_vars), integer dtype
    y: None or integer ndarray of shape (num_obs,) with non-negative labels

    Extra keyword arguments (multi-GPU, one process per GPU): `gene_lo`, `P_total`
    and `comm` declare that X holds only genes [gene_lo, gene_lo + X.shape[1]).
    """

    def __init__(self, X, y=None, gene_lo=0, P_total=None, comm=None):
        # infoset.py:191-196.  A matrix that already is canonical CSC (the normal case) is
        # taken as it is — its arrays are uploaded, packed and CHECKED on the device
        # (csrc/ingest.cu); only a matrix that fails the check goes through scipy's
        # sum_duplicates / eliminate_zeros on the host, as in the reference.
        Xc = X if scipy.sparse.issparse(X) and X.format == "csc" else scipy.sparse.csc_matrix(X)
        assert np.issubdtype(Xc.dtype, np.integer) or Xc.dtype == np.bool_, \
            "X should be integer dtype"
        N, P = Xc.shape
        kw = dict(gene_lo=gene_lo, P_total=P_total, comm=comm)
        try:
            eng = Engine.from_csc_arrays(Xc.indptr, Xc.indices, Xc.data, N, **kw)
        except NotCanonical as e:
            if e.flags & 1:
                raise ValueError("row index out of range")
            Xc = Xc.copy()
            Xc.sum_duplicates()
            Xc.eliminate_zeros()
            assert Xc.has_canonical_format, "Sparse matrix not in canonical format"
            _check_codes_u8(Xc.data)
            eng = Engine.from_csc_arrays(Xc.indptr, Xc.indices, Xc.data, N, **kw)
        self._X = Xc
        self._attach(eng)
        self.set_y(y)

    def _attach(self, eng):
        self._engine = eng
        self.N = eng.N
        self.P = eng.P_total
        # infoset.py:199 int(np.log2(X.max()) + 1); an all-zero matrix raises there too
        if eng.K < 2:
            raise OverflowError("cannot convert float infinity to integer")
        self._shift = eng.shift
        self.has_y = False
        self.y = None

    @classmethod
    def _from_engine(cls, eng, y=None):
        self = cls.__new__(cls)
        self._X = None
        self._attach(eng)
        self.set_y(y)
        return self

    @classmethod
    def _from_device_codes(cls, codes, **kw):
        P, N = codes.shape
        return cls._from_engine(Engine.from_dense_codes(codes, N, **kw))

    @property
    def X(self):
        if self._X is None:
            self._X = self._engine.to_scipy_csc()
        return self._X

    def set_y(self, y):
        """Attach / replace / drop (None) the target labels (infoset.py:202-216)."""
        self.has_y = y is not None
        self.y = y
        if not self.has_y:
            self._engine.set_y(None)
            return
        if isinstance(y, torch.Tensor):
            yd = y.to(self._engine.device)
            assert not yd.dtype.is_floating_point, "y should be integer dtype"
            assert tuple(yd.shape) == (self.N,)
            assert int(yd.min().item()) >= 0, \
                "y should have non-negative integers only; try using np.unique with return_inverse"
            yd = yd.to(torch.int32).contiguous()
            self.y = y
        else:
            if scipy.sparse.issparse(y):
                y = y.toarray().flatten()
            y = np.asarray(y)
            self.y = y
            assert np.issubdtype(y.dtype, np.integer), "y should be integer dtype"
            assert y.shape == (self.N,)
            assert y.min() >= 0, \
                "y should have non-negative integers only; try using np.unique with return_inverse"
            if y.max() == 0:
                raise OverflowError("cannot convert float infinity to integer")  # infoset.py:216
            yd = torch.from_numpy(y.astype(np.int32)).to(self._engine.device)
        self._engine.set_y(yd)
        self.classsizes = self._engine.classsizes.cpu().numpy().astype(np.float64)
        self._ybits = int(self.classsizes.size - 1).bit_length()

    def entropy(self, cols):
        """Shannon entropy of the joint of `cols` (-1 as FIRST entry = y; infoset.py:218-249)."""
        return self._engine.entropy_scalar(cols)

    def entropy_wrt(self, cols):
        """H(cols, x_i) for every feature i (infoset.py:251-303) as float64 ndarray."""
        out = self._engine.entropy_wrt_device(cols)
        return self._gather(out)

    def _gather(self, t):
        comm = self._engine.comm
        if comm is not None:
            t = comm.all_gather_shards(t)
        return t.cpu().numpy()

    def _selector_backend(self):
        return _SelectorBackend(self._engine, self._engine.P)


class InformationSet:
    """Dense discrete gene-expression matrix (infoset.py:86-173).

    Args
    ----
    X: ndarray (num_obs, num_vars) of dtype int; if `has_y`, the LAST column is the
       target label column.
    The matrix (label column included, exactly like the reference, which also sizes
    `_shift` from it, infoset.py:112) is compacted on the GPU into the packed by-gene
    form, so dense and sparse sets share one set of kernels.
    """

    def __init__(self, X, has_y=False):
        X = np.asarray(X)
        assert np.issubdtype(X.dtype, np.integer), "X should be integer dtype"
        assert X.ndim == 2
        _check_codes_u8(X)
        self.has_y = bool(has_y)
        self.X = X.astype(np.int64)
        self.N, self.P = X.shape
        if X.max() <= 0:
            raise OverflowError("cannot convert float infinity to integer")  # infoset.py:112
        self._shift = int(X.max()).bit_length()
        dev = _device()
        codes = torch.from_numpy(np.ascontiguousarray(X.T.astype(np.uint8))).to(dev)
        self._engine = Engine.from_dense_codes(codes, self.N)
        self._n_feats = self.P - 1 if self.has_y else self.P  # infoset.py:162
        self._codes = codes if self.has_y else None  # (kept for the selectors' engine below)
        self._feat_engine = None

    def entropy(self, cols):
        """Entropy of the joint of `cols`; negative indices count from the end (numpy)."""
        return self._engine.entropy_scalar(self._cols(cols))

    def entropy_wrt(self, cols):
        out = self._engine.entropy_wrt_device(self._cols(cols))
        return out[: self._n_feats].cpu().numpy()

    def _cols(self, cols):
        cols = [int(c) for c in np.asarray(cols, dtype=np.int64).ravel()]
        # every column (the label column too) is an ordinary column here
        return [c + self.P if c < 0 else c for c in cols]

    def _selector_backend(self):
        """Greedy selectors on a set with a label column run on a second engine that holds the
        feature columns only and takes the label column as y: the same joint counts and the
        same (y, x_j, x_i) bin order as the generic column path (the label column is the most
        significant field of the packed state, infoset.py:133-136), but through the fused
        step — three launches per step, CUDA graph, no host synchronisation — instead of a
        state-table build with a host sync per step."""
        if not self.has_y:
            return _DenseNoY(self._engine, self._n_feats)
        if self._feat_engine is None and self._n_feats > 0:
            try:
                eng = Engine.from_dense_codes(self._codes[: self._n_feats].contiguous(), self.N)
                eng.set_y(self._codes[self._n_feats].to(torch.int32))
                self._feat_engine = eng
            except (ValueError, OverflowError, NotImplementedError):
                self._feat_engine = False  # (all-zero features, too many classes: generic path)
        if self._feat_engine:
            return _SelectorBackend(self._feat_engine, self._n_feats, ycol=None)
        return _SelectorBackend(self._engine, self._n_feats, ycol=self.P - 1)


class _DenseNoY(_SelectorBackend):
    """Dense set without labels: unsupervised selectors only, generic column path."""

    def __init__(self, engine, n_feats):
        super().__init__(engine, n_feats, ycol=-1)

    def base_vectors(self, supervised):
        assert not supervised
        return self.engine.entropy_wrt_device([]).contiguous(), None, None
