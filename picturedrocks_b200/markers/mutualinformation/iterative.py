"""Greedy mutual-information selectors, device-resident
(reference: picturedrocks/markers/mutualinformation/iterative.py:24-134).

Class names, constructor signature, attributes (`infoset`, `S`, `base_score`,
`penalty`, `score`) and methods (`add`, `remove`, `autoselect`) follow the
reference.  Scores live on the GPU; `score` / `penalty` / `base_score` are
materialised as float64 ndarrays on access, `S` as a Python list.
`autoselect(n)` runs n greedy steps without any host synchronisation; on the V-layout
path (K <= 5) a step is three launches (+ one 16-byte all-gather when genes are sharded):

    prb_pick_scatter   argmax of the candidate records, commit, column of the winner
    prb_stream_v       one streaming pass (joint histograms)
    prb_finalize_update  entropies, score update, this rank's next candidate record

Score algebra is evaluated in the reference's order (iterative.py:29-33, 50-58,
78-86, 114-120) so results are bit-identical given bit-identical entropies.
"""
import ctypes
import os
import weakref

import numpy as np
import torch

from ... import _native as nat
from .._iterative import IterativeFeatureSelection, UnivariateMixin

_KIND = {"CIFE": 0, "JMI": 1, "CIFEUnsup": 2}
# capture the greedy step as a CUDA graph (set False to launch every kernel eagerly)
USE_CUDA_GRAPH = os.environ.get("PRB_CUDA_GRAPH", "1") != "0"
GRAPH_MIN_STEPS = int(os.environ.get("PRB_GRAPH_MIN_STEPS", "8"))
_LIVE = weakref.WeakSet()  # selectors that may hold a captured graph


def release_graphs():
    """Drop every captured CUDA graph (must precede destroy_process_group: a live graph
    that contains NCCL nodes makes the communicator teardown hang)."""
    for fs in list(_LIVE):
        fs._graph = None


def _backend(infoset):
    try:
        return infoset._selector_backend()
    except AttributeError:
        raise TypeError(
            "picturedrocks_b200 selectors need a picturedrocks_b200 information set "
            "(GPU-resident); got %r" % type(infoset).__name__)


class _DeviceScores:
    """base/penalty/score vectors of one selector, resident on the GPU."""

    def __init__(self, infoset, supervised):
        self.be = _backend(infoset)
        e = self.be.engine
        self.engine = e
        self.comm = e.comm
        self.dev = e.device
        self.P = self.be.n_feats  # local candidates
        self.gene_lo = e.gene_lo
        self.P_total = e.P_total if self.be.ycol is None else self.be.n_feats
        self.supervised = supervised
        Hx, Hy, Hyx = self.be.base_vectors(supervised)
        f64 = dict(dtype=torch.float64, device=self.dev)
        if supervised:
            self.base = torch.empty(self.P, **f64)
            nat.call("prb_mi_base", nat.ptr(Hx), ctypes.c_double(Hy), nat.ptr(Hyx), self.P,
                     nat.ptr(self.base), nat.stream())
        else:
            self.base = Hx.clone()
        # replicated H(x_j), H(y,x_j) lookup vectors (== the reference's scalar
        # entropy([j]) / entropy([-1, j]) bit-for-bit: same bins, same order)
        self.Hx_all = self._replicate(Hx)
        self.Hyx_all = self._replicate(Hyx) if supervised else None
        self.penalty = torch.zeros(self.P, **f64)
        self.score = self.base.clone()
        self.selected = torch.zeros(max(self.P, 1), dtype=torch.uint8, device=self.dev)
        # V-layout path: the whole step runs in engine.fused_step (three launches)
        if e.vmode and self.be.ycol is None and e.P:
            e._v()  # (the layout decides whether the hit table can be v-major)
        self.fused = (e.vmode and (e.K1 <= 4 or e.wide_ok(supervised)) and self.be.ycol is None
                      and (self.comm is None or min(self.comm.shard_sizes) > 0))
        self.n_blocks = max(1, -(-self.P // 256))
        n_cand = self.n_blocks
        if self.fused:
            C = e.C if supervised else 1
            n_cand = max(n_cand, int(nat.load().prb_finalize_update_blocks(self.P, C)),
                         int(nat.load().prb_wide_chain_blocks(self.P)))
        self.cand_score = torch.zeros(n_cand, **f64)
        self.cand_idx = torch.full((n_cand,), -1, dtype=torch.int64, device=self.dev)
        # one 16-byte (score, idx) record: a single all-gather moves both
        self.best = torch.zeros(2, **f64)
        self.best.view(torch.int64)[1] = -1
        self.best_score = self.best[0:1]
        self.best_idx = self.best.view(torch.int64)[1:2]
        # fused path: the candidate records the next pick reduces — this rank's own, or one
        # per rank (all-gathered at the end of every step)
        self.n_rec = 1 if self.comm is None else self.comm.world
        self.records = self.best if self.comm is None else torch.zeros(2 * self.n_rec, **f64)
        self.records_valid = False
        # ... or written by the peers straight into this rank's buffer over NVLink (dist.PeerRecords)
        self.peer = None
        if self.fused and self.comm is not None:
            from ...dist import PeerRecords

            if PeerRecords.available(self.comm):
                self.peer = PeerRecords(self.comm)
        self.S_dev = torch.zeros(self.P_total + 1, dtype=torch.int32, device=self.dev)
        self.n_sel = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.cand_valid = False

    def _replicate(self, v):
        if self.comm is None:
            return v
        return self.comm.all_gather_shards(v)

    def gather(self, v):
        if self.comm is not None:
            v = self.comm.all_gather_shards(v)
        out = v.cpu().numpy()
        if self.peer is not None:
            self.peer.check()  # (a record that never arrived: raise instead of returning garbage)
        return out

    # -- device steps ---------------------------------------------------------
    def _share_best(self):
        """records <- every rank's (score, index) record."""
        if self.peer is not None:
            nat.call("prb_peer_share", nat.ptr(self.best), nat.ptr(self.peer.peer_ptrs),
                     self.comm.world, self.comm.rank, nat.ptr(self.peer.xtag), nat.stream())
        elif self.comm is not None:
            import torch.distributed as dist

            dist.all_gather_into_tensor(self.records, self.best, group=self.comm.group)
        self.records_valid = True

    def fused_step(self, kind, is_remove=False, pick=True):
        """V-layout path: one greedy step (pick = True) or one add / remove of the gene in
        engine.gene_ptr, see Engine.fused_step."""
        if pick and not self.records_valid:
            # first step after construction: candidates straight from the score vector
            s = nat.stream()
            if self.P:
                nat.call("prb_argmax_blocks", nat.ptr(self.score), self.P,
                         ctypes.c_int32(self.gene_lo), nat.ptr(self.cand_score),
                         nat.ptr(self.cand_idx), s)
                nat.call("prb_best_record", nat.ptr(self.cand_score), nat.ptr(self.cand_idx),
                         self.n_blocks, nat.ptr(self.best), s)
            self._share_best()
        self.engine.fused_step(self, _KIND[kind], is_remove, pick)
        if self.peer is None:
            self._share_best()
        else:
            # with peer buffers the kernel's tail has sent the record — also after an add /
            # remove that came first: the next pick must wait for THAT exchange, a second send
            # without a wait in between would break the one-ahead rule of csrc/peer.cu
            self.records_valid = True

    def pick_and_commit(self):
        """best_idx <- argmax(score) (global index; lowest index on ties), then commit it:
        gene_ptr <- best, S.append(best), selected[best] = 1."""
        s = nat.stream()
        e = self.engine
        if not self.cand_valid:
            nat.call("prb_argmax_blocks", nat.ptr(self.score), self.P, ctypes.c_int32(self.gene_lo),
                     nat.ptr(self.cand_score), nat.ptr(self.cand_idx), s)
        commit = (nat.ptr(e.gene_ptr), nat.ptr(self.S_dev), nat.ptr(self.n_sel),
                  nat.ptr(self.selected), ctypes.c_int32(self.gene_lo), self.P, s)
        if self.comm is None:
            nat.call("prb_argmax_commit", nat.ptr(self.cand_score), nat.ptr(self.cand_idx),
                     self.n_blocks, 1, nat.ptr(self.best_score), nat.ptr(self.best_idx), *commit)
            return
        nat.call("prb_argmax_final", nat.ptr(self.cand_score), nat.ptr(self.cand_idx),
                 self.n_blocks, 1, nat.ptr(self.best_score), nat.ptr(self.best_idx), s)
        g = self.comm.all_gather_cat(self.best)  # [world][score, idx]
        nat.call("prb_argmax_commit", nat.ptr(g), nat.ptr(g) + 8, self.comm.world, 2,
                 nat.ptr(self.best_score), nat.ptr(self.best_idx), *commit)

    def apply(self, kind, is_remove, host_ind=None):
        """penalty/score update for the gene in engine.gene_ptr."""
        if self.fused:
            self.fused_step(kind, is_remove, pick=False)
            return
        Hij, Hyij = self.be.step(self.supervised, host_ind)
        nat.call("prb_selector_update", _KIND[kind], int(is_remove), nat.ptr(self.base),
                 nat.ptr(self.penalty), nat.ptr(self.score), nat.ptr(self.selected), nat.ptr(Hij),
                 nat.ptr(Hyij), nat.ptr(self.Hx_all), nat.ptr(self.Hyx_all),
                 nat.ptr(self.engine.gene_ptr), nat.ptr(self.n_sel), self.P,
                 ctypes.c_int32(self.gene_lo), nat.ptr(self.cand_score), nat.ptr(self.cand_idx),
                 nat.stream())
        self.cand_valid = True


class _GreedyDevice(IterativeFeatureSelection):
    """Shared machinery of CIFE / JMI / CIFEUnsup."""

    _kind = None
    _supervised = True

    def __init__(self, infoset):
        if self._supervised:
            assert infoset.has_y, "Information Set must have target labels"
        self.infoset = infoset
        self._d = _DeviceScores(infoset, self._supervised)
        self._S = []
        self._pending = 0  # selections made on device, not yet mirrored in _S
        self._graph = None
        self._graph_epoch = -1
        self.graph_kernels_per_step = 0
        _LIVE.add(self)

    # -- host-visible state -----------------------------------------------------
    @property
    def S(self):
        if self._pending:
            n = len(self._S) + self._pending
            self._S = [int(v) for v in self._d.S_dev[:n].cpu().numpy()]
            self._pending = 0
            if self._d.peer is not None:
                self._d.peer.check()  # a peer that never delivered a record: raise, do not return garbage
        return self._S

    @S.setter
    def S(self, value):
        self._S = list(value)
        self._pending = 0
        self._push_S()

    @property
    def score(self):
        return self._d.gather(self._d.score)

    @property
    def penalty(self):
        return self._d.gather(self._d.penalty)

    @property
    def base_score(self):
        return self._d.gather(self._d.base)

    def _push_S(self):
        d = self._d
        S = self._S
        d.n_sel.fill_(len(S))
        if S:
            d.S_dev[: len(S)] = torch.tensor(S, dtype=torch.int32, device=d.dev)
        d.selected.zero_()
        loc = [s - d.gene_lo for s in S if 0 <= s - d.gene_lo < d.P]
        if loc:
            d.selected[torch.tensor(loc, dtype=torch.int64, device=d.dev)] = 1

    def _check(self, ind):
        ind = int(ind)
        if ind < 0:
            ind += self._d.P_total
        if not (0 <= ind < self._d.P_total):
            raise IndexError("feature index out of range")
        return ind

    # -- reference API ------------------------------------------------------------
    def add(self, ind):
        ind = self._check(ind)
        self.S.append(ind)
        self._push_S()
        self._d.engine.gene_ptr.fill_(ind)
        self._d.apply(self._kind, False, ind)

    def remove(self, ind):
        ind = self._check(ind)
        self.S.remove(ind)  # ValueError if absent, like list.remove in the reference
        self._push_S()
        self._d.engine.gene_ptr.fill_(ind)
        self._d.apply(self._kind, True, ind)

    def _one_step(self):
        d = self._d
        if d.fused:
            d.fused_step(self._kind)
            return
        d.pick_and_commit()
        d.apply(self._kind, False)

    def _capture(self):
        """Capture one greedy step (kernels + collectives) on a side stream."""
        d = self._d
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        n0 = nat.LAUNCHES
        with torch.cuda.stream(side):
            g.capture_begin(capture_error_mode="thread_local")
            try:
                self._one_step()
            finally:
                g.capture_end()
        torch.cuda.current_stream().wait_stream(side)
        self.graph_kernels_per_step = nat.LAUNCHES - n0
        self._graph, self._graph_epoch = g, d.engine.epoch

    def _graph_ok(self):
        d = self._d
        return (d.be.ycol is None and USE_CUDA_GRAPH
                and getattr(d.engine, "profile_events", None) is None)

    def autoselect(self, n_feats):
        """Greedy selection of n_feats features, entirely on the device.  After one eager
        step the step is captured once as a CUDA graph (kernels + the two NCCL collectives)
        and replayed, so the host neither synchronises nor pays per-launch overhead."""
        n_feats = int(n_feats)
        d = self._d
        if len(self._S) + self._pending + n_feats > d.P_total:
            raise ValueError("cannot select more features than there are")
        left = n_feats
        if self._graph_ok():
            if self._graph is not None and self._graph_epoch != d.engine.epoch:
                self._graph = None
            if self._graph is None and left >= GRAPH_MIN_STEPS:
                self._one_step()  # warm: lazy buffers, NCCL, function attributes
                left -= 1
                self._capture()
            if self._graph is not None:
                for _ in range(left):
                    self._graph.replay()
                left = 0
        for _ in range(left):
            self._one_step()
        self._pending += n_feats


class CIFE(_GreedyDevice):
    """Conditional infomax feature extraction: score = I(x;y) - Σ_j I(x;x_j;y)."""

    _kind = "CIFE"


class JMI(_GreedyDevice):
    """Joint mutual information: score = I(x;y) - mean_j I(x;x_j;y)."""

    _kind = "JMI"


class CIFEUnsup(_GreedyDevice):
    """Unsupervised CIFE: score = H(x) - Σ_j I(x;x_j)."""

    _kind = "CIFEUnsup"
    _supervised = False


class MIMixin:
    """Univariate supervised base score I(x_i; y) (iterative.py:24-35)."""

    def __init__(self, infoset):
        assert infoset.has_y, "Information Set must have target labels"
        self.infoset = infoset
        self.S = []
        d = _DeviceScores(infoset, True)
        self.base_score = d.gather(d.base)
        self.penalty = np.zeros(len(self.base_score))
        self.score = self.base_score.copy()


class EntropyMixin:
    """Univariate unsupervised base score H(x_i) (iterative.py:38-44)."""

    def __init__(self, infoset):
        self.infoset = infoset
        self.S = []
        d = _DeviceScores(infoset, False)
        self.base_score = d.gather(d.base)
        self.penalty = np.zeros(len(self.base_score))
        self.score = self.base_score.copy()


class MIM(MIMixin, UnivariateMixin, IterativeFeatureSelection):
    """Mutual information maximisation (no redundancy term)."""


class UniEntropy(EntropyMixin, UnivariateMixin, IterativeFeatureSelection):
    """Highest-entropy features."""
