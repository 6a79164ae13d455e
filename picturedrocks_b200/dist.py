"""Multi-GPU plumbing: genes are sharded across ranks (one process per GPU).

`entropy_wrt` is independent per candidate gene (the reference's serial loop,
infoset.py:399-409), so each rank streams only its contiguous gene range.  Per
greedy step the ranks exchange exactly two things:
  1. the selected gene's dense code vector x_j[N] (uint8): the owner scatters its
     column into a zeroed buffer, everyone max-all-reduces it (fixed size, so the
     host never needs to know which rank owns the winner);
  2. one (score, global index) candidate per rank: all-gather, then every rank runs
     the same deterministic reduce (max score, lowest index on ties).
Both are enqueued on the compute stream's NCCL communicator without host syncs.

EXPERIMENTAL (PRB_PEER_EXCHANGE=1, NCCL backend only; off by default — written in round 1,
not yet run on hardware): `PeerExchange` does both exchanges with the kernels of
csrc/peer.cu over NVLink peer memory (a symmetric buffer from
torch.distributed._symmetric_memory) instead of the two NCCL collectives.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def partition_genes(P_total, world):
    """Contiguous, equal-count gene ranges [(lo, hi), ...] for `world` ranks."""
    base, rem = divmod(int(P_total), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class Comm:
    def __init__(self, group=None, shard_sizes=None):
        assert dist.is_initialized(), "torch.distributed is not initialised"
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shard_sizes = list(shard_sizes) if shard_sizes is not None else None
        self.peer = None  # PeerExchange, see enable_peer()

    def enable_peer(self, N):
        """Set up the peer-memory exchange for matrices of N cells if PRB_PEER_EXCHANGE=1
        (returns the PeerExchange, or None when it is switched off / not applicable)."""
        if (self.peer is None and os.environ.get("PRB_PEER_EXCHANGE", "0") == "1"
                and self.shard_sizes is not None and dist.get_backend(self.group) == "nccl"):
            self.peer = PeerExchange(self, N)
        assert self.peer is None or self.peer.N == int(N), "one PeerExchange per cell count"
        return self.peer

    def max_int(self, v, device=None):
        t = torch.tensor([int(v)], dtype=torch.int64,
                         device=device if device is not None else self._dev())
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def _dev(self):
        return (torch.device("cuda", torch.cuda.current_device())
                if dist.get_backend(self.group) == "nccl" else torch.device("cpu"))

    def allreduce_max_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def allreduce_sum_(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_cat(self, t):
        """Concatenate equally sized per-rank tensors in rank order."""
        out = torch.empty((self.world * t.numel(),), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous().view(-1), group=self.group)
        return out

    def all_gather_shards(self, t):
        """Concatenate per-gene vectors of the (possibly unequal) gene shards."""
        sizes = self.shard_sizes
        assert sizes is not None and t.numel() == sizes[self.rank]
        if len(set(sizes)) == 1:
            return self.all_gather_cat(t)
        m = max(sizes)
        pad = torch.zeros(m, dtype=t.dtype, device=t.device)
        pad[: t.numel()] = t
        g = self.all_gather_cat(pad).view(self.world, m)
        return torch.cat([g[r, : sizes[r]] for r in range(self.world)])


class PeerExchange:
    """Symmetric buffer + device-side step state of the peer-memory exchange
    (csrc/peer.cu; call order in include/prb200.h)."""

    def __init__(self, comm, N):
        import torch.distributed._symmetric_memory as symm

        from . import _native as nat

        self.nat, self.N, self.rank, self.world = nat, int(N), comm.rank, comm.world
        nbytes, off_col, n_pad = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        nat.call("prb_peer_layout", self.N, self.world, ctypes.byref(nbytes), ctypes.byref(off_col),
                 ctypes.byref(n_pad))
        dev = torch.device("cuda", torch.cuda.current_device())
        group = comm.group if comm.group is not None else dist.group.WORLD
        self.buf = symm.empty(nbytes.value, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64,
                                  device=dev)
        self.col = self.buf[off_col.value: off_col.value + n_pad.value]  # this rank's x_j area
        self.n_pad = n_pad.value
        lo = [0]
        for n in comm.shard_sizes:
            lo.append(lo[-1] + int(n))
        self.shard_lo = torch.tensor(lo, dtype=torch.int32, device=dev)
        self.seq = torch.zeros(1, dtype=torch.int32, device=dev)
        self.dirty = torch.zeros(1, dtype=torch.int32, device=dev)
        self.records = torch.zeros(2 * self.world, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group=comm.group)  # every buffer is zeroed before the first record arrives

    def gather_candidates(self, cand_score, cand_idx, n_cand):
        """records[W][2] <- every rank's (score, index) candidate (advances the step)."""
        nat = self.nat
        nat.call("prb_peer_gather16", nat.ptr(cand_score), nat.ptr(cand_idx), int(n_cand),
                 nat.ptr(self.peers), self.rank, self.world, self.N, nat.ptr(self.seq),
                 nat.ptr(self.records), nat.stream())
        return self.records

    def clean(self):
        nat = self.nat
        nat.call("prb_peer_clean", nat.ptr(self.buf), self.N, self.world, nat.ptr(self.dirty),
                 nat.stream())

    def fetch_column(self, gene_ptr, xj_local):
        """xj_local <- the winning gene's dense codes, pulled from its owner."""
        nat = self.nat
        assert xj_local.numel() >= self.n_pad
        nat.call("prb_peer_column", nat.ptr(self.peers), self.rank, self.world, self.N,
                 nat.ptr(gene_ptr), nat.ptr(self.shard_lo), nat.ptr(self.seq), nat.ptr(xj_local),
                 nat.ptr(self.dirty), nat.stream())


def shutdown():
    """Release captured graphs, then tear the process group down."""
    import gc

    from .markers.mutualinformation.iterative import release_graphs

    release_graphs()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
