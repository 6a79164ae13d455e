"""Multi-GPU plumbing: genes are sharded across ranks (one process per GPU).

`entropy_wrt` is independent per candidate gene (the reference's serial loop,
infoset.py:399-409), so each rank streams only its contiguous gene range.  At
set-up every rank receives a copy of the WHOLE packed CSC over NVLink (one all-gather,
10.5 GB at BASELINE config 4 — 6 % of a B200's HBM) and of the per-gene count tables, so
the selected gene's column and state sizes are read locally.  Per greedy step the ranks
exchange exactly one thing: one 16-byte (score, global index) candidate per rank —
all-gather, then every rank runs the same deterministic reduce (max score, lowest index
on ties = np.argmax, picturedrocks/markers/_iterative.py:65).  Only the streamed layout
and the score vectors stay sharded.
Both are enqueued on the compute stream's NCCL communicator without host syncs.
"""
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


def partition_genes_weighted(weights, world):
    """Contiguous gene ranges [(lo, hi), ...] of nearly equal total weight — the stored entries
    per gene: a rank's step time is its shard's nnz, not its gene count (config 4 cut into 8
    equal-count ranges leaves the fullest shard 5.9 % above the mean, and every step waits for
    it).  Every rank gets at least one gene when there are enough."""
    import numpy as np

    w = np.asarray(weights, dtype=np.float64).ravel()
    P, world = w.size, int(world)
    c = np.concatenate([[0.0], np.cumsum(w)])
    cuts = np.searchsorted(c, c[-1] * np.arange(1, world) / world, side="left")
    bounds = np.concatenate([[0], cuts, [P]]).astype(np.int64)
    if P >= world:  # strictly increasing bounds
        for r in range(1, world):
            bounds[r] = min(max(bounds[r], bounds[r - 1] + 1), P - (world - r))
    bounds = np.maximum.accumulate(bounds)
    return [(int(bounds[r]), int(bounds[r + 1])) for r in range(world)]


class Comm:
    def __init__(self, group=None, shard_sizes=None):
        assert dist.is_initialized(), "torch.distributed is not initialised"
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.shard_sizes = list(shard_sizes) if shard_sizes is not None else None

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

    def all_gather_shards(self, t, width=1):
        """Concatenate per-gene vectors (width = values per gene) of the (possibly unequal)
        gene shards."""
        sizes = self.shard_sizes
        assert sizes is not None and t.numel() == sizes[self.rank] * width
        if len(set(sizes)) == 1:
            return self.all_gather_cat(t)
        m = max(sizes) * width
        pad = torch.zeros(m, dtype=t.dtype, device=t.device)
        pad[: t.numel()] = t.reshape(-1)
        g = self.all_gather_cat(pad).view(self.world, m)
        return torch.cat([g[r, : sizes[r] * width] for r in range(self.world)])

    def all_gather_slots(self, buf, slot):
        """In-place all-gather: buf holds `world` slots of `slot` elements and this rank's is
        filled in; afterwards every slot holds its rank's data."""
        assert buf.numel() >= self.world * slot and buf.is_contiguous()
        out = buf[: self.world * slot]
        dist.all_gather_into_tensor(out, out[self.rank * slot: (self.rank + 1) * slot],
                                    group=self.group)
        return buf


class PeerRecords:
    """The (score, index) records of all ranks, exchanged through NVLink peer memory instead
    of an NCCL all-gather (csrc/peer.cu): every rank owns a buffer uint64[2][W][4] that all its
    peers map through CUDA IPC (torch's own handle exchange) and write into directly.
    `available(comm)`: NCCL backend, one node, PRB_PEER != 0."""

    @staticmethod
    def available(comm):
        import os

        return (comm is not None and comm.world > 1 and dist.get_backend(comm.group) == "nccl"
                and os.environ.get("PRB_PEER", "1") != "0" and torch.cuda.is_available())

    def __init__(self, comm):
        W, dev = comm.world, torch.device("cuda", torch.cuda.current_device())
        self.comm = comm
        # an allocation of its own (not a slice of a cached block other tensors share)
        self.local = torch.zeros(2 * W * 4, dtype=torch.int64, device=dev)
        handle = self.local.untyped_storage()._share_cuda_()
        handles = [None] * W
        dist.all_gather_object(handles, handle, group=comm.group)
        self._peer_storages, ptrs = [], []
        for r in range(W):
            if r == comm.rank:
                ptrs.append(self.local.data_ptr())
                continue
            h = handles[r]
            st = torch.UntypedStorage._new_shared_cuda(dev.index, *h[1:])
            self._peer_storages.append(st)
            ptrs.append(st.data_ptr())
        self.peer_ptrs = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.xtag = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group=comm.group)  # every buffer is mapped and zeroed before the first store

    def check(self):
        if int(self.err.item()):
            raise RuntimeError("a peer rank did not deliver its candidate record within 4 s")


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
