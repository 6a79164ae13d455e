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
