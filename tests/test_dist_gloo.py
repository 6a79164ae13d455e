"""world_size-2 gloo tests (CPU) of the multi-GPU exchange protocol in
picturedrocks_b200/dist.py: gene partition, owner-scatter + max-all-reduce of the selected
gene's dense code vector, all-gather of (score, index) candidates with the np.argmax tie
rule.  The per-shard entropies come from the CPU oracle (allowed in tests); the greedy
sequence of the sharded run must equal the single-process oracle sequence."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from picturedrocks_b200.dist import Comm, partition_genes


def test_partition_genes():
    for P, w in ((10, 3), (27998, 8), (5, 8), (16, 2)):
        parts = partition_genes(P, w)
        assert parts[0][0] == 0 and parts[-1][1] == P
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pick(scores, idxs):
    """max score, NaN maximal, lowest index on ties (np.argmax over the concatenation)."""
    best = None
    for s, i in zip(scores, idxs):
        if i < 0:
            continue
        if best is None:
            best = (s, i)
            continue
        bs, bi = best
        if np.isnan(s) or np.isnan(bs):
            better = (np.isnan(s) and not np.isnan(bs)) or (np.isnan(s) and np.isnan(bs) and i < bi)
        else:
            better = s > bs or (s == bs and i < bi)
        if better:
            best = (s, i)
    return best


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from picturedrocks_b200.synth import DiscretisedSpec

        spec = DiscretisedSpec(N=1500, P=91, C=5, K=5, mean_density=0.2, seed=11)
        y = spec.cpu_labels()
        parts = partition_genes(spec.P, world)
        lo, hi = parts[rank]
        comm = Comm(shard_sizes=[b - a for a, b in parts])
        assert comm.max_int(rank + 3) == world + 2
        Xl = spec.cpu_csc(range(lo, hi), y)  # this rank's gene shard only
        Xl.data = Xl.data.astype(np.int64)
        local = O.SparseInformationSet(Xl, y)
        # replicated lookup vectors via the shard all-gather
        Hx = comm.all_gather_shards(torch.from_numpy(local.entropy_wrt(np.arange(0)))).numpy()
        Hyx = comm.all_gather_shards(torch.from_numpy(local.entropy_wrt(np.array([-1])))).numpy()
        Hy = local.entropy(np.array([-1]))
        base = (Hx[lo:hi] + Hy) - Hyx[lo:hi]
        score, penalty, S = base.copy(), np.zeros(hi - lo), []
        for _ in range(6):
            j_loc = int(np.argmax(score)) if hi > lo else -1
            cs = comm.all_gather_cat(torch.tensor([score[j_loc] if j_loc >= 0 else 0.0],
                                                  dtype=torch.float64)).numpy()
            ci = comm.all_gather_cat(torch.tensor([lo + j_loc if j_loc >= 0 else -1],
                                                  dtype=torch.int64)).numpy()
            _, j = _pick(cs, ci)
            S.append(int(j))
            # owner scatters the dense code vector, everyone max-reduces it
            xj = torch.zeros(spec.N, dtype=torch.uint8)
            if lo <= j < hi:
                col = Xl.getcol(j - lo)
                xj[torch.from_numpy(col.indices.astype(np.int64))] = torch.from_numpy(
                    col.data.astype(np.uint8))
            comm.allreduce_max_(xj)
            assert np.array_equal(xj.numpy(), spec.cpu_codes(j, y))
            # entropies of the local shard w.r.t. the replicated x_j
            nzr = np.flatnonzero(xj.numpy()).astype(np.int64)
            vals = xj.numpy()[nzr].astype(np.int64)
            Hij = O.sparse_entropy_wrt(Xl.indices, Xl.indptr, Xl.data, nzr, vals, spec.N,
                                       hi - lo, 1, local._shift, 0, np.arange(0), np.arange(0))
            Hyij = O.sparse_entropy_wrt(Xl.indices, Xl.indptr, Xl.data, nzr, vals, spec.N,
                                        hi - lo, 1, local._shift, local._ybits, y,
                                        local.classsizes)
            d = (((base + Hx[j]) - Hij) - Hyx[j]) + Hyij
            penalty += d
            score -= d
            if lo <= j < hi:
                score[j - lo] = -np.inf
        full = comm.all_gather_shards(torch.from_numpy(score)).numpy()
        if rank == 0:
            ret["S"] = S
            ret["score"] = full
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_protocol_matches_single_process():
    from oracle import oracle as O
    from picturedrocks_b200.synth import DiscretisedSpec

    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    spec = DiscretisedSpec(N=1500, P=91, C=5, K=5, mean_density=0.2, seed=11)
    y = spec.cpu_labels()
    X = spec.cpu_csc(range(spec.P), y)
    ref = O.CIFE(O.SparseInformationSet(X, y))
    ref.autoselect(6)
    assert ret["S"] == [int(s) for s in ref.S]
    assert np.array_equal(ret["score"], ref.score)


def test_partition_genes_weighted():
    from picturedrocks_b200.dist import partition_genes_weighted

    rng = np.random.RandomState(0)
    w = rng.gamma(0.5, 100.0, 2801).astype(np.int64)
    for world in (1, 2, 4, 8):
        parts = partition_genes_weighted(w, world)
        assert parts[0][0] == 0 and parts[-1][1] == w.size
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert all(hi > lo for lo, hi in parts)
        tot = [w[lo:hi].sum() for lo, hi in parts]
        assert max(tot) <= np.mean(tot) + w.max()  # within one gene of the mean
    # fewer genes than ranks, all-zero weights
    parts = partition_genes_weighted(np.zeros(10), 3)
    assert parts[0][0] == 0 and parts[-1][1] == 10 and all(hi > lo for lo, hi in parts)


def _fold_worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from picturedrocks_b200.data import ExpressionData
        from picturedrocks_b200.performance import FoldTester

        rng = np.random.RandomState(5)
        X = rng.poisson(1.0, size=(60, 12)).astype(np.float64)
        y = rng.randint(0, 3, 60)
        import pandas as pd

        ad = ExpressionData(X, obs=pd.DataFrame({"y": y}))
        ft = FoldTester(ad)
        ft.makefolds(k=5, random=False)
        calls = []

        def select(train):  # a host stand-in for mi_selector: the 3 genes with the largest mean
            calls.append(train.n_obs)
            return [int(g) for g in np.argsort(-np.asarray(train.X).mean(0), kind="stable")[:3]]

        ft.selectmarkers(select, comm=Comm())
        assert len(calls) == len(range(rank, 5, world))  # this rank ran only its folds
        if rank == 0:
            ret["markers"] = [list(m) for m in ft.markers]
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_folds_distributed_over_ranks():
    """FoldTester.selectmarkers(select, comm): rank r runs the folds r, r + world, ...; every
    rank ends up with all marker lists, equal to the single-process result."""
    from picturedrocks_b200.data import ExpressionData
    from picturedrocks_b200.performance import FoldTester

    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fold_worker, args=(world, port, ret), nprocs=world, join=True)
    rng = np.random.RandomState(5)
    X = rng.poisson(1.0, size=(60, 12)).astype(np.float64)
    y = rng.randint(0, 3, 60)
    import pandas as pd

    ft = FoldTester(ExpressionData(X, obs=pd.DataFrame({"y": y})))
    ft.makefolds(k=5, random=False)
    ft.selectmarkers(lambda tr: [int(g) for g in np.argsort(-np.asarray(tr.X).mean(0), kind="stable")[:3]])
    assert ret["markers"] == [list(m) for m in ft.markers]
