"""Gene-sharded selection over NCCL (needs >= 2 GPUs; skipped on a 1-GPU box)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        import picturedrocks_b200 as pr
        from picturedrocks_b200.dist import Comm, partition_genes_weighted
        from picturedrocks_b200.synth import DiscretisedSpec

        spec = DiscretisedSpec(N=40000, P=1203, C=9, K=5, mean_density=0.08, seed=21)
        parts = partition_genes_weighted(spec.device_gene_nnz(), world)  # unequal gene counts
        comm = Comm(shard_sizes=[b - a for a, b in parts])
        eng, yd = spec.device_engine(parts[rank][0], parts[rank][1], comm=comm)
        inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
        out = {}
        for kind in ("CIFE", "JMI", "CIFEUnsup"):
            fs = getattr(pr.markers, kind)(inf)
            fs.autoselect(12)  # >= GRAPH_MIN_STEPS: the step (incl. NCCL) runs as a CUDA graph
            fs.remove(fs.S[2])
            for g in (7, 19, 640, 3, 1100):  # consecutive steps that send a record, pick none
                fs.add(g)
            fs.remove(19)
            fs.autoselect(2)
            out[kind] = (list(fs.S), fs.score)
        fs = pr.markers.CIFE(inf)  # add / remove BEFORE the first pick
        fs.add(5)
        fs.add(900)
        fs.remove(5)
        fs.autoselect(3)
        out["add_first"] = (list(fs.S), fs.score)
        out["wrt"] = inf.entropy_wrt([-1, 3])
        out["wrt2"] = inf.entropy_wrt([5, 700])
        out["ent"] = inf.entropy([-1, 700])
        if rank == 0:
            ret.update(out)
    finally:
        from picturedrocks_b200.dist import shutdown

        shutdown()


@pytest.mark.timeout(600)
def test_two_rank_selection_equals_single_gpu():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import picturedrocks_b200 as pr
    from picturedrocks_b200.synth import DiscretisedSpec

    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    spec = DiscretisedSpec(N=40000, P=1203, C=9, K=5, mean_density=0.08, seed=21)
    eng, yd = spec.device_engine()
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    for kind in ("CIFE", "JMI", "CIFEUnsup"):
        fs = getattr(pr.markers, kind)(inf)
        fs.autoselect(12)
        fs.remove(fs.S[2])
        for g in (7, 19, 640, 3, 1100):
            fs.add(g)
        fs.remove(19)
        fs.autoselect(2)
        assert ret[kind][0] == list(fs.S), kind
        assert np.array_equal(ret[kind][1], fs.score), kind
    fs = pr.markers.CIFE(inf)
    fs.add(5)
    fs.add(900)
    fs.remove(5)
    fs.autoselect(3)
    assert ret["add_first"][0] == list(fs.S)
    assert np.array_equal(ret["add_first"][1], fs.score)
    assert np.array_equal(ret["wrt"], inf.entropy_wrt([-1, 3]))
    assert np.array_equal(ret["wrt2"], inf.entropy_wrt([5, 700]))
    assert ret["ent"] == inf.entropy([-1, 700])
