#!/usr/bin/env python
"""Golden gene sequences at the FULL sizes of BASELINE.json's configurations, from the CPU oracle.

    python tests/golden/make_golden_large.py [case ...]      # default: every case

Runs on host cores only (no GPU, no /root/reference): the oracle (oracle/oracle.c, pinned
against the unmodified reference by tests/golden/make_golden.py + tests/test_oracle_golden.py
and, live, by tests/test_oracle_vs_reference.py) performs the reference's greedy loop
(picturedrocks/markers/_iterative.py:53-66 -> mutualinformation/iterative.py:48-59, 76-87)
with its per-gene loop sharded over OpenMP threads — "reference kernel x ncores harness"
in SURVEY.md §8(d)'s words, not the reference's own speed.  Inputs are bench.py's synthetic
workloads (same generators as the GPU side: picturedrocks_b200/synth.py, restated in C in
oracle/synth_cpu.c for the large ones).

Cases (bench.py WORKLOADS):
  c4      1,306,127 x 27,998, 60 classes : CIFE, 72 steps   <- the north-star acceptance test
  c3         68,579 x 32,738, 10 classes : CIFE, 60 steps
  c2          3,005 x 19,972,  9 classes : makeinfoset (k=5), MIM top-50 + JMI top-50
  c1          2,730 x  3,451, 19 classes : makeinfoset (k=5), CIFE top-20  (also in config1.npz)
  c5_k5   1,000,000 x 20,000, 20 classes : JMI, 110 steps
  c5_k10    100,000 x 50,000, 20 classes, 10 bins : JMI, 110 steps
  c5_k20    100,000 x 50,000, 20 classes, 20 bins : JMI, 110 steps
Each pack holds the sequence S, the final score / penalty / base vectors, the winner's
score and its margin over the runner-up at every step (how close the nearest tie was).
"""
import json
import os
import sys
import time
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from bench import WORKLOADS  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "c4": [("CIFE", 72)],
    "c3": [("CIFE", 60)],
    "c2": [("MIM", 50), ("JMI", 50)],
    "c1": [("CIFE", 20)],
    "c5_k5": [("JMI", 110)],
    "c5_k10": [("JMI", 110)],
    "c5_k20": [("JMI", 110)],
}


def make_infoset(w, threads):
    if w.get("dense"):
        import pandas as pd

        from picturedrocks_b200.synth import poisson_gamma

        X, y = poisson_gamma(w["N"], w["P"], w["C"], w["seed"])
        adata = types.SimpleNamespace(X=X, obs=pd.DataFrame({"y": y}))
        return O.makeinfoset(adata, True, k=w["K"], nthreads=threads), None
    from picturedrocks_b200.synth import DiscretisedSpec

    spec = DiscretisedSpec(N=w["N"], P=w["P"], C=w["C"], K=w["K"], mean_density=w["mean_density"],
                           seed=w["seed"])
    indptr, rows, codes, y = O.synth_csc(spec, nthreads=threads)
    return O.FlatSparseInformationSet(indptr, rows, codes, spec.N, y, nthreads=threads), int(indptr[-1])


def greedy(fs, n):
    """BASE:64-66 with the winner's score and its margin recorded."""
    best, margin = [], []
    for _ in range(n):
        j = int(np.argmax(fs.score))
        top2 = np.partition(fs.score, -2)[-2:]
        best.append(float(top2[1]))
        margin.append(float(top2[1] - top2[0]))
        fs.add(j)
    return np.array(best), np.array(margin)


def run_case(name, threads):
    w = WORKLOADS[name]
    t0 = time.time()
    inf, nnz = make_infoset(w, threads)
    t_gen = time.time() - t0
    pack = {}
    meta = {k: w[k] for k in ("N", "P", "C", "K", "seed") if k in w}
    meta.update(mean_density=w.get("mean_density"), dense=bool(w.get("dense")), nnz=nnz,
                workload=w["name"], threads=threads, numpy=np.__version__,
                how="CPU oracle (oracle/oracle.c), per-gene loop over %d OpenMP threads" % threads)
    for kind, n in CASES[name]:
        t1 = time.time()
        fs = getattr(O, kind)(inf)
        pack[kind + "_base"] = np.array(fs.base_score)
        if kind in ("MIM", "UniEntropy"):
            fs.autoselect(n)
        else:
            pack[kind + "_best"], pack[kind + "_margin"] = greedy(fs, n)
        pack[kind + "_S"] = np.array([int(s) for s in fs.S], dtype=np.int64)
        pack[kind + "_score"] = np.array(fs.score)
        pack[kind + "_penalty"] = np.array(fs.penalty)
        meta[kind + "_seconds"] = round(time.time() - t1, 2)
        print("%s %s: %d steps in %.1f s (input %.1f s); S[:8] = %s; smallest margin %.3g"
              % (name, kind, n, time.time() - t1, t_gen, list(pack[kind + "_S"][:8]),
                 pack.get(kind + "_margin", np.array([np.nan])).min()), flush=True)
    pack["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(OUT, "large_%s.npz" % name), **pack)


def main():
    names = sys.argv[1:] or list(CASES)
    threads = O.host_cores()
    for name in names:
        run_case(name, threads)


if __name__ == "__main__":
    main()
