"""Latency of the interactive protocol (reference:
picturedrocks/markers/interactive.py:142-247): what the widget does per click —
`featsel.add(ind)` or `featsel.remove(ind)` followed by a read of `featsel.score` on the host
(np.argsort / boolean masks) and of `featsel.S`.  One GPU, resident matrix.

    python profiles/interactive_latency.py [workload]      # default c4

Prints one JSON line: median / max wall-clock milliseconds per call, host-synchronous
(the call returns a numpy score vector, so device work and the D2H copy are included).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import picturedrocks_b200 as pr  # noqa: E402
from bench import WORKLOADS, make_spec  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c4"
    w = WORKLOADS[name]
    eng, yd = make_spec(w).device_engine(0, w["P"])
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    out = {"workload": w["name"], "unit": "ms per call (host wall clock, incl. score D2H)"}
    for kind in ("CIFE", "JMI"):
        fs = getattr(pr.markers, kind)(inf)
        fs.autoselect(3)                      # a few genes selected, buffers warm
        order = np.argsort(fs.score)[::-1]
        picks = [int(g) for g in order[:12]]
        t_add, t_rem = [], []
        for g in picks:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fs.add(g)
            s = fs.score                      # numpy float64[P], -inf for selected genes
            t_add.append(1e3 * (time.perf_counter() - t0))
            assert s[g] == -np.inf and fs.S[-1] == g
        for g in picks[::-1]:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fs.remove(g)
            s = fs.score
            t_rem.append(1e3 * (time.perf_counter() - t0))
            assert np.isfinite(s[g]) and g not in fs.S
        out[kind] = {"add_median": float(np.median(t_add)), "add_max": float(np.max(t_add)),
                     "remove_median": float(np.median(t_rem)), "remove_max": float(np.max(t_rem)),
                     "calls": len(picks)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
