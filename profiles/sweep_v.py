"""Tuning sweep of the V-layout plan knobs on one GPU (run on the B200 box through gpurun).

    python profiles/sweep_v.py c4_shard8 "" "TILE_CELLS=62300" "TILE_CELLS=62300,HELP_MIN=64" ...

One resident matrix, one process: for every knob set (PRB_V_<NAME> environment overrides read
by engine._build_v / prb_stream_v) the V layout is rebuilt, a fresh CIFE selector runs
3 warm-up + 40 timed greedy steps (CUDA graph), then 10 eager steps with CUDA events around
every k_stream_v launch.  Prints one line per knob set.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import picturedrocks_b200 as pr  # noqa: E402
import picturedrocks_b200.markers.mutualinformation.iterative as it  # noqa: E402
from bench import WORKLOADS, make_spec  # noqa: E402


def main():
    # "c4" or "c4:P=14000" (override fields of the named workload)
    name, _, over = sys.argv[1].partition(":")
    w = dict(WORKLOADS[name])
    for kv in filter(None, over.split(",")):
        k, v = kv.split("=")
        w[k] = type(w[k])(v)
    sets = sys.argv[2:] or [""]
    it.GRAPH_MIN_STEPS = 2
    eng, yd = make_spec(w).device_engine(0, w["P"])
    ref_S = None
    for knobs in sets:
        for k in [k for k in os.environ if k.startswith("PRB_V_") or k.startswith("PRB_ELL_")]:
            del os.environ[k]
        for kv in filter(None, knobs.split(",")):
            k, v = kv.split("=")
            os.environ[("PRB_" if k.startswith("ELL_") else "PRB_V_") + k] = v
        eng.v = None
        eng.epoch += 1
        eng.ell2 = eng.ell and os.environ.get("PRB_ELL_V", "2") == "2"
        eng.vmajor = (eng.ell2 and eng.K1 <= 4
                      and os.environ.get("PRB_ELL_VMAJOR", "1") != "0")
        eng._hits = None
        inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
        fs = pr.markers.CIFE(inf)
        fs.autoselect(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fs.autoselect(40)
        e1.record()
        torch.cuda.synchronize()
        step_ms = e0.elapsed_time(e1) / 40
        eng.profile_events = []
        fs.autoselect(10)
        torch.cuda.synchronize()
        k_ms = float(np.mean([a.elapsed_time(b) for a, b in eng.profile_events]))
        eng.profile_events = None
        S = fs.S
        if ref_S is None:
            ref_S = S
        v = eng.v
        print("%-44s step %.4f ms  kernel %.4f ms  rows %d items %d R %d cap %s tiles %d segs %d %s"
              % (knobs or "(default)", step_ms, k_ms, v["rows"], v["n_items"], v.get("item_rows", 0),
                 v.get("cap"), v["plan"]["n_st"], v["plan"]["n_seg"],
                 "same S" if S == ref_S else "S DIFFERS"),
              flush=True)
        del fs, inf


if __name__ == "__main__":
    main()
