"""Per-item start / end times of one k_stream_ell2 launch (run on the B200 box through gpurun).

    python profiles/ell2_timing.py c4 [KNOB=V,...]

Builds the workload, runs 3 greedy CIFE steps, then one step with the timing buffer of
prb_stream_ell2 attached.  Prints per CTA the rows / bundles of its chunk and when its last
item ended, and a fit of the item durations  t = a * rows + b * bundles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import picturedrocks_b200 as pr  # noqa: E402
import picturedrocks_b200.markers.mutualinformation.iterative as it  # noqa: E402
from bench import WORKLOADS, make_spec  # noqa: E402


def main():
    w = dict(WORKLOADS[sys.argv[1]])
    for kv in filter(None, (sys.argv[2] if len(sys.argv) > 2 else "").split(",")):
        k, v = kv.split("=")
        os.environ[("PRB_" if k.startswith("ELL_") else "PRB_V_") + k] = v
    it.GRAPH_MIN_STEPS = 1 << 30  # eager steps
    eng, yd = make_spec(w).device_engine(0, w["P"])
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    fs = pr.markers.CIFE(inf)
    fs.autoselect(3)
    v = eng.v
    n_ctas, n_items = v["n_ctas"], v["n_items"]
    tm = torch.zeros(n_items * 2, dtype=torch.int64, device=eng.device)
    eng.ell_timing = tm
    fs.autoselect(1)
    torch.cuda.synchronize()
    eng.ell_timing = None
    t = tm.cpu().numpy().reshape(n_items, 2)
    t0 = t[:, 0].min()
    start, end = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3
    item_row0 = v["item_row0"].cpu().numpy().astype(np.int64)
    phase = v["phase"].cpu().numpy().reshape(-1, 4)
    cta_ph0 = v["cta_ph0"].cpu().numpy()
    # bundles per item: header rows carry bit 31 patterns; count them from the stream lengths
    ell = v["ell"]
    # a header row is where a bundle starts: recover the bundle starts by walking the stream
    # on the host would be slow; use the planner's srow0 if it was kept
    srow0 = v.get("srow0")
    rows = np.diff(item_row0)
    dur = end - start
    if srow0 is not None:
        srow0 = srow0.cpu().numpy().astype(np.int64)
        nb = np.diff(np.searchsorted(srow0, item_row0, side="left"))
    else:
        nb = np.zeros_like(rows)
    if os.environ.get("PRB_TIMING_DUMP"):
        np.savez_compressed(os.environ["PRB_TIMING_DUMP"], start=start, end=end, rows=rows, nb=nb,
                            phase=phase, cta_ph0=cta_ph0, item_row0=item_row0,
                            cell0=np.asarray(v["plan"]["cell0"]), st=np.asarray(v["plan"]["st"]))
    print("launch: %.1f us; items %d" % (end.max(), n_items))
    A = np.stack([rows, nb], 1).astype(np.float64)
    coef, *_ = np.linalg.lstsq(A, dur, rcond=None)
    print("fit  item us = %.5f * rows + %.5f * bundles   (=> one bundle costs %.1f rows)"
          % (coef[0], coef[1], coef[1] / coef[0]))
    print("cta tile0 tiles items    rows bundles  rows/bundle   end_us")
    ends = []
    for c in range(n_ctas):
        a, b = cta_ph0[c], cta_ph0[c + 1]
        if b <= a:
            continue
        i0, i1 = phase[a, 2], phase[b - 1, 2] + phase[b - 1, 3]
        ends.append((end[i0:i1].max(), c, phase[a, 0], (phase[a:b, 1] & 255).sum(), i1 - i0,
                     rows[i0:i1].sum(), nb[i0:i1].sum()))
    ends.sort()
    for e, c, t0_, nt, ni, r, nbb in ends[:8] + ends[-16:]:
        print("%3d %5d %5d %5d %7d %7d %12.1f %8.1f" % (c, t0_, nt, ni, r, nbb, r / max(nbb, 1), e))
    # slowest CTA: its items in order of position
    e, c = ends[-1][0], ends[-1][1]
    a, b = cta_ph0[c], cta_ph0[c + 1]
    i0, i1 = phase[a, 2], phase[b - 1, 2] + phase[b - 1, 3]
    print("items of the slowest CTA %d (every 16th): idx rows bundles start dur us/row" % c)
    for i in range(i0, i1, 16):
        print("  %5d %6d %6d %8.1f %8.1f %8.4f" % (i - i0, rows[i], nb[i], start[i], dur[i],
                                                  dur[i] / max(rows[i], 1)))


if __name__ == "__main__":
    main()
