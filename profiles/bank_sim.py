"""CPU simulation of the shared-memory bank conflicts of k_stream_ell2's probes (no GPU needed).

A data row of the streamed lane-run layout holds, per lane, two 16-bit entries = byte addresses
in the 64 KB tile window; the kernel probes them with two LDS.U8 gathers per row.  A gather costs
as many wavefronts as the fullest bank has DISTINCT 4-byte words (bank = (addr >> 2) & 31).  The
order of the entries inside a lane-run is free (counts do not depend on it), so the layout builder
chooses it.  This script compares, on synthetic bundles shaped like config 4's (cells uniform in
a class segment, lane-run lengths from a gamma distribution, 32 lane-runs of similar length per
bundle):

  ascending   the CSC order (cells ascending)                      = no builder
  rotation    what k_ell_build_banked ships: per lane, chunks of 512 entries counting-sorted by
              bank, step s of lane l takes bank (l + s) mod 32, else the next non-empty bank
  matching    a per-step assignment ACROSS the lanes of the bundle: lanes with the fewest
              candidate banks choose first, each takes a bank no other lane has taken in this
              step (the one it holds most entries of), else its fullest bank
  bound       1.0 (every gather conflict-free)

Usage: python profiles/bank_sim.py [bundles] > profiles/rNN_bank_sim.txt"""
import sys

import numpy as np


def wavefronts(addr_rows):
    """addr_rows: int[steps][32] byte addresses (-1 = padding, probes one common slot)."""
    total = 0
    for row in addr_rows:
        words = {}
        for a in row:
            a = 32767 if a < 0 else int(a)
            words.setdefault((a >> 2) & 31, set()).add(a >> 2)
        total += max(len(w) for w in words.values())
    return total


def order_ascending(runs, steps):
    out = -np.ones((steps, 32), dtype=np.int64)
    for l, r in enumerate(runs):
        out[: len(r), l] = np.sort(r)
    return out


def order_rotation(runs, steps, chunk=512):
    out = -np.ones((steps, 32), dtype=np.int64)
    for l, r in enumerate(runs):
        r = np.sort(r)
        for c0 in range(0, len(r), chunk):
            part = r[c0: c0 + chunk]
            queues = [list(part[((part >> 2) & 31) == q]) for q in range(32)]
            for s in range(len(part)):
                q0 = (l + c0 + s) & 31
                q = next((q0 + d) & 31 for d in range(32) if queues[(q0 + d) & 31])
                out[c0 + s, l] = queues[q].pop(0)
    return out


def order_matching(runs, steps):
    out = -np.ones((steps, 32), dtype=np.int64)
    queues = [[list(np.sort(r[((r >> 2) & 31) == q])) for q in range(32)] for r in runs]
    left = [len(r) for r in runs]
    for s in range(steps):
        lanes = [l for l in range(32) if left[l] > 0]
        # lanes with few non-empty banks choose first
        lanes.sort(key=lambda l: sum(1 for q in queues[l] if q))
        taken = set()
        for l in lanes:
            cand = [q for q in range(32) if queues[l][q]]
            free = [q for q in cand if q not in taken]
            # keep the lane's queues level: take from its fullest bank
            q = max(free or cand, key=lambda q: len(queues[l][q]))
            taken.add(q)
            out[s, l] = queues[l][q].pop(0)
            left[l] -= 1
    return out


def synth_bundle(rng, mean_len, seg_cells=21800, n_segments=2):
    """32 lane-runs of similar length; every lane-run lives in one class segment of the tile."""
    base_len = max(2, int(rng.gamma(0.7, mean_len / 0.7)))
    base_len = min(base_len, 1024)
    runs = []
    for _ in range(32):
        n = max(1, min(1024, int(base_len * rng.uniform(0.9, 1.0))))
        seg0 = int(rng.integers(0, n_segments)) * (32768 - seg_cells)
        runs.append(seg0 + rng.choice(seg_cells, size=n, replace=False))
    return runs


def main():
    n_bundles = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(0)
    print("# wavefronts per LDS.U8 gather (1.0 = conflict-free); %d synthetic bundles per line" % n_bundles)
    print("# %-22s %10s %10s %10s" % ("mean lane-run length", "ascending", "rotation", "matching"))
    for mean_len in (16, 32, 64, 128, 512):
        tot = {"a": 0, "r": 0, "m": 0}
        gathers = 0
        for _ in range(n_bundles):
            runs = synth_bundle(rng, mean_len)
            steps = max(len(r) for r in runs)
            steps += steps & 1
            gathers += steps
            tot["a"] += wavefronts(order_ascending(runs, steps))
            tot["r"] += wavefronts(order_rotation(runs, steps))
            tot["m"] += wavefronts(order_matching(runs, steps))
        print("  %-22d %10.2f %10.2f %10.2f" % (mean_len, tot["a"] / gathers, tot["r"] / gathers,
                                                tot["m"] / gathers))


if __name__ == "__main__":
    main()
