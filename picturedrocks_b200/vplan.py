"""Host-side planning for the V layout (csrc/vstream.cu): pure functions, no CUDA.

* `segment_plan`  — cells (relabelled so that classes are contiguous) are cut into
  SEGMENTS, a class or a piece of a class that fits the shared-memory tile, and consecutive
  segments are packed into SUPER-TILES (what one CTA stages in shared memory at a time).
* `item_cuts`     — the rows of a super-tile's block are cut into work items: large ones
  first, smaller ones towards the end, so that the tail of a launch is one small item long.
"""
import numpy as np


def segment_plan(class_sizes, cap, max_segs):
    """class_sizes: cells per class in label order (one entry = N without labels); cap: cells a
    super-tile may span; max_segs: segments per super-tile.
    Returns (cell0, seg_class, st): cell0[n_seg+1] ascending segment boundaries (last = N),
    seg_class[n_seg], st[n_st+1] with super-tile t = segments [st[t], st[t+1])."""
    cap, max_segs = int(cap), int(max_segs)
    assert cap >= 1 and max_segs >= 1
    cell0, seg_class, pos = [], [], 0
    for c, n in enumerate(int(v) for v in class_sizes):
        pieces = -(-n // cap)
        for k in range(pieces):
            cell0.append(pos + n * k // pieces)
            seg_class.append(c)
        pos += n
    N = pos
    cell0.append(N)
    n_seg = len(seg_class)
    n_st_min = max(1, -(-N // cap))
    target = -(-N // n_st_min)  # aim at equally filled super-tiles
    st, cur_cells, cur_n = [0], 0, 0
    for i in range(n_seg):
        sz = cell0[i + 1] - cell0[i]
        if cur_n and (cur_cells + sz > cap or cur_n == max_segs or cur_cells >= target):
            st.append(i)
            cur_cells, cur_n = 0, 0
        cur_cells += sz
        cur_n += 1
    st.append(n_seg)
    return cell0, seg_class, st


def item_rows(total_rows, warps, forced=0):
    """Rows of the large work items, a multiple of 32.  Every item costs a pipeline restart
    (favours few, large items), while large items spread the concurrently streamed ranges
    over more memory and lengthen the tail (favours many, small ones).  Measured optimum on
    B200 (profiles/r01_item_rows_sweep.txt; rows per warp -> rows per item: 481 -> 64,
    2 254 -> 384..640, 4 652 -> 768, 9 198 -> 1 024, 18 159 -> 1 536):
    R = min(4.7 * rpw^0.59, rpw / 6)."""
    if forced:
        return (int(forced) + 31) & ~31
    rpw = max(int(total_rows), 0) / max(int(warps), 1)
    R = int(min(4.7 * rpw ** 0.59, rpw / 6.0))
    return max(32, (R + 16) & ~31)


def item_cuts(lo, hi, R):
    """First rows of the work items of the block [lo, hi): the first 3/4 of the block in items
    of R rows, then up to 92 % in items of R/2, the rest in items of R/4."""
    parts, pos = [], lo
    for frac, size in ((0.75, R), (0.92, max(min(R, 256), R // 2)), (1.0, max(min(R, 128), R // 4))):
        stop = lo + int((hi - lo) * frac) if frac < 1.0 else hi
        if pos < stop:
            parts.append(np.arange(pos, stop, size, dtype=np.int64))
            pos = int(parts[-1][-1]) + size
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.int64)



ELL2_SPLIT = 1 << 30    # header bit: the lane-run is a piece of a longer run
ELL2_TILE = 1 << 15     # positions per super-tile in the padded cell space


def ell_plan(lens, run_start, n_seg, st, cap, seg_id=None, id_shift=22):
    """Index plan of the lane-run layout (csrc/vell.cu), torch tensors on the device of `lens`.
    lens int64[n_runs]: entries of run (vg, s) at index vg * n_seg + s; run_start int64[n_runs]:
    its first entry in the flat run-major scratch; st: super-tile t = segments [st[t], st[t+1]);
    cap: entries per lane-run (even).
    Every run is cut into ceil(len / cap) lane-runs of nearly equal length; the lane-runs of a
    super-tile are sorted by length (descending) and grouped 32 at a time into bundles of
    ceil(longest / 2) rows.  Returns a dict: n_bundles; brow0 int64
    [n_bundles+1] first row of every bundle; nb_t / tile_b0 int64[n_st] bundles per tile and
    the first one; lr_src int64 / lr_len int32 / lr_hdr int32 [32 * n_bundles] per lane: first
    entry in the scratch, length, header word (segment-in-tile << 24 | vg; -1 = empty lane).
    seg_id (int64[n_seg], the streamed layout of csrc/vell2.cu): header word = seg_id of the
    segment << id_shift | vg, bit 30 = piece of a run that was cut, all ones below id_shift =
    empty lane (bit 31 is filled in by prb_ell_build)."""
    import torch

    dev = lens.device
    i64 = dict(dtype=torch.int64, device=dev)
    n_st = len(st) - 1
    n_runs = lens.numel()
    nsub = (lens + (cap - 1)) // cap
    n_lr = int(nsub.sum().item()) if n_runs else 0
    first = torch.cumsum(nsub, 0) - nsub
    run_of = torch.repeat_interleave(torch.arange(n_runs, **i64), nsub)
    sub = torch.arange(n_lr, **i64) - first[run_of]
    ln, ns = lens[run_of], nsub[run_of]
    base = ln // torch.clamp(ns, min=1)
    rem = ln - base * ns
    sublen = base + (sub < rem).to(torch.int64)
    src = run_start[run_of] + sub * base + torch.minimum(sub, rem)
    seg = run_of % n_seg
    st_t = torch.tensor(list(st), **i64)
    seg_tile = torch.repeat_interleave(torch.arange(n_st, **i64), st_t[1:] - st_t[:-1])
    tile = seg_tile[seg]
    if seg_id is None:
        hdr = ((seg - st_t[tile]) << 24) | (run_of // n_seg)
    else:
        hdr = (seg_id[seg] << id_shift) | (run_of // n_seg) | ((ns > 1).to(torch.int64) * ELL2_SPLIT)
    del run_of, sub, ln, ns, base, rem, seg
    order = torch.argsort((tile << 32) | ((1 << 32) - 1 - sublen))
    tile_s = tile[order]
    cnt_t = torch.bincount(tile, minlength=n_st)
    nb_t = (cnt_t + 31) // 32
    tile_b0 = torch.cumsum(nb_t, 0) - nb_t
    n_bundles = int(nb_t.sum().item())
    lr0_t = torch.cumsum(cnt_t, 0) - cnt_t
    slot = tile_b0[tile_s] * 32 + (torch.arange(n_lr, **i64) - lr0_t[tile_s])
    n_slots = max(n_bundles, 1) * 32
    lr_src = torch.zeros(n_slots, **i64)
    lr_len = torch.zeros(n_slots, dtype=torch.int32, device=dev)
    lr_hdr = torch.full((n_slots,), -1 if seg_id is None else (1 << id_shift) - 1,
                        dtype=torch.int32, device=dev)
    lr_src[slot] = src[order]
    lr_len[slot] = sublen[order].to(torch.int32)
    lr_hdr[slot] = hdr[order].to(torch.int32)
    L = (lr_len.view(-1, 32)[:n_bundles, 0].to(torch.int64) + 1) // 2  # lane 0 is the longest
    brow0 = torch.zeros(n_bundles + 1, **i64)
    torch.cumsum(L, 0, out=brow0[1:])
    return dict(n_bundles=n_bundles, brow0=brow0, nb_t=nb_t, tile_b0=tile_b0, lr_src=lr_src,
                lr_len=lr_len, lr_hdr=lr_hdr)


def ell_items(brow0, tile_b, R):
    """Work items of the lane-run kernel: ranges of whole bundles of one super-tile with about R
    rows (item_cuts: large first, small last).  brow0 int64[n_bundles+1] (host), tile_b
    [n_st+1] first bundle of every tile.  Returns (item_off int32[n_st+1], item_b0 int32
    [n_items + n_st]): the first bundles of tile t's items, then one sentinel (= tile_b[t+1])."""
    item_off, item_b0 = [0], []
    for t in range(len(tile_b) - 1):
        b_lo, b_hi = int(tile_b[t]), int(tile_b[t + 1])
        bcut = np.zeros(0, dtype=np.int64)
        if b_hi > b_lo:
            cuts = item_cuts(int(brow0[b_lo]), int(brow0[b_hi]), R)
            bcut = np.unique(np.searchsorted(brow0[b_lo:b_hi], cuts, side="left") + b_lo)
            bcut = bcut[bcut < b_hi]
            if bcut.size == 0 or bcut[0] != b_lo:
                bcut = np.concatenate([[b_lo], bcut])
        item_b0.append(bcut)
        item_b0.append(np.array([b_hi], dtype=np.int64))
        item_off.append(item_off[-1] + bcut.size)
    return np.asarray(item_off, dtype=np.int32), np.concatenate(item_b0).astype(np.int32)


def guided_cuts(cost, n_claimers, min_cost, guide=2.0, max_cost=None):
    """Cut positions (indices into `cost`) of one phase into work items.
    cost: float64/int64[n+1] ascending cumulative cost at the n+1 bundle boundaries of the phase.
    Items of max_cost while plenty is left — the warps of a launch then read a compact window of
    the stream (large items spread the ~4 700 concurrently streamed ranges over gigabytes, which
    costs a third of the bandwidth, profiles/r01_item_rows_sweep.txt) — then guided
    self-scheduling: an item takes 1/(guide * n_claimers) of what is left, never less than
    min_cost, so that the warps of a CTA finish together.  The ideal positions are snapped to
    bundle boundaries.  Returns the first boundary of every item (ascending, unique, first = 0,
    all < n)."""
    n = len(cost) - 1
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    c0, total = float(cost[0]), float(cost[-1] - cost[0])
    d = max(float(guide) * n_claimers, 1.0)
    min_cost = max(float(min_cost), 1.0)
    max_cost = max(float(max_cost), min_cost) if max_cost else total
    left = [np.array([total])]
    if total > max_cost * d:  # items of max_cost until left / d drops below it
        n_f = int((total - max_cost * d) // max_cost) + 1
        left.append(total - max_cost * np.arange(1, n_f + 1))
    cur = float(left[-1][-1])
    if cur > min_cost * d:    # guided part
        n_g = int(np.ceil(np.log(min_cost * d / cur) / np.log1p(-1.0 / d)))
        left.append(cur * np.power(1.0 - 1.0 / d, np.arange(1, n_g + 1)))
        cur = float(left[-1][-1])
    left.append(np.arange(cur - min_cost, 0.0, -min_cost))  # tail of min_cost items
    pos = (total + c0) - np.concatenate(left)  # ascending
    cuts = np.searchsorted(cost, pos, side="left")
    cuts = cuts[cuts < n]
    if cuts.size:
        cuts = cuts[np.concatenate(([True], cuts[1:] != cuts[:-1]))]
    if cuts.size == 0 or cuts[0] != 0:
        cuts = np.concatenate([[0], cuts])
    return cuts.astype(np.int64)


def ell2_chunks(srow0, tile_b, n_ctas, warps, max_items, bundle_cost=8, min_rows=40, guide=2.0,
                max_rows=512, pool_frac=0.15, pool_unit=8000, pool_min=150000):
    """Work plan of k_stream_ell2 (csrc/vell2.cu).  srow0 int64[nb+1]: first STREAM row of every
    bundle (header rows included); tile_b int64[n_st+1]: first bundle of every super-tile.
    The bundles are cut into n_ctas stretches of equal cost (rows + bundle_cost per bundle).  The
    first 1 - pool_frac of a stretch is the CTA's own chunk; the rest goes, in units of about
    pool_unit cost, into a POOL that CTAs draw from (largest units first) once their own chunk
    is done — equal cost is not equal time (the same chunk runs up to 25 % faster or slower
    depending on the SM and on which part of the hit table it writes).  Chunks below pool_min
    cost get no pool.  A chunk or unit is cut
    into phases (its bundles of at most two consecutive tiles) and a phase into work items of
    at most max_rows rows (guided_cuts).
    Returns (item_row0 int32[n_items+1] in stream order; phase int32[n_ph][4] = first tile,
    tiles (| 256: claim the items last to first), first item, items — the phases of the chunks CTA by CTA, then the pool's;
    cta_ph0 int32[n_ctas+1]; pool0 = index of the first pool phase)."""
    srow0 = np.asarray(srow0, dtype=np.int64)
    tile_b = np.asarray(tile_b, dtype=np.int64)
    nb, n_st = len(srow0) - 1, len(tile_b) - 1
    cost = (srow0 + bundle_cost * np.arange(nb + 1, dtype=np.int64)).astype(np.float64)
    total = float(cost[-1])

    # pieces in stream order: (owner CTA or -1 for the pool, first bundle, end bundle)
    per = total / n_ctas
    # (a pool phase costs a claim, a tile load and two barriers: measured worth it only for
    # chunks of >= ~150 000 rows — config 4 on one GPU: -2.5 % —, a loss on a 1/8 gene shard)
    n_units = int(min(4, round(pool_frac * per / pool_unit))) if (pool_frac > 0 and per >= pool_min) else 0
    k = np.arange(n_units + 1, dtype=np.float64)
    rel = np.concatenate([[1.0 - pool_frac + pool_frac * kk / n_units for kk in k[:-1]], [1.0]]) \
        if n_units else np.array([1.0])
    ends = (np.arange(n_ctas, dtype=np.float64)[:, None] + rel[None, :]).reshape(-1) * per
    eb = np.searchsorted(cost, ends, side="left")
    eb = np.maximum.accumulate(np.minimum(eb, nb))
    eb[-1] = nb
    pieces, b = [], 0
    for i, e in enumerate(eb):
        e = int(e)
        if e > b:
            c, u = divmod(i, n_units + 1)
            pieces.append((c if u == 0 else -1, b, e))
        b = max(b, e)
    # phases and items of every piece, items numbered in stream order
    item_row0, n_it = [], 0
    own = [[] for _ in range(n_ctas)]
    pool = []
    for o, b, b1 in pieces:
        while b < b1:
            t = int(np.searchsorted(tile_b, b, side="right")) - 1
            end = min(b1, int(tile_b[min(t + 2, n_st)]))
            tiles = 2 if (t + 1 < n_st and int(tile_b[t + 1]) < end) else 1
            # bundles are sorted by length inside a tile, so a phase that crosses a tile boundary
            # runs short -> long: its items are cut and claimed from the far end, the longest
            # bundles (44 us of one warp's time each at 512 rows) must not come last
            rev = tiles == 2
            seg = cost[b:end + 1]
            if rev:
                seg = seg[-1] - seg[::-1]
            mc, xc = float(min_rows), float(max_rows)
            while True:
                cuts = guided_cuts(seg, warps, mc, guide, xc)
                if cuts.size <= max_items:
                    break
                mc, xc = mc * 1.25, xc * 1.25
            if rev:
                cuts = np.concatenate(([0], (end - b) - cuts[:0:-1]))
            item_row0.append(srow0[b + cuts])
            ph = (t, tiles | (256 if rev else 0), n_it, cuts.size)
            if o >= 0:
                own[o].append(ph)
            else:
                pool.append((-(cost[end] - cost[b]), len(pool), ph))
            n_it += cuts.size
            b = end
    phases, cta_ph0 = [], [0]
    for c in range(n_ctas):
        phases.extend(own[c])
        cta_ph0.append(len(phases))
    pool0 = len(phases)
    phases.extend(ph for _, _, ph in sorted(pool))  # largest first
    item_row0.append(srow0[nb: nb + 1])
    return (np.concatenate(item_row0).astype(np.int32),
            np.asarray(phases, dtype=np.int32).reshape(-1, 4),
            np.asarray(cta_ph0, dtype=np.int32), pool0)
