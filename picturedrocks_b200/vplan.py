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

