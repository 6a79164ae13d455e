"""Host logic of the V layout (picturedrocks_b200/vplan.py): segment plan, super-tiles and
work items.  Pure functions — runs without a GPU."""
import numpy as np
import pytest

from picturedrocks_b200 import vplan

CASES = [
    ([1306127], 65504, 31),                                     # no labels, config-4 cells
    ([21769] * 59 + [21756], 65504, 31),                        # 60 equal classes
    ([1, 0, 5, 100000, 3, 0, 70000], 65504, 31),                # empty and oversized classes
    ([10] * 256, 65504, 31),                                    # many tiny classes: max_segs binds
    ([300, 200, 100], 256, 31),                                 # tiny tile: every class split
    ([37], 128, 31),
]


@pytest.mark.parametrize("sizes,cap,max_segs", CASES)
def test_segment_plan_invariants(sizes, cap, max_segs):
    cell0, seg_class, st = vplan.segment_plan(sizes, cap, max_segs)
    N = sum(sizes)
    n_seg = len(seg_class)
    assert cell0[0] == 0 and cell0[-1] == N and len(cell0) == n_seg + 1
    assert all(b > a for a, b in zip(cell0, cell0[1:]))         # no empty segment
    # a segment never crosses a class boundary and carries the right class
    bounds = np.concatenate(([0], np.cumsum(sizes)))
    for s in range(n_seg):
        c = seg_class[s]
        assert bounds[c] <= cell0[s] and cell0[s + 1] <= bounds[c + 1]
    # segments of a class are contiguous and cover it
    for c, n in enumerate(sizes):
        cov = sum(cell0[s + 1] - cell0[s] for s in range(n_seg) if seg_class[s] == c)
        assert cov == n
    # super-tiles: consecutive, bounded span and segment count
    assert st[0] == 0 and st[-1] == n_seg and all(b > a for a, b in zip(st, st[1:]))
    for t in range(len(st) - 1):
        assert st[t + 1] - st[t] <= max_segs
        assert cell0[st[t + 1]] - cell0[st[t]] <= cap


def test_segment_plan_balances_super_tiles():
    cell0, _, st = vplan.segment_plan([21769] * 60, 65504, 31)
    spans = [cell0[st[t + 1]] - cell0[st[t]] for t in range(len(st) - 1)]
    assert len(spans) == 20 and max(spans) == min(spans) == 3 * 21769


@pytest.mark.parametrize("rows,warps,lo,hi", [
    (86_001_274, 4736, 1408, 1664),     # config 4 on one GPU: measured optimum 1 536
    (43_563_220, 4736, 896, 1152),      # half / quarter / eighth of its genes
    (22_035_146, 4736, 576, 832),
    (10_678_142, 4736, 256, 640),
    (2_278_450, 4736, 64, 96),          # config 3
    (130_591, 4736, 32, 32), (100, 4736, 32, 32), (0, 4736, 32, 32)])
def test_item_rows(rows, warps, lo, hi):
    R = vplan.item_rows(rows, warps)
    assert R % 32 == 0 and lo <= R <= hi
    assert vplan.item_rows(rows, warps, forced=320) == 320


@pytest.mark.parametrize("lo,hi,R", [(0, 4_300_000, 1152), (1000, 1000 + 545_000, 256),
                                     (7, 40, 32), (5, 5, 64), (0, 1, 32)])
def test_item_cuts_cover_the_block(lo, hi, R):
    cuts = vplan.item_cuts(lo, hi, R)
    if hi == lo:
        assert cuts.size == 0
        return
    assert cuts[0] == lo and np.all(np.diff(cuts) > 0) and cuts[-1] < hi
    sizes = np.diff(np.concatenate((cuts, [hi])))
    assert sizes.sum() == hi - lo and sizes.max() <= R
    # large items first, small items last (only the very last one may be a remainder)
    assert np.all(np.diff(sizes[:-1]) <= 0)
    if hi - lo > 64 * R:
        assert sizes[0] == R and sizes[-2] <= max(128, R // 4)

