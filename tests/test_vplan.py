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



def _emulate_lane_run_layout(lens, n_seg, st, cap, R, rng):
    """numpy emulation of prb_v_scatter_flat + vplan.ell_plan + prb_ell_build + the item walk
    of k_stream_ell: every entry of every run must be visited exactly once, by a lane whose
    header names the run, through items that cover every bundle exactly once."""
    import torch

    lens_t = torch.from_numpy(lens.astype(np.int64))
    run_start = torch.cumsum(lens_t, 0) - lens_t
    total = int(lens.sum())
    tmp = np.arange(total, dtype=np.int64)  # entry e of the scratch "is" the number e
    ep = vplan.ell_plan(lens_t, run_start, n_seg, st, cap)
    nb, brow0 = ep["n_bundles"], ep["brow0"].numpy()
    lr_src, lr_len, lr_hdr = ep["lr_src"].numpy(), ep["lr_len"].numpy(), ep["lr_hdr"].numpy()
    assert brow0[0] == 0 and np.all(np.diff(brow0) >= 1)
    rows = int(brow0[-1])
    ell = np.full((rows, 32, 2), -7, dtype=np.int64)  # prb_ell_build
    hdr_rows = np.zeros((max(nb, 1), 32), dtype=np.int64)
    for b in range(nb):
        L = brow0[b + 1] - brow0[b]
        n = lr_len[b * 32:(b + 1) * 32]
        assert L == (n[0] + 1) // 2 and np.all(n <= n[0]) and n[0] <= cap
        hdr_rows[b] = lr_hdr[b * 32:(b + 1) * 32]
        for l in range(32):
            e = tmp[lr_src[b * 32 + l]: lr_src[b * 32 + l] + n[l]]
            col = np.full(2 * L, -1, dtype=np.int64)
            col[: n[l]] = e
            ell[brow0[b]: brow0[b + 1], l, :] = col.reshape(L, 2)
    assert not np.any(ell == -7)
    # items
    tile_b = np.concatenate([ep["tile_b0"].numpy(), [nb]])
    item_off, item_b0 = vplan.ell_items(brow0, tile_b, R)
    seen_b = np.zeros(nb, dtype=np.int64)
    owner = np.full(total, -1, dtype=np.int64)
    n_st = len(st) - 1
    for t in range(n_st):
        for k in range(item_off[t + 1] - item_off[t]):
            b0, b1 = item_b0[item_off[t] + t + k], item_b0[item_off[t] + t + k + 1]
            assert tile_b[t] <= b0 < b1 <= tile_b[t + 1]
            for b in range(b0, b1):
                seen_b[b] += 1
                for l in range(32):
                    h = hdr_rows[b, l]
                    ents = ell[brow0[b]: brow0[b + 1], l].reshape(-1)
                    ents = ents[ents >= 0]
                    if h == -1:
                        assert ents.size == 0
                        continue
                    vg, seg_local = h & 0xFFFFFF, h >> 24
                    run = vg * n_seg + st[t] + seg_local
                    assert st[t] + seg_local < st[t + 1]
                    assert np.all(owner[ents] == -1)
                    owner[ents] = run
    assert np.all(seen_b == 1)
    want = np.repeat(np.arange(lens.size), lens)
    assert np.array_equal(owner, want)
    return rows, nb


@pytest.mark.parametrize("cap,R", [(2, 32), (8, 32), (64, 96), (128, 32), (1024, 992)])
def test_lane_run_plan_covers_every_entry_once(cap, R):
    rng = np.random.RandomState(cap + R)
    n_seg, st = 7, [0, 3, 4, 7]
    PV = 53
    lens = (rng.gamma(0.4, 60.0, PV * n_seg)).astype(np.int64)
    lens[rng.rand(lens.size) < 0.3] = 0
    lens[5] = 3000  # a run much longer than the cap
    _emulate_lane_run_layout(lens, n_seg, st, cap, R, rng)


def test_lane_run_plan_degenerate():
    rng = np.random.RandomState(0)
    _emulate_lane_run_layout(np.zeros(12, dtype=np.int64), 3, [0, 3], 64, 32, rng)  # empty matrix
    _emulate_lane_run_layout(np.array([0, 0, 1, 0], dtype=np.int64), 2, [0, 1, 2], 64, 32, rng)
    _emulate_lane_run_layout(np.full(40, 1, dtype=np.int64), 1, [0, 1], 2, 32, rng)


# ---- the streamed lane-run layout (csrc/vell2.cu): chunks, phases, pool, items ---------------
def _check_ell2_plan(srow0, tile_b, n_ctas, warps, max_items, **kw):
    item_row0, phase, cta_ph0, pool0 = vplan.ell2_chunks(srow0, tile_b, n_ctas, warps, max_items, **kw)
    nb = len(srow0) - 1
    assert item_row0[0] == srow0[0] and item_row0[-1] == srow0[-1]
    assert np.all(np.diff(item_row0) > 0) or nb == 0
    assert set(item_row0.tolist()) <= set(np.asarray(srow0).tolist())  # items start at bundle starts
    assert len(cta_ph0) == n_ctas + 1 and cta_ph0[0] == 0 and cta_ph0[-1] == pool0 <= len(phase)
    assert np.all(np.diff(cta_ph0) >= 0)
    covered = np.zeros(len(item_row0) - 1, dtype=np.int64)
    for t0, nt, i0, n in phase:
        assert nt & 255 in (1, 2) and (nt >> 8) == (1 if nt & 255 == 2 else 0)
        nt &= 255
        assert 1 <= n <= max_items
        covered[i0: i0 + n] += 1
        # every bundle of the phase lies in tile t0 or (nt == 2) t0 + 1
        b0 = np.searchsorted(srow0, item_row0[i0])
        b1 = np.searchsorted(srow0, item_row0[i0 + n])
        assert tile_b[t0] <= b0 and b1 <= tile_b[min(t0 + nt, len(tile_b) - 1)]
    assert np.all(covered == 1)
    # pool phases: largest first
    bc = kw.get("bundle_cost", 8)
    cost = [int(item_row0[i0 + n] - item_row0[i0])
            + bc * int(np.searchsorted(srow0, item_row0[i0 + n]) - np.searchsorted(srow0, item_row0[i0]))
            for _, _, i0, n in phase[pool0:]]
    assert all(a >= b for a, b in zip(cost, cost[1:]))
    return phase, cta_ph0, pool0


def test_ell2_chunks_cover_every_bundle_once():
    rng = np.random.RandomState(3)
    L = rng.randint(1, 300, size=40000)
    srow0 = np.concatenate([[0], np.cumsum(L + 1)]).astype(np.int64)
    tile_b = np.array([0, 7000, 7000, 16000, 30000, 40000])
    ph, cp, p0 = _check_ell2_plan(srow0, tile_b, 148, 32, 624)
    assert p0 == len(ph)  # chunks of ~40 000 rows: below pool_min, no pool
    ph, cp, p0 = _check_ell2_plan(srow0, tile_b, 148, 32, 624, pool_min=0, pool_unit=500)
    assert p0 < len(ph)


def test_ell2_chunks_small_and_degenerate():
    rng = np.random.RandomState(4)
    for nb in (0, 1, 5, 200, 3000):
        L = rng.randint(1, 40, size=nb)
        srow0 = np.concatenate([[0], np.cumsum(L + 1)]).astype(np.int64)
        # many small tiles: chunks span more than two tiles -> several phases per CTA
        n_t = 17
        tile_b = np.unique(np.concatenate([[0, nb], rng.randint(0, nb + 1, size=n_t)]))
        tile_b = np.sort(np.concatenate([tile_b, tile_b[1:2]]))  # an empty tile
        _check_ell2_plan(srow0, tile_b, 148, 32, 624, pool_min=0, pool_unit=50)
        _check_ell2_plan(srow0, tile_b, 4, 32, 16, pool_min=0, pool_unit=50)


def test_guided_cuts_sizes():
    cost = np.arange(0, 300001, 150, dtype=np.int64)  # bundles of 150 rows
    cuts = vplan.guided_cuts(cost, 32, 40, 2.0, 512)
    sizes = np.diff(np.concatenate([cost[cuts], cost[-1:]]))
    assert cuts[0] == 0 and sizes.sum() == 300000
    assert sizes.max() <= 512 + 150 and sizes[:100].min() >= 450  # capped items first
    assert sizes[-1] <= 300  # small ones last


def test_ell_plan_streamed_headers():
    import torch

    n_seg, st = 5, [0, 2, 5]
    lens = torch.tensor([0, 3, 2500, 1, 7, 9, 0, 0, 40, 2], dtype=torch.int64)  # 2 virtual genes
    run_start = torch.cumsum(lens, 0) - lens
    seg_id = torch.tensor([4, 5, 6, 7, 8], dtype=torch.int64)
    ep = vplan.ell_plan(lens, run_start, n_seg, st, 1024, seg_id=seg_id, id_shift=20)
    hdr = ep["lr_hdr"].numpy().astype(np.int64) & 0xFFFFFFFF
    ln = ep["lr_len"].numpy()
    live = ln > 0
    assert np.all((hdr[~live] & 0xFFFFF) == 0xFFFFF)  # empty lanes
    vg = hdr[live] & 0xFFFFF
    sid = (hdr[live] >> 20) & 0x3FF
    split = (hdr[live] >> 30) & 1
    assert set(vg.tolist()) <= {0, 1} and set(sid.tolist()) <= {4, 5, 6, 7, 8}
    # the run of 2 500 entries is cut into three flagged pieces, nothing else is flagged
    assert split.sum() == 3 and np.all(ln[live][split == 1] >= 833)
    assert ln[live].sum() == int(lens.sum())


def _emulate_streamed_layout(lens, n_seg, st, seg_class, cap, n_ctas, rng, **kw):
    """numpy emulation of the streamed lane-run layout (csrc/vell2.cu): prb_v_scatter_flat +
    vplan.ell_plan(seg_id=...) + prb_ell_build(hdr_in_stream=1) + vplan.ell2_chunks, then the walk
    of k_stream_ell2 — CTA by CTA, phase by phase, items in claim order (last to first in a
    phase that crosses a tile boundary), bundle by bundle with the row count taken from bit 31
    of the 32 header words.  Every entry of every run must be visited exactly once, by a lane
    whose header names the run's (segment, virtual gene); the tiles of a phase must be the ones
    its bundles belong to."""
    import torch

    id_shift = 20
    lens_t = torch.from_numpy(lens.astype(np.int64))
    run_start = torch.cumsum(lens_t, 0) - lens_t
    total = int(lens.sum())
    seg_id = torch.arange(n_seg, dtype=torch.int64)
    ep = vplan.ell_plan(lens_t, run_start, n_seg, st, cap, seg_id=seg_id, id_shift=id_shift)
    nb = ep["n_bundles"]
    brow0 = ep["brow0"].numpy()
    srow0 = brow0 + np.arange(nb + 1)  # one header row per bundle
    lr_src, lr_len = ep["lr_src"].numpy(), ep["lr_len"].numpy()
    lr_hdr = ep["lr_hdr"].numpy().astype(np.int64) & 0xFFFFFFFF
    rows = int(srow0[-1])
    stream = np.full((rows, 32, 2), -7, dtype=np.int64)  # data rows: two entries per lane
    header = np.full((rows, 32), -1, dtype=np.int64)     # header rows: one word per lane
    for b in range(nb):
        L = int(srow0[b + 1] - srow0[b] - 1)
        n = lr_len[b * 32:(b + 1) * 32]
        assert L == (n[0] + 1) // 2 >= 1 and np.all(n <= n[0]) and n[0] <= cap
        lbits = (L >> np.arange(32)) & 1
        header[srow0[b]] = (lr_hdr[b * 32:(b + 1) * 32] & 0x7FFFFFFF) | (lbits << 31)
        for l in range(32):
            col = np.full(2 * L, -1, dtype=np.int64)
            col[: n[l]] = np.arange(lr_src[b * 32 + l], lr_src[b * 32 + l] + n[l])
            stream[srow0[b] + 1: srow0[b + 1], l, :] = col.reshape(L, 2)
    tile_b = np.concatenate([ep["tile_b0"].numpy(), [nb]])
    item_row0, phase, cta_ph0, pool0 = vplan.ell2_chunks(srow0, tile_b, n_ctas, 32, 624, **kw)
    seg_tile = np.repeat(np.arange(len(st) - 1), np.diff(st))
    owner = np.full(total, -1, dtype=np.int64)
    seen_rows = np.zeros(rows, dtype=np.int64)

    def walk_phase(ph):
        t0, nt, i0, n_it = (int(v) for v in phase[ph])
        rev, nt = nt >> 8, nt & 255
        resident = set(range(t0, t0 + nt))
        for k in range(n_it):
            kk = n_it - 1 - k if rev else k
            r, r_end = int(item_row0[i0 + kk]), int(item_row0[i0 + kk + 1])
            while r < r_end:  # a bundle: header row, then L data rows
                h = header[r]
                assert np.all(h >= 0), "an item must start at a header row"
                L = int(sum(((int(h[l]) >> 31) & 1) << l for l in range(32)))
                seen_rows[r: r + 1 + L] += 1
                for l in range(32):
                    vg = int(h[l]) & ((1 << id_shift) - 1)
                    seg = (int(h[l]) & 0x3FFFFFFF) >> id_shift
                    ents = stream[r + 1: r + 1 + L, l].reshape(-1)
                    ents = ents[ents >= 0]
                    if vg == (1 << id_shift) - 1:
                        assert ents.size == 0
                        continue
                    assert seg_tile[seg] in resident
                    assert np.all(owner[ents] == -1)
                    owner[ents] = vg * n_seg + seg
                    split = (int(h[l]) >> 30) & 1
                    assert split == int(lens[vg * n_seg + seg] > cap)
                r += 1 + L
            assert r == r_end

    for c in range(n_ctas):
        for ph in range(int(cta_ph0[c]), int(cta_ph0[c + 1])):
            walk_phase(ph)
    for ph in range(pool0, len(phase)):
        walk_phase(ph)
    assert np.all(seen_rows == 1)
    assert np.array_equal(owner, np.repeat(np.arange(lens.size), lens))


@pytest.mark.parametrize("cap,n_ctas,kw", [
    (2, 7, {}), (8, 148, {}), (64, 148, dict(pool_min=0, pool_unit=40)),
    (1024, 5, dict(pool_min=0, pool_unit=200, min_rows=4, max_rows=16))])
def test_streamed_layout_walk_covers_every_entry_once(cap, n_ctas, kw):
    rng = np.random.RandomState(cap + n_ctas)
    n_seg, st = 7, [0, 1, 3, 4, 7]  # four tiles
    seg_class = [0, 1, 1, 2, 3, 3, 3]
    PV = 61
    lens = (rng.gamma(0.4, 60.0, PV * n_seg)).astype(np.int64)
    lens[rng.rand(lens.size) < 0.3] = 0
    lens[5] = 3000  # a run much longer than the cap: pieces flagged "split"
    _emulate_streamed_layout(lens, n_seg, st, seg_class, cap, n_ctas, rng, **kw)


def test_streamed_layout_walk_degenerate():
    rng = np.random.RandomState(0)
    _emulate_streamed_layout(np.zeros(12, dtype=np.int64), 3, [0, 3], [0, 0, 0], 64, 4, rng)
    _emulate_streamed_layout(np.array([0, 0, 1, 0], dtype=np.int64), 2, [0, 1, 2], [0, 1], 64, 4, rng)
    _emulate_streamed_layout(np.full(40, 1, dtype=np.int64), 1, [0, 1], [0], 2, 148, rng)
