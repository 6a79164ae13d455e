"""Device-resident information set: packed by-gene CSC + count tables in HBM.

Host-side orchestration of the kernels in csrc/ through the C-ABI
(include/prb200.h).  PyTorch tensors own the device memory and provide the
stream; every computation on the path is a libprb200.so kernel.

Reference being replaced (picturedrocks/markers/mutualinformation/infoset.py):
  SparseInformationSet.__init__/set_y :188-216, entropy :218-249,
  entropy_wrt :251-303, _sparse_entropy_wrt :306-410, _sparse_entropy :413-428,
  _sparse_make_master_col :431-442.

Multi-GPU: an Engine streams only the genes [gene_lo, gene_lo + P) of a
P_total-gene matrix (one process per GPU), but keeps a replicated copy of the
whole packed CSC (one all-gather over NVLink at set-up), so the selected gene's
column is always read locally; the only per-step exchange is the 16-byte
(score, index) all-gather of the argmax — see dist.py.
"""
import ctypes
import os

import numpy as np
import torch

from . import _native as nat
from . import vplan

_I16_MAX_BINS = 65536
# K <= 21: the V layouts (cells relabelled by class, register counters): the streamed lane-run
# layout of csrc/vell2.cu (default), the lane-run layout of csrc/vell.cu (PRB_ELL_V=1) or, for
# K <= 5, the warp-per-run layout of csrc/vstream.cu (PRB_V_KERNEL=rows).  PRB_NO_VLAYOUT=1
# forces the generic shared-histogram kernel (csrc/stream.cu) everywhere.
USE_VLAYOUT = os.environ.get("PRB_NO_VLAYOUT", "0") != "1"
ELL_MAX_K1 = 20
PACKED_PAD = 4  # words of slack after the last stored entry
SX_NONE = 255   # csrc/common.cuh PRB_SX_NONE: shift code of a cell with x_j = 0


def _i32(x):
    return ctypes.c_int32(int(x))


class NotCanonical(ValueError):
    """The CSC arrays handed to Engine.from_csc_arrays are not a canonical matrix of bin codes
    (flags of csrc/ingest.cu: 1 row id out of range, 2 code outside [0, 255], 4 stored zero,
    8 rows of a gene not strictly ascending)."""

    def __init__(self, flags):
        super().__init__("CSC arrays are not canonical (flags %d)" % flags)
        self.flags = flags


class Engine:
    def __init__(self, indptr, packed, N, K, gene_lo=0, P_total=None, comm=None, term=None):
        """indptr: int64[P+1] device tensor (local offsets, indptr[0] == 0);
        packed: int32 device tensor holding the uint32 words code<<24 | row;
        term: optional host float64[N+1] term table (see host_term_table)."""
        nat.require_cuda()
        nat.load()
        assert indptr.dtype == torch.int64 and indptr.is_cuda and indptr.is_contiguous()
        assert packed.dtype == torch.int32 and packed.is_cuda and packed.is_contiguous()
        self.device = indptr.device
        self.indptr = indptr
        self.packed = packed
        self.N = int(N)
        self.P = indptr.numel() - 1
        self.gene_lo = int(gene_lo)
        self.P_total = int(P_total) if P_total is not None else self.P
        self.comm = comm
        if self.N <= 0 or self.N > (1 << 24):
            raise ValueError("number of cells must be in (0, 2^24] (packed row field)")
        self.K = int(K)
        if comm is not None:
            self.K = comm.max_int(self.K)
        if not (2 <= self.K <= 256):
            raise ValueError("bin codes must lie in [0, 255] with at least one non-zero code")
        self.K1 = self.K - 1
        self.shift = (self.K - 1).bit_length()  # infoset.py:199  int(log2(max)+1)
        self.nnz = int(indptr[-1].item())
        if packed.numel() < self.nnz + PACKED_PAD:  # keep 16 bytes of slack behind the stream
            packed = torch.cat([packed[: self.nnz], packed.new_zeros(PACKED_PAD)])
            self.packed = packed
        self._replicate_columns()
        sm = ctypes.c_int()
        smem = ctypes.c_int()
        nat.call("prb_device_info", ctypes.byref(sm), ctypes.byref(smem))
        self.sm_count, self.smem_optin = sm.value, smem.value
        # fp64 term table, built on the host (glibc log2), resident in HBM
        tt = term if term is not None else self.host_term_table(self.N)
        self.term = torch.from_numpy(tt).to(self.device)
        # per-cell state workspaces
        dev = self.device
        # per-cell joint state (uint8 or uint16 view), padded for 128-bit tile copies
        self.state_buf = torch.zeros(2 * self.N + 64, dtype=torch.uint8, device=dev)
        self.xj = torch.zeros(self.N + 16, dtype=torch.uint8, device=dev)  # + slack: 16-byte copies
        # V path: per-cell shift code of the selected gene (csrc/vstream.cu), all "x_j = 0"
        # between steps: prb_pick_scatter writes a column, the step's tail resets it
        self.sx = torch.full((self.N + 64,), SX_NONE, dtype=torch.uint8, device=dev)
        self.mpacked = None
        self.hsize = torch.zeros(_I16_MAX_BINS, dtype=torch.int32, device=dev)
        self.hs_y = torch.zeros(_I16_MAX_BINS, dtype=torch.int32, device=dev)
        self.hs_begin = torch.zeros(258, dtype=torch.int32, device=dev)
        self.hs_ysum = torch.zeros(257, dtype=torch.int32, device=dev)
        self.ha = torch.zeros(256, dtype=torch.int32, device=dev)
        self.n_hs_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(8192, dtype=torch.int32, device=dev)
        self.ticket = torch.zeros(4, dtype=torch.int32, device=dev)
        self.ell_dbg = torch.zeros(16, dtype=torch.int32, device=dev)  # see prb_stream_ell
        self._plans = {}
        self._tps = {}
        kern = os.environ.get("PRB_V_KERNEL", "ell")
        self.ell = USE_VLAYOUT and self.K1 <= ELL_MAX_K1 and (kern != "rows" or self.K1 > 4)
        # streamed lane-run layout (csrc/vell2.cu), the default; PRB_ELL_V=1: the layout of
        # csrc/vell.cu (global work queue, side arrays for the bundle headers)
        self.ell2 = (self.ell and os.environ.get("PRB_ELL_V", "2") == "2"
                     and self.P * self.K1 < (1 << 22) - 1)
        # ... whose hit table is v-major for K <= 5 (16-bit counts, hits[g][segment][v][x_j - 1],
        # plain 8-byte stores); the bundle header then carries the segment beside the virtual gene
        self.vmajor = (self.ell2 and self.K1 <= 4
                       and os.environ.get("PRB_ELL_VMAJOR", "1") != "0")
        self.vmode = USE_VLAYOUT and (self.ell or self.K1 <= 4)
        # warp-per-run layout: 16-bit entries, 64 per row (prb_v_scatter16 / prb_stream_v16);
        # PRB_V16=0 selects the 32-bit entries of round 1 (1.40 vs 1.80 ms at config 4)
        self.v16 = self.vmode and not self.ell and os.environ.get("PRB_V16", "1") == "1"
        if self.vmode:
            nat.call("prb_v_check_smem_base")
        self.v = None  # V layout (csrc/vstream.cu), built lazily for the current labels
        self.newpos = None    # V path: int32[N] new position of every cell (None = identity)
        self.y_sorted = None  # V path: labels in the new order
        self.classsizes_host = None
        self.gene_ptr = torch.zeros(1, dtype=torch.int32, device=dev)
        self._hits = None
        self._table = None
        self.epoch = 0  # bumped whenever device buffers a captured graph points at change
        self.clsz_all = torch.tensor([self.N], dtype=torch.int64, device=dev)
        self._cnt_rep = {}  # replicated count tables (all genes), by "labels or not"
        # labels
        self.has_y = False
        self.y = None
        self.C = 1
        self.classsizes = None
        self.cnt_y = None
        self._cnt_x = None  # label-free count table, built on first use

    def _replicate_columns(self):
        """col_packed / col_begin / col_end: the packed CSC of ALL genes, addressed by the
        global gene id.  One GPU: the local arrays.  Gene-sharded: every rank's shard is
        all-gathered once (in place, equal slots) over NVLink, so that the column of the
        selected gene (infoset.py:289 -> 431-442) never has to be exchanged per step."""
        if self.comm is None or self.comm.world == 1:
            self.col_packed = self.packed
            self.col_begin, self.col_end = self.indptr[:-1], self.indptr[1:]
            return
        comm = self.comm
        slot = (comm.max_int(self.nnz) + PACKED_PAD + 3) & ~3
        full = torch.empty(comm.world * slot + PACKED_PAD, dtype=torch.int32, device=self.device)
        mine = full[comm.rank * slot: (comm.rank + 1) * slot]
        mine[: self.nnz] = self.packed[: self.nnz]
        mine[self.nnz:].zero_()
        full[comm.world * slot:].zero_()
        comm.all_gather_slots(full, slot)
        self.packed = mine  # the local shard is a view of the replicated copy
        bounds = comm.all_gather_shards(self.indptr[:-1] + comm.rank * slot), \
            comm.all_gather_shards(self.indptr[1:] + comm.rank * slot)
        self.col_packed = full
        self.col_begin, self.col_end = bounds[0].contiguous(), bounds[1].contiguous()

    def cnt_all(self, with_y):
        """Count table of ALL genes [P_total][C][K-1] (replicated over the ranks)."""
        cnt = self.cnt_y if with_y else self.cnt_x
        if self.comm is None or self.comm.world == 1:
            return cnt
        key = bool(with_y)
        if key not in self._cnt_rep:
            width = (self.C if with_y else 1) * self.K1
            self._cnt_rep[key] = self.comm.all_gather_shards(cnt[: self.P * width], width)
        return self._cnt_rep[key]

    @staticmethod
    def host_term_table(N):
        """T[c] = -(c/N)*log2(c/N) for c = 0..N on the HOST (glibc log2; infoset.py:392-397)."""
        tt = np.empty(int(N) + 1, dtype=np.float64)
        nat.call("prb_term_table", int(N), tt.ctypes.data)
        return tt

    # ------------------------------------------------------------------ tables
    @property
    def cnt_x(self):
        """cnt_x[g][v-1] = #cells with code v in gene g (independent of the labels)."""
        if self._cnt_x is None:
            if self.vmode and self.v is not None and self.P:
                n_seg = self.v["plan"]["n_seg"]  # sum of the run lengths over the segments
                L = self.v["seglen"][: self.P * self.K1 * n_seg].view(self.P * self.K1, n_seg)
                self._cnt_x = L.sum(1, dtype=torch.int32).contiguous()
            else:
                self._cnt_x = self._count_table(None, 1)
        return self._cnt_x

    def _plan(self, n_states):
        """(tile_cells, n_tiles, teams, state_bytes) of a streaming launch with n_states states."""
        pl = self._plans.get(n_states)
        if pl is None:
            tile, nt, teams, sb, smem = (ctypes.c_int64(), ctypes.c_int(), ctypes.c_int(),
                                         ctypes.c_int(), ctypes.c_int64())
            nat.call("prb_stream_plan", self.N, n_states, self.K1, ctypes.byref(tile),
                     ctypes.byref(nt), ctypes.byref(teams), ctypes.byref(sb), ctypes.byref(smem))
            pl = (tile.value, nt.value, teams.value, sb.value)
            if nt.value > self.counters.numel():
                raise NotImplementedError("too many cell tiles")
            self._plans[n_states] = pl
        return pl

    def _tile_ptrs(self, tile_cells, n_tiles):
        key = (tile_cells, n_tiles)
        tp = self._tps.get(key)
        if tp is None:
            tp = torch.empty((n_tiles + 1) * max(self.P, 1), dtype=torch.int64, device=self.device)
            if self.P:
                nat.call("prb_tile_ptrs", nat.ptr(self.packed), nat.ptr(self.indptr), self.P,
                         tile_cells, n_tiles, nat.ptr(tp), nat.stream())
            self._tps[key] = tp
        return tp

    def _store_state(self, values, sb):
        """values: int32 device tensor [N] of state ids + 1 -> state_buf as uint8/uint16."""
        if sb == 1:
            self.state_buf[: self.N] = values.to(torch.uint8)
        else:
            self.state_buf[: 2 * self.N] = values.to(torch.int16).view(torch.uint8)

    def _count_table(self, y, C):
        """cnt[g][c][v-1] = #cells of class c with code v in gene g.  V layout: the run
        lengths; generic path: one streaming pass in which every cell is a "master row"
        whose state is its class."""
        if C * self.K1 > _I16_MAX_BINS:
            raise NotImplementedError("classes x codes exceeds 65536 histogram bins")
        if self.vmode:
            return self._count_table_v(y is not None, C)
        cnt = torch.zeros(max(self.P * C * self.K1, 1), dtype=torch.int32, device=self.device)
        sb = self._plan(C)[3]
        if y is None:
            self._store_state(torch.ones(self.N, dtype=torch.int32, device=self.device), sb)
        else:
            self._store_state(y + 1, sb)
        self._stream(C, 0, cnt)
        return cnt

    def set_y(self, y):
        """y: int32 device tensor [N] of labels in 0..C-1, or None (infoset.py:202-216).
        The packed CSC keeps its cell ids; the V layout (csrc/vstream.cu) is built with the
        cells RELABELLED so that classes are contiguous (newpos = a stable sort of the
        labels — entropies do not depend on the order of cells), and the per-cell vectors of
        that path (x_j, its shift codes, y_sorted) live in the new order."""
        self.epoch += 1
        self.v = None  # new labels: new segments
        self._cnt_rep.pop(True, None)
        if y is None:
            self.has_y, self.y, self.C, self.classsizes, self.cnt_y = False, None, 1, None, None
            self.classsizes_host = self.newpos = self.y_sorted = None
            return
        assert y.dtype == torch.int32 and y.is_cuda and y.numel() == self.N
        C = int(y.max().item()) + 1
        if C > 256:
            raise NotImplementedError("more than 256 classes")
        self.has_y, self.y, self.C = True, y.contiguous(), C
        self.classsizes = torch.bincount(self.y, minlength=C).to(torch.int64)
        self.classsizes_host = [int(v) for v in self.classsizes.cpu().numpy()]
        if self.vmode:
            self.y_sorted, order = torch.sort(self.y, stable=True)
            self.newpos = torch.empty(self.N, dtype=torch.int32, device=self.device)
            self.newpos[order] = torch.arange(self.N, dtype=torch.int32, device=self.device)
        self.cnt_y = self._count_table(self.y, C)

    # ------------------------------------------------ V layout (csrc/vstream.cu), K <= 5
    def _v_plan(self):
        """Segment plan: one segment per class (split when a class exceeds the shared-memory
        tile; a single class [0, N) without labels) packed into super-tiles."""
        classes = self.has_y
        cap, max_segs = ctypes.c_int64(), ctypes.c_int()
        nat.call("prb_v_tile_capacity", ctypes.byref(cap), ctypes.byref(max_segs))
        if self.ell2:  # tiles of the two-tile window; the class travels in the bundle header
            tc = ctypes.c_int()
            nat.call("prb_ell2_geometry", ctypes.byref(tc), None, None, None)
            cap, max_segs = ctypes.c_int64(tc.value), ctypes.c_int(1 << 30)
        cap = max(128, min(int(os.environ.get("PRB_V_TILE_CELLS", cap.value)), cap.value))
        sizes = self.classsizes_host if classes else [self.N]
        cell0, cls, st = vplan.segment_plan(sizes, cap, max_segs.value)
        assert cell0[-1] == self.N
        n_seg = len(cls)
        n_st = len(st) - 1
        if n_st > self.counters.numel():
            raise NotImplementedError("too many cell tiles")
        max_cells = max(cell0[st[t + 1]] - cell0[st[t]] for t in range(n_st))
        i32 = dict(dtype=torch.int32, device=self.device)
        return dict(n_seg=n_seg, n_st=n_st, max_cells=max_cells, cell0=cell0, st=st,
                    seg_cell0=torch.tensor(cell0, **i32), st_seg0=torch.tensor(st, **i32),
                    seg_class=torch.tensor(cls, **i32) if classes else None)

    def _build_v(self):
        """Re-order the packed CSC into the V layout for the current segment plan: cells
        relabelled by class, (gene, code) virtual genes, runs per segment padded to 32-entry
        rows, stored tile-major; plus the work-item tables of csrc/vstream.cu."""
        pl = self._v_plan()
        P, K1, n_seg, dev = self.P, self.K1, pl["n_seg"], self.device
        newpos = self.newpos if self.has_y else None
        PV = max(P * K1, 1)
        if n_seg * K1 > 65536:
            raise NotImplementedError("too many cell segments")
        seglen = torch.zeros(PV * n_seg, dtype=torch.int32, device=dev)
        # per-cell table of the two layout kernels, indexed by the ORIGINAL cell id:
        # bits 0..23 relabelled position, bits 24..31 segment (seg16 when there are > 256)
        cellmap = seg16 = None
        pl["pos"] = newpos
        if n_seg > 1 or newpos is not None:
            pos = newpos if newpos is not None else torch.arange(self.N, dtype=torch.int32,
                                                                 device=dev)
            cellmap = pos
            if n_seg > 1:
                if n_seg * K1 > 65536:
                    raise NotImplementedError("too many cell segments")
                seg = torch.bucketize(pos, pl["seg_cell0"][1:-1].contiguous(), right=True)
                if self.ell2 and pl["n_st"] > 1:
                    # padded cell space: super-tile t owns [t * 32768, t * 32768 + its cells)
                    st_t = torch.tensor(pl["st"], dtype=torch.int64, device=dev)
                    seg_tile = torch.repeat_interleave(
                        torch.arange(pl["n_st"], dtype=torch.int64, device=dev), st_t[1:] - st_t[:-1])
                    shift = (torch.arange(pl["n_st"], dtype=torch.int64, device=dev) * vplan.ELL2_TILE
                             - pl["seg_cell0"].to(torch.int64)[st_t[:-1]])
                    pos = (pos + shift[seg_tile[seg]].to(torch.int32)).contiguous()
                    pl["pos"] = pos
                    cellmap = pos & 0xFFFFFF  # (only the low 16 bits reach the layout)
                if n_seg <= 256:
                    cellmap = (cellmap | (seg << 24).to(torch.int32)).contiguous()
                else:
                    seg16 = seg.to(torch.int16)  # bit pattern of a uint16
        nat.call("prb_v_count", nat.ptr(self.packed), nat.ptr(self.indptr), P, nat.ptr(cellmap),
                 nat.ptr(seg16), n_seg, K1, nat.ptr(seglen), nat.ptr(self.counters), nat.stream())
        if self.ell2:
            return self._build_ell2(pl, seglen, cellmap, seg16)
        if self.ell:
            return self._build_ell(pl, seglen, cellmap, seg16)
        # rows of every run, scanned in storage order: (super-tile, virtual gene, segment)
        span = nat.load().prb_v_tile_span()
        cell0, st = pl["cell0"], pl["st"]
        epr_log = 6 if self.v16 else 5  # entries per 128-byte row: 64 (16-bit) or 32
        nr = (seglen.view(PV, n_seg) + ((1 << epr_log) - 1)) >> epr_log
        run_end = torch.empty((PV, n_seg), dtype=torch.int32, device=dev)
        tile_row0, base = [0], 0
        pad = np.zeros(n_seg, dtype=np.int64)
        for t in range(pl["n_st"]):
            a, b = st[t], st[t + 1]
            cs = torch.cumsum(nr[:, a:b].reshape(-1), 0, dtype=torch.int64) + base
            base = int(cs[-1].item())
            if base >= (1 << 31):
                raise NotImplementedError("more than 2^31 rows in the V layout")
            run_end[:, a:b] = cs.view(PV, b - a).to(torch.int32)
            tile_row0.append(base)
            # padding words point at the cell slot just past the run's super-tile
            pad[a:b] = 0xFF000000 | (cell0[b] & (span - 1))
        v_rows = base
        seg_pad = torch.from_numpy(pad.astype(np.uint32).view(np.int32)).to(dev)
        vpacked = torch.empty(v_rows * 32 + PACKED_PAD, dtype=torch.int32, device=dev)
        nat.call("prb_v_scatter16" if self.v16 else "prb_v_scatter", nat.ptr(self.packed),
                 nat.ptr(self.indptr), P, nat.ptr(cellmap),
                 nat.ptr(seg16), n_seg, K1, nat.ptr(seglen), nat.ptr(run_end),
                 nat.ptr(seg_pad), nat.ptr(vpacked), nat.ptr(self.counters), nat.stream())
        # work items: consecutive row ranges of every super-tile's block (size: the measured
        # rule of vplan.item_rows); the first 3/4 of a block in large items, then two rounds
        # of smaller ones, so that the tail of a launch is short
        warps = nat.load().prb_v_warps()
        R = vplan.item_rows(v_rows, warps, int(os.environ.get("PRB_V_ITEM_ROWS", 0)))
        item_off, row0, vg0 = [0], [], []
        for t in range(pl["n_st"]):
            cuts = vplan.item_cuts(tile_row0[t], tile_row0[t + 1], R)
            hi = tile_row0[t + 1]
            row0.append(cuts)
            row0.append(np.array([hi], dtype=np.int64))  # sentinel
            item_off.append(item_off[-1] + cuts.size)
        row0 = np.concatenate(row0).astype(np.int32)
        row0_d = torch.from_numpy(row0).to(dev)
        for t in range(pl["n_st"]):
            a, b = item_off[t] + t, item_off[t + 1] + t
            if b > a:
                vg0.append(torch.searchsorted(run_end[:, st[t + 1] - 1].contiguous(), row0_d[a:b],
                                              right=True).to(torch.int32))
        item_vg0 = (torch.cat(vg0) if vg0 else torch.zeros(1, dtype=torch.int32, device=dev))
        self.v = dict(plan=pl, vpacked=vpacked, run_end=run_end, seglen=seglen, rows=v_rows,
                      relabelled=self.has_y,
                      tile_row0=torch.tensor(tile_row0, dtype=torch.int32, device=dev),
                      item_rows=R, n_items=item_off[-1], item_vg0=item_vg0.contiguous(),
                      item_row0=row0_d,
                      item_off=torch.tensor(item_off, dtype=torch.int32, device=dev))
        self.sx.fill_(SX_NONE)
        self.epoch += 1
        return self.v

    def _build_ell(self, pl, seglen, cellmap, seg16):
        """The lane-run layout of csrc/vell.cu from the run lengths seglen[P*K1][n_seg]: flat
        run-major scratch -> runs cut into lane-runs of at most CAP entries -> sorted by
        (super-tile, length descending) -> bundles of 32 -> transposed rows; work items =
        ranges of whole bundles."""
        P, K1, n_seg, dev = self.P, self.K1, pl["n_seg"], self.device
        n_st, st, cell0 = pl["n_st"], pl["st"], pl["cell0"]
        i64 = dict(dtype=torch.int64, device=dev)
        span = 1 << 16
        lens = seglen[: P * K1 * n_seg].to(torch.int64)
        n_runs = lens.numel()
        run_start = torch.cumsum(lens, 0) - lens
        total = int(lens.sum().item()) if n_runs else 0
        assert total == self.nnz
        tmp = torch.empty(total + 16, dtype=torch.int16, device=dev)
        nat.call("prb_v_scatter_flat", nat.ptr(self.packed), nat.ptr(self.indptr), P,
                 nat.ptr(cellmap), nat.ptr(seg16), n_seg, K1, nat.ptr(seglen), nat.ptr(run_start),
                 nat.ptr(tmp), nat.ptr(self.counters), nat.stream())
        warps = nat.load().prb_ell_warps()
        R = vplan.item_rows(total // 64 + 1, warps, int(os.environ.get("PRB_V_ITEM_ROWS", 0)))
        cap = int(os.environ.get("PRB_ELL_CAP", 0)) or 1024
        cap = min(cap, 32768) & ~1
        ep = vplan.ell_plan(lens, run_start, n_seg, st, cap)
        n_bundles, brow0, nb_t, tile_b0 = ep["n_bundles"], ep["brow0"], ep["nb_t"], ep["tile_b0"]
        lr_src, lr_len, lr_hdr = ep["lr_src"], ep["lr_len"], ep["lr_hdr"]
        del ep
        rows = int(brow0[-1].item())
        if rows >= (1 << 31):
            raise NotImplementedError("more than 2^31 rows in the lane-run layout")
        brow0 = brow0.to(torch.int32)
        # padding entries point at the cell slot just past the bundle's super-tile
        pad_t = torch.tensor([cell0[st[t + 1]] & (span - 1) for t in range(n_st)], **i64)
        bundle_pad = torch.repeat_interleave(pad_t, nb_t).to(torch.int16)
        ell = torch.empty(rows * 32 + PACKED_PAD, dtype=torch.int32, device=dev)
        bhdr = torch.empty(max(n_bundles, 1) * 32, dtype=torch.int32, device=dev)
        nat.call("prb_ell_build", nat.ptr(tmp), nat.ptr(lr_src), nat.ptr(lr_len), nat.ptr(lr_hdr),
                 nat.ptr(brow0), nat.ptr(bundle_pad), n_bundles, nat.ptr(ell), nat.ptr(bhdr), 0,
                 nat.stream())
        # work items: ranges of whole bundles with about R rows (large first, small last)
        tb = torch.cat([tile_b0, tile_b0.new_tensor([n_bundles])]).cpu().numpy()
        item_off, item_b0 = vplan.ell_items(brow0.cpu().numpy().astype(np.int64), tb, R)
        torch.cuda.current_stream().synchronize()  # tmp / lr_* are released below
        del tmp, lr_src, lr_len, lr_hdr
        self.v = dict(plan=pl, kind="ell", ell=ell, bhdr=bhdr, brow0=brow0, n_bundles=n_bundles, seglen=seglen,
                      rows=rows, relabelled=self.has_y, item_rows=R, cap=cap, n_items=int(item_off[-1]),
                      item_b0=torch.from_numpy(item_b0).to(dev),
                      item_off=torch.from_numpy(item_off).to(dev))
        self.sx.fill_(SX_NONE)
        self.epoch += 1
        return self.v

    def _build_ell2(self, pl, seglen, cellmap, seg16):
        """The streamed lane-run layout of csrc/vell2.cu: the lane-runs and bundles of
        _build_ell, every bundle preceded by its header row, the stream cut into one chunk per
        CTA / phases of two tiles / guided work items (vplan.ell2_chunks)."""
        P, K1, n_seg, dev = self.P, self.K1, pl["n_seg"], self.device
        n_st, st = pl["n_st"], pl["st"]
        i64 = dict(dtype=torch.int64, device=dev)
        geo = [ctypes.c_int() for _ in range(4)]
        nat.call("prb_ell2_geometry", *[ctypes.byref(g) for g in geo])
        _, n_ctas, warps, max_items = [g.value for g in geo]
        if self.C > 256:
            raise NotImplementedError("more than 256 classes")
        lens = seglen[: P * K1 * n_seg].to(torch.int64)
        n_runs = lens.numel()
        run_start = torch.cumsum(lens, 0) - lens
        total = int(lens.sum().item()) if n_runs else 0
        assert total == self.nnz
        tmp = torch.empty(total + 16, dtype=torch.int16, device=dev)
        nat.call("prb_v_scatter_flat", nat.ptr(self.packed), nat.ptr(self.indptr), P,
                 nat.ptr(cellmap), nat.ptr(seg16), n_seg, K1, nat.ptr(seglen), nat.ptr(run_start),
                 nat.ptr(tmp), nat.ptr(self.counters), nat.stream())
        # lane-runs of at most `cap` entries; small matrices get shorter ones so that a CTA's
        # chunk still holds several bundles
        cap = int(os.environ.get("PRB_ELL_CAP", 0))
        if not cap:
            cap = int(min(1024, max(64, (total // 64 // n_ctas // 4) * 2)))
        cap = min(cap, 32768) & ~1
        vg_bits = max(8, int(P * K1 + 1).bit_length())
        if self.vmajor and n_seg > (1 << (30 - vg_bits)):
            self.vmajor = False  # (segment and virtual gene do not fit one header word)
        if self.vmajor:  # the slot of a run in hits: its segment
            seg_id, id_shift = torch.arange(n_seg, **i64), vg_bits
        else:            # ... its class
            seg_id = (pl["seg_class"].to(torch.int64) if pl["seg_class"] is not None
                      else torch.zeros(n_seg, **i64))
            id_shift = 22
        ep = vplan.ell_plan(lens, run_start, n_seg, st, cap, seg_id=seg_id, id_shift=id_shift)
        n_bundles, nb_t, tile_b0 = ep["n_bundles"], ep["nb_t"], ep["tile_b0"]
        lr_src, lr_len, lr_hdr = ep["lr_src"], ep["lr_len"], ep["lr_hdr"]
        srow0 = ep["brow0"] + torch.arange(n_bundles + 1, **i64)  # + one header row per bundle
        del ep
        rows = int(srow0[-1].item())
        if rows >= (1 << 31):
            raise NotImplementedError("more than 2^31 rows in the lane-run layout")
        # padding entries point at slot 32767 of the bundle's half of the tile window
        pad_t = ((torch.arange(n_st, **i64) & 1) << 15) | (vplan.ELL2_TILE - 1)
        bundle_pad = torch.repeat_interleave(pad_t, nb_t).to(torch.int16)
        ell = torch.empty(rows * 32 + PACKED_PAD, dtype=torch.int32, device=dev)
        srow0_32 = srow0.to(torch.int32)
        # (the planner's inputs come to the host BEFORE the builder is launched: the host plans
        # the chunks while the kernel transposes the lane-runs)
        tb = torch.cat([tile_b0, tile_b0.new_tensor([n_bundles])]).cpu().numpy()
        srow0_host = srow0.cpu().numpy()
        nat.call("prb_ell_build", nat.ptr(tmp), nat.ptr(lr_src), nat.ptr(lr_len), nat.ptr(lr_hdr),
                 nat.ptr(srow0_32), nat.ptr(bundle_pad), n_bundles, nat.ptr(ell), None, 1,
                 nat.stream())
        env = os.environ.get
        item_row0, phase, cta_ph0, pool0 = vplan.ell2_chunks(
            srow0_host, tb, n_ctas, warps, max_items,
            bundle_cost=int(env("PRB_ELL_BCOST", 8)), min_rows=int(env("PRB_ELL_MIN_ROWS", 40)),
            guide=float(env("PRB_ELL_GUIDE", 2.0)), max_rows=int(env("PRB_ELL_MAX_ROWS", 512)),
            pool_frac=float(env("PRB_ELL_POOL", 0.15)), pool_unit=float(env("PRB_ELL_POOL_UNIT", 8000)),
            pool_min=float(env("PRB_ELL_POOL_MIN", 150000)))
        torch.cuda.current_stream().synchronize()  # tmp / lr_* are released below
        del tmp, lr_src, lr_len, lr_hdr
        # 5 < K <= 21: the fused step on a v-major table with slot = class (finalize_wide.cu) needs
        # every class in ONE segment (two segments would store to the same row)
        seg_cls = pl["seg_class"].cpu().tolist() if pl["seg_class"] is not None else [0] * n_seg
        wide = (K1 > 4 and not self.vmajor and len(set(seg_cls)) == n_seg
                and os.environ.get("PRB_WIDE_FUSED", "1") != "0"
                and bool(nat.load().prb_wide_supported(self.C, self.K))
                and P * int(nat.load().prb_wide_terms_stride(self.C, self.K)) * 8 <= (24 << 30))
        cls_slot0 = all_slot0 = None
        if self.vmajor:  # segments of class y = a range of slots of the hit table
            all_slot0 = torch.tensor([0, n_seg], dtype=torch.int32, device=dev)
            cls_slot0 = all_slot0
            if pl["seg_class"] is not None:
                cls_slot0 = torch.searchsorted(
                    pl["seg_class"].contiguous(),
                    torch.arange(self.C + 1, dtype=torch.int32, device=dev)).to(torch.int32)
        self.v = dict(plan=pl, kind="ell2", ell=ell, n_bundles=n_bundles, seglen=seglen, rows=rows,
                      id_shift=id_shift, cls_slot0=cls_slot0, all_slot0=all_slot0, wide=wide,
                      relabelled=pl["pos"] is not None, pos=pl["pos"], item_rows=0, cap=cap,
                      n_items=int(item_row0.size - 1), n_ctas=n_ctas, pool0=pool0,
                      pool1=int(phase.shape[0]),
                      srow0=srow0_32 if env("PRB_ELL_KEEP_PLAN") else None,
                      item_row0=torch.from_numpy(item_row0).to(dev),
                      phase=torch.from_numpy(phase).to(dev).contiguous(),
                      cta_ph0=torch.from_numpy(cta_ph0).to(dev))
        need = n_st * vplan.ELL2_TILE + 64
        if self.sx.numel() < need:
            self.sx = torch.empty(need, dtype=torch.uint8, device=dev)
        self.sx.fill_(SX_NONE)
        self.epoch += 1
        return self.v

    def _v(self):
        return self.v if self.v is not None else self._build_v()

    def _hits_v(self, with_y, v):
        """The zeroed hit table for one pass of the V path + what the kernels need to address
        it: (hits, cls_slot0, n_slot).  v-major table (csrc/vell2.cu): one slot per SEGMENT
        (16-bit counts); the finalize kernels add the slots [cls_slot0[y], cls_slot0[y+1]) of
        class y — all of them in a call that does not condition on the labels."""
        if not self.vmajor:
            C = self.C if with_y else 1
            return self.hits_buffer(C * self.K1 * self.K1), None, C
        n_seg = v["plan"]["n_seg"]
        cls_slot0 = v["cls_slot0"] if with_y else v["all_slot0"]
        row_words = 2 * ((self.K1 + 3) // 4)  # 32-bit words of one (segment, v) row
        return self.hits_buffer(n_seg * self.K1 * row_words), cls_slot0, n_seg

    def _vpos(self, v):
        """Position of every cell in the per-cell vectors of the V path (sx): the class-sorted
        order, for the streamed lane-run layout in its padded cell space; None = identity."""
        if "pos" in v:
            return v["pos"]
        return self.newpos if v["relabelled"] else None

    def _count_table_v(self, with_classes, C):
        """cnt[g][c][v-1] from run lengths (no streaming pass over the hot kernel).  Without
        labels only the per-(gene, code) totals are needed: one counting pass with a single
        segment; the V layout itself is built when the labels (or the first step) arrive."""
        dev, P, K1 = self.device, self.P, self.K1
        if not with_classes:
            seglen = torch.zeros(max(P * K1, 1), dtype=torch.int32, device=dev)
            nat.call("prb_v_count", nat.ptr(self.packed), nat.ptr(self.indptr), P, None, None, 1,
                     K1, nat.ptr(seglen), nat.ptr(self.counters), nat.stream())
            return seglen
        v = self._v()
        pl = v["plan"]
        L = v["seglen"][: P * K1 * pl["n_seg"]].view(P, K1, pl["n_seg"])
        cnt = torch.zeros((P, K1, C), dtype=torch.int32, device=dev)
        cnt.index_add_(2, pl["seg_class"].long(), L)
        cnt = cnt.permute(0, 2, 1).contiguous().view(-1)
        return cnt if cnt.numel() else torch.zeros(1, dtype=torch.int32, device=dev)

    def _stream_v(self, with_y, C, hits, vm=None):
        """The streaming pass (counters were zeroed by prb_pick_scatter)."""
        ev = getattr(self, "profile_events", None)
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self._stream_v_launch(with_y, C, hits, vm)
            e1.record()
            ev.append((e0, e1))
        else:
            self._stream_v_launch(with_y, C, hits, vm)

    def _stream_v_launch(self, with_y, C, hits, vm=None):
        """vm: v-major 16-bit hit table (default: the engine's choice for K <= 5; the wide
        fused step passes True with slot = class)."""
        v = self._v()
        pl = v["plan"]
        if self.ell2:
            vm = self.vmajor if vm is None else vm
            n_slot = v["plan"]["n_seg"] if self.vmajor else C
            nat.call("prb_stream_ell2", nat.ptr(v["ell"]), nat.ptr(v["item_row0"]),
                     nat.ptr(v["phase"]), nat.ptr(v["cta_ph0"]), v["n_ctas"], v["pool0"],
                     v["pool1"], nat.ptr(self.counters), nat.ptr(self.sx),
                     self.K1, n_slot, v["id_shift"], nat.ptr(hits), int(vm),
                     self.P * self.K1, nat.ptr(getattr(self, "ell_timing", None)), nat.stream())
            return
        if self.ell:
            nat.call("prb_stream_ell", nat.ptr(v["ell"]), nat.ptr(v["bhdr"]), nat.ptr(v["brow0"]),
                     v["n_bundles"],
                     self.K1, nat.ptr(pl["st_seg0"]), nat.ptr(pl["seg_cell0"]),
                     nat.ptr(pl["seg_class"]) if with_y else None, pl["n_st"], pl["max_cells"],
                     nat.ptr(self.sx), C, nat.ptr(hits), nat.ptr(self.counters),
                     nat.ptr(v["item_off"]), nat.ptr(v["item_b0"]), v["n_items"], 1,
                     self.P * self.K1, nat.ptr(self.ell_dbg), nat.stream())
            return
        nat.call("prb_stream_v16" if self.v16 else "prb_stream_v", nat.ptr(v["vpacked"]),
                 nat.ptr(v["run_end"]),
                 nat.ptr(v["tile_row0"]), pl["n_seg"], self.P, self.K1, v["rows"],
                 nat.ptr(pl["st_seg0"]),
                 nat.ptr(pl["seg_cell0"]), nat.ptr(pl["seg_class"]) if with_y else None,
                 pl["n_st"], pl["max_cells"], nat.ptr(self.sx), C, nat.ptr(hits),
                 nat.ptr(self.counters), nat.ptr(v["item_off"]), nat.ptr(v["item_row0"]),
                 nat.ptr(v["item_vg0"]), v["n_items"], 1, nat.stream())

    def _pick_scatter(self, with_y, records, n_rec, sel):
        """Head of a V-path step (csrc/step.cu): winner of the candidate records (n_rec = 0:
        the gene already in gene_ptr), commit, state sizes, work counters, column -> sx."""
        v = self._v()
        C = self.C if with_y else 1
        newpos = self._vpos(v)
        peer = getattr(sel, "peer", None) if sel is not None else None
        nat.call("prb_pick_scatter", nat.ptr(records), n_rec, nat.ptr(self.gene_ptr),
                 nat.ptr(sel.S_dev) if sel else None, nat.ptr(sel.n_sel) if sel else None,
                 nat.ptr(sel.selected) if sel else None, _i32(self.gene_lo), self.P,
                 nat.ptr(self.col_packed), nat.ptr(self.col_begin), nat.ptr(self.col_end),
                 nat.ptr(newpos), nat.ptr(self.sx), nat.ptr(self.cnt_all(with_y)), C, self.K,
                 nat.ptr(self.hsize), nat.ptr(self.hs_begin), nat.ptr(self.hs_y),
                 nat.ptr(self.hs_ysum), nat.ptr(self.ha), nat.ptr(self.counters),
                 v["plan"]["n_st"], nat.ptr(peer.local) if peer else None,
                 peer.comm.world if peer else 0, nat.ptr(peer.xtag) if peer else None,
                 nat.ptr(peer.err) if peer else None, nat.stream())

    def fused_step(self, sel, kind, is_remove, pick):
        """One greedy step of a selector on the V path, three launches, no host sync:
        prb_pick_scatter -> prb_stream_v -> prb_finalize_update (entropies of INFOSET:392-397,
        the score update of ITER:48-59 / 76-87 / 112-121, this rank's argmax record and the
        reset of sx).  pick = False: add(ind) / remove(ind) with the gene in gene_ptr."""
        with_y = sel.supervised
        C, y, cnt, clsz = self._labels(with_y)
        n_bins = C * self.K1 * self.K1
        if n_bins > _I16_MAX_BINS:
            raise NotImplementedError("C*(K-1)^2 exceeds 65536 histogram bins")
        self.cnt_all(with_y)  # (collective on first use: before any kernel of the step)
        v = self._v()
        if self.P == 0:
            return
        self._pick_scatter(with_y, sel.records, sel.n_rec if pick else 0, sel)
        newpos = self._vpos(v)
        peer = getattr(sel, "peer", None)  # the step's record goes out from the kernel's tail
        if self.K1 > 4:
            # 5 < K <= 21 (csrc/finalize_wide.cu): v-major table with slot = class, the ordered
            # term list of every gene, then the thread-per-gene chain with the fused tail
            assert self.wide_ok(with_y)
            row_words = 2 * ((self.K1 + 3) // 4)
            hits = self.hits_buffer(C * self.K1 * row_words)
            self._stream_v(with_y, C, hits, vm=True)
            terms, terms2, n_terms = self._wide_buffers(C)
            nat.call("prb_finalize_update_wide", nat.ptr(hits), nat.ptr(cnt), nat.ptr(clsz),
                     nat.ptr(self.hsize), nat.ptr(self.hs_ysum), nat.ptr(self.ha), C, self.K,
                     nat.ptr(self.term), self.N, self.P, nat.ptr(terms), nat.ptr(terms2),
                     nat.ptr(n_terms), kind, int(is_remove), nat.ptr(sel.base),
                     nat.ptr(sel.penalty), nat.ptr(sel.score), nat.ptr(sel.selected),
                     nat.ptr(sel.Hx_all), nat.ptr(sel.Hyx_all), nat.ptr(self.gene_ptr),
                     nat.ptr(sel.n_sel), _i32(self.gene_lo), nat.ptr(sel.cand_score),
                     nat.ptr(sel.cand_idx), nat.ptr(sel.best), nat.ptr(self.ticket),
                     nat.ptr(self.col_packed), nat.ptr(self.col_begin), nat.ptr(self.col_end),
                     nat.ptr(newpos), nat.ptr(self.sx),
                     nat.ptr(peer.peer_ptrs) if peer else None, peer.comm.world if peer else 0,
                     peer.comm.rank if peer else 0, nat.ptr(peer.xtag) if peer else None,
                     nat.stream())
            return
        hits, cls_slot0, n_slot = self._hits_v(with_y, v)
        self._stream_v(with_y, C, hits)
        nat.call("prb_finalize_update", nat.ptr(hits), nat.ptr(cnt), nat.ptr(clsz),
                 nat.ptr(self.hsize), nat.ptr(self.hs_ysum), nat.ptr(self.ha), C, self.K,
                 nat.ptr(self.term), self.N, self.P, kind, int(is_remove), nat.ptr(sel.base),
                 nat.ptr(sel.penalty), nat.ptr(sel.score), nat.ptr(sel.selected),
                 nat.ptr(sel.Hx_all), nat.ptr(sel.Hyx_all), nat.ptr(self.gene_ptr),
                 nat.ptr(sel.n_sel), _i32(self.gene_lo), nat.ptr(sel.cand_score),
                 nat.ptr(sel.cand_idx), nat.ptr(sel.best), nat.ptr(self.ticket),
                 nat.ptr(self.col_packed), nat.ptr(self.col_begin), nat.ptr(self.col_end),
                 nat.ptr(newpos), nat.ptr(self.sx), int(self.vmajor), nat.ptr(cls_slot0), n_slot,
                 nat.ptr(peer.peer_ptrs) if peer else None, peer.comm.world if peer else 0,
                 peer.comm.rank if peer else 0, nat.ptr(peer.xtag) if peer else None,
                 nat.stream())

    def wide_ok(self, with_y):
        """The fused step of the 5 < K <= 21 path is available: the layout keeps every class
        in one segment (and a call without labels sees a single segment)."""
        if not (self.vmode and self.ell2 and self.K1 > 4 and self.P):
            return False
        v = self._v()
        return bool(v.get("wide")) and (with_y or v["plan"]["n_seg"] == 1)

    def _wide_buffers(self, C):
        key = (C, self.P, self.K)
        if getattr(self, "_wide_key", None) != key:
            self._wide = None
            self.epoch += 1  # (a captured step points at the old buffers)
            stride = int(nat.load().prb_wide_terms_stride(C, self.K))
            self._wide = (torch.empty(self.P * stride, dtype=torch.float64, device=self.device),
                          torch.empty(self.P * self.K * self.K, dtype=torch.float64,
                                      device=self.device),
                          torch.empty(self.P, dtype=torch.int32, device=self.device))
            self._wide_key = key
        return self._wide

    # --------------------------------------------------------------- primitives
    def hits_buffer(self, n_bins):
        need = max(self.P * n_bins, 1)
        if self._hits is None or self._hits.numel() < need:
            self._hits = None
            self.epoch += 1
            self._hits = torch.zeros(need, dtype=torch.int32, device=self.device)
        return self._hits

    def _stream(self, n_states, direct_C, hits):
        ev = getattr(self, "profile_events", None)
        if ev is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            self._stream_launch(n_states, direct_C, hits)
            e1.record()
            ev.append((e0, e1))
        else:
            self._stream_launch(n_states, direct_C, hits)

    def _stream_launch(self, n_states, direct_C, hits):
        tile, nt, _, _ = self._plan(n_states)
        tp = self._tile_ptrs(tile, nt)
        nat.call("prb_stream_hits", nat.ptr(self.packed), nat.ptr(tp), self.P, self.nnz,
                 nat.ptr(self.state_buf), self.N, n_states, self.K1, direct_C, nat.ptr(hits),
                 nat.ptr(self.counters), nat.stream())

    def _finalize(self, hits, cnt, clsz, C, n_hs_cap, n_hs_ptr, out_full, out_marg, direct,
                  vm=(0, None, 1)):
        """vm = (v-major hits, cls_slot0, n_slot) of a V-path pass (_hits_v)."""
        wide = os.environ.get("PRB_WIDE_FINALIZE", "1") != "0" and self.K1 > 4
        if hits is None or (direct and not wide):
            # single-gene layout with K <= 5 (or no master column at all): lanes / one
            # thread per gene.  More bins per class: the warp-per-gene kernel below, whose
            # loads and term gathers run in parallel over the lanes.
            nat.call("prb_finalize_direct", nat.ptr(hits), nat.ptr(cnt), nat.ptr(clsz),
                     nat.ptr(self.hsize), nat.ptr(self.hs_ysum), nat.ptr(self.ha), C, self.K,
                     nat.ptr(self.term), self.N, self.P, nat.ptr(out_full), nat.ptr(out_marg),
                     int(vm[0]), nat.ptr(vm[1]), vm[2], nat.stream())
            return
        nat.call("prb_finalize", nat.ptr(hits), nat.ptr(cnt), nat.ptr(clsz), nat.ptr(self.hs_begin),
                 nat.ptr(self.hs_y), nat.ptr(self.hsize), nat.ptr(self.hs_ysum),
                 nat.ptr(self.ha) if direct else None, nat.ptr(n_hs_ptr), n_hs_cap, C, self.K,
                 nat.ptr(self.term), self.N, self.P, nat.ptr(out_full), nat.ptr(out_marg),
                 nat.stream())

    def _labels(self, with_y):
        if with_y:
            if not self.has_y:
                raise AssertionError("information set has no y")
            return self.C, self.y, self.cnt_y, self.classsizes
        return 1, None, self.cnt_x, self.clsz_all

    def scatter_selected_gene(self):
        """Generic path: x_j[N] <- dense codes of gene *gene_ptr, read from the replicated
        packed CSC (no exchange)."""
        self.xj.zero_()
        nat.call("prb_scatter_gene_u8", nat.ptr(self.col_packed), nat.ptr(self.col_begin),
                 nat.ptr(self.col_end), nat.ptr(self.gene_ptr), self.P_total, None,
                 nat.ptr(self.xj), nat.stream())

    def direct_step(self, with_y, out_full, out_marg):
        """One fused pass for the gene in gene_ptr:
        out_full = H([y,] x_j, x_i); out_marg = H(x_j, x_i) (only with_y)."""
        C, y, cnt, clsz = self._labels(with_y)
        n_hs = C * self.K1
        n_bins = n_hs * self.K1
        if n_bins > _I16_MAX_BINS:
            raise NotImplementedError("C*(K-1)^2 exceeds 65536 histogram bins")
        if self.vmode:
            if self.P == 0:
                return
            v = self._v()
            self._pick_scatter(with_y, None, 0, None)
            hits, cls_slot0, n_slot = self._hits_v(with_y, v)
            self._stream_v(with_y, C, hits)
            self._finalize(hits, cnt, clsz, C, n_hs, None, out_full, out_marg, True,
                           vm=(int(self.vmajor), cls_slot0, n_slot))
            nat.call("prb_v_unscatter", nat.ptr(self.gene_ptr), nat.ptr(self.col_packed),
                     nat.ptr(self.col_begin), nat.ptr(self.col_end),
                     nat.ptr(self._vpos(v)), nat.ptr(self.sx),
                     nat.stream())
            return
        self.scatter_selected_gene()
        sb = self._plan(n_hs)[3]
        nat.call("prb_state_direct", nat.ptr(self.xj), nat.ptr(y), self.N, C, self.K,
                 nat.ptr(self.state_buf), sb, nat.ptr(self.hsize), nat.ptr(self.hs_begin),
                 nat.ptr(self.hs_y), nat.ptr(self.hs_ysum), nat.ptr(self.ha), nat.stream())
        hits = self.hits_buffer(n_bins)
        self._stream(n_hs, C, hits)
        self._finalize(hits, cnt, clsz, C, n_hs, None, out_full, out_marg, True)

    def _build_mpacked(self, cols):
        """Joint-state packing of several columns (infoset.py:431-442): first column is
        the most significant digit; each digit is `shift` bits wide."""
        m = len(cols)
        mbits = self.shift * m
        if mbits > 30:
            raise NotImplementedError("joint state of %d columns needs more than 30 bits" % m)
        if self.mpacked is None:
            self.mpacked = torch.zeros(self.N, dtype=torch.int32, device=self.device)
        self.mpacked.zero_()
        for pos, c in enumerate(cols):
            nat.call("prb_scatter_gene_or32", nat.ptr(self.col_packed), nat.ptr(self.col_begin),
                     nat.ptr(self.col_end), c, self.shift * (m - 1 - pos), nat.ptr(self.mpacked),
                     nat.stream())
        return mbits

    def _table_scratch(self, n_keys):
        if n_keys > (1 << 24):
            raise NotImplementedError("joint state space larger than 2^24 keys")
        if self._table is None or self._table.numel() < n_keys:
            self._table = torch.zeros(n_keys, dtype=torch.int32, device=self.device)
        return self._table

    def _norm_cols(self, cols):
        cols = [int(c) for c in np.asarray(cols, dtype=np.int64).ravel()]
        with_y = len(cols) > 0 and cols[0] == -1  # infoset.py:278
        if with_y:
            cols = cols[1:]
        out = []
        for c in cols:
            if c < 0:
                c += self.P_total  # scipy/numpy negative column index
            if not (0 <= c < self.P_total):
                raise IndexError("column index out of range")
            out.append(c)
        return with_y, out

    # -------------------------------------------------------------- public API
    def entropy_wrt_device(self, cols):
        """float64 device vector [P]: H(cols, x_i) for every local gene i."""
        with_y, cols = self._norm_cols(cols)
        C, y, cnt, clsz = self._labels(with_y)
        out = torch.empty(self.P, dtype=torch.float64, device=self.device)
        m = len(cols)
        if m == 0:
            self._finalize(None, cnt, clsz, C, 0, None, out, None, False)
        elif m == 1:
            self.gene_ptr.fill_(cols[0])
            self.direct_step(with_y, out, None)
        else:
            mbits = self._build_mpacked(cols)
            n_keys = C << mbits
            table = self._table_scratch(n_keys)
            nat.call("prb_state_table", nat.ptr(self.mpacked), nat.ptr(y), self.N, C, mbits,
                     nat.ptr(table), nat.ptr(self.hsize), nat.ptr(self.hs_begin),
                     nat.ptr(self.hs_y), nat.ptr(self.hs_ysum), nat.ptr(self.n_hs_dev),
                     self.hsize.numel(), nat.stream())
            n_hs = int(self.n_hs_dev.item())
            if n_hs * self.K1 > _I16_MAX_BINS:
                raise NotImplementedError(
                    "%d joint states x %d codes exceeds 65536 histogram bins" % (n_hs, self.K1))
            if n_hs == 0:
                self._finalize(None, cnt, clsz, C, 0, None, out, None, False)
            else:
                sb = self._plan(n_hs)[3]
                nat.call("prb_state_assign", nat.ptr(self.mpacked), nat.ptr(y), self.N, mbits,
                         nat.ptr(table), nat.ptr(self.state_buf), sb, nat.stream())
                hits = self.hits_buffer(n_hs * self.K1)
                self._stream(n_hs, 0, hits)
                self._finalize(hits, cnt, clsz, C, n_hs, None, out, None, False)
        return out

    def entropy_scalar(self, cols):
        """H(cols) as a Python float (infoset.py:218-249)."""
        with_y, cols = self._norm_cols(cols)
        C, y, _, _ = self._labels(with_y)
        mbits = self._build_mpacked(cols) if cols else 0
        n_keys = C << mbits
        table = self._table_scratch(n_keys)
        out = torch.empty(1, dtype=torch.float64, device=self.device)
        nat.call("prb_entropy_keys", nat.ptr(self.mpacked) if cols else None, nat.ptr(y), self.N,
                 mbits, n_keys, nat.ptr(table), nat.ptr(self.term), nat.ptr(out), nat.stream())
        return float(out.item())

    # ------------------------------------------------------------ constructors
    @classmethod
    def from_csc_arrays(cls, indptr, indices, data, N, chunk_entries=None, **kw):
        """A CSC matrix of bin codes as the caller holds it — indptr[P+1], row ids (int32 or
        int64) and codes (any integer dtype) of the stored entries; numpy arrays / CPU tensors
        (pinned or pageable) or device tensors — packed on the device (csrc/ingest.cu).
        Host arrays are uploaded in chunks on a copy stream, double-buffered, and every chunk is
        packed and validated while the next one is on the bus; the canonical-form check of
        infoset.py:191-196 (sorted rows, no duplicates, no stored zeros) runs on the device too.
        Raises NotCanonical when the caller has to canonicalise on the host first."""
        nat.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())

        def as_tensor(a):
            if isinstance(a, torch.Tensor):
                return a
            a = np.ascontiguousarray(a)
            if a.dtype.kind == "u" and a.dtype.itemsize > 1:  # torch has no uint16/32/64 copies
                a = a.view(np.dtype("i%d" % a.dtype.itemsize))
            if a.dtype == np.bool_:
                a = a.view(np.uint8)
            return torch.from_numpy(a)

        code_signed = int(not (isinstance(data, np.ndarray) and data.dtype.kind in "ub")
                          and not (isinstance(data, torch.Tensor) and data.dtype == torch.uint8))
        indptr_d = as_tensor(indptr).to(device=dev, dtype=torch.int64, non_blocking=True).contiguous()
        rows_h, codes_h = as_tensor(indices), as_tensor(data)
        if rows_h.dtype not in (torch.int32, torch.int64):
            rows_h = rows_h.to(torch.int64)
        if codes_h.dtype.is_floating_point:
            raise AssertionError("X should be integer dtype")
        nnz = rows_h.numel()
        if nnz == 0:  # infoset.py:199 int(np.log2(X.max()) + 1) of an all-zero matrix
            raise OverflowError("cannot convert float infinity to integer")
        rb, cb = rows_h.element_size(), codes_h.element_size()
        packed = torch.empty(nnz + PACKED_PAD, dtype=torch.int32, device=dev)
        packed[nnz:].zero_()  # the slack words; everything below is written by the pack kernel
        flags = torch.zeros(1, dtype=torch.int32, device=dev)
        ms = torch.cuda.current_stream()
        if rows_h.is_cuda:
            nat.call("prb_pack_chunk", nat.ptr(rows_h), rb, nat.ptr(codes_h), cb, code_signed, nnz,
                     N, nat.ptr(packed), nat.ptr(flags), nat.stream())
        else:
            ce = int(chunk_entries or os.environ.get("PRB_INGEST_CHUNK", 1 << 25))
            cs = torch.cuda.Stream()
            stage = [(torch.empty(min(ce, nnz), dtype=rows_h.dtype, device=dev),
                      torch.empty(min(ce, nnz), dtype=codes_h.dtype, device=dev)) for _ in range(2)]
            copied = [torch.cuda.Event() for _ in range(2)]
            packed_ev = [torch.cuda.Event() for _ in range(2)]
            cs.wait_stream(ms)
            for k, a in enumerate(range(0, nnz, ce)):
                b, buf = min(a + ce, nnz), k % 2
                sr, sc = stage[buf]
                if k >= 2:
                    cs.wait_event(packed_ev[buf])  # the staging buffer is free again
                with torch.cuda.stream(cs):
                    sr[: b - a].copy_(rows_h[a:b], non_blocking=True)
                    sc[: b - a].copy_(codes_h[a:b], non_blocking=True)
                    copied[buf].record(cs)
                ms.wait_event(copied[buf])
                nat.call("prb_pack_chunk", nat.ptr(sr), rb, nat.ptr(sc), cb, code_signed, b - a, N,
                         nat.ptr(packed) + 4 * a, nat.ptr(flags), nat.stream())
                packed_ev[buf].record(ms)
            for t in stage:  # (the allocator must not hand the buffers out before the copies ran)
                t[0].record_stream(cs)
                t[1].record_stream(cs)
        nat.call("prb_check_columns", nat.ptr(packed), nat.ptr(indptr_d), indptr_d.numel() - 1,
                 nat.ptr(flags), nat.stream())
        # host work while the (asynchronous) upload is in flight
        kw.setdefault("term", cls.host_term_table(N))
        kmax = packed.view(torch.uint8)[3: 4 * nnz: 4].max()
        f = int(flags.item())
        if f:
            raise NotCanonical(f)
        return cls(indptr_d, packed, N, int(kmax.item()) + 1, **kw)

    @classmethod
    def from_dense_codes(cls, codes, N, **kw):
        """codes: uint8 device tensor [P, ld] (column-major cells x genes, ld >= N)."""
        P, ld = codes.shape
        dev = codes.device
        counts = torch.zeros(P, dtype=torch.int64, device=dev)
        nat.call("prb_dense_count_nnz", nat.ptr(codes), N, P, ld, nat.ptr(counts), nat.stream())
        indptr = torch.zeros(P + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=indptr[1:])
        nnz = int(indptr[-1].item())
        packed = torch.zeros(nnz + PACKED_PAD, dtype=torch.int32, device=dev)
        nat.call("prb_dense_fill_csc", nat.ptr(codes), N, P, ld, nat.ptr(indptr), nat.ptr(packed),
                 nat.stream())
        K = (int(codes[:, :N].max().item()) + 1) if P else 1
        return cls(indptr, packed, N, K, **kw)

    def to_scipy_csc(self, n_genes=None):
        """Host scipy.sparse.csc_matrix (int64 data like the reference's) of the first
        n_genes local genes (default all)."""
        import scipy.sparse

        G = self.P if n_genes is None else min(int(n_genes), self.P)
        ip = self.indptr[: G + 1].cpu().numpy()
        n = int(ip[-1])
        rows = torch.empty(max(n, 1), dtype=torch.int32, device=self.device)
        codes = torch.empty(max(n, 1), dtype=torch.uint8, device=self.device)
        nat.call("prb_unpack_csc", nat.ptr(self.packed), n, nat.ptr(rows), nat.ptr(codes),
                 nat.stream())
        return scipy.sparse.csc_matrix(
            (codes[:n].cpu().numpy().astype(np.int64), rows[:n].cpu().numpy(), ip),
            shape=(self.N, G))
