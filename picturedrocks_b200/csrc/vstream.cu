// vstream.cu — the hot loop for K <= 5 bins on the "V layout": 4 instructions per entry.
//
// Replaces the per-gene two-pointer merge of _sparse_entropy_wrt
// (picturedrocks/markers/mutualinformation/infoset.py:365-391) for the selectors' calls
// entropy_wrt([j]) and entropy_wrt([-1, j]) (iterative.py:53, 55); the class-conditional
// count tables behind entropy_wrt([]) / entropy_wrt([-1]) (iterative.py:30, 32) fall out
// of the layout's own segment lengths (no streaming pass at all).
//
// V layout (a private re-ordering of the packed by-gene CSC; a histogram does not depend
// on the order in which cells are visited).  Cells are relabelled so that classes are
// contiguous (newpos, a stable sort of the labels; only the V layout and the per-cell
// vectors of this path live in the new order) and cut into SEGMENTS (a class, or a piece of one that fits a shared-memory tile).
// Every gene g is split into K-1 VIRTUAL GENES (g, v), one per non-zero code v; the
// entries of a virtual gene are grouped by segment (in arbitrary order inside a segment), and every
// (g, v, segment) run is padded to a multiple of 32 entries (one 128-byte ROW, one entry
// per lane).  Inside a run the class AND the candidate's code are known, so the joint
// histogram of (y, x_j, x_i) restricted to it has only K-1 <= 4 bins (the value of the
// selected gene x_j): they live in ONE register as four 8-bit fields, and an entry costs
//     AND (cell id mod tile)  ->  LDS.U8 (shift code of x_j)  ->  SHL  ->  ADD.
// At the end of a run the fields are summed over the warp (5 shuffles) and added to
// hits[g][class][x_j-1][v-1].  No shared-memory atomics, no per-entry decode of the code.
//
// HBM-bound by design: 4 bytes per stored entry (+ the row padding), read exactly once
// through a per-warp ring of TMA bulk copies (cp.async.bulk + mbarrier); one persistent
// CTA per SM; (super-tile, fixed-size row range) work items claimed dynamically per warp.
#include <stdlib.h>

#include "vcommon.cuh"

namespace prb {

constexpr int V_MAX_SEGS = 31;  // segments per super-tile (32 boundaries = one warp load)

template <int SPAN, int STAGE_ROWS, int STAGES, int WARPS>
struct VCfg {
    static constexpr int span = SPAN, stage_rows = STAGE_ROWS, stages = STAGES, warps = WARPS;
    static constexpr int threads = WARPS * 32;
    static constexpr int stage_bytes = STAGE_ROWS * 128;
    static constexpr int ring_bytes = WARPS * STAGES * stage_bytes;
    static constexpr int ctl_bytes = 256;  // s_flag + per-segment class offsets
    static constexpr int smem_bytes = SPAN + ring_bytes + WARPS * STAGES * 8 + ctl_bytes;
};

// the two 16-bit entries of a V16 word
__device__ __forceinline__ uint32_t v_pair(uint32_t w)
{
    return v_shl1(v_probe(w & 0xFFFFu)) + v_shl1(v_probe(w >> 16));
}

// Sum the four 8-bit fields of acc over the warp (two REDUX.SUM on 16-bit field pairs: a
// per-lane field is <= 255, so 32 lanes fit 16 bits) and add field a to out[a*K1].
// woff = lane*K1 for the writer lanes 0..K1-1, -1 elsewhere.
__device__ __forceinline__ void v_flush(uint32_t &acc, int32_t *out, int woff, int lane)
{
    const uint32_t s02 = __reduce_add_sync(0xffffffffu, acc & 0x00FF00FFu);         // fields 0, 2
    const uint32_t s13 = __reduce_add_sync(0xffffffffu, (acc >> 8) & 0x00FF00FFu);  // fields 1, 3
    acc = 0;
    const uint32_t sv = (lane & 1) ? s13 : s02;
    const uint32_t cnt = (lane & 2) ? (sv >> 16) : (sv & 0xFFFFu);
    if (woff >= 0 && cnt) atomicAdd(out + woff, (int)cnt);
}

// Segment plan (host-built, see engine.py):
//   seg_cell0 int32[n_seg+1] : cell boundaries of the segments (ascending, last = N)
//   seg_class int32[n_seg]   : class of each segment, or NULL (everything counts as class 0)
//   st_seg0   int32[n_st+1]  : super-tile t = segments [st_seg0[t], st_seg0[t+1]), <= 31 of
//                              them, spanning <= SPAN - 32 cells
// V layout, TILE-MAJOR: rows (32 entries, 128 bytes) ordered by (super-tile, virtual gene,
// segment), so the virtual genes of one work item are ONE contiguous row range:
//   run_end   int32[PV][n_seg] : row just past run (vg, s);  PV = P*K1
//   tile_row0 int32[n_st+1]    : first row of every super-tile's block
// Work items: consecutive row ranges of a super-tile's block, large first and small last
// (the tail of a launch is one item long);
//   item_off  int32[n_st+1]        : first item id of every super-tile
//   item_row0 int32[n_items+n_st]  : first row of every item; after the items of tile t comes
//                                    one sentinel (= end of the block), i.e. item k of tile t
//                                    is entry item_off[t] + t + k
//   item_vg0  int32[n_items]       : first virtual gene with a run ending past the item's first row
// Padding words of a run point at the cell slot just past their super-tile, which the tile
// loader sets to "x_j = 0", so every row is counted the same way (no lane masks).
// sx: uint8[N] shift codes 8*(x_j-1), >= 32 where x_j = 0 (prb_state_v).
// hits: int32[P][C][K1][K1], added to.
// EPW = entries per 32-bit word of the layout: 1 (word = code<<24 | cell, of which this kernel
// only reads the cell's low 16 bits) or 2 (16-bit "V16" layout, the default of this kernel (PRB_V16=0: 32-bit entries): a row holds
// 64 entries of 16 bits, the cell id inside its super-tile; half the bytes per entry).
template <typename CFG, int EPW>
__global__ void __launch_bounds__(CFG::threads, 1)
k_stream_v(const uint32_t *__restrict__ vpacked, const int32_t *__restrict__ run_end,
           const int32_t *__restrict__ tile_row0, int n_seg, int64_t PV, int K1,
           const int32_t *__restrict__ st_seg0, const int32_t *__restrict__ seg_cell0,
           const int32_t *__restrict__ seg_class, int n_st, const uint8_t *__restrict__ sx, int C,
           int32_t *__restrict__ hits, int *counters, const int32_t *__restrict__ item_off,
           const int32_t *__restrict__ item_row0, const int32_t *__restrict__ item_vg0,
           int help_min)
{
    constexpr int SPAN = CFG::span, SR = CFG::stage_rows, STAGES = CFG::stages;
    constexpr int STAGE_BYTES = CFG::stage_bytes, THREADS = CFG::threads;
    constexpr int FB = 255 / EPW;  // rows after which a lane's 8-bit fields could overflow
    // The kernel has NO static shared memory, so the dynamic region starts at a fixed offset of
    // the CTA's shared window and the tile (its first SPAN bytes) is addressed by the bare
    // cell index: a probe is LDS.U16 (index from the ring) -> LDS.U8 [index + imm].
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *s_ctl = (int *)(smem_raw + SPAN + CFG::ring_bytes + CFG::warps * STAGES * 8);
    int &s_flag = s_ctl[0];
    int *s_cls = s_ctl + 1;  // [V_MAX_SEGS + 1]
    uint8_t *st = (uint8_t *)smem_raw;
    if (v_smem_u32(smem_raw) != V_SMEM_BASE) __trap();  // layout assumption of v_probe()
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t *ring = (const uint32_t *)(smem_raw + SPAN + warp * (STAGES * STAGE_BYTES)) + lane;
    const uint32_t ring_s = v_smem_u32(smem_raw + SPAN + warp * (STAGES * STAGE_BYTES));
    const uint32_t bar_s = v_smem_u32(smem_raw + SPAN + CFG::ring_bytes) + warp * (STAGES * 8);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) v_mbar_init(bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t parity = 0;  // bit i: phase parity to wait for on barrier i
    const int KK = K1 * K1;
    const int woff = lane < K1 ? lane * K1 : -1;  // writer lanes of v_flush
    uint32_t acc = 0;

    // A CTA starts on its home super-tile (blockIdx mod n_st) and drains it together with the
    // other CTAs that call it home.  Afterwards it helps the super-tile with the most
    // unclaimed items — but only when that is worth staging another 64 KB tile for
    // (>= help_min items left), or when the tile has no home CTA at all (small grids).
    for (bool first = true;; first = false) {
        __syncthreads();  // every warp has left the previous super-tile
        if (warp == 0) {
            int best = -1, best_t = -1;
            if (first) {
                best_t = (int)(blockIdx.x % (unsigned)n_st);
                best = 1;
            } else {
                for (int tt = lane; tt < n_st; tt += 32) {
                    const int rem = item_off[tt + 1] - item_off[tt] - *(volatile int *)(counters + tt);
                    const bool homeless = tt >= (int)gridDim.x;
                    if (rem > 0 && (rem >= help_min || homeless) && rem > best) {
                        best = rem;
                        best_t = tt;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const int b2 = __shfl_xor_sync(0xffffffffu, best, o);
                    const int t2 = __shfl_xor_sync(0xffffffffu, best_t, o);
                    if (b2 > best || (b2 == best && t2 > best_t)) {
                        best = b2;
                        best_t = t2;
                    }
                }
            }
            if (lane == 0) s_flag = best > 0 ? best_t : -1;
        }
        __syncthreads();
        const int t = s_flag;
        if (t < 0) break;
        const int item0 = item_off[t], n_items = item_off[t + 1] - item0;
        const int s0 = st_seg0[t], s1 = st_seg0[t + 1];
        const int nsg = s1 - s0;
        if (tid < nsg) s_cls[tid] = (seg_class ? seg_class[s0 + tid] : 0) * KK;
        const int c1 = seg_cell0[s1];
        {
            // shift codes of the super-tile: global -> shared in 16-byte chunks from the aligned
            // address below its first cell (the buffer is padded), at (cell id mod SPAN)
            const int64_t c0 = seg_cell0[s0];
            const int64_t a0 = c0 & ~int64_t(15);
            const int nvec = (int)((c1 - a0 + 15) >> 4);
            const uint4 *src = (const uint4 *)(sx + a0);
            const int v0 = (int)(a0 >> 4);
            uint4 *dst = (uint4 *)smem_raw;
            for (int i = tid; i < nvec; i += THREADS)
                dst[(v0 + i) & (SPAN / 16 - 1)] = __ldg(src + i);
        }
        __syncthreads();
        if (tid == 0) st[c1 & (SPAN - 1)] = PRB_SX_NONE;  // the slot padding words point at: x_j = 0
        __syncthreads();
        // work items are fixed-size row ranges of the super-tile's block (balanced whatever
        // the density of a gene); the next one is claimed while the current one is counted
        int claim = lane == 0 ? atomicAdd(counters + t, 1) : 0;
        for (;;) {
            const int item = __shfl_sync(0xffffffffu, claim, 0);
            if (item >= n_items) break;
            if (lane == 0) claim = atomicAdd(counters + t, 1);
            // rows [row_begin, row_begin + nrows) of the layout (item t's own entry of
            // item_row0 is followed by a sentinel = end of the tile's block)
            const int row_begin = __ldg(item_row0 + item0 + t + item);
            const int nrows = __ldg(item_row0 + item0 + t + item + 1) - row_begin;
            const int nstg = (nrows + SR - 1) / SR;
            // stage b of the item lives in ring slot b % STAGES; row 0 <-> ring slot 0 (all
            // copies of the previous item were consumed)
            const uint32_t *src_next = vpacked + (int64_t)row_begin * 32;  // next stage to issue
            int rows_unissued = nrows;
            int issue_slot = 0;
            auto issue = [&]() {
                const int rows = min(SR, rows_unissued);
                v_bulk_load(ring_s + issue_slot * STAGE_BYTES, src_next, (uint32_t)rows * 128u,
                            bar_s + 8 * issue_slot);
                src_next += SR * 32;
                rows_unissued -= SR;
                issue_slot = issue_slot + 1 == STAGES ? 0 : issue_slot + 1;
            };
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < STAGES; ++i)
                    if (i < nstg) issue();
            }
            // first virtual gene with a run that ends past the item's first row
            const int64_t vg_a = __ldg(item_vg0 + item0 + item);
            int wait_slot = 0;  // ring slot of the next stage to wait for
            int r = 0;          // next row to count (relative to row_begin)
            int limit = 0;      // rows below it are in shared memory (next stage boundary)
            int fb = FB;        // row at which the 8-bit fields could overflow
            const uint32_t *rp = ring;  // lane's word of row r
            int64_t g = vg_a / K1;
            int v = (int)(vg_a - g * K1);
            int32_t ends_next = lane < nsg ? __ldg(run_end + vg_a * n_seg + s0 + lane) : 0;
#pragma unroll 1
            for (int64_t vg = vg_a; r < nrows; ++vg) {
                const int32_t ends = ends_next - row_begin;
                if (vg + 1 < PV && lane < nsg)
                    ends_next = __ldg(run_end + (vg + 1) * n_seg + s0 + lane);
                int32_t *og = hits + g * (int64_t)C * KK + v;
                if (++v == K1) {
                    v = 0;
                    ++g;
                }
#pragma unroll 1
                for (int seg = 0; seg < nsg; ++seg) {
                    // the part of run (vg, seg) inside the item: rows [r, e)
                    const int e = min(__shfl_sync(0xffffffffu, ends, seg), nrows);
                    if (e <= r) continue;  // empty, or entirely before the item
                    int32_t *oseg = og + s_cls[seg];
                    for (;;) {
                        const int stop = min(min(e, limit), fb);
                        const uint32_t *rp_end = rp + (stop - r) * 32;
                        if constexpr (EPW == 1) {
#pragma unroll 1
                            for (; rp + 96 < rp_end; rp += 128) {
                                const uint32_t w0 = rp[0], w1 = rp[32], w2 = rp[64], w3 = rp[96];
                                const uint32_t x0 = v_probe(w0 & (SPAN - 1)), x1 = v_probe(w1 & (SPAN - 1));
                                const uint32_t x2 = v_probe(w2 & (SPAN - 1)), x3 = v_probe(w3 & (SPAN - 1));
                                acc += v_shl1(x0) + v_shl1(x1) + v_shl1(x2) + v_shl1(x3);
                            }
#pragma unroll 1
                            for (; rp < rp_end; rp += 32) acc += v_shl1(v_probe(rp[0] & (SPAN - 1)));
                        } else {
                            static_assert(EPW == 1 || SPAN == 65536, "16-bit entries index a 64 K tile");
#pragma unroll 1
                            for (; rp + 96 < rp_end; rp += 128) {
                                const uint32_t w0 = rp[0], w1 = rp[32], w2 = rp[64], w3 = rp[96];
                                acc += v_pair(w0) + v_pair(w1) + v_pair(w2) + v_pair(w3);
                            }
#pragma unroll 1
                            for (; rp < rp_end; rp += 32) acc += v_pair(rp[0]);
                        }
                        r = stop;
                        if (r == e) break;
                        if (r == fb) {  // the 8-bit fields are full
                            v_flush(acc, oseg, woff, lane);
                            fb = r + FB;
                        }
                        if (r == limit) {
                            // ring event at a stage boundary: hand the stage just counted back
                            // to the copy engine, wait for the stage that holds row r
                            if (r) {
                                __syncwarp();  // every lane has read the stage
                                if (lane == 0 && rows_unissued > 0) issue();
                            }
                            v_mbar_wait(bar_s + 8 * wait_slot, (parity >> wait_slot) & 1u);
                            parity ^= 1u << wait_slot;
                            rp = ring + wait_slot * (SR * 32);  // row r opens this stage
                            wait_slot = wait_slot + 1 == STAGES ? 0 : wait_slot + 1;
                            limit = min(nrows, r + SR);
                        }
                    }
                    v_flush(acc, oseg, woff, lane);  // hits are added to: partial runs are fine
                    fb = r + FB;
                }
            }
            __syncwarp();  // the ring is free for the next item
        }
    }
}

// ---- building the V layout -------------------------------------------------------------
// Both passes: one block per gene (dynamic queue), the gene's entries read with several
// independent loads per thread in flight, per-(code, segment) counters of the gene in shared
// memory.  Per-cell table cellmap uint32[N] (indexed by the ORIGINAL cell id): bits 0..23 =
// relabelled cell position newpos[cell], bits 24..31 = segment of the cell when there are at
// most 256 segments — ONE gather per entry gives both; with more segments seg16 uint16[N]
// carries the segment instead.  cellmap == NULL: identity positions, a single segment.
// (Both passes are bound by the number of distinct cache lines an entry touches: the table
// gather and, in pass 2, the store.)
constexpr int VL_WARPS = 8, VL_THREADS = VL_WARPS * 32, VL_U = 4;

__device__ __forceinline__ int vl_seg(uint32_t cm, uint32_t cell, const uint16_t *seg16)
{
    return seg16 ? (int)seg16[cell] : (int)(cm >> 24);
}

// pass 1: entries per (code, segment) -> seglen[P*K1][n_seg]
__global__ void __launch_bounds__(VL_THREADS)
k_vcount(const uint32_t *__restrict__ packed, const int64_t *__restrict__ indptr, int64_t P,
         const uint32_t *__restrict__ cellmap, const uint16_t *__restrict__ seg16, int n_seg,
         int K1, int32_t *__restrict__ seglen, int *queue)
{
    extern __shared__ int vb_cnt[];  // [nk]
    __shared__ int s_gene;
    const int tid = threadIdx.x;
    const int nk = K1 * n_seg;
    for (;;) {
        __syncthreads();  // the previous gene's counters were written out
        if (tid == 0) s_gene = atomicAdd(queue, 1);
        for (int i = tid; i < nk; i += VL_THREADS) vb_cnt[i] = 0;
        __syncthreads();
        const int64_t g = s_gene;
        if (g >= P) break;
        const int64_t b = indptr[g], e = indptr[g + 1];
        for (int64_t i0 = b + tid; i0 < e; i0 += VL_THREADS * VL_U) {
            uint32_t w[VL_U];
#pragma unroll
            for (int u = 0; u < VL_U; ++u) {
                const int64_t i = i0 + u * VL_THREADS;
                w[u] = i < e ? packed[i] : 0u;  // a stored entry has a non-zero code byte
            }
            int seg[VL_U];
#pragma unroll
            for (int u = 0; u < VL_U; ++u) {
                const uint32_t cell = w[u] & ROW_MASK;
                seg[u] = (w[u] && n_seg > 1) ? vl_seg(seg16 ? 0u : cellmap[cell], cell, seg16) : 0;
            }
#pragma unroll
            for (int u = 0; u < VL_U; ++u)
                if (w[u]) atomicAdd(vb_cnt + ((w[u] >> 24) - 1u) * n_seg + seg[u], 1);
        }
        __syncthreads();
        for (int i = tid; i < nk; i += VL_THREADS) seglen[g * nk + i] = vb_cnt[i];
    }
}

// pass 2: scatter of the relabelled entries (word = code | newpos[cell]) into the padded
// runs.  Run (vg, s) occupies rows [run_end - ceil(len/32), run_end).  A gene is processed in
// chunks of VS_CHUNK entries that are counting-sorted by (code, segment) in shared memory
// (slot inside the chunk's group = old value of a shared counter, so the order of the
// entries inside a run is arbitrary — the streaming kernel only counts them) and leave the
// chunk group by group: consecutive threads write consecutive words.  The padding words of
// segment s are seg_pad[s] = 0xFF000000 | (a cell slot the kernel keeps at "x_j = 0").
// E16: the V16 layout — rows of 64 entries, an entry is the low
// 16 bits of the word (the cell id inside its super-tile), runs padded to 64 entries.
constexpr int VS_U = 8, VS_CHUNK = VL_THREADS * VS_U;

template <bool E16>
__global__ void __launch_bounds__(VL_THREADS)
k_vscatter(const uint32_t *__restrict__ packed, const int64_t *__restrict__ indptr, int64_t P,
           const uint32_t *__restrict__ cellmap, const uint16_t *__restrict__ seg16, int n_seg,
           int K1, const int32_t *__restrict__ seglen, const int32_t *__restrict__ run_end,
           const uint32_t *__restrict__ seg_pad, uint32_t *__restrict__ vpacked, int *queue,
           const int64_t *__restrict__ run_start)
{
    // run_start != NULL (E16 only): FLAT output for vell.cu — run (vg, s) starts at entry
    // run_start[vg * n_seg + s] of a 16-bit array, runs back to back, no padding
    extern __shared__ __align__(8) unsigned char vs_raw[];
    __shared__ int s_gene;
    __shared__ int s_wsum[VL_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nk = K1 * n_seg;
    const int per = (nk + VL_THREADS - 1) / VL_THREADS;  // keys per thread in the scan
    int64_t *next = (int64_t *)vs_raw;       // [nk] next free slot of every run
    int *ccnt = (int *)(next + nk);          // [nk] entries of the chunk per key
    int *cstart = ccnt + nk;                 // [nk] first slot of the key inside the chunk
    uint32_t *stage = (uint32_t *)(cstart + nk);       // [VS_CHUNK] words, grouped by key
    uint16_t *skey = (uint16_t *)(stage + VS_CHUNK);   // [VS_CHUNK] key of every staged word
    for (;;) {
        __syncthreads();  // the previous gene's padding pass is done with next[]
        if (tid == 0) s_gene = atomicAdd(queue, 1);
        __syncthreads();
        const int64_t g = s_gene;
        if (g >= P) break;
        const int64_t b = indptr[g], e = indptr[g + 1];
        constexpr int EPR_LOG = E16 ? 6 : 5, EPR = 1 << EPR_LOG;  // entries per 128-byte row
        for (int i = tid; i < nk; i += VL_THREADS) {
            if (run_start) {
                next[i] = run_start[g * nk + i];
            } else {
                const int nr = (seglen[g * nk + i] + EPR - 1) >> EPR_LOG;
                next[i] = (int64_t)(run_end[g * nk + i] - nr) * EPR;
            }
            ccnt[i] = 0;
        }
        __syncthreads();
        for (int64_t c0 = b; c0 < e; c0 += VS_CHUNK) {
            const int n_chunk = (int)min((int64_t)VS_CHUNK, e - c0);
            // A: key of every entry, its slot inside the chunk's group of that key
            uint32_t w[VS_U];
#pragma unroll
            for (int u = 0; u < VS_U; ++u) {
                const int i = u * VL_THREADS + tid;
                w[u] = i < n_chunk ? packed[c0 + i] : 0u;  // a stored entry has a non-zero code byte
            }
            uint32_t word[VS_U];
            int key[VS_U], rk[VS_U];
#pragma unroll
            for (int u = 0; u < VS_U; ++u) {
                const uint32_t cell = w[u] & ROW_MASK;
                const uint32_t cm = (w[u] && cellmap) ? cellmap[cell] : cell;
                const int seg = (w[u] && n_seg > 1) ? vl_seg(cm, cell, seg16) : 0;
                key[u] = (int)((w[u] >> 24) - 1u) * n_seg + seg;
                word[u] = (w[u] & ~ROW_MASK) | (cm & ROW_MASK);
            }
#pragma unroll
            for (int u = 0; u < VS_U; ++u)
                rk[u] = w[u] ? atomicAdd(ccnt + key[u], 1) : 0;
            __syncthreads();
            // exclusive scan of the chunk's counts over the keys (thread t: keys [t*per, ..))
            {
                int sum = 0;
                for (int k = tid * per; k < min(nk, (tid + 1) * per); ++k) sum += ccnt[k];
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_wsum[warp] = inc;
                __syncthreads();
                int run = inc - sum;
                for (int wv = 0; wv < warp; ++wv) run += s_wsum[wv];
                for (int k = tid * per; k < min(nk, (tid + 1) * per); ++k) {
                    cstart[k] = run;
                    run += ccnt[k];
                }
            }
            __syncthreads();
            // B: words into the staging buffer, grouped by key
#pragma unroll
            for (int u = 0; u < VS_U; ++u)
                if (w[u]) {
                    const int p = cstart[key[u]] + rk[u];
                    stage[p] = word[u];
                    skey[p] = (uint16_t)key[u];
                }
            __syncthreads();
            // C: out, group by group (consecutive slots of a key -> consecutive words)
            for (int p = tid; p < n_chunk; p += VL_THREADS) {
                const int k = skey[p];
                if (E16)
                    ((uint16_t *)vpacked)[next[k] + (p - cstart[k])] = (uint16_t)stage[p];
                else
                    vpacked[next[k] + (p - cstart[k])] = stage[p];
            }
            __syncthreads();
            for (int k = tid; k < nk; k += VL_THREADS) {
                next[k] += ccnt[k];
                ccnt[k] = 0;
            }
            __syncthreads();
        }
        // padding up to the end of the run's last row
        if (!run_start)
        for (int i = tid; i < nk; i += VL_THREADS) {
            const int s = i % n_seg;
            const int64_t end = (int64_t)run_end[g * nk + i] * EPR;
            const uint32_t padw = seg_pad[s];
            for (int64_t p = next[i]; p < end; ++p) {
                if (E16)
                    ((uint16_t *)vpacked)[p] = (uint16_t)padw;
                else
                    vpacked[p] = padw;
            }
        }
    }
}

}  // namespace prb

using namespace prb;

// Kernel configurations (cells per super-tile, rows per stage, stages per warp, warps per
// CTA); all fit the 227 KB of shared memory.  PRB_V_CFG selects one for tuning runs.
using VCfg0 = VCfg<1 << 16, 20, 2, 32>;  //  64 KB of shift codes + 32 warps x 2 x 2.5 KB
using VCfg1 = VCfg<1 << 16, 16, 2, 32>;  //  64 KB + 32 x 2 x 2 KB
using VCfg2 = VCfg<1 << 16, 26, 2, 24>;  //  64 KB + 24 x 2 x 3.25 KB
using VCfg3 = VCfg<1 << 16, 13, 3, 32>;  //  64 KB + 32 x 3 x 1.625 KB

static int v_cfg()
{
    static const int cfg = [] {
        const char *e = getenv("PRB_V_CFG");
        const int v = e ? atoi(e) : 0;
        return v >= 0 && v <= 3 ? v : 0;
    }();
    return cfg;
}

// items a super-tile must still have unclaimed for a CTA of another tile to come and help
// (staging a 64 KB tile has to pay off); PRB_V_HELP_MIN overrides it for tuning runs
static int v_help_min(int warps)
{
    const char *e = getenv("PRB_V_HELP_MIN");
    const int forced = e ? atoi(e) : 0;
    return forced > 0 ? forced : 8 * warps;
}

static int v_span()
{
    switch (v_cfg()) {
        case 1: return VCfg1::span;
        case 2: return VCfg2::span;
        case 3: return VCfg3::span;
        default: return VCfg0::span;
    }
}

extern "C" int prb_v_tile_span(void) { return v_span(); }

// warps of one launch of the active configuration (to size the work items)
extern "C" int prb_v_warps(void)
{
    int w = VCfg0::warps;
    switch (v_cfg()) {
        case 1: w = VCfg1::warps; break;
        case 2: w = VCfg2::warps; break;
        case 3: w = VCfg3::warps; break;
    }
    return w * devinfo().sm_count;
}

extern "C" int prb_v_tile_capacity(int64_t *max_cells, int *max_segs)
{
    // a super-tile's shift codes sit in shared memory at (cell id mod SPAN); 16 cells of
    // slack for the aligned 16-byte copies + the slot the padding words point at
    if (max_cells) *max_cells = v_span() - 32;
    if (max_segs) *max_segs = V_MAX_SEGS;
    return 0;
}

extern "C" int prb_v_count(const uint32_t *packed, const int64_t *indptr, int64_t P,
                           const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                           int32_t *seglen,
                           int32_t *queue, prb_stream_t stream)
{
    if (P == 0) return 0;
    PRB_REQUIRE(K1 >= 1 && K1 <= 20 && n_seg >= 1, "bad arguments");
    PRB_REQUIRE(n_seg == 1 || seg16 != nullptr || (cellmap != nullptr && n_seg <= 256),
                "the segment of every cell is required (cellmap bits 24..31, or seg16)");
    PRB_REQUIRE(n_seg <= 65536, "too many segments");
    const size_t smem = (size_t)K1 * n_seg * 4;
    PRB_REQUIRE((int64_t)smem <= devinfo().smem_optin - 1024, "too many segments");
    PRB_CUDA(cudaFuncSetAttribute(k_vcount, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PRB_CUDA(cudaMemsetAsync(queue, 0, sizeof(int32_t), S(stream)));
    const int grid = (int)hmin(P, (int64_t)devinfo().sm_count * 8);
    k_vcount<<<grid, VL_THREADS, smem, S(stream)>>>(packed, indptr, P, cellmap, seg16, n_seg, K1,
                                                    seglen, queue);
    PRB_LAUNCH_CHECK("k_vcount");
    return 0;
}

template <bool E16>
static int v_scatter_impl(const uint32_t *packed, const int64_t *indptr, int64_t P,
                          const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                          const int32_t *seglen, const int32_t *run_end, const uint32_t *seg_pad,
                          uint32_t *vpacked, int32_t *queue, prb_stream_t stream,
                          const int64_t *run_start = nullptr)
{
    if (P == 0) return 0;
    PRB_REQUIRE(K1 >= 1 && K1 <= (run_start ? 20 : 4) && n_seg >= 1, "bad arguments");
    PRB_REQUIRE(n_seg == 1 || seg16 != nullptr || (cellmap != nullptr && n_seg <= 256),
                "the segment of every cell is required (cellmap bits 24..31, or seg16)");
    PRB_REQUIRE(n_seg <= 65536 / K1, "too many segments");
    const size_t nk = (size_t)K1 * n_seg;
    const size_t smem = nk * 16 + (size_t)VS_CHUNK * 6;
    PRB_REQUIRE((int64_t)smem <= devinfo().smem_optin - 1024, "too many segments");
    PRB_CUDA(cudaFuncSetAttribute(k_vscatter<E16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    PRB_CUDA(cudaMemsetAsync(queue, 0, sizeof(int32_t), S(stream)));
    const int grid = (int)hmin(P, (int64_t)devinfo().sm_count * 8);
    k_vscatter<E16><<<grid, VL_THREADS, smem, S(stream)>>>(packed, indptr, P, cellmap, seg16, n_seg,
                                                           K1, seglen, run_end, seg_pad, vpacked,
                                                           queue, run_start);
    PRB_LAUNCH_CHECK("k_vscatter");
    return 0;
}

extern "C" int prb_v_scatter(const uint32_t *packed, const int64_t *indptr, int64_t P,
                             const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                             const int32_t *seglen, const int32_t *run_end,
                             const uint32_t *seg_pad, uint32_t *vpacked, int32_t *queue,
                             prb_stream_t stream)
{
    return v_scatter_impl<false>(packed, indptr, P, cellmap, seg16, n_seg, K1, seglen, run_end,
                                 seg_pad, vpacked, queue, stream);
}

// V16 layout: run_end counts rows of 64 entries.
extern "C" int prb_v_scatter16(const uint32_t *packed, const int64_t *indptr, int64_t P,
                               const uint32_t *cellmap, const uint16_t *seg16, int n_seg, int K1,
                               const int32_t *seglen, const int32_t *run_end,
                               const uint32_t *seg_pad, uint32_t *vpacked, int32_t *queue,
                               prb_stream_t stream)
{
    return v_scatter_impl<true>(packed, indptr, P, cellmap, seg16, n_seg, K1, seglen, run_end,
                                seg_pad, vpacked, queue, stream);
}

// FLAT variant for the lane-run layout (vell.cu): relabelled 16-bit entries (the cell id inside
// its super-tile) of run (vg, s) at tmp16[run_start[vg * n_seg + s] ..), in arbitrary order.
extern "C" int prb_v_scatter_flat(const uint32_t *packed, const int64_t *indptr, int64_t P,
                                  const uint32_t *cellmap, const uint16_t *seg16, int n_seg,
                                  int K1, const int32_t *seglen, const int64_t *run_start,
                                  uint16_t *tmp16, int32_t *queue, prb_stream_t stream)
{
    PRB_REQUIRE(run_start != nullptr && tmp16 != nullptr, "null pointer");
    return v_scatter_impl<true>(packed, indptr, P, cellmap, seg16, n_seg, K1, seglen, nullptr,
                                nullptr, (uint32_t *)tmp16, queue, stream, run_start);
}

template <int EPW>
static int stream_v_impl(const uint32_t *vpacked, const int32_t *run_end, const int32_t *tile_row0,
                         int n_seg, int64_t P, int K1, int64_t v_rows, const int32_t *st_seg0,
                         const int32_t *seg_cell0, const int32_t *seg_class, int n_st,
                         int64_t max_tile_cells, const uint8_t *sx, int C, int32_t *hits,
                         int32_t *counters, const int32_t *item_off, const int32_t *item_row0,
                         const int32_t *item_vg0, int64_t n_items, int counters_clean,
                         prb_stream_t stream)
{
    if (P == 0 || v_rows == 0) return 0;
    PRB_REQUIRE(K1 >= 1 && K1 <= 4, "the register-counter path needs K <= 5");
    PRB_REQUIRE(n_seg >= 1 && n_st >= 1 && C >= 1 && sx != nullptr, "bad segment plan");
    PRB_REQUIRE(item_off && item_row0 && item_vg0 && n_items >= 1, "bad item plan");
    PRB_REQUIRE(max_tile_cells <= v_span() - 32, "super-tile spans too many cells");
    if (!counters_clean)  // (prb_pick_scatter zeroes them in the greedy step)
        PRB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * n_st, S(stream)));
    const int64_t PV = P * K1;
    auto launch = [&](auto kern, int smem, int warps) -> int {
        const int grid = (int)hmax(1, hmin(devinfo().sm_count, ceil_div(n_items, warps)));
        PRB_REQUIRE(smem + 512 <= devinfo().smem_optin, "not enough shared memory per block");
        PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, warps * 32, smem, S(stream)>>>(vpacked, run_end, tile_row0, n_seg, PV, K1,
                                                    st_seg0, seg_cell0, seg_class, n_st, sx, C, hits,
                                                    counters, item_off, item_row0, item_vg0,
                                                    v_help_min(warps));
        return 0;
    };
    switch (v_cfg()) {
        case 1:
            if (launch(k_stream_v<VCfg1, EPW>, VCfg1::smem_bytes, VCfg1::warps)) return 1;
            break;
        case 2:
            if (launch(k_stream_v<VCfg2, EPW>, VCfg2::smem_bytes, VCfg2::warps)) return 1;
            break;
        case 3:
            if (launch(k_stream_v<VCfg3, EPW>, VCfg3::smem_bytes, VCfg3::warps)) return 1;
            break;
        default:
            if (launch(k_stream_v<VCfg0, EPW>, VCfg0::smem_bytes, VCfg0::warps)) return 1;
    }
    PRB_LAUNCH_CHECK("k_stream_v");
    return 0;
}

extern "C" int prb_stream_v(const uint32_t *vpacked, const int32_t *run_end,
                            const int32_t *tile_row0, int n_seg, int64_t P, int K1,
                            int64_t v_rows, const int32_t *st_seg0, const int32_t *seg_cell0,
                            const int32_t *seg_class, int n_st, int64_t max_tile_cells,
                            const uint8_t *sx, int C, int32_t *hits, int32_t *counters,
                            const int32_t *item_off, const int32_t *item_row0,
                            const int32_t *item_vg0, int64_t n_items, int counters_clean,
                            prb_stream_t stream)
{
    return stream_v_impl<1>(vpacked, run_end, tile_row0, n_seg, P, K1, v_rows, st_seg0, seg_cell0,
                            seg_class, n_st, max_tile_cells, sx, C, hits, counters, item_off,
                            item_row0, item_vg0, n_items, counters_clean, stream);
}

// The same pass over the V16 layout built by
// prb_v_scatter16 (rows of 64 16-bit entries).
extern "C" int prb_stream_v16(const uint32_t *vpacked, const int32_t *run_end,
                              const int32_t *tile_row0, int n_seg, int64_t P, int K1,
                              int64_t v_rows, const int32_t *st_seg0, const int32_t *seg_cell0,
                              const int32_t *seg_class, int n_st, int64_t max_tile_cells,
                              const uint8_t *sx, int C, int32_t *hits, int32_t *counters,
                              const int32_t *item_off, const int32_t *item_row0,
                              const int32_t *item_vg0, int64_t n_items, int counters_clean,
                              prb_stream_t stream)
{
    PRB_REQUIRE(v_span() == 65536, "16-bit entries need 64 K-cell tiles");
    return stream_v_impl<2>(vpacked, run_end, tile_row0, n_seg, P, K1, v_rows, st_seg0, seg_cell0,
                            seg_class, n_st, max_tile_cells, sx, C, hits, counters, item_off,
                            item_row0, item_vg0, n_items, counters_clean, stream);
}
