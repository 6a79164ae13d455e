// vell.cu — the hot loop on the "lane-run" layout: one lane = one run, no per-run bookkeeping.
//
// Same job as vstream.cu — the per-gene two-pointer merge of _sparse_entropy_wrt
// (picturedrocks/markers/mutualinformation/infoset.py:365-391) for the selectors' calls
// entropy_wrt([j]) / entropy_wrt([-1, j]) (iterative.py:53, 55) — on a layout whose cost does
// not depend on how short the runs are (small N, many classes, gene shards of a multi-GPU run,
// K = 10 / 20 bins), where the warp-per-run kernel of vstream.cu spends most of its
// instructions on run boundaries.
//
// As in vstream.cu the cells are relabelled so that classes are contiguous, cut into segments
// and super-tiles, and every gene g is split into K-1 virtual genes (g, v).  A LANE-RUN is a
// piece of at most CAP entries of one run (virtual gene vg, segment s): inside it the class y
// and the candidate's code v are fixed, so the joint histogram restricted to it has only the
// K-1 values of the selected gene x_j as bins — 8-bit fields of (K-1+3)/4 registers of ONE
// lane.  32 lane-runs of the same super-tile and of similar length form a BUNDLE, stored
// transposed (ELLPACK): row r of a bundle holds entries 2r and 2r+1 of each of its 32 lane-runs
// (a 32-bit word per lane: two 16-bit cell ids inside the super-tile); a separate header row
// per bundle names the runs (per lane: segment-in-tile << 24 | virtual gene, 0xFFFFFFFF =
// empty lane).  So a warp reads
// conflict-free 128-byte rows exactly as before, an entry still costs
//     LDS.U8 (shift code of x_j at that cell)  ->  SHL  ->  ADD,
// but the 32 lanes count for 32 different runs, nothing is reduced across the warp, and the
// whole per-run work is one flush per BUNDLE: every lane adds its <= K-1 counts to
// hits[g][y][x_j-1][v-1] (RED).  Lane-runs are sorted by length inside a super-tile, so the
// padding of a bundle to its longest lane-run is a few per cent.
//
// HBM-bound by design: 2 bytes per stored entry, read exactly once through the same per-warp
// ring of TMA bulk copies (cp.async.bulk + mbarrier) as vstream.cu; one persistent CTA per SM;
// work items = ranges of whole bundles, claimed dynamically per warp.
#include <stdlib.h>

#include "vcommon.cuh"

namespace prb {

constexpr uint32_t ELL_NONE = 0xFFFFFFFFu;

template <int SPAN, int STAGE_ROWS, int STAGES, int WARPS>
struct ECfg {
    static constexpr int span = SPAN, stage_rows = STAGE_ROWS, stages = STAGES, warps = WARPS;
    static constexpr int threads = WARPS * 32;
    static constexpr int stage_bytes = STAGE_ROWS * 128;
    static constexpr int ring_bytes = WARPS * STAGES * stage_bytes;
    static constexpr int ctl_bytes = 256;  // s_flag + per-segment class offsets
    static constexpr int smem_bytes = SPAN + ring_bytes + WARPS * STAGES * 8 + ctl_bytes;
};

constexpr int E_MAX_SEGS = 31;  // segments per super-tile (the plan of vstream.cu)

// Segment plan as in vstream.cu (seg_cell0 / seg_class / st_seg0).  Layout:
//   ell     uint32[32 * rows]  : the data rows of the bundles, back to back
//   bhdr    uint32[n_bundles][32] : header words
//   brow0   int32[n_bundles+1] : first row of every bundle
// Work items: ranges of whole bundles of one super-tile (large first, small last);
//   item_off int32[n_st+1]        : first item id of every super-tile
//   item_b0  int32[n_items+n_st]  : first bundle of every item; after the items of tile t comes
//                                   one sentinel (= first bundle of the next tile), i.e. item k
//                                   of tile t is entry item_off[t] + t + k
// sx: uint8[N] shift codes 8*(x_j-1), PRB_SX_NONE where x_j = 0.  Padding entries point at the
// cell slot just past their super-tile, which the tile loader sets to PRB_SX_NONE.
// hits: int32[P][C][K1][K1], added to.
// NREG = ceil(K1 / 4) counter registers per lane.
template <typename CFG, int NREG>
__global__ void __launch_bounds__(CFG::threads, 1)
k_stream_ell(const uint32_t *__restrict__ ell, const uint32_t *__restrict__ bhdr,
             const int32_t *__restrict__ brow0, int K1,
             const int32_t *__restrict__ st_seg0, const int32_t *__restrict__ seg_cell0,
             const int32_t *__restrict__ seg_class, int n_st, const uint8_t *__restrict__ sx, int C,
             int32_t *__restrict__ hits, int *counters, const int32_t *__restrict__ item_off,
             const int32_t *__restrict__ item_b0, int help_min, uint32_t PV, int32_t *dbg)
{
    constexpr int SPAN = CFG::span, SR = CFG::stage_rows, STAGES = CFG::stages;
    constexpr int STAGE_BYTES = CFG::stage_bytes, THREADS = CFG::threads;
    constexpr int FB = 127;  // rows (2 entries per lane) after which an 8-bit field could overflow
    // NO static shared memory (see v_probe): the tile is the first SPAN bytes of the window
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *s_ctl = (int *)(smem_raw + SPAN + CFG::ring_bytes + CFG::warps * STAGES * 8);
    int &s_flag = s_ctl[0];
    int *s_cls = s_ctl + 1;  // [E_MAX_SEGS + 1]
    uint8_t *st = (uint8_t *)smem_raw;
    if (v_smem_u32(smem_raw) != V_SMEM_BASE) __trap();  // checked by the host wrapper beforehand
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

    // home super-tile first, then help the tile with the most unclaimed items (vstream.cu)
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
        if (tid < s1 - s0) s_cls[tid] = (seg_class ? seg_class[s0 + tid] : 0) * KK;
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
        if (tid == 0) st[c1 & (SPAN - 1)] = PRB_SX_NONE;  // the slot padding entries point at
        __syncthreads();
        int claim = lane == 0 ? atomicAdd(counters + t, 1) : 0;
        for (;;) {
            const int item = __shfl_sync(0xffffffffu, claim, 0);
            if (item >= n_items) break;
            if (lane == 0) claim = atomicAdd(counters + t, 1);
            const int b_begin = __ldg(item_b0 + item0 + t + item);
            const int b_end = __ldg(item_b0 + item0 + t + item + 1);
            const int row_begin = __ldg(brow0 + b_begin);
            const int nrows = __ldg(brow0 + b_end) - row_begin;
            const int nstg = (nrows + SR - 1) / SR;
            const uint32_t *src_next = ell + (int64_t)row_begin * 32;  // next stage to issue
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
            int wait_slot = 0;  // ring slot of the next stage to wait for
            int r = 0;          // next row to read (relative to row_begin)
            int limit = 0;      // rows below it are in shared memory (next stage boundary)
            const uint32_t *rp = ring;  // lane's word of row r
            // ring event at a stage boundary: hand the stage just read back to the copy
            // engine, wait for the stage that holds row r
            auto next_stage = [&]() {
                if (r) {
                    __syncwarp();  // every lane has read the stage
                    if (lane == 0 && rows_unissued > 0) issue();
                }
                v_mbar_wait(bar_s + 8 * wait_slot, (parity >> wait_slot) & 1u);
                parity ^= 1u << wait_slot;
                rp = ring + wait_slot * (SR * 32);
                wait_slot = wait_slot + 1 == STAGES ? 0 : wait_slot + 1;
                limit = min(nrows, r + SR);
            };
            for (int b = b_begin; b < b_end; b += 32) {
                // last rows of the next (up to) 32 bundles, one coalesced load
                const int nb = min(32, b_end - b);
                const int32_t ends = lane < nb ? __ldg(brow0 + b + 1 + lane) - row_begin : 0;
#pragma unroll 1
                for (int i = 0; i < nb; ++i) {
                    const int e = __shfl_sync(0xffffffffu, ends, i);
                    // this lane's run: read from global memory (not through the ring — a word
                    // that is only needed at the flush could still be in flight when its ring
                    // slot is handed back to the copy engine), consumed after the last row
                    const uint32_t hdr = __ldg(bhdr + (int64_t)(b + i) * 32 + lane);
                    uint32_t acc[NREG], wide[2 * NREG];
#pragma unroll
                    for (int q = 0; q < NREG; ++q) acc[q] = wide[2 * q] = wide[2 * q + 1] = 0u;
                    int fb = r + FB;
                    auto count = [&](uint32_t w) {
                        const uint32_t x0 = v_probe(w & 0xFFFFu), x1 = v_probe(w >> 16);
#pragma unroll
                        for (int q = 0; q < NREG; ++q)
                            acc[q] += v_shl1(x0 - 32u * q) + v_shl1(x1 - 32u * q);
                    };
                    while (r < e) {
                        if (r == limit) next_stage();
                        const int stop = min(min(e, limit), fb);
                        const uint32_t *rp_end = rp + (stop - r) * 32;
#pragma unroll 1
                        for (; rp + 96 < rp_end; rp += 128) {
                            const uint32_t w0 = rp[0], w1 = rp[32], w2 = rp[64], w3 = rp[96];
                            count(w0);
                            count(w1);
                            count(w2);
                            count(w3);
                        }
#pragma unroll 1
                        for (; rp < rp_end; rp += 32) count(rp[0]);
                        r = stop;
                        if (r == fb) {  // the 8-bit fields are full: move them to 16-bit ones
#pragma unroll
                            for (int q = 0; q < NREG; ++q) {
                                wide[2 * q] += acc[q] & 0x00FF00FFu;
                                wide[2 * q + 1] += (acc[q] >> 8) & 0x00FF00FFu;
                                acc[q] = 0u;
                            }
                            fb = r + FB;
                        }
                    }
                    if (hdr != ELL_NONE && ((hdr & 0x00FFFFFFu) >= PV || (int)(hdr >> 24) >= s1 - s0)) {
                        // a header that names no run of this tile: never write through it
                        if (dbg && atomicAdd(dbg, 1) == 0) {
                            dbg[1] = (int)blockIdx.x, dbg[2] = warp, dbg[3] = lane, dbg[4] = t;
                            dbg[5] = item, dbg[6] = b + i, dbg[7] = r, dbg[8] = e, dbg[9] = (int)hdr;
                            dbg[10] = row_begin, dbg[11] = nrows, dbg[12] = limit, dbg[13] = b_begin;
                            dbg[14] = b_end, dbg[15] = n_items;
                        }
                    } else if (hdr != ELL_NONE) {
                        const uint32_t vg = hdr & 0x00FFFFFFu;
                        const uint32_t g = vg / (uint32_t)K1, v = vg - g * (uint32_t)K1;
                        int32_t *out = hits + (int64_t)g * C * KK + s_cls[hdr >> 24] + v;
#pragma unroll
                        for (int q = 0; q < NREG; ++q) {
                            const uint32_t w02 = wide[2 * q] + (acc[q] & 0x00FF00FFu);
                            const uint32_t w13 = wide[2 * q + 1] + ((acc[q] >> 8) & 0x00FF00FFu);
                            const uint32_t c[4] = {w02 & 0xFFFFu, w13 & 0xFFFFu, w02 >> 16, w13 >> 16};
#pragma unroll
                            for (int f = 0; f < 4; ++f)
                                if (4 * q + f < K1 && c[f]) atomicAdd(out + (4 * q + f) * K1, (int)c[f]);
                        }
                    }
                }
            }
            __syncwarp();  // the ring is free for the next item
        }
    }
}

// ---- building the layout ---------------------------------------------------------------------
// One warp per bundle; lane l copies lane-run b*32+l (lr_len entries from tmp[lr_src ..), the
// flat run-major scratch written by prb_v_scatter_flat) into its column of the bundle, two
// entries per row word, padded with the bundle's pad entry (the "x_j = 0" slot of its tile).
constexpr int EB_WARPS = 8;

// header word of a lane: vell.cu keeps it in the side array bhdr; the streamed layout of
// vell2.cu (hdr_rows = 1) stores it as the row in front of the bundle's data rows, bit 31 of
// lane l carrying bit l of the number of data rows L.  `data` = the lane's word of data row 0.
__device__ __forceinline__ void ell_put_header(uint32_t hdr, int L, int lane, int hdr_rows,
                                               uint32_t *data, uint32_t *side)
{
    if (hdr_rows)
        *(data - 32) = (hdr & 0x7FFFFFFFu) | ((((uint32_t)L >> lane) & 1u) << 31);
    else
        *side = hdr;
}

__global__ void __launch_bounds__(EB_WARPS * 32)
k_ell_build(const uint16_t *__restrict__ tmp, const int64_t *__restrict__ lr_src,
            const int32_t *__restrict__ lr_len, const uint32_t *__restrict__ lr_hdr,
            const int32_t *__restrict__ brow0, const uint16_t *__restrict__ bundle_pad,
            int64_t n_bundles, uint32_t *__restrict__ ell, uint32_t *__restrict__ bhdr, int hdr_rows)
{
    const int lane = threadIdx.x & 31;
    const int64_t wstride = (int64_t)gridDim.x * EB_WARPS;
    for (int64_t b = blockIdx.x * (int64_t)EB_WARPS + (threadIdx.x >> 5); b < n_bundles; b += wstride) {
        const int64_t src = lr_src[b * 32 + lane];
        const int n = lr_len[b * 32 + lane];
        const int64_t row0 = brow0[b] + hdr_rows;
        const int L = (int)(brow0[b + 1] - row0);
        const uint32_t pad = bundle_pad[b];
        uint32_t *out = ell + row0 * 32 + lane;
        ell_put_header(lr_hdr[b * 32 + lane], L, lane, hdr_rows, out, bhdr + b * 32 + lane);
        const uint16_t *in = tmp + src;
        for (int r = 0; r < L; ++r) {
            const uint32_t e0 = 2 * r < n ? in[2 * r] : pad;
            const uint32_t e1 = 2 * r + 1 < n ? in[2 * r + 1] : pad;
            out[(int64_t)r * 32] = e0 | (e1 << 16);
        }
    }
}

// Bank-aware variant.  The streaming kernel's cost is the shared-memory gather of the shift
// codes: the 32 lanes of a row probe 32 unrelated cells, and cells whose 4-byte words fall into
// the same bank are served one after the other (3.5 wavefronts per gather for random cells).
// The order of the entries inside a lane-run is free, so this builder deals them out such that
// lane l probes bank (l + s) mod 32 at step s whenever it still has an entry there (next
// non-empty bank otherwise): each chunk of EB_CH entries of a lane-run is counting-sorted by
// bank in shared memory and emitted round-robin over the 32 bank queues, the starting bank
// rotated by the lane and the step.  (Simulation: 1.7 wavefronts per gather at 512 entries per
// chunk; measured with ncu, see DESIGN.md.)
constexpr int EB_CH = 512;                 // entries per chunk
constexpr int EB_STRIDE = EB_CH + 2;       // u16 per lane in shared memory (bank skew)
constexpr int EBB_WARPS = 2;

__global__ void __launch_bounds__(EBB_WARPS * 32)
k_ell_build_banked(const uint16_t *__restrict__ tmp, const int64_t *__restrict__ lr_src,
                   const int32_t *__restrict__ lr_len, const uint32_t *__restrict__ lr_hdr,
                   const int32_t *__restrict__ brow0, const uint16_t *__restrict__ bundle_pad,
                   int64_t n_bundles, uint32_t *__restrict__ ell, uint32_t *__restrict__ bhdr,
                   int hdr_rows)
{
    extern __shared__ __align__(16) unsigned char eb_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // per warp: ent[32][EB_STRIDE] (bank-grouped entries of the chunk), ptr[32][33], end[32][33]
    uint16_t *ent = (uint16_t *)eb_raw + (size_t)warp * (32 * EB_STRIDE + 2 * 32 * 33) + lane * EB_STRIDE;
    uint16_t *ptr = (uint16_t *)eb_raw + (size_t)warp * (32 * EB_STRIDE + 2 * 32 * 33) + 32 * EB_STRIDE + lane * 33;
    uint16_t *end = ptr + 32 * 33;
    const int64_t wstride = (int64_t)gridDim.x * EBB_WARPS;
    for (int64_t b = blockIdx.x * (int64_t)EBB_WARPS + warp; b < n_bundles; b += wstride) {
        const int64_t src = lr_src[b * 32 + lane];
        const int n = lr_len[b * 32 + lane];
        const int64_t row0 = brow0[b] + hdr_rows;
        const int L = (int)(brow0[b + 1] - row0);
        const uint32_t pad = bundle_pad[b];
        ell_put_header(lr_hdr[b * 32 + lane], L, lane, hdr_rows, ell + row0 * 32 + lane,
                       bhdr + b * 32 + lane);
        const uint16_t *in = tmp + src;
        for (int c0 = 0; c0 < 2 * L; c0 += EB_CH) {
            const int m = max(0, min(n - c0, EB_CH));  // this lane's entries in the chunk
            // counting sort by bank (every lane works on its own rows of ptr / end / ent)
            for (int q = 0; q < 32; ++q) end[q] = 0;
            for (int k = 0; k < m; ++k) end[(in[c0 + k] >> 2) & 31] += 1;
            uint32_t mask = 0;
            int run = 0;
            for (int q = 0; q < 32; ++q) {
                const int c = end[q];
                if (c) mask |= 1u << q;
                ptr[q] = (uint16_t)run;
                run += c;
                end[q] = (uint16_t)run;
            }
            for (int k = 0; k < m; ++k) {
                const uint16_t e = in[c0 + k];
                const int q = (e >> 2) & 31;
                const int at = ptr[q];
                ent[at] = e;
                ptr[q] = (uint16_t)(at + 1);
            }
            // rewind the queue heads: head of bank q = end of bank q-1
            for (int q = 31; q > 0; --q) ptr[q] = end[q - 1];
            ptr[0] = 0;
            // deal the entries out: step s wants bank (lane + c0 + s) mod 32
            const int rows = min(L - c0 / 2, EB_CH / 2);
            uint32_t *out = ell + (row0 + c0 / 2) * 32 + lane;
            for (int r = 0; r < rows; ++r) {
                uint32_t w = 0;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int s = 2 * r + half;
                    uint32_t e = pad;
                    if (s < m) {
                        const int q0 = (lane + c0 + s) & 31;
                        const int q = (q0 + __ffs(__funnelshift_r(mask, mask, q0)) - 1) & 31;
                        const int at = ptr[q];
                        e = ent[at];
                        ptr[q] = (uint16_t)(at + 1);
                        if (at + 1 == end[q]) mask &= ~(1u << q);
                    }
                    w |= e << (16 * half);
                }
                out[(int64_t)r * 32] = w;
            }
        }
    }
}

// dynamic shared memory base of a kernel without static shared memory (v_probe's assumption)
__global__ void k_smem_base(uint32_t *out)
{
    extern __shared__ __align__(128) unsigned char probe_raw[];
    *out = v_smem_u32(probe_raw);
}

}  // namespace prb

using namespace prb;

using ECfg0 = ECfg<1 << 16, 20, 2, 32>;  //  64 KB of shift codes + 32 warps x 2 x 2.5 KB
using ECfg1 = ECfg<1 << 16, 10, 2, 32>;  //  64 KB + 32 x 2 x 1.25 KB
using ECfg2 = ECfg<1 << 16, 13, 3, 32>;  //  64 KB + 32 x 3 x 1.625 KB

static int e_cfg()
{
    static const int cfg = [] {
        const char *e = getenv("PRB_ELL_CFG");
        const int v = e ? atoi(e) : 0;
        return v >= 0 && v <= 2 ? v : 0;
    }();
    return cfg;
}

// The kernels address the shared-memory tile by the bare cell index + V_SMEM_BASE (v_probe):
// verify once per process that dynamic shared memory really starts there (0 = ok).
extern "C" int prb_v_check_smem_base(void)
{
    static int state = -1;  // -1 unknown, 0 ok, 1 mismatch
    if (state < 0) {
        uint32_t *d = nullptr, h = 0;
        PRB_CUDA(cudaMalloc(&d, sizeof(uint32_t)));
        k_smem_base<<<1, 1, 1024>>>(d);
        cudaError_t e = cudaMemcpy(&h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost);
        cudaFree(d);
        if (e != cudaSuccess) return fail("prb_v_check_smem_base", cudaGetErrorString(e));
        state = h == (uint32_t)V_SMEM_BASE ? 0 : 1;
    }
    PRB_REQUIRE(state == 0,
                "dynamic shared memory does not start at the window offset the V-layout kernels "
                "assume (V_SMEM_BASE); rebuild with the value this driver / toolkit uses");
    return 0;
}

extern "C" int prb_ell_build(const uint16_t *tmp, const int64_t *lr_src, const int32_t *lr_len,
                             const uint32_t *lr_hdr, const int32_t *brow0,
                             const uint16_t *bundle_pad, int64_t n_bundles, uint32_t *ell,
                             uint32_t *bhdr, int hdr_in_stream, prb_stream_t stream)
{
    if (n_bundles == 0) return 0;
    PRB_REQUIRE(tmp && lr_src && lr_len && lr_hdr && brow0 && bundle_pad && ell &&
                    (bhdr || hdr_in_stream),
                "null pointer");
    const int hdr_rows = hdr_in_stream ? 1 : 0;
    static const int banked = [] {
        const char *e = getenv("PRB_ELL_BANKED");  // 0: plain transposition (tuning / A-B runs)
        return e ? atoi(e) : 1;
    }();
    if (banked) {
        const int smem = EBB_WARPS * (32 * EB_STRIDE + 2 * 32 * 33) * 2;
        PRB_CUDA(cudaFuncSetAttribute(k_ell_build_banked,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int grid = (int)hmin(ceil_div(n_bundles, EBB_WARPS), (int64_t)devinfo().sm_count * 3);
        k_ell_build_banked<<<grid, EBB_WARPS * 32, smem, S(stream)>>>(
            tmp, lr_src, lr_len, lr_hdr, brow0, bundle_pad, n_bundles, ell, bhdr, hdr_rows);
        PRB_LAUNCH_CHECK("k_ell_build_banked");
        return 0;
    }
    const int grid = (int)hmin(ceil_div(n_bundles, EB_WARPS), (int64_t)devinfo().sm_count * 16);
    k_ell_build<<<grid, EB_WARPS * 32, 0, S(stream)>>>(tmp, lr_src, lr_len, lr_hdr, brow0,
                                                       bundle_pad, n_bundles, ell, bhdr, hdr_rows);
    PRB_LAUNCH_CHECK("k_ell_build");
    return 0;
}

// warps of one launch (to size the work items)
extern "C" int prb_ell_warps(void) { return ECfg0::warps * devinfo().sm_count; }

extern "C" int prb_stream_ell(const uint32_t *ell, const uint32_t *bhdr, const int32_t *brow0,
                              int64_t n_bundles, int K1,
                              const int32_t *st_seg0, const int32_t *seg_cell0,
                              const int32_t *seg_class, int n_st, int64_t max_tile_cells,
                              const uint8_t *sx, int C, int32_t *hits, int32_t *counters,
                              const int32_t *item_off, const int32_t *item_b0, int64_t n_items,
                              int counters_clean, int64_t PV, int32_t *dbg, prb_stream_t stream)
{
    if (n_bundles == 0 || n_items == 0) return 0;
    PRB_REQUIRE(K1 >= 1 && K1 <= 20, "the lane-run path needs K <= 21");
    PRB_REQUIRE(n_st >= 1 && C >= 1 && sx && ell && bhdr && brow0 && hits && counters,
                "bad arguments");
    PRB_REQUIRE(item_off && item_b0, "bad item plan");
    PRB_REQUIRE(max_tile_cells <= ECfg0::span - 32, "super-tile spans too many cells");
    if (!counters_clean)  // (prb_pick_scatter zeroes them in the greedy step)
        PRB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * n_st, S(stream)));
    static const int help_forced = [] {
        const char *e = getenv("PRB_V_HELP_MIN");
        return e ? atoi(e) : 0;
    }();
    auto launch = [&](auto kern, int smem, int warps) -> int {
        const int grid = (int)hmax(1, hmin(devinfo().sm_count, ceil_div(n_items, warps)));
        PRB_REQUIRE(smem + 512 <= devinfo().smem_optin, "not enough shared memory per block");
        PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, warps * 32, smem, S(stream)>>>(ell, bhdr, brow0, K1, st_seg0, seg_cell0, seg_class,
                                                    n_st, sx, C, hits, counters, item_off, item_b0,
                                                    help_forced > 0 ? help_forced : 8 * warps,
                                                    (uint32_t)PV, dbg);
        return 0;
    };
    const int nreg = (K1 + 3) / 4;
#define PRB_EL(CFG)                                                                      \
    switch (nreg) {                                                                      \
        case 1: if (launch(k_stream_ell<CFG, 1>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        case 2: if (launch(k_stream_ell<CFG, 2>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        case 3: if (launch(k_stream_ell<CFG, 3>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        case 4: if (launch(k_stream_ell<CFG, 4>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        default: if (launch(k_stream_ell<CFG, 5>, CFG::smem_bytes, CFG::warps)) return 1;       \
    }
    switch (e_cfg()) {
        case 1: PRB_EL(ECfg1) break;
        case 2: PRB_EL(ECfg2) break;
        default: PRB_EL(ECfg0)
    }
#undef PRB_EL
    PRB_LAUNCH_CHECK("k_stream_ell");
    return 0;
}
