// swar.cu — the hot loop for K <= 5 bins: class-sorted cells + register counters.
//
// Replaces the per-gene two-pointer merge of _sparse_entropy_wrt
// (picturedrocks/markers/mutualinformation/infoset.py:365-391) for the selectors' calls
// entropy_wrt([j]) and entropy_wrt([-1, j]) (iterative.py:53, 55) and the class-conditional
// count tables behind entropy_wrt([]) / entropy_wrt([-1]) (iterative.py:30, 32).
//
// Cells are kept SORTED BY CLASS (prb_class_sort; a relabelling of cells changes no
// entropy).  Inside one gene the stored entries of a class are then one contiguous run,
// the joint histogram of (y, x_j, x_i) restricted to that run has only (K-1)^2 <= 16
// bins, and those fit in REGISTERS: every thread counts its entries into sixteen 4-bit
// fields (two 32-bit words, one shift + add per entry), which are widened to 8-bit
// fields before they can overflow, summed over the warp with a 9-shuffle transpose-reduce
// at the end of the run (or every 31 blocks) and added to hits[g][class][x_j-1][x_i-1].
// No shared-memory atomics, no per-team histograms; shared memory holds only the
// per-cell x_j shift code (int8) of one "super-tile" of cells (up to ~220k cells, whole
// classes where possible), so a probe is ONE shared byte load.
//
// HBM-bound by design: 4 bytes per stored entry, read exactly once through a per-warp ring
// of 1 KB TMA bulk copies (cp.async.bulk + mbarrier, 3 blocks in flight per warp, 96 KB
// per SM); one persistent CTA per SM; (super-tile, gene-chunk) work items claimed
// dynamically per warp.
#include <stdlib.h>

#include "common.cuh"

namespace prb {

constexpr int SW_THREADS = 1024;
constexpr int SW_U = 8;                 // words per lane and block
constexpr int SW_BLOCK = 32 * SW_U;     // entries per warp and block
constexpr int SW_MAX_SEGS = 31;         // class segments per super-tile (32 boundaries = 1 warp load)

// PTX shl clamps: a shift amount >= 32 (including "negative" ones seen as unsigned) gives 0.
__device__ __forceinline__ uint32_t shl1(uint32_t s)
{
    uint32_t r;
    asm("shl.b32 %0, 1, %1;" : "=r"(r) : "r"(s));
    return r;
}

struct Counters {
    uint32_t A, B;      // 16 x 4-bit fields: bin b = (x_j-1)*4 + (x_i-1); A = bins 0..7
    uint32_t a8[4];     // 16 x 8-bit fields
    int n8;             // spills into a8 since the last flush (warp-uniform)

    __device__ __forceinline__ void clear()
    {
        A = B = 0;
        n8 = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) a8[i] = 0;
    }
    // after at most 8 increments per field
    __device__ __forceinline__ void spill4()
    {
        a8[0] += A & 0x0F0F0F0Fu;
        a8[1] += (A >> 4) & 0x0F0F0F0Fu;
        a8[2] += B & 0x0F0F0F0Fu;
        a8[3] += (B >> 4) & 0x0F0F0F0Fu;
        A = B = 0;
    }
};

// Widen the sixteen 8-bit fields to 16 bits, sum them over the warp (transpose-reduce:
// 4 + 2 + 1 + 1 + 1 shuffles) and add them to out[(x_j-1)*K1 + x_i-1].  Per-lane fields
// are <= 255, so 32 lanes cannot carry into the neighbouring 16-bit field.
__device__ __forceinline__ void flush(Counters &c, int32_t *out, int K1, int lane)
{
    uint32_t a16[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        a16[2 * r] = c.a8[r] & 0x00FF00FFu;
        a16[2 * r + 1] = (c.a8[r] >> 8) & 0x00FF00FFu;
        c.a8[r] = 0;
    }
    c.n8 = 0;
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    uint32_t r4[4], r2[2], r1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t send = b4 ? a16[i] : a16[i + 4];
        const uint32_t keep = b4 ? a16[i + 4] : a16[i];
        r4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint32_t send = b3 ? r4[i] : r4[i + 2];
        const uint32_t keep = b3 ? r4[i + 2] : r4[i];
        r2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    {
        const uint32_t send = b2 ? r2[0] : r2[1];
        const uint32_t keep = b2 ? r2[1] : r2[0];
        r1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
    r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
    // lane holds register q = lane >> 2 (q = 4*b4 + 2*b3 + b2): low half = bin lo(q), high = lo(q)+4
    const int q = lane >> 2;
    const int lo = ((q & 4) << 1) | ((q & 1) << 1) | ((q >> 1) & 1);  // {0,2,1,3,8,10,9,11}
    const int sub = lane & 3;
    if (sub < 2) {
        const int bin = lo + 4 * sub;
        const uint32_t cnt = sub ? (r1 >> 16) : (r1 & 0xFFFFu);
        const int a = bin >> 2, v = bin & 3;
        if (cnt && a < K1 && v < K1) atomicAdd(out + a * K1 + v, (int)cnt);
    }
}

// The super-tile's shift codes live in shared memory at index (cell id mod SPAN), SPAN a
// power of two: a super-tile spans at most SPAN consecutive cells, so the mapping is
// injective and a probe is one AND plus one LDS from a constant base.
template <bool COUNT, bool MASKED, int SPAN>
__device__ __forceinline__ void acc(uint32_t w, bool ok, const int8_t *__restrict__ st, Counters &c)
{
    // field offset = 4*code - 4 (+ 16*(x_j-1), folded into the shift code); multiply-adds
    // keep this off the (busier) shift/logic pipe
    uint32_t sh = (w >> 24) * 4u + (COUNT ? 0xFFFFFFFCu : (uint32_t)(int)st[w & (SPAN - 1)]);
    if (MASKED) sh = ok ? sh : 127u;  // >= 64: counts nothing
    c.A += shl1(sh);
    if (!COUNT) c.B += shl1(sh - 32u);
}

// ---- per-warp TMA ring: the packed stream goes global -> shared with 1-D bulk copies ----
// (cp.async.bulk, one instruction per stage, completion on an mbarrier), STAGES stages of
// UNITS x 1 KB in flight per warp without holding registers or LSU issue slots.
constexpr int SW_WARPS = SW_THREADS / 32;

template <int SPAN, int UNITS, int STAGES>
struct SwarCfg {
    static constexpr int span = SPAN, units = UNITS, stages = STAGES;
    static constexpr int stage_entries = UNITS * SW_BLOCK;
    static constexpr int stage_bytes = stage_entries * 4;
    static constexpr int ring_bytes = SW_WARPS * STAGES * stage_bytes;
    static constexpr int smem_bytes = SPAN + ring_bytes + SW_WARPS * STAGES * 8;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}

// State of one warp walking one gene through the class segments of a super-tile.
// Positions are entry indices relative to the gene's first entry (indptr[g]).
struct Walk {
    int seg, nsg, seg_end, gend;
    int32_t bl;        // lane l: boundary l of this gene inside the super-tile
    int32_t *og;       // out + g*C*stride
    int32_t *oseg;     // og + class(seg)*stride
    int lo;            // first position not yet counted
    bool dirty;
};

// Count one unit of 256 entries starting at position blk (16-byte aligned in memory, so the
// first unit of a gene may start below the gene's first entry of this super-tile and the
// last one may run past its end).  Lane l holds entries blk + 128*h + 4*l + j in w[4*h + j].
template <bool COUNT, int SPAN>
__device__ __forceinline__ void process_block(const uint32_t (&w)[SW_U], int blk,
                                              const int8_t *__restrict__ st, Counters &c,
                                              Walk &k, const int *s_cls, int out_stride, int K1,
                                              int lane)
{
    const int nxt = blk + SW_BLOCK;
    if (k.lo <= blk && k.seg_end > nxt) {
        // the whole unit lies inside the current class segment
#pragma unroll
        for (int u = 0; u < SW_U; ++u) acc<COUNT, false, SPAN>(w[u], true, st, c);
        k.dirty = true;
    } else {
        const int blk_end = min(nxt, k.gend);
        for (;;) {
            const int hi = min(k.seg_end, blk_end);
            if (hi > k.lo) {
                const int rlo = k.lo - blk - 4 * lane, rhi = hi - blk - 4 * lane;
#pragma unroll
                for (int u = 0; u < SW_U; ++u) {
                    const int off = (u >> 2) * 128 + (u & 3);
                    acc<COUNT, true, SPAN>(w[u], off >= rlo && off < rhi, st, c);
                }
                k.dirty = true;
                k.lo = hi;
            }
            if (k.seg_end > blk_end) break;
            // the segment ends inside (or at the end of) this unit
            if (k.dirty) {
                c.spill4();
                flush(c, k.oseg, K1, lane);
                k.dirty = false;
            }
            if (++k.seg >= k.nsg) break;
            k.seg_end = __shfl_sync(0xffffffffu, k.bl, k.seg + 1);
            k.oseg = k.og + s_cls[k.seg] * out_stride;
        }
    }
    c.spill4();
    if (++c.n8 == 31) flush(c, k.oseg, K1, lane);  // <= 31*8 = 248 per lane and 8-bit field
}

// Segment plan (built on the host by the caller, see engine.py):
//   seg_cell0 int32[n_seg+1] : cell boundaries of the class segments (ascending, last = N)
//   seg_class int32[n_seg]   : class of each segment, or NULL (everything counts as class 0)
//   st_seg0   int32[n_st+1]  : super-tile t = segments [st_seg0[t], st_seg0[t+1]), <= 31 of
//                              them, spanning <= SPAN - 16 cells
//   tpf       int32[P][n_seg+1] : first stored entry of gene g (relative to indptr[g]) whose
//                                 cell id is >= seg_cell0[s]
// packed must be readable up to the next multiple of 16 bytes past its last entry.
// out: hits int32[P][C][K1][K1]  (COUNT: cnt int32[P][C][K1], every cell has x_j = 1)
template <bool COUNT, typename CFG>
__global__ void __launch_bounds__(SW_THREADS, 1)
k_stream_swar(const uint32_t *__restrict__ packed, const int64_t *__restrict__ indptr,
              const int32_t *__restrict__ tpf, int n_seg, int64_t P,
              const int32_t *__restrict__ st_seg0, const int32_t *__restrict__ seg_cell0,
              const int32_t *__restrict__ seg_class, int n_st, const int8_t *__restrict__ sx,
              int K1, int C, int32_t *__restrict__ out, int *counters, int genes_per_item)
{
    constexpr int SPAN = CFG::span, UNITS = CFG::units, STAGES = CFG::stages;
    constexpr int STAGE_ENTRIES = CFG::stage_entries, STAGE_BYTES = CFG::stage_bytes;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ int s_flag;
    __shared__ int s_cls[SW_MAX_SEGS + 1];
    const int8_t *st = (const int8_t *)smem_raw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char *ring = smem_raw + SPAN + warp * (STAGES * STAGE_BYTES);
    const uint32_t ring_s = smem_u32(ring);
    const uint32_t bar_s = smem_u32(smem_raw + SPAN + CFG::ring_bytes) + warp * (STAGES * 8);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) mbar_init(bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    int stage = 0;        // next stage to consume (== next to fill once the ring is primed)
    uint32_t parity = 0;  // bit i: phase parity to wait for on stage i
    const int n_items = (int)((P + genes_per_item - 1) / genes_per_item);
    const int out_stride = COUNT ? K1 : K1 * K1;
    const int nb = n_seg + 1;
    Counters c;
    c.clear();

    for (int tt = 0; tt < n_st; ++tt) {
        const int t = (int)((blockIdx.x + tt) % (unsigned)n_st);
        __syncthreads();  // every warp has left the previous super-tile
        if (tid == 0) s_flag = *(volatile int *)(counters + t) < n_items;
        __syncthreads();
        if (!s_flag) continue;  // already drained by other CTAs
        const int s0 = st_seg0[t], s1 = st_seg0[t + 1];
        const int nsg = s1 - s0;
        if (tid < nsg) s_cls[tid] = seg_class ? seg_class[s0 + tid] : 0;
        if (!COUNT) {
            // per-cell shift codes of the super-tile: global -> shared in 16-byte chunks, from
            // the aligned address below its first cell (the buffer is padded), placed at
            // (cell id mod SPAN)
            const int64_t c0 = seg_cell0[s0], c1 = seg_cell0[s1];
            const int64_t a0 = c0 & ~int64_t(15);
            const int nvec = (int)((c1 - a0 + 15) >> 4);
            const uint4 *src = (const uint4 *)(sx + a0);
            const int v0 = (int)(a0 >> 4);
            uint4 *dst = (uint4 *)smem_raw;
            for (int i = tid; i < nvec; i += SW_THREADS)
                dst[(v0 + i) & (SPAN / 16 - 1)] = __ldg(src + i);
        }
        __syncthreads();
        for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(counters + t, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            if (item >= n_items) break;
            const int64_t g_end = min((int64_t)(item + 1) * genes_per_item, P);
            for (int64_t g = (int64_t)item * genes_per_item; g < g_end; ++g) {
                Walk k;
                // boundaries of this gene inside the super-tile: lane l holds boundary s0 + l
                k.bl = lane <= nsg ? __ldg(tpf + g * nb + s0 + lane) : 0;
                const int gpos = __shfl_sync(0xffffffffu, k.bl, 0);
                k.gend = __shfl_sync(0xffffffffu, k.bl, nsg);
                if (gpos >= k.gend) continue;
                k.nsg = nsg;
                k.seg = 0;
                k.seg_end = __shfl_sync(0xffffffffu, k.bl, 1);
                k.og = out + g * (int64_t)C * out_stride;
                k.oseg = k.og + s_cls[0] * out_stride;
                k.lo = gpos;
                k.dirty = false;
                const int64_t base = __ldg(indptr + g);
                // stages are 16-byte aligned in memory: start at or below the first entry
                const int ea = gpos - (int)((base + gpos) & 3);
                const uint32_t *src = packed + base + ea;
                const int n_ent = k.gend - ea;
                const int nstg = (n_ent + STAGE_ENTRIES - 1) / STAGE_ENTRIES;
                auto issue = [&](int b, int s) {
                    const int rem = n_ent - b * STAGE_ENTRIES;
                    const uint32_t bytes =
                        rem >= STAGE_ENTRIES ? STAGE_BYTES : ((rem * 4 + 15) & ~15);
                    bulk_load(ring_s + s * STAGE_BYTES, src + (int64_t)b * STAGE_ENTRIES, bytes,
                              bar_s + 8 * s);
                };
                // all copies of the previous gene were consumed: stages [stage, stage + STAGES)
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < STAGES; ++i)
                        if (i < nstg) issue(i, (stage + i) % STAGES);
                }
                // one copy of the counting code (runtime stage / unit indices): the body is
                // large and must stay inside the instruction cache
#pragma unroll 1
                for (int b = 0; b < nstg; ++b) {
                    mbar_wait(bar_s + 8 * stage, (parity >> stage) & 1u);
                    parity ^= 1u << stage;
                    const int rem = n_ent - b * STAGE_ENTRIES;  // > 0
                    const int nu = min(UNITS, (rem + SW_BLOCK - 1) / SW_BLOCK);
                    const uint4 *rb = (const uint4 *)(ring + stage * STAGE_BYTES);
#pragma unroll 1
                    for (int h = 0; h < nu; ++h) {
                        const uint4 q0 = rb[h * 64 + lane], q1 = rb[h * 64 + 32 + lane];
                        const uint32_t w[SW_U] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                        if (h == nu - 1) {
                            __syncwarp();  // every lane has read the stage: refill it
                            if (lane == 0 && b + STAGES < nstg) issue(b + STAGES, stage);
                        }
                        process_block<COUNT, SPAN>(w, ea + b * STAGE_ENTRIES + h * SW_BLOCK, st, c,
                                                   k, s_cls, out_stride, K1, lane);
                    }
                    stage = stage + 1 == STAGES ? 0 : stage + 1;
                }
                if (k.dirty) flush(c, k.oseg, K1, lane);
            }
        }
    }
}

// tpf[g][s] = first entry of gene g (relative to indptr[g]) whose cell id is >= seg_cell0[s]
__global__ void k_seg_ptrs(const uint32_t *__restrict__ packed, const int64_t *__restrict__ indptr,
                           int64_t P, const int32_t *__restrict__ seg_cell0, int n_seg,
                           int32_t *__restrict__ tpf)
{
    const int nb = n_seg + 1;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P * nb) return;
    const int64_t g = i / nb;
    const int s = (int)(i - g * nb);
    const int64_t b = indptr[g];
    int64_t lo = b, hi = indptr[g + 1];
    if (s == 0) {
        tpf[i] = 0;
        return;
    }
    if (s == n_seg) {
        tpf[i] = (int32_t)(hi - b);
        return;
    }
    const uint32_t row0 = (uint32_t)seg_cell0[s];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((packed[mid] & ROW_MASK) < row0)
            lo = mid + 1;
        else
            hi = mid;
    }
    tpf[i] = (int32_t)(lo - b);
}

// Stable partition of every gene's entries by the class of their cell, with the cell ids
// relabelled: out word = code | newpos[row].  One warp per gene (dynamic queue), two passes:
// class counts -> exclusive offsets -> ordered scatter (rank inside a 32-entry group by
// __match_any_sync, so the ascending order inside a class is kept).
constexpr int CS_WARPS = 8;

__global__ void __launch_bounds__(CS_WARPS * 32)
k_class_sort(const uint32_t *__restrict__ in, const int64_t *__restrict__ indptr, int64_t P,
             const int32_t *__restrict__ newpos, const int32_t *__restrict__ cls, int C,
             uint32_t *__restrict__ out, int *queue)
{
    extern __shared__ int off_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *off = off_all + warp * C;
    for (;;) {
        int gi = 0;
        if (lane == 0) gi = atomicAdd(queue, 1);
        const int64_t g = __shfl_sync(0xffffffffu, gi, 0);
        if (g >= P) break;
        const int64_t b = indptr[g], e = indptr[g + 1];
        for (int i = lane; i < C; i += 32) off[i] = 0;
        __syncwarp();
        for (int64_t i = b + lane; i < e; i += 32) atomicAdd(off + cls[in[i] & ROW_MASK], 1);
        __syncwarp();
        int carry = 0;
        for (int c0 = 0; c0 < C; c0 += 32) {
            const int i = c0 + lane;
            const int v = i < C ? off[i] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (i < C) off[i] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        __syncwarp();
        for (int64_t i0 = b; i0 < e; i0 += 32) {
            const int64_t i = i0 + lane;
            const bool valid = i < e;
            const unsigned act = __ballot_sync(0xffffffffu, valid);
            if (valid) {
                const uint32_t w = in[i];
                const uint32_t row = w & ROW_MASK;
                const int cc = cls[row];
                const unsigned m = __match_any_sync(act, cc);
                const int rank = __popc(m & ((1u << lane) - 1u));
                const int base = off[cc];
                out[b + base + rank] = (w & ~ROW_MASK) | (uint32_t)newpos[row];
                __syncwarp(act);
                if (rank == 0) off[cc] = base + __popc(m);
            }
            __syncwarp();
        }
    }
}

}  // namespace prb

using namespace prb;

// Kernel configurations (cells per super-tile, 1 KB units per stage, stages per warp); all fit
// the 227 KB of shared memory.  PRB_SWAR_CFG selects one for tuning runs.
using Cfg0 = SwarCfg<1 << 17, 1, 3>;  // 128 KB of shift codes + 32 warps x 3 x 1 KB
using Cfg1 = SwarCfg<1 << 16, 2, 2>;  //  64 KB + 32 x 2 x 2 KB
using Cfg2 = SwarCfg<1 << 16, 1, 5>;  //  64 KB + 32 x 5 x 1 KB
using Cfg3 = SwarCfg<1 << 16, 1, 3>;  //  64 KB + 32 x 3 x 1 KB (tile-size probe)

static int swar_cfg()
{
    static const int cfg = [] {
        const char *e = getenv("PRB_SWAR_CFG");
        const int v = e ? atoi(e) : 1;  // measured best on config 4: 64 KB tile, 2 x 2 KB ring
        return v >= 0 && v <= 3 ? v : 1;
    }();
    return cfg;
}

static int swar_span()
{
    switch (swar_cfg()) {
        case 1: return Cfg1::span;
        case 2: return Cfg2::span;
        case 3: return Cfg3::span;
        default: return Cfg0::span;
    }
}

extern "C" int prb_swar_tile_capacity(int64_t *max_cells, int *max_segs)
{
    // a super-tile's shift codes sit in shared memory at (cell id mod SPAN); 16 cells of
    // slack for the aligned 16-byte copies
    if (max_cells) *max_cells = swar_span() - 16;
    if (max_segs) *max_segs = SW_MAX_SEGS;
    return 0;
}

extern "C" int prb_seg_ptrs(const uint32_t *packed, const int64_t *indptr, int64_t P,
                            const int32_t *seg_cell0, int n_seg, int32_t *tpf, prb_stream_t stream)
{
    PRB_REQUIRE(P > 0 && n_seg >= 1, "bad arguments");
    const int64_t n = P * (int64_t)(n_seg + 1);
    k_seg_ptrs<<<(unsigned)ceil_div(n, 256), 256, 0, S(stream)>>>(packed, indptr, P, seg_cell0,
                                                                 n_seg, tpf);
    PRB_LAUNCH_CHECK("k_seg_ptrs");
    return 0;
}

extern "C" int prb_stream_swar(const uint32_t *packed, const int64_t *indptr, const int32_t *tpf,
                               int n_seg, int64_t P, int64_t nnz, const int32_t *st_seg0,
                               const int32_t *seg_cell0, const int32_t *seg_class, int n_st,
                               int64_t max_tile_cells, const int8_t *sx, int K1, int C,
                               int32_t *out, int32_t *counters, prb_stream_t stream)
{
    if (P == 0 || nnz == 0) return 0;
    PRB_REQUIRE(K1 >= 1 && K1 <= 4, "the register-counter path needs K <= 5");
    PRB_REQUIRE(n_seg >= 1 && n_st >= 1 && C >= 1, "bad segment plan");
    const bool count = sx == nullptr;
    PRB_REQUIRE(max_tile_cells <= swar_span() - 16, "super-tile spans too many cells");
    PRB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * n_st, S(stream)));
    // ~16 work items per warp and super-tile
    const int64_t warps_per_tile = hmax(1, (int64_t)devinfo().sm_count * SW_WARPS / n_st);
    const int gpi = (int)hmin(64, hmax(1, P / (warps_per_tile * 16)));
    const int64_t n_items = ceil_div(P, gpi) * n_st;
    const int grid = (int)hmax(1, hmin(devinfo().sm_count, ceil_div(n_items, SW_WARPS)));
    auto launch = [&](auto kern, int smem) -> int {
        PRB_REQUIRE(smem + 512 <= devinfo().smem_optin, "not enough shared memory per block");
        PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, SW_THREADS, smem, S(stream)>>>(packed, indptr, tpf, n_seg, P, st_seg0,
                                                    seg_cell0, seg_class, n_st, sx, K1, C, out,
                                                    counters, gpi);
        return 0;
    };
#define PRB_SW(CFG)                                                                    \
    if (count ? launch(k_stream_swar<true, CFG>, CFG::smem_bytes)                      \
              : launch(k_stream_swar<false, CFG>, CFG::smem_bytes))                    \
        return 1;                                                                      \
    break
    switch (swar_cfg()) {
        case 1: PRB_SW(Cfg1);
        case 2: PRB_SW(Cfg2);
        case 3: PRB_SW(Cfg3);
        default: PRB_SW(Cfg0);
    }
#undef PRB_SW
    PRB_LAUNCH_CHECK("k_stream_swar");
    return 0;
}

extern "C" int prb_class_sort(const uint32_t *packed_in, const int64_t *indptr, int64_t P,
                              const int32_t *newpos, const int32_t *cls, int C,
                              uint32_t *packed_out, int32_t *queue, prb_stream_t stream)
{
    if (P == 0) return 0;
    PRB_REQUIRE(C >= 1 && C <= 4096, "class count out of range");
    PRB_CUDA(cudaMemsetAsync(queue, 0, sizeof(int32_t), S(stream)));
    const int grid = (int)hmin(ceil_div(P, CS_WARPS), (int64_t)devinfo().sm_count * 8);
    k_class_sort<<<grid, CS_WARPS * 32, (size_t)CS_WARPS * C * 4, S(stream)>>>(
        packed_in, indptr, P, newpos, cls, C, packed_out, queue);
    PRB_LAUNCH_CHECK("k_class_sort");
    return 0;
}
