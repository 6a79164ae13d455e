// vell2.cu — the hot loop on the STREAMED lane-run layout: same lane-runs and bundles as vell.cu,
// but everything a warp needs arrives through its ring of TMA bulk copies, and the work of a
// launch is laid out per CTA beforehand.
//
// Same job as vell.cu / vstream.cu — the per-gene two-pointer merge of _sparse_entropy_wrt
// (picturedrocks/markers/mutualinformation/infoset.py:365-391) for the selectors' calls
// entropy_wrt([j]) / entropy_wrt([-1, j]) (iterative.py:53, 55).  What changed against vell.cu,
// each item taken from its ncu profile (profiles/r02a_k_stream_ell_*_ncu_summary.txt):
//   * 7.6 % of the warp time was the global atomic that claims a work item, 8.2 % the CTA
//     barrier at a change of super-tile, up to 36 % (1/8 gene shard) idle SMs at the end of the
//     launch.  Here the stream is cut into one CHUNK of equal cost per CTA when the layout is
//     built; a chunk is cut into items of decreasing size that the 32 warps of the CTA claim
//     from a SHARED-memory counter, their first rows sitting in shared memory as well.
//   * Super-tiles hold at most 32 752 cells and live at 32 K-aligned positions of the (padded)
//     cell space, so the 16-bit entry is the address inside a 64 KB window of TWO tiles (bit 15 =
//     parity of the tile): a chunk that crosses a tile boundary has both tiles resident and no
//     barrier at all.  (A chunk that spans more tiles runs in phases of two.)
//   * The bundle header (per lane: class << 23 | virtual gene; bit 31 of lane l = bit l of the
//     bundle's number of data rows) is the first row of the bundle IN the stream: no side
//     arrays, no global loads in the loop.
//   * The ring never drains between work items: the issuing side claims the next item as soon
//     as the current one is fully in flight.
//   * A stage that lies inside one bundle (the common case) is counted by straight-line code.
//   * The flush of a bundle was 38 % of the launch (measured by leaving it out: 0.81 ms
//     instead of 1.30 ms at config 4): 27 M scalar REDs to random sectors of the 107 MB hits
//     table, every one a read-modify-write that usually misses L2 because the stream keeps
//     flushing the cache.  Now hits is stored v-major with one slot per SEGMENT —
//     hits[g][segment][v][x_j - 1], 16-bit counts (a segment has at most 32 752 cells), the
//     values of x_j contiguous — so a lane-run that holds a WHOLE run (nearly all do) writes
//     its counts with ONE plain 8-byte store (the table is all zero between steps and every run
//     has exactly one writer; the finalize kernels add the segments of a class); only the
//     pieces of a run longer than the lane-run cap still use REDs.  The table is half the size
//     (54 MB at config 4), its stores carry an L2 evict-last policy and the stream is loaded
//     evict-first, so the table stays in L2 from its first write to the finalize kernel.
// Counting itself is unchanged: LDS.U8 (shift code of x_j at the entry's cell) -> SHL -> ADD
// into 8-bit fields of one register per four values of x_j, widened to 16-bit fields every
// <= 127 rows.
//
// HBM-bound by design: 2 bytes per stored entry, read exactly once.
#include <stdlib.h>

#include "vcommon.cuh"

namespace prb {

constexpr int E2_HALF = 1 << 15;        // bytes (= cell slots) of one half of the tile window
constexpr int E2_MAX_ITEMS = 624;       // work items of one phase
// header word of a lane: bit 31 = bit `lane` of the bundle's number of data rows, bit 30 = the
// lane-run is a piece of a longer run (add, do not store), bits id_shift..29 = slot of the run
// in a gene's block of hits (its segment, or its class), bits 0..id_shift-1 virtual gene (all
// ones: empty lane)
constexpr uint32_t E2_SPLIT = 0x40000000u;

template <int STAGE_ROWS, int STAGES, int WARPS>
struct E2Cfg {
    static constexpr int stage_rows = STAGE_ROWS, stages = STAGES, warps = WARPS;
    static constexpr int threads = WARPS * 32;
    static constexpr int stage_bytes = STAGE_ROWS * 128;
    static constexpr int ring_bytes = WARPS * STAGES * stage_bytes;
    static constexpr int bar_bytes = WARPS * STAGES * 8;
    static constexpr int ctl_bytes = (E2_MAX_ITEMS + 1 + 4) * 4;  // s_next + first rows of the items
    static constexpr int smem_bytes = 2 * E2_HALF + ring_bytes + bar_bytes + ctl_bytes;
};

// ell        uint32[32 * rows]: the stream — per bundle one header row, then L data rows (word of
//            lane l = entries 2r, 2r+1 of its lane-run: addresses inside the tile window)
// item_row0  int32[n_items + 1]: first row of every work item (items are consecutive)
// phase      int4[n_phases]: x = first super-tile, y = tiles resident (1 or 2) | 256 when the
//            items are to be claimed last to first, z = first item, w = number of items
//            (<= E2_MAX_ITEMS)
// cta_ph0    int32[gridDim.x + 1]: phases of every CTA's own chunk; the phases [pool0, pool1) are
//            a pool the CTAs draw from afterwards through *pool_ctr (zero on entry), largest
//            first: equal cost is not equal time
// sx         uint8[n_st * 32768 (+ slack)]: shift codes in the padded cell space; slot 32767 of
//            every tile stays PRB_SX_NONE (what padding entries point at)
// hits       VM: uint16[P][n_slot][K1][4 * NREG] (v-major, see above; a slot = a segment, so a
//            count is at most 32 752), all zero on entry; the four counts of register q of a
//            run sit in the order x_j - 1 = 4q, 4q+2, 4q+1, 4q+3 (two 32-bit words as the lane
//            holds them);
//            !VM: int32[P][n_slot][K1][K1] (x_j-major, the layout of vell.cu), added to.
//            The slot of a run is the id field of its header when n_slot > 1, else 0.
//            PV = P * K1.
template <typename CFG, int NREG, bool VM, bool TIMING>
__global__ void __launch_bounds__(CFG::threads, 1)
k_stream_ell2(const uint32_t *__restrict__ ell, const int32_t *__restrict__ item_row0,
              const int4 *__restrict__ phase, const int32_t *__restrict__ cta_ph0, int pool0,
              int pool1, int *pool_ctr, const uint8_t *__restrict__ sx, int K1, int n_slot, int id_shift,
              int32_t *__restrict__ hits, uint32_t PV, long long *timing)
{
    constexpr int SR = CFG::stage_rows, STAGES = CFG::stages, STAGE_BYTES = CFG::stage_bytes;
    constexpr int THREADS = CFG::threads;
    constexpr int FB = 127;  // rows (2 entries per lane) after which an 8-bit field could overflow
    static_assert(SR <= 40, "a stage must fit the 8-bit counter fields several times");
    // NO static shared memory (see v_probe): the tile window is the first 64 KB of the window
    extern __shared__ __align__(128) unsigned char smem_raw[];
    int *s_ctl = (int *)(smem_raw + 2 * E2_HALF + CFG::ring_bytes + CFG::bar_bytes);
    int &s_next = s_ctl[0];
    int *s_row0 = s_ctl + 4;  // [E2_MAX_ITEMS + 1]
    if (v_smem_u32(smem_raw) != V_SMEM_BASE) __trap();  // checked by the host wrapper beforehand
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t *ring =
        (const uint32_t *)(smem_raw + 2 * E2_HALF + warp * (STAGES * STAGE_BYTES)) + lane;
    const uint32_t ring_s = v_smem_u32(smem_raw + 2 * E2_HALF + warp * (STAGES * STAGE_BYTES));
    const uint32_t bar_s = v_smem_u32(smem_raw + 2 * E2_HALF + CFG::ring_bytes) + warp * (STAGES * 8);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < STAGES; ++i) v_mbar_init(bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t n_issued = 0, n_consumed = 0;  // stages of this warp, over the whole launch
    const uint64_t policy = v_policy_evict_first();
    const uint64_t policy_last = v_policy_evict_last();
    const int KK = K1 * K1;
    const uint32_t vg_mask = (1u << id_shift) - 1u;
    const int ph0 = cta_ph0[blockIdx.x], ph1 = cta_ph0[blockIdx.x + 1];
    // the claim of a pool phase is a global atomic: thread 0 issues it one phase ahead
    int pool_next = 0;
    if (tid == 0 && pool1 > pool0) pool_next = pool0 + atomicAdd(pool_ctr, 1);
    bool first_phase = true;
    for (int ph = ph0;; ++ph) {
        if (ph >= ph1 && ph < pool0) ph = pool0;  // own chunk done
        if (ph >= pool0) {
            if (pool1 <= pool0) break;
            __syncthreads();  // (also: every warp has left the previous phase)
            if (tid == 0) {
                s_ctl[1] = pool_next;
                if (pool_next < pool1) pool_next = pool0 + atomicAdd(pool_ctr, 1);
            }
            __syncthreads();
            ph = s_ctl[1];
            if (ph >= pool1) break;
        }
        const int4 pd = __ldg(phase + ph);
        if (!first_phase) __syncthreads();  // every warp has left the previous phase
        first_phase = false;
        for (int i = tid; i <= pd.w; i += THREADS) s_row0[i] = __ldg(item_row0 + pd.z + i);
        if (tid == 0) s_next = 0;
        __syncthreads();
        const int n_tiles = pd.y & 0xFF;
        const bool rev = (pd.y >> 8) != 0;  // items are claimed from the far end of the phase
        const int n_items = pd.w;
        // issuing side: rows [.., is_left) of the item it works on are not in flight yet; an item
        // it has claimed ahead of the counting side waits in (b_row0, b_n)
        int b_n = 0, is_left = 0;
        int b_k = 0;  // (TIMING: index of the item claimed ahead)
        bool exhausted = false;
        const uint32_t *is_ptr = ell;
        auto try_issue = [&]() {
            while (n_issued - n_consumed < (uint32_t)STAGES) {
                if (is_left == 0) {
                    if (b_n != 0 || exhausted) break;
                    int k = 0;
                    if (lane == 0) k = atomicAdd(&s_next, 1);
                    k = __shfl_sync(0xffffffffu, k, 0);
                    if (k >= n_items) {
                        exhausted = true;
                        break;
                    }
                    if (rev) k = n_items - 1 - k;
                    const int r0 = s_row0[k];
                    if (TIMING) b_k = pd.z + k;
                    b_n = s_row0[k + 1] - r0;
                    is_ptr = ell + (int64_t)r0 * 32;
                    is_left = b_n;
                    if (b_n == 0) continue;  // (never produced by the planner)
                }
                const int rows = min(SR, is_left);
                if (lane == 0) {
                    const uint32_t slot = n_issued % STAGES;
                    v_bulk_load_hint(ring_s + slot * STAGE_BYTES, is_ptr, (uint32_t)rows * 128u,
                                     bar_s + 8 * slot, policy);
                }
                is_ptr += rows * 32;
                is_left -= rows;
                ++n_issued;
            }
        };
        try_issue();  // the stream starts to flow while the tiles are staged
        for (int i = tid; i < n_tiles * (E2_HALF / 16); i += THREADS) {
            const int t = pd.x + i / (E2_HALF / 16), o = i % (E2_HALF / 16);
            ((uint4 *)smem_raw)[(t & 1) * (E2_HALF / 16) + o] =
                __ldg((const uint4 *)(sx + (int64_t)t * E2_HALF) + o);
        }
        __syncthreads();
        // counting side
        uint32_t hdr = 0xFFFFFFFFu;
        uint32_t acc[NREG], wide[2 * NREG];
#pragma unroll
        for (int q = 0; q < NREG; ++q) acc[q] = wide[2 * q] = wide[2 * q + 1] = 0u;
        auto count = [&](uint32_t w) {
            const uint32_t x0 = v_probe(w & 0xFFFFu), x1 = v_probe(w >> 16);
#pragma unroll
            for (int q = 0; q < NREG; ++q) acc[q] += v_shl1(x0 - 32u * q) + v_shl1(x1 - 32u * q);
        };
        auto widen = [&]() {
#pragma unroll
            for (int q = 0; q < NREG; ++q) {
                wide[2 * q] += acc[q] & 0x00FF00FFu;
                wide[2 * q + 1] += (acc[q] >> 8) & 0x00FF00FFu;
                acc[q] = 0u;
            }
        };
        while (b_n != 0) {
            int a_left = b_n;  // rows of the current item not counted yet
            b_n = 0;
            const int a_k = b_k;
            if (TIMING && lane == 0) {  // tuning runs: start / end of every item (globaltimer, ns)
                long long t0;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
                timing[2 * a_k] = t0;
            }
            int bl = 0;   // data rows left in the current bundle (items start at a header row)
            int fbc = 0;  // rows counted into the 8-bit fields since they were last emptied
            while (a_left > 0) {
                const int n = min(SR, a_left);
                const uint32_t slot = n_consumed % STAGES;
                v_mbar_wait(bar_s + 8 * slot, (n_consumed / STAGES) & 1u);
                const uint32_t *rp = ring + slot * (SR * 32);
                int n_left = n;
                while (n_left > 0) {
                    if (bl == 0) {  // header row
                        hdr = rp[0];
                        rp += 32;
                        --n_left;
                        bl = (int)__ballot_sync(0xffffffffu, hdr >> 31);
                        continue;
                    }
                    int m;
                    if (n_left == SR && bl >= SR && fbc <= FB - SR) {
                        m = SR;  // the whole stage lies inside the bundle: straight-line code
#pragma unroll
                        for (int i = 0; i < SR; ++i) count(rp[i * 32]);
                    } else {
                        m = min(min(n_left, bl), FB - fbc);
                        const uint32_t *rp_end = rp + m * 32;
                        const uint32_t *p = rp;
#pragma unroll 1
                        for (; p + 96 < rp_end; p += 128) {
                            const uint32_t w0 = p[0], w1 = p[32], w2 = p[64], w3 = p[96];
                            count(w0);
                            count(w1);
                            count(w2);
                            count(w3);
                        }
#pragma unroll 1
                        for (; p < rp_end; p += 32) count(p[0]);
                    }
                    rp += m * 32;
                    n_left -= m;
                    bl -= m;
                    fbc += m;
                    if (bl == 0) {  // end of the bundle: every lane adds its counts to hits
                        const uint32_t vg = hdr & vg_mask;
                        if (vg < PV) {
                            const uint32_t g = vg / (uint32_t)K1, v = vg - g * (uint32_t)K1;
                            const uint32_t slot = n_slot > 1 ? (hdr & ~(E2_SPLIT | 0x80000000u)) >> id_shift : 0u;
                            const int64_t row = ((int64_t)g * n_slot + slot) * K1 + v;
#pragma unroll
                            for (int q = 0; q < NREG; ++q) {
                                // 16-bit counts of x_j - 1 = 4q (low) | 4q+2 (high) and 4q+1 | 4q+3
                                const uint32_t w02 = wide[2 * q] + (acc[q] & 0x00FF00FFu);
                                const uint32_t w13 = wide[2 * q + 1] + ((acc[q] >> 8) & 0x00FF00FFu);
                                if (VM) {
                                    uint32_t *out = reinterpret_cast<uint32_t *>(hits) + row * (2 * NREG) + 2 * q;
                                    if (!(hdr & E2_SPLIT)) {
                                        v_store2_keep(out, w02, w13, policy_last);
                                    } else {  // a piece of a run: the fields cannot carry into each other
                                        if (w02) atomicAdd(out, w02);
                                        if (w13) atomicAdd(out + 1, w13);
                                    }
                                } else {
                                    int32_t *out = hits + ((int64_t)g * n_slot + slot) * KK + v;
                                    const uint32_t c[4] = {w02 & 0xFFFFu, w13 & 0xFFFFu, w02 >> 16, w13 >> 16};
#pragma unroll
                                    for (int f = 0; f < 4; ++f)
                                        if (4 * q + f < K1 && c[f]) atomicAdd(out + (4 * q + f) * K1, (int)c[f]);
                                }
                            }
                        }
#pragma unroll
                        for (int q = 0; q < NREG; ++q) acc[q] = wide[2 * q] = wide[2 * q + 1] = 0u;
                        fbc = 0;
                    } else if (fbc > FB - SR) {
                        widen();
                        fbc = 0;
                    }
                }
                a_left -= n;
                ++n_consumed;
                __syncwarp();  // every lane has read the stage: its slot goes back to the copy engine
                try_issue();
            }
            if (TIMING && lane == 0) {
                long long t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                timing[2 * a_k + 1] = t1;
            }
        }
    }
}

}  // namespace prb

using namespace prb;

using E2Cfg0 = E2Cfg<20, 2, 32>;  // 64 KB tile window + 32 warps x 2 x 2.5 KB
using E2Cfg1 = E2Cfg<13, 3, 32>;  // 64 KB + 32 x 3 x 1.625 KB
using E2Cfg2 = E2Cfg<10, 4, 32>;  // 64 KB + 32 x 4 x 1.25 KB

static int e2_cfg()
{
    static const int cfg = [] {
        const char *e = getenv("PRB_ELL_CFG");
        const int v = e ? atoi(e) : 0;
        return v >= 0 && v <= 2 ? v : 0;
    }();
    return cfg;
}

// cells a super-tile of the streamed layout may hold, CTAs of a launch, warps per CTA, items per
// phase: what the host-side planner (vplan.ell2_chunks) needs
extern "C" int prb_ell2_geometry(int *tile_cells, int *ctas, int *warps, int *max_items)
{
    if (tile_cells) *tile_cells = E2_HALF - 16;
    if (ctas) *ctas = devinfo().sm_count;
    if (warps) *warps = E2Cfg0::warps;
    if (max_items) *max_items = E2_MAX_ITEMS;
    return 0;
}

extern "C" int prb_stream_ell2(const uint32_t *ell, const int32_t *item_row0, const int32_t *phase,
                               const int32_t *cta_ph0, int n_ctas, int pool0, int pool1,
                               int32_t *pool_ctr, const uint8_t *sx, int K1,
                               int n_slot, int id_shift, int32_t *hits, int vmajor,
                               int64_t PV, int64_t *timing, prb_stream_t stream)
{
    if (n_ctas == 0) return 0;
    PRB_REQUIRE(K1 >= 1 && K1 <= 20, "the lane-run path needs K <= 21");
    PRB_REQUIRE(ell && item_row0 && phase && cta_ph0 && sx && hits, "null pointer");
    PRB_REQUIRE(id_shift >= 8 && id_shift <= 29 && n_slot >= 1 && n_slot <= (1 << (30 - id_shift)),
                "bad header format");
    PRB_REQUIRE(PV >= 1 && PV < (int64_t)(1 << id_shift) - 1,
                "too many virtual genes for the header's gene field");
    PRB_REQUIRE(n_ctas >= 1 && n_ctas <= 65535, "bad grid");
    PRB_REQUIRE(pool1 <= pool0 || pool_ctr != nullptr, "the pool needs its counter");
    auto launch = [&](auto kern, int smem, int warps) -> int {
        PRB_REQUIRE(smem <= devinfo().smem_optin, "not enough shared memory per block");
        PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<n_ctas, warps * 32, smem, S(stream)>>>(ell, item_row0, (const int4 *)phase, cta_ph0,
                                                      pool0, pool1, pool_ctr, sx,
                                                      K1, n_slot, id_shift, hits, (uint32_t)PV,
                                                      (long long *)timing);
        return 0;
    };
    const int nreg = (K1 + 3) / 4;
    if (timing) {  // tuning runs (profiles/ell2_timing.py)
        PRB_REQUIRE(nreg == 1 && vmajor, "item timing is built for K <= 5, v-major hits only");
        if (launch(k_stream_ell2<E2Cfg0, 1, true, true>, E2Cfg0::smem_bytes, E2Cfg0::warps)) return 1;
        PRB_LAUNCH_CHECK("k_stream_ell2");
        return 0;
    }
#define PRB_E2V(CFG, VM)                                                                            \
    switch (nreg) {                                                                                 \
        case 1: if (launch(k_stream_ell2<CFG, 1, VM, false>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        case 2: if (launch(k_stream_ell2<CFG, 2, VM, false>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        case 3: if (launch(k_stream_ell2<CFG, 3, VM, false>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        case 4: if (launch(k_stream_ell2<CFG, 4, VM, false>, CFG::smem_bytes, CFG::warps)) return 1; break; \
        default: if (launch(k_stream_ell2<CFG, 5, VM, false>, CFG::smem_bytes, CFG::warps)) return 1;       \
    }
#define PRB_E2(CFG)          \
    if (vmajor) {            \
        PRB_E2V(CFG, true)   \
    } else {                 \
        PRB_E2V(CFG, false)  \
    }
    switch (e2_cfg()) {
        case 1: PRB_E2(E2Cfg1) break;
        case 2: PRB_E2(E2Cfg2) break;
        default: PRB_E2(E2Cfg0)
    }
#undef PRB_E2V
#undef PRB_E2
    PRB_LAUNCH_CHECK("k_stream_ell2");
    return 0;
}
