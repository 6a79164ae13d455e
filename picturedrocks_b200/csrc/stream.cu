// stream.cu — the hot loop: one streaming pass over the packed by-gene CSC that
// builds, for every gene, the joint histogram of (cell state, bin code).
//
// Replaces the per-gene two-pointer merge of _sparse_entropy_wrt
// (picturedrocks/markers/mutualinformation/infoset.py:365-391).  The reference
// visits the union of the gene's rows and the master column's rows; here the
// master column is a per-cell bitmap staged in shared memory, only the
// INTERSECTION (both non-zero) is histogrammed, and every other bin is inferred
// by subtraction from per-gene class-conditional tables in finalize.cu.
//
// HBM-bound by design: 4 bytes per stored entry, read exactly once with
// 128-bit coalesced loads; one persistent CTA per SM; dynamic ranges of the nnz
// stream are claimed per team (1..32 warps) so skewed genes balance.
#include <stdlib.h>

#include "common.cuh"

namespace prb {

thread_local char g_err[512] = "";

const DevInfo &devinfo()
{
    static DevInfo info = [] {
        DevInfo d{148, 227 * 1024};
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) {
            cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
            cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        }
        return d;
    }();
    return info;
}

constexpr int CTA_THREADS = 1024;
constexpr int SLOT_WORDS = 64;  // per-team work slots (+ the tile-skip flag) at the front of smem

// Shared-memory plan of one streaming launch.
//   [slots 256 B][remap u16 x bins][per-cell state of one cell TILE][teams x hist int32]
// The per-cell state (0 = not a master row, else joint-state id + 1; uint8 when the ids
// fit, else uint16) of a tile of cells is staged in shared memory, so the probe of every
// stored entry is ONE shared load and hits need no global gather.  Cells are tiled when N
// does not fit; rows are ascending inside a gene, so a gene's entries of one tile are a
// contiguous sub-range found once by binary search (prb_tile_ptrs).
struct StreamPlan {
    int teams;
    int team_shift;  // log2(threads per team)
    int state_bytes;
    int n_bins;
    int bins_pad;
    int64_t remap_bytes;
    int64_t tile_cells;
    int64_t tile_bytes;
    int n_tiles;
    int64_t smem_bytes;
};

static int plan_stream(int64_t N, int64_t n_states, int K1, StreamPlan *p)
{
    const int64_t smem_max = devinfo().smem_optin;
    const int64_t n_bins = n_states * K1;
    if (n_bins <= 0 || n_bins > 65536) return fail("prb_stream", "bin count outside (0, 65536]");
    p->n_bins = (int)n_bins;
    p->state_bytes = (n_states + 1 <= 256) ? 1 : 2;
    p->bins_pad = (int)((n_bins + 3) & ~int64_t(3));
    p->remap_bytes = ((int64_t)p->bins_pad * 2 + 15) & ~int64_t(15);
    const int64_t hist_bytes = (int64_t)p->bins_pad * 4;
    const int64_t fixed = SLOT_WORDS * 4 + p->remap_bytes;
    const int64_t full = ((N * p->state_bytes + 127) & ~int64_t(127));
    const int64_t want = full < 128 * 1024 ? full : 128 * 1024;
    // warp-sized teams need no named barriers; larger teams use barrier ids 1..8
    static const int cand[] = {32, 8, 4, 2, 1};
    int teams = 0;
    int64_t avail = 0;
    static const int forced = [] {
        const char *e = getenv("PRB_STREAM_TEAMS");  // tuning knob: 32, 8, 4, 2 or 1
        return e ? atoi(e) : 0;
    }();
    for (int c : cand) {
        if (forced && c > forced) continue;
        avail = smem_max - fixed - c * hist_bytes;
        if (avail >= want) {
            teams = c;
            break;
        }
    }
    if (!teams) {
        avail = smem_max - fixed - hist_bytes;
        if (avail < 4096) return fail("prb_stream", "joint histogram does not fit in shared memory");
        teams = 1;
    }
    p->teams = teams;
    int shift = 10;
    for (int t = teams; t > 1; t >>= 1) --shift;
    p->team_shift = shift;
    int64_t tile = (avail / p->state_bytes) & ~int64_t(127);
    if (tile >= N) {
        p->tile_cells = N;
        p->n_tiles = 1;
    } else {
        p->n_tiles = (int)ceil_div(N, tile);
        // equalise the tiles
        p->tile_cells = (ceil_div(N, p->n_tiles) + 127) & ~int64_t(127);
        p->n_tiles = (int)ceil_div(N, p->tile_cells);
    }
    p->tile_bytes = (p->tile_cells * p->state_bytes + 15) & ~int64_t(15);
    p->smem_bytes = fixed + p->tile_bytes + (int64_t)teams * hist_bytes;
    return 0;
}

__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void team_sync(int team, int team_shift)
{
    if (team_shift == 5)
        __syncwarp();
    else if (team_shift == 10)
        __syncthreads();
    else
        asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "r"(1 << team_shift) : "memory");
}

template <typename ST>
__device__ __forceinline__ void count_one(uint32_t w, const ST *__restrict__ st, int n_states,
                                          int *hist)
{
    const uint32_t s = st[w & ROW_MASK];
    if (s) atomicAdd(hist + (s - 1u) + (uint32_t)n_states * ((w >> 24) - 1u), 1);
}

// 8 stored entries: all state probes (shared loads) are issued before the first shared
// atomic so their latencies overlap.
template <typename ST>
__device__ __forceinline__ void count8(const uint4 qa, const uint4 qb, const ST *__restrict__ st,
                                       int n_states, int *hist)
{
    const uint32_t w[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = st[w[i] & ROW_MASK];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        if (s[i]) atomicAdd(hist + (s[i] - 1u) + (uint32_t)n_states * ((w[i] >> 24) - 1u), 1);
}

template <typename ST>
__global__ void __launch_bounds__(CTA_THREADS, 1)
k_stream(const uint32_t *__restrict__ packed, const int64_t *__restrict__ tp, int64_t P,
         int n_tiles, int64_t tile_cells, int64_t N, const ST *__restrict__ state, int n_states,
         int K1, int direct_C, int n_bins, int bins_pad, int64_t remap_bytes, int64_t tile_bytes,
         int32_t *__restrict__ hits, int *counters, int genes_per_item, int team_shift)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int *slots = (int *)smem_raw;
    uint16_t *remap = (uint16_t *)(smem_raw + SLOT_WORDS * 4);
    ST *tile_state = (ST *)(smem_raw + SLOT_WORDS * 4 + remap_bytes);
    int *hist_base = (int *)(smem_raw + SLOT_WORDS * 4 + remap_bytes + tile_bytes);
    const int tid = threadIdx.x;
    const int team_threads = 1 << team_shift;
    const int team = tid >> team_shift;
    const int tl = tid & (team_threads - 1);
    int *hist = hist_base + (int64_t)team * bins_pad;

    // shared bin (state + n_states*(code-1): class id fastest, so the popular low codes of
    // different classes fall in different banks) -> global bin (state-major, code fastest)
    for (int b = tid; b < n_bins; b += CTA_THREADS) {
        const int v = b / n_states, s = b - v * n_states;
        const int h = direct_C ? (s % direct_C) * K1 + s / direct_C : s;
        remap[b] = (uint16_t)(h * K1 + v);
    }
    for (int i = tl; i < bins_pad; i += team_threads) hist[i] = 0;
    const int n_items = (int)((P + genes_per_item - 1) / genes_per_item);

    for (int tt = 0; tt < n_tiles; ++tt) {
        const int tile = (int)((blockIdx.x + tt) % (unsigned)n_tiles);
        __syncthreads();  // every team has left the previous tile
        if (tid == 0) slots[SLOT_WORDS - 1] = *(volatile int *)(counters + tile) < n_items;
        __syncthreads();
        if (!slots[SLOT_WORDS - 1]) continue;  // tile already drained by other CTAs
        const int64_t cell0 = (int64_t)tile * tile_cells;
        const int64_t ncell = min(tile_cells, N - cell0);
        {
            // tile state: global -> shared, 128-bit copies (buffers are padded to 16 B)
            const int64_t nvec = (ncell * (int64_t)sizeof(ST) + 15) >> 4;
            const uint4 *src = (const uint4 *)(state + cell0);
            uint4 *dst = (uint4 *)tile_state;
            for (int64_t i = tid; i < nvec; i += CTA_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
        const ST *st = tile_state - cell0;  // indexed by absolute row id
        const int64_t *tp0 = tp + (int64_t)tile * P, *tp1 = tp0 + P;
        for (;;) {
            int item;
            if (team_shift == 5) {
                item = 0;
                if (tl == 0) item = atomicAdd(counters + tile, 1);
                item = __shfl_sync(0xffffffffu, item, 0);
            } else {
                if (tl == 0) slots[team] = atomicAdd(counters + tile, 1);
                team_sync(team, team_shift);
                item = slots[team];
            }
            if (item >= n_items) break;
            const int64_t g_end = min((int64_t)(item + 1) * genes_per_item, P);
            for (int64_t g = (int64_t)item * genes_per_item; g < g_end; ++g) {
                const int64_t pos = __ldg(tp0 + g), seg_end = __ldg(tp1 + g);
                if (pos >= seg_end) continue;
                // head (unaligned prefix, <= 3 words)
                const int64_t a0 = min((pos + 3) & ~int64_t(3), seg_end);
                if (pos + tl < a0) count_one<ST>(__ldg(packed + pos + tl), st, n_states, hist);
                // body: 128-bit loads, 4 in flight per thread
                const int64_t nvec = (seg_end - a0) >> 2;
                const uint4 *vp = (const uint4 *)(packed + a0);
                int64_t v = tl;
                for (; v + 3 * (int64_t)team_threads < nvec; v += 4 * (int64_t)team_threads) {
                    const uint4 q0 = ld_stream(vp + v);
                    const uint4 q1 = ld_stream(vp + v + team_threads);
                    const uint4 q2 = ld_stream(vp + v + 2 * team_threads);
                    const uint4 q3 = ld_stream(vp + v + 3 * team_threads);
                    count8<ST>(q0, q1, st, n_states, hist);
                    count8<ST>(q2, q3, st, n_states, hist);
                }
                for (; v + (int64_t)team_threads < nvec; v += 2 * (int64_t)team_threads) {
                    const uint4 q0 = ld_stream(vp + v);
                    const uint4 q1 = ld_stream(vp + v + team_threads);
                    count8<ST>(q0, q1, st, n_states, hist);
                }
                for (; v < nvec; v += team_threads) {
                    const uint4 q = ld_stream(vp + v);
                    count_one<ST>(q.x, st, n_states, hist);
                    count_one<ST>(q.y, st, n_states, hist);
                    count_one<ST>(q.z, st, n_states, hist);
                    count_one<ST>(q.w, st, n_states, hist);
                }
                // tail (<= 3 words)
                const int64_t a1 = a0 + (nvec << 2);
                if (a1 + tl < seg_end) count_one<ST>(__ldg(packed + a1 + tl), st, n_states, hist);
                team_sync(team, team_shift);
                // flush this gene segment's non-empty bins
                int32_t *out = hits + g * n_bins;
                for (int b = tl; b < n_bins; b += team_threads) {
                    const int c = hist[b];
                    if (c) {
                        atomicAdd(out + remap[b], c);
                        hist[b] = 0;
                    }
                }
                team_sync(team, team_shift);
            }
        }
    }
}

// tp[t][g] = first position of gene g whose row id is >= t * tile_cells
__global__ void k_tile_ptrs(const uint32_t *__restrict__ packed,
                            const int64_t *__restrict__ indptr, int64_t P, int64_t tile_cells,
                            int n_tiles, int64_t *__restrict__ tp)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= P * (int64_t)(n_tiles + 1)) return;
    const int64_t t = i / P, g = i - t * P;
    int64_t lo = indptr[g], hi = indptr[g + 1];
    if (t == 0) {
        tp[i] = lo;
        return;
    }
    if (t == n_tiles) {
        tp[i] = hi;
        return;
    }
    const uint32_t row0 = (uint32_t)(t * tile_cells);
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((packed[mid] & ROW_MASK) < row0)
            lo = mid + 1;
        else
            hi = mid;
    }
    tp[i] = lo;
}

__global__ void k_pack(const int32_t *__restrict__ rows, const uint8_t *__restrict__ codes,
                       int64_t nnz, uint32_t *__restrict__ packed)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += stride)
        packed[i] = ((uint32_t)codes[i] << 24) | ((uint32_t)rows[i] & ROW_MASK);
}

__global__ void k_unpack(const uint32_t *__restrict__ packed, int64_t nnz,
                         int32_t *__restrict__ rows, uint8_t *__restrict__ codes)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const uint32_t w = packed[i];
        rows[i] = (int32_t)(w & ROW_MASK);
        codes[i] = (uint8_t)(w >> 24);
    }
}

// One block per gene; 8 warps own contiguous slices of the cells so the packed
// words come out in ascending row order.
constexpr int DENSE_WARPS = 8;

template <bool FILL>
__global__ void __launch_bounds__(DENSE_WARPS * 32)
k_dense_csc(const uint8_t *__restrict__ codes, int64_t N, int64_t ld,
            const int64_t *__restrict__ indptr, int64_t *__restrict__ nnz_per_gene,
            uint32_t *__restrict__ packed)
{
    __shared__ int64_t warp_count[DENSE_WARPS];
    const int64_t g = blockIdx.x;
    const uint8_t *col = codes + g * ld;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slice = ceil_div(ceil_div(N, DENSE_WARPS), 32) * 32;
    const int64_t b = min(warp * slice, N), e = min(b + slice, N);
    int64_t cnt = 0;
    for (int64_t base = b; base < e; base += 32) {
        const int64_t i = base + lane;
        const bool nz = i < e && col[i] != 0;
        cnt += __popc(__ballot_sync(0xffffffffu, nz));
    }
    if (lane == 0) warp_count[warp] = cnt;
    __syncthreads();
    if (!FILL) {
        if (threadIdx.x == 0) {
            int64_t t = 0;
            for (int w = 0; w < DENSE_WARPS; ++w) t += warp_count[w];
            nnz_per_gene[g] = t;
        }
        return;
    }
    int64_t off = indptr[g];
    for (int w = 0; w < warp; ++w) off += warp_count[w];
    for (int64_t base = b; base < e; base += 32) {
        const int64_t i = base + lane;
        const uint32_t c = i < e ? col[i] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, c != 0);
        if (c) packed[off + __popc(m & ((1u << lane) - 1u))] = (c << 24) | (uint32_t)i;
        off += __popc(m);
    }
}

}  // namespace prb

using namespace prb;

extern "C" int prb_abi_version(void) { return PRB_ABI_VERSION; }
extern "C" const char *prb_last_error(void) { return prb::g_err; }

extern "C" int prb_device_info(int *sm_count, int *smem_optin_bytes)
{
    int dev = 0;
    PRB_CUDA(cudaGetDevice(&dev));
    if (sm_count) *sm_count = devinfo().sm_count;
    if (smem_optin_bytes) *smem_optin_bytes = devinfo().smem_optin;
    return 0;
}

extern "C" int prb_pack_csc(const int32_t *rows, const uint8_t *codes, int64_t nnz,
                            uint32_t *packed, prb_stream_t stream)
{
    if (nnz == 0) return 0;
    const int blocks = (int)hmin(ceil_div(nnz, 256), devinfo().sm_count * 16);
    k_pack<<<blocks, 256, 0, S(stream)>>>(rows, codes, nnz, packed);
    PRB_LAUNCH_CHECK("k_pack");
    return 0;
}

extern "C" int prb_unpack_csc(const uint32_t *packed, int64_t nnz, int32_t *rows, uint8_t *codes,
                              prb_stream_t stream)
{
    if (nnz == 0) return 0;
    const int blocks = (int)hmin(ceil_div(nnz, 256), devinfo().sm_count * 16);
    k_unpack<<<blocks, 256, 0, S(stream)>>>(packed, nnz, rows, codes);
    PRB_LAUNCH_CHECK("k_unpack");
    return 0;
}

extern "C" int prb_dense_count_nnz(const uint8_t *codes, int64_t N, int64_t P, int64_t ld,
                                   int64_t *nnz_per_gene, prb_stream_t stream)
{
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS && ld >= N, "bad shape");
    if (P == 0) return 0;
    k_dense_csc<false><<<(unsigned)P, DENSE_WARPS * 32, 0, S(stream)>>>(codes, N, ld, nullptr,
                                                                         nnz_per_gene, nullptr);
    PRB_LAUNCH_CHECK("k_dense_csc<count>");
    return 0;
}

extern "C" int prb_dense_fill_csc(const uint8_t *codes, int64_t N, int64_t P, int64_t ld,
                                  const int64_t *indptr, uint32_t *packed, prb_stream_t stream)
{
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS && ld >= N, "bad shape");
    if (P == 0) return 0;
    k_dense_csc<true><<<(unsigned)P, DENSE_WARPS * 32, 0, S(stream)>>>(codes, N, ld, indptr,
                                                                        nullptr, packed);
    PRB_LAUNCH_CHECK("k_dense_csc<fill>");
    return 0;
}

extern "C" int prb_stream_plan(int64_t N, int64_t n_states, int K1, int64_t *tile_cells,
                               int *n_tiles, int *teams_per_cta, int *state_bytes,
                               int64_t *smem_bytes)
{
    StreamPlan p;
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS, "N out of range for the packed row field");
    if (plan_stream(N, n_states, K1, &p)) return 1;
    if (tile_cells) *tile_cells = p.tile_cells;
    if (n_tiles) *n_tiles = p.n_tiles;
    if (teams_per_cta) *teams_per_cta = p.teams;
    if (state_bytes) *state_bytes = p.state_bytes;
    if (smem_bytes) *smem_bytes = p.smem_bytes;
    return 0;
}

extern "C" int prb_tile_ptrs(const uint32_t *packed, const int64_t *indptr, int64_t P,
                             int64_t tile_cells, int n_tiles, int64_t *tp, prb_stream_t stream)
{
    PRB_REQUIRE(P > 0 && n_tiles >= 1 && tile_cells > 0, "bad arguments");
    const int64_t n = P * (int64_t)(n_tiles + 1);
    k_tile_ptrs<<<(unsigned)ceil_div(n, 256), 256, 0, S(stream)>>>(packed, indptr, P, tile_cells,
                                                                  n_tiles, tp);
    PRB_LAUNCH_CHECK("k_tile_ptrs");
    return 0;
}

extern "C" int prb_stream_hits(const uint32_t *packed, const int64_t *tp, int64_t P, int64_t nnz,
                               const void *state, int64_t N, int64_t n_states, int K1,
                               int direct_C, int32_t *hits, int32_t *counters,
                               prb_stream_t stream)
{
    if (P == 0 || nnz == 0) return 0;
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS, "N out of range for the packed row field");
    PRB_REQUIRE(K1 >= 1 && K1 < PRB_MAX_CODES, "bad K");
    PRB_REQUIRE(direct_C == 0 || n_states % direct_C == 0, "direct layout needs n_states = C*(K-1)");
    StreamPlan p;
    if (plan_stream(N, n_states, K1, &p)) return 1;
    PRB_CUDA(cudaMemsetAsync(counters, 0, sizeof(int32_t) * p.n_tiles, S(stream)));
    // ~16 work items per team and tile
    const int64_t teams_per_tile = hmax(1, (int64_t)devinfo().sm_count * p.teams / p.n_tiles);
    const int gpi = (int)hmin(64, hmax(1, P / (teams_per_tile * 16)));
    const int64_t n_items = ceil_div(P, gpi) * p.n_tiles;
    const int grid = (int)hmax(1, hmin(devinfo().sm_count, ceil_div(n_items, p.teams)));
    auto launch = [&](auto kern, auto *st) -> int {
        PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)p.smem_bytes));
        kern<<<grid, CTA_THREADS, p.smem_bytes, S(stream)>>>(
            packed, tp, P, p.n_tiles, p.tile_cells, N, st, (int)n_states, K1, direct_C, p.n_bins,
            p.bins_pad, p.remap_bytes, p.tile_bytes, hits, counters, gpi, p.team_shift);
        PRB_LAUNCH_CHECK("k_stream");
        return 0;
    };
    if (p.state_bytes == 1) return launch(k_stream<uint8_t>, (const uint8_t *)state);
    return launch(k_stream<uint16_t>, (const uint16_t *)state);
}
