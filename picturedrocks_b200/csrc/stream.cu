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
constexpr int SLOT_WORDS = 64;  // per-team work slots at the front of smem

struct StreamPlan {
    int teams;
    int team_shift;  // log2(threads per team)
    int64_t bm_words;
    int64_t bm_words_pad;
    int bins_pad;
    bool bm_smem;
    int64_t smem_bytes;
};

static int plan_stream(int64_t N, int64_t n_bins, int use_bitmap, StreamPlan *p)
{
    const int64_t smem_max = devinfo().smem_optin;
    p->bm_words = use_bitmap ? ceil_div(N, 32) : 0;
    p->bm_words_pad = (p->bm_words + 3) & ~int64_t(3);
    p->bins_pad = (int)((n_bins + 3) & ~int64_t(3));
    if (p->bins_pad < 4) p->bins_pad = 4;
    const int64_t hist_bytes = (int64_t)p->bins_pad * 4;
    const int64_t fixed = SLOT_WORDS * 4;
    int64_t avail = smem_max - fixed - p->bm_words_pad * 4;
    p->bm_smem = use_bitmap && avail >= hist_bytes;
    if (!p->bm_smem) avail = smem_max - fixed;
    if (avail < hist_bytes) return fail("prb_stream", "joint histogram does not fit in shared memory");
    int teams = 1;
    while (teams < 32 && (int64_t)(teams * 2) * hist_bytes <= avail) teams *= 2;
    p->teams = teams;
    int shift = 10;  // 1024 threads
    for (int t = teams; t > 1; t >>= 1) --shift;
    p->team_shift = shift;
    p->smem_bytes = fixed + (p->bm_smem ? p->bm_words_pad * 4 : 0) + (int64_t)teams * hist_bytes;
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
        asm volatile("bar.sync %0, %1;" ::"r"(team), "r"(1 << team_shift) : "memory");
}

template <bool FILTER>
__device__ __forceinline__ void count_one(uint32_t w, const uint32_t *__restrict__ bmp,
                                          const uint16_t *__restrict__ state16, int *hist)
{
    bool hit = true;
    if (FILTER) {
        const uint32_t bits = bmp[(w & ROW_MASK) >> 5];
        hit = (bits >> (w & 31u)) & 1u;
    }
    if (hit) {
        const uint32_t base = state16 ? (uint32_t)__ldg(state16 + (w & ROW_MASK)) : 0u;
        atomicAdd(hist + base + (w >> 24) - 1u, 1);
    }
}

template <bool FILTER, bool BM_SMEM>
__global__ void __launch_bounds__(CTA_THREADS, 1)
k_stream(const uint32_t *__restrict__ packed, const int64_t *__restrict__ indptr,
         const int32_t *__restrict__ range_gene, int64_t n_ranges, int64_t range_len,
         int64_t nnz, const uint32_t *__restrict__ bitmap, int64_t bm_words,
         int64_t bm_words_pad, const uint16_t *__restrict__ state16, int n_bins, int bins_pad,
         int32_t *__restrict__ hits, int *work_counter, int team_shift)
{
    extern __shared__ __align__(16) uint32_t smem[];
    int *slots = (int *)smem;
    uint32_t *bm = smem + SLOT_WORDS;
    int *hist_base = (int *)(bm + (BM_SMEM ? bm_words_pad : 0));
    const int tid = threadIdx.x;
    const int team_threads = 1 << team_shift;
    const int team = tid >> team_shift;
    const int tl = tid & (team_threads - 1);
    int *hist = hist_base + (int64_t)team * bins_pad;

    if (FILTER && BM_SMEM) {
        // bitmap: global -> shared, 128-bit copies (bm_words_pad is a multiple of 4;
        // the global buffer is allocated padded).
        const uint4 *src = (const uint4 *)bitmap;
        uint4 *dst = (uint4 *)bm;
        for (int64_t i = tid; i < (bm_words_pad >> 2); i += CTA_THREADS) dst[i] = __ldg(src + i);
    }
    for (int i = tl; i < bins_pad; i += team_threads) hist[i] = 0;
    __syncthreads();
    const uint32_t *bmp = BM_SMEM ? bm : bitmap;
    (void)bm_words;

    for (;;) {
        int r;
        if (team_shift == 5) {
            r = 0;
            if (tl == 0) r = atomicAdd(work_counter, 1);
            r = __shfl_sync(0xffffffffu, r, 0);
        } else {
            if (tl == 0) slots[team] = atomicAdd(work_counter, 1);
            team_sync(team, team_shift);
            r = slots[team];
        }
        if (r >= n_ranges) break;
        int64_t pos = (int64_t)r * range_len;
        const int64_t end = min(pos + range_len, nnz);
        int g = range_gene[r];
        while (pos < end) {
            int64_t gend = __ldg(indptr + g + 1);
            while (gend <= pos) {
                ++g;
                gend = __ldg(indptr + g + 1);
            }
            const int64_t seg_end = min(gend, end);
            // head (unaligned prefix, <= 3 words)
            const int64_t a0 = min((pos + 3) & ~int64_t(3), seg_end);
            if (pos + tl < a0) count_one<FILTER>(__ldg(packed + pos + tl), bmp, state16, hist);
            // body: 128-bit loads, 4 in flight per thread
            const int64_t nvec = (seg_end - a0) >> 2;
            const uint4 *vp = (const uint4 *)(packed + a0);
            int64_t v = tl;
            for (; v + 3 * (int64_t)team_threads < nvec; v += 4 * (int64_t)team_threads) {
                const uint4 q0 = ld_stream(vp + v);
                const uint4 q1 = ld_stream(vp + v + team_threads);
                const uint4 q2 = ld_stream(vp + v + 2 * team_threads);
                const uint4 q3 = ld_stream(vp + v + 3 * team_threads);
                count_one<FILTER>(q0.x, bmp, state16, hist);
                count_one<FILTER>(q0.y, bmp, state16, hist);
                count_one<FILTER>(q0.z, bmp, state16, hist);
                count_one<FILTER>(q0.w, bmp, state16, hist);
                count_one<FILTER>(q1.x, bmp, state16, hist);
                count_one<FILTER>(q1.y, bmp, state16, hist);
                count_one<FILTER>(q1.z, bmp, state16, hist);
                count_one<FILTER>(q1.w, bmp, state16, hist);
                count_one<FILTER>(q2.x, bmp, state16, hist);
                count_one<FILTER>(q2.y, bmp, state16, hist);
                count_one<FILTER>(q2.z, bmp, state16, hist);
                count_one<FILTER>(q2.w, bmp, state16, hist);
                count_one<FILTER>(q3.x, bmp, state16, hist);
                count_one<FILTER>(q3.y, bmp, state16, hist);
                count_one<FILTER>(q3.z, bmp, state16, hist);
                count_one<FILTER>(q3.w, bmp, state16, hist);
            }
            for (; v < nvec; v += team_threads) {
                const uint4 q = ld_stream(vp + v);
                count_one<FILTER>(q.x, bmp, state16, hist);
                count_one<FILTER>(q.y, bmp, state16, hist);
                count_one<FILTER>(q.z, bmp, state16, hist);
                count_one<FILTER>(q.w, bmp, state16, hist);
            }
            // tail (<= 3 words)
            const int64_t a1 = a0 + (nvec << 2);
            if (a1 + tl < seg_end) count_one<FILTER>(__ldg(packed + a1 + tl), bmp, state16, hist);
            team_sync(team, team_shift);
            // flush this gene segment's non-empty bins
            int32_t *out = hits + (int64_t)g * n_bins;
            for (int b = tl; b < n_bins; b += team_threads) {
                const int c = hist[b];
                if (c) {
                    atomicAdd(out + b, c);
                    hist[b] = 0;
                }
            }
            team_sync(team, team_shift);
            pos = seg_end;
        }
    }
}

__global__ void k_range_genes(const int64_t *__restrict__ indptr, int64_t P, int64_t range_len,
                              int64_t n_ranges, int32_t *__restrict__ range_gene)
{
    const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n_ranges) return;
    const int64_t pos = r * range_len;
    // largest g in [0, P) with indptr[g] <= pos < indptr[g+1]  (upper_bound - 1)
    int64_t lo = 0, hi = P;  // search over indptr[1..P]
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (indptr[mid + 1] <= pos)
            lo = mid + 1;
        else
            hi = mid;
    }
    range_gene[r] = (int32_t)min(lo, P - 1);
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

extern "C" int prb_range_genes(const int64_t *indptr, int64_t P, int64_t nnz, int64_t range_len,
                               int64_t n_ranges, int32_t *range_gene, prb_stream_t stream)
{
    (void)nnz;
    if (n_ranges == 0) return 0;
    PRB_REQUIRE(P > 0 && range_len > 0, "bad arguments");
    k_range_genes<<<(unsigned)ceil_div(n_ranges, 256), 256, 0, S(stream)>>>(indptr, P, range_len,
                                                                            n_ranges, range_gene);
    PRB_LAUNCH_CHECK("k_range_genes");
    return 0;
}

extern "C" int prb_stream_smem_bytes(int64_t N, int64_t n_bins, int use_bitmap,
                                     int *teams_per_cta, int64_t *smem_bytes)
{
    StreamPlan p;
    if (plan_stream(N, n_bins, use_bitmap, &p)) return 1;
    if (teams_per_cta) *teams_per_cta = p.teams;
    if (smem_bytes) *smem_bytes = p.smem_bytes;
    return 0;
}

extern "C" int prb_stream_hits(const uint32_t *packed, const int64_t *indptr,
                               const int32_t *range_gene, int64_t n_ranges, int64_t range_len,
                               int64_t nnz, const uint32_t *bitmap, const uint16_t *state16,
                               int64_t N, int64_t n_bins, int32_t *hits, int32_t *work_counter,
                               prb_stream_t stream)
{
    if (n_ranges == 0 || nnz == 0) return 0;
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS, "N out of range for the packed row field");
    PRB_REQUIRE(n_bins > 0 && n_bins < (int64_t(1) << 30), "bad bin count");
    PRB_REQUIRE(n_ranges < (int64_t(1) << 31) - 65536, "too many ranges");
    StreamPlan p;
    if (plan_stream(N, n_bins, bitmap != nullptr, &p)) return 1;
    const int grid = (int)hmin(devinfo().sm_count, ceil_div(n_ranges, p.teams));
    auto launch = [&](auto kern) -> int {
        PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)p.smem_bytes));
        kern<<<grid, CTA_THREADS, p.smem_bytes, S(stream)>>>(
            packed, indptr, range_gene, n_ranges, range_len, nnz, bitmap, p.bm_words,
            p.bm_words_pad, state16, (int)n_bins, p.bins_pad, hits, work_counter, p.team_shift);
        PRB_LAUNCH_CHECK("k_stream");
        return 0;
    };
    if (bitmap == nullptr) return launch(k_stream<false, false>);
    if (p.bm_smem) return launch(k_stream<true, true>);
    return launch(k_stream<true, false>);
}
