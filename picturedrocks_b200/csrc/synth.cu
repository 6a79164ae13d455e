// synth.cu — synthetic discretised expression matrices, generated on device.
//
// Bench / test utility (BASELINE.json names shapes only; the recipe is ours).
// Counter-based: every (gene, cell) entry is a pure function of (seed, gene,
// cell, y[cell]) in 64-bit integer arithmetic, so picturedrocks_b200/synth.py
// reproduces any subset of genes bit-for-bit on the CPU for parity checks.
//
//   gene key   gk  = mix(seed ^ gene_eff * GOLD)
//   density    base_q24 = ((u^8 * dmax_q24) >> 32) + dmin_q24,  u = hi32(mix(gk ^ A))
//   class mod  ck  = mix(gk ^ (B + class * GOLD));  thr = min(base * MULT[ck & 7] / 4, cap)
//   entry      h   = mix(gk + cell * M2);  non-zero iff (h >> 40) < thr
//   code       1 + #{t in 1..K-2 : ((h >> 8) & 0xffff) >= cum[(ck >> 3) & 3][t-1]}
//   planted    gene % 4001 == 17 duplicates gene-1; gene % 5003 == 29 is all-zero.
#include "common.cuh"

namespace prb {

constexpr uint64_t GOLD = 0x9E3779B97F4A7C15ull;
constexpr uint64_t M2 = 0xD1B54A32D192ED03ull;
constexpr uint64_t SALT_A = 0x1234567ull;
constexpr uint64_t SALT_B = 0xC1A55ull;
constexpr uint64_t SALT_Y = 0xA5A5A5A5A5A5A5A5ull;

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

__global__ void k_synth_labels(int32_t *__restrict__ y, int64_t N, int C, uint64_t seed)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += stride)
        y[i] = (int32_t)((mix64((seed ^ SALT_Y) + (uint64_t)i * GOLD) >> 33) % (uint64_t)C);
}

constexpr int SYN_WARPS = 8;

template <bool FILL>
__global__ void __launch_bounds__(SYN_WARPS * 32)
k_synth(const int32_t *__restrict__ y, int64_t N, int64_t gene_lo, int C, int K,
        uint32_t mean_q24, uint64_t seed, const uint32_t *__restrict__ cum,
        const int64_t *__restrict__ indptr, int64_t *__restrict__ nnz_per_gene,
        uint32_t *__restrict__ packed)
{
    __shared__ uint32_t thr[256];
    __shared__ uint8_t pidx[256];
    __shared__ int64_t warp_count[SYN_WARPS];
    const int64_t gl = blockIdx.x;
    const int64_t gene = gene_lo + gl;
    const int64_t gene_eff = (gene % 4001 == 17 && gene > 0) ? gene - 1 : gene;
    const bool all_zero = gene_eff % 5003 == 29;
    const uint64_t gk = mix64(seed ^ ((uint64_t)gene_eff * GOLD));
    {
        const uint64_t u32 = mix64(gk ^ SALT_A) >> 32;
        const uint64_t u2 = (u32 * u32) >> 32, u4 = (u2 * u2) >> 32, u8 = (u4 * u4) >> 32;
        const uint64_t cap = (uint64_t)(0.95 * 16777216.0);
        uint64_t dmax = 9ull * mean_q24;
        if (dmax > (uint64_t)(0.9 * 16777216.0)) dmax = (uint64_t)(0.9 * 16777216.0);
        const uint64_t base = ((u8 * dmax) >> 32) + (mean_q24 >> 4);
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const uint64_t ck = mix64(gk ^ (SALT_B + (uint64_t)c * GOLD));
            const uint64_t mult = (0xC8644321ull >> (4 * (ck & 7))) & 15;  // {1,2,3,4,4,6,8,12}
            uint64_t t = base * mult / 4;
            if (t > cap) t = cap;
            thr[c] = all_zero ? 0u : (uint32_t)t;
            pidx[c] = (uint8_t)((ck >> 3) & 3);
        }
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t slice = ceil_div(ceil_div(N, SYN_WARPS), 32) * 32;
    const int64_t b = min(warp * slice, N), e = min(b + slice, N);
    auto code_at = [&](int64_t cell) -> uint32_t {
        const int c = y[cell];
        const uint64_t h = mix64(gk + (uint64_t)cell * M2);
        if ((uint32_t)(h >> 40) >= thr[c]) return 0u;
        const uint32_t r16 = (uint32_t)(h >> 8) & 0xffffu;
        const uint32_t *cm = cum + pidx[c] * (K - 2);
        uint32_t code = 1;
        for (int t = 0; t < K - 2; ++t) code += r16 >= cm[t];
        return code;
    };
    int64_t cnt = 0;
    for (int64_t base = b; base < e; base += 32) {
        const int64_t i = base + lane;
        const bool nz = i < e && code_at(i) != 0;
        cnt += __popc(__ballot_sync(0xffffffffu, nz));
    }
    if (lane == 0) warp_count[warp] = cnt;
    __syncthreads();
    if (!FILL) {
        if (threadIdx.x == 0) {
            int64_t t = 0;
            for (int w = 0; w < SYN_WARPS; ++w) t += warp_count[w];
            nnz_per_gene[gl] = t;
        }
        return;
    }
    int64_t off = indptr[gl];
    for (int w = 0; w < warp; ++w) off += warp_count[w];
    for (int64_t base = b; base < e; base += 32) {
        const int64_t i = base + lane;
        const uint32_t c = i < e ? code_at(i) : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, c != 0);
        if (c) packed[off + __popc(m & ((1u << lane) - 1u))] = (c << 24) | (uint32_t)i;
        off += __popc(m);
    }
}

}  // namespace prb

using namespace prb;

extern "C" int prb_synth_labels(int32_t *y, int64_t N, int C, uint64_t seed, prb_stream_t stream)
{
    PRB_REQUIRE(N > 0 && C >= 1 && C <= 256, "bad N or C");
    k_synth_labels<<<(unsigned)hmin(ceil_div(N, 256), 4096), 256, 0, S(stream)>>>(y, N, C,
                                                                                          seed);
    PRB_LAUNCH_CHECK("k_synth_labels");
    return 0;
}

static int synth_check(int64_t N, int64_t P, int C, int K)
{
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS, "N out of range");
    PRB_REQUIRE(P > 0 && P < (int64_t(1) << 31), "P out of range");
    PRB_REQUIRE(C >= 1 && C <= 256, "C must be in [1, 256]");
    PRB_REQUIRE(K >= 2 && K <= PRB_MAX_CODES, "K must be in [2, 256]");
    return 0;
}

extern "C" int prb_synth_count(const int32_t *y, int64_t N, int64_t gene_lo, int64_t P, int C,
                               int K, uint32_t mean_density_q24, uint64_t seed,
                               const uint32_t *cum, int64_t *nnz_per_gene, prb_stream_t stream)
{
    if (synth_check(N, P, C, K)) return 1;
    k_synth<false><<<(unsigned)P, SYN_WARPS * 32, 0, S(stream)>>>(
        y, N, gene_lo, C, K, mean_density_q24, seed, cum, nullptr, nnz_per_gene, nullptr);
    PRB_LAUNCH_CHECK("k_synth<count>");
    return 0;
}

extern "C" int prb_synth_fill(const int32_t *y, int64_t N, int64_t gene_lo, int64_t P, int C,
                              int K, uint32_t mean_density_q24, uint64_t seed,
                              const uint32_t *cum, const int64_t *indptr, uint32_t *packed,
                              prb_stream_t stream)
{
    if (synth_check(N, P, C, K)) return 1;
    k_synth<true><<<(unsigned)P, SYN_WARPS * 32, 0, S(stream)>>>(
        y, N, gene_lo, C, K, mean_density_q24, seed, cum, indptr, nullptr, packed);
    PRB_LAUNCH_CHECK("k_synth<fill>");
    return 0;
}
