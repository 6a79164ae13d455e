// discretize.cu — quantile_discretize on device.
//
// Replaces the per-gene Python loop of quantile_discretize
// (picturedrocks/markers/mutualinformation/infoset.py:50-83): for every gene,
//   edge_0 = np.percentile(col, 100/k)
//   edge_i = np.percentile(col[col > edge_{i-1}], 100/(k-i))   (stop when the
//            column max <= last edge or nothing is left, :74-79)
//   codes  = np.digitize(col, edges, right=True)               (:81)
// np.percentile / np.digitize are third-party (numpy, un-pinned in the
// reference's requirements.txt:4); the arithmetic below follows numpy 2.3.5
// numpy/lib/_function_base_impl.py ('linear': virtual index (n-1)*q, _lerp with
// the t >= 0.5 branch) in the INPUT dtype: a float32 column keeps q, the virtual
// index, gamma and the lerp in float32.  Order statistics come from a full
// per-gene sort (the reference uses np.partition; same order statistics).
// Output: column-major uint8 codes in HBM (and, from stream.cu, the packed CSC).
#include <cub/device/device_segmented_radix_sort.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include "common.cuh"

namespace prb {

template <typename T>
__global__ void k_transpose(const T *__restrict__ X, int64_t N, int64_t P, T *__restrict__ out)
{
    __shared__ T tile[32][33];
    const int64_t p0 = (int64_t)blockIdx.x * 32, n0 = (int64_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t n = n0 + r, p = p0 + threadIdx.x;
        if (n < N && p < P) tile[r][threadIdx.x] = X[n * P + p];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int64_t p = p0 + r, n = n0 + threadIdx.x;
        if (n < N && p < P) out[p * N + n] = tile[threadIdx.x][r];
    }
}

struct SegOffset {
    int64_t N;
    __host__ __device__ int64_t operator()(int64_t seg) const { return seg * N; }
};

// ---- ascending columns behind an accessor -------------------------------------------
// Dense: the sorted column itself.  Sparse: the sorted stored values of a gene plus its
// implicit zeros — rank r of the full column is v[r] below the zeros, 0 inside the block of
// N - nnz implicit zeros (placed right after the negative values), v[r - (N - nnz)] above
// (infoset.py:66-67 densifies sparse input, so the zeros take part in the quantiles).
template <typename T>
struct DenseCol {
    const T *s;
    __device__ __forceinline__ T at(int64_t r) const { return s[r]; }
};

template <typename T>
struct SparseCol {
    const T *v;
    int64_t n_neg, n_zero;  // stored values < 0; implicit zeros
    __device__ __forceinline__ T at(int64_t r) const
    {
        if (r < n_neg) return v[r];
        if (r < n_neg + n_zero) return T(0);
        return v[r - n_zero];
    }
};

// ---- numpy percentile(method="linear") on ranks [start, start + n) of an ascending column
template <typename C>
__device__ __forceinline__ float pct(const C &c, int64_t start, int64_t n, double q_py, float)
{
    const float q = __fdiv_rn((float)q_py, 100.0f);
    const float nm1 = (float)(n - 1);
    const float vi = __fmul_rn(nm1, q);
    if (vi >= nm1) return c.at(start + n - 1);
    const float pf = floorf(vi);
    const int64_t p = (int64_t)pf;
    const float g = __fsub_rn(vi, pf);
    const float a = c.at(start + p), b = c.at(start + p + 1);
    const float d = __fsub_rn(b, a);
    if (g >= 0.5f) return __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g)));
    return __fadd_rn(a, __fmul_rn(d, g));
}

template <typename C>
__device__ __forceinline__ double pct(const C &c, int64_t start, int64_t n, double q_py, double)
{
    const double q = __ddiv_rn(q_py, 100.0);
    const double nm1 = (double)(n - 1);
    const double vi = __dmul_rn(nm1, q);
    if (vi >= nm1) return c.at(start + n - 1);
    const double pf = floor(vi);
    const int64_t p = (int64_t)pf;
    const double g = __dsub_rn(vi, pf);
    const double a = c.at(start + p), b = c.at(start + p + 1);
    const double d = __dsub_rn(b, a);
    if (g >= 0.5) return __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, g)));
    return __dadd_rn(a, __dmul_rn(d, g));
}

template <typename C>
__device__ __forceinline__ double pct(const C &c, int64_t start, int64_t n, double q_py, int64_t)
{
    const double q = __ddiv_rn(q_py, 100.0);
    const double nm1 = (double)(n - 1);
    const double vi = __dmul_rn(nm1, q);
    if (vi >= nm1) return (double)c.at(start + n - 1);
    const double pf = floor(vi);
    const int64_t p = (int64_t)pf;
    const double g = __dsub_rn(vi, pf);
    const int64_t a = c.at(start + p), b = c.at(start + p + 1);
    const double d = (double)(b - a);
    if (g >= 0.5) return __dsub_rn((double)b, __dmul_rn(d, __dsub_rn(1.0, g)));
    return __dadd_rn((double)a, __dmul_rn(d, g));
}

// <= k-1 edges of one ascending column of N values (infoset.py:69-80)
template <typename T, typename E, typename C>
__device__ __forceinline__ int qd_edges_col(const C &c, int64_t N, int k, E *e)
{
    const E cmax = (E)c.at(N - 1);
    int ne = 0;
    int64_t start = 0;
    for (int i = 0; i < k - 1; ++i) {
        if (ne > 0) {
            const E m = e[ne - 1];
            int64_t lo = start, hi = N;  // first rank with value > m
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((E)c.at(mid) > m)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            start = lo;
            if (cmax <= m || start >= N) break;
        }
        e[ne++] = pct(c, start, N - start, __ddiv_rn(100.0, (double)(k - i)), T());
    }
    return ne;
}

// one thread per gene: <= k-1 binary searches on the sorted column
template <typename T, typename E>
__global__ void k_qd_edges(const T *__restrict__ sorted, int64_t N, int64_t P, int k,
                           E *__restrict__ edges, int32_t *__restrict__ n_edges)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= P) return;
    n_edges[j] = qd_edges_col<T, E>(DenseCol<T>{sorted + j * N}, N, k, edges + j * (int64_t)(k - 1));
}

// sparse genes: sorted stored values of gene j at sorted_vals[indptr[j] - indptr[0] ...);
// zero_code[j] = code of the implicit zeros (#edges < 0; 0 for non-negative data)
template <typename T, typename E>
__global__ void k_qd_edges_sparse(const T *__restrict__ sorted_vals,
                                  const int64_t *__restrict__ indptr, int64_t N, int64_t P, int k,
                                  E *__restrict__ edges, int32_t *__restrict__ n_edges,
                                  int32_t *__restrict__ zero_code)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= P) return;
    const T *v = sorted_vals + (indptr[j] - indptr[0]);
    const int64_t nnz = indptr[j + 1] - indptr[j];
    int64_t lo = 0, hi = nnz;  // stored values < 0
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (v[mid] < T(0))
            lo = mid + 1;
        else
            hi = mid;
    }
    E *e = edges + j * (int64_t)(k - 1);
    const int ne = qd_edges_col<T, E>(SparseCol<T>{v, lo, N - nnz}, N, k, e);
    n_edges[j] = ne;
    int z = 0;
    while (z < ne && e[z] < E(0)) ++z;
    zero_code[j] = z;
}

// stored entries of a by-gene CSC -> packed words code<<24 | row, entries whose code is 0
// dropped (the reference's eliminate_zeros on the coded matrix, infoset.py:195).  One block
// per gene, 8 warps own contiguous slices so rows stay ascending; FILL = false counts.
constexpr int SP_WARPS = 8;

template <typename T, typename E, bool FILL>
__global__ void __launch_bounds__(SP_WARPS * 32)
k_qd_pack_sparse(const int32_t *__restrict__ rows, const T *__restrict__ vals,
                 const int64_t *__restrict__ indptr, int k, const E *__restrict__ edges,
                 const int32_t *__restrict__ n_edges, const int64_t *__restrict__ out_indptr,
                 int64_t *__restrict__ nnz_per_gene, uint32_t *__restrict__ packed)
{
    __shared__ int64_t warp_count[SP_WARPS];
    __shared__ E e[PRB_MAX_CODES];
    const int64_t j = blockIdx.x;
    const int ne = n_edges[j];
    for (int i = threadIdx.x; i < ne; i += blockDim.x) e[i] = edges[j * (int64_t)(k - 1) + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b0 = indptr[j], n = indptr[j + 1] - b0;
    const int64_t slice = ceil_div(ceil_div(n, SP_WARPS), 32) * 32;
    const int64_t b = min(warp * slice, n), en = min(b + slice, n);
    auto code_of = [&](int64_t i) -> uint32_t {
        const E x = (E)vals[b0 + i];
        int c = 0;
        while (c < ne && e[c] < x) ++c;
        return (uint32_t)c;
    };
    int64_t cnt = 0;
    for (int64_t base = b; base < en; base += 32) {
        const int64_t i = base + lane;
        const bool nz = i < en && code_of(i) != 0;
        cnt += __popc(__ballot_sync(0xffffffffu, nz));
    }
    if (lane == 0) warp_count[warp] = cnt;
    __syncthreads();
    if (!FILL) {
        if (threadIdx.x == 0) {
            int64_t t = 0;
            for (int w = 0; w < SP_WARPS; ++w) t += warp_count[w];
            nnz_per_gene[j] = t;
        }
        return;
    }
    int64_t off = out_indptr[j];
    for (int w = 0; w < warp; ++w) off += warp_count[w];
    for (int64_t base = b; base < en; base += 32) {
        const int64_t i = base + lane;
        const uint32_t c = i < en ? code_of(i) : 0u;
        const unsigned m = __ballot_sync(0xffffffffu, c != 0);
        if (c) packed[off + __popc(m & ((1u << lane) - 1u))] = (c << 24) | (uint32_t)rows[b0 + i];
        off += __popc(m);
    }
}

template <typename T, typename E>
__global__ void __launch_bounds__(256)
k_qd_digitize(const T *__restrict__ Xt, int64_t N, int k, const E *__restrict__ edges,
              const int32_t *__restrict__ n_edges, uint8_t *__restrict__ codes, int64_t ld)
{
    extern __shared__ __align__(16) unsigned char sh_raw[];
    E *e = reinterpret_cast<E *>(sh_raw);
    const int64_t j = blockIdx.y;
    const int ne = n_edges[j];
    for (int i = threadIdx.x; i < ne; i += blockDim.x) e[i] = edges[j * (int64_t)(k - 1) + i];
    __syncthreads();
    const T *col = Xt + j * N;
    uint8_t *out = codes + j * ld;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += stride) {
        const E x = (E)col[i];
        int c = 0;
        while (c < ne && e[c] < x) ++c;
        out[i] = (uint8_t)c;
    }
}

}  // namespace prb

using namespace prb;

extern "C" int prb_transpose(const void *X, int elem_bytes, int64_t N, int64_t P, void *out,
                             prb_stream_t stream)
{
    PRB_REQUIRE(elem_bytes == 4 || elem_bytes == 8, "elem_bytes must be 4 or 8");
    if (N == 0 || P == 0) return 0;
    const dim3 grid((unsigned)ceil_div(P, 32), (unsigned)ceil_div(N, 32)), block(32, 8);
    PRB_REQUIRE(ceil_div(N, 32) <= 65535, "N too large for one transpose launch");
    if (elem_bytes == 4)
        k_transpose<uint32_t><<<grid, block, 0, S(stream)>>>((const uint32_t *)X, N, P,
                                                             (uint32_t *)out);
    else
        k_transpose<uint64_t><<<grid, block, 0, S(stream)>>>((const uint64_t *)X, N, P,
                                                             (uint64_t *)out);
    PRB_LAUNCH_CHECK("k_transpose");
    return 0;
}

template <typename T>
static int sort_columns_t(const T *in, T *out, int64_t N, int64_t P, void *ws, size_t *ws_bytes,
                          cudaStream_t st)
{
    auto begins = thrust::make_transform_iterator(thrust::counting_iterator<int64_t>(0), SegOffset{N});
    PRB_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(ws, *ws_bytes, in, out, (int)(N * P), (int)P,
                                                     begins, begins + 1, 0, (int)sizeof(T) * 8,
                                                     st));
    return 0;
}

static int sort_dispatch(const void *in, void *out, int dtype, int64_t N, int64_t P, void *ws,
                         size_t *ws_bytes, cudaStream_t st)
{
    PRB_REQUIRE(N > 0 && P > 0 && N * P < (int64_t(1) << 31), "N*P must be in (0, 2^31)");
    switch (dtype) {
        case 0: return sort_columns_t((const float *)in, (float *)out, N, P, ws, ws_bytes, st);
        case 1: return sort_columns_t((const double *)in, (double *)out, N, P, ws, ws_bytes, st);
        case 2: return sort_columns_t((const int64_t *)in, (int64_t *)out, N, P, ws, ws_bytes, st);
    }
    return fail("prb_sort_columns", "dtype must be 0 (f32), 1 (f64) or 2 (i64)");
}

extern "C" int prb_sort_columns_workspace(int dtype, int64_t N, int64_t P, int64_t *bytes)
{
    size_t b = 0;
    if (sort_dispatch(nullptr, nullptr, dtype, N, P, nullptr, &b, 0)) return 1;
    *bytes = (int64_t)b + 256;
    return 0;
}

extern "C" int prb_sort_columns(const void *cols_in, void *cols_out, int dtype, int64_t N,
                                int64_t P, void *workspace, int64_t workspace_bytes,
                                prb_stream_t stream)
{
    PRB_REQUIRE(workspace != nullptr && workspace_bytes > 0, "workspace required");
    size_t b = (size_t)workspace_bytes;
    return sort_dispatch(cols_in, cols_out, dtype, N, P, workspace, &b, S(stream));
}

extern "C" int prb_qd_edges(const void *sorted, int dtype, int64_t N, int64_t P, int k,
                            void *edges, int32_t *n_edges, prb_stream_t stream)
{
    PRB_REQUIRE(k >= 2 && k <= PRB_MAX_CODES, "k must be in [2, 256]");
    PRB_REQUIRE(N > 0, "empty matrix");
    if (P == 0) return 0;
    const unsigned grid = (unsigned)ceil_div(P, 128);
    switch (dtype) {
        case 0:
            k_qd_edges<float, float><<<grid, 128, 0, S(stream)>>>((const float *)sorted, N, P, k,
                                                                  (float *)edges, n_edges);
            break;
        case 1:
            k_qd_edges<double, double><<<grid, 128, 0, S(stream)>>>((const double *)sorted, N, P,
                                                                    k, (double *)edges, n_edges);
            break;
        case 2:
            k_qd_edges<int64_t, double><<<grid, 128, 0, S(stream)>>>((const int64_t *)sorted, N,
                                                                     P, k, (double *)edges,
                                                                     n_edges);
            break;
        default:
            return fail("prb_qd_edges", "bad dtype");
    }
    PRB_LAUNCH_CHECK("k_qd_edges");
    return 0;
}

extern "C" int prb_qd_digitize(const void *Xt, int dtype, int64_t N, int64_t P, int k,
                               const void *edges, const int32_t *n_edges, uint8_t *codes,
                               int64_t ld, prb_stream_t stream)
{
    PRB_REQUIRE(k >= 2 && k <= PRB_MAX_CODES, "k must be in [2, 256]");
    PRB_REQUIRE(ld >= N, "ld < N");
    if (P == 0 || N == 0) return 0;
    PRB_REQUIRE(P <= 65535, "at most 65535 genes per digitize launch");
    const dim3 grid((unsigned)hmin(ceil_div(N, 256), 64), (unsigned)P);
    switch (dtype) {
        case 0:
            k_qd_digitize<float, float><<<grid, 256, (k - 1) * 4, S(stream)>>>(
                (const float *)Xt, N, k, (const float *)edges, n_edges, codes, ld);
            break;
        case 1:
            k_qd_digitize<double, double><<<grid, 256, (k - 1) * 8, S(stream)>>>(
                (const double *)Xt, N, k, (const double *)edges, n_edges, codes, ld);
            break;
        case 2:
            k_qd_digitize<int64_t, double><<<grid, 256, (k - 1) * 8, S(stream)>>>(
                (const int64_t *)Xt, N, k, (const double *)edges, n_edges, codes, ld);
            break;
        default:
            return fail("prb_qd_digitize", "bad dtype");
    }
    PRB_LAUNCH_CHECK("k_qd_digitize");
    return 0;
}

// ---- sparse input: by-gene CSC of raw values (cells x genes) -------------------------------
template <typename T>
static int sort_segments_t(const T *in, T *out, int64_t n_items, int64_t n_segs,
                           const int64_t *offsets, void *ws, size_t *ws_bytes, cudaStream_t st)
{
    PRB_CUDA(cub::DeviceSegmentedRadixSort::SortKeys(ws, *ws_bytes, in, out, (int)n_items,
                                                     (int)n_segs, offsets, offsets + 1, 0,
                                                     (int)sizeof(T) * 8, st));
    return 0;
}

static int sort_segments_dispatch(const void *in, void *out, int dtype, int64_t n_items,
                                  int64_t n_segs, const int64_t *offsets, void *ws,
                                  size_t *ws_bytes, cudaStream_t st)
{
    PRB_REQUIRE(n_items >= 0 && n_items < (int64_t(1) << 31) && n_segs > 0 &&
                    n_segs < (int64_t(1) << 31),
                "at most 2^31 values / segments per call");
    switch (dtype) {
        case 0: return sort_segments_t((const float *)in, (float *)out, n_items, n_segs, offsets, ws, ws_bytes, st);
        case 1: return sort_segments_t((const double *)in, (double *)out, n_items, n_segs, offsets, ws, ws_bytes, st);
        case 2: return sort_segments_t((const int64_t *)in, (int64_t *)out, n_items, n_segs, offsets, ws, ws_bytes, st);
    }
    return fail("prb_sort_segments", "dtype must be 0 (f32), 1 (f64) or 2 (i64)");
}

extern "C" int prb_sort_segments_workspace(int dtype, int64_t n_items, int64_t n_segs,
                                           int64_t *bytes)
{
    size_t b = 0;
    if (sort_segments_dispatch(nullptr, nullptr, dtype, n_items, n_segs, nullptr, nullptr, &b, 0))
        return 1;
    *bytes = (int64_t)b + 256;
    return 0;
}

extern "C" int prb_sort_segments(const void *vals_in, void *vals_out, int dtype, int64_t n_items,
                                 int64_t n_segs, const int64_t *offsets, void *workspace,
                                 int64_t workspace_bytes, prb_stream_t stream)
{
    PRB_REQUIRE(workspace != nullptr && workspace_bytes > 0, "workspace required");
    if (n_items == 0) return 0;
    size_t b = (size_t)workspace_bytes;
    return sort_segments_dispatch(vals_in, vals_out, dtype, n_items, n_segs, offsets, workspace, &b,
                                  S(stream));
}

extern "C" int prb_qd_edges_sparse(const void *sorted_vals, const int64_t *indptr, int dtype,
                                   int64_t N, int64_t P, int k, void *edges, int32_t *n_edges,
                                   int32_t *zero_code, prb_stream_t stream)
{
    PRB_REQUIRE(k >= 2 && k <= PRB_MAX_CODES, "k must be in [2, 256]");
    PRB_REQUIRE(N > 0, "empty matrix");
    if (P == 0) return 0;
    const unsigned grid = (unsigned)ceil_div(P, 128);
    switch (dtype) {
        case 0:
            k_qd_edges_sparse<float, float><<<grid, 128, 0, S(stream)>>>(
                (const float *)sorted_vals, indptr, N, P, k, (float *)edges, n_edges, zero_code);
            break;
        case 1:
            k_qd_edges_sparse<double, double><<<grid, 128, 0, S(stream)>>>(
                (const double *)sorted_vals, indptr, N, P, k, (double *)edges, n_edges, zero_code);
            break;
        case 2:
            k_qd_edges_sparse<int64_t, double><<<grid, 128, 0, S(stream)>>>(
                (const int64_t *)sorted_vals, indptr, N, P, k, (double *)edges, n_edges, zero_code);
            break;
        default:
            return fail("prb_qd_edges_sparse", "bad dtype");
    }
    PRB_LAUNCH_CHECK("k_qd_edges_sparse");
    return 0;
}

template <bool FILL>
static int pack_sparse(const int32_t *rows, const void *vals, const int64_t *indptr, int dtype,
                       int64_t P, int k, const void *edges, const int32_t *n_edges,
                       const int64_t *out_indptr, int64_t *nnz_per_gene, uint32_t *packed,
                       cudaStream_t st)
{
    PRB_REQUIRE(k >= 2 && k <= PRB_MAX_CODES, "k must be in [2, 256]");
    if (P == 0) return 0;
    const unsigned grid = (unsigned)P;
    switch (dtype) {
        case 0:
            k_qd_pack_sparse<float, float, FILL><<<grid, SP_WARPS * 32, 0, st>>>(
                rows, (const float *)vals, indptr, k, (const float *)edges, n_edges, out_indptr,
                nnz_per_gene, packed);
            break;
        case 1:
            k_qd_pack_sparse<double, double, FILL><<<grid, SP_WARPS * 32, 0, st>>>(
                rows, (const double *)vals, indptr, k, (const double *)edges, n_edges, out_indptr,
                nnz_per_gene, packed);
            break;
        case 2:
            k_qd_pack_sparse<int64_t, double, FILL><<<grid, SP_WARPS * 32, 0, st>>>(
                rows, (const int64_t *)vals, indptr, k, (const double *)edges, n_edges, out_indptr,
                nnz_per_gene, packed);
            break;
        default:
            return fail("prb_qd_pack_sparse", "bad dtype");
    }
    PRB_LAUNCH_CHECK("k_qd_pack_sparse");
    return 0;
}

extern "C" int prb_qd_count_sparse(const int32_t *rows, const void *vals, const int64_t *indptr,
                                   int dtype, int64_t P, int k, const void *edges,
                                   const int32_t *n_edges, int64_t *nnz_per_gene,
                                   prb_stream_t stream)
{
    return pack_sparse<false>(rows, vals, indptr, dtype, P, k, edges, n_edges, nullptr,
                              nnz_per_gene, nullptr, S(stream));
}

extern "C" int prb_qd_fill_sparse(const int32_t *rows, const void *vals, const int64_t *indptr,
                                  int dtype, int64_t P, int k, const void *edges,
                                  const int32_t *n_edges, const int64_t *out_indptr,
                                  uint32_t *packed, prb_stream_t stream)
{
    return pack_sparse<true>(rows, vals, indptr, dtype, P, k, edges, n_edges, out_indptr, nullptr,
                             packed, S(stream));
}
