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

// ---- numpy percentile(method="linear") on an ascending array ----------------
__device__ __forceinline__ float pct(const float *s, int64_t n, double q_py)
{
    const float q = __fdiv_rn((float)q_py, 100.0f);
    const float nm1 = (float)(n - 1);
    const float vi = __fmul_rn(nm1, q);
    if (vi >= nm1) return s[n - 1];
    const float pf = floorf(vi);
    const int64_t p = (int64_t)pf;
    const float g = __fsub_rn(vi, pf);
    const float a = s[p], b = s[p + 1];
    const float d = __fsub_rn(b, a);
    if (g >= 0.5f) return __fsub_rn(b, __fmul_rn(d, __fsub_rn(1.0f, g)));
    return __fadd_rn(a, __fmul_rn(d, g));
}

__device__ __forceinline__ double pct(const double *s, int64_t n, double q_py)
{
    const double q = __ddiv_rn(q_py, 100.0);
    const double nm1 = (double)(n - 1);
    const double vi = __dmul_rn(nm1, q);
    if (vi >= nm1) return s[n - 1];
    const double pf = floor(vi);
    const int64_t p = (int64_t)pf;
    const double g = __dsub_rn(vi, pf);
    const double a = s[p], b = s[p + 1];
    const double d = __dsub_rn(b, a);
    if (g >= 0.5) return __dsub_rn(b, __dmul_rn(d, __dsub_rn(1.0, g)));
    return __dadd_rn(a, __dmul_rn(d, g));
}

__device__ __forceinline__ double pct(const int64_t *s, int64_t n, double q_py)
{
    const double q = __ddiv_rn(q_py, 100.0);
    const double nm1 = (double)(n - 1);
    const double vi = __dmul_rn(nm1, q);
    if (vi >= nm1) return (double)s[n - 1];
    const double pf = floor(vi);
    const int64_t p = (int64_t)pf;
    const double g = __dsub_rn(vi, pf);
    const int64_t a = s[p], b = s[p + 1];
    const double d = (double)(b - a);
    if (g >= 0.5) return __dsub_rn((double)b, __dmul_rn(d, __dsub_rn(1.0, g)));
    return __dadd_rn((double)a, __dmul_rn(d, g));
}

// one thread per gene: <= k-1 binary searches on the sorted column
template <typename T, typename E>
__global__ void k_qd_edges(const T *__restrict__ sorted, int64_t N, int64_t P, int k,
                           E *__restrict__ edges, int32_t *__restrict__ n_edges)
{
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= P) return;
    const T *s = sorted + j * N;
    E *e = edges + j * (int64_t)(k - 1);
    const E cmax = (E)s[N - 1];
    int ne = 0;
    int64_t start = 0;
    for (int i = 0; i < k - 1; ++i) {
        if (ne > 0) {
            const E m = e[ne - 1];
            int64_t lo = start, hi = N;  // first index with s[idx] > m
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if ((E)s[mid] > m)
                    hi = mid;
                else
                    lo = mid + 1;
            }
            start = lo;
            if (cmax <= m || start >= N) break;
        }
        e[ne++] = pct(s + start, N - start, __ddiv_rn(100.0, (double)(k - i)));
    }
    n_edges[j] = ne;
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
