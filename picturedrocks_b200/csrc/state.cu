// state.cu — joint-state ("master column") construction, resident on device.
//
// Replaces _sparse_make_master_col (infoset.py:431-442: X[:, cols] @ powers of
// 2^shift, eliminate_zeros, sort_indices) and the y-packing of
// SparseInformationSet.entropy (infoset.py:225-249).  Instead of a sparse
// packed column the device keeps, per cell,
//   state : uint8 / uint16, 0 = "every selected column is zero here" (not a master
//           row), else 1 + id of the cell's (y, master) joint state
// where ids are ORDER-PRESERVING w.r.t. the reference's packed id
// (y most significant, then cols[0], ..., cols[m-1]) so that entropies can be
// accumulated in the reference's ascending-bin order.
#include <vector>

#include "common.cuh"

namespace prb {

constexpr int SCATTER_BLOCKS_PER_SM = 4;

// col_begin / col_end: first / past-the-last word of every gene's column in `packed`
// (indptr[:-1] / indptr[1:] of a contiguous CSC, or the slots of the replicated copy)
__global__ void k_scatter_u8(const uint32_t *__restrict__ packed,
                             const int64_t *__restrict__ col_begin,
                             const int64_t *__restrict__ col_end,
                             const int32_t *__restrict__ gene_ptr, int64_t P,
                             const int32_t *__restrict__ newpos, uint8_t *__restrict__ xj)
{
    const int64_t g = *gene_ptr;
    if (g < 0 || g >= P) return;
    const int64_t b = col_begin[g], e = col_end[g];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = b + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += stride) {
        const uint32_t w = __ldg(packed + i);
        const uint32_t cell = w & ROW_MASK;
        xj[newpos ? (uint32_t)newpos[cell] : cell] = (uint8_t)(w >> 24);
    }
}

__global__ void k_scatter_or32(const uint32_t *__restrict__ packed,
                               const int64_t *__restrict__ col_begin,
                               const int64_t *__restrict__ col_end, int64_t gene, int shift_bits,
                               uint32_t *__restrict__ mpacked)
{
    const int64_t b = col_begin[gene], e = col_end[gene];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = b + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < e; i += stride) {
        const uint32_t w = __ldg(packed + i);
        atomicOr(mpacked + (w & ROW_MASK), (w >> 24) << shift_bits);
    }
}

// ---- direct single-column state -------------------------------------------
// state = 1 + (x_j - 1)*C + y  (class id fastest: bank-friendly shared histograms);
// hsize is indexed y*(K-1) + (x_j - 1) (the reference's bin order).
template <typename ST>
__global__ void __launch_bounds__(256)
k_state_direct(const uint8_t *__restrict__ xj, const int32_t *__restrict__ y, int64_t N, int K1,
               int C, ST *__restrict__ state, int32_t *__restrict__ hsize, int use_smem_hist)
{
    extern __shared__ int sh_hist[];
    const int n_hs = C * K1;
    if (use_smem_hist) {
        for (int i = threadIdx.x; i < n_hs; i += blockDim.x) sh_hist[i] = 0;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < N; cell += stride) {
        const int a = xj[cell];
        uint32_t s = 0;
        if (a) {
            const int yy = y ? y[cell] : 0;
            s = 1u + (uint32_t)((a - 1) * C + yy);
            atomicAdd(use_smem_hist ? sh_hist + yy * K1 + a - 1 : hsize + yy * K1 + a - 1, 1);
        }
        state[cell] = (ST)s;
    }
    if (use_smem_hist) {
        __syncthreads();
        for (int i = threadIdx.x; i < n_hs; i += blockDim.x)
            if (sh_hist[i]) atomicAdd(hsize + i, sh_hist[i]);
    }
}

__global__ void k_state_direct_sums(const int32_t *__restrict__ hsize, int C, int K1,
                                    int32_t *__restrict__ hs_begin, int32_t *__restrict__ hs_y,
                                    int32_t *__restrict__ hs_ysum, int32_t *__restrict__ ha,
                                    int32_t *work_counter)
{
    const int t = threadIdx.x, nt = blockDim.x;
    for (int yy = t; yy <= C; yy += nt) hs_begin[yy] = yy * K1;
    for (int h = t; h < C * K1; h += nt) hs_y[h] = h / K1;
    for (int yy = t; yy < C; yy += nt) {
        int s = 0;
        for (int a = 0; a < K1; ++a) s += hsize[yy * K1 + a];
        hs_ysum[yy] = s;
    }
    for (int a = t; a < K1; a += nt) {
        int s = 0;
        for (int yy = 0; yy < C; ++yy) s += hsize[yy * K1 + a];
        ha[a] = s;
    }
    (void)work_counter;
}

// ---- generic multi-column state through a presence table --------------------
__global__ void k_table_count(const uint32_t *__restrict__ mpacked, const int32_t *__restrict__ y,
                              int64_t N, int mbits, int32_t *__restrict__ table, int count_zero)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += stride) {
        const uint32_t m = mpacked ? mpacked[i] : 0u;
        if (m || count_zero) {
            const int64_t key = ((int64_t)(y ? y[i] : 0) << mbits) | m;
            atomicAdd(table + key, 1);
        }
    }
}

// small-table variant: privatised in shared memory (avoids same-address L2 atomics)
__global__ void __launch_bounds__(256)
k_table_count_smem(const uint32_t *__restrict__ mpacked, const int32_t *__restrict__ y, int64_t N,
                   int mbits, int32_t *__restrict__ table, int n_keys, int count_zero)
{
    extern __shared__ int sh_tab[];
    for (int i = threadIdx.x; i < n_keys; i += blockDim.x) sh_tab[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += stride) {
        const uint32_t m = mpacked ? mpacked[i] : 0u;
        if (m || count_zero) atomicAdd(sh_tab + ((((y ? y[i] : 0)) << mbits) | m), 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_keys; i += blockDim.x)
        if (sh_tab[i]) atomicAdd(table + i, sh_tab[i]);
}

// single block: ordered compaction of the present keys -> ranks.
// table[key] (count) is replaced by rank(key) (or -1 if absent).
__global__ void __launch_bounds__(1024)
k_table_rank(int32_t *__restrict__ table, int64_t n_keys, int C, int mbits,
             int32_t *__restrict__ hsize, int32_t *__restrict__ hs_begin,
             int32_t *__restrict__ hs_y, int32_t *__restrict__ hs_ysum,
             int32_t *__restrict__ n_hs_out, int64_t hsize_cap, int32_t *work_counter)
{
    __shared__ int warp_tot[32];
    __shared__ int base_s;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if (t == 0) base_s = 0;
    for (int yy = t; yy < C; yy += 1024) hs_ysum[yy] = 0;
    __syncthreads();
    for (int64_t k0 = 0; k0 < n_keys; k0 += 1024) {
        const int64_t key = k0 + t;
        const int cnt = key < n_keys ? table[key] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, cnt > 0);
        if (lane == 0) warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = base_s;
        for (int w = 0; w < warp; ++w) off += warp_tot[w];
        const int r = off + __popc(m & ((1u << lane) - 1u));
        if (key < n_keys) {
            // first key of a class: its rank is the class's hs_begin
            if ((key & ((int64_t(1) << mbits) - 1)) == 0) hs_begin[key >> mbits] = r;
            if (cnt > 0) {
                if (r < hsize_cap) {
                    hsize[r] = cnt;
                    hs_y[r] = (int32_t)(key >> mbits);
                }
                atomicAdd(hs_ysum + (key >> mbits), cnt);
                table[key] = r;
            } else {
                table[key] = -1;
            }
        }
        __syncthreads();
        if (t == 0) {
            int tot = 0;
            for (int w = 0; w < 32; ++w) tot += warp_tot[w];
            base_s += tot;
        }
        __syncthreads();
    }
    if (t == 0) {
        hs_begin[C] = base_s;
        *n_hs_out = base_s;
        (void)work_counter;
    }
}

template <typename ST>
__global__ void __launch_bounds__(256)
k_table_assign(const uint32_t *__restrict__ mpacked, const int32_t *__restrict__ y, int64_t N,
               int mbits, const int32_t *__restrict__ table, ST *__restrict__ state)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t cell = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; cell < N; cell += stride) {
        const uint32_t m = mpacked[cell];
        uint32_t s = 0;
        if (m) {
            const int64_t key = ((int64_t)(y ? y[cell] : 0) << mbits) | m;
            s = 1u + (uint32_t)table[key];
        }
        state[cell] = (ST)s;
    }
}

// one warp: Σ T[count] over the table in ascending key order, zero bins skipped.
__global__ void k_table_entropy(const int32_t *__restrict__ table, int64_t n_keys,
                                const double *__restrict__ term, double *__restrict__ out)
{
    const int lane = threadIdx.x;
    double h = 0.0;
    for (int64_t k0 = 0; k0 < n_keys; k0 += 32) {
        const int64_t key = k0 + lane;
        const int cnt = key < n_keys ? table[key] : 0;
        const double tv = cnt > 0 ? term[cnt] : 0.0;
        unsigned nz = __ballot_sync(0xffffffffu, cnt > 0);
        while (nz) {
            const int l = __ffs(nz) - 1;
            nz &= nz - 1;
            h = __dadd_rn(h, __shfl_sync(0xffffffffu, tv, l));
        }
    }
    if (lane == 0) *out = h;
}

}  // namespace prb

using namespace prb;

static inline int cell_grid(int64_t N, int threads)
{
    return (int)hmax(1, hmin(ceil_div(N, threads), (int64_t)devinfo().sm_count * 8));
}

extern "C" int prb_scatter_gene_u8(const uint32_t *packed, const int64_t *col_begin,
                                   const int64_t *col_end, const int32_t *gene_ptr, int64_t P,
                                   const int32_t *newpos, uint8_t *xj, prb_stream_t stream)
{
    k_scatter_u8<<<devinfo().sm_count * SCATTER_BLOCKS_PER_SM, 256, 0, S(stream)>>>(
        packed, col_begin, col_end, gene_ptr, P, newpos, xj);
    PRB_LAUNCH_CHECK("k_scatter_u8");
    return 0;
}

extern "C" int prb_scatter_gene_or32(const uint32_t *packed, const int64_t *col_begin,
                                     const int64_t *col_end, int64_t gene, int shift_bits,
                                     uint32_t *mpacked, prb_stream_t stream)
{
    PRB_REQUIRE(shift_bits >= 0 && shift_bits < 32, "shift out of range");
    k_scatter_or32<<<devinfo().sm_count * SCATTER_BLOCKS_PER_SM, 256, 0, S(stream)>>>(
        packed, col_begin, col_end, gene, shift_bits, mpacked);
    PRB_LAUNCH_CHECK("k_scatter_or32");
    return 0;
}

extern "C" int prb_state_direct(const uint8_t *xj, const int32_t *y, int64_t N, int C, int K,
                                void *state, int state_bytes, int32_t *hsize, int32_t *hs_begin,
                                int32_t *hs_y, int32_t *hs_ysum, int32_t *ha, prb_stream_t stream)
{
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS, "N out of range");
    PRB_REQUIRE(C >= 1 && K >= 2 && K <= PRB_MAX_CODES, "bad C or K");
    const int K1 = K - 1;
    const int64_t n_hs = (int64_t)C * K1;
    PRB_REQUIRE(n_hs * K1 <= 65536, "C*(K-1)^2 exceeds 65536 histogram bins");
    PRB_REQUIRE(state_bytes == 2 || (state_bytes == 1 && n_hs + 1 <= 256), "state width too small");
    PRB_CUDA(cudaMemsetAsync(hsize, 0, sizeof(int32_t) * n_hs, S(stream)));
    const int use_smem = n_hs <= 8192;
    const int grid = cell_grid(N, 256);
    const size_t sm = use_smem ? n_hs * 4 : 0;
    if (state_bytes == 1)
        k_state_direct<uint8_t><<<grid, 256, sm, S(stream)>>>(xj, y, N, K1, C, (uint8_t *)state,
                                                              hsize, use_smem);
    else
        k_state_direct<uint16_t><<<grid, 256, sm, S(stream)>>>(xj, y, N, K1, C, (uint16_t *)state,
                                                               hsize, use_smem);
    PRB_LAUNCH_CHECK("k_state_direct");
    k_state_direct_sums<<<1, 256, 0, S(stream)>>>(hsize, C, K1, hs_begin, hs_y, hs_ysum, ha,
                                                  nullptr);
    PRB_LAUNCH_CHECK("k_state_direct_sums");
    return 0;
}

static int table_count(const uint32_t *mpacked, const int32_t *y, int64_t N, int mbits,
                       int64_t n_keys, int32_t *table, int count_zero, cudaStream_t st)
{
    PRB_CUDA(cudaMemsetAsync(table, 0, sizeof(int32_t) * n_keys, st));
    if (n_keys <= 8192) {
        k_table_count_smem<<<cell_grid(N, 256 * 8), 256, n_keys * 4, st>>>(
            mpacked, y, N, mbits, table, (int)n_keys, count_zero);
    } else {
        k_table_count<<<cell_grid(N, 256), 256, 0, st>>>(mpacked, y, N, mbits, table, count_zero);
    }
    PRB_LAUNCH_CHECK("k_table_count");
    return 0;
}

extern "C" int prb_state_table(const uint32_t *mpacked, const int32_t *y, int64_t N, int C,
                               int mbits, int32_t *table, int32_t *hsize, int32_t *hs_begin,
                               int32_t *hs_y, int32_t *hs_ysum, int32_t *n_hs_out,
                               int64_t hsize_cap, prb_stream_t stream)
{
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS, "N out of range");
    PRB_REQUIRE(C >= 1, "bad C");
    PRB_REQUIRE(mbits >= 1 && mbits <= 30, "master column needs 1..30 bits");
    const int64_t n_keys = (int64_t)C << mbits;
    PRB_REQUIRE(n_keys <= (int64_t(1) << 24), "joint state space too large for the presence table");
    if (table_count(mpacked, y, N, mbits, n_keys, table, 0, S(stream))) return 1;
    k_table_rank<<<1, 1024, 0, S(stream)>>>(table, n_keys, C, mbits, hsize, hs_begin, hs_y,
                                            hs_ysum, n_hs_out, hsize_cap, nullptr);
    PRB_LAUNCH_CHECK("k_table_rank");
    return 0;
}

extern "C" int prb_state_assign(const uint32_t *mpacked, const int32_t *y, int64_t N, int mbits,
                                const int32_t *table, void *state, int state_bytes,
                                prb_stream_t stream)
{
    PRB_REQUIRE(state_bytes == 1 || state_bytes == 2, "state_bytes must be 1 or 2");
    if (state_bytes == 1)
        k_table_assign<uint8_t><<<cell_grid(N, 256), 256, 0, S(stream)>>>(mpacked, y, N, mbits,
                                                                          table, (uint8_t *)state);
    else
        k_table_assign<uint16_t><<<cell_grid(N, 256), 256, 0, S(stream)>>>(
            mpacked, y, N, mbits, table, (uint16_t *)state);
    PRB_LAUNCH_CHECK("k_table_assign");
    return 0;
}

extern "C" int prb_entropy_keys(const uint32_t *mpacked, const int32_t *y, int64_t N, int mbits,
                                int64_t n_keys, int32_t *table, const double *term_table,
                                double *out, prb_stream_t stream)
{
    PRB_REQUIRE(N > 0, "empty");
    PRB_REQUIRE(n_keys >= 1 && n_keys <= (int64_t(1) << 24), "key space too large");
    if (table_count(mpacked, y, N, mbits, n_keys, table, 1, S(stream))) return 1;
    k_table_entropy<<<1, 32, 0, S(stream)>>>(table, n_keys, term_table, out);
    PRB_LAUNCH_CHECK("k_table_entropy");
    return 0;
}

// ---- HOST-buffer drop-in for _sparse_entropy (infoset.py:413-428) -----------------------
namespace prb {
__global__ void k_count_values(const int64_t *__restrict__ data, int64_t nnz, int64_t max_val,
                               int32_t *__restrict__ table, int *bad)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nnz; i += stride) {
        const int64_t v = data[i];
        if (v < 0 || v > max_val)
            *bad = 1;
        else
            atomicAdd(table + v, 1);
    }
}
__global__ void k_add_zero_bin(int32_t *table, int32_t add) { table[0] += add; }
}  // namespace prb

extern "C" int prb_sparse_entropy_host(const int64_t *data, int64_t nnz, int64_t n_rows,
                                       int64_t max_val, double *out)
{
    // counts[0] = n_rows; every stored entry moves one count from bin 0 to bin data[i];
    // h = sum over the bins in ascending order of -(c/n_rows) * log2(c/n_rows)
    PRB_REQUIRE(n_rows > 0 && nnz >= 0 && nnz <= n_rows, "bad sizes");
    PRB_REQUIRE(max_val >= 0 && max_val < (int64_t(1) << 24), "value space too large");
    PRB_REQUIRE(out != nullptr && (nnz == 0 || data != nullptr), "null pointer");
    const int64_t n_keys = max_val + 1;
    int64_t *d_data = nullptr;
    int32_t *d_table = nullptr;
    double *d_term = nullptr, *d_out = nullptr;
    int *d_bad = nullptr;
    std::vector<double> h_term((size_t)n_rows + 1);
    if (prb_term_table(n_rows, h_term.data())) return 1;
    auto cleanup = [&]() {
        cudaFree(d_data);
        cudaFree(d_table);
        cudaFree(d_term);
        cudaFree(d_out);
        cudaFree(d_bad);
    };
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
        return r == cudaSuccess;
    };
    int bad = 0;
    if (ok(cudaMalloc(&d_data, sizeof(int64_t) * (size_t)(nnz ? nnz : 1))) &&
        ok(cudaMalloc(&d_table, sizeof(int32_t) * (size_t)n_keys)) &&
        ok(cudaMalloc(&d_term, sizeof(double) * (size_t)(n_rows + 1))) &&
        ok(cudaMalloc(&d_out, sizeof(double))) && ok(cudaMalloc(&d_bad, sizeof(int))) &&
        ok(cudaMemcpy(d_data, data, sizeof(int64_t) * (size_t)nnz, cudaMemcpyHostToDevice)) &&
        ok(cudaMemcpy(d_term, h_term.data(), sizeof(double) * (size_t)(n_rows + 1),
                      cudaMemcpyHostToDevice)) &&
        ok(cudaMemset(d_table, 0, sizeof(int32_t) * (size_t)n_keys)) &&
        ok(cudaMemset(d_bad, 0, sizeof(int)))) {
        if (nnz)
            k_count_values<<<cell_grid(nnz, 256), 256>>>(d_data, nnz, max_val, d_table, d_bad);
        k_add_zero_bin<<<1, 1>>>(d_table, (int32_t)(n_rows - nnz));
        k_table_entropy<<<1, 32>>>(d_table, n_keys, d_term, d_out);
        ok(cudaGetLastError());
        ok(cudaMemcpy(out, d_out, sizeof(double), cudaMemcpyDeviceToHost));
        ok(cudaMemcpy(&bad, d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    }
    cleanup();
    if (e != cudaSuccess) return fail(__func__, cudaGetErrorString(e));
    if (bad) return fail(__func__, "a value lies outside [0, max_val]");
    return 0;
}
