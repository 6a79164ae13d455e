// ingest.cu — a canonical CSC matrix of bin codes, as the caller holds it, -> packed words.
//
// Replaces the host-side canonicalisation of SparseInformationSet.__init__
// (picturedrocks/markers/mutualinformation/infoset.py:188-200: csc_matrix(X),
// sum_duplicates, eliminate_zeros, the has_canonical_format assert) for input that already
// IS canonical CSC — the normal case: the pass over the matrix happens on the device, chunk
// by chunk while the next chunk is still on the PCIe bus, instead of on one host core.
//   prb_pack_chunk    row ids (int32 | int64) + codes (any integer width) of a range of
//                     stored entries -> packed words code << 24 | row; flags entries the
//                     packed form cannot hold or the reference would have dropped
//   prb_check_columns rows strictly ascending inside every gene (sorted, no duplicates)
// A matrix that fails either check goes through scipy's canonicalisation on the host, exactly
// as in the reference, and is packed again.
#include "common.cuh"

namespace prb {

enum { BAD_ROW = 1, BAD_CODE = 2, ZERO_CODE = 4, UNSORTED = 8 };

__device__ __forceinline__ int64_t load_int(const void *p, int64_t i, int bytes, int is_signed)
{
    switch (bytes) {
        case 1: return is_signed ? (int64_t)((const int8_t *)p)[i] : (int64_t)((const uint8_t *)p)[i];
        case 2: return is_signed ? (int64_t)((const int16_t *)p)[i] : (int64_t)((const uint16_t *)p)[i];
        case 4: return is_signed ? (int64_t)((const int32_t *)p)[i] : (int64_t)((const uint32_t *)p)[i];
        default: return ((const int64_t *)p)[i];
    }
}

__global__ void __launch_bounds__(256)
k_pack_chunk(const void *__restrict__ rows, int row_bytes, const void *__restrict__ codes,
             int code_bytes, int code_signed, int64_t n, int64_t N, uint32_t *__restrict__ packed,
             int32_t *__restrict__ flags)
{
    int bad = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t r = row_bytes == 4 ? (int64_t)((const int32_t *)rows)[i] : ((const int64_t *)rows)[i];
        const int64_t c = load_int(codes, i, code_bytes, code_signed);
        if (r < 0 || r >= N) bad |= BAD_ROW;
        if (c < 0 || c > 255) bad |= BAD_CODE;
        if (c == 0) bad |= ZERO_CODE;
        packed[i] = ((uint32_t)c << 24) | ((uint32_t)r & ROW_MASK);
    }
    if (bad) atomicOr(flags, bad);
}

__global__ void __launch_bounds__(256)
k_check_columns(const uint32_t *__restrict__ packed, const int64_t *__restrict__ indptr, int64_t P,
                int32_t *__restrict__ flags)
{
    // one warp per gene (grid-stride): every entry is compared with its successor
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    int bad = 0;
    for (int64_t g = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); g < P; g += warps) {
        const int64_t b = indptr[g], e = indptr[g + 1];
        for (int64_t i = b + lane; i + 1 < e; i += 32)
            if ((packed[i] & ROW_MASK) >= (packed[i + 1] & ROW_MASK)) bad = UNSORTED;
    }
    if (bad) atomicOr(flags, bad);
}

// ---- CSR (cells-major, how AnnData stores counts) -> by-gene CSC, a counting transpose ------
// Replaces scipy's csc_matrix(X) of infoset.py:191 for CSR input.  The rows are cut into B
// blocks; T[b][c] = stored entries of column c in row block b (k_csr_count), turned into the
// first output slot of (b, c) by a scan down every column on top of the column offsets
// (k_csr_offsets); k_csr_fill then lets block b walk its rows IN ORDER, all threads on the
// entries of one row (their columns are distinct in a canonical CSR), each taking the next slot
// of its column — so the rows come out ascending inside every gene, as the layout builders and
// the reference's two-pointer merge (infoset.py:365-391) need, without any sort.
__global__ void __launch_bounds__(256)
k_csr_count(const int64_t *__restrict__ crow, const void *__restrict__ col, int col_bytes,
            int64_t N, int64_t P, int64_t rows_per_block, int32_t *__restrict__ T,
            int32_t *__restrict__ flags)
{
    const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
    if (r0 >= r1) return;
    int32_t *Tb = T + (int64_t)blockIdx.x * P;
    int bad = 0;
    for (int64_t i = crow[r0] + threadIdx.x; i < crow[r1]; i += blockDim.x) {
        const int64_t c = col_bytes == 4 ? (int64_t)((const int32_t *)col)[i] : ((const int64_t *)col)[i];
        if (c < 0 || c >= P)
            bad = BAD_ROW;
        else
            atomicAdd(Tb + c, 1);
    }
    if (bad) atomicOr(flags, bad);
}

// indptr int64[P + 1]: column offsets (exclusive scan of the column totals, by the caller)
__global__ void __launch_bounds__(256)
k_csr_offsets(int32_t *__restrict__ T, int B, int64_t P, const int64_t *__restrict__ indptr,
              int64_t *__restrict__ T64)
{
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= P) return;
    int64_t run = indptr[c];
    for (int b = 0; b < B; ++b) {
        const int32_t t = T[(int64_t)b * P + c];
        T64[(int64_t)b * P + c] = run;
        run += t;
    }
}

__global__ void __launch_bounds__(256)
k_csr_fill(const int64_t *__restrict__ crow, const void *__restrict__ col, int col_bytes,
           const void *__restrict__ vals, int val_bytes, int64_t N, int64_t P,
           int64_t rows_per_block, unsigned long long *__restrict__ T64,
           int32_t *__restrict__ rows_out, void *__restrict__ vals_out)
{
    const int64_t r0 = blockIdx.x * rows_per_block, r1 = min(N, r0 + rows_per_block);
    unsigned long long *Tb = T64 + (int64_t)blockIdx.x * P;
    for (int64_t r = r0; r < r1; ++r) {
        const int64_t e0 = crow[r], e1 = crow[r + 1];
        for (int64_t i = e0 + threadIdx.x; i < e1; i += blockDim.x) {
            const int64_t c = col_bytes == 4 ? (int64_t)((const int32_t *)col)[i] : ((const int64_t *)col)[i];
            const int64_t at = (int64_t)atomicAdd(Tb + c, 1ull);  // (columns of a row are distinct)
            rows_out[at] = (int32_t)r;
            if (val_bytes == 4)
                ((uint32_t *)vals_out)[at] = ((const uint32_t *)vals)[i];
            else
                ((uint64_t *)vals_out)[at] = ((const uint64_t *)vals)[i];
        }
        __syncthreads();  // row r has taken its slots before row r + 1 asks for the next ones
    }
}

}  // namespace prb

using namespace prb;

extern "C" int prb_csr_count(const int64_t *crow, const void *col, int col_bytes, int64_t N,
                             int64_t P, int B, int32_t *T, int32_t *flags, prb_stream_t stream)
{
    if (N == 0 || P == 0) return 0;
    PRB_REQUIRE(crow && col && T && flags && B >= 1 && (col_bytes == 4 || col_bytes == 8),
                "bad arguments");
    PRB_CUDA(cudaMemsetAsync(T, 0, sizeof(int32_t) * (size_t)B * P, S(stream)));
    k_csr_count<<<B, 256, 0, S(stream)>>>(crow, col, col_bytes, N, P, ceil_div(N, B), T, flags);
    PRB_LAUNCH_CHECK("k_csr_count");
    return 0;
}

extern "C" int prb_csr_fill(const int64_t *crow, const void *col, int col_bytes, const void *vals,
                            int val_bytes, int64_t N, int64_t P, int B, int32_t *T,
                            const int64_t *indptr, int64_t *T64, int32_t *rows_out, void *vals_out,
                            prb_stream_t stream)
{
    if (N == 0 || P == 0) return 0;
    PRB_REQUIRE(crow && col && vals && T && indptr && T64 && rows_out && vals_out && B >= 1,
                "bad arguments");
    PRB_REQUIRE((col_bytes == 4 || col_bytes == 8) && (val_bytes == 4 || val_bytes == 8),
                "column ids must be int32 / int64, values 4 or 8 bytes wide");
    k_csr_offsets<<<(unsigned)ceil_div(P, 256), 256, 0, S(stream)>>>(T, B, P, indptr, T64);
    PRB_LAUNCH_CHECK("k_csr_offsets");
    k_csr_fill<<<B, 256, 0, S(stream)>>>(crow, col, col_bytes, vals, val_bytes, N, P, ceil_div(N, B),
                                         (unsigned long long *)T64, rows_out, vals_out);
    PRB_LAUNCH_CHECK("k_csr_fill");
    return 0;
}

extern "C" int prb_pack_chunk(const void *rows, int row_bytes, const void *codes, int code_bytes,
                              int code_signed, int64_t n, int64_t N, uint32_t *packed,
                              int32_t *flags, prb_stream_t stream)
{
    if (n == 0) return 0;
    PRB_REQUIRE(row_bytes == 4 || row_bytes == 8, "row ids must be int32 or int64");
    PRB_REQUIRE(code_bytes == 1 || code_bytes == 2 || code_bytes == 4 || code_bytes == 8,
                "codes must be 1-, 2-, 4- or 8-byte integers");
    PRB_REQUIRE(N > 0 && N <= PRB_MAX_ROWS && rows && codes && packed && flags, "bad arguments");
    const int blocks = (int)hmin(ceil_div(n, 256), (int64_t)devinfo().sm_count * 16);
    k_pack_chunk<<<blocks, 256, 0, S(stream)>>>(rows, row_bytes, codes, code_bytes, code_signed, n,
                                                N, packed, flags);
    PRB_LAUNCH_CHECK("k_pack_chunk");
    return 0;
}

extern "C" int prb_check_columns(const uint32_t *packed, const int64_t *indptr, int64_t P,
                                 int32_t *flags, prb_stream_t stream)
{
    if (P == 0) return 0;
    PRB_REQUIRE(packed && indptr && flags, "null pointer");
    const int blocks = (int)hmin(ceil_div(P, 8), (int64_t)devinfo().sm_count * 16);
    k_check_columns<<<blocks, 256, 0, S(stream)>>>(packed, indptr, P, flags);
    PRB_LAUNCH_CHECK("k_check_columns");
    return 0;
}
