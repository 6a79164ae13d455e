// host.cu — host-side entry points: the fp64 term table and the synchronous
// HOST-buffer drop-in for the reference's njit seam
//   _sparse_entropy_wrt(dindices, dindptr, ddata, mindices, mdata, n_rows,
//                       n_feats, n_mcols, shift, ybits, y, classsizes)
// (picturedrocks/markers/mutualinformation/infoset.py:306-410).
#include <math.h>

#include <vector>

#include "common.cuh"

using namespace prb;

extern "C" int prb_term_table(int64_t n_rows, double *out_host)
{
    PRB_REQUIRE(n_rows > 0 && out_host != nullptr, "bad arguments");
    // infoset.py:392-396: counts = counts / n_rows ; h += -e * np.log2(e)
    // (glibc log2 == the libm call numba emits; compiled without FMA contraction)
    const double n = (double)n_rows;
    out_host[0] = 0.0;
    for (int64_t c = 1; c <= n_rows; ++c) {
        volatile double e = (double)c / n;
        volatile double l = log2(e);
        out_host[c] = -e * l;
    }
    return 0;
}

namespace {
struct DevBuf {
    void *p = nullptr;
    ~DevBuf()
    {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes)
    {
        PRB_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
        return 0;
    }
    template <typename T>
    T *as()
    {
        return (T *)p;
    }
};
}  // namespace

extern "C" int prb_sparse_entropy_wrt_host(const int32_t *dindices, const int64_t *dindptr,
                                           const int64_t *ddata, const int64_t *mindices,
                                           const int64_t *mdata, int64_t m_nnz, int64_t n_rows,
                                           int64_t n_feats, int64_t n_mcols, int64_t shift,
                                           int64_t ybits, const int64_t *y,
                                           const double *classsizes, int64_t n_classes,
                                           double *out)
{
    PRB_REQUIRE(n_rows > 0 && n_rows <= PRB_MAX_ROWS, "n_rows out of range");
    PRB_REQUIRE(n_feats > 0, "n_feats must be positive");
    PRB_REQUIRE(shift >= 1 && shift <= 8, "shift must be in [1, 8]");
    PRB_REQUIRE(n_mcols >= 0 && shift * n_mcols <= 30, "master column wider than 30 bits");
    const int K = 1 << shift, K1 = K - 1;
    const int C = ybits ? (int)n_classes : 1;
    PRB_REQUIRE(C >= 1 && (int64_t)C * K1 <= 65536, "too many classes");
    const int64_t N = n_rows, P = n_feats, nnz = dindptr[P];
    const int mbits = (int)(shift * n_mcols);
    cudaStream_t st = 0;

    // ---- host staging --------------------------------------------------------
    std::vector<uint32_t> h_packed((size_t)nnz);
    for (int64_t i = 0; i < nnz; ++i) {
        if (ddata[i] <= 0 || ddata[i] >= K) return fail(__func__, "code outside [1, 2^shift)");
        h_packed[(size_t)i] = ((uint32_t)ddata[i] << 24) | (uint32_t)dindices[i];
    }
    std::vector<int32_t> h_y((size_t)N, 0);
    std::vector<uint16_t> h_ystate((size_t)N, 0);
    std::vector<int64_t> h_cls((size_t)C, N);
    if (ybits) {
        for (int64_t i = 0; i < N; ++i) {
            if (y[i] < 0 || y[i] >= C) return fail(__func__, "label outside [0, n_classes)");
            h_y[(size_t)i] = (int32_t)y[i];
            h_ystate[(size_t)i] = (uint16_t)(y[i] * K1);
        }
        for (int c = 0; c < C; ++c) h_cls[(size_t)c] = (int64_t)classsizes[c];
    }
    std::vector<uint32_t> h_m((size_t)N, 0u);
    for (int64_t t = 0; t < m_nnz; ++t) h_m[(size_t)mindices[t]] = (uint32_t)mdata[t];
    std::vector<double> h_term((size_t)N + 1);
    if (prb_term_table(N, h_term.data())) return 1;

    // ---- device buffers --------------------------------------------------------
    const int64_t range_len = 4096;
    const int64_t n_ranges = ceil_div(nnz, range_len);
    const int64_t bm_words = (ceil_div(N, 32) + 3) & ~int64_t(3);
    DevBuf d_packed, d_indptr, d_rg, d_y, d_ystate, d_cls, d_m, d_term, d_cnt, d_ctr, d_out;
    DevBuf d_table, d_bitmap, d_state, d_hsize, d_hsb, d_hsy, d_hsys, d_nhs, d_hits;
    if (d_packed.alloc(4 * (size_t)nnz) || d_indptr.alloc(8 * (size_t)(P + 1)) ||
        d_rg.alloc(4 * (size_t)n_ranges) || d_y.alloc(4 * (size_t)N) ||
        d_ystate.alloc(2 * (size_t)N) || d_cls.alloc(8 * (size_t)C) ||
        d_m.alloc(4 * (size_t)N) || d_term.alloc(8 * (size_t)(N + 1)) ||
        d_cnt.alloc(4 * (size_t)P * C * K1) || d_ctr.alloc(16) || d_out.alloc(8 * (size_t)P))
        return 1;
    PRB_CUDA(cudaMemcpyAsync(d_packed.p, h_packed.data(), 4 * (size_t)nnz, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_indptr.p, dindptr, 8 * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_y.p, h_y.data(), 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_ystate.p, h_ystate.data(), 2 * (size_t)N, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_cls.p, h_cls.data(), 8 * (size_t)C, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_m.p, h_m.data(), 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_term.p, h_term.data(), 8 * (size_t)(N + 1), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4 * (size_t)P * C * K1, st));
    PRB_CUDA(cudaMemsetAsync(d_ctr.p, 0, 16, st));
    if (prb_range_genes(d_indptr.as<int64_t>(), P, nnz, range_len, n_ranges, d_rg.as<int32_t>(), st))
        return 1;
    // class-conditional code tables (every stored entry, state = y)
    if (prb_stream_hits(d_packed.as<uint32_t>(), d_indptr.as<int64_t>(), d_rg.as<int32_t>(),
                        n_ranges, range_len, nnz, nullptr, ybits ? d_ystate.as<uint16_t>() : nullptr,
                        N, (int64_t)C * K1, d_cnt.as<int32_t>(), d_ctr.as<int32_t>(), st))
        return 1;
    int32_t n_hs = 0;
    if (mbits > 0 && m_nnz > 0) {
        const int64_t n_keys = (int64_t)C << mbits;
        const int64_t hs_cap = N < n_keys ? N : n_keys;
        if (d_table.alloc(4 * (size_t)n_keys) || d_bitmap.alloc(4 * (size_t)bm_words) ||
            d_state.alloc(2 * (size_t)N) || d_hsize.alloc(4 * (size_t)hs_cap) ||
            d_hsb.alloc(4 * (size_t)(C + 1)) || d_hsy.alloc(4 * (size_t)hs_cap) ||
            d_hsys.alloc(4 * (size_t)C) || d_nhs.alloc(16))
            return 1;
        PRB_CUDA(cudaMemsetAsync(d_bitmap.p, 0, 4 * (size_t)bm_words, st));
        if (prb_state_table(d_m.as<uint32_t>(), ybits ? d_y.as<int32_t>() : nullptr, N, C, K, mbits,
                            d_table.as<int32_t>(), d_bitmap.as<uint32_t>(), d_state.as<uint16_t>(),
                            d_hsize.as<int32_t>(), d_hsb.as<int32_t>(), d_hsy.as<int32_t>(),
                            d_hsys.as<int32_t>(), d_nhs.as<int32_t>(), d_ctr.as<int32_t>(), hs_cap,
                            st))
            return 1;
        PRB_CUDA(cudaMemcpyAsync(&n_hs, d_nhs.p, 4, cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        PRB_REQUIRE((int64_t)n_hs * K1 <= 65536, "too many joint states for the uint16 state range");
        // state16 was written as rank*K1 truncated to 16 bits only if the check above holds
    }
    if (n_hs > 0) {
        if (d_hits.alloc(4 * (size_t)P * n_hs * K1)) return 1;
        PRB_CUDA(cudaMemsetAsync(d_hits.p, 0, 4 * (size_t)P * n_hs * K1, st));
        if (prb_stream_hits(d_packed.as<uint32_t>(), d_indptr.as<int64_t>(), d_rg.as<int32_t>(),
                            n_ranges, range_len, nnz, d_bitmap.as<uint32_t>(),
                            d_state.as<uint16_t>(), N, (int64_t)n_hs * K1, d_hits.as<int32_t>(),
                            d_ctr.as<int32_t>(), st))
            return 1;
    }
    if (prb_finalize(n_hs > 0 ? d_hits.as<int32_t>() : nullptr, d_cnt.as<int32_t>(),
                     d_cls.as<int64_t>(), d_hsb.as<int32_t>(), d_hsy.as<int32_t>(),
                     d_hsize.as<int32_t>(), d_hsys.as<int32_t>(), nullptr, nullptr, n_hs, C, K,
                     d_term.as<double>(), N, P, d_out.as<double>(), nullptr, st))
        return 1;
    PRB_CUDA(cudaMemcpyAsync(out, d_out.p, 8 * (size_t)P, cudaMemcpyDeviceToHost, st));
    PRB_CUDA(cudaStreamSynchronize(st));
    return 0;
}
