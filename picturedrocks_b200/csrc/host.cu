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
    // count-table pass: every cell is a master row whose state is its class (+1)
    int64_t tile = 0, smem = 0;
    int n_tiles = 0, teams = 0, sb_y = 1;
    if (prb_stream_plan(N, C, K1, &tile, &n_tiles, &teams, &sb_y, &smem)) return 1;
    std::vector<int32_t> h_y((size_t)N, 0);
    std::vector<uint8_t> h_ystate((size_t)N * sb_y + 64, 0);
    std::vector<int64_t> h_cls((size_t)C, N);
    for (int64_t i = 0; i < N; ++i) {
        int64_t yy = 0;
        if (ybits) {
            yy = y[i];
            if (yy < 0 || yy >= C) return fail(__func__, "label outside [0, n_classes)");
            h_y[(size_t)i] = (int32_t)yy;
        }
        if (sb_y == 1)
            h_ystate[(size_t)i] = (uint8_t)(yy + 1);
        else
            ((uint16_t *)h_ystate.data())[i] = (uint16_t)(yy + 1);
    }
    if (ybits)
        for (int c = 0; c < C; ++c) h_cls[(size_t)c] = (int64_t)classsizes[c];
    std::vector<uint32_t> h_m((size_t)N, 0u);
    for (int64_t t = 0; t < m_nnz; ++t) h_m[(size_t)mindices[t]] = (uint32_t)mdata[t];
    std::vector<double> h_term((size_t)N + 1);
    if (prb_term_table(N, h_term.data())) return 1;

    // ---- device buffers --------------------------------------------------------
    DevBuf d_packed, d_indptr, d_tp, d_y, d_state, d_cls, d_m, d_term, d_cnt, d_ctr, d_out;
    DevBuf d_table, d_hsize, d_hsb, d_hsy, d_hsys, d_nhs, d_hits;
    if (d_packed.alloc(4 * (size_t)nnz) || d_indptr.alloc(8 * (size_t)(P + 1)) ||
        d_tp.alloc(8 * (size_t)P * (n_tiles + 1)) || d_y.alloc(4 * (size_t)N) ||
        d_state.alloc(2 * (size_t)N + 64) || d_cls.alloc(8 * (size_t)C) ||
        d_m.alloc(4 * (size_t)N) || d_term.alloc(8 * (size_t)(N + 1)) ||
        d_cnt.alloc(4 * (size_t)P * C * K1) || d_ctr.alloc(4 * 8192) || d_out.alloc(8 * (size_t)P) ||
        d_hsb.alloc(4 * (size_t)(C + 1)) || d_hsys.alloc(4 * (size_t)C) || d_nhs.alloc(16))
        return 1;
    PRB_CUDA(cudaMemcpyAsync(d_packed.p, h_packed.data(), 4 * (size_t)nnz, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_indptr.p, dindptr, 8 * (size_t)(P + 1), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_y.p, h_y.data(), 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_state.p, h_ystate.data(), (size_t)N * sb_y, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_cls.p, h_cls.data(), 8 * (size_t)C, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_m.p, h_m.data(), 4 * (size_t)N, cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemcpyAsync(d_term.p, h_term.data(), 8 * (size_t)(N + 1), cudaMemcpyHostToDevice, st));
    PRB_CUDA(cudaMemsetAsync(d_cnt.p, 0, 4 * (size_t)P * C * K1, st));
    PRB_REQUIRE(n_tiles <= 8192, "too many cell tiles");
    if (prb_tile_ptrs(d_packed.as<uint32_t>(), d_indptr.as<int64_t>(), P, tile, n_tiles,
                      d_tp.as<int64_t>(), st))
        return 1;
    if (prb_stream_hits(d_packed.as<uint32_t>(), d_tp.as<int64_t>(), P, nnz, d_state.p, N, C, K1,
                        0, d_cnt.as<int32_t>(), d_ctr.as<int32_t>(), st))
        return 1;
    int32_t n_hs = 0;
    if (mbits > 0 && m_nnz > 0) {
        const int64_t n_keys = (int64_t)C << mbits;
        const int64_t hs_cap = N < n_keys ? N : n_keys;
        if (d_table.alloc(4 * (size_t)n_keys) || d_hsize.alloc(4 * (size_t)hs_cap) ||
            d_hsy.alloc(4 * (size_t)hs_cap))
            return 1;
        if (prb_state_table(d_m.as<uint32_t>(), ybits ? d_y.as<int32_t>() : nullptr, N, C, mbits,
                            d_table.as<int32_t>(), d_hsize.as<int32_t>(), d_hsb.as<int32_t>(),
                            d_hsy.as<int32_t>(), d_hsys.as<int32_t>(), d_nhs.as<int32_t>(), hs_cap,
                            st))
            return 1;
        PRB_CUDA(cudaMemcpyAsync(&n_hs, d_nhs.p, 4, cudaMemcpyDeviceToHost, st));
        PRB_CUDA(cudaStreamSynchronize(st));
        PRB_REQUIRE((int64_t)n_hs * K1 <= 65536, "too many joint states (65536 histogram bins)");
    }
    if (n_hs > 0) {
        int sb = 1, nt2 = 0;
        int64_t tile2 = 0;
        if (prb_stream_plan(N, n_hs, K1, &tile2, &nt2, &teams, &sb, &smem)) return 1;
        PRB_REQUIRE(nt2 <= 8192, "too many cell tiles");
        DevBuf d_tp2;
        if (d_tp2.alloc(8 * (size_t)P * (nt2 + 1)) || d_hits.alloc(4 * (size_t)P * n_hs * K1))
            return 1;
        PRB_CUDA(cudaMemsetAsync(d_hits.p, 0, 4 * (size_t)P * n_hs * K1, st));
        if (prb_tile_ptrs(d_packed.as<uint32_t>(), d_indptr.as<int64_t>(), P, tile2, nt2,
                          d_tp2.as<int64_t>(), st) ||
            prb_state_assign(d_m.as<uint32_t>(), ybits ? d_y.as<int32_t>() : nullptr, N, mbits,
                             d_table.as<int32_t>(), d_state.p, sb, st) ||
            prb_stream_hits(d_packed.as<uint32_t>(), d_tp2.as<int64_t>(), P, nnz, d_state.p, N,
                            n_hs, K1, 0, d_hits.as<int32_t>(), d_ctr.as<int32_t>(), st))
            return 1;
        PRB_CUDA(cudaStreamSynchronize(st));  // d_tp2 is freed on scope exit
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
