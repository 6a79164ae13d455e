// selector.cu — MIM/JMI/CIFE score algebra and the greedy argmax, on device.
//
// Replaces the numpy expressions of
//   picturedrocks/markers/mutualinformation/iterative.py:24-35 (base score),
//   :47-72 (CIFE), :75-100 (JMI), :111-134 (CIFEUnsup)
// and np.argmax in picturedrocks/markers/_iterative.py:64-66.
// All arithmetic is fp64 with explicit round-to-nearest adds in the reference's
// left-to-right order, so scores are bit-identical given bit-identical entropies.
#include "select.cuh"

namespace prb {

__global__ void k_mi_base(const double *__restrict__ Hx, double Hy, const double *__restrict__ Hyx,
                          int64_t P, double *__restrict__ base)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < P) base[i] = __dsub_rn(__dadd_rn(Hx[i], Hy), Hyx[i]);
}

__global__ void __launch_bounds__(SEL_BLOCK)
k_selector_update(int kind, int is_remove, const double *__restrict__ base,
                  double *__restrict__ penalty, double *__restrict__ score,
                  const uint8_t *__restrict__ selected, const double *__restrict__ Hij,
                  const double *__restrict__ Hyij, const double *__restrict__ Hx_all,
                  const double *__restrict__ Hyx_all, const int32_t *__restrict__ gene_ptr,
                  const int32_t *__restrict__ n_sel_ptr, int64_t P, int32_t gene_lo,
                  double *__restrict__ cand_score, int64_t *__restrict__ cand_idx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int j = *gene_ptr;
    const double Hj = Hx_all[j];
    double sc = 0.0;
    int64_t idx = -1;
    if (i < P) {
        const double Hyj = kind == PRB_SEL_CIFE_UNSUP ? 0.0 : Hyx_all[j];
        double pen = penalty[i];
        sc = sel_update(kind, is_remove, base[i], Hj, Hij[i], Hyj,
                        kind == PRB_SEL_CIFE_UNSUP ? 0.0 : Hyij[i], &pen, score[i], *n_sel_ptr,
                        selected[i] != 0);
        penalty[i] = pen;
        score[i] = sc;
        idx = gene_lo + i;
    }
    block_argmax<SEL_BLOCK>(sc, idx);
    if (threadIdx.x == 0) {
        cand_score[blockIdx.x] = sc;
        cand_idx[blockIdx.x] = idx;
    }
}

__global__ void __launch_bounds__(SEL_BLOCK)
k_argmax_blocks(const double *__restrict__ score, int64_t P, int32_t gene_lo,
                double *__restrict__ cand_score, int64_t *__restrict__ cand_idx)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    double sc = i < P ? score[i] : 0.0;
    int64_t idx = i < P ? gene_lo + i : -1;
    block_argmax<SEL_BLOCK>(sc, idx);
    if (threadIdx.x == 0) {
        cand_score[blockIdx.x] = sc;
        cand_idx[blockIdx.x] = idx;
    }
}

__global__ void __launch_bounds__(SEL_BLOCK)
k_argmax_final(const double *__restrict__ cand_score, const int64_t *__restrict__ cand_idx,
               int64_t n_cand, int64_t stride, double *__restrict__ best_score,
               int64_t *__restrict__ best_idx)
{
    double sc = 0.0;
    int64_t idx = -1;
    for (int64_t c = threadIdx.x; c < n_cand; c += SEL_BLOCK) {
        const double s2 = cand_score[c * stride];
        const int64_t i2 = cand_idx[c * stride];
        if (i2 >= 0 && (idx < 0 || better(s2, i2, sc, idx))) {
            sc = s2;
            idx = i2;
        }
    }
    block_argmax<SEL_BLOCK>(sc, idx);
    if (threadIdx.x == 0) {
        *best_score = sc;
        *best_idx = idx;
    }
}

// k_argmax_final followed by k_commit_add in one launch (the greedy step's first kernel).
__global__ void __launch_bounds__(SEL_BLOCK)
k_argmax_commit(const double *__restrict__ cand_score, const int64_t *__restrict__ cand_idx,
                int64_t n_cand, int64_t stride, double *__restrict__ best_score,
                int64_t *__restrict__ best_idx, int32_t *gene_ptr, int32_t *S, int32_t *n_sel_ptr,
                uint8_t *selected, int32_t gene_lo, int64_t P)
{
    double sc = 0.0;
    int64_t idx = -1;
    for (int64_t c = threadIdx.x; c < n_cand; c += SEL_BLOCK) {
        const double s2 = cand_score[c * stride];
        const int64_t i2 = cand_idx[c * stride];
        if (i2 >= 0 && (idx < 0 || better(s2, i2, sc, idx))) {
            sc = s2;
            idx = i2;
        }
    }
    block_argmax<SEL_BLOCK>(sc, idx);
    if (threadIdx.x == 0) {
        *best_score = sc;
        *best_idx = idx;
        *gene_ptr = (int32_t)idx;
        const int n = *n_sel_ptr;
        S[n] = (int32_t)idx;
        *n_sel_ptr = n + 1;
        const int64_t l = idx - gene_lo;
        if (l >= 0 && l < P) selected[l] = 1;
    }
}

__global__ void k_commit_add(const int64_t *__restrict__ best_idx, int32_t *gene_ptr, int32_t *S,
                             int32_t *n_sel_ptr, uint8_t *selected, int32_t gene_lo, int64_t P)
{
    const int64_t idx = *best_idx;
    *gene_ptr = (int32_t)idx;
    const int n = *n_sel_ptr;
    S[n] = (int32_t)idx;
    *n_sel_ptr = n + 1;
    const int64_t l = idx - gene_lo;
    if (l >= 0 && l < P) selected[l] = 1;
}

}  // namespace prb

using namespace prb;

extern "C" int prb_mi_base(const double *Hx, double Hy, const double *Hyx, int64_t P, double *base,
                           prb_stream_t stream)
{
    if (P == 0) return 0;
    k_mi_base<<<(unsigned)ceil_div(P, 256), 256, 0, S(stream)>>>(Hx, Hy, Hyx, P, base);
    PRB_LAUNCH_CHECK("k_mi_base");
    return 0;
}

extern "C" int prb_selector_update(int kind, int is_remove, const double *base, double *penalty,
                                   double *score, const uint8_t *selected, const double *Hij,
                                   const double *Hyij, const double *Hx_all,
                                   const double *Hyx_all, const int32_t *gene_ptr,
                                   const int32_t *n_sel_ptr, int64_t P, int32_t gene_lo,
                                   double *cand_score, int64_t *cand_idx, prb_stream_t stream)
{
    PRB_REQUIRE(kind == PRB_SEL_CIFE || kind == PRB_SEL_JMI || kind == PRB_SEL_CIFE_UNSUP,
                "unknown selector kind");
    PRB_REQUIRE(kind == PRB_SEL_CIFE_UNSUP || (Hyij && Hyx_all), "supervised update needs Hyij/Hyx");
    if (P == 0) return 0;
    k_selector_update<<<(unsigned)ceil_div(P, SEL_BLOCK), SEL_BLOCK, 0, S(stream)>>>(
        kind, is_remove, base, penalty, score, selected, Hij, Hyij, Hx_all, Hyx_all, gene_ptr,
        n_sel_ptr, P, gene_lo, cand_score, cand_idx);
    PRB_LAUNCH_CHECK("k_selector_update");
    return 0;
}

extern "C" int prb_argmax_blocks(const double *score, int64_t P, int32_t gene_lo,
                                 double *cand_score, int64_t *cand_idx, prb_stream_t stream)
{
    if (P == 0) return 0;
    k_argmax_blocks<<<(unsigned)ceil_div(P, SEL_BLOCK), SEL_BLOCK, 0, S(stream)>>>(
        score, P, gene_lo, cand_score, cand_idx);
    PRB_LAUNCH_CHECK("k_argmax_blocks");
    return 0;
}

extern "C" int prb_argmax_final(const double *cand_score, const int64_t *cand_idx, int64_t n_cand,
                                int64_t stride, double *best_score, int64_t *best_idx,
                                prb_stream_t stream)
{
    PRB_REQUIRE(stride >= 1, "stride must be >= 1");
    k_argmax_final<<<1, SEL_BLOCK, 0, S(stream)>>>(cand_score, cand_idx, n_cand, stride,
                                                   best_score, best_idx);
    PRB_LAUNCH_CHECK("k_argmax_final");
    return 0;
}

extern "C" int prb_argmax_commit(const double *cand_score, const int64_t *cand_idx, int64_t n_cand,
                                 int64_t stride, double *best_score, int64_t *best_idx,
                                 int32_t *gene_ptr, int32_t *S_out, int32_t *n_sel_ptr,
                                 uint8_t *selected, int32_t gene_lo, int64_t P,
                                 prb_stream_t stream)
{
    PRB_REQUIRE(stride >= 1, "stride must be >= 1");
    k_argmax_commit<<<1, SEL_BLOCK, 0, prb::S(stream)>>>(cand_score, cand_idx, n_cand, stride,
                                                         best_score, best_idx, gene_ptr, S_out,
                                                         n_sel_ptr, selected, gene_lo, P);
    PRB_LAUNCH_CHECK("k_argmax_commit");
    return 0;
}

extern "C" int prb_commit_add(const int64_t *best_idx, int32_t *gene_ptr, int32_t *S_out,
                              int32_t *n_sel_ptr, uint8_t *selected, int32_t gene_lo, int64_t P,
                              prb_stream_t stream)
{
    k_commit_add<<<1, 1, 0, prb::S(stream)>>>(best_idx, gene_ptr, S_out, n_sel_ptr, selected,
                                              gene_lo, P);
    PRB_LAUNCH_CHECK("k_commit_add");
    return 0;
}
