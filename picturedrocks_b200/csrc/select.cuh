// select.cuh — device helpers shared by selector.cu and the fused step kernels (step.cu,
// finalize.cu): the score algebra of one add()/remove() and the block-wide np.argmax.
#pragma once
#include <math.h>

#include "common.cuh"

namespace prb {

constexpr int SEL_BLOCK = 256;

// One add()/remove() update of gene i for the selected gene j, in the reference's
// left-to-right order (picturedrocks/markers/mutualinformation/iterative.py):
//   delta = (((base + Hj) - Hij) - Hyj) + Hyij      CIFE :50-56, JMI :78-84
//   delta = (base + Hj) - Hij                        CIFEUnsup :114-118
//   add:    penalty += delta; CIFE/CIFEUnsup: score -= delta (:57-58, :119-120)
//   remove: penalty -= delta; CIFE/CIFEUnsup: score = base - penalty (:70-71, :131-132)
//   JMI:    score = base - penalty / len(S)          (:86, :99)
//   score[S] = -inf                                  (:59, :72, :87, :100)
// Returns the new score; *penalty is updated in place.
__device__ __forceinline__ double sel_update(int kind, int is_remove, double b, double Hj,
                                             double Hij, double Hyj, double Hyij, double *penalty,
                                             double score_old, int n_sel, bool is_selected)
{
    double d;
    if (kind == PRB_SEL_CIFE_UNSUP)
        d = __dsub_rn(__dadd_rn(b, Hj), Hij);
    else
        d = __dadd_rn(__dsub_rn(__dsub_rn(__dadd_rn(b, Hj), Hij), Hyj), Hyij);
    const double pen = is_remove ? __dsub_rn(*penalty, d) : __dadd_rn(*penalty, d);
    *penalty = pen;
    double sc;
    if (kind == PRB_SEL_JMI)
        sc = __dsub_rn(b, __ddiv_rn(pen, (double)n_sel));
    else if (is_remove)
        sc = __dsub_rn(b, pen);
    else
        sc = __dsub_rn(score_old, d);
    return is_selected ? -INFINITY : sc;
}

// block-wide argmax with np.argmax semantics (first maximum; NaN maximal); idx < 0 = no
// candidate.  Every thread of the block must call it; the result is valid in thread 0.
// NT = threads per block (a multiple of 32, <= 1024).
template <int NT>
__device__ __forceinline__ void block_argmax(double &s, int64_t &idx)
{
    __shared__ double ss[NT / 32];
    __shared__ int64_t si[NT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        const double s2 = __shfl_down_sync(0xffffffffu, s, o);
        const int64_t i2 = __shfl_down_sync(0xffffffffu, idx, o);
        if (i2 >= 0 && (idx < 0 || better(s2, i2, s, idx))) {
            s = s2;
            idx = i2;
        }
    }
    if (lane == 0) {
        ss[warp] = s;
        si[warp] = idx;
    }
    __syncthreads();
    if (warp == 0) {
        s = lane < NT / 32 ? ss[lane] : 0.0;
        idx = lane < NT / 32 ? si[lane] : -1;
        for (int o = 16; o > 0; o >>= 1) {
            const double s2 = __shfl_down_sync(0xffffffffu, s, o);
            const int64_t i2 = __shfl_down_sync(0xffffffffu, idx, o);
            if (i2 >= 0 && (idx < 0 || better(s2, i2, s, idx))) {
                s = s2;
                idx = i2;
            }
        }
    }
    __syncthreads();  // ss / si may be reused by a second call
}

}  // namespace prb
