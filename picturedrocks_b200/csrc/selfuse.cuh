// selfuse.cuh — the greedy step's tail that the finalize kernels fuse in (prb_finalize_update).
#pragma once
#include "select.cuh"

namespace prb {

// SEL: the selector update of ITER:48-59 / 76-87 / 112-121 for the gene just finished (score
// algebra of select.cuh), the block's argmax candidate, the reduction of all candidates to ONE
// (score, index) record by the last block to finish (np.argmax of BASE:65 over this rank's
// genes), and the reset of the selected gene's entries of sx to "x_j = 0" (the streaming pass is
// over), so the next step's scatter starts from a clean vector.
struct SelFuse {
    int kind, is_remove;
    const double *base;
    double *penalty, *score;
    const uint8_t *selected;
    const double *Hx_all, *Hyx_all;
    const int32_t *gene_ptr, *n_sel_ptr;
    int32_t gene_lo;
    double *cand_score;   // [gridDim.x] scratch
    int64_t *cand_idx;    // [gridDim.x] scratch
    double *best;         // [2]: score, index (int64 bits)
    int *ticket;          // zero on entry and on exit
    // column of the selected gene (packed CSC holding every gene: gene ids are global)
    const uint32_t *col_packed;
    const int64_t *col_begin, *col_end;
    const int32_t *newpos;
    uint8_t *sx;
    // gene-sharded run: the record goes straight into the peers' buffers (peer.cu), W > 0
    unsigned long long *const *peer_bufs;
    int W, rank;
    int *xtag;
};

// the streaming pass has consumed sx: hand it back all "x_j = 0" (every thread of the grid)
__device__ __forceinline__ void sel_reset_sx(const SelFuse &sel)
{
    const int64_t j = *sel.gene_ptr;
    const int64_t cb = sel.col_begin[j], ce = sel.col_end[j];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = cb + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < ce; i += stride) {
        const uint32_t cell = __ldg(sel.col_packed + i) & ROW_MASK;
        sel.sx[sel.newpos ? (uint32_t)sel.newpos[cell] : cell] = (uint8_t)PRB_SX_NONE;
    }
}

// (sc, idx): this thread's candidate (idx < 0: none).  Block argmax, candidate of the block,
// and — in the last block to arrive — the rank's record, stored into the peers' buffers too
// when the run is gene-sharded (k_peer_share of peer.cu, fused in).  Every thread of the block
// must call it.
template <int NT>
__device__ __forceinline__ void sel_tail(const SelFuse &sel, double sc, int64_t idx)
{
    block_argmax<NT>(sc, idx);
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        sel.cand_score[blockIdx.x] = sc;
        sel.cand_idx[blockIdx.x] = idx;
        __threadfence();
        s_last = atomicAdd(sel.ticket, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    sc = 0.0;
    idx = -1;
    for (int c = threadIdx.x; c < (int)gridDim.x; c += NT) {
        const double s2 = __ldcg(sel.cand_score + c);
        const int64_t i2 = __ldcg((const long long *)sel.cand_idx + c);
        if (i2 >= 0 && (idx < 0 || better(s2, i2, sc, idx))) {
            sc = s2;
            idx = i2;
        }
    }
    block_argmax<NT>(sc, idx);
    __shared__ double s_sc;
    __shared__ int64_t s_idx;
    if (threadIdx.x == 0) {
        sel.best[0] = sc;
        ((int64_t *)sel.best)[1] = idx;
        *sel.ticket = 0;
        s_sc = sc;
        s_idx = idx;
    }
    if (sel.W > 0) {
        __syncthreads();
        const int tag = *sel.xtag + 1;
        if ((int)threadIdx.x < sel.W) {
            unsigned long long *slot =
                sel.peer_bufs[threadIdx.x] + ((size_t)(tag & 1) * sel.W + sel.rank) * 4;
            asm volatile("st.global.v2.b64 [%0], {%1, %2};" ::"l"(slot),
                         "l"((unsigned long long)__double_as_longlong(s_sc)),
                         "l"((unsigned long long)s_idx)
                         : "memory");
            __threadfence_system();
            asm volatile("st.release.sys.global.b64 [%0], %1;" ::"l"(slot + 2),
                         "l"((unsigned long long)tag)
                         : "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0) *sel.xtag = tag;
    }
}

}  // namespace prb
