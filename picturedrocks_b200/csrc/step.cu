// step.cu — the head of a greedy step on the V-layout path, one launch.
//
// Replaces, for K <= 5, the host-side sequence of the reference's autoselect()/add()
// (picturedrocks/markers/_iterative.py:64-66: np.argmax(score) -> add(best);
// mutualinformation/iterative.py:48-49: S.append(ind)) and the construction of the master
// column of the selected gene (infoset.py:289 -> 431-442) with ONE kernel:
//   1. every block reduces the n_rec (score, index) candidate records — one per rank after
//      the all-gather, or this rank's own record — with np.argmax's rule (max score, NaN
//      maximal, lowest index on ties), so all blocks and all ranks agree on the winner j
//      without another launch;
//   2. block 0 commits it: gene_ptr <- j, S[n_sel++] <- j, selected[j - gene_lo] <- 1, and
//      copies the sizes of the (class, x_j) states from the count table of gene j
//      (hsize[y][a] = cnt[j][y][a]: the cells of class y with x_j = a+1), their per-class and
//      per-value sums; block 1 zeroes the streaming kernel's work counters;
//   3. all blocks scatter gene j's column into the per-cell shift-code vector sx
//      (vstream.cu: 8*(x_j-1) at newpos[cell]); sx is all "x_j = 0" on entry — the step's
//      tail (prb_finalize_update / prb_v_unscatter) resets the entries again.
// The column is read from a packed CSC that holds EVERY gene (multi-GPU: the replicated
// copy, see engine.py), so no rank has to send x_j to the others.
#include "select.cuh"

namespace prb {

constexpr int PICK_THREADS = 256;

__global__ void __launch_bounds__(PICK_THREADS)
k_pick_scatter(const double *__restrict__ rec, int n_rec, int32_t *gene_ptr, int32_t *S_out,
               int32_t *n_sel_ptr, uint8_t *selected, int32_t gene_lo, int64_t P_local,
               const uint32_t *__restrict__ col_packed, const int64_t *__restrict__ col_begin,
               const int64_t *__restrict__ col_end, const int32_t *__restrict__ newpos,
               uint8_t *__restrict__ sx, const int32_t *__restrict__ cnt_all, int C, int K1,
               int32_t *hsize, int32_t *hs_begin, int32_t *hs_y, int32_t *hs_ysum, int32_t *ha,
               int *counters, int n_counters, const unsigned long long *peer_local, int n_peers,
               const int *xtag, int *peer_err)
{
    __shared__ int s_j;
    __shared__ double s_rec[2 * PICK_THREADS / 4];
    if (peer_local) {
        // the records arrive in this rank's peer buffer (peer.cu): wait for the tag of the
        // current exchange in every rank's slot, then take a private copy.  A step that does
        // not pick (add / remove: n_rec = 0) waits all the same: every rank then runs at most
        // one exchange ahead of the slowest, which is what the two halves of the buffer allow.
        const int tag = *xtag;
        const unsigned long long *half = peer_local + (size_t)(tag & 1) * n_peers * 4;
        if (threadIdx.x < n_peers) {
            const unsigned long long *slot = half + threadIdx.x * 4;
            unsigned long long t = 0, t0 = 0;
            for (unsigned spin = 0;; ++spin) {
                asm volatile("ld.acquire.sys.global.b64 %0, [%1];" : "=l"(t) : "l"(slot + 2) : "memory");
                if (t == (unsigned long long)tag) break;
                if ((spin & 1023u) == 1023u) {  // a peer that never answers must not hang the GPU
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                    if (t0 == 0) t0 = now;
                    if (now - t0 > 4000000000ull) {
                        atomicExch(peer_err, 1);
                        break;
                    }
                }
            }
            s_rec[2 * threadIdx.x] = __longlong_as_double((long long)slot[0]);
            ((unsigned long long *)s_rec)[2 * threadIdx.x + 1] = slot[1];
        }
        __syncthreads();
        if (n_rec > 0) rec = s_rec;
    }
    if (threadIdx.x == 0) {
        int64_t j;
        if (n_rec > 0) {
            double sc = rec[0];
            j = ((const int64_t *)rec)[1];
            for (int r = 1; r < n_rec; ++r) {
                const double s2 = rec[2 * r];
                const int64_t i2 = ((const int64_t *)rec)[2 * r + 1];
                if (i2 >= 0 && (j < 0 || better(s2, i2, sc, j))) {
                    sc = s2;
                    j = i2;
                }
            }
        } else {
            j = *gene_ptr;  // add(ind) / remove(ind) / entropy_wrt([j]): the host chose the gene
        }
        s_j = (int)j;
    }
    __syncthreads();
    const int64_t j = s_j;
    if (j < 0) return;  // no candidate at all (empty matrix)
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0 && n_rec > 0) {
            *gene_ptr = (int32_t)j;
            const int n = *n_sel_ptr;
            S_out[n] = (int32_t)j;
            *n_sel_ptr = n + 1;
            const int64_t l = j - gene_lo;
            if (l >= 0 && l < P_local) selected[l] = 1;
        }
        const int n_hs = C * K1;
        const int32_t *cj = cnt_all + j * n_hs;
        for (int i = threadIdx.x; i < n_hs; i += PICK_THREADS) {
            hsize[i] = cj[i];
            hs_y[i] = i / K1;  // the generic finalize kernel's view of the direct layout
        }
        for (int yy = threadIdx.x; yy <= C; yy += PICK_THREADS) hs_begin[yy] = yy * K1;
        for (int yy = threadIdx.x; yy < C; yy += PICK_THREADS) {
            int s = 0;
            for (int a = 0; a < K1; ++a) s += cj[yy * K1 + a];
            hs_ysum[yy] = s;
        }
        if (threadIdx.x < K1) {
            int s = 0;
            for (int yy = 0; yy < C; ++yy) s += cj[yy * K1 + threadIdx.x];
            ha[threadIdx.x] = s;
        }
    }
    if (blockIdx.x == (gridDim.x > 1 ? 1 : 0))
        for (int i = threadIdx.x; i < n_counters; i += PICK_THREADS) counters[i] = 0;
    const int64_t b = col_begin[j], e = col_end[j];
    const int64_t stride = (int64_t)gridDim.x * PICK_THREADS;
    for (int64_t i = b + blockIdx.x * (int64_t)PICK_THREADS + threadIdx.x; i < e; i += stride) {
        const uint32_t w = __ldg(col_packed + i);
        const uint32_t cell = w & ROW_MASK;
        sx[newpos ? (uint32_t)newpos[cell] : cell] = (uint8_t)(8u * ((w >> 24) - 1u));
    }
}

// sx[newpos[cell]] <- "x_j = 0" for the stored entries of gene *gene_ptr (the API path of
// entropy_wrt([j]); the selectors' step does this inside prb_finalize_update)
__global__ void __launch_bounds__(PICK_THREADS)
k_unscatter(const int32_t *__restrict__ gene_ptr, const uint32_t *__restrict__ col_packed,
            const int64_t *__restrict__ col_begin, const int64_t *__restrict__ col_end,
            const int32_t *__restrict__ newpos, uint8_t *__restrict__ sx)
{
    const int64_t j = *gene_ptr;
    const int64_t b = col_begin[j], e = col_end[j];
    const int64_t stride = (int64_t)gridDim.x * PICK_THREADS;
    for (int64_t i = b + blockIdx.x * (int64_t)PICK_THREADS + threadIdx.x; i < e; i += stride) {
        const uint32_t cell = __ldg(col_packed + i) & ROW_MASK;
        sx[newpos ? (uint32_t)newpos[cell] : cell] = (uint8_t)PRB_SX_NONE;
    }
}

// reduce the per-block candidates of prb_argmax_blocks to this rank's (score, index) record
__global__ void __launch_bounds__(SEL_BLOCK)
k_best_record(const double *__restrict__ cand_score, const int64_t *__restrict__ cand_idx,
              int64_t n_cand, double *__restrict__ best)
{
    double sc = 0.0;
    int64_t idx = -1;
    for (int64_t c = threadIdx.x; c < n_cand; c += SEL_BLOCK) {
        const double s2 = cand_score[c];
        const int64_t i2 = cand_idx[c];
        if (i2 >= 0 && (idx < 0 || better(s2, i2, sc, idx))) {
            sc = s2;
            idx = i2;
        }
    }
    block_argmax<SEL_BLOCK>(sc, idx);
    if (threadIdx.x == 0) {
        best[0] = sc;
        ((int64_t *)best)[1] = idx;
    }
}

}  // namespace prb

using namespace prb;

extern "C" int prb_pick_scatter(const double *records, int n_rec, int32_t *gene_ptr, int32_t *S_out,
                                int32_t *n_sel_ptr, uint8_t *selected, int32_t gene_lo,
                                int64_t P_local, const uint32_t *col_packed,
                                const int64_t *col_begin, const int64_t *col_end,
                                const int32_t *newpos, uint8_t *sx, const int32_t *cnt_all, int C,
                                int K, int32_t *hsize, int32_t *hs_begin, int32_t *hs_y,
                                int32_t *hs_ysum, int32_t *ha, int32_t *counters, int n_counters,
                                const uint64_t *peer_local, int n_peers, const int32_t *xtag,
                                int32_t *peer_err, prb_stream_t stream)
{
    PRB_REQUIRE(n_rec >= 0 && (n_rec == 0 || records != nullptr || peer_local != nullptr),
                "bad candidate records");
    PRB_REQUIRE(peer_local == nullptr ||
                    (xtag && peer_err && n_peers >= 1 && n_peers <= PICK_THREADS / 4 &&
                     (n_rec == 0 || n_rec == n_peers)),
                "bad peer buffer arguments");
    PRB_REQUIRE(C >= 1 && K >= 2 && K <= 21, "the V-layout path needs 2 <= K <= 21");
    PRB_REQUIRE(gene_ptr && col_packed && col_begin && col_end && sx && cnt_all && hsize &&
                    hs_begin && hs_y && hs_ysum && ha,
                "null pointer");
    PRB_REQUIRE(n_rec == 0 || (S_out && n_sel_ptr && selected), "commit targets are required");
    k_pick_scatter<<<devinfo().sm_count * 4, PICK_THREADS, 0, S(stream)>>>(
        records, n_rec, gene_ptr, S_out, n_sel_ptr, selected, gene_lo, P_local, col_packed,
        col_begin, col_end, newpos, sx, cnt_all, C, K - 1, hsize, hs_begin, hs_y, hs_ysum, ha,
        counters, counters ? n_counters : 0, (const unsigned long long *)peer_local, n_peers, xtag,
        peer_err);
    PRB_LAUNCH_CHECK("k_pick_scatter");
    return 0;
}

extern "C" int prb_v_unscatter(const int32_t *gene_ptr, const uint32_t *col_packed,
                               const int64_t *col_begin, const int64_t *col_end,
                               const int32_t *newpos, uint8_t *sx, prb_stream_t stream)
{
    PRB_REQUIRE(gene_ptr && col_packed && col_begin && col_end && sx, "null pointer");
    k_unscatter<<<devinfo().sm_count * 4, PICK_THREADS, 0, S(stream)>>>(gene_ptr, col_packed,
                                                                        col_begin, col_end, newpos,
                                                                        sx);
    PRB_LAUNCH_CHECK("k_unscatter");
    return 0;
}

extern "C" int prb_best_record(const double *cand_score, const int64_t *cand_idx, int64_t n_cand,
                               double *best, prb_stream_t stream)
{
    PRB_REQUIRE(cand_score && cand_idx && best && n_cand >= 0, "bad arguments");
    k_best_record<<<1, SEL_BLOCK, 0, S(stream)>>>(cand_score, cand_idx, n_cand, best);
    PRB_LAUNCH_CHECK("k_best_record");
    return 0;
}
