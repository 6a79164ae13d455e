// finalize_wide.cu — the tail of a greedy step for 5 < K <= 21 on the v-major hit table of the
// streamed lane-run kernel (vell2.cu), bit-compatible with the reference, in two kernels.
//
// Same job as k_finalize_direct_grp<.., SEL, VM> (finalize.cu): the tail of entropy_wrt_i in
// _sparse_entropy_wrt (picturedrocks/markers/mutualinformation/infoset.py:392-397; terms from
// the host-built fp64 table, added one at a time in ascending packed-bin order (y, x_j, x_i)),
// the selector update (iterative.py:48-59, 76-87) and the rank's argmax record.  With many bins
// the joint histogram of a gene is large and almost empty (20 bins, 20 classes: 8 000 bins, of
// which the 20 * 39 marginal ones and ~250 of the 7 220 hit bins are non-zero), and the
// 1 000-odd dependent DADDs of the ordered sum are the only inherently serial part.  The generic
// warp-per-gene kernel (k_finalize) walks all bins through ballot / shuffle rounds: 3.7 ms of a
// 4.3 ms step at config 5 / 20 bins.  Here:
//
//   k_wide_terms  (warp per gene, everything parallel over lanes, loads coalesced)
//     pass 1  the gene's block of the table, 8-byte loads: for every non-zero count of
//             (class y, x_j = a+1, x_i = v+1)  bm[y][a] |= 1 << v,  rs[y][a] += c,
//             cs[y][v] += c,  m2[a][v] += c  (shared-memory atomics, ~250 per gene)
//     scan    E[y][a] = number of non-zero hit bins before row (y, a) in bin order
//     terms   every bin knows its POSITION in the ordered sum:
//               (y,0,0)           y*(1+2*K1) + E[y][0]
//               (y,0,v)           ... + 1 + v
//               (y,a,0)           y*(1+2*K1) + 1 + K1 + a + E[y][a]
//               (y,a,v) non-zero  ... + 1 + popc(bm[y][a] & ((1 << v) - 1))
//             so lanes gather the terms of the marginal bins (derived by subtraction, as in
//             finalize.cu) and, in a second pass over the table (L1 / L2 hits), of the hit bins,
//             and store them at their positions of terms[g][.]; zero bins store +0.0, which adds
//             exactly; the table is zeroed where it was not.  The K*K bins of H(x_j, x_i)
//             follow from m2 and go to terms2[g][.].
//   k_wide_chain  (thread per gene; a warp stages 32 genes x 32 terms through shared memory so
//     that the loads are coalesced) adds the terms in order, applies the selector update and
//     runs the fused tail (selfuse.cuh).  32 genes share a warp, so a gene's chain costs ~30
//     warp instructions instead of the thousands of a warp-per-gene ballot loop.
//
// Needs every class in ONE segment (slot = class); the engine falls back to the generic path
// otherwise.
#include <stdlib.h>

#include "selfuse.cuh"

namespace prb {

constexpr int WT_WARPS = 8, WT_THREADS = WT_WARPS * 32;  // k_wide_terms
constexpr int WT_LIST = 384;  // word pairs per compaction round of k_wide_terms
constexpr int WT_U = 8;       // table loads a lane keeps in flight
constexpr int WC_WARPS = 4, WC_THREADS = WC_WARPS * 32;  // k_wide_chain (34 KB of staging tiles)

__device__ __forceinline__ int wt_scan_excl(int v, int lane, int &total)
{
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
}

// hits: uint32 words of uint16 hits[P][C][K1][4 * NQ] (slot = class).  Shared memory per warp
// (ints): bm[CK] E[CK] rs[CK] cs[CK] z0[C] m2[K1*K1] xc[K1] cr[CK] (pad), CK = C * K1, then the
// compaction list lp[WT_LIST] lw uint2[WT_LIST].
// terms: double[P][stride], stride >= C*(1+2*K1) + C*K1*K1; terms2: double[P][K*K] (want_marg);
// n_terms: int32[P].
__global__ void __launch_bounds__(WT_THREADS)
k_wide_terms(uint32_t *__restrict__ hits, const int32_t *__restrict__ cnt,
             const int64_t *__restrict__ clsz, const int32_t *__restrict__ hsize,
             const int32_t *__restrict__ hs_ysum, const int32_t *__restrict__ ha, int C, int K1,
             const double *__restrict__ term, int64_t N, int64_t P, double *__restrict__ terms,
             int64_t stride, double *__restrict__ terms2, int32_t *__restrict__ n_terms,
             int want_marg, int per_warp_ints)
{
    extern __shared__ int wt_sh[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int CK = C * K1, NQ = (K1 + 3) >> 2, RW = 2 * NQ, K = K1 + 1, D = 1 + 2 * K1;
    int *bm = wt_sh + (size_t)warp * (per_warp_ints + 3 * WT_LIST);
    int *E = bm + CK;
    int *rs = E + CK;
    int *cs = rs + CK;
    int *z0 = cs + CK;
    int *m2 = z0 + C;
    int *xc = m2 + K1 * K1;
    const int n_zero = 4 * CK + C + K1 * K1 + K1;  // everything up to here is zeroed per gene
    int *cr = xc + K1;  // the gene's row of the count table
    // compaction list of non-zero word pairs (index, words), behind the per-gene tables
    int *lp = bm + per_warp_ints;
    uint2 *lw = reinterpret_cast<uint2 *>(lp + WT_LIST);
    const bool marg = want_marg != 0;
    auto T = [&](int c) -> double { return c > 0 ? __ldg(term + c) : 0.0; };
    const int n_pairs = CK * NQ;  // (w02, w13) word pairs of a gene's block
    // x / d for 0 <= x < 2^16 as a multiply-shift (all indices here are far below that)
    const uint32_t mNQ = 0xFFFFFFFFu / (uint32_t)NQ + 1u, mK1 = 0xFFFFFFFFu / (uint32_t)K1 + 1u;
    const uint32_t mD = 0xFFFFFFFFu / (uint32_t)D + 1u, mK = 0xFFFFFFFFu / (uint32_t)K + 1u;
    auto dv = [](int x, uint32_t m) { return (int)__umulhi((uint32_t)x, m); };
    for (int64_t g = (int64_t)blockIdx.x * WT_WARPS + warp; g < P; g += (int64_t)gridDim.x * WT_WARPS) {
        for (int i = lane; i < (n_zero + 3) / 4; i += 32)  // (cr behind it is rewritten below)
            reinterpret_cast<int4 *>(bm)[i] = make_int4(0, 0, 0, 0);
        __syncwarp();
        for (int i = lane; i < CK; i += 32) cr[i] = __ldg(cnt + g * (int64_t)CK + i);
        __syncwarp();
        const int *crow = cr;
        uint32_t *hg = hits + g * (int64_t)CK * RW;
        // ---- pass 1 (non-zero word pairs compacted first: every lane then has work)
        int rounds = 0, n_last = 0;
        for (int p0 = 0; p0 < n_pairs;) {
            int n_list = 0;
            ++rounds;
            for (; p0 < n_pairs && n_list <= WT_LIST - 32 * WT_U; p0 += 32 * WT_U) {
                uint2 wu[WT_U];  // WT_U loads in flight before the first ballot needs one
#pragma unroll
                for (int u = 0; u < WT_U; ++u) {
                    const int p = p0 + 32 * u + lane;
                    wu[u] = *reinterpret_cast<const uint2 *>(hg + 2 * min(p, n_pairs - 1));
                }
#pragma unroll
                for (int u = 0; u < WT_U; ++u) {
                    const int p = p0 + 32 * u + lane;
                    const uint2 w = wu[u];
                    const bool nz = p < n_pairs && (w.x | w.y) != 0u;
                    const unsigned m = __ballot_sync(0xffffffffu, nz);
                    if (nz) {
                        const int at = n_list + __popc(m & ((1u << lane) - 1u));
                        lp[at] = p;
                        lw[at] = w;
                    }
                    n_list += __popc(m);
                }
            }
            __syncwarp();
            for (int i = lane; i < n_list; i += 32) {
                const int p = lp[i];
                const uint2 w = lw[i];
                const int row = dv(p, mNQ), q = p - row * NQ;  // row = y * K1 + v
                const int y = dv(row, mK1), v = row - y * K1;
                const int cf[4] = {(int)(w.x & 0xFFFFu), (int)(w.y & 0xFFFFu), (int)(w.x >> 16),
                                   (int)(w.y >> 16)};
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const int c = cf[f], a = 4 * q + f;
                    if (c == 0 || a >= K1) continue;
                    atomicOr(bm + y * K1 + a, 1 << v);
                    atomicAdd(rs + y * K1 + a, c);
                    atomicAdd(cs + y * K1 + v, c);
                    if (marg) atomicAdd(m2 + a * K1 + v, c);
                }
            }
            n_last = n_list;
            __syncwarp();
        }
        const bool keep_list = rounds == 1;  // the list still holds every non-zero pair of the gene
        __syncwarp();
        // ---- scan of the non-zero hit bins per (y, a) row, per-class totals of the x_j = 0 group
        int run = 0;
        for (int i0 = 0; i0 < CK; i0 += 32) {
            const int i = i0 + lane;
            const int n = i < CK ? __popc((unsigned)bm[i]) : 0;
            int tot;
            const int ex = wt_scan_excl(n, lane, tot);
            if (i < CK) E[i] = run + ex;
            run += tot;
        }
        const int n_hit = run;
        for (int y = lane; y < C; y += 32) {
            int s = 0;
            for (int v = 0; v < K1; ++v) s += crow[y * K1 + v] - cs[y * K1 + v];
            z0[y] = s;
        }
        if (marg) {
            for (int v = lane; v < K1; v += 32) {
                int s = 0;
                for (int y = 0; y < C; ++y) s += crow[y * K1 + v];
                xc[v] = s;
            }
        }
        __syncwarp();
        double *tg = terms + g * stride;
        // ---- terms of the marginal bins: D = 1 + 2*K1 per class (four bins per lane and round:
        // the four gathers are in flight together)
        for (int i0 = 0; i0 < C * D; i0 += 128) {
            int cc[4], pp[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + 32 * u + lane;
                cc[u] = 0, pp[u] = -1;
                if (i < C * D) {
                    const int y = dv(i, mD), t = i - y * D;
                    if (t == 0) {
                        cc[u] = (int)(clsz[y] - hs_ysum[y] - z0[y]);
                        pp[u] = y * D + E[y * K1];
                    } else if (t <= K1) {
                        cc[u] = crow[y * K1 + t - 1] - cs[y * K1 + t - 1];
                        pp[u] = y * D + E[y * K1] + t;
                    } else {
                        const int a = t - 1 - K1;
                        cc[u] = hsize[y * K1 + a] - rs[y * K1 + a];
                        pp[u] = y * D + 1 + K1 + a + E[y * K1 + a];
                    }
                }
            }
            double tv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) tv[u] = __ldg(term + max(cc[u], 0));
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (pp[u] >= 0) tg[pp[u]] = cc[u] > 0 ? tv[u] : 0.0;
        }
        // ---- terms of the hit bins: from the list of pass 1 when one round held them all, else
        // a second pass over the block (L1 / L2 hits; compacted the same way); the block is left
        // all zero
        for (int p0 = 0; p0 < n_pairs;) {
            int n_list = 0;
            if (keep_list) {
                n_list = n_last;
                p0 = n_pairs;
                for (int i = lane; i < n_list; i += 32)
                    *reinterpret_cast<uint2 *>(hg + 2 * lp[i]) = make_uint2(0u, 0u);
            }
            for (; p0 < n_pairs && n_list <= WT_LIST - 32 * WT_U; p0 += 32 * WT_U) {
                uint2 wu[WT_U];  // WT_U loads in flight before the first ballot needs one
#pragma unroll
                for (int u = 0; u < WT_U; ++u) {
                    const int p = p0 + 32 * u + lane;
                    wu[u] = *reinterpret_cast<const uint2 *>(hg + 2 * min(p, n_pairs - 1));
                }
#pragma unroll
                for (int u = 0; u < WT_U; ++u) {
                    const int p = p0 + 32 * u + lane;
                    const uint2 w = wu[u];
                    const bool nz = p < n_pairs && (w.x | w.y) != 0u;
                    const unsigned m = __ballot_sync(0xffffffffu, nz);
                    if (nz) {
                        *reinterpret_cast<uint2 *>(hg + 2 * p) = make_uint2(0u, 0u);
                        const int at = n_list + __popc(m & ((1u << lane) - 1u));
                        lp[at] = p;
                        lw[at] = w;
                    }
                    n_list += __popc(m);
                }
            }
            __syncwarp();
            for (int i = lane; i < n_list; i += 32) {
                const int p = lp[i];
                const uint2 w = lw[i];
                const int row = dv(p, mNQ), q = p - row * NQ;
                const int y = dv(row, mK1), v = row - y * K1;
                const int cf[4] = {(int)(w.x & 0xFFFFu), (int)(w.y & 0xFFFFu), (int)(w.x >> 16),
                                   (int)(w.y >> 16)};
#pragma unroll
                for (int f = 0; f < 4; ++f) {
                    const int c = cf[f], a = 4 * q + f;
                    if (c == 0 || a >= K1) continue;
                    const int pos = y * D + 2 + K1 + a + E[y * K1 + a] +
                                    __popc((unsigned)bm[y * K1 + a] & ((1u << v) - 1u));
                    tg[pos] = T(c);
                }
            }
            __syncwarp();
        }
        if (lane == 0) n_terms[g] = C * D + n_hit;
        // ---- H(x_j, x_g): bins (a, v) in order, a = 0: x_j = 0 (as k_finalize's marginal part)
        if (marg) {
            int sum_ha = 0, sum_xc = 0, sum_m2 = 0;
            for (int a = 0; a < K1; ++a) {
                sum_ha += ha[a];
                sum_xc += xc[a];
            }
            for (int i = lane; i < K1 * K1; i += 32) sum_m2 += m2[i];
            sum_m2 = __reduce_add_sync(0xffffffffu, sum_m2);
            double *t2 = terms2 + g * (int64_t)(K * K);
            for (int t = lane; t < K * K; t += 32) {
                const int a = dv(t, mK), v = t - a * K;
                int c;
                if (a == 0) {
                    if (v == 0) {
                        c = (int)(N - sum_ha - (sum_xc - sum_m2));
                    } else {
                        int s = 0;
                        for (int aa = 0; aa < K1; ++aa) s += m2[aa * K1 + v - 1];
                        c = xc[v - 1] - s;
                    }
                } else if (v == 0) {
                    int s = 0;
                    for (int vv = 0; vv < K1; ++vv) s += m2[(a - 1) * K1 + vv];
                    c = ha[a - 1] - s;
                } else {
                    c = m2[(a - 1) * K1 + v - 1];
                }
                t2[t] = T(c);
            }
        }
        __syncwarp();
    }
}

// One thread per gene adds its terms in order; the 32 genes of a warp are staged 32 terms at a
// time through shared memory (coalesced 256-byte loads per gene).  SEL: selector update + the
// fused tail; else out_full / out_marg.
template <bool SEL>
__global__ void __launch_bounds__(WC_THREADS)
k_wide_chain(const double *__restrict__ terms, int64_t stride, const double *__restrict__ terms2,
             const int32_t *__restrict__ n_terms, int KK2, int64_t P, double *__restrict__ out_full,
             double *__restrict__ out_marg, int want_marg, SelFuse sel)
{
    __shared__ double tile[WC_WARPS][32][33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if constexpr (SEL) sel_reset_sx(sel);
    const int64_t g0 = ((int64_t)blockIdx.x * WC_WARPS + warp) * 32;
    const int64_t g = g0 + lane;
    const bool active = g < P;
    double (*tw)[33] = tile[warp];
    auto chain = [&](const double *base, int64_t st, int n_mine) -> double {
        int n_max = n_mine;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) n_max = max(n_max, __shfl_xor_sync(0xffffffffu, n_max, o));
        double h = 0.0;
        double nxt[32];  // the next chunk's terms of the 32 genes, in flight while this one is added
        unsigned ok = 0;  // bit j: nxt[j] holds a term of gene j (else: a dummy load)
        auto fetch = [&](int c0) {
            ok = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int nj = __shfl_sync(0xffffffffu, n_mine, j);
                const bool v = c0 + lane < nj;
                ok |= (unsigned)v << j;
                // (unconditional load from a safe address: a predicated load would be followed by
                // a select that waits for it, and the 32 loads would no longer be in flight together)
                nxt[j] = __ldg(v ? base + (g0 + j) * st + c0 + lane : base);
            }
        };
        if (n_max > 0) fetch(0);
        for (int c0 = 0; c0 < n_max; c0 += 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) tw[j][lane] = ((ok >> j) & 1u) ? nxt[j] : 0.0;
            __syncwarp();
            if (c0 + 32 < n_max) fetch(c0 + 32);
#pragma unroll 8
            for (int i = 0; i < 32; ++i) h = __dadd_rn(h, tw[lane][i]);
            __syncwarp();
        }
        return h;
    };
    const int n_mine = active ? n_terms[g] : 0;
    const double h = chain(terms, stride, n_mine);
    double h2 = 0.0;
    if (want_marg) h2 = chain(terms2, KK2, active ? KK2 : 0);
    if constexpr (SEL) {
        double sc = 0.0;
        int64_t idx = -1;
        if (active) {
            const int j = *sel.gene_ptr;
            const bool unsup = sel.kind == PRB_SEL_CIFE_UNSUP;
            double pen = sel.penalty[g];
            sc = sel_update(sel.kind, sel.is_remove, sel.base[g], sel.Hx_all[j], unsup ? h : h2,
                            unsup ? 0.0 : sel.Hyx_all[j], unsup ? 0.0 : h, &pen, sel.score[g],
                            *sel.n_sel_ptr, sel.selected[g] != 0);
            sel.penalty[g] = pen;
            sel.score[g] = sc;
            idx = sel.gene_lo + g;
        }
        sel_tail<WC_THREADS>(sel, sc, idx);
    } else if (active) {
        out_full[g] = h;
        if (want_marg) out_marg[g] = h2;
    }
}

}  // namespace prb

using namespace prb;

static int wt_per_warp_ints(int C, int K1)
{
    const int n = 5 * C * K1 + C + K1 * K1 + K1;
    return (n + 3) & ~3;  // (16-byte rows: the tables are zeroed with int4 stores, the list behind
                          // them holds 8-byte entries)
}

// doubles per gene of the term buffer, bytes of shared memory of k_wide_terms (0: does not fit)
extern "C" int64_t prb_wide_terms_stride(int C, int K)
{
    const int K1 = K - 1;
    return (int64_t)C * (1 + 2 * K1) + (int64_t)C * K1 * K1;
}

extern "C" int prb_wide_supported(int C, int K)
{
    const int K1 = K - 1;
    if (K1 < 1 || K1 > 20 || C < 1) return 0;
    return (int64_t)(wt_per_warp_ints(C, K1) + 3 * WT_LIST) * 4 * WT_WARPS <= devinfo().smem_optin - 1024;
}

extern "C" int64_t prb_wide_chain_blocks(int64_t P) { return ceil_div(hmax(P, 1), WC_THREADS); }

extern "C" int prb_finalize_update_wide(
    int32_t *hits, const int32_t *cnt, const int64_t *classsizes, const int32_t *hsize,
    const int32_t *hs_ysum, const int32_t *ha, int C, int K, const double *table, int64_t N,
    int64_t P, double *terms, double *terms2, int32_t *n_terms, int kind, int is_remove,
    const double *base, double *penalty, double *score, const uint8_t *selected,
    const double *Hx_all, const double *Hyx_all, const int32_t *gene_ptr, const int32_t *n_sel_ptr,
    int32_t gene_lo, double *cand_score, int64_t *cand_idx, double *best, int32_t *ticket,
    const uint32_t *col_packed, const int64_t *col_begin, const int64_t *col_end,
    const int32_t *newpos, uint8_t *sx, const uint64_t *const *peer_bufs, int W, int rank,
    int32_t *xtag, prb_stream_t stream)
{
    PRB_REQUIRE(P >= 1 && prb_wide_supported(C, K), "unsupported C / K for the wide finalize path");
    PRB_REQUIRE(kind == PRB_SEL_CIFE || kind == PRB_SEL_JMI || kind == PRB_SEL_CIFE_UNSUP,
                "unknown selector kind");
    PRB_REQUIRE(hits && cnt && classsizes && hsize && hs_ysum && ha && table && terms && n_terms,
                "null pointer");
    PRB_REQUIRE(base && penalty && score && selected && Hx_all && gene_ptr && n_sel_ptr &&
                    cand_score && cand_idx && best && ticket && col_packed && col_begin && col_end && sx,
                "null pointer");
    const int want_marg = kind != PRB_SEL_CIFE_UNSUP;
    PRB_REQUIRE(!want_marg || (terms2 && Hyx_all), "supervised update needs terms2 and Hyx_all");
    PRB_REQUIRE(peer_bufs == nullptr || (W >= 1 && W <= 256 && rank >= 0 && rank < W && xtag),
                "bad peer buffer arguments");
    const int K1 = K - 1;
    const int pwi = wt_per_warp_ints(C, K1);
    const int smem = (pwi + 3 * WT_LIST) * 4 * WT_WARPS;
    PRB_CUDA(cudaFuncSetAttribute(k_wide_terms, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int64_t stride = prb_wide_terms_stride(C, K);
    const int grid_a = (int)hmin(ceil_div(P, WT_WARPS), (int64_t)devinfo().sm_count * 8);
    k_wide_terms<<<grid_a, WT_THREADS, smem, S(stream)>>>(
        reinterpret_cast<uint32_t *>(hits), cnt, classsizes, hsize, hs_ysum, ha, C, K1, table, N, P,
        terms, stride, terms2, n_terms, want_marg, pwi);
    PRB_LAUNCH_CHECK("k_wide_terms");
    SelFuse sel{kind, is_remove, base, penalty, score, selected, Hx_all, Hyx_all, gene_ptr,
                n_sel_ptr, gene_lo, cand_score, cand_idx, best, ticket, col_packed, col_begin,
                col_end, newpos, sx, (unsigned long long *const *)peer_bufs, peer_bufs ? W : 0,
                rank, xtag};
    k_wide_chain<true><<<(unsigned)prb_wide_chain_blocks(P), WC_THREADS, 0, S(stream)>>>(
        terms, stride, terms2, n_terms, K * K, P, nullptr, nullptr, want_marg, sel);
    PRB_LAUNCH_CHECK("k_wide_chain");
    return 0;
}
