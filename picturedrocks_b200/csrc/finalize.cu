// finalize.cu — histograms -> entropies, bit-compatible with the reference.
//
// Replaces the tail of entropy_wrt_i in _sparse_entropy_wrt
// (picturedrocks/markers/mutualinformation/infoset.py:392-397):
//     counts = counts / n_rows ; for e in counts: if e > 0: h += -e * np.log2(e)
// The per-bin term -(c/N)*log2(c/N) comes from a HOST-built fp64 table (glibc
// log2, the same libm call numba emits), and terms are added one at a time in
// ascending packed-bin order (y, master, code) with __dadd_rn, so the fp64 sum
// is the same sequence of roundings as the reference's loop.
//
// One warp per gene.  Bins that the streaming pass did not count are inferred
// by subtraction (the reference's "pre-count zeros through classsizes" trick,
// infoset.py:354-359, 390-391, taken one step further):
//   hit bins      (y, m>0, v>0)  = hits
//   (y, m>0, 0)   = hsize[state] - Σ_v hits
//   (y, 0,   v>0) = cnt[g][y][v] - Σ_{states of y} hits[.][v]
//   (y, 0,   0)   = classsize[y] - Σ hsize(states of y) - Σ_v (y,0,v)
#include <stdlib.h>

#include "selfuse.cuh"

namespace prb {

__device__ __forceinline__ double ordered_add(double h, int c, const double *__restrict__ term)
{
    const double tv = c > 0 ? __ldg(term + c) : 0.0;
    unsigned nz = __ballot_sync(0xffffffffu, c > 0);
    while (nz) {
        const int l = __ffs(nz) - 1;
        nz &= nz - 1;
        h = __dadd_rn(h, __shfl_sync(0xffffffffu, tv, l));
    }
    return h;
}

__global__ void __launch_bounds__(256)
k_finalize(int32_t *__restrict__ hits, int64_t hits_stride, const int32_t *__restrict__ cnt,
           const int64_t *__restrict__ clsz, const int32_t *__restrict__ hs_begin,
           const int32_t *__restrict__ hs_y, const int32_t *__restrict__ hsize,
           const int32_t *__restrict__ hs_ysum, const int32_t *__restrict__ ha,
           const int32_t *__restrict__ n_hs_ptr, int n_hs_cap, int C, int K,
           const double *__restrict__ term, int64_t N, int64_t P, double *__restrict__ out_full,
           double *__restrict__ out_marg, int scratch_ints)
{
    extern __shared__ int sh[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const int K1 = K - 1;
    const bool marg = out_marg != nullptr;
    int n_hs = 0;
    if (hits) n_hs = n_hs_ptr ? min(*n_hs_ptr, n_hs_cap) : n_hs_cap;

    int *rs = sh + (int64_t)warp * scratch_ints;  // [n_hs_cap]
    int *csy = rs + n_hs_cap;                      // [C*K1]
    int *z0 = csy + C * K1;                        // [C]
    int *m2 = z0 + C;                              // [K1*K1]  (marg only)
    int *xc = m2 + (marg ? K1 * K1 : 0);           // [K1]     (marg only)

    for (int64_t g = (int64_t)blockIdx.x * wpb + warp; g < P; g += (int64_t)gridDim.x * wpb) {
        for (int i = lane; i < scratch_ints; i += 32) rs[i] = 0;
        __syncwarp();
        int32_t *hrow = hits ? hits + g * hits_stride : nullptr;
        const int32_t *crow = cnt + g * (int64_t)C * K1;
        // phase A: row / column sums of the hit histogram
        for (int t = lane; t < n_hs * K1; t += 32) {
            const int x = hrow[t];
            if (x) {
                const int h = t / K1, v = t - h * K1;
                atomicAdd(rs + h, x);
                atomicAdd(csy + hs_y[h] * K1 + v, x);
                if (marg) atomicAdd(m2 + (h % K1) * K1 + v, x);
            }
        }
        __syncwarp();
        // phase A2: per-class totals of the master==0 group
        for (int t = lane; t < C * K1; t += 32) {
            const int c0 = crow[t];
            const int yy = t / K1;
            if (c0 - csy[t]) atomicAdd(z0 + yy, c0 - csy[t]);
            if (marg && c0) atomicAdd(xc + (t - yy * K1), c0);
        }
        __syncwarp();
        // phase B: ordered sum over (y, master, code)
        double h_acc = 0.0;
        for (int yy = 0; yy < C; ++yy) {
            const int hb = n_hs ? hs_begin[yy] : 0;
            const int he = n_hs ? min(hs_begin[yy + 1], n_hs) : 0;
            const int nb = (1 + max(he - hb, 0)) * K;
            for (int t0 = 0; t0 < nb; t0 += 32) {
                const int t = t0 + lane;
                int c = 0;
                if (t < nb) {
                    const int slot = t / K, v = t - slot * K;
                    if (slot == 0) {
                        if (v == 0)
                            c = (int)(clsz[yy] - (n_hs ? hs_ysum[yy] : 0) - z0[yy]);
                        else
                            c = crow[yy * K1 + v - 1] - csy[yy * K1 + v - 1];
                    } else {
                        const int h = hb + slot - 1;
                        if (v == 0) {
                            c = hsize[h] - rs[h];
                        } else {
                            c = hrow[h * K1 + v - 1];
                            if (c) hrow[h * K1 + v - 1] = 0;  // leave the buffer clean
                        }
                    }
                }
                h_acc = ordered_add(h_acc, c, term);
            }
        }
        if (lane == 0) out_full[g] = h_acc;
        if (marg) {
            // direct layout: master value a = (state % K1) + 1; sum over y.
            int sum_ha = 0, sum_xc = 0, sum_m2 = 0;
            for (int a = 0; a < K1; ++a) {
                sum_ha += (n_hs && ha) ? ha[a] : 0;
                sum_xc += xc[a];
            }
            for (int i = 0; i < K1 * K1; ++i) sum_m2 += m2[i];
            double h2 = 0.0;
            for (int t0 = 0; t0 < K * K; t0 += 32) {
                const int t = t0 + lane;
                int c = 0;
                if (t < K * K) {
                    const int a = t / K, v = t - a * K;
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
                        c = ((n_hs && ha) ? ha[a - 1] : 0) - s;
                    } else {
                        c = m2[(a - 1) * K1 + v - 1];
                    }
                }
                h2 = ordered_add(h2, c, term);
            }
            if (lane == 0) out_marg[g] = h2;
        }
        __syncwarp();
    }
}


// ---------------------------------------------------------------------------
// Direct (single selected gene) layout, one THREAD per gene: state id of a hit
// cell is y*(K-1) + x_j - 1, so hits[g] is a [C][K-1][K-1] cube and every bin of
// the (y, x_j, x_i) histogram follows from it, cnt[g][y][.], hsize[y][.] and
// classsizes by subtraction.  Each lane walks its own gene sequentially in the
// reference's bin order, so the ordered fp64 sum needs no cross-lane traffic and
// all 32 lanes stay busy.  TK1 > 0 fixes K-1 at compile time (loops unroll and
// the term-table gathers of a class are issued together).
// Scratch (shared memory, one int column per thread, stride = blockDim.x):
//   cs[K1] rs[K1] m2[K1*K1] xc[K1]
// ---------------------------------------------------------------------------
template <int TK1>
__global__ void __launch_bounds__(128)
k_finalize_direct(int32_t *__restrict__ hits, int64_t hits_stride,
                  const int32_t *__restrict__ cnt, const int64_t *__restrict__ clsz,
                  const int32_t *__restrict__ hsize, const int32_t *__restrict__ hs_ysum,
                  const int32_t *__restrict__ ha, int C, int k1_rt,
                  const double *__restrict__ term, int64_t N, int64_t P,
                  double *__restrict__ out_full, double *__restrict__ out_marg)
{
    extern __shared__ int sh[];
    const int K1 = TK1 ? TK1 : k1_rt;
    const int nt = blockDim.x;
    int *cs = sh + threadIdx.x;
    int *rs = cs + K1 * nt;
    int *m2 = rs + K1 * nt;
    int *xc = m2 + K1 * K1 * nt;
    const bool marg = out_marg != nullptr;
    const int64_t g = blockIdx.x * (int64_t)nt + threadIdx.x;
    if (g >= P) return;
    if (marg) {
#pragma unroll
        for (int i = 0; i < K1 * K1; ++i) m2[i * nt] = 0;
#pragma unroll
        for (int i = 0; i < K1; ++i) xc[i * nt] = 0;
    }
    const int32_t *crow = cnt + g * (int64_t)C * K1;
    int32_t *hrow = hits ? hits + g * hits_stride : nullptr;
    double h = 0.0;
    if (TK1 > 0) {
        // compile-time K-1: the class's hit square lives in registers (one pass over
        // global memory, 128-bit loads when the square is a multiple of 4 ints)
        constexpr int KK = TK1 > 0 ? TK1 * TK1 : 1;
        constexpr int TK = TK1 > 0 ? TK1 : 1;
        for (int y = 0; y < C; ++y, crow += TK) {
            int x[KK];
            int c0[TK];
#pragma unroll
            for (int v = 0; v < TK; ++v) c0[v] = crow[v];
            bool any = false;
            if (hrow) {
                if (KK % 4 == 0) {
#pragma unroll
                    for (int q = 0; q < KK / 4; ++q) {
                        const int4 t = *reinterpret_cast<const int4 *>(hrow + 4 * q);
                        x[4 * q] = t.x, x[4 * q + 1] = t.y, x[4 * q + 2] = t.z, x[4 * q + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < KK; ++i) x[i] = hrow[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < KK; ++i) x[i] = 0;
            }
            int cs[TK], rs[TK];
#pragma unroll
            for (int v = 0; v < TK; ++v) cs[v] = 0;
#pragma unroll
            for (int a = 0; a < TK; ++a) {
                rs[a] = 0;
#pragma unroll
                for (int v = 0; v < TK; ++v) {
                    const int t = x[a * TK + v];
                    rs[a] += t;
                    cs[v] += t;
                    any |= t != 0;
                    if (marg) m2[(a * TK + v) * nt] += t;
                }
            }
            int tot0 = 0;
#pragma unroll
            for (int v = 0; v < TK; ++v) {
                tot0 += c0[v] - cs[v];
                if (marg) xc[v * nt] += c0[v];
            }
            // term-table gathers of the whole class are independent: issue them together
            double t0[TK + 1], ta[TK * (TK + 1)];
            t0[0] = __ldg(term + (int)(clsz[y] - (hrow ? hs_ysum[y] : 0) - tot0));
#pragma unroll
            for (int v = 0; v < TK; ++v) t0[v + 1] = __ldg(term + (c0[v] - cs[v]));
            if (hrow) {
#pragma unroll
                for (int a = 0; a < TK; ++a) {
                    ta[a * (TK + 1)] = __ldg(term + (hsize[y * TK + a] - rs[a]));
#pragma unroll
                    for (int v = 0; v < TK; ++v)
                        ta[a * (TK + 1) + v + 1] = __ldg(term + x[a * TK + v]);
                }
            }
#pragma unroll
            for (int v = 0; v <= TK; ++v) h = __dadd_rn(h, t0[v]);
            if (hrow) {
#pragma unroll
                for (int i = 0; i < TK * (TK + 1); ++i) h = __dadd_rn(h, ta[i]);
                if (any) {  // leave the buffer clean for the next pass
                    if (KK % 4 == 0) {
#pragma unroll
                        for (int q = 0; q < KK / 4; ++q)
                            *reinterpret_cast<int4 *>(hrow + 4 * q) = make_int4(0, 0, 0, 0);
                    } else {
#pragma unroll
                        for (int i = 0; i < KK; ++i) hrow[i] = 0;
                    }
                }
                hrow += KK;
            }
        }
    } else
    for (int y = 0; y < C; ++y, crow += K1) {
        // pass 1: row / column sums of this class's hit square
#pragma unroll
        for (int v = 0; v < K1; ++v) cs[v * nt] = 0;
        if (hrow) {
#pragma unroll
            for (int a = 0; a < K1; ++a) {
                int r = 0;
#pragma unroll
                for (int v = 0; v < K1; ++v) {
                    const int x = hrow[a * K1 + v];
                    r += x;
                    cs[v * nt] += x;
                    if (marg) m2[(a * K1 + v) * nt] += x;
                }
                rs[a * nt] = r;
            }
        }
        int tot0 = 0;
#pragma unroll
        for (int v = 0; v < K1; ++v) {
            const int c0 = crow[v];
            tot0 += c0 - cs[v * nt];
            if (marg) xc[v * nt] += c0;
        }
        // master == 0 group: (y,0,0), (y,0,v)
        const int c00 = (int)(clsz[y] - (hrow ? hs_ysum[y] : 0) - tot0);
        h = __dadd_rn(h, __ldg(term + c00));
#pragma unroll
        for (int v = 0; v < K1; ++v) h = __dadd_rn(h, __ldg(term + (crow[v] - cs[v * nt])));
        // master == a groups
        if (hrow) {
#pragma unroll
            for (int a = 0; a < K1; ++a) {
                h = __dadd_rn(h, __ldg(term + (hsize[y * K1 + a] - rs[a * nt])));
#pragma unroll
                for (int v = 0; v < K1; ++v) {
                    const int x = hrow[a * K1 + v];
                    h = __dadd_rn(h, __ldg(term + x));
                    if (x) hrow[a * K1 + v] = 0;  // leave the buffer clean
                }
            }
            hrow += K1 * K1;
        }
    }
    out_full[g] = h;
    if (marg) {
        int sum_ha = 0, sum_xc = 0, sum_m2 = 0;
        for (int a = 0; a < K1; ++a) {
            sum_ha += (hits && ha) ? ha[a] : 0;
            sum_xc += xc[a * nt];
        }
        for (int i = 0; i < K1 * K1; ++i) sum_m2 += m2[i * nt];
        double h2 = __ldg(term + (int)(N - sum_ha - (sum_xc - sum_m2)));
        h2 = __dadd_rn(0.0, h2);
        for (int v = 0; v < K1; ++v) {
            int s = 0;
            for (int a = 0; a < K1; ++a) s += m2[(a * K1 + v) * nt];
            h2 = __dadd_rn(h2, __ldg(term + (xc[v * nt] - s)));
        }
        for (int a = 0; a < K1; ++a) {
            int s = 0;
            for (int v = 0; v < K1; ++v) s += m2[(a * K1 + v) * nt];
            h2 = __dadd_rn(h2, __ldg(term + (((hits && ha) ? ha[a] : 0) - s)));
            for (int v = 0; v < K1; ++v) h2 = __dadd_rn(h2, __ldg(term + m2[(a * K1 + v) * nt]));
        }
        out_marg[g] = h2;
    }
}


// ---------------------------------------------------------------------------
// Direct layout, compile-time K-1, L lanes per gene.  The L lanes of a group take L
// consecutive classes: each lane loads its class's hit square, derives the 1+K1+K1*(K1+1)
// bin counts by subtraction and gathers their terms — this part (most of the
// instructions and all of the memory latency) runs on all lanes in parallel.  The ordered
// sum then walks the L classes in order: every lane extends the running sum by ITS terms,
// and the value of the lane that owns the next class in order is broadcast with a shuffle,
// so the sequence of fp64 roundings is exactly the reference's (y, x_j, x_i) bin order.
// Classes past C contribute +0.0 terms (exact no-ops).
// ---------------------------------------------------------------------------
// SEL: the greedy step's tail fused in (prb_finalize_update) — the selector update of
// ITER:48-59 / 76-87 / 112-121 for the gene just finished (score algebra of select.cuh), the
// block's argmax candidate, the reduction of all candidates to ONE (score, index) record by
// the last block to finish (np.argmax of BASE:65 over this rank's genes), and the reset of
// the selected gene's entries of sx to "x_j = 0" (the streaming pass is over), so the next
// step's scatter starts from a clean vector.  Entropies are not written out then.



// VM: hits is the v-major table of the streamed lane-run kernel (vell2.cu), uint16
// hits[g][slot][v][4] with the counts of x_j - 1 = 0, 2, 1, 3 in the last dimension (one
// 8-byte store per run), instead of int32 hits[g][y][x_j - 1][v]; the slots of class y are
// [cls_slot0[y], cls_slot0[y+1]) — one per segment of the class, added up here — or just y when
// cls_slot0 is NULL.  hits_stride counts 32-bit words in both cases.
template <int TK1, int L, bool SEL, bool VM>
__global__ void __launch_bounds__(256, 3)
k_finalize_direct_grp(int32_t *__restrict__ hits, int64_t hits_stride,
                      const int32_t *__restrict__ cnt, const int64_t *__restrict__ clsz,
                      const int32_t *__restrict__ hsize, const int32_t *__restrict__ hs_ysum,
                      const int32_t *__restrict__ ha, int C, const double *__restrict__ term,
                      int64_t N, int64_t P, double *__restrict__ out_full,
                      double *__restrict__ out_marg, int want_marg, SelFuse sel,
                      const int32_t *__restrict__ cls_slot0)
{
    constexpr int TK = TK1, KK = TK1 * TK1, NT = 1 + TK + TK * (TK + 1);
    constexpr int SQ = VM ? TK * 4 : KK;  // counts of one class's hit square in registers
    constexpr int SQW = VM ? TK * 2 : KK; // 32-bit words of one slot's square in memory
    auto at = [](int a, int v) { return VM ? v * 4 + a : a * TK + v; };
    const int lane = threadIdx.x & 31;
    const int sub = lane & (L - 1);
    const int grp0 = lane & ~(L - 1);
    const int64_t g_raw = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / L;
    const bool active = g_raw < P;
    const int64_t g = active ? g_raw : P - 1;  // keep whole warps alive for the shuffles
    const bool marg = want_marg != 0;
    if constexpr (SEL) sel_reset_sx(sel);
    const int32_t *crow = cnt + g * (int64_t)C * TK;
    int32_t *hrow = hits ? hits + g * hits_stride : nullptr;
    // (a shared-memory copy of the first 4 096 terms was measured: no gain at config 4, where the
    // gathers then pay bank conflicts instead of L2 sectors, and a loss for small gene counts)
    auto T = [&](int c) -> double { return c ? __ldg(term + c) : 0.0; };
    int m2[KK], xc[TK];
#pragma unroll
    for (int i = 0; i < KK; ++i) m2[i] = 0;
#pragma unroll
    for (int i = 0; i < TK; ++i) xc[i] = 0;
    double h = 0.0;
    for (int y0 = 0; y0 < C; y0 += L) {
        const int y = y0 + sub;
        double t[NT];
        if (y < C) {
            int x[SQ], c0[TK];
#pragma unroll
            for (int v = 0; v < TK; ++v) c0[v] = crow[y * TK + v];
            bool any = false;
            const int sl0 = (VM && cls_slot0) ? cls_slot0[y] : y;
            const int sl1 = (VM && cls_slot0) ? cls_slot0[y + 1] : y + 1;
            if (hrow && VM) {
#pragma unroll
                for (int i = 0; i < SQ; ++i) x[i] = 0;
                for (int sl = sl0; sl < sl1; ++sl) {  // the segments of the class
                    uint32_t *hr = reinterpret_cast<uint32_t *>(hrow) + (int64_t)sl * SQW;
                    uint32_t w[SQW];
                    if (SQW % 4 == 0) {
#pragma unroll
                        for (int q = 0; q < SQW / 4; ++q) {
                            const uint4 t4 = *reinterpret_cast<const uint4 *>(hr + 4 * q);
                            w[4 * q] = t4.x, w[4 * q + 1] = t4.y, w[4 * q + 2] = t4.z, w[4 * q + 3] = t4.w;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < SQW / 2; ++q) {
                            const uint2 t2 = *reinterpret_cast<const uint2 *>(hr + 2 * q);
                            w[2 * q] = t2.x, w[2 * q + 1] = t2.y;
                        }
                    }
                    uint32_t nz = 0;
#pragma unroll
                    for (int v = 0; v < TK; ++v) {
                        const uint32_t w02 = w[2 * v], w13 = w[2 * v + 1];
                        nz |= w02 | w13;
                        x[v * 4 + 0] += (int)(w02 & 0xFFFFu);
                        x[v * 4 + 2] += (int)(w02 >> 16);
                        x[v * 4 + 1] += (int)(w13 & 0xFFFFu);
                        x[v * 4 + 3] += (int)(w13 >> 16);
                    }
                    if (nz && active) {  // leave the buffer clean for the next pass
#pragma unroll
                        for (int q = 0; q < SQW / 2; ++q)
                            *reinterpret_cast<uint2 *>(hr + 2 * q) = make_uint2(0u, 0u);
                    }
                }
            } else if (hrow) {
                int32_t *hr = hrow + (int64_t)sl0 * SQ;
                if (SQ % 4 == 0) {
#pragma unroll
                    for (int q = 0; q < SQ / 4; ++q) {
                        const int4 w = *reinterpret_cast<const int4 *>(hr + 4 * q);
                        x[4 * q] = w.x, x[4 * q + 1] = w.y, x[4 * q + 2] = w.z, x[4 * q + 3] = w.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < SQ; ++i) x[i] = hr[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < SQ; ++i) x[i] = 0;
            }
            int cs[TK], rs[TK];
#pragma unroll
            for (int v = 0; v < TK; ++v) cs[v] = 0;
#pragma unroll
            for (int a = 0; a < TK; ++a) {
                rs[a] = 0;
#pragma unroll
                for (int v = 0; v < TK; ++v) {
                    const int q = x[at(a, v)];
                    rs[a] += q;
                    cs[v] += q;
                    any |= q != 0;
                    m2[a * TK + v] += q;
                }
            }
            int tot0 = 0;
#pragma unroll
            for (int v = 0; v < TK; ++v) {
                tot0 += c0[v] - cs[v];
                xc[v] += c0[v];
            }
            t[0] = T((int)(clsz[y] - (hrow ? hs_ysum[y] : 0) - tot0));
#pragma unroll
            for (int v = 0; v < TK; ++v) t[1 + v] = T(c0[v] - cs[v]);
#pragma unroll
            for (int a = 0; a < TK; ++a) {
                t[1 + TK + a * (TK + 1)] = T((hrow ? hsize[y * TK + a] : 0) - rs[a]);
#pragma unroll
                for (int v = 0; v < TK; ++v) t[2 + TK + a * (TK + 1) + v] = T(x[at(a, v)]);
            }
            if (!VM && any && active) {  // leave the buffer clean for the next pass
                int32_t *hr = hrow + (int64_t)sl0 * SQ;
                if (SQ % 4 == 0) {
#pragma unroll
                    for (int q = 0; q < SQ / 4; ++q)
                        *reinterpret_cast<int4 *>(hr + 4 * q) = make_int4(0, 0, 0, 0);
                } else {
#pragma unroll
                    for (int i = 0; i < SQ; ++i) hr[i] = 0;
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < NT; ++i) t[i] = 0.0;
        }
        // ordered accumulation over the L classes of this round
#pragma unroll
        for (int l2 = 0; l2 < L; ++l2) {
            double hh = h;
#pragma unroll
            for (int i = 0; i < NT; ++i) hh = __dadd_rn(hh, t[i]);
            h = __shfl_sync(0xffffffffu, hh, grp0 + l2);
        }
    }
    if (!SEL && active && sub == 0) out_full[g] = h;
    double h2_out = 0.0;
    if (marg) {
        // class sums of the group -> every lane
#pragma unroll
        for (int o = 1; o < L; o <<= 1) {
#pragma unroll
            for (int i = 0; i < KK; ++i) m2[i] += __shfl_xor_sync(0xffffffffu, m2[i], o);
#pragma unroll
            for (int i = 0; i < TK; ++i) xc[i] += __shfl_xor_sync(0xffffffffu, xc[i], o);
        }
        if (active && sub == 0) {
            int sum_ha = 0, sum_xc = 0, sum_m2 = 0, hav[TK];
#pragma unroll
            for (int a = 0; a < TK; ++a) {
                hav[a] = (hits && ha) ? ha[a] : 0;
                sum_ha += hav[a];
                sum_xc += xc[a];
            }
#pragma unroll
            for (int i = 0; i < KK; ++i) sum_m2 += m2[i];
            double h2 = __dadd_rn(0.0, T((int)(N - sum_ha - (sum_xc - sum_m2))));
#pragma unroll
            for (int v = 0; v < TK; ++v) {
                int s = 0;
#pragma unroll
                for (int a = 0; a < TK; ++a) s += m2[a * TK + v];
                h2 = __dadd_rn(h2, T(xc[v] - s));
            }
#pragma unroll
            for (int a = 0; a < TK; ++a) {
                int s = 0;
#pragma unroll
                for (int v = 0; v < TK; ++v) s += m2[a * TK + v];
                h2 = __dadd_rn(h2, T(hav[a] - s));
#pragma unroll
                for (int v = 0; v < TK; ++v) h2 = __dadd_rn(h2, T(m2[a * TK + v]));
            }
            if (SEL)
                h2_out = h2;
            else
                out_marg[g] = h2;
        }
    }
    if constexpr (SEL) {
        // supervised: h = H(y, x_j, x_g), h2 = H(x_j, x_g); unsupervised: h = H(x_j, x_g)
        double sc = 0.0;
        int64_t idx = -1;
        if (active && sub == 0) {
            const int j = *sel.gene_ptr;
            const bool unsup = sel.kind == PRB_SEL_CIFE_UNSUP;
            double pen = sel.penalty[g];
            sc = sel_update(sel.kind, sel.is_remove, sel.base[g], sel.Hx_all[j], unsup ? h : h2_out,
                            unsup ? 0.0 : sel.Hyx_all[j], unsup ? 0.0 : h, &pen, sel.score[g],
                            *sel.n_sel_ptr, sel.selected[g] != 0);
            sel.penalty[g] = pen;
            sel.score[g] = sc;
            idx = sel.gene_lo + g;
        }
        sel_tail<256>(sel, sc, idx);
    }
}

}  // namespace prb

using namespace prb;

extern "C" int prb_finalize(int32_t *hits, const int32_t *cnt, const int64_t *classsizes,
                            const int32_t *hs_begin, const int32_t *hs_y, const int32_t *hsize,
                            const int32_t *hs_ysum, const int32_t *ha, const int32_t *n_hs_ptr,
                            int64_t n_hs_cap, int C, int K, const double *table, int64_t N,
                            int64_t P, double *out_full, double *out_marg, prb_stream_t stream)
{
    if (P == 0) return 0;
    PRB_REQUIRE(C >= 1 && K >= 2 && K <= PRB_MAX_CODES, "bad C or K");
    PRB_REQUIRE(out_full != nullptr, "out_full is required");
    if (!hits) n_hs_cap = 0;
    const int K1 = K - 1;
    const int64_t scratch = n_hs_cap + (int64_t)C * K1 + C + (out_marg ? (int64_t)K1 * K1 + K1 : 0);
    const int64_t smem_max = devinfo().smem_optin - 1024;
    PRB_REQUIRE(scratch * 4 <= smem_max, "finalize scratch does not fit in shared memory");
    int wpb = 8;
    while (wpb > 1 && scratch * 4 * wpb > smem_max) wpb >>= 1;
    // keep a few blocks per SM resident when scratch is small
    const int64_t smem = scratch * 4 * wpb;
    PRB_CUDA(cudaFuncSetAttribute(k_finalize, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_max));
    const int grid = (int)hmin(ceil_div(P, wpb), (int64_t)devinfo().sm_count * 8);
    k_finalize<<<grid, wpb * 32, smem, S(stream)>>>(
        hits, n_hs_cap * K1, cnt, classsizes, hs_begin, hs_y, hsize, hs_ysum, ha, n_hs_ptr,
        (int)n_hs_cap, C, K, table, N, P, out_full, out_marg, (int)scratch);
    PRB_LAUNCH_CHECK("k_finalize");
    return 0;
}


// lanes per gene of k_finalize_direct_grp: 8 when there are enough genes to fill the GPU, more
// for small gene counts (a multi-GPU shard), where the kernel is bound by the per-gene chain
static int grp_lanes(int64_t P, int C)
{
    static const int forced = [] {
        const char *e = getenv("PRB_FINALIZE_L");
        return e ? atoi(e) : 0;
    }();
    int L = P * 8 >= (int64_t)devinfo().sm_count * 256 * 3 ? 8 : 16;
    if (forced == 8 || forced == 16 || forced == 32) L = forced;
    if (C <= 8) L = 8;
    return L;
}

template <int TK1>
static int launch_direct(int32_t *hits, int64_t stride /* x_j-major */, const int32_t *cnt, const int64_t *clsz,
                         const int32_t *hsize, const int32_t *hs_ysum, const int32_t *ha, int C,
                         int K1, const double *table, int64_t N, int64_t P, double *out_full,
                         double *out_marg, int vmajor, const int32_t *cls_slot0, int n_slot,
                         cudaStream_t st)
{
    if constexpr (TK1 > 0) {
        const int threads = 256;
        const int L = grp_lanes(P, C);
        const int wm = out_marg != nullptr;
        if (vmajor) stride = (int64_t)n_slot * TK1 * 2;
#define PRB_FG(LL, VM)                                                                                  \
    k_finalize_direct_grp<TK1, LL, false, VM><<<(unsigned)ceil_div(P * LL, threads), threads, 0, st>>>( \
        hits, stride, cnt, clsz, hsize, hs_ysum, ha, C, table, N, P, out_full, out_marg, wm,            \
        SelFuse{}, cls_slot0)
        if (vmajor) {
            if (L == 32)
                PRB_FG(32, true);
            else if (L == 16)
                PRB_FG(16, true);
            else
                PRB_FG(8, true);
        } else {
            if (L == 32)
                PRB_FG(32, false);
            else if (L == 16)
                PRB_FG(16, false);
            else
                PRB_FG(8, false);
        }
#undef PRB_FG
        PRB_LAUNCH_CHECK("k_finalize_direct_grp");
        return 0;
    }
    const int64_t per_thread = (int64_t)(3 * K1 + K1 * K1) * 4;
    int threads = 128;
    const int64_t smem_max = devinfo().smem_optin - 1024;
    while (threads > 32 && per_thread * threads > smem_max) threads >>= 1;
    PRB_REQUIRE(per_thread * threads <= smem_max, "finalize scratch does not fit in shared memory");
    PRB_REQUIRE(!vmajor, "v-major hits need K <= 5 here");
    // spread genes over all SMs: small blocks when there are few genes
    while (threads > 32 && ceil_div(P, threads) < 2 * devinfo().sm_count) threads >>= 1;
    auto kern = k_finalize_direct<TK1>;
    PRB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_max));
    kern<<<(unsigned)ceil_div(P, threads), threads, per_thread * threads, st>>>(
        hits, stride, cnt, clsz, hsize, hs_ysum, ha, C, K1, table, N, P, out_full, out_marg);
    PRB_LAUNCH_CHECK("k_finalize_direct");
    return 0;
}

extern "C" int prb_finalize_direct(int32_t *hits, const int32_t *cnt, const int64_t *classsizes,
                                   const int32_t *hsize, const int32_t *hs_ysum,
                                   const int32_t *ha, int C, int K, const double *table,
                                   int64_t N, int64_t P, double *out_full, double *out_marg,
                                   int vmajor, const int32_t *cls_slot0, int n_slot,
                                   prb_stream_t stream)
{
    if (P == 0) return 0;
    PRB_REQUIRE(!vmajor || n_slot >= 1, "v-major hits need the number of slots per gene");
    PRB_REQUIRE(C >= 1 && K >= 2 && K <= PRB_MAX_CODES, "bad C or K");
    PRB_REQUIRE(out_full != nullptr, "out_full is required");
    const int K1 = K - 1;
    const int64_t stride = (int64_t)C * K1 * K1;
#define PRB_FD(T)                                                                              \
    return launch_direct<T>(hits, stride, cnt, classsizes, hsize, hs_ysum, ha, C, K1, table, N, \
                            P, out_full, out_marg, vmajor, cls_slot0, n_slot, S(stream))
    switch (K1) {
        case 1: PRB_FD(1);
        case 2: PRB_FD(2);
        case 3: PRB_FD(3);
        case 4: PRB_FD(4);
        default: PRB_FD(0);
    }
#undef PRB_FD
}


// ---- the greedy step's tail in one launch (V-layout path, K <= 5) ---------------------------
extern "C" int64_t prb_finalize_update_blocks(int64_t P, int C)
{
    return ceil_div(hmax(P, 1) * grp_lanes(P, C), 256);
}

template <int TK1>
static int launch_update(int32_t *hits, const int32_t *cnt, const int64_t *clsz,
                         const int32_t *hsize, const int32_t *hs_ysum, const int32_t *ha, int C,
                         const double *table, int64_t N, int64_t P, int want_marg,
                         const SelFuse &sel, int vmajor, const int32_t *cls_slot0, int n_slot,
                         cudaStream_t st)
{
    const int threads = 256;
    const int L = grp_lanes(P, C);
    const int64_t stride = vmajor ? (int64_t)n_slot * TK1 * 2 : (int64_t)C * TK1 * TK1;
#define PRB_FU(LL, VM)                                                                                 \
    k_finalize_direct_grp<TK1, LL, true, VM><<<(unsigned)ceil_div(P * LL, threads), threads, 0, st>>>( \
        hits, stride, cnt, clsz, hsize, hs_ysum, ha, C, table, N, P, nullptr, nullptr, want_marg,      \
        sel, cls_slot0)
    if (vmajor) {
        if (L == 32)
            PRB_FU(32, true);
        else if (L == 16)
            PRB_FU(16, true);
        else
            PRB_FU(8, true);
    } else {
        if (L == 32)
            PRB_FU(32, false);
        else if (L == 16)
            PRB_FU(16, false);
        else
            PRB_FU(8, false);
    }
#undef PRB_FU
    PRB_LAUNCH_CHECK("k_finalize_update");
    return 0;
}

extern "C" int prb_finalize_update(int32_t *hits, const int32_t *cnt, const int64_t *classsizes,
                                   const int32_t *hsize, const int32_t *hs_ysum,
                                   const int32_t *ha, int C, int K, const double *table,
                                   int64_t N, int64_t P, int kind, int is_remove,
                                   const double *base, double *penalty, double *score,
                                   const uint8_t *selected, const double *Hx_all,
                                   const double *Hyx_all, const int32_t *gene_ptr,
                                   const int32_t *n_sel_ptr, int32_t gene_lo, double *cand_score,
                                   int64_t *cand_idx, double *best, int32_t *ticket,
                                   const uint32_t *col_packed, const int64_t *col_begin,
                                   const int64_t *col_end, const int32_t *newpos, uint8_t *sx,
                                   int vmajor, const int32_t *cls_slot0, int n_slot,
                                   const uint64_t *const *peer_bufs, int W, int rank,
                                   int32_t *xtag, prb_stream_t stream)
{
    PRB_REQUIRE(P >= 1 && C >= 1 && K >= 2 && K <= 5, "the fused step tail needs 2 <= K <= 5");
    PRB_REQUIRE(peer_bufs == nullptr || (W >= 1 && W <= 256 && rank >= 0 && rank < W && xtag),
                "bad peer buffer arguments");
    PRB_REQUIRE(!vmajor || n_slot >= 1, "v-major hits need the number of slots per gene");
    PRB_REQUIRE(kind == PRB_SEL_CIFE || kind == PRB_SEL_JMI || kind == PRB_SEL_CIFE_UNSUP,
                "unknown selector kind");
    PRB_REQUIRE(hits && base && penalty && score && selected && Hx_all && gene_ptr && n_sel_ptr,
                "null pointer");
    PRB_REQUIRE(kind == PRB_SEL_CIFE_UNSUP || Hyx_all, "supervised update needs Hyx_all");
    PRB_REQUIRE(cand_score && cand_idx && best && ticket && col_packed && col_begin && col_end && sx,
                "null pointer");
    SelFuse sel{kind, is_remove, base, penalty, score, selected, Hx_all, Hyx_all, gene_ptr,
                n_sel_ptr, gene_lo, cand_score, cand_idx, best, ticket, col_packed, col_begin,
                col_end, newpos, sx, (unsigned long long *const *)peer_bufs, peer_bufs ? W : 0,
                rank, xtag};
    const int want_marg = kind != PRB_SEL_CIFE_UNSUP;
#define PRB_LU(T)                                                                               \
    return launch_update<T>(hits, cnt, classsizes, hsize, hs_ysum, ha, C, table, N, P, want_marg, \
                            sel, vmajor, cls_slot0, n_slot, S(stream))
    switch (K - 1) {
        case 1: PRB_LU(1);
        case 2: PRB_LU(2);
        case 3: PRB_LU(3);
        default: PRB_LU(4);
    }
#undef PRB_LU
}
