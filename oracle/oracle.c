/*
 * oracle.c — CPU restatement of the PicturedRocks mutual-information hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the
 * product (picturedrocks_b200) never does.
 *
 * Every function restates one function of the reference, cited as
 * <file>:<lines> relative to the reference tree:
 *   INFOSET = picturedrocks/markers/mutualinformation/infoset.py
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against
 * golden vectors generated from the unmodified reference (numba JIT, numpy
 * 2.3.5) by tests/golden/make_golden.py, and tests/test_oracle_vs_reference.py
 * compares it with the live reference when /root/reference is present.
 *
 * Arithmetic notes (why results are bit-identical to the numba-JIT reference):
 *   - counts are int64, exact;
 *   - each entropy term is  e = count / n_rows ; h += -e * log2(e)  with glibc
 *     log2 (numba lowers np.log2 on a float64 scalar to the libm call), summed
 *     sequentially in ascending packed-bin order, zero bins skipped
 *     (INFOSET:392-397, 137-141, 423-427);
 *   - compile with -ffp-contract=off so no FMA is formed.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- shared tail: Σ -p·log2 p over a counts array, ascending index ---------- */
static double entropy_of_counts(const int64_t *counts, int64_t nbins, int64_t n_rows)
{
    double h = 0.0;
    const double n = (double)n_rows;
    for (int64_t b = 0; b < nbins; ++b) {
        if (counts[b] > 0) {
            double e = (double)counts[b] / n;
            h += -e * log2(e);
        }
    }
    return h;
}

/* ---------------------------------------------------------------------------
 * _sparse_entropy_wrt  (INFOSET:306-410)
 *
 * One gene = one two-pointer merge of the gene's CSC column (row ids ascending)
 * with the master column (row ids ascending, values already packed from the
 * selected columns).  Bin id = (y << yshift) + (master << shift) + code.
 * Rows where both are zero are never visited: they are pre-counted in bin
 * (y,0,0) through classsizes (INFOSET:354-359) and moved out one at a time
 * (INFOSET:390-391).
 * --------------------------------------------------------------------------- */
#define DEFINE_SPARSE_WRT(NAME, IDX_T, DAT_T)                                              \
    int NAME(const IDX_T *dindices, const int64_t *dindptr, const DAT_T *ddata,            \
             const int64_t *mindices, const int64_t *mdata_in, int64_t m_nnz,              \
             int64_t n_rows, int64_t feat_begin, int64_t feat_end, int64_t n_mcols,        \
             int64_t shift, int64_t ybits, const int64_t *y, const double *classsizes,     \
             int64_t n_classes, double *out, int nthreads)                                 \
    {                                                                                      \
        const int64_t nbins = (int64_t)1 << (shift * (n_mcols + 1) + ybits);               \
        const int64_t yshift = shift * (n_mcols + 1);                                      \
        int64_t *mdata = (int64_t *)malloc(sizeof(int64_t) * (size_t)(m_nnz > 0 ? m_nnz : 1)); \
        if (!mdata) return 1;                                                              \
        for (int64_t t = 0; t < m_nnz; ++t) mdata[t] = mdata_in[t] << shift; /* :350 */    \
        int failed = 0;                                                                    \
        (void)nthreads;                                                                    \
        _Pragma("omp parallel num_threads(nthreads > 0 ? nthreads : 1)")                   \
        {                                                                                  \
            int64_t *counts = (int64_t *)malloc(sizeof(int64_t) * (size_t)nbins);          \
            if (!counts) {                                                                 \
                _Pragma("omp atomic write") failed = 1;                                    \
            } else {                                                                       \
                _Pragma("omp for schedule(dynamic, 8)")                                    \
                for (int64_t g = feat_begin; g < feat_end; ++g) {                          \
                    const IDX_T *cind = dindices + dindptr[g];                             \
                    const DAT_T *cdat = ddata + dindptr[g];                                \
                    const int64_t cn = dindptr[g + 1] - dindptr[g];                        \
                    memset(counts, 0, sizeof(int64_t) * (size_t)nbins);       /* :353 */   \
                    if (ybits) {                                                           \
                        for (int64_t c = 0; c < n_classes; ++c)                            \
                            counts[c << yshift] = (int64_t)classsizes[c];     /* :356 */   \
                    } else {                                                               \
                        counts[0] = n_rows;                                   /* :359 */   \
                    }                                                                      \
                    int64_t ci = 0, mi = 0;                                                \
                    while (ci < cn || mi < m_nnz) {                           /* :365 */   \
                        int64_t curval, row;                                               \
                        if (mi >= m_nnz || (ci < cn && (int64_t)cind[ci] < mindices[mi])) {\
                            curval = (int64_t)cdat[ci]; row = (int64_t)cind[ci]; ++ci;     \
                        } else if (ci >= cn || (int64_t)cind[ci] > mindices[mi]) {         \
                            curval = mdata[mi]; row = mindices[mi]; ++mi;                  \
                        } else {                                                           \
                            curval = (int64_t)cdat[ci] + mdata[mi];                        \
                            row = (int64_t)cind[ci]; ++ci; ++mi;                           \
                        }                                                                  \
                        const int64_t yval = ybits ? (y[row] << yshift) : 0;  /* :386 */   \
                        counts[yval] -= 1;                                    /* :390 */   \
                        counts[yval + curval] += 1;                           /* :391 */   \
                    }                                                                      \
                    out[g - feat_begin] = entropy_of_counts(counts, nbins, n_rows);        \
                }                                                                          \
                free(counts);                                                              \
            }                                                                              \
        }                                                                                  \
        free(mdata);                                                                       \
        return failed;                                                                     \
    }

DEFINE_SPARSE_WRT(orc_sparse_entropy_wrt_i32_i64, int32_t, int64_t)
DEFINE_SPARSE_WRT(orc_sparse_entropy_wrt_i64_i64, int64_t, int64_t)
DEFINE_SPARSE_WRT(orc_sparse_entropy_wrt_i32_u8, int32_t, uint8_t)

/* ---------------------------------------------------------------------------
 * _sparse_entropy  (INFOSET:413-428): entropy of one packed sparse column;
 * the zero bin is n_rows minus the number of stored entries.
 * --------------------------------------------------------------------------- */
double orc_sparse_entropy(const int64_t *data, int64_t nnz, int64_t n_rows, int64_t max_val)
{
    const int64_t nbins = max_val + 1;
    int64_t *counts = (int64_t *)calloc((size_t)nbins, sizeof(int64_t));
    if (!counts) return NAN;
    counts[0] = n_rows;
    for (int64_t t = 0; t < nnz; ++t) {
        counts[0] -= 1;
        counts[data[t]] += 1;
    }
    double h = entropy_of_counts(counts, nbins, n_rows);
    free(counts);
    return h;
}

/* ---------------------------------------------------------------------------
 * InformationSet.entropy  (INFOSET:114-142): dense, C-order int64 X[N, ld];
 * cols may hold negative indices (numpy wrap-around: -1 is the y column).
 * First column is the most significant digit.
 * --------------------------------------------------------------------------- */
double orc_dense_entropy(const int64_t *X, int64_t N, int64_t P, const int64_t *cols,
                         int64_t n_cols, int64_t shift)
{
    const int64_t nbins = (int64_t)1 << (shift * n_cols);
    int64_t *counts = (int64_t *)calloc((size_t)nbins, sizeof(int64_t));
    if (!counts) return NAN;
    for (int64_t i = 0; i < N; ++i) {
        int64_t ind = 0;
        for (int64_t j = 0; j < n_cols; ++j) {
            int64_t c = cols[j] < 0 ? cols[j] + P : cols[j];
            ind = (ind << shift) | X[i * P + c];
        }
        counts[ind] += 1;
    }
    double h = entropy_of_counts(counts, nbins, N);
    free(counts);
    return h;
}

/* InformationSet.entropy_wrt  (INFOSET:144-173): entropy(cols ⊕ [i]) for every
 * feature i in [0, n_feats). */
int orc_dense_entropy_wrt(const int64_t *X, int64_t N, int64_t P, const int64_t *cols,
                          int64_t n_cols, int64_t shift, int64_t n_feats, double *out,
                          int nthreads)
{
    int failed = 0;
    (void)nthreads;
#pragma omp parallel num_threads(nthreads > 0 ? nthreads : 1)
    {
        int64_t *tmpl = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_cols + 1));
        if (!tmpl) {
#pragma omp atomic write
            failed = 1;
        } else {
            for (int64_t j = 0; j < n_cols; ++j) tmpl[j] = cols[j];
#pragma omp for schedule(dynamic, 4)
            for (int64_t i = 0; i < n_feats; ++i) {
                tmpl[n_cols] = i;
                out[i] = orc_dense_entropy(X, N, P, tmpl, n_cols + 1, shift);
            }
            free(tmpl);
        }
    }
    return failed;
}

/* ---------------------------------------------------------------------------
 * quantile_discretize  (INFOSET:50-83) with numpy 2.3.5 semantics for
 * np.percentile(method="linear") and np.digitize(right=True); numpy is a
 * third-party dependency not vendored in the reference (requirements.txt:4,
 * un-pinned; golden vectors were generated under numpy 2.3.5):
 *   numpy/lib/_function_base_impl.py
 *     percentile: q = q / a.dtype.type(100) for float arrays  → q in a's dtype
 *     'linear' virtual index = (n - 1) * q                     (same dtype as q)
 *     _get_indexes: prev = floor(vi), next = prev + 1, both → last when vi >= n-1
 *     _get_gamma:   gamma = vi - prev                          (dtype of vi)
 *     _lerp:        a + (b-a)*g, replaced by b - (b-a)*(1-g) where g >= 0.5
 * dtype rule: float32 column → all arithmetic float32; float64 → float64;
 * integer column → q/vi/gamma float64, (b-a) integer, result float64.
 *
 * Per column: edge_0 = percentile(col, 100/k); edge_i = percentile(values
 * strictly above edge_{i-1}, 100/(k-i)); stop when the column max <= last edge
 * or nothing is left.  code = number of edges strictly below the value.
 * The reference selects order statistics with np.partition; a full sort gives
 * the same order statistics.
 * --------------------------------------------------------------------------- */
#define DEFINE_CMP(NAME, T)                                   \
    static int NAME(const void *a, const void *b)             \
    {                                                         \
        T x = *(const T *)a, y = *(const T *)b;               \
        return (x > y) - (x < y);                             \
    }
DEFINE_CMP(cmp_f32, float)
DEFINE_CMP(cmp_f64, double)
DEFINE_CMP(cmp_i64, int64_t)

static float pct_f32(const float *s, int64_t n, double q_py)
{
    volatile float q = (float)q_py / 100.0f;
    volatile float vi = (float)(n - 1) * q;
    if (vi >= (float)(n - 1)) return s[n - 1];
    float pf = floorf(vi);
    int64_t p = (int64_t)pf;
    volatile float g = vi - pf;
    float a = s[p], b = s[p + 1];
    volatile float d = b - a;
    volatile float r;
    if (g >= 0.5f) {
        volatile float omg = 1.0f - g;
        volatile float t = d * omg;
        r = b - t;
    } else {
        volatile float t = d * g;
        r = a + t;
    }
    return r;
}

static double pct_f64(const double *s, int64_t n, double q_py)
{
    double q = q_py / 100.0;
    double vi = (double)(n - 1) * q;
    if (vi >= (double)(n - 1)) return s[n - 1];
    double pf = floor(vi);
    int64_t p = (int64_t)pf;
    double g = vi - pf;
    double a = s[p], b = s[p + 1];
    double d = b - a;
    if (g >= 0.5) return b - d * (1.0 - g);
    return a + d * g;
}

static double pct_i64(const int64_t *s, int64_t n, double q_py)
{
    double q = q_py / 100.0;
    double vi = (double)(n - 1) * q;
    if (vi >= (double)(n - 1)) return (double)s[n - 1];
    double pf = floor(vi);
    int64_t p = (int64_t)pf;
    double g = vi - pf;
    int64_t a = s[p], b = s[p + 1];
    int64_t d = b - a;
    if (g >= 0.5) return (double)b - (double)d * (1.0 - g);
    return (double)a + (double)d * g;
}

/* X is C-order [N, P] (as numpy hands it over); codes out is column-major
 * uint8 [P][N]; edges out is [P][k-1] in E_T with n_edges[P]. */
#define DEFINE_QD(NAME, T, E_T, CMP, PCT)                                                  \
    int NAME(const T *X, int64_t N, int64_t P, int64_t k, uint8_t *codes, E_T *edges,      \
             int32_t *n_edges, int nthreads)                                               \
    {                                                                                      \
        int failed = 0;                                                                    \
        (void)nthreads;                                                                    \
        if (k < 2 || k > 256) return 2;                                                    \
        _Pragma("omp parallel num_threads(nthreads > 0 ? nthreads : 1)")                   \
        {                                                                                  \
            T *s = (T *)malloc(sizeof(T) * (size_t)(N > 0 ? N : 1));                       \
            if (!s) {                                                                      \
                _Pragma("omp atomic write") failed = 1;                                    \
            } else {                                                                       \
                _Pragma("omp for schedule(dynamic, 16)")                                   \
                for (int64_t j = 0; j < P; ++j) {                                          \
                    for (int64_t i = 0; i < N; ++i) s[i] = X[i * P + j];                   \
                    qsort(s, (size_t)N, sizeof(T), CMP);                                   \
                    E_T *e = edges + j * (k - 1);                                          \
                    int32_t ne = 0;                                                        \
                    int64_t start = 0;                                                     \
                    const T cmax = s[N - 1];                                               \
                    for (int64_t i = 0; i < k - 1; ++i) {                                  \
                        if (ne > 0) {                                                      \
                            E_T m = e[ne - 1];                                             \
                            while (start < N && !((E_T)s[start] > m)) ++start;             \
                            if ((E_T)cmax <= m || start >= N) break;                       \
                        }                                                                  \
                        e[ne++] = PCT(s + start, N - start, 100.0 / (double)(k - i));      \
                    }                                                                      \
                    n_edges[j] = ne;                                                       \
                    uint8_t *cj = codes + j * N;                                           \
                    for (int64_t i = 0; i < N; ++i) {                                      \
                        E_T x = (E_T)X[i * P + j];                                         \
                        int c = 0;                                                         \
                        while (c < ne && e[c] < x) ++c;                                    \
                        cj[i] = (uint8_t)c;                                                \
                    }                                                                      \
                }                                                                          \
                free(s);                                                                   \
            }                                                                              \
        }                                                                                  \
        return failed;                                                                     \
    }

DEFINE_QD(orc_quantile_discretize_f32, float, float, cmp_f32, pct_f32)
DEFINE_QD(orc_quantile_discretize_f64, double, double, cmp_f64, pct_f64)
DEFINE_QD(orc_quantile_discretize_i64, int64_t, double, cmp_i64, pct_i64)

/* The fp64 term table the GPU path consumes: T[c] = -(c/N)·log2(c/N), T[0]=0.
 * Exposed so tests can check the product's host-built table against glibc. */
void orc_term_table(int64_t n_rows, double *out)
{
    out[0] = 0.0;
    for (int64_t c = 1; c <= n_rows; ++c) {
        double e = (double)c / (double)n_rows;
        out[c] = -e * log2(e);
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
