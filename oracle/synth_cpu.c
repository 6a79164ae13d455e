/*
 * synth_cpu.c — the benchmark's synthetic discretised matrices, generated on the host.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT (part of liboracle.so; only tests/, bench.py's CPU
 * legs and tests/golden/make_golden_large.py use it).  BASELINE.json names shapes only;
 * the recipe is this repo's (picturedrocks_b200/synth.py documents it and holds the numpy
 * version `DiscretisedSpec.cpu_codes`, picturedrocks_b200/csrc/synth.cu the device one).
 * This file is the same 64-bit integer arithmetic in plain C + OpenMP so that the FULL
 * config-4 matrix (1,306,127 x 27,998, 2.6e9 stored entries) can be handed to the CPU
 * oracle in about a minute: by-gene CSC with int32 row ids and uint8 codes, rows
 * ascending inside a gene.  tests/test_synth_cpu.py checks it against the numpy version.
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GOLD 0x9E3779B97F4A7C15ull
#define M2 0xD1B54A32D192ED03ull
#define SALT_A 0x1234567ull
#define SALT_B 0xC1A55ull
#define SALT_Y 0xA5A5A5A5A5A5A5A5ull

static inline uint64_t mix64(uint64_t z)
{
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

void orc_synth_labels(int64_t *y, int64_t N, int64_t C, uint64_t seed)
{
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i)
        y[i] = (int64_t)((mix64((seed ^ SALT_Y) + (uint64_t)i * GOLD) >> 33) % (uint64_t)C);
}

typedef struct {
    uint64_t gk;
    uint32_t thr[256];
    uint8_t pidx[256];
} gene_params;

static void gene_setup(gene_params *gp, int64_t gene, int64_t C, uint32_t mean_q24, uint64_t seed)
{
    static const uint64_t MULT[8] = {1, 2, 3, 4, 4, 6, 8, 12};
    const int64_t gene_eff = (gene % 4001 == 17 && gene > 0) ? gene - 1 : gene; /* planted duplicate */
    const int all_zero = gene_eff % 5003 == 29;                                  /* planted all-zero */
    const uint64_t gk = mix64(seed ^ ((uint64_t)gene_eff * GOLD));
    const uint64_t u32 = mix64(gk ^ SALT_A) >> 32;
    const uint64_t u2 = (u32 * u32) >> 32, u4 = (u2 * u2) >> 32, u8 = (u4 * u4) >> 32;
    const uint64_t cap = (uint64_t)(0.95 * 16777216.0);
    uint64_t dmax = 9ull * mean_q24;
    if (dmax > (uint64_t)(0.9 * 16777216.0)) dmax = (uint64_t)(0.9 * 16777216.0);
    const uint64_t base = ((u8 * dmax) >> 32) + (mean_q24 >> 4);
    gp->gk = gk;
    for (int64_t c = 0; c < C; ++c) {
        const uint64_t ck = mix64(gk ^ (SALT_B + (uint64_t)c * GOLD));
        uint64_t t = base * MULT[ck & 7] / 4;
        if (t > cap) t = cap;
        gp->thr[c] = all_zero ? 0u : (uint32_t)t;
        gp->pidx[c] = (uint8_t)((ck >> 3) & 3);
    }
}

static inline uint32_t code_at(const gene_params *gp, int64_t cell, int64_t c, int K,
                               const uint32_t *cum)
{
    const uint64_t h = mix64(gp->gk + (uint64_t)cell * M2);
    if ((uint32_t)(h >> 40) >= gp->thr[c]) return 0u;
    const uint32_t r16 = (uint32_t)(h >> 8) & 0xffffu;
    const uint32_t *cm = cum + gp->pidx[c] * (K - 2);
    uint32_t code = 1;
    for (int t = 0; t < K - 2; ++t) code += r16 >= cm[t];
    return code;
}

/* pass 1: stored entries of genes [gene_lo, gene_lo + P), or of the P listed genes when
 * gene_ids is not NULL */
int orc_synth_count(const int64_t *y, int64_t N, int64_t gene_lo, int64_t P, int64_t C, int64_t K,
                    uint32_t mean_q24, uint64_t seed, const uint32_t *cum, int64_t *nnz_per_gene,
                    int nthreads, const int64_t *gene_ids)
{
    if (C < 1 || C > 256 || K < 2 || K > 256) return 2;
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t g = 0; g < P; ++g) {
        gene_params gp;
        gene_setup(&gp, gene_ids ? gene_ids[g] : gene_lo + g, C, mean_q24, seed);
        int64_t n = 0;
        for (int64_t i = 0; i < N; ++i) n += code_at(&gp, i, y[i], (int)K, cum) != 0;
        nnz_per_gene[g] = n;
    }
    return 0;
}

/* pass 2: rows[indptr[g] ..] ascending cell ids, codes[..] their codes */
int orc_synth_fill(const int64_t *y, int64_t N, int64_t gene_lo, int64_t P, int64_t C, int64_t K,
                   uint32_t mean_q24, uint64_t seed, const uint32_t *cum, const int64_t *indptr,
                   int32_t *rows, uint8_t *codes, int nthreads, const int64_t *gene_ids)
{
    if (C < 1 || C > 256 || K < 2 || K > 256) return 2;
    (void)nthreads;
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads > 0 ? nthreads : 1)
    for (int64_t g = 0; g < P; ++g) {
        gene_params gp;
        gene_setup(&gp, gene_ids ? gene_ids[g] : gene_lo + g, C, mean_q24, seed);
        int64_t o = indptr[g];
        for (int64_t i = 0; i < N; ++i) {
            const uint32_t c = code_at(&gp, i, y[i], (int)K, cum);
            if (c) {
                rows[o] = (int32_t)i;
                codes[o] = (uint8_t)c;
                ++o;
            }
        }
    }
    return 0;
}
