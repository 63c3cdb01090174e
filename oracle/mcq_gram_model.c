/*
 * mcq_gram_model.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Bit-level CPU model of the arithmetic the CUDA search kernel performs
 * (quantization_b200/csrc/search.cu).  The product re-expresses one pass of the
 * reference's _refine_indexes (quantization.py:308-547) through two tables
 *
 *     P[b, r] = <x_b, c_r>            r = n*K + k     (one GEMM per frame)
 *     G[r, s] = <c_r, c_s>                            (once per parameter version)
 *
 * so that no per-frame delta vectors are ever formed.  In exact arithmetic this is
 * identical to the reference; in fp32 it differs from it only at near-ties.  The
 * model exists so that the GPU kernel can be checked BIT FOR BIT (same P, same G in,
 * same indexes out), which separates "kernel bug" from "fp32 near-tie".  The parity
 * statement against the reference itself is made with oracle/mcq_oracle.c.
 *
 * Only tests/ may load this library.
 *
 * Arithmetic contract shared with the kernel (all fp32, no contraction except the
 * explicit fmaf, which is exact doubling + one rounding):
 *
 *   level 1 (quantization.py:401-418, constants common to a codebook dropped):
 *     t[n,k]     = sum_{m != n, ascending m} G[row_m, n*K+k]       row_m = m*K + idx[m]
 *     v[n,k]     = fmaf(2, t[n,k] - P[n*K+k], G[n*K+k, n*K+k])
 *     delta[n,k] = v[n,k] - v[n, idx[n]]      ( = score - |x_err|^2 of the reference)
 *   reduce (:470-503): ascending by (delta, candidate position), keep the first newK.
 *   combine (:504-547), groups e = 2m, o = 2m+1 covering codebooks Ae, Ao:
 *     D_ab(p,q)  = ((G[a,p ; b,q] - G[a,p ; b,old_b]) - G[a,old_a ; b,q]) + G[a,old_a ; b,old_b]
 *     dot(i,j)   = sum_{b in Ao asc} ( sum_{a in Ae asc} D_ab(k_i[a], k_j[b]) )
 *     delta(i,j) = fmaf(2, dot(i,j), delta_e[i] + delta_o[j])        flat index i*Kcur + j
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GM_OK 0
#define GM_EINVAL -1
#define GM_EUNSUPPORTED -2
#define GM_ENOMEM -3

typedef struct {
    float v;
    int i;
} sv_t;

static int cmp_sv(const void *a, const void *b) {
    const sv_t *x = (const sv_t *)a, *y = (const sv_t *)b;
    if (x->v < y->v) return -1;
    if (x->v > y->v) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

static int k_cutoff(int base, int L) { /* quantization.py:455-463 */
    int c = base;
    while (L >= 4) {
        L /= 4;
        c *= 2;
    }
    return c < 128 ? c : 128;
}

static int is_pow2(long n) { return n > 0 && (n & (n - 1)) == 0; }

/* G = Cs Cs^T, P = X Cs^T: double accumulation, one rounding. */
int mcq_gm_gram(const float *Cs, int NK, int D, float *G, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(dynamic, 8)
    for (int r = 0; r < NK; ++r)
        for (int s = r; s < NK; ++s) {
            double acc = 0.0;
            const float *a = Cs + (size_t)r * D, *b = Cs + (size_t)s * D;
            for (int d = 0; d < D; ++d) acc += (double)a[d] * (double)b[d];
            G[(size_t)r * NK + s] = (float)acc;
            G[(size_t)s * NK + r] = (float)acc;
        }
    return GM_OK;
}

int mcq_gm_xct(const float *x, long B, int D, const float *Cs, int NK, float *P, int nthreads) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel for schedule(static)
    for (long b = 0; b < B; ++b)
        for (int r = 0; r < NK; ++r) {
            double acc = 0.0;
            const float *a = x + (size_t)b * D, *c = Cs + (size_t)r * D;
            for (int d = 0; d < D; ++d) acc += (double)a[d] * (double)c[d];
            P[(size_t)b * NK + r] = (float)acc;
        }
    return GM_OK;
}

typedef struct {
    float *sc, *sc2;
    int *ci, *ci2;
    sv_t *sv;
} gscratch_t;

static int plan(int N, int K, long *max_cand) {
    int base = K <= 16 ? 8 : 16, n = N, k = K, L = 1;
    long mc = (long)N * K;
    for (;;) {
        int cut = k_cutoff(base, L);
        if (n == 1 && k == 1) break;
        if (k > cut || n == 1) {
            k = n == 1 ? 1 : cut;
        } else {
            if (L == 1 && k == K) return GM_EUNSUPPORTED;
            n /= 2;
            k = k * k;
            L *= 2;
            if ((long)n * k > mc) mc = (long)n * k;
        }
    }
    *max_cand = mc;
    return GM_OK;
}

static void search_frame(const float *P, const float *G, int *idx, int N, int K, gscratch_t *s) {
    const size_t NK = (size_t)N * K;
    float *sc = s->sc, *sc2 = s->sc2;
    int *ci = s->ci, *ci2 = s->ci2;
    int old[64];
    for (int n = 0; n < N; ++n) old[n] = idx[n];
    /* level 1 */
    for (int n = 0; n < N; ++n) {
        for (int k = 0; k < K; ++k) {
            size_t col = (size_t)n * K + k;
            float t = 0.0f;
            for (int m = 0; m < N; ++m)
                if (m != n) t = t + G[((size_t)m * K + old[m]) * NK + col];
            float cross = t - P[col];
            sc[col] = fmaf(2.0f, cross, G[col * NK + col]);
            ci[col] = k;
        }
        float vold = sc[(size_t)n * K + old[n]];
        for (int k = 0; k < K; ++k) sc[(size_t)n * K + k] = sc[(size_t)n * K + k] - vold;
    }
    int Ncur = N, Kcur = K, L = 1;
    const int base = K <= 16 ? 8 : 16;
    for (;;) {
        int cut = k_cutoff(base, L);
        if (Ncur == 1 && Kcur == 1) {
            for (int l = 0; l < L; ++l) idx[l] = ci[l];
            return;
        }
        if (Kcur > cut || Ncur == 1) {
            int newK = Ncur == 1 ? 1 : cut;
            for (int n = 0; n < Ncur; ++n) {
                sv_t *sv = s->sv;
                for (int k = 0; k < Kcur; ++k) {
                    sv[k].v = sc[(size_t)n * Kcur + k];
                    sv[k].i = k;
                }
                qsort(sv, Kcur, sizeof(sv_t), cmp_sv);
                for (int j = 0; j < newK; ++j) {
                    sc2[(size_t)n * newK + j] = sv[j].v;
                    for (int l = 0; l < L; ++l)
                        ci2[((size_t)n * newK + j) * L + l] = ci[((size_t)n * Kcur + sv[j].i) * L + l];
                }
            }
            float *t = sc; sc = sc2; sc2 = t;
            int *ti = ci; ci = ci2; ci2 = ti;
            Kcur = newK;
        } else {
            int newN = Ncur / 2, newK = Kcur * Kcur, newL = 2 * L;
            for (int m = 0; m < newN; ++m) {
                int ae0 = (2 * m) * L, ao0 = (2 * m + 1) * L; /* first codebook of the even / odd group */
                for (int i = 0; i < Kcur; ++i) {
                    const int *ei = ci + ((size_t)(2 * m) * Kcur + i) * L;
                    for (int j = 0; j < Kcur; ++j) {
                        const int *oj = ci + ((size_t)(2 * m + 1) * Kcur + j) * L;
                        float dot = 0.0f;
                        for (int lb = 0; lb < L; ++lb) {
                            int b = ao0 + lb;
                            size_t cq = (size_t)b * K + oj[lb], co = (size_t)b * K + old[b];
                            float w = 0.0f;
                            for (int la = 0; la < L; ++la) {
                                int a = ae0 + la;
                                size_t rp = ((size_t)a * K + ei[la]) * NK, ro = ((size_t)a * K + old[a]) * NK;
                                float d = ((G[rp + cq] - G[rp + co]) - G[ro + cq]) + G[ro + co];
                                w = w + d;
                            }
                            dot = dot + w;
                        }
                        size_t flat = (size_t)m * newK + (size_t)i * Kcur + j;
                        float base_s = sc[(size_t)(2 * m) * Kcur + i] + sc[(size_t)(2 * m + 1) * Kcur + j];
                        sc2[flat] = fmaf(2.0f, dot, base_s);
                        int *dst = ci2 + flat * newL;
                        for (int l = 0; l < L; ++l) dst[l] = ei[l];
                        for (int l = 0; l < L; ++l) dst[L + l] = oj[l];
                    }
                }
            }
            float *t = sc; sc = sc2; sc2 = t;
            int *ti = ci; ci = ci2; ci2 = ti;
            Ncur = newN; Kcur = newK; L = newL;
        }
    }
}

/* iters passes of the search for every frame.  P (B, N*K), G (N*K, N*K), idx_in/idx_out (B,N) int64. */
int mcq_gm_search(const float *P, const float *G, long B, int N, int K, int iters, const int64_t *idx_in,
                  int64_t *idx_out, int nthreads) {
    if (B < 0 || !is_pow2(N) || !is_pow2(K) || N > 64 || iters < 0) return GM_EINVAL;
    if (N > 1 && K < 16) return GM_EUNSUPPORTED;
    long max_cand = 0;
    int rc = plan(N, K, &max_cand);
    if (rc) return rc;
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#else
    (void)nthreads;
#endif
#pragma omp parallel
    {
        gscratch_t s;
        s.sc = (float *)malloc(sizeof(float) * max_cand);
        s.sc2 = (float *)malloc(sizeof(float) * max_cand);
        s.ci = (int *)malloc(sizeof(int) * max_cand * N);
        s.ci2 = (int *)malloc(sizeof(int) * max_cand * N);
        s.sv = (sv_t *)malloc(sizeof(sv_t) * max_cand);
        if (!s.sc || !s.sc2 || !s.ci || !s.ci2 || !s.sv) {
#pragma omp atomic write
            err = GM_ENOMEM;
        } else {
#pragma omp for schedule(dynamic, 16)
            for (long b = 0; b < B; ++b) {
                int idx[64];
                for (int n = 0; n < N; ++n) idx[n] = (int)idx_in[(size_t)b * N + n];
                for (int it = 0; it < iters; ++it) search_frame(P + (size_t)b * N * K, G, idx, N, K, &s);
                for (int n = 0; n < N; ++n) idx_out[(size_t)b * N + n] = idx[n];
            }
        }
        free(s.sc); free(s.sc2); free(s.ci); free(s.ci2); free(s.sv);
    }
    return err;
}
