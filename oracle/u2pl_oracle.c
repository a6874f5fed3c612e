/*
 * oracle/u2pl_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C, scalar, single thread) of the integer/ordering
 * sensitive part of U2PL's per-step hot path:
 *
 *   - per-pixel softmax entropy          reference u2pl/utils/loss_helper.py:35-36
 *                                        and train_semi.py:402-403
 *   - np.percentile on float32 data      call sites loss_helper.py:38-40,
 *                                        train_semi.py:405-407,412-415; algorithm =
 *                                        numpy 2.3.5 lib/_function_base_impl.py
 *                                        (_quantile / _lerp, "linear" method)
 *   - reliable / unreliable partition    loss_helper.py:41-44, train_semi.py:408-418
 *   - per-pixel class rank window        loss_helper.py:91-97,127-134
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (u2pl_b200/) never does.
 *
 * ARITHMETIC CONTRACT.  Index sets can only be
 * bit-exact if the entropies they are cut from are bit-exact, and libm /
 * ATen / CUDA expf,logf all differ in the last ulp.  The framework therefore
 * fixes the arithmetic: exp and log are defined below purely in terms of
 * IEEE-754 binary32 round-to-nearest-even add, mul, fma, div and integer bit
 * operations, in a fixed order.  This file and u2pl_b200/csrc/arith.cuh are two
 * independent implementations of that same contract (compile this file with
 * -ffp-contract=off so the compiler never fuses on its own).
 * Against the reference's own libm-based entropies the contract differs by
 * <= ~2e-7 absolute (tests/test_oracle_golden.py pins that on fixtures made
 * by importing the reference, oracle/make_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

#define U2PL_MAGIC 12582912.0f            /* 1.5 * 2^23 */
#define U2PL_LOG2E 1.44269502162933349609375f
#define U2PL_LN2   0.693147182464599609375f

/* exp(d) for d <= 0.  d is clamped at -87 (exp(-87) ~ 1.6e-38, still normal). */
float u2pl_oracle_expf(float d)
{
    d = fmaxf(d, -87.0f);
    float t  = fmaf(d, U2PL_LOG2E, U2PL_MAGIC);   /* low mantissa bits hold k = rint(d*log2e) */
    float kf = t - U2PL_MAGIC;
    float r  = fmaf(kf, -U2PL_LN2, d);            /* |r| <= ln2/2 (+eps) */
    float q  = 0.0013933652080595493f;
    q = fmaf(q, r, 0.008363181725144386f);
    q = fmaf(q, r, 0.04166646674275398f);
    q = fmaf(q, r, 0.16666576266288757f);
    q = fmaf(q, r, 0.5f);
    float r2 = r * r;
    float p  = fmaf(r2, q, r);
    p = p + 1.0f;                                  /* e^r in [0.707, 1.415] */
    return u2f(f2u(p) + (f2u(t) << 23));           /* scale by 2^k via the exponent field */
}

/* log(y) for y a positive normal float (here y = prob + 1e-10 in [1e-10, 1]). */
float u2pl_oracle_logf(float y)
{
    uint32_t ix = f2u(y);
    int32_t  e  = (int32_t)(ix - 0x3f3504f3u) >> 23;      /* y = m * 2^e, m in [sqrt(.5), sqrt(2)) */
    float m  = u2f(ix - ((uint32_t)e << 23));
    float ef = u2f(0x4B400000u + (uint32_t)e) - U2PL_MAGIC;
    float f  = m - 1.0f;
    float R  = 0.08507229387760162f;
    R = fmaf(R, f, -0.14198024570941925f);
    R = fmaf(R, f, 0.1495114266872406f);
    R = fmaf(R, f, -0.16587895154953003f);
    R = fmaf(R, f, 0.1996057629585266f);
    R = fmaf(R, f, -0.2500097155570984f);
    R = fmaf(R, f, 0.33333972096443176f);
    float f2 = f * f;
    float u  = fmaf(f, R, -0.5f);
    float tt = f2 * u;
    float l  = f + tt;                                    /* log1p(f) = f - f^2/2 + f^3 R(f) */
    return fmaf(ef, U2PL_LN2, l);
}

/*
 * entropy[b, i] = - sum_c p_c * log(p_c + 1e-10),  p = softmax over the C axis
 * of logits[b, :, i] (NCHW contiguous, HW = pixels per image).
 * Restates loss_helper.py:35-36 under the arithmetic contract:
 *   m = max_c x_c; e_c = EXP(x_c - m); S = sum_c e_c (c ascending);
 *   rinv = 1/S (IEEE div); p_c = e_c*rinv; acc = fma(p_c, LOG(p_c + 1e-10), acc); ent = -acc.
 */
void u2pl_oracle_entropy(const float *logits, int64_t B, int64_t C, int64_t HW, float *ent)
{
    float *e = (float *)malloc(sizeof(float) * (size_t)C);
    for (int64_t b = 0; b < B; ++b) {
        for (int64_t i = 0; i < HW; ++i) {
            const float *x = logits + b * C * HW + i;
            float m = x[0];
            for (int64_t c = 1; c < C; ++c) m = fmaxf(m, x[c * HW]);
            float S = 0.0f;
            for (int64_t c = 0; c < C; ++c) { e[c] = u2pl_oracle_expf(x[c * HW] - m); S = S + e[c]; }
            float rinv = 1.0f / S;
            float acc = 0.0f;
            for (int64_t c = 0; c < C; ++c) {
                float p = e[c] * rinv;
                float l = u2pl_oracle_logf(p + 1e-10f);
                acc = fmaf(p, l, acc);
            }
            ent[b * HW + i] = -acc;
        }
    }
    free(e);
}

/* softmax probabilities under the same contract (used by the low-res prep oracle). */
void u2pl_oracle_softmax(const float *logits, int64_t B, int64_t C, int64_t HW, float *prob)
{
    for (int64_t b = 0; b < B; ++b) {
        for (int64_t i = 0; i < HW; ++i) {
            const float *x = logits + b * C * HW + i;
            float *p = prob + b * C * HW + i;
            float m = x[0];
            for (int64_t c = 1; c < C; ++c) m = fmaxf(m, x[c * HW]);
            float S = 0.0f;
            for (int64_t c = 0; c < C; ++c) { p[c * HW] = u2pl_oracle_expf(x[c * HW] - m); S = S + p[c * HW]; }
            float rinv = 1.0f / S;
            for (int64_t c = 0; c < C; ++c) p[c * HW] = p[c * HW] * rinv;
        }
    }
}

static int cmp_float(const void *a, const void *b)
{
    float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/*
 * np.percentile(vals (float32), q) with numpy 2.x semantics: the virtual index
 * is computed in float32 (q/float32(100), (n-1)*q32), "linear" method, two-sided
 * lerp.  Returns 0 on success, -1 if n == 0.  Also reports the two order-statistic
 * ranks that were used.  Sort-based (the reference uses np.partition; same values).
 */
int u2pl_oracle_percentile(const float *vals, int64_t n, float q, float *out,
                           int64_t *rank_lo, int64_t *rank_hi)
{
    if (n <= 0) return -1;
    float *s = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(s, vals, sizeof(float) * (size_t)n);
    qsort(s, (size_t)n, sizeof(float), cmp_float);
    float q32 = q / 100.0f;
    float nm1 = (float)(n - 1);
    float v   = nm1 * q32;
    int64_t lo, hi;
    if (v >= nm1) { lo = n - 1; hi = n - 1; }
    else { lo = (int64_t)floorf(v); hi = lo + 1; }
    float g = v - floorf(v);
    float a = s[lo], b = s[hi];
    float d = b - a;
    float r = a + d * g;
    if (g >= 0.5f) r = b - d * (1.0f - g);
    *out = r;
    if (rank_lo) *rank_lo = lo;
    if (rank_hi) *rank_hi = hi;
    free(s);
    return 0;
}

/*
 * rank of class `cls` in a descending stable sort of prob[0..C) with stride:
 * number of classes j with p_j > p_cls, or p_j == p_cls and j < cls.
 * Restates the membership test of loss_helper.py:91-97,127-134 without a sort.
 */
int u2pl_oracle_rank_of(const float *prob, int64_t C, int64_t stride, int64_t cls)
{
    float pc = prob[cls * stride];
    int r = 0;
    for (int64_t j = 0; j < C; ++j) {
        float pj = prob[j * stride];
        if (pj > pc || (pj == pc && j < cls)) ++r;
    }
    return r;
}
