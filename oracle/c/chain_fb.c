/* CPU restatement (plain C) of the forward-backward recursions of the hot path.
 * TEST / BASELINE INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * What it restates: Kaldi's chain denominator (scaled-probability leaky-HMM, float), chain
 * numerator (log domain, double) and lattice forward-backward (log domain, double), which the
 * reference reaches through PyKaldi at ops/ops.py:60 and ops/ops.py:265 (Kaldi sources are not
 * vendored in the reference: chain-denominator.cc, chain-numerator.cc, lattice-functions.cc).
 * Loop structure follows SURVEY.md Appendix B/C; one thread per utterance, OpenMP across utterances
 * (the reference calls Kaldi once per utterance, bin/train_chain.py:261-275).
 * Checked against the numpy oracle in tests/test_oracle_c.py.  Parity unpinned by the reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline double log_add(double a, double b) {
    if (a < b) { double t = a; a = b; b = t; }
    if (b == -INFINITY) return a;
    return a + log1p(exp(b - a));
}

/* Denominator forward-backward for one sequence (float arithmetic like Kaldi's CUDA/CPU path).
 * fwd CSR: off[S+1], prob[A], pdf[A], dst[A].  ll [T,N] row-major.  gamma [T,N] output (+= scale*gamma).
 * Returns log Z_den. */
double pk2o_den_fb(int S, int N, int T, const int32_t* off, const float* prob, const int32_t* pdf,
                   const int32_t* dst, const float* init, const float* ll, float leaky, float scale,
                   float* gamma) {
    float* alpha = (float*)malloc(sizeof(float) * (size_t)(T + 1) * S);
    float* asum = (float*)malloc(sizeof(float) * (size_t)(T + 1));
    float* e = (float*)malloc(sizeof(float) * (size_t)N);
    float* beta = (float*)malloc(sizeof(float) * (size_t)S);
    float* betan = (float*)malloc(sizeof(float) * (size_t)S);
    double logsum = 0.0;
    float tot = 0.f;
    for (int j = 0; j < S; ++j) { alpha[j] = init[j]; tot += init[j]; }
    asum[0] = tot;
    for (int j = 0; j < S; ++j) alpha[j] += leaky * tot * init[j];
    for (int t = 1; t <= T; ++t) {
        const float* row = ll + (size_t)(t - 1) * N;
        for (int p = 0; p < N; ++p) { float v = row[p]; v = v < -30.f ? -30.f : (v > 30.f ? 30.f : v); e[p] = expf(v); }
        float* a = alpha + (size_t)t * S;
        const float* ap = alpha + (size_t)(t - 1) * S;
        memset(a, 0, sizeof(float) * S);
        for (int i = 0; i < S; ++i) {
            const float ai = ap[i];
            for (int k = off[i]; k < off[i + 1]; ++k) a[dst[k]] += ai * prob[k] * e[pdf[k]];
        }
        const float inv = 1.0f / asum[t - 1];
        tot = 0.f;
        for (int j = 0; j < S; ++j) { a[j] *= inv; tot += a[j]; }
        asum[t] = tot;
        for (int j = 0; j < S; ++j) a[j] += leaky * tot * init[j];
        logsum += log((double)asum[t - 1]);
    }
    float totp = 0.f;
    for (int j = 0; j < S; ++j) totp += alpha[(size_t)T * S + j];
    const double logz = log((double)totp) + logsum;
    float isum = 0.f;
    for (int j = 0; j < S; ++j) isum += init[j];
    const float bT = (1.0f / totp) * (1.0f + leaky * isum);
    for (int j = 0; j < S; ++j) beta[j] = bT;
    for (int t = T - 1; t >= 0; --t) {
        const float* row = ll + (size_t)t * N;
        for (int p = 0; p < N; ++p) { float v = row[p]; v = v < -30.f ? -30.f : (v > 30.f ? 30.f : v); e[p] = expf(v); }
        const float* a = alpha + (size_t)t * S;
        const float inv = 1.0f / asum[t];
        float* g = gamma + (size_t)t * N;
        float dot = 0.f;
        for (int i = 0; i < S; ++i) {
            float acc = 0.f;
            const float ai = a[i];
            for (int k = off[i]; k < off[i + 1]; ++k) {
                const float x = prob[k] * e[pdf[k]] * beta[dst[k]] * inv;
                acc += x;
                g[pdf[k]] += scale * ai * x;
            }
            betan[i] = acc;
            dot += acc * init[i];
        }
        for (int i = 0; i < S; ++i) beta[i] = betan[i] + leaky * dot;
    }
    free(alpha); free(asum); free(e); free(beta); free(betan);
    return logz;
}

/* Log-domain forward-backward over an epsilon-free, topologically sorted FST with time-stamped
 * states (numerator graph).  Arcs sorted by src.  gamma[t,pdf] += scale * posterior.  Returns log Z. */
double pk2o_num_fb(int S, int A, int N, int T, int start, const int32_t* src, const int32_t* dst,
                   const int32_t* pdf, const float* w, const float* final_cost, const int32_t* times,
                   const float* ll, float scale, float* gamma) {
    double* la = (double*)malloc(sizeof(double) * S);
    double* lb = (double*)malloc(sizeof(double) * S);
    for (int s = 0; s < S; ++s) la[s] = -INFINITY;
    la[start] = 0.0;
    for (int k = 0; k < A; ++k)
        la[dst[k]] = log_add(la[dst[k]], la[src[k]] + (double)ll[(size_t)times[src[k]] * N + pdf[k]] - (double)w[k]);
    double z = -INFINITY;
    for (int s = 0; s < S; ++s) {
        lb[s] = isfinite(final_cost[s]) ? -(double)final_cost[s] : -INFINITY;
        if (isfinite(final_cost[s])) z = log_add(z, la[s] + lb[s]);
    }
    for (int k = A - 1; k >= 0; --k) {
        const int t = times[src[k]];
        const double sc = (double)ll[(size_t)t * N + pdf[k]] - (double)w[k] + lb[dst[k]];
        lb[src[k]] = log_add(lb[src[k]], sc);
        gamma[(size_t)t * N + pdf[k]] += scale * (float)exp(la[src[k]] + sc - z);
    }
    (void)T;
    free(la); free(lb);
    return z;
}

/* Lattice forward-backward + MMI merge.  Arcs sorted by src (topological), tid 0 =
 * epsilon.  post [T,N] output = (num - den) merged posteriors, dropped frames zero.  Returns total like. */
double pk2o_lat_mmi(int S, int A, int N, int T, const int32_t* src, const int32_t* dst, const int32_t* tid,
                    const float* gc, const float* final_cost, const int32_t* times, const int32_t* tid2pdf,
                    const int32_t* num_ali, const uint8_t* keep, const float* ll, float lm, float ac, float* post) {
    double* alpha = (double*)malloc(sizeof(double) * S);
    double* beta = (double*)malloc(sizeof(double) * S);
    for (int s = 0; s < S; ++s) alpha[s] = -INFINITY;
    alpha[0] = 0.0;
#define ARC_LIKE(k) (-(double)(lm * gc[k]) + (tid[k] ? (double)ac * (double)ll[(size_t)times[src[k]] * N + tid2pdf[tid[k]]] : 0.0))
    for (int k = 0; k < A; ++k) alpha[dst[k]] = log_add(alpha[dst[k]], alpha[src[k]] + ARC_LIKE(k));
    double tot = -INFINITY;
    for (int s = 0; s < S; ++s) {
        beta[s] = isfinite(final_cost[s]) ? -(double)(lm * final_cost[s]) : -INFINITY;
        if (isfinite(final_cost[s])) tot = log_add(tot, alpha[s] + beta[s]);
    }
    memset(post, 0, sizeof(float) * (size_t)T * N);
    for (int k = A - 1; k >= 0; --k) {
        const double ab = beta[dst[k]] + ARC_LIKE(k);
        beta[src[k]] = log_add(beta[src[k]], ab);
        if (tid[k]) {
            const int t = times[src[k]];
            if (keep[t]) post[(size_t)t * N + tid2pdf[tid[k]]] -= (float)exp(alpha[src[k]] + ab - tot);
        }
    }
#undef ARC_LIKE
    for (int t = 0; t < T; ++t)
        if (keep[t]) post[(size_t)t * N + tid2pdf[num_ali[t]]] += 1.0f;
    free(alpha); free(beta);
    return tot;
}

/* Batch drivers: OpenMP across utterances. */
void pk2o_chain_batch(int n_seq, int S, int N, const int32_t* off, const float* prob, const int32_t* pdf,
                      const int32_t* dst, const float* init, const float* const* ll, const int32_t* T,
                      float leaky, float* const* grad, double* logz_den) {
#pragma omp parallel for schedule(dynamic)
    for (int b = 0; b < n_seq; ++b)
        logz_den[b] = pk2o_den_fb(S, N, T[b], off, prob, pdf, dst, init, ll[b], leaky, 1.0f, grad[b]);
}
