/*
 * pb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the numeric kernels on Poreplex's per-read raw-signal
 * hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path
 * (poreplex_b200/) never does.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the hyeshik/poreplex tree).  Two of the reference's numeric dependencies are
 * third-party packages that are NOT vendored in that tree and cannot be
 * installed here:
 *
 *   pomegranate >= 0.10 (setup.py:74)   HMM Viterbi  -> orc_viterbi()
 *   tensorflow  >= 1.8  (setup.py:79)   LSTM predict -> orc_lstm_*(), orc_*_predict()
 *
 * Their published algorithms are restated here (Keras 2.2.4 LSTM / LSTMCell cell
 * equations, Eigen's float tanh / logistic / exp kernels that TF's CPU backend
 * uses, pomegranate's log-space Viterbi and pair_lse).  PARITY FOR THOSE TWO IS
 * UNPINNED against the real packages (no golden vectors exist anywhere in the
 * reference: tests/test_commandline.py is an empty test).  What IS pinned: every
 * piece of the reference's own Python runs verbatim on top of these kernels
 * (oracle/refshim, tests/golden/make_golden.py) and the event detector is
 * checked against the reference's own C compiled into oracle/_ref.
 *
 * Arithmetic contract (shared with the CUDA kernels, see DESIGN.md "Numerics"):
 *   - compiled with -ffp-contract=off; a fused multiply-add happens only where
 *     fmaf()/fma() is written;
 *   - all transcendental functions are built from + - * / fma and bit casts, so
 *     that a GPU can reproduce them bit for bit.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "pb_oracle.h"

/* ------------------------------------------------------------------------- */
/* bit casts                                                                 */
/* ------------------------------------------------------------------------- */
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

/* ------------------------------------------------------------------------- */
/* f32 activation functions: Eigen's packet kernels as used by TF's CPU ops   */
/* ------------------------------------------------------------------------- */

/* Eigen/src/Core/MathFunctionsImpl.h generic_fast_tanh_float (3.3.x): clamp to
 * [-9, 9], odd/even rational 13/6 polynomial, pmadd Horner, one division. */
static inline float clampf(float x, float lo, float hi)
{
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}

static inline float tanhf_eigen(float a)
{
    float x = clampf(a, -9.0f, 9.0f);
    const float x2 = x * x;
    float p = fmaf(x2, -2.76076847742355e-16f, 2.00018790482477e-13f);
    p = fmaf(x2, p, -8.60467152213735e-11f);
    p = fmaf(x2, p, 5.12229709037114e-08f);
    p = fmaf(x2, p, 1.48572235717979e-05f);
    p = fmaf(x2, p, 6.37261928875436e-04f);
    p = fmaf(x2, p, 4.89352455891786e-03f);
    p = x * p;
    float q = fmaf(x2, 1.19825839466702e-06f, 1.18534705686654e-04f);
    q = fmaf(x2, q, 2.26843463243900e-03f);
    q = fmaf(x2, q, 4.89352518554385e-03f);
    return p / q;
}

float orc_tanhf(float a) { return tanhf_eigen(a); }

/* Eigen/src/Core/functors/UnaryFunctors.h scalar_logistic_op<float>::packetOp:
 * clamp to [-18, 18], rational 9/10 polynomial, + 0.5, clamp to [0, 1]. */
static inline float sigmoidf_eigen(float a)
{
    float x = clampf(a, -18.0f, 18.0f);
    const float x2 = x * x;
    float p = fmaf(x2, 4.37031012579801e-11f, 1.15627324459942e-07f);
    p = fmaf(x2, p, 6.08574864600143e-05f);
    p = fmaf(x2, p, 8.51377133304701e-03f);
    p = fmaf(x2, p, 2.48287947061529e-01f);
    p = x * p;
    float q = fmaf(x2, 6.10247389755681e-13f, 5.76102136993427e-09f);
    q = fmaf(x2, q, 6.29106785017040e-06f);
    q = fmaf(x2, q, 1.70198817374094e-03f);
    q = fmaf(x2, q, 1.16817656904453e-01f);
    q = fmaf(x2, q, 9.93151921023180e-01f);
    float r = p / q + 0.5f;
    return clampf(r, 0.0f, 1.0f);
}

float orc_sigmoidf(float a) { return sigmoidf_eigen(a); }

/* Eigen pexp<float> (Cephes expf): range reduction by ln2 split in two,
 * degree-5 polynomial, scale by 2^n through the exponent field. */
float orc_expf(float a)
{
    float x = clampf(a, -88.3762626647949f, 88.3762626647950f);
    float fx = floorf(fmaf(x, 1.44269504088896341f, 0.5f));
    float r = fmaf(-fx, 0.693359375f, x);
    r = fmaf(-fx, -2.12194440e-4f, r);
    const float z = r * r;
    float y = 1.9875691500E-4f;
    y = fmaf(y, r, 1.3981999507E-3f);
    y = fmaf(y, r, 8.3334519073E-3f);
    y = fmaf(y, r, 4.1665795894E-2f);
    y = fmaf(y, r, 1.6666665459E-1f);
    y = fmaf(y, r, 5.0000001201E-1f);
    y = fmaf(y, z, r);
    y = y + 1.0f;
    int n = (int)fx + 127;
    if (n < 0) n = 0;
    if (n > 254) n = 254;
    return y * u2f((uint32_t)n << 23);
}

/* ------------------------------------------------------------------------- */
/* f64 exp / log for pair_lse (pomegranate calls libc cexp/clog; libm is not   */
/* bit-reproducible on a GPU, so both sides use this fixed recipe, |err|<2ulp) */
/* ------------------------------------------------------------------------- */

/* exp(x) for x in [-745, 0]; x < -40 returns 0 (then exp(x)+1 == 1 exactly,
 * which is all pair_lse needs). */
double orc_exp_neg(double x)
{
    if (!(x >= -40.0)) return 0.0;
    double k = floor(x * 1.4426950408889634074 + 0.5);
    double r = fma(-k, 6.93147180369123816490e-01, x);
    r = fma(-k, 1.90821492927058770002e-10, r);
    /* Taylor, degree 13, |r| <= 0.3466 */
    double p = 1.0 / 6227020800.0;
    p = fma(p, r, 1.0 / 479001600.0);
    p = fma(p, r, 1.0 / 39916800.0);
    p = fma(p, r, 1.0 / 3628800.0);
    p = fma(p, r, 1.0 / 362880.0);
    p = fma(p, r, 1.0 / 40320.0);
    p = fma(p, r, 1.0 / 5040.0);
    p = fma(p, r, 1.0 / 720.0);
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    int64_t ki = (int64_t)k;
    return u2d(d2u(p) + ((uint64_t)ki << 52));
}

/* log(w) for w in [1, 2]: fold to [sqrt(1/2), sqrt(2)], atanh series. */
double orc_log_1to2(double w)
{
    double e = 0.0;
    if (w > 1.4142135623730951) { w = w * 0.5; e = 1.0; }
    const double f = w - 1.0;
    const double s = f / (2.0 + f);
    const double z = s * s;
    double p = 1.0 / 23.0;
    p = fma(p, z, 1.0 / 21.0);
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    /* log(m) = 2s + 2s*z*p */
    const double t = (s + s);
    double r = fma(t * z, p, t);
    return fma(e, 6.93147180559945286227e-01, r);
}

/* pomegranate utils.pyx pair_lse(x, y): +inf / -inf short cuts, then
 * max + log(exp(min - max) + 1). */
double orc_pair_lse(double x, double y)
{
    if (x == INFINITY || y == INFINITY) return INFINITY;
    if (x == -INFINITY) return y;
    if (y == -INFINITY) return x;
    if (x > y) return x + orc_log_1to2(orc_exp_neg(y - x) + 1.0);
    return y + orc_log_1to2(orc_exp_neg(x - y) + 1.0);
}

/* ------------------------------------------------------------------------- */
/* A1  int16 DAC -> pA      fast5_file.py:122-131                             */
/*     np.array(range / digitisation * (raw + offset), dtype=float32):        */
/*     gain is one Python-float division, the product is fp64, one cast.      */
/* ------------------------------------------------------------------------- */
void orc_dac_to_pa(const int16_t *raw, int64_t n, double gain, double offset, float *out)
{
    for (int64_t i = 0; i < n; i++)
        out[i] = (float)(gain * ((double)raw[i] + offset));
}

/* ------------------------------------------------------------------------- */
/* A2/A4  mean-pool by `stride`   signal_loader.py:224-225, 244-247           */
/*     reshape(-1, stride).mean(axis=1, dtype=float32): numpy's pairwise-sum   */
/*     inner loop (8 accumulators, tail added sequentially) then / f32(stride) */
/* ------------------------------------------------------------------------- */
static float np_pairwise_sum_f32(const float *a, int n)
{
    if (n < 8) {
        float res = 0.0f;
        for (int i = 0; i < n; i++) res += a[i];
        return res;
    }
    if (n <= 128) {
        float r[8];
        int i;
        for (i = 0; i < 8; i++) r[i] = a[i];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
}

void orc_pool_mean(const float *x, int64_t npooled, int stride, float *out)
{
    const float div = (float)stride;
    for (int64_t i = 0; i < npooled; i++)
        out[i] = np_pairwise_sum_f32(x + i * stride, stride) / div;
}

/* A4  np.poly1d(f32[scale, shift])(x): Horner in f32, unfused.
 *     signal_loader.py:258-262 */
void orc_scale(const float *x, int64_t n, float scale, float shift, float *out)
{
    for (int64_t i = 0; i < n; i++) {
        float y = scale * x[i];
        out[i] = y + shift;
    }
}

/* ------------------------------------------------------------------------- */
/* LSTM (Keras 2.2.4-tf)                                                      */
/*   gate column order [i | f | c | o]; h = c = 0 initially; no masking.      */
/*   K.dot() is restated as an fmaf chain from 0 with k ascending.            */
/*   implementation=1 (LSTM layers of the scaler):                            */
/*       z = ((x.W + b) + h.U)                  recurrent.py LSTMCell.call     */
/*   implementation=2 (LSTMCell of the demultiplexer):                        */
/*       z = ((x.W + h.U) + b)                                                */
/* ------------------------------------------------------------------------- */
static void dot_chain(const float *restrict v, int K, const float *restrict M, int ncol,
                      float *restrict acc)
{
    /* acc[j] = fma chain over k ascending, from 0.  Column-blocked so that the 64
     * running sums stay in vector registers; the order of operations per column
     * is unchanged. */
    int j0 = 0;
    for (; j0 + 64 <= ncol; j0 += 64) {
        float a[64];
        for (int j = 0; j < 64; j++) a[j] = 0.0f;
        for (int k = 0; k < K; k++) {
            const float vk = v[k];
            const float *row = M + (size_t)k * ncol + j0;
            for (int j = 0; j < 64; j++) a[j] = fmaf(vk, row[j], a[j]);
        }
        for (int j = 0; j < 64; j++) acc[j0 + j] = a[j];
    }
    for (int j = j0; j < ncol; j++) {
        float a = 0.0f;
        for (int k = 0; k < K; k++) a = fmaf(v[k], M[(size_t)k * ncol + j], a);
        acc[j] = a;
    }
}

void orc_lstm_step(const orc_lstm *L, const float *x, float *h, float *c)
{
    const int H = L->units, G = 4 * L->units;
    float xw[ORC_MAX_GATES], hu[ORC_MAX_GATES], z[ORC_MAX_GATES];
    if (L->in_dim == 1) {
        for (int j = 0; j < G; j++) xw[j] = x[0] * L->W[j];
    } else {
        dot_chain(x, L->in_dim, L->W, G, xw);
    }
    dot_chain(h, H, L->U, G, hu);
    if (L->impl == 1) {
        for (int j = 0; j < G; j++) z[j] = (xw[j] + L->b[j]) + hu[j];
    } else {
        for (int j = 0; j < G; j++) z[j] = (xw[j] + hu[j]) + L->b[j];
    }
    /* gate activations (loops kept branch-free so that gcc vectorises them) */
    float ig[ORC_MAX_UNITS], fg[ORC_MAX_UNITS], cg[ORC_MAX_UNITS], og[ORC_MAX_UNITS];
    for (int u = 0; u < H; u++) ig[u] = sigmoidf_eigen(z[u]);
    for (int u = 0; u < H; u++) fg[u] = sigmoidf_eigen(z[H + u]);
    for (int u = 0; u < H; u++) cg[u] = tanhf_eigen(z[2 * H + u]);
    for (int u = 0; u < H; u++) og[u] = sigmoidf_eigen(z[3 * H + u]);
    for (int u = 0; u < H; u++) {
        const float a = fg[u] * c[u];
        const float b = ig[u] * cg[u];
        c[u] = a + b;                             /* two products, one add, unfused */
    }
    for (int u = 0; u < H; u++) h[u] = og[u] * tanhf_eigen(c[u]);
}

/* Run a layer over x[T][in_dim]; reverse != 0 consumes t = T-1 .. 0 and stores
 * its outputs back at position t (Bidirectional re-reverses the backward
 * outputs before concat).  hseq may be NULL. */
void orc_lstm_seq(const orc_lstm *L, const float *x, int T, int reverse,
                  float *hseq, float *hlast)
{
    float h[ORC_MAX_UNITS], c[ORC_MAX_UNITS];
    memset(h, 0, sizeof h);
    memset(c, 0, sizeof c);
    for (int s = 0; s < T; s++) {
        const int t = reverse ? T - 1 - s : s;
        orc_lstm_step(L, x + (size_t)t * L->in_dim, h, c);
        if (hseq) memcpy(hseq + (size_t)t * L->units, h, sizeof(float) * L->units);
    }
    if (hlast) memcpy(hlast, h, sizeof(float) * L->units);
}

static void dense(const float *h, int K, const float *W, const float *b, int n, float *out)
{
    dot_chain(h, K, W, n, out);
    for (int j = 0; j < n; j++) out[j] = out[j] + b[j];
}

/* A3  scaler network: LSTM(48, seq) -> LSTM(48, last) -> Dense(2).
 *     signal_loader.py:96-97; architecture from scaler-r3.hdf5 model_config. */
void orc_scaler_predict(const orc_scaler *S, const float *head, int T, float z[2])
{
    float h1[ORC_MAX_UNITS] = {0}, c1[ORC_MAX_UNITS] = {0};
    float h2[ORC_MAX_UNITS] = {0}, c2[ORC_MAX_UNITS] = {0};
    for (int t = 0; t < T; t++) {
        orc_lstm_step(&S->l1, head + t, h1, c1);
        orc_lstm_step(&S->l2, h1, h2, c2);
    }
    dense(h2, S->l2.units, S->Wd, S->bd, 2, z);
}

/* A7  demux network: Bidirectional(LSTMCell 48, concat) -> LSTMCell 64 ->
 *     Dense(5) + softmax.  barcoding.py:106-107; demux-tetra-r4.hdf5. */
void orc_demux_predict(const orc_demux *D, const float *win, int T, float *probs)
{
    const int H1 = D->fwd.units;
    float *hcat = (float *)malloc(sizeof(float) * (size_t)T * 2 * H1);
    float *tmp = (float *)malloc(sizeof(float) * (size_t)T * H1);
    orc_lstm_seq(&D->fwd, win, T, 0, tmp, NULL);
    for (int t = 0; t < T; t++) memcpy(hcat + (size_t)t * 2 * H1, tmp + (size_t)t * H1, sizeof(float) * H1);
    orc_lstm_seq(&D->bwd, win, T, 1, tmp, NULL);
    for (int t = 0; t < T; t++) memcpy(hcat + (size_t)t * 2 * H1 + H1, tmp + (size_t)t * H1, sizeof(float) * H1);
    float hl[ORC_MAX_UNITS];
    orc_lstm_seq(&D->l2, hcat, T, 0, NULL, hl);
    float logit[ORC_MAX_CLASSES], e[ORC_MAX_CLASSES];
    const int n = D->n_classes;
    dense(hl, D->l2.units, D->Wd, D->bd, n, logit);
    /* softmax: exp(x - max), sequential sum, multiply by the reciprocal */
    float m = logit[0];
    for (int j = 1; j < n; j++) m = fmaxf(m, logit[j]);
    float s = 0.0f;
    for (int j = 0; j < n; j++) { e[j] = orc_expf(logit[j] - m); s += e[j]; }
    const float rs = 1.0f / s;
    for (int j = 0; j < n; j++) probs[j] = e[j] * rs;
    free(tmp);
    free(hcat);
}

/* ------------------------------------------------------------------------- */
/* A5  HMM Viterbi   (pomegranate hmm.pyx _viterbi, restated; Appendix D of    */
/*     SURVEY.md).  States are given in baked order (sorted by name).         */
/*     signal_analyzer.py:352,389 ; worker_persistence.py:95-121              */
/* ------------------------------------------------------------------------- */
void orc_hmm_emissions(const orc_hmm *M, double x, double *e)
{
    for (int s = 0; s < M->n_states; s++) {
        if (M->n_comp[s] == 1) {
            /* NormalDistribution._log_probability */
            const double d = x - M->mu[s][0];
            e[s] = M->lsp[s][0] - (d * d) * M->inv2s2[s][0];
        } else {
            /* GeneralMixtureModel._log_probability */
            double acc = -INFINITY;
            for (int j = 0; j < M->n_comp[s]; j++) {
                const double d = x - M->mu[s][j];
                const double lp = M->lsp[s][j] - (d * d) * M->inv2s2[s][j];
                acc = orc_pair_lse(acc, lp + M->logw[s][j]);
            }
            e[s] = acc;
        }
    }
}

double orc_viterbi(const orc_hmm *M, const float *x, int T, int32_t *path)
{
    const int S = M->n_states;
    if (T <= 0) return -INFINITY;
    uint8_t *bp = (uint8_t *)malloc((size_t)T * S);
    double v[ORC_MAX_STATES] = {0}, nv[ORC_MAX_STATES], e[ORC_MAX_STATES];
    /* t = 0: only edges out of the silent start state carry probability */
    orc_hmm_emissions(M, (double)x[0], e);
    for (int s = 0; s < S; s++) {
        v[s] = -INFINITY;
        if (M->log_start[s] > -INFINITY) {
            const double cand = (0.0 + M->log_start[s]) + e[s];
            if (cand > v[s]) v[s] = cand;
        }
        bp[s] = 0xFF;
    }
    for (int t = 1; t < T; t++) {
        orc_hmm_emissions(M, (double)x[t], e);
        for (int l = 0; l < S; l++) {
            double best = -INFINITY;
            int arg = 0xFF;
            for (int k = M->in_begin[l]; k < M->in_begin[l + 1]; k++) {
                const double cand = (v[M->in_src[k]] + M->in_logp[k]) + e[l];
                if (cand > best) { best = cand; arg = M->in_src[k]; }
            }
            nv[l] = best;
            bp[(size_t)t * S + l] = (uint8_t)arg;
        }
        memcpy(v, nv, sizeof(double) * S);
    }
    int end = 0;
    double best = v[0];
    for (int s = 1; s < S; s++) if (v[s] > best) { best = v[s]; end = s; }
    if (best == -INFINITY) { free(bp); return -INFINITY; }
    int cur = end;
    for (int t = T - 1; t >= 0; t--) {
        path[t] = cur;
        if (t > 0) cur = bp[(size_t)t * S + cur];
    }
    free(bp);
    return best;
}

/* signal_analyzer.py:355-362: run-length groups of the state path; a later run of
 * the same state overwrites an earlier one (dict assignment). seg[s] = (first,
 * last) inclusive, (-1, -1) when the state never occurs. */
void orc_segments_from_path(const int32_t *path, int T, int n_states, int32_t *seg)
{
    for (int s = 0; s < n_states; s++) seg[2 * s] = seg[2 * s + 1] = -1;
    int t = 0;
    while (t < T) {
        int s = path[t], first = t;
        while (t + 1 < T && path[t + 1] == s) t++;
        seg[2 * s] = first;
        seg[2 * s + 1] = t;
        t++;
    }
}

/* ------------------------------------------------------------------------- */
/* A6  barcode window   barcoding.py:77-101                                    */
/* ------------------------------------------------------------------------- */
static int cmp_f32(const void *a, const void *b)
{
    const float x = *(const float *)a, y = *(const float *)b;
    return (x > y) - (x < y);
}

/* np.median of f32: middle element, or mean of the two middle ones computed as
 * np.mean does it: f32 add, then / 2 in f32. */
static float np_median_f32(const float *x, int n, float *scratch)
{
    if (n <= 0) return NAN;
    memcpy(scratch, x, sizeof(float) * n);
    qsort(scratch, n, sizeof(float), cmp_f32);
    if (n & 1) return scratch[n / 2];
    return (scratch[n / 2 - 1] + scratch[n / 2]) / 2.0f;
}

/* returns 1 and fills out[trimlen] if the adapter is accepted, else 0 */
int orc_barcode_window(const float *adapter, int len, int minlen, int maxlen,
                       int trimlen, float pad, float *out)
{
    if (!(minlen <= len && len <= maxlen) || len <= 0) return 0;
    const float *sig = adapter;
    int n = len;
    if (len > trimlen) { sig = adapter + (len - trimlen); n = trimlen; }
    float *tmp = (float *)malloc(sizeof(float) * 2 * n);
    float *dev = tmp + n;
    const float med = np_median_f32(sig, n, tmp);
    for (int i = 0; i < n; i++) dev[i] = fabsf(sig[i] - med);
    float *tmp2 = (float *)malloc(sizeof(float) * n);
    const float mad = np_median_f32(dev, n, tmp2);
    free(tmp2);
    /* max(0.01, mad * 1.4826): f32 multiply (NEP 50), compare against f32(0.01) */
    const float scaled = mad * 1.4826f;
    const float div = (scaled > 0.01f) ? scaled : 0.01f;
    const int npad = trimlen - n;
    for (int i = 0; i < npad; i++) out[i] = pad;
    for (int i = 0; i < n; i++) out[npad + i] = (sig[i] - med) / div;
    free(tmp);
    return 1;
}

/* barcoding.py:72-75,108-118: argmax - decoys, threshold in fp64, bisect_right */
void orc_barcode_decide(const float *probs, int n_classes, int n_decoy,
                        const double *calib, int n_calib, double score_threshold,
                        int32_t *barcode, int32_t *guess, int32_t *phred)
{
    int arg = 0;
    float score = probs[0];
    for (int j = 1; j < n_classes; j++) if (probs[j] > score) { score = probs[j]; arg = j; }
    const int bcid = arg - n_decoy;
    const double sc = (double)score;
    *barcode = (bcid >= 0 && sc >= score_threshold) ? bcid : -1;
    *guess = bcid;
    if (sc <= 0.0) { *phred = 0; return; }
    int lo = 0, hi = n_calib;          /* bisect_right */
    while (lo < hi) {
        const int mid = (lo + hi) / 2;
        if (sc < calib[mid]) hi = mid; else lo = mid + 1;
    }
    *phred = lo;
}

/* ------------------------------------------------------------------------- */
/* A9  event detector  src/contrib/scrappie/event_detection.c (restated; the   */
/*     reference's own C is also compiled into oracle/_ref and compared).      */
/* ------------------------------------------------------------------------- */
static void ev_tstat(const double *sum, const double *sumsq, int64_t n, int64_t w, float *t)
{
    /* event_detection.c:61-117 */
    const float eta = FLT_MIN;
    const float wf = (float)w;
    for (int64_t i = 0; i < n; i++) t[i] = 0.0f;
    if (n < 2 * w || w < 2) return;
    for (int64_t i = w; i <= n - w; i++) {
        double sum1 = sum[i], sumsq1 = sumsq[i];
        if (i > w) { sum1 -= sum[i - w]; sumsq1 -= sumsq[i - w]; }
        const float sum2 = (float)(sum[i + w] - sum[i]);
        const float sumsq2 = (float)(sumsq[i + w] - sumsq[i]);
        const float mean1 = (float)(sum1 / wf);
        const float mean2 = sum2 / wf;
        float cv = (float)(sumsq1 / wf - (double)(mean1 * mean1)
                           + (double)(sumsq2 / wf) - (double)(mean2 * mean2));
        cv = fmaxf(cv, eta);
        const float dm = mean2 - mean1;
        t[i] = (float)(fabs((double)dm) / sqrt((double)(cv / wf)));
    }
    /* the reference's "fudge boundaries" loop zeroes [0, w) and [n - w, n) and the
     * main loop then overwrites i = n - w; with t[] pre-zeroed the net effect
     * is identical: computed values on [w, n - w], zero elsewhere. */
}

typedef struct {
    float threshold; int64_t window; int64_t masked_to; int64_t peak_pos;
    float peak_value; int valid;
} ev_det;

int64_t orc_detect_events(const float *x, int64_t n, int64_t w1, int64_t w2,
                          float thr1, float thr2, float peak_height,
                          orc_event *ev, int64_t max_events)
{
    if (n <= 0) return 0;
    double *sum = (double *)calloc((size_t)n + 1, sizeof(double));
    double *sumsq = (double *)calloc((size_t)n + 1, sizeof(double));
    float *t1 = (float *)calloc((size_t)n, sizeof(float));
    float *t2 = (float *)calloc((size_t)n, sizeof(float));
    int64_t *peaks = (int64_t *)calloc((size_t)n, sizeof(int64_t));
    /* event_detection.c:35-49: data[i]*data[i] is an f32 product */
    for (int64_t i = 0; i < n; i++) {
        sum[i + 1] = sum[i] + x[i];
        sumsq[i + 1] = sumsq[i] + (double)(x[i] * x[i]);
    }
    ev_tstat(sum, sumsq, n, w1, t1);
    ev_tstat(sum, sumsq, n, w2, t2);
    /* event_detection.c:124-201 */
    ev_det det[2] = {
        { thr1, w1, 0, -1, FLT_MAX, 0 },
        { thr2, w2, 0, -1, FLT_MAX, 0 } };
    const float *sig[2] = { t1, t2 };
    int64_t npk = 0;
    for (int64_t i = 0; i < n; i++) {
        for (int k = 0; k < 2; k++) {
            ev_det *d = &det[k];
            if (d->masked_to >= i) continue;
            const float cur = sig[k][i];
            if (d->peak_pos == -1) {
                if (cur < d->peak_value) d->peak_value = cur;
                else if (cur - d->peak_value > peak_height) { d->peak_value = cur; d->peak_pos = i; }
            } else {
                if (cur > d->peak_value) { d->peak_value = cur; d->peak_pos = i; }
                if (k == 0 && d->peak_value > d->threshold) {
                    det[1].masked_to = d->peak_pos + d->window;
                    det[1].peak_pos = -1;
                    det[1].peak_value = FLT_MAX;
                    det[1].valid = 0;
                }
                if (d->peak_value - cur > peak_height && d->peak_value > d->threshold) d->valid = 1;
                if (d->valid && (i - d->peak_pos) > d->window / 2) {
                    peaks[npk++] = d->peak_pos;
                    d->peak_pos = -1;
                    d->peak_value = cur;
                    d->valid = 0;
                }
            }
        }
    }
    /* event_detection.c:216-271; size_t arithmetic kept (end - start wraps if
     * the two detectors emitted peaks out of position order) */
    int64_t ne = 1;
    for (int64_t i = 0; i < n; i++) if (peaks[i] > 0 && peaks[i] < n) ne++;
    int64_t nout = ne < max_events ? ne : max_events;
    for (int64_t e = 0; e < nout; e++) {
        uint64_t s, en;
        if (e == 0) { s = 0; en = (uint64_t)peaks[0]; }
        else if (e < ne - 1) { s = (uint64_t)peaks[e - 1]; en = (uint64_t)peaks[e]; }
        else { s = (uint64_t)peaks[ne - 2]; en = (uint64_t)n; }
        orc_event *o = &ev[e];
        o->start = s;
        o->length = (float)(uint64_t)(en - s);
        o->mean = (float)(sum[en] - sum[s]) / o->length;
        const float deltasqr = (float)(sumsq[en] - sumsq[s]);
        const float var = deltasqr / o->length - o->mean * o->mean;
        o->stdv = sqrtf(fmaxf(var, 0.0f));
    }
    free(peaks); free(t2); free(t1); free(sumsq); free(sum);
    return ne;
}

/* ------------------------------------------------------------------------- */
/* whole-read pipeline (stages A..D of SignalAnalyzer.process,                */
/* signal_analyzer.py:82-134,230-286) for the numeric outputs; used by the     */
/* parity tests and as the timed CPU baseline.                                */
/* ------------------------------------------------------------------------- */
void orc_process_read(const orc_model *M, const int16_t *raw, int64_t n,
                      double gain, double offset, int flags, orc_result *R)
{
    const int stride = M->stride;
    memset(R, 0, sizeof *R);
    R->status = ORC_OKAY;
    R->barcode = -1; R->guess = -1; R->phred = -1;
    for (int s = 0; s < ORC_MAX_STATES; s++) R->seg[s][0] = R->seg[s][1] = -1;

    /* A2 load_padded_signal_head  signal_loader.py:212-231 */
    int64_t headlen = n < M->scaler_length ? n : M->scaler_length;
    headlen -= headlen % stride;
    if (headlen < M->scaler_min_length) { R->status = ORC_SCALER_SIGNAL_TOO_SHORT; return; }
    const int64_t npool_all = n / stride;
    float *pa = (float *)malloc(sizeof(float) * (size_t)(npool_all * stride + 1));
    float *pooled = (float *)malloc(sizeof(float) * (size_t)(npool_all + 1));
    orc_dac_to_pa(raw, npool_all * stride, gain, offset, pa);
    orc_pool_mean(pa, npool_all, stride, pooled);
    free(pa);

    const int Thead = M->scaler_length / stride;
    float *head = (float *)calloc((size_t)Thead, sizeof(float));
    const int nh = (int)(headlen / stride);
    memcpy(head + (Thead - nh), pooled, sizeof(float) * nh);   /* left zero-pad */

    /* A3 fit_scalers  signal_loader.py:89-109 */
    float z[2];
    orc_scaler_predict(&M->scaler, head, Thead, z);
    free(head);
    const double sc64 = M->scale_std * (double)z[0] + M->scale_mean;
    const double sh64 = M->shift_std * (double)z[1] + M->shift_mean;
    R->z[0] = z[0]; R->z[1] = z[1];
    R->scale = (float)sc64; R->shift = (float)sh64;
    if (!(sc64 >= M->qc_scale[0] && sc64 <= M->qc_scale[1] &&
          sh64 >= M->qc_shift[0] && sh64 <= M->qc_shift[1])) {
        R->status = ORC_SCALING_QC_FAIL;
        free(pooled);
        return;
    }

    /* A4 + A5 load_signal(pool) and detect_segments  signal_analyzer.py:346-364 */
    int64_t T = npool_all;
    const int64_t scan = M->scan_limit / stride;
    if (T > scan) T = scan;
    float *sig = (float *)malloc(sizeof(float) * (size_t)(T + 1));
    orc_scale(pooled, T, R->scale, R->shift, sig);
    int32_t *path = (int32_t *)malloc(sizeof(int32_t) * (size_t)(T + 1));
    R->viterbi_logp = orc_viterbi(&M->seg, sig, (int)T, path);
    int32_t seg[2 * ORC_MAX_STATES];
    if (R->viterbi_logp == -INFINITY) {
        R->status = ORC_UNKNOWN_ERROR;        /* reference: viterbi returns (-inf, None) */
        free(path); free(sig); free(pooled);
        return;
    }
    orc_segments_from_path(path, (int)T, M->seg.n_states, seg);
    for (int s = 0; s < M->seg.n_states; s++) { R->seg[s][0] = seg[2 * s]; R->seg[s][1] = seg[2 * s + 1]; }
    free(path);
    const int ad = M->adapter_state;
    if (R->seg[ad][0] < 0) {
        R->status = ORC_ADAPTER_NOT_DETECTED;
        free(sig); free(pooled);
        return;
    }

    /* A6 + A7 push / predict  barcoding.py:83-118 */
    if (flags & ORC_FLAG_BARCODING) {
        const int a0 = R->seg[ad][0], a1 = R->seg[ad][1];
        float win[ORC_MAX_WINDOW];
        if (orc_barcode_window(sig + a0, a1 - a0 + 1, M->demux_minlen, M->demux_maxlen,
                               M->demux_trimlen, M->demux_pad, win)) {
            R->pushed = 1;
            orc_demux_predict(&M->demux, win, M->demux_trimlen, R->probs);
            orc_barcode_decide(R->probs, M->demux.n_classes, M->n_decoy, M->calib,
                               M->n_calib, M->score_threshold,
                               &R->barcode, &R->guess, &R->phred);
        }
    }
    free(sig);
    free(pooled);
}

void orc_process_batch(const orc_model *M, const int16_t *raw, const int64_t *offsets,
                       const int64_t *lengths, const double *gain, const double *offset,
                       int64_t N, int flags, int nthreads, orc_result *R)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int64_t i = 0; i < N; i++)
        orc_process_read(M, raw + offsets[i], lengths[i], gain[i], offset[i], flags, &R[i]);
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
