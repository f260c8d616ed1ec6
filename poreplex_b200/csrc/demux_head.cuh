// demux_head.cuh -- Dense(n_classes) + softmax + the decision rule of barcoding.py:103-118,
// shared by the exact layer-2 kernel (kernels_lstm.cu) and the tensor-core path
// (kernels_lstm_tc.cu) so both make their calls with the same arithmetic.
#pragma once
#include "pb_internal.h"
#include "pb_math.cuh"

namespace pb {

struct DemuxCall {
    float probs[PB2_MAX_CLASSES];
    float logit[PB2_MAX_CLASSES];
    int arg;            // argmax class (first maximum wins)
    float best;         // its probability
    int barcode, guess, score;
};

// h[k * hstride], k < H2: final hidden state of layer 2 for one read
template <int H2>
__device__ __forceinline__ void demux_head(const float *h, int hstride, const float *__restrict__ Wd,
                                           const float *__restrict__ bd, int nc, int n_decoy,
                                           double score_threshold, const double *calibration,
                                           int n_calibration, DemuxCall &out)
{
    float e[PB2_MAX_CLASSES];
#pragma unroll
    for (int j = 0; j < PB2_MAX_CLASSES; j++) out.logit[j] = 0.f;
    for (int k = 0; k < H2; k++) {
        const float hk = h[k * hstride];
#pragma unroll
        for (int j = 0; j < PB2_MAX_CLASSES; j++)
            if (j < nc) out.logit[j] = pb::ffma(hk, Wd[k * nc + j], out.logit[j]);
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < PB2_MAX_CLASSES; j++)
        if (j < nc) { out.logit[j] = pb::fadd(out.logit[j], bd[j]); m = fmaxf(m, out.logit[j]); }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < PB2_MAX_CLASSES; j++)
        if (j < nc) { e[j] = pb::exp_eigen(pb::fsub(out.logit[j], m)); sum = pb::fadd(sum, e[j]); }
    const float rs = pb::fdiv(1.0f, sum);
    int arg = 0;
    float best = -1.f;
#pragma unroll
    for (int j = 0; j < PB2_MAX_CLASSES; j++) {
        float p = 0.f;
        if (j < nc) {
            p = pb::fmul(e[j], rs);
            if (p > best) { best = p; arg = j; }
        }
        out.probs[j] = p;
    }
    out.arg = arg;
    out.best = best;
    // barcoding.py:108-118
    const int bcid = arg - n_decoy;
    const double sc = (double)best;
    out.barcode = (bcid >= 0 && sc >= score_threshold) ? bcid : -1;
    out.guess = bcid;
    int lo = 0;
    if (sc > 0.0) {                 // bisect_right(calibration, score)
        int hi = n_calibration;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (sc < calibration[mid]) hi = mid; else lo = mid + 1;
        }
    }
    out.score = lo;
}

// Can (arg, barcode, score) change if every class logit RELATIVE TO THE CALLED CLASS,
// d_j = l_j - l_arg, moves by at most `delta`?  p_best = 1 / (1 + sum_j exp(d_j)) then moves by
// at most delta p (1 - p); 1e-6 covers the f32 rounding of the softmax itself.  Returns true
// when the call is safe.  (`delta` is calibrated on exactly this quantity, the largest
// |d_j(tensor core) - d_j(exact)| of a window, see tools/tc_diag.py.)
__device__ __forceinline__ bool demux_call_is_safe(const DemuxCall &c, int nc, double delta,
                                                   double score_threshold,
                                                   const double *calibration, int n_calibration)
{
    float second = -INFINITY;
#pragma unroll
    for (int j = 0; j < PB2_MAX_CLASSES; j++)
        if (j < nc && j != c.arg) second = fmaxf(second, c.logit[j]);
    if ((double)c.logit[c.arg] - (double)second <= delta) return false;
    const double s = (double)c.best;
    const double tol = delta * s * (1.0 - s) + 1e-6;
    if (fabs(s - score_threshold) <= tol) return false;
    for (int i = 0; i < n_calibration; i++)
        if (fabs(s - calibration[i]) <= tol) return false;
    return true;
}

}  // namespace pb
