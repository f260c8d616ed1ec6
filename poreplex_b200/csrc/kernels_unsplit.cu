// kernels_unsplit.cu -- chimera filter (A12): SignalAnalysis.detect_unsplit_read
// (poreplex/signal_analyzer.py:366-443) + the derived event columns of load_events
// (:320-325) + utils.union_intervals (utils.py:28-39); restated in
// oracle/unsplit_restated.py and checked there against the reference running verbatim.
//
//   k_unsplit_windows  one thread per (read, window): the 8 s window every 3 s over the
//                      event table is decoded with the unsplit-read HMM (fp64 Viterbi,
//                      same core as the segmentation), the state path is run-length
//                      grouped and leader*->adapter runs that pass the duration cut-offs
//                      become candidate intervals.  Windows k and k+3 never overlap, so
//                      three back-pointer planes indexed by global event index suffice.
//   k_unsplit_decide   one thread per read: candidates up to the first empty window are
//                      sorted and merged, high-quality bases (max p_model_state per base
//                      > limit) are counted per sub-read, and the two limits decide.
#include "pb_internal.h"
#include "viterbi_core.cuh"

namespace pb {

constexpr int UW_MAX_CAND = 8;         // candidate intervals per window
constexpr int UD_MAX_INTERVALS = 64;   // candidate intervals per read
constexpr int UW_THREADS = 64;

struct UnsplitArgs {
    const int64_t *ev_off;     // [n + 1]
    const int64_t *start;      // [total]
    const float *mean;         // [total] unscaled
    const int32_t *move;       // [total]
    const double *p_state;     // [total]
    const double *rate;        // [n]
    const float *scale_shift;  // [n][2]
    const int32_t *status;     // [n]
    const int32_t *segments;   // [n][8][2]
    int64_t n;
    int stride, seg_adapter_state;
    int st_adapter, st_leader_high, st_leader_low;     // unsplit-model baked indices
    pb2_unsplit_params P;
    int max_windows;
    uint32_t *planes;          // [3][total]
    int64_t total_events;
    int64_t *cand;             // [n][max_windows][1 + 2 * UW_MAX_CAND]
    int32_t *flag;             // [n]
};

__device__ __forceinline__ int64_t ev_end(const UnsplitArgs &A, int64_t g, int64_t g_last) {
    // events['end'] = start + hstack(diff(start), [1])
    return g < g_last ? A.start[g + 1] : A.start[g] + 1;
}

__global__ void __launch_bounds__(UW_THREADS)
k_unsplit_windows(const HmmDev M, const HmmMask K, const UnsplitArgs A)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (r >= A.n) return;
    constexpr int CW = 1 + 2 * UW_MAX_CAND;
    int64_t *out = A.cand + (r * A.max_windows + k) * CW;
    out[0] = -2;                                   // not a window
    const int64_t e0 = A.ev_off[r], e1 = A.ev_off[r + 1];
    if (A.status[r] != PB2_ST_OKAY || e1 <= e0) return;
    const int adapter_last = A.segments[(r * PB2_MAX_STATES + A.seg_adapter_state) * 2 + 1];
    if (adapter_last < 0) return;
    const double rate = A.rate[r];
    const int64_t payload_start = ((int64_t)adapter_last + 1) * A.stride;
    const int64_t window_size = (int64_t)pb::dmul(A.P.window_size, rate);
    const int64_t window_step = (int64_t)pb::dmul(A.P.window_step, rate);
    const int64_t strict_duration = (int64_t)pb::dmul(A.P.strict_duration, rate);
    const int64_t cut_full[2] = {(int64_t)pb::dmul(A.P.loosen_full_length, rate),
                                 (int64_t)pb::dmul(A.P.strict_full_length, rate)};
    const int64_t cut_dna[2] = {(int64_t)pb::dmul(A.P.loosen_dna_length, rate),
                                (int64_t)pb::dmul(A.P.strict_dna_length, rate)};
    const int64_t g_last = e1 - 1;
    const int64_t last_end = A.start[g_last] + 1;
    if (window_step <= 0) return;
    const int64_t left = payload_start + (int64_t)k * window_step;
    if (left >= last_end) return;                  // range(payload_start, last_end, step)
    // evblock = events[start.between(left, left + window_size)]  (inclusive both ends)
    int64_t lo = e0, hi = e1;
    { int64_t a = e0, b = e1; while (a < b) { const int64_t m = (a + b) >> 1; if (A.start[m] < left) a = m + 1; else b = m; } lo = a; }
    { int64_t a = lo, b = e1; const int64_t right = left + window_size;
      while (a < b) { const int64_t m = (a + b) >> 1; if (A.start[m] <= right) a = m + 1; else b = m; } hi = a; }
    const int64_t T = hi - lo;
    if (T < 1) { out[0] = -1; return; }            // empty window: the reference breaks here
    const float scale = A.scale_shift[2 * r], shift = A.scale_shift[2 * r + 1];
    uint32_t *bp = A.planes + (int64_t)(k % 3) * A.total_events + lo;

    double v[PB2_MAX_STATES], e[PB2_MAX_STATES];
    hmm_emissions(M, (double)pb::fadd(pb::fmul(scale, A.mean[lo]), shift), e);
    viterbi_init(M, v, e);
    for (int64_t t = 1; t < T; t++) {
        hmm_emissions(M, (double)pb::fadd(pb::fmul(scale, A.mean[lo + t]), shift), e);
        bp[t] = viterbi_step(M, K, v, e);
    }
    double best;
    int cur = viterbi_end(M, v, best);
    if (best == pb::neg_inf()) { out[0] = -3; return; }   // reference: TypeError -> unknown_error
    for (int64_t t = T - 1; t >= 0; t--) {
        const uint32_t w = bp[t];
        bp[t] = (uint32_t)cur;
        if (t > 0) cur = (w >> (3 * cur)) & 7;
    }
    // run-length groups, leader* -> adapter (signal_analyzer.py:392-421)
    int64_t ncand = 0;
    int64_t leader_start = -1;
    int64_t t = 0;
    while (t < T) {
        const int s = (int)bp[t];
        const int64_t first = t;
        while (t + 1 < T && (int)bp[t + 1] == s) t++;
        const int64_t last = t;
        t++;
        if (s != A.st_adapter && s != A.st_leader_high && s != A.st_leader_low) { leader_start = -1; continue; }
        if (leader_start < 0) leader_start = first;
        if (s != A.st_adapter) continue;
        const int64_t adapter_end = ev_end(A, lo + last, g_last);
        const int64_t leader_in_read = A.start[lo + leader_start];
        const int64_t total_duration = adapter_end - leader_in_read;
        const int64_t adapter_duration = adapter_end - A.start[lo + first];
        const int strict = (leader_in_read - payload_start) <= strict_duration;
        if (total_duration >= cut_full[strict] && adapter_duration >= cut_dna[strict]) {
            if (ncand < UW_MAX_CAND) { out[1 + 2 * ncand] = leader_in_read; out[2 + 2 * ncand] = 1 + adapter_end; }
            ncand++;
        }
        leader_start = -1;
    }
    out[0] = ncand;
}

__device__ __forceinline__ int64_t count_hq(const UnsplitArgs &A, int64_t e0, int64_t e1,
                                            double lo_v, double hi_v)
{
    // events[start.between(lo, hi)] then groupby('pos')['p_model_state'].max() > limit;
    // pos = cumsum(move), so consecutive events share a base while move == 0
    int64_t a = e0, b = e1;
    while (a < b) { const int64_t m = (a + b) >> 1; if ((double)A.start[m] < lo_v) a = m + 1; else b = m; }
    const int64_t first = a;
    b = e1;
    while (a < b) { const int64_t m = (a + b) >> 1; if ((double)A.start[m] <= hi_v) a = m + 1; else b = m; }
    const int64_t last = a;             // [first, last)
    int64_t n = 0;
    int64_t i = first;
    while (i < last) {
        double best = A.p_state[i];
        int64_t j = i;
        while (j + 1 < last && A.move[j + 1] == 0) { j++; best = A.p_state[j] > best ? A.p_state[j] : best; }
        n += best > A.P.basecount_quality_limit;
        i = j + 1;
    }
    return n;
}

__global__ void __launch_bounds__(UW_THREADS)
k_unsplit_decide(const UnsplitArgs A)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= A.n) return;
    A.flag[r] = 0;
    constexpr int CW = 1 + 2 * UW_MAX_CAND;
    const int64_t e0 = A.ev_off[r], e1 = A.ev_off[r + 1];
    if (A.status[r] != PB2_ST_OKAY || e1 <= e0) return;
    int64_t ib[UD_MAX_INTERVALS], ie[UD_MAX_INTERVALS];
    int ni = 0;
    bool overflow = false;
    for (int k = 0; k < A.max_windows; k++) {
        const int64_t *c = A.cand + (r * A.max_windows + k) * CW;
        if (c[0] == -2 || c[0] == -1) break;       // past the last window / empty window
        if (c[0] == -3) { A.flag[r] = -1; return; }
        if (c[0] > UW_MAX_CAND) overflow = true;
        const int nc = c[0] > UW_MAX_CAND ? UW_MAX_CAND : (int)c[0];
        for (int q = 0; q < nc; q++) {
            if (ni < UD_MAX_INTERVALS) { ib[ni] = c[1 + 2 * q]; ie[ni] = c[2 + 2 * q]; ni++; }
            else overflow = true;
        }
    }
    if (overflow) { A.flag[r] = -2; return; }      // reported as an error, never guessed
    if (ni == 0) return;
    // sorted(iset): lexicographic insertion sort
    for (int i = 1; i < ni; i++) {
        const int64_t b = ib[i], e = ie[i];
        int j = i - 1;
        while (j >= 0 && (ib[j] > b || (ib[j] == b && ie[j] > e))) { ib[j + 1] = ib[j]; ie[j + 1] = ie[j]; j--; }
        ib[j + 1] = b; ie[j + 1] = e;
    }
    // union_intervals (utils.py:28-39)
    int nm = 0;
    for (int i = 0; i < ni; i++) {
        if (nm > 0 && ie[nm - 1] >= ib[i]) {
            if (ie[nm - 1] < ie[i]) ie[nm - 1] = ie[i];
            continue;
        }
        ib[nm] = ib[i]; ie[nm] = ie[i]; nm++;
    }
    const int adapter_last = A.segments[(r * PB2_MAX_STATES + A.seg_adapter_state) * 2 + 1];
    const int64_t payload_start = ((int64_t)adapter_last + 1) * A.stride;
    // sub-reads between [0, payload_start], merged adapters, [inf, inf]
    const double inf = pb::pos_inf();
    const int64_t sub0 = count_hq(A, e0, e1, (double)payload_start, (double)ib[0]);
    int64_t total = 0;
    for (int i = 0; i < nm; i++) {
        const double hi_v = (i + 1 < nm) ? (double)ib[i + 1] : inf;
        total += count_hq(A, e0, e1, (double)ie[i], hi_v);
    }
    const bool is_unsplit = (double)total > A.P.subread_basecount_limit ||
        pb::ddiv((double)(total + 1), (double)(sub0 + 1)) > A.P.subread_baseratio_limit;
    A.flag[r] = is_unsplit ? 1 : 0;
}

// ---------------------------------------------------------------------------
// k_event_means: the `mean` column of convert_events_guppy (fast5_file.py:209-230):
// pA of raw[first : first + stride * E] (clamped to the read), scipy medfilt(5) with zero
// padding at the ends of THAT slice, NaN padding of a short last block, then the float32
// pairwise row mean of reshape(E, stride).  One thread per event row.
// ---------------------------------------------------------------------------
__global__ void k_event_means(const int16_t *__restrict__ raw, const int64_t *__restrict__ raw_off,
                              const int64_t *__restrict__ raw_len, const double *__restrict__ range,
                              const double *__restrict__ digitisation,
                              const double *__restrict__ offset, const int64_t *__restrict__ ev_off,
                              const int64_t *__restrict__ first_sample, int stride, int64_t n,
                              int64_t total, float *__restrict__ mean)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total) return;
    int64_t a = 0, b = n;                      // read r with ev_off[r] <= g < ev_off[r + 1]
    while (a + 1 < b) { const int64_t m = (a + b) >> 1; if (ev_off[m] <= g) a = m; else b = m; }
    const int64_t r = a;
    const int64_t E = ev_off[r + 1] - ev_off[r];
    const int64_t first = first_sample[r];
    int64_t end = first + (int64_t)stride * E;
    if (end > raw_len[r]) end = raw_len[r];
    const double gain = pb::ddiv(range[r], digitisation[r]);
    const double off = offset[r];
    const int16_t *x = raw + raw_off[r];
    const int64_t base = first + (g - ev_off[r]) * stride;
    float v[32];
    for (int j = 0; j < stride; j++) {
        const int64_t p = base + j;
        if (p >= end) { v[j] = NAN; continue; }
        float w[5];
#pragma unroll
        for (int q = 0; q < 5; q++) {
            const int64_t pp = p - 2 + q;
            w[q] = (pp < first || pp >= end) ? 0.0f : pb::dac_to_pa((int)x[pp], gain, off);
        }
        v[j] = pb::median5(w[0], w[1], w[2], w[3], w[4]);
    }
    mean[g] = pb::pool_mean_generic(v, stride);
}

int launch_unsplit(pb2_context *ctx, const pb2_batch *batch, const pb2_event_tables &ev_in,
                   int64_t n, const float *scale_shift, const int32_t *status,
                   const int32_t *segments, int32_t max_windows, int32_t *flag, cudaStream_t st)
{
    if (n <= 0) return PB2_OK;
    if (max_windows < 1) max_windows = 1;
    pb2_event_tables ev = ev_in;
    if (!ev.mean) {
        if (!batch || !ev.first_sample || ev.block_stride < 1 || ev.block_stride > 32)
            return fail(ctx, PB2_EINVAL, "event means missing and cannot be derived "
                        "(need batch, first_sample and 1 <= block_stride <= 32)");
        float *m = (float *)ws_get(ctx, ctx->ws_evmean, sizeof(float) * (size_t)(ev.n_events_total + 1));
        if (!m) return PB2_ENOMEM;
        if (ev.n_events_total > 0) {
            PB_LAUNCH(ctx, K_EVENT_MEANS, "k_event_means", st,
                k_event_means<<<(unsigned)((ev.n_events_total + 127) / 128), 128, 0, st>>>(
                    batch->raw, batch->raw_offsets, batch->raw_lengths, batch->range,
                    batch->digitisation, batch->offset, ev.event_offsets, ev.first_sample,
                    ev.block_stride, n, ev.n_events_total, m));
        }
        ev.mean = m;
    }
    UnsplitArgs A = {};
    A.ev_off = ev.event_offsets; A.start = ev.start; A.mean = ev.mean; A.move = ev.move;
    A.p_state = ev.p_model_state; A.rate = ev.sampling_rate;
    A.scale_shift = scale_shift; A.status = status; A.segments = segments; A.n = n;
    A.stride = ctx->scaler.stride; A.seg_adapter_state = ctx->adapter_state;
    A.st_adapter = ctx->unsplit_states[0]; A.st_leader_high = ctx->unsplit_states[1];
    A.st_leader_low = ctx->unsplit_states[2];
    A.P = ctx->unsplit;
    A.max_windows = max_windows;
    A.total_events = ev.n_events_total;
    constexpr int CW = 1 + 2 * UW_MAX_CAND;
    char *base = (char *)ws_get(ctx, ctx->ws_unsplit,
                                sizeof(uint32_t) * 3 * (size_t)(ev.n_events_total + 1) +
                                sizeof(int64_t) * (size_t)n * max_windows * CW + 64);
    if (!base) return PB2_ENOMEM;
    A.cand = (int64_t *)base;
    A.planes = (uint32_t *)(base + sizeof(int64_t) * (size_t)n * max_windows * CW);
    A.flag = flag;
    HmmMask K;
    make_mask(ctx->unsplit_hmm, K);
    dim3 grid((unsigned)((n + UW_THREADS - 1) / UW_THREADS), (unsigned)max_windows);
    PB_LAUNCH(ctx, K_UNSPLIT_WINDOWS, "k_unsplit_windows", st,
        k_unsplit_windows<<<grid, UW_THREADS, 0, st>>>(ctx->unsplit_hmm, K, A));
    PB_LAUNCH(ctx, K_UNSPLIT_DECIDE, "k_unsplit_decide", st,
        k_unsplit_decide<<<(unsigned)((n + UW_THREADS - 1) / UW_THREADS), UW_THREADS, 0, st>>>(A));
    return PB2_OK;
}

}  // namespace pb
