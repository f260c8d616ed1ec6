// kernels_polya.cu -- poly(A) dwell measurement (A8 + A9): PolyASignalAnalyzer.__call__
// (poreplex/polya.py:50-187) with csupport.detect_events (src/csupport.c:70-124,
// src/contrib/scrappie/event_detection.c) inlined as a stream.  One thread per read runs
// pb::polya_analyze (polya_core.cuh); reads are independent, so a batch of 10^5..10^6
// reads fills the machine while every read keeps the reference's exact sequential
// arithmetic (fp64 prefix sums, state-machine peak detector, first-maximum interval).
#include "pb_internal.h"
#include "polya_core.cuh"

namespace pb {

static_assert(sizeof(PolyaParams) == sizeof(pb2_polya_params), "PolyaParams mirrors the ABI struct");
static_assert(sizeof(PolyaResult) == sizeof(pb2_polya_result), "PolyaResult mirrors the ABI struct");

constexpr int POLYA_THREADS = 64;

// NESTED: the literal one-loop-per-walk formulation (polya_analyze_nested), kept for comparison
// (POREPLEX_B200_POLYA_NESTED=1); the default is the single-loop one (polya_core.cuh).
// MINB: resident blocks per SM the register allocation aims at (POREPLEX_B200_POLYA_MINB = 4, 6, 8;
// measured per 1 M reads: 4 -> 221 registers, 8 warps per SM, 155 ms; 6 -> 168 registers and a few
// spilled words, 12 warps, 124 ms: the default).
template <bool NESTED, int MINB>
__global__ void __launch_bounds__(POLYA_THREADS, MINB)
k_polya(const PolyaParams P, const int16_t *__restrict__ raw,
        const int64_t *__restrict__ raw_offsets, const int64_t *__restrict__ raw_lengths,
        const double *__restrict__ range, const double *__restrict__ digitisation,
        const double *__restrict__ offset, const float *__restrict__ scale_shift,
        const int32_t *__restrict__ status, const int32_t *__restrict__ segments, int64_t n,
        int adapter_state, int polya_state, PolyaResult *__restrict__ out,
        EventCacheSlot *__restrict__ cache, int cache_cap)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    PolyaResult &R = out[r];
    R.found = 0; R.n_spikes = 0; R.begin = 0; R.end = 0; R.dwell_samples = 0;
    R.extensions = 0; R.flags = 0;
    if (status[r] != PB2_ST_OKAY) return;
    const int32_t *seg = segments + r * PB2_MAX_STATES * 2;
    // signal_analyzer.py:251-256: polya-tail segment, else open range after the adapter
    int32_t rb, re;
    if (polya_state >= 0 && seg[2 * polya_state] >= 0) {
        rb = seg[2 * polya_state];
        re = seg[2 * polya_state + 1];
    } else {
        rb = seg[2 * adapter_state + 1] + 1;
        re = -1;
    }
    const double gain = pb::ddiv(range[r], digitisation[r]);
    // event replay cache: slot k of read r at cache[k * n + r] (coalesced across the warp)
    if (NESTED)
        polya_analyze_nested(P, raw + raw_offsets[r], raw_lengths[r], gain, offset[r], scale_shift[2 * r],
                             scale_shift[2 * r + 1], rb, re, R, cache ? cache + r : nullptr, n, cache_cap);
    else
        polya_analyze(P, raw + raw_offsets[r], raw_lengths[r], gain, offset[r], scale_shift[2 * r],
                      scale_shift[2 * r + 1], rb, re, R, cache ? cache + r : nullptr, n, cache_cap);
}

int launch_polya(pb2_context *ctx, const pb2_batch &b, const float *scale_shift,
                 const int32_t *status, const int32_t *segments, pb2_polya_result *out,
                 cudaStream_t st)
{
    if (b.n_reads <= 0) return PB2_OK;
    PolyaParams P;
    memcpy(&P, &ctx->polya, sizeof P);
    // replay cache: up to 192 events per read, bounded to ~3 GiB
    int cap = 192;
    while (cap > 0 && (size_t)cap * (size_t)b.n_reads * sizeof(EventCacheSlot) > ((size_t)3 << 30)) cap /= 2;
    EventCacheSlot *cache = cap >= 16
        ? (EventCacheSlot *)ws_get(ctx, ctx->ws_polya, (size_t)cap * (size_t)b.n_reads * sizeof(EventCacheSlot))
        : nullptr;
    if (!cache) { cap = 0; cudaGetLastError(); }
    static const bool nested = [] { const char *e = getenv("POREPLEX_B200_POLYA_NESTED"); return e && e[0] == '1'; }();
    static const int minb = [] { const char *e = getenv("POREPLEX_B200_POLYA_MINB"); return e ? atoi(e) : 6; }();
    const unsigned grid = (unsigned)((b.n_reads + POLYA_THREADS - 1) / POLYA_THREADS);
#define PB_POLYA(NESTED, MINB)                                                                   \
    PB_LAUNCH(ctx, K_POLYA, "k_polya", st,                                                       \
        (k_polya<NESTED, MINB><<<grid, POLYA_THREADS, 0, st>>>(                                  \
            P, b.raw, b.raw_offsets, b.raw_lengths, b.range, b.digitisation, b.offset, scale_shift, \
            status, segments, b.n_reads, ctx->adapter_state, ctx->polya_state,                   \
            reinterpret_cast<PolyaResult *>(out), cache, cap)))
    if (nested) PB_POLYA(true, 4);
    else if (minb == 4) PB_POLYA(false, 4);
    else if (minb == 8) PB_POLYA(false, 8);
    else PB_POLYA(false, 6);
#undef PB_POLYA
    return PB2_OK;
}

// ---------------------------------------------------------------------------
// k_detect_events: csupport.detect_events (src/csupport.c:70-124 ->
// src/contrib/scrappie/event_detection.c:273-324) for a batch of float32 signals, the
// reference's one native entry point as a stand-alone call.  One thread per signal runs the
// same streaming detector k_polya uses (EventStreamT over the plain signal).  The number of
// events is not known up front: a first pass (FILL = false) counts, the caller turns counts
// into offsets, a second pass writes the 28-byte records of csupport.c:156-159
// (start u8, length f4, mean f4, stdv f4, pos i4 = -1, state i4 = -1).
// ---------------------------------------------------------------------------
constexpr int DETECT_THREADS = 64;

template <int RING, bool FILL>
__global__ void __launch_bounds__(DETECT_THREADS)
k_detect_events(const PolyaParams P, const float *__restrict__ signal,
                const int64_t *__restrict__ offsets, const int64_t *__restrict__ lengths, int64_t n,
                int64_t *__restrict__ counts, const int64_t *__restrict__ event_offsets,
                uint32_t *__restrict__ records)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int64_t len = lengths[r];
    int64_t k = 0;
    if (len > 0) {                                 // empty signal: no events (csupport raises)
        PlainSource src;
        src.x = signal + offsets[r]; src.n = len; src.next = 0;
        PlainEventStream<RING> es;
        EventRings<RING> rings;
        es.use(rings);
        es.begin(src, P);
        Event ev[2];
        const int64_t base = FILL ? event_offsets[r] : 0;
        while (!es.finished()) {
            const int m = es.step(ev);
            for (int q = 0; q < m; q++) {
                if (FILL) {
                    uint32_t *w = records + (base + k) * 7;
                    w[0] = (uint32_t)ev[q].start;
                    w[1] = (uint32_t)(ev[q].start >> 32);
                    w[2] = __float_as_uint(ev[q].length);
                    w[3] = __float_as_uint(ev[q].mean);
                    w[4] = __float_as_uint(ev[q].stdv);
                    w[5] = 0xFFFFFFFFu;
                    w[6] = 0xFFFFFFFFu;
                }
                k++;
            }
        }
    }
    if (!FILL) counts[r] = k;
}

int launch_detect_events(pb2_context *ctx, const float *signal, const int64_t *offsets,
                         const int64_t *lengths, int64_t n, const pb2_detector_params &dp,
                         int64_t *counts, const int64_t *event_offsets, void *records,
                         cudaStream_t st)
{
    if (n <= 0) return PB2_OK;
    if (dp.window_length1 < 1 || dp.window_length2 < 1)
        return fail(ctx, PB2_EINVAL, "detect_events: window lengths must be >= 1");
    const int64_t wmax = dp.window_length1 > dp.window_length2 ? dp.window_length1 : dp.window_length2;
    if (2 * wmax + 2 > 512)
        return fail(ctx, PB2_EUNSUPPORTED, "detect_events: window length %lld > 255", (long long)wmax);
    PolyaParams P = {};
    P.w1 = (int32_t)dp.window_length1; P.w2 = (int32_t)dp.window_length2;
    P.thr1 = dp.threshold1; P.thr2 = dp.threshold2; P.peak_height = dp.peak_height;
    const unsigned grid = (unsigned)((n + DETECT_THREADS - 1) / DETECT_THREADS);
    const bool small = 2 * wmax + 2 <= 64;
    uint32_t *rec = (uint32_t *)records;
#define PB_DETECT(RING, FILL)                                                                  \
    PB_LAUNCH(ctx, K_MISC, "k_detect_events", st,                                              \
        k_detect_events<RING, FILL><<<grid, DETECT_THREADS, 0, st>>>(P, signal, offsets, lengths, n,  \
                                                                    counts, event_offsets, rec))
    if (records) {
        if (!event_offsets) return fail(ctx, PB2_EINVAL, "detect_events: records without event_offsets");
        if (small) PB_DETECT(64, true); else PB_DETECT(512, true);
    } else {
        if (!counts) return fail(ctx, PB2_EINVAL, "detect_events: neither counts nor records");
        if (small) PB_DETECT(64, false); else PB_DETECT(512, false);
    }
#undef PB_DETECT
    return PB2_OK;
}

}  // namespace pb
