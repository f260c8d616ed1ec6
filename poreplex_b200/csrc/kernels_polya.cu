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

__global__ void __launch_bounds__(POLYA_THREADS)
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
    PB_LAUNCH(ctx, K_POLYA, "k_polya", st,
        k_polya<<<(unsigned)((b.n_reads + POLYA_THREADS - 1) / POLYA_THREADS), POLYA_THREADS, 0, st>>>(
            P, b.raw, b.raw_offsets, b.raw_lengths, b.range, b.digitisation, b.offset, scale_shift,
            status, segments, b.n_reads, ctx->adapter_state, ctx->polya_state,
            reinterpret_cast<PolyaResult *>(out), cache, cap));
    return PB2_OK;
}

}  // namespace pb
