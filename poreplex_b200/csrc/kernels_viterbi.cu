// kernels_viterbi.cu -- HMM Viterbi segmentation (A5).
//
// Reference: SignalAnalysis.detect_segments (signal_analyzer.py:346-364) calling
// pomegranate's HiddenMarkovModel.viterbi (restated in oracle/pb_oracle.c
// orc_viterbi).  fp64 log space, candidates evaluated as
// ((v[src] + logT) + e[dst]) in baked source order, strict '>' (first maximum wins),
// final state = first argmax.
//
// Mapping: one thread per read.  The recurrence is strictly sequential in time and
// every read is independent, so a batch of 10^5..10^6 reads fills the machine with
// the oracle's exact operation order intact (no re-association).  Each thread streams
// its pooled samples, keeps the 8 state scores in registers, and writes one packed
// back-pointer word per time step to a [t][read] scratch matrix (coalesced across the
// warp); the traceback walks that matrix backwards and emits run-length segments.
#include "pb_internal.h"
#include "pb_math.cuh"
#include "viterbi_core.cuh"

namespace pb {

constexpr int VT_THREADS = 128;

// ---------------------------------------------------------------------------
// k_segment: scale (signal_loader.py:262, unfused) + Viterbi + run-length segments
// (signal_analyzer.py:355-362: a later run of a state overwrites an earlier one).
// ---------------------------------------------------------------------------
// Topology of the stock segmentation model (rna-r941.cfg:61-101 in pomegranate's baked, i.e.
// alphabetical, state order: adapter, leader-high, leader-low, polya-tail, pre-leader,
// transcript): in-edges 0<-{0,1} 1<-{1,2} 2<-{2,4} 3<-{0,3} 4<-{4} 5<-{0,3,5}.  k_segment is
// compiled once for it (EDGES != 0) and once generically; launch_segment picks by comparing masks.
constexpr uint32_t SEG_STOCK_NCOMP = 0x211112;      // mixture components: adapter 2, transcript 2, others 1
constexpr uint64_t SEG_STOCK_EDGES =
    (0x03ull << 0) | (0x06ull << 8) | (0x14ull << 16) | (0x09ull << 24) | (0x10ull << 32) | (0x29ull << 40);

// One launch decodes the same reads under up to three (scale, shift) sets -- blockIdx.y selects:
// the tensor-core path decodes every read at the three corners of its uncertainty triangle
// (DESIGN.md 3a), and one grid of 3 n threads fills the machine where three grids of n do not
// (a chunk of a host batch is a fraction of a wave of this kernel).
struct SegSets {
    const float *scale_shift[3];
    int32_t *status[3];
    int32_t *segments[3];
};

template <uint64_t EDGES, int NS>
__global__ void __launch_bounds__(VT_THREADS)
k_segment(const HmmDev M, const HmmMask K, const int64_t *__restrict__ raw_offsets,
          const int64_t *__restrict__ raw_lengths, const float *__restrict__ pooled,
          const SegSets S, int64_t r0, int64_t n_chunk, int stride,
          int scan_limit, int adapter_state, uint32_t *__restrict__ bp, int64_t bp_set_stride,
          float *__restrict__ pooled_scaled_out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunk) return;
    const float *__restrict__ scale_shift = S.scale_shift[blockIdx.y];
    int32_t *status = S.status[blockIdx.y];
    int32_t *__restrict__ segments = S.segments[blockIdx.y];
    bp += (int64_t)blockIdx.y * bp_set_stride;
    const int64_t r = r0 + i;
    int32_t *seg = segments + r * PB2_MAX_STATES * 2;
#pragma unroll
    for (int s = 0; s < PB2_MAX_STATES * 2; s++) seg[s] = -1;
    if (status[r] != PB2_ST_OKAY) return;
    int64_t T64 = raw_lengths[r] / stride;
    const int T = (int)(T64 > scan_limit ? scan_limit : T64);
    if (T <= 0) { status[r] = PB2_ST_UNKNOWN_ERROR; return; }
    const float scale = scale_shift[2 * r], shift = scale_shift[2 * r + 1];
    const int64_t po = pooled_offset(raw_offsets[r], stride);
    const float *x = pooled + po;

    double v[PB2_MAX_STATES], e[PB2_MAX_STATES];
    {
        const float y = pb::fadd(pb::fmul(scale, x[0]), shift);
        if (pooled_scaled_out) pooled_scaled_out[po] = y;
        if (EDGES != 0) hmm_emissions_topo<SEG_STOCK_NCOMP, NS>(M, (double)y, e);
        else hmm_emissions(M, (double)y, e);
        viterbi_init(M, v, e);
    }
    for (int t = 1; t < T; t++) {
        const float y = pb::fadd(pb::fmul(scale, x[t]), shift);
        if (pooled_scaled_out) pooled_scaled_out[po + t] = y;
        if (EDGES != 0) hmm_emissions_topo<SEG_STOCK_NCOMP, NS>(M, (double)y, e);
        else hmm_emissions(M, (double)y, e);
        bp[(int64_t)t * n_chunk + i] = (EDGES != 0) ? viterbi_step_topo<EDGES, NS>(K, v, e)
                                                    : viterbi_step(M, K, v, e);
    }
    double best;
    int cur = viterbi_end(M, v, best);
    if (best == pb::neg_inf()) { status[r] = PB2_ST_UNKNOWN_ERROR; return; }

    // traceback: emit (first, last) of the LAST run of every state
    uint32_t seen = 0;
    int run_last = T - 1;
    for (int t = T - 1; t >= 0; t--) {
        int prev = -1;
        if (t > 0) prev = (bp[(int64_t)t * n_chunk + i] >> (3 * cur)) & 7;
        if (t == 0 || prev != cur) {
            if (!(seen & (1u << cur))) {
                seen |= 1u << cur;
                seg[2 * cur] = t;
                seg[2 * cur + 1] = run_last;
            }
            run_last = t - 1;
            cur = prev;
        }
    }
    if (!(seen & (1u << adapter_state))) status[r] = PB2_ST_ADAPTER_NOT_DETECTED;
}

static int launch_segment_sets(pb2_context *ctx, const pb2_batch &b, const float *pooled,
                               const SegSets &S, int n_sets, float *pooled_scaled_out, cudaStream_t st)
{
    if (b.n_reads <= 0) return PB2_OK;
    HmmMask K;
    make_mask(ctx->seg_hmm, K);
    const bool stock_topology = ctx->seg_hmm.n_states == 6 && pack_edges(K) == SEG_STOCK_EDGES &&
                                pack_ncomp(ctx->seg_hmm) == SEG_STOCK_NCOMP && !ctx->generic_viterbi;
    int64_t Tmax = ctx->scan_limit_pooled;
    if (b.max_raw_length > 0 && b.max_raw_length / ctx->scaler.stride < Tmax)
        Tmax = b.max_raw_length / ctx->scaler.stride;
    if (Tmax < 1) Tmax = 1;
    // back-pointer scratch is [set][Tmax][chunk] words; bound it to ~2 GiB per set and launch
    int64_t chunk = ((int64_t)2 << 30) / (Tmax * 4);
    chunk = (chunk / VT_THREADS) * VT_THREADS;
    if (chunk < VT_THREADS) chunk = VT_THREADS;
    if (chunk > b.n_reads) chunk = b.n_reads;
    uint32_t *bp = (uint32_t *)ws_get(ctx, ctx->ws_bp, (size_t)chunk * Tmax * 4 * n_sets);
    if (!bp) return PB2_ENOMEM;
    const int64_t set_stride = chunk * Tmax;
    for (int64_t r0 = 0; r0 < b.n_reads; r0 += chunk) {
        const int64_t nc = (b.n_reads - r0 < chunk) ? b.n_reads - r0 : chunk;
        const dim3 grid((unsigned)((nc + VT_THREADS - 1) / VT_THREADS), (unsigned)n_sets);
        if (stock_topology) {
            PB_LAUNCH(ctx, K_SEGMENT, "k_segment<stock>", st,
                k_segment<SEG_STOCK_EDGES, 6><<<grid, VT_THREADS, 0, st>>>(
                ctx->seg_hmm, K, b.raw_offsets, b.raw_lengths, pooled, S, r0, nc,
                ctx->scaler.stride, (int)Tmax, ctx->adapter_state, bp, set_stride, pooled_scaled_out));
        } else {
            PB_LAUNCH(ctx, K_SEGMENT, "k_segment", st,
                k_segment<0, 0><<<grid, VT_THREADS, 0, st>>>(
                ctx->seg_hmm, K, b.raw_offsets, b.raw_lengths, pooled, S, r0, nc,
                ctx->scaler.stride, (int)Tmax, ctx->adapter_state, bp, set_stride, pooled_scaled_out));
        }
    }
    return PB2_OK;
}

int launch_segment(pb2_context *ctx, const pb2_batch &b, const float *pooled,
                   const float *scale_shift, int32_t *status, int32_t *segments,
                   float *pooled_scaled_out, cudaStream_t st)
{
    SegSets S = {{scale_shift, nullptr, nullptr}, {status, nullptr, nullptr}, {segments, nullptr, nullptr}};
    return launch_segment_sets(ctx, b, pooled, S, 1, pooled_scaled_out, st);
}

// the three corner decodings of the tensor-core path in one launch; ss3 = [3][n][2]
int launch_segment3(pb2_context *ctx, const pb2_batch &b, const float *pooled, const float *ss3,
                    int32_t *const status[3], int32_t *const segments[3], cudaStream_t st)
{
    const int64_t n = b.n_reads;
    SegSets S = {{ss3, ss3 + 2 * n, ss3 + 4 * n}, {status[0], status[1], status[2]},
                 {segments[0], segments[1], segments[2]}};
    return launch_segment_sets(ctx, b, pooled, S, 3, nullptr, st);
}

// ---------------------------------------------------------------------------
// k_viterbi_paths: HiddenMarkovModel.viterbi over dense rows, full state path out.
// Used by the parity tests and by windowed decoding (unsplit-read model).
// The path buffer itself doubles as back-pointer storage: path[r][t] first holds the
// packed word of step t, then is overwritten by the decoded state during traceback.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(VT_THREADS)
k_viterbi_paths(const HmmDev M, const HmmMask K, const float *__restrict__ x,
                const int32_t *__restrict__ lengths, int64_t n, int ld,
                int32_t *__restrict__ path, double *__restrict__ logp)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int T = lengths[r] < ld ? lengths[r] : ld;
    const float *xr = x + r * ld;
    int32_t *pr = path + r * ld;
    if (T <= 0) { if (logp) logp[r] = pb::neg_inf(); return; }
    double v[PB2_MAX_STATES], e[PB2_MAX_STATES];
    hmm_emissions(M, (double)xr[0], e);
    viterbi_init(M, v, e);
    for (int t = 1; t < T; t++) {
        hmm_emissions(M, (double)xr[t], e);
        pr[t] = (int32_t)viterbi_step(M, K, v, e);
    }
    double best;
    int cur = viterbi_end(M, v, best);
    if (logp) logp[r] = best;
    if (best == pb::neg_inf()) {
        for (int t = 0; t < T; t++) pr[t] = -1;
        return;
    }
    for (int t = T - 1; t >= 0; t--) {
        const uint32_t w = (uint32_t)pr[t];
        pr[t] = cur;
        if (t > 0) cur = (w >> (3 * cur)) & 7;
    }
}

int launch_viterbi_paths(pb2_context *ctx, const HmmDev &hmm, const float *x,
                         const int32_t *lengths, int64_t n, int32_t ld, int32_t *path,
                         double *logp, cudaStream_t st)
{
    if (n <= 0) return PB2_OK;
    HmmMask K;
    make_mask(hmm, K);
    PB_LAUNCH(ctx, K_VITERBI_PATHS, "k_viterbi_paths", st,
        k_viterbi_paths<<<(unsigned)((n + VT_THREADS - 1) / VT_THREADS), VT_THREADS, 0, st>>>(
        hmm, K, x, lengths, n, ld, path, logp));
    return PB2_OK;
}

}  // namespace pb
