// kernels_events.cu -- event-table derivation from guppy `Move` tables on the device
// (SURVEY.md section 8f rank 3).  Replaces, for a whole batch of reads at once,
//   Fast5Reader.construct_events_from_moves   fast5_file.py:183-207
//   Fast5Reader.convert_events_guppy          fast5_file.py:209-230
//   SignalAnalysis.load_events (derived cols) signal_analyzer.py:311-326
// which the reference runs per read in pandas / numpy.
//
// Columns and their arithmetic (all as the reference's numpy computes them):
//   start   = arange(first_sample, first_sample + stride * E, stride)
//   length  = stride
//   mean    = float32 pairwise row mean of medfilt(pA, 5) reshaped [E][stride]   (k_event_stats)
//   stdv    = np.std of the same row: sqrt(pairwise_sum((x - mean)^2) / stride), float32
//   move    = as given
//   pos     = cumsum(move)                                  (signal_analyzer.py:321)
//   end     = start + diff(start), last + 1                 (signal_analyzer.py:323-324)
//   p_model_state = qual[cumsum(move) - 1 + posshift], qual = 1 - 10 ** -((q - 33) / 10) taken
//             from a 256-entry float64 table the host builds with the reference's own numpy
//             expression (libm pow is not reproducible on a GPU; the table is exact)
//   model_state   = 5 characters of the reversed sequence (U -> T; '__' padded on both sides
//             for flip-flop models) at cumsum(move) - 1
//   scaled_mean   = float32 poly1d(scale, shift)(mean): unfused multiply, add
// Per-read errors the reference raises as exceptions come back as codes:
//   1 "Move table is encoded with an unknown kmer-size."     fast5_file.py:197
//   2 "Numbers of events and raw data strides does not match." fast5_file.py:221
#include "pb_internal.h"
#include "pb_math.cuh"

namespace pb {

struct EventArgs {
    const int16_t *raw; const int64_t *raw_off, *raw_len;
    const double *range, *digitisation, *offset;
    const int64_t *ev_off, *first_sample;
    const int32_t *move;
    const uint8_t *qstring, *sequence; const int64_t *seq_off;
    const double *qual_table;
    const float *scale_shift;
    int stride; int64_t n, total;
    pb2_event_columns out;
};

// one thread per event row: mean, stdv, start, end, length, scaled_mean
__global__ void k_event_stats(const EventArgs A)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= A.total) return;
    int64_t a = 0, b = A.n;                    // read r with ev_off[r] <= g < ev_off[r + 1]
    while (a + 1 < b) { const int64_t m = (a + b) >> 1; if (A.ev_off[m] <= g) a = m; else b = m; }
    const int64_t r = a;
    const int64_t E = A.ev_off[r + 1] - A.ev_off[r];
    const int64_t first = A.first_sample[r];
    int64_t end = first + (int64_t)A.stride * E;
    if (end > A.raw_len[r]) end = A.raw_len[r];
    const double gain = pb::ddiv(A.range[r], A.digitisation[r]);
    const double off = A.offset[r];
    const int16_t *x = A.raw + A.raw_off[r];
    const int64_t k = g - A.ev_off[r];
    const int64_t base = first + k * A.stride;
    float v[32];
    for (int j = 0; j < A.stride; j++) {
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
    const float mean = pb::pool_mean_generic(v, A.stride);
    if (A.out.mean) A.out.mean[g] = mean;
    if (A.out.stdv) {
        // numpy _var: x = arr - arrmean; x = x * x; sum(x) / n; sqrt -- every step float32
        float d[32];
        for (int j = 0; j < A.stride; j++) { const float t = pb::fsub(v[j], mean); d[j] = pb::fmul(t, t); }
        A.out.stdv[g] = pb::fsqrt(pb::pool_mean_generic(d, A.stride));
    }
    if (A.out.scaled_mean && A.scale_shift)
        A.out.scaled_mean[g] = pb::fadd(pb::fmul(A.scale_shift[2 * r], mean), A.scale_shift[2 * r + 1]);
    if (A.out.start) A.out.start[g] = base;
    if (A.out.end) A.out.end[g] = (k + 1 < E) ? base + A.stride : base + 1;
    if (A.out.length) A.out.length[g] = A.stride;
}

// one warp per read: pos = cumsum(move) by warp scans over chunks of 32 events, then the
// k-mer size check and the two look-ups that hang off pos
__global__ void k_event_pos(const EventArgs A)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= A.n) return;
    const int64_t e0 = A.ev_off[r], E = A.ev_off[r + 1] - e0;
    int64_t carry = 0;
    for (int64_t c = 0; c < E; c += 32) {
        const int64_t i = c + lane;
        int64_t v = (i < E) ? (int64_t)A.move[e0 + i] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t u = __shfl_up_sync(0xFFFFFFFFu, v, d);
            if (lane >= d) v += u;
        }
        if (i < E && A.out.pos) A.out.pos[e0 + i] = carry + v;
        carry += __shfl_sync(0xFFFFFFFFu, v, 31);
    }
    int err = 0;
    // convert_events_guppy: rows of the padded raw slice must equal the number of events
    {
        const int64_t first = A.first_sample[r];
        int64_t end = first + (int64_t)A.stride * E;
        if (end > A.raw_len[r]) end = A.raw_len[r];
        const int64_t nraw = end > first ? end - first : 0;
        const int64_t rows = (nraw + A.stride - 1) / A.stride;
        if (rows != E) err = 2;
    }
    int shift = 0;
    bool flipflop = false;
    int64_t slen = 0;
    if (A.seq_off) {
        slen = A.seq_off[r + 1] - A.seq_off[r];
        const int64_t kmer = slen - carry + 1;
        if (kmer == 5) shift = 2;
        else if (kmer == 1) flipflop = true;
        else if (!err) err = 1;
    }
    if (lane == 0 && A.out.error) A.out.error[r] = err;
    if (err == 1 || !A.seq_off || !A.out.pos) return;
    const uint8_t *q = A.qstring ? A.qstring + A.seq_off[r] : nullptr;
    const uint8_t *s = A.sequence ? A.sequence + A.seq_off[r] : nullptr;
    __syncwarp();
    for (int64_t i = lane; i < E; i += 32) {
        const int64_t p = A.out.pos[e0 + i] - 1;                 // moves.cumsum() - 1
        if (A.out.p_model_state && q && A.qual_table) {
            const int64_t qi = p + shift;
            A.out.p_model_state[e0 + i] = (qi >= 0 && qi < slen) ? A.qual_table[q[qi]] : NAN;
        }
        if (A.out.model_state && s) {
            // revseq = sequence[::-1].replace('U', 'T'), '__' + revseq + '__' for flip-flop
            for (int j = 0; j < 5; j++) {
                int64_t k = p + j - (flipflop ? 2 : 0);
                uint8_t ch = '_';
                if (k >= 0 && k < slen) { ch = s[slen - 1 - k]; if (ch == 'U') ch = 'T'; }
                else if (!flipflop) ch = 0;                      // slice past the end: shorter string
                A.out.model_state[(e0 + i) * 5 + j] = ch;
            }
        }
    }
}

int launch_derive_events(pb2_context *ctx, const pb2_batch &b, const pb2_event_tables &ev,
                         const pb2_basecalls *bc, const float *scale_shift,
                         const pb2_event_columns &out, cudaStream_t st)
{
    const int64_t n = b.n_reads;
    if (n <= 0) return PB2_OK;
    if (!ev.event_offsets || !ev.first_sample || ev.block_stride < 1 || ev.block_stride > 32)
        return fail(ctx, PB2_EINVAL, "derive_event_tables: need event_offsets, first_sample and "
                                     "1 <= block_stride <= 32");
    if ((out.pos || out.p_model_state || out.model_state) && !ev.move)
        return fail(ctx, PB2_EINVAL, "derive_event_tables: pos / p_model_state / model_state need move");
    if ((out.p_model_state || out.model_state) && !out.pos)
        return fail(ctx, PB2_EINVAL, "derive_event_tables: p_model_state / model_state need the pos column");
    EventArgs A = {};
    A.raw = b.raw; A.raw_off = b.raw_offsets; A.raw_len = b.raw_lengths;
    A.range = b.range; A.digitisation = b.digitisation; A.offset = b.offset;
    A.ev_off = ev.event_offsets; A.first_sample = ev.first_sample; A.move = ev.move;
    A.stride = ev.block_stride; A.n = n; A.total = ev.n_events_total;
    if (bc) { A.qstring = bc->qstring; A.sequence = bc->sequence; A.seq_off = bc->seq_offsets; A.qual_table = bc->qual_table; }
    A.scale_shift = scale_shift;
    A.out = out;
    if (A.total > 0 && (out.mean || out.stdv || out.scaled_mean || out.start || out.end || out.length)) {
        PB_LAUNCH(ctx, K_EVENT_MEANS, "k_event_stats", st,
            k_event_stats<<<(unsigned)((A.total + 127) / 128), 128, 0, st>>>(A));
    }
    if (out.pos || out.error) {
        PB_LAUNCH(ctx, K_EVENT_POS, "k_event_pos", st,
            k_event_pos<<<(unsigned)((n * 32 + 127) / 128), 128, 0, st>>>(A));
    }
    return PB2_OK;
}

}  // namespace pb
