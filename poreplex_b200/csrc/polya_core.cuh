// polya_core.cuh -- poly(A) dwell measurement for ONE read, as a streaming state machine.
//
// Reference: poreplex/polya.py (PolyASignalAnalyzer.__call__, call_polya,
// try_recalibrate_shifted_signal, calc_internal_polya_stdv, find_best_polya_interval) and
// its native helper csupport.detect_events -> src/contrib/scrappie/event_detection.c.
// Semantics (float32 Series arithmetic, NEP-50 comparisons, numpy pairwise sums, int64
// scores) follow oracle/polya_restated.py, which is checked against the reference's own
// polya.py running verbatim.
//
// Shape of the computation.  The reference materialises the window's signal, prefix
// sums, t-statistics, event table and two E x E score matrices.  Here one thread walks
// the window as a stream and keeps O(1) state:
//   raw int16 -> pA (fp64) -> scaled f32 -> 7-tap median (zero padded)
//     -> fp64 prefix sums held in a 64-entry ring -> short/long t-statistics
//     -> two-detector peak state machine -> events, one at a time (EventStream::next)
// The O(E^2) interval search is an exact O(E) scan (same first-row-major-maximum rule);
// sums that the reference takes with numpy's pairwise algorithm are reproduced by a
// push-style evaluator that is told n up front (PairwiseSum).  Values needed after the
// search (interval sums, spikes, the longest event's sample stdv) are recomputed by
// re-streaming the window, so no per-event storage exists at all.
//
// Everything is __host__ __device__ so that tests/hostcheck can run this exact code on
// the CPU against the oracle without a GPU.
#pragma once
#include <cfloat>
#include "pb_math.cuh"

namespace pb {

struct PolyaParams {
    int32_t stride;                  // rough_signal_stride (15)
    int32_t refinement_expansion;    // 200
    int32_t openend_unit;            // openend_expansion // stride (66)
    int32_t max_extension;           // maximum_openend_extension (50)
    int32_t w1, w2;                  // event_detection window lengths (7, 20)
    float thr1, thr2, peak_height;   // 3, 8, 4
    float cutoff_lo, cutoff_hi;      // f32(polya_mean_cutoff)
    float mean_loc;                  // f32(polya_mean_dist[0])
    float trigger;                   // f32(polya_mean_trigger_recalibration * sd)
    float half_range;                // f32(sd * z_cutoff)
    float stdv_max;                  // f32(polya_stdv_max)
    double stdv_lo, stdv_hi;         // polya_stdv_range
    int32_t spike_tolerance;         // 110
    double spike_weight;             // 1.5
    int32_t recal_max_dist;          // 100
    float recal_min_length;          // 750
    float recal_max_stdv;            // 5
};

constexpr int POLYA_MAX_SPIKES = 48;

struct PolyaResult {
    int32_t found;                   // 1: set_polya_tail was called
    int32_t n_spikes;                // may exceed POLYA_MAX_SPIKES (then truncated)
    int64_t begin, end;              // raw-sample coordinates in the read
    int64_t dwell_samples;           // dwell_time = dwell_samples / sampling_rate
    int32_t extensions;              // open-end extensions used (diagnostic)
    int32_t flags;                   // bit 0: anchor buffer overflow
    float spikes[POLYA_MAX_SPIKES][4];   // (length, mean[k-1], mean[k], mean[k+1])
};

// ---- numpy pairwise float32 sum, push style (n known up front) ------------------
struct PairwiseSum {
    struct Frame { int64_t right_n; float left; int have_left; };
    Frame stack[40];
    int depth;
    int64_t leaf_n, leaf_i;
    float r[8], res;
    bool finished;
    float total;

    PB_HD void descend(int64_t n) {
        while (n > 128) {
            int64_t n2 = n / 2;
            n2 -= n2 % 8;
            stack[depth].right_n = n - n2;
            stack[depth].have_left = 0;
            stack[depth].left = 0.f;
            depth++;
            n = n2;
        }
        leaf_n = n;
        leaf_i = 0;
        res = 0.f;
    }
    PB_HD void begin(int64_t n) {
        depth = 0;
        finished = (n <= 0);
        total = 0.f;
        if (!finished) descend(n);
    }
    PB_HD void leaf_done(float v) {
        for (;;) {
            if (depth == 0) { total = v; finished = true; return; }
            Frame &f = stack[depth - 1];
            if (!f.have_left) {
                f.left = v;
                f.have_left = 1;
                descend(f.right_n);
                return;
            }
            v = pb::fadd(f.left, v);
            depth--;
        }
    }
    PB_HD void push(float x) {
        if (finished) return;
        const int64_t n = leaf_n, i = leaf_i;
        if (n < 8) {
            res = pb::fadd(res, x);
        } else if (i < 8) {
            r[i] = x;
        } else if (i < n - (n % 8)) {
            r[i & 7] = pb::fadd(r[i & 7], x);
        } else {
            if (i == n - (n % 8))
                res = pb::fadd(pb::fadd(pb::fadd(r[0], r[1]), pb::fadd(r[2], r[3])),
                               pb::fadd(pb::fadd(r[4], r[5]), pb::fadd(r[6], r[7])));
            res = pb::fadd(res, x);
        }
        leaf_i = i + 1;
        if (leaf_i == n) {
            if (n >= 8 && (n % 8) == 0)
                res = pb::fadd(pb::fadd(pb::fadd(r[0], r[1]), pb::fadd(r[2], r[3])),
                               pb::fadd(pb::fadd(r[4], r[5]), pb::fadd(r[6], r[7])));
            leaf_done(res);
        }
    }
};

// ---- median of 7 (scipy.signal.medfilt kernel 7 = 4th smallest) ------------------
PB_HD void cswap(float &a, float &b) { const float lo = a < b ? a : b; b = a < b ? b : a; a = lo; }
PB_HD float median7(float p0, float p1, float p2, float p3, float p4, float p5, float p6) {
    // 13-exchange median-of-7 network (checked exhaustively in tests/test_hostcheck_cpu.py)
    cswap(p0, p5); cswap(p0, p3); cswap(p1, p6);
    cswap(p2, p4); cswap(p0, p1); cswap(p3, p5);
    cswap(p2, p6); cswap(p2, p3); cswap(p3, p6);
    cswap(p4, p5); cswap(p1, p4); cswap(p1, p3);
    cswap(p3, p4);
    return p3;
}

struct Event {
    uint64_t start;
    float length, mean, stdv;
    int64_t end;          // (int64)(start + length) as events['end']
};

// ---- windowed signal source: raw -> pA -> scale -> medfilt(7) ---------------------
struct WindowSource {
    const int16_t *raw;   // read base
    double gain, offset;
    float scale, shift;
    int64_t w0, n;        // window [w0, w0 + n) in read coordinates
    int64_t next;         // index (window coords) of the next filtered sample to emit
    float y[7];           // y[k] = scaled sample at window index next - 3 + k (0 outside)

    PB_HD float scaled_at(int64_t i) const {       // i in window coords; zero padding
        if (i < 0 || i >= n) return 0.0f;
        const float pa = pb::dac_to_pa((int)raw[w0 + i], gain, offset);
        return pb::fadd(pb::fmul(scale, pa), shift);
    }
    PB_HD void seek(int64_t i) {
        next = i;
        for (int k = 0; k < 7; k++) y[k] = scaled_at(i - 3 + k);
    }
    PB_HD float pop() {                             // medfilt7 value at index `next`
        const float m = median7(y[0], y[1], y[2], y[3], y[4], y[5], y[6]);
        for (int k = 0; k < 6; k++) y[k] = y[k + 1];
        next++;
        y[6] = scaled_at(next + 3);
        return m;
    }
};

// ---- a float32 signal as it is: the input csupport.detect_events itself takes ------
struct PlainSource {
    const float *x;
    int64_t n;
    int64_t next;
    PB_HD void seek(int64_t i) { next = i; }
    PB_HD float pop() { return x[next++]; }       // EventStream never pops past n
};

// ---- event_detection.c as an iterator --------------------------------------------
struct Detector {
    float threshold;
    int64_t window, masked_to, peak_pos;
    float peak_value;
    int valid;
    double snapS, snapQ;          // prefix sums at peak_pos
};

// RING: entries of the prefix-sum ring, a power of two >= 2 * max(window) + 2 (index i needs
// S[i - w] .. S[i + w] and the fill runs one ahead).
// The rings live OUTSIDE the stream object (EventRings, handed over with use()): they are indexed
// dynamically, and an object with one dynamically indexed member is kept in local memory as a whole
// -- with the rings as members every scalar of the detector state (indices, peak machines, window
// source) was a local-memory load / store on each use.
template <int RING = 64>
struct EventRings { double S[RING], Q[RING]; };

template <class Source, int RING = 64>
struct EventStreamT {
    static constexpr int64_t MASK = RING - 1;
    Source src;
    int64_t n;
    int64_t head;                 // prefix sums S[0..head] are in the ring
    double *S, *Q;                // EventRings<RING> of the caller
    PB_HD void use(EventRings<RING> &r) { S = r.S; Q = r.Q; }
    int64_t det_i;
    Detector d[2];
    float peak_height;
    int64_t w[2], wmax;
    // event emission
    uint64_t prev_pos; double prevS, prevQ;
    Event pend_ev[2]; int n_pend;          // only used by next()
    int64_t n_peaks;
    bool tail_emitted;

    PB_HD void begin(const Source &s, const PolyaParams &P) {
        src = s;
        src.seek(0);
        n = s.n;
        head = 0;
        S[0] = 0.0; Q[0] = 0.0;
        det_i = 0;
        w[0] = P.w1; w[1] = P.w2;
        wmax = P.w1 > P.w2 ? P.w1 : P.w2;
        peak_height = P.peak_height;
        for (int k = 0; k < 2; k++) {
            d[k].threshold = k ? P.thr2 : P.thr1;
            d[k].window = w[k];
            d[k].masked_to = 0;
            d[k].peak_pos = -1;
            d[k].peak_value = FLT_MAX;
            d[k].valid = 0;
            d[k].snapS = d[k].snapQ = 0.0;
        }
        prev_pos = 0; prevS = 0.0; prevQ = 0.0;
        n_pend = 0;
        n_peaks = 0;
        tail_emitted = false;
    }
    PB_HD void fill_to(int64_t idx) {              // make S[idx] available
        if (idx > n) idx = n;
        while (head < idx) {
            const float m = src.pop();
            const double s = pb::dadd(S[head & MASK], (double)m);
            const double q = pb::dadd(Q[head & MASK], (double)pb::fmul(m, m));
            head++;
            S[head & MASK] = s;
            Q[head & MASK] = q;
        }
    }
    // compute_tstat (event_detection.c:61-117) at index i for window wl
    PB_HD float tstat(int64_t i, int64_t wl) const {
        if (n < 2 * wl || wl < 2) return 0.0f;
        if (i < wl || i > n - wl) return 0.0f;
        const float wf = (float)wl;
        double sum1 = S[i & MASK], sumsq1 = Q[i & MASK];
        if (i > wl) {
            sum1 = pb::dsub(sum1, S[(i - wl) & MASK]);
            sumsq1 = pb::dsub(sumsq1, Q[(i - wl) & MASK]);
        }
        const float sum2 = (float)pb::dsub(S[(i + wl) & MASK], S[i & MASK]);
        const float sumsq2 = (float)pb::dsub(Q[(i + wl) & MASK], Q[i & MASK]);
        const float mean1 = (float)pb::ddiv(sum1, (double)wf);
        const float mean2 = pb::fdiv(sum2, wf);
        double cv = pb::dsub(pb::ddiv(sumsq1, (double)wf), (double)pb::fmul(mean1, mean1));
        cv = pb::dadd(cv, (double)pb::fdiv(sumsq2, wf));
        cv = pb::dsub(cv, (double)pb::fmul(mean2, mean2));
        float combined = (float)cv;
        combined = fmaxf(combined, FLT_MIN);
        const float dm = pb::fsub(mean2, mean1);
        return (float)pb::ddiv(fabs((double)dm), sqrt((double)pb::fdiv(combined, wf)));
    }
    PB_HD void make_event(uint64_t s, double sS, double sQ, uint64_t e, double eS, double eQ,
                          Event &ev) const {
        // create_event (event_detection.c:216-236); size_t arithmetic kept
        ev.start = s;
        ev.length = (float)(uint64_t)(e - s);
        ev.mean = pb::fdiv((float)pb::dsub(eS, sS), ev.length);
        const float deltasqr = (float)pb::dsub(eQ, sQ);
        const float var = pb::fsub(pb::fdiv(deltasqr, ev.length), pb::fmul(ev.mean, ev.mean));
        ev.stdv = sqrtf(fmaxf(var, 0.0f));
        ev.end = (int64_t)pb::dadd((double)s, (double)ev.length);
    }
    PB_HD bool finished() const { return tail_emitted; }

    // Advance by ONE detector index (or, after the last index, emit the closing event).
    // Writes the 0..2 events completed by this step to out[] in table order and returns
    // their count.  Stepping sample by sample keeps the 32 reads of a warp in lock step;
    // the consumers below are written as "for every step, for every event" loops.
    PB_HD int step(Event *out) {
        if (det_i >= n) {
            if (tail_emitted) return 0;
            tail_emitted = true;
            fill_to(n);
            if (n_peaks == 0) {
                // create_events with no peak: create_event(0, peaks[0] = 0)
                make_event(0, 0.0, 0.0, 0, 0.0, 0.0, out[0]);
            } else {
                make_event(prev_pos, prevS, prevQ, (uint64_t)n, S[n & MASK], Q[n & MASK], out[0]);
            }
            return 1;
        }
        // short_long_peak_detector, one index (event_detection.c:124-201)
        const int64_t i = det_i++;
        fill_to(i + wmax + 1);
        int nout = 0;
        for (int k = 0; k < 2; k++) {
            Detector &D = d[k];
            if (D.masked_to >= i) continue;
            const float cur = tstat(i, w[k]);
            if (D.peak_pos == -1) {
                if (cur < D.peak_value) {
                    D.peak_value = cur;
                } else if (pb::fsub(cur, D.peak_value) > peak_height) {
                    D.peak_value = cur;
                    D.peak_pos = i;
                    D.snapS = S[i & MASK]; D.snapQ = Q[i & MASK];
                }
            } else {
                if (cur > D.peak_value) {
                    D.peak_value = cur;
                    D.peak_pos = i;
                    D.snapS = S[i & MASK]; D.snapQ = Q[i & MASK];
                }
                if (k == 0 && D.peak_value > D.threshold) {
                    d[1].masked_to = D.peak_pos + D.window;
                    d[1].peak_pos = -1;
                    d[1].peak_value = FLT_MAX;
                    d[1].valid = 0;
                }
                if (pb::fsub(D.peak_value, cur) > peak_height && D.peak_value > D.threshold)
                    D.valid = 1;
                if (D.valid && (i - D.peak_pos) > D.window / 2) {
                    make_event(prev_pos, prevS, prevQ, (uint64_t)D.peak_pos, D.snapS, D.snapQ,
                               out[nout]);
                    nout++;
                    prev_pos = (uint64_t)D.peak_pos; prevS = D.snapS; prevQ = D.snapQ;
                    n_peaks++;
                    D.peak_pos = -1;
                    D.peak_value = cur;
                    D.valid = 0;
                }
            }
        }
        return nout;
    }

    // pull-style access (host tests): next event in table order, false when exhausted
    PB_HD bool next(Event &ev) {
        for (;;) {
            if (n_pend > 0) {
                ev = pend_ev[0];
                pend_ev[0] = pend_ev[1];
                n_pend--;
                return true;
            }
            if (finished()) return false;
            n_pend = step(pend_ev);
        }
    }
};

using EventStream = EventStreamT<WindowSource>;      // poly(A): raw -> pA -> scale -> medfilt(7)
template <int RING> using PlainEventStream = EventStreamT<PlainSource, RING>;   // csupport.detect_events(signal)

// ---- event iteration with a replay cache --------------------------------------------
// The first complete walk over a window records its events (16 bytes each) in a per-read
// slice of a global scratch; later walks over the same window replay them instead of
// re-running median filter, prefix sums, t-statistics and the peak detector.  A window
// with more events than the slice holds simply keeps streaming.
struct EventCacheSlot { uint32_t start; float length, mean, stdv; };

struct EvIter {
    EventStream es;
    EventCacheSlot *cache;        // element k of this read at cache[k * stride]
    int64_t stride;
    int cap, count, pos;
    bool valid, replay, overflow;

    PB_HD void attach(EventCacheSlot *c, int64_t stride_, int cap_) {
        cache = c; stride = stride_; cap = c ? cap_ : 0; count = 0; valid = false;
    }
    PB_HD void invalidate() { valid = false; }
    PB_HD void begin(const WindowSource &src, const PolyaParams &P) {
        pos = 0;
        replay = valid;
        if (!replay) { es.begin(src, P); count = 0; overflow = (cap == 0); }
    }
    PB_HD bool finished() const { return replay ? pos >= count : es.finished(); }
    PB_HD int step(Event *out) {
        if (replay) {
            const EventCacheSlot c = cache[(int64_t)pos * stride];
            pos++;
            out[0].start = c.start; out[0].length = c.length; out[0].mean = c.mean;
            out[0].stdv = c.stdv;
            out[0].end = (int64_t)pb::dadd((double)c.start, (double)c.length);
            return 1;
        }
        const int k = es.step(out);
        for (int q = 0; q < k; q++) {
            if (count < cap) {
                EventCacheSlot c;
                c.start = (uint32_t)out[q].start; c.length = out[q].length;
                c.mean = out[q].mean; c.stdv = out[q].stdv;
                cache[(int64_t)count * stride] = c;
            } else {
                overflow = true;
            }
            count++;
        }
        if (es.finished()) valid = !overflow;
        return k;
    }
};

PB_HD bool between_f32(float x, float lo, float hi) { return x >= lo && x <= hi; }

// ---- the whole of PolyASignalAnalyzer for one read --------------------------------
// rough_begin / rough_end: pooled-sample range from the segmentation (rough_end < 0 =
// None: no polya-tail state, polya.py:53-56,69-70).
//
// Two formulations of the same computation.  polya_analyze_nested follows the reference's
// control flow literally: one loop per walk over the window's events (anchors, count, sum,
// interval search, second pass), each with its own copy of the event-stream step.  On the GPU
// that costs a warp dearly: the reads of a warp take different routes (12 % recalibrate first,
// some extend the window), every route runs its walks at its own code address, and the warp
// executes the routes one after the other -- ncu: 9.9 of 32 lanes active per instruction, 37 %
// of all instructions issued by a copy of the detector that 4 lanes were in.  polya_analyze
// (below it) is the same computation as ONE loop with ONE event-stream step: which walk a read
// is in is data (`walk`), so all lanes that are walking step together whatever their route, and
// the sample loops of calc_internal_polya_stdv are two more walk kinds of the same loop.  Both
// are kept: tests/hostcheck requires them to agree with the oracle and with each other.
PB_HD void polya_analyze_nested(const PolyaParams &P, const int16_t *raw, int64_t full_length,
                         double gain, double offset, float scale, float shift,
                         int32_t rough_begin, int32_t rough_end_in, PolyaResult &R,
                         EventCacheSlot *cache = nullptr, int64_t cache_stride = 1,
                         int cache_cap = 0)
{
    R.found = 0; R.n_spikes = 0; R.begin = 0; R.end = 0; R.dwell_samples = 0;
    R.extensions = 0; R.flags = 0;
    const int64_t stride = P.stride;
    int64_t rough_end_cur = rough_end_in;         // < 0 : None
    bool have_range = false;
    float lo = P.cutoff_lo, hi = P.cutoff_hi;
    int ext_depth = 0;
    EvIter es;
    EventRings<64> rings;
    es.es.use(rings);
    es.attach(cache, cache_stride, cache_cap);
    Event evs[2];

    for (;;) {                                    // one iteration per __call__ (window)
        int64_t rough_end = rough_end_cur;
        if (rough_end < 0 || rough_end - rough_begin < P.openend_unit)
            rough_end = (int64_t)rough_begin + P.openend_unit;
        int64_t insp_begin = (int64_t)rough_begin * stride - P.refinement_expansion;
        if (insp_begin < 0) insp_begin = 0;
        int64_t insp_end = (rough_end + 1) * stride + P.refinement_expansion;
        if (insp_end > full_length) insp_end = full_length;
        const int64_t adapter_end = (int64_t)rough_begin * stride - insp_begin;
        WindowSource src;
        src.raw = raw; src.gain = gain; src.offset = offset; src.scale = scale; src.shift = shift;
        src.w0 = insp_begin; src.n = insp_end - insp_begin;
        if (src.n <= 0) return;                   // csupport raises on an empty signal
        es.invalidate();                          // new window: recorded events are stale
        if (!have_range) { lo = P.cutoff_lo; hi = P.cutoff_hi; }
        bool recal_mode = rough_end_cur < 0;
        bool extend = false;
        int guard = 0;

        for (;;) {                                // call_polya / try_recalibrate ping-pong
            if (++guard > 8) return;              // cannot happen; never spin
            if (recal_mode) {
                // try_recalibrate_shifted_signal (polya.py:127-148)
                float a_ml[64], a_len[64];
                int na = 0;
                es.begin(src, P);
                while (!es.finished()) {
                    const int ne = es.step(evs);
                    for (int q = 0; q < ne; q++) {
                        const Event &ev = evs[q];
                        if ((int64_t)ev.start <= adapter_end + P.recal_max_dist &&
                            ev.end > adapter_end && ev.stdv < P.recal_max_stdv) {
                            if (na < 64) { a_ml[na] = pb::fmul(ev.mean, ev.length); a_len[na] = ev.length; }
                            na++;
                        }
                    }
                }
                if (na == 0) return;
                if (na > 64) { R.flags |= 1; return; }
                PairwiseSum s1, s2;
                s1.begin(na); s2.begin(na);
                for (int k = 0; k < na; k++) { s1.push(a_ml[k]); s2.push(a_len[k]); }
                const float pm = pb::fdiv(s1.total, s2.total);
                lo = pb::fsub(pm, P.half_range);
                hi = pb::fadd(pm, P.half_range);
                have_range = true;
                // events[is_polya]['length'].sum() >= min_length
                int64_t npol = 0;
                es.begin(src, P);
                while (!es.finished()) {
                    const int ne = es.step(evs);
                    for (int q = 0; q < ne; q++) npol += between_f32(evs[q].mean, lo, hi);
                }
                PairwiseSum sl;
                sl.begin(npol);
                es.begin(src, P);
                while (!es.finished()) {
                    const int ne = es.step(evs);
                    for (int q = 0; q < ne; q++)
                        if (between_f32(evs[q].mean, lo, hi)) sl.push(evs[q].length);
                }
                if (!(sl.total >= P.recal_min_length)) return;
                recal_mode = false;
            }
            // ---- find_best_polya_interval as an O(E) scan (polya.py:156-187)
            int64_t best = 0, best_i = -1, best_j = -1, best_npol = 0;
            int64_t n_events = 0;
            {
                bool alive = false;
                int64_t Sv = 0, minP = 0, minI = -1, minCnt = 0, Pfx = 0, cnt = 0;
                es.begin(src, P);
                int64_t j = 0;
                while (!es.finished()) {
                  const int ne = es.step(evs);
                  for (int q = 0; q < ne; q++) {
                    const Event &ev = evs[q];
                    const bool ip = between_f32(ev.mean, lo, hi);
                    const double L = (double)ev.length;
                    const double v = ip ? L : -L;
                    const int64_t m = (v > 0) ? (int64_t)v : (int64_t)pb::dmul(v, P.spike_weight);
                    const int64_t s = ip ? 1 : (int64_t)(-L);
                    const int64_t Pprev = Pfx, cprev = cnt;
                    Pfx += m;
                    cnt += ip;
                    if (alive) {
                        Sv = (Sv < 0) ? -1 : (s > 0 ? P.spike_tolerance : Sv + s);
                        if (Sv < 0) alive = false;
                    }
                    const int64_t Sjj = (s > 0) ? P.spike_tolerance : s;
                    if (Sjj >= 0) {
                        if (!alive) { alive = true; Sv = Sjj; minP = Pprev; minI = j; minCnt = cprev; }
                        else if (Pprev < minP) { minP = Pprev; minI = j; minCnt = cprev; }
                    }
                    if (alive && Sv > 0) {
                        const int64_t val = Pfx - minP;
                        if (val > best) { best = val; best_i = minI; best_j = j; best_npol = cnt - minCnt; }
                    }
                    j++;
                  }
                }
                n_events = j;
            }
            const bool has_best = best > 0;
            if (has_best && best_j == n_events - 1 && insp_end < full_length &&
                ext_depth < P.max_extension) {
                extend = true;
                break;
            }
            if (!has_best) { recal_mode = true; continue; }
            // ---- second pass over the chosen interval
            const int64_t n_int = best_j - best_i + 1;
            PairwiseSum s_ml, s_len, s_dw;
            s_ml.begin(n_int); s_len.begin(n_int); s_dw.begin(best_npol);
            uint64_t long_start = 0; float long_len = -1.0f;
            uint64_t first_start = 0; int64_t last_end = 0;
            int nsp = 0;
            float prev_mean = 0.f;
            int pending_spike = -1;               // spike waiting for its right neighbour
            {
                es.begin(src, P);
                int64_t j = 0;
                bool stop = false;
                while (!es.finished() && !stop) {
                  const int ne = es.step(evs);
                  for (int q = 0; q < ne; q++) {
                    const Event &ev = evs[q];
                    if (j > best_j) { stop = true; break; }
                    if (j >= best_i) {
                        const bool ip = between_f32(ev.mean, lo, hi);
                        s_ml.push(pb::fmul(ev.mean, ev.length));
                        s_len.push(ev.length);
                        if (ip) s_dw.push(ev.length);
                        if (ev.length > long_len) { long_len = ev.length; long_start = ev.start; }
                        if (j == best_i) first_start = ev.start;
                        if (j == best_j)
                            last_end = (int64_t)pb::dadd((double)ev.start, (double)ev.length);
                        if (pending_spike >= 0) {
                            if (pending_spike < POLYA_MAX_SPIKES) R.spikes[pending_spike][3] = ev.mean;
                            pending_spike = -1;
                        }
                        if (!ip) {
                            if (nsp < POLYA_MAX_SPIKES) {
                                R.spikes[nsp][0] = ev.length;
                                R.spikes[nsp][1] = (j > best_i) ? prev_mean : NAN;
                                R.spikes[nsp][2] = ev.mean;
                                R.spikes[nsp][3] = NAN;
                            }
                            pending_spike = nsp;
                            nsp++;
                        }
                        prev_mean = ev.mean;
                    }
                    j++;
                  }
                }
            }
            if (!have_range) {
                // is_polya_signal_shifted (polya.py:88-93)
                const float lvl = pb::fdiv(s_ml.total, s_len.total);
                if (fabsf(pb::fsub(lvl, P.mean_loc)) > P.trigger) { recal_mode = true; continue; }
            }
            // calc_internal_polya_stdv of the longest event (polya.py:150-154)
            const int64_t ilen = (int64_t)long_len;
            const int64_t sb = (int64_t)pb::dadd((double)long_start, pb::dmul((double)ilen, P.stdv_lo));
            const int64_t se = (int64_t)pb::dadd((double)long_start, pb::dmul((double)ilen, P.stdv_hi));
            bool sd_ok = false;
            if (se - sb > 2) {
                const int64_t b = sb < 0 ? 0 : sb, e = se > src.n ? src.n : se;   // numpy slicing clips
                const int64_t cntn = e - b;
                if (cntn > 0) {
                    WindowSource ws = src;
                    PairwiseSum sm;
                    sm.begin(cntn);
                    ws.seek(b);
                    for (int64_t q = 0; q < cntn; q++) sm.push(ws.pop());
                    const float mu = pb::fdiv(sm.total, (float)cntn);
                    PairwiseSum sv;
                    sv.begin(cntn);
                    ws.seek(b);
                    for (int64_t q = 0; q < cntn; q++) {
                        const float dx = pb::fsub(ws.pop(), mu);
                        sv.push(pb::fmul(dx, dx));
                    }
                    const float sd = sqrtf(pb::fdiv(sv.total, (float)cntn));
                    sd_ok = sd < P.stdv_max;
                }
            }
            if (sd_ok) {
                R.found = 1;
                R.begin = (int64_t)first_start + insp_begin;
                R.end = last_end + insp_begin;
                R.dwell_samples = (int64_t)s_dw.total;
                R.n_spikes = nsp;
                R.extensions = ext_depth;
                return;
            }
            if (!have_range) { recal_mode = true; continue; }
            return;
        }
        if (!extend) return;
        rough_end_cur = rough_end + P.openend_unit;
        ext_depth++;
    }
}

// ---- the same as one loop (see above) ------------------------------------------------
PB_HD void polya_analyze(const PolyaParams &P, const int16_t *raw, int64_t full_length,
                         double gain, double offset, float scale, float shift,
                         int32_t rough_begin, int32_t rough_end_in, PolyaResult &R,
                         EventCacheSlot *cache = nullptr, int64_t cache_stride = 1,
                         int cache_cap = 0)
{
    R.found = 0; R.n_spikes = 0; R.begin = 0; R.end = 0; R.dwell_samples = 0;
    R.extensions = 0; R.flags = 0;
    const int64_t stride = P.stride;
    int64_t rough_end_cur = rough_end_in;         // < 0 : None
    bool have_range = false;
    float lo = P.cutoff_lo, hi = P.cutoff_hi;
    int ext_depth = 0;
    EvIter es;
    EventRings<64> rings;
    es.es.use(rings);
    es.attach(cache, cache_stride, cache_cap);
    Event evs[2];

    // what the loop does next: set up a window, start a call_polya / recalibration round, or
    // advance the walk in progress by one event-stream step (W_ANCHORS .. W_SECOND) or by one
    // sample (W_SD_MEAN, W_SD_VAR)
    enum { DO_WINDOW = 0, DO_ROUND, W_ANCHORS, W_COUNT, W_SUM, W_FIND, W_SECOND, W_SD_MEAN, W_SD_VAR };
    int walk = DO_WINDOW;

    // window
    WindowSource src;
    src.raw = raw; src.gain = gain; src.offset = offset; src.scale = scale; src.shift = shift;
    src.w0 = 0; src.n = 0; src.next = 0;
    for (int k = 0; k < 7; k++) src.y[k] = 0.f;
    int64_t rough_end = 0, insp_begin = 0, insp_end = 0, adapter_end = 0;
    bool recal_mode = false;
    int guard = 0;
    // try_recalibrate_shifted_signal
    float a_ml[64], a_len[64];
    int na = 0;
    int64_t npol = 0;
    PairwiseSum sl;
    // find_best_polya_interval
    int64_t best = 0, best_i = -1, best_j = -1, best_npol = 0, j = 0;
    bool alive = false;
    int64_t Sv = 0, minP = 0, minI = -1, minCnt = 0, Pfx = 0, cnt = 0;
    // second pass
    PairwiseSum s_ml, s_len, s_dw;
    uint64_t long_start = 0; float long_len = -1.0f;
    uint64_t first_start = 0; int64_t last_end = 0;
    int nsp = 0;
    float prev_mean = 0.f;
    int pending_spike = -1;
    bool stop = false;
    // calc_internal_polya_stdv
    WindowSource ws = src;
    PairwiseSum sd_sum;
    int64_t sd_b = 0, sd_n = 0, sd_q = 0;
    float sd_mu = 0.f;

    for (;;) {
        if (walk >= W_ANCHORS && walk <= W_SECOND) {
            if (!es.finished() && !stop) {
                // ---- one step of the event stream, events to the walk's consumer
                const int ne = es.step(evs);
                for (int q = 0; q < ne; q++) {
                    const Event &ev = evs[q];
                    if (walk == W_FIND) {
                        // find_best_polya_interval as an O(E) scan (polya.py:156-187)
                        const bool ip = between_f32(ev.mean, lo, hi);
                        const double L = (double)ev.length;
                        const double v = ip ? L : -L;
                        const int64_t m = (v > 0) ? (int64_t)v : (int64_t)pb::dmul(v, P.spike_weight);
                        const int64_t s = ip ? 1 : (int64_t)(-L);
                        const int64_t Pprev = Pfx, cprev = cnt;
                        Pfx += m;
                        cnt += ip;
                        if (alive) {
                            Sv = (Sv < 0) ? -1 : (s > 0 ? P.spike_tolerance : Sv + s);
                            if (Sv < 0) alive = false;
                        }
                        const int64_t Sjj = (s > 0) ? P.spike_tolerance : s;
                        if (Sjj >= 0) {
                            if (!alive) { alive = true; Sv = Sjj; minP = Pprev; minI = j; minCnt = cprev; }
                            else if (Pprev < minP) { minP = Pprev; minI = j; minCnt = cprev; }
                        }
                        if (alive && Sv > 0) {
                            const int64_t val = Pfx - minP;
                            if (val > best) { best = val; best_i = minI; best_j = j; best_npol = cnt - minCnt; }
                        }
                        j++;
                    } else if (walk == W_SECOND) {
                        if (j > best_j) { stop = true; break; }
                        if (j >= best_i) {
                            const bool ip = between_f32(ev.mean, lo, hi);
                            s_ml.push(pb::fmul(ev.mean, ev.length));
                            s_len.push(ev.length);
                            if (ip) s_dw.push(ev.length);
                            if (ev.length > long_len) { long_len = ev.length; long_start = ev.start; }
                            if (j == best_i) first_start = ev.start;
                            if (j == best_j)
                                last_end = (int64_t)pb::dadd((double)ev.start, (double)ev.length);
                            if (pending_spike >= 0) {
                                if (pending_spike < POLYA_MAX_SPIKES) R.spikes[pending_spike][3] = ev.mean;
                                pending_spike = -1;
                            }
                            if (!ip) {
                                if (nsp < POLYA_MAX_SPIKES) {
                                    R.spikes[nsp][0] = ev.length;
                                    R.spikes[nsp][1] = (j > best_i) ? prev_mean : NAN;
                                    R.spikes[nsp][2] = ev.mean;
                                    R.spikes[nsp][3] = NAN;
                                }
                                pending_spike = nsp;
                                nsp++;
                            }
                            prev_mean = ev.mean;
                        }
                        j++;
                    } else if (walk == W_ANCHORS) {
                        // try_recalibrate_shifted_signal (polya.py:127-148): anchor events
                        if ((int64_t)ev.start <= adapter_end + P.recal_max_dist &&
                            ev.end > adapter_end && ev.stdv < P.recal_max_stdv) {
                            if (na < 64) { a_ml[na] = pb::fmul(ev.mean, ev.length); a_len[na] = ev.length; }
                            na++;
                        }
                    } else if (walk == W_COUNT) {
                        npol += between_f32(ev.mean, lo, hi);
                    } else {                      // W_SUM
                        if (between_f32(ev.mean, lo, hi)) sl.push(ev.length);
                    }
                }
                continue;
            }
            // ---- the walk is over: what follows it in the reference's control flow
            if (walk == W_ANCHORS) {
                if (na == 0) return;
                if (na > 64) { R.flags |= 1; return; }
                PairwiseSum s1, s2;
                s1.begin(na); s2.begin(na);
                for (int k = 0; k < na; k++) { s1.push(a_ml[k]); s2.push(a_len[k]); }
                const float pm = pb::fdiv(s1.total, s2.total);
                lo = pb::fsub(pm, P.half_range);
                hi = pb::fadd(pm, P.half_range);
                have_range = true;
                // events[is_polya]['length'].sum() >= min_length: count, then sum
                npol = 0;
                walk = W_COUNT;
                es.begin(src, P);
            } else if (walk == W_COUNT) {
                sl.begin(npol);
                walk = W_SUM;
                es.begin(src, P);
            } else if (walk == W_SUM) {
                if (!(sl.total >= P.recal_min_length)) return;
                recal_mode = false;
                best = 0; best_i = -1; best_j = -1; best_npol = 0; j = 0;
                alive = false; Sv = 0; minP = 0; minI = -1; minCnt = 0; Pfx = 0; cnt = 0;
                walk = W_FIND;
                es.begin(src, P);
            } else if (walk == W_FIND) {
                const int64_t n_events = j;
                const bool has_best = best > 0;
                if (has_best && best_j == n_events - 1 && insp_end < full_length &&
                    ext_depth < P.max_extension) {
                    rough_end_cur = rough_end + P.openend_unit;      // extend the window
                    ext_depth++;
                    walk = DO_WINDOW;
                } else if (!has_best) {
                    recal_mode = true;
                    walk = DO_ROUND;
                } else {
                    // second pass over the chosen interval
                    const int64_t n_int = best_j - best_i + 1;
                    s_ml.begin(n_int); s_len.begin(n_int); s_dw.begin(best_npol);
                    long_start = 0; long_len = -1.0f; first_start = 0; last_end = 0;
                    nsp = 0; prev_mean = 0.f; pending_spike = -1;
                    j = 0; stop = false;
                    walk = W_SECOND;
                    es.begin(src, P);
                }
            } else {                              // W_SECOND
                stop = false;
                bool again = false;
                if (!have_range) {
                    // is_polya_signal_shifted (polya.py:88-93)
                    const float lvl = pb::fdiv(s_ml.total, s_len.total);
                    if (fabsf(pb::fsub(lvl, P.mean_loc)) > P.trigger) again = true;
                }
                if (again) {
                    recal_mode = true;
                    walk = DO_ROUND;
                } else {
                    // calc_internal_polya_stdv of the longest event (polya.py:150-154)
                    const int64_t ilen = (int64_t)long_len;
                    const int64_t sb = (int64_t)pb::dadd((double)long_start, pb::dmul((double)ilen, P.stdv_lo));
                    const int64_t se = (int64_t)pb::dadd((double)long_start, pb::dmul((double)ilen, P.stdv_hi));
                    sd_n = 0;
                    if (se - sb > 2) {
                        const int64_t b = sb < 0 ? 0 : sb, e = se > src.n ? src.n : se;   // numpy slicing clips
                        sd_b = b;
                        sd_n = e - b;
                    }
                    if (sd_n > 0) {
                        ws = src;
                        sd_sum.begin(sd_n);
                        ws.seek(sd_b);
                        sd_q = 0;
                        walk = W_SD_MEAN;
                    } else {
                        // no usable stdv: not a poly(A) call
                        if (!have_range) { recal_mode = true; walk = DO_ROUND; }
                        else return;
                    }
                }
            }
            continue;
        }
        if (walk == W_SD_MEAN || walk == W_SD_VAR) {
            if (sd_q < sd_n) {
                // ---- one sample of the longest event
                const float x = ws.pop();
                if (walk == W_SD_MEAN) {
                    sd_sum.push(x);
                } else {
                    const float dx = pb::fsub(x, sd_mu);
                    sd_sum.push(pb::fmul(dx, dx));
                }
                sd_q++;
                continue;
            }
            if (walk == W_SD_MEAN) {
                sd_mu = pb::fdiv(sd_sum.total, (float)sd_n);
                sd_sum.begin(sd_n);
                ws.seek(sd_b);
                sd_q = 0;
                walk = W_SD_VAR;
                continue;
            }
            const float sd = sqrtf(pb::fdiv(sd_sum.total, (float)sd_n));
            if (sd < P.stdv_max) {
                R.found = 1;
                R.begin = (int64_t)first_start + insp_begin;
                R.end = last_end + insp_begin;
                R.dwell_samples = (int64_t)s_dw.total;
                R.n_spikes = nsp;
                R.extensions = ext_depth;
                return;
            }
            if (!have_range) { recal_mode = true; walk = DO_ROUND; continue; }
            return;
        }
        if (walk == DO_WINDOW) {
            // ---- one __call__ (window)
            rough_end = rough_end_cur;
            if (rough_end < 0 || rough_end - rough_begin < P.openend_unit)
                rough_end = (int64_t)rough_begin + P.openend_unit;
            insp_begin = (int64_t)rough_begin * stride - P.refinement_expansion;
            if (insp_begin < 0) insp_begin = 0;
            insp_end = (rough_end + 1) * stride + P.refinement_expansion;
            if (insp_end > full_length) insp_end = full_length;
            adapter_end = (int64_t)rough_begin * stride - insp_begin;
            src.w0 = insp_begin; src.n = insp_end - insp_begin;
            if (src.n <= 0) return;                   // csupport raises on an empty signal
            es.invalidate();                          // new window: recorded events are stale
            if (!have_range) { lo = P.cutoff_lo; hi = P.cutoff_hi; }
            recal_mode = rough_end_cur < 0;
            guard = 0;
            walk = DO_ROUND;
        }
        // ---- DO_ROUND: call_polya / try_recalibrate ping-pong
        if (++guard > 8) return;                      // cannot happen; never spin
        if (recal_mode) {
            na = 0;
            walk = W_ANCHORS;
        } else {
            best = 0; best_i = -1; best_j = -1; best_npol = 0; j = 0;
            alive = false; Sv = 0; minP = 0; minI = -1; minCnt = 0; Pfx = 0; cnt = 0;
            walk = W_FIND;
        }
        stop = false;
        es.begin(src, P);
    }
}

}  // namespace pb
