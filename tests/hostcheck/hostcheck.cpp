// hostcheck.cpp -- TEST-ONLY host build of the __host__ __device__ cores in
// poreplex_b200/csrc (compiled with g++, no CUDA).  Lets the CPU test-suite run the very
// code the kernels execute against the oracle.  Never used by the product path.
#include <cstdint>
#include <cmath>
#include "../../poreplex_b200/csrc/polya_core.cuh"

// csupport.detect_events on a plain float32 signal (the stream k_detect_events runs);
// ring: 64 or 512 prefix-sum entries
template <int RING>
static int64_t detect_plain(const float *x, int64_t n, const pb::PolyaParams *P, uint64_t *start,
                            float *length, float *mean, float *stdv, int64_t cap)
{
    pb::PlainSource src;
    src.x = x; src.n = n; src.next = 0;
    pb::PlainEventStream<RING> es;
    pb::EventRings<RING> rings;
    es.use(rings);
    es.begin(src, *P);
    pb::Event ev;
    int64_t k = 0;
    while (es.next(ev)) {
        if (k < cap) { start[k] = ev.start; length[k] = ev.length; mean[k] = ev.mean; stdv[k] = ev.stdv; }
        k++;
    }
    return k;
}

extern "C" {

float hc_median7(const float *v) { return pb::median7(v[0], v[1], v[2], v[3], v[4], v[5], v[6]); }

float hc_median5(const float *v) { return pb::median5(v[0], v[1], v[2], v[3], v[4]); }

float hc_pairwise_sum(const float *a, int64_t n)
{
    pb::PairwiseSum s;
    s.begin(n);
    for (int64_t i = 0; i < n; i++) s.push(a[i]);
    return s.total;
}

// events of one window into caller arrays; returns the count
int64_t hc_detect_events(const int16_t *raw, int64_t n, double gain, double offset, float scale,
                         float shift, const pb::PolyaParams *P, uint64_t *start, float *length,
                         float *mean, float *stdv, int64_t cap)
{
    pb::WindowSource src;
    src.raw = raw; src.gain = gain; src.offset = offset; src.scale = scale; src.shift = shift;
    src.w0 = 0; src.n = n;
    pb::EventStream es;
    pb::EventRings<64> rings;
    es.use(rings);
    es.begin(src, *P);
    pb::Event ev;
    int64_t k = 0;
    while (es.next(ev)) {
        if (k < cap) { start[k] = ev.start; length[k] = ev.length; mean[k] = ev.mean; stdv[k] = ev.stdv; }
        k++;
    }
    return k;
}

int64_t hc_detect_events_plain(const float *x, int64_t n, const pb::PolyaParams *P, int ring,
                               uint64_t *start, float *length, float *mean, float *stdv, int64_t cap)
{
    return ring == 64 ? detect_plain<64>(x, n, P, start, length, mean, stdv, cap)
                      : detect_plain<512>(x, n, P, start, length, mean, stdv, cap);
}

void hc_polya(const pb::PolyaParams *P, const int16_t *raw, int64_t full_length, double gain,
              double offset, float scale, float shift, int32_t rough_begin, int32_t rough_end,
              pb::PolyaResult *R)
{
    pb::polya_analyze(*P, raw, full_length, gain, offset, scale, shift, rough_begin, rough_end, *R);
}

// same, with the event replay cache enabled (capacity `cap` events)
void hc_polya_cached(const pb::PolyaParams *P, const int16_t *raw, int64_t full_length, double gain,
                     double offset, float scale, float shift, int32_t rough_begin,
                     int32_t rough_end, pb::PolyaResult *R, int cap)
{
    pb::EventCacheSlot *buf = new pb::EventCacheSlot[cap > 0 ? cap : 1];
    pb::polya_analyze(*P, raw, full_length, gain, offset, scale, shift, rough_begin, rough_end, *R,
                      buf, 1, cap);
    delete[] buf;
}

// the literal (one loop per walk) formulation, which the kernel's single-loop one must equal
void hc_polya_nested(const pb::PolyaParams *P, const int16_t *raw, int64_t full_length, double gain,
                     double offset, float scale, float shift, int32_t rough_begin, int32_t rough_end,
                     pb::PolyaResult *R, int cap)
{
    pb::EventCacheSlot *buf = cap > 0 ? new pb::EventCacheSlot[cap] : nullptr;
    pb::polya_analyze_nested(*P, raw, full_length, gain, offset, scale, shift, rough_begin, rough_end,
                             *R, buf, 1, cap);
    delete[] buf;
}

int hc_sizeof_params(void) { return (int)sizeof(pb::PolyaParams); }
int hc_sizeof_result(void) { return (int)sizeof(pb::PolyaResult); }
}
