// kernels_signal.cu -- HBM-bound element kernels of the signal path:
//   k_pool      int16 DAC -> pA (fp64 affine) -> mean-pool by `stride`   (A1, A2/A4)
//   k_windows   adapter slice -> median/MAD normalised barcode window    (A6)
//   k_finalize  per-read label / default barcode fields                  (A13)
//   k_counts    (label, barcode, status) histogram                       (io.py:274-278)
#include "pb_internal.h"
#include "pb_math.cuh"

namespace pb {

// ---------------------------------------------------------------------------
// k_pool: one warp per read chunk; 32 pooled samples (= one 32*stride int16 tile, 960 bytes for
// stride 15) per iteration.
//   raw tiles  staged in shared memory by the TMA unit: lane 0 issues one bulk asynchronous copy
//              per tile (cp.async.bulk.shared.global, completion counted in bytes on an
//              mbarrier) POOL_STAGES tiles ahead of the arithmetic, so a warp always has
//              several 960-byte requests in flight instead of one load-use round trip per tile
//              (reads start 16-byte aligned and a tile is a multiple of 16 bytes for every
//              stride; the ragged last tile of a read is rounded up to 16 bytes when that stays
//              inside the buffer, else it is fetched with plain loads);
//              What bounds the kernel after that is not the stream but the conversion: per raw
//              sample one I2F.F64 and one F2F.F32.F64 (XU pipe, 16 lanes / SM / clock: 1.8 ms
//              per 4e9 samples) next to the DADD / DMUL.  Measured and rejected on B200: the
//              conversions by hand (2^52 magic add + integer rounding: 4.98 ms instead of
//              2.46), integer addition of whole-number offsets (2.60 ms).
//   lane l     converts its `stride` samples in fp64 (fast5_file.py:130-131), sums
//              them in numpy's pairwise order (signal_loader.py:224-225) and stores
//              one f32 -- 128-byte coalesced store per warp.
// grid.x covers reads (one warp per read), grid.y covers chunks of POOL_CHUNK pooled
// samples so that long reads are spread over many warps.
// ---------------------------------------------------------------------------
constexpr int POOL_WARPS = 8;
constexpr int POOL_CHUNK = 512;         // pooled samples per (warp, blockIdx.y)
constexpr int POOL_MAX_STRIDE = 32;
constexpr int POOL_STAGES = 4;

__device__ __forceinline__ uint32_t pool_smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void pool_bar_init(uint64_t *bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pool_smem_u32(bar)) : "memory");
}
// arm the barrier for `bytes` and start the bulk copy global -> shared that will complete it
__device__ __forceinline__ void pool_bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(pool_smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(pool_smem_u32(dst)), "l"(src), "r"(bytes), "r"(pool_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void pool_bar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra WAIT_%=;\n\t}"
                 ::"r"(pool_smem_u32(bar)), "r"(parity) : "memory");
}

template <int STRIDE>
__global__ void __launch_bounds__(POOL_WARPS * 32)
k_pool(const int16_t *__restrict__ raw, const int64_t *__restrict__ raw_offsets,
       const int64_t *__restrict__ raw_lengths, const double *__restrict__ range,
       const double *__restrict__ digitisation, const double *__restrict__ offset,
       int64_t n_reads, int64_t n_raw_total, int limit_pooled, float *__restrict__ pooled)
{
    constexpr int TILE = 32 * STRIDE;                 // int16 elements per warp tile
    static_assert((TILE * 2) % 16 == 0, "a tile must be a whole number of 16-byte units");
    __shared__ __align__(128) int16_t stage[POOL_WARPS][POOL_STAGES][TILE];
    __shared__ __align__(8) uint64_t bars[POOL_WARPS][POOL_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * POOL_WARPS + warp;
    if (lane == 0)
        for (int sidx = 0; sidx < POOL_STAGES; sidx++) pool_bar_init(&bars[warp][sidx]);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (r >= n_reads) return;
    const int64_t roff = raw_offsets[r];
    const int64_t rlen = raw_lengths[r];
    int64_t T = rlen / STRIDE;
    if (T > limit_pooled) T = limit_pooled;
    const int64_t t_begin = (int64_t)blockIdx.y * POOL_CHUNK;
    if (t_begin >= T) return;
    const int64_t t_end = (t_begin + POOL_CHUNK < T) ? t_begin + POOL_CHUNK : T;
    const double gain = pb::ddiv(range[r], digitisation[r]);   // range / digitisation
    const double off = offset[r];
    const int64_t pout = pooled_offset(roff, STRIDE);
    const int ntiles = (int)((t_end - t_begin + 31) / 32);

    // tile i of this warp: raw elements [e0, e0 + nel); fetched by the TMA unit when its
    // 16-byte rounded extent lies inside the buffer
    auto issue = [&](int i) {
        const int64_t t0 = t_begin + (int64_t)i * 32;
        const int64_t e0 = roff + t0 * STRIDE;
        const int64_t np = (t_end - t0 < 32) ? t_end - t0 : 32;
        const int64_t nel = (np * STRIDE + 7) & ~(int64_t)7;
        int16_t *dst = stage[warp][i % POOL_STAGES];
        if (e0 + nel <= n_raw_total) {
            if (lane == 0) pool_bulk_load(dst, raw + e0, (uint32_t)(nel * 2), &bars[warp][i % POOL_STAGES]);
        } else {
            for (int64_t j = lane; j < nel; j += 32) dst[j] = (e0 + j < n_raw_total) ? raw[e0 + j] : (int16_t)0;
            __syncwarp();
            if (lane == 0)          // complete the phase by hand so the consumer's wait is uniform
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
                             ::"r"(pool_smem_u32(&bars[warp][i % POOL_STAGES])) : "memory");
        }
    };
    for (int i = 0; i < POOL_STAGES && i < ntiles; i++) issue(i);

    for (int i = 0; i < ntiles; i++) {
        const int sidx = i % POOL_STAGES;
        pool_bar_wait(&bars[warp][sidx], (uint32_t)((i / POOL_STAGES) & 1));
        const int16_t *mine = stage[warp][sidx];
        const int64_t t = t_begin + (int64_t)i * 32 + lane;
        if (t < t_end) {
            float a[STRIDE];
#pragma unroll
            for (int j = 0; j < STRIDE; j++)
                a[j] = pb::dac_to_pa((int)mine[lane * STRIDE + j], gain, off);
            pooled[pout + t] = pb::pool_mean<STRIDE>(a);
        }
        __syncwarp();                                  // every lane is done with this stage
        if (i + POOL_STAGES < ntiles) issue(i + POOL_STAGES);
    }
}

int launch_pool(pb2_context *ctx, const pb2_batch &b, int stride, float *pooled,
                cudaStream_t st)
{
    if (b.n_reads <= 0) return PB2_OK;
    if (stride != 15)
        return fail(ctx, PB2_EUNSUPPORTED, "pooling stride %d not built (15 only)", stride);
    const int limit = ctx->scan_limit_pooled > ctx->scaler.length / stride
                          ? ctx->scan_limit_pooled : ctx->scaler.length / stride;
    // without a host-side maximum, cover the whole scan limit
    int64_t maxT = limit;
    if (b.max_raw_length > 0 && b.max_raw_length / stride < maxT) maxT = b.max_raw_length / stride;
    if (maxT < 1) maxT = 1;
    dim3 grid((unsigned)((b.n_reads + POOL_WARPS - 1) / POOL_WARPS),
              (unsigned)((maxT + POOL_CHUNK - 1) / POOL_CHUNK));
    PB_LAUNCH(ctx, K_POOL, "k_pool", st,
        k_pool<15><<<grid, POOL_WARPS * 32, 0, st>>>(b.raw, b.raw_offsets, b.raw_lengths, b.range,
                                                 b.digitisation, b.offset, b.n_reads,
                                                 b.n_raw_total, limit, pooled));
    return PB2_OK;
}

// ---------------------------------------------------------------------------
// k_windows: BarcodeDemultiplexer.push (barcoding.py:77-101).  One warp per read.
// The <= trim_length samples of the adapter tail are scaled (unfused mul, add),
// held in shared memory, and both medians are found by exact rank counting
// (rank = #smaller + #equal-with-lower-index), which reproduces np.median's
// order statistics without sorting.
// ---------------------------------------------------------------------------
constexpr int WIN_WARPS = 4;

// k-th smallest (0-based) of v[0..n) by a bitwise search on the order-preserving integer image of
// the floats: the answer is the largest key K with #{key < K} <= k.  32 counting passes of n / 32
// elements per lane instead of n rank counts of n elements each (n <= 300: ~9x fewer
// instructions); the value returned is an element of v, exactly the order statistic np.median
// / np.partition pick.
__device__ __forceinline__ uint32_t float_key(float x) {
    const uint32_t u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}
__device__ __forceinline__ uint32_t warp_select_key(const float *v, int n, int k, int lane)
{
    uint32_t res = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t cand = res | (1u << bit);
        int c = 0;
        for (int i = lane; i < n; i += 32) c += float_key(v[i]) < cand;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c <= k) res = cand;
    }
    return res;
}

__device__ __forceinline__ float warp_median(const float *v, int n, int lane,
                                              float *slot /* [2] shared, per warp; unused */)
{
    (void)slot;
    const int k_hi = n >> 1, k_lo = (n & 1) ? k_hi : k_hi - 1;
    const uint32_t key_lo = warp_select_key(v, n, k_lo, lane);
    if (n & 1) return key_float(key_lo);
    // the next order statistic: the same value if it occurs often enough, else the smallest
    // element above it
    int le = 0;
    uint32_t above = 0xFFFFFFFFu;
    for (int i = lane; i < n; i += 32) {
        const uint32_t kx = float_key(v[i]);
        le += kx <= key_lo;
        if (kx > key_lo && kx < above) above = kx;
    }
    le = __reduce_add_sync(0xffffffffu, le);
    above = __reduce_min_sync(0xffffffffu, above);
    const float lo = key_float(key_lo);
    const float hi = (le > k_hi) ? lo : key_float(above);
    return pb::fdiv(pb::fadd(lo, hi), 2.0f);        // np.mean of the two middle values
}

// Slot assignment for the compacted window list: an exclusive prefix sum of the accept test of
// BarcodeDemultiplexer.push (barcoding.py:77-84), so that slots follow read order.
// k_window_accept: the test + the prefix inside each block of ACCEPT_BLOCK reads + block totals;
// k_window_slot_bases: exclusive scan of the block totals (one block) and the grand total.
constexpr int ACCEPT_BLOCK = 1024;

__device__ __forceinline__ bool window_accepted(const int32_t *__restrict__ status,
                                                const int32_t *__restrict__ segments, int64_t r,
                                                int adapter_state, int min_len, int max_len)
{
    if (status[r] != PB2_ST_OKAY) return false;
    const int a0 = segments[(r * PB2_MAX_STATES + adapter_state) * 2 + 0];
    const int a1 = segments[(r * PB2_MAX_STATES + adapter_state) * 2 + 1];
    const int len = a1 - a0 + 1;
    return (a0 >= 0) && (len > 0) && (min_len <= len) && (len <= max_len);
}

// exclusive prefix of `v` over the ACCEPT_BLOCK threads of the block; total in *block_total
__device__ __forceinline__ int block_exclusive_sum(int v, int *wsum /* [32] shared */, int *block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    __syncthreads();                       // wsum may still be read from a previous call
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = wsum[lane];
        int wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += o;
        }
        wsum[lane] = wi - w;
        if (lane == 31) *block_total = wi;
    }
    __syncthreads();
    return wsum[warp] + incl - v;
}

__global__ void __launch_bounds__(ACCEPT_BLOCK)
k_window_accept(const int32_t *__restrict__ status, const int32_t *__restrict__ segments,
                int64_t n_reads, int adapter_state, int min_len, int max_len,
                int32_t *__restrict__ slot_of, int32_t *__restrict__ block_sums)
{
    __shared__ int wsum[32];
    __shared__ int total;
    const int64_t r = (int64_t)blockIdx.x * ACCEPT_BLOCK + threadIdx.x;
    const int ok = (r < n_reads &&
                    window_accepted(status, segments, r, adapter_state, min_len, max_len)) ? 1 : 0;
    const int ex = block_exclusive_sum(ok, wsum, &total);
    if (r < n_reads) slot_of[r] = ex;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(ACCEPT_BLOCK)
k_window_slot_bases(int32_t *__restrict__ block_sums, int n_blocks, int *__restrict__ slot_count)
{
    __shared__ int wsum[32];
    __shared__ int total;
    int carry = 0;
    for (int base = 0; base < n_blocks; base += ACCEPT_BLOCK) {
        const int i = base + threadIdx.x;
        const int v = i < n_blocks ? block_sums[i] : 0;
        const int ex = block_exclusive_sum(v, wsum, &total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
        __syncthreads();                   // everyone has read `total` before it is rewritten
    }
    if (threadIdx.x == 0) *slot_count = carry;
}

__global__ void __launch_bounds__(WIN_WARPS * 32)
k_windows(const int64_t *__restrict__ raw_offsets, const float *__restrict__ pooled,
          const float *__restrict__ scale_shift, const int32_t *__restrict__ status,
          const int32_t *__restrict__ segments, int64_t n_reads, int stride,
          int adapter_state, int min_len, int max_len, int trim_len, float pad_value,
          float *__restrict__ windows, int32_t *__restrict__ pushed,
          const int32_t *__restrict__ slot_of, const int32_t *__restrict__ slot_base,
          int32_t *__restrict__ slot_read)
{
    __shared__ float sx[WIN_WARPS][PB2_WINDOW_MAX];
    __shared__ float sd[WIN_WARPS][PB2_WINDOW_MAX];
    __shared__ float slot[WIN_WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r = (int64_t)blockIdx.x * WIN_WARPS + warp;
    if (r >= n_reads) return;
    const bool compact = slot_of != nullptr;
    float *out = windows + r * trim_len;
    int ok = 0;
    int a0 = -1, a1 = -1;
    if (status[r] == PB2_ST_OKAY) {
        a0 = segments[(r * PB2_MAX_STATES + adapter_state) * 2 + 0];
        a1 = segments[(r * PB2_MAX_STATES + adapter_state) * 2 + 1];
        const int len = a1 - a0 + 1;
        ok = (a0 >= 0) && (len > 0) && (min_len <= len) && (len <= max_len);
    }
    if (!ok) {
        if (lane == 0) pushed[r] = 0;
        if (!compact)
            for (int i = lane; i < trim_len; i += 32) out[i] = 0.0f;
        return;
    }
    if (compact) {
        // accepted windows are packed densely IN READ ORDER (slot = number of accepted reads
        // before r, from k_window_accept + k_window_slot_bases): which windows share a 128-slot
        // tile of the tensor-core classifier, and with it every float it produces, is then a
        // function of the batch alone, not of the order in which warps happened to run
        const int slot = slot_base[r / ACCEPT_BLOCK] + slot_of[r];
        if (lane == 0) slot_read[slot] = (int32_t)r;
        out = windows + (int64_t)slot * trim_len;
    }
    const int len = a1 - a0 + 1;
    const int n = len > trim_len ? trim_len : len;
    const int first = a1 - n + 1;
    const float scale = scale_shift[2 * r], shift = scale_shift[2 * r + 1];
    const float *src = pooled + pooled_offset(raw_offsets[r], stride) + first;
    float *x = sx[warp], *d = sd[warp];
    for (int i = lane; i < n; i += 32) x[i] = pb::fadd(pb::fmul(scale, src[i]), shift);
    __syncwarp();
    const float med = warp_median(x, n, lane, slot[warp]);
    for (int i = lane; i < n; i += 32) d[i] = fabsf(pb::fsub(x[i], med));
    __syncwarp();
    const float mad = warp_median(d, n, lane, slot[warp]);
    const float scaled = pb::fmul(mad, 1.4826f);
    const float div = scaled > 0.01f ? scaled : 0.01f;
    const int npad = trim_len - n;
    for (int i = lane; i < npad; i += 32) out[i] = pad_value;
    for (int i = lane; i < n; i += 32) out[npad + i] = pb::fdiv(pb::fsub(x[i], med), div);
    if (lane == 0) pushed[r] = 1;
}

int launch_windows(pb2_context *ctx, const pb2_batch &b, const float *pooled,
                   const float *scale_shift, const int32_t *status, const int32_t *segments,
                   float *windows, int32_t *pushed, int *slot_count, int32_t *slot_read,
                   cudaStream_t st)
{
    if (b.n_reads <= 0) return PB2_OK;
    const DemuxDev &d = ctx->demux;
    if (d.trim_length > PB2_WINDOW_MAX)
        return fail(ctx, PB2_EUNSUPPORTED, "signal_trim_length %d > %d", d.trim_length,
                    PB2_WINDOW_MAX);
    const int32_t *slot_of = nullptr, *slot_base = nullptr;
    if (slot_count) {
        const int n_blocks = (int)((b.n_reads + ACCEPT_BLOCK - 1) / ACCEPT_BLOCK);
        int32_t *so = (int32_t *)ws_get(ctx, ctx->ws_slotof,
                                        sizeof(int32_t) * ((size_t)b.n_reads + n_blocks));
        if (!so) return PB2_ENOMEM;
        int32_t *sums = so + b.n_reads;
        PB_LAUNCH(ctx, K_MISC, "k_window_accept", st,
            k_window_accept<<<n_blocks, ACCEPT_BLOCK, 0, st>>>(status, segments, b.n_reads,
                                                           ctx->adapter_state, d.min_length,
                                                           d.max_length, so, sums));
        PB_LAUNCH(ctx, K_MISC, "k_window_slot_bases", st,
            k_window_slot_bases<<<1, ACCEPT_BLOCK, 0, st>>>(sums, n_blocks, slot_count));
        slot_of = so;
        slot_base = sums;
    }
    const unsigned grid = (unsigned)((b.n_reads + WIN_WARPS - 1) / WIN_WARPS);
    PB_LAUNCH(ctx, K_WINDOWS, "k_windows", st,
        k_windows<<<grid, WIN_WARPS * 32, 0, st>>>(b.raw_offsets, pooled, scale_shift, status,
                                               segments, b.n_reads, ctx->scaler.stride,
                                               ctx->adapter_state, d.min_length, d.max_length,
                                               d.trim_length, d.pad_value, windows, pushed,
                                               slot_of, slot_base, slot_read));
    return PB2_OK;
}

// ---------------------------------------------------------------------------
// k_finalize: label per read (signal_analyzer.py:281-286; a read stopped before
// stage C keeps label None) and "not classified" defaults for the barcode fields.
// ---------------------------------------------------------------------------
__global__ void k_finalize(int64_t n, const int32_t *__restrict__ status,
                           const int32_t *__restrict__ pushed, int32_t *__restrict__ label,
                           int32_t *__restrict__ barcode, int32_t *__restrict__ guess,
                           int32_t *__restrict__ score)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = status[i];
    int lab;
    if (s == PB2_ST_OKAY) lab = PB2_LABEL_PASS;
    else if (s == PB2_ST_UNSPLIT_READ) lab = PB2_LABEL_ARTIFACT;
    else if (s == PB2_ST_SCALER_SIGNAL_TOO_SHORT || s == PB2_ST_SCALING_QC_FAIL ||
             s == PB2_ST_UNKNOWN_ERROR || s == PB2_ST_DISAPPEARED ||
             s == PB2_ST_IRREGULAR_FAST5) lab = PB2_LABEL_NONE;
    else lab = PB2_LABEL_FAIL;
    if (label) label[i] = lab;
    if (barcode && (pushed == nullptr || !pushed[i])) {
        barcode[i] = -1;
        if (guess) guess[i] = INT32_MIN;
        if (score) score[i] = -1;
    }
}

int launch_finalize(pb2_context *ctx, int64_t n, uint32_t flags, int32_t *status,
                    int32_t *label, int32_t *barcode, int32_t *guess, int32_t *score,
                    cudaStream_t st, const int32_t *pushed_mask)
{
    if (n <= 0) return PB2_OK;
    // the accept mask of the window stage: the caller's array, else the context's scratch
    const int32_t *pushed = (flags & PB2_FLAG_BARCODING)
                                ? (pushed_mask ? pushed_mask : (const int32_t *)ctx->ws_pushed.ptr)
                                : nullptr;
    PB_LAUNCH(ctx, K_FINALIZE, "k_finalize", st,
        k_finalize<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, status, pushed, label, barcode,
                                                           guess, score));
    return PB2_OK;
}

// ---------------------------------------------------------------------------
// k_counts: FinalSummaryTracker.feed_results (io.py:274-278) as a histogram.
// Shared-memory bins per block, one global atomic per non-empty bin.
// ---------------------------------------------------------------------------
constexpr int N_BINS = PB2_N_LABEL * PB2_N_BARCODE_SLOTS * PB2_N_STATUS;

__global__ void __launch_bounds__(256)
k_counts(const int32_t *__restrict__ status, const int32_t *__restrict__ label,
         const int32_t *__restrict__ barcode, int64_t n, unsigned long long *counts)
{
    __shared__ unsigned int bins[N_BINS];
    for (int i = threadIdx.x; i < N_BINS; i += blockDim.x) bins[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int s = status[i];
        const int l = label[i];
        const int b = barcode ? barcode[i] : -1;
        if (s < 0 || s >= PB2_N_STATUS || l < 0 || l >= PB2_N_LABEL) continue;
        const int slot = (b >= 0 && b < PB2_N_BARCODE_SLOTS - 1) ? b + 1 : 0;
        atomicAdd(&bins[(l * PB2_N_BARCODE_SLOTS + slot) * PB2_N_STATUS + s], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N_BINS; i += blockDim.x)
        if (bins[i]) atomicAdd(&counts[i], (unsigned long long)bins[i]);
}

int launch_counts(pb2_context *ctx, const int32_t *status, const int32_t *label,
                  const int32_t *barcode, int64_t n, int64_t *counts, cudaStream_t st)
{
    PB_CUDA(ctx, cudaMemsetAsync(counts, 0, sizeof(int64_t) * N_BINS, st));
    if (n <= 0) return PB2_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    PB_LAUNCH(ctx, K_COUNTS, "k_counts", st,
        k_counts<<<(unsigned)blocks, 256, 0, st>>>(status, label, barcode, n,
                                               reinterpret_cast<unsigned long long *>(counts)));
    return PB2_OK;
}

}  // namespace pb
