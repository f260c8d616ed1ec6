// kernels_lstm.cu -- exact-f32 LSTM inference for the scaler (A3) and the barcode
// demultiplexer (A7).
//
// Reference: keras Model.predict at signal_loader.py:96-97 and barcoding.py:106-107
// (TensorFlow is not vendored; cell equations and activation kernels restated in
// oracle/pb_oracle.c).  Bit-exact contract: every pre-activation is an fmaf chain
// from 0 with k ascending per K.dot(), combined in Keras' order for the layer's
// `implementation`; activations are pb::sigmoid_eigen / pb::tanh_eigen.
//
// Work decomposition (all kernels): a CTA owns a tile of TB = 32 reads and steps them
// through time together.  A thread owns one UNIT PAIR (2 hidden units x 4 gates) for
// RG = 4 reads: 32 independent fma chains, fed per k by one 16-byte load of the 4
// reads' h[k] (smem, k-major) and two 16-byte loads of its 8 weights (smem, re-laid
// out as [k][unit_pair][gate][2]).  The chains use packed f32x2 FMAs (FFMA2 on
// sm_100), each half being an IEEE fma identical to the scalar one.  The cell state
// c lives in registers of the owning thread; h goes through double-buffered shared
// memory with ONE block barrier per time step.
#include "pb_internal.h"
#include "pb_math.cuh"
#include "demux_head.cuh"

namespace pb {

constexpr int TB = 32;            // reads per CTA tile
constexpr int RG = 4;             // reads per thread
constexpr int NRG = TB / RG;      // read groups per tile (8)

// ---- read groups --------------------------------------------------------------
// A CTA holds GROUPS independent tiles ("read groups") that share the weight matrices in
// shared memory but step through time on their own: each group synchronises with a
// named barrier of its own, so while one group is in its activation phase (scalar FMA /
// MUFU / ALU pipes) another is in its dot-product phase (packed FFMA2 pipe).
__device__ __forceinline__ void group_sync(int group, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(nthreads) : "memory");
}

// ---- packed helpers ---------------------------------------------------------
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    return __ffma2_rn(a, b, c);
}

struct Acc {                       // z[r][gate] for units (2up, 2up+1)
    float2 v[RG][4];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int r = 0; r < RG; r++)
#pragma unroll
            for (int g = 0; g < 4; g++) v[r][g] = make_float2(0.f, 0.f);
    }
};

// acc[r][g] = fma chain over k = 0..K-1 of h[k][r] * W[k][g][unit]   (from 0)
//   Wt : smem, [K][NUP][4][2] floats        hs : smem, [K][TB] floats
template <int K, int NUP>
__device__ __forceinline__ void dot_tile(const float *__restrict__ Wt,
                                         const float *__restrict__ hs, int up, int rg,
                                         Acc &acc)
{
    acc.zero();
    const float4 *w = reinterpret_cast<const float4 *>(Wt) + up * 2;
    const float4 *h = reinterpret_cast<const float4 *>(hs) + rg;
#pragma unroll 4
    for (int k = 0; k < K; k++) {
        const float4 hv = h[k * (TB / 4)];
        const float4 w0 = w[k * NUP * 2];
        const float4 w1 = w[k * NUP * 2 + 1];
        const float2 wi = make_float2(w0.x, w0.y), wf = make_float2(w0.z, w0.w);
        const float2 wc = make_float2(w1.x, w1.y), wo = make_float2(w1.z, w1.w);
        const float hr[RG] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int r = 0; r < RG; r++) {
            const float2 hh = make_float2(hr[r], hr[r]);
            acc.v[r][0] = ffma2(hh, wi, acc.v[r][0]);
            acc.v[r][1] = ffma2(hh, wf, acc.v[r][1]);
            acc.v[r][2] = ffma2(hh, wc, acc.v[r][2]);
            acc.v[r][3] = ffma2(hh, wo, acc.v[r][3]);
        }
    }
}

// same chain continued: acc keeps its incoming value, k runs over [K0, K1)
template <int K0, int K1, int NUP>
__device__ __forceinline__ void dot_tile_range(const float *__restrict__ Wt,
                                               const float *__restrict__ hs, int up, int rg,
                                               Acc &acc)
{
    const float4 *w = reinterpret_cast<const float4 *>(Wt) + up * 2;
    const float4 *h = reinterpret_cast<const float4 *>(hs) + rg;
#pragma unroll 4
    for (int k = K0; k < K1; k++) {
        const float4 hv = h[k * (TB / 4)];
        const float4 w0 = w[k * NUP * 2];
        const float4 w1 = w[k * NUP * 2 + 1];
        const float2 wi = make_float2(w0.x, w0.y), wf = make_float2(w0.z, w0.w);
        const float2 wc = make_float2(w1.x, w1.y), wo = make_float2(w1.z, w1.w);
        const float hr[RG] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int r = 0; r < RG; r++) {
            const float2 hh = make_float2(hr[r], hr[r]);
            acc.v[r][0] = ffma2(hh, wi, acc.v[r][0]);
            acc.v[r][1] = ffma2(hh, wf, acc.v[r][1]);
            acc.v[r][2] = ffma2(hh, wc, acc.v[r][2]);
            acc.v[r][3] = ffma2(hh, wo, acc.v[r][3]);
        }
    }
}

// re-layout a [K][4H] row-major Keras matrix into smem [K][NUP][4][2]
template <int K, int H>
__device__ __forceinline__ void load_weights(const float *__restrict__ g, float *Wt)
{
    constexpr int NUP = H / 2;
    for (int idx = threadIdx.x; idx < K * 4 * H; idx += blockDim.x) {
        const int k = idx / (4 * H), col = idx % (4 * H);
        const int gate = col / H, u = col % H;
        Wt[((k * NUP + (u >> 1)) * 4 + gate) * 2 + (u & 1)] = g[idx];
    }
}

// gate bias / input-kernel values of this thread's unit pair: [gate] -> (u0, u1)
template <int H>
__device__ __forceinline__ void load_pair(const float *__restrict__ v, int up, float2 (&out)[4])
{
#pragma unroll
    for (int g = 0; g < 4; g++) out[g] = make_float2(v[g * H + 2 * up], v[g * H + 2 * up + 1]);
}

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    return make_float2(pb::fadd(a.x, b.x), pb::fadd(a.y, b.y));
}
__device__ __forceinline__ float2 fmul2s(float s, float2 b) {
    return make_float2(pb::fmul(s, b.x), pb::fmul(s, b.y));
}

// cell update for both units of the pair, one read; returns h
template <bool EXACT>
__device__ __forceinline__ float2 cell_pair(const float2 (&z)[4], float2 &c, bool &risk)
{
    float2 h;
    pb::lstm_cell<EXACT>(z[0].x, z[1].x, z[2].x, z[3].x, c.x, h.x, risk);
    pb::lstm_cell<EXACT>(z[0].y, z[1].y, z[2].y, z[3].y, c.y, h.y, risk);
    return h;
}

// store this thread's new h for its unit pair and 4 reads: hs[k = unit][read]
__device__ __forceinline__ void store_h(float *hs, int up, int rg, const float2 (&h)[RG])
{
    *reinterpret_cast<float4 *>(hs + (2 * up) * TB + rg * RG) =
        make_float4(h[0].x, h[1].x, h[2].x, h[3].x);
    *reinterpret_cast<float4 *>(hs + (2 * up + 1) * TB + rg * RG) =
        make_float4(h[0].y, h[1].y, h[2].y, h[3].y);
}

// ============================================================================
// Scaler: LSTM(H, seq, impl 1) -> LSTM(H, last, impl 1) -> Dense(2)
// ============================================================================
struct ScalerArgs {
    const float *x;                // float buffer holding the (unscaled) pooled signals
    const int64_t *xoff;           // [n] element offset of each read's first real sample
    const int32_t *nreal;          // [n] real head samples (0 = skip read)
    int64_t n;
    int thead;                     // time steps including the left zero padding
    const float *W1, *U1, *b1, *W2, *U2, *b2, *Wd, *bd;
    const float *zero_prefix;      // [thead+1][4][H] or nullptr
    float *prefix_dump;            // when set: record the state of read 0 after each step
    float *z_out;                  // [n][2] raw network outputs (optional)
    // output transform + QC (signal_loader.py:98-109)
    double scale_std, scale_mean, shift_std, shift_mean;
    double qc_scale_lo, qc_scale_hi, qc_shift_lo, qc_shift_hi;
    int32_t *status;               // may be nullptr (heads API)
    float *scale_shift;            // may be nullptr
};

// EXACT = false: branch-free divisions (pb::div_posq); EXACT = true: IEEE __fdiv_rn
// everywhere (verification mode, pb2_set_exact_division, and the zero-prefix table).
constexpr int SCALER_GROUPS = 2;

template <int H, int GROUPS, bool EXACT>
__global__ void __launch_bounds__(GROUPS * (H / 2) * NRG, 1)
k_scaler_lstm(const ScalerArgs A)
{
    constexpr int NUP = H / 2;
    constexpr int GT = NUP * NRG;             // threads per read group
    bool risk = false;
    extern __shared__ __align__(16) float smem[];
    float *U1t = smem;                        // [H][NUP][4][2]
    float *W2t = U1t + H * 4 * H;
    float *U2t = W2t + H * 4 * H;
    __shared__ int s_nreal_all[GROUPS][TB];
    __shared__ int s_maxreal_all[GROUPS];

    const int group = threadIdx.x / GT;
    const int tid = threadIdx.x % GT;
    const int rg = tid % NRG, up = tid / NRG;
    float *h1s = U2t + H * 4 * H + group * (4 * H * TB);   // [2][H][TB]
    float *h2s = h1s + 2 * H * TB;                         // [2][H][TB]
    int *s_nreal = s_nreal_all[group];
    const int64_t tile = (int64_t)blockIdx.x * GROUPS + group;
    const int64_t tile0 = tile * TB;

    load_weights<H, H>(A.U1, U1t);
    load_weights<H, H>(A.W2, W2t);
    load_weights<H, H>(A.U2, U2t);
    if (tid == 0) s_maxreal_all[group] = 0;
    __syncthreads();
    if (tile0 >= A.n) return;                 // whole group past the end (never syncs again)
    if (tid < TB) {
        const int64_t r = tile0 + tid;
        const int nr = (r < A.n) ? A.nreal[r] : 0;
        s_nreal[tid] = nr;
        atomicMax(&s_maxreal_all[group], nr);
    }
    group_sync(group, GT);
    const int t_start = A.zero_prefix ? (A.thead - s_maxreal_all[group]) : 0;

    // per-thread constants
    float2 w1[4], b1[4], b2[4];
    load_pair<H>(A.W1, up, w1);
    load_pair<H>(A.b1, up, b1);
    load_pair<H>(A.b2, up, b2);
    // per-read input addressing
    const float *xp[RG];
    int pad[RG];
#pragma unroll
    for (int r = 0; r < RG; r++) {
        const int64_t rd = tile0 + rg * RG + r;
        const int nr = s_nreal[rg * RG + r];
        pad[r] = A.thead - nr;
        xp[r] = A.x + ((rd < A.n && nr > 0) ? A.xoff[rd] : 0) - pad[r];
    }

    // initial state: zeros, or the tabulated state after t_start zero-input steps
    float2 c1[RG], c2[RG];
    {
        float2 c1i = make_float2(0.f, 0.f), c2i = c1i, h1i = c1i, h2i = c1i;
        if (A.zero_prefix && t_start > 0) {
            const float *tb = A.zero_prefix + (size_t)t_start * 4 * H;
            h1i = make_float2(tb[0 * H + 2 * up], tb[0 * H + 2 * up + 1]);
            c1i = make_float2(tb[1 * H + 2 * up], tb[1 * H + 2 * up + 1]);
            h2i = make_float2(tb[2 * H + 2 * up], tb[2 * H + 2 * up + 1]);
            c2i = make_float2(tb[3 * H + 2 * up], tb[3 * H + 2 * up + 1]);
        }
        float2 h1v[RG], h2v[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) { c1[r] = c1i; c2[r] = c2i; h1v[r] = h1i; h2v[r] = h2i; }
        store_h(h1s, up, rg, h1v);
        store_h(h2s, up, rg, h2v);
    }
    group_sync(group, GT);

    int cur1 = 0, cur2 = 0;
    Acc acc;
    for (int t = t_start; t < A.thead; t++) {
        // ---- layer 1, step t: z = ((x*W1 + b1) + h1.U1)
        float xv[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) xv[r] = (t >= pad[r]) ? __ldg(xp[r] + t) : 0.0f;
        dot_tile<H, NUP>(U1t, h1s + cur1 * H * TB, up, rg, acc);
        float2 hn[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) {
            float2 z[4];
#pragma unroll
            for (int g = 0; g < 4; g++)
                z[g] = fadd2(fadd2(fmul2s(xv[r], w1[g]), b1[g]), acc.v[r][g]);
            hn[r] = cell_pair<EXACT>(z, c1[r], risk);
        }
        store_h(h1s + (cur1 ^ 1) * H * TB, up, rg, hn);
        cur1 ^= 1;
        group_sync(group, GT);
        // ---- layer 2, step t: z = ((h1.W2 + b2) + h2.U2)
        dot_tile<H, NUP>(W2t, h1s + cur1 * H * TB, up, rg, acc);
        float2 zx[RG][4];
#pragma unroll
        for (int r = 0; r < RG; r++)
#pragma unroll
            for (int g = 0; g < 4; g++) zx[r][g] = fadd2(acc.v[r][g], b2[g]);
        dot_tile<H, NUP>(U2t, h2s + cur2 * H * TB, up, rg, acc);
#pragma unroll
        for (int r = 0; r < RG; r++) {
            float2 z[4];
#pragma unroll
            for (int g = 0; g < 4; g++) z[g] = fadd2(zx[r][g], acc.v[r][g]);
            hn[r] = cell_pair<EXACT>(z, c2[r], risk);
        }
        store_h(h2s + (cur2 ^ 1) * H * TB, up, rg, hn);
        cur2 ^= 1;
        if (A.prefix_dump && tile == 0 && rg == 0) {
            // state of read 0 after t + 1 steps; h1 was stored above into buffer cur1
            float *tb = A.prefix_dump + (size_t)(t + 1) * 4 * H;
            const float *h1n = h1s + cur1 * H * TB;
            tb[0 * H + 2 * up] = h1n[(2 * up) * TB];
            tb[0 * H + 2 * up + 1] = h1n[(2 * up + 1) * TB];
            tb[1 * H + 2 * up] = c1[0].x;  tb[1 * H + 2 * up + 1] = c1[0].y;
            tb[2 * H + 2 * up] = hn[0].x;  tb[2 * H + 2 * up + 1] = hn[0].y;
            tb[3 * H + 2 * up] = c2[0].x;  tb[3 * H + 2 * up + 1] = c2[0].y;
        }
        // no barrier needed here: the next writers of h1 target the buffer last read
        // before the barrier above, and h2's new buffer is read only after the next one
    }
    group_sync(group, GT);

    // ---- Dense(2) + output transform + QC, one thread per read
    if (tid < TB) {
        const int64_t r = tile0 + tid;
        if (r < A.n && s_nreal[tid] > 0) {
            const float *h2f = h2s + cur2 * H * TB;
            float z0 = 0.f, z1 = 0.f;
            for (int k = 0; k < H; k++) {
                const float hk = h2f[k * TB + tid];
                z0 = pb::ffma(hk, A.Wd[2 * k], z0);
                z1 = pb::ffma(hk, A.Wd[2 * k + 1], z1);
            }
            z0 = pb::fadd(z0, A.bd[0]);
            z1 = pb::fadd(z1, A.bd[1]);
            if (A.z_out) { A.z_out[2 * r] = z0; A.z_out[2 * r + 1] = z1; }
            if (A.scale_shift) {
                // poly1d([std, mean])(z) promotes to fp64 under numpy 2
                const double sc = pb::dadd(pb::dmul(A.scale_std, (double)z0), A.scale_mean);
                const double sh = pb::dadd(pb::dmul(A.shift_std, (double)z1), A.shift_mean);
                A.scale_shift[2 * r] = (float)sc;
                A.scale_shift[2 * r + 1] = (float)sh;
                const bool ok = sc >= A.qc_scale_lo && sc <= A.qc_scale_hi &&
                                sh >= A.qc_shift_lo && sh <= A.qc_shift_hi;
                if (A.status) A.status[r] = ok ? PB2_ST_OKAY : PB2_ST_SCALING_QC_FAIL;
            }
        }
    }
    (void)risk;
}

// load_padded_signal_head bookkeeping (signal_loader.py:212-222): how many pooled
// samples feed the scaler, or scaler_signal_too_short.
__global__ void k_scaler_prepare(const int64_t *__restrict__ raw_offsets,
                                 const int64_t *__restrict__ raw_lengths, int64_t n, int stride,
                                 int length, int min_length, int32_t *__restrict__ status,
                                 int64_t *__restrict__ xoff, int32_t *__restrict__ nreal,
                                 float *__restrict__ scale_shift)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t sl = raw_lengths[i] < length ? raw_lengths[i] : length;
    sl -= sl % stride;
    xoff[i] = pooled_offset(raw_offsets[i], stride);
    scale_shift[2 * i] = 0.f;
    scale_shift[2 * i + 1] = 0.f;
    if (sl < min_length || sl <= 0) {
        status[i] = PB2_ST_SCALER_SIGNAL_TOO_SHORT;
        nreal[i] = 0;
    } else {
        status[i] = PB2_ST_OKAY;
        nreal[i] = (int32_t)(sl / stride);
    }
}

int launch_scaler_prepare(pb2_context *ctx, const pb2_batch &b, int32_t *status, float *scale_shift,
                          int64_t *xoff, int32_t *nreal, cudaStream_t st)
{
    const ScalerDev &S = ctx->scaler;
    PB_LAUNCH(ctx, K_SCALER_PREPARE, "k_scaler_prepare", st,
        k_scaler_prepare<<<(unsigned)((b.n_reads + 255) / 256), 256, 0, st>>>(
        b.raw_offsets, b.raw_lengths, b.n_reads, S.stride, S.length, S.min_length, status, xoff,
        nreal, scale_shift));
    return PB2_OK;
}

template <int H, int GROUPS>
static size_t scaler_smem() { return sizeof(float) * (3 * H * 4 * H + GROUPS * 4 * H * TB); }

static int run_scaler(pb2_context *ctx, ScalerArgs &A, cudaStream_t st, bool exact_only = false)
{
    const ScalerDev &S = ctx->scaler;
    if (S.l1.units != 48 || S.l2.units != 48 || S.l1.in_dim != 1 || S.l2.in_dim != 48 ||
        S.l1.impl != 1 || S.l2.impl != 1)
        return fail(ctx, PB2_EUNSUPPORTED, "scaler network shape not built "
                    "(LSTM(48,impl1) x2 expected)");
    A.W1 = S.l1.kernel; A.U1 = S.l1.recurrent; A.b1 = S.l1.bias;
    A.W2 = S.l2.kernel; A.U2 = S.l2.recurrent; A.b2 = S.l2.bias;
    A.Wd = S.dense_kernel; A.bd = S.dense_bias;
    A.scale_std = S.scale_std; A.scale_mean = S.scale_mean;
    A.shift_std = S.shift_std; A.shift_mean = S.shift_mean;
    A.qc_scale_lo = S.qc_scale_lo; A.qc_scale_hi = S.qc_scale_hi;
    A.qc_shift_lo = S.qc_shift_lo; A.qc_shift_hi = S.qc_shift_hi;
    constexpr int G = SCALER_GROUPS;
    const size_t smem = scaler_smem<48, G>();
    bool &attr_done = ctx->attr_scaler;      // per context: the attribute is per device
    if (!attr_done) {
        PB_CUDA(ctx, cudaFuncSetAttribute(k_scaler_lstm<48, G, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PB_CUDA(ctx, cudaFuncSetAttribute(k_scaler_lstm<48, G, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    const int64_t tiles = (A.n + TB - 1) / TB;
    const unsigned grid = (unsigned)((tiles + G - 1) / G);
    if (exact_only || ctx->exact_division) {
        PB_LAUNCH(ctx, K_SCALER_LSTM, "k_scaler_lstm<exact>", st,
            k_scaler_lstm<48, G, true><<<grid, G * 24 * NRG, smem, st>>>(A));
    } else {
        PB_LAUNCH(ctx, K_SCALER_LSTM, "k_scaler_lstm", st,
            k_scaler_lstm<48, G, false><<<grid, G * 24 * NRG, smem, st>>>(A));
    }
    return PB2_OK;
}

int build_zero_prefix(pb2_context *ctx)
{
    ScalerDev &S = ctx->scaler;
    const int thead = S.length / S.stride;
    const int H = S.l1.units;
    const size_t bytes = sizeof(float) * (size_t)(thead + 1) * 4 * H;
    if (S.zero_prefix) cudaFree(S.zero_prefix);
    S.zero_prefix = nullptr;
    float *table = nullptr;
    PB_CUDA(ctx, cudaMalloc(&table, bytes));
    PB_CUDA(ctx, cudaMemset(table, 0, bytes));
    float *zeros = nullptr;
    int64_t *xoff = nullptr;
    int32_t *nreal = nullptr;
    PB_CUDA(ctx, cudaMalloc(&zeros, sizeof(float) * thead));
    PB_CUDA(ctx, cudaMemset(zeros, 0, sizeof(float) * thead));
    PB_CUDA(ctx, cudaMalloc(&xoff, sizeof(int64_t)));
    PB_CUDA(ctx, cudaMemset(xoff, 0, sizeof(int64_t)));
    PB_CUDA(ctx, cudaMalloc(&nreal, sizeof(int32_t)));
    PB_CUDA(ctx, cudaMemcpy(nreal, &thead, sizeof(int32_t), cudaMemcpyHostToDevice));
    ScalerArgs A = {};
    A.x = zeros; A.xoff = xoff; A.nreal = nreal; A.n = 1; A.thead = thead;
    A.zero_prefix = nullptr; A.prefix_dump = table;
    int rc = run_scaler(ctx, A, 0, /*exact_only=*/true);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(zeros); cudaFree(xoff); cudaFree(nreal);
    if (rc != PB2_OK) { cudaFree(table); return rc; }
    if (e != cudaSuccess) { cudaFree(table); return check_cuda(ctx, e, "zero-prefix build"); }
    S.zero_prefix = table;
    S.zero_prefix_steps = thead;
    return PB2_OK;
}

int launch_scaler(pb2_context *ctx, const pb2_batch &b, const float *pooled, int32_t *status,
                  float *scale_shift, float *z_out, cudaStream_t st)
{
    if (b.n_reads <= 0) return PB2_OK;
    const ScalerDev &S = ctx->scaler;
    int64_t *xoff = (int64_t *)ws_get(ctx, ctx->ws_misc, (size_t)b.n_reads * 12);
    if (!xoff) return PB2_ENOMEM;
    int32_t *nreal = (int32_t *)(xoff + b.n_reads);
    PB_LAUNCH(ctx, K_SCALER_PREPARE, "k_scaler_prepare", st,
        k_scaler_prepare<<<(unsigned)((b.n_reads + 255) / 256), 256, 0, st>>>(
        b.raw_offsets, b.raw_lengths, b.n_reads, S.stride, S.length, S.min_length, status, xoff,
        nreal, scale_shift));
    ScalerArgs A = {};
    A.x = pooled; A.xoff = xoff; A.nreal = nreal; A.n = b.n_reads;
    A.thead = S.length / S.stride;
    A.zero_prefix = S.zero_prefix; A.prefix_dump = nullptr; A.z_out = z_out;
    A.status = status; A.scale_shift = scale_shift;
    return run_scaler(ctx, A, st);
}

__global__ void k_iota_heads(int64_t n, int thead, int64_t *xoff, int32_t *nreal)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { xoff[i] = i * thead; nreal[i] = thead; }
}

int launch_scaler_heads(pb2_context *ctx, const float *heads, int64_t n, float *z_out,
                        cudaStream_t st)
{
    if (n <= 0) return PB2_OK;
    const ScalerDev &S = ctx->scaler;
    int64_t *xoff = (int64_t *)ws_get(ctx, ctx->ws_misc, (size_t)n * 12);
    if (!xoff) return PB2_ENOMEM;
    int32_t *nreal = (int32_t *)(xoff + n);
    const int thead = S.length / S.stride;
    PB_LAUNCH(ctx, K_MISC, "k_iota_heads", st,
        k_iota_heads<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, thead, xoff, nreal));
    ScalerArgs A = {};
    A.x = heads; A.xoff = xoff; A.nreal = nreal; A.n = n; A.thead = thead;
    A.zero_prefix = nullptr;       // explicit heads: run every step as keras does
    A.z_out = z_out;
    return run_scaler(ctx, A, st);
}

// ============================================================================
// Demultiplexer, layer 1: Bidirectional(LSTMCell(H1, impl 2)); one direction per
// blockIdx.y.  Output h(t) of both directions goes to a scratch laid out for layer 2:
//   G[tile][t][dir*H1 + unit][TB]
// ============================================================================
struct DemuxArgs {
    const float *windows;          // [n][T]
    int64_t n;
    int T;
    const float *Wf, *Uf, *bf, *Wb, *Ub, *bb;     // layer 1 (fwd, bwd)
    const float *W2, *U2, *b2;                    // layer 2
    const float *Wd, *bd;                         // dense
    float *G;                                     // layer-1 outputs
    const int32_t *pushed;                        // may be nullptr
    const int *slot_count;                        // compacted input: valid rows (or nullptr)
    const int32_t *slot_read;                     // compacted input: row -> read (or nullptr)
    int64_t row0;                                 // first row of this pass
    int n_classes, n_decoy;
    int n_calibration;
    double score_threshold;
    float *class_probs; int32_t *barcode, *guess, *score;
    // left-pad skipping (all nullptr / 0 = off).  While every read of a tile is still in its
    // -1000 left padding, the forward layer-1 state is a read-independent function of t:
    //   pad_state [T+1][2][H1]   h, c after n pad steps (built once by this same kernel)
    //   pad_prefix[T][H2/2][4][2] layer-2 chain x.W2 over the forward half (k < H1) at pad step t
    //   tile_tstart[tiles]       per tile: min over its reads of the leading pad count
    const float *pad_state;
    const float *pad_prefix;
    int *tile_tstart;
    float *pad_dump;               // table construction: record (h, c) of row 0 after each step
    float pad_value;
};

__constant__ double c_calibration[PB2_MAX_CALIB];

template <int H1, bool EXACT>
__global__ void __launch_bounds__((H1 / 2) * NRG, 3)
k_demux_l1(const DemuxArgs A)
{
    constexpr int NUP = H1 / 2;
    bool risk = false;
    extern __shared__ __align__(16) float smem[];
    float *Ut = smem;                         // [H1][NUP][4][2]
    float *hs = Ut + H1 * 4 * H1;             // [2][H1][TB]
    const int tid = threadIdx.x;
    const int rg = tid % NRG, up = tid / NRG;
    const int dir = blockIdx.y;
    const int64_t tile = blockIdx.x;
    const int64_t tile0 = tile * TB;
    int64_t n_eff = A.n;                      // rows of this pass that hold a window
    if (A.slot_count) {
        n_eff = (int64_t)*A.slot_count - A.row0;
        if (n_eff > A.n) n_eff = A.n;
    }
    if (tile0 >= n_eff) return;
    const float *W = dir ? A.Wb : A.Wf, *U = dir ? A.Ub : A.Uf, *bias = dir ? A.bb : A.bf;

    load_weights<H1, H1>(U, Ut);
    float2 w[4], b[4];
    load_pair<H1>(W, up, w);
    load_pair<H1>(bias, up, b);
    const float *xp[RG];
#pragma unroll
    for (int r = 0; r < RG; r++) {
        int64_t rd = tile0 + rg * RG + r;
        if (rd >= n_eff) rd = n_eff - 1;
        xp[r] = A.windows + rd * A.T;
    }
    // forward direction: steps that are left padding for EVERY read of the tile are not
    // computed -- their (h, c) come from the pad-state table and G is filled from it
    __shared__ int s_tstart;
    int t_start = 0;
    if (dir == 0 && A.pad_state) {
        if (tid == 0) s_tstart = A.T;
        __syncthreads();
        if (tid < TB) {
            int64_t rd = tile0 + tid;
            if (rd < n_eff) {
                const float *row = A.windows + rd * A.T;
                int np = 0;
                while (np < A.T && row[np] == A.pad_value) np++;
                atomicMin(&s_tstart, np);
            }
        }
        __syncthreads();
        t_start = s_tstart;
        if (t_start >= A.T) t_start = A.T - 1;      // keep at least the last step live
        if (tid == 0 && A.tile_tstart) A.tile_tstart[tile] = t_start;
    }
    float *Gt = A.G + (size_t)tile * A.T * (2 * H1) * TB;
    float2 c[RG], hz[RG];
    {
        float2 h0 = make_float2(0.f, 0.f), c0 = h0;
        if (t_start > 0) {
            const float *tb = A.pad_state + (size_t)t_start * 2 * H1;
            h0 = make_float2(tb[2 * up], tb[2 * up + 1]);
            c0 = make_float2(tb[H1 + 2 * up], tb[H1 + 2 * up + 1]);
            for (int t = 0; t < t_start; t++) {     // h after pad step t = table[t + 1]
                const float *tt = A.pad_state + (size_t)(t + 1) * 2 * H1;
                const float2 hv = make_float2(tt[2 * up], tt[2 * up + 1]);
                float2 hq[RG];
#pragma unroll
                for (int r = 0; r < RG; r++) hq[r] = hv;
                store_h(Gt + ((size_t)t * 2 * H1) * TB, up, rg, hq);
            }
        }
#pragma unroll
        for (int r = 0; r < RG; r++) { c[r] = c0; hz[r] = h0; }
    }
    store_h(hs, up, rg, hz);
    __syncthreads();

    int cur = 0;
    Acc acc;
    for (int s = t_start; s < A.T; s++) {
        const int t = dir ? (A.T - 1 - s) : s;
        float xv[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) xv[r] = __ldg(xp[r] + t);
        dot_tile<H1, NUP>(Ut, hs + cur * H1 * TB, up, rg, acc);
        float2 hn[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) {
            float2 z[4];
#pragma unroll
            for (int g = 0; g < 4; g++)       // implementation 2: ((x.W + h.U) + b)
                z[g] = fadd2(fadd2(fmul2s(xv[r], w[g]), acc.v[r][g]), b[g]);
            hn[r] = cell_pair<EXACT>(z, c[r], risk);
        }
        store_h(hs + (cur ^ 1) * H1 * TB, up, rg, hn);
        store_h(Gt + ((size_t)t * 2 * H1 + dir * H1) * TB, up, rg, hn);
        if (A.pad_dump && dir == 0 && tile == 0 && rg == 0) {
            float *tt = A.pad_dump + (size_t)(s + 1) * 2 * H1;
            tt[2 * up] = hn[0].x;  tt[2 * up + 1] = hn[0].y;
            tt[H1 + 2 * up] = c[0].x;  tt[H1 + 2 * up + 1] = c[0].y;
        }
        cur ^= 1;
        __syncthreads();
    }
    (void)risk;
}

// pad_prefix[t][up][gate][2] = fma chain over k < H1 of pad_state[t+1].h[k] * W2[k][col]
// (exactly the first H1 terms of layer 2's x.W chain at a pad step)
template <int H1, int H2>
__global__ void k_pad_prefix(const float *__restrict__ pad_state, const float *__restrict__ W2,
                             int T, float *__restrict__ prefix)
{
    const int t = blockIdx.x;
    const int col = threadIdx.x;              // 0 .. 4*H2-1
    if (t >= T || col >= 4 * H2) return;
    const float *h = pad_state + (size_t)(t + 1) * 2 * H1;
    float acc = 0.f;
    for (int k = 0; k < H1; k++) acc = pb::ffma(h[k], W2[(size_t)k * 4 * H2 + col], acc);
    const int gate = col / H2, u = col % H2;
    prefix[(((size_t)t * (H2 / 2) + (u >> 1)) * 4 + gate) * 2 + (u & 1)] = acc;
}

// ============================================================================
// Demultiplexer, layer 2: LSTMCell(H2, impl 2) over concat(fwd, bwd), then
// Dense(n_classes) + softmax + the decision rule of barcoding.py:108-118.
// ============================================================================
constexpr int DEMUX_L2_GROUPS = 2;

template <int H1, int H2, int GROUPS, bool EXACT>
__global__ void __launch_bounds__(GROUPS * (H2 / 2) * NRG, 1)
k_demux_l2(const DemuxArgs A)
{
    constexpr int NUP = H2 / 2;
    constexpr int GT = NUP * NRG;             // threads per read group
    bool risk = false;
    constexpr int KX = 2 * H1;
    extern __shared__ __align__(16) float smem[];
    float *Wt = smem;                         // [KX][NUP][4][2]
    float *Ut = Wt + KX * 4 * H2;             // [H2][NUP][4][2]
    const int group = threadIdx.x / GT;
    const int tid = threadIdx.x % GT;
    float *hs = Ut + H2 * 4 * H2 + group * (2 * H2 * TB + KX * TB);   // [2][H2][TB]
    float *xs = hs + 2 * H2 * TB;             // [KX][TB]
    const int rg = tid % NRG, up = tid / NRG;
    const int64_t tile = (int64_t)blockIdx.x * GROUPS + group;
    const int64_t tile0 = tile * TB;

    int64_t n_eff = A.n;
    if (A.slot_count) {
        n_eff = (int64_t)*A.slot_count - A.row0;
        if (n_eff > A.n) n_eff = A.n;
    }
    if ((int64_t)blockIdx.x * GROUPS * TB >= n_eff) return;      // whole CTA idle
    load_weights<KX, H2>(A.W2, Wt);
    load_weights<H2, H2>(A.U2, Ut);
    __syncthreads();
    if (tile0 >= n_eff) return;
    float2 b[4];
    load_pair<H2>(A.b2, up, b);
    float2 c[RG], hz[RG];
#pragma unroll
    for (int r = 0; r < RG; r++) { c[r] = make_float2(0.f, 0.f); hz[r] = c[r]; }
    store_h(hs, up, rg, hz);

    const float *Gt = A.G + (size_t)tile * A.T * KX * TB;
    const int t_pad = (A.pad_prefix && A.tile_tstart) ? A.tile_tstart[tile] : 0;
    int cur = 0;
    Acc acc;
    for (int t = 0; t < A.T; t++) {
        group_sync(group, GT);                      // xs free (and h stores of step t-1 visible)
        const float4 *src = reinterpret_cast<const float4 *>(Gt + (size_t)t * KX * TB);
        for (int i = tid; i < KX * TB / 4; i += GT)
            reinterpret_cast<float4 *>(xs)[i] = __ldg(src + i);
        group_sync(group, GT);
        if (t < t_pad) {
            // every read of the tile is in its left padding: the forward half of the input
            // (k < H1) is the same for all of them, its partial chain is tabulated
            const float4 *pf = reinterpret_cast<const float4 *>(
                A.pad_prefix + ((size_t)t * NUP + up) * 8);
            const float4 p0 = __ldg(pf), p1 = __ldg(pf + 1);
#pragma unroll
            for (int r = 0; r < RG; r++) {
                acc.v[r][0] = make_float2(p0.x, p0.y); acc.v[r][1] = make_float2(p0.z, p0.w);
                acc.v[r][2] = make_float2(p1.x, p1.y); acc.v[r][3] = make_float2(p1.z, p1.w);
            }
            dot_tile_range<H1, KX, NUP>(Wt, xs, up, rg, acc);
        } else {
            dot_tile<KX, NUP>(Wt, xs, up, rg, acc);
        }
        float2 zx[RG][4];
#pragma unroll
        for (int r = 0; r < RG; r++)
#pragma unroll
            for (int g = 0; g < 4; g++) zx[r][g] = acc.v[r][g];
        dot_tile<H2, NUP>(Ut, hs + cur * H2 * TB, up, rg, acc);
        float2 hn[RG];
#pragma unroll
        for (int r = 0; r < RG; r++) {
            float2 z[4];
#pragma unroll
            for (int g = 0; g < 4; g++) z[g] = fadd2(fadd2(zx[r][g], acc.v[r][g]), b[g]);
            hn[r] = cell_pair<EXACT>(z, c[r], risk);
        }
        store_h(hs + (cur ^ 1) * H2 * TB, up, rg, hn);
        cur ^= 1;
    }
    (void)risk;
    group_sync(group, GT);

    if (tid < TB) {
        const int64_t row = tile0 + tid;
        // results go to the read that owns the row
        const int64_t r = (row < n_eff && A.slot_read) ? (int64_t)A.slot_read[A.row0 + row] : row;
        if (row < n_eff && (A.pushed == nullptr || A.pushed[r])) {
            DemuxCall call;
            demux_head<H2>(hs + cur * H2 * TB + tid, TB, A.Wd, A.bd, A.n_classes, A.n_decoy,
                           A.score_threshold, c_calibration, A.n_calibration, call);
            if (A.class_probs) {
#pragma unroll
                for (int j = 0; j < PB2_MAX_CLASSES; j++)
                    A.class_probs[r * PB2_MAX_CLASSES + j] = call.probs[j];
            }
            if (A.barcode) A.barcode[r] = call.barcode;
            if (A.guess) A.guess[r] = call.guess;
            if (A.score) A.score[r] = call.score;
        }
    }
}

int launch_demux_exact(pb2_context *ctx, const float *windows, const int32_t *pushed, int64_t n,
                       const int *slot_count, const int32_t *slot_read,
                       float *class_probs, int32_t *barcode, int32_t *guess, int32_t *score,
                       cudaStream_t st)
{
    if (n <= 0) return PB2_OK;
    const DemuxDev &D = ctx->demux;
    if (D.fwd.units != 48 || D.bwd.units != 48 || D.l2.units != 64 || D.fwd.in_dim != 1 ||
        D.l2.in_dim != 96 || D.fwd.impl != 2 || D.bwd.impl != 2 || D.l2.impl != 2)
        return fail(ctx, PB2_EUNSUPPORTED, "demux network shape not built "
                    "(Bidirectional(LSTMCell 48) -> LSTMCell 64, impl 2 expected)");
    constexpr int H1 = 48, H2 = 64;
    const int T = D.trim_length;
    const int64_t tiles = (n + TB - 1) / TB;
    // layer-1 scratch is tiles*T*96*32 floats (~3.5 MiB per tile at T=300): bound it
    const size_t per_tile = sizeof(float) * (size_t)T * 2 * H1 * TB;
    int64_t tiles_per_pass = (int64_t)(((size_t)4 << 30) / per_tile);
    if (tiles_per_pass > tiles) tiles_per_pass = tiles;
    float *G = (float *)ws_get(ctx, ctx->ws_h1, per_tile * (size_t)tiles_per_pass);
    if (!G) return PB2_ENOMEM;

    bool &attr_done = ctx->attr_demux;
    const size_t smem1 = sizeof(float) * (H1 * 4 * H1 + 2 * H1 * TB);
    constexpr int G2 = DEMUX_L2_GROUPS;
    const size_t smem2 = sizeof(float) * (2 * H1 * 4 * H2 + H2 * 4 * H2 +
                                          G2 * (2 * H2 * TB + 2 * H1 * TB));
    if (!attr_done) {
        PB_CUDA(ctx, cudaFuncSetAttribute(k_demux_l1<H1, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        PB_CUDA(ctx, cudaFuncSetAttribute(k_demux_l1<H1, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        PB_CUDA(ctx, cudaFuncSetAttribute(k_demux_l2<H1, H2, G2, false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        PB_CUDA(ctx, cudaFuncSetAttribute(k_demux_l2<H1, H2, G2, true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        attr_done = true;
    }
    const bool use_pad = D.pad_state && D.pad_prefix && !ctx->no_pad_skip;
    int *tstart = (int *)ws_get(ctx, ctx->ws_tstart, sizeof(int) * (size_t)tiles_per_pass);
    if (!tstart) return PB2_ENOMEM;
    PB_CUDA(ctx, cudaMemcpyToSymbolAsync(c_calibration, D.calibration,
                                         sizeof(double) * PB2_MAX_CALIB, 0,
                                         cudaMemcpyHostToDevice, st));
    for (int64_t t0 = 0; t0 < tiles; t0 += tiles_per_pass) {
        const int64_t nt = (tiles - t0 < tiles_per_pass) ? tiles - t0 : tiles_per_pass;
        const int64_t r0 = t0 * TB;
        DemuxArgs A = {};
        A.windows = windows + r0 * T;
        A.n = (n - r0 < nt * TB) ? n - r0 : nt * TB;
        A.T = T;
        A.Wf = D.fwd.kernel; A.Uf = D.fwd.recurrent; A.bf = D.fwd.bias;
        A.Wb = D.bwd.kernel; A.Ub = D.bwd.recurrent; A.bb = D.bwd.bias;
        A.W2 = D.l2.kernel; A.U2 = D.l2.recurrent; A.b2 = D.l2.bias;
        A.Wd = D.dense_kernel; A.bd = D.dense_bias;
        A.G = G;
        A.slot_count = slot_count; A.slot_read = slot_read; A.row0 = r0;
        A.pad_state = use_pad ? D.pad_state : nullptr;
        A.pad_prefix = use_pad ? D.pad_prefix : nullptr;
        A.tile_tstart = use_pad ? tstart : nullptr;
        A.pad_dump = nullptr;
        A.pad_value = D.pad_value;
        A.n_classes = D.n_classes; A.n_decoy = D.n_decoy;
        A.n_calibration = D.n_calibration; A.score_threshold = D.score_threshold;
        if (slot_read) {                 // outputs are indexed by read, not by row
            A.pushed = nullptr;
            A.class_probs = class_probs; A.barcode = barcode; A.guess = guess; A.score = score;
        } else {
            A.pushed = pushed ? pushed + r0 : nullptr;
            A.class_probs = class_probs ? class_probs + r0 * PB2_MAX_CLASSES : nullptr;
            A.barcode = barcode ? barcode + r0 : nullptr;
            A.guess = guess ? guess + r0 : nullptr;
            A.score = score ? score + r0 : nullptr;
        }
        if (ctx->exact_division) {
            PB_LAUNCH(ctx, K_DEMUX_L1, "k_demux_l1<exact>", st,
                k_demux_l1<H1, true><<<dim3((unsigned)nt, 2), (H1 / 2) * NRG, smem1, st>>>(A));
            PB_LAUNCH(ctx, K_DEMUX_L2, "k_demux_l2<exact>", st,
                k_demux_l2<H1, H2, G2, true><<<(unsigned)((nt + G2 - 1) / G2), G2 * (H2 / 2) * NRG, smem2, st>>>(A));
        } else {
            PB_LAUNCH(ctx, K_DEMUX_L1, "k_demux_l1", st,
                k_demux_l1<H1, false><<<dim3((unsigned)nt, 2), (H1 / 2) * NRG, smem1, st>>>(A));
            PB_LAUNCH(ctx, K_DEMUX_L2, "k_demux_l2", st,
                k_demux_l2<H1, H2, G2, false><<<(unsigned)((nt + G2 - 1) / G2), G2 * (H2 / 2) * NRG, smem2, st>>>(A));
        }
    }
    return PB2_OK;
}

// Tables for left-pad skipping, built with the exact kernels themselves: one all-pad window
// is stepped through the forward layer-1 kernel while its (h, c) are recorded, then the
// layer-2 chain prefixes are accumulated in the kernel's own fma order.
int build_pad_tables(pb2_context *ctx)
{
    DemuxDev &D = ctx->demux;
    cudaFree(D.pad_state); cudaFree(D.pad_prefix);
    D.pad_state = D.pad_prefix = nullptr;
    if (D.fwd.units != 48 || D.l2.units != 64 || D.fwd.in_dim != 1 || D.l2.in_dim != 96)
        return PB2_OK;                       // unsupported shape is reported by launch_demux
    constexpr int H1 = 48, H2 = 64;
    const int T = D.trim_length;
    float *state = nullptr, *prefix = nullptr, *win = nullptr, *G = nullptr;
    PB_CUDA(ctx, cudaMalloc(&state, sizeof(float) * (size_t)(T + 1) * 2 * H1));
    PB_CUDA(ctx, cudaMemset(state, 0, sizeof(float) * (size_t)(T + 1) * 2 * H1));
    PB_CUDA(ctx, cudaMalloc(&prefix, sizeof(float) * (size_t)T * 4 * H2));
    PB_CUDA(ctx, cudaMalloc(&G, sizeof(float) * (size_t)T * 2 * H1 * TB));
    std::vector<float> hw((size_t)T, D.pad_value);
    PB_CUDA(ctx, cudaMalloc(&win, sizeof(float) * (size_t)T));
    PB_CUDA(ctx, cudaMemcpy(win, hw.data(), sizeof(float) * (size_t)T, cudaMemcpyHostToDevice));
    DemuxArgs A = {};
    A.windows = win; A.n = 1; A.T = T;
    A.Wf = D.fwd.kernel; A.Uf = D.fwd.recurrent; A.bf = D.fwd.bias;
    A.Wb = D.bwd.kernel; A.Ub = D.bwd.recurrent; A.bb = D.bwd.bias;
    A.G = G; A.pad_dump = state; A.pad_value = D.pad_value;
    const size_t smem1 = sizeof(float) * (H1 * 4 * H1 + 2 * H1 * TB);
    PB_CUDA(ctx, cudaFuncSetAttribute(k_demux_l1<H1, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    k_demux_l1<H1, true><<<dim3(1, 1), (H1 / 2) * NRG, smem1, 0>>>(A);     // forward only
    k_pad_prefix<H1, H2><<<T, 4 * H2, 0, 0>>>(state, D.l2.kernel, T, prefix);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(win); cudaFree(G);
    if (e != cudaSuccess) { cudaFree(state); cudaFree(prefix); return check_cuda(ctx, e, "pad tables"); }
    D.pad_state = state;
    D.pad_prefix = prefix;
    return PB2_OK;
}

}  // namespace pb

namespace pb {
// G[tile][t][K][TB] -> out[row][t][K]
__global__ void k_debug_untile(const float *__restrict__ G, int64_t n, int T, int K, float *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * T * K) return;
    const int k = (int)(i % K);
    const int t = (int)((i / K) % T);
    const int64_t row = i / ((int64_t)K * T);
    out[i] = G[(((row / TB) * T + t) * K + k) * TB + row % TB];
}

// Verification: layer-1 outputs of the exact kernels for the first n <= 4096 windows,
// out[n][T][2*H1] (forward | backward), every position stepped (no pad skipping).
int debug_demux_l1(pb2_context *ctx, const float *windows, int64_t n, float *out, cudaStream_t st)
{
    const DemuxDev &D = ctx->demux;
    constexpr int H1 = 48;
    if (D.fwd.units != H1 || n <= 0 || n > 4096) return fail(ctx, PB2_EINVAL, "debug_demux_l1: bad size");
    const int T = D.trim_length;
    const int64_t tiles = (n + TB - 1) / TB;
    const size_t per_tile = sizeof(float) * (size_t)T * 2 * H1 * TB;
    float *G = (float *)ws_get(ctx, ctx->ws_h1, per_tile * (size_t)tiles);
    if (!G) return PB2_ENOMEM;
    const size_t smem1 = sizeof(float) * (H1 * 4 * H1 + 2 * H1 * TB);
    PB_CUDA(ctx, cudaFuncSetAttribute(k_demux_l1<H1, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    DemuxArgs A = {};
    A.windows = windows; A.n = n; A.T = T;
    A.Wf = D.fwd.kernel; A.Uf = D.fwd.recurrent; A.bf = D.fwd.bias;
    A.Wb = D.bwd.kernel; A.Ub = D.bwd.recurrent; A.bb = D.bwd.bias;
    A.G = G; A.pad_value = D.pad_value;
    k_demux_l1<H1, true><<<dim3((unsigned)tiles, 2), (H1 / 2) * NRG, smem1, st>>>(A);
    PB_LAUNCH_CHECK(ctx, "k_demux_l1<debug>");
    k_debug_untile<<<(unsigned)((n * T * 2 * H1 + 255) / 256), 256, 0, st>>>(G, n, T, 2 * H1, out);
    PB_LAUNCH_CHECK(ctx, "k_debug_untile");
    return PB2_OK;
}

// The demultiplexer as the rest of the library calls it: tensor-core path with exact
// re-check unless a verification mode asks for the exact kernels only.
int launch_demux(pb2_context *ctx, const float *windows, const int32_t *pushed, int64_t n,
                 const int *slot_count, const int32_t *slot_read,
                 float *class_probs, int32_t *barcode, int32_t *guess, int32_t *score,
                 cudaStream_t st)
{
    if (ctx->fast_lstm && !ctx->exact_division)
        return launch_demux_tc(ctx, windows, pushed, n, slot_count, slot_read, class_probs, barcode,
                               guess, score, nullptr, nullptr, nullptr, /*recheck=*/true, st);
    return launch_demux_exact(ctx, windows, pushed, n, slot_count, slot_read, class_probs, barcode,
                              guess, score, st);
}
}  // namespace pb
