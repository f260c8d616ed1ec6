// tc_core.cuh -- tcgen05 / TMEM / mbarrier primitives for the tensor-core LSTM path
// (sm_100a only; inline PTX, no CUTLASS).
//
// Layout conventions used by every kernel that includes this file
//   * one CTA owns a tile of 128 reads = the 128 TMEM lanes (M = 128, cta_group::1);
//   * accumulators D[128][N] are fp32, one TMEM column per output column;
//   * the A operand (recurrent state h, or a layer input x) lives in TMEM as packed fp16
//     pairs: element (read m, k) is half (k & 1) of column (k >> 1) of lane m;
//   * the B operand (weights) lives in shared memory, fp16, K-major, no swizzle:
//     8x8 "core matrices" of 128 contiguous bytes (8 rows n x 16 bytes of k), laid out
//     [k / 8][n / 8] -> stride between row groups SBO = 128 B, stride between the two
//     k-halves of one K = 16 slab LBO = N * 16 B;
//   * products are fp16 x fp16 -> fp32 exact; an fp32 operand v is split v = hi + lo with
//     hi = fp16(v), lo = fp16(v - hi) and a product a*b is evaluated as the three MMAs
//     a_hi*b_hi + a_hi*b_lo + a_lo*b_hi (relative error about 3 * 2^-22).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace pb {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier ------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}"
                 ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait.  A protocol bug must not hang the GPU: after about a second the CTA is
// declared dead (*dead = 1), every later wait returns at once and the kernel runs to its
// end with garbage that the host rejects through the error word.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, volatile int *dead) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (*dead) return;
        if (clock64() - t0 > 2000000000LL) { *dead = 1; return; }
    }
}

// ---- fences --------------------------------------------------------------------
__device__ __forceinline__ void fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {     // generic st.shared -> MMA reads
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMEM allocation (one full warp) ---------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}

// ---- descriptors -----------------------------------------------------------------
// instruction descriptor, kind::f16: A, B fp16 (K-major), D fp32, M x N
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major, Blackwell version field = 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

// ---- MMA issue (one thread) ------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                       uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
                 : "memory");
}
// arrive on `bar` when every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM <-> registers (warp w touches lanes 32*(w%4) .. +31; thread = lane) ------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c,
                                         uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                 ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- fp32 -> (hi, lo) fp16 split ---------------------------------------------------
// packs two values: returns hi pair, writes lo pair (element 0 in the low half)
__device__ __forceinline__ uint32_t split2(float a, float b, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    lo = *reinterpret_cast<const uint32_t *>(&l);
    return *reinterpret_cast<const uint32_t *>(&h);
}

// ---- weights -> shared memory (canonical K-major, no swizzle) ----------------------
// Accumulator column of (unit u, gate g): the four gates of a PAIR of units sit in eight
// consecutive columns as (i,i', f,f', c,c', o,o'), so a thread reading its columns with
// tcgen05.ld gets aligned register pairs for the packed f32x2 gate arithmetic.
__host__ __device__ constexpr int gate_col(int u, int gate) { return (u >> 1) * 8 + gate * 2 + (u & 1); }

// Chunk-major unit order (NP > 0): a gate thread owns UPT = H / NP consecutive units and works
// through them in chunks of 8; the accumulator keeps chunk 0 of every thread part first, then
// chunk 1, ... so that "all first chunks" is ONE contiguous column group that a single set of
// MMAs produces (and can overwrite as soon as every thread has read its first chunk).
// NP = 0: natural order.
template <int H, int NP>
__host__ __device__ constexpr int unit_slot(int u) {
    if (NP == 0) return u;
    const int upt = H / NP, p = u / upt, w = u % upt;
    return (w / 8) * (NP * 8) + p * 8 + (w % 8);
}

// src: Keras matrix [K][4H] row-major, gate blocks i|f|c|o.  Output column n =
// gate_col(unit_slot(u), gate).  B_hi / B_lo: fp16 [K/8][N/8][8 (n)][8 (k)], N = 4H.
template <int K, int H, int NP = 0>
__device__ __forceinline__ void load_b_split(const float *__restrict__ src, __half *b_hi,
                                             __half *b_lo, int tid, int nthreads) {
    constexpr int N = 4 * H;
    for (int idx = tid; idx < K * N; idx += nthreads) {
        const int k = idx / N, col = idx % N;          // coalesced read of src
        const int gate = col / H, u = col % H;
        const int n = gate_col(unit_slot<H, NP>(u), gate);
        const float v = src[idx];
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const int o = ((k >> 3) * (N >> 3) + (n >> 3)) * 64 + (n & 7) * 8 + (k & 7);
        b_hi[o] = hi;
        b_lo[o] = lo;
    }
}

// issue the three split products for one operand pair over K (multiple of 16):
//   D (+)= A_hi B_hi + A_hi B_lo + A_lo B_hi
// a_hi / a_lo: TMEM column addresses of the packed fp16 A halves; b_hi / b_lo: smem byte
// addresses of the canonical B matrices with N output columns.
// full = false issues only the leading product A_hi B_hi (relative error about 2^-11): the
// deliberately perturbed evaluation used to measure how sensitive a read's result is.
// N = columns of the whole B matrix in shared memory (sets the stride between its k-chunks),
// NSUB = columns this call produces: d_tmem / b_hi / b_lo already point at the first of them
// (a column block of B starts (n0 / 8) * 128 bytes into each k-chunk).
template <int K, int N, int NSUB = N>
__device__ __forceinline__ void issue_split_gemm(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo,
                                                 uint32_t b_hi, uint32_t b_lo, bool &first,
                                                 bool full = true) {
    constexpr uint32_t LBO = N * 16, SBO = 128;
    constexpr uint32_t idesc = idesc_f16(128, NSUB);
#pragma unroll
    for (int j = 0; j < K / 16; j++) {
        const uint64_t dh = smem_desc(b_hi + j * 2 * LBO, LBO, SBO);
        const uint64_t dl = smem_desc(b_lo + j * 2 * LBO, LBO, SBO);
        mma_ts(d_tmem, a_hi + j * 8, dh, idesc, first ? 0u : 1u);
        first = false;
        if (full) {
            mma_ts(d_tmem, a_hi + j * 8, dl, idesc, 1u);
            mma_ts(d_tmem, a_lo + j * 8, dh, idesc, 1u);
        }
    }
}

}  // namespace tc
}  // namespace pb
