// tc_probe.cu -- stand-alone check of the tcgen05 building blocks in csrc/tc_core.cuh:
// D[128][N] = A[128][K] * B[K][N] with A in TMEM (fp16 hi/lo), B in shared memory (fp16 hi/lo,
// canonical no-swizzle K-major), three split products, read back with tcgen05.ld.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tools/tc_probe.cu
//   ./tc_probe [variant]      variant 0 = layout as designed, 1 = LBO/SBO swapped
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../poreplex_b200/csrc/tc_core.cuh"

using namespace pb::tc;

template <int H, int K>
__global__ void __launch_bounds__(128, 1)
k_probe(const float *__restrict__ A, const float *__restrict__ W, float *__restrict__ D,
        int variant, int *err)
{
    constexpr int N = 4 * H;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __half *b_hi = reinterpret_cast<__half *>(smem_raw);
    __half *b_lo = b_hi + K * N;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ int s_dead;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); s_dead = 0; }
    if (warp == 0) tmem_alloc(&s_tmem, 512);
    load_b_split<K, H>(W, b_hi, b_lo, tid, blockDim.x);
    fence_proxy_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tbase = s_tmem;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t d_col = 0, ahi_col = N, alo_col = N + K / 2;
    // A row of this thread -> TMEM (packed fp16 hi / lo)
    for (int k0 = 0; k0 < K; k0 += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            hi[j] = split2(A[tid * K + k0 + 2 * j], A[tid * K + k0 + 2 * j + 1], lo[j]);
        tmem_st4(tbase + lane_base + ahi_col + k0 / 2, hi[0], hi[1], hi[2], hi[3]);
        tmem_st4(tbase + lane_base + alo_col + k0 / 2, lo[0], lo[1], lo[2], lo[3]);
    }
    tmem_st_wait();
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        bool first = true;
        if (variant == 0) {
            issue_split_gemm<K, N>(tbase + d_col, tbase + ahi_col, tbase + alo_col,
                                   smem_u32(b_hi), smem_u32(b_lo), first);
        } else {
            constexpr uint32_t LBO = N * 16, SBO = 128;
            constexpr uint32_t idesc = idesc_f16(128, N);
            for (int j = 0; j < K / 16; j++) {
                const uint64_t dh = smem_desc(smem_u32(b_hi) + j * 2 * LBO, SBO, LBO);
                const uint64_t dl = smem_desc(smem_u32(b_lo) + j * 2 * LBO, SBO, LBO);
                mma_ts(tbase + d_col, tbase + ahi_col + j * 8, dh, idesc, first ? 0u : 1u);
                first = false;
                mma_ts(tbase + d_col, tbase + ahi_col + j * 8, dl, idesc, 1u);
                mma_ts(tbase + d_col, tbase + alo_col + j * 8, dh, idesc, 1u);
            }
        }
        mma_commit(&bar);
    }
    mbar_wait(&bar, 0, &s_dead);
    fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tbase + lane_base + d_col + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j++) D[tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 512);
    if (tid == 0 && s_dead) *err = 1;
}

template <int H, int K>
static int run(int variant)
{
    constexpr int N = 4 * H;
    std::vector<float> A(128 * K), W(K * N), D(128 * N);
    srand(1234 + H + K);
    for (auto &v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (auto &v : W) v = ((float)rand() / RAND_MAX * 2.f - 1.f) * 1.5f;
    float *dA, *dW, *dD;
    int *derr, herr = 0;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMalloc(&derr, 4); cudaMemset(derr, 0, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = (size_t)K * N * 2 * 2;
    cudaFuncSetAttribute(k_probe<H, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_probe<H, K><<<1, 128, smem>>>(dA, dW, dD, variant, derr);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("H=%d K=%d variant %d: CUDA error %s\n", H, K, variant, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < N; n++) {
            const int u = (n / 8) * 2 + (n & 1), g = (n % 8) / 2;      // inverse of gate_col
            double ref = 0;
            for (int k = 0; k < K; k++) ref += (double)A[m * K + k] * (double)W[k * N + g * H + u];
            maxerr = fmax(maxerr, fabs(ref - (double)D[m * N + n]));
            maxref = fmax(maxref, fabs(ref));
        }
    printf("H=%d K=%d variant %d: max |err| %.3e (max |ref| %.3f) timeout=%d -> %s\n", H, K, variant,
           maxerr, maxref, herr, (maxerr < 1e-4 && !herr) ? "OK" : "MISMATCH");
    return (maxerr < 1e-4 && !herr) ? 0 : 1;
}

int main(int argc, char **argv)
{
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    int rc = 0;
    rc |= run<48, 48>(variant);
    rc |= run<48, 96>(variant);
    rc |= run<64, 64>(variant);
    rc |= run<64, 96>(variant);
    return rc;
}
