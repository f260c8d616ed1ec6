// kernels_svb.cu -- streamvbyte-16 + zigzag + delta decoding of raw signals on the device.
//
// FAST5 files written by MinKNOW store the int16 `Signal` dataset through ONT's VBZ filter
// (HDF5 filter 32020): zstd( streamvbyte16( zigzag( delta(samples) ) ) ).  The reference reads it
// through h5py + the vbz plugin on the CPU (fast5_file.py:122-128).  Here the host only removes the
// zstd stage (poreplex_b200/csrc_host/fast5_loader.cpp) and hands the streamvbyte body over: about
// 1.13 bytes per sample instead of 2 cross the PCIe bus, and the rest of the decoder runs here.
//
// Stream layout for `count` samples (VBZ version 1, integer size 2):
//   keys   ceil(count / 8) bytes, bit k of byte g = 1 if value 8 g + k takes two bytes, else one
//   data   the values, little endian, 1 or 2 bytes each
//   value  v = zigzag(d) = (d << 1) ^ (d >> 15) of the 16-bit difference d = x[i] - x[i-1]
//          (x[-1] = 0, arithmetic modulo 2^16)
//
// One warp per read.  Per iteration a lane owns one key byte = 8 samples: its data offset is an
// exclusive warp scan of 8 + popcount(key), its values are decoded from at most 16 bytes, the
// delta prefix is a second warp scan of the lanes' sums; 8 int16 leave as one 16-byte store.
#include "pb_internal.h"

namespace pb {

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total)
{
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}

__global__ void __launch_bounds__(256)
k_svb16_decode(const uint8_t *__restrict__ packed, const int64_t *__restrict__ packed_offsets,
               const int64_t *__restrict__ raw_offsets, const int64_t *__restrict__ raw_lengths,
               int64_t n_reads, int16_t *__restrict__ raw, int32_t *__restrict__ error)
{
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n_reads) return;
    const int64_t count = raw_lengths[r];
    if (count <= 0) return;
    const uint8_t *keys = packed + packed_offsets[r];
    const int64_t avail = packed_offsets[r + 1] - packed_offsets[r];
    const int64_t ngroups = (count + 7) >> 3;
    const uint8_t *data = keys + ngroups;
    int16_t *out = raw + raw_offsets[r];
    int64_t data_pos = 0;                       // bytes of `data` consumed by earlier iterations
    uint32_t prev = 0;                          // x[i-1] (mod 2^16) carried across iterations
    bool bad = false;
    for (int64_t g0 = 0; g0 < ngroups; g0 += 32) {
        const int64_t g = g0 + lane;
        int nvalid = 0;
        unsigned key = 0;
        if (g < ngroups) {
            nvalid = (count - 8 * g >= 8) ? 8 : (int)(count - 8 * g);
            key = keys[g] & ((1u << nvalid) - 1u);
        }
        const int nbytes = nvalid + __popc(key);
        int total;
        const int off = warp_excl_scan(nbytes, lane, total);
        // the stream must hold every byte it promises (a damaged file must not read out of bounds)
        if (ngroups + data_pos + total > avail) { bad = true; break; }
        const uint8_t *p = data + data_pos + off;
        uint32_t x[8];
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            uint32_t v = 0;
            if (k < nvalid) {
                v = *p++;
                if ((key >> k) & 1u) v |= (uint32_t)(*p++) << 8;
            }
            const uint32_t d = (v >> 1) ^ (0u - (v & 1u));          // zigzag^-1 (mod 2^16 below)
            run += d;
            x[k] = run;
        }
        int tot2;
        const int base = warp_excl_scan((int)(run & 0xFFFFu), lane, tot2);
        const uint32_t start = prev + (uint32_t)base;
        if (nvalid == 8) {
            uint4 w;
            w.x = ((start + x[0]) & 0xFFFFu) | ((start + x[1]) << 16);
            w.y = ((start + x[2]) & 0xFFFFu) | ((start + x[3]) << 16);
            w.z = ((start + x[4]) & 0xFFFFu) | ((start + x[5]) << 16);
            w.w = ((start + x[6]) & 0xFFFFu) | ((start + x[7]) << 16);
            *reinterpret_cast<uint4 *>(out + 8 * g) = w;                // reads start 16-byte aligned
        } else {
            for (int k = 0; k < nvalid; k++) out[8 * g + k] = (int16_t)(uint16_t)(start + x[k]);
        }
        prev += (uint32_t)tot2;
        data_pos += total;
    }
    if (bad && lane == 0 && error) atomicExch(error, 1);
}

int launch_svb16_decode(pb2_context *ctx, const uint8_t *packed, const int64_t *packed_offsets,
                        const int64_t *raw_offsets, const int64_t *raw_lengths, int64_t n,
                        int16_t *raw, int32_t *error, cudaStream_t st)
{
    if (n <= 0) return PB2_OK;
    PB_LAUNCH(ctx, K_SVB_DECODE, "k_svb16_decode", st,
        k_svb16_decode<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(
            packed, packed_offsets, raw_offsets, raw_lengths, n, raw, error));
    return PB2_OK;
}

}  // namespace pb
