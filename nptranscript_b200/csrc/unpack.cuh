// unpack.cuh — expansion of the compact transfer format (npt_packed_batch, include/npt.h) into the npt_batch layout the
// walks read.  Pure copy kernels: every output element depends on one input element at the same index, so there are no
// scans and no per-read work; both run at the HBM rate (852 B in + 1 681 B out per SARS-CoV-2 dRNA read).
//
// Why it exists: through npt_submit the path is bound by the host-to-device link (ncu / bench r01: 7.05 GB per 4 Mi reads at
// 52 GB/s = 135 of 140 ms), and half of those bytes are zeros — BAM spends 32 bits on a CIGAR element whose length fits 12
// and 4 bits on a base that is one of A C G T.
#pragma once
#include <cstdint>

namespace npt {

// 16-bit words -> 32-bit BAM words ((len<<4)|op zero-extends to itself); escaped elements are overwritten by patch_wide_kernel.
// thread = 8 elements; `in` and `out` are 16-byte aligned and padded to a multiple of 8 elements.
__global__ void unpack_cigar_kernel(const uint16_t* __restrict__ in, uint32_t* __restrict__ out, long long n) {
  const long long n8 = (n + 7) >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in) + i);
    uint4* o = reinterpret_cast<uint4*>(out) + 2 * i;
    o[0] = make_uint4(v.x & 0xFFFFu, v.x >> 16, v.y & 0xFFFFu, v.y >> 16);
    o[1] = make_uint4(v.z & 0xFFFFu, v.z >> 16, v.w & 0xFFFFu, v.w >> 16);
  }
}
__global__ void patch_wide_kernel(const uint32_t* __restrict__ idx, const uint32_t* __restrict__ word, long long n, uint32_t* out, long long n_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if ((long long)idx[i] < n_out) out[idx[i]] = word[i];
}

// two seq2 bytes (8 bases, 2 bits each, first base in the top bits of the first byte; x = first byte | second byte << 8) ->
// four seq4 bytes as one little-endian word (BAM codes A C G T = 1 2 4 8 = 1 << code, first base in the high nibble)
__device__ __forceinline__ uint32_t unpack_8_bases(uint32_t x) {
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;                      // nibble f of a 16-bit half = 2-bit field f of that input byte
  const uint32_t b0 = x & 0x11111111u, b1 = (x >> 1) & 0x11111111u;
  const uint32_t hot = (~(b0 | b1) & 0x11111111u) | ((b0 & ~b1) << 1) | ((b1 & ~b0) << 2) | ((b0 & b1) << 3);
  return __byte_perm(hot, 0, 0x2301);                    // the byte with the first two bases of a seq2 byte comes first in memory
}
// thread = 8 seq2 bytes -> 16 seq4 bytes; both streams 16-byte aligned, seq4 padded to a multiple of 16 bytes
__global__ void unpack_seq_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n_out_bytes) {
  const long long n16 = (n_out_bytes + 15) >> 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(in) + i);
    reinterpret_cast<uint4*>(out)[i] = make_uint4(unpack_8_bases(v.x & 0xFFFFu), unpack_8_bases(v.x >> 16), unpack_8_bases(v.y & 0xFFFFu),
                                                  unpack_8_bases(v.y >> 16));
  }
}
__global__ void patch_exceptions_kernel(const unsigned long long* __restrict__ idx, const uint8_t* __restrict__ val, long long n, uint8_t* out,
                                        long long n_out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (idx[i] < (unsigned long long)n_out) out[idx[i]] = val[i];
}

}  // namespace npt
