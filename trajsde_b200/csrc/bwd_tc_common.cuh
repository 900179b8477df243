// Device helpers shared by the tensor-core backward kernels (euler_bwd_tc.cu, gru_bwd_tc.cu): MN-major UMMA descriptors over the
// [rows][64 x f16] SW128 operand tiles, and this thread's 32-channel half of an operand-tile row.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace trajsde {
namespace bwdtc {

using namespace tc;

// MN-major SW128 operand descriptor: tile stored [K rows][64 x f16 = 128 B]; LBO = byte stride between 64-element MN groups
// (the next tile of a stack), SBO = 1024 B between 8-row K groups; one MMA (K = 16 rows) advances the start address by 2048 B.
// Validated by bench_micro/mnmajor_test.cu.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptors, D = f32, A = B = f16: both operands MN-major / A K-major with B MN-major
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_k_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// this thread's 32 channels (4 swizzled 16-byte chunks) of an operand-tile row
__device__ __forceinline__ void st_row32(uint8_t* tile_row, uint32_t row, uint32_t hh, const float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t p[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) p[e] = pack_f16x2(v[q * 8 + 2 * e], v[q * 8 + 2 * e + 1]);
    *reinterpret_cast<uint4*>(tile_row + (((hh * 4 + q) ^ (row & 7u)) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
  }
}
__device__ __forceinline__ void unpack_f16x2(uint32_t w, float& lo, float& hi) {
  lo = __half2float(__ushort_as_half((unsigned short)(w & 0xffffu)));
  hi = __half2float(__ushort_as_half((unsigned short)(w >> 16)));
}
__device__ __forceinline__ void ld_row32(const uint8_t* tile_row, uint32_t row, uint32_t hh, float (&v)[32]) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(tile_row + (((hh * 4 + q) ^ (row & 7u)) << 4));
    unpack_f16x2(u.x, v[q * 8 + 0], v[q * 8 + 1]);
    unpack_f16x2(u.y, v[q * 8 + 2], v[q * 8 + 3]);
    unpack_f16x2(u.z, v[q * 8 + 4], v[q * 8 + 5]);
    unpack_f16x2(u.w, v[q * 8 + 6], v[q * 8 + 7]);
  }
}

// Coalesced row loads + transposition to "thread owns its row": lane L of a warp that owns tile rows r0 .. r0+31 loads, for
// i = 0..7, the 16-byte chunk (L & 7) of row r0 + 4 i + (L >> 3) of its 32-channel half (one instruction = four full 128-byte row
// segments); to_own_row() then exchanges the registers through 4 KB of per-warp shared-memory staging.
__device__ __forceinline__ void load_rows_coalesced(const float* slab, int64_t row_stride, int64_t first_row, int64_t n_rows, int col,
                                                    int lane, float4 (&dst)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = first_row + 4 * i + (lane >> 3);
    dst[i] = r < n_rows ? ld_nc_f4(slab + r * row_stride + col + (lane & 7) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
__device__ __forceinline__ void to_own_row(uint8_t* stage4k, int lane, float4 (&v)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t rl = 4 * i + (lane >> 3);
    *reinterpret_cast<float4*>(stage4k + rl * 128 + ((((uint32_t)lane & 7u) ^ (rl & 7u)) << 4)) = v[i];
  }
  __syncwarp();
#pragma unroll
  for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(stage4k + lane * 128 + (((uint32_t)q ^ ((uint32_t)lane & 7u)) << 4));
  __syncwarp();
}

// Column sums over the warp's 32 rows: on return lane L holds sum over lanes of v[L] (butterfly transpose-reduce, 31 shuffles).
__device__ __forceinline__ float colsum32(const float (&v)[32], int lane) {
  float a16[16], a8[8], a4[4], a2[2];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
    a16[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 8;
    const float send = up ? a16[i] : a16[i + 8], keep = up ? a16[i + 8] : a16[i];
    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 4;
    const float send = up ? a8[i] : a8[i + 4], keep = up ? a8[i + 4] : a8[i];
    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 2;
    const float send = up ? a4[i] : a4[i + 2], keep = up ? a4[i + 2] : a4[i];
    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  const bool up = lane & 1;
  const float send = up ? a2[0] : a2[1], keep = up ? a2[1] : a2[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// Weight-gradient flush: d[0..31] = v * scale (+ d[0..31]) for this thread's 32 consecutive floats of the CTA's partial vector.
// The lanes of a warp write 128-byte rows that lie 256+ bytes apart, so the cost is the number of store (and, when accumulating,
// load) instructions: 16-byte accesses where the row is 16-byte aligned, 8-byte ones otherwise (rows of the 66-wide W1 gradients
// start on odd multiples of 8 bytes) instead of 32 scalar read-modify-writes (measured: 27 k -> 9 k clk per launch of the one-step
// backward, bench_micro/bwd_fixed_cost.py).
__device__ __forceinline__ void flush_row32(float* d, const uint32_t (&v)[32], float scale, bool accumulate) {
  // accumulate: vector reductions (red.global.add.v4/v2.f32) instead of load + add + store — no round trip to L2 in front of the
  // stores.  Every address is owned by ONE thread of ONE CTA and launches are stream-ordered, so the sums stay bit-reproducible.
  if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 o = make_float4(__uint_as_float(v[4 * q]) * scale, __uint_as_float(v[4 * q + 1]) * scale,
                                   __uint_as_float(v[4 * q + 2]) * scale, __uint_as_float(v[4 * q + 3]) * scale);
      if (accumulate) atomicAdd(reinterpret_cast<float4*>(d + 4 * q), o);
      else *reinterpret_cast<float4*>(d + 4 * q) = o;
    }
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float2 o = make_float2(__uint_as_float(v[2 * q]) * scale, __uint_as_float(v[2 * q + 1]) * scale);
      if (accumulate) atomicAdd(reinterpret_cast<float2*>(d + 2 * q), o);
      else *reinterpret_cast<float2*>(d + 2 * q) = o;
    }
  }
}

}  // namespace bwdtc
}  // namespace trajsde
