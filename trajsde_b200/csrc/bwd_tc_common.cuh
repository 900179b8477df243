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

// Same loads through L2 only (ld.global.cg): for rows another CTA of the SAME launch has written (the single-launch encoder sweep), where
// the non-coherent path of ld_nc_f4 could return a stale line.
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void load_rows_coalesced_cg(const float* slab, int64_t row_stride, int64_t first_row, int64_t n_rows, int col,
                                                       int lane, float4 (&dst)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = first_row + 4 * i + (lane >> 3);
    dst[i] = r < n_rows ? ld_cg_f4(slab + r * row_stride + col + (lane & 7) * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- single-launch encoder sweep (enc_bwd_sweep.cu): per-tile progress counters in global memory -------------------------------------
// The GRU role and the SDE role(s) of the merged kernel hand a 128-row tile back and forth once per iteration of the recurrence:
//   gru_done[tile]      = number of iterations whose GRU backward has written dL/dy1 of the tile       (S - i after iteration i)
//   sde_done[pass][tile] = number of iterations whose SDE-step backward has written the carried adjoint of the tile (that pass's rows)
// Writers: every epilogue thread stores its rows, CTA-local named barrier, then ONE thread publishes with st.release.gpu (the release is
// cumulative over the stores ordered before it by the barrier — the split-K semaphore pattern; a __threadfence() per thread in front of the
// barrier cost 2.5 us per GRU tile).
// Readers: every thread polls with ld.acquire.gpu and then reads the rows with ld.global.cg.  All CTAs of the launch are co-resident (grid
// <= SM count, one CTA per SM), so the polls cannot starve the producers; a bounded spin (about 30 s) plus a launch-wide abort word turns a
// scheduling accident into a status bit instead of a hung GPU.
struct SweepCtl {
  int S;                       // iterations of the recurrence (0: not a sweep)
  int32_t* gru_done;           // [num_tiles]
  int32_t* sde_done[2];        // [num_tiles] per pass (sde_done[1] == NULL: single diffusion net)
  int32_t* abort_word;         // launch-wide: set by the first wait that timed out
  int32_t* status;             // TrajsdeEncBwdArgs.status or NULL
  // SDE role: where iteration i finds its state and writes its result
  const float* h0;             // state of iteration 0
  const float* latent;         // [S][rows][64]: state of iteration i > 0 = latent[i-1]
  float* carry;                // [rows][64]: dL/dy0 of iteration i > 0
  float* grad_h0;              // dL/dy0 of iteration 0
  int64_t slab;                // rows * 64
};
__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int32_t* p, int32_t v) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
#ifdef TRAJSDE_SWEEP_TIMELINE
// debug build only (bench_micro/enc_bwd_ab.py --timeline): per CTA, thread 0: [0] kernel clocks, [1] clocks inside sweep_wait, [2] inside sweep_publish, [3] waits that blocked
static __device__ long long g_sweep_tl[160 * 4];
#define SWEEP_TL_ADD(slot, v) do { if (threadIdx.x == 0) g_sweep_tl[blockIdx.x * 4 + (slot)] += (v); } while (0)
#else
#define SWEEP_TL_ADD(slot, v) do { } while (0)
#endif
__device__ __forceinline__ void sweep_wait_(const SweepCtl& sw, const int32_t* counter, int32_t target) {
  if (ld_acquire_gpu(counter) >= target) return;
  SWEEP_TL_ADD(3, 1);
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (ld_acquire_gpu(counter) < target) {
    __nanosleep(40);
    if ((++spins & 63u) == 0) {
      if (*reinterpret_cast<volatile int32_t*>(sw.abort_word) != 0) return;
      // ~36 s at 1.9 GHz: give up, flag the call.  Long on purpose: CTAs of this launch may legitimately sit behind another kernel of the
      // process that holds a few SMs (an NCCL all-reduce on a side stream waiting for a straggler rank) and start late
      if (clock64() - t0 > (1ll << 36)) {
        atomicExch(sw.abort_word, 1);
        if (sw.status) atomicOr(sw.status, TRAJSDE_STATUS_SWEEP_TIMEOUT);
        return;
      }
    }
  }
}
__device__ __forceinline__ void sweep_wait(const SweepCtl& sw, const int32_t* counter, int32_t target) {
#ifdef TRAJSDE_SWEEP_TIMELINE
  const long long t0 = clock64();
#endif
  sweep_wait_(sw, counter, target);
#ifdef TRAJSDE_SWEEP_TIMELINE
  SWEEP_TL_ADD(1, clock64() - t0);
#endif
}
// all `nthreads` callers have stored their rows of the tile: publish `value` on `counter`
__device__ __forceinline__ void sweep_publish(int32_t* counter, int32_t value, uint32_t bar_id, uint32_t nthreads, bool leader) {
#ifdef TRAJSDE_SWEEP_TIMELINE
  const long long t0 = clock64();
#endif
  named_bar_sync(bar_id, nthreads);
  if (leader) st_release_gpu(counter, value);
#ifdef TRAJSDE_SWEEP_TIMELINE
  SWEEP_TL_ADD(2, clock64() - t0);
#endif
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
