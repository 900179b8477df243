// Shared device/host helpers for the trajsde_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/trajsde_b200.h"

#define TS_DIM 64
#define TS_IN1 66  // y (64) + sin t + cos t: row length of net[0].weight

namespace trajsde {

// ---- thread-local error string (C-ABI: trajsde_last_error_string) -------------------------------------------------
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define TS_CUDA_CHECK(expr)                                                                              \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return ::trajsde::set_error(TRAJSDE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                                  __FILE__, __LINE__);                                                   \
  } while (0)

// ---- Philox4x32-10 + Box–Muller ----------------------------------------------------------------------------------
// Stream definition: key = (seed_lo, seed_hi); counter = (grow_lo, grow_hi, step, chunk) where
// grow = global row id, step = step_offset + k, chunk = channel/4.  One call -> 4 uint32 -> 4 increments N(0, h) for
// channels 4*chunk .. 4*chunk+3 via two Box–Muller pairs.
__host__ __device__ __forceinline__ uint32_t ts_mulhi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    if (r > 0) {
      k.x += W0;
      k.y += W1;
    }
    uint32_t hi0 = ts_mulhi(M0, c.x), lo0 = M0 * c.x;
    uint32_t hi1 = ts_mulhi(M1, c.z), lo1 = M1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  }
  return c;
}

// uniform in (0,1]: (x + 0.5) * 2^-32 evaluated in fp32 (never 0).
__device__ __forceinline__ float ts_u01(uint32_t x) { return fmaf((float)x, 2.3283064365386963e-10f, 1.1641532182693481e-10f); }

// 4 Brownian increments N(0, h) for (global row, step, chunk): Box–Muller on the four Philox words, radius scaled by sqrt(h).
// Every kernel (and trajsde_philox_dw, which dumps the stream for replays) calls THIS function, and every product is an explicit
// round-to-nearest multiply, so the increments are bit-identical wherever they are drawn.  The transcendental steps are single
// MUFU instructions (lg2 / sqrt / sin / cos .approx): u = (x + 0.5) 2^-32 lies in [1.2e-10, 1], so no denormal or range handling is
// needed, and the angle 2 pi u is formed directly by the integer -> float conversion FMA.
__device__ __forceinline__ float ts_lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ts_sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ts_sincos_approx(float x, float& s, float& c) {
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
}
// Effective Philox key of a call: the host-side seed plus, when the caller keeps its seed in device memory (TrajsdeNoise.seed_dev: CUDA-graph
// replays draw fresh noise by bumping that word between replays), the device word.
__device__ __forceinline__ uint64_t ts_noise_seed(const TrajsdeNoise& n) { return n.seed + (n.seed_dev ? *n.seed_dev : 0ull); }

__device__ __forceinline__ float4 philox_dw4(uint64_t seed, uint64_t grow, uint32_t step, uint32_t chunk, float sqrt_h) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)grow, (uint32_t)(grow >> 32), step, chunk),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  // sqrt(-2 ln u) = sqrt(-2 ln2 * lg2 u)
  const float r0 = __fmul_rn(ts_sqrt_approx(__fmul_rn(ts_lg2_approx(ts_u01(r.x)), -1.3862943611198906f)), sqrt_h);
  const float r1 = __fmul_rn(ts_sqrt_approx(__fmul_rn(ts_lg2_approx(ts_u01(r.z)), -1.3862943611198906f)), sqrt_h);
  float s0, c0, s1, c1;
  ts_sincos_approx(__fmaf_rn((float)r.y, 1.4629180792671596e-09f, 7.314590396335798e-10f), s0, c0);   // 2 pi (x + 0.5) 2^-32
  ts_sincos_approx(__fmaf_rn((float)r.w, 1.4629180792671596e-09f, 7.314590396335798e-10f), s1, c1);
  return make_float4(__fmul_rn(r0, c0), __fmul_rn(r0, s0), __fmul_rn(r1, c1), __fmul_rn(r1, s1));
}

// ---- small math helpers -------------------------------------------------------------------------------------------
__device__ __forceinline__ float ts_sigmoid_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float ts_tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// 256-bit global accesses (sm_100: LDG.256 / STG.256, 32-byte aligned): a thread that owns 128 contiguous bytes of a row moves them as four
// whole 32-byte sectors instead of eight half-used ones (a warp's 16-byte stores to 32 different rows touch 32 sectors for 512 bytes)
__device__ __forceinline__ void st_f8(float* p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z),
               "f"(b.w) : "memory");
}
__device__ __forceinline__ void st_cs_f8(float* p, float4 a, float4 b) {
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y),
               "f"(b.z), "f"(b.w) : "memory");
}
__device__ __forceinline__ void ld_nc_f8(const float* p, float4& a, float4& b) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void st_cs_f4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- launchers implemented in the kernel translation units --------------------------------------------------------
int launch_euler_fwd_exact(const TrajsdeEulerFwdArgs& a, cudaStream_t s);
int launch_euler_bwd_exact(const TrajsdeEulerBwdArgs& a, cudaStream_t s);
int64_t euler_bwd_exact_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual);
int launch_euler_bwd_tc(const TrajsdeEulerBwdArgs& a, cudaStream_t s);
int64_t euler_bwd_tc_workspace_bytes(int64_t rows, int32_t n_steps);
int launch_euler_fwd_tc(const TrajsdeEulerFwdArgs& a, cudaStream_t s);
int64_t euler_fwd_tc_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual);
int launch_enc_fwd_tc(const TrajsdeEncFwdArgs& a, cudaStream_t s);
int64_t enc_fwd_tc_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual);
// internal pieces of the tensor-core backward (euler_bwd_tc.cu), reused by the encoder backward
int bwd_tc_pack(const TrajsdeEulerBwdArgs& a, uint8_t* img, cudaStream_t s);
int bwd_tc_absmax(const float* x, int slabs, int64_t rows, int64_t slab_stride, int64_t row_stride, uint32_t* amax_bits, cudaStream_t s);
int bwd_tc_grid(int64_t rows, bool dual = false);
int bwd_tc_main(const TrajsdeEulerBwdArgs& a, const uint8_t* img0, const uint8_t* img1, const uint32_t* amax_bits, float* part0, float* part1,
                int accumulate, cudaStream_t s, bool pdl = false, const int32_t* row_map = nullptr, const int32_t* n_active = nullptr);
int launch_euler_bwd_reduce(const float* part0, const float* part1, int n0, int n1, const TrajsdeMlpGrad& gf, const TrajsdeMlpGrad& gg,
                            const TrajsdeMlpGrad& ga, cudaStream_t s);
// GRU jump backward (gru_bwd.cu) and the encoder-recurrence backward driver (enc_bwd.cu)
int gru_bwd_grid(int64_t rows);
int launch_gru_bwd(const TrajsdeGru& w, int64_t rows, const float* y1, const float* aa_out, int64_t slab, const uint8_t* obs_mask,
                   int64_t obs_mask_row_stride, const int32_t* slot, int iter, const float* carry, const float* grad_latent, float* grad_y1,
                   float* grad_aa_out, float* partial, cudaStream_t s);
int gru_bwd_tc_pack(const TrajsdeGru& w, uint8_t* img, cudaStream_t s);
int launch_gru_bwd_tc(int64_t rows, const float* y1, const float* aa_out, int64_t slab, const uint8_t* obs_mask, int64_t obs_mask_row_stride,
                      const int32_t* slot, int iter, const float* carry, const float* grad_latent, float* grad_y1, float* grad_aa_out,
                      const uint8_t* img, const uint32_t* amax_bits, float* partial, float* h_out_fwd_only, cudaStream_t s, bool pdl = false);
int launch_gru_standalone(const TrajsdeGruArgs& a, bool backward, cudaStream_t s);
int64_t gru_standalone_workspace_bytes(int64_t rows);
int launch_gru_bwd_reduce(const float* partial, int n, const TrajsdeGruGrad& g, cudaStream_t s);
int launch_enc_bwd(const TrajsdeEncBwdArgs& a, cudaStream_t s);
// single-launch form of the sweep (enc_bwd_sweep.cu)
int enc_bwd_sweep_grid(int64_t rows, bool dual, int* gg, int* gs);
int64_t enc_bwd_sweep_counter_bytes(int64_t rows);
int launch_enc_bwd_sweep(const TrajsdeEncBwdArgs& a, const TrajsdeEulerBwdArgs& b, const uint8_t* img0, const uint8_t* img1, const uint8_t* gru_img,
                         const uint32_t* amax_bits, float* part0, float* part1, float* gru_part, float* gbuf, float* carry, int32_t* counters,
                         int Gg, int Gs, cudaStream_t s);
int64_t enc_bwd_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual);
int launch_heads_fwd(const TrajsdeHeadsArgs& a, cudaStream_t s);
int64_t heads_workspace_bytes();
int launch_compact_rows(const uint8_t* flags, int64_t rows, int32_t* row_map, int32_t* n_active, cudaStream_t s);
int launch_aggr_embed(const TrajsdeAggrArgs& a, bool backward, cudaStream_t s);
int launch_pi_head(const TrajsdePiArgs& a, cudaStream_t s);
int64_t aggr_workspace_bytes(int64_t n_modes, int64_t n_actors);
int launch_l2_loss(const TrajsdeL2Args& a, bool backward, cudaStream_t s);
int64_t l2_workspace_bytes(int64_t n_actors);
int launch_diff_bce(const TrajsdeBceArgs& a, cudaStream_t s);
int64_t bce_workspace_bytes();
int launch_heads_bwd(const TrajsdeHeadsBwdArgs& a, cudaStream_t s);
int64_t heads_bwd_workspace_bytes();
int launch_philox_dw(const TrajsdeSchedule& sched, const TrajsdeNoise& noise, int64_t rows, float* out, cudaStream_t s);

}  // namespace trajsde
