// Tensor-core Euler–Maruyama forward (TRAJSDE_MODE_TC_F16): persistent sm_100a kernel, tcgen05 + TMEM + TMA.
//
// Same reference functions as euler_exact.cu (models/utils/sdeint.py:340-384,477-485,544; dec_hivt_nusargo_sde.py:119-127,
// 154-158,180-195; enc_hivt_nusargo_sde_sep2.py:390-398,436-440,462-482), different machine mapping:
//
//   * one CTA per SM, two 128-row tiles ("slots") in flight so one slot's tcgen05.mma overlaps the other's epilogue;
//   * per slot: 8 epilogue warps — thread = (row, 32-channel half); its fp32 state stays in registers for ALL steps —
//     one MMA-issuer warp (one elected thread issues tcgen05.mma / tcgen05.commit) and one IO warp (TMA loads/stores,
//     per-step bias staging), so epilogue threads never block on a copy they did not need;
//   * drift/diffusion weights are packed once per call (fp16, 128B-swizzled K-major UMMA tiles) and staged once per CTA
//     into shared memory with a bulk TMA copy; the five 64x64 layers of a step are three dependent MMA phases:
//         P1: [z1f | z1g (| z1g_alt)] = y  . [W1y ; V1y (; V1y_alt)]^T        M=128, N=128 (192), K=64
//         P2:  z2f = h1f . W2^T ,  z2g = h1g . V2^T (, z2g_alt = h1g . V2alt^T) M=128, N=64 each
//         P3:  f   = h2f . W3^T                                                 M=128, N=64
//     accumulators live in TMEM (fp32) and are read back with tcgen05.ld; bias + MUFU tanh + fp16 pack happen in
//     registers and the next operand tile is written straight back to swizzled shared memory;
//   * g's last layer (64 -> 1), the sigmoid, the Euler update y' = y + f h + g dW and the output interpolation are fused
//     into the P3 epilogue in fp32;
//   * HBM traffic: y0 tile in (TMA), dW tile per step in (TMA, when caller-supplied), ys tile per output out (TMA store
//     from the same staging buffer), optional states tile per step out.  Nothing else touches global memory.
//   * dual diffusion (encoder): each row evaluates only ITS net's activations (the other net's MMA rows are ignored), so
//     the routing by nus_mask costs MMA columns but no extra MUFU work.
#include "common.cuh"
#include "tc_common.cuh"

namespace trajsde {

using namespace tc;

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_SLOTS = 2;
constexpr int EPI_WARPS_PER_SLOT = 8;                       // 4 TMEM lane quadrants x 2 column halves
constexpr int EPI_THREADS_PER_SLOT = EPI_WARPS_PER_SLOT * 32;
constexpr int NUM_EPI_WARPS = NUM_SLOTS * EPI_WARPS_PER_SLOT;
constexpr int WARP_MMA0 = NUM_EPI_WARPS;                    // warps 16,17: MMA issuers (slot 0,1)
constexpr int WARP_IO0 = NUM_EPI_WARPS + NUM_SLOTS;         // warps 18,19: IO (slot 0,1)
constexpr int NUM_THREADS = (NUM_EPI_WARPS + 2 * NUM_SLOTS) * 32;  // 640
constexpr int EPI_REGS = 112, AUX_REGS = 32;              // setmaxnreg split: epilogue warpgroups grow, MMA/IO warpgroup shrinks

// ---- packed weight image (bytes) -----------------------------------------------------------------------------------------
constexpr uint32_t IMG_B1 = 0;            // [192 rows][64] f16 SW128: W1y | V1y | V1y_alt
constexpr uint32_t IMG_W2 = 24576;        // [64][64]
constexpr uint32_t IMG_V2 = 32768;
constexpr uint32_t IMG_V2A = 40960;
constexpr uint32_t IMG_W3 = 49152;
constexpr uint32_t IMG_VEC = 57344;       // fp32 vectors
constexpr uint32_t IMG_BYTES = 59392;     // 58 KB
// single-diffusion variants (no alt net): the alt tiles make room for a BIAS tile whose K=16 slices carry every bias of the step
// through the tensor core (see "biases through the MMA" below)
constexpr uint32_t SIMG_B1 = 0;           // [128 rows][64] f16 SW128: W1y | V1y
constexpr uint32_t SIMG_BIAS = 16384;     // [128 rows][64] f16 SW128: slice 0 layer-1 bias + time columns, slice 1 b2 | c2, slice 2 b3
constexpr uint32_t SIMG_W2 = 32768, SIMG_V2 = 40960, SIMG_W3 = 49152;
#ifndef TRAJSDE_FWD_SEPARATE_DW
#define TRAJSDE_FWD_SEPARATE_DW 1       // supplied-dW single-diffusion variant: dW tile in the A1f|A1g region instead of the output staging buffer (A/B)
#endif
#ifndef TRAJSDE_FWD_BIAS_MMA
#define TRAJSDE_FWD_BIAS_MMA 1          // in-kernel-noise variants only (see BIAS_MMA); 0 adds the biases in the epilogues everywhere (A/B)
#endif
// fp32 vector slots (float index inside IMG_VEC)
constexpr int VEC_B2 = 0, VEC_C2 = 64, VEC_C2A = 128, VEC_B3 = 192, VEC_W3G = 256, VEC_W3GA = 320, VEC_C3 = 384, VEC_C3A = 385;
constexpr int BIAS1_LD = 192;             // per-step layer-1 bias row: b1f | c1 | c1_alt (time features folded in)

// ---- shared memory map -----------------------------------------------------------------------------------------------------
constexpr uint32_t SLOT_BYTES = 81920;    // A0 16K | A1f 16K | A1g 16K | X 32K
constexpr uint32_t OFF_A0 = 0, OFF_A1F = 16384, OFF_A1G = 32768, OFF_X = 49152;
constexpr uint32_t SMEM_SLOTS = IMG_BYTES;
constexpr uint32_t SMEM_RING = SMEM_SLOTS + NUM_SLOTS * SLOT_BYTES;      // per slot: 3 x [192] fp32 layer-1 bias ring
constexpr int RING_LD = BIAS1_LD + 8;           // bias row + {h, sqrt h, w0, w1, first output, #outputs, pad, pad}
constexpr uint32_t RING_BYTES = 3 * RING_LD * 4;
constexpr uint32_t SMEM_GPART = SMEM_RING + NUM_SLOTS * RING_BYTES;      // per slot: [2 halves][128 rows] fp32 partial g dots
constexpr uint32_t GPART_BYTES = 2 * TILE_M * 4;
constexpr uint32_t SMEM_BARS = SMEM_GPART + NUM_SLOTS * GPART_BYTES;
constexpr uint32_t SMEM_TOTAL = SMEM_BARS + 256;   // 1 + 2 x 10 mbarriers, TMEM base pointer
constexpr uint32_t SMEM_ALLOC = SMEM_TOTAL + 1024;  // slack for manual 1024-B alignment
static_assert(SMEM_ALLOC <= 232448, "exceeds 227 KB of shared memory per CTA");

struct TcParams {
  TrajsdeEulerFwdArgs a;
  const uint8_t* img;      // packed weight image in workspace
  const float* bias1;      // [S][192]
  int num_tiles;
  int dual;
};

// ---- weight packing -----------------------------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(TrajsdeEulerFwdArgs a, uint8_t* __restrict__ img, float* __restrict__ bias1, int dual) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  __half* b1 = reinterpret_cast<__half*>(img + IMG_B1);
  for (int idx = tid; idx < (dual ? 192 : 128) * 64; idx += nth) {
    const int n = idx >> 6, k = idx & 63, net = n >> 6, r = n & 63;
    const float* w = net == 0 ? a.drift.w1 : net == 1 ? a.diffusion.w1 : a.diffusion_alt.w1;
    *reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(b1) + sw128_off_h(n, k)) = __float2half_rn(w[r * TS_IN1 + k]);
  }
  for (int idx = tid; idx < 4 * 64 * 64; idx += nth) {
    const int m = idx >> 12, n = (idx >> 6) & 63, k = idx & 63;
    if (m == 2 && !dual) continue;
    const float* w = m == 0 ? a.drift.w2 : m == 1 ? a.diffusion.w2 : m == 2 ? a.diffusion_alt.w2 : a.drift.w3;
    const uint32_t off = dual ? (m == 0 ? IMG_W2 : m == 1 ? IMG_V2 : m == 2 ? IMG_V2A : IMG_W3) : (m == 0 ? SIMG_W2 : m == 1 ? SIMG_V2 : SIMG_W3);
    *reinterpret_cast<__half*>(img + off + sw128_off_h(n, k)) = __float2half_rn(w[n * 64 + k]);
  }
  if (!dual) {
    // BIAS tile: B operand rows n = output channel (0..63 drift, 64..127 diffusion), K slices of 16; the matching A slice holds, for
    // every row of the tile, (1, 1, s_hi, s_lo, s_hi, c_hi, c_lo, c_hi, 0...) with s = sin t0, c = cos t0 split into an fp16 head and
    // an fp16 remainder — so slice 0 adds b1 + W1[:,64] sin t0 + W1[:,65] cos t0 to ~2^-22, slices 1 / 2 add b2 | c2 and b3
    for (int idx = tid; idx < 128 * 64; idx += nth) {
      const int n = idx >> 6, k = idx & 63, slice = k >> 4, j = k & 15, net = n >> 6, r = n & 63;
      const TrajsdeMlp& m = net == 0 ? a.drift : a.diffusion;
      float v = 0.f;
      auto hi = [](float x) { return __half2float(__float2half_rn(x)); };
      if (slice == 0 && j < 8) {
        const float b = m.b1[r], ws = m.w1[r * TS_IN1 + 64], wc = m.w1[r * TS_IN1 + 65];
        v = j == 0 ? hi(b) : j == 1 ? b - hi(b) : (j == 2 || j == 3) ? hi(ws) : j == 4 ? ws - hi(ws) : (j == 5 || j == 6) ? hi(wc) : wc - hi(wc);
      } else if (slice == 1 && j < 2) {
        const float b = m.b2[r];
        v = j == 0 ? hi(b) : b - hi(b);
      } else if (slice == 2 && j < 2 && net == 0) {
        const float b = a.drift.b3[r];
        v = j == 0 ? hi(b) : b - hi(b);
      }
      *reinterpret_cast<__half*>(img + SIMG_BIAS + sw128_off_h(n, k)) = __float2half_rn(v);
    }
  }
  float* vec = reinterpret_cast<float*>(img + IMG_VEC);
  for (int i = tid; i < 512; i += nth) {
    float v = 0.f;
    const int c = i & 63;
    if (i < 64) v = a.drift.b2[c];
    else if (i < 128) v = a.diffusion.b2[c];
    else if (i < 192) v = dual ? a.diffusion_alt.b2[c] : 0.f;
    else if (i < 256) v = a.drift.b3[c];
    else if (i < 320) v = a.diffusion.w3[c];
    else if (i < 384) v = dual ? a.diffusion_alt.w3[c] : 0.f;
    else if (i == VEC_C3) v = a.diffusion.b3[0];
    else if (i == VEC_C3A) v = dual ? a.diffusion_alt.b3[0] : 0.f;
    vec[i] = v;
  }
  const int S = a.sched.n_steps;
  for (int idx = tid; idx < S * BIAS1_LD; idx += nth) {
    const int k = idx / BIAS1_LD, n = idx % BIAS1_LD, net = n >> 6, r = n & 63;
    const TrajsdeMlp* m = net == 0 ? &a.drift : net == 1 ? &a.diffusion : (dual ? &a.diffusion_alt : nullptr);
    float v = 0.f;
    if (m) {
      const float sn = a.sched.step_tab[4 * k + 2], cs = a.sched.step_tab[4 * k + 3];
      v = fmaf(m->w1[r * TS_IN1 + 65], cs, fmaf(m->w1[r * TS_IN1 + 64], sn, m->b1[r]));
    }
    bias1[idx] = v;
  }
}

// ---- epilogue helpers -------------------------------------------------------------------------------------------------------
// 32 bias values of this thread's column half -> registers.  Issued BEFORE the tcgen05.ld of the accumulators so the shared-memory
// latency hides behind the TMEM read instead of stalling every group of eight tanh (the asm volatile tcgen05.ld keeps the order).
struct Bias32 { float4 v[8]; };
__device__ __forceinline__ Bias32 ld_bias32(const float* __restrict__ bias) {
  Bias32 b;
#pragma unroll
  for (int q = 0; q < 8; ++q) b.v[q] = *reinterpret_cast<const float4*>(bias + 4 * q);
  return b;
}
// tanh(acc + bias) for 32 accumulator columns, packed to fp16 and stored as 4 x 16-byte chunks of an operand row.
__device__ __forceinline__ void act32_to_operand(const uint32_t (&v)[32], const Bias32& b, uint8_t* tile_row_base, uint32_t row, uint32_t chunk0) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 ba = b.v[2 * q], bb = b.v[2 * q + 1];
    uint32_t p[4];
    p[0] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 0]) + ba.x), ts_tanh_approx(__uint_as_float(v[q * 8 + 1]) + ba.y));
    p[1] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 2]) + ba.z), ts_tanh_approx(__uint_as_float(v[q * 8 + 3]) + ba.w));
    p[2] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 4]) + bb.x), ts_tanh_approx(__uint_as_float(v[q * 8 + 5]) + bb.y));
    p[3] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 6]) + bb.z), ts_tanh_approx(__uint_as_float(v[q * 8 + 7]) + bb.w));
    *reinterpret_cast<uint4*>(tile_row_base + (((chunk0 + q) ^ (row & 7u)) << 4)) = make_uint4(p[0], p[1], p[2], p[3]);
  }
}

// Same burst, result written as the A operand of the next MMA in TENSOR memory (16 packed columns of this thread's lane): no
// st.shared, no fence.proxy.async (= MEMBAR.ALL.CTA) — 260 clk less per dependent phase (bench_micro/tmem_a_test.cu).
__device__ __forceinline__ void act32_to_tmem(const uint32_t (&v)[32], const Bias32& b, uint32_t taddr) {
  uint32_t p[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 ba = b.v[2 * q], bb = b.v[2 * q + 1];
    p[4 * q + 0] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 0]) + ba.x), ts_tanh_approx(__uint_as_float(v[q * 8 + 1]) + ba.y));
    p[4 * q + 1] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 2]) + ba.z), ts_tanh_approx(__uint_as_float(v[q * 8 + 3]) + ba.w));
    p[4 * q + 2] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 4]) + bb.x), ts_tanh_approx(__uint_as_float(v[q * 8 + 5]) + bb.y));
    p[4 * q + 3] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[q * 8 + 6]) + bb.z), ts_tanh_approx(__uint_as_float(v[q * 8 + 7]) + bb.w));
  }
  tmem_st_32x32b_x16(taddr, p);
  tc_wait_st();
}

// Bias-free variant: the MMA chain already added the bias (BIAS tile), so the burst is LDTM -> 32 x MUFU.TANH -> 16 x F2FP -> STTM.
__device__ __forceinline__ void tanh32_to_tmem(const uint32_t (&v)[32], uint32_t taddr) {
  uint32_t p[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) p[e] = pack_f16x2(ts_tanh_approx(__uint_as_float(v[2 * e])), ts_tanh_approx(__uint_as_float(v[2 * e + 1])));
  tmem_st_32x32b_x16(taddr, p);
  tc_wait_st();
}

// This row's diffusion-net column block -> registers.  In a warp whose rows use both diffusion nets the alt block is fetched
// too and selected per lane (tcgen05.ld is warp-collective: its address must be warp-uniform).
template <bool DUAL>
__device__ __forceinline__ void ld_g(uint32_t tm_g_uniform, uint32_t tm_g_alt, bool w_mixed, bool use_alt, uint32_t (&vg)[32]) {
  tmem_ld_32x32b_x32(tm_g_uniform, vg);
  if (DUAL && w_mixed) {
    uint32_t v2[32];
    tmem_ld_32x32b_x32(tm_g_alt, v2);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) vg[j] = use_alt ? v2[j] : vg[j];
  } else {
    tc_wait_ld();
  }
}

struct Epi3Ctx {
  uint8_t* x_row;       // output staging row (and, unless dw_row points elsewhere, where the step's dW chunks are)
  uint8_t* dw_row;      // this thread's dW chunks of the step
  uint8_t* a0_row;
  uint8_t* st_row;
  const float* b3;      // vec + VEC_B3 + hh*32
  uint32_t tm_y, tm_f;  // TMEM addresses of this thread's 32 state / drift columns
  uint32_t tm_oa;       // TMEM address of this thread's 16 packed operand columns (TMEM_A variants)
  uint32_t row, hh;
  float h, g, w0, w1;
  bool has_out, save_states, valid;
  // slow path only
  int ob, nout;
  const float* out_w;
  float* ys_row;        // a.ys + grow * row_stride + hh*32
  int64_t ys_t_stride;
};

// Epilogue 3: f = z3 + b3 ; y' = y + f h + g dW (dW already in this thread's X chunks) ; outputs in place ; Y, A0 <- y'.
// Y_LOADED: the caller issued the tcgen05.ld of the state columns into `yv` before it waited for P3 (the read runs under that wait).
template <bool MULTI, bool TMEM_A, bool Y_LOADED = false, bool BIAS_IN_MMA = TMEM_A>
__device__ __forceinline__ void epi3_update(const Epi3Ctx& c, uint32_t (&yv)[32]) {
  uint32_t fv[32];
  if (!Y_LOADED) tmem_ld_32x32b_x32(c.tm_y, yv);
  tmem_ld_32x32b_x32(c.tm_f, fv);
  tc_wait_ld();
  if (c.save_states) {  // Y[k] -> states staging (aliases A1f|A1g: both consumed by P2 already)
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<uint4*>(c.st_row + ((q ^ (c.row & 7u)) << 4)) = make_uint4(yv[4 * q], yv[4 * q + 1], yv[4 * q + 2], yv[4 * q + 3]);
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float4* xp = reinterpret_cast<float4*>(c.x_row + ((q ^ (c.row & 7u)) << 4));
    const float4 dw = *reinterpret_cast<const float4*>(c.dw_row + ((q ^ (c.row & 7u)) << 4));
    float f0 = __uint_as_float(fv[4 * q]), f1 = __uint_as_float(fv[4 * q + 1]), f2 = __uint_as_float(fv[4 * q + 2]), f3 = __uint_as_float(fv[4 * q + 3]);
    if (!BIAS_IN_MMA) {                                      // otherwise b3 came through the MMA (BIAS tile, slice 2)
      const float4 b3 = *reinterpret_cast<const float4*>(c.b3 + 4 * q);
      f0 += b3.x; f1 += b3.y; f2 += b3.z; f3 += b3.w;
    }
    const float y0 = __uint_as_float(yv[4 * q]), y1 = __uint_as_float(yv[4 * q + 1]);
    const float y2 = __uint_as_float(yv[4 * q + 2]), y3 = __uint_as_float(yv[4 * q + 3]);
    const float n0 = fmaf(c.g, dw.x, fmaf(f0, c.h, y0));
    const float n1 = fmaf(c.g, dw.y, fmaf(f1, c.h, y1));
    const float n2 = fmaf(c.g, dw.z, fmaf(f2, c.h, y2));
    const float n3 = fmaf(c.g, dw.w, fmaf(f3, c.h, y3));
    if (c.has_out)
      *xp = make_float4(fmaf(c.w1, n0, c.w0 * y0), fmaf(c.w1, n1, c.w0 * y1), fmaf(c.w1, n2, c.w0 * y2), fmaf(c.w1, n3, c.w0 * y3));
    if (MULTI) {  // rare: one step completes several outputs (zero-step intervals) -> direct global stores
      if (c.valid) {
        for (int o = 1; o < c.nout; ++o) {
          const float v0 = c.out_w[2 * (c.ob + o)], v1 = c.out_w[2 * (c.ob + o) + 1];
          *reinterpret_cast<float4*>(c.ys_row + (int64_t)(c.ob + o + 1) * c.ys_t_stride + 4 * q) =
              make_float4(fmaf(v1, n0, v0 * y0), fmaf(v1, n1, v0 * y1), fmaf(v1, n2, v0 * y2), fmaf(v1, n3, v0 * y3));
        }
      }
    }
    yv[4 * q] = __float_as_uint(n0); yv[4 * q + 1] = __float_as_uint(n1);
    yv[4 * q + 2] = __float_as_uint(n2); yv[4 * q + 3] = __float_as_uint(n3);
  }
  tmem_st_32x32b_x32(c.tm_y, yv);
  if (TMEM_A) {
    uint32_t pk[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) pk[e] = pack_f16x2(__uint_as_float(yv[2 * e]), __uint_as_float(yv[2 * e + 1]));
    tmem_st_32x32b_x16(c.tm_oa, pk);
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) pk[e] = pack_f16x2(__uint_as_float(yv[q * 8 + 2 * e]), __uint_as_float(yv[q * 8 + 2 * e + 1]));
      *reinterpret_cast<uint4*>(c.a0_row + (((c.hh * 4 + q) ^ (c.row & 7u)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
  tc_wait_st();
}

#ifdef TRAJSDE_FWD_TIMELINE
// debug build only (bench_micro/fwd_timeline.py): clocks of thread 0 of CTA 0 — [0] kernel, [1] waiting for X to be released (previous
// step's output store), [2] waiting for the step's dW tile to land, [3] waiting for P3, [4] steps
__device__ long long g_fwd_tl[8];
__device__ long long g_fwd_seg[16];   // clocks between consecutive FWD_MARKs of a step (single-diffusion variants), thread 0 of CTA 0
#define FWD_TL(i, expr) do { const long long _t0 = clock64(); expr; if (threadIdx.x == 0 && blockIdx.x == 0) g_fwd_tl[i] += clock64() - _t0; } while (0)
#define FWD_MARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _t = clock64(); g_fwd_seg[i] += _t - seg_prev; seg_prev = _t; } } while (0)
#else
#define FWD_TL(i, expr) do { expr; } while (0)
#define FWD_MARK(i) do { } while (0)
#endif

template <bool HAS_DW, bool DUAL>
__global__ void __launch_bounds__(NUM_THREADS, 1)
euler_fwd_tc_kernel(const TcParams p, const __grid_constant__ CUtensorMap tm_y0, const __grid_constant__ CUtensorMap tm_dw,
                    const __grid_constant__ CUtensorMap tm_ys, const __grid_constant__ CUtensorMap tm_st) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const TrajsdeEulerFwdArgs& a = p.a;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int S = a.sched.n_steps;
#ifdef TRAJSDE_FWD_TIMELINE
  const long long tl_start = clock64();
  long long seg_prev = tl_start;
#endif
  // contiguous, balanced tile range of this CTA (sizes differ by at most one tile); its two slots take alternate tiles
  const int tiles_q = p.num_tiles / (int)gridDim.x, tiles_r = p.num_tiles % (int)gridDim.x;
  const int tile_lo = (int)blockIdx.x * tiles_q + min((int)blockIdx.x, tiles_r);
  const int tile_hi = tile_lo + tiles_q + ((int)blockIdx.x < tiles_r ? 1 : 0);
  const bool save_states = a.states != nullptr;
  const uint64_t noise_seed = HAS_DW ? 0ull : ts_noise_seed(a.noise);

  // mbarriers.  [0] weights; per slot s (stride 80 B): opnd[2] (256 epilogue arrivals each), acc[2] (tcgen05.commit), tma
  // (TMA tx), xfull (256: staging written / y0 consumed), xfree (IO: stores have read the staging buffers), ring[2] (IO:
  // step entry filled).  opnd/acc/ring come in pairs used alternately: a parity wait can only tell apart ONE outstanding phase,
  // and the fine-grained hand-offs below put two phases of the same kind in flight.
  const uint32_t bar_w = base + SMEM_BARS;
  auto bar_opnd = [&](int s, int i) { return base + SMEM_BARS + 8u + 80u * s + 8u * i; };
  auto bar_acc = [&](int s, int i) { return base + SMEM_BARS + 24u + 80u * s + 8u * i; };
  auto bar_tma = [&](int s) { return base + SMEM_BARS + 40u + 80u * s; };
  auto bar_xfull = [&](int s) { return base + SMEM_BARS + 48u + 80u * s; };
  auto bar_xfree = [&](int s) { return base + SMEM_BARS + 56u + 80u * s; };
  auto bar_ring = [&](int s, int i) { return base + SMEM_BARS + 64u + 80u * s + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + SMEM_BARS + 192);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int s = 0; s < NUM_SLOTS; ++s) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(bar_opnd(s, i), EPI_THREADS_PER_SLOT);
        mbar_init(bar_acc(s, i), 1);
        mbar_init(bar_ring(s, i), 32);                     // every lane of the IO warp releases its own ring stores
      }
      mbar_init(bar_tma(s), 1);
      mbar_init(bar_xfull(s), EPI_THREADS_PER_SLOT);
      mbar_init(bar_xfree(s), 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tm_y0);
    tma_prefetch_desc(&tm_dw);
    tma_prefetch_desc(&tm_ys);
    tma_prefetch_desc(&tm_st);
  }
  if (warp == WARP_MMA0) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (threadIdx.x == 0) {  // stage the packed weights once per CTA (bulk TMA copy)
    mbar_arrive_expect_tx(bar_w, IMG_BYTES);
    bulk_load_1d(base, p.img, IMG_BYTES, bar_w);
  }

  const float* vec = reinterpret_cast<const float*>(sm + IMG_VEC);

  if (warp < NUM_EPI_WARPS) {
    // =============================================== EPILOGUE WARPS ===============================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));   // 16x32x112 + 4x32x32 = 640 x 96: inc only draws on what dec released
    const int slot = warp / EPI_WARPS_PER_SLOT;
    const int wq = warp % EPI_WARPS_PER_SLOT;
    const int quad = wq & 3;                               // TMEM lane quadrant (= warp index % 4)
    const uint32_t hh = wq >> 2;                           // which 32-channel half of the row this thread owns
    const uint32_t row = quad * 32 + lane;                 // row inside the tile == TMEM lane
    uint8_t* slot_sm = sm + SMEM_SLOTS + slot * SLOT_BYTES;
    uint8_t* a0_row = slot_sm + OFF_A0 + row * 128;
    uint8_t* a1f_row = slot_sm + OFF_A1F + row * 128;
    uint8_t* a1g_row = slot_sm + OFF_A1G + row * 128;
    uint8_t* x_row = slot_sm + OFF_X + hh * 16384 + row * 128;      // this thread's 32 channels: 8 swizzled 16-B chunks
    const float* ring = reinterpret_cast<const float*>(sm + SMEM_RING + slot * RING_BYTES);
    float* gpart = reinterpret_cast<float*>(sm + SMEM_GPART + slot * GPART_BYTES);
    // TMEM columns of a slot: [0,192) P1/P2 accumulators (P3 reuses [0,64)), [192,256) the resident fp32 state Y
    const uint32_t tm_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + slot * 256 + hh * 32;
    // single-diffusion variants keep the MMA A operands in the TMEM columns the dual variant needs for its second diffusion net:
    // OA = [128,160): y -> h1f -> h2f -> y' (each written after the MMA that read the previous content has completed), OB = [160,192): h1g
    constexpr bool TMEM_A = !DUAL;
    constexpr bool BIAS_MMA = TMEM_A && !HAS_DW && (TRAJSDE_FWD_BIAS_MMA != 0);   // measured: +2 % with in-kernel noise, -2 % with supplied dW
    const uint32_t tm_oa = tmem_base + ((uint32_t)(quad * 32) << 16) + slot * 256 + 128 + hh * 16, tm_ob = tm_oa + 32;
    const uint32_t pair_bar = 1 + slot * 4 + quad;         // named barrier shared by the two warps that own the same rows
    uint32_t par_accA = 0, par_accB = 0, par_tma = 0, par_xfree = 0;
    uint32_t gstep = 0;
    Epi3Ctx c3;
    // Supplied dW, operands in tensor memory, no states to stage: the A1f|A1g region is free and takes the dW tile, so the IO warp can load
    // the next step's tile as soon as epilogue 3 has read this one — without waiting for the output store to drain X
    const bool sep_dw = HAS_DW && !DUAL && !save_states && (TRAJSDE_FWD_SEPARATE_DW != 0);
    c3.x_row = x_row;
    c3.dw_row = sep_dw ? slot_sm + OFF_A1F + hh * 16384 + row * 128 : x_row;
    c3.a0_row = a0_row;
    c3.st_row = slot_sm + OFF_A1F + hh * 16384 + row * 128;
    c3.b3 = vec + VEC_B3 + hh * 32;
    c3.tm_y = tm_lane + 192;
    c3.tm_oa = tm_oa;
    c3.tm_f = tm_lane;
    c3.row = row;
    c3.hh = hh;
    c3.save_states = save_states;
    c3.out_w = a.sched.out_w;
    c3.ys_t_stride = a.ys_t_stride;

    if (BIAS_MMA) {
      // biases through the MMA: the A0 tile (unused as an operand tile here: y / h1f / h2f live in tensor memory) holds the bias
      // operand rows.  K slice (gstep & 1) of every row = (1, 1, s_hi, s_lo, s_hi, c_hi, c_lo, c_hi, 0 x 8) for the step's sin t0 /
      // cos t0; the fifth MMA of every phase multiplies it with the BIAS tile's slice for that layer.  Zero everything once.
#pragma unroll
      for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(a0_row + (((hh * 4 + q) ^ (row & 7u)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    }
    mbar_wait(bar_w, 0);

    for (int tile = tile_lo + slot; tile < tile_hi; tile += NUM_SLOTS) {
      const int64_t row0 = (int64_t)tile * TILE_M;
      const int64_t grow = row0 + row;
      const bool valid = grow < a.rows;
      const bool use_alt = DUAL && valid && (a.alt_mask[grow] == 0);
      const int gcol = use_alt ? 128 : 64;                 // column block / bias offset of this row's diffusion net
      const bool w_all_alt = DUAL && __all_sync(0xffffffffu, use_alt);
      const bool w_mixed = DUAL && !w_all_alt && __any_sync(0xffffffffu, use_alt);
      const uint32_t ucol = w_all_alt ? 128 : 64;          // warp-uniform TMEM column block (mixed warps also read 128)
      const float* c2v = vec + (use_alt ? VEC_C2A : VEC_C2) + hh * 32;
      const float* w3v = vec + (use_alt ? VEC_W3GA : VEC_W3G) + hh * 32;
      const float c3b = vec[use_alt ? VEC_C3A : VEC_C3];
      c3.valid = valid;
      c3.ys_row = a.ys + grow * a.ys_row_stride + hh * 32;

      // ---- tile prologue: y0 tile (TMA -> X) -> Y (TMEM) + A0 ---------------------------------------------------------
      mbar_wait(bar_tma(slot), par_tma);
      par_tma ^= 1;
      {
        uint32_t yv[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 v = *reinterpret_cast<const uint4*>(x_row + ((q ^ (row & 7u)) << 4));
          yv[4 * q] = v.x; yv[4 * q + 1] = v.y; yv[4 * q + 2] = v.z; yv[4 * q + 3] = v.w;
        }
        tmem_st_32x32b_x32(tm_lane + 192, yv);
        if (TMEM_A) {
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) pk[e] = pack_f16x2(__uint_as_float(yv[2 * e]), __uint_as_float(yv[2 * e + 1]));
          tmem_st_32x32b_x16(tm_oa, pk);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) pk[e] = pack_f16x2(__uint_as_float(yv[q * 8 + 2 * e]), __uint_as_float(yv[q * 8 + 2 * e + 1]));
            *reinterpret_cast<uint4*>(a0_row + (((hh * 4 + q) ^ (row & 7u)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        tc_wait_st();
      }
      if (BIAS_MMA) {                                      // bias operand slice of this tile's first step (its ring entry is published)
        mbar_wait(bar_ring(slot, gstep & 1), (gstep >> 1) & 1);
        if (hh == 0)
          *reinterpret_cast<uint4*>(a0_row + (((2u * (gstep & 1u)) ^ (row & 7u)) << 4)) =
              *reinterpret_cast<const uint4*>(ring + (gstep % 3u) * RING_LD);
      }
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(bar_opnd(slot, 0));                      // A0 ready -> P1 of step 0
      mbar_arrive(bar_xfull(slot));                        // y0 consumed: IO may store X as ys[0] and then refill it

      for (int k = 0; k < S; ++k, ++gstep) {
        if constexpr (TMEM_A) {
          // ===================== single-diffusion variants: operands in tensor memory, biases through the MMA =====================
          mbar_wait(bar_ring(slot, gstep & 1), (gstep >> 1) & 1);   // this step's scalars are in the ring
          FWD_MARK(0);   // loop tail + ring wait
          const float* ent = ring + (gstep % 3u) * RING_LD;
          const float4 sc = *reinterpret_cast<const float4*>(ent + BIAS1_LD);      // h, sqrt(h), w0, w1
          const int2 so = *reinterpret_cast<const int2*>(ent + BIAS1_LD + 4);      // first output index, #outputs of this step
          // In-kernel noise: the step's 8 Philox calls of this thread are spread over the three gaps in which the thread would
          // otherwise only wait for the tensor core (after epilogue 1, inside epilogue 2, before epilogue 3), so the draw never
          // forms one 500-instruction block in front of the P3 wait.
          auto draw = [&](int q0, int q1) {
#pragma unroll
            for (int q = q0; q < q1; ++q)
              *reinterpret_cast<float4*>(x_row + ((q ^ (row & 7u)) << 4)) =
                  philox_dw4(noise_seed, (uint64_t)grow + a.noise.row_offset, a.noise.step_offset + (uint32_t)k, (uint32_t)(hh * 8 + q), sc.y);
          };

          // ---- epilogue 1: h1f = tanh(z1f) -> OA (P2f may start), h1g = tanh(z1g) -> OB (biases are in the accumulators) ----------
          mbar_wait(bar_acc(slot, 0), par_accA);
          FWD_MARK(1);   // wait P1
          par_accA ^= 1;
          tc_fence_after();
          {
#ifdef TRAJSDE_FWD_E1_PREFETCH
            uint32_t v[32], vg[32];
            tmem_ld_32x32b_x32(tm_lane, v);                  // both layer-1 accumulator blocks in flight at once: the second read's
            tmem_ld_32x32b_x32(tm_lane + 64, vg);            // latency hides behind the first block's tanh burst
            tc_wait_ld();
            tanh32_to_tmem(v, tm_oa);
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 1));                  // h1f ready -> P2f
            tanh32_to_tmem(vg, tm_ob);
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 0));                  // h1g ready -> P2g
#else
            uint32_t v[32];
            Bias32 bf;
            if (!BIAS_MMA) bf = ld_bias32(ent + hh * 32);
            tmem_ld_32x32b_x32(tm_lane, v);
            tc_wait_ld();
            if (BIAS_MMA) tanh32_to_tmem(v, tm_oa);
            else act32_to_tmem(v, bf, tm_oa);
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 1));                  // h1f ready -> P2f
            if (!BIAS_MMA) bf = ld_bias32(ent + 64 + hh * 32);
            tmem_ld_32x32b_x32(tm_lane + 64, v);
            tc_wait_ld();
            if (BIAS_MMA) tanh32_to_tmem(v, tm_ob);
            else act32_to_tmem(v, bf, tm_ob);
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 0));                  // h1g ready -> P2g
#endif
          }
          if (!HAS_DW) {
            mbar_wait(bar_xfree(slot), par_xfree);           // stores of the previous step have finished reading X / the states staging
            par_xfree ^= 1;
            draw(0, 3);
          }
          FWD_MARK(2);   // epilogue 1 (+ draw)

          // ---- epilogue 2: h2f = tanh(z2f) -> OA (P3 may start) ; partial of w3 . tanh(z2g) -----------------------------------------
          mbar_wait(bar_acc(slot, 1), par_accB);
          FWD_MARK(3);   // wait P2f
          par_accB ^= 1;
          tc_fence_after();
          {
            uint32_t v[32];
            Bias32 b2;
            if (!BIAS_MMA) b2 = ld_bias32(vec + VEC_B2 + hh * 32);
            tmem_ld_32x32b_x32(tm_lane, v);
            tc_wait_ld();
            if (BIAS_MMA) tanh32_to_tmem(v, tm_oa);
            else act32_to_tmem(v, b2, tm_oa);
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 1));                  // h2f ready -> P3
            if (!HAS_DW) draw(3, 5);
            FWD_MARK(4);   // epilogue 2a (+ draw)
            mbar_wait(bar_acc(slot, 0), par_accA);
            FWD_MARK(5);   // wait P2g
            par_accA ^= 1;
            tc_fence_after();
            tmem_ld_32x32b_x32(tm_lane + 64, v);
            tc_wait_ld();
            float gd = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 w = *reinterpret_cast<const float4*>(w3v + j);
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if (!BIAS_MMA) b = *reinterpret_cast<const float4*>(c2v + j);
              gd = fmaf(ts_tanh_approx(BIAS_MMA ? __uint_as_float(v[j]) : __uint_as_float(v[j]) + b.x), w.x, gd);
              gd = fmaf(ts_tanh_approx(BIAS_MMA ? __uint_as_float(v[j + 1]) : __uint_as_float(v[j + 1]) + b.y), w.y, gd);
              gd = fmaf(ts_tanh_approx(BIAS_MMA ? __uint_as_float(v[j + 2]) : __uint_as_float(v[j + 2]) + b.z), w.z, gd);
              gd = fmaf(ts_tanh_approx(BIAS_MMA ? __uint_as_float(v[j + 3]) : __uint_as_float(v[j + 3]) + b.w), w.w, gd);
            }
            gpart[hh * TILE_M + row] = gd;
          }
          FWD_MARK(6);   // epilogue 2b

          // ---- epilogue 3 ---------------------------------------------------------------------------------------------------------
          if (HAS_DW) {
            FWD_TL(1, mbar_wait(bar_xfree(slot), par_xfree));
            par_xfree ^= 1;
            FWD_TL(2, mbar_wait(bar_tma(slot), par_tma));      // dW tile of this step has landed in X
            par_tma ^= 1;
          } else {
            draw(5, 8);
          }
          uint32_t yv[32];
#ifdef TRAJSDE_FWD_Y_PREFETCH
          tmem_ld_32x32b_x32(c3.tm_y, yv);                     // the state read runs under the P3 wait
          constexpr bool YPRE = true;
#else
          constexpr bool YPRE = false;
#endif
          FWD_MARK(7);   // X free + dW landed (or draw)
          FWD_TL(3, mbar_wait(bar_acc(slot, 1), par_accB));
          FWD_MARK(8);   // wait P3
#ifdef TRAJSDE_FWD_TIMELINE
          if (threadIdx.x == 0 && blockIdx.x == 0) g_fwd_tl[4] += 1;
#endif
          par_accB ^= 1;
          tc_fence_after();
          named_bar_sync(pair_bar, 64);                        // partner warp's partial g dot is in smem
          const float g = __fdividef(1.0f, 1.0f + __expf(-((gpart[row] + gpart[TILE_M + row]) + c3b)));
          c3.h = sc.x;
          c3.g = g;
          c3.w0 = sc.z;
          c3.w1 = sc.w;
          c3.ob = so.x;
          c3.nout = so.y;
          c3.has_out = so.y > 0;
          if (so.y > 1) epi3_update<true, true, YPRE, BIAS_MMA>(c3, yv);
          else epi3_update<false, true, YPRE, BIAS_MMA>(c3, yv);
          if (k == S - 1 && hh == 0 && a.g_last && valid) a.g_last[grow] = g;
          if (BIAS_MMA && k + 1 < S) {                         // bias operand slice of the next step (other parity: nobody reads it now)
            mbar_wait(bar_ring(slot, (gstep + 1) & 1), ((gstep + 1) >> 1) & 1);
            if (hh == 0)
              *reinterpret_cast<uint4*>(a0_row + (((2u * ((gstep + 1) & 1u)) ^ (row & 7u)) << 4)) =
                  *reinterpret_cast<const uint4*>(ring + ((gstep + 1) % 3u) * RING_LD);
          }
          FWD_MARK(9);   // epilogue 3
          fence_proxy_async();
          tc_fence_before();
          if (k + 1 < S) mbar_arrive(bar_opnd(slot, 0));     // y' (tensor memory) + bias slice ready -> P1 of step k+1
          mbar_arrive(bar_xfull(slot));                      // X (outputs) / states staging written -> IO warp stores them
          FWD_MARK(10);  // fences + arrives
        } else {
          // ===================== dual-diffusion variants: operands through swizzled shared memory =====================================
          if (save_states && !TMEM_A) {                      // previous step's states store must have drained A1f|A1g before epilogue 1
            mbar_wait(bar_xfree(slot), par_xfree);           // writes h1f there (the TMEM-operand variants write that staging area only
            par_xfree ^= 1;                                  // in epilogue 3 and wait there)
          }
          mbar_wait(bar_ring(slot, gstep & 1), (gstep >> 1) & 1);   // this step's bias row + scalars are in the ring
          const float* ent = ring + (gstep % 3u) * RING_LD;
          const float4 sc = *reinterpret_cast<const float4*>(ent + BIAS1_LD);      // h, sqrt(h), w0, w1
          const int2 so = *reinterpret_cast<const int2*>(ent + BIAS1_LD + 4);      // first output index, #outputs of this step

          // ---- epilogue 1: h1f = tanh(z1f + b1f(t)) -> A1f (P2f may start), h1g = tanh(z1g + c1(t)) -> A1g -------------------
          mbar_wait(bar_acc(slot, 0), par_accA);
          par_accA ^= 1;
          tc_fence_after();
          {
            uint32_t v[32];
            Bias32 bf = ld_bias32(ent + hh * 32);
            tmem_ld_32x32b_x32(tm_lane, v);
            tc_wait_ld();
            if (TMEM_A) {
              act32_to_tmem(v, bf, tm_oa);
            } else {
              act32_to_operand(v, bf, a1f_row, row, hh * 4);
              fence_proxy_async();
            }
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 1));                  // A1f ready -> P2f
            bf = ld_bias32(ent + gcol + hh * 32);
            ld_g<DUAL>(tm_lane + ucol, tm_lane + 128, w_mixed, use_alt, v);
            if (TMEM_A) {
              act32_to_tmem(v, bf, tm_ob);
            } else {
              act32_to_operand(v, bf, a1g_row, row, hh * 4);
              fence_proxy_async();
            }
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 0));                  // A1g ready -> P2g
          }

          // ---- epilogue 2: h2f = tanh(z2f + b2) -> A0 (P3 may start) ; partial of w3 . tanh(z2g + c2) ----------------------------
          mbar_wait(bar_acc(slot, 1), par_accB);
          par_accB ^= 1;
          tc_fence_after();
          {
            uint32_t v[32];
            const Bias32 b2 = ld_bias32(vec + VEC_B2 + hh * 32);
            tmem_ld_32x32b_x32(tm_lane, v);
            tc_wait_ld();
            if (TMEM_A) {
              act32_to_tmem(v, b2, tm_oa);
            } else {
              act32_to_operand(v, b2, a0_row, row, hh * 4);
              fence_proxy_async();
            }
            tc_fence_before();
            mbar_arrive(bar_opnd(slot, 1));                  // h2f ready -> P3
            mbar_wait(bar_acc(slot, 0), par_accA);
            par_accA ^= 1;
            tc_fence_after();
            ld_g<DUAL>(tm_lane + ucol, tm_lane + 128, w_mixed, use_alt, v);
            float gd = 0.f;
  #pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(c2v + j);
              const float4 w = *reinterpret_cast<const float4*>(w3v + j);
              gd = fmaf(ts_tanh_approx(__uint_as_float(v[j]) + b.x), w.x, gd);
              gd = fmaf(ts_tanh_approx(__uint_as_float(v[j + 1]) + b.y), w.y, gd);
              gd = fmaf(ts_tanh_approx(__uint_as_float(v[j + 2]) + b.z), w.z, gd);
              gd = fmaf(ts_tanh_approx(__uint_as_float(v[j + 3]) + b.w), w.w, gd);
            }
            gpart[hh * TILE_M + row] = gd;
          }

          // ---- epilogue 3 -----------------------------------------------------------------------------------------------------
          if (!save_states || TMEM_A) {                        // stores of the previous step have finished reading X (and the states staging)
            mbar_wait(bar_xfree(slot), par_xfree);
            par_xfree ^= 1;
          }
          if (HAS_DW) {
            mbar_wait(bar_tma(slot), par_tma);                 // dW tile of this step has landed in X
            par_tma ^= 1;
          } else {                                             // draw this thread's 32 increments into X while P3 runs
  #pragma unroll 2
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(x_row + ((q ^ (row & 7u)) << 4)) =
                  philox_dw4(noise_seed, (uint64_t)grow + a.noise.row_offset, a.noise.step_offset + (uint32_t)k, (uint32_t)(hh * 8 + q), sc.y);
          }
          mbar_wait(bar_acc(slot, 1), par_accB);
          par_accB ^= 1;
          tc_fence_after();
          named_bar_sync(pair_bar, 64);                        // partner warp's partial g dot is in smem
          const float g = __fdividef(1.0f, 1.0f + __expf(-((gpart[row] + gpart[TILE_M + row]) + c3b)));
          c3.h = sc.x;
          c3.g = g;
          c3.w0 = sc.z;
          c3.w1 = sc.w;
          c3.ob = so.x;
          c3.nout = so.y;
          c3.has_out = so.y > 0;
          {
            uint32_t yv[32];
            if (so.y > 1) epi3_update<true, TMEM_A>(c3, yv);
            else epi3_update<false, TMEM_A>(c3, yv);
          }
          if (k == S - 1 && hh == 0 && a.g_last && valid) a.g_last[grow] = g;
          fence_proxy_async();
          tc_fence_before();
          if (k + 1 < S) mbar_arrive(bar_opnd(slot, 0));     // A0 = y' ready -> P1 of step k+1
          mbar_arrive(bar_xfull(slot));                      // X (outputs) / states staging written -> IO warp stores them
        }
      }
    }
  } else if (warp < WARP_IO0) {
    // =============================================== MMA ISSUER WARPS ===============================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
    const int slot = warp - WARP_MMA0;
    // The whole warp runs this loop (warp-uniform control flow and operands -> descriptors live in uniform registers); one
    // elected lane issues the tcgen05.mma / tcgen05.commit instructions.
    {
      const uint32_t slot_u32 = base + SMEM_SLOTS + slot * SLOT_BYTES;
      const uint32_t d_base = tmem_base + slot * 256;
      const uint32_t idesc_p1 = umma_idesc_f16(TILE_M, DUAL ? 192u : 128u);
      const uint32_t idesc_64 = umma_idesc_f16(TILE_M, 64);
      // all operand descriptors share the high word (SBO / version / swizzle); only the start-address field differs
      const uint64_t dhi = umma_desc_sw128(0);
      auto D = [&](uint32_t addr) { return dhi | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      const uint32_t aA0 = slot_u32 + OFF_A0, aA1f = slot_u32 + OFF_A1F, aA1g = slot_u32 + OFF_A1G;
      constexpr bool TMEM_A = !DUAL;
      constexpr bool BIAS_MMA = TMEM_A && !HAS_DW && (TRAJSDE_FWD_BIAS_MMA != 0);
      const uint32_t t_oa = d_base + 128, t_ob = d_base + 160;   // TMEM operand columns (lane field 0: all 128 rows)
      const uint32_t aB1 = base + (TMEM_A ? SIMG_B1 : IMG_B1), aW2 = base + (TMEM_A ? SIMG_W2 : IMG_W2), aV2 = base + (TMEM_A ? SIMG_V2 : IMG_V2),
                     aV2a = base + IMG_V2A, aW3 = base + (TMEM_A ? SIMG_W3 : IMG_W3);
      // biases through the MMA (TMEM_A variants): fifth K=16 block of every phase = bias operand slice (A0 tile, slice gstep & 1)
      // x BIAS tile slice of the layer (slice 0: layer 1 incl. time columns; slice 1: b2 | c2 at rows 0 / 64; slice 2: b3)
      const uint32_t aBias = base + SIMG_BIAS;
      uint32_t par_op0 = 0, par_op1 = 0;
      uint32_t gstep = 0;
      mbar_wait(bar_w, 0);
      for (int tile = tile_lo + slot; tile < tile_hi; tile += NUM_SLOTS) {
        for (int k = 0; k < S; ++k, ++gstep) {
          const uint64_t dAb = D(aA0 + 32u * (gstep & 1u));
          // P1: [z1f | z1g (| z1g_alt)] = y . [W1y ; V1y (; V1y_alt)]^T
          mbar_wait(bar_opnd(slot, 0), par_op0);
          par_op0 ^= 1;
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (TMEM_A) tc_mma_f16_ts(d_base, t_oa + 8 * kk, D(aB1 + 32 * kk), idesc_p1, kk > 0);
              else tc_mma_f16(d_base, D(aA0 + 32 * kk), D(aB1 + 32 * kk), idesc_p1, kk > 0);
            }
            if (BIAS_MMA) tc_mma_f16(d_base, dAb, D(aBias), idesc_p1, 1);
            tc_commit(bar_acc(slot, 0));
          }
          __syncwarp();
          // P2f: z2f = h1f . W2^T  (overwrites the z1f columns the epilogue has already consumed)
          mbar_wait(bar_opnd(slot, 1), par_op1);
          par_op1 ^= 1;
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (TMEM_A) tc_mma_f16_ts(d_base, t_oa + 8 * kk, D(aW2 + 32 * kk), idesc_64, kk > 0);
              else tc_mma_f16(d_base, D(aA1f + 32 * kk), D(aW2 + 32 * kk), idesc_64, kk > 0);
            }
            if (BIAS_MMA) tc_mma_f16(d_base, dAb, D(aBias + 32), idesc_64, 1);
            tc_commit(bar_acc(slot, 1));
          }
          __syncwarp();
          // P2g: z2g = h1g . V2^T (, z2g_alt = h1g . V2alt^T)
          mbar_wait(bar_opnd(slot, 0), par_op0);
          par_op0 ^= 1;
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (TMEM_A) tc_mma_f16_ts(d_base + 64, t_ob + 8 * kk, D(aV2 + 32 * kk), idesc_64, kk > 0);
              else tc_mma_f16(d_base + 64, D(aA1g + 32 * kk), D(aV2 + 32 * kk), idesc_64, kk > 0);
            }
            if (BIAS_MMA) tc_mma_f16(d_base + 64, dAb, D(aBias + 64 * 128 + 32), idesc_64, 1);
            if (DUAL) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) tc_mma_f16(d_base + 128, D(aA1g + 32 * kk), D(aV2a + 32 * kk), idesc_64, kk > 0);
            }
            tc_commit(bar_acc(slot, 0));
          }
          __syncwarp();
          // P3: drift output f = h2f . W3^T (reuses the z2f columns, already consumed by epilogue 2)
          mbar_wait(bar_opnd(slot, 1), par_op1);
          par_op1 ^= 1;
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              if (TMEM_A) tc_mma_f16_ts(d_base, t_oa + 8 * kk, D(aW3 + 32 * kk), idesc_64, kk > 0);
              else tc_mma_f16(d_base, D(aA0 + 32 * kk), D(aW3 + 32 * kk), idesc_64, kk > 0);
            }
            if (BIAS_MMA) tc_mma_f16(d_base, dAb, D(aBias + 64), idesc_64, 1);
            tc_commit(bar_acc(slot, 1));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // =============================================== IO WARPS ===============================================================
    // TMA in/out for the slot + staging of each step's layer-1 bias row and scalars into a 3-deep smem ring.
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
    const int slot = warp - WARP_IO0;
    const uint32_t slot_u32 = base + SMEM_SLOTS + slot * SLOT_BYTES;
    float* ring = reinterpret_cast<float*>(sm + SMEM_RING + slot * RING_BYTES);
    uint32_t par_xfull = 0;
    uint32_t gstep = 0;
    const bool sep_dw = HAS_DW && !DUAL && !save_states && (TRAJSDE_FWD_SEPARATE_DW != 0);
    const uint32_t dw_dst = slot_u32 + (sep_dw ? OFF_A1F : OFF_X);
    int my_tiles = 0;
    for (int tile = tile_lo + slot; tile < tile_hi; tile += NUM_SLOTS) ++my_tiles;
    const uint32_t total_steps = (uint32_t)my_tiles * (uint32_t)S;
    // Ring entry of global step g -> ring[g % 3], announced on bar_ring[g & 1].  Entry g+1 is published at the start of IO
    // iteration g, i.e. once the epilogue has finished step g-1: its slot (last read in step g-2) is free and its barrier's
    // previous phase (entry g-1) has been consumed, so every barrier has at most one outstanding phase.
    struct Ent { float2 b[3]; float4 sc; int2 so; uint4 aw; };
    constexpr bool BIAS_MMA = !DUAL && !HAS_DW && (TRAJSDE_FWD_BIAS_MMA != 0);
    auto ent_fetch = [&](uint32_t g, Ent& e) {
      const int k = (int)(g % (uint32_t)S);
      if (!BIAS_MMA) {
        const float2* src = reinterpret_cast<const float2*>(p.bias1 + (size_t)k * BIAS1_LD);
#pragma unroll
        for (int i = 0; i < 3; ++i) e.b[i] = __ldg(src + lane + 32 * i);
      }
      if (lane == 0) {
        const float4 st = __ldg(reinterpret_cast<const float4*>(a.sched.step_tab) + k);
        if (BIAS_MMA) { // bias operand row of the step: (1, 1, s_hi, s_lo, s_hi, c_hi, c_lo, c_hi) as fp16 (head + remainder of sin t0, cos t0)
          const float sh = __half2float(__float2half_rn(st.z)), ch = __half2float(__float2half_rn(st.w));
          e.aw = make_uint4(pack_f16x2(1.f, 1.f), pack_f16x2(sh, st.z - sh), pack_f16x2(sh, ch), pack_f16x2(st.w - ch, ch));
        }
        const int ob = a.sched.out_begin[k], oe = a.sched.out_begin[k + 1];
        float w0 = 0.f, w1 = 1.f;
        if (oe > ob) {
          w0 = a.sched.out_w[2 * ob];
          w1 = a.sched.out_w[2 * ob + 1];
        }
        e.sc = make_float4(st.y, sqrtf(st.y), w0, w1);
        e.so = make_int2(ob, oe - ob);
      }
    };
    auto ent_publish = [&](uint32_t g, const Ent& e) {
      float* dst = ring + (g % 3u) * RING_LD;
      if (!BIAS_MMA) {
#pragma unroll
        for (int i = 0; i < 3; ++i) reinterpret_cast<float2*>(dst)[lane + 32 * i] = e.b[i];
      }
      if (lane == 0) {
        if (BIAS_MMA) *reinterpret_cast<uint4*>(dst) = e.aw;
        *reinterpret_cast<float4*>(dst + BIAS1_LD) = e.sc;
        *reinterpret_cast<int2*>(dst + BIAS1_LD + 4) = e.so;
      }
      // every lane release-arrives for its own stores; the __syncwarp paces the warp: lane 0 blocks on bar_xfull below, and without
      // the rendezvous lanes 1..31 would run ahead and publish entries whose ring slots are still in use
      __syncwarp();
      mbar_arrive(bar_ring(slot, g & 1));
    };
    if (total_steps > 0) {
      Ent e;
      ent_fetch(0, e);
      ent_publish(0, e);
    }
    for (int tile = tile_lo + slot; tile < tile_hi; tile += NUM_SLOTS) {
      const int row0 = tile * TILE_M;
      if (lane == 0) {
        mbar_arrive_expect_tx(bar_tma(slot), 32768);
        tma_load_3d(slot_u32 + OFF_X, &tm_y0, bar_tma(slot), 0, row0, 0);
        tma_load_3d(slot_u32 + OFF_X + 16384, &tm_y0, bar_tma(slot), 32, row0, 0);
        mbar_wait(bar_xfull(slot), par_xfull);             // every epilogue thread holds its y0
        tma_store_3d(&tm_ys, slot_u32 + OFF_X, 0, row0, 0);              // ys[0] = y0
        tma_store_3d(&tm_ys, slot_u32 + OFF_X + 16384, 32, row0, 0);
        tma_store_commit();
        tma_store_wait_read0();
        mbar_arrive(bar_xfree(slot));
        if (HAS_DW) {
          mbar_arrive_expect_tx(bar_tma(slot), 32768);
          tma_load_3d(dw_dst, &tm_dw, bar_tma(slot), 0, row0, 0);
          tma_load_3d(dw_dst + 16384, &tm_dw, bar_tma(slot), 32, row0, 0);
        }
      }
      par_xfull ^= 1;
      for (int k = 0; k < S; ++k, ++gstep) {
        if (gstep + 1 < total_steps) {
          Ent e;
          ent_fetch(gstep + 1, e);
          ent_publish(gstep + 1, e);
        }
        if (lane == 0) {
          const int ob = a.sched.out_begin[k], oe = a.sched.out_begin[k + 1];
          mbar_wait(bar_xfull(slot), par_xfull);           // epilogue 3 of step k finished writing X / states staging
          if (HAS_DW && sep_dw && k + 1 < S) {               // ... and reading its dW tile: the next one goes out ahead of the stores
            mbar_arrive_expect_tx(bar_tma(slot), 32768);
            tma_load_3d(dw_dst, &tm_dw, bar_tma(slot), 0, row0, k + 1);
            tma_load_3d(dw_dst + 16384, &tm_dw, bar_tma(slot), 32, row0, k + 1);
          }
          if (oe > ob) {
            tma_store_3d(&tm_ys, slot_u32 + OFF_X, 0, row0, ob + 1);
            tma_store_3d(&tm_ys, slot_u32 + OFF_X + 16384, 32, row0, ob + 1);
          }
          if (save_states) {
            tma_store_3d(&tm_st, slot_u32 + OFF_A1F, 0, row0, k);
            tma_store_3d(&tm_st, slot_u32 + OFF_A1F + 16384, 32, row0, k);
          }
          tma_store_commit();
          tma_store_wait_read0();
          if (k + 1 < S) {
            mbar_arrive(bar_xfree(slot));
            if (HAS_DW && !sep_dw) {
              mbar_arrive_expect_tx(bar_tma(slot), 32768);
              tma_load_3d(slot_u32 + OFF_X, &tm_dw, bar_tma(slot), 0, row0, k + 1);
              tma_load_3d(slot_u32 + OFF_X + 16384, &tm_dw, bar_tma(slot), 32, row0, k + 1);
            }
          }
        }
        par_xfull ^= 1;
      }
    }
    if (lane == 0) tma_store_wait_all0();
  }

  tc_fence_before();
  __syncthreads();
#ifdef TRAJSDE_FWD_TIMELINE
  if (threadIdx.x == 0 && blockIdx.x == 0) g_fwd_tl[0] += clock64() - tl_start;
#endif
  if (warp == WARP_MMA0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- host side -----------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 3-D fp32 tensor {64 channels, rows, slabs}; box {32, 128, 1}; 128B swizzle; OOB rows are zero-filled / clipped.
int make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t slabs, int64_t row_stride, int64_t slab_stride) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TRAJSDE_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(slabs > 0 ? slabs : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)(slab_stride > 0 ? slab_stride : row_stride * rows) * 4};
  cuuint32_t box[3] = {32, TILE_M, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TRAJSDE_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return TRAJSDE_OK;
}

}  // namespace

int tc_make_map(CUtensorMap* m, const float* ptr, int64_t rows, int64_t slabs, int64_t row_stride, int64_t slab_stride) {
  return make_map(m, ptr, rows, slabs, row_stride, slab_stride);
}

int64_t euler_fwd_tc_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual) {
  (void)rows;
  (void)dual;
  return (int64_t)IMG_BYTES + (int64_t)n_steps * BIAS1_LD * 4 + 256;
}

int launch_euler_fwd_tc(const TrajsdeEulerFwdArgs& a, cudaStream_t s) {
  int dev = 0, sms = 0;
  TS_CUDA_CHECK(cudaGetDevice(&dev));
  TS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0)
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "workspace must be 256-byte aligned");
  if (a.rows >= (int64_t)1 << 31) return set_error(TRAJSDE_ERR_UNSUPPORTED, "rows >= 2^31 unsupported in TC mode");
  TcParams p;
  p.a = a;
  p.img = static_cast<const uint8_t*>(a.workspace);
  p.bias1 = reinterpret_cast<const float*>(static_cast<const uint8_t*>(a.workspace) + IMG_BYTES);
  p.num_tiles = (int)((a.rows + TILE_M - 1) / TILE_M);
  p.dual = a.alt_mask != nullptr;
  if (p.num_tiles == 0) return TRAJSDE_OK;

  tc_pack_kernel<<<24, 256, 0, s>>>(a, static_cast<uint8_t*>(a.workspace),
                                    reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + IMG_BYTES), p.dual);
  TS_CUDA_CHECK(cudaGetLastError());

  CUtensorMap tm_y0, tm_dw, tm_ys, tm_st;
  const int T = a.sched.n_outputs + 1;
  int rc;
  if ((rc = make_map(&tm_y0, a.y0, a.rows, 1, a.y0_row_stride, 0)) != 0) return rc;
  if ((rc = make_map(&tm_ys, a.ys, a.rows, T, a.ys_row_stride, a.ys_t_stride)) != 0) return rc;
  if (a.noise.dw) {
    if ((rc = make_map(&tm_dw, a.noise.dw, a.rows, a.sched.n_steps, 64, 0)) != 0) return rc;
  } else {
    tm_dw = tm_y0;
  }
  if (a.states) {
    if ((rc = make_map(&tm_st, a.states, a.rows, a.sched.n_steps, 64, 0)) != 0) return rc;
  } else {
    tm_st = tm_y0;
  }
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC);
    if (e != cudaSuccess) return e;
    kern<<<grid, NUM_THREADS, SMEM_ALLOC, s>>>(p, tm_y0, tm_dw, tm_ys, tm_st);
    return cudaGetLastError();
  };
  const bool hd = a.noise.dw != nullptr, du = p.dual != 0;
  TS_CUDA_CHECK(hd ? (du ? launch(euler_fwd_tc_kernel<true, true>) : launch(euler_fwd_tc_kernel<true, false>))
                   : (du ? launch(euler_fwd_tc_kernel<false, true>) : launch(euler_fwd_tc_kernel<false, false>)));
  return TRAJSDE_OK;
}

}  // namespace trajsde

#ifdef TRAJSDE_FWD_TIMELINE
extern "C" int trajsde_debug_fwd_timeline(long long* out8) {
  long long zero[8] = {0};
  if (cudaMemcpyFromSymbol(out8, trajsde::g_fwd_tl, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(trajsde::g_fwd_tl, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
extern "C" int trajsde_debug_fwd_segments(long long* out16) {
  long long zero[16] = {0};
  if (cudaMemcpyFromSymbol(out16, trajsde::g_fwd_seg, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(trajsde::g_fwd_seg, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif
