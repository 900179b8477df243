// Fused encoder recurrence (TRAJSDE_MODE_TC_F16): n_steps x [one Euler–Maruyama step of the dual-diffusion latent SDE +
// GRU_Unit observation jump] in ONE persistent sm_100a kernel — the loop body of LocalEncoderSDESepPara2.forward
// (models/encoders/enc_hivt_nusargo_sde_sep2.py:128-182) that the reference runs as 21 x [sdeint_dual (models/utils/sdeint.py:
// 110-197) + GRU_Unit.forward (models/utils/ode_utils.py:136-152)] ~ 191 aten launches per iteration with two host syncs.
//
// Machine mapping: one 128-row tile per CTA at a time; 16 epilogue warps (thread = one row x one 16-channel quarter) + one
// MMA-issuer warp.  The fp32 latent state of the tile lives in TMEM for all iterations; all weights (SDE 56 KB + GRU 72 KB,
// fp16 128B-swizzled UMMA tiles) plus the per-iteration layer-1 bias rows are staged once per CTA with bulk TMA copies.
// Per iteration eight dependent tcgen05 phases, handed over through alternating mbarrier pairs exactly like euler_tc.cu:
//   P1  [z1f|z1g|z1g_alt] = y . [W1y;V1y;V1y_alt]^T      P2f z2f = h1f . W2^T      P2g z2g = h1g . V2^T (+alt)    P3 f = h2f . W3^T
//   G1  [zu|zr] = [y1|x] . [U1;R1]^T (K=128)             G2  u' = tu . U2^T, r' = tr . R2^T
//   G3  zn = [x|r*y1] . N1^T (K=128)                     G4  n = tn . N2^T
// with  y1 = y + f h + g dW,  u = sigmoid(u'), r = sigmoid(r'),  h' = (1-u) n + u y1,  state <- mask ? h' : y1.
// HBM traffic per row-iteration: x in (256 B), latent out (256 B), g out (4 B), mask (1 B) [+ dW in (256 B) when supplied].
#include "common.cuh"
#include "tc_common.cuh"

namespace trajsde {

using namespace tc;

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_EPI_WARPS = 8;           // thread = one row x one 32-channel half
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = NUM_EPI_THREADS + 128;   // + one warpgroup: MMA-issuer warp and three idle warps (setmaxnreg is per warpgroup)
constexpr int EPI_REGS = 216, AUX_REGS = 64;
constexpr int S_MAX = 32;

// ---- packed image (bytes) -------------------------------------------------------------------------------------------------
constexpr uint32_t IMG_B1 = 0, IMG_W2 = 24576, IMG_V2 = 32768, IMG_V2A = 40960, IMG_W3 = 49152;
constexpr uint32_t IMG_UR1H = 57344, IMG_UR1X = 73728, IMG_U2 = 90112, IMG_R2 = 98304;
constexpr uint32_t IMG_N1X = 106496, IMG_N1RH = 114688, IMG_N2 = 122880;
constexpr uint32_t IMG_VEC = 131072;     // 1024 fp32
constexpr uint32_t IMG_BIAS1 = 135168;   // [S_MAX][192] fp32
constexpr uint32_t IMG_BYTES = IMG_BIAS1 + S_MAX * 192 * 4;   // 159744
constexpr int VEC_B2 = 0, VEC_C2 = 64, VEC_C2A = 128, VEC_B3 = 192, VEC_W3G = 256, VEC_W3GA = 320, VEC_C3 = 384, VEC_C3A = 385;
constexpr int VEC_UB1 = 512, VEC_RB1 = 576, VEC_UB2 = 640, VEC_RB2 = 704, VEC_NB1 = 768, VEC_NB2 = 832;

// ---- shared memory map: weights + bias rows only — every MMA A operand lives in tensor memory ---------------------------------
constexpr uint32_t OFF_GPART = IMG_BYTES;                    // [2 halves][128 rows] fp32 partial diffusion dots
constexpr uint32_t OFF_BARS = OFF_GPART + 2 * TILE_M * 4;
constexpr uint32_t SMEM_TOTAL = OFF_BARS + 128;
constexpr uint32_t SMEM_ALLOC = SMEM_TOTAL + 1024;
static_assert(SMEM_ALLOC <= 232448, "exceeds 227 KB of shared memory per CTA");

// ---- TMEM columns ----------------------------------------------------------------------------------------------------------------
// [0,192)   accumulators: SDE phases (P1 192 wide in the dual variant), then reused by the GRU phases (G1 [zu|zr] 128, G2 u'|r', G3 zn, G4 n)
// [192,256) fp32 latent state Y
// [256,384) A operands, fp16 pairs (32 columns each): AH (y / h2f / y1 / r*y1 / h'), AX (x), A1F (h1f / tu / tn), A1G (h1g / tr)
constexpr uint32_t TM_Y = 192, TM_AH = 256, TM_AX = 288, TM_A1F = 320, TM_A1G = 352;

struct EncParams {
  TrajsdeEncFwdArgs a;
  const uint8_t* img;
  int num_tiles;
  int dual;
};

__device__ __forceinline__ void put_h(uint8_t* img, uint32_t off, int n, int k, float v) {
  *reinterpret_cast<__half*>(img + off + sw128_off_h(n, k)) = __float2half_rn(v);
}

__global__ void enc_pack_kernel(TrajsdeEncFwdArgs a, uint8_t* __restrict__ img, int dual) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int idx = tid; idx < 192 * 64; idx += nth) {
    const int n = idx >> 6, k = idx & 63, net = n >> 6, r = n & 63;
    const float* w = net == 0 ? a.drift.w1 : net == 1 ? a.diffusion.w1 : (dual ? a.diffusion_alt.w1 : nullptr);
    put_h(img, IMG_B1, n, k, w ? w[r * TS_IN1 + k] : 0.f);
  }
  for (int idx = tid; idx < 64 * 64; idx += nth) {
    const int n = idx >> 6, k = idx & 63;
    put_h(img, IMG_W2, n, k, a.drift.w2[n * 64 + k]);
    put_h(img, IMG_V2, n, k, a.diffusion.w2[n * 64 + k]);
    put_h(img, IMG_V2A, n, k, dual ? a.diffusion_alt.w2[n * 64 + k] : 0.f);
    put_h(img, IMG_W3, n, k, a.drift.w3[n * 64 + k]);
    put_h(img, IMG_U2, n, k, a.gru.u2[n * 64 + k]);
    put_h(img, IMG_R2, n, k, a.gru.r2[n * 64 + k]);
    put_h(img, IMG_N2, n, k, a.gru.n2[n * 64 + k]);
    put_h(img, IMG_N1X, n, k, a.gru.n1[n * 128 + k]);
    put_h(img, IMG_N1RH, n, k, a.gru.n1[n * 128 + 64 + k]);
    // [U1;R1]: rows 0..63 update gate, 64..127 reset gate; K-half 0 multiplies h_cur, K-half 1 the input
    put_h(img, IMG_UR1H, n, k, a.gru.u1[n * 128 + k]);
    put_h(img, IMG_UR1H, n + 64, k, a.gru.r1[n * 128 + k]);
    put_h(img, IMG_UR1X, n, k, a.gru.u1[n * 128 + 64 + k]);
    put_h(img, IMG_UR1X, n + 64, k, a.gru.r1[n * 128 + 64 + k]);
  }
  float* vec = reinterpret_cast<float*>(img + IMG_VEC);
  for (int i = tid; i < 1024; i += nth) {
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
    else if (i >= VEC_UB1 && i < VEC_UB1 + 64) v = a.gru.ub1[c];
    else if (i >= VEC_RB1 && i < VEC_RB1 + 64) v = a.gru.rb1[c];
    else if (i >= VEC_UB2 && i < VEC_UB2 + 64) v = a.gru.ub2[c];
    else if (i >= VEC_RB2 && i < VEC_RB2 + 64) v = a.gru.rb2[c];
    else if (i >= VEC_NB1 && i < VEC_NB1 + 64) v = a.gru.nb1[c];
    else if (i >= VEC_NB2 && i < VEC_NB2 + 64) v = a.gru.nb2[c];
    vec[i] = v;
  }
  float* bias1 = reinterpret_cast<float*>(img + IMG_BIAS1);
  const int S = a.sched.n_steps;
  for (int idx = tid; idx < S_MAX * 192; idx += nth) {
    const int k = idx / 192, n = idx % 192, net = n >> 6, r = n & 63;
    const TrajsdeMlp* m = net == 0 ? &a.drift : net == 1 ? &a.diffusion : (dual ? &a.diffusion_alt : nullptr);
    float v = 0.f;
    if (m && k < S) {
      const float sn = a.sched.step_tab[4 * k + 2], cs = a.sched.step_tab[4 * k + 3];
      v = fmaf(m->w1[r * TS_IN1 + 65], cs, fmaf(m->w1[r * TS_IN1 + 64], sn, m->b1[r]));
    }
    bias1[idx] = v;
  }
}

// fp16 pairs of this thread's 32 values -> its 16 packed operand columns
__device__ __forceinline__ void st_operand32(uint32_t taddr, const float (&t)[32]) {
  uint32_t p[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) p[e] = pack_f16x2(t[2 * e], t[2 * e + 1]);
  tmem_st_32x32b_x16(taddr, p);
}
template <bool SIGMOID>
__device__ __forceinline__ float act1(float x) {
  return SIGMOID ? fmaf(0.5f, ts_tanh_approx(0.5f * x), 0.5f) : ts_tanh_approx(x);
}
// tanh(acc + bias) for 32 accumulator columns -> operand columns
__device__ __forceinline__ void tanh32_to_operand(const uint32_t (&v)[32], const float* __restrict__ bias, uint32_t taddr) {
  float t[32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * q);
    t[4 * q] = ts_tanh_approx(__uint_as_float(v[4 * q]) + b.x);
    t[4 * q + 1] = ts_tanh_approx(__uint_as_float(v[4 * q + 1]) + b.y);
    t[4 * q + 2] = ts_tanh_approx(__uint_as_float(v[4 * q + 2]) + b.z);
    t[4 * q + 3] = ts_tanh_approx(__uint_as_float(v[4 * q + 3]) + b.w);
  }
  st_operand32(taddr, t);
}

template <bool DUAL>
__device__ __forceinline__ void ld_g32(uint32_t tm_uniform, uint32_t tm_alt, bool w_mixed, bool use_alt, uint32_t (&v)[32]) {
  tmem_ld_32x32b_x32(tm_uniform, v);
  if (DUAL && w_mixed) {
    uint32_t v2[32];
    tmem_ld_32x32b_x32(tm_alt, v2);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = use_alt ? v2[j] : v[j];
  } else {
    tc_wait_ld();
  }
}

#ifdef TRAJSDE_ENC_TIMELINE
// debug build only (bench_micro/enc_fwd_timeline.py): thread 0 of CTA 0: [0] kernel clocks, [1] clocks waiting for the tensor core (all eight
// accumulator barriers of an iteration), [2] iterations
__device__ long long g_enc_tl[4];
__device__ long long g_enc_seg[20];   // [2k] work before hand-shake k, [2k+1] the wait itself (k = P1, P2f, P2g, P3, G1, G2, G3, G4), [16] iteration tail
#define ENC_TL_WAIT(k, expr) do { const long long _t0 = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) g_enc_seg[2 * (k)] += _t0 - seg_prev; expr; \
    seg_prev = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) { g_enc_tl[1] += seg_prev - _t0; g_enc_seg[2 * (k) + 1] += seg_prev - _t0; } } while (0)
#else
#define ENC_TL_WAIT(k, expr) do { expr; } while (0)
#endif

template <bool HAS_DW, bool DUAL>
__global__ void __launch_bounds__(NUM_THREADS, 1) enc_fwd_tc_kernel(const EncParams p) {
#ifdef TRAJSDE_ENC_TIMELINE
  const long long tl_start = clock64();
  long long seg_prev = tl_start;
#endif
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const TrajsdeEncFwdArgs& a = p.a;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int S = a.sched.n_steps;
  const uint64_t noise_seed = HAS_DW ? 0ull : ts_noise_seed(a.noise);

  const uint32_t bar_w = base + OFF_BARS;
  auto bar_opnd = [&](int i) { return base + OFF_BARS + 8u + 8u * i; };
  auto bar_acc = [&](int i) { return base + OFF_BARS + 24u + 8u * i; };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + OFF_BARS + 64);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_opnd(i), NUM_EPI_THREADS);
      mbar_init(bar_acc(i), 1);
    }
    mbar_fence_init();
  }
  if (warp == NUM_EPI_WARPS) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (threadIdx.x == 0) {  // stage weights + bias rows once per CTA (bulk TMA copies)
    mbar_arrive_expect_tx(bar_w, IMG_BYTES);
    constexpr uint32_t HALF = 79872;                      // two copies: keeps each bulk copy well below 128 KB
    bulk_load_1d(base, p.img, HALF, bar_w);
    bulk_load_1d(base + HALF, p.img + HALF, IMG_BYTES - HALF, bar_w);
  }

  const float* vec = reinterpret_cast<const float*>(sm + IMG_VEC);
  const float* bias1_tab = reinterpret_cast<const float*>(sm + IMG_BIAS1);

  if (warp < NUM_EPI_WARPS) {
    // =============================================== EPILOGUE WARPS ===============================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    const int quad = warp & 3;                             // TMEM lane quadrant (= warp index % 4)
    const uint32_t hh = (uint32_t)warp >> 2;               // 32-channel half of the row owned by this thread
    const uint32_t row = quad * 32 + lane;
    float* gpart = reinterpret_cast<float*>(sm + OFF_GPART);
    const uint32_t tml = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t tm = tml + hh * 32;                     // accumulator / state columns of this thread
    const uint32_t o_ah = tml + TM_AH + hh * 16, o_ax = tml + TM_AX + hh * 16, o_a1f = tml + TM_A1F + hh * 16, o_a1g = tml + TM_A1G + hh * 16;
    const uint32_t pair_bar = 1 + quad;                    // named barrier of the two warps that share these rows
    uint32_t par_accA = 0, par_accB = 0;
    // per-iteration scalars live in lane registers (S <= 32): lane i holds the slot index and the step size of iteration i
    const int lane_slot = lane < S ? a.slot[lane] : 0;
    const float lane_h = lane < S ? a.sched.step_tab[4 * lane + 1] : 0.f;
    mbar_wait(bar_w, 0);

    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int64_t grow = (int64_t)tile * TILE_M + row;
      const bool valid = grow < a.rows;
      const int64_t lrow = valid ? grow : a.rows - 1;        // row the unconditional loads of this thread read
      const bool use_alt = DUAL && valid && (a.alt_mask[grow] == 0);
      const int gcol = use_alt ? 128 : 64;
      const bool w_all_alt = DUAL && __all_sync(0xffffffffu, use_alt);
      const bool w_mixed = DUAL && !w_all_alt && __any_sync(0xffffffffu, use_alt);
      const uint32_t ucol = w_all_alt ? 128 : 64;
      const float* c2v = vec + (use_alt ? VEC_C2A : VEC_C2) + hh * 32;
      const float* w3v = vec + (use_alt ? VEC_W3GA : VEC_W3G) + hh * 32;
      const float c3b = vec[use_alt ? VEC_C3A : VEC_C3];

      // ---- prologue: h0 -> Y (TMEM) + AH --------------------------------------------------------------------------------
      {
        uint32_t yv[32];
        float t[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (valid) v = *reinterpret_cast<const float4*>(a.h0 + grow * a.h0_row_stride + hh * 32 + 4 * q);
          t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) yv[j] = __float_as_uint(t[j]);
        tmem_st_32x32b_x32(tm + TM_Y, yv);
        st_operand32(o_ah, t);
        // GRU input x of iteration 0 -> AX.  Later iterations fetch theirs one GRU phase ahead (after the Euler update of the previous
        // iteration) and store it once G3 has read the old one, so epilogue 1 never waits for a global load.
        const int slot0 = __shfl_sync(0xffffffffu, lane_slot, 0);
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f), w = v;
          if (valid) ld_nc_f8(a.aa_out + ((int64_t)slot0 * a.rows + grow) * 64 + hh * 32 + 4 * q, v, w);
          t[4 * q] = v.x; t[4 * q + 1] = v.y; t[4 * q + 2] = v.z; t[4 * q + 3] = v.w;
          t[4 * q + 4] = w.x; t[4 * q + 5] = w.y; t[4 * q + 6] = w.z; t[4 * q + 7] = w.w;
        }
        st_operand32(o_ax, t);
        tc_wait_st();
      }
      tc_fence_before();
      mbar_arrive(bar_opnd(0));                            // AH ready -> P1 of iteration 0

      // observation mask of this row for all iterations as one bit mask (21 independent byte loads here instead of one dependent
      // load at the top of every iteration)
      uint32_t obs_bits = 0;
      {
        const uint8_t* mrow = a.obs_mask + (valid ? grow : 0) * a.obs_mask_row_stride;   // rows past the end read row 0 and are never stored
        for (int it = 0; it < S; ++it) {
          const int sl = __shfl_sync(0xffffffffu, lane_slot, it);                      // warp-uniform: outside any divergent branch
          obs_bits |= (mrow[sl] != 0 ? 1u : 0u) << it;
        }
      }

      for (int it = 0; it < S; ++it) {
#ifdef TRAJSDE_ENC_TIMELINE
        if (threadIdx.x == 0 && blockIdx.x == 0) g_enc_tl[2] += 1;
#endif
        const float h = __shfl_sync(0xffffffffu, lane_h, it);                // step size / slot index of iteration `it`: held by lane `it`
        const int slot_next = __shfl_sync(0xffffffffu, lane_slot, (it + 1) & 31);
        const float* b1row = bias1_tab + it * 192;
        // ---- early global loads of this iteration: Brownian increments ------------------------------------------------------
        // Supplied increments of this iteration.  Unconditional loads from a clamped row — rows past the end read the last row and are
        // never stored; a predicated load into zero-initialised registers makes the compiler merge the two values right behind the load.
        // (Issuing them one iteration ahead changes nothing: what the kernel pays for its per-thread row accesses is L1 tag throughput —
        // every 256-bit access of a warp touches 32 different lines — and that ~2 k clk per iteration shows up at whichever global
        // access comes first, bench_micro/enc_fwd_timeline.py.)
        float4 dwv[8];
        const bool observed = (obs_bits >> it) & 1u;
        if (HAS_DW) {
          const float* ds = a.noise.dw + ((int64_t)it * a.rows + lrow) * 64 + hh * 32;
#pragma unroll
          for (int q = 0; q < 8; q += 2) ld_nc_f8(ds + 4 * q, dwv[q], dwv[q + 1]);
        }

        // ---- epilogue 1 ------------------------------------------------------------------------------------------------------
        ENC_TL_WAIT(0, mbar_wait(bar_acc(0), par_accA));                   // P1
        par_accA ^= 1;
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tm, v);
          tc_wait_ld();
          tanh32_to_operand(v, b1row + hh * 32, o_a1f);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_opnd(1));                        // h1f -> P2f
          ld_g32<DUAL>(tm + ucol, tm + 128, w_mixed, use_alt, v);
          tanh32_to_operand(v, b1row + gcol + hh * 32, o_a1g);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_opnd(0));                        // h1g -> P2g
        }
        // ---- epilogue 2 ------------------------------------------------------------------------------------------------------
        ENC_TL_WAIT(1, mbar_wait(bar_acc(1), par_accB));                   // P2f
        par_accB ^= 1;
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tm, v);
          tc_wait_ld();
          tanh32_to_operand(v, vec + VEC_B2 + hh * 32, o_ah);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_opnd(1));                        // h2f -> P3
          ENC_TL_WAIT(2, mbar_wait(bar_acc(0), par_accA));                 // P2g
          par_accA ^= 1;
          tc_fence_after();
          ld_g32<DUAL>(tm + ucol, tm + 128, w_mixed, use_alt, v);
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
        // ---- epilogue 3: Euler update ----------------------------------------------------------------------------------------------
        if (!HAS_DW) {
          const float sqrt_h = sqrtf(h);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            dwv[q] = philox_dw4(noise_seed, (uint64_t)grow + a.noise.row_offset, a.noise.step_offset + (uint32_t)it,
                                (uint32_t)(hh * 8 + q), sqrt_h);
          }
        }
        ENC_TL_WAIT(3, mbar_wait(bar_acc(1), par_accB));                   // P3
        par_accB ^= 1;
        tc_fence_after();
        named_bar_sync(pair_bar, 64);                      // both partial diffusion dots of every row are in smem
        const float g = __fdividef(1.0f, 1.0f + __expf(-((gpart[row] + gpart[TILE_M + row]) + c3b)));
        {
          uint32_t yv[32], fv[32];
          tmem_ld_32x32b_x32(tm + TM_Y, yv);
          tmem_ld_32x32b_x32(tm, fv);
          tc_wait_ld();
          float t[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b3 = *reinterpret_cast<const float4*>(vec + VEC_B3 + hh * 32 + 4 * q);
            t[4 * q] = fmaf(g, dwv[q].x, fmaf(__uint_as_float(fv[4 * q]) + b3.x, h, __uint_as_float(yv[4 * q])));
            t[4 * q + 1] = fmaf(g, dwv[q].y, fmaf(__uint_as_float(fv[4 * q + 1]) + b3.y, h, __uint_as_float(yv[4 * q + 1])));
            t[4 * q + 2] = fmaf(g, dwv[q].z, fmaf(__uint_as_float(fv[4 * q + 2]) + b3.z, h, __uint_as_float(yv[4 * q + 2])));
            t[4 * q + 3] = fmaf(g, dwv[q].w, fmaf(__uint_as_float(fv[4 * q + 3]) + b3.w, h, __uint_as_float(yv[4 * q + 3])));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) yv[j] = __float_as_uint(t[j]);
          tmem_st_32x32b_x32(tm + TM_Y, yv);               // Y <- y1
          st_operand32(o_ah, t);
          tc_wait_st();
          tc_fence_before();
          mbar_arrive(bar_opnd(0));                        // y1 -> G1 (x is in AX since the previous iteration / the tile prologue)
          // the global stores go out behind the hand-over: G1 runs while they drain
          if (hh == 0 && valid) a.g_out[(int64_t)it * a.rows + grow] = g;
          if (a.y1_out && valid) {                         // pre-GRU state, saved for the backward call
            float* dst = a.y1_out + ((int64_t)it * a.rows + grow) * 64 + hh * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_cs_f8(dst + 8 * q, make_float4(t[8 * q], t[8 * q + 1], t[8 * q + 2], t[8 * q + 3]), make_float4(t[8 * q + 4], t[8 * q + 5], t[8 * q + 6], t[8 * q + 7]));
          }
        }
        // next iteration's GRU input: in flight during G1..G3, stored to AX after G3 (its last reader this iteration)
        float4 xn[8];
        {                                                    // (slot_next of the last iteration wraps to a valid slot; its x is never used)
          const float* xs = a.aa_out + ((int64_t)slot_next * a.rows + lrow) * 64 + hh * 32;
#pragma unroll
          for (int q = 0; q < 8; q += 2) ld_nc_f8(xs + 4 * q, xn[q], xn[q + 1]);
        }
        // ---- GRU epilogue 1: tu = tanh(zu + ub1), tr = tanh(zr + rb1) -----------------------------------------------------------------
        ENC_TL_WAIT(4, mbar_wait(bar_acc(0), par_accA));                   // G1
        par_accA ^= 1;
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tm, v);
          tc_wait_ld();
          tanh32_to_operand(v, vec + VEC_UB1 + hh * 32, o_a1f);
          tmem_ld_32x32b_x32(tm + 64, v);
          tc_wait_ld();
          tanh32_to_operand(v, vec + VEC_RB1 + hh * 32, o_a1g);
          tc_wait_st();
        }
        tc_fence_before();
        mbar_arrive(bar_opnd(1));                          // tu, tr -> G2
        // ---- GRU epilogue 2: u = sigmoid(u' + ub2) (kept), r = sigmoid(r' + rb2), r * y1 -> AH ---------------------------------------------
        float u[32];
        ENC_TL_WAIT(5, mbar_wait(bar_acc(1), par_accB));                   // G2
        par_accB ^= 1;
        tc_fence_after();
        {
          uint32_t v[32], yv[32];
          tmem_ld_32x32b_x32(tm, v);
          tmem_ld_32x32b_x32(tm + TM_Y, yv);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = act1<true>(__uint_as_float(v[j]) + vec[VEC_UB2 + hh * 32 + j]);
          tmem_ld_32x32b_x32(tm + 64, v);
          tc_wait_ld();
          float t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j)
            t[j] = act1<true>(__uint_as_float(v[j]) + vec[VEC_RB2 + hh * 32 + j]) * __uint_as_float(yv[j]);
          st_operand32(o_ah, t);                           // AH (y1 as fp16) was consumed by G1
          tc_wait_st();
        }
        tc_fence_before();
        mbar_arrive(bar_opnd(0));                          // r*y1 -> G3
        // ---- GRU epilogue 3: tn = tanh(zn + nb1) --------------------------------------------------------------------------------------------
        ENC_TL_WAIT(6, mbar_wait(bar_acc(0), par_accA));                   // G3
        par_accA ^= 1;
        tc_fence_after();
        {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tm, v);
          tc_wait_ld();
          tanh32_to_operand(v, vec + VEC_NB1 + hh * 32, o_a1f);
          float t[32];                                     // x of the next iteration -> AX (G3 of this iteration has completed)
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            t[4 * q] = xn[q].x; t[4 * q + 1] = xn[q].y; t[4 * q + 2] = xn[q].z; t[4 * q + 3] = xn[q].w;
          }
          st_operand32(o_ax, t);
          tc_wait_st();
        }
        tc_fence_before();
        mbar_arrive(bar_opnd(1));                          // tn -> G4
        // ---- GRU epilogue 4: h' = (1-u) (n + nb2) + u y1 ; masked ; state, operand, latent ---------------------------------------------------
        ENC_TL_WAIT(7, mbar_wait(bar_acc(1), par_accB));                   // G4
        par_accB ^= 1;
        tc_fence_after();
        {
          uint32_t v[32], yv[32];
          tmem_ld_32x32b_x32(tm, v);
          tmem_ld_32x32b_x32(tm + TM_Y, yv);
          tc_wait_ld();
          float t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float y1 = __uint_as_float(yv[j]);
            const float n = __uint_as_float(v[j]) + vec[VEC_NB2 + hh * 32 + j];
            const float hn = fmaf(u[j], y1, (1.0f - u[j]) * n);
            t[j] = observed ? hn : y1;
            yv[j] = __float_as_uint(t[j]);
          }
          tmem_st_32x32b_x32(tm + TM_Y, yv);
          st_operand32(o_ah, t);                           // AH (r*y1) was consumed by G3
          tc_wait_st();
          tc_fence_before();
          if (it + 1 < S) mbar_arrive(bar_opnd(0));        // h' -> P1 of the next iteration
          if (valid) {                                     // the latent store goes out behind the hand-over
            float* dst = a.latent + ((int64_t)it * a.rows + grow) * 64 + hh * 32;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              st_cs_f8(dst + 8 * q, make_float4(t[8 * q], t[8 * q + 1], t[8 * q + 2], t[8 * q + 3]), make_float4(t[8 * q + 4], t[8 * q + 5], t[8 * q + 6], t[8 * q + 7]));
          }
        }
#ifdef TRAJSDE_ENC_TIMELINE
        { const long long _t = clock64(); if (threadIdx.x == 0 && blockIdx.x == 0) g_enc_seg[16] += _t - seg_prev; seg_prev = _t; }
#endif
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
    if (warp == NUM_EPI_WARPS) {
      // =============================================== MMA ISSUER =====================================================
      // warp-uniform loop (descriptors stay in uniform registers); one elected lane issues tcgen05.mma / tcgen05.commit.
      // A operands come from tensor memory (fp16 pairs, 8 columns per K = 16 instruction), B operands from the staged weight tiles.
      const uint32_t idesc_p1 = umma_idesc_f16(TILE_M, DUAL ? 192u : 128u);
      const uint32_t idesc_64 = umma_idesc_f16(TILE_M, 64);
      const uint32_t idesc_128 = umma_idesc_f16(TILE_M, 128);
      const uint64_t dhi = umma_desc_sw128(0);
      auto D = [&](uint32_t addr) { return dhi | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      const uint32_t d0 = tmem_base;
      const uint32_t aAH = d0 + TM_AH, aAX = d0 + TM_AX, aA1f = d0 + TM_A1F, aA1g = d0 + TM_A1G;
      uint32_t par_op0 = 0, par_op1 = 0;
      mbar_wait(bar_w, 0);
      auto wait0 = [&]() { mbar_wait(bar_opnd(0), par_op0); par_op0 ^= 1; tc_fence_after(); };
      auto wait1 = [&]() { mbar_wait(bar_opnd(1), par_op1); par_op1 ^= 1; tc_fence_after(); };
      auto mma4 = [&](uint32_t d, uint32_t a_tmem, uint32_t baddr, uint32_t idesc, bool acc_first) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc_mma_f16_ts(d, a_tmem + 8 * kk, D(base + baddr + 32 * kk), idesc, (acc_first || kk > 0) ? 1u : 0u);
      };
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int it = 0; it < S; ++it) {
          wait0();                                                                                         // P1
          if (elect_one()) { mma4(d0, aAH, IMG_B1, idesc_p1, false); tc_commit(bar_acc(0)); }
          __syncwarp();
          wait1();                                                                                         // P2f
          if (elect_one()) { mma4(d0, aA1f, IMG_W2, idesc_64, false); tc_commit(bar_acc(1)); }
          __syncwarp();
          wait0();                                                                                         // P2g
          if (elect_one()) {
            mma4(d0 + 64, aA1g, IMG_V2, idesc_64, false);
            if (DUAL) mma4(d0 + 128, aA1g, IMG_V2A, idesc_64, false);
            tc_commit(bar_acc(0));
          }
          __syncwarp();
          wait1();                                                                                         // P3
          if (elect_one()) { mma4(d0, aAH, IMG_W3, idesc_64, false); tc_commit(bar_acc(1)); }
          __syncwarp();
          wait0();                                                                                         // G1: [zu|zr] -> [0,128)
          if (elect_one()) {
            mma4(d0, aAH, IMG_UR1H, idesc_128, false);
            mma4(d0, aAX, IMG_UR1X, idesc_128, true);
            tc_commit(bar_acc(0));
          }
          __syncwarp();
          wait1();                                                                                         // G2: u' -> [0,64), r' -> [64,128)
          if (elect_one()) {
            mma4(d0, aA1f, IMG_U2, idesc_64, false);
            mma4(d0 + 64, aA1g, IMG_R2, idesc_64, false);
            tc_commit(bar_acc(1));
          }
          __syncwarp();
          wait0();                                                                                         // G3: zn -> [0,64)
          if (elect_one()) {
            mma4(d0, aAX, IMG_N1X, idesc_64, false);
            mma4(d0, aAH, IMG_N1RH, idesc_64, true);
            tc_commit(bar_acc(0));
          }
          __syncwarp();
          wait1();                                                                                         // G4: n -> [0,64)
          if (elect_one()) { mma4(d0, aA1f, IMG_N2, idesc_64, false); tc_commit(bar_acc(1)); }
          __syncwarp();
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef TRAJSDE_ENC_TIMELINE
  if (threadIdx.x == 0 && blockIdx.x == 0) g_enc_tl[0] += clock64() - tl_start;
#endif
  if (warp == NUM_EPI_WARPS) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

int64_t enc_fwd_tc_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual) {
  (void)rows; (void)n_steps; (void)dual;
  return (int64_t)IMG_BYTES + 256;
}

int launch_enc_fwd_tc(const TrajsdeEncFwdArgs& a, cudaStream_t s) {
  int dev = 0, sms = 0;
  TS_CUDA_CHECK(cudaGetDevice(&dev));
  TS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (a.sched.n_steps > S_MAX) return set_error(TRAJSDE_ERR_UNSUPPORTED, "fused encoder supports at most %d iterations, got %d", S_MAX, a.sched.n_steps);
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0) return set_error(TRAJSDE_ERR_UNSUPPORTED, "workspace must be 256-byte aligned");
  EncParams p;
  p.a = a;
  p.img = static_cast<const uint8_t*>(a.workspace);
  p.num_tiles = (int)((a.rows + TILE_M - 1) / TILE_M);
  p.dual = a.alt_mask != nullptr;
  if (p.num_tiles == 0) return TRAJSDE_OK;
  enc_pack_kernel<<<32, 256, 0, s>>>(a, static_cast<uint8_t*>(a.workspace), p.dual);
  TS_CUDA_CHECK(cudaGetLastError());
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC);
    if (e != cudaSuccess) return e;
    kern<<<grid, NUM_THREADS, SMEM_ALLOC, s>>>(p);
    return cudaGetLastError();
  };
  const bool hd = a.noise.dw != nullptr, du = p.dual != 0;
  TS_CUDA_CHECK(hd ? (du ? launch(enc_fwd_tc_kernel<true, true>) : launch(enc_fwd_tc_kernel<true, false>))
                   : (du ? launch(enc_fwd_tc_kernel<false, true>) : launch(enc_fwd_tc_kernel<false, false>)));
  return TRAJSDE_OK;
}

}  // namespace trajsde

#ifdef TRAJSDE_ENC_TIMELINE
extern "C" int trajsde_debug_enc_segments(long long* out20) {
  long long zero[20] = {0};
  if (cudaMemcpyFromSymbol(out20, trajsde::g_enc_seg, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(trajsde::g_enc_seg, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
extern "C" int trajsde_debug_enc_timeline(long long* out4) {
  long long zero[4] = {0};
  if (cudaMemcpyFromSymbol(out4, trajsde::g_enc_tl, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(trajsde::g_enc_tl, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif
