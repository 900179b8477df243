// Device body of the tensor-core GRU_Unit backward (see gru_bwd_tc.cu for the description).  A header so that the single-launch
// encoder-recurrence sweep (enc_bwd_sweep.cu) can run it as one ROLE of a merged kernel next to the SDE-step backward.
#pragma once
#include "bwd_common.cuh"
#include "bwd_tc_common.cuh"

namespace trajsde {
namespace grutc {

using namespace tc;
using namespace bwd;
using namespace bwdtc;

constexpr int TILE_M = 128;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = NUM_EPI_THREADS + 128;
constexpr int EPI_REGS = 216, AUX_REGS = 64;

// ---- packed weight image: fp16 SW128 tiles [N rows][64 k] + fp32 bias vectors ------------------------------------------------
constexpr uint32_t IMG_UR1H = 0;        // [128][64]: rows 0..63 U1[:, :64], rows 64..127 R1[:, :64]   (h_cur part)
constexpr uint32_t IMG_UR1X = 16384;    // [128][64]: U1[:, 64:], R1[:, 64:]                            (input part)
constexpr uint32_t IMG_U2 = 32768, IMG_R2 = 40960;
constexpr uint32_t IMG_N1X = 49152, IMG_N1RH = 57344;   // N1[:, :64] (input part), N1[:, 64:] (r*h part); adjacent
constexpr uint32_t IMG_N2 = 65536;
constexpr uint32_t IMG_VEC = 73728;
constexpr int VEC_UB1 = 0, VEC_RB1 = 64, VEC_NB1 = 128, VEC_UB2 = 192, VEC_RB2 = 256, VEC_NB2 = 320;
constexpr uint32_t IMG_BYTES = IMG_VEC + 384 * 4;   // 75264
static_assert(IMG_BYTES <= GRU_TC_IMG_BYTES, "image larger than its workspace slot");

// ---- shared memory map -----------------------------------------------------------------------------------------------------------
constexpr uint32_t OFF_TILES = 75776;
constexpr uint32_t TILE_BYTES = 16384;
constexpr int T_Y1 = 0, T_X = 1;        // [y1|x] = tiles 0,1 ; [x|r y1] = tiles 1,2
constexpr int T_RY = 2, T_TN = 3;       // later d_u', d_r':  [d_u'|d_r'] = tiles 2,3
constexpr int T_TU = 4, T_TR = 5;       // [tu|tr] = tiles 4,5 (also the row-transposition staging at tile start)
constexpr int T_DZN = 6, T_DN = 7;      // [dz_n|d_n] = tiles 6,7 ; later [dz_u|dz_r]
constexpr uint32_t OFF_BSUM = OFF_TILES + 8 * TILE_BYTES;          // [8 warps][6][32] fp32 column sums
constexpr uint32_t OFF_BARS = OFF_BSUM + 8 * 6 * 32 * 4;
constexpr uint32_t SMEM_TOTAL = OFF_BARS + 64;
constexpr uint32_t SMEM_ALLOC = SMEM_TOTAL + 1024;
static_assert(SMEM_ALLOC <= 232448, "exceeds 227 KB of shared memory per CTA");

// ---- TMEM columns -------------------------------------------------------------------------------------------------------------------
constexpr uint32_t TM_W = 0;            // 128 working columns (later: dU1 | dR1 accumulators)
constexpr uint32_t TM_GN1 = 128;        // 128: lanes 0..63 dN1[m][0..127]
constexpr uint32_t TM_GN2 = 256;        // 64:  lanes 64..127 dN2[m-64][0..63]
constexpr uint32_t TM_G2 = 320;         // 128: lanes 0..63 x [0,64) dU2, lanes 64..127 x [64,128) dR2
// 128 + 128 + 64 + 128 = 448 columns; the dU1|dR1 product (128 more) is computed into the working columns once the tile's last
// epilogue has read them, and flushed to the partial vector per tile.

struct GruTcParams {
  int64_t rows;
  const float* y1;
  const float* x;              // aa_out base; slab slot[iter]
  int64_t x_slab;
  const uint8_t* obs_mask;
  int64_t obs_mask_row_stride;
  const int32_t* slot;
  int iter;
  const float* carry;
  const float* grad_latent;
  float* grad_y1;
  float* grad_x;               // grad_aa_out base or NULL
  const uint8_t* img;
  const uint32_t* amax_bits;
  float* partial;              // [grid][GRU_G_PAD], accumulated into
  int num_tiles;
  int fwd_only;                // 1: stop after the recompute phases and write h' (the stand-alone GRU_Unit forward, trajsde_gru_fwd)
  float* h_out;                // [rows,64], fwd_only
};

#ifdef TRAJSDE_GRU_TIMELINE
// debug build only (bench_micro/enc_bwd_ab.py --gru-timeline): clocks of thread 0 of CTA 0 between consecutive marks of a GRU tile
static __device__ long long g_gru_seg[24];
#define GRU_MARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _t = clock64(); g_gru_seg[i] += _t - seg_prev; seg_prev = _t; } } while (0)
#else
#define GRU_MARK(i) do { } while (0)
#endif

__device__ __forceinline__ float sigmoid_mufu(float x) { return fmaf(0.5f, ts_tanh_approx(0.5f * x), 0.5f); }

// SWEEP: the GRU role of the single-launch encoder sweep — every iteration i = S-1 .. 0 of the recurrence over this CTA's tiles, pointers
// of iteration i derived from the bases in `p` (y1, grad_latent: slab i; aa_out slab slot[i]), the carried adjoint read once the SDE role(s)
// have published it, dL/dy1 published for them.  (cta, ncta): this CTA's index among the CTAs of the role.
template <bool SWEEP>
__device__ __forceinline__ void gru_bwd_tc_body(const GruTcParams& p, const SweepCtl& sw, const int cta, const int ncta, uint8_t* smem_raw) {
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int tiles_q = p.num_tiles / ncta, tiles_r = p.num_tiles % ncta;
  const int tile_lo = cta * tiles_q + min(cta, tiles_r);
  const int tile_hi = tile_lo + tiles_q + (cta < tiles_r ? 1 : 0);
  const int it_hi = SWEEP ? sw.S - 1 : p.iter, it_lo = SWEEP ? 0 : p.iter;

  const uint32_t bar_w = base + OFF_BARS, bar_opnd = bar_w + 8, bar_acc = bar_w + 16, bar_wg = bar_w + 24;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + OFF_BARS + 32);
  auto tile_u32 = [&](int t) { return base + OFF_TILES + (uint32_t)t * TILE_BYTES; };

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_opnd, NUM_EPI_THREADS);
    mbar_init(bar_acc, 1);
    mbar_init(bar_wg, 1);
    mbar_fence_init();
  }
  if (warp == NUM_EPI_WARPS) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // nothing above touches global memory; everything below may depend on the previous kernel of the stream
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_w, IMG_BYTES);
    bulk_load_1d(base, p.img, IMG_BYTES, bar_w);
  }
  const float* vec = reinterpret_cast<const float*>(sm + IMG_VEC);

  if (warp < NUM_EPI_WARPS) {
    // =============================================== EPILOGUE WARPS ===============================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    const int quad = warp & 3;
    const uint32_t hh = (uint32_t)warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t tm = tmem_base + ((uint32_t)(quad * 32) << 16) + hh * 32;
    auto trow = [&](int t) { return sm + OFF_TILES + (uint32_t)t * TILE_BYTES + row * 128; };
    uint8_t* stage = sm + OFF_TILES + T_TU * TILE_BYTES + (uint32_t)warp * 4096;

    float sigma = 1.f, inv_sigma = 1.f;
    {
      const float amax = __uint_as_float(*p.amax_bits);
      if (amax > 0.f) {
        int e;
        frexpf(amax, &e);
        e = max(-100, min(100, -e - 3));
        sigma = ldexpf(1.f, e);
        inv_sigma = ldexpf(1.f, -e);
      }
    }
#ifdef TRAJSDE_GRU_TIMELINE
    long long seg_prev = clock64();
#endif
    float bs_n2 = 0.f, bs_n1 = 0.f, bs_u2 = 0.f, bs_r2 = 0.f, bs_u1 = 0.f, bs_r1 = 0.f;   // column (hh*32 + lane) sums over this warp's rows
    uint32_t hs = 0, ntile = 0;
    mbar_wait(bar_w, 0);

    for (int it = it_hi; it >= it_lo; --it)
    for (int tile = tile_lo; tile < tile_hi; ++tile, ++ntile) {
      const int slot = p.slot[it];
      const float* xin = p.x + (int64_t)slot * p.x_slab;
      float* gx = p.grad_x ? p.grad_x + (int64_t)slot * p.x_slab : nullptr;
      const float* y1p = SWEEP ? p.y1 + (int64_t)it * p.x_slab : p.y1;
      const float* carry = SWEEP ? (it == it_hi ? nullptr : p.carry) : p.carry;
      const float* glat = SWEEP && p.grad_latent ? p.grad_latent + (int64_t)it * p.x_slab : p.grad_latent;
      GRU_MARK(0);   // previous tile's flush tail / loop
      const int64_t row0 = (int64_t)tile * TILE_M + quad * 32;
      const int64_t grow = (int64_t)tile * TILE_M + row;
      const bool valid = grow < p.rows;
      const bool obs = valid && p.obs_mask[grow * p.obs_mask_row_stride + slot] != 0;

      // ---- tile start: rows in (coalesced), y1 / x operand tiles, a = sigma (carry + dL/dlatent) ------------------------------------
      float y1[32], a[32];
      {
        float4 ly[8], lx[8], lc[8], lg[8];
        load_rows_coalesced(y1p, 64, row0, p.rows, hh * 32, lane, ly);
        load_rows_coalesced(xin, 64, row0, p.rows, hh * 32, lane, lx);
        if (glat) load_rows_coalesced(glat, 64, row0, p.rows, hh * 32, lane, lg);
        if (SWEEP) {
          if (carry) {                                             // written by the SDE role(s) of this launch during iteration it + 1
            sweep_wait(sw, sw.sde_done[0] + tile, sw.S - 1 - it);
            if (sw.sde_done[1]) sweep_wait(sw, sw.sde_done[1] + tile, sw.S - 1 - it);
            load_rows_coalesced_cg(carry, 64, row0, p.rows, hh * 32, lane, lc);
          }
        } else if (carry) {
          load_rows_coalesced(carry, 64, row0, p.rows, hh * 32, lane, lc);
        }
        if (ntile > 0 && !p.fwd_only) mbar_wait(bar_wg, (ntile - 1) & 1);   // previous tile's weight-gradient MMAs have read every tile
        to_own_row(stage, lane, ly);
        to_own_row(stage, lane, lx);
        if (carry) to_own_row(stage, lane, lc);
        if (glat) to_own_row(stage, lane, lg);
        float t[32];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          y1[4 * q] = ly[q].x; y1[4 * q + 1] = ly[q].y; y1[4 * q + 2] = ly[q].z; y1[4 * q + 3] = ly[q].w;
          t[4 * q] = lx[q].x; t[4 * q + 1] = lx[q].y; t[4 * q + 2] = lx[q].z; t[4 * q + 3] = lx[q].w;
          float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
          if (carry) s = lc[q];
          if (glat) { s.x += lg[q].x; s.y += lg[q].y; s.z += lg[q].z; s.w += lg[q].w; }
          a[4 * q] = valid ? s.x * sigma : 0.f; a[4 * q + 1] = valid ? s.y * sigma : 0.f;
          a[4 * q + 2] = valid ? s.z * sigma : 0.f; a[4 * q + 3] = valid ? s.w * sigma : 0.f;
        }
        st_row32(trow(T_X), row, hh, t);
        st_row32(trow(T_Y1), row, hh, y1);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // y1, x -> F1
      GRU_MARK(1);   // tile start: waits for the other roles, row loads, transposes, operand stores

      uint32_t v[32];
      float t[32];
      // ---- F1: tu, tr ------------------------------------------------------------------------------------------------------------------
      GRU_MARK(2);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(3);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = ts_tanh_approx(__uint_as_float(v[j]) + vec[VEC_UB1 + hh * 32 + j]);
      st_row32(trow(T_TU), row, hh, t);
      tmem_ld_32x32b_x32(tm + TM_W + 64, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = ts_tanh_approx(__uint_as_float(v[j]) + vec[VEC_RB1 + hh * 32 + j]);
      st_row32(trow(T_TR), row, hh, t);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // tu, tr -> F2
      // ---- F2: u, r ; r*y1 ---------------------------------------------------------------------------------------------------------------
      float u[32], r[32];
      GRU_MARK(4);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(5);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) u[j] = sigmoid_mufu(__uint_as_float(v[j]) + vec[VEC_UB2 + hh * 32 + j]);
      tmem_ld_32x32b_x32(tm + TM_W + 64, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        r[j] = sigmoid_mufu(__uint_as_float(v[j]) + vec[VEC_RB2 + hh * 32 + j]);
        t[j] = r[j] * y1[j];
      }
      st_row32(trow(T_RY), row, hh, t);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // r*y1 -> F3
      // ---- F3: tn ------------------------------------------------------------------------------------------------------------------------
      GRU_MARK(6);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(7);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = ts_tanh_approx(__uint_as_float(v[j]) + vec[VEC_NB1 + hh * 32 + j]);
      st_row32(trow(T_TN), row, hh, t);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // tn -> F4
      // ---- F4: n ; d_n = a (1-u) ; d_u' = a (y1 - n) u (1-u) ; d_y1 = a u   (a = 0 on unobserved rows, which pass dL/dh' straight on) ------
      float dy1[32], dup[32];
      GRU_MARK(8);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(9);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      tc_wait_ld();
      if (p.fwd_only) {                                            // h' = mask ? (1-u) n + u y1 : y1
        tc_fence_before();
        if (valid) {
          float* dst = p.h_out + grow * 64 + hh * 32;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 4 * q + e;
              const float n = __uint_as_float(v[j]) + vec[VEC_NB2 + hh * 32 + j];
              o[e] = obs ? fmaf(u[j], y1[j], (1.f - u[j]) * n) : y1[j];
            }
            *reinterpret_cast<float4*>(dst + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
          }
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float n = __uint_as_float(v[j]) + vec[VEC_NB2 + hh * 32 + j];
        const float ae = obs ? a[j] : 0.f;
        t[j] = ae * (1.f - u[j]);
        dup[j] = ae * (y1[j] - n) * u[j] * (1.f - u[j]);
        dy1[j] = obs ? ae * u[j] : a[j];
      }
      st_row32(trow(T_DN), row, hh, t);
      bs_n2 += colsum32(t, lane);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // d_n -> B1
      // ---- B1: dz_n = d_tn (1 - tn^2) -------------------------------------------------------------------------------------------------------
      GRU_MARK(10);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(11);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      ld_row32(trow(T_TN), row, hh, t);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]) * fmaf(-t[j], t[j], 1.f);
      st_row32(trow(T_DZN), row, hh, t);
      bs_n1 += colsum32(t, lane);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // dz_n -> dN1, dN2 products, then B2
      // ---- B2: d_x (part), d(r y1) -> d_r', d_y1 ; d_u', d_r' tiles ---------------------------------------------------------------------------
      float dx[32];
      GRU_MARK(12);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(13);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) dx[j] = __uint_as_float(v[j]);
      tmem_ld_32x32b_x32(tm + TM_W + 64, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float dry = __uint_as_float(v[j]);
        dy1[j] = fmaf(dry, r[j], dy1[j]);
        t[j] = dry * y1[j] * r[j] * (1.f - r[j]);                  // d_r'
      }
      st_row32(trow(T_TN), row, hh, t);                            // tile 3: d_r'
      bs_r2 += colsum32(t, lane);
      st_row32(trow(T_RY), row, hh, dup);                          // tile 2: d_u'
      bs_u2 += colsum32(dup, lane);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // d_u', d_r' -> B3
      // ---- B3: dz_u = d_tu (1 - tu^2), dz_r = d_tr (1 - tr^2) ------------------------------------------------------------------------------------
      GRU_MARK(14);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(15);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      ld_row32(trow(T_TU), row, hh, t);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]) * fmaf(-t[j], t[j], 1.f);
      st_row32(trow(T_DZN), row, hh, t);                           // tile 6: dz_u
      bs_u1 += colsum32(t, lane);
      tmem_ld_32x32b_x32(tm + TM_W + 64, v);
      ld_row32(trow(T_TR), row, hh, t);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]) * fmaf(-t[j], t[j], 1.f);
      st_row32(trow(T_DN), row, hh, t);                            // tile 7: dz_r
      bs_r1 += colsum32(t, lane);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // dz_u, dz_r -> B4
      // ---- B4: d_y1, d_x complete -> global ----------------------------------------------------------------------------------------------------
      GRU_MARK(16);
      mbar_wait(bar_acc, hs & 1); ++hs;
      GRU_MARK(17);
      tc_fence_after();
      tmem_ld_32x32b_x32(tm + TM_W, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) dy1[j] = (dy1[j] + __uint_as_float(v[j])) * inv_sigma;
      tmem_ld_32x32b_x32(tm + TM_W + 64, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) dx[j] = (dx[j] + __uint_as_float(v[j])) * inv_sigma;
      tc_fence_before();
      mbar_arrive(bar_opnd);                                       // working columns read -> dU2|dR2, dU1|dR1 products
      if (valid) {
        float* d1 = p.grad_y1 + grow * 64 + hh * 32;               // 256-bit stores: whole 32-byte sectors (see st_f8)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          st_f8(d1 + 8 * q, make_float4(dy1[8 * q], dy1[8 * q + 1], dy1[8 * q + 2], dy1[8 * q + 3]),
                make_float4(dy1[8 * q + 4], dy1[8 * q + 5], dy1[8 * q + 6], dy1[8 * q + 7]));
        if (gx) {
          float* d2 = gx + grow * 64 + hh * 32;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            st_f8(d2 + 8 * q, make_float4(dx[8 * q], dx[8 * q + 1], dx[8 * q + 2], dx[8 * q + 3]),
                  make_float4(dx[8 * q + 4], dx[8 * q + 5], dx[8 * q + 6], dx[8 * q + 7]));
        }
      }
      if (SWEEP) sweep_publish(sw.gru_done + tile, sw.S - it, 6, NUM_EPI_THREADS, threadIdx.x == 0);   // dL/dy1 of the tile is out
      // ---- dU1 | dR1 live in the working columns: flush them for this tile before the next F1 overwrites them -------------------------------------
      GRU_MARK(18);  // B4 epilogue: stores + publish
      mbar_wait(bar_wg, ntile & 1);
      GRU_MARK(19);  // wait for the weight-gradient products
      tc_fence_after();
      {
        float* out = p.partial + (size_t)cta * GRU_G_PAD;
        const bool lo = quad < 2;
        const int m = (int)row & 63;
        float* d = out + (lo ? GRU_U1 : GRU_R1) + m * 128;
        tmem_ld_32x32b_x32(tm + TM_W, v);                          // columns hh*32 .. of [0,64): h_cur part
        tc_wait_ld();
        flush_row32(d + hh * 32, v, inv_sigma, true);
        tmem_ld_32x32b_x32(tm + TM_W + 64, v);                     // [64,128): input part
        tc_wait_ld();
        flush_row32(d + 64 + hh * 32, v, inv_sigma, true);
      }
      tc_fence_before();
      GRU_MARK(20);  // per-tile dU1 | dR1 flush
    }

    // ================= remaining weight-gradient accumulators + bias sums of this CTA -> partial =================================================
    if (!p.fwd_only) {
      float* out = p.partial + (size_t)cta * GRU_G_PAD;
      const bool lo = quad < 2;
      const int m = (int)row & 63;
      uint32_t v[32];
      tc_fence_after();
      if (lo) {
        float* d = out + GRU_N1 + m * 128;
        tmem_ld_32x32b_x32(tm + TM_GN1, v);
        tc_wait_ld();
        flush_row32(d + hh * 32, v, inv_sigma, true);
        tmem_ld_32x32b_x32(tm + TM_GN1 + 64, v);
        tc_wait_ld();
        flush_row32(d + 64 + hh * 32, v, inv_sigma, true);
        d = out + GRU_U2 + m * 64 + hh * 32;
        tmem_ld_32x32b_x32(tm + TM_G2, v);
        tc_wait_ld();
        flush_row32(d, v, inv_sigma, true);
      } else {
        float* d = out + GRU_N2 + m * 64 + hh * 32;
        tmem_ld_32x32b_x32(tm + TM_GN2, v);
        tc_wait_ld();
        flush_row32(d, v, inv_sigma, true);
        d = out + GRU_R2 + m * 64 + hh * 32;
        tmem_ld_32x32b_x32(tm + TM_G2 + 64, v);
        tc_wait_ld();
        flush_row32(d, v, inv_sigma, true);
      }
      // bias gradients: this lane's column (hh*32 + lane) summed over the warp's rows -> combine the four row quadrants
      float* bsum = reinterpret_cast<float*>(sm + OFF_BSUM) + warp * 6 * 32;
      bsum[0 * 32 + lane] = bs_u1; bsum[1 * 32 + lane] = bs_u2; bsum[2 * 32 + lane] = bs_r1;
      bsum[3 * 32 + lane] = bs_r2; bsum[4 * 32 + lane] = bs_n1; bsum[5 * 32 + lane] = bs_n2;
      named_bar_sync(5, NUM_EPI_THREADS);
      for (int idx = threadIdx.x; idx < 6 * 64; idx += NUM_EPI_THREADS) {
        const int which = idx >> 6, c = idx & 63, h2 = c >> 5, l2 = c & 31;
        const float* b0 = reinterpret_cast<const float*>(sm + OFF_BSUM);
        float s = 0.f;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) s += b0[((h2 * 4 + qd) * 6 + which) * 32 + l2];
        const int off = which == 0 ? GRU_UB1 : which == 1 ? GRU_UB2 : which == 2 ? GRU_RB1 : which == 3 ? GRU_RB2 : which == 4 ? GRU_NB1 : GRU_NB2;
        out[off + c] += s * inv_sigma;
      }
    }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
    if (warp == NUM_EPI_WARPS) {
      // =============================================== MMA ISSUER WARP ===============================================
      const uint32_t ik_128 = umma_idesc_f16(TILE_M, 128), ik_64 = umma_idesc_f16(TILE_M, 64);
      const uint32_t ikm_128 = umma_idesc_f16_k_mn(TILE_M, 128), ikm_64 = umma_idesc_f16_k_mn(TILE_M, 64);
      const uint32_t imm_128 = umma_idesc_f16_mn(TILE_M, 128), imm_64 = umma_idesc_f16_mn(TILE_M, 64);
      const uint64_t khi = umma_desc_sw128(0), m16 = umma_desc_mn_sw128(0, 16384), m8 = umma_desc_mn_sw128(0, 8192);
      auto KD = [&](uint32_t addr) { return khi | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      auto M16 = [&](uint32_t addr) { return m16 | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      auto M8 = [&](uint32_t addr) { return m8 | (uint64_t)((addr & 0x3FFFFu) >> 4); };
      // forward: D (+)= A[128 rows][64] . B[N][64]^T  (both K-major)
      auto mma_kk = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool acc_first) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc_mma_f16(d, KD(a_addr + 32 * kk), KD(b_addr + 32 * kk), idesc, (acc_first || kk > 0) ? 1u : 0u);
      };
      // backward through a layer: D (+)= A[128 rows][64 k] . W[k][n]  (A K-major, B = forward weight tile read MN-major; lbo8: the
      // two 64-column groups of B are 8 KB apart, else 16 KB)
      auto mma_km = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool lbo8, bool acc_first) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_f16(d, KD(a_addr + 32 * kk), lbo8 ? M8(b_addr + 2048 * kk) : M16(b_addr + 2048 * kk), idesc, (acc_first || kk > 0) ? 1u : 0u);
      };
      // weight gradient: D (+)= [A0|A1]^T . [B0(|B1)]  over the 128 rows (both MN-major, stacks of adjacent 16 KB tiles)
      auto mma_mm = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool acc_first) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) tc_mma_f16(d, M16(a_addr + 2048 * kk), M16(b_addr + 2048 * kk), idesc, (acc_first || kk > 0) ? 1u : 0u);
      };
      const uint32_t d0 = tmem_base;
      uint32_t hs = 0;
      bool wg_acc = false;
      auto wait_opnd = [&]() { mbar_wait(bar_opnd, hs & 1); ++hs; tc_fence_after(); };
      mbar_wait(bar_w, 0);
      for (int it = it_hi; it >= it_lo; --it)
      for (int tile = tile_lo; tile < tile_hi; ++tile) {
        wait_opnd();                                                // F1: [zu|zr] = y1 . [U1h;R1h]^T + x . [U1x;R1x]^T
        if (elect_one()) {
          mma_kk(d0 + TM_W, tile_u32(T_Y1), base + IMG_UR1H, ik_128, false);
          mma_kk(d0 + TM_W, tile_u32(T_X), base + IMG_UR1X, ik_128, true);
          tc_commit(bar_acc);
        }
        __syncwarp();
        wait_opnd();                                                // F2: u' = tu . U2^T, r' = tr . R2^T
        if (elect_one()) {
          mma_kk(d0 + TM_W, tile_u32(T_TU), base + IMG_U2, ik_64, false);
          mma_kk(d0 + TM_W + 64, tile_u32(T_TR), base + IMG_R2, ik_64, false);
          tc_commit(bar_acc);
        }
        __syncwarp();
        wait_opnd();                                                // F3: zn = x . N1x^T + (r y1) . N1rh^T
        if (elect_one()) {
          mma_kk(d0 + TM_W, tile_u32(T_X), base + IMG_N1X, ik_64, false);
          mma_kk(d0 + TM_W, tile_u32(T_RY), base + IMG_N1RH, ik_64, true);
          tc_commit(bar_acc);
        }
        __syncwarp();
        wait_opnd();                                                // F4: n = tn . N2^T
        if (elect_one()) {
          mma_kk(d0 + TM_W, tile_u32(T_TN), base + IMG_N2, ik_64, false);
          tc_commit(bar_acc);
        }
        __syncwarp();
        if (p.fwd_only) continue;
        wait_opnd();                                                // B1: d_tn = d_n . N2
        if (elect_one()) {
          mma_km(d0 + TM_W, tile_u32(T_DN), base + IMG_N2, ikm_64, false, false);
          tc_commit(bar_acc);
        }
        __syncwarp();
        wait_opnd();                                                // dN1 / dN2 products first (their tiles are recycled by the next
        if (elect_one()) {                                          // epilogue), then B2: [d_x | d(r y1)] = dz_n . [N1x | N1rh]
          mma_mm(d0 + TM_GN1, tile_u32(T_DZN), tile_u32(T_X), imm_128, wg_acc);
          mma_mm(d0 + TM_GN2, tile_u32(T_DZN), tile_u32(T_TN), imm_64, wg_acc);
          mma_km(d0 + TM_W, tile_u32(T_DZN), base + IMG_N1X, ikm_128, true, false);
          tc_commit(bar_acc);
        }
        __syncwarp();
        wait_opnd();                                                // B3: d_tu = d_u' . U2 ; d_tr = d_r' . R2
        if (elect_one()) {
          mma_km(d0 + TM_W, tile_u32(T_RY), base + IMG_U2, ikm_64, false, false);
          mma_km(d0 + TM_W + 64, tile_u32(T_TN), base + IMG_R2, ikm_64, false, false);
          tc_commit(bar_acc);
        }
        __syncwarp();
        wait_opnd();                                                // B4: [d_y1 | d_x] = dz_u . [U1h | U1x] + dz_r . [R1h | R1x]
        if (elect_one()) {
          mma_km(d0 + TM_W, tile_u32(T_DZN), base + IMG_UR1H, ikm_128, false, false);
          mma_km(d0 + TM_W, tile_u32(T_DN), base + IMG_UR1H + 64 * 128, ikm_128, false, true);
          tc_commit(bar_acc);
        }
        __syncwarp();
        wait_opnd();                                                // working columns read: dU2|dR2 (accumulated over tiles), dU1|dR1 (per tile)
        if (elect_one()) {
          mma_mm(d0 + TM_G2, tile_u32(T_RY), tile_u32(T_TU), imm_128, wg_acc);
          mma_mm(d0 + TM_W, tile_u32(T_DZN), tile_u32(T_Y1), imm_128, false);
          tc_commit(bar_wg);
        }
        __syncwarp();
        wg_acc = true;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == NUM_EPI_WARPS) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace grutc
}  // namespace trajsde
