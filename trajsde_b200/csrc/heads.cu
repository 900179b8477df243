// Fused decoder heads over the solver's outputs: for every (row, t) latent x = sol_y[row, t, :64]
//     loc   = W2d . relu(LayerNorm(W1d x + b1d)) + b2d          (self.decoder, models/decoders/dec_hivt_nusargo_sde.py:50-54, 96)
//     scale = W2s . relu(LayerNorm(W1s x + b1s)) + b2s          (self.scale,   dec_hivt_nusargo_sde.py:56-61, 98; ELU/+1/+min_scale stay with the caller)
// The reference runs ~10 aten kernels per head, each a full pass over the [M, 60, 64] tensor (3.15 GB at configs[1]); here the
// latents are read ONCE (TMA, 128B-swizzled fp32 tiles), both heads' first layers are one tcgen05 product per 128-point tile
// ([128 x 64] . [W1d; W1s]^T, fp16 operands from tensor memory, fp32 accumulate), and LayerNorm / ReLU / the 64 -> 2 projections run
// in the epilogue with thread = (point, head): the whole normalisation is register-local.  HBM per point: 256 B in, 16 B out.
//
//   * persistent CTAs, two per SM (84 KB of shared memory, 256 TMEM columns each); 8 epilogue warps + one MMA-issuer warp + one IO warp;
//   * a tile = 128 consecutive rows at one output time; a CTA walks a contiguous range of tiles in the order that is unit-stride
//     in memory (t fastest for rows_major latents, rows fastest for time-major ones);
//   * software pipeline: the operand of tile i+1 (fp32 smem -> fp16 pairs -> TMEM) is staged before the accumulator of tile i is
//     drained, and MMA(i+1) is released as soon as those 64 registers are loaded, so it runs under tile i's LayerNorm arithmetic.
#include "common.cuh"
#include "tc_common.cuh"

namespace trajsde {

using namespace tc;

namespace {

constexpr int TILE_M = 128;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int WARP_MMA = NUM_EPI_WARPS;                 // warp 8: MMA issuer, warp 9: IO (TMA loads)
constexpr int NUM_THREADS = (NUM_EPI_WARPS + 2) * 32;   // 320
constexpr int NBUF = 2;

// packed image: [128][64] fp16 SW128 K-major B operand (rows 0..63 head 0, 64..127 head 1) + fp32 vectors
constexpr uint32_t IMG_W1 = 0;
constexpr uint32_t IMG_WB = 16384;       // [128][64] fp16 SW128, only k = 0, 1 used: centred layer-1 bias as hi + lo fp16 (A operand there is 1, 1)
constexpr uint32_t IMG_VEC = 32768;
// vector slots (float index), per head h: + h * 64 (w2: + h * 128)
constexpr int VEC_G = 128, VEC_BETA = 256, VEC_W2 = 384, VEC_B2 = 640;   // [0,128) unused (the bias rides the MMA); w2: [head][2][64]; b2: [head][2]
constexpr uint32_t IMG_BYTES = IMG_VEC + 648 * 4;        // 35360
constexpr uint32_t OFF_X = 35840;                        // 1024-aligned: NBUF x 32 KB fp32 tiles (two 16 KB boxes of 32 channels)
constexpr uint32_t X_BYTES = 32768;
constexpr uint32_t OFF_BARS = OFF_X + NBUF * X_BYTES;    // w, full[NBUF], empty[NBUF], opnd, acc
constexpr uint32_t SMEM_TOTAL = OFF_BARS + 128;
constexpr uint32_t SMEM_ALLOC = SMEM_TOTAL + 1024;
static_assert(2 * (SMEM_ALLOC + 1024) <= 232448, "two CTAs per SM must fit");

// TMEM columns (256 allocated): [0,128) accumulator (head 0 | head 1), [128,224) two operand buffers of 48 columns: 32 of fp16 pairs
// (K = 64 channels) + 16 whose first column is (1, 1) and the rest 0 (K block of the bias product)
constexpr uint32_t TM_ACC = 0, TM_A = 128, TM_A_STRIDE = 48;

struct HeadsParams {
  TrajsdeHeadsArgs a;
  const uint8_t* img;
  int row_tiles;
  uint32_t num_tiles;    // row_tiles * n_t
  int t_fastest;         // tile index = rt * n_t + t   (else t * row_tiles + rt)
};

__global__ void heads_pack_kernel(TrajsdeHeadsArgs a, uint8_t* __restrict__ img) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  // LayerNorm subtracts the channel mean of z = W1 x + b1, which is linear in x: mean_n(z) = (mean_n W1[n,:]) . x + mean(b1).  The image
  // holds the CENTRED layer W1 - 1 (mean_n W1)^T, b1 - mean(b1), so the accumulator already is z - mean(z) and the epilogue needs
  // neither the sum over channels nor the subtraction (the fp16 rounding of the centred weights leaves a residual mean of order
  // 2^-11 |z|, far below the operand rounding itself).
  for (int idx = tid; idx < 128 * 64; idx += nth) {
    const int n = idx >> 6, k = idx & 63, h = n >> 6;
    float v = 0.f;
    if (h < a.n_heads) {
      const float* w = a.head[h].w1;
      float cm = 0.f;
      for (int m = 0; m < 64; ++m) cm += w[m * 64 + k];
      v = w[(n & 63) * 64 + k] - cm * (1.0f / 64.0f);
    }
    *reinterpret_cast<__half*>(img + IMG_W1 + sw128_off_h(n, k)) = __float2half_rn(v);
    // bias tile: k = 0 holds the fp16 head of the centred bias, k = 1 its fp16 remainder (their sum is exact to 2^-22), k >= 2 zero
    float bv = 0.f;
    if (h < a.n_heads && k < 2) {
      float bm = 0.f;
      for (int m = 0; m < 64; ++m) bm += a.head[h].b1[m];
      const float bc = a.head[h].b1[n & 63] - bm * (1.0f / 64.0f);
      const float hi = __half2float(__float2half_rn(bc));
      bv = k == 0 ? hi : bc - hi;
    }
    *reinterpret_cast<__half*>(img + IMG_WB + sw128_off_h(n, k)) = __float2half_rn(bv);
  }
  float* vec = reinterpret_cast<float*>(img + IMG_VEC);
  for (int i = tid; i < 648; i += nth) {
    float v = 0.f;
    if (i < VEC_W2) {
      const int h = (i >> 6) & 1, c = i & 63;
      if (h < a.n_heads && i >= VEC_G) v = i < VEC_BETA ? a.head[h].ln_g[c] : a.head[h].ln_b[c];
    } else if (i < VEC_B2) {
      const int h = (i - VEC_W2) >> 7, j = (i - VEC_W2) & 127;
      if (h < a.n_heads) v = a.head[h].w2[j];
    } else if (i < VEC_B2 + 4) {
      const int h = (i - VEC_B2) >> 1;
      if (h < a.n_heads) v = a.head[h].b2[(i - VEC_B2) & 1];
    }
    vec[i] = v;
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 2) heads_fwd_kernel(const HeadsParams p, const __grid_constant__ CUtensorMap tm_x) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  const TrajsdeHeadsArgs& a = p.a;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;

  // contiguous, balanced tile range of this CTA
  const uint32_t tq = p.num_tiles / gridDim.x, tr = p.num_tiles % gridDim.x;
  const uint32_t tile_lo = blockIdx.x * tq + min(blockIdx.x, tr);
  const int n_my = (int)(tq + (blockIdx.x < tr ? 1u : 0u));
  const uint32_t fast_n = p.t_fastest ? (uint32_t)a.n_t : (uint32_t)p.row_tiles;
  auto tile_coord = [&](int i, int& rt, int& t) {
    const uint32_t g = tile_lo + (uint32_t)i, hi = g / fast_n, lo = g - hi * fast_n;
    rt = (int)(p.t_fastest ? hi : lo);
    t = (int)(p.t_fastest ? lo : hi);
  };

  const uint32_t bar_w = base + OFF_BARS;
  auto bar_full = [&](int b) { return base + OFF_BARS + 8u + 8u * b; };
  auto bar_empty = [&](int b) { return base + OFF_BARS + 8u + 8u * NBUF + 8u * b; };
  const uint32_t bar_opnd = base + OFF_BARS + 8u + 16u * NBUF, bar_acc = bar_opnd + 8u;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + OFF_BARS + 96);

  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    for (int b = 0; b < NBUF; ++b) {
      mbar_init(bar_full(b), 1);
      mbar_init(bar_empty(b), NUM_EPI_THREADS);
    }
    mbar_init(bar_opnd, NUM_EPI_THREADS);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_x);
  }
  if (warp == WARP_MMA) tmem_alloc(smem_u32(tmem_ptr_smem), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_w, IMG_BYTES);
    bulk_load_1d(base, p.img, IMG_BYTES, bar_w);
  }
  const float* vec = reinterpret_cast<const float*>(sm + IMG_VEC);

  if (warp < NUM_EPI_WARPS) {
    // =============================================== EPILOGUE WARPS ===============================================
    const int quad = warp & 3;                              // TMEM lane quadrant
    const uint32_t hh = (uint32_t)warp >> 2;                // operand staging: 32-channel half of the row; epilogue: head index
    const uint32_t row = quad * 32 + lane;
    const uint32_t tml = tmem_base + ((uint32_t)(quad * 32) << 16);
    const bool head_on = (int)hh < a.n_heads;
    const bool cat4 = (a.flags & TRAJSDE_HEADS_FLAG_CAT4) != 0;   // both heads write one [rows, n_t, 4] result (dec…sde.py:98-100)
    float* outp = head_on ? (cat4 ? a.out[0] + 2 * hh : a.out[hh]) : nullptr;
    const int64_t ostride = cat4 ? 4 : 2;
    uint32_t par_acc = 0;

    // fp32 tile (this thread's 32 channels) -> fp16 pairs -> operand buffer `ab` in tensor memory; releases the smem buffer
    auto stage_operand = [&](int i) {
      const int b = i % NBUF;
      mbar_wait(bar_full(b), (uint32_t)(i / NBUF) & 1u);   // buffer b is filled for the (i / NBUF)-th time
      const uint8_t* xr = sm + OFF_X + b * X_BYTES + hh * 16384 + row * 128;
      uint32_t pk[16];
#pragma unroll
      for (uint32_t q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(xr + ((q ^ (row & 7u)) << 4));
        pk[2 * q] = pack_f16x2(v.x, v.y);
        pk[2 * q + 1] = pack_f16x2(v.z, v.w);
      }
      tmem_st_32x32b_x16(tml + TM_A + (uint32_t)(i & 1) * TM_A_STRIDE + hh * 16, pk);
      mbar_arrive(bar_empty(b));                           // the IO warp may refill this buffer
      tc_wait_st();
    };

    if (hh == 0) {                                           // constant K block of the bias product: (1, 1) then zeros, both buffers
      uint32_t ones[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) ones[j] = 0u;
      ones[0] = 0x3C003C00u;                                 // fp16 pair (1.0, 1.0)
      tmem_st_32x32b_x16(tml + TM_A + 32, ones);
      tmem_st_32x32b_x16(tml + TM_A + TM_A_STRIDE + 32, ones);
      tc_wait_st();
    }
    mbar_wait(bar_w, 0);
    if (n_my > 0) {
      stage_operand(0);
      tc_fence_before();
      mbar_arrive(bar_opnd);
    }
    for (int i = 0; i < n_my; ++i) {
      int rt, t;
      tile_coord(i, rt, t);
      if (i + 1 < n_my) stage_operand(i + 1);
      mbar_wait(bar_acc, par_acc);
      par_acc ^= 1;
      tc_fence_after();
      float v[64];
      {
        uint32_t u0[32], u1[32];
        tmem_ld_32x32b_x32(tml + TM_ACC + hh * 64, u0);
        tmem_ld_32x32b_x32(tml + TM_ACC + hh * 64 + 32, u1);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] = __uint_as_float(u0[j]);
          v[32 + j] = __uint_as_float(u1[j]);
        }
      }
      tc_fence_before();
      if (i + 1 < n_my) mbar_arrive(bar_opnd);             // accumulator drained, next operand staged -> MMA(i+1)
      if (head_on) {
        // the accumulator holds z - mean(z), centred bias included (see heads_pack_kernel): variance = mean of squares; four
        // independent partial sums keep the dependent chains short next to the 4-cycle FMA latency
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          s0 = fmaf(v[j], v[j], s0); s1 = fmaf(v[j + 1], v[j + 1], s1); s2 = fmaf(v[j + 2], v[j + 2], s2); s3 = fmaf(v[j + 3], v[j + 3], s3);
        }
        const float rstd = rsqrtf(((s0 + s1) + (s2 + s3)) * (1.0f / 64.0f) + a.ln_eps);
        const float* gm = vec + VEC_G + hh * 64;
        const float* bt = vec + VEC_BETA + hh * 64;
        const float* w2 = vec + VEC_W2 + hh * 128;
        float o0a = vec[VEC_B2 + hh * 2], o1a = vec[VEC_B2 + hh * 2 + 1], o0b = 0.f, o1b = 0.f;
#pragma unroll
        for (int j = 0; j < 64; j += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(gm + j);
          const float4 b4 = *reinterpret_cast<const float4*>(bt + j);
          const float4 wa = *reinterpret_cast<const float4*>(w2 + j);
          const float4 wb = *reinterpret_cast<const float4*>(w2 + 64 + j);
          const float r0 = fmaxf(fmaf(v[j] * rstd, g4.x, b4.x), 0.f);
          const float r1 = fmaxf(fmaf(v[j + 1] * rstd, g4.y, b4.y), 0.f);
          const float r2 = fmaxf(fmaf(v[j + 2] * rstd, g4.z, b4.z), 0.f);
          const float r3 = fmaxf(fmaf(v[j + 3] * rstd, g4.w, b4.w), 0.f);
          o0a = fmaf(r0, wa.x, o0a); o0b = fmaf(r1, wa.y, o0b); o0a = fmaf(r2, wa.z, o0a); o0b = fmaf(r3, wa.w, o0b);
          o1a = fmaf(r0, wb.x, o1a); o1b = fmaf(r1, wb.y, o1b); o1a = fmaf(r2, wb.z, o1a); o1b = fmaf(r3, wb.w, o1b);
        }
        float o0 = o0a + o0b, o1 = o1a + o1b;
        if (cat4 && hh == 1) {                               // F.elu_(scale) + 1.0, then + min_scale (same order of the fp32 adds)
          o0 = ((o0 > 0.f ? o0 : expf(o0) - 1.f) + 1.0f) + a.min_scale;
          o1 = ((o1 > 0.f ? o1 : expf(o1) - 1.f) + 1.0f) + a.min_scale;
        }
        const int64_t grow = (int64_t)rt * TILE_M + row;
        if (grow < a.rows) *reinterpret_cast<float2*>(outp + (grow * a.n_t + t) * ostride) = make_float2(o0, o1);
      }
    }
  } else if (warp == WARP_MMA) {
    // =============================================== MMA ISSUER =====================================================
    const uint32_t idesc = umma_idesc_f16(TILE_M, 128);
    const uint64_t dhi = umma_desc_sw128(0);
    auto D = [&](uint32_t addr) { return dhi | (uint64_t)((addr & 0x3FFFFu) >> 4); };
    uint32_t par_op = 0;
    mbar_wait(bar_w, 0);
    for (int i = 0; i < n_my; ++i) {
      mbar_wait_suspend(bar_opnd, par_op);                 // MMA(i) was released a whole LayerNorm epilogue ahead of its consumer
      par_op ^= 1;
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          tc_mma_f16_ts(tmem_base + TM_ACC, tmem_base + TM_A + (uint32_t)(i & 1) * TM_A_STRIDE + 8 * kk, D(base + IMG_W1 + 32 * kk), idesc, kk > 0);
        tc_mma_f16_ts(tmem_base + TM_ACC, tmem_base + TM_A + (uint32_t)(i & 1) * TM_A_STRIDE + 32, D(base + IMG_WB), idesc, 1);   // + 1 . b1'^T
        tc_commit(bar_acc);
      }
      __syncwarp();
    }
  } else {
    // =============================================== IO WARP ==========================================================
    if (lane == 0) {
      for (int i = 0; i < n_my; ++i) {
        const int b = i % NBUF;
        if (i >= NBUF) mbar_wait_suspend(bar_empty(b), (uint32_t)(i / NBUF - 1) & 1u);   // its previous content has been staged
        int rt, t;
        tile_coord(i, rt, t);
        const uint32_t dst = base + OFF_X + b * X_BYTES;
        mbar_arrive_expect_tx(bar_full(b), X_BYTES);
        tma_load_3d(dst, &tm_x, bar_full(b), 0, rt * TILE_M, t);
        tma_load_3d(dst + 16384, &tm_x, bar_full(b), 32, rt * TILE_M, t);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

int64_t heads_workspace_bytes() { return (int64_t)IMG_BYTES + 256; }

int launch_heads_fwd(const TrajsdeHeadsArgs& a, cudaStream_t s) {
  int dev = 0, sms = 0;
  TS_CUDA_CHECK(cudaGetDevice(&dev));
  TS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0)
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "workspace must be 256-byte aligned");
  if (a.rows >= (int64_t)1 << 31) return set_error(TRAJSDE_ERR_UNSUPPORTED, "rows >= 2^31 unsupported");
  HeadsParams p;
  p.a = a;
  p.img = static_cast<const uint8_t*>(a.workspace);
  p.row_tiles = (int)((a.rows + TILE_M - 1) / TILE_M);
  if ((int64_t)p.row_tiles * a.n_t >= (int64_t)1 << 31) return set_error(TRAJSDE_ERR_UNSUPPORTED, "rows x n_t too large");
  p.num_tiles = (uint32_t)p.row_tiles * (uint32_t)a.n_t;
  p.t_fastest = a.x_t_stride < a.x_row_stride ? 1 : 0;
  if (p.num_tiles == 0) return TRAJSDE_OK;

  heads_pack_kernel<<<16, 256, 0, s>>>(a, static_cast<uint8_t*>(a.workspace));
  TS_CUDA_CHECK(cudaGetLastError());
  CUtensorMap tm_x;
  int rc;
  if ((rc = tc_make_map(&tm_x, a.x, a.rows, a.n_t, a.x_row_stride, a.x_t_stride)) != 0) return rc;
  TS_CUDA_CHECK(cudaFuncSetAttribute(heads_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
  const uint32_t max_ctas = 2u * (uint32_t)sms;
  const int grid = (int)(p.num_tiles < max_ctas ? p.num_tiles : max_ctas);
  heads_fwd_kernel<<<grid, NUM_THREADS, SMEM_ALLOC, s>>>(p, tm_x);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

}  // namespace trajsde
