// Device body of the tensor-core Euler-Maruyama backward (see euler_bwd_tc.cu for the description).  A header so that the single-launch
// encoder-recurrence sweep (enc_bwd_sweep.cu) can run it as one ROLE of a merged kernel next to the GRU backward.
#pragma once
#include "bwd_common.cuh"
#include "bwd_tc_common.cuh"

namespace trajsde {
namespace sdetc {

using namespace tc;
using namespace bwd;
using namespace bwdtc;

constexpr int TILE_M = 128;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = NUM_EPI_THREADS + 128;  // + one warpgroup: MMA-issuer warp and three idle warps (setmaxnreg is per warpgroup)
constexpr int EPI_REGS = 216, AUX_REGS = 64;        // 256 x 216 + 128 x 64 = 63488 <= 64 K registers

// ---- packed weight image (bytes): fp16 SW128 K-major B operands + fp32 vectors ------------------------------------------------
constexpr uint32_t IMG_B1 = 0;                        // [128][64]: rows 0..63 W1y, 64..127 V1y     (P1, N = 128)
constexpr uint32_t IMG_W2 = 16384, IMG_V2 = 24576;    // forward orientation                       (P2)
constexpr uint32_t IMG_W3T = 32768, IMG_W2T = 40960, IMG_V2T = 49152, IMG_W1YT = 57344, IMG_V1YT = 65536;   // B[n][k] = W[k][n]
constexpr uint32_t IMG_VEC = 73728;
constexpr int VEC_B1 = 0, VEC_W1S = 64, VEC_W1C = 128, VEC_B2 = 192, VEC_C1 = 256, VEC_C2 = 448,   // [320,448): diffusion time columns, addressed as VEC_C1 + VEC_W1S / VEC_W1C
              VEC_W3G = 512, VEC_C3 = 576;
constexpr uint32_t IMG_BYTES = IMG_VEC + 640 * 4;     // 76288
static_assert(IMG_BYTES <= BWD_TC_IMG_BYTES, "image larger than its workspace slot");

// ---- shared memory map ------------------------------------------------------------------------------------------------------
constexpr uint32_t OFF_TILES = 76800;                 // eight [128 rows][64] fp16 SW128 tiles, 16 KB each
constexpr uint32_t TILE_BYTES = 16384;
// tile roles (adjacency matters: M=128 / N=128 MN-major stacks are two consecutive tiles)
constexpr int T_H2F = 0;      // h2f, later dz1f          [dz1f|dz1g] = tiles 0,1
constexpr int T_DF = 1;       // df,  later dz1g          [df|y]      = tiles 1,2
constexpr int T_Y = 2;
constexpr int T_DZ2F = 3;     //                          [dz2f|dz2g] = tiles 3,4
constexpr int T_DZ2G = 4;
constexpr int T_H1F = 5;      //                          [h1f|h1g]   = tiles 5,6
constexpr int T_H1G = 6;
constexpr int T_TIME = 7;     // column 0 = 1, 1 = sin t_k, 2 = cos t_k, rest 0
constexpr uint32_t OFF_XCHG = OFF_TILES + 8 * TILE_BYTES;          // q[2][128], pd[2][128] fp32
constexpr int SCHED_MAX = 128;                        // schedule tables staged in shared memory when they fit (else read from global)
constexpr uint32_t OFF_STAB = OFF_XCHG + 4 * TILE_M * 4;            // float4 step_tab[SCHED_MAX]
constexpr uint32_t OFF_OBEG = OFF_STAB + SCHED_MAX * 16;           // int out_begin[SCHED_MAX + 4]
constexpr uint32_t OFF_OUTW = OFF_OBEG + (SCHED_MAX + 4) * 4;      // float2 out_w[SCHED_MAX]
constexpr uint32_t OFF_BROW = OFF_OUTW + SCHED_MAX * 8;            // float bias1[128]: layer-1 bias rows (f | g) of the current step
constexpr uint32_t OFF_DWT = OFF_BROW + 128 * 4;      // Philox variant: [128 rows][64] fp16 increments of the step, drawn by the aux warps
constexpr uint32_t OFF_BARS = OFF_DWT + TILE_BYTES;   // w, opnd, acc, wg, dwfull, dwempty
constexpr uint32_t SMEM_TOTAL = OFF_BARS + 64;
constexpr int NUM_DW_WARPS = 3, NUM_DW_THREADS = NUM_DW_WARPS * 32;   // the three warps of the issuer's warpgroup that were idle
constexpr uint32_t SMEM_ALLOC = SMEM_TOTAL + 1024;
static_assert(SMEM_ALLOC <= 232448, "exceeds 227 KB of shared memory per CTA");

// ---- TMEM columns -------------------------------------------------------------------------------------------------------------
constexpr uint32_t TM_R0 = 0, TM_R1 = 64, TM_E = 128;
constexpr uint32_t TM_WGA = 192;    // 128 cols: lanes 0..63 x [0,64) = dW2, lanes 64..127 x [64,128) = dV2
constexpr uint32_t TM_WGB = 320;    // 64 cols: lanes 0..63 dW1y, 64..127 dV1y
constexpr uint32_t TM_WGC = 384;    // 64 cols: lanes 0..63 dW3
constexpr uint32_t TM_SUM1 = 448, TM_SUM2 = 464, TM_SUM3 = 480;   // 16 cols each: col 0 = column sums (, 1 = x sin, 2 = x cos)

struct BwdTcParams {
  TrajsdeEulerBwdArgs a;
  // blockIdx.y selects the pass: dual diffusion runs both nets' passes in one launch (rows are independent)
  const uint8_t* img[2];       // packed weight image of the pass (drift + that pass's diffusion net)
  float* partial[2];           // [gridDim.x][G_PAD] of the pass
  int filter[2];               // 0: all rows; 1: only rows with alt_mask != 0; 2: only rows with alt_mask == 0
  const uint32_t* amax_bits;   // max |grad| as float bits (absmax pre-pass)
  // zero-row skipping (TRAJSDE_BWD_FLAG_SKIP_ZERO_ROWS): rows whose incoming gradients are all zero have a zero adjoint at every step and
  // contribute exactly nothing to any result, so the sweep runs over the compacted list row_map[0 .. *n_active) of the other rows
  const int32_t* row_map;      // device [rows] or NULL (identity)
  const int32_t* n_active;     // device scalar or NULL (all rows)
  int num_tiles;
  int accumulate;              // add into the CTA's partial vector instead of overwriting it (multi-launch accumulation)
};

#ifdef TRAJSDE_BWD_TIMELINE
static __device__ long long g_bwd_tl[16];
#define TL_MARK(i) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long _t = clock64(); g_bwd_tl[i] += _t - tl_prev; tl_prev = _t; } } while (0)
#else
#define TL_MARK(i) do { } while (0)
#endif

// SWEEP: the SDE-step role of the single-launch encoder sweep (enc_bwd_sweep.cu) — p.a describes ONE Euler step (n_steps = n_outputs = 1);
// the CTA runs it for every iteration i = S-1 .. 0 of the recurrence over its tiles, with the per-iteration pieces derived here: step-table
// row i of a.sched.step_tab (sw.S rows), state h0 / latent[i-1], increments dw + i slabs or Philox step step_offset + i, grad_g_last + i rows,
// result into the carry buffer (grad_h0 for i = 0).  dL/dy1 (a.grad_ys slab 1) is read once the GRU role has published the tile.
// (cta, ncta): this CTA's index among the CTAs of its pass.
template <bool HAS_DW, bool SWEEP>
__device__ __forceinline__ void euler_bwd_tc_body(const BwdTcParams& p, const SweepCtl& sw, const int cta, const int ncta, const int pass,
                                                  uint8_t* smem_raw) {
#ifdef TRAJSDE_BWD_TIMELINE
  long long tl_prev = clock64();
#endif
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);

  const TrajsdeEulerBwdArgs& a = p.a;
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int S = a.sched.n_steps;
  const int it_hi = SWEEP ? sw.S - 1 : 0, it_lo = 0;

  const uint32_t bar_w = base + OFF_BARS, bar_opnd = bar_w + 8, bar_acc = bar_w + 16, bar_wg = bar_w + 24;
  const uint32_t bar_dwfull = bar_w + 32, bar_dwempty = bar_w + 40;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sm + OFF_BARS + 48);
  auto tile_u32 = [&](int t) { return base + OFF_TILES + (uint32_t)t * TILE_BYTES; };

  pdl_launch_dependents();
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_opnd, NUM_EPI_THREADS);
    mbar_init(bar_acc, 1);
    mbar_init(bar_wg, 1);
    mbar_init(bar_dwfull, NUM_DW_THREADS);
    mbar_init(bar_dwempty, NUM_EPI_THREADS);
    mbar_fence_init();
  }
  if (warp == NUM_EPI_WARPS) tmem_alloc(smem_u32(tmem_ptr_smem), 512);
  // time tile: zero once (only chunk 0 of every row is rewritten per step)
  for (uint32_t i = threadIdx.x; i < TILE_BYTES / 16 && threadIdx.x < NUM_EPI_THREADS; i += NUM_EPI_THREADS)   // epilogue threads: they fence.proxy.async later
    reinterpret_cast<uint4*>(sm + OFF_TILES + T_TIME * TILE_BYTES)[i] = make_uint4(0u, 0u, 0u, 0u);
  pdl_wait();   // nothing above touches global memory; everything below may depend on the previous kernel of the stream
  // rows of this sweep: all of them, or the compacted list of rows with a non-zero incoming gradient (count known on the device only)
  const int32_t* __restrict__ rmap = p.row_map;
  const int64_t n_rows = p.n_active ? (int64_t)*p.n_active : a.rows;
  const uint64_t noise_seed = HAS_DW ? 0ull : ts_noise_seed(a.noise);
  const int num_tiles = (int)((n_rows + TILE_M - 1) / TILE_M);
  const int tiles_q = num_tiles / ncta, tiles_r = num_tiles % ncta;
  const int tile_lo = cta * tiles_q + min(cta, tiles_r);
  const int tile_hi = tile_lo + tiles_q + (cta < tiles_r ? 1 : 0);
  // schedule tables -> shared memory (every step reads them; three dependent global round trips otherwise)
  const int n_tab = SWEEP ? sw.S : S;                          // step-table rows (the sweep indexes them by iteration)
  const bool sched_in_smem = n_tab <= SCHED_MAX && a.sched.n_outputs <= SCHED_MAX;
  const float4* stab = reinterpret_cast<const float4*>(a.sched.step_tab);
  const int* obeg = a.sched.out_begin;
  const float2* outw = reinterpret_cast<const float2*>(a.sched.out_w);
  if (sched_in_smem) {
    float4* s_stab = reinterpret_cast<float4*>(sm + OFF_STAB);
    int* s_obeg = reinterpret_cast<int*>(sm + OFF_OBEG);
    float2* s_outw = reinterpret_cast<float2*>(sm + OFF_OUTW);
    for (int i = threadIdx.x; i < n_tab; i += NUM_THREADS) s_stab[i] = stab[i];
    for (int i = threadIdx.x; i <= S; i += NUM_THREADS) s_obeg[i] = obeg[i];
    for (int i = threadIdx.x; i < a.sched.n_outputs; i += NUM_THREADS) s_outw[i] = outw[i];
    stab = s_stab;
    obeg = s_obeg;
    outw = s_outw;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_w, IMG_BYTES);
    bulk_load_1d(base, p.img[pass], IMG_BYTES, bar_w);
  }
  const float* vec = reinterpret_cast<const float*>(sm + IMG_VEC);

  if (warp < NUM_EPI_WARPS) {
    // =============================================== EPILOGUE WARPS ===============================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    const int quad = warp & 3;
    const uint32_t hh = (uint32_t)warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t tm = tmem_base + ((uint32_t)(quad * 32) << 16) + hh * 32;   // this thread's lane / 32-column half
    auto trow = [&](int t) { return sm + OFF_TILES + (uint32_t)t * TILE_BYTES + row * 128; };
    float* qbuf = reinterpret_cast<float*>(sm + OFF_XCHG);
    float* pdbuf = qbuf + 2 * TILE_M;
    float* brow = reinterpret_cast<float*>(sm + OFF_BROW);
    const int eid = warp * 32 + lane;                      // 0..255
    const uint32_t pair_bar = 1 + quad;

    // adjoint scale: a power of two that puts max|grad| into [2^-4, 2^-3)
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
    float dw3g_acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) dw3g_acc[j] = 0.f;
    float dc3_acc = 0.f;
    float adj_peak = 0.f;   // largest |scaled adjoint| this thread has carried (range check of the fp16 delta operands)
    uint32_t hs = 0;        // hand-shake counter: acc barrier parity
    uint32_t gstep = 0;     // steps processed by this CTA: wg barrier parity

    mbar_wait(bar_w, 0);
    const float c3 = vec[VEC_C3];
    TL_MARK(13);  // kernel prologue: barriers, TMEM, tables, weight image

    for (int it = it_hi; it >= it_lo; --it)
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      const float* states_p = SWEEP ? (it == 0 ? sw.h0 : sw.latent + (int64_t)(it - 1) * sw.slab) : a.states;
      const float* dw_p = SWEEP && HAS_DW ? a.noise.dw + (int64_t)it * sw.slab : a.noise.dw;
      const float* gg_p = SWEEP && a.grad_g_last ? a.grad_g_last + (int64_t)it * a.rows : a.grad_g_last;
      float* gy0_p = SWEEP ? (it == 0 ? sw.grad_h0 : sw.carry) : a.grad_y0;
      const int64_t srow = (int64_t)tile * TILE_M + row;       // position in the (possibly compacted) row list
      bool valid = srow < n_rows;
      const int64_t grow = valid && rmap ? (int64_t)rmap[srow] : srow;   // row of the tensors
      if (valid && p.filter[pass]) valid = (a.alt_mask[grow] != 0) == (p.filter[pass] == 1);   // other net's rows: adjoint stays zero
      // Row prefetch, coalesced: lane L of this warp loads, for i = 0..7, the 16-byte chunk (L & 7) of tile row 32 quad + 4 i + (L >> 3)
      // of its 32-channel half (one instruction = four full 128-byte row segments instead of 32 scattered 16-byte pieces); the
      // registers are transposed to "thread = own row" through 4 KB of per-warp staging when the step consumes them.
      float4 py[8], pdw[8], pgy[8];
      const int64_t lrow0 = (int64_t)tile * TILE_M + quad * 32 + (lane >> 3);     // + 4 i
      const int lcol = hh * 32 + (lane & 7) * 4;
      int32_t lmap[8];                                         // tensor rows behind the eight list positions this lane fetches (-1: none)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = lrow0 + 4 * i;
        lmap[i] = r < n_rows ? (rmap ? rmap[r] : (int32_t)r) : -1;
      }
      auto load_rows = [&](const float* slab, int64_t row_stride, float4 (&dst)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = lmap[i] >= 0 ? ld_nc_f4(slab + (int64_t)lmap[i] * row_stride + lcol) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      auto load_rows_cg = [&](const float* slab, int64_t row_stride, float4 (&dst)[8]) {   // rows written by another CTA of this launch
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = lmap[i] >= 0 ? ld_cg_f4(slab + (int64_t)lmap[i] * row_stride + lcol) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      uint8_t* stage = sm + OFF_TILES + T_H1F * TILE_BYTES + (uint32_t)warp * 4096;   // h1f|h1g tiles are idle at step start
      auto to_own_row = [&](float4 (&v)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t rl = 4 * i + (lane >> 3);
          *reinterpret_cast<float4*>(stage + rl * 128 + ((((uint32_t)lane & 7u) ^ (rl & 7u)) << 4)) = v[i];
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(stage + lane * 128 + (((uint32_t)q ^ ((uint32_t)lane & 7u)) << 4));
        __syncwarp();
      };
      auto prefetch_y_dw = [&](int k) {
        load_rows(states_p + (int64_t)k * a.rows * 64, 64, py);
        if (HAS_DW) load_rows(dw_p + (int64_t)k * a.rows * 64, 64, pdw);
      };
      auto prefetch_gy = [&](int k) {
        const int ob = obeg[k], oe = obeg[k + 1];
        if (a.grad_ys && oe > ob) {
          if (SWEEP) load_rows_cg(a.grad_ys + (int64_t)(ob + 1) * a.grad_ys_t_stride, a.grad_ys_row_stride, pgy);
          else load_rows(a.grad_ys + (int64_t)(ob + 1) * a.grad_ys_t_stride, a.grad_ys_row_stride, pgy);
        }
      };
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        py[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        pdw[q] = py[q];
        pgy[q] = py[q];
      }
      prefetch_y_dw(S - 1);
      if (SWEEP) sweep_wait(sw, sw.gru_done + tile, sw.S - it);   // the GRU role has written dL/dy1 of this tile for iteration `it`
      prefetch_gy(S - 1);
      float adj[32];          // A = dL/dY[k+1] before the output terms, scaled by sigma
#pragma unroll
      for (int j = 0; j < 32; ++j) adj[j] = 0.f;

      for (int k = S - 1; k >= 0; --k, ++gstep) {
        const float4 stp = stab[SWEEP ? it : k];
        const float h = stp.y, sn = stp.z, cs = stp.w;
        const int ob = obeg[k], oe = obeg[k + 1];

        TL_MARK(0);   // (loop overhead / previous e5 tail)
        // ================= step start: y -> operand (P1 starts) ; A' = A + sum w1 gy ; E = A' + sum w0 gy ; df ; q = A'.dW =================
        float e_[32];
        {
          to_own_row(py);
          if (HAS_DW) to_own_row(pdw);
          if (a.grad_ys && oe > ob) to_own_row(pgy);
          if (!valid) {                                            // padding / other net's rows: state and gradient read as zero
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              py[q] = make_float4(0.f, 0.f, 0.f, 0.f);
              pgy[q] = py[q];
            }
          }
          TL_MARK(1);   // SS part 1: transposes
          // previous step's trailing weight-gradient MMAs must have finished reading Y / DF(dz1g) / TIME
          if (gstep > 0) mbar_wait(bar_wg, (gstep - 1) & 1);
          TL_MARK(2);   // wait bar_wg
          // P1 needs only y (and the bias rows): hand it over first, so that its hand-shake and MMAs run under the adjoint update below
          float t[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            t[4 * q] = py[q].x; t[4 * q + 1] = py[q].y; t[4 * q + 2] = py[q].z; t[4 * q + 3] = py[q].w;
          }
          st_row32(trow(T_Y), row, hh, t);
          if (eid < 128) {                                       // layer-1 bias rows of this step (time features folded in)
            const int c = eid & 63, o = eid < 64 ? 0 : VEC_C1;
            brow[eid] = fmaf(vec[o + VEC_W1C + c], cs, fmaf(vec[o + VEC_W1S + c], sn, vec[o + VEC_B1 + c]));
          }
          if (hh == 0)
            *reinterpret_cast<uint4*>(trow(T_TIME) + ((0u ^ (row & 7u)) << 4)) = make_uint4(pack_f16x2(1.f, sn), pack_f16x2(cs, 0.f), 0u, 0u);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_opnd);                                   // y, time tile -> P1
        {
#pragma unroll
          for (int j = 0; j < 32; ++j) e_[j] = adj[j];
          if (a.grad_ys && oe > ob) {
            const float w0 = outw[ob].x, w1 = outw[ob].y;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float g4[4] = {pgy[q].x * sigma, pgy[q].y * sigma, pgy[q].z * sigma, pgy[q].w * sigma};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                adj[4 * q + e] = fmaf(w1, g4[e], adj[4 * q + e]);
                e_[4 * q + e] = fmaf(w1 + w0, g4[e], e_[4 * q + e]);
              }
            }
            for (int o = ob + 1; o < oe; ++o) {   // rare: several outputs interpolate inside the same step
              const float v0 = outw[o].x, v1 = outw[o].y;
              if (valid) {
                const float* gs = a.grad_ys + (int64_t)(o + 1) * a.grad_ys_t_stride + grow * a.grad_ys_row_stride + hh * 32;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 g = ld_nc_f4(gs + 4 * q);
                  const float g4[4] = {g.x * sigma, g.y * sigma, g.z * sigma, g.w * sigma};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    adj[4 * q + e] = fmaf(v1, g4[e], adj[4 * q + e]);
                    e_[4 * q + e] = fmaf(v1 + v0, g4[e], e_[4 * q + e]);
                  }
                }
              }
            }
          }
          // q = A' . dW (this thread's half)
          float qp = 0.f;
          if (HAS_DW) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              qp = fmaf(adj[4 * q], pdw[q].x, qp);
              qp = fmaf(adj[4 * q + 1], pdw[q].y, qp);
              qp = fmaf(adj[4 * q + 2], pdw[q].z, qp);
              qp = fmaf(adj[4 * q + 3], pdw[q].w, qp);
            }
          } else {
            // the increments of this step, drawn again by the aux warps while the previous step ran (fp16: q is a 64-term dot product
            // that feeds fp16 delta operands anyway)
            mbar_wait(bar_dwfull, gstep & 1);
            float t[32];
            ld_row32(sm + OFF_DWT + row * 128, row, hh, t);
            mbar_arrive(bar_dwempty);
#pragma unroll
            for (int j = 0; j < 32; ++j) qp = fmaf(adj[j], t[j], qp);
          }
          qbuf[hh * TILE_M + row] = valid ? qp : 0.f;
          // E -> TMEM (read back by the last epilogue of this step)
          {
            uint32_t ev[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) ev[j] = __float_as_uint(e_[j]);
            tmem_st_32x32b_x32(tm + TM_E, ev);
          }
          // df -> operand tile (read by D1 and its trailing dW3 product; made visible by the fence of epilogue 1)
          float t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = h * adj[j];
          st_row32(trow(T_DF), row, hh, t);
          tc_wait_st();
        }
        TL_MARK(3);   // SS part 2 + fence + arrive + prefetch issue
        // ================= epilogue 1: h1f, h1g ==================================================================================
        mbar_wait(bar_acc, hs & 1);
        ++hs;
        TL_MARK(4);   // wait P1
        tc_fence_after();
        {
          uint32_t v[32];
          float t[32];
          tmem_ld_32x32b_x32(tm + TM_R0, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = ts_tanh_approx(__uint_as_float(v[j]) + brow[hh * 32 + j]);
          st_row32(trow(T_H1F), row, hh, t);
          tmem_ld_32x32b_x32(tm + TM_R1, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = ts_tanh_approx(__uint_as_float(v[j]) + brow[64 + hh * 32 + j]);
          st_row32(trow(T_H1G), row, hh, t);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_opnd);                                   // h1f, h1g -> P2
        // next step's rows.  Every fence.proxy.async is a MEMBAR.ALL.CTA that waits for outstanding loads, so they are issued where
        // the distance to the next fence is longest: the P2 hand-shake plus the 128 tanh of epilogue 2 (~2.3 k clk) lie ahead here,
        // against ~1 k clk when they were issued at the step start
        if (k > 0) {
          prefetch_y_dw(k - 1);
          prefetch_gy(k - 1);
        }

        TL_MARK(5);   // e1
        // ================= epilogue 2: h2f ; h2g, g, ds, dz2g =====================================================================
        mbar_wait(bar_acc, hs & 1);
        ++hs;
        TL_MARK(6);   // wait P2
        tc_fence_after();
        {
          uint32_t v[32];
          float t[32];
          tmem_ld_32x32b_x32(tm + TM_R0, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = ts_tanh_approx(__uint_as_float(v[j]) + vec[VEC_B2 + hh * 32 + j]);
          st_row32(trow(T_H2F), row, hh, t);
          tmem_ld_32x32b_x32(tm + TM_R1, v);
          tc_wait_ld();
          float pd = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            t[j] = ts_tanh_approx(__uint_as_float(v[j]) + vec[VEC_C2 + hh * 32 + j]);
            pd = fmaf(t[j], vec[VEC_W3G + hh * 32 + j], pd);
          }
          pdbuf[hh * TILE_M + row] = pd;
          named_bar_sync(pair_bar, 64);                          // partner half's pd and q are in smem
          const float s = (pdbuf[row] + pdbuf[TILE_M + row]) + c3;
          const float g = __fdividef(1.0f, 1.0f + __expf(-s));
          float dg = qbuf[row] + qbuf[TILE_M + row];
          if (k == S - 1 && gg_p && valid) dg = fmaf(gg_p[grow], sigma, dg);
          const float ds = dg * g * (1.f - g);
          if (hh == 0) dc3_acc += ds;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            dw3g_acc[j] = fmaf(ds, t[j], dw3g_acc[j]);
            t[j] = ds * vec[VEC_W3G + hh * 32 + j] * fmaf(-t[j], t[j], 1.f);
          }
          st_row32(trow(T_DZ2G), row, hh, t);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_opnd);                                   // h2f, dz2g -> D1 (+ trailing dW3 / db3)

        TL_MARK(7);   // e2
        // ================= epilogue 3: dz2f = dh2f (1 - h2f^2) =====================================================================
        mbar_wait(bar_acc, hs & 1);
        ++hs;
        TL_MARK(8);   // wait D1
        tc_fence_after();
        {
          uint32_t v[32];
          float t[32];
          tmem_ld_32x32b_x32(tm + TM_R0, v);
          ld_row32(trow(T_H2F), row, hh, t);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]) * fmaf(-t[j], t[j], 1.f);
          st_row32(trow(T_DZ2F), row, hh, t);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_opnd);                                   // dz2f -> D2 (+ trailing dW2|dV2 / db2|dc2)

        TL_MARK(9);   // e3
        // ================= epilogue 4: dz1f = dh1f (1 - h1f^2) -> tile H2F ; dz1g = dh1g (1 - h1g^2) -> tile DF ==========================
        mbar_wait(bar_acc, hs & 1);
        ++hs;
        TL_MARK(10);  // wait D2
        tc_fence_after();
        {
          uint32_t v[32];
          float t[32];
          tmem_ld_32x32b_x32(tm + TM_R0, v);
          ld_row32(trow(T_H1F), row, hh, t);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]) * fmaf(-t[j], t[j], 1.f);
          st_row32(trow(T_H2F), row, hh, t);
          tmem_ld_32x32b_x32(tm + TM_R1, v);
          ld_row32(trow(T_H1G), row, hh, t);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]) * fmaf(-t[j], t[j], 1.f);
          st_row32(trow(T_DF), row, hh, t);
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_opnd);                                   // dz1f, dz1g -> D3 (+ trailing dW1y|dV1y / db1|dc1 / time columns)

        TL_MARK(11);  // e4
        // ================= epilogue 5: A[k] = E + dy =================================================================================
        mbar_wait(bar_acc, hs & 1);
        ++hs;
        TL_MARK(12);  // wait D3
        tc_fence_after();
        {
          uint32_t v[32], ev[32];
          tmem_ld_32x32b_x32(tm + TM_R0, v);
          tmem_ld_32x32b_x32(tm + TM_E, ev);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            adj[j] = valid ? __uint_as_float(ev[j]) + __uint_as_float(v[j]) : 0.f;
            adj_peak = fmaxf(adj_peak, fabsf(adj[j]));
          }
        }
        tc_fence_before();
      }
      // ---- grad_y0 = A[0] / sigma (+ grad_ys[0]: ys[0] = y0) -----------------------------------------------------------------
      if (valid) {
        float* dst = gy0_p + grow * 64 + hh * 32;
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          float4 o[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            o[e] = make_float4(adj[4 * (q + e)] * inv_sigma, adj[4 * (q + e) + 1] * inv_sigma, adj[4 * (q + e) + 2] * inv_sigma, adj[4 * (q + e) + 3] * inv_sigma);
            if (a.grad_ys) {
              const float4 g = ld_nc_f4(a.grad_ys + grow * a.grad_ys_row_stride + hh * 32 + 4 * (q + e));
              o[e].x += g.x; o[e].y += g.y; o[e].z += g.z; o[e].w += g.w;
            }
          }
          st_f8(dst + 4 * q, o[0], o[1]);                       // 32-byte sectors (grad_y0 / the carry buffer are 256-byte aligned rows)
        }
      }
      if (SWEEP) sweep_publish(sw.sde_done[pass] + tile, sw.S - it, 6, NUM_EPI_THREADS, threadIdx.x == 0);   // carried adjoint of the tile is out
    }

    TL_MARK(0);   // last step's tail + grad_y0 stores
    if (a.status && !(adj_peak <= 16384.f)) atomicOr(a.status, TRAJSDE_STATUS_ADJOINT_RANGE);   // also catches NaN
    // ================= weight-gradient partials of this CTA ===============================================================================
    if (gstep > 0) mbar_wait(bar_wg, (gstep - 1) & 1);          // every MMA of the CTA has completed
    tc_fence_after();
    float* out = p.partial[pass] + (size_t)cta * G_PAD;
    const bool acc_out = p.accumulate != 0;
    auto put = [&](int idx, float v) { out[idx] = acc_out ? out[idx] + v : v; };
    if (tile_lo == tile_hi) {                                  // no tile for this CTA (the tile count is a device value): a zero partial vector
      if (!acc_out)
        for (int i = eid; i < G_PAD; i += NUM_EPI_THREADS) out[i] = 0.f;
    } else {
    const bool lo = quad < 2;                                    // TMEM lanes 0..63: drift net, 64..127: diffusion net
    const int m = (int)row & 63;
    {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tm + TM_WGA + (lo ? 0u : 64u), v);      // dW2 / dV2
      tc_wait_ld();
      float* d = out + (lo ? G_FW2 : G_GW2) + m * 64 + hh * 32;
      flush_row32(d, v, inv_sigma, acc_out);
      tmem_ld_32x32b_x32(tm + TM_WGB, v);                        // dW1y / dV1y
      tc_wait_ld();
      d = out + (lo ? G_FW1 : G_GW1) + m * TS_IN1 + hh * 32;
      flush_row32(d, v, inv_sigma, acc_out);
      tmem_ld_32x32b_x32(tm + TM_WGC, v);                        // dW3 (lanes 0..63)
      tc_wait_ld();
      if (lo) {
        d = out + G_FW3 + m * 64 + hh * 32;
        flush_row32(d, v, inv_sigma, acc_out);
      }
      const uint32_t tm0 = tmem_base + ((uint32_t)(quad * 32) << 16);
      uint32_t s1[16], s2[16], s3[16];
      tmem_ld_32x32b_x16(tm0 + TM_SUM1, s1);
      tmem_ld_32x32b_x16(tm0 + TM_SUM2, s2);
      tmem_ld_32x32b_x16(tm0 + TM_SUM3, s3);
      tc_wait_ld();
      if (hh == 0) {
        put((lo ? G_FB1 : G_GB1) + m, __uint_as_float(s1[0]) * inv_sigma);
        put((lo ? G_FW1 : G_GW1) + m * TS_IN1 + 64, __uint_as_float(s1[1]) * inv_sigma);
        put((lo ? G_FW1 : G_GW1) + m * TS_IN1 + 65, __uint_as_float(s1[2]) * inv_sigma);
        put((lo ? G_FB2 : G_GB2) + m, __uint_as_float(s2[0]) * inv_sigma);
        if (lo) put(G_FB3 + m, __uint_as_float(s3[0]) * inv_sigma);
      }
    }
    // w3 / c3 of the diffusion net: per-thread running sums -> column sums over the warp's 32 rows (butterfly), then over the four
    // row quadrants through 1 KB of shared memory
    {
      float* red = reinterpret_cast<float*>(sm + OFF_TILES);     // [8 warps][32] + [8] fp32 scratch (operand tiles are idle now)
      red[warp * 32 + lane] = colsum32(dw3g_acc, lane);          // column hh*32 + lane, rows of this warp
      float cs = dc3_acc;                                        // hh == 0 threads carry the c3 sums, the others 0
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) cs += __shfl_xor_sync(0xffffffffu, cs, off);
      if (lane == 0) red[256 + warp] = cs;
      named_bar_sync(5, NUM_EPI_THREADS);
      if (threadIdx.x < 64) {
        const int h2 = threadIdx.x >> 5, l2 = threadIdx.x & 31;
        const float s4 = (red[(h2 * 4 + 0) * 32 + l2] + red[(h2 * 4 + 1) * 32 + l2]) + (red[(h2 * 4 + 2) * 32 + l2] + red[(h2 * 4 + 3) * 32 + l2]);
        put(G_GW3 + threadIdx.x, s4 * inv_sigma);
      } else if (threadIdx.x == 64) {
        put(G_GB3, ((red[256] + red[257]) + (red[258] + red[259])) * inv_sigma);
      }
    }
    }
    TL_MARK(14);  // weight-gradient flush
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
    if (!HAS_DW && warp > NUM_EPI_WARPS) {
      // =============================================== INCREMENT PRODUCERS (Philox variant) =========================================
      // The three warps next to the MMA issuer regenerate the forward's Brownian increments of step k into the fp16 tile OFF_DWT while
      // the epilogue warps are busy with step k+1; the epilogue reads its row for q = A'.dW and hands the tile back.  Keyed by (global
      // row, step, channel / 4) like every other draw of the stream.
      const uint32_t t = (uint32_t)(warp - NUM_EPI_WARPS - 1) * 32u + (uint32_t)lane;      // 0..95
      uint32_t n = 0;                                                                    // tiles-steps produced: dwempty parity
      for (int it = it_hi; it >= it_lo; --it)
      for (int tile = tile_lo; tile < tile_hi; ++tile) {
        const int64_t srow0 = (int64_t)tile * TILE_M;
        const uint32_t soff = a.noise.step_offset + (SWEEP ? (uint32_t)it : 0u);
        for (int k = S - 1; k >= 0; --k, ++n) {
          const float sqrt_h = sqrtf(stab[SWEEP ? it : k].y);
          if (n > 0) mbar_wait(bar_dwempty, (n - 1) & 1);                                // every epilogue thread has read the previous tile
          for (uint32_t item = t; item < 2u * TILE_M; item += NUM_DW_THREADS) {
            const uint32_t r = item & (TILE_M - 1), h2 = item >> 7;
            uint8_t* tr = sm + OFF_DWT + r * 128;
            const int64_t sr = srow0 + r;
            const uint64_t grow_r = (uint64_t)(rmap && sr < n_rows ? (int64_t)rmap[sr] : sr) + a.noise.row_offset;
#pragma unroll 1
            for (uint32_t c = 0; c < 4; ++c) {                                           // 16-byte chunk = 8 channels = two Philox calls
              const float4 n0 = philox_dw4(noise_seed, grow_r, soff + (uint32_t)k, h2 * 8 + 2 * c, sqrt_h);
              const float4 n1 = philox_dw4(noise_seed, grow_r, soff + (uint32_t)k, h2 * 8 + 2 * c + 1, sqrt_h);
              *reinterpret_cast<uint4*>(tr + (((h2 * 4 + c) ^ (r & 7u)) << 4)) =
                  make_uint4(pack_f16x2(n0.x, n0.y), pack_f16x2(n0.z, n0.w), pack_f16x2(n1.x, n1.y), pack_f16x2(n1.z, n1.w));
            }
          }
          mbar_arrive(bar_dwfull);
        }
      }
    }
    if (warp == NUM_EPI_WARPS) {
    // =============================================== MMA ISSUER WARP ===============================================
    // warp-uniform loop (descriptors in uniform registers); one elected lane issues tcgen05.mma / tcgen05.commit
    const uint32_t idesc_128 = umma_idesc_f16(TILE_M, 128), idesc_64 = umma_idesc_f16(TILE_M, 64);
    const uint32_t imn_128 = umma_idesc_f16_mn(TILE_M, 128), imn_64 = umma_idesc_f16_mn(TILE_M, 64), imn_16 = umma_idesc_f16_mn(TILE_M, 16);
    const uint64_t khi = umma_desc_sw128(0), mhi = umma_desc_mn_sw128(0, TILE_BYTES);
    auto KD = [&](uint32_t addr) { return khi | (uint64_t)((addr & 0x3FFFFu) >> 4); };
    auto MD = [&](uint32_t addr) { return mhi | (uint64_t)((addr & 0x3FFFFu) >> 4); };
    // K-major product: D[128 x N] (+)= A[128 rows][64] . B[N][64]^T
    auto mma_k = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool acc_first) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc_mma_f16(d, KD(a_addr + 32 * kk), KD(b_addr + 32 * kk), idesc, (acc_first || kk > 0) ? 1u : 0u);
    };
    // MN-major product over the 128 rows: D[128 x N] (+)= [A0|A1]^T . B
    auto mma_mn = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool acc_first) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) tc_mma_f16(d, MD(a_addr + 2048 * kk), MD(b_addr + 2048 * kk), idesc, (acc_first || kk > 0) ? 1u : 0u);
    };
    const uint32_t d0 = tmem_base;
    uint32_t hs = 0;
    bool wg_acc = false;
    mbar_wait(bar_w, 0);
    for (int it = it_hi; it >= it_lo; --it)
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      for (int k = S - 1; k >= 0; --k) {
        // P1
        mbar_wait(bar_opnd, hs & 1); ++hs;
        tc_fence_after();
        if (elect_one()) {
          mma_k(d0 + TM_R0, tile_u32(T_Y), base + IMG_B1, idesc_128, false);
          tc_commit(bar_acc);
        }
        __syncwarp();
        // P2
        mbar_wait(bar_opnd, hs & 1); ++hs;
        tc_fence_after();
        if (elect_one()) {
          mma_k(d0 + TM_R0, tile_u32(T_H1F), base + IMG_W2, idesc_64, false);
          mma_k(d0 + TM_R1, tile_u32(T_H1G), base + IMG_V2, idesc_64, false);
          tc_commit(bar_acc);
        }
        __syncwarp();
        // D1: dh2f, dh1g ; trailing: dW3 += df^T h2f, db3 += df^T 1
        mbar_wait(bar_opnd, hs & 1); ++hs;
        tc_fence_after();
        if (elect_one()) {
          mma_k(d0 + TM_R0, tile_u32(T_DF), base + IMG_W3T, idesc_64, false);
          mma_k(d0 + TM_R1, tile_u32(T_DZ2G), base + IMG_V2T, idesc_64, false);
          tc_commit(bar_acc);
#ifndef TRAJSDE_BWD_NO_WGRAD
          mma_mn(d0 + TM_WGC, tile_u32(T_DF), tile_u32(T_H2F), imn_64, wg_acc);
          mma_mn(d0 + TM_SUM3, tile_u32(T_DF), tile_u32(T_TIME), imn_16, wg_acc);
#endif
        }
        __syncwarp();
        // D2: dh1f ; trailing: dW2|dV2 += [dz2f|dz2g]^T [h1f|h1g], db2|dc2
        mbar_wait(bar_opnd, hs & 1); ++hs;
        tc_fence_after();
        if (elect_one()) {
          mma_k(d0 + TM_R0, tile_u32(T_DZ2F), base + IMG_W2T, idesc_64, false);
          tc_commit(bar_acc);
#ifndef TRAJSDE_BWD_NO_WGRAD
          mma_mn(d0 + TM_WGA, tile_u32(T_DZ2F), tile_u32(T_H1F), imn_128, wg_acc);
          mma_mn(d0 + TM_SUM2, tile_u32(T_DZ2F), tile_u32(T_TIME), imn_16, wg_acc);
#endif
        }
        __syncwarp();
        // D3: dy = dz1f . W1y + dz1g . V1y ; trailing: dW1y|dV1y += [dz1f|dz1g]^T y, db1|dc1 and the time columns
        mbar_wait(bar_opnd, hs & 1); ++hs;
        tc_fence_after();
        if (elect_one()) {
          mma_k(d0 + TM_R0, tile_u32(T_H2F), base + IMG_W1YT, idesc_64, false);
          mma_k(d0 + TM_R0, tile_u32(T_DF), base + IMG_V1YT, idesc_64, true);
          tc_commit(bar_acc);
#ifndef TRAJSDE_BWD_NO_WGRAD
          mma_mn(d0 + TM_WGB, tile_u32(T_H2F), tile_u32(T_Y), imn_64, wg_acc);
          mma_mn(d0 + TM_SUM1, tile_u32(T_H2F), tile_u32(T_TIME), imn_16, wg_acc);
#endif
          tc_commit(bar_wg);
        }
        __syncwarp();
        wg_acc = true;
      }
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

}  // namespace sdetc
}  // namespace trajsde
