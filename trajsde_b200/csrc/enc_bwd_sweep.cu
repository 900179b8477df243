// Single-launch backward of the fused encoder recurrence (the default path of trajsde_enc_bwd; enc_bwd.cu drives it).
//
// The per-iteration form enqueues 2 S kernels (GRU backward, SDE-step backward; S = 21 in the reference configuration), each a dependent
// wave with its own prologue (barriers, TMEM, 75 KB weight image), weight-gradient flush and drain: ~50 us per iteration for ~15 us of
// tile work at 21,504 rows.  Here ONE persistent kernel runs the whole reverse sweep.  Its CTAs take ROLES by block index:
//
//   role 0      GRU_Unit backward (gru_bwd_tc_body.cuh)            CTAs [0, Gg)
//   role 1 (2)  SDE-step backward (euler_bwd_tc_body.cuh), pass 0 (and pass 1 = the other diffusion net's rows)   CTAs [Gg, Gg + Gs) ([.., Gg + 2 Gs))
//
// A GRU tile costs about twice an SDE-step tile (8 dependent tensor-core phases against 5; measured 54 k vs 26 k clk, bench_micro/
// enc_bwd_ab.py --timeline), so the GRU role gets half of the SMs and each SDE pass a quarter (one pass: two thirds / one third).
// Every CTA of a role owns a contiguous range of 128-row tiles and walks it once per iteration i = S-1 .. 0, in ascending order.  A tile moves
// GRU(i) -> SDE(i) -> GRU(i-1) -> ... through two per-tile progress counters in global memory (bwd_tc_common.cuh: SweepCtl): the rows of
// dL/dy1 and of the carried adjoint go through L2 (written once, read once, 32 KB per tile), the weight images are loaded once per CTA, and the
// weight-gradient accumulators stay in TMEM for all S iterations, flushed once.  With more than one tile per CTA the roles pipeline (the GRU
// role works on tile j+1 of iteration i while the SDE roles take tile j); with one tile per CTA the chain is the dependent one, minus the
// launches.  Gg + 2 Gs <= SM count and one CTA per SM (215 KB of shared memory each) make the launch co-resident, which the counters rely on.
#include "euler_bwd_tc_body.cuh"
#include "gru_bwd_tc_body.cuh"

namespace trajsde {

using namespace bwd;
using bwdtc::SweepCtl;

namespace {

constexpr uint32_t SWEEP_SMEM = grutc::SMEM_ALLOC > sdetc::SMEM_ALLOC ? grutc::SMEM_ALLOC : sdetc::SMEM_ALLOC;
static_assert(grutc::NUM_THREADS == sdetc::NUM_THREADS, "the roles share one block shape");

template <bool HAS_DW>
__global__ void __launch_bounds__(grutc::NUM_THREADS, 1) enc_bwd_sweep_kernel(const grutc::GruTcParams pg, const sdetc::BwdTcParams ps,
                                                                               const SweepCtl sw, const int Gg, const int Gs) {
  extern __shared__ uint8_t smem_raw[];
  const int role = (int)blockIdx.x < Gg ? 0 : 1 + ((int)blockIdx.x - Gg) / Gs;   // CTA-uniform
  const int cta = role == 0 ? (int)blockIdx.x : ((int)blockIdx.x - Gg) % Gs;
#ifdef TRAJSDE_SWEEP_TIMELINE
  const long long t0 = clock64();
#endif
  if (role == 0) grutc::gru_bwd_tc_body<true>(pg, sw, cta, Gg, smem_raw);
  else sdetc::euler_bwd_tc_body<HAS_DW, true>(ps, sw, cta, Gs, role - 1, smem_raw);
#ifdef TRAJSDE_SWEEP_TIMELINE
  if (threadIdx.x == 0) bwdtc::g_sweep_tl[blockIdx.x * 4] += clock64() - t0;
#endif
}

}  // namespace

// CTAs of the GRU role (*gg) and of each SDE pass (*gs): SMs split 2 : 1 : 1 (dual diffusion) or 2 : 1, at most one CTA per tile
int enc_bwd_sweep_grid(int64_t rows, bool dual, int* gg, int* gs) {
  int dev = 0, sms = 0;
  *gg = *gs = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  const int64_t tiles = (rows + grutc::TILE_M - 1) / grutc::TILE_M;
  int s = sms / (dual ? 4 : 3);
  if (s < 1) s = 1;
  int g = sms - (dual ? 2 : 1) * s;
  if (g > MAX_PARTIALS) g = MAX_PARTIALS;
  if (g < 1) g = 1;
  *gs = (int)(tiles < s ? tiles : s);
  *gg = (int)(tiles < g ? tiles : g);
  return 0;
}

int64_t enc_bwd_sweep_counter_bytes(int64_t rows) { return (3 * ((rows + grutc::TILE_M - 1) / grutc::TILE_M) + 64) * 4; }

// `b`: ONE Euler step as enc_bwd.cu describes it (n_steps = n_outputs = 1, grad_ys = gbuf), with the per-iteration members holding the BASES:
// sched.step_tab = the S-row table, noise = the call's (dw base / step_offset of iteration 0), grad_g_last = grad_g base.
int launch_enc_bwd_sweep(const TrajsdeEncBwdArgs& a, const TrajsdeEulerBwdArgs& b, const uint8_t* img0, const uint8_t* img1, const uint8_t* gru_img,
                         const uint32_t* amax_bits, float* part0, float* part1, float* gru_part, float* gbuf, float* carry, int32_t* counters,
                         int Gg, int Gs, cudaStream_t s) {
  if (a.rows >= (int64_t)1 << 31) return set_error(TRAJSDE_ERR_UNSUPPORTED, "rows >= 2^31 unsupported in TC mode");
  const bool dual = img1 != nullptr;
  const int S = a.sched.n_steps;
  const int64_t slab = a.rows * 64;
  const int num_tiles = (int)((a.rows + grutc::TILE_M - 1) / grutc::TILE_M);
  if (Gg <= 0 || Gs <= 0) return TRAJSDE_OK;

  grutc::GruTcParams pg;
  pg.rows = a.rows;
  pg.y1 = a.y1;
  pg.x = a.aa_out;
  pg.x_slab = slab;
  pg.obs_mask = a.obs_mask;
  pg.obs_mask_row_stride = a.obs_mask_row_stride;
  pg.slot = a.slot;
  pg.iter = 0;
  pg.carry = carry;
  pg.grad_latent = a.grad_latent;
  pg.grad_y1 = gbuf + slab;
  pg.grad_x = a.grad_aa_out;
  pg.img = gru_img;
  pg.amax_bits = amax_bits;
  pg.partial = gru_part;
  pg.num_tiles = num_tiles;
  pg.fwd_only = 0;
  pg.h_out = nullptr;

  sdetc::BwdTcParams ps;
  ps.a = b;
  ps.img[0] = img0;
  ps.img[1] = img1;
  ps.partial[0] = part0;
  ps.partial[1] = part1;
  ps.filter[0] = dual ? 1 : 0;
  ps.filter[1] = 2;
  ps.amax_bits = amax_bits;
  ps.row_map = nullptr;
  ps.n_active = nullptr;
  ps.num_tiles = num_tiles;
  ps.accumulate = 0;

  SweepCtl sw;
  sw.S = S;
  sw.gru_done = counters;
  sw.sde_done[0] = counters + num_tiles;
  sw.sde_done[1] = dual ? counters + 2 * num_tiles : nullptr;
  sw.abort_word = counters + 3 * num_tiles;
  sw.status = a.status;
  sw.h0 = a.h0;
  sw.latent = a.latent;
  sw.carry = carry;
  sw.grad_h0 = a.grad_h0;
  sw.slab = slab;

  TS_CUDA_CHECK(cudaMemsetAsync(counters, 0, (size_t)enc_bwd_sweep_counter_bytes(a.rows), s));
  const int grid = Gg + Gs * (dual ? 2 : 1);
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM);
    if (e != cudaSuccess) return e;
    kern<<<grid, grutc::NUM_THREADS, SWEEP_SMEM, s>>>(pg, ps, sw, Gg, Gs);
    return cudaGetLastError();
  };
  TS_CUDA_CHECK(a.noise.dw ? launch(enc_bwd_sweep_kernel<true>) : launch(enc_bwd_sweep_kernel<false>));
  return TRAJSDE_OK;
}

}  // namespace trajsde

#ifdef TRAJSDE_SWEEP_TIMELINE
// debug build only: per-CTA clocks of the sweep kernel since the last call (then reset)
extern "C" int trajsde_debug_sweep_timeline(long long* out640) {
  static long long zero[640];
  if (cudaMemcpyFromSymbol(out640, trajsde::bwdtc::g_sweep_tl, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(trajsde::bwdtc::g_sweep_tl, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif

#ifdef TRAJSDE_GRU_TIMELINE
extern "C" int trajsde_debug_gru_segments(long long* out24) {
  static long long zero[24];
  if (cudaMemcpyFromSymbol(out24, trajsde::grutc::g_gru_seg, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(trajsde::grutc::g_gru_seg, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif

