// Backward of the fused encoder recurrence (trajsde_enc_bwd): what torch.autograd computes through the loop body of
// LocalEncoderSDESepPara2.forward (models/encoders/enc_hivt_nusargo_sde_sep2.py:128-182): n_steps x [sdeint_dual one Euler step
// (models/utils/sdeint.py:110-197) + GRU_Unit jump (models/utils/ode_utils.py:136-152)].
//
// One C-ABI call = one reverse sweep over the iterations.  Default: ONE persistent launch whose CTAs take the GRU / SDE-step roles and
// hand tiles to each other (enc_bwd_sweep.cu).  With TRAJSDE_BWD_FLAG_PER_STEP_LAUNCHES / _EXACT_KERNELS, or more than 128 iterations, the
// same sweep enqueued from the host without any synchronisation, two launches per iteration:
//   for i = S-1 .. 0:   a_h   = carry (dL/dy0 of iteration i+1's SDE step) + dL/d latent[i]
//                       GRU backward (gru_bwd_tc.cu; gru_bwd.cu = fp32 validation kernel):  a_h -> a_y1, dL/d aa_out[slot_i], GRU partials
//                       SDE step backward (euler_bwd_tc.cu):  a_y1, dL/dg[i] -> carry, SDE weight-gradient partials (one pass per diffusion net, same launch)
// Partials accumulate in the workspace across the sweep and are reduced once, in fixed order (bit-reproducible).
#include "bwd_common.cuh"

namespace trajsde {

using namespace bwd;

namespace {

__global__ void enc_bwd_tables_kernel(int32_t* out_begin, float* out_w) {
  out_begin[0] = 0;
  out_begin[1] = 1;
  out_w[0] = 0.f;   // ys[1] = 0 * Y[0] + 1 * Y[1]
  out_w[1] = 1.f;
}

struct Ws {
  int32_t* out_begin;
  float* out_w;
  uint32_t* amax;
  uint8_t* img0;
  uint8_t* img1;
  uint8_t* gru_img;
  float* gbuf;     // [2][rows][64]: slab 0 zero (dL/d ys[0]), slab 1 = dL/d y1 of the current iteration
  float* carry;    // [rows][64]
  float* part0;
  float* part1;
  float* gru_part;
  int32_t* counters;   // per-tile progress counters of the single-launch sweep
  int64_t bytes;
};

int64_t align256(int64_t x) { return (x + 255) & ~255ll; }
constexpr int64_t SWEEP_MAX_ROWS = 40960;   // measured: 2,688 rows 1.35 -> 0.98 ms, 21,504 rows 2.44 -> 2.03 ms, 86,016 rows 6.3 -> 7.0 ms (bench_micro/enc_bwd_ab.py)

Ws carve(void* base, int64_t rows) {
  Ws w;
  uint8_t* p = static_cast<uint8_t*>(base);
  int64_t off = 0;
  auto take = [&](int64_t n) { uint8_t* q = p ? p + off : nullptr; off += align256(n); return q; };
  w.out_begin = reinterpret_cast<int32_t*>(take(16));
  w.out_w = reinterpret_cast<float*>(take(16));
  w.amax = reinterpret_cast<uint32_t*>(take(16));
  w.img0 = take(BWD_TC_IMG_BYTES);
  w.img1 = take(BWD_TC_IMG_BYTES);
  w.gru_img = take(GRU_TC_IMG_BYTES);
  w.gbuf = reinterpret_cast<float*>(take(2 * rows * 64 * 4));
  w.carry = reinterpret_cast<float*>(take(rows * 64 * 4));
  w.part0 = reinterpret_cast<float*>(take((int64_t)MAX_PARTIALS * G_PAD * 4));
  w.part1 = reinterpret_cast<float*>(take((int64_t)MAX_PARTIALS * G_PAD * 4));
  w.gru_part = reinterpret_cast<float*>(take((int64_t)MAX_PARTIALS * GRU_G_PAD * 4));
  w.counters = reinterpret_cast<int32_t*>(take(enc_bwd_sweep_counter_bytes(rows)));
  w.bytes = off;
  return w;
}

}  // namespace

int64_t enc_bwd_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual) {
  (void)n_steps;
  (void)dual;
  return carve(nullptr, rows).bytes;
}

int launch_enc_bwd(const TrajsdeEncBwdArgs& a, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0) return set_error(TRAJSDE_ERR_UNSUPPORTED, "workspace must be 256-byte aligned");
  const Ws w = carve(a.workspace, a.rows);
  const int S = a.sched.n_steps;
  const int64_t slab = a.rows * 64;
  const bool dual = a.alt_mask != nullptr;
  const bool gru_tc = !(a.flags & TRAJSDE_BWD_FLAG_EXACT_KERNELS);   // fp32 GRU kernel kept for A/B validation
  // single launch: up to SWEEP_MAX_ROWS (beyond it every per-iteration kernel fills the device for long enough that the launches cost
  // nothing, and the sweep's fixed split of the SMs between the roles can only lose); 128: step-table rows staged in shared memory
  const bool sweep = gru_tc && !(a.flags & TRAJSDE_BWD_FLAG_PER_STEP_LAUNCHES) && S <= 128 && a.rows <= SWEEP_MAX_ROWS;
  int Gg = 0, Gs = 0;
  if (sweep && enc_bwd_sweep_grid(a.rows, dual, &Gg, &Gs) != 0) return set_error(TRAJSDE_ERR_CUDA, "cudaGetDevice failed");
  const int grid = sweep ? Gs : bwd_tc_grid(a.rows, dual), ggrid = sweep ? Gg : gru_tc ? bwd_tc_grid(a.rows) : gru_bwd_grid(a.rows);
  int rc;

  enc_bwd_tables_kernel<<<1, 1, 0, s>>>(w.out_begin, w.out_w);
  TS_CUDA_CHECK(cudaGetLastError());
  TS_CUDA_CHECK(cudaMemsetAsync(w.amax, 0, 4, s));
  TS_CUDA_CHECK(cudaMemsetAsync(w.gbuf, 0, slab * 4, s));
  TS_CUDA_CHECK(cudaMemsetAsync(w.part0, 0, (size_t)grid * G_PAD * 4, s));
  if (dual) TS_CUDA_CHECK(cudaMemsetAsync(w.part1, 0, (size_t)grid * G_PAD * 4, s));
  TS_CUDA_CHECK(cudaMemsetAsync(w.gru_part, 0, (size_t)ggrid * GRU_G_PAD * 4, s));
  // one loss scale for the whole sweep, from the incoming gradients (the carried adjoint has 2^13 of head-room above it)
  if (a.grad_latent && (rc = bwd_tc_absmax(a.grad_latent, S, a.rows, slab, 64, w.amax, s)) != 0) return rc;
  if (a.grad_g && (rc = bwd_tc_absmax(a.grad_g, 1, (int64_t)S * a.rows, 0, 0, w.amax, s)) != 0) return rc;

  if (gru_tc && (rc = gru_bwd_tc_pack(a.gru, w.gru_img, s)) != 0) return rc;

  TrajsdeEulerBwdArgs b;
  memset(&b, 0, sizeof(b));
  b.struct_bytes = sizeof(b);
  b.mode = TRAJSDE_MODE_TC_F16;
  b.rows = a.rows;
  b.dim = 64;
  b.sched.n_steps = 1;
  b.sched.n_outputs = 1;
  b.sched.out_begin = w.out_begin;
  b.sched.out_w = w.out_w;
  b.drift = a.drift;
  b.diffusion = a.diffusion;
  b.diffusion_alt = a.diffusion_alt;
  b.alt_mask = a.alt_mask;
  b.noise = a.noise;
  b.grad_ys = w.gbuf;
  b.grad_ys_t_stride = slab;
  b.grad_ys_row_stride = 64;
  b.status = a.status;
  if ((rc = bwd_tc_pack(b, w.img0, s)) != 0) return rc;
  if (dual) {
    TrajsdeEulerBwdArgs b2 = b;
    b2.diffusion = a.diffusion_alt;
    if ((rc = bwd_tc_pack(b2, w.img1, s)) != 0) return rc;
  }

  if (sweep) {
    b.sched.step_tab = a.sched.step_tab;
    b.grad_g_last = a.grad_g;
    if ((rc = launch_enc_bwd_sweep(a, b, w.img0, dual ? w.img1 : nullptr, w.gru_img, w.amax, w.part0, w.part1, w.gru_part, w.gbuf, w.carry,
                                   w.counters, Gg, Gs, s)) != 0)
      return rc;
  }
  for (int i = S - 1; i >= 0 && !sweep; --i) {
    const float* carry_in = i == S - 1 ? nullptr : w.carry;
    const float* glat = a.grad_latent ? a.grad_latent + (int64_t)i * slab : nullptr;
    rc = gru_tc ? launch_gru_bwd_tc(a.rows, a.y1 + (int64_t)i * slab, a.aa_out, slab, a.obs_mask, a.obs_mask_row_stride, a.slot, i, carry_in,
                                    glat, w.gbuf + slab, a.grad_aa_out, w.gru_img, w.amax, w.gru_part, nullptr, s, i < S - 1)
                : launch_gru_bwd(a.gru, a.rows, a.y1 + (int64_t)i * slab, a.aa_out, slab, a.obs_mask, a.obs_mask_row_stride, a.slot, i,
                                 carry_in, glat, w.gbuf + slab, a.grad_aa_out, w.gru_part, s);
    if (rc != 0) return rc;
    b.sched.step_tab = a.sched.step_tab + 4 * i;
    b.noise.dw = a.noise.dw ? a.noise.dw + (int64_t)i * slab : nullptr;
    b.noise.step_offset = a.noise.step_offset + (uint32_t)i;
    b.states = i == 0 ? a.h0 : a.latent + (int64_t)(i - 1) * slab;
    b.grad_g_last = a.grad_g ? a.grad_g + (int64_t)i * a.rows : nullptr;
    b.grad_y0 = i == 0 ? a.grad_h0 : w.carry;
    // both nets' passes in one launch; launched with programmatic stream serialisation: its prologue (barriers, TMEM) overlaps the GRU
    // kernel's tail, and it lets the next iteration's GRU kernel do the same (the chain is dependent, the prologues are not)
    if ((rc = bwd_tc_main(b, w.img0, dual ? w.img1 : nullptr, w.amax, w.part0, w.part1, 1, s, true)) != 0) return rc;
  }
  if ((rc = launch_euler_bwd_reduce(w.part0, dual ? w.part1 : nullptr, grid, dual ? grid : 0, a.grad_drift, a.grad_diffusion,
                                    a.grad_diffusion_alt, s)) != 0)
    return rc;
  return launch_gru_bwd_reduce(w.gru_part, ggrid, a.grad_gru, s);
}

// ---- stand-alone GRU_Unit forward / backward (trajsde_gru_fwd / trajsde_gru_bwd): the jump as its own operator, for the drop-in path
// that keeps the reference's encoder loop (install() rebinds gru_unit.forward) --------------------------------------------------------
int64_t gru_standalone_workspace_bytes(int64_t rows) {
  (void)rows;
  return align256(GRU_TC_IMG_BYTES) + 256 + 256 + align256((int64_t)MAX_PARTIALS * GRU_G_PAD * 4);
}

int launch_gru_standalone(const TrajsdeGruArgs& a, bool backward, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0) return set_error(TRAJSDE_ERR_UNSUPPORTED, "workspace must be 256-byte aligned");
  uint8_t* ws = static_cast<uint8_t*>(a.workspace);
  uint8_t* img = ws;
  uint32_t* amax = reinterpret_cast<uint32_t*>(ws + align256(GRU_TC_IMG_BYTES));
  int32_t* slot0 = reinterpret_cast<int32_t*>(ws + align256(GRU_TC_IMG_BYTES) + 256);
  float* part = reinterpret_cast<float*>(ws + align256(GRU_TC_IMG_BYTES) + 512);
  const int grid = bwd_tc_grid(a.rows);
  int rc;
  TS_CUDA_CHECK(cudaMemsetAsync(amax, 0, 512, s));                 // amax word and the single slot index (0)
  if ((rc = gru_bwd_tc_pack(a.gru, img, s)) != 0) return rc;
  if (!backward)
    return launch_gru_bwd_tc(a.rows, a.h_cur, a.x, 0, a.mask, 1, slot0, 0, nullptr, nullptr, nullptr, nullptr, img, amax, part, a.h_next, s);
  TS_CUDA_CHECK(cudaMemsetAsync(part, 0, (size_t)grid * GRU_G_PAD * 4, s));
  if ((rc = bwd_tc_absmax(a.grad_h_next, 1, a.rows, 0, 64, amax, s)) != 0) return rc;
  if ((rc = launch_gru_bwd_tc(a.rows, a.h_cur, a.x, 0, a.mask, 1, slot0, 0, nullptr, a.grad_h_next, a.grad_h_cur, a.grad_x, img, amax, part,
                              nullptr, s)) != 0)
    return rc;
  return launch_gru_bwd_reduce(part, grid, a.grad_gru, s);
}

}  // namespace trajsde
