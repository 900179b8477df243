// C-ABI entry points (include/trajsde_b200.h): argument validation + dispatch.  No device allocation, no sync.
#include <string.h>

#include "common.cuh"

namespace trajsde {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; }   // tensors the kernels touch with 256-bit accesses

static int check_mlp(const TrajsdeMlp& m, const char* name) {
  if (!m.w1 || !m.b1 || !m.w2 || !m.b2 || !m.w3 || !m.b3)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "%s: null weight pointer", name);
  return TRAJSDE_OK;
}

static int check_sched(const TrajsdeSchedule& s) {
  if (s.n_steps <= 0 || s.n_outputs < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "schedule: n_steps=%d n_outputs=%d", s.n_steps, s.n_outputs);
  if (!s.step_tab || !s.out_begin || (s.n_outputs > 0 && !s.out_w))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "schedule: null table pointer");
  if (!aligned16(s.step_tab)) return set_error(TRAJSDE_ERR_UNSUPPORTED, "schedule: step_tab must be 16-byte aligned");
  return TRAJSDE_OK;
}

static int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(TRAJSDE_ERR_NO_DEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return set_error(TRAJSDE_ERR_NO_DEVICE, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (major != 10) return set_error(TRAJSDE_ERR_NO_DEVICE, "device compute capability %d.x: this library is sm_100a only", major);
  return TRAJSDE_OK;
}

}  // namespace trajsde

using namespace trajsde;

extern "C" {

int trajsde_abi_version(void) { return TRAJSDE_ABI_VERSION; }

const char* trajsde_last_error_string(void) { return last_error_buf(); }

int trajsde_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return set_error(TRAJSDE_ERR_NO_DEVICE, "no CUDA device");
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return set_error(TRAJSDE_ERR_CUDA, "cudaDeviceGetAttribute failed");
  return sms;
}

int64_t trajsde_euler_fwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual) {
  if (mode == TRAJSDE_MODE_EXACT_F32) return 0;
  if (mode == TRAJSDE_MODE_TC_F16) return euler_fwd_tc_workspace_bytes(rows, n_steps, dual);
  return set_error(TRAJSDE_ERR_UNSUPPORTED, "unknown mode %d", mode);
}

int trajsde_euler_fwd(const TrajsdeEulerFwdArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeEulerFwdArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeEulerFwdArgs));
  if (a->dim != TRAJSDE_DIM) return set_error(TRAJSDE_ERR_UNSUPPORTED, "dim %d unsupported (only 64)", a->dim);
  if (a->rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows %lld < 0", (long long)a->rows);
  int rc;
  if ((rc = check_sched(a->sched)) != 0) return rc;
  if ((rc = check_mlp(a->drift, "drift")) != 0) return rc;
  if ((rc = check_mlp(a->diffusion, "diffusion")) != 0) return rc;
  if (a->alt_mask && (rc = check_mlp(a->diffusion_alt, "diffusion_alt")) != 0) return rc;
  if (a->rows == 0) return TRAJSDE_OK;
  if (!a->y0 || !a->ys) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "y0/ys null");
  if (!aligned16(a->y0) || !aligned16(a->ys) || (a->y0_row_stride & 3) || (a->ys_row_stride & 3) || (a->ys_t_stride & 3) ||
      (a->noise.dw && !aligned16(a->noise.dw)) || (a->states && !aligned16(a->states)))
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "y0/ys/dw/states must be 16-byte aligned with strides multiple of 4 elements");
  if (a->y0_row_stride < 64 || a->ys_row_stride < 64)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "row strides must be >= 64");
  if ((rc = check_device()) != 0) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  switch (a->mode) {
    case TRAJSDE_MODE_EXACT_F32:
      return launch_euler_fwd_exact(*a, s);
    case TRAJSDE_MODE_TC_F16: {
      int64_t need = euler_fwd_tc_workspace_bytes(a->rows, a->sched.n_steps, a->alt_mask != nullptr);
      if (need < 0) return (int)need;
      if (a->workspace_bytes < need || (need > 0 && !a->workspace))
        return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
      return launch_euler_fwd_tc(*a, s);
    }
    default:
      return set_error(TRAJSDE_ERR_UNSUPPORTED, "unknown mode %d", a->mode);
  }
}

int64_t trajsde_euler_bwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual) {
  if (mode == TRAJSDE_MODE_EXACT_F32) return euler_bwd_exact_workspace_bytes(rows, n_steps, dual);
  if (mode == TRAJSDE_MODE_TC_F16) {
    const int64_t ex = euler_bwd_exact_workspace_bytes(rows, n_steps, dual), tc = euler_bwd_tc_workspace_bytes(rows, n_steps);
    return ex > tc ? ex : tc;
  }
  return set_error(TRAJSDE_ERR_UNSUPPORTED, "unknown mode %d", mode);
}

int trajsde_euler_bwd(const TrajsdeEulerBwdArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeEulerBwdArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeEulerBwdArgs));
  if (a->dim != TRAJSDE_DIM) return set_error(TRAJSDE_ERR_UNSUPPORTED, "dim %d unsupported (only 64)", a->dim);
  if (a->rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows < 0");
  int rc;
  if ((rc = check_sched(a->sched)) != 0) return rc;
  if ((rc = check_mlp(a->drift, "drift")) != 0) return rc;
  if ((rc = check_mlp(a->diffusion, "diffusion")) != 0) return rc;
  if (a->alt_mask && (rc = check_mlp(a->diffusion_alt, "diffusion_alt")) != 0) return rc;
  if (!a->grad_drift.w1 || !a->grad_drift.b1 || !a->grad_drift.w2 || !a->grad_drift.b2 || !a->grad_drift.w3 || !a->grad_drift.b3 ||
      !a->grad_diffusion.w1 || !a->grad_diffusion.b1 || !a->grad_diffusion.w2 || !a->grad_diffusion.b2 || !a->grad_diffusion.w3 ||
      !a->grad_diffusion.b3)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "null gradient pointer");
  if (a->alt_mask && (!a->grad_diffusion_alt.w1 || !a->grad_diffusion_alt.b1 || !a->grad_diffusion_alt.w2 ||
                      !a->grad_diffusion_alt.b2 || !a->grad_diffusion_alt.w3 || !a->grad_diffusion_alt.b3))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "null grad_diffusion_alt pointer with alt_mask set");
  if (a->rows > 0 && (!a->states || !a->grad_y0)) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "states/grad_y0 null");
  if (a->grad_y0 && !aligned32(a->grad_y0)) return set_error(TRAJSDE_ERR_UNSUPPORTED, "grad_y0 must be 32-byte aligned");
  if (a->grad_ys && ((a->grad_ys_row_stride & 3) || (a->grad_ys_t_stride & 3) || !aligned16(a->grad_ys)))
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "grad_ys must be 16-byte aligned with strides multiple of 4 elements");
  if (a->mode != TRAJSDE_MODE_EXACT_F32 && a->mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "unknown mode %d", a->mode);
  const bool use_tc = a->mode == TRAJSDE_MODE_TC_F16 && !(a->flags & TRAJSDE_BWD_FLAG_EXACT_KERNELS);
  int64_t need = use_tc ? euler_bwd_tc_workspace_bytes(a->rows, a->sched.n_steps)
                        : euler_bwd_exact_workspace_bytes(a->rows, a->sched.n_steps, a->alt_mask != nullptr);
  if (a->workspace_bytes < need || (need > 0 && !a->workspace))
    return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
  if ((rc = check_device()) != 0) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
  return use_tc ? launch_euler_bwd_tc(*a, s) : launch_euler_bwd_exact(*a, s);
}

int64_t trajsde_enc_fwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual) {
  if (mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "fused encoder exists in TC_F16 mode only (mode %d)", mode);
  return enc_fwd_tc_workspace_bytes(rows, n_steps, dual);
}

int trajsde_enc_fwd(const TrajsdeEncFwdArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeEncFwdArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeEncFwdArgs));
  if (a->dim != TRAJSDE_DIM) return set_error(TRAJSDE_ERR_UNSUPPORTED, "dim %d unsupported (only 64)", a->dim);
  if (a->mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "fused encoder exists in TC_F16 mode only (mode %d)", a->mode);
  if (a->rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows < 0");
  if (a->sched.n_steps <= 0 || !a->sched.step_tab || !aligned16(a->sched.step_tab))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "schedule: n_steps=%d or step_tab null/misaligned", a->sched.n_steps);
  int rc;
  if ((rc = check_mlp(a->drift, "drift")) != 0) return rc;
  if ((rc = check_mlp(a->diffusion, "diffusion")) != 0) return rc;
  if (a->alt_mask && (rc = check_mlp(a->diffusion_alt, "diffusion_alt")) != 0) return rc;
  const TrajsdeGru& g = a->gru;
  if (!g.u1 || !g.ub1 || !g.u2 || !g.ub2 || !g.r1 || !g.rb1 || !g.r2 || !g.rb2 || !g.n1 || !g.nb1 || !g.n2 || !g.nb2)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "gru: null weight pointer");
  if (a->rows == 0) return TRAJSDE_OK;
  if (!a->h0 || !a->aa_out || !a->slot || !a->obs_mask || !a->latent || !a->g_out)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "h0/aa_out/slot/obs_mask/latent/g_out null");
  if (!aligned16(a->h0) || (a->h0_row_stride & 3) || !aligned32(a->aa_out) || !aligned32(a->latent) ||
      (a->noise.dw && !aligned32(a->noise.dw)) || (a->y1_out && !aligned32(a->y1_out)))
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "h0 must be 16-byte aligned (row stride multiple of 4 elements), aa_out/latent/dw/y1_out 32-byte aligned");
  int64_t need = enc_fwd_tc_workspace_bytes(a->rows, a->sched.n_steps, a->alt_mask != nullptr);
  if (a->workspace_bytes < need || !a->workspace)
    return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
  if ((rc = check_device()) != 0) return rc;
  return launch_enc_fwd_tc(*a, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int64_t trajsde_enc_bwd_workspace_bytes(int32_t mode, int64_t rows, int32_t n_steps, int32_t dual) {
  if (mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "fused encoder exists in TC_F16 mode only (mode %d)", mode);
  if (rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows < 0");
  return enc_bwd_workspace_bytes(rows, n_steps, dual);
}

int trajsde_enc_bwd(const TrajsdeEncBwdArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeEncBwdArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeEncBwdArgs));
  if (a->dim != TRAJSDE_DIM) return set_error(TRAJSDE_ERR_UNSUPPORTED, "dim %d unsupported (only 64)", a->dim);
  if (a->mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "fused encoder exists in TC_F16 mode only (mode %d)", a->mode);
  if (a->rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows < 0");
  if (a->sched.n_steps <= 0 || !a->sched.step_tab || !aligned16(a->sched.step_tab))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "schedule: n_steps=%d or step_tab null/misaligned", a->sched.n_steps);
  int rc;
  if ((rc = check_mlp(a->drift, "drift")) != 0) return rc;
  if ((rc = check_mlp(a->diffusion, "diffusion")) != 0) return rc;
  if (a->alt_mask && (rc = check_mlp(a->diffusion_alt, "diffusion_alt")) != 0) return rc;
  const TrajsdeGru& g = a->gru;
  if (!g.u1 || !g.ub1 || !g.u2 || !g.ub2 || !g.r1 || !g.rb1 || !g.r2 || !g.rb2 || !g.n1 || !g.nb1 || !g.n2 || !g.nb2)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "gru: null weight pointer");
  const TrajsdeGruGrad& gg = a->grad_gru;
  if (!gg.u1 || !gg.ub1 || !gg.u2 || !gg.ub2 || !gg.r1 || !gg.rb1 || !gg.r2 || !gg.rb2 || !gg.n1 || !gg.nb1 || !gg.n2 || !gg.nb2)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "grad_gru: null pointer");
  const TrajsdeMlpGrad* mg[3] = {&a->grad_drift, &a->grad_diffusion, &a->grad_diffusion_alt};
  for (int i = 0; i < (a->alt_mask ? 3 : 2); ++i)
    if (!mg[i]->w1 || !mg[i]->b1 || !mg[i]->w2 || !mg[i]->b2 || !mg[i]->w3 || !mg[i]->b3)
      return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "null gradient pointer");
  if (a->rows == 0) {
    // an empty shard still owes its caller defined parameter gradients (they are accumulated into .grad and all-reduced): zeros
    if ((rc = check_device()) != 0) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(cuda_stream);
    for (int i = 0; i < 3; ++i) {
      // an empty mask tensor has a NULL data pointer: the alt net's gradients are zeroed whenever the caller passed buffers for them
      if (i == 2 && (!mg[i]->w1 || !mg[i]->b1 || !mg[i]->w2 || !mg[i]->b2 || !mg[i]->w3 || !mg[i]->b3)) break;
      const size_t n3 = i == 0 ? 64 * 64 : 64, nb3 = i == 0 ? 64 : 1;
      TS_CUDA_CHECK(cudaMemsetAsync(mg[i]->w1, 0, sizeof(float) * 64 * 66, s));
      TS_CUDA_CHECK(cudaMemsetAsync(mg[i]->b1, 0, sizeof(float) * 64, s));
      TS_CUDA_CHECK(cudaMemsetAsync(mg[i]->w2, 0, sizeof(float) * 64 * 64, s));
      TS_CUDA_CHECK(cudaMemsetAsync(mg[i]->b2, 0, sizeof(float) * 64, s));
      TS_CUDA_CHECK(cudaMemsetAsync(mg[i]->w3, 0, sizeof(float) * n3, s));
      TS_CUDA_CHECK(cudaMemsetAsync(mg[i]->b3, 0, sizeof(float) * nb3, s));
    }
    float* const w128[3] = {gg.u1, gg.r1, gg.n1};
    float* const w64[3] = {gg.u2, gg.r2, gg.n2};
    float* const bias[6] = {gg.ub1, gg.ub2, gg.rb1, gg.rb2, gg.nb1, gg.nb2};
    for (int i = 0; i < 3; ++i) {
      TS_CUDA_CHECK(cudaMemsetAsync(w128[i], 0, sizeof(float) * 64 * 128, s));
      TS_CUDA_CHECK(cudaMemsetAsync(w64[i], 0, sizeof(float) * 64 * 64, s));
    }
    for (int i = 0; i < 6; ++i) TS_CUDA_CHECK(cudaMemsetAsync(bias[i], 0, sizeof(float) * 64, s));
    return TRAJSDE_OK;
  }
  if (!a->h0 || !a->aa_out || !a->slot || !a->obs_mask || !a->latent || !a->y1 || !a->grad_h0)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "h0/aa_out/slot/obs_mask/latent/y1/grad_h0 null");
  if (!aligned16(a->h0) || !aligned16(a->aa_out) || !aligned16(a->latent) || !aligned16(a->y1) || !aligned32(a->grad_h0) ||
      (a->grad_latent && !aligned16(a->grad_latent)) || (a->noise.dw && !aligned16(a->noise.dw)) ||
      (a->grad_aa_out && !aligned32(a->grad_aa_out)))
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "tensors must be 16-byte aligned (grad_h0, grad_aa_out: 32-byte)");
  int64_t need = enc_bwd_workspace_bytes(a->rows, a->sched.n_steps, a->alt_mask != nullptr);
  if (a->workspace_bytes < need || !a->workspace)
    return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
  if ((rc = check_device()) != 0) return rc;
  return launch_enc_bwd(*a, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int64_t trajsde_gru_workspace_bytes(int32_t mode, int64_t rows) {
  if (mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "the fused GRU jump exists in TC_F16 mode only (mode %d)", mode);
  if (rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows < 0");
  return gru_standalone_workspace_bytes(rows);
}

static int gru_call(const TrajsdeGruArgs* a, void* cuda_stream, bool backward) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeGruArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeGruArgs));
  if (a->dim != TRAJSDE_DIM) return set_error(TRAJSDE_ERR_UNSUPPORTED, "dim %d unsupported (only 64)", a->dim);
  if (a->mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "the fused GRU jump exists in TC_F16 mode only (mode %d)", a->mode);
  if (a->rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows < 0");
  const TrajsdeGru& g = a->gru;
  if (!g.u1 || !g.ub1 || !g.u2 || !g.ub2 || !g.r1 || !g.rb1 || !g.r2 || !g.rb2 || !g.n1 || !g.nb1 || !g.n2 || !g.nb2)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "gru: null weight pointer");
  if (backward) {
    const TrajsdeGruGrad& gg = a->grad_gru;
    if (!gg.u1 || !gg.ub1 || !gg.u2 || !gg.ub2 || !gg.r1 || !gg.rb1 || !gg.r2 || !gg.rb2 || !gg.n1 || !gg.nb1 || !gg.n2 || !gg.nb2)
      return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "grad_gru: null pointer");
  }
  if (a->rows > 0) {
    if (!a->h_cur || !a->x || !a->mask) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "h_cur/x/mask null");
    if (!backward && !a->h_next) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "h_next null");
    if (backward && (!a->grad_h_next || !a->grad_h_cur || !a->grad_x)) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "grad_h_next/grad_h_cur/grad_x null");
    if (!aligned16(a->h_cur) || !aligned16(a->x) || (a->h_next && !aligned16(a->h_next)) || (a->grad_h_next && !aligned16(a->grad_h_next)) ||
        (a->grad_h_cur && !aligned16(a->grad_h_cur)) || (a->grad_x && !aligned16(a->grad_x)))
      return set_error(TRAJSDE_ERR_UNSUPPORTED, "tensors must be 16-byte aligned");
  }
  int64_t need = gru_standalone_workspace_bytes(a->rows);
  if (a->workspace_bytes < need || !a->workspace)
    return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
  int rc;
  if ((rc = check_device()) != 0) return rc;
  return launch_gru_standalone(*a, backward, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int trajsde_gru_fwd(const TrajsdeGruArgs* a, void* cuda_stream) { return gru_call(a, cuda_stream, false); }
int trajsde_gru_bwd(const TrajsdeGruArgs* a, void* cuda_stream) { return gru_call(a, cuda_stream, true); }

int64_t trajsde_heads_workspace_bytes(int32_t mode) {
  if (mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "the fused decoder heads exist in TC_F16 mode only (mode %d)", mode);
  return heads_workspace_bytes();
}

int trajsde_heads_fwd(const TrajsdeHeadsArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeHeadsArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeHeadsArgs));
  if (a->dim != TRAJSDE_DIM) return set_error(TRAJSDE_ERR_UNSUPPORTED, "dim %d unsupported (only 64)", a->dim);
  if (a->mode != TRAJSDE_MODE_TC_F16) return set_error(TRAJSDE_ERR_UNSUPPORTED, "the fused decoder heads exist in TC_F16 mode only (mode %d)", a->mode);
  if (a->rows < 0 || a->n_t < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows / n_t < 0");
  if (a->n_heads != 1 && a->n_heads != 2) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_heads %d (1 or 2)", a->n_heads);
  for (int h = 0; h < a->n_heads; ++h) {
    const TrajsdeHead& hd = a->head[h];
    if (!hd.w1 || !hd.b1 || !hd.ln_g || !hd.ln_b || !hd.w2 || !hd.b2)
      return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "head[%d]: null parameter pointer", h);
    const bool needs_out = h == 0 || !(a->flags & TRAJSDE_HEADS_FLAG_CAT4);
    if (a->rows > 0 && a->n_t > 0 && needs_out && !a->out[h]) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "out[%d] null", h);
  }
  if ((a->flags & TRAJSDE_HEADS_FLAG_CAT4) && a->n_heads != 2)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "TRAJSDE_HEADS_FLAG_CAT4 needs n_heads == 2 (loc and scale)");
  if ((a->flags & TRAJSDE_HEADS_FLAG_CAT4) && a->out[0] && (reinterpret_cast<uintptr_t>(a->out[0]) & 7u))
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "out[0] must be 8-byte aligned");
  if (a->rows == 0 || a->n_t == 0) return TRAJSDE_OK;
  if (!a->x || !aligned16(a->x) || (a->x_row_stride & 3) != 0 || (a->x_t_stride & 3) != 0 || a->x_row_stride < 64 || a->x_t_stride < 64)
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "x must be 16-byte aligned with row / t strides that are multiples of 4 elements and >= 64");
  const int64_t need = heads_workspace_bytes();
  if (a->workspace_bytes < need || !a->workspace)
    return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
  int rc;
  if ((rc = check_device()) != 0) return rc;
  return launch_heads_fwd(*a, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int64_t trajsde_heads_bwd_workspace_bytes(int32_t mode) {
  (void)mode;
  return heads_bwd_workspace_bytes();
}

int trajsde_heads_bwd(const TrajsdeHeadsBwdArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeHeadsBwdArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeHeadsBwdArgs));
  if (a->dim != TRAJSDE_DIM) return set_error(TRAJSDE_ERR_UNSUPPORTED, "dim %d unsupported (only 64)", a->dim);
  if (a->rows < 0 || a->n_t < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows / n_t < 0");
  if (a->n_heads != 1 && a->n_heads != 2) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_heads %d (1 or 2)", a->n_heads);
  for (int h = 0; h < a->n_heads; ++h) {
    const TrajsdeHead& hd = a->head[h];
    const TrajsdeHeadGrad& g = a->grad_head[h];
    if (!hd.w1 || !hd.b1 || !hd.ln_g || !hd.ln_b || !hd.w2 || !hd.b2)
      return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "head[%d]: null parameter pointer", h);
    if (!g.w1 || !g.b1 || !g.ln_g || !g.ln_b || !g.w2 || !g.b2)
      return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "grad_head[%d]: null pointer", h);
    if (a->grad_out[h] && (reinterpret_cast<uintptr_t>(a->grad_out[h]) & 7u))
      return set_error(TRAJSDE_ERR_UNSUPPORTED, "grad_out[%d] must be 8-byte aligned", h);
  }
  if ((a->flags & TRAJSDE_HEADS_FLAG_CAT4) && a->n_heads != 2)
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "TRAJSDE_HEADS_FLAG_CAT4 needs n_heads == 2 (loc and scale)");
  const bool empty = a->rows == 0 || a->n_t == 0;
  if (!empty) {
    if (!a->x || !aligned16(a->x) || (a->x_row_stride & 3) != 0 || (a->x_t_stride & 3) != 0)
      return set_error(TRAJSDE_ERR_UNSUPPORTED, "x must be 16-byte aligned with row / t strides that are multiples of 4 elements");
    if (!a->grad_x || !aligned16(a->grad_x) || (a->gx_row_stride & 3) != 0 || (a->gx_t_stride & 3) != 0)
      return set_error(TRAJSDE_ERR_UNSUPPORTED, "grad_x must be 16-byte aligned with row / t strides that are multiples of 4 elements");
  }
  const int64_t need = heads_bwd_workspace_bytes();
  if (a->workspace_bytes < need || !a->workspace)
    return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
  int rc;
  if ((rc = check_device()) != 0) return rc;
  return launch_heads_bwd(*a, reinterpret_cast<cudaStream_t>(cuda_stream));     // rows == 0: the reduce still writes zero gradients
}

int64_t trajsde_aggr_embed_workspace_bytes(int64_t n_modes, int64_t n_actors) {
  if (n_modes < 0 || n_actors < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_modes / n_actors < 0");
  return aggr_workspace_bytes(n_modes, n_actors);
}

static int aggr_call(const TrajsdeAggrArgs* a, void* cuda_stream, bool backward) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeAggrArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeAggrArgs));
  if (a->n_modes < 0 || a->n_actors < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_modes / n_actors < 0");
  if (!a->w || !a->b || !a->ln_g || !a->ln_b) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "null parameter pointer");
  const bool empty = a->n_modes == 0 || a->n_actors == 0;
  if (!empty && (!a->global_embed || !a->local_embed || !aligned16(a->global_embed) || !aligned16(a->local_embed)))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "global_embed / local_embed null or misaligned");
  if (!backward) {
    if (!empty && (!a->out || !aligned16(a->out))) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "out null or misaligned");
  } else {
    if (!a->grad_w || !a->grad_b || !a->grad_ln_g || !a->grad_ln_b) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "null gradient pointer");
    if (!empty && (!a->grad_out || !a->grad_global || !a->grad_local || !aligned16(a->grad_out) || !aligned16(a->grad_global) || !aligned16(a->grad_local)))
      return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "grad_out / grad_global / grad_local null or misaligned");
    const int64_t need = aggr_workspace_bytes(a->n_modes, a->n_actors);
    if (a->workspace_bytes < need || !a->workspace || (reinterpret_cast<uintptr_t>(a->workspace) & 255u))
      return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes (256-byte aligned)", (long long)a->workspace_bytes, (long long)need);
  }
  int rc;
  if ((rc = check_device()) != 0) return rc;
  return launch_aggr_embed(*a, backward, reinterpret_cast<cudaStream_t>(cuda_stream));
}
int trajsde_aggr_embed_fwd(const TrajsdeAggrArgs* a, void* cuda_stream) { return aggr_call(a, cuda_stream, false); }
int trajsde_aggr_embed_bwd(const TrajsdeAggrArgs* a, void* cuda_stream) { return aggr_call(a, cuda_stream, true); }

int trajsde_pi_head_fwd(const TrajsdePiArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdePiArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdePiArgs));
  if (a->n_modes < 0 || a->n_actors < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_modes / n_actors < 0");
  if (!a->w1 || !a->b1 || !a->ln_g || !a->ln_b || !a->w2 || !a->b2) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "null parameter pointer");
  const bool empty = a->n_modes == 0 || a->n_actors == 0;
  if (!empty && (!a->global_embed || !a->local_embed || !aligned16(a->global_embed) || !aligned16(a->local_embed)))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "global_embed / local_embed null or misaligned");
  if (!empty && !a->out) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "out null");
  int rc;
  if ((rc = check_device()) != 0) return rc;
  return launch_pi_head(*a, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int64_t trajsde_l2_loss_workspace_bytes(int64_t n_actors) {
  if (n_actors < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_actors < 0");
  return l2_workspace_bytes(n_actors);
}

static int l2_call(const TrajsdeL2Args* a, void* cuda_stream, bool backward) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeL2Args))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeL2Args));
  if (a->n_modes <= 0 || a->n_actors < 0 || a->n_t < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_modes <= 0 or n_actors / n_t < 0");
  if (a->loc_stride < 2) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "loc_stride < 2");
  if (!a->count || !a->best_mode) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "count / best_mode null");
  const bool empty = a->n_actors == 0 || a->n_t == 0;
  if (!empty && (!a->loc || !a->target || !a->reg_mask)) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "loc / target / reg_mask null");
  if (!backward) {
    if (!a->loss) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "loss null");
    const int64_t need = l2_workspace_bytes(a->n_actors);
    if (a->workspace_bytes < need || !a->workspace)
      return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)need);
  } else if (!a->grad_loss || (!empty && !a->grad_loc) || a->grad_loc_stride < 2) {
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "grad_loss / grad_loc null or grad_loc_stride < 2");
  }
  int rc;
  if ((rc = check_device()) != 0) return rc;
  return launch_l2_loss(*a, backward, reinterpret_cast<cudaStream_t>(cuda_stream));
}
int trajsde_l2_loss_fwd(const TrajsdeL2Args* a, void* cuda_stream) { return l2_call(a, cuda_stream, false); }
int trajsde_l2_loss_bwd(const TrajsdeL2Args* a, void* cuda_stream) { return l2_call(a, cuda_stream, true); }

int64_t trajsde_diff_bce_workspace_bytes(void) { return bce_workspace_bytes(); }

int trajsde_diff_bce(const TrajsdeBceArgs* a, void* cuda_stream) {
  if (!a) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "args == NULL");
  if (a->struct_bytes != sizeof(TrajsdeBceArgs))
    return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "struct_bytes %u != %zu (ABI mismatch)", a->struct_bytes, sizeof(TrajsdeBceArgs));
  if (a->n_in < 0 || a->n_out < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "n_in / n_out < 0");
  if ((a->n_in > 0 && !a->diff_in) || (a->n_out > 0 && !a->diff_out) || !a->loss) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "diff_in / diff_out / loss null");
  if (a->workspace_bytes < bce_workspace_bytes() || !a->workspace)
    return set_error(TRAJSDE_ERR_WORKSPACE, "workspace %lld < required %lld bytes", (long long)a->workspace_bytes, (long long)bce_workspace_bytes());
  int rc;
  if ((rc = check_device()) != 0) return rc;
  return launch_diff_bce(*a, reinterpret_cast<cudaStream_t>(cuda_stream));
}

int trajsde_philox_dw(const TrajsdeSchedule* sched, const TrajsdeNoise* noise, int64_t rows, float* dw_out, void* cuda_stream) {
  if (!sched || !noise) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "sched/noise null");
  int rc;
  if ((rc = check_sched(*sched)) != 0) return rc;
  if (rows < 0) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "rows < 0");
  if (rows == 0) return TRAJSDE_OK;
  if (!dw_out || !aligned16(dw_out)) return set_error(TRAJSDE_ERR_INVALID_ARGUMENT, "dw_out null or misaligned");
  if ((rc = check_device()) != 0) return rc;
  return launch_philox_dw(*sched, *noise, rows, dw_out, reinterpret_cast<cudaStream_t>(cuda_stream));
}

}  // extern "C"
