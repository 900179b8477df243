// Exact-fp32 Euler–Maruyama forward (validation path, TRAJSDE_MODE_EXACT_F32).
//
// Replaces, for one tile of rows and ALL solver steps inside one kernel, what the reference does with ~60 aten launches
// per step: BaseSDESolver.integrate (models/utils/sdeint.py:340-384), Euler_private.step (:477-485), prod_diagonal (:544),
// FFunc/GFunc forward (dec_hivt_nusargo_sde.py:119-127,154-158; enc…sep2.py:390-398,436-440), the dual-g row routing of
// the encoder's LSDEFunc.g (enc…sep2.py:470-482) and torchsde's linear_interp.
//
// Layout: persistent CTAs (256 threads) loop over 64-row tiles.  Weights stay in shared memory in nn.Linear layout
// [out][in] (row stride 68 floats), the state lives in registers (each thread owns a 4-row x 4-column patch, columns
// tx + 16*j) and in a shared activation buffer feeding the next layer.  All arithmetic is fp32 FFMA with k-ascending
// accumulation; tanh/sigmoid are libm-accurate; the Euler update and the output interpolation use un-contracted
// mul/add in the reference's evaluation order so that they round like the reference.
#include "common.cuh"

namespace trajsde {

namespace {

constexpr int EX_ROWS = 64;
constexpr int EX_THREADS = 256;
constexpr int LDS_ = 68;  // padded row stride (floats) of weight and activation tiles: conflict-free float4 access

// vector table slots (64 floats each)
enum { V_B1 = 0, V_W1S, V_W1C, V_B2, V_B3, V_C1, V_V1S, V_V1C, V_C2, V_W3G, V_AC1, V_AV1S, V_AV1C, V_AC2, V_AW3G, V_COUNT };

struct ExSmemLayout {
  // offsets in floats
  static constexpr int W_SZ = 64 * LDS_;
  static constexpr int w1y = 0, w2 = W_SZ, w3 = 2 * W_SZ, v1y = 3 * W_SZ, v2 = 4 * W_SZ, av1y = 5 * W_SZ, av2 = 6 * W_SZ;
  static constexpr int vec = 7 * W_SZ;
  static constexpr int act0 = vec + V_COUNT * 64;
  static constexpr int act1 = act0 + EX_ROWS * LDS_;
  static constexpr int act2 = act1 + EX_ROWS * LDS_;
  static constexpr int grow = act2 + EX_ROWS * LDS_;
  static constexpr int total = grow + EX_ROWS;
};

// out[r][n] = init + sum_k act[r][k] * W[n][k]; thread patch rows ty*4+i, cols tx+16*j.  k ascending.
__device__ __forceinline__ void dense64(const float* __restrict__ act, const float* __restrict__ w, int ty, int tx,
                                        float (&acc)[4][4]) {
#pragma unroll 2
  for (int k = 0; k < 64; k += 4) {
    float4 a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(act + (ty * 4 + i) * LDS_ + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4*>(w + (tx + 16 * j) * LDS_ + k);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
        acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
        acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
        acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
      }
  }
}

__device__ __forceinline__ void store_patch(float* act, int ty, int tx, const float (&v)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) act[(ty * 4 + i) * LDS_ + tx + 16 * j] = v[i][j];
}

// Stage one nn.Linear weight [64][ld] (first 64 columns) into smem [64][LDS_].
__device__ __forceinline__ void stage_weight(float* dst, const float* __restrict__ src, int ld, int tid) {
  for (int idx = tid; idx < 64 * 64; idx += EX_THREADS) {
    int n = idx >> 6, k = idx & 63;
    dst[n * LDS_ + k] = src[(size_t)n * ld + k];
  }
}

struct ExParams {
  TrajsdeEulerFwdArgs a;
  int num_tiles;
};

__global__ void __launch_bounds__(EX_THREADS, 1) euler_fwd_exact_kernel(const ExParams p) {
  extern __shared__ __align__(16) float smem[];
  using L = ExSmemLayout;
  const TrajsdeEulerFwdArgs& a = p.a;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const bool dual = a.alt_mask != nullptr;

  // ---- stage weights once per CTA --------------------------------------------------------------------------------
  stage_weight(smem + L::w1y, a.drift.w1, TS_IN1, tid);
  stage_weight(smem + L::w2, a.drift.w2, 64, tid);
  stage_weight(smem + L::w3, a.drift.w3, 64, tid);
  stage_weight(smem + L::v1y, a.diffusion.w1, TS_IN1, tid);
  stage_weight(smem + L::v2, a.diffusion.w2, 64, tid);
  if (dual) {
    stage_weight(smem + L::av1y, a.diffusion_alt.w1, TS_IN1, tid);
    stage_weight(smem + L::av2, a.diffusion_alt.w2, 64, tid);
  }
  if (tid < 64) {
    float* v = smem + L::vec;
    v[V_B1 * 64 + tid] = a.drift.b1[tid];
    v[V_W1S * 64 + tid] = a.drift.w1[tid * TS_IN1 + 64];
    v[V_W1C * 64 + tid] = a.drift.w1[tid * TS_IN1 + 65];
    v[V_B2 * 64 + tid] = a.drift.b2[tid];
    v[V_B3 * 64 + tid] = a.drift.b3[tid];
    v[V_C1 * 64 + tid] = a.diffusion.b1[tid];
    v[V_V1S * 64 + tid] = a.diffusion.w1[tid * TS_IN1 + 64];
    v[V_V1C * 64 + tid] = a.diffusion.w1[tid * TS_IN1 + 65];
    v[V_C2 * 64 + tid] = a.diffusion.b2[tid];
    v[V_W3G * 64 + tid] = a.diffusion.w3[tid];
    if (dual) {
      v[V_AC1 * 64 + tid] = a.diffusion_alt.b1[tid];
      v[V_AV1S * 64 + tid] = a.diffusion_alt.w1[tid * TS_IN1 + 64];
      v[V_AV1C * 64 + tid] = a.diffusion_alt.w1[tid * TS_IN1 + 65];
      v[V_AC2 * 64 + tid] = a.diffusion_alt.b2[tid];
      v[V_AW3G * 64 + tid] = a.diffusion_alt.w3[tid];
    }
  }
  const float c3_main = a.diffusion.b3[0];
  const float c3_alt = dual ? a.diffusion_alt.b3[0] : 0.f;
  __syncthreads();

  const float* vec = smem + L::vec;
  float* act0 = smem + L::act0;
  float* act1 = smem + L::act1;
  float* act2 = smem + L::act2;
  float* grow = smem + L::grow;
  const int S = a.sched.n_steps;
  const uint64_t noise_seed = a.noise.dw ? 0ull : ts_noise_seed(a.noise);

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * EX_ROWS;
    float y[4][4];
    bool valid[4], use_alt[4];
    int64_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      r[i] = row0 + ty * 4 + i;
      valid[i] = r[i] < a.rows;
      use_alt[i] = dual && valid[i] && (a.alt_mask[r[i]] == 0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        y[i][j] = valid[i] ? a.y0[r[i] * a.y0_row_stride + tx + 16 * j] : 0.f;
        if (valid[i]) a.ys[r[i] * a.ys_row_stride + tx + 16 * j] = y[i][j];  // ys[0] = y0
      }
    }
    bool t_main = false, t_alt = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      t_main |= valid[i] && !use_alt[i];
      t_alt |= use_alt[i];
    }
    __syncthreads();  // previous tile finished with act0/grow
    store_patch(act0, ty, tx, y);
    const int any_main = __syncthreads_or(t_main);
    const int any_alt = __syncthreads_or(t_alt);

    for (int k = 0; k < S; ++k) {
      const float4 st = *reinterpret_cast<const float4*>(a.sched.step_tab + 4 * k);
      const float h = st.y, sn = st.z, cs = st.w;
      if (a.states) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (valid[i])
#pragma unroll
            for (int j = 0; j < 4; ++j) a.states[((int64_t)k * a.rows + r[i]) * 64 + tx + 16 * j] = y[i][j];
      }
      float acc[4][4], f[4][4];
      // ---- drift: f = W3 tanh(W2 tanh(W1y y + b1 + w1s sin + w1c cos) + b2) + b3 ------------------------------------
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = tx + 16 * j;
        const float b = fmaf(vec[V_W1C * 64 + n], cs, fmaf(vec[V_W1S * 64 + n], sn, vec[V_B1 * 64 + n]));
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = b;
      }
      dense64(act0, smem + L::w1y, ty, tx, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = tanhf(acc[i][j]);
      store_patch(act1, ty, tx, acc);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = vec[V_B2 * 64 + tx + 16 * j];
      dense64(act1, smem + L::w2, ty, tx, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = tanhf(acc[i][j]);
      store_patch(act2, ty, tx, acc);
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i][j] = vec[V_B3 * 64 + tx + 16 * j];
      dense64(act2, smem + L::w3, ty, tx, f);

      // ---- diffusion: g = sigmoid(w3 . tanh(V2 tanh(V1y y + c1 + ..) + c2) + c3), per-row net selection --------------
#pragma unroll 1
      for (int net = 0; net < 2; ++net) {
        if (net == 0 ? !any_main : !any_alt) continue;  // CTA-uniform
        const int vb = net == 0 ? V_C1 : V_AC1;
        const float* w1 = smem + (net == 0 ? L::v1y : L::av1y);
        const float* w2 = smem + (net == 0 ? L::v2 : L::av2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = tx + 16 * j;
          const float b = fmaf(vec[(vb + 2) * 64 + n], cs, fmaf(vec[(vb + 1) * 64 + n], sn, vec[vb * 64 + n]));
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = b;
        }
        dense64(act0, w1, ty, tx, acc);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = tanhf(acc[i][j]);
        store_patch(act1, ty, tx, acc);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i][j] = vec[(vb + 3) * 64 + tx + 16 * j];
        dense64(act1, w2, ty, tx, acc);
        float part[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          part[i] = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) part[i] = fmaf(tanhf(acc[i][j]), vec[(vb + 4) * 64 + tx + 16 * j], part[i]);
#pragma unroll
          for (int off = 8; off >= 1; off >>= 1) part[i] += __shfl_xor_sync(0xffffffffu, part[i], off);
        }
        if (tx == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (use_alt[i] == (net == 1)) grow[ty * 4 + i] = ts_sigmoid_exact(part[i] + (net == 0 ? c3_main : c3_alt));
        }
        __syncthreads();
      }

      // ---- Euler update (sdeint.py:484) + outputs (linear_interp) ------------------------------------------------------
      const float sqrt_h = sqrtf(h);
      const int ob = a.sched.out_begin[k], oe = a.sched.out_begin[k + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float g = grow[ty * 4 + i];
        float yn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float dw;
          if (a.noise.dw) {
            dw = valid[i] ? a.noise.dw[((int64_t)k * a.rows + r[i]) * 64 + tx + 16 * j] : 0.f;
          } else {
            const int c = tx + 16 * j;
            const float4 n4 = philox_dw4(noise_seed, (uint64_t)r[i] + a.noise.row_offset,
                                         a.noise.step_offset + (uint32_t)k, (uint32_t)(c >> 2), sqrt_h);
            dw = (c & 3) == 0 ? n4.x : (c & 3) == 1 ? n4.y : (c & 3) == 2 ? n4.z : n4.w;
          }
          yn[j] = __fadd_rn(__fadd_rn(y[i][j], __fmul_rn(f[i][j], h)), __fmul_rn(g, dw));
        }
        if (valid[i]) {
          for (int o = ob; o < oe; ++o) {
            const float w0 = a.sched.out_w[2 * o], w1 = a.sched.out_w[2 * o + 1];
            float* dst = a.ys + (int64_t)(o + 1) * a.ys_t_stride + r[i] * a.ys_row_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) dst[tx + 16 * j] = __fadd_rn(__fmul_rn(w0, y[i][j]), __fmul_rn(w1, yn[j]));
          }
          if (k == S - 1 && a.g_last && tx == 0) a.g_last[r[i]] = g;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) y[i][j] = yn[j];
      }
      store_patch(act0, ty, tx, y);
      __syncthreads();
    }
  }
}

__global__ void philox_dw_kernel(TrajsdeSchedule sched, TrajsdeNoise noise, int64_t rows, float* __restrict__ out) {
  const int64_t total = (int64_t)sched.n_steps * rows * 16;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t chunk = (uint32_t)(idx & 15);
    const int64_t rk = idx >> 4;
    const int64_t row = rk % rows;
    const int k = (int)(rk / rows);
    const float sqrt_h = sqrtf(sched.step_tab[4 * k + 1]);
    *reinterpret_cast<float4*>(out + idx * 4) =
        philox_dw4(ts_noise_seed(noise), (uint64_t)row + noise.row_offset, noise.step_offset + (uint32_t)k, chunk, sqrt_h);
  }
}

}  // namespace

int launch_euler_fwd_exact(const TrajsdeEulerFwdArgs& a, cudaStream_t s) {
  int dev = 0, sms = 0;
  TS_CUDA_CHECK(cudaGetDevice(&dev));
  TS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem_bytes = (size_t)ExSmemLayout::total * sizeof(float);
  TS_CUDA_CHECK(cudaFuncSetAttribute(euler_fwd_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  ExParams p;
  p.a = a;
  p.num_tiles = (int)((a.rows + EX_ROWS - 1) / EX_ROWS);
  if (p.num_tiles == 0) return TRAJSDE_OK;
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  euler_fwd_exact_kernel<<<grid, EX_THREADS, smem_bytes, s>>>(p);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

int launch_philox_dw(const TrajsdeSchedule& sched, const TrajsdeNoise& noise, int64_t rows, float* out, cudaStream_t s) {
  const int64_t total = (int64_t)sched.n_steps * rows * 16;
  if (total == 0) return TRAJSDE_OK;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  philox_dw_kernel<<<(int)blocks, 256, 0, s>>>(sched, noise, rows, out);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

}  // namespace trajsde
