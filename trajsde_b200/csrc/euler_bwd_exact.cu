// Exact-fp32 backward of the fused Euler–Maruyama solve: discretise-then-optimise (what torch.autograd computes through the
// reference solver, config `adjoint: false`, configs/nusargo/hivt_nuSArgo_sdesepenc_sdedec.yml:41).
//
// Forward being differentiated (models/utils/sdeint.py:340-384,477-485,544 + nets dec…sde.py:119-127,154-158 / enc…sep2.py:
// 390-398,436-440,470-482):      Y[k+1] = Y[k] + h_k f(t_k, Y[k]) + g(t_k, Y[k]) dW_k ,   ys[j+1] = w0_j Y[k_j] + w1_j Y[k_j+1]
//
// Two kernels, both recomputing activations from the states Y[k] saved by the forward call:
//   dgrad  : per 32-row tile, reverse sweep over all steps with the adjoint A[k+1] = dL/dY[k+1] resident in registers;
//            A[k] = A[k+1] + h J_f^T A[k+1] + (A[k+1] . dW_k) grad g + sum_j w0_j gy_j ;  writes grad_y0 and every A[k+1].
//   wgrad  : fully parallel over (step, tile) units: recomputes the layer deltas from (Y[k], A[k+1], dW_k), accumulates all
//            weight / bias / time-column gradients in registers across units, writes one partial vector per CTA;
//   reduce : sums the per-CTA partials in a fixed order (deterministic: no float atomics).
// Dual diffusion (encoder): the host runs dgrad+wgrad once per diffusion net with the rows of the other net masked out (rows
// are independent), and one reduce over both partial sets.
#include "bwd_common.cuh"

namespace trajsde {

using namespace bwd;

namespace {

constexpr int BW_ROWS = 32;
constexpr int BW_THREADS = 256;
constexpr int LD = 68;               // padded row stride (floats) of weight / activation tiles
constexpr int WSZ = 64 * LD;         // one staged 64x64 matrix
constexpr int ASZ = BW_ROWS * LD;    // one activation tile

struct BwdParams {
  TrajsdeEulerBwdArgs a;
  TrajsdeMlp g;            // diffusion net of this pass
  int filter;              // 0: all rows; 1: only rows with alt_mask != 0; 2: only rows with alt_mask == 0
  float* adj;              // [S, rows, 64] workspace: A[k+1] of every step
  float* partial;          // [grid, G_PAD] per-CTA partial gradients of this pass (wgrad)
  int num_tiles;
  int accumulate_y0;       // second pass adds into grad_y0 / adj instead of overwriting (disjoint rows => plain write suffices)
};

__device__ __forceinline__ void stage(float* dst, const float* __restrict__ src, int ld, bool transpose, int tid) {
  for (int idx = tid; idx < 64 * 64; idx += BW_THREADS) {
    const int n = idx >> 6, k = idx & 63;
    const float v = src[(size_t)n * ld + k];
    if (transpose) dst[k * LD + n] = v;   // dst[m][n] = W[n][m]
    else dst[n * LD + k] = v;
  }
}

// acc[i][j] += sum_k act[ty*2+i][k] * w[tx+16j][k]
__device__ __forceinline__ void dense32(const float* __restrict__ act, const float* __restrict__ w, int ty, int tx, float (&acc)[2][4]) {
#pragma unroll 4
  for (int k = 0; k < 64; k += 4) {
    float4 av[2], bv[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) av[i] = *reinterpret_cast<const float4*>(act + (ty * 2 + i) * LD + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(w + (tx + 16 * j) * LD + k);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc[i][j] = fmaf(av[i].x, bv[j].x, acc[i][j]);
        acc[i][j] = fmaf(av[i].y, bv[j].y, acc[i][j]);
        acc[i][j] = fmaf(av[i].z, bv[j].z, acc[i][j]);
        acc[i][j] = fmaf(av[i].w, bv[j].w, acc[i][j]);
      }
  }
}

__device__ __forceinline__ void put(float* buf, int ty, int tx, const float (&v)[2][4]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) buf[(ty * 2 + i) * LD + tx + 16 * j] = v[i][j];
}

__device__ __forceinline__ void fill(float (&acc)[2][4], const float* vec, int tx) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float b = vec[tx + 16 * j];
    acc[0][j] = b;
    acc[1][j] = b;
  }
}

__device__ __forceinline__ float noise_elem(const TrajsdeNoise& nz, int64_t rows, int k, int64_t r, int c, float sqrt_h) {
  if (nz.dw) return nz.dw[((int64_t)k * rows + r) * 64 + c];
  const float4 n4 = philox_dw4(ts_noise_seed(nz), (uint64_t)r + nz.row_offset, nz.step_offset + (uint32_t)k, (uint32_t)(c >> 2), sqrt_h);
  return (c & 3) == 0 ? n4.x : (c & 3) == 1 ? n4.y : (c & 3) == 2 ? n4.z : n4.w;
}

// vector slots in smem
enum { S_B1 = 0, S_W1S, S_W1C, S_B2, S_C1, S_V1S, S_V1C, S_C2, S_W3G, S_COUNT };

// Recompute the activations of one step for a 32-row tile whose state is in s_y: h1f,h2f,h1g,h2g -> smem, g -> s_g.
// Weights: w1y, w2, v1y, v2 in nn.Linear orientation.  All threads must call; contains __syncthreads.
__device__ __forceinline__ void recompute(const float* s_y, const float* w1y, const float* w2, const float* v1y, const float* v2,
                                          const float* vec, float sn, float cs, float c3, float* s_h1f, float* s_h2f,
                                          float* s_h1g, float* s_h2g, float* s_g, int ty, int tx) {
  float acc[2][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = tx + 16 * j;
    const float b = fmaf(vec[S_W1C * 64 + n], cs, fmaf(vec[S_W1S * 64 + n], sn, vec[S_B1 * 64 + n]));
    acc[0][j] = b;
    acc[1][j] = b;
  }
  dense32(s_y, w1y, ty, tx, acc);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = tanhf(acc[i][j]);
  put(s_h1f, ty, tx, acc);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = tx + 16 * j;
    const float b = fmaf(vec[S_V1C * 64 + n], cs, fmaf(vec[S_V1S * 64 + n], sn, vec[S_C1 * 64 + n]));
    acc[0][j] = b;
    acc[1][j] = b;
  }
  dense32(s_y, v1y, ty, tx, acc);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = tanhf(acc[i][j]);
  put(s_h1g, ty, tx, acc);
  __syncthreads();
  fill(acc, vec + S_B2 * 64, tx);
  dense32(s_h1f, w2, ty, tx, acc);
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = tanhf(acc[i][j]);
  put(s_h2f, ty, tx, acc);
  fill(acc, vec + S_C2 * 64, tx);
  dense32(s_h1g, v2, ty, tx, acc);
  float part[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    part[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      acc[i][j] = tanhf(acc[i][j]);
      part[i] = fmaf(acc[i][j], vec[S_W3G * 64 + tx + 16 * j], part[i]);
    }
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) part[i] += __shfl_xor_sync(0xffffffffu, part[i], off);
  }
  put(s_h2g, ty, tx, acc);
  if (tx == 0) {
    s_g[ty * 2] = ts_sigmoid_exact(part[0] + c3);
    s_g[ty * 2 + 1] = ts_sigmoid_exact(part[1] + c3);
  }
  __syncthreads();
}

__device__ __forceinline__ void load_vectors(float* vec, const TrajsdeMlp& f, const TrajsdeMlp& g, int tid) {
  if (tid < 64) {
    vec[S_B1 * 64 + tid] = f.b1[tid];
    vec[S_W1S * 64 + tid] = f.w1[tid * TS_IN1 + 64];
    vec[S_W1C * 64 + tid] = f.w1[tid * TS_IN1 + 65];
    vec[S_B2 * 64 + tid] = f.b2[tid];
    vec[S_C1 * 64 + tid] = g.b1[tid];
    vec[S_V1S * 64 + tid] = g.w1[tid * TS_IN1 + 64];
    vec[S_V1C * 64 + tid] = g.w1[tid * TS_IN1 + 65];
    vec[S_C2 * 64 + tid] = g.b2[tid];
    vec[S_W3G * 64 + tid] = g.w3[tid];
  }
}

// ======================================================================================================================
// dgrad: reverse sweep, adjoint in registers
// ======================================================================================================================
struct DgSmem {
  static constexpr int w1y = 0, w2 = WSZ, v1y = 2 * WSZ, v2 = 3 * WSZ;                      // forward orientation
  static constexpr int w3t = 4 * WSZ, w2t = 5 * WSZ, w1yt = 6 * WSZ, v2t = 7 * WSZ, v1yt = 8 * WSZ;  // transposed (dgrad)
  static constexpr int vec = 9 * WSZ;
  static constexpr int y = vec + S_COUNT * 64, h1f = y + ASZ, h2f = h1f + ASZ, h1g = h2f + ASZ, h2g = h1g + ASZ;
  static constexpr int d = h2g + ASZ, d2 = d + ASZ;
  static constexpr int g = d2 + ASZ, ds = g + BW_ROWS;
  static constexpr int total = ds + BW_ROWS;
};

__global__ void __launch_bounds__(BW_THREADS, 1) euler_bwd_dgrad_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float sm[];
  using L = DgSmem;
  const TrajsdeEulerBwdArgs& a = p.a;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  stage(sm + L::w1y, a.drift.w1, TS_IN1, false, tid);
  stage(sm + L::w2, a.drift.w2, 64, false, tid);
  stage(sm + L::v1y, p.g.w1, TS_IN1, false, tid);
  stage(sm + L::v2, p.g.w2, 64, false, tid);
  stage(sm + L::w3t, a.drift.w3, 64, true, tid);
  stage(sm + L::w2t, a.drift.w2, 64, true, tid);
  stage(sm + L::w1yt, a.drift.w1, TS_IN1, true, tid);
  stage(sm + L::v2t, p.g.w2, 64, true, tid);
  stage(sm + L::v1yt, p.g.w1, TS_IN1, true, tid);
  load_vectors(sm + L::vec, a.drift, p.g, tid);
  const float c3 = p.g.b3[0];
  __syncthreads();
  const float* vec = sm + L::vec;
  const int S = a.sched.n_steps;

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * BW_ROWS;
    int64_t r[2];
    bool act[2];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      r[i] = row0 + ty * 2 + i;
      act[i] = r[i] < a.rows;
      if (act[i] && p.filter) act[i] = (a.alt_mask[r[i]] != 0) == (p.filter == 1);
      any |= act[i];
    }
    if (!__syncthreads_or(any)) continue;   // tile has no row of this pass's net
    float adj[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) adj[i][j] = 0.f;

    for (int k = S - 1; k >= 0; --k) {
      const float4 st = *reinterpret_cast<const float4*>(a.sched.step_tab + 4 * k);
      const float h = st.y, sn = st.z, cs = st.w, sqrt_h = sqrtf(h);
      const int ob = a.sched.out_begin[k], oe = a.sched.out_begin[k + 1];
      float dwv[2][4];
      // A[k+1] += sum w1_o gy[o+1] ; publish A[k+1] ; stage Y[k]
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = tx + 16 * j;
          float yv = 0.f;
          dwv[i][j] = 0.f;
          if (act[i]) {
            if (a.grad_ys)
              for (int o = ob; o < oe; ++o)
                adj[i][j] = fmaf(a.sched.out_w[2 * o + 1], a.grad_ys[(int64_t)(o + 1) * a.grad_ys_t_stride + r[i] * a.grad_ys_row_stride + c],
                                 adj[i][j]);
            p.adj[((int64_t)k * a.rows + r[i]) * 64 + c] = adj[i][j];
            yv = a.states[((int64_t)k * a.rows + r[i]) * 64 + c];
            dwv[i][j] = noise_elem(a.noise, a.rows, k, r[i], c, sqrt_h);
          }
          sm[L::y + (ty * 2 + i) * LD + c] = yv;
        }
      }
      __syncthreads();
      recompute(sm + L::y, sm + L::w1y, sm + L::w2, sm + L::v1y, sm + L::v2, vec, sn, cs, c3, sm + L::h1f, sm + L::h2f,
                sm + L::h1g, sm + L::h2g, sm + L::g, ty, tx);
      // ---- diffusion scalar: ds = (A . dW + [k == S-1] gg) * g (1 - g) ------------------------------------------------------
      float q[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        q[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) q[i] = fmaf(adj[i][j], dwv[i][j], q[i]);
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) q[i] += __shfl_xor_sync(0xffffffffu, q[i], off);
      }
      if (tx == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float g = sm[L::g + ty * 2 + i];
          float dg = q[i];
          if (k == S - 1 && a.grad_g_last && act[i]) dg += a.grad_g_last[r[i]];
          sm[L::ds + ty * 2 + i] = act[i] ? dg * g * (1.f - g) : 0.f;
        }
      }
      // ---- drift chain: df = h A -> d ; dz2f -> d2 ; dz1f -> d ; dy_f ---------------------------------------------------------
      float t[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) t[i][j] = h * adj[i][j];
      put(sm + L::d, ty, tx, t);
      __syncthreads();
      float acc[2][4] = {};
      dense32(sm + L::d, sm + L::w3t, ty, tx, acc);                       // dh2f
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hh = sm[L::h2f + (ty * 2 + i) * LD + tx + 16 * j];
          t[i][j] = acc[i][j] * (1.f - hh * hh);                          // dz2f
        }
      put(sm + L::d2, ty, tx, t);
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      dense32(sm + L::d2, sm + L::w2t, ty, tx, acc);                      // dh1f
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hh = sm[L::h1f + (ty * 2 + i) * LD + tx + 16 * j];
          t[i][j] = acc[i][j] * (1.f - hh * hh);                          // dz1f
        }
      __syncthreads();                                                    // everyone done reading d (df)
      put(sm + L::d, ty, tx, t);
      // dz2g = ds w3 (1 - h2g^2) -> d2 (d2's dz2f has been consumed: the sync above follows the dense that read it)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float dsr = sm[L::ds + ty * 2 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hh = sm[L::h2g + (ty * 2 + i) * LD + tx + 16 * j];
          t[i][j] = dsr * vec[S_W3G * 64 + tx + 16 * j] * (1.f - hh * hh);
        }
      }
      put(sm + L::d2, ty, tx, t);
      __syncthreads();
      float dy[2][4] = {};
      dense32(sm + L::d, sm + L::w1yt, ty, tx, dy);                       // dy_f
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      dense32(sm + L::d2, sm + L::v2t, ty, tx, acc);                      // dh1g
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float hh = sm[L::h1g + (ty * 2 + i) * LD + tx + 16 * j];
          t[i][j] = acc[i][j] * (1.f - hh * hh);                          // dz1g
        }
      __syncthreads();                                                    // d (dz1f) consumed
      put(sm + L::d, ty, tx, t);
      __syncthreads();
      dense32(sm + L::d, sm + L::v1yt, ty, tx, dy);                       // + dy_g
      // ---- A[k] ------------------------------------------------------------------------------------------------------------------
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = adj[i][j] + dy[i][j];
          if (act[i] && a.grad_ys) {
            const int c = tx + 16 * j;
            for (int o = ob; o < oe; ++o)
              v = fmaf(a.sched.out_w[2 * o], a.grad_ys[(int64_t)(o + 1) * a.grad_ys_t_stride + r[i] * a.grad_ys_row_stride + c], v);
          }
          adj[i][j] = act[i] ? v : 0.f;
        }
      __syncthreads();                                                    // smem tiles free for the next step
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (act[i])
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = tx + 16 * j;
          float v = adj[i][j];
          if (a.grad_ys) v += a.grad_ys[r[i] * a.grad_ys_row_stride + c];  // ys[0] = y0
          a.grad_y0[r[i] * 64 + c] = v;
        }
  }
}

// ======================================================================================================================
// wgrad: parallel over (step, tile) units
// ======================================================================================================================
struct WgSmem {
  static constexpr int w1y = 0, w2 = WSZ, v1y = 2 * WSZ, v2 = 3 * WSZ;
  static constexpr int w3t = 4 * WSZ, w2t = 5 * WSZ, v2t = 6 * WSZ;
  static constexpr int vec = 7 * WSZ;
  static constexpr int y = vec + S_COUNT * 64, h1f = y + ASZ, h2f = h1f + ASZ, h1g = h2f + ASZ, h2g = h1g + ASZ;
  static constexpr int df = h2g + ASZ, dz2f = df + ASZ, dz1f = dz2f + ASZ, dz2g = dz1f + ASZ, dz1g = dz2g + ASZ;
  static constexpr int g = dz1g + ASZ, ds = g + BW_ROWS;
  static constexpr int total = ds + BW_ROWS;
};

// gW[tn*4+i][tm*4+j] += sum_r d[r][tn*4+i] * x[r][tm*4+j]; colsum[i] += sum_r d[r][tn*4+i]
__device__ __forceinline__ void outer32(const float* __restrict__ d, const float* __restrict__ x, int tn, int tm, float (&gw)[4][4],
                                        float (&colsum)[4]) {
#pragma unroll 4
  for (int r = 0; r < BW_ROWS; ++r) {
    const float4 dv = *reinterpret_cast<const float4*>(d + r * LD + tn * 4);
    const float4 xv = *reinterpret_cast<const float4*>(x + r * LD + tm * 4);
    const float da[4] = {dv.x, dv.y, dv.z, dv.w};
    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      colsum[i] += da[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) gw[i][j] = fmaf(da[i], xa[j], gw[i][j]);
    }
  }
}

__global__ void __launch_bounds__(BW_THREADS, 1) euler_bwd_wgrad_kernel(const BwdParams p) {
  extern __shared__ __align__(16) float sm[];
  using L = WgSmem;
  const TrajsdeEulerBwdArgs& a = p.a;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // dot-style mapping; outer-style uses tn = ty, tm = tx
  stage(sm + L::w1y, a.drift.w1, TS_IN1, false, tid);
  stage(sm + L::w2, a.drift.w2, 64, false, tid);
  stage(sm + L::v1y, p.g.w1, TS_IN1, false, tid);
  stage(sm + L::v2, p.g.w2, 64, false, tid);
  stage(sm + L::w3t, a.drift.w3, 64, true, tid);
  stage(sm + L::w2t, a.drift.w2, 64, true, tid);
  stage(sm + L::v2t, p.g.w2, 64, true, tid);
  load_vectors(sm + L::vec, a.drift, p.g, tid);
  const float c3 = p.g.b3[0];
  __syncthreads();
  const float* vec = sm + L::vec;
  const int S = a.sched.n_steps;

  float gW3[4][4] = {}, gW2[4][4] = {}, gW1[4][4] = {}, gV2[4][4] = {}, gV1[4][4] = {};
  float gb3[4] = {}, gb2[4] = {}, gb1[4] = {}, gw1s[4] = {}, gw1c[4] = {};
  float gc2[4] = {}, gc1[4] = {}, gv1s[4] = {}, gv1c[4] = {};
  float gw3g[4] = {};
  float gc3 = 0.f;

  const int64_t units = (int64_t)p.num_tiles * S;
  for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
    const int tile = (int)(u / S), k = (int)(u % S);
    const int64_t row0 = (int64_t)tile * BW_ROWS;
    int64_t r[2];
    bool act[2];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      r[i] = row0 + ty * 2 + i;
      act[i] = r[i] < a.rows;
      if (act[i] && p.filter) act[i] = (a.alt_mask[r[i]] != 0) == (p.filter == 1);
      any |= act[i];
    }
    if (!__syncthreads_or(any)) continue;
    const float4 st = *reinterpret_cast<const float4*>(a.sched.step_tab + 4 * k);
    const float h = st.y, sn = st.z, cs = st.w, sqrt_h = sqrtf(h);
    float adj[2][4], q[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      q[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tx + 16 * j;
        float yv = 0.f;
        adj[i][j] = 0.f;
        if (act[i]) {
          yv = a.states[((int64_t)k * a.rows + r[i]) * 64 + c];
          adj[i][j] = p.adj[((int64_t)k * a.rows + r[i]) * 64 + c];
          q[i] = fmaf(adj[i][j], noise_elem(a.noise, a.rows, k, r[i], c, sqrt_h), q[i]);
        }
        sm[L::y + (ty * 2 + i) * LD + c] = yv;
        sm[L::df + (ty * 2 + i) * LD + c] = h * adj[i][j];
      }
#pragma unroll
      for (int off = 8; off >= 1; off >>= 1) q[i] += __shfl_xor_sync(0xffffffffu, q[i], off);
    }
    __syncthreads();
    recompute(sm + L::y, sm + L::w1y, sm + L::w2, sm + L::v1y, sm + L::v2, vec, sn, cs, c3, sm + L::h1f, sm + L::h2f, sm + L::h1g,
              sm + L::h2g, sm + L::g, ty, tx);
    if (tx == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float g = sm[L::g + ty * 2 + i];
        float dg = q[i];
        if (k == S - 1 && a.grad_g_last && act[i]) dg += a.grad_g_last[r[i]];
        sm[L::ds + ty * 2 + i] = act[i] ? dg * g * (1.f - g) : 0.f;
      }
    }
    float t[2][4], acc[2][4] = {};
    dense32(sm + L::df, sm + L::w3t, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hh = sm[L::h2f + (ty * 2 + i) * LD + tx + 16 * j];
        t[i][j] = acc[i][j] * (1.f - hh * hh);
      }
    put(sm + L::dz2f, ty, tx, t);
    __syncthreads();   // dz2f + ds visible
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float dsr = sm[L::ds + ty * 2 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hh = sm[L::h2g + (ty * 2 + i) * LD + tx + 16 * j];
        t[i][j] = dsr * vec[S_W3G * 64 + tx + 16 * j] * (1.f - hh * hh);
        acc[i][j] = 0.f;
      }
    }
    put(sm + L::dz2g, ty, tx, t);
    dense32(sm + L::dz2f, sm + L::w2t, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hh = sm[L::h1f + (ty * 2 + i) * LD + tx + 16 * j];
        t[i][j] = acc[i][j] * (1.f - hh * hh);
        acc[i][j] = 0.f;
      }
    put(sm + L::dz1f, ty, tx, t);
    __syncthreads();   // dz2g visible
    dense32(sm + L::dz2g, sm + L::v2t, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hh = sm[L::h1g + (ty * 2 + i) * LD + tx + 16 * j];
        t[i][j] = acc[i][j] * (1.f - hh * hh);
      }
    put(sm + L::dz1g, ty, tx, t);
    __syncthreads();   // all deltas visible
    // ---- accumulate (outer-style mapping: n = ty*4+i, m = tx*4+j) ----------------------------------------------------------------
    float cs3[4] = {}, cs2[4] = {}, cs1[4] = {}, csg2[4] = {}, csg1[4] = {};
    outer32(sm + L::df, sm + L::h2f, ty, tx, gW3, cs3);
    outer32(sm + L::dz2f, sm + L::h1f, ty, tx, gW2, cs2);
    outer32(sm + L::dz1f, sm + L::y, ty, tx, gW1, cs1);
    outer32(sm + L::dz2g, sm + L::h1g, ty, tx, gV2, csg2);
    outer32(sm + L::dz1g, sm + L::y, ty, tx, gV1, csg1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      gb3[i] += cs3[i];                // dL/df = h A = df, so b3 sees the column sums of df
      gb2[i] += cs2[i];
      gb1[i] += cs1[i];
      gw1s[i] = fmaf(sn, cs1[i], gw1s[i]);
      gw1c[i] = fmaf(cs, cs1[i], gw1c[i]);
      gc2[i] += csg2[i];
      gc1[i] += csg1[i];
      gv1s[i] = fmaf(sn, csg1[i], gv1s[i]);
      gv1c[i] = fmaf(cs, csg1[i], gv1c[i]);
    }
    if (ty == 0) {
      for (int r2 = 0; r2 < BW_ROWS; ++r2) {
        const float dsr = sm[L::ds + r2];
        const float4 hv = *reinterpret_cast<const float4*>(sm + L::h2g + r2 * LD + tx * 4);
        gw3g[0] = fmaf(dsr, hv.x, gw3g[0]);
        gw3g[1] = fmaf(dsr, hv.y, gw3g[1]);
        gw3g[2] = fmaf(dsr, hv.z, gw3g[2]);
        gw3g[3] = fmaf(dsr, hv.w, gw3g[3]);
        if (tx == 0) gc3 += dsr;
      }
    }
    __syncthreads();   // tiles free for the next unit
  }

  // ---- write this CTA's partial vector ---------------------------------------------------------------------------------------------
  float* out = p.partial + (size_t)blockIdx.x * G_PAD;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = tx * 4 + j;
      out[G_FW3 + n * 64 + m] = gW3[i][j];
      out[G_FW2 + n * 64 + m] = gW2[i][j];
      out[G_FW1 + n * TS_IN1 + m] = gW1[i][j];
      out[G_GW2 + n * 64 + m] = gV2[i][j];
      out[G_GW1 + n * TS_IN1 + m] = gV1[i][j];
    }
    if (tx == 0) {
      out[G_FB3 + n] = gb3[i];
      out[G_FB2 + n] = gb2[i];
      out[G_FB1 + n] = gb1[i];
      out[G_FW1 + n * TS_IN1 + 64] = gw1s[i];
      out[G_FW1 + n * TS_IN1 + 65] = gw1c[i];
      out[G_GB2 + n] = gc2[i];
      out[G_GB1 + n] = gc1[i];
      out[G_GW1 + n * TS_IN1 + 64] = gv1s[i];
      out[G_GW1 + n * TS_IN1 + 65] = gv1c[i];
    }
  }
  if (ty == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) out[G_GW3 + tx * 4 + j] = gw3g[j];
    if (tx == 0) out[G_GB3] = gc3;
  }
}

int bwd_grid(int64_t rows, int sms) {
  const int64_t tiles = (rows + BW_ROWS - 1) / BW_ROWS;
  return (int)(tiles < sms ? tiles : sms);
}

}  // namespace

int64_t euler_bwd_exact_workspace_bytes(int64_t rows, int32_t n_steps, int32_t dual) {
  (void)dual;
  const int64_t adj = (int64_t)n_steps * rows * 64 * 4;
  const int64_t part = 2ll * 160 * G_PAD * 4;   // up to 160 CTAs per pass, two passes
  return ((adj + 255) & ~255ll) + part + 256;
}

int launch_euler_bwd_exact(const TrajsdeEulerBwdArgs& a, cudaStream_t s) {
  int dev = 0, sms = 0;
  TS_CUDA_CHECK(cudaGetDevice(&dev));
  TS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (sms > 160) sms = 160;
  const bool dual = a.alt_mask != nullptr;
  const int64_t adj_bytes = (((int64_t)a.sched.n_steps * a.rows * 64 * 4) + 255) & ~255ll;
  float* adj = static_cast<float*>(a.workspace);
  float* part = reinterpret_cast<float*>(static_cast<uint8_t*>(a.workspace) + adj_bytes);
  const int num_tiles = (int)((a.rows + BW_ROWS - 1) / BW_ROWS);
  const int grid = num_tiles > 0 ? bwd_grid(a.rows, sms) : 0;
  const size_t dg_smem = (size_t)DgSmem::total * 4, wg_smem = (size_t)WgSmem::total * 4;
  TS_CUDA_CHECK(cudaFuncSetAttribute(euler_bwd_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dg_smem));
  TS_CUDA_CHECK(cudaFuncSetAttribute(euler_bwd_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem));
  const int passes = dual ? 2 : 1;
  int wg_grid = 0;
  if (grid > 0) {
    const int64_t units = (int64_t)num_tiles * a.sched.n_steps;
    wg_grid = (int)(units < sms ? units : sms);
    for (int pass = 0; pass < passes; ++pass) {
      BwdParams p;
      p.a = a;
      p.g = pass == 0 ? a.diffusion : a.diffusion_alt;
      p.filter = dual ? (pass == 0 ? 1 : 2) : 0;
      p.adj = adj;
      p.partial = part + (size_t)pass * 160 * G_PAD;
      p.num_tiles = num_tiles;
      p.accumulate_y0 = 0;
      euler_bwd_dgrad_kernel<<<grid, BW_THREADS, dg_smem, s>>>(p);
      TS_CUDA_CHECK(cudaGetLastError());
      euler_bwd_wgrad_kernel<<<wg_grid, BW_THREADS, wg_smem, s>>>(p);
      TS_CUDA_CHECK(cudaGetLastError());
    }
  }
  euler_bwd_reduce_kernel<<<(G_TOTAL + 255) / 256, 256, 0, s>>>(part, dual ? part + (size_t)160 * G_PAD : nullptr, wg_grid,
                                                              dual ? wg_grid : 0, a.grad_drift, a.grad_diffusion, a.grad_diffusion_alt, 0);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

}  // namespace trajsde
