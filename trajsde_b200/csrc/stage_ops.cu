// The decoder stage's prologue and the two training losses around the SDE solve, each as a fused fp32 kernel:
//
//   aggr_embed   hidden_0 = ReLU(LayerNorm(W [global_embed ; local_embed] + b))            models/decoders/dec_hivt_nusargo_sde.py:26-29, 82-85
//                rows = modes x actors; the reference materialises cat(global, local.expand(modes, ...)) [modes, N, 128] and runs
//                Linear / LayerNorm / ReLU as separate passes — here one pass reads both embeddings and writes y0 [modes*N, 64].
//                Backward: dL/dglobal, dL/dlocal (summed over the modes in a fixed order), dL/dW, db, dgamma, dbeta.
//   pi           mode scores pi[n, m] = w2 . ReLU(LayerNorm(W1 [local_embed[n] ; global_embed[m, n]] + b1)) + b2      dec…sde.py:63-67, 92-94
//                same inputs and first layer shape as aggr_embed (the cat order is swapped): the forward kernel with a 64 -> 1 projection
//                behind it instead of the store.  Forward only — no loss of the reference configuration reads pi (losses L2 + DiffBCE,
//                yml:78-80; pi feeds the test-time metrics), so its rarely needed backward stays with autograd (trajsde_b200/stage.py).
//   L2           winner-takes-all displacement loss                                          losses/L2.py:10-27
//                per actor: best mode = argmin_m mean_t (masked) ||y - loc_m|| ; loss = mean over valid (actor, slot) of the best mode's
//                displacement.  One warp per actor; the backward writes (loc - y) / ||loc - y|| / count for the best mode only.
//   DiffBCE      BCE(diff_in, 0) + BCE(diff_out, 1), mean reduction, log clamped at -100     losses/diff_BCE.py:11-16, enc…sep2.py:194-195
//
// All three are bandwidth-trivial next to the solve (57 MB / 200 MB / KBs at BASELINE configs[1]); they exist to delete the ~25 aten
// launches and the [modes, N, 128] / [modes, N, T] temporaries the reference spends on them.  Reductions use per-block partials and a
// fixed-order final sum (bit-reproducible, no float atomics).
#include "common.cuh"

namespace trajsde {

namespace {

constexpr int SO_THREADS = 256;
constexpr int SO_TILE = 64;
constexpr int LDX = 132;     // padded leading dimension of the [64][128] input tile
constexpr int LDZ = 68;      // padded leading dimension of the [64][64] tiles
constexpr int AG_W = 0, AG_B = 8192, AG_G = 8256, AG_BETA = 8320, AG_N = 8384, AG_PAD = 8384;

__device__ __forceinline__ float4 ld4s(const float* p) { return *reinterpret_cast<const float4*>(p); }

// X tile [64 rows][128]: columns 0..63 = global_embed[row], 64..127 = local_embed[row % n_actors]  (cat order of dec…sde.py:82);
// `rmap` (backward): the tile's rows are entries row0 .. row0 + 63 of a compacted row list
__device__ __forceinline__ void aggr_load_x(const TrajsdeAggrArgs& a, int64_t row0, int64_t rows, float* xs, int tid, const int32_t* rmap = nullptr) {
  const int pt = tid >> 2, qq = tid & 3;                     // 4 threads per row, 32 columns each
  const bool ok = row0 + pt < rows;
  const int64_t r = ok && rmap ? (int64_t)rmap[row0 + pt] : row0 + pt;
  const float* src = qq < 2 ? a.global_embed + r * 64 + 32 * qq : a.local_embed + (r % a.n_actors) * 64 + 32 * (qq - 2);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(xs + pt * LDX + 32 * qq + 4 * i) = ok ? ld4s(src + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
}

// Z[64][64] = X[64][128] W^T + b, thread -> rows 4 pm + pp, channels kq + 16 jj
__device__ __forceinline__ void aggr_gemm_z(const float* xs, const float* ws, const float* vb, float* zs, int tid) {
  const int pm = tid >> 4, kq = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int pp = 0; pp < 4; ++pp)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) acc[pp][jj] = vb[kq + 16 * jj];
#pragma unroll 4
  for (int k = 0; k < 128; k += 4) {
    float4 xv[4], wv[4];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) xv[pp] = ld4s(xs + (4 * pm + pp) * LDX + k);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) wv[jj] = ld4s(ws + (kq + 16 * jj) * LDX + k);
#pragma unroll
    for (int pp = 0; pp < 4; ++pp)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        acc[pp][jj] = fmaf(xv[pp].x, wv[jj].x, fmaf(xv[pp].y, wv[jj].y, fmaf(xv[pp].z, wv[jj].z, fmaf(xv[pp].w, wv[jj].w, acc[pp][jj]))));
  }
#pragma unroll
  for (int pp = 0; pp < 4; ++pp)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) zs[(4 * pm + pp) * LDZ + kq + 16 * jj] = acc[pp][jj];
}

// per-row LayerNorm statistics of this thread's 16 channels (4 threads per row): z <- z-hat, returns rstd
__device__ __forceinline__ float ln_rows16(float (&z)[16], float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += z[i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  const float mean = s * (1.0f / 64.0f);
  float v2 = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    z[i] -= mean;
    v2 = fmaf(z[i], z[i], v2);
  }
  v2 += __shfl_xor_sync(0xffffffffu, v2, 1);
  v2 += __shfl_xor_sync(0xffffffffu, v2, 2);
  const float rstd = 1.0f / sqrtf(v2 * (1.0f / 64.0f) + eps);
#pragma unroll
  for (int i = 0; i < 16; ++i) z[i] *= rstd;
  return rstd;
}

// swap_halves: the weight's input columns are [local | global] (pi head) while the X tile is always staged [global | local]
__device__ __forceinline__ void aggr_stage_weights(const TrajsdeAggrArgs& a, float* ws, float* vecs, int tid, bool swap_halves = false) {
  for (int i = tid; i < 64 * 128; i += SO_THREADS) ws[(i >> 7) * LDX + (swap_halves ? (i & 127) ^ 64 : (i & 127))] = a.w[i];
  for (int i = tid; i < 192; i += SO_THREADS) vecs[i] = i < 64 ? a.b[i] : i < 128 ? a.ln_g[i - 64] : a.ln_b[i - 128];
}

// PI: `a` carries the pi head's first layer; w2 [64], b2 [1] its projection; pi_out [n_actors, n_modes] (the .squeeze(-1).t() of :94)
template <bool PI>
__global__ void __launch_bounds__(SO_THREADS, 2) aggr_embed_fwd_kernel(const TrajsdeAggrArgs a, const float* __restrict__ w2,
                                                                        const float* __restrict__ b2, float* __restrict__ pi_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* ws = reinterpret_cast<float*>(smem);      // [64][LDX]
  float* xs = ws + 64 * LDX;                       // [64][LDX]
  float* zs = xs + SO_TILE * LDX;                  // [64][LDZ]
  float* vecs = zs + SO_TILE * LDZ;                // b | gamma | beta
  const int tid = threadIdx.x;
  const int64_t rows = (int64_t)a.n_modes * a.n_actors;
  aggr_stage_weights(a, ws, vecs, tid, PI);
  const int pt = tid >> 2, qq = tid & 3;
  float w2r[16];
  if (PI) {
#pragma unroll
    for (int i = 0; i < 16; ++i) w2r[i] = w2[16 * qq + i];
  }
  for (int64_t row0 = (int64_t)blockIdx.x * SO_TILE; row0 < rows; row0 += (int64_t)gridDim.x * SO_TILE) {
    __syncthreads();
    aggr_load_x(a, row0, rows, xs, tid);
    __syncthreads();
    aggr_gemm_z(xs, ws, vecs, zs, tid);
    __syncthreads();
    float z[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 v = ld4s(zs + pt * LDZ + 16 * qq + 4 * i);
      z[4 * i] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w;
    }
    ln_rows16(z, a.ln_eps);
    if (PI) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) s = fmaf(fmaxf(fmaf(vecs[64 + 16 * qq + i], z[i], vecs[128 + 16 * qq + i]), 0.f), w2r[i], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      const int64_t r = row0 + pt;
      if (qq == 0 && r < rows) pi_out[(r % a.n_actors) * a.n_modes + r / a.n_actors] = s + b2[0];
      continue;
    }
    if (row0 + pt < rows) {
      float* dst = a.out + (row0 + pt) * 64 + 16 * qq;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 o;
        o.x = fmaxf(fmaf(vecs[64 + 16 * qq + 4 * i], z[4 * i], vecs[128 + 16 * qq + 4 * i]), 0.f);
        o.y = fmaxf(fmaf(vecs[64 + 16 * qq + 4 * i + 1], z[4 * i + 1], vecs[128 + 16 * qq + 4 * i + 1]), 0.f);
        o.z = fmaxf(fmaf(vecs[64 + 16 * qq + 4 * i + 2], z[4 * i + 2], vecs[128 + 16 * qq + 4 * i + 2]), 0.f);
        o.w = fmaxf(fmaf(vecs[64 + 16 * qq + 4 * i + 3], z[4 * i + 3], vecs[128 + 16 * qq + 4 * i + 3]), 0.f);
        *reinterpret_cast<float4*>(dst + 4 * i) = o;
      }
    }
  }
}

// backward: recompute z, LayerNorm / ReLU backward, dX = dZ W (global half -> grad_global rows, local half -> per-row scratch that
// aggr_reduce_modes_kernel sums over the modes), dW += dZ^T X, column sums for db / dgamma / dbeta; one partial vector per block
// flags[r] = dL/dout[r] has a non-zero entry (4 threads per row).  The rows of a decoder batch that received no gradient — nine modes in
// ten under the reference's winner-takes-all L2 — contribute nothing to any result of the backward and are left out of it.
__global__ void aggr_row_flags_kernel(const float* __restrict__ grad_out, int64_t rows, uint8_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = i >> 2;
  bool nz = false;
  if (r < rows) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 v = ld4s(grad_out + r * 64 + 16 * (i & 3) + 4 * q);
      nz = nz || v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, nz);
  if (r < rows && (i & 3) == 0) flags[r] = ((m >> (threadIdx.x & 28)) & 0xFu) ? 1 : 0;
}

__global__ void __launch_bounds__(SO_THREADS, 1) aggr_embed_bwd_kernel(const TrajsdeAggrArgs a, float* __restrict__ local_rows,
                                                                        float* __restrict__ partial, const int32_t* __restrict__ rmap,
                                                                        const int32_t* __restrict__ n_active) {
  extern __shared__ __align__(16) uint8_t smem[];
  float* ws = reinterpret_cast<float*>(smem);
  float* xs = ws + 64 * LDX;
  float* zh = xs + SO_TILE * LDX;                  // z-hat
  float* dr = zh + SO_TILE * LDZ;                  // dL/d(pre-ReLU)
  float* dz = dr + SO_TILE * LDZ;
  float* vecs = dz + SO_TILE * LDZ;
  const int tid = threadIdx.x;
  const int64_t rows = (int64_t)*n_active;             // length of the compacted list rmap[0 .. rows): the rows with a gradient
  aggr_stage_weights(a, ws, vecs, tid);
  const int pt = tid >> 2, qq = tid & 3, pm = tid >> 4, kq = tid & 15;
  float gw[4][8];                                  // dW[pm + 16 jj][4 kq .. +3] and [64 + 4 kq .. +3]
#pragma unroll
  for (int jj = 0; jj < 4; ++jj)
#pragma unroll
    for (int i = 0; i < 8; ++i) gw[jj][i] = 0.f;
  float gcol = 0.f;                                // threads 0..191: db | dgamma | dbeta of channel tid % 64
  for (int64_t row0 = (int64_t)blockIdx.x * SO_TILE; row0 < rows; row0 += (int64_t)gridDim.x * SO_TILE) {
    __syncthreads();
    aggr_load_x(a, row0, rows, xs, tid, rmap);
    __syncthreads();
    aggr_gemm_z(xs, ws, vecs, zh, tid);
    __syncthreads();
    {
      float z[16], dzh[16], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = ld4s(zh + pt * LDZ + 16 * qq + 4 * i);
        z[4 * i] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w;
      }
      const float rstd = ln_rows16(z, a.ln_eps);
      const bool ok = row0 + pt < rows;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = 16 * qq + i;
        const float go = ok ? a.grad_out[(int64_t)rmap[row0 + pt] * 64 + c] : 0.f;
        const float pre = fmaf(vecs[64 + c], z[i], vecs[128 + c]);
        const float drv = pre > 0.f ? go : 0.f;
        dzh[i] = drv * vecs[64 + c];
        s1 += dzh[i];
        s2 = fmaf(dzh[i], z[i], s2);
        zh[pt * LDZ + c] = z[i];
        dr[pt * LDZ + c] = drv;
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
      s1 *= (1.0f / 64.0f);
      s2 *= (1.0f / 64.0f);
#pragma unroll
      for (int i = 0; i < 16; ++i) dz[pt * LDZ + 16 * qq + i] = rstd * (dzh[i] - s1 - z[i] * s2);
    }
    __syncthreads();
    // dX = dZ W : thread -> rows 4 pm + pp, input channels 4 kq .. +3 (global half) and 64 + 4 kq .. +3 (local half)
    {
      float4 ag[4], al[4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) ag[pp] = al[pp] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int j = 0; j < 64; ++j) {
        const float4 wg = ld4s(ws + j * LDX + 4 * kq), wl = ld4s(ws + j * LDX + 64 + 4 * kq);
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const float d = dz[(4 * pm + pp) * LDZ + j];
          ag[pp].x = fmaf(d, wg.x, ag[pp].x); ag[pp].y = fmaf(d, wg.y, ag[pp].y); ag[pp].z = fmaf(d, wg.z, ag[pp].z); ag[pp].w = fmaf(d, wg.w, ag[pp].w);
          al[pp].x = fmaf(d, wl.x, al[pp].x); al[pp].y = fmaf(d, wl.y, al[pp].y); al[pp].z = fmaf(d, wl.z, al[pp].z); al[pp].w = fmaf(d, wl.w, al[pp].w);
        }
      }
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        if (row0 + 4 * pm + pp < rows) {
          const int64_t r = rmap[row0 + 4 * pm + pp];
          *reinterpret_cast<float4*>(a.grad_global + r * 64 + 4 * kq) = ag[pp];
          *reinterpret_cast<float4*>(local_rows + r * 64 + 4 * kq) = al[pp];
        }
      }
    }
    // dW += dZ^T X
#pragma unroll 2
    for (int pp = 0; pp < SO_TILE; ++pp) {
      const float4 xg = ld4s(xs + pp * LDX + 4 * kq), xl = ld4s(xs + pp * LDX + 64 + 4 * kq);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float d = dz[pp * LDZ + pm + 16 * jj];
        gw[jj][0] = fmaf(d, xg.x, gw[jj][0]); gw[jj][1] = fmaf(d, xg.y, gw[jj][1]); gw[jj][2] = fmaf(d, xg.z, gw[jj][2]); gw[jj][3] = fmaf(d, xg.w, gw[jj][3]);
        gw[jj][4] = fmaf(d, xl.x, gw[jj][4]); gw[jj][5] = fmaf(d, xl.y, gw[jj][5]); gw[jj][6] = fmaf(d, xl.z, gw[jj][6]); gw[jj][7] = fmaf(d, xl.w, gw[jj][7]);
      }
    }
    if (tid < 192) {
      const int v = tid >> 6, c = tid & 63;
      float s = 0.f;
      if (v == 0) for (int pp = 0; pp < SO_TILE; ++pp) s += dz[pp * LDZ + c];
      else if (v == 1) for (int pp = 0; pp < SO_TILE; ++pp) s = fmaf(dr[pp * LDZ + c], zh[pp * LDZ + c], s);
      else for (int pp = 0; pp < SO_TILE; ++pp) s += dr[pp * LDZ + c];
      gcol += s;
    }
  }
  float* o = partial + (size_t)blockIdx.x * AG_PAD;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    *reinterpret_cast<float4*>(o + AG_W + (pm + 16 * jj) * 128 + 4 * kq) = make_float4(gw[jj][0], gw[jj][1], gw[jj][2], gw[jj][3]);
    *reinterpret_cast<float4*>(o + AG_W + (pm + 16 * jj) * 128 + 64 + 4 * kq) = make_float4(gw[jj][4], gw[jj][5], gw[jj][6], gw[jj][7]);
  }
  if (tid < 192) o[AG_B + tid] = gcol;
}

__global__ void aggr_reduce_kernel(const float* __restrict__ partial, int n_blocks, TrajsdeAggrArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= AG_N) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partial[(size_t)b * AG_PAD + i];
  if (i < AG_B) a.grad_w[i] = s;
  else if (i < AG_G) a.grad_b[i - AG_B] = s;
  else if (i < AG_BETA) a.grad_ln_g[i - AG_G] = s;
  else a.grad_ln_b[i - AG_BETA] = s;
}

// dL/dlocal[n] = sum over the modes of the per-row local halves, in mode order
__global__ void aggr_reduce_modes_kernel(const float* __restrict__ local_rows, int n_modes, int64_t n_actors, float* __restrict__ grad_local) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_actors * 16) return;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int m = 0; m < n_modes; ++m) {
    const float4 v = ld4s(local_rows + ((int64_t)m * n_actors) * 64 + i * 4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
  }
  *reinterpret_cast<float4*>(grad_local + i * 4) = s;
}

// ---- L2 loss (losses/L2.py:10-27) ---------------------------------------------------------------------------------------------------
// one warp per actor; loc element (m, n, t, c) at loc + ((m * N + n) * T + t) * loc_stride + c ; partial[block] = {sum, count}
__global__ void __launch_bounds__(256) l2_loss_fwd_kernel(const TrajsdeL2Args a, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t n = (int64_t)blockIdx.x * 8 + wib;
  float loss = 0.f, cnt = 0.f;
  if (n < a.n_actors) {
    float best_ade = 0.f;
    int best = 0;
    for (int m = 0; m < a.n_modes; ++m) {
      float s = 0.f;
      for (int t = lane; t < a.n_t; t += 32) {
        if (!a.reg_mask[n * a.n_t + t]) continue;                                    // ade[:, ~reg_mask] = 0   (:17-18)
        const float* p = a.loc + (((int64_t)m * a.n_actors + n) * a.n_t + t) * a.loc_stride;
        const float dx = a.target[(n * a.n_t + t) * 2] - p[0], dy = a.target[(n * a.n_t + t) * 2 + 1] - p[1];
        s += sqrtf(fmaf(dx, dx, dy * dy));
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
      if (m == 0 || s < best_ade) {                                                    // argmin keeps the first minimum   (:19)
        best_ade = s;
        best = m;
      }
    }
    if (lane == 0) a.best_mode[n] = best;
    for (int t = lane; t < a.n_t; t += 32) cnt += a.reg_mask[n * a.n_t + t] ? 1.f : 0.f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
    loss = best_ade;                                                                   // = sum over the valid slots of the best mode's l2
  }
  __shared__ float sl[8], sc[8];
  if (lane == 0) {
    sl[wib] = n < a.n_actors ? loss : 0.f;
    sc[wib] = n < a.n_actors ? cnt : 0.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float l = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) {
      l += sl[i];
      c += sc[i];
    }
    partial[2 * blockIdx.x] = l;
    partial[2 * blockIdx.x + 1] = c;
  }
}

// fixed-order two-level sum of the per-block partials: thread i adds entries i, i + 256, ... ; then a shared-memory tree
__global__ void __launch_bounds__(256) l2_loss_finish_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ loss,
                                                             float* __restrict__ count) {
  __shared__ double sl[256], sc[256];
  double l = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < n_blocks; i += 256) {
    l += partial[2 * i];
    c += partial[2 * i + 1];
  }
  sl[threadIdx.x] = l;
  sc[threadIdx.x] = c;
  __syncthreads();
  for (int off = 128; off >= 1; off >>= 1) {
    if ((int)threadIdx.x < off) {
      sl[threadIdx.x] += sl[threadIdx.x + off];
      sc[threadIdx.x] += sc[threadIdx.x + off];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *count = (float)sc[0];
    *loss = sc[0] > 0.0 ? (float)(sl[0] / sc[0]) : 0.f;                               // reg_mask.sum() == 0 -> 0   (:22-27)
  }
}

// dL/dloc of the best mode: (loc - y) / ||loc - y|| * grad_loss / count on valid slots; every other entry stays zero (caller zero-fills)
__global__ void __launch_bounds__(256) l2_loss_bwd_kernel(const TrajsdeL2Args a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n_actors * a.n_t) return;
  if (!a.reg_mask[i]) return;
  const float cnt = *a.count;
  if (!(cnt > 0.f)) return;
  const int64_t n = i / a.n_t, t = i - n * a.n_t;
  const int m = a.best_mode[n];
  const int64_t off = (((int64_t)m * a.n_actors + n) * a.n_t + t) * a.loc_stride;
  const float dx = a.loc[off] - a.target[i * 2], dy = a.loc[off + 1] - a.target[i * 2 + 1];
  const float l2 = sqrtf(fmaf(dx, dx, dy * dy));
  const float sgrad = *a.grad_loss / cnt;
  float* g = a.grad_loc + (((int64_t)m * a.n_actors + n) * a.n_t + t) * a.grad_loc_stride;
  g[0] = l2 > 0.f ? dx / l2 * sgrad : 0.f;                                            // torch.norm backward at 0: subgradient 0
  g[1] = l2 > 0.f ? dy / l2 * sgrad : 0.f;
}

// ---- DiffBCE (losses/diff_BCE.py:11-16): loss = mean(-clamp(log(1 - d_in), -100)) + mean(-clamp(log(d_out), -100)) + its gradients --------
__global__ void __launch_bounds__(256) diff_bce_kernel(const TrajsdeBceArgs a, float* __restrict__ partial) {
  float s = 0.f;
  const float inv_in = a.n_in > 0 ? 1.0f / (float)a.n_in : 0.f, inv_out = a.n_out > 0 ? 1.0f / (float)a.n_out : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n_in + a.n_out; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < a.n_in) {                                          // label 0: -log(1 - d)
      const float d = a.diff_in[i], lg = log1pf(-d);
      s -= fmaxf(lg, -100.f) * inv_in;
      if (a.grad_in) a.grad_in[i] = lg > -100.f ? inv_in / (1.f - d) : 0.f;
    } else {                                                   // label 1: -log(d)
      const float d = a.diff_out[i - a.n_in], lg = logf(d);
      s -= fmaxf(lg, -100.f) * inv_out;
      if (a.grad_out) a.grad_out[i - a.n_in] = lg > -100.f ? -inv_out / d : 0.f;
    }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  __shared__ float sw[8];
  if ((threadIdx.x & 31) == 0) sw[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sw[i];
    partial[blockIdx.x] = t;
  }
}

__global__ void diff_bce_finish_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ loss) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double l = 0.0;
  for (int i = 0; i < n_blocks; ++i) l += partial[i];
  *loss = (float)l;
}

constexpr size_t AGGR_FWD_SMEM = (2 * 64 * LDX + SO_TILE * LDZ + 192) * sizeof(float);
constexpr size_t AGGR_BWD_SMEM = (2 * 64 * LDX + 3 * SO_TILE * LDZ + 192) * sizeof(float);

int sm_count() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return sms;
}

}  // namespace

static int64_t a256(int64_t x) { return (x + 255) & ~(int64_t)255; }

// partials | per-row local halves [rows][64] | n_active (256 B) | row flags [rows] | row map [rows]
int64_t aggr_workspace_bytes(int64_t n_modes, int64_t n_actors) {
  const int64_t rows = n_modes * n_actors;
  return a256((int64_t)sm_count() * AG_PAD * 4) + a256(rows * 64 * 4) + 256 + a256(rows) + a256(rows * 4) + 256;
}

int launch_aggr_embed(const TrajsdeAggrArgs& a, bool backward, cudaStream_t s) {
  const int64_t rows = (int64_t)a.n_modes * a.n_actors;
  const int sms = sm_count();
  if (sms <= 0) return set_error(TRAJSDE_ERR_CUDA, "device attributes unavailable");
  const int64_t tiles = (rows + SO_TILE - 1) / SO_TILE;
  if (!backward) {
    if (rows == 0) return TRAJSDE_OK;
    TS_CUDA_CHECK(cudaFuncSetAttribute(aggr_embed_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AGGR_FWD_SMEM));
    aggr_embed_fwd_kernel<false><<<(int)(tiles < 2 * sms ? tiles : 2 * sms), SO_THREADS, AGGR_FWD_SMEM, s>>>(a, nullptr, nullptr, nullptr);
    TS_CUDA_CHECK(cudaGetLastError());
    return TRAJSDE_OK;
  }
  uint8_t* wsb = static_cast<uint8_t*>(a.workspace);
  float* partial = reinterpret_cast<float*>(wsb);
  float* local_rows = reinterpret_cast<float*>(wsb + a256((int64_t)sms * AG_PAD * 4));
  int32_t* n_active = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(local_rows) + a256(rows * 64 * 4));
  uint8_t* flags = reinterpret_cast<uint8_t*>(n_active) + 256;
  int32_t* rmap = reinterpret_cast<int32_t*>(flags + a256(rows));
  const int grid = (int)(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
  TS_CUDA_CHECK(cudaMemsetAsync(n_active, 0, 4, s));
  if (rows > 0) {
    // rows without a gradient: dL/dglobal = 0, no contribution to dL/dlocal or to the parameter gradients -> only the others are visited
    TS_CUDA_CHECK(cudaMemsetAsync(a.grad_global, 0, sizeof(float) * 64 * (size_t)rows, s));
    TS_CUDA_CHECK(cudaMemsetAsync(local_rows, 0, sizeof(float) * 64 * (size_t)rows, s));
    aggr_row_flags_kernel<<<(int)((rows * 4 + 255) / 256), 256, 0, s>>>(a.grad_out, rows, flags);
    TS_CUDA_CHECK(cudaGetLastError());
    int rc = launch_compact_rows(flags, rows, rmap, n_active, s);
    if (rc != 0) return rc;
  }
  TS_CUDA_CHECK(cudaFuncSetAttribute(aggr_embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AGGR_BWD_SMEM));
  aggr_embed_bwd_kernel<<<grid, SO_THREADS, AGGR_BWD_SMEM, s>>>(a, local_rows, partial, rmap, n_active);
  TS_CUDA_CHECK(cudaGetLastError());
  aggr_reduce_kernel<<<(AG_N + 255) / 256, 256, 0, s>>>(partial, grid, a);
  TS_CUDA_CHECK(cudaGetLastError());
  if (a.n_actors > 0) {
    aggr_reduce_modes_kernel<<<(int)((a.n_actors * 16 + 255) / 256), 256, 0, s>>>(local_rows, a.n_modes, a.n_actors, a.grad_local);
    TS_CUDA_CHECK(cudaGetLastError());
  }
  return TRAJSDE_OK;
}

int launch_pi_head(const TrajsdePiArgs& p, cudaStream_t s) {
  const int64_t rows = (int64_t)p.n_modes * p.n_actors;
  if (rows == 0) return TRAJSDE_OK;
  const int sms = sm_count();
  if (sms <= 0) return set_error(TRAJSDE_ERR_CUDA, "device attributes unavailable");
  TrajsdeAggrArgs a;
  memset(&a, 0, sizeof(a));
  a.n_modes = p.n_modes;
  a.n_actors = p.n_actors;
  a.global_embed = p.global_embed;
  a.local_embed = p.local_embed;
  a.w = p.w1;
  a.b = p.b1;
  a.ln_g = p.ln_g;
  a.ln_b = p.ln_b;
  a.ln_eps = p.ln_eps;
  const int64_t tiles = (rows + SO_TILE - 1) / SO_TILE;
  TS_CUDA_CHECK(cudaFuncSetAttribute(aggr_embed_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AGGR_FWD_SMEM));
  aggr_embed_fwd_kernel<true><<<(int)(tiles < 2 * sms ? tiles : 2 * sms), SO_THREADS, AGGR_FWD_SMEM, s>>>(a, p.w2, p.b2, p.out);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

int64_t l2_workspace_bytes(int64_t n_actors) { return ((n_actors + 7) / 8 + 1) * 8 + 256; }

int launch_l2_loss(const TrajsdeL2Args& a, bool backward, cudaStream_t s) {
  if (!backward) {
    const int blocks = (int)((a.n_actors + 7) / 8);
    float* partial = static_cast<float*>(a.workspace);
    if (blocks > 0) {
      l2_loss_fwd_kernel<<<blocks, 256, 0, s>>>(a, partial);
      TS_CUDA_CHECK(cudaGetLastError());
    }
    l2_loss_finish_kernel<<<1, 256, 0, s>>>(partial, blocks, a.loss, a.count);
    TS_CUDA_CHECK(cudaGetLastError());
    return TRAJSDE_OK;
  }
  const int64_t n = a.n_actors * a.n_t;
  if (n > 0) {
    l2_loss_bwd_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(a);
    TS_CUDA_CHECK(cudaGetLastError());
  }
  return TRAJSDE_OK;
}

int64_t bce_workspace_bytes() { return 1024 * 4 + 256; }

int launch_diff_bce(const TrajsdeBceArgs& a, cudaStream_t s) {
  const int64_t n = a.n_in + a.n_out;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  float* partial = static_cast<float*>(a.workspace);
  if (blocks > 0) {
    diff_bce_kernel<<<blocks, 256, 0, s>>>(a, partial);
    TS_CUDA_CHECK(cudaGetLastError());
  }
  diff_bce_finish_kernel<<<1, 32, 0, s>>>(partial, blocks, a.loss);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

}  // namespace trajsde
