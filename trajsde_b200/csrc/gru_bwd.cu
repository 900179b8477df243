// Backward of one GRU_Unit observation jump of the encoder recurrence (models/utils/ode_utils.py:136-152, called at
// models/encoders/enc_hivt_nusargo_sde_sep2.py:165-169), exact fp32 on CUDA cores:
//     u = sigmoid(U2 tanh(U1 [y1,x] + ub1) + ub2)      r = sigmoid(R2 tanh(R1 [y1,x] + rb1) + rb2)
//     n = N2 tanh(N1 [x, r*y1] + nb1) + nb2            h' = mask ? (1-u) n + u y1 : y1
// Per 32-row tile the kernel recomputes the gates from (y1, x), back-propagates a = dL/dh' to dL/dy1 and dL/dx, and adds the
// tile's weight / bias gradients into the CTA's private partial vector (plain read-modify-write: no float atomics, fixed order,
// bit-reproducible).  The partial persists across the launches of one encoder backward (one launch per loop iteration);
// gru_bwd_reduce_kernel sums the CTAs' partials once at the end.
#include "bwd_common.cuh"

namespace trajsde {

using namespace bwd;

namespace {

constexpr int GB_ROWS = 32;
constexpr int GB_THREADS = 256;
constexpr int LD1 = 132;   // padded row stride of 128-wide matrices / tiles
constexpr int LD2 = 68;    // padded row stride of 64-wide matrices / tiles

// shared memory map (floats)
constexpr int S_U1 = 0, S_R1 = S_U1 + 64 * LD1, S_N1 = S_R1 + 64 * LD1;
constexpr int S_U2 = S_N1 + 64 * LD1, S_R2 = S_U2 + 64 * LD2, S_N2 = S_R2 + 64 * LD2;
constexpr int S_VEC = S_N2 + 64 * LD2;                     // ub1 rb1 nb1 ub2 rb2 nb2
constexpr int S_IN1 = S_VEC + 6 * 64;                      // [32][LD1]: y1 | x
constexpr int S_RY = S_IN1 + GB_ROWS * LD1;                // [32][LD2]: r*y1        (later d_zr)
constexpr int S_TU = S_RY + GB_ROWS * LD2, S_TR = S_TU + GB_ROWS * LD2, S_TN = S_TR + GB_ROWS * LD2;   // TN later d_zu
constexpr int S_B0 = S_TN + GB_ROWS * LD2, S_B1 = S_B0 + GB_ROWS * LD2;
constexpr int S_TOTAL = S_B1 + GB_ROWS * LD2;
static_assert(S_TOTAL * 4 <= 232448, "exceeds 227 KB of shared memory per CTA");

struct GruBwdParams {
  TrajsdeGru w;
  int64_t rows;
  const float* y1;           // [rows,64] state after the SDE step of this iteration (h_cur of the GRU)
  const float* x;            // aa_out base [n_slots, rows, 64]: the GRU input of this iteration is slab slot[iter]
  int64_t x_slab;            // elements per slot slab
  const uint8_t* obs_mask;   // [rows, n_slots] bool
  int64_t obs_mask_row_stride;
  const int32_t* slot;       // device: slot index of this iteration = slot[iter]
  int iter;
  const float* carry;        // [rows,64] dL/dh' flowing back from the next iteration's SDE step, or NULL
  const float* grad_latent;  // [rows,64] dL/d latent[iter] from the loss, or NULL
  float* grad_y1;            // out [rows,64]
  float* grad_x;             // out [rows,64] (grad_aa_out + slot*rows*64 resolved in-kernel), may be NULL
  int64_t grad_x_slab;       // elements per slot slab of grad_x
  float* partial;            // [grid][GRU_G_PAD]
  int num_tiles;
};

// [64][K] row-major global matrix -> padded smem rows, 16 bytes per access (K = 64 << k_shift6)
__device__ __forceinline__ void stage_w(float* dst, const float* __restrict__ src, int k_dim, int ld, int tid) {
  const int q_per_row = k_dim >> 2, sh = k_dim == 128 ? 5 : 4;
  if (reinterpret_cast<uintptr_t>(src) & 15u) {   // parameter living at an odd offset of a flat buffer: scalar copy
    for (int idx = tid; idx < 64 * k_dim; idx += GB_THREADS) dst[(idx >> (sh + 2)) * ld + (idx & (k_dim - 1))] = src[idx];
    return;
  }
  for (int idx = tid; idx < 64 * q_per_row; idx += GB_THREADS) {
    const int n = idx >> sh, k4 = idx & (q_per_row - 1);
    *reinterpret_cast<float4*>(dst + n * ld + 4 * k4) = __ldg(reinterpret_cast<const float4*>(src) + idx);
  }
}

// acc[i][j] += sum_{k<K} act[(ty*2+i)*lda + k] * w[(tx+16j)*ldw + k]          (y = act . W^T, W in nn.Linear orientation)
__device__ __forceinline__ void dense_nk(const float* __restrict__ act, int lda, const float* __restrict__ w, int ldw, int K, int ty, int tx,
                                         float (&acc)[2][4]) {
#pragma unroll 4
  for (int k = 0; k < K; k += 4) {
    float4 av[2], bv[4];
#pragma unroll
    for (int i = 0; i < 2; ++i) av[i] = *reinterpret_cast<const float4*>(act + (ty * 2 + i) * lda + k);
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = *reinterpret_cast<const float4*>(w + (tx + 16 * j) * ldw + k);
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

// acc[i][j] += sum_{k<64} d[(ty*2+i)*LD2 + k] * w[k*ldw + col0 + tx + 16j]      (dx = d . W: back through a Linear layer)
__device__ __forceinline__ void dense_kn(const float* __restrict__ d, const float* __restrict__ w, int ldw, int col0, int ty, int tx,
                                         float (&acc)[2][4]) {
#pragma unroll 2
  for (int k = 0; k < 64; k += 4) {
    float4 dv[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) dv[i] = *reinterpret_cast<const float4*>(d + (ty * 2 + i) * LD2 + k);
    const float da[2][4] = {{dv[0].x, dv[0].y, dv[0].z, dv[0].w}, {dv[1].x, dv[1].y, dv[1].z, dv[1].w}};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float b[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = w[(k + kk) * ldw + col0 + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(da[i][kk], b[j], acc[i][j]);
    }
  }
}

__device__ __forceinline__ void put_tile(float* buf, int ty, int tx, const float (&v)[2][4]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) buf[(ty * 2 + i) * LD2 + tx + 16 * j] = v[i][j];
}

// partial[w_off + n*ldo + col0 + m] += sum_r d[r][n] * x[r*ldx + m]   for n = 4ty..4ty+3, m = 4tx..4tx+3;
// with_bias: partial[b_off + n] += sum_r d[r][n]   (tx == 0 threads)
__device__ __forceinline__ void wgrad_tile(const float* __restrict__ d, const float* __restrict__ x, int ldx, float* __restrict__ out, int w_off,
                                           int ldo, int col0, int b_off, bool with_bias, int ty, int tx) {
  float gw[4][4] = {}, cs[4] = {};
#pragma unroll 4
  for (int r = 0; r < GB_ROWS; ++r) {
    const float4 dv = *reinterpret_cast<const float4*>(d + r * LD2 + ty * 4);
    const float4 xv = *reinterpret_cast<const float4*>(x + r * ldx + tx * 4);
    const float da[4] = {dv.x, dv.y, dv.z, dv.w};
    const float xa[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cs[i] += da[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) gw[i][j] = fmaf(da[i], xa[j], gw[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float4* p = reinterpret_cast<float4*>(out + w_off + (ty * 4 + i) * ldo + col0 + tx * 4);
    float4 v = *p;
    v.x += gw[i][0]; v.y += gw[i][1]; v.z += gw[i][2]; v.w += gw[i][3];
    *p = v;
    if (with_bias && tx == 0) out[b_off + ty * 4 + i] += cs[i];
  }
}

__device__ __forceinline__ float sigmoidf_exact(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(GB_THREADS, 1) gru_bwd_kernel(const GruBwdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  stage_w(sm + S_U1, p.w.u1, 128, LD1, tid);
  stage_w(sm + S_R1, p.w.r1, 128, LD1, tid);
  stage_w(sm + S_N1, p.w.n1, 128, LD1, tid);
  stage_w(sm + S_U2, p.w.u2, 64, LD2, tid);
  stage_w(sm + S_R2, p.w.r2, 64, LD2, tid);
  stage_w(sm + S_N2, p.w.n2, 64, LD2, tid);
  if (tid < 64) {
    sm[S_VEC + tid] = p.w.ub1[tid];
    sm[S_VEC + 64 + tid] = p.w.rb1[tid];
    sm[S_VEC + 128 + tid] = p.w.nb1[tid];
    sm[S_VEC + 192 + tid] = p.w.ub2[tid];
    sm[S_VEC + 256 + tid] = p.w.rb2[tid];
    sm[S_VEC + 320 + tid] = p.w.nb2[tid];
  }
  const int slot = p.slot[p.iter];
  float* gx = p.grad_x ? p.grad_x + (int64_t)slot * p.grad_x_slab : nullptr;
  const float* xin = p.x + (int64_t)slot * p.x_slab;
  float* out = p.partial + (size_t)blockIdx.x * GRU_G_PAD;
  const float* vec = sm + S_VEC;
  __syncthreads();

  for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    const int64_t row0 = (int64_t)tile * GB_ROWS;
    // ---- 1. load this thread's elements: rows ty*2+i, channels tx+16j ------------------------------------------------------
    float y1[2][4], a[2][4];
    bool obs[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = row0 + ty * 2 + i;
      const bool in = r < p.rows;
      obs[i] = in && p.obs_mask[r * p.obs_mask_row_stride + slot] != 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = tx + 16 * j;
        float yv = 0.f, xv = 0.f, av = 0.f;
        if (in) {
          yv = p.y1[r * 64 + c];
          xv = xin[r * 64 + c];
          if (p.carry) av = p.carry[r * 64 + c];
          if (p.grad_latent) av += p.grad_latent[r * 64 + c];
        }
        y1[i][j] = yv;
        a[i][j] = av;
        sm[S_IN1 + (ty * 2 + i) * LD1 + c] = yv;
        sm[S_IN1 + (ty * 2 + i) * LD1 + 64 + c] = xv;
      }
    }
    __syncthreads();
    // ---- 2. tu = tanh(U1 [y1,x] + ub1), tr = tanh(R1 [y1,x] + rb1) ---------------------------------------------------------------
    float tu[2][4], tr[2][4], acc[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = vec[tx + 16 * j];
    dense_nk(sm + S_IN1, LD1, sm + S_U1, LD1, 128, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) tu[i][j] = tanhf(acc[i][j]);
    put_tile(sm + S_TU, ty, tx, tu);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = vec[64 + tx + 16 * j];
    dense_nk(sm + S_IN1, LD1, sm + S_R1, LD1, 128, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) tr[i][j] = tanhf(acc[i][j]);
    put_tile(sm + S_TR, ty, tx, tr);
    __syncthreads();
    // ---- 3. u, r ; r*y1 --------------------------------------------------------------------------------------------------------------------
    float u[2][4], rg[2][4], t[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = vec[192 + tx + 16 * j];
    dense_nk(sm + S_TU, LD2, sm + S_U2, LD2, 64, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) u[i][j] = sigmoidf_exact(acc[i][j]);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = vec[256 + tx + 16 * j];
    dense_nk(sm + S_TR, LD2, sm + S_R2, LD2, 64, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        rg[i][j] = sigmoidf_exact(acc[i][j]);
        t[i][j] = rg[i][j] * y1[i][j];
      }
    put_tile(sm + S_RY, ty, tx, t);
    __syncthreads();
    // ---- 4. tn = tanh(N1 [x, r*y1] + nb1) -------------------------------------------------------------------------------------------------
    float tn[2][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = vec[128 + tx + 16 * j];
    dense_nk(sm + S_IN1 + 64, LD1, sm + S_N1, LD1, 64, ty, tx, acc);
    dense_nk(sm + S_RY, LD2, sm + S_N1 + 64, LD1, 64, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) tn[i][j] = tanhf(acc[i][j]);
    put_tile(sm + S_TN, ty, tx, tn);
    __syncthreads();
    // ---- 5. n ; direct terms: d_n = a (1-u), d_u = a (y1 - n), d_y1 = a u  (a = 0 on unobserved rows, which pass dL/dh' straight to y1) ----
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = vec[320 + tx + 16 * j];
    dense_nk(sm + S_TN, LD2, sm + S_N2, LD2, 64, ty, tx, acc);
    float dy1[2][4], dup[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float ae = obs[i] ? a[i][j] : 0.f;
        t[i][j] = ae * (1.f - u[i][j]);                                  // d_n
        dup[i][j] = ae * (y1[i][j] - acc[i][j]) * u[i][j] * (1.f - u[i][j]);   // d_u' = d_u u (1-u)
        dy1[i][j] = obs[i] ? ae * u[i][j] : a[i][j];
      }
    put_tile(sm + S_B0, ty, tx, t);
    __syncthreads();
    // ---- 6. d_zn = (d_n . N2) (1 - tn^2) -> B1 ; dN2 += d_n^T tn -----------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = 0.f;
    dense_kn(sm + S_B0, sm + S_N2, LD2, 0, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) t[i][j] = acc[i][j] * (1.f - tn[i][j] * tn[i][j]);
    put_tile(sm + S_B1, ty, tx, t);
    wgrad_tile(sm + S_B0, sm + S_TN, LD2, out, GRU_N2, 64, 0, GRU_NB2, true, ty, tx);
    __syncthreads();
    // ---- 7. [d_x | d(r*y1)] = d_zn . N1 ; d_r' ; dN1 += d_zn^T [x, r*y1] ------------------------------------------------------------------------
    float dx[2][4] = {}, drp[2][4];
    dense_kn(sm + S_B1, sm + S_N1, LD1, 0, ty, tx, dx);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = 0.f;
    dense_kn(sm + S_B1, sm + S_N1, LD1, 64, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dy1[i][j] = fmaf(acc[i][j], rg[i][j], dy1[i][j]);
        drp[i][j] = acc[i][j] * y1[i][j] * rg[i][j] * (1.f - rg[i][j]);     // d_r' = d_r r (1-r)
      }
    wgrad_tile(sm + S_B1, sm + S_IN1 + 64, LD1, out, GRU_N1, 128, 0, GRU_NB1, true, ty, tx);
    wgrad_tile(sm + S_B1, sm + S_RY, LD2, out, GRU_N1, 128, 64, 0, false, ty, tx);
    __syncthreads();                                                     // B0, B1, TN, RY are free
    // ---- 8. d_u', d_r' -> B0, B1 --------------------------------------------------------------------------------------------------------------
    put_tile(sm + S_B0, ty, tx, dup);
    put_tile(sm + S_B1, ty, tx, drp);
    __syncthreads();
    // ---- 9. d_zu = (d_u' . U2)(1 - tu^2) -> TN buffer ; d_zr = (d_r' . R2)(1 - tr^2) -> RY buffer ; dU2, dR2 ---------------------------------------
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = 0.f;
    dense_kn(sm + S_B0, sm + S_U2, LD2, 0, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) t[i][j] = acc[i][j] * (1.f - tu[i][j] * tu[i][j]);
    put_tile(sm + S_TN, ty, tx, t);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = 0.f;
    dense_kn(sm + S_B1, sm + S_R2, LD2, 0, ty, tx, acc);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) t[i][j] = acc[i][j] * (1.f - tr[i][j] * tr[i][j]);
    put_tile(sm + S_RY, ty, tx, t);
    wgrad_tile(sm + S_B0, sm + S_TU, LD2, out, GRU_U2, 64, 0, GRU_UB2, true, ty, tx);
    wgrad_tile(sm + S_B1, sm + S_TR, LD2, out, GRU_R2, 64, 0, GRU_RB2, true, ty, tx);
    __syncthreads();
    // ---- 10. [d_y1 | d_x] += d_zu . U1 + d_zr . R1 ; dU1, dR1 ; outputs --------------------------------------------------------------------------
    dense_kn(sm + S_TN, sm + S_U1, LD1, 0, ty, tx, dy1);
    dense_kn(sm + S_RY, sm + S_R1, LD1, 0, ty, tx, dy1);
    dense_kn(sm + S_TN, sm + S_U1, LD1, 64, ty, tx, dx);
    dense_kn(sm + S_RY, sm + S_R1, LD1, 64, ty, tx, dx);
    wgrad_tile(sm + S_TN, sm + S_IN1, LD1, out, GRU_U1, 128, 0, GRU_UB1, true, ty, tx);
    wgrad_tile(sm + S_TN, sm + S_IN1 + 64, LD1, out, GRU_U1, 128, 64, 0, false, ty, tx);
    wgrad_tile(sm + S_RY, sm + S_IN1, LD1, out, GRU_R1, 128, 0, GRU_RB1, true, ty, tx);
    wgrad_tile(sm + S_RY, sm + S_IN1 + 64, LD1, out, GRU_R1, 128, 64, 0, false, ty, tx);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = row0 + ty * 2 + i;
      if (r < p.rows) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = tx + 16 * j;
          p.grad_y1[r * 64 + c] = dy1[i][j];
          if (gx) gx[r * 64 + c] = dx[i][j];
        }
      }
    }
    __syncthreads();                                                     // tiles free for the next row tile
  }
}

// grads = sum over CTAs (fixed order) of the partial vectors
__global__ void gru_bwd_reduce_kernel(const float* __restrict__ part, int n, TrajsdeGruGrad g) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= GRU_G_TOTAL) return;
  float s = 0.f;
  for (int c = 0; c < n; ++c) s += part[(size_t)c * GRU_G_PAD + i];
  const int gate = i / GRU_GATE, o = i - gate * GRU_GATE;
  float* w1 = gate == 0 ? g.u1 : gate == 1 ? g.r1 : g.n1;
  float* b1 = gate == 0 ? g.ub1 : gate == 1 ? g.rb1 : g.nb1;
  float* w2 = gate == 0 ? g.u2 : gate == 1 ? g.r2 : g.n2;
  float* b2 = gate == 0 ? g.ub2 : gate == 1 ? g.rb2 : g.nb2;
  if (o < 8192) w1[o] = s;
  else if (o < 8256) b1[o - 8192] = s;
  else if (o < 12352) w2[o - 8256] = s;
  else b2[o - 12352] = s;
}

}  // namespace

int gru_bwd_grid(int64_t rows) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (sms > MAX_PARTIALS) sms = MAX_PARTIALS;
  const int64_t tiles = (rows + GB_ROWS - 1) / GB_ROWS;
  return (int)(tiles < sms ? tiles : sms);
}

int launch_gru_bwd(const TrajsdeGru& w, int64_t rows, const float* y1, const float* aa_out, int64_t slab, const uint8_t* obs_mask,
                   int64_t obs_mask_row_stride, const int32_t* slot, int iter, const float* carry, const float* grad_latent, float* grad_y1,
                   float* grad_aa_out, float* partial, cudaStream_t s) {
  GruBwdParams p;
  p.w = w;
  p.rows = rows;
  p.y1 = y1;
  p.x = aa_out;
  p.x_slab = slab;
  p.obs_mask = obs_mask;
  p.obs_mask_row_stride = obs_mask_row_stride;
  p.slot = slot;
  p.iter = iter;
  p.carry = carry;
  p.grad_latent = grad_latent;
  p.grad_y1 = grad_y1;
  p.grad_x = grad_aa_out;
  p.grad_x_slab = slab;
  p.partial = partial;
  p.num_tiles = (int)((rows + GB_ROWS - 1) / GB_ROWS);
  const int grid = gru_bwd_grid(rows);
  if (grid <= 0) return TRAJSDE_OK;
  TS_CUDA_CHECK(cudaFuncSetAttribute(gru_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_TOTAL * 4));
  gru_bwd_kernel<<<grid, GB_THREADS, S_TOTAL * 4, s>>>(p);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

int launch_gru_bwd_reduce(const float* partial, int n, const TrajsdeGruGrad& g, cudaStream_t s) {
  gru_bwd_reduce_kernel<<<(GRU_G_TOTAL + 255) / 256, 256, 0, s>>>(partial, n, g);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

}  // namespace trajsde
