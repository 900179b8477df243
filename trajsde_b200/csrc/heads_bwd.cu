// Backward of the decoder heads (training): what torch.autograd computes through
//     out_h = W2_h . relu(LayerNorm(W1_h x + b1_h)) + b2_h        h = 0 (self.decoder), 1 (self.scale)
// of SDEDecoder.forward (models/decoders/dec_hivt_nusargo_sde.py:50-61, 96, 98) for every (row, t) latent x: dL/dx (summed over the
// heads) and the gradients of every head parameter, from dL/dout_h [rows, n_t, 2].
//
// The reference's training losses make that gradient SPARSE: L2 (losses/L2.py:12-20) looks at `loc` only — the scale head receives no
// gradient at all — and only at the best of the 10 modes of an actor, on its valid future slots; so ~5 % of the (point, head) pairs
// carry a non-zero dL/dout.  The kernel therefore scans dL/dout (16 B per point), compacts the active points of each head in shared
// memory (ballot + prefix), and runs the dense work only on tiles of 64 ACTIVE points.  Points without a gradient cost 16 B of traffic;
// their dL/dx stays whatever the caller put there (zeros).  Exact fp32 arithmetic on the CUDA cores (FFMA, libm rsqrt): this is the
// validation-grade twin of the tensor-core forward (heads.cu), fast because of the sparsity, not because of the pipe.
//
//   per tile of 64 points and head:   Z = X W1^T + b1        (register-blocked 4 points x 4 channels per thread)
//                                     LayerNorm / ReLU / 64 -> 2 projection backward per point (4 threads per point, shuffles)
//                                     dX += dZ W1            dW1 += dZ^T X (16 accumulators per thread, kept across all tiles)
//                                     column sums over the tile's points for db1, dgamma, dbeta, dW2 (thread = (vector, channel))
//   per block: one partial vector [2 heads][4418]; a fixed-order reduce over the blocks writes the gradients (bit-reproducible).
#include "common.cuh"

namespace trajsde {

namespace {

constexpr int HB_THREADS = 256;
constexpr int HB_TILE = 64;                 // active points per tile
constexpr int HB_LD = 68;                   // padded leading dimension (floats) of the 64-wide shared-memory tiles
constexpr int HG_W1 = 0, HG_B1 = 4096, HG_G = 4160, HG_BETA = 4224, HG_W2 = 4288, HG_B2 = 4416, HG_N = 4418, HG_PAD = 4420;

struct HeadsBwdParams {
  TrajsdeHeadsBwdArgs a;
  float* partial;        // [grid][2][HG_PAD]
  int64_t n_points;      // rows * n_t
  int gstride;           // floats between consecutive points of a.grad_out[h]: 2, or 4 in CAT4 mode (a.grad_out[1] = a.grad_out[0] + 2 then)
  int elu_head1;         // CAT4: head 1's incoming gradient goes through the derivative of elu(raw) + 1 + min_scale first
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__global__ void __launch_bounds__(HB_THREADS, 2) heads_bwd_kernel(const HeadsBwdParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const TrajsdeHeadsBwdArgs& a = p.a;
  float* w1s = reinterpret_cast<float*>(smem);                 // [2 heads][64][HB_LD]  W1 row-major (j, k)
  float* xs = w1s + 2 * 64 * HB_LD;                            // [64 points][HB_LD]
  float* zh = xs + HB_TILE * HB_LD;                            // z-hat (normalised pre-activation), later unused
  float* dr = zh + HB_TILE * HB_LD;                            // dL/d(pre-ReLU)
  float* dz = dr + HB_TILE * HB_LD;                            // dL/dz
  float* vecs = dz + HB_TILE * HB_LD;                          // per head: b1[64] g[64] beta[64] w2[128] -> 320 floats
  float* dsc = vecs + 2 * 320;                                 // [64 points][2]: dL/dout of the tile's points
  int64_t* qpt = reinterpret_cast<int64_t*>(dsc + 2 * HB_TILE);   // [2 heads][HB_TILE + HB_THREADS] queued active point ids
  int* qn = reinterpret_cast<int*>(qpt + 2 * (HB_TILE + HB_THREADS));   // [2] queue lengths
  int* wsum = qn + 2;                                          // [8] per-warp counts
  int64_t* rlist = reinterpret_cast<int64_t*>(wsum + 14);      // [HB_THREADS] flagged rows of the current 256-row chunk (row_flags mode)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_heads = a.n_heads;
  for (int i = tid; i < 2 * 64 * 64; i += HB_THREADS) {
    const int h = i >> 12, j = (i >> 6) & 63, k = i & 63;
    w1s[(h * 64 + j) * HB_LD + k] = h < n_heads ? a.head[h].w1[j * 64 + k] : 0.f;
  }
  for (int i = tid; i < 2 * 320; i += HB_THREADS) {
    const int h = i / 320, r = i % 320;
    float v = 0.f;
    if (h < n_heads) v = r < 64 ? a.head[h].b1[r] : r < 128 ? a.head[h].ln_g[r - 64] : r < 192 ? a.head[h].ln_b[r - 128] : a.head[h].w2[r - 192];
    vecs[i] = v;
  }
  if (tid < 2) qn[tid] = 0;
  __syncthreads();

  // accumulators kept across all tiles of this block
  float gw1[2][16];                    // dW1[j0 + 16 jj][k0 .. k0+3], j0 = tid / 16, k0 = 4 (tid % 16)      per head
  float gcol[2] = {0.f, 0.f};          // thread = (vector v = tid / 64, channel c = tid % 64): v = 0 db1, 1 dgamma, 2 dbeta, 3 dW2[0]
  float gcol2[2] = {0.f, 0.f};         // v == 3 threads also carry dW2[1][c]; v == 0 threads of c < 2 carry db2[c]
  float gx_peak = 0.f;                 // largest |value| this thread has written to grad_x (-> a.grad_amax: the solver backward's loss scale)
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 16; ++i) gw1[h][i] = 0.f;

  const int pm = tid >> 4, kq = tid & 15;                    // GEMM mapping: points 4 pm .. 4 pm + 3 ; channels kq + 16 jj (B) / 4 kq .. 4 kq + 3 (D, E)
  const int pt = tid >> 2, qq = tid & 3;                     // per-point mapping: point pt, channels 16 qq .. 16 qq + 15

  auto process = [&](int h, int n) {                         // dense backward of the first n (<= 64) queued points of head h
    const float* W = w1s + h * 64 * HB_LD;
    const float* vb1 = vecs + h * 320, *vg = vb1 + 64, *vbeta = vb1 + 128, *vw2 = vb1 + 192;
    const int64_t* q = qpt + h * (HB_TILE + HB_THREADS);
    // ---- X tile (gather: one 256-byte row per point, 4 threads x 64 B) + the points' dL/dout ------------------------------------
    {
      const bool ok = pt < n;
      const int64_t pid = ok ? q[pt] : 0;
      const int64_t r = pid / a.n_t, t = pid - r * a.n_t;
      const float* src = a.x + r * a.x_row_stride + t * a.x_t_stride + 16 * qq;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(xs + pt * HB_LD + 16 * qq + 4 * i) = ok ? ld4(src + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (qq == 0) {
        const float2 d = ok ? *reinterpret_cast<const float2*>(a.grad_out[h] + pid * p.gstride) : make_float2(0.f, 0.f);
        dsc[2 * pt] = d.x;
        dsc[2 * pt + 1] = d.y;
      }
    }
    __syncthreads();
    // ---- Z = X W1^T + b1 : thread -> points 4 pm + pp, channels kq + 16 jj ------------------------------------------------------------
    {
      float acc[4][4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) acc[pp][jj] = vb1[kq + 16 * jj];
#pragma unroll 4
      for (int k = 0; k < 64; k += 4) {
        float4 xv[4], wv[4];
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) xv[pp] = ld4(xs + (4 * pm + pp) * HB_LD + k);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) wv[jj] = ld4(W + (kq + 16 * jj) * HB_LD + k);
#pragma unroll
        for (int pp = 0; pp < 4; ++pp)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            acc[pp][jj] = fmaf(xv[pp].x, wv[jj].x, fmaf(xv[pp].y, wv[jj].y, fmaf(xv[pp].z, wv[jj].z, fmaf(xv[pp].w, wv[jj].w, acc[pp][jj]))));
      }
#pragma unroll
      for (int pp = 0; pp < 4; ++pp)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) zh[(4 * pm + pp) * HB_LD + kq + 16 * jj] = acc[pp][jj];
    }
    __syncthreads();
    // ---- per point: LayerNorm forward, ReLU, projection backward, LayerNorm backward (4 threads per point) ------------------------
    {
      float z[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = ld4(zh + pt * HB_LD + 16 * qq + 4 * i);
        z[4 * i] = v.x; z[4 * i + 1] = v.y; z[4 * i + 2] = v.z; z[4 * i + 3] = v.w;
      }
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
      const float rstd = 1.0f / sqrtf(v2 * (1.0f / 64.0f) + a.ln_eps);
      float d0 = dsc[2 * pt], d1 = dsc[2 * pt + 1];
      if (p.elu_head1 && h == 1) {                             // block-uniform.  d/draw [elu(raw) + 1 + min_scale] = raw > 0 ? 1 : exp(raw)
        float r0 = 0.f, r1 = 0.f;                              // raw = W2 relu(LayerNorm(z)) + b2, recomputed (4 threads per point)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = 16 * qq + i;
          const float actv = fmaxf(fmaf(vg[c], z[i] * rstd, vbeta[c]), 0.f);
          r0 = fmaf(vw2[c], actv, r0);
          r1 = fmaf(vw2[64 + c], actv, r1);
        }
        r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
        r0 += __shfl_xor_sync(0xffffffffu, r0, 2);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 2);
        r0 += a.head[1].b2[0];
        r1 += a.head[1].b2[1];
        d0 *= r0 > 0.f ? 1.f : expf(r0);
        d1 *= r1 > 0.f ? 1.f : expf(r1);
        __syncwarp();                                          // the point's four threads (same warp) have read the incoming gradient
        if (qq == 0) {                                         // the dW2 / db2 sums below use dL/draw
          dsc[2 * pt] = d0;
          dsc[2 * pt + 1] = d1;
        }
      }
      float dzh[16], s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = 16 * qq + i;
        z[i] *= rstd;                                                     // z-hat
        const float pre = fmaf(vg[c], z[i], vbeta[c]);
        const float da = fmaf(vw2[c], d0, vw2[64 + c] * d1);
        const float drv = pre > 0.f ? da : 0.f;
        dzh[i] = drv * vg[c];
        s1 += dzh[i];
        s2 = fmaf(dzh[i], z[i], s2);
        zh[pt * HB_LD + c] = z[i];
        dr[pt * HB_LD + c] = drv;
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
      s1 *= (1.0f / 64.0f);
      s2 *= (1.0f / 64.0f);
#pragma unroll
      for (int i = 0; i < 16; ++i) dz[pt * HB_LD + 16 * qq + i] = rstd * (dzh[i] - s1 - z[i] * s2);
    }
    __syncthreads();
    // ---- dX += dZ W1 : thread -> points 4 pm + pp, input channels 4 kq .. 4 kq + 3 ; read-modify-write of the caller's dL/dx rows --------
    {
      float4 acc[4];
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) acc[pp] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int j = 0; j < 64; ++j) {
        const float4 wv = ld4(W + j * HB_LD + 4 * kq);
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
          const float d = dz[(4 * pm + pp) * HB_LD + j];
          acc[pp].x = fmaf(d, wv.x, acc[pp].x); acc[pp].y = fmaf(d, wv.y, acc[pp].y);
          acc[pp].z = fmaf(d, wv.z, acc[pp].z); acc[pp].w = fmaf(d, wv.w, acc[pp].w);
        }
      }
#pragma unroll
      for (int pp = 0; pp < 4; ++pp) {
        const int pi = 4 * pm + pp;
        if (pi < n) {
          const int64_t pid = q[pi];
          const int64_t r = pid / a.n_t, t = pid - r * a.n_t;
          float4* dst = reinterpret_cast<float4*>(a.grad_x + r * a.gx_row_stride + t * a.gx_t_stride + 4 * kq);
          float4 o = *dst;                                               // a point may be active in both heads: accumulate
          o.x += acc[pp].x; o.y += acc[pp].y; o.z += acc[pp].z; o.w += acc[pp].w;
          *dst = o;
          gx_peak = fmaxf(gx_peak, fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w))));
        }
      }
    }
    // ---- dW1 += dZ^T X : thread -> output channels pm + 16 jj, input channels 4 kq .. 4 kq + 3 -----------------------------------------
#pragma unroll 4
    for (int pp = 0; pp < HB_TILE; ++pp) {
      const float4 xv = ld4(xs + pp * HB_LD + 4 * kq);
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float d = dz[pp * HB_LD + pm + 16 * jj];
        gw1[h][4 * jj] = fmaf(d, xv.x, gw1[h][4 * jj]); gw1[h][4 * jj + 1] = fmaf(d, xv.y, gw1[h][4 * jj + 1]);
        gw1[h][4 * jj + 2] = fmaf(d, xv.z, gw1[h][4 * jj + 2]); gw1[h][4 * jj + 3] = fmaf(d, xv.w, gw1[h][4 * jj + 3]);
      }
    }
    // ---- column sums over the tile's points: thread = (vector tid / 64, channel tid % 64) ------------------------------------------------
    {
      const int v = tid >> 6, c = tid & 63;
      float s = 0.f, s2 = 0.f;
      if (v == 0) {
        for (int pp = 0; pp < HB_TILE; ++pp) s += dz[pp * HB_LD + c];                                 // db1
        if (c < 2) for (int pp = 0; pp < HB_TILE; ++pp) s2 += dsc[2 * pp + c];                       // db2
      } else if (v == 1) {
        for (int pp = 0; pp < HB_TILE; ++pp) s = fmaf(dr[pp * HB_LD + c], zh[pp * HB_LD + c], s);      // dgamma
      } else if (v == 2) {
        for (int pp = 0; pp < HB_TILE; ++pp) s += dr[pp * HB_LD + c];                                 // dbeta
      } else {
        const float g = vg[c], b = vbeta[c];
        for (int pp = 0; pp < HB_TILE; ++pp) {                                                       // dW2[0][c], dW2[1][c]
          const float act = fmaxf(fmaf(g, zh[pp * HB_LD + c], b), 0.f);
          s = fmaf(dsc[2 * pp], act, s);
          s2 = fmaf(dsc[2 * pp + 1], act, s2);
        }
      }
      gcol[h] += s;
      gcol2[h] += s2;
    }
    __syncthreads();                                           // tiles and queue entries are free again
  };

  // ---- scan this block's share of the points, compact the active ones per head, process full tiles as they form -----------------------------
  // one step: 256 candidate points (has / pid per thread) -> per head: ballot-compact the active ones into the queue, drain full tiles
  auto scan256 = [&](bool has, int64_t pid) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h >= n_heads || !a.grad_out[h]) continue;
      bool act = false;
      if (has) {
        const float2 d = *reinterpret_cast<const float2*>(a.grad_out[h] + pid * p.gstride);
        act = d.x != 0.f || d.y != 0.f;
      }
      const unsigned m = __ballot_sync(0xffffffffu, act);
      if (lane == 0) wsum[warp] = __popc(m);
      __syncthreads();
      int off = qn[h], tot = 0;
      for (int w = 0; w < HB_THREADS / 32; ++w) {
        if (w < warp) off += wsum[w];
        tot += wsum[w];
      }
      if (act) qpt[h * (HB_TILE + HB_THREADS) + off + __popc(m & ((1u << lane) - 1u))] = pid;
      __syncthreads();
      int n = qn[h] + tot;
      int done = 0;
      while (n - done >= HB_TILE) {                            // warp-uniform: every thread sees the same counts
        if (done > 0) {                                        // shift the queue so the tile starts at entry 0
          int64_t keep = 0;
          const bool mv = tid < n - done;
          if (mv) keep = qpt[h * (HB_TILE + HB_THREADS) + done + tid];
          __syncthreads();
          if (mv) qpt[h * (HB_TILE + HB_THREADS) + tid] = keep;
          __syncthreads();
          n -= done;
          done = 0;
        }
        process(h, HB_TILE);
        done = HB_TILE;
      }
      if (done > 0) {
        int64_t keep = 0;
        const bool mv = tid < n - done;
        if (mv) keep = qpt[h * (HB_TILE + HB_THREADS) + done + tid];
        __syncthreads();
        if (mv) qpt[h * (HB_TILE + HB_THREADS) + tid] = keep;
        n -= done;
      }
      __syncthreads();
      if (tid == 0) qn[h] = n;
      __syncthreads();
    }
  };
  if (a.row_flags) {
    // the caller's row flags are known (heads_row_flags_kernel ran before): walk this block's contiguous range of ROWS 256 at a time, compact
    // the flagged ones and scan only their points — under a winner-takes-all loss nine rows in ten are skipped without reading a gradient
    const int64_t rper = (a.rows + gridDim.x - 1) / gridDim.x;
    const int64_t r_lo = (int64_t)blockIdx.x * rper, r_hi = r_lo + rper < a.rows ? r_lo + rper : a.rows;
    for (int64_t rc = r_lo; rc < r_hi; rc += HB_THREADS) {
      const int64_t r = rc + tid;
      const bool f = r < r_hi && a.row_flags[r] != 0;
      const unsigned m = __ballot_sync(0xffffffffu, f);
      __syncthreads();                                         // previous chunk's row list / wsum are no longer read
      if (lane == 0) wsum[warp] = __popc(m);
      __syncthreads();
      int off = 0, nr = 0;
      for (int w = 0; w < HB_THREADS / 32; ++w) {
        if (w < warp) off += wsum[w];
        nr += wsum[w];
      }
      if (f) rlist[off + __popc(m & ((1u << lane) - 1u))] = r;
      __syncthreads();
      const int64_t nv = (int64_t)nr * a.n_t;
      for (int64_t vb = 0; vb < nv; vb += HB_THREADS) {
        const int64_t vp = vb + tid;
        const bool has = vp < nv;
        int64_t pid = 0;
        if (has) {
          const int64_t li = vp / a.n_t;
          pid = rlist[li] * a.n_t + (vp - li * a.n_t);
        }
        scan256(has, pid);
      }
    }
  } else {
    const int64_t per = (p.n_points + gridDim.x - 1) / gridDim.x;
    const int64_t p_lo = (int64_t)blockIdx.x * per, p_hi = p_lo + per < p.n_points ? p_lo + per : p.n_points;
    for (int64_t base = p_lo; base < p_hi; base += HB_THREADS) scan256(base + tid < p_hi, base + tid);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h >= n_heads || !a.grad_out[h]) continue;
    const int n = qn[h];
    if (n > 0) process(h, n);                                  // the remainder (< 64 points; the tile is padded with inert points)
  }

  if (a.grad_amax) {                                           // non-negative floats order like their bit patterns
    if (!(gx_peak <= 3.0e38f)) gx_peak = 3.0e38f;              // inf / nan: clamp (the result is garbage either way)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) gx_peak = fmaxf(gx_peak, __shfl_xor_sync(0xffffffffu, gx_peak, off));
    if (lane == 0 && gx_peak > 0.f) atomicMax(reinterpret_cast<unsigned int*>(a.grad_amax), __float_as_uint(gx_peak));
  }
  // ---- this block's partial vector -----------------------------------------------------------------------------------------------------
  float* out = p.partial + (size_t)blockIdx.x * 2 * HG_PAD;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float* o = out + h * HG_PAD;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
      *reinterpret_cast<float4*>(o + HG_W1 + (pm + 16 * jj) * 64 + 4 * kq) = make_float4(gw1[h][4 * jj], gw1[h][4 * jj + 1], gw1[h][4 * jj + 2], gw1[h][4 * jj + 3]);
    const int v = tid >> 6, c = tid & 63;
    if (v == 0) {
      o[HG_B1 + c] = gcol[h];
      if (c < 2) o[HG_B2 + c] = gcol2[h];
    } else if (v == 1) {
      o[HG_G + c] = gcol[h];
    } else if (v == 2) {
      o[HG_BETA + c] = gcol[h];
    } else {
      o[HG_W2 + c] = gcol[h];
      o[HG_W2 + 64 + c] = gcol2[h];
    }
  }
}

// row_flags mode, pass 1: flags[r] = 1 if any point of row r carries a gradient in any head (flags zeroed by a memset before)
// Four independent points per thread and iteration: the scan is a pure stream over dL/dout (196 MB at BASELINE configs[1]) and needs the
// loads in flight, not arithmetic (one point per iteration: 171 us; this form: HBM speed).
__global__ void heads_row_flags_kernel(const TrajsdeHeadsBwdArgs a, int64_t n_points, int gstride) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t pid0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pid0 < n_points; pid0 += 4 * stride) {
    float2 d[4][2];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t pid = pid0 + k * stride;
#pragma unroll
      for (int h = 0; h < 2; ++h)
        d[k][h] = (pid < n_points && h < a.n_heads && a.grad_out[h]) ? *reinterpret_cast<const float2*>(a.grad_out[h] + pid * gstride)
                                                                      : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool act = d[k][0].x != 0.f || d[k][0].y != 0.f || d[k][1].x != 0.f || d[k][1].y != 0.f;
      if (act) a.row_flags[(pid0 + k * stride) / a.n_t] = 1;   // benign race: every writer stores the same value
    }
  }
}

// pass 2: zero the n_t x 64 gradient entries of the flagged rows (one warp per row; the other rows stay untouched)
__global__ void heads_zero_rows_kernel(const TrajsdeHeadsBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < a.rows; r += nwarps) {
    if (!a.row_flags[r]) continue;
    float* base = a.grad_x + r * a.gx_row_stride;
    for (int i = lane; i < a.n_t * 16; i += 32)
      *reinterpret_cast<float4*>(base + (int64_t)(i >> 4) * a.gx_t_stride + 4 * (i & 15)) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// fixed-order sum over the blocks' partial vectors -> the caller's gradient tensors (written, not accumulated)
__global__ void heads_bwd_reduce_kernel(const float* __restrict__ partial, int n_blocks, TrajsdeHeadsBwdArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * HG_N) return;
  const int h = i / HG_N, r = i % HG_N;
  if (h >= a.n_heads) return;
  float s = 0.f;
  for (int b = 0; b < n_blocks; ++b) s += partial[(size_t)b * 2 * HG_PAD + h * HG_PAD + r];
  const TrajsdeHeadGrad& g = a.grad_head[h];
  if (r < HG_B1) g.w1[r] = s;
  else if (r < HG_G) g.b1[r - HG_B1] = s;
  else if (r < HG_BETA) g.ln_g[r - HG_G] = s;
  else if (r < HG_W2) g.ln_b[r - HG_BETA] = s;
  else if (r < HG_B2) g.w2[r - HG_W2] = s;
  else g.b2[r - HG_B2] = s;
}

constexpr size_t HB_SMEM = (2 * 64 * HB_LD + 4 * HB_TILE * HB_LD + 2 * 320 + 2 * HB_TILE) * sizeof(float) + 2 * (HB_TILE + HB_THREADS) * sizeof(int64_t) +
                           16 * sizeof(int) + HB_THREADS * sizeof(int64_t);

int heads_bwd_grid() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return 2 * sms;
}

}  // namespace

int64_t heads_bwd_workspace_bytes() { return (int64_t)heads_bwd_grid() * 2 * HG_PAD * 4 + 256; }

int launch_heads_bwd(const TrajsdeHeadsBwdArgs& a, cudaStream_t s) {
  HeadsBwdParams p;
  p.a = a;
  p.partial = static_cast<float*>(a.workspace);
  p.n_points = a.rows * (int64_t)a.n_t;
  const bool cat4 = (a.flags & TRAJSDE_HEADS_FLAG_CAT4) != 0;
  p.gstride = cat4 ? 4 : 2;
  p.elu_head1 = cat4 ? 1 : 0;
  if (cat4) p.a.grad_out[1] = a.grad_out[0] ? a.grad_out[0] + 2 : nullptr;   // channels 2..3 of dL/d out['loc']
  int grid = heads_bwd_grid();
  if (grid <= 0) return set_error(TRAJSDE_ERR_CUDA, "device attributes unavailable");
  const int64_t chunks = (p.n_points + HB_THREADS - 1) / HB_THREADS;
  if (chunks < grid) grid = (int)(chunks > 0 ? chunks : 1);
  if (a.grad_amax) TS_CUDA_CHECK(cudaMemsetAsync(a.grad_amax, 0, 4, s));
  if (a.row_flags && a.rows > 0) {
    TS_CUDA_CHECK(cudaMemsetAsync(a.row_flags, 0, (size_t)a.rows, s));
    if (p.n_points > 0) {
      heads_row_flags_kernel<<<grid, HB_THREADS, 0, s>>>(p.a, p.n_points, p.gstride);
      TS_CUDA_CHECK(cudaGetLastError());
      heads_zero_rows_kernel<<<grid, HB_THREADS, 0, s>>>(a);
      TS_CUDA_CHECK(cudaGetLastError());
    }
  }
  TS_CUDA_CHECK(cudaFuncSetAttribute(heads_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HB_SMEM));
  heads_bwd_kernel<<<grid, HB_THREADS, HB_SMEM, s>>>(p);
  TS_CUDA_CHECK(cudaGetLastError());
  heads_bwd_reduce_kernel<<<(2 * HG_N + 255) / 256, 256, 0, s>>>(p.partial, grid, a);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

}  // namespace trajsde
