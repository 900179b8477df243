// Tensor-core backward of the fused Euler–Maruyama solve (TRAJSDE_MODE_TC_F16, single diffusion net): discretise-then-optimise,
// i.e. what torch.autograd computes through the reference solver (config `adjoint: false`, configs/nusargo/
// hivt_nuSArgo_sdesepenc_sdedec.yml:41; solver models/utils/sdeint.py:340-384,477-485,544; nets dec_hivt_nusargo_sde.py:119-127,
// 154-158,180-195).  Same math as euler_bwd_exact.cu, different machine mapping: ONE persistent sm_100a kernel fuses the
// reverse adjoint sweep (dgrad) and the weight gradients (wgrad):
//
//   * one CTA per SM, one 128-row tile at a time; 8 epilogue warps (thread = row x 32-channel half) + one MMA-issuer warp;
//   * per step, five dependent tcgen05 phases (fp16 operands, fp32 accumulate in TMEM), recomputing the activations from the
//     state Y[k] the forward call saved:
//        P1  [z1f|z1g] = y . [W1y;V1y]^T                  P2  z2f = h1f . W2^T ,  z2g = h1g . V2^T
//        D1  dh2f = df . W3 ,  dh1g = dz2g . V2           D2  dh1f = dz2f . W2           D3  dy = dz1f . W1y + dz1g . V1y
//     with  df = h A',  ds = (A'.dW) g (1-g),  dz2g = ds w3 (1-h2g^2),  dz* = dh* (1-h*^2),  A = A' + dy + sum_j w0_j gy_j;
//   * the weight gradients are MN-major tcgen05.mma products over the SAME shared-memory tiles (contraction over the 128 rows),
//     accumulated in TMEM across all steps and tiles of the CTA and written once, as one partial vector per CTA:
//        dW2|dV2 += [dz2f|dz2g]^T [h1f|h1g]      dW1y|dV1y += [dz1f|dz1g]^T y      dW3 += df^T h2f
//     bias / time-column gradients are the same products against a [1, sin t, cos t] tile; they trail the dgrad MMAs of their
//     phase, off the critical path;
//   * the adjoint is carried SCALED by a power of two chosen from max|grad| (absmax pre-pass) so that the fp16 delta operands
//     neither overflow nor underflow; everything leaves the kernel unscaled, in fp32;
//   * HBM traffic per row-step: state 256 B + dW 256 B (or Philox regenerated in-kernel) + grad_ys 256 B, read once.
// Partials are summed by the fixed-order reduce of bwd_common.cuh (bit-reproducible, no float atomics).
#include "euler_bwd_tc_body.cuh"

namespace trajsde {

using namespace tc;
using namespace bwd;
using namespace bwdtc;
using namespace sdetc;

namespace {


// ---- pre-passes -------------------------------------------------------------------------------------------------------------------
__global__ void bwd_tc_pack_kernel(TrajsdeEulerBwdArgs a, uint8_t* __restrict__ img) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  auto put = [&](uint32_t off, int n, int k, float v) { *reinterpret_cast<__half*>(img + off + sw128_off_h(n, k)) = __float2half_rn(v); };
  for (int idx = tid; idx < 64 * 64; idx += nth) {
    const int n = idx >> 6, k = idx & 63;
    put(IMG_B1, n, k, a.drift.w1[n * TS_IN1 + k]);
    put(IMG_B1, n + 64, k, a.diffusion.w1[n * TS_IN1 + k]);
    put(IMG_W2, n, k, a.drift.w2[n * 64 + k]);
    put(IMG_V2, n, k, a.diffusion.w2[n * 64 + k]);
    put(IMG_W3T, n, k, a.drift.w3[k * 64 + n]);
    put(IMG_W2T, n, k, a.drift.w2[k * 64 + n]);
    put(IMG_V2T, n, k, a.diffusion.w2[k * 64 + n]);
    put(IMG_W1YT, n, k, a.drift.w1[k * TS_IN1 + n]);
    put(IMG_V1YT, n, k, a.diffusion.w1[k * TS_IN1 + n]);
  }
  float* vec = reinterpret_cast<float*>(img + IMG_VEC);
  for (int i = tid; i < 640; i += nth) {
    const int c = i & 63;
    float v = 0.f;
    switch (i >> 6) {
      case 0: v = a.drift.b1[c]; break;
      case 1: v = a.drift.w1[c * TS_IN1 + 64]; break;
      case 2: v = a.drift.w1[c * TS_IN1 + 65]; break;
      case 3: v = a.drift.b2[c]; break;
      case 4: v = a.diffusion.b1[c]; break;
      case 5: v = a.diffusion.w1[c * TS_IN1 + 64]; break;
      case 6: v = a.diffusion.w1[c * TS_IN1 + 65]; break;
      case 7: v = a.diffusion.b2[c]; break;
      case 8: v = a.diffusion.w3[c]; break;
      default: v = c == 0 ? a.diffusion.b3[0] : 0.f; break;
    }
    vec[i] = v;
  }
}

// max |x| over `slabs` slabs of [rows][64] floats (slab / row strides in elements) or, with row_stride == 0, over a flat array of
// `rows` floats -> atomicMax on *amax_bits (non-negative floats order like their bit patterns)
// `block_step` > 1: only every block_step-th block of 32 rows is scanned (all slabs, all channels of those rows) — the result feeds a
// power-of-two loss scale that has 2^17 of head-room above it (status bit at 2^14, fp16 overflow at 2^16 of a value mapped to 2^-3), so an
// estimate within a few binades of the true maximum is as good as the maximum, and the scan of a [61, 204800, 64] gradient drops from
// 0.47 ms to 0.06 ms.  `only_if_zero`: the full scan that follows a sampled one; it returns at once unless the sample saw nothing but
// zeros (e.g. every sampled row is a padded agent), so a tiny gradient is never scaled by 1 and flushed to zero in fp16.
__global__ void bwd_tc_absmax_kernel(const float* __restrict__ x, int slabs, int64_t rows, int64_t slab_stride, int64_t row_stride,
                                     uint32_t* __restrict__ amax_bits, int block_step, int only_if_zero) {
  if (only_if_zero && *reinterpret_cast<volatile uint32_t*>(amax_bits) != 0u) return;
  float m = 0.f;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  if (row_stride == 0) {
    for (int64_t i = tid; i < rows; i += nth) m = fmaxf(m, fabsf(x[i]));
  } else {
    const int64_t n_blocks = (rows + 31) / 32, n_sampled = (n_blocks + block_step - 1) / block_step;
    const int64_t per_slab = n_sampled * 512;   // float4 units: 32 rows x 16 per sampled block
    for (int t = 0; t < slabs; ++t) {
      const float* slab = x + (int64_t)t * slab_stride;
      for (int64_t i = tid; i < per_slab; i += nth) {
        const int64_t r = (i >> 9) * block_step * 32 + ((i >> 4) & 31);
        if (r >= rows) continue;
        const float4 v = ld_nc_f4(slab + r * row_stride + 4 * (i & 15));
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
      }
    }
  }
  if (!(m <= 3.0e38f)) m = 3.0e38f;   // inf / nan in the incoming gradient: clamp (the result is garbage either way)
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}

// Zero-row skipping, pass 1: a scan of grad_ys (every slab) that yields BOTH max |grad| (for the loss scale) and a per-row flag "some
// incoming gradient of this row is non-zero".  16 threads per row (one float4 each), 16 rows per 256-thread block and iteration: whole
// 256-byte rows per slab, coalesced in either storage layout.  `grad_g` (sdeint_dual's second output) marks its rows too.
//   sample == 1: only every 8th block of 32 rows is scanned; counts[0] += active sampled rows, counts[1] += sampled rows (0.06 ms at
//                204,800 rows x 61 slabs);
//   sample == 0: the full scan — unless the sample found more than half of its rows active: then the cotangent is (close to) dense,
//                skipping would save little, and every row is flagged without reading the gradient again (the sampled maximum is good
//                enough for the power-of-two loss scale, see bwd_tc_absmax_kernel).
__global__ void bwd_row_activity_kernel(const float* __restrict__ x, int slabs, int64_t rows, int64_t slab_stride, int64_t row_stride,
                                        const float* __restrict__ grad_g, uint32_t* __restrict__ amax_bits, uint8_t* __restrict__ flags,
                                        uint32_t* __restrict__ counts, int sample, const uint8_t* __restrict__ given) {
  // `given`: the caller's row flags (TrajsdeEulerBwdArgs.row_flags): rows with given[r] == 0 are not even read; the others are scanned
  // for max |grad| only
  const int sub = threadIdx.x & 15;
  const bool dense = !given && !sample && 2u * reinterpret_cast<volatile uint32_t*>(counts)[0] > reinterpret_cast<volatile uint32_t*>(counts)[1];
  const int64_t n_items = sample ? ((rows + 255) / 256) * 32 : rows;       // sampled: 32 rows out of every 256
  float bm = 0.f;
  uint32_t n_act = 0, n_seen = 0;
  for (int64_t it = (int64_t)blockIdx.x * 16 + (threadIdx.x >> 4); it < n_items; it += (int64_t)gridDim.x * 16) {   // uniform per 16-lane group
    const int64_t r = sample ? (it >> 5) * 256 + (it & 31) : it;
    if (r >= rows) continue;
    if (dense) {
      if (sub == 0) flags[r] = 1;
      continue;
    }
    if (given && !given[r]) {
      if (sub == 0) flags[r] = 0;
      continue;
    }
    float m = 0.f;
    const float* px = x + r * row_stride + 4 * sub;
    int t = 0;
    for (; t + 8 <= slabs; t += 8) {                         // eight independent 16-byte loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ld_nc_f4(px + (int64_t)(t + j) * slab_stride);
#pragma unroll
      for (int j = 0; j < 8; ++j) m = fmaxf(m, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
    }
    for (; t < slabs; ++t) {
      const float4 v = ld_nc_f4(px + (int64_t)t * slab_stride);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    if (grad_g && sub == 0) m = fmaxf(m, fabsf(grad_g[r]));
    if (!(m <= 3.0e38f)) m = 3.0e38f;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off, 16));
    if (sub == 0) {
      if (!sample) flags[r] = (given || m > 0.f) ? 1 : 0;
      n_act += m > 0.f ? 1u : 0u;
      n_seen += 1u;
    }
    bm = fmaxf(bm, m);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, off));
  if ((threadIdx.x & 31) == 0 && bm > 0.f) atomicMax(amax_bits, __float_as_uint(bm));
  if (sample) {
    n_act += __shfl_xor_sync(0xffffffffu, n_act, 16);
    n_seen += __shfl_xor_sync(0xffffffffu, n_seen, 16);
    if ((threadIdx.x & 31) == 0 && n_seen) {
      atomicAdd(counts, n_act);
      atomicAdd(counts + 1, n_seen);
    }
  }
}

// pass 2: order-preserving compaction of the flagged rows (one 1024-thread block: count per contiguous chunk, block scan, write)
__global__ void __launch_bounds__(1024, 1) bwd_compact_rows_kernel(const uint8_t* __restrict__ flags, int64_t rows, int32_t* __restrict__ row_map,
                                                                    int32_t* __restrict__ n_active) {
  __shared__ int32_t part[1024];
  // chunks are multiples of 16 rows so that the flags go through 16-byte loads when the array is 16-byte aligned (byte loads took 82 us
  // at 204,800 rows — a latency chain of 200 dependent-address loads per thread)
  const int64_t chunk = (((rows + 1023) / 1024) + 15) & ~(int64_t)15, lo = (int64_t)threadIdx.x * chunk;
  const int64_t hi = lo >= rows ? lo : (lo + chunk < rows ? lo + chunk : rows);
  const bool vec = (reinterpret_cast<uintptr_t>(flags) & 15u) == 0;
  int32_t c = 0;
  {
    int64_t r = lo;
    if (vec)
      for (; r + 16 <= hi; r += 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(flags + r);
        c += (__popc(__vcmpne4(v.x, 0u)) + __popc(__vcmpne4(v.y, 0u)) + __popc(__vcmpne4(v.z, 0u)) + __popc(__vcmpne4(v.w, 0u))) >> 3;
      }
    for (; r < hi; ++r) c += flags[r] != 0;
  }
  part[threadIdx.x] = c;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {          // Hillis–Steele inclusive scan
    const int32_t v = (int)threadIdx.x >= off ? part[threadIdx.x - off] : 0;
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  int32_t pos = part[threadIdx.x] - c;
  {
    int64_t r = lo;
    if (vec)
      for (; r + 16 <= hi; r += 16) {
        const uint4 v = *reinterpret_cast<const uint4*>(flags + r);
        if ((v.x | v.y | v.z | v.w) == 0u) continue;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if ((w[i >> 2] >> (8 * (i & 3))) & 0xffu) row_map[pos++] = (int32_t)(r + i);
      }
    for (; r < hi; ++r)
      if (flags[r]) row_map[pos++] = (int32_t)r;
  }
  if (threadIdx.x == 1023) *n_active = part[1023];
}


template <bool HAS_DW>
__global__ void __launch_bounds__(NUM_THREADS, 1) euler_bwd_tc_kernel(const BwdTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  euler_bwd_tc_body<HAS_DW, false>(p, SweepCtl{}, (int)blockIdx.x, (int)gridDim.x, (int)blockIdx.y, smem_raw);
}

}  // namespace

// order-preserving compaction of flagged rows (shared with the aggr_embed backward, stage_ops.cu)
int launch_compact_rows(const uint8_t* flags, int64_t rows, int32_t* row_map, int32_t* n_active, cudaStream_t s) {
  bwd_compact_rows_kernel<<<1, 1024, 0, s>>>(flags, rows, row_map, n_active);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

// ---- internal launch API (also used by enc_bwd.cu) ---------------------------------------------------------------------------------
int bwd_tc_pack(const TrajsdeEulerBwdArgs& a, uint8_t* img, cudaStream_t s) {
  bwd_tc_pack_kernel<<<16, 256, 0, s>>>(a, img);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

int bwd_tc_absmax(const float* x, int slabs, int64_t rows, int64_t slab_stride, int64_t row_stride, uint32_t* amax_bits, cudaStream_t s) {
  if (!x || rows <= 0) return TRAJSDE_OK;
  int dev = 0, sms = 0;
  TS_CUDA_CHECK(cudaGetDevice(&dev));
  TS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // large row-major gradients: sampled scan (every 8th block of 32 rows), then the full scan that only runs if the sample was all zero
  const bool sampled = row_stride != 0 && rows * (int64_t)slabs >= ((int64_t)1 << 16);
  for (int pass = 0; pass < (sampled ? 2 : 1); ++pass) {
    const int step = sampled && pass == 0 ? 8 : 1;
    const int64_t work = row_stride == 0 ? rows : ((rows + 31) / 32 + step - 1) / step * 512;
    int64_t blocks = (work + 511) / 512;
    if (blocks > 2 * sms) blocks = 2 * sms;
    bwd_tc_absmax_kernel<<<(int)blocks, 512, 0, s>>>(x, slabs, rows, slab_stride, row_stride, amax_bits, step, sampled && pass == 1 ? 1 : 0);
    TS_CUDA_CHECK(cudaGetLastError());
  }
  return TRAJSDE_OK;
}

// CTAs per pass.  Dual diffusion launches two passes (gridDim.y = 2) of one-CTA-per-SM kernels: half the SMs per pass keeps both passes
// resident at once — one wave with one prologue and one weight-gradient flush per CTA instead of two waves.
int bwd_tc_grid(int64_t rows, bool dual) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (sms > MAX_PARTIALS) sms = MAX_PARTIALS;
  if (dual && sms > 1) sms /= 2;
  const int64_t tiles = (rows + TILE_M - 1) / TILE_M;
  return (int)(tiles < sms ? tiles : sms);
}

// one fused dgrad+wgrad launch over all steps of a.sched; partial[grid][G_PAD] written (or accumulated into).
// img1 != NULL: dual diffusion — pass 0 (img0, part0) takes the rows with alt_mask != 0, pass 1 (img1 packed with a.diffusion_alt,
// part1) the rows with alt_mask == 0; both passes run in the same launch (gridDim.y = 2).
int bwd_tc_main(const TrajsdeEulerBwdArgs& a, const uint8_t* img0, const uint8_t* img1, const uint32_t* amax_bits, float* part0, float* part1,
                int accumulate, cudaStream_t s, bool pdl, const int32_t* row_map, const int32_t* n_active) {
  if ((reinterpret_cast<uintptr_t>(img0) & 15u) != 0 || (reinterpret_cast<uintptr_t>(img1) & 15u) != 0)
    return set_error(TRAJSDE_ERR_UNSUPPORTED, "workspace must be 256-byte aligned");
  if (a.rows >= (int64_t)1 << 31) return set_error(TRAJSDE_ERR_UNSUPPORTED, "rows >= 2^31 unsupported in TC mode");
  const bool dual = img1 != nullptr;
  BwdTcParams p;
  p.a = a;
  p.img[0] = img0;
  p.img[1] = img1;
  p.partial[0] = part0;
  p.partial[1] = part1;
  p.filter[0] = dual ? 1 : 0;
  p.filter[1] = 2;
  p.amax_bits = amax_bits;
  p.row_map = row_map;
  p.n_active = n_active;
  p.num_tiles = (int)((a.rows + TILE_M - 1) / TILE_M);
  p.accumulate = accumulate;
  const int grid = bwd_tc_grid(a.rows, dual);
  if (grid <= 0) return TRAJSDE_OK;
  auto launch = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, dual ? 2 : 1);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_ALLOC;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, p);
  };
  TS_CUDA_CHECK(a.noise.dw ? launch(euler_bwd_tc_kernel<true>) : launch(euler_bwd_tc_kernel<false>));
  return TRAJSDE_OK;
}

int launch_euler_bwd_reduce(const float* part0, const float* part1, int n0, int n1, const TrajsdeMlpGrad& gf, const TrajsdeMlpGrad& gg,
                            const TrajsdeMlpGrad& ga, cudaStream_t s) {
  euler_bwd_reduce_kernel<<<(G_TOTAL + 255) / 256, 256, 0, s>>>(part0, part1, n0, n1, gf, gg, ga, 0);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

static int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }

int64_t euler_bwd_tc_workspace_bytes(int64_t rows, int32_t n_steps) {
  (void)n_steps;
  return 2 * (int64_t)BWD_TC_IMG_BYTES + 256 + align256(2 * (int64_t)MAX_PARTIALS * G_PAD * 4) + 256 + align256(rows) + align256(4 * rows);
}

// workspace: img0 | img1 | amax (256 B) | partial0 | partial1 | n_active (256 B) | row flags [rows] | row_map [rows] (zero-row skipping)
int launch_euler_bwd_tc(const TrajsdeEulerBwdArgs& a, cudaStream_t s) {
  if ((reinterpret_cast<uintptr_t>(a.workspace) & 255u) != 0) return set_error(TRAJSDE_ERR_UNSUPPORTED, "workspace must be 256-byte aligned");
  uint8_t* ws = static_cast<uint8_t*>(a.workspace);
  uint8_t* img0 = ws;
  uint8_t* img1 = ws + BWD_TC_IMG_BYTES;
  uint32_t* amax = reinterpret_cast<uint32_t*>(ws + 2 * BWD_TC_IMG_BYTES);
  float* part0 = reinterpret_cast<float*>(ws + 2 * BWD_TC_IMG_BYTES + 256);
  float* part1 = part0 + (size_t)MAX_PARTIALS * G_PAD;
  const bool dual = a.alt_mask != nullptr;
  const int grid = bwd_tc_grid(a.rows, dual);
  // zero-row skipping: a winner-takes-all loss (losses/L2.py:17-20: only the best of the 10 modes of an actor receives a gradient) leaves
  // ~90 % of the decoder rows with an all-zero incoming gradient; their adjoint is zero at every step, so the sweep visits the others only
  const bool skip_zero = (a.flags & TRAJSDE_BWD_FLAG_SKIP_ZERO_ROWS) != 0 && !dual && a.grad_ys != nullptr;
  uint8_t* tail = reinterpret_cast<uint8_t*>(part0) + align256(2 * (int64_t)MAX_PARTIALS * G_PAD * 4);
  int32_t* n_active = reinterpret_cast<int32_t*>(tail);
  uint8_t* row_flags = tail + 256;
  int32_t* row_map = reinterpret_cast<int32_t*>(row_flags + align256(a.rows));
  int rc;
  if (grid > 0) {
    TS_CUDA_CHECK(cudaMemsetAsync(amax, 0, 4, s));
    if (skip_zero) {
      int dev = 0, sms = 0;
      TS_CUDA_CHECK(cudaGetDevice(&dev));
      TS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      uint32_t* counts = reinterpret_cast<uint32_t*>(n_active) + 2;        // {active, seen} sampled rows
      TS_CUDA_CHECK(cudaMemsetAsync(counts, 0, 8, s));
      const bool handed = a.row_flags && a.grad_amax && !a.grad_g_last;                          // flags AND the gradient's magnitude come from the caller:
      if (handed) TS_CUDA_CHECK(cudaMemcpyAsync(amax, a.grad_amax, 4, cudaMemcpyDeviceToDevice, s));   // nothing to scan
      for (int sample = a.row_flags ? 0 : 1; sample >= 0 && !handed; --sample) {   // caller's flags: no sampling pass, flagged rows only
        const int64_t want = ((sample ? ((a.rows + 255) / 256) * 32 : a.rows) + 15) / 16;
        bwd_row_activity_kernel<<<(int)(want < 8 * sms ? want : 8 * sms), 256, 0, s>>>(a.grad_ys, a.sched.n_outputs + 1, a.rows, a.grad_ys_t_stride,
                                                                                     a.grad_ys_row_stride, a.grad_g_last, amax, row_flags, counts, sample,
                                                                                     a.row_flags);
        TS_CUDA_CHECK(cudaGetLastError());
      }
      bwd_compact_rows_kernel<<<1, 1024, 0, s>>>(handed ? a.row_flags : row_flags, a.rows, row_map, n_active);
      TS_CUDA_CHECK(cudaGetLastError());
      TS_CUDA_CHECK(cudaMemsetAsync(a.grad_y0, 0, sizeof(float) * 64 * (size_t)a.rows, s));   // skipped rows: dL/dy0 = 0
    } else {
      if (a.grad_ys && (rc = bwd_tc_absmax(a.grad_ys, a.sched.n_outputs + 1, a.rows, a.grad_ys_t_stride, a.grad_ys_row_stride, amax, s)) != 0) return rc;
      if (a.grad_g_last && (rc = bwd_tc_absmax(a.grad_g_last, 1, a.rows, 0, 0, amax, s)) != 0) return rc;
    }
    if ((rc = bwd_tc_pack(a, img0, s)) != 0) return rc;
    if (dual) {   // second pass: the rows of the other diffusion net (rows are independent; the drift gradients of both passes add up)
      TrajsdeEulerBwdArgs b = a;
      b.diffusion = a.diffusion_alt;
      if ((rc = bwd_tc_pack(b, img1, s)) != 0) return rc;
    }
    if ((rc = bwd_tc_main(a, img0, dual ? img1 : nullptr, amax, part0, part1, 0, s, false, skip_zero ? row_map : nullptr,
                          skip_zero ? n_active : nullptr)) != 0)
      return rc;
  }
  euler_bwd_reduce_kernel<<<(G_TOTAL + 255) / 256, 256, 0, s>>>(part0, dual ? part1 : nullptr, grid, dual ? grid : 0, a.grad_drift,
                                                              a.grad_diffusion, a.grad_diffusion_alt, 0);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

}  // namespace trajsde

#ifdef TRAJSDE_BWD_TIMELINE
// debug build only (bench_micro/bwd_timeline.py): accumulated clocks per phase of thread 0 / CTA 0, then reset
extern "C" int trajsde_debug_bwd_timeline(long long* out16) {
  long long zero[16] = {0};
  if (cudaMemcpyFromSymbol(out16, trajsde::sdetc::g_bwd_tl, sizeof(zero)) != cudaSuccess) return -1;
  return cudaMemcpyToSymbol(trajsde::sdetc::g_bwd_tl, zero, sizeof(zero)) == cudaSuccess ? 0 : -1;
}
#endif
