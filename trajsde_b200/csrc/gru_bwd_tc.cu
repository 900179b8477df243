// Tensor-core backward of one GRU_Unit observation jump of the encoder recurrence (models/utils/ode_utils.py:136-152, called at
// models/encoders/enc_hivt_nusargo_sde_sep2.py:165-169); same math as gru_bwd.cu, mapped like euler_bwd_tc.cu:
//
//   u = sigmoid(U2 tanh(U1 [y1,x] + ub1) + ub2)    r = sigmoid(R2 tanh(R1 [y1,x] + rb1) + rb2)
//   n = N2 tanh(N1 [x, r*y1] + nb1) + nb2          h' = mask ? (1-u) n + u y1 : y1
//
//   * persistent CTAs, one 128-row tile at a time; 8 epilogue warps (thread = row x 32-channel half) + one MMA-issuer warp;
//   * eight dependent tcgen05 phases per tile (fp16 operands, fp32 accumulate): F1 [zu|zr], F2 [u'|r'], F3 zn, F4 n (recompute),
//     B1 d_tn = d_n N2, B2 [d_x|d(r y1)] = d_zn N1, B3 [d_tu|d_tr] = [d_u' U2 | d_r' R2], B4 [d_y1|d_x] += d_zu U1 + d_zr R1;
//     the backward phases read the FORWARD weight tiles as MN-major B operands (no transposed copies in shared memory);
//   * weight gradients = MN-major products over the operand tiles (contraction over the 128 rows), accumulated in TMEM across the
//     CTA's tiles and added once per launch into the CTA's private partial vector (plain read-modify-write, fixed order):
//       [dz_n|d_n]^T [x|r y1] -> dN1 ,  [dz_n|d_n]^T tn -> dN2 ,  [d_u'|d_r']^T [tu|tr] -> dU2, dR2 ,  [dz_u|dz_r]^T [y1|x] -> dU1, dR1
//     bias gradients are column sums taken in registers (butterfly transpose-reduce over the warp's 32 rows);
//   * deltas are carried scaled by the sweep's power-of-two loss scale (same amax word as the SDE backward); outputs are unscaled.
#include "gru_bwd_tc_body.cuh"

namespace trajsde {

using namespace tc;
using namespace bwd;
using namespace bwdtc;
using namespace grutc;

namespace {


__global__ void gru_tc_pack_kernel(TrajsdeGru w, uint8_t* __restrict__ img) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  auto put = [&](uint32_t off, int n, int k, float v) { *reinterpret_cast<__half*>(img + off + sw128_off_h(n, k)) = __float2half_rn(v); };
  for (int idx = tid; idx < 64 * 64; idx += nth) {
    const int n = idx >> 6, k = idx & 63;
    put(IMG_UR1H, n, k, w.u1[n * 128 + k]);
    put(IMG_UR1H, n + 64, k, w.r1[n * 128 + k]);
    put(IMG_UR1X, n, k, w.u1[n * 128 + 64 + k]);
    put(IMG_UR1X, n + 64, k, w.r1[n * 128 + 64 + k]);
    put(IMG_U2, n, k, w.u2[n * 64 + k]);
    put(IMG_R2, n, k, w.r2[n * 64 + k]);
    put(IMG_N1X, n, k, w.n1[n * 128 + k]);
    put(IMG_N1RH, n, k, w.n1[n * 128 + 64 + k]);
    put(IMG_N2, n, k, w.n2[n * 64 + k]);
  }
  float* vec = reinterpret_cast<float*>(img + IMG_VEC);
  for (int i = tid; i < 384; i += nth) {
    const int c = i & 63;
    const float* src = i < 64 ? w.ub1 : i < 128 ? w.rb1 : i < 192 ? w.nb1 : i < 256 ? w.ub2 : i < 320 ? w.rb2 : w.nb2;
    vec[i] = src[c];
  }
}




__global__ void __launch_bounds__(NUM_THREADS, 1) gru_bwd_tc_kernel(const GruTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  gru_bwd_tc_body<false>(p, SweepCtl{}, (int)blockIdx.x, (int)gridDim.x, smem_raw);
}

}  // namespace

int gru_bwd_tc_pack(const TrajsdeGru& w, uint8_t* img, cudaStream_t s) {
  gru_tc_pack_kernel<<<16, 256, 0, s>>>(w, img);
  TS_CUDA_CHECK(cudaGetLastError());
  return TRAJSDE_OK;
}

int launch_gru_bwd_tc(int64_t rows, const float* y1, const float* aa_out, int64_t slab, const uint8_t* obs_mask, int64_t obs_mask_row_stride,
                      const int32_t* slot, int iter, const float* carry, const float* grad_latent, float* grad_y1, float* grad_aa_out,
                      const uint8_t* img, const uint32_t* amax_bits, float* partial, float* h_out_fwd_only, cudaStream_t s, bool pdl) {
  GruTcParams p;
  p.fwd_only = h_out_fwd_only != nullptr;
  p.h_out = h_out_fwd_only;
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
  p.img = img;
  p.amax_bits = amax_bits;
  p.partial = partial;
  p.num_tiles = (int)((rows + TILE_M - 1) / TILE_M);
  const int grid = bwd_tc_grid(rows);
  if (grid <= 0) return TRAJSDE_OK;
  TS_CUDA_CHECK(cudaFuncSetAttribute(gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_ALLOC));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_ALLOC;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  TS_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gru_bwd_tc_kernel, p));
  return TRAJSDE_OK;
}

}  // namespace trajsde
