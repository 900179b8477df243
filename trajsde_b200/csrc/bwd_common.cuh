// Shared by the exact and tensor-core backward kernels: flat gradient-vector layout and the deterministic partial reduce.
#pragma once
#include "common.cuh"

namespace trajsde {
namespace bwd {

// flat gradient vector layout (21,121 floats per (f, g) pair; g_alt appended for dual)
constexpr int G_FW1 = 0, G_FB1 = 4224, G_FW2 = 4288, G_FB2 = 8384, G_FW3 = 8448, G_FB3 = 12544;
constexpr int G_GW1 = 12608, G_GB1 = 16832, G_GW2 = 16896, G_GB2 = 20992, G_GW3 = 21056, G_GB3 = 21120;
constexpr int G_TOTAL = 21121, G_PAD = 21124;

constexpr int MAX_PARTIALS = 160;   // per-CTA partial vectors per pass
constexpr int64_t BWD_TC_IMG_BYTES = 76800;
constexpr int64_t GRU_TC_IMG_BYTES = 75776;   // packed GRU weight image of gru_bwd_tc.cu   // packed weight image of euler_bwd_tc.cu (rounded up to 1 KB)

// GRU_Unit gradient vector layout: per gate (update, reset, new_state): w1[64,128] b1[64] w2[64,64] b2[64]
constexpr int GRU_GATE = 8192 + 64 + 4096 + 64;   // 12416
constexpr int GRU_U1 = 0, GRU_UB1 = 8192, GRU_U2 = 8256, GRU_UB2 = 12352;
constexpr int GRU_R1 = GRU_GATE, GRU_RB1 = GRU_GATE + 8192, GRU_R2 = GRU_GATE + 8256, GRU_RB2 = GRU_GATE + 12352;
constexpr int GRU_N1 = 2 * GRU_GATE, GRU_NB1 = 2 * GRU_GATE + 8192, GRU_N2 = 2 * GRU_GATE + 8256, GRU_NB2 = 2 * GRU_GATE + 12352;
constexpr int GRU_G_TOTAL = 3 * GRU_GATE, GRU_G_PAD = 3 * GRU_GATE;   // 37248

// grads[i] = sum over CTAs (fixed order) of partial sets; set 0 -> (f, g); set 1 (dual) -> (f, g_alt)
static __global__ void euler_bwd_reduce_kernel(const float* __restrict__ part0, const float* __restrict__ part1, int n0, int n1,
                                        TrajsdeMlpGrad gf, TrajsdeMlpGrad gg, TrajsdeMlpGrad ga, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G_TOTAL) return;
  float s0 = 0.f, s1 = 0.f;
  for (int c = 0; c < n0; ++c) s0 += part0[(size_t)c * G_PAD + i];
  for (int c = 0; c < n1; ++c) s1 += part1[(size_t)c * G_PAD + i];
  if (i < G_GW1) {
    const float v = s0 + s1;
    float* d = i < G_FB1 ? gf.w1 + (i - G_FW1) : i < G_FW2 ? gf.b1 + (i - G_FB1) : i < G_FB2 ? gf.w2 + (i - G_FW2)
             : i < G_FW3 ? gf.b2 + (i - G_FB2) : i < G_FB3 ? gf.w3 + (i - G_FW3) : gf.b3 + (i - G_FB3);
    *d = accumulate ? *d + v : v;
  } else {
    for (int set = 0; set < 2; ++set) {
      if (set == 1 && !part1) break;
      const TrajsdeMlpGrad& t = set == 0 ? gg : ga;
      const float v = set == 0 ? s0 : s1;
      float* d = i < G_GB1 ? t.w1 + (i - G_GW1) : i < G_GW2 ? t.b1 + (i - G_GB1) : i < G_GB2 ? t.w2 + (i - G_GW2)
               : i < G_GW3 ? t.b2 + (i - G_GB2) : i < G_GB3 ? t.w3 + (i - G_GW3) : t.b3;
      *d = accumulate ? *d + v : v;
    }
  }
}


}  // namespace bwd
}  // namespace trajsde
