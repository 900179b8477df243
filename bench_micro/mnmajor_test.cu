// Validates MN-major UMMA operands on tiles stored as [K rows][64 x f16 = 128 B] with 128B swizzle (the layout every operand
// tile of the solver kernels already has): D[M=128][N=128] = sum_r A[r][M] * B[r][N], A = two adjacent tiles, B = two adjacent tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I trajsde_b200/csrc -o bench_micro/mnmajor_test bench_micro/mnmajor_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "../trajsde_b200/csrc/tc_common.cuh"
using namespace trajsde::tc;

__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;   // LBO: stride between 64-element MN groups
  d |= (uint64_t)(1024u >> 4) << 32;                  // SBO: stride between 8-row K groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int n_dim, int a_bf16) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  // tiles: A0 (cols 0..63 of a), A1 (cols 64..127), B0, B1: each [128 rows][64] f16 swizzled; a,b are [128][128] fp32 row-major
  for (int idx = tid; idx < 128 * 128; idx += blockDim.x) {
    const int r = idx >> 7, c = idx & 127, t = c >> 6, cc = c & 63;
    if (a_bf16) *reinterpret_cast<__nv_bfloat16*>(sm + t * 16384 + sw128_off_h(r, cc)) = __float2bfloat16_rn(a[idx]);
    else *reinterpret_cast<__half*>(sm + t * 16384 + sw128_off_h(r, cc)) = __float2half_rn(a[idx]);
    *reinterpret_cast<__half*>(sm + 32768 + t * 16384 + sw128_off_h(r, cc)) = __float2half_rn(b[idx]);
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_ptr), 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (tid == 0) {
    // idesc: D f32, A/B f16, a_major = b_major = MN (bits 15,16), N, M = 128
    const uint32_t idesc = (1u << 4) | ((a_bf16 ? 1u : 0u) << 7) | (1u << 15) | (1u << 16) | (((uint32_t)n_dim >> 3) << 17) | ((128u >> 4) << 24);
    for (int kk = 0; kk < 8; ++kk) {   // K = 128 rows, 16 per instruction -> +2048 B per step
      const uint64_t da = desc_mn_sw128(base + kk * 2048, 16384);
      const uint64_t db = desc_mn_sw128(base + 32768 + kk * 2048, 16384);
      tc_mma_f16(tm, da, db, idesc, kk > 0);
    }
    tc_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c0 = 0; c0 < n_dim; c0 += 32) {
    tmem_ld_32x32b_x32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc_wait_ld();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + (tid & 31)) * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}

int main() {
  std::vector<float> a(128 * 128), b(128 * 128), ref(128 * 128), got(128 * 128);
  srand(1);
  for (auto& x : a) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : b) x = (rand() % 2001 - 1000) / 1000.f;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 128; ++n) {
      double s = 0;
      for (int r = 0; r < 128; ++r) s += (double)__half2float(__float2half_rn(a[r * 128 + m])) * __half2float(__float2half_rn(b[r * 128 + n]));
      ref[m * 128 + n] = (float)s;
    }
  float *da, *db, *dout;
  cudaMalloc(&da, 65536); cudaMalloc(&db, 65536); cudaMalloc(&dout, 65536);
  cudaMemcpy(da, a.data(), 65536, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), 65536, cudaMemcpyHostToDevice);
  for (int a_bf16 = 0; a_bf16 < 2; ++a_bf16)
  for (int n_dim : {128, 64, 16}) {
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        double s = 0;
        for (int r = 0; r < 128; ++r) {
          const float av = a_bf16 ? __bfloat162float(__float2bfloat16_rn(a[r * 128 + m])) : __half2float(__float2half_rn(a[r * 128 + m]));
          s += (double)av * __half2float(__float2half_rn(b[r * 128 + n]));
        }
        ref[m * 128 + n] = (float)s;
      }
    cudaMemset(dout, 0, 65536);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    k<<<1, 128, 70000>>>(da, db, dout, n_dim, a_bf16);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(got.data(), dout, 65536, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < n_dim; ++n) maxerr = fmax(maxerr, fabs(got[m * 128 + n] - ref[m * 128 + n]));
    printf("A=%s B=f16 MN-major M=128 N=%d K=128: max abs err %.3e  (%s)  sample got %.4f ref %.4f\n", a_bf16 ? "bf16" : "f16", n_dim, maxerr, cudaGetErrorString(e), got[5 * 128 + 7], ref[5 * 128 + 7]);
  }
  return 0;
}
