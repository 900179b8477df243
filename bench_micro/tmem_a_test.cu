// A operand from tensor memory (tcgen05.mma ... [d_tmem], [a_tmem], b_desc): epilogue threads write the fp16 activations of their row
// with tcgen05.st (two halves per 32-bit column) instead of st.shared + fence.proxy.async.  Checks numerics against the CPU and
// measures clocks per dependent phase for both operand paths.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I trajsde_b200/csrc -o bench_micro/tmem_a_test bench_micro/tmem_a_test.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../trajsde_b200/csrc/tc_common.cuh"
using namespace trajsde::tc;

__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// mode 0: A through shared memory; mode 1: A through tensor memory.  D[128 x 64] = A[128 x 64] . B[64 x 64]^T, repeated `iters` times
// as a dependent chain (the epilogue re-writes A from the previous D, scaled, so every phase depends on the one before).
__global__ void __launch_bounds__(288, 1) k(int mode, int iters, const float* __restrict__ a0, const float* __restrict__ b, float* __restrict__ out,
                                            long long* out_clk) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_opnd = smem_u32(&bars[0]), bar_acc = smem_u32(&bars[1]);
  for (int i = tid; i < 64 * 64; i += blockDim.x) {
    const int n = i >> 6, kk = i & 63;
    *reinterpret_cast<__half*>(sm + 16384 + sw128_off_h(n, kk)) = __float2half_rn(b[i]);
  }
  if (tid == 0) {
    mbar_init(bar_opnd, 256);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_ptr), 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;   // D: columns [0,64) ; A (packed fp16): columns [64,96)
  if (warp < 8) {
    const int quad = warp & 3, hh = warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t lane_base = tm + ((uint32_t)(quad * 32) << 16);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = a0[row * 64 + hh * 32 + j];
    uint32_t par = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      uint32_t p[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) p[e] = pack_f16x2(v[2 * e], v[2 * e + 1]);
      if (mode == 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(sm + row * 128 + ((((uint32_t)(hh * 4 + q)) ^ (row & 7u)) << 4)) = make_uint4(p[4 * q], p[4 * q + 1], p[4 * q + 2], p[4 * q + 3]);
        fence_proxy_async();
      } else {
        tmem_st_32x32b_x16(lane_base + 64 + hh * 16, p);   // this thread's 32 halves = 16 packed columns
        tc_wait_st();
      }
      tc_fence_before();
      mbar_arrive(bar_opnd);
      mbar_wait(bar_acc, par);
      par ^= 1;
      tc_fence_after();
      uint32_t u[32];
      tmem_ld_32x32b_x32(lane_base + hh * 32, u);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]) * 0.25f;   // next phase's A
      if (it == 0) {
#pragma unroll
        for (int j = 0; j < 32; ++j) out[row * 64 + hh * 32 + j] = __uint_as_float(u[j]);
      }
    }
    if (tid == 0) out_clk[0] = clock64() - t0;
  } else if (warp == 8) {
    const uint32_t idesc = umma_idesc_f16(128, 64);
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
      mbar_wait(bar_opnd, par);
      par ^= 1;
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (mode == 0) tc_mma_f16(tm, umma_desc_sw128(base + 32 * kk), umma_desc_sw128(base + 16384 + 32 * kk), idesc, kk > 0);
          else tc_mma_f16_ts(tm, tm + 64 + 8 * kk, umma_desc_sw128(base + 16384 + 32 * kk), idesc, kk > 0);
        }
        tc_commit(bar_acc);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 128);
}

int main() {
  std::vector<float> a(128 * 64), b(64 * 64), ref(128 * 64), got(128 * 64);
  srand(3);
  for (auto& x : a) x = (rand() % 2001 - 1000) / 1000.f;
  for (auto& x : b) x = (rand() % 2001 - 1000) / 4000.f;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int kk = 0; kk < 64; ++kk) s += (double)__half2float(__float2half_rn(a[m * 64 + kk])) * __half2float(__float2half_rn(b[n * 64 + kk]));
      ref[m * 64 + n] = (float)s;
    }
  float *da, *db, *dout;
  long long* dclk;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4); cudaMalloc(&dout, got.size() * 4); cudaMalloc(&dclk, 64);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(dout, 0, got.size() * 4);
      k<<<1, 288, 40000>>>(mode, iters, da, db, dout, dclk);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long clk = 0;
    cudaMemcpy(&clk, dclk, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(got.data(), dout, got.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (size_t i = 0; i < got.size(); ++i) maxerr = fmax(maxerr, fabs(got[i] - ref[i]));
    printf("A through %s: %8.1f clk/phase, first-phase max abs err %.3e (sample got %.4f ref %.4f)\n", mode == 0 ? "shared memory" : "tensor memory",
           (double)clk / iters, maxerr, got[5 * 64 + 7], ref[5 * 64 + 7]);
  }
  return 0;
}
