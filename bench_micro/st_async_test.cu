// Can epilogue threads hand an operand tile to tcgen05.mma with st.async (async-proxy store + mbarrier complete_tx) instead of
// st.shared + fence.proxy.async + mbarrier.arrive?  Checks numerics (every phase writes NEW data, D = A . I^T must equal it) and
// measures clocks per phase for both protocols.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I trajsde_b200/csrc -o bench_micro/st_async_test bench_micro/st_async_test.cu
#include <cstdio>
#include <cstdlib>
#include "../trajsde_b200/csrc/tc_common.cuh"
using namespace trajsde::tc;

__device__ __forceinline__ void st_async_v4(uint32_t dst, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(dst), "r"(a), "r"(b),
               "r"(c), "r"(d), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(288, 1) k(int mode, int iters, long long* out_clk, int* out_bad) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_opnd = smem_u32(&bars[0]), bar_acc = smem_u32(&bars[1]);
  // B = identity [64][64] (K-major SW128) at sm + 16384
  for (int i = tid; i < 64 * 64; i += blockDim.x) {
    const int n = i >> 6, kk = i & 63;
    *reinterpret_cast<__half*>(sm + 16384 + sw128_off_h(n, kk)) = __float2half_rn(n == kk ? 1.f : 0.f);
  }
  if (tid == 0) {
    mbar_init(bar_opnd, mode == 0 ? 256 : 1);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_ptr), 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp < 8) {
    const int quad = warp & 3, hh = warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t taddr = tm + ((uint32_t)(quad * 32) << 16) + hh * 32;
    uint8_t* a_row = sm + row * 128;
    const uint32_t a_row_u32 = base + row * 128;
    int bad = 0;
    uint32_t par = 0;
    long long t0 = clock64();
    for (int it = 0; it <= iters; ++it) {
      // write A[row][hh*32 + j] = ((it * 7 + row + j) & 63) - 32   (exact in fp16)
      if (it < iters) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t p[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = q * 8 + 2 * e;
            p[e] = pack_f16x2((float)(((it * 7 + (int)row + hh * 32 + j) & 63) - 32), (float)(((it * 7 + (int)row + hh * 32 + j + 1) & 63) - 32));
          }
          const uint32_t off = (((uint32_t)(hh * 4 + q)) ^ (row & 7u)) << 4;
          if (mode == 0) *reinterpret_cast<uint4*>(a_row + off) = make_uint4(p[0], p[1], p[2], p[3]);
          else st_async_v4(a_row_u32 + off, p[0], p[1], p[2], p[3], bar_opnd);
        }
        if (mode == 0) fence_proxy_async();
        tc_fence_before();
        if (mode == 0) mbar_arrive(bar_opnd);
      }
      if (it > 0) {   // D of the previous phase must equal the A written in the previous phase
        // (the write above for phase `it` may only be issued once the MMA of phase it-1 has READ A: it has, we waited for its commit)
      }
      if (it < iters) {
        mbar_wait(bar_acc, par);
        par ^= 1;
        tc_fence_after();
        uint32_t u[32];
        tmem_ld_32x32b_x32(taddr, u);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float want = (float)(((it * 7 + (int)row + hh * 32 + j) & 63) - 32);
          bad += (__uint_as_float(u[j]) != want);
        }
      }
    }
    if (tid == 0) out_clk[0] = clock64() - t0;
    atomicAdd(out_bad, bad);
  } else if (warp == 8) {
    const uint32_t idesc = umma_idesc_f16(128, 64);
    uint32_t par = 0;
    for (int it = 0; it < iters; ++it) {
      if (mode == 1 && elect_one()) mbar_arrive_expect_tx(bar_opnd, 16384);
      __syncwarp();
      mbar_wait(bar_opnd, par);
      par ^= 1;
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc_mma_f16(tm, umma_desc_sw128(base + 32 * kk), umma_desc_sw128(base + 16384 + 32 * kk), idesc, kk > 0);
        tc_commit(bar_acc);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 64);
}

int main() {
  long long* d_clk;
  int* d_bad;
  cudaMalloc(&d_clk, 64);
  cudaMalloc(&d_bad, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int iters = 4000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d_bad, 0, 4);
      k<<<1, 288, 40000>>>(mode, iters, d_clk, d_bad);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    }
    long long clk = 0;
    int bad = 0;
    cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost);
    printf("%s: %8.1f clk/phase, mismatching accumulator elements %d of %d\n",
           mode == 0 ? "st.shared + fence.proxy.async + mbarrier.arrive" : "st.async (complete_tx on the operand mbarrier)      ", (double)clk / iters, bad,
           iters * 128 * 64);
  }
  return 0;
}
