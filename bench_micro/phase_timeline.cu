// Where the ~1000 clk of one MMA-phase hand-shake go: clock64() stamps inside one CTA (8 epilogue warps + 1 MMA warp).
//   mode 0: separate MMA-issuer thread, mbarrier hand-off both ways (the round-1 protocol)
//   mode 1: no MMA warp: epilogue warps meet at a named barrier (bar.sync) and thread 0 issues the MMAs itself
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I trajsde_b200/csrc -o bench_micro/phase_timeline bench_micro/phase_timeline.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../trajsde_b200/csrc/tc_common.cuh"
using namespace trajsde::tc;

constexpr int ITERS = 512;

struct Cfg { int mode; int n_dim; int n_mma; int fence_proxy; int sts; };

__global__ void __launch_bounds__(288, 1) k(Cfg c, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bars[2];
  __shared__ long long acc_stamp[8];   // sums: see host print
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_opnd = smem_u32(&bars[0]), bar_acc = smem_u32(&bars[1]);
  for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid < 8) acc_stamp[tid] = 0;
  if (tid == 0) {
    mbar_init(bar_opnd, 256);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&tmem_ptr), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  const uint32_t idesc = umma_idesc_f16(128, (uint32_t)c.n_dim);
  if (warp < 8) {
    const int quad = warp & 3, part = warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t taddr = tm + ((uint32_t)(quad * 32) << 16) + part * 32;
    uint8_t* a_row = sm + row * 128;
    uint32_t par = 0;
    long long s_wait = 0, s_ld = 0, s_work = 0, s_total = 0, s_issue = 0;
    long long t_arrive = clock64();
    if (c.mode == 0) mbar_arrive(bar_opnd);
    else {
      named_bar_sync(1, 256);
      if (tid == 0) {
        tc_fence_after();
        for (int kk = 0; kk < c.n_mma; ++kk) tc_mma_f16(tm, umma_desc_sw128(base + 32 * kk), umma_desc_sw128(base + 16384 + 32 * kk), idesc, kk > 0);
        tc_commit(bar_acc);
      }
    }
    for (int it = 0; it < ITERS; ++it) {
      mbar_wait(bar_acc, par);
      par ^= 1;
      tc_fence_after();
      const long long t1 = clock64();
      uint32_t u[32];
      tmem_ld_32x32b_x32(taddr, u);
      tc_wait_ld();
      const long long t2 = clock64();
      if (c.sts) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(a_row + (((part * 4 + q) ^ (row & 7u)) << 4)) = make_uint4(u[8 * q] & 0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
      }
      if (c.fence_proxy) fence_proxy_async();
      tc_fence_before();
      const long long t3 = clock64();
      s_wait += t1 - t_arrive;
      s_ld += t2 - t1;
      s_work += t3 - t2;
      t_arrive = clock64();
      if (c.mode == 0) mbar_arrive(bar_opnd);
      else {
        named_bar_sync(1, 256);
        if (tid == 0 && it + 1 < ITERS) {
          tc_fence_after();
          const long long ta = clock64();
          for (int kk = 0; kk < c.n_mma; ++kk) tc_mma_f16(tm, umma_desc_sw128(base + 32 * kk), umma_desc_sw128(base + 16384 + 32 * kk), idesc, kk > 0);
          tc_commit(bar_acc);
          s_issue += clock64() - ta;
        }
      }
      s_total += 1;
    }
    if (tid == 0) { out[0] = s_wait; out[1] = s_ld; out[2] = s_work; out[3] = s_issue; }
  } else if (warp == 8 && lane == 0 && c.mode == 0) {
    uint32_t par = 0;
    long long s_issue = 0;
    for (int it = 0; it <= ITERS; ++it) {
      mbar_wait(bar_opnd, par);
      par ^= 1;
      tc_fence_after();
      if (it == ITERS) break;
      const long long ta = clock64();
      for (int kk = 0; kk < c.n_mma; ++kk) tc_mma_f16(tm, umma_desc_sw128(base + 32 * kk), umma_desc_sw128(base + 16384 + 32 * kk), idesc, kk > 0);
      tc_commit(bar_acc);
      s_issue += clock64() - ta;
    }
    out[3] = s_issue;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  struct Named { const char* name; Cfg c; };
  std::vector<Named> v = {
      {"mma-warp  N=64  4 mma fence sts", {0, 64, 4, 1, 1}},
      {"mma-warp  N=64  1 mma fence sts", {0, 64, 1, 1, 1}},
      {"mma-warp  N=64  0 mma fence sts", {0, 64, 0, 1, 1}},
      {"mma-warp  N=128 4 mma fence sts", {0, 128, 4, 1, 1}},
      {"mma-warp  N=256 4 mma fence sts", {0, 256, 4, 1, 1}},
      {"mma-warp  N=64  4 mma no fence no sts", {0, 64, 4, 0, 0}},
      {"self-issue N=64  4 mma fence sts", {1, 64, 4, 1, 1}},
      {"self-issue N=64  1 mma fence sts", {1, 64, 1, 1, 1}},
      {"self-issue N=64  0 mma fence sts", {1, 64, 0, 1, 1}},
      {"self-issue N=128 4 mma fence sts", {1, 128, 4, 1, 1}},
  };
  printf("%-40s %10s %10s %10s %10s %10s\n", "variant (clk per phase, thread 0)", "arrive->acc", "ldtm", "sts+fence", "mma issue", "sum");
  for (auto& nv : v) {
    long long h[4] = {0, 0, 0, 0};
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(d, 0, 64);
      k<<<1, 288, 60000>>>(nv.c, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", nv.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("%-40s %10.1f %10.1f %10.1f %10.1f %10.1f\n", nv.name, (double)h[0] / ITERS, (double)h[1] / ITERS, (double)h[2] / ITERS,
           (double)h[3] / ITERS, (double)(h[0] + h[1] + h[2]) / ITERS);
  }
  return 0;
}
