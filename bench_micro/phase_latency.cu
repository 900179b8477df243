// Latency of one "MMA phase" hand-shake of the solver kernels, in isolation: epilogue warps (tcgen05.ld -> optional MUFU work ->
// st.shared operand -> fence.proxy.async -> mbarrier arrive) <-> one MMA-issuer thread (mbarrier wait -> 4 x tcgen05.mma ->
// tcgen05.commit).  Prints SM clocks per phase for several protocol variants; the numbers size the per-step dependency chain of
// euler_tc.cu / enc_tc.cu / euler_bwd_tc.cu.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I trajsde_b200/csrc -o bench_micro/phase_latency bench_micro/phase_latency.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../trajsde_b200/csrc/tc_common.cuh"
using namespace trajsde::tc;

struct Cfg {
  int n_warps;       // epilogue warps (8 or 16): 8 -> thread = (row, 32-col half), 16 -> (row, 16-col quarter)
  int arrive_mode;   // 0: every thread arrives (count = threads), 1: __syncwarp + elected lane (count = warps)
  int fence_proxy;   // fence.proxy.async before the arrive
  int n_dim;         // MMA N
  int n_tanh;        // MUFU.TANH per thread per phase
  int sts;           // write the operand chunks (st.shared.v4) per phase
  int iters;
  int membar;        // extra __threadfence_block() per phase
  int wait_mode;     // 0: every epilogue warp polls the mbarrier; 1: warp 0 polls, the others wait at a named barrier
};

__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(544, 1) k(Cfg c, long long* out_clk, float* sink) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - raw);
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_opnd = smem_u32(&bars[0]), bar_acc = smem_u32(&bars[1]);
  const int n_epi = c.n_warps * 32;
  for (int i = tid; i < 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;  // fp16 1.0
  if (tid == 0) {
    mbar_init(bar_opnd, c.arrive_mode ? c.n_warps : n_epi);
    mbar_init(bar_acc, 1);
    mbar_fence_init();
  }
  if (warp == c.n_warps) tmem_alloc(smem_u32(&tmem_ptr), 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_ptr;
  long long t0 = 0;
  if (warp < c.n_warps) {
    const int quad = warp & 3, part = warp >> 2;
    const uint32_t row = quad * 32 + lane;
    const uint32_t width = c.n_warps == 8 ? 32 : 16;
    const uint32_t taddr = tm + ((uint32_t)(quad * 32) << 16) + part * width;
    uint8_t* a_row = sm + row * 128;
    float acc = 0.f;
    uint32_t par = 0;
    // prime: first arrive so the MMA thread can start
    if (c.arrive_mode) { __syncwarp(); if (lane == 0) mbar_arrive(bar_opnd); } else mbar_arrive(bar_opnd);
    if (tid == 0) t0 = clock64();
    for (int it = 0; it < c.iters; ++it) {
      if (c.wait_mode == 0) {
        mbar_wait(bar_acc, par);
      } else {
        if (warp == 0) mbar_wait(bar_acc, par);
        named_bar_sync(1, n_epi);
      }
      par ^= 1;
      tc_fence_after();
      float v[32];
      if (width == 32) {
        uint32_t u[32];
        tmem_ld_32x32b_x32(taddr, u);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(u[j]);
      } else {
        uint32_t u[16];
        tmem_ld_32x32b_x16(taddr, u);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j] = __uint_as_float(u[j]); v[j + 16] = v[j]; }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < c.n_tanh) v[j] = tanh_approx(v[j] + 0.25f);
      if (c.sts) {
        const int nchunk = width / 8;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < nchunk)
            *reinterpret_cast<uint4*>(a_row + (((part * nchunk + q) ^ (row & 7u)) << 4)) =
                make_uint4(pack_f16x2(v[8 * q], v[8 * q + 1]), pack_f16x2(v[8 * q + 2], v[8 * q + 3]), pack_f16x2(v[8 * q + 4], v[8 * q + 5]),
                           pack_f16x2(v[8 * q + 6], v[8 * q + 7]));
      } else {
        acc += v[0] + v[7] + v[31];
      }
      if (c.membar) __threadfence_block();
      if (c.fence_proxy) fence_proxy_async();
      tc_fence_before();
      if (c.arrive_mode) { __syncwarp(); if (lane == 0) mbar_arrive(bar_opnd); } else mbar_arrive(bar_opnd);
    }
    if (tid == 0) out_clk[0] = clock64() - t0;
    if (acc == 123.456f) sink[tid] = acc;
  } else if (warp == c.n_warps && lane == 0) {
    const uint32_t idesc = umma_idesc_f16(128, (uint32_t)c.n_dim);
    uint32_t par = 0;
    for (int it = 0; it <= c.iters; ++it) {
      mbar_wait(bar_opnd, par);
      par ^= 1;
      tc_fence_after();
      if (it == c.iters) break;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc_mma_f16(tm, umma_desc_sw128(base + 32 * kk), umma_desc_sw128(base + 16384 + 32 * kk), idesc, kk > 0);
      tc_commit(bar_acc);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == c.n_warps) tmem_dealloc(tm, 256);
}

int main() {
  long long* d_clk;
  float* sink;
  cudaMalloc(&d_clk, 64);
  cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  struct Named { const char* name; Cfg c; };
  std::vector<Named> v = {
      {"8w  all-arrive  proxyfence N=64  no work       ", {8, 0, 1, 64, 0, 0, 2000, 0, 0}},
      {"8w  elect-arrive proxyfence N=64 no work       ", {8, 1, 1, 64, 0, 0, 2000, 0, 0}},
      {"8w  elect-arrive no fence  N=64  no work       ", {8, 1, 0, 64, 0, 0, 2000, 0, 0}},
      {"8w  all-arrive  no fence   N=64  no work       ", {8, 0, 0, 64, 0, 0, 2000, 0, 0}},
      {"8w  all-arrive  proxyfence N=128 no work       ", {8, 0, 1, 128, 0, 0, 2000, 0, 0}},
      {"8w  all-arrive  proxyfence N=64  sts           ", {8, 0, 1, 64, 0, 1, 2000, 0, 0}},
      {"8w  elect-arrive proxyfence N=64 sts           ", {8, 1, 1, 64, 0, 1, 2000, 0, 0}},
      {"8w  all-arrive  proxyfence N=64  sts+32tanh    ", {8, 0, 1, 64, 32, 1, 2000, 0, 0}},
      {"8w  elect-arrive proxyfence N=64 sts+32tanh    ", {8, 1, 1, 64, 32, 1, 2000, 0, 0}},
      {"8w  all-arrive  proxyfence+membar N=64 sts     ", {8, 0, 1, 64, 0, 1, 2000, 1, 0}},
      {"16w all-arrive  proxyfence N=64  no work       ", {16, 0, 1, 64, 0, 0, 2000, 0, 0}},
      {"16w elect-arrive proxyfence N=64 no work       ", {16, 1, 1, 64, 0, 0, 2000, 0, 0}},
      {"16w all-arrive  proxyfence N=64  sts           ", {16, 0, 1, 64, 0, 1, 2000, 0, 0}},
      {"16w elect-arrive proxyfence N=64 sts           ", {16, 1, 1, 64, 0, 1, 2000, 0, 0}},
      {"16w all-arrive  proxyfence N=64  sts+16tanh    ", {16, 0, 1, 64, 16, 1, 2000, 0, 0}},
      {"16w elect-arrive proxyfence N=64 sts+16tanh    ", {16, 1, 1, 64, 16, 1, 2000, 0, 0}},
      {"16w elect-arrive proxyfence N=128 sts+16tanh   ", {16, 1, 1, 128, 16, 1, 2000, 0, 0}},
      {"8w  all-arrive one-warp-polls N=64 sts+32tanh  ", {8, 0, 1, 64, 32, 1, 2000, 0, 1}},
      {"8w  all-arrive one-warp-polls N=64 no work     ", {8, 0, 1, 64, 0, 0, 2000, 0, 1}},
      {"16w all-arrive one-warp-polls N=64 sts+16tanh  ", {16, 0, 1, 64, 16, 1, 2000, 0, 1}},
      {"16w all-arrive one-warp-polls N=64 no work     ", {16, 0, 1, 64, 0, 0, 2000, 0, 1}},
      {"16w elect-arrive one-warp-polls N=64 sts+16tanh", {16, 1, 1, 64, 16, 1, 2000, 0, 1}},
  };
  for (auto& nv : v) {
    for (int rep = 0; rep < 2; ++rep) {
      k<<<1, (nv.c.n_warps + 1) * 32, 60000>>>(nv.c, d_clk, sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: %s\n", nv.name, cudaGetErrorString(e)); return 1; }
    }
    long long clk = 0;
    cudaMemcpy(&clk, d_clk, 8, cudaMemcpyDeviceToHost);
    printf("%s %8.1f clk/phase\n", nv.name, (double)clk / nv.c.iters);
  }
  return 0;
}
