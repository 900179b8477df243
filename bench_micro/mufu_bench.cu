// Microbenchmark: MUFU tanh throughput per SM for f32 / f16x2 / bf16x2 forms (decides the epilogue design).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_micro/mufu_bench bench_micro/mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters) {
  float a0 = threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t h0 = threadIdx.x * 7 + 0x3c003800u, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3;
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a0));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a1));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a2));
      asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 2) {
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h3));
    } else if (MODE == 3) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 4) {  // FMA-pipe reference: 4 dependent-chain FFMAs
      a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
    } else if (MODE == 5) {  // cvt.rn.f16x2.f32 pack throughput
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(a0), "f"(a1));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(a1), "f"(a2));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(a2), "f"(a3));
      asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h3) : "f"(a3), "f"(a0));
      a0 += __uint_as_float(h0 & 1); a1 += __uint_as_float(h1 & 1);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}

template <int MODE>
void run(const char* name, int per_instr) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 8 * 256 * sizeof(float));
  const int iters = 20000;
  k<MODE><<<sms * 8, 256>>>(out, 100);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sms * 8, 256>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double instr = (double)sms * 8 * 256 * iters * 4;             // thread-level instructions
  double elems = instr * per_instr;
  printf("%-22s %8.3f ms  %.2f Telem/s  = %.1f elem/clk/SM at max clock %.0f MHz (%.1f at 1.5GHz)\n", name, ms, elems / ms * 1e-9,
         elems / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz * 1e-3, elems / (ms * 1e-3) / sms / 1.5e9);
  cudaFree(out);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("tanh.approx.f16x2", 2);
  run<2>("tanh.approx.bf16x2", 2);
  run<3>("ex2.approx.f32", 1);
  run<4>("ffma", 1);
  run<5>("cvt.rn.f16x2.f32(+2 fadd)", 2);
  return 0;
}
