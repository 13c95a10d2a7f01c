// Microbenchmark: throughput of ex2.approx in f32, f16x2 and bf16x2 form (results per clock per SM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float a0 = seed + threadIdx.x * 1e-3f, a1 = a0 + 0.1f, a2 = a0 + 0.2f, a3 = a0 + 0.3f;
  uint32_t h0 = 0x3c003c00u + threadIdx.x, h1 = h0 + 1, h2 = h0 + 2, h3 = h0 + 3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
    } else if (MODE == 1) {
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h3));
    } else {
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h0));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h1));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h2));
      asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h3));
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (float)(t1 - t0);
  out[1 + blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(h0 ^ h1 ^ h2 ^ h3);
}

int main() {
  float* d;
  cudaMalloc(&d, sizeof(float) * (1 + 148 * 1024));
  const int iters = 4096;
  const char* names[3] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2", "ex2.approx.ftz.bf16x2"};
  for (int mode = 0; mode < 3; ++mode) {
    for (int threads = 128; threads <= 1024; threads *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(d, iters, 0.5f);
        if (mode == 1) k<1><<<148, threads>>>(d, iters, 0.5f);
        if (mode == 2) k<2><<<148, threads>>>(d, iters, 0.5f);
        cudaDeviceSynchronize();
      }
      float clk;
      cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
      const double instr = 4.0 * iters * threads;          // lane-instructions per SM
      const double results = instr * (mode == 0 ? 1 : 2);
      printf("%-24s threads/SM %4d: %.2f lane-instr/clk/SM, %.2f results/clk/SM\n", names[mode], threads, instr / clk,
             results / clk);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
