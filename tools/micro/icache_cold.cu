// How long does a kernel take that executes N straight-line instructions exactly once per warp?  Measures the cost of
// instruction fetch for code that is not in the SM's instruction caches: (a) the same kernel launched back to back
// (code in L2 and possibly in the SM caches), (b) after another big kernel ran in between (SM caches replaced),
// (c) after a write of 512 MB (L2 flushed: code comes from HBM).   nvcc -arch=sm_100a -O3 -o icache_cold icache_cold.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int N, int SALT>
__global__ void straight(float* out, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
#pragma unroll
  for (int i = 0; i < N / 4; ++i) {
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x0) : "f"(a), "f"(b));
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x1) : "f"(a), "f"(b));
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x2) : "f"(a), "f"(b));
    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x3) : "f"(a), "f"(b));
  }
  if (x0 + x1 + x2 + x3 == 12345.f) out[threadIdx.x] = x0;
}

__global__ void fill(float* p, size_t n, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template <int N>
void run(float* out, float* big, size_t nbig, int threads) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  float warm = 0, other = 0, cold = 0;
  const int reps = 10;
  for (int r = 0; r < reps; ++r) {
    straight<N, 0><<<148, threads>>>(out, 1.0001f, 0.5f);
    cudaEventRecord(e0); straight<N, 0><<<148, threads>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); warm += ms;
    straight<16384, 1><<<148, threads>>>(out, 1.0001f, 0.5f);          // 256 KB of other code through the SM caches
    cudaEventRecord(e0); straight<N, 0><<<148, threads>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); other += ms;
    fill<<<148 * 8, 256>>>(big, nbig, 1.f);
    cudaEventRecord(e0); straight<N, 0><<<148, threads>>>(out, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1); cold += ms;
  }
  printf("N=%6d instr (%4d KB) threads=%4d: back-to-back %7.2f us   after other code %7.2f us   after L2 flush %7.2f us\n", N, N * 16 / 1024,
         threads, warm / reps * 1e3, other / reps * 1e3, cold / reps * 1e3);
}

int main() {
  float* out; cudaMalloc(&out, 4096);
  const size_t nbig = (size_t)128 << 20;   // 512 MB
  float* big; cudaMalloc(&big, nbig * 4);
  for (int threads : {32, 512}) {
    run<256>(out, big, nbig, threads);
    run<1024>(out, big, nbig, threads);
    run<2048>(out, big, nbig, threads);
    run<4096>(out, big, nbig, threads);
    run<8192>(out, big, nbig, threads);
  }
  return 0;
}
