// Micro-benchmark: does alternating kernels with different shared-memory carve-outs cost time on B200?
//   A: 1 CTA/SM, 200 KB dynamic smem (like the tcgen05 GEMMs);  B: small kernel, 38 KB dynamic smem (like pool_sparse)
// Sequences of 2000 launches on one stream, timed with events:  B only | A only | A,B alternating (default carve-out)
// | A,B alternating with B pinned to cudaSharedmemCarveoutMaxShared.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void kA(float* p) { extern __shared__ float s[]; s[threadIdx.x] = p[threadIdx.x]; __syncthreads(); if (s[0] == 123.f) p[0] = 1.f; }
__global__ void kB(float* p) { extern __shared__ float s[]; s[threadIdx.x] = p[threadIdx.x]; __syncthreads(); if (s[0] == 123.f) p[1] = 1.f; }
__global__ void kC(float* p) { if (p[threadIdx.x] == 123.f) p[2] = 1.f; }   // no shared memory at all
static float run(int mode, float* d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaDeviceSynchronize(); cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) {
    if (mode == 0) kB<<<128, 512, 38 * 1024>>>(d);
    if (mode == 1) kA<<<148, 256, 200 * 1024>>>(d);
    if (mode == 2) { kA<<<148, 256, 200 * 1024>>>(d); kB<<<128, 512, 38 * 1024>>>(d); }
    if (mode == 3) { kA<<<148, 256, 200 * 1024>>>(d); kC<<<128, 512>>>(d); }
    if (mode == 4) kC<<<128, 512>>>(d);
  }
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1e3f / iters;
}
int main() {
  float* d; cudaMalloc(&d, 1 << 20); cudaMemset(d, 0, 1 << 20);
  cudaFuncSetAttribute(kA, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(kB, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int it = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    printf("default carve-out:  B only %.2f us | A only %.2f us | A,B pair %.2f us | A,C pair %.2f us | C only %.2f us\n",
           run(0, d, it), run(1, d, it), run(2, d, it), run(3, d, it), run(4, d, it));
  }
  cudaFuncSetAttribute(kB, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute(kC, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  for (int rep = 0; rep < 2; ++rep) {
    printf("B,C pinned to max:  B only %.2f us | A only %.2f us | A,B pair %.2f us | A,C pair %.2f us | C only %.2f us\n",
           run(0, d, it), run(1, d, it), run(2, d, it), run(3, d, it), run(4, d, it));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
