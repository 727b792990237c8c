// Shared helpers for the libpcuda kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pcuda.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpcuda is written for sm_100a (B200) only"
#endif

namespace pcuda {

// ---- error plumbing (host) --------------------------------------------------------------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);  // returns cudaGetLastError() as a positive code (0 = ok)
void count_launch(int n = 1);        // bench accounting: kernels launched by this library
int sm_count();
int tuning(int key);
cudaError_t smem_optin_impl(const void* fn, int bytes);   // per (kernel, device) opt-in to large dynamic shared memory
template <class F>
static inline cudaError_t smem_optin(F* fn, int bytes) {
  return smem_optin_impl(reinterpret_cast<const void*>(fn), bytes);
}

enum TuneKey { TUNE_ENTROPY_PRECISE = 0, TUNE_CHAMFER_ROWS = 1, TUNE_MLP_FORCE_FP32 = 2, TUNE_MLP_TC_MASK = 3, TUNE_MLP_EPI_DEBUG = 4, TUNE_CHAMFER_SEED = 5, TUNE_MLP_NO_FORK = 6, TUNE_COMM_BLOCKS = 7, TUNE_FC_NO_CLUSTER = 8, TUNE_FC_CLUSTER_SIZE = 9, TUNE_NKEYS = 12 };

#define PCUDA_REQUIRE(cond, code, ...)                 \
  do {                                                 \
    if (!(cond)) return ::pcuda::fail(code, __VA_ARGS__); \
  } while (0)

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// streaming 128-bit global access: read-once / write-once data should not displace L2 lines
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  __stcs(reinterpret_cast<float4*>(p), v);
}

// Order-preserving map float -> uint32 (total order, -0 < +0), and back.
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pcuda
