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

enum TuneKey { TUNE_ENTROPY_PRECISE = 0, TUNE_CHAMFER_ROWS = 1, TUNE_MLP_FORCE_FP32 = 2, TUNE_MLP_TC_MASK = 3, TUNE_MLP_EPI_DEBUG = 4, TUNE_CHAMFER_SEED = 5, TUNE_MLP_NO_FORK = 6, TUNE_COMM_BLOCKS = 7, TUNE_PDL = 8, TUNE_COMM_NO_SMALL_P2P = 9, TUNE_SPARSE_SORTED = 10, TUNE_CHAMFER_BWD_TWO_PASS = 11, TUNE_NKEYS = 12 };

// ---- programmatic dependent launch -----------------------------------------------------------------
// The hot path at the reference's shapes is a chain of ~70 dependent launches of 3-20 us per D4 pass: what a step
// costs is launch-to-launch latency, not work.  Every libpcuda kernel calls pdl_entry() — `griddepcontrol.wait` (block
// until the preceding kernel of the stream has completed and its writes are visible) followed by
// `griddepcontrol.launch_dependents` (the next kernel may be scheduled once all CTAs of this one got here) — and kernels
// with work that does not depend on their predecessor (staging weights, barrier / TMEM set-up) do that work BEFORE
// pdl_entry() and are launched with the programmatic-stream-serialisation attribute (PCUDA_LAUNCH_PDL): ordering and
// visibility of everything after pdl_entry() are exactly those of ordinary stream order, while launch latency and the
// prologue overlap the predecessor.  Without the attribute both instructions are no-ops.  Under stream capture the edges
// become programmatic dependency edges of the CUDA graph.
__device__ __forceinline__ void pdl_trigger() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// Every kernel starts with pdl_entry(): wait for the grid before it in the stream (a no-op unless this launch carries
// the programmatic-serialization attribute), THEN let the grid after it start early.  In this order at most two kernels
// of a stream overlap, and when a dependent starts, everything before its predecessor has completed: code a kernel runs
// BEFORE pdl_entry() may read anything except what its immediate predecessor writes (parameters, its own shared memory,
// barrier / TMEM set-up), and must not write global memory.  Two kinds of kernel call pdl_wait() alone and so release
// their successor only when they complete: the optimiser kernels (they write the parameters others stage early) and the
// kernels that spin on a peer GPU (comm.cu: a successor resident while they wait can deadlock two ranks).
__device__ __forceinline__ void pdl_entry() {
  pdl_wait();
  pdl_trigger();
}

// kPrologue marks launches of kernels that do useful work before pdl_entry() (weight staging, barrier / TMEM set-up).
// pcuda_tune(8, v): 0 default = every launch carries the programmatic-serialization attribute, 1 = none, 4 = only the
// kPrologue ones, 3 = only launches of at least two waves of CTAs.
// Measured (tools/ab_pdl.py, graph replay, L2 flushed, 40 steps, two rounds), cfg2:
//   trigger-before-wait in every kernel, every launch (profiles/r2_ab_pdl.txt):      off 0.475 / 0.478 ms   on 0.495 / 0.496
//   wait-then-trigger (this code), fc_fwd staging its weights early (r2_ab_pdl2.txt): off 0.457 / 0.459   fc_fwd only 0.447 /
//   0.453   every launch 0.426 / 0.424
// With the trigger first, dependents of dependents pile up on the SMs the concurrent branches of the step need; with the
// wait first at most one successor per stream is resident early, and launch latency plus prologue are hidden.
template <bool kPrologue, class... KArgs, class... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int mode = tuning(TUNE_PDL);
  const long long ctas = static_cast<long long>(grid.x) * grid.y * grid.z;
  cfg.numAttrs = (mode == 0 || mode == 2 || (mode == 4 && kPrologue) || (mode == 3 && ctas >= 2ll * sm_count())) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// PCUDA_LAUNCH((kernel<T...>), grid, block, smem, stream, args...): parenthesise kernel names that contain commas
#define PCUDA_LAUNCH(kernel, grid, block, smem, st, ...) \
  (void)::pcuda::launch_k<false>(kernel, dim3(grid), dim3(block), static_cast<size_t>(smem), st, __VA_ARGS__)
#define PCUDA_LAUNCH_PDL(kernel, grid, block, smem, st, ...) \
  (void)::pcuda::launch_k<true>(kernel, dim3(grid), dim3(block), static_cast<size_t>(smem), st, __VA_ARGS__)

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
