// Entropy / self-information map, forward and backward — one fused HBM pass each.
//
// Replaces the inline tensor expressions of the reference train loops
//   train_mscmrseg.py:222,265      m = -1.0 * sigmoid(z) * log(sigmoid(z) + 1e-7)
//   train_mmwhs.py:213-217,224,242 p = softmax(z,1) | sigmoid(z); m = -1.0*p*log(p+1e-7)/log(C)
//   train_mmwhs.py:225,243         mean(sum(m, dim=1))
// (6-9 element-wise launches + softmax in the reference, each a full round trip of [B,C,H,W]).
//
// Layout: NCHW contiguous, so for a fixed pixel the C class values are HW floats apart and the
// pixels of one class are contiguous.  Each thread owns 4 consecutive pixels (one float4 per
// class): every warp-level access is a 512-byte contiguous run, the class reduction of the
// softmax stays in registers, and the only cross-thread reduction is the optional mean-entropy
// scalar (warp shuffle -> one double atomic per block -> last block finalises).
//
// Roofline: HBM. fwd 8 B/element (+4 if p is emitted), bwd 12 B/element (+4 with grad_p).
#include <type_traits>

#include "pcuda_common.cuh"

namespace pcuda {
namespace {

constexpr int kThreads = 256;

struct MeanWs {
  double sum;
  unsigned int ticket;
  unsigned int pad;
};
static_assert(sizeof(MeanWs) == PCUDA_ENTROPY_WS_BYTES, "workspace layout");

// Two arithmetic flavours, selected per call (pcuda_tune key 0):
//   FAST (default)  MUFU-based: ex2.approx / lg2.approx / rcp.approx.  exp and the reciprocal are
//                   within ~2 ulp; lg2.approx has an ABSOLUTE error of 2^-22 in log2 on [0.5, 2],
//                   i.e. <= 1.7e-7 on a map entry — inside the 1e-5*|m| + 5e-7 parity band
//                   (the reference itself is 1.6e-7 off fp64) at ~11 instructions
//                   per element, which leaves the kernel HBM-bound.
//   precise         libdevice expf/logf and IEEE reciprocal (~45 instructions per element,
//                   issue-bound on B200); kept as the cross-check.
__device__ __forceinline__ float mufu_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mufu_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

template <bool FAST>
__device__ __forceinline__ float pc_log(float t) {
  if constexpr (!FAST) return logf(t);
  else return mufu_lg2(t) * kLn2;
}
// exp(x - mx): softmax numerator
template <bool FAST>
__device__ __forceinline__ float pc_exp_sub(float x, float mx) {
  if constexpr (!FAST) return expf(x - mx);
  else return mufu_ex2((x - mx) * kLog2e);
}
template <bool FAST>
__device__ __forceinline__ float pc_rcp(float x) {
  if constexpr (!FAST) return __frcp_rn(x);
  else return mufu_rcp(x);
}
template <bool FAST>
__device__ __forceinline__ float pc_sigmoid(float z) {
  // 1/(1+exp(-z)) as ATen evaluates it; exp overflow -> inf -> 0 is the reference behaviour too.
  if constexpr (!FAST) return __frcp_rn(1.0f + expf(-z));
  else return mufu_rcp(1.0f + mufu_ex2(-z * kLog2e));
}

// Finalise the mean-entropy scalar: every block adds its partial, the last one to take a ticket
// converts and resets the workspace so the next call finds it zeroed.
__device__ __forceinline__ void block_mean_commit(float local, MeanWs* ws, float* mean_out,
                                                  double inv_count) {
  __shared__ float warp_part[kThreads / 32];
  float w = warp_sum(local);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) warp_part[wid] = w;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) s += static_cast<double>(warp_part[i]);
    atomicAdd(&ws->sum, s);
    __threadfence();
    unsigned int t = atomicAdd(&ws->ticket, 1u);
    if (t == gridDim.x - 1) {
      __threadfence();
      double total = atomicAdd(&ws->sum, 0.0);
      *mean_out = static_cast<float>(total * inv_count);
      ws->sum = 0.0;
      ws->ticket = 0u;
      __threadfence();
    }
  }
}

// ---- forward, vectorised: C compile-time, 4 pixels per thread ---------------------------------
template <int ACT, int C, bool WRITE_P, bool WITH_MEAN, bool FAST>
__global__ void __launch_bounds__(kThreads)
entropy_fwd_vec4(const float* __restrict__ z, float* __restrict__ m, float* __restrict__ p_out,
                 float* __restrict__ mean_out, MeanWs* __restrict__ ws, int64_t n_quads,
                 int64_t quads_per_img, int64_t HW, float inv_norm, float smooth,
                 double inv_count) {
  pdl_entry();
  float local = 0.0f;
  // (image, quad-in-image) advance incrementally: no 64-bit division inside the streaming loop
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  int64_t q = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x;
  int64_t b = q / quads_per_img;
  int64_t r = q - b * quads_per_img;
  for (; q < n_quads; q += stride, r += stride) {
    while (r >= quads_per_img) { r -= quads_per_img; ++b; }
    const int64_t base = b * C * HW + r * 4;
    float4 v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = ld_stream4(z + base + c * HW);

    if constexpr (ACT == PCUDA_ACT_SOFTMAX) {
      float4 mx = v[0];
#pragma unroll
      for (int c = 1; c < C; ++c) {
        mx.x = fmaxf(mx.x, v[c].x); mx.y = fmaxf(mx.y, v[c].y);
        mx.z = fmaxf(mx.z, v[c].z); mx.w = fmaxf(mx.w, v[c].w);
      }
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c].x = pc_exp_sub<FAST>(v[c].x, mx.x); v[c].y = pc_exp_sub<FAST>(v[c].y, mx.y);
        v[c].z = pc_exp_sub<FAST>(v[c].z, mx.z); v[c].w = pc_exp_sub<FAST>(v[c].w, mx.w);
        s.x += v[c].x; s.y += v[c].y; s.z += v[c].z; s.w += v[c].w;
      }
      // ATen divides exp by the sum; a correctly rounded reciprocal + multiply is within 1 ulp.
      const float4 inv = make_float4(pc_rcp<FAST>(s.x), pc_rcp<FAST>(s.y), pc_rcp<FAST>(s.z), pc_rcp<FAST>(s.w));
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c].x *= inv.x; v[c].y *= inv.y; v[c].z *= inv.z; v[c].w *= inv.w;
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c].x = pc_sigmoid<FAST>(v[c].x); v[c].y = pc_sigmoid<FAST>(v[c].y);
        v[c].z = pc_sigmoid<FAST>(v[c].z); v[c].w = pc_sigmoid<FAST>(v[c].w);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if constexpr (WRITE_P) st_stream4(p_out + base + c * HW, v[c]);
      float4 e;
      // reference order: ((-1.0 * p) * log(p + smooth)) / ln C
      e.x = (-v[c].x) * pc_log<FAST>(v[c].x + smooth) * inv_norm;
      e.y = (-v[c].y) * pc_log<FAST>(v[c].y + smooth) * inv_norm;
      e.z = (-v[c].z) * pc_log<FAST>(v[c].z + smooth) * inv_norm;
      e.w = (-v[c].w) * pc_log<FAST>(v[c].w + smooth) * inv_norm;
      st_stream4(m + base + c * HW, e);
      if constexpr (WITH_MEAN) local += (e.x + e.y) + (e.z + e.w);
    }
  }
  if constexpr (WITH_MEAN) block_mean_commit(local, ws, mean_out, inv_count);
}

// ---- forward, generic: runtime C, one pixel per thread, no alignment requirement ---------------
template <int ACT, bool FAST>
__global__ void __launch_bounds__(kThreads)
entropy_fwd_generic(const float* __restrict__ z, float* __restrict__ m, float* __restrict__ p_out,
                    float* __restrict__ mean_out, MeanWs* __restrict__ ws, int64_t n_pix, int C,
                    int64_t HW, float inv_norm, float smooth, double inv_count) {
  pdl_entry();
  float local = 0.0f;
  for (int64_t q = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; q < n_pix;
       q += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int64_t b = q / HW;
    const int64_t base = b * C * HW + (q - b * HW);
    float mx = -INFINITY, inv = 1.0f;
    if (ACT == PCUDA_ACT_SOFTMAX) {
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[base + c * HW]);
      float s = 0.0f;
      for (int c = 0; c < C; ++c) s += pc_exp_sub<FAST>(z[base + c * HW], mx);
      inv = pc_rcp<FAST>(s);
    }
    for (int c = 0; c < C; ++c) {
      const float zz = z[base + c * HW];
      const float p = (ACT == PCUDA_ACT_SOFTMAX) ? pc_exp_sub<FAST>(zz, mx) * inv : pc_sigmoid<FAST>(zz);
      if (p_out) p_out[base + c * HW] = p;
      const float e = (-p) * pc_log<FAST>(p + smooth) * inv_norm;
      m[base + c * HW] = e;
      local += e;
    }
  }
  if (mean_out) block_mean_commit(local, ws, mean_out, inv_count);
}

// ---- backward ---------------------------------------------------------------------------------
// t_c = -k * g_c * (log(p_c+s) + p_c/(p_c+s)) [+ grad_p_c],   g_c = grad_m_c + grad_mean/(B*HW)
// softmax: dz_c = p_c * (t_c - sum_j t_j p_j)      sigmoid: dz_c = t_c * p_c * (1 - p_c)
template <bool FAST>
__device__ __forceinline__ float entropy_t(float p, float g, float inv_norm, float smooth) {
  const float ps = p + smooth;
  return -inv_norm * g * (pc_log<FAST>(ps) + p * pc_rcp<FAST>(ps));
}

template <int ACT, int C, bool HAS_GM, bool HAS_GP, bool FAST>
__global__ void __launch_bounds__(kThreads)
entropy_bwd_vec4(const float* __restrict__ z, const float* __restrict__ grad_m,
                 const float* __restrict__ grad_p, const float* __restrict__ grad_mean,
                 float* __restrict__ grad_z, int64_t n_quads, int64_t quads_per_img, int64_t HW,
                 float inv_norm, float smooth, float inv_count) {
  pdl_entry();
  const float gs = grad_mean ? __ldg(grad_mean) * inv_count : 0.0f;
  // (image, quad-in-image) advance incrementally: no 64-bit division inside the streaming loop
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  int64_t q = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x;
  int64_t b = q / quads_per_img;
  int64_t r = q - b * quads_per_img;
  for (; q < n_quads; q += stride, r += stride) {
    while (r >= quads_per_img) { r -= quads_per_img; ++b; }
    const int64_t base = b * C * HW + r * 4;
    float4 v[C];
#pragma unroll
    for (int c = 0; c < C; ++c) v[c] = ld_stream4(z + base + c * HW);
    if constexpr (ACT == PCUDA_ACT_SOFTMAX) {
      float4 mx = v[0];
#pragma unroll
      for (int c = 1; c < C; ++c) {
        mx.x = fmaxf(mx.x, v[c].x); mx.y = fmaxf(mx.y, v[c].y);
        mx.z = fmaxf(mx.z, v[c].z); mx.w = fmaxf(mx.w, v[c].w);
      }
      float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c].x = pc_exp_sub<FAST>(v[c].x, mx.x); v[c].y = pc_exp_sub<FAST>(v[c].y, mx.y);
        v[c].z = pc_exp_sub<FAST>(v[c].z, mx.z); v[c].w = pc_exp_sub<FAST>(v[c].w, mx.w);
        s.x += v[c].x; s.y += v[c].y; s.z += v[c].z; s.w += v[c].w;
      }
      const float4 inv = make_float4(pc_rcp<FAST>(s.x), pc_rcp<FAST>(s.y), pc_rcp<FAST>(s.z), pc_rcp<FAST>(s.w));
      float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 t[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        v[c].x *= inv.x; v[c].y *= inv.y; v[c].z *= inv.z; v[c].w *= inv.w;
        float4 g = make_float4(gs, gs, gs, gs);
        if constexpr (HAS_GM) {
          const float4 gm = ld_stream4(grad_m + base + c * HW);
          g.x += gm.x; g.y += gm.y; g.z += gm.z; g.w += gm.w;
        }
        t[c].x = entropy_t<FAST>(v[c].x, g.x, inv_norm, smooth);
        t[c].y = entropy_t<FAST>(v[c].y, g.y, inv_norm, smooth);
        t[c].z = entropy_t<FAST>(v[c].z, g.z, inv_norm, smooth);
        t[c].w = entropy_t<FAST>(v[c].w, g.w, inv_norm, smooth);
        if constexpr (HAS_GP) {
          const float4 gp = ld_stream4(grad_p + base + c * HW);
          t[c].x += gp.x; t[c].y += gp.y; t[c].z += gp.z; t[c].w += gp.w;
        }
        dot.x = fmaf(t[c].x, v[c].x, dot.x); dot.y = fmaf(t[c].y, v[c].y, dot.y);
        dot.z = fmaf(t[c].z, v[c].z, dot.z); dot.w = fmaf(t[c].w, v[c].w, dot.w);
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float4 o;
        o.x = v[c].x * (t[c].x - dot.x); o.y = v[c].y * (t[c].y - dot.y);
        o.z = v[c].z * (t[c].z - dot.z); o.w = v[c].w * (t[c].w - dot.w);
        st_stream4(grad_z + base + c * HW, o);
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float4 g = make_float4(gs, gs, gs, gs);
        if constexpr (HAS_GM) {
          const float4 gm = ld_stream4(grad_m + base + c * HW);
          g.x += gm.x; g.y += gm.y; g.z += gm.z; g.w += gm.w;
        }
        float4 gp = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (HAS_GP) gp = ld_stream4(grad_p + base + c * HW);
        float4 o;
        float p;
        p = pc_sigmoid<FAST>(v[c].x); o.x = (entropy_t<FAST>(p, g.x, inv_norm, smooth) + gp.x) * p * (1.0f - p);
        p = pc_sigmoid<FAST>(v[c].y); o.y = (entropy_t<FAST>(p, g.y, inv_norm, smooth) + gp.y) * p * (1.0f - p);
        p = pc_sigmoid<FAST>(v[c].z); o.z = (entropy_t<FAST>(p, g.z, inv_norm, smooth) + gp.z) * p * (1.0f - p);
        p = pc_sigmoid<FAST>(v[c].w); o.w = (entropy_t<FAST>(p, g.w, inv_norm, smooth) + gp.w) * p * (1.0f - p);
        st_stream4(grad_z + base + c * HW, o);
      }
    }
  }
}

template <int ACT, bool FAST>
__global__ void __launch_bounds__(kThreads)
entropy_bwd_generic(const float* __restrict__ z, const float* __restrict__ grad_m,
                    const float* __restrict__ grad_p, const float* __restrict__ grad_mean,
                    float* __restrict__ grad_z, int64_t n_pix, int C, int64_t HW, float inv_norm,
                    float smooth, float inv_count) {
  pdl_entry();
  const float gs = grad_mean ? __ldg(grad_mean) * inv_count : 0.0f;
  for (int64_t q = blockIdx.x * static_cast<int64_t>(kThreads) + threadIdx.x; q < n_pix;
       q += static_cast<int64_t>(gridDim.x) * kThreads) {
    const int64_t b = q / HW;
    const int64_t base = b * C * HW + (q - b * HW);
    float mx = -INFINITY, inv = 1.0f, dot = 0.0f;
    if (ACT == PCUDA_ACT_SOFTMAX) {
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, z[base + c * HW]);
      float s = 0.0f;
      for (int c = 0; c < C; ++c) s += pc_exp_sub<FAST>(z[base + c * HW], mx);
      inv = pc_rcp<FAST>(s);
      for (int c = 0; c < C; ++c) {
        const float p = pc_exp_sub<FAST>(z[base + c * HW], mx) * inv;
        const float g = gs + (grad_m ? grad_m[base + c * HW] : 0.0f);
        float t = entropy_t<FAST>(p, g, inv_norm, smooth);
        if (grad_p) t += grad_p[base + c * HW];
        dot = fmaf(t, p, dot);
      }
    }
    for (int c = 0; c < C; ++c) {
      const float zz = z[base + c * HW];
      const float p = (ACT == PCUDA_ACT_SOFTMAX) ? pc_exp_sub<FAST>(zz, mx) * inv : pc_sigmoid<FAST>(zz);
      const float g = gs + (grad_m ? grad_m[base + c * HW] : 0.0f);
      float t = entropy_t<FAST>(p, g, inv_norm, smooth);
      if (grad_p) t += grad_p[base + c * HW];
      grad_z[base + c * HW] = (ACT == PCUDA_ACT_SOFTMAX) ? p * (t - dot) : t * p * (1.0f - p);
    }
  }
}

// ---- launch plumbing ---------------------------------------------------------------------------
template <typename K>
int grid_for(K kernel, int64_t work_items) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0) != cudaSuccess ||
      per_sm <= 0)
    per_sm = 4;
  const int64_t need = (work_items + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(sm_count()) * per_sm;
  int64_t g = need < cap ? need : cap;
  return static_cast<int>(g < 1 ? 1 : g);
}

struct FwdArgs {
  const float* z; float* m; float* p; float* mean_out; MeanWs* ws;
  int B, C; int64_t HW; float inv_norm, smooth; cudaStream_t st;
};

template <int ACT, int C, bool WP, bool WM, bool FAST>
int launch_fwd_vec4(const FwdArgs& a, int64_t n_quads, int64_t qpi) {
  auto k = entropy_fwd_vec4<ACT, C, WP, WM, FAST>;
  const int grid = grid_for(k, n_quads);
  PCUDA_LAUNCH(k, grid, kThreads, 0, a.st, a.z, a.m, a.p, a.mean_out, a.ws, n_quads, qpi, a.HW, a.inv_norm,
                                 a.smooth, 1.0 / (static_cast<double>(a.B) * a.HW));
  count_launch(); return check_launch("entropy_fwd_vec4");
}

template <int ACT, int C, bool FAST>
int dispatch_fwd_flags(const FwdArgs& a, int64_t n_quads, int64_t qpi) {
  const bool wp = a.p != nullptr, wm = a.mean_out != nullptr;
  if (wp && wm) return launch_fwd_vec4<ACT, C, true, true, FAST>(a, n_quads, qpi);
  if (wp) return launch_fwd_vec4<ACT, C, true, false, FAST>(a, n_quads, qpi);
  if (wm) return launch_fwd_vec4<ACT, C, false, true, FAST>(a, n_quads, qpi);
  return launch_fwd_vec4<ACT, C, false, false, FAST>(a, n_quads, qpi);
}

template <int ACT, bool FAST>
int dispatch_fwd_c(const FwdArgs& a, int Ceff, int64_t n_quads, int64_t qpi) {
  switch (Ceff) {
    case 1: return dispatch_fwd_flags<ACT, 1, FAST>(a, n_quads, qpi);
    case 2: return dispatch_fwd_flags<ACT, 2, FAST>(a, n_quads, qpi);
    case 3: return dispatch_fwd_flags<ACT, 3, FAST>(a, n_quads, qpi);
    case 4: return dispatch_fwd_flags<ACT, 4, FAST>(a, n_quads, qpi);
    case 5: return dispatch_fwd_flags<ACT, 5, FAST>(a, n_quads, qpi);
    case 6: return dispatch_fwd_flags<ACT, 6, FAST>(a, n_quads, qpi);
    case 7: return dispatch_fwd_flags<ACT, 7, FAST>(a, n_quads, qpi);
    case 8: return dispatch_fwd_flags<ACT, 8, FAST>(a, n_quads, qpi);
    default: return -100;
  }
}

struct BwdArgs {
  const float* z; const float* gm; const float* gp; const float* gmean; float* gz;
  int B, C; int64_t HW; float inv_norm, smooth; cudaStream_t st;
};

template <int ACT, int C, bool GM, bool GP, bool FAST>
int launch_bwd_vec4(const BwdArgs& a, int64_t n_quads, int64_t qpi) {
  auto k = entropy_bwd_vec4<ACT, C, GM, GP, FAST>;
  const int grid = grid_for(k, n_quads);
  PCUDA_LAUNCH(k, grid, kThreads, 0, a.st, a.z, a.gm, a.gp, a.gmean, a.gz, n_quads, qpi, a.HW, a.inv_norm,
                                 a.smooth, static_cast<float>(1.0 / (static_cast<double>(a.B) * a.HW)));
  count_launch(); return check_launch("entropy_bwd_vec4");
}

template <int ACT, int C, bool FAST>
int dispatch_bwd_flags(const BwdArgs& a, int64_t n_quads, int64_t qpi) {
  const bool gm = a.gm != nullptr, gp = a.gp != nullptr;
  if (gm && gp) return launch_bwd_vec4<ACT, C, true, true, FAST>(a, n_quads, qpi);
  if (gm) return launch_bwd_vec4<ACT, C, true, false, FAST>(a, n_quads, qpi);
  if (gp) return launch_bwd_vec4<ACT, C, false, true, FAST>(a, n_quads, qpi);
  return launch_bwd_vec4<ACT, C, false, false, FAST>(a, n_quads, qpi);
}

template <int ACT, bool FAST>
int dispatch_bwd_c(const BwdArgs& a, int Ceff, int64_t n_quads, int64_t qpi) {
  switch (Ceff) {
    case 1: return dispatch_bwd_flags<ACT, 1, FAST>(a, n_quads, qpi);
    case 2: return dispatch_bwd_flags<ACT, 2, FAST>(a, n_quads, qpi);
    case 3: return dispatch_bwd_flags<ACT, 3, FAST>(a, n_quads, qpi);
    case 4: return dispatch_bwd_flags<ACT, 4, FAST>(a, n_quads, qpi);
    case 5: return dispatch_bwd_flags<ACT, 5, FAST>(a, n_quads, qpi);
    case 6: return dispatch_bwd_flags<ACT, 6, FAST>(a, n_quads, qpi);
    case 7: return dispatch_bwd_flags<ACT, 7, FAST>(a, n_quads, qpi);
    case 8: return dispatch_bwd_flags<ACT, 8, FAST>(a, n_quads, qpi);
    default: return -100;
  }
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" int pcuda_entropy_fwd(const float* z, float* m, float* p, float* mean_out, void* ws,
                                 int B, int C, int64_t HW, int activation, float inv_norm,
                                 float smooth, pcuda_stream_t stream) {
  PCUDA_REQUIRE(B >= 0 && C >= 1 && HW >= 0, PCUDA_E_SHAPE, "entropy_fwd: bad shape B=%d C=%d HW=%lld", B, C, (long long)HW);
  PCUDA_REQUIRE(C <= 16, PCUDA_E_UNSUPPORTED, "entropy_fwd: C=%d > 16", C);
  PCUDA_REQUIRE(activation == PCUDA_ACT_SIGMOID || activation == PCUDA_ACT_SOFTMAX, PCUDA_E_UNSUPPORTED, "entropy_fwd: activation %d", activation);
  PCUDA_REQUIRE(!mean_out || ws, PCUDA_E_NULL, "entropy_fwd: mean_out needs ws");
  const int64_t n_pix = static_cast<int64_t>(B) * HW;
  if (n_pix == 0) {
    // torch.mean of an empty tensor is NaN; keep that visible rather than leaving garbage.
    if (mean_out) {
      const float nanv = __builtin_nanf("");
      cudaMemcpyAsync(mean_out, &nanv, sizeof(float), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
    }
    return 0;
  }
  PCUDA_REQUIRE(z && m, PCUDA_E_NULL, "entropy_fwd: z/m is NULL");
  FwdArgs a{z, m, p, mean_out, static_cast<MeanWs*>(ws), B, C, HW, inv_norm, smooth, static_cast<cudaStream_t>(stream)};
  const bool fast = tuning(TUNE_ENTROPY_PRECISE) == 0;
  const bool ptr_ok = aligned16(z) && aligned16(m) && (!p || aligned16(p));
  int rc = -100;
  if (activation == PCUDA_ACT_SIGMOID) {
    // element-wise: treat [B,C,HW] as one flat single-"class" image
    const int64_t n = n_pix * C;
    if (ptr_ok && (n % 4 == 0)) {
      FwdArgs f = a; f.C = 1; f.HW = n; f.B = 1;
      // mean divides by B*HW of the ORIGINAL shape
      auto run = [&](auto fastc) {
        constexpr bool F = decltype(fastc)::value;
        const bool wp = p != nullptr, wm = mean_out != nullptr;
        const double invc = 1.0 / static_cast<double>(n_pix);
        auto go = [&](auto k) {
          const int grid = grid_for(k, n / 4);
          PCUDA_LAUNCH(k, grid, kThreads, 0, f.st, z, m, p, mean_out, f.ws, n / 4, n / 4, n, inv_norm, smooth, invc);
          count_launch(); return check_launch("entropy_fwd_vec4(sigmoid)");
        };
        if (wp && wm) return go(entropy_fwd_vec4<PCUDA_ACT_SIGMOID, 1, true, true, F>);
        if (wp) return go(entropy_fwd_vec4<PCUDA_ACT_SIGMOID, 1, true, false, F>);
        if (wm) return go(entropy_fwd_vec4<PCUDA_ACT_SIGMOID, 1, false, true, F>);
        return go(entropy_fwd_vec4<PCUDA_ACT_SIGMOID, 1, false, false, F>);
      };
      return fast ? run(std::true_type{}) : run(std::false_type{});
    }
  } else if (ptr_ok && (HW % 4 == 0) && C <= 8) {
    rc = fast ? dispatch_fwd_c<PCUDA_ACT_SOFTMAX, true>(a, C, n_pix / 4, HW / 4)
              : dispatch_fwd_c<PCUDA_ACT_SOFTMAX, false>(a, C, n_pix / 4, HW / 4);
    if (rc != -100) return rc;
  }
  // generic path: any C <= 16, any alignment
  const double invc = 1.0 / static_cast<double>(n_pix);
  auto go = [&](auto k) {
    const int grid = grid_for(k, n_pix);
    PCUDA_LAUNCH(k, grid, kThreads, 0, a.st, z, m, p, mean_out, a.ws, n_pix, C, HW, inv_norm, smooth, invc);
    count_launch(); return check_launch("entropy_fwd_generic");
  };
  if (activation == PCUDA_ACT_SOFTMAX)
    return fast ? go(entropy_fwd_generic<PCUDA_ACT_SOFTMAX, true>) : go(entropy_fwd_generic<PCUDA_ACT_SOFTMAX, false>);
  return fast ? go(entropy_fwd_generic<PCUDA_ACT_SIGMOID, true>) : go(entropy_fwd_generic<PCUDA_ACT_SIGMOID, false>);
}

extern "C" int pcuda_entropy_bwd(const float* z, const float* grad_m, const float* grad_p,
                                 const float* grad_mean, float* grad_z, int B, int C, int64_t HW,
                                 int activation, float inv_norm, float smooth,
                                 pcuda_stream_t stream) {
  PCUDA_REQUIRE(B >= 0 && C >= 1 && HW >= 0, PCUDA_E_SHAPE, "entropy_bwd: bad shape B=%d C=%d HW=%lld", B, C, (long long)HW);
  PCUDA_REQUIRE(C <= 16, PCUDA_E_UNSUPPORTED, "entropy_bwd: C=%d > 16", C);
  PCUDA_REQUIRE(activation == PCUDA_ACT_SIGMOID || activation == PCUDA_ACT_SOFTMAX, PCUDA_E_UNSUPPORTED, "entropy_bwd: activation %d", activation);
  const int64_t n_pix = static_cast<int64_t>(B) * HW;
  if (n_pix == 0) return 0;
  PCUDA_REQUIRE(z && grad_z, PCUDA_E_NULL, "entropy_bwd: z/grad_z is NULL");
  BwdArgs a{z, grad_m, grad_p, grad_mean, grad_z, B, C, HW, inv_norm, smooth, static_cast<cudaStream_t>(stream)};
  const bool fast = tuning(TUNE_ENTROPY_PRECISE) == 0;
  const bool ptr_ok = aligned16(z) && aligned16(grad_z) && (!grad_m || aligned16(grad_m)) && (!grad_p || aligned16(grad_p));
  const float invc = static_cast<float>(1.0 / static_cast<double>(n_pix));
  if (activation == PCUDA_ACT_SIGMOID) {
    const int64_t n = n_pix * C;
    if (ptr_ok && (n % 4 == 0)) {
      auto run = [&](auto fastc) {
        constexpr bool F = decltype(fastc)::value;
        auto go = [&](auto k) {
          const int grid = grid_for(k, n / 4);
          PCUDA_LAUNCH(k, grid, kThreads, 0, a.st, z, grad_m, grad_p, grad_mean, grad_z, n / 4, n / 4, n, inv_norm, smooth, invc);
          count_launch(); return check_launch("entropy_bwd_vec4(sigmoid)");
        };
        const bool gm = grad_m != nullptr, gp = grad_p != nullptr;
        if (gm && gp) return go(entropy_bwd_vec4<PCUDA_ACT_SIGMOID, 1, true, true, F>);
        if (gm) return go(entropy_bwd_vec4<PCUDA_ACT_SIGMOID, 1, true, false, F>);
        if (gp) return go(entropy_bwd_vec4<PCUDA_ACT_SIGMOID, 1, false, true, F>);
        return go(entropy_bwd_vec4<PCUDA_ACT_SIGMOID, 1, false, false, F>);
      };
      return fast ? run(std::true_type{}) : run(std::false_type{});
    }
  } else if (ptr_ok && (HW % 4 == 0) && C <= 8) {
    int rc = fast ? dispatch_bwd_c<PCUDA_ACT_SOFTMAX, true>(a, C, n_pix / 4, HW / 4)
                  : dispatch_bwd_c<PCUDA_ACT_SOFTMAX, false>(a, C, n_pix / 4, HW / 4);
    if (rc != -100) return rc;
  }
  auto go = [&](auto k) {
    const int grid = grid_for(k, n_pix);
    PCUDA_LAUNCH(k, grid, kThreads, 0, a.st, z, grad_m, grad_p, grad_mean, grad_z, n_pix, C, HW, inv_norm, smooth, invc);
    count_launch(); return check_launch("entropy_bwd_generic");
  };
  if (activation == PCUDA_ACT_SOFTMAX)
    return fast ? go(entropy_bwd_generic<PCUDA_ACT_SOFTMAX, true>) : go(entropy_bwd_generic<PCUDA_ACT_SOFTMAX, false>);
  return fast ? go(entropy_bwd_generic<PCUDA_ACT_SIGMOID, true>) : go(entropy_bwd_generic<PCUDA_ACT_SIGMOID, false>);
}
