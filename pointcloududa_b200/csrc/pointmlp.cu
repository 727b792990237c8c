// PointNet shared MLP (Conv1d k=1 + train-mode BatchNorm1d + ReLU stack) with fused global max-pool.
//
// Replaces, for the point-cloud discriminator D4 of the reference,
//   networks/PointNetCls.py:41-44   (STN3d trunk)      :84-87 (STNkd trunk)
//   networks/PointNetCls.py:143-162 (PointNetfeat trunk + torch.max(x, 2))
// and their autograd backward.  This file holds the FP32 CUDA-core implementation (parity anchor
// and the K=3 / K=8 layers); pointmlp_tc.cu holds the tcgen05 tensor-core GEMM for wide layers.
//
// Data layout: activations are POINT-MAJOR  y[m, c],  m = b*N + n  (channels contiguous), so a
// point is a K-contiguous GEMM row and weights [Cout, Cin] are K-contiguous too.
//
// What is never materialised:
//   * the pooled (last) layer's [B, 1024, N] activation: its GEMM epilogue reduces straight to
//     per-channel BN statistics and a per-(cloud, channel) arg-max.  BN is a monotone per-channel
//     affine map, so max_n relu?(bn(y)) = relu?(bn(max_n y)) for gamma >= 0 (min_n for gamma < 0).
//   * any post-BN / post-ReLU activation: consumers re-apply (y - mean)*invstd*gamma + beta and
//     ReLU while loading their operand tile.
//   * the dense [B*N, 1024] gradient of the pooled layer: d(out)/d(y) is (#clouds) non-zeros per
//     channel plus a BN correction that is rank-one + a multiple of y itself, so dgrad/wgrad
//     collapse to   da = sparse - u' - Q a,   dW = sparse - diag(kappa) W Ghat
//     with Q = W^T diag(kappa) W  [Cin x Cin]  and  Ghat the centred Gram matrix of the layer's
//     input — an 8x cut in backward FLOPs for 128 -> 1024 (SURVEY.md §9 gives the dense formulas).
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "pointmlp_common.cuh"
#include "pointmlp_tc.cuh"

namespace pcuda {
// comm.cu: sum of raw fp64 BatchNorm statistics over the ranks of a communicator (cross-rank BatchNorm mode)
int comm_world(pcuda_comm_t* c);
int comm_sum_f64(pcuda_comm_t* c, double* buf, int64_t count, cudaStream_t st);

namespace {

constexpr int TM = 64;   // tile rows
constexpr int TN = 64;   // tile cols
constexpr int TK = 16;   // contraction chunk
constexpr int LD = 68;   // padded leading dimension of the smem tiles
constexpr int kThreads = 256;

// ---- shared 64x64x16 FP32 micro-kernel ---------------------------------------------------------
struct Tiles {
  float As[TK][LD];
  float Bs[TK][LD];
};

__device__ __forceinline__ void mma_chunk(const Tiles& t, float (&acc)[4][4], int ty, int tx) {
#pragma unroll
  for (int kk = 0; kk < TK; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(&t.As[kk][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&t.Bs[kk][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// row-operand chunk: rows = 64 tile rows, contraction k0..k0+15.  f4(row, k) -> 4 consecutive k.
template <class F>
__device__ __forceinline__ void stage_rows(float (&S)[TK][LD], F f4, int k0) {
  const int r = threadIdx.x >> 2, kq = (threadIdx.x & 3) * 4;
  const float4 v = f4(r, k0 + kq);
  S[kq + 0][r] = v.x; S[kq + 1][r] = v.y; S[kq + 2][r] = v.z; S[kq + 3][r] = v.w;
}
// contraction-major chunk: f4(kk_global, col) -> 4 consecutive output columns
template <class F>
__device__ __forceinline__ void stage_cols(float (&S)[TK][LD], F f4, int k0) {
  const int kk = threadIdx.x >> 4, c4 = (threadIdx.x & 15) * 4;
  const float4 v = f4(k0 + kk, c4);
  *reinterpret_cast<float4*>(&S[kk][c4]) = v;
}


// ---- forward layer -------------------------------------------------------------------------------
// grid (B * tiles_per_sample, ceil(Cout/64)).  Tiles never straddle two clouds.
template <bool POOL>
__global__ void __launch_bounds__(kThreads)
mlp_fwd_kernel(ActSrc src, const float* __restrict__ W, const float* __restrict__ bias, int Cout,
               int N, int tiles_per_sample, float* __restrict__ y_out, double* __restrict__ stats,
               const float* __restrict__ pivot, const float* __restrict__ gamma,
               unsigned long long* __restrict__ keys) {
  pdl_entry();
  __shared__ Tiles t;
  __shared__ unsigned long long skey[TN];
  const int b = blockIdx.x / tiles_per_sample;
  const int n0 = (blockIdx.x - b * tiles_per_sample) * TM;
  const int c0 = blockIdx.y * TN;
  const int K = src.C;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (threadIdx.x < TN) skey[threadIdx.x] = 0ull;

  float acc[4][4] = {};
  auto a4 = [&](int r, int k) -> float4 {
    if (n0 + r >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
    return load_act4(src, static_cast<int64_t>(b) * N + n0 + r, k);
  };
  auto w4 = [&](int r, int k) -> float4 {  // W[c0+r, k..k+3]
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = c0 + r;
    if (c < Cout) {
      if ((K & 3) == 0 && k + 3 < K) v = *reinterpret_cast<const float4*>(W + static_cast<int64_t>(c) * K + k);
      else { float* pv = &v.x; for (int i = 0; i < 4 && k + i < K; ++i) pv[i] = W[static_cast<int64_t>(c) * K + k + i]; }
    }
    return v;
  };
  for (int k0 = 0; k0 < K; k0 += TK) {
    __syncthreads();
    stage_rows(t.As, a4, k0);
    stage_rows(t.Bs, w4, k0);
    __syncthreads();
    mma_chunk(t, acc, ty, tx);
  }

  // epilogue: bias, store, BN statistics, pooled arg-max
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
  unsigned long long ck[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + tx * 4 + j;
    if (c >= Cout) continue;
    const float bj = bias ? bias[c] : 0.f;
    const float g = POOL ? gamma[c] : 0.f;
    // statistics are accumulated on (y - pivot_c), pivot_c = y of point 0: E[y^2]-E[y]^2 on raw
    // values loses |mean|^2/var digits in fp32, centred sums do not (torch uses Welford)
    const float pv = pivot[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty * 4 + i;
      if (n >= N) continue;
      const float v = acc[i][j] + bj;
      acc[i][j] = v;
      const float vc = v - pv;
      cs[j] += vc;
      cq[j] = fmaf(vc, vc, cq[j]);
      if (POOL) {
        // track max of sign(gamma)*y; gamma == 0 -> every n ties and the first index wins
        const float vv = g > 0.f ? v : (g < 0.f ? -v : 0.f);
        const unsigned long long key = pool_key(vv, n);
        ck[j] = key > ck[j] ? key : ck[j];
      }
    }
  }
  if (y_out != nullptr) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + ty * 4 + i;
      if (n >= N) continue;
      float* row = y_out + (static_cast<int64_t>(b) * N + n) * Cout;
      const int c = c0 + tx * 4;
      if ((Cout & 3) == 0 && c + 3 < Cout) {
        *reinterpret_cast<float4*>(row + c) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      } else {
        for (int j = 0; j < 4 && c + j < Cout; ++j) row[c + j] = acc[i][j];
      }
    }
  }
  // column sums of the 16 row groups are combined in a FIXED order through the (now free) operand
  // tiles: float atomics would make the statistics, and through them every bf16 rounding downstream,
  // depend on the scheduling order
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    t.As[ty][tx * 4 + j] = cs[j];
    t.Bs[ty][tx * 4 + j] = cq[j];
    if (POOL) atomicMax(&skey[tx * 4 + j], ck[j]);
  }
  __syncthreads();
  if (threadIdx.x < TN && c0 + threadIdx.x < Cout) {
    const int c = c0 + threadIdx.x;
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int g = 0; g < TK; ++g) { s += static_cast<double>(t.As[g][threadIdx.x]); q += static_cast<double>(t.Bs[g][threadIdx.x]); }
    atomicAdd(&stats[c], s);
    atomicAdd(&stats[Cout + c], q);
    if (POOL) atomicMax(&keys[static_cast<int64_t>(b) * Cout + c], skey[threadIdx.x]);
  }
}

// ---- narrow layers (cin <= 8: the 3->64 / 3->8 / 8->64 first layers) ------------------------------------
// With a contraction of 3 the layer is a pure HBM stream (write y forward, read dy backward), so the
// 64x64x16 tile kernels above waste a factor 5 in staging.  Here a thread owns 4 consecutive channels of
// one point: x[m, 0..K) and its 4 x K weights sit in registers, y / dy move as coalesced float4.
constexpr int kNarrowK = 8;   // widest narrow layer; the kernels are instantiated for KN = 4 (cin <= 4: the 3 -> 64 / 3 -> 8
                              // layers, i.e. all of them in the default network) and KN = 8
constexpr int kNarrowThreads = 256;

// all K <= 8 input channels of point m (one 64-bit division instead of one per channel)
template <int KN>
__device__ __forceinline__ void load_point(const ActSrc& s, int64_t m, float (&xv)[KN]) {
  if (s.y != nullptr) {
#pragma unroll
    for (int k = 0; k < KN; ++k)
      xv[k] = k < s.C ? bn_act(s.y[m * s.C + k], s.mean[k], s.invstd[k], s.gamma[k], s.beta[k], s.relu) : 0.f;
  } else {
    const int64_t b = m / s.N, n = m - b * s.N;
    if (s.trans != nullptr) {          // C <= 4 (validated by the host): transformed on load
      float t[4];
      load_raw_point(s, b, n, t);
#pragma unroll
      for (int k = 0; k < KN; ++k) xv[k] = k < 4 ? t[k < 4 ? k : 0] : 0.f;
      return;
    }
    const float* px = s.x + b * s.sxb + n * s.sxn;
#pragma unroll
    for (int k = 0; k < KN; ++k) xv[k] = k < s.C ? __ldg(px + k * s.sxc) : 0.f;
  }
}

// forward: y = W a + b, centred BN statistics; CTA 0 also publishes the pivot (y of point 0)
template <int KN>
__global__ void __launch_bounds__(kNarrowThreads)
mlp_fwd_narrow_kernel(ActSrc src, const float* __restrict__ W, const float* __restrict__ bias, int Cout, int64_t M,
                      float* __restrict__ y_out, double* __restrict__ stats, float* __restrict__ pivot) {
  __shared__ float red[kNarrowThreads / 16][2][64];   // up to 16 channel groups... sized below by cgroups <= 16... see launch
  const int K = src.C;
  const int cgroups = Cout >> 2;                       // <= 64 (Cout <= 256): thread = (point slot, channel group)
  const int tpp = cgroups;                             // threads per point
  const int pslots = kNarrowThreads / tpp;             // points per CTA pass
  const int cg = threadIdx.x % tpp, ps = threadIdx.x / tpp;
  const int c0 = cg * 4;
  float w[4][KN], bj[4], pv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bj[j] = bias ? bias[c0 + j] : 0.f;
#pragma unroll
    for (int k = 0; k < KN; ++k) w[j][k] = k < K ? W[static_cast<int64_t>(c0 + j) * K + k] : 0.f;
  }
  pdl_entry();                                         // weights and bias (parameters) were requested before the wait
  {  // pivot = pre-BN output at global point 0 (same definition as bn_finalize_pivot_kernel)
    float x0[KN];
    load_point<KN>(src, 0, x0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float d = 0.f;
#pragma unroll
      for (int k = 0; k < KN; ++k) d = fmaf(w[j][k], x0[k], d);
      pv[j] = d + bj[j];
    }
    if (blockIdx.x == 0 && ps == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) pivot[c0 + j] = pv[j];
    }
  }
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
  if (ps < pslots) {
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * pslots + ps; m < M; m += static_cast<int64_t>(gridDim.x) * pslots) {
      float xv[KN];
      load_point<KN>(src, m, xv);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < KN; ++k) d = fmaf(w[j][k], xv[k], d);
        v[j] = d + bj[j];
        const float vc = v[j] - pv[j];
        cs[j] += vc;
        cq[j] = fmaf(vc, vc, cq[j]);
      }
      *reinterpret_cast<float4*>(y_out + m * Cout + c0) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  // fixed-order reduction over the point slots of the CTA, then one fp64 atomic per channel and CTA
  float* rs = &red[0][0][0];            // [pslots][2][Cout] floats, pslots * Cout = 256 -> 2 KB
  if (ps < pslots) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { rs[(ps * 2 + 0) * Cout + c0 + j] = cs[j]; rs[(ps * 2 + 1) * Cout + c0 + j] = cq[j]; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cout; c += kNarrowThreads) {
    double sd = 0.0, qd = 0.0;
    for (int g = 0; g < pslots; ++g) { sd += static_cast<double>(rs[(g * 2 + 0) * Cout + c]); qd += static_cast<double>(rs[(g * 2 + 1) * Cout + c]); }
    atomicAdd(&stats[c], sd);
    atomicAdd(&stats[Cout + c], qd);
  }
}

// weight gradient: partial[blockIdx.x][c][k] = sum over this CTA's points of dy[m,c] * a_prev[m,k]
template <int KN>
__global__ void __launch_bounds__(kNarrowThreads)
wgrad_narrow_kernel(DySrc dys, ActSrc aprev, int64_t M, float* __restrict__ partial) {
  pdl_entry();
  extern __shared__ float red_w[];     // [pslots][C][K]
  const int C = dys.C, K = aprev.C;
  const int tpp = C >> 2, pslots = kNarrowThreads / tpp;
  const int cg = threadIdx.x % tpp, ps = threadIdx.x / tpp;
  const int c0 = cg * 4;
  float acc[4][KN];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KN; ++k) acc[j][k] = 0.f;
  if (ps < pslots) {
    const DyConst4 dk = dy_const4(dys, c0);
    for (int64_t m = static_cast<int64_t>(blockIdx.x) * pslots + ps; m < M; m += static_cast<int64_t>(gridDim.x) * pslots) {
      const float4 d = load_dy4c(dys, dk, m, c0);
      const float dv[4] = {d.x, d.y, d.z, d.w};
      float xv[KN];
      load_point<KN>(aprev, m, xv);
#pragma unroll
      for (int k = 0; k < KN; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j][k] = fmaf(dv[j], xv[k], acc[j][k]);
    }
  }
  if (ps < pslots) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < KN; ++k)
        if (k < K) red_w[(static_cast<int64_t>(ps) * C + c0 + j) * K + k] = acc[j][k];
  }
  __syncthreads();
  float* outp = partial + static_cast<int64_t>(blockIdx.x) * C * K;
  for (int i = threadIdx.x; i < C * K; i += kNarrowThreads) {
    double t = 0.0;
    for (int g = 0; g < pslots; ++g) t += static_cast<double>(red_w[static_cast<int64_t>(g) * C * K + i]);
    outp[i] = static_cast<float>(t);
  }
}

// gradient w.r.t. the network input: grad_x[b, k, n] = sum_c dy[m, c] W[c, k]   (K <= 8, C <= 128: the
// tpp = C/4 <= 32 threads of a point sit in one warp and combine with shuffles)
template <int KN>
__global__ void __launch_bounds__(kNarrowThreads)
dgrad_input_narrow_kernel(DySrc dys, const float* __restrict__ W, int K, int N, int64_t M, float* __restrict__ grad_x) {
  const int C = dys.C;
  const int tpp = C >> 2, pslots = kNarrowThreads / tpp;
  const int cg = threadIdx.x % tpp, ps = threadIdx.x / tpp;
  const int c0 = cg * 4;
  float w[4][KN];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < KN; ++k) w[j][k] = k < K ? W[static_cast<int64_t>(c0 + j) * K + k] : 0.f;
  pdl_entry();                                           // the weights (parameters) were requested before the wait
  const int64_t stride = static_cast<int64_t>(gridDim.x) * pslots;
  const int64_t iters = (M + stride - 1) / stride;       // uniform trip count: the shuffles below need full warps
  const DyConst4 dk = dy_const4(dys, c0);
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t m = it * stride + static_cast<int64_t>(blockIdx.x) * pslots + ps;
    const bool ok = m < M && ps < pslots;
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) d = load_dy4c(dys, dk, m, c0);
    float g[KN];
#pragma unroll
    for (int k = 0; k < KN; ++k) g[k] = d.x * w[0][k] + d.y * w[1][k] + d.z * w[2][k] + d.w * w[3][k];
    for (int o = tpp >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int k = 0; k < KN; ++k) g[k] += __shfl_xor_sync(0xffffffffu, g[k], o);
    }
    if (ok && cg == 0) {
      const int64_t b = m / N, n = m - b * N;
#pragma unroll
      for (int k = 0; k < KN; ++k)
        if (k < K) grad_x[(b * K + k) * N + n] = g[k];
    }
  }
}

// Backward of the input transform fused into the first layer's operand load (a = T^T x per point): with
// g = dL/da [B, C, N] (written by the first layer's dgrad into the workspace),
//   grad_x[b, j, n] = sum_k T[b][j][k] g[b, k, n]          (torch.bmm backward w.r.t. the cloud)
//   grad_T[b][j][k] = sum_n x[b, j, n] g[b, k, n]          (w.r.t. the transform: into STN3d's head)
// One CTA per cloud, fixed-order reduction: deterministic.  C <= 4.
__global__ void __launch_bounds__(256)
input_transform_bwd_kernel(const float* __restrict__ x, int64_t sxb, int64_t sxc, int64_t sxn, const float* __restrict__ T,
                           const float* __restrict__ g, int C, int N, float* __restrict__ grad_x, float* __restrict__ grad_T) {
  pdl_entry();
  __shared__ double red[8][16];
  const int b = blockIdx.x;
  const float* Tb = T + static_cast<int64_t>(b) * C * C;
  float t[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) t[j][k] = (j < C && k < C) ? Tb[j * C + k] : 0.f;
  float acc[4][4] = {};
  for (int n = threadIdx.x; n < N; n += 256) {
    float gv[4], xv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) gv[k] = k < C ? g[(static_cast<int64_t>(b) * C + k) * N + n] : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) xv[j] = j < C ? __ldg(x + b * sxb + j * sxc + n * sxn) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (grad_x != nullptr && j < C) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) v = fmaf(t[j][k], gv[k], v);
        grad_x[(static_cast<int64_t>(b) * C + j) * N + n] = v;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[j][k] = fmaf(xv[j], gv[k], acc[j][k]);
    }
  }
  if (grad_T == nullptr) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const double v = warp_sum(static_cast<double>(acc[j][k]));
      if (lane == 0) red[warp][j * 4 + k] = v;
    }
  __syncthreads();
  if (threadIdx.x < 16) {
    const int j = threadIdx.x >> 2, k = threadIdx.x & 3;
    if (j < C && k < C) {
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      grad_T[static_cast<int64_t>(b) * C * C + j * C + k] = static_cast<float>(v);
    }
  }
}

bool narrow_ok(int cin, int cout) {
  // threads per point = cout/4 must divide the CTA (a power of two <= 64)
  const int tpp = cout >> 2;
  return cin <= kNarrowK && (cout & 3) == 0 && tpp >= 1 && tpp <= 64 && (tpp & (tpp - 1)) == 0;
}
int narrow_grid(int64_t M, int cout) {
  const int pslots = kNarrowThreads / (cout >> 2);
  const int64_t want = (M + pslots - 1) / pslots;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 4;
  return static_cast<int>(std::max<int64_t>(1, std::min(want, cap)));
}

// One block.  Phase 1 (if stats != nullptr): finalise BatchNorm of layer l-1 from its centred sums.
// Phase 2 (if W_next != nullptr): pivot of layer l = its pre-BN output at global point 0, which
// needs layer l-1's finalised parameters — hence the same kernel, separated by a block barrier.
__global__ void __launch_bounds__(1024)
bn_finalize_pivot_kernel(const double* __restrict__ stats, const float* __restrict__ pivot_prev, int C,
                         double count, float eps, float momentum, int train, float* __restrict__ save_mean,
                         float* __restrict__ save_invstd, float* __restrict__ running_mean,
                         float* __restrict__ running_var, ActSrc src_next, const float* __restrict__ W_next,
                         const float* __restrict__ bias_next, int C_next, float* __restrict__ pivot_next) {
  pdl_entry();
  if (stats != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (train) {
        const double d = stats[c] / count;
        const double mean = static_cast<double>(pivot_prev[c]) + d;
        double var = stats[C + c] / count - d * d;
        var = var > 0.0 ? var : 0.0;
        save_mean[c] = static_cast<float>(mean);
        save_invstd[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
        if (running_mean) running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
        if (running_var) {
          const double unbiased = var * (count / (count - 1.0));
          running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
        }
      } else {
        save_mean[c] = running_mean[c];
        save_invstd[c] = static_cast<float>(1.0 / sqrt(static_cast<double>(running_var[c]) + static_cast<double>(eps)));
      }
    }
  }
  if (W_next == nullptr) return;
  __threadfence_block();
  __syncthreads();
  const int K = src_next.C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int c = warp; c < C_next; c += nwarps) {
    float dot = 0.f;
    for (int k = lane; k < K; k += 32) dot = fmaf(W_next[static_cast<int64_t>(c) * K + k], load_act1(src_next, 0, k), dot);
    dot = warp_sum(dot);
    if (lane == 0) pivot_next[c] = dot + (bias_next ? bias_next[c] : 0.f);
  }
}

__global__ void pool_finalize_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ mean,
                                     const float* __restrict__ invstd, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, int relu, int B, int C,
                                     float* __restrict__ out, int32_t* __restrict__ arg, BnRaw raw) {
  pdl_entry();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * C) return;
  const int c = static_cast<int>(i % C);
  const unsigned long long key = keys[i];
  const float vv = ordered_to_float(static_cast<uint32_t>(key >> 32));
  const int n = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull));
  const float g = gamma[c];
  float mu, is;
  if (raw.stats != nullptr) bn_raw_channel(raw, c, C, i < C, mu, is);      // finalise-on-read; cloud 0's threads write
  else { mu = mean[c]; is = invstd[c]; }
  const float y = g > 0.f ? vv : (g < 0.f ? -vv : mu);
  out[i] = bn_act(y, mu, is, g, beta[c], relu);
  arg[i] = g == 0.f ? 0 : n;   // gamma == 0: every point ties at beta and the first one wins
}

// dense output (pool == 0): out[b, c, n] = act(y[(b,n), c]); 32x32 smem transpose
__global__ void dense_out_kernel(ActSrc src, int B, float* __restrict__ out) {
  pdl_entry();
  __shared__ float tile[32][33];
  const int N = src.N, C = src.C;
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (n < N && c < C) ? load_act1(src, static_cast<int64_t>(b) * N + n, c) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, n = n0 + threadIdx.x;
    if (c < C && n < N) out[(static_cast<int64_t>(b) * C + c) * N + n] = tile[threadIdx.x][r];
  }
}

// ---- backward: pooled last layer ---------------------------------------------------------------
// One warp per (cloud, channel): masked upstream gradient and yhat at the selected point.  Only grad_out comes from the
// kernel before this one: the pre-BN value at the selected point (a K-long dot product behind a dependent arg -> row
// fetch) is recomputed from forward-pass state before the grid-dependency wait.
__global__ void pool_sel_kernel(ActSrc src, const float* __restrict__ W, const float* __restrict__ bias,
                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, int relu, const float* __restrict__ out,
                                const int32_t* __restrict__ arg, const float* __restrict__ grad_out, int B, int N,
                                int C, float* __restrict__ coef, float* __restrict__ gsel, float* __restrict__ gyh) {
  const int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const bool valid = w < static_cast<int64_t>(B) * C;
  const int c = valid ? static_cast<int>(w % C) : 0;
  const int64_t b = valid ? w / C : 0;
  const int K = src.C;
  float yhat = 0.f, ov = 0.f, sc = 0.f;
  if (valid) {
    const int64_t m = b * N + arg[w];
    ov = out[w];
    const float mu = mean[c], is = invstd[c], bs = bias ? bias[c] : 0.f;
    sc = gamma[c] * is;
    float dot = 0.f;
    for (int k = lane; k < K; k += 32) dot = fmaf(W[static_cast<int64_t>(c) * K + k], load_act1(src, m, k), dot);
    dot = warp_sum(dot);
    yhat = (dot + bs - mu) * is;
  }
  pdl_entry();
  if (!valid) return;
  float g = grad_out[w];
  if (relu && !(ov > 0.f)) g = 0.f;
  if (lane == 0) {
    coef[w] = sc * g;
    gsel[w] = g;
    gyh[w] = g != 0.f ? g * yhat : 0.f;
  }
}

// One thread per channel: fixed-order sums over the B clouds.
__global__ void pool_coef_kernel(const float* __restrict__ gsel, const float* __restrict__ gyh,
                                 const float* __restrict__ invstd, const float* __restrict__ gamma, int B, int C,
                                 double count, int train, float* __restrict__ alpha, float* __restrict__ kappa,
                                 float* __restrict__ grad_gamma, float* __restrict__ grad_beta,
                                 float* __restrict__ grad_bias, double* __restrict__ xsums, int phase) {
  // phase 0: everything.  Cross-rank BatchNorm: phase 1 = this rank's sums (-> its parameter gradients, and xsums[2C]
  // for the all-reduce), phase 2 = alpha / kappa from the reduced xsums and the global count.
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dbeta = 0.0, dgamma = 0.0;
  if (phase == 2) {
    dbeta = xsums[c];
    dgamma = xsums[C + c];
  } else {
    for (int b = 0; b < B; ++b) {
      dbeta += static_cast<double>(gsel[static_cast<int64_t>(b) * C + c]);
      dgamma += static_cast<double>(gyh[static_cast<int64_t>(b) * C + c]);
    }
  }
  const double sc = static_cast<double>(gamma[c]) * static_cast<double>(invstd[c]);
  if (phase != 2) {
    if (grad_gamma) grad_gamma[c] = static_cast<float>(dgamma);
    if (grad_beta) grad_beta[c] = static_cast<float>(dbeta);
    if (grad_bias) grad_bias[c] = train ? 0.f : static_cast<float>(sc * dbeta);
  }
  if (phase == 1) {
    xsums[c] = dbeta;
    xsums[C + c] = dgamma;
    return;
  }
  alpha[c] = train ? static_cast<float>(sc * dbeta / count) : 0.f;
  kappa[c] = train ? static_cast<float>(sc * static_cast<double>(invstd[c]) * dgamma / count) : 0.f;
}

// column sums of an activation: partial[s, k] over the s-th slice of points.
// grid (S, ceil(K/32)), 256 threads: lane = channel, the 8 warps interleave the rows of the slice
__global__ void __launch_bounds__(256)
act_colsum_kernel(ActSrc src, int64_t M, int64_t chunk, double* __restrict__ partial) {
  pdl_entry();
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.y * 32 + lane;
  const int64_t m0 = blockIdx.x * chunk;
  const int64_t m1 = m0 + chunk < M ? m0 + chunk : M;
  double s = 0.0;
  if (k < src.C) {
    float f = 0.f;
    if (src.y != nullptr) {
      const float sc = src.invstd[k] * src.gamma[k];
      const float sh = src.beta[k] - src.mean[k] * sc;
      int64_t m = m0 + warp;
      for (; m + 24 < m1; m += 32) {                       // 4 rows in flight per warp
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = src.y[(m + 8 * u) * src.C + k];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float t = fmaf(v[u], sc, sh);
          f += src.relu ? fmaxf(t, 0.f) : t;
        }
        s += static_cast<double>(f); f = 0.f;
      }
      for (; m < m1; m += 8) {
        const float t = fmaf(src.y[m * src.C + k], sc, sh);
        f += src.relu ? fmaxf(t, 0.f) : t;
      }
    } else {
      for (int64_t m = m0 + warp; m < m1; m += 8) f += load_act1(src, m, k);
    }
    s += static_cast<double>(f);
  }
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && k < src.C) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += part[w][lane];
    partial[static_cast<int64_t>(blockIdx.x) * src.C + k] = t;
  }
}

// OUT[s, r, c] = sum_{m in slice s} P[m, r] * R[m, c]   (contraction over points; 64x64 tile / CTA)
// MODE 0: P = dy (DySrc), R = act (wgrad).  MODE 1: P = R = act (Gram matrix).
template <int MODE>
__global__ void __launch_bounds__(kThreads)
point_contract_kernel(DySrc dys, ActSrc pact, ActSrc ract, int64_t M, int64_t chunk, int CR, int CC,
                      float* __restrict__ partial) {
  pdl_entry();
  __shared__ Tiles t;
  const int r0 = blockIdx.y * TM, c0 = blockIdx.z * TN;
  const int64_t m0 = blockIdx.x * chunk;
  const int64_t m1 = m0 + chunk < M ? m0 + chunk : M;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4] = {};
  for (int64_t mm = m0; mm < m1; mm += TK) {
    auto p4 = [&](int kk, int c4) -> float4 {
      const int64_t m = mm + kk;
      if (m >= m1 || r0 + c4 >= CR) return make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) return load_dy4(dys, m, r0 + c4);
      return load_act4(pact, m, r0 + c4);
    };
    auto q4 = [&](int kk, int c4) -> float4 {
      const int64_t m = mm + kk;
      if (m >= m1 || c0 + c4 >= CC) return make_float4(0.f, 0.f, 0.f, 0.f);
      return load_act4(ract, m, c0 + c4);
    };
    __syncthreads();
    stage_cols(t.As, p4, 0);
    stage_cols(t.Bs, q4, 0);
    __syncthreads();
    mma_chunk(t, acc, ty, tx);
  }
  float* outp = partial + static_cast<int64_t>(blockIdx.x) * CR * CC;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= CR) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c < CC) outp[static_cast<int64_t>(r) * CC + c] = acc[i][j];
    }
  }
}

// out[i] = scale * sum_s partial[s, i]  (fixed order, double accumulation)
template <typename TOut>
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int64_t n, int S, TOut* __restrict__ out) {
  pdl_entry();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < S; ++k) s += static_cast<double>(partial[static_cast<int64_t>(k) * n + i]);
  out[i] = static_cast<TOut>(s);
}

// The same for MANY partials of FEW outputs (the narrow layers' weight gradient: one partial per SM of 64 x 3 values): a
// thread walking 148 slices alone is 37 dependent round trips.  The 8 warps of a CTA split the slices (contiguous
// ranges, summed in warp order: deterministic), 32 outputs per CTA.
template <typename TOut>
__global__ void __launch_bounds__(256) reduce_partials_split_kernel(const float* __restrict__ partial, int64_t n, int S, TOut* __restrict__ out) {
  pdl_entry();
  __shared__ double part[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = blockIdx.x * 32ll + lane;
  const int per = (S + 7) / 8, s0 = warp * per, s1 = min(S, s0 + per);
  double t = 0.0;
  if (i < n) {
#pragma unroll 8
    for (int k = s0; k < s1; ++k) t += static_cast<double>(partial[static_cast<int64_t>(k) * n + i]);
  }
  part[warp][lane] = t;
  __syncthreads();
  if (warp == 0 && i < n) {
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) r += part[w][lane];
    out[i] = static_cast<TOut>(r);
  }
}

// abar[k] = (sum_s colsum_partial[s,k]) / M.  One warp per channel, lanes interleave the slices; the
// butterfly reduction has a fixed order, so the result is deterministic.
__global__ void __launch_bounds__(256)
abar_kernel(const double* __restrict__ colsum_partial, int S, int K, double count, float* __restrict__ abar,
            double* __restrict__ abar_d) {
  pdl_entry();
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= K) return;
  double a = 0.0;
  for (int s = lane; s < S; s += 32) a += colsum_partial[static_cast<int64_t>(s) * K + k];
  a = warp_sum(a);
  if (lane == 0) {
    a /= count;          // count == 1: the raw column sum (cross-rank BatchNorm: divided after the all-reduce)
    abar[k] = static_cast<float>(a);
    abar_d[k] = a;
  }
}

// Cross-rank BatchNorm: the sums of (y - pivot), (y - pivot)^2 of this rank are re-expressed around a pivot every rank
// shares (the layer's bias, or 0) before they are summed over the ranks:  with d = pivot - pivot',
//   sum (y - pivot') = S + n d,     sum (y - pivot')^2 = Q + 2 d S + n d^2.
__global__ void stats_rebase_kernel(double* __restrict__ stats, const float* __restrict__ pivot_from,
                                    const float* __restrict__ pivot_to, int C, double count) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double d = static_cast<double>(pivot_from[c]) - (pivot_to ? static_cast<double>(pivot_to[c]) : 0.0);
  const double S = stats[c], Q = stats[C + c];
  stats[c] = S + count * d;
  stats[C + c] = Q + 2.0 * d * S + count * d * d;
}

// Ghat[k,k'] = G[k,k'] - M abar_k abar_k'
__global__ void gram_center_kernel(const double* __restrict__ abar_d, const double* __restrict__ G, int K,
                                   double count, float* __restrict__ Ghat) {
  pdl_entry();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(K) * K) return;
  const int k = static_cast<int>(i / K), k2 = static_cast<int>(i % K);
  Ghat[i] = static_cast<float>(G[i] - count * abar_d[k] * abar_d[k2]);
}

// The three kernels above in one launch (single rank: abar needs no exchange): every CTA sums the column-sum partials
// into abar (CTA 0 publishes it), then its 256 elements of G = sum_s partial[s] and centres them.
__global__ void __launch_bounds__(256)
gram_finish_kernel(const float* __restrict__ gpartial, int Sg, const double* __restrict__ colsum_partial, int Sc, int K,
                   double count, float* __restrict__ abar, double* __restrict__ abar_d, float* __restrict__ Ghat) {
  pdl_entry();
  extern __shared__ double ab[];                   // [K]
  const int64_t kk = static_cast<int64_t>(K) * K;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  double g = 0.0;
  if (i < kk)
    for (int s = 0; s < Sg; ++s) g += static_cast<double>(gpartial[static_cast<int64_t>(s) * kk + i]);
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double a = 0.0;
    for (int s = 0; s < Sc; ++s) a += colsum_partial[static_cast<int64_t>(s) * K + k];
    a /= count;
    ab[k] = a;
    if (blockIdx.x == 0) { abar[k] = static_cast<float>(a); abar_d[k] = a; }
  }
  __syncthreads();
  if (i >= kk) return;
  const int k = static_cast<int>(i / K), k2 = static_cast<int>(i % K);
  Ghat[i] = static_cast<float>(g - count * ab[k] * ab[k2]);
}

// Q[k,k'] = sum_c kappa_c W[c,k] W[c,k'].  grid (K/32, K/32, CS), 256 threads, each 2x2 outputs; slice z
// of the channels goes to partial[z] (reduced in a fixed order by q_finish_kernel / reduce_partials_kernel).  The 32 x 32
// output tile walks over its channels in chunks of 32 staged in shared memory.  The tiles of W (parameters) of the
// first kQPre chunks are requested before the grid-dependency wait and all at once; kappa / alpha after it.
// upartial != nullptr: the CTAs of the first tile column also emit  upartial[z][k] = sum_{c in slice z} alpha_c W[c,k],
// the first term of u' (they hold those tiles of W anyway).
constexpr int kPoolQSplits = 8;
constexpr int kQPre = 4;
__global__ void __launch_bounds__(256)
pool_q_kernel(const float* __restrict__ W, const float* __restrict__ kappa, const float* __restrict__ alpha, int C, int K,
              float* __restrict__ partial, float* __restrict__ upartial) {
  __shared__ float Wa[32][33], Wb[32][33];
  __shared__ float ured[8][32];
  const int k0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int per = (C + gridDim.z - 1) / gridDim.z;
  const int cbeg = blockIdx.z * per, cend = min(C, cbeg + per);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 5, lc = threadIdx.x & 31;   // loader: 8 rows x 32 columns per pass
  const bool with_u = upartial != nullptr && blockIdx.y == 0;
  float wa[kQPre][4], wb[kQPre][4];
#pragma unroll
  for (int it = 0; it < kQPre; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = cbeg + 32 * it + lr + 8 * i;
      const bool ok = c < cend;
      wa[it][i] = (ok && k0 + lc < K) ? W[static_cast<int64_t>(c) * K + k0 + lc] : 0.f;
      wb[it][i] = (ok && j0 + lc < K) ? W[static_cast<int64_t>(c) * K + j0 + lc] : 0.f;
    }
  pdl_entry();
  float kp[kQPre][4], al[kQPre][4];
#pragma unroll
  for (int it = 0; it < kQPre; ++it)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = cbeg + 32 * it + lr + 8 * i;
      kp[it][i] = c < cend ? kappa[c] : 0.f;
      al[it][i] = (with_u && c < cend) ? alpha[c] : 0.f;
    }
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float uacc = 0.f;
  auto chunk = [&](const float (&ra)[4], const float (&rb)[4]) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) { Wa[lr + 8 * i][lc] = ra[i]; Wb[lr + 8 * i][lc] = rb[i]; }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; ++r) {
      const float a0 = Wa[r][ty * 2], a1 = Wa[r][ty * 2 + 1], b0 = Wb[r][tx * 2], b1 = Wb[r][tx * 2 + 1];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
  };
#pragma unroll
  for (int it = 0; it < kQPre; ++it) {
    if (cbeg + 32 * it >= cend) break;
    float ra[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ra[i] = kp[it][i] * wa[it][i]; uacc = fmaf(al[it][i], wa[it][i], uacc); }
    chunk(ra, wb[it]);
  }
  for (int c0 = cbeg + 32 * kQPre; c0 < cend; c0 += 32) {      // slices longer than kQPre chunks
    float ra[4], rb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = c0 + lr + 8 * i;
      const bool ok = c < cend;
      const float wk = (ok && k0 + lc < K) ? W[static_cast<int64_t>(c) * K + k0 + lc] : 0.f;
      ra[i] = ok ? kappa[c] * wk : 0.f;
      if (with_u && ok) uacc = fmaf(alpha[c], wk, uacc);
      rb[i] = (ok && j0 + lc < K) ? W[static_cast<int64_t>(c) * K + j0 + lc] : 0.f;
    }
    chunk(ra, rb);
  }
  float* Q = partial + static_cast<int64_t>(blockIdx.z) * K * K;
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int k = k0 + ty * 2 + i, k2 = j0 + tx * 2 + j;
      if (k < K && k2 < K) Q[static_cast<int64_t>(k) * K + k2] = acc[i][j];
    }
  if (with_u) {
    ured[lr][lc] = uacc;
    __syncthreads();
    if (lr == 0 && k0 + lc < K) {
      float t = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) t += ured[r][lc];
      upartial[static_cast<int64_t>(blockIdx.z) * K + k0 + lc] = t;
    }
  }
}

// u'[k] = sum_c alpha_c W[c,k] - sum_k' Q[k',k] abar_k'.  grid K/32, 1024 threads: lane = k, the 32 warps
// interleave the summation index; partial sums are combined in a fixed order.
__global__ void __launch_bounds__(1024)
pool_u_kernel(const float* __restrict__ W, const float* __restrict__ alpha, const float* __restrict__ Q,
              const float* __restrict__ abar, int C, int K, float* __restrict__ u) {
  pdl_entry();
  __shared__ double part[32][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  double s = 0.0;
  if (k < K) {
    float f = 0.f;
#pragma unroll 8
    for (int c = warp; c < C; c += 32) f = fmaf(alpha[c], W[static_cast<int64_t>(c) * K + k], f);
    s = static_cast<double>(f);
    f = 0.f;
#pragma unroll 8
    for (int k2 = warp; k2 < K; k2 += 32) f = fmaf(Q[static_cast<int64_t>(k2) * K + k], abar[k2], f);
    s -= static_cast<double>(f);
  }
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0 && k < K) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 32; ++w) t += part[w][lane];
    u[k] = static_cast<float>(t);
  }
}

// Sparse part of the pooled layer's dgrad, applied AFTER the dense low-rank GEMM (tc::pool_dgrad wrote
// dz_prev = mask * (-u - a Q) and its column sums):
//   S[m, :] = sum_{c : arg[b,c] == n, coef != 0} coef[b,c] * W[c, :]            (ascending c: deterministic)
//   dz_prev[m, :] += mask(m, :) * S[m, :]     and the dbeta / dgamma sums of the previous layer likewise
// (or grad_x[b, :, n] += S[m, :] when the previous "layer" is the network input).
// Only <= min(C, N) points per cloud are touched.  The kernel is a chain of dependent L2 round trips, not
// bandwidth (measured at B=8, N=300: one selected point costs ~10k cycles when done naively, a point
// selected by 140 channels 33k), so it is organised to keep the chains short and even:
//   * prologue (per CTA): sarg[c] = selected point (or -1), first[n] = lowest channel selecting point n,
//     then the HEAD channels (first[sarg[c]] == c: one per touched point, one writer per point, no
//     atomics on dz_prev) are compacted into a list, so the warps of the grid take heads round-robin
//     — at most ceil(heads / warps) each, instead of whatever a static channel split happens to give;
//   * per head: the row of dz_prev / y_prev and the BN constants are requested FIRST, then the channels
//     selecting the same point are compacted with ballots over sarg (4 groups of 32 per step), then their
//     rows of W stream with kSparseInFlight independent 128-bit loads in flight, summed in ascending order.
// grid (B, parts), 512 threads; dynamic smem: see pool_sparse_smem().
// KPL: output columns per lane, 4 (K <= 128) or 16 (K <= 512); lane owns [KPL*lane, KPL*lane + KPL).
// IF: rows of W in flight per lane; MINB = 2 caps the registers at 64 so two CTAs share an SM (many points
// selected by one channel each, N >= C: more warps hide more latency than deeper batches do).
template <int KPL, int IF, int MINB>
__global__ void __launch_bounds__(512, MINB)
pool_sparse_kernel(const int32_t* __restrict__ arg, const float* __restrict__ coef, const float* __restrict__ W,
                   int C, int N, int K, DgradOut o) {
  pdl_entry();
  constexpr int kSparseInFlight = IF;
  extern __shared__ int sm_i[];
  const int nchunk = (C + 31) >> 5;
  int* sarg = sm_i;                                        // [C]
  int* first = sarg + C;                                   // [N]
  int* cnt = first + N;                                    // [nchunk + 1] heads per 32-channel chunk -> exclusive offsets
  unsigned* hmask = reinterpret_cast<unsigned*>(cnt + nchunk + 1);   // [nchunk]
  uint16_t* heads = reinterpret_cast<uint16_t*>(hmask + nchunk);     // [C]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  uint16_t* lst = heads + ((C + 1) & ~1) + static_cast<size_t>(warp) * C;   // [nwarp][C]
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    sarg[c] = coef[static_cast<int64_t>(b) * C + c] != 0.f ? arg[static_cast<int64_t>(b) * C + c] : -1;
  for (int n = threadIdx.x; n < N; n += blockDim.x) first[n] = 0x7fffffff;
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    if (sarg[c] >= 0) atomicMin(&first[sarg[c]], c);
  __syncthreads();
  // head list, ascending: per-chunk ballots, a one-warp scan of the chunk counts, then the scatter
  for (int ch = warp; ch < nchunk; ch += nwarp) {
    const int c = ch * 32 + lane;
    const bool is_head = c < C && sarg[c] >= 0 && first[sarg[c]] == c;
    const unsigned m = __ballot_sync(0xffffffffu, is_head);
    if (lane == 0) { hmask[ch] = m; cnt[ch] = __popc(m); }
  }
  __syncthreads();
  if (warp == 0) {
    int base = 0;
    for (int c0 = 0; c0 < nchunk; c0 += 32) {
      const int v = c0 + lane < nchunk ? cnt[c0 + lane] : 0;
      int incl = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (c0 + lane < nchunk) cnt[c0 + lane] = base + incl - v;
      base += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) cnt[nchunk] = base;
  }
  __syncthreads();
  for (int ch = warp; ch < nchunk; ch += nwarp) {
    const unsigned m = hmask[ch];
    if ((m >> lane) & 1u) heads[cnt[ch] + __popc(m & ((1u << lane) - 1u))] = static_cast<uint16_t>(ch * 32 + lane);
  }
  __syncthreads();
  const int H = cnt[nchunk];

  const bool to_x = o.grad_x != nullptr;
  const int k0 = KPL * lane;
  // rows of W / dz_prev / y_prev (and the per-channel BN vectors) are 16-byte aligned runs
  const bool vec = (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(o.y_prev) |
                                     reinterpret_cast<uintptr_t>(o.dz_prev)) & 15u) == 0;
  float ws[KPL], wq[KPL];                        // this warp's share of the previous layer's dbeta / dgamma sums
#pragma unroll
  for (int jj = 0; jj < KPL; ++jj) { ws[jj] = 0.f; wq[jj] = 0.f; }
  // per-channel BN constants of this lane's columns: loaded once when they fit in registers
  constexpr bool kKeepBN = KPL == 4 && MINB == 1;
  constexpr int KB = kKeepBN ? KPL : 1;
  float mean_k[KB], istd_k[KB], gam_k[KB], bet_k[KB];
#pragma unroll
  for (int jj = 0; jj < KB; ++jj) {
    const bool ok = kKeepBN && !to_x && k0 + jj < K;
    mean_k[jj] = ok ? o.mean[k0 + jj] : 0.f;
    istd_k[jj] = ok ? o.invstd[k0 + jj] : 0.f;
    gam_k[jj] = ok ? o.gamma[k0 + jj] : 0.f;
    bet_k[jj] = ok ? o.beta[k0 + jj] : 0.f;
  }
  // heads are in ascending channel order, which is roughly descending popularity (a point many channels select is
  // very likely selected by an early one): consecutive heads go to different CTAs, so every warp of the grid gets one
  // early (long) and one late (short) head instead of one CTA getting the 16 longest
  for (int h = warp * gridDim.y + blockIdx.y; h < H; h += nwarp * gridDim.y) {
    const int c = heads[h];
    const int a = sarg[c];
    const int64_t mrow = static_cast<int64_t>(b) * N + a;
    // the read-modify-write operands first: their latency overlaps everything below
    float yv[KPL], dv[KPL];
#pragma unroll
    for (int jj = 0; jj < KPL; ++jj) { yv[jj] = 0.f; dv[jj] = 0.f; }
    if (!to_x) {
      if (vec) {
#pragma unroll
        for (int v = 0; v < KPL / 4; ++v) {
          if (k0 + 4 * v < K) {
            const float4 ty = *(reinterpret_cast<const float4*>(o.y_prev + mrow * K + k0) + v);
            const float4 td = *(reinterpret_cast<const float4*>(o.dz_prev + mrow * K + k0) + v);
            yv[4 * v + 0] = ty.x; yv[4 * v + 1] = ty.y; yv[4 * v + 2] = ty.z; yv[4 * v + 3] = ty.w;
            dv[4 * v + 0] = td.x; dv[4 * v + 1] = td.y; dv[4 * v + 2] = td.z; dv[4 * v + 3] = td.w;
          }
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < KPL; ++jj) {
          const int k = k0 + jj;
          if (k < K) { yv[jj] = o.y_prev[mrow * K + k]; dv[jj] = o.dz_prev[mrow * K + k]; }
        }
      }
    }
    // pass 1: the channels selecting point a, ascending, into this warp's list (4 groups of 32 per step)
    int m = 0;
    for (int c0 = c & ~31; c0 < C; c0 += 128) {
      unsigned mk[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int cc = c0 + 32 * u + lane;
        mk[u] = __ballot_sync(0xffffffffu, cc >= c && cc < C && sarg[cc] == a);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if ((mk[u] >> lane) & 1u) lst[m + __popc(mk[u] & ((1u << lane) - 1u))] = static_cast<uint16_t>(c0 + 32 * u + lane);
        m += __popc(mk[u]);
      }
    }
    __syncwarp();
    // pass 2: S = sum coef * W[row, :], rows in ascending order, kSparseInFlight rows loaded before the first use
    float acc[KPL];
#pragma unroll
    for (int jj = 0; jj < KPL; ++jj) acc[jj] = 0.f;
    for (int i = 0; i < m; i += kSparseInFlight) {
      float cf[kSparseInFlight];
      float wv[kSparseInFlight][KPL];
#pragma unroll
      for (int u = 0; u < kSparseInFlight; ++u) {
        const int cs = i + u < m ? static_cast<int>(lst[i + u]) : -1;
        cf[u] = cs >= 0 ? coef[static_cast<int64_t>(b) * C + cs] : 0.f;
        const float* wr = W + static_cast<int64_t>(cs >= 0 ? cs : 0) * K + k0;
        if (vec) {
#pragma unroll
          for (int v = 0; v < KPL / 4; ++v) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cs >= 0 && k0 + 4 * v < K) t = __ldg(reinterpret_cast<const float4*>(wr) + v);
            wv[u][4 * v + 0] = t.x; wv[u][4 * v + 1] = t.y; wv[u][4 * v + 2] = t.z; wv[u][4 * v + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int jj = 0; jj < KPL; ++jj) wv[u][jj] = (cs >= 0 && k0 + jj < K) ? __ldg(wr + jj) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < kSparseInFlight; ++u)
#pragma unroll
        for (int jj = 0; jj < KPL; ++jj) acc[jj] = fmaf(cf[u], wv[u][jj], acc[jj]);
    }
    __syncwarp();                                 // the list is rewritten for the next head
    if (to_x) {
#pragma unroll
      for (int jj = 0; jj < KPL; ++jj) {
        const int k = k0 + jj;
        if (k < K) o.grad_x[(static_cast<int64_t>(b) * K + k) * N + a] += acc[jj];
      }
    } else {
#pragma unroll
      for (int jj = 0; jj < KPL; ++jj) {
        const int k = k0 + jj;
        float v = 0.f;
        if (k < K) {
          const float mu = kKeepBN ? mean_k[jj % KB] : o.mean[k], is = kKeepBN ? istd_k[jj % KB] : o.invstd[k];
          const float ga = kKeepBN ? gam_k[jj % KB] : o.gamma[k], be = kKeepBN ? bet_k[jj % KB] : o.beta[k];
          const float yh = (yv[jj] - mu) * is;
          const bool on = !o.relu || fmaf(yh, ga, be) > 0.f;
          v = on ? acc[jj] : 0.f;
          ws[jj] += v;
          wq[jj] = fmaf(v, yh, wq[jj]);
        }
        dv[jj] += v;
      }
      if (vec) {
#pragma unroll
        for (int v = 0; v < KPL / 4; ++v)
          if (k0 + 4 * v < K)
            *(reinterpret_cast<float4*>(o.dz_prev + mrow * K + k0) + v) =
                make_float4(dv[4 * v + 0], dv[4 * v + 1], dv[4 * v + 2], dv[4 * v + 3]);
      } else {
#pragma unroll
        for (int jj = 0; jj < KPL; ++jj)
          if (k0 + jj < K) o.dz_prev[mrow * K + k0 + jj] = dv[jj];
      }
    }
  }
  if (to_x) return;
  // fixed-order combination of the warps' sums, one fp64 atomic per channel and CTA
  __syncthreads();
  float* red = reinterpret_cast<float*>(sm_i);            // [2][nwarp][K]
#pragma unroll
  for (int jj = 0; jj < KPL; ++jj) {
    const int k = k0 + jj;
    if (k < K) { red[warp * K + k] = ws[jj]; red[(nwarp + warp) * K + k] = wq[jj]; }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double sd = 0.0, qd = 0.0;
    for (int w = 0; w < nwarp; ++w) { sd += static_cast<double>(red[w * K + k]); qd += static_cast<double>(red[(nwarp + w) * K + k]); }
    if (sd != 0.0 || qd != 0.0) { atomicAdd(&o.sums[k], sd); atomicAdd(&o.sums[K + k], qd); }
  }
}

// dynamic shared memory of pool_sparse_kernel: sarg[C] first[N] cnt[nchunk+1] hmask[nchunk] ints, heads[C] +
// 16 match lists of C uint16; re-used as [2][16][K] floats by the final reduction
static size_t pool_sparse_smem(int C, int N, int K) {
  const size_t nchunk = (static_cast<size_t>(C) + 31) / 32;
  const size_t a = sizeof(int) * (static_cast<size_t>(C) + N + 2 * nchunk + 1) +
                   sizeof(uint16_t) * (((static_cast<size_t>(C) + 1) & ~static_cast<size_t>(1)) + 16 * static_cast<size_t>(C));
  return std::max(a, sizeof(float) * 2 * 16 * static_cast<size_t>(K));
}

// ---- pool_sparse_sorted_kernel: the same update for K <= 128 and fewer points than channels (the 300-point clouds) ----
// There a handful of points are selected by 100+ channels each and most by one to three, so "one warp per point" leaves
// the kernel waiting for the few warps that walk ceil(m / IF) dependent round trips (measured, B=8 N=300: median CTA
// done after 6 us, the last after 14-27 us).  Here the channels of a cloud are first sorted by selected point (stable
// counting sort in shared memory: ascending channel inside a point), the sorted rows are cut into gridDim.y CTA ranges at
// point boundaries and every CTA range into 16 equal warp segments regardless of point boundaries, so every warp loads
// <= ceil(range / 16) rows of W, normally in ONE round trip:
//   phase A  a warp sums its segment point by point; a point that lies inside the segment goes to sbuf[point], the (at
//            most two) points cut by a segment end go to part[2 * warp + {0: cut on the left, 1: cut on the right only}];
//   phase B  warp w owns points w, w + 16, ... of the CTA range: S = sbuf[point], or the partials of the warps the point
//            spans in warp order (= ascending channel: deterministic), then the masked read-modify-write of dz_prev and
//            the BN sums.  The rows of y_prev / dz_prev of a warp's first two points are requested before phase A.
// grid (B, parts), 512 threads.  Needs N <= 1024, C <= 65535, K % 4 == 0, 16-byte aligned rows.
constexpr int kSsWarps = 16;
constexpr int kSsIF = 12;      // rows of W in flight per lane
constexpr int kSsPre = 2;      // points per warp whose read-modify-write operands are requested up front

struct SsLayout { size_t sarg, start, wtot, gpoint, order, cm, bn, part, sbuf, bytes; };
__host__ __device__ inline SsLayout ss_layout(int C, int N, int K, int sb_rows) {
  const size_t nchunk = (static_cast<size_t>(C) + 31) / 32;
  auto even = [](size_t v) { return (v + 1) & ~static_cast<size_t>(1); };
  SsLayout l{};
  size_t off = 0;
  l.sarg = off;   off += sizeof(int) * static_cast<size_t>(C);
  l.start = off;  off += sizeof(uint32_t) * (static_cast<size_t>(N) + 1);
  l.wtot = off;   off += sizeof(uint32_t) * 32;
  l.gpoint = off; off += sizeof(uint16_t) * even(N);
  l.order = off;  off += sizeof(uint16_t) * even(C);
  l.cm = off;     off += sizeof(uint16_t) * even(nchunk * N);
  off = (off + 15) & ~static_cast<size_t>(15);
  l.bn = off;     off += sizeof(float) * 4 * static_cast<size_t>(K);                 // mean | invstd | gamma | beta of the previous layer
  l.part = off;   off += sizeof(float) * 2 * kSsWarps * static_cast<size_t>(K);      // also the final [2][16][K] reduction
  l.sbuf = off;   off += sizeof(float) * static_cast<size_t>(sb_rows) * K;
  l.bytes = off;
  return l;
}

__global__ void __launch_bounds__(32 * kSsWarps, 1)
pool_sparse_sorted_kernel(const int32_t* __restrict__ arg, const float* __restrict__ coef, const float* __restrict__ W,
                          int C, int N, int K, int sb_rows, DgradOut o) {
  pdl_entry();
  extern __shared__ __align__(16) unsigned char ss_raw[];
  const SsLayout L = ss_layout(C, N, K, sb_rows);
  int* sarg = reinterpret_cast<int*>(ss_raw + L.sarg);                 // [C] selected point, -1 when coef == 0
  uint32_t* start = reinterpret_cast<uint32_t*>(ss_raw + L.start);     // [N + 1] low 16: sorted rows before point n; high 16: non-empty points before n
  uint32_t* wtot = reinterpret_cast<uint32_t*>(ss_raw + L.wtot);
  uint16_t* gpoint = reinterpret_cast<uint16_t*>(ss_raw + L.gpoint);   // [#non-empty] their point ids, ascending
  uint16_t* order = reinterpret_cast<uint16_t*>(ss_raw + L.order);     // [T] channels sorted by (point, channel)
  uint16_t* cm = reinterpret_cast<uint16_t*>(ss_raw + L.cm);           // [nchunk][N] channels of chunk ch selecting point n -> prefix over ch
  float* part = reinterpret_cast<float*>(ss_raw + L.part);
  float* sbuf = reinterpret_cast<float*>(ss_raw + L.sbuf);
  const int nchunk = (C + 31) >> 5;
  const int b = blockIdx.x, P = gridDim.y, p = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt = (1u << lane) - 1u;
  const int k0 = 4 * lane;
  const bool kin = k0 < K;
  // per-channel BN constants of the previous layer: requested first, used last
  float* bn = reinterpret_cast<float*>(ss_raw + L.bn);
  for (int i = tid; i < 4 * K; i += blockDim.x) {
    const int which = i / K, k = i - which * K;
    bn[i] = which == 0 ? o.mean[k] : which == 1 ? o.invstd[k] : which == 2 ? o.gamma[k] : o.beta[k];
  }
  for (int c = tid; c < C; c += blockDim.x)
    sarg[c] = coef[static_cast<int64_t>(b) * C + c] != 0.f ? arg[static_cast<int64_t>(b) * C + c] : -1;
  {
    uint32_t* cz = reinterpret_cast<uint32_t*>(cm);
    const int nw = (nchunk * N + 1) >> 1;
    for (int i = tid; i < nw; i += blockDim.x) cz[i] = 0u;
  }
  __syncthreads();
  // channels of each 32-channel chunk per point (the lowest lane of a group of equal points writes the group size)
  for (int ch = warp; ch < nchunk; ch += kSsWarps) {
    const int c = ch * 32 + lane;
    const int a = c < C ? sarg[c] : -1;
    const unsigned mk = __match_any_sync(0xffffffffu, a);
    if (a >= 0 && lane == __ffs(mk) - 1) cm[ch * N + a] = static_cast<uint16_t>(__popc(mk));
  }
  __syncthreads();
  // per point: exclusive prefix over the chunks (in place) and the total
  for (int n = tid; n < N; n += blockDim.x) {
    uint32_t run = 0;
    for (int ch = 0; ch < nchunk; ++ch) {
      const uint32_t t = cm[ch * N + n];
      cm[ch * N + n] = static_cast<uint16_t>(run);
      run += t;
    }
    start[n] = run | (run ? 0x10000u : 0u);
  }
  __syncthreads();
  // exclusive scan over the points of (rows, non-empty flag) packed in one word: two points per thread
  {
    const int i0 = 2 * tid;
    const uint32_t v0 = i0 < N ? start[i0] : 0u, v1 = i0 + 1 < N ? start[i0 + 1] : 0u;
    const uint32_t s = v0 + v1;
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    uint32_t base = 0;
    for (int w = 0; w < warp; ++w) base += wtot[w];
    const uint32_t excl = base + incl - s;
    if (i0 < N) start[i0] = excl;
    if (i0 + 1 < N) start[i0 + 1] = excl + v0;
    if (tid == 32 * kSsWarps - 1) start[N] = base + incl;
  }
  __syncthreads();
  for (int ch = warp; ch < nchunk; ch += kSsWarps) {
    const int c = ch * 32 + lane;
    const int a = c < C ? sarg[c] : -1;
    const unsigned mk = __match_any_sync(0xffffffffu, a);
    if (a >= 0) order[(start[a] & 0xffffu) + cm[ch * N + a] + __popc(mk & lt)] = static_cast<uint16_t>(c);
  }
  for (int n = tid; n < N; n += blockDim.x)
    if ((start[n + 1] >> 16) != (start[n] >> 16)) gpoint[start[n] >> 16] = static_cast<uint16_t>(n);
  __syncthreads();
  const int T = static_cast<int>(start[N] & 0xffffu), Gtot = static_cast<int>(start[N] >> 16);
  // this CTA's rows [R0, R1): cuts at q * T / P moved down to the start of the point they fall in
  auto cut = [&](int q) -> int {
    if (q >= P) return T;
    const int r = static_cast<int>(static_cast<int64_t>(q) * T / P);
    return r >= T ? T : static_cast<int>(start[sarg[order[r]]] & 0xffffu);
  };
  const int R0 = cut(p), R1 = cut(p + 1);
  const int len = R1 - R0;
  float ws[4] = {0.f, 0.f, 0.f, 0.f}, wq[4] = {0.f, 0.f, 0.f, 0.f};
  if (len > 0) {
    const int g0 = static_cast<int>(start[sarg[order[R0]]] >> 16);
    const int G = (R1 < T ? static_cast<int>(start[sarg[order[R1]]] >> 16) : Gtot) - g0;
    if (G > sb_rows) __trap();                      // cannot happen: a range holds <= ceil(T / P) + 1 points
    const int per = (len + kSsWarps - 1) / kSsWarps;
    const int s0 = R0 + min(len, warp * per), s1 = R0 + min(len, warp * per + per);
    // the rows this warp will read-modify-write in phase B: in flight during phase A
    float4 yp[kSsPre], dp[kSsPre];
#pragma unroll
    for (int t = 0; t < kSsPre; ++t) {
      yp[t] = make_float4(0.f, 0.f, 0.f, 0.f); dp[t] = yp[t];
      const int gi = warp + t * kSsWarps;
      if (gi < G && kin) {
        const int64_t mrow = static_cast<int64_t>(b) * N + gpoint[g0 + gi];
        yp[t] = *reinterpret_cast<const float4*>(o.y_prev + mrow * K + k0);
        dp[t] = *reinterpret_cast<const float4*>(o.dz_prev + mrow * K + k0);
      }
    }
    // phase A
    {
      int cur = -1;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      auto emit = [&]() {
        const int st = static_cast<int>(start[cur] & 0xffffu), en = static_cast<int>(start[cur + 1] & 0xffffu);
        float* dst = st < s0 ? part + static_cast<size_t>(2 * warp) * K
                   : en > s1 ? part + static_cast<size_t>(2 * warp + 1) * K
                             : sbuf + static_cast<size_t>(static_cast<int>(start[cur] >> 16) - g0) * K;
        if (kin) *reinterpret_cast<float4*>(dst + k0) = acc;
      };
      for (int i = s0; i < s1; i += kSsIF) {
        float cf[kSsIF];
        float4 wv[kSsIF];
#pragma unroll
        for (int u = 0; u < kSsIF; ++u) {
          const int cs = i + u < s1 ? static_cast<int>(order[i + u]) : -1;
          cf[u] = cs >= 0 ? coef[static_cast<int64_t>(b) * C + cs] : 0.f;
          wv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cs >= 0 && kin) wv[u] = __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(cs) * K + k0));
        }
#pragma unroll
        for (int u = 0; u < kSsIF; ++u) {
          if (i + u >= s1) break;
          const int a = sarg[order[i + u]];
          if (a != cur) {
            if (cur >= 0) emit();
            cur = a;
            acc = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          acc.x = fmaf(cf[u], wv[u].x, acc.x); acc.y = fmaf(cf[u], wv[u].y, acc.y);
          acc.z = fmaf(cf[u], wv[u].z, acc.z); acc.w = fmaf(cf[u], wv[u].w, acc.w);
        }
      }
      if (cur >= 0) emit();
    }
    __syncthreads();
    // phase B
    auto do_point = [&](int gi, float4 y4, float4 d4) {
      const int a = gpoint[g0 + gi];
      const int st = static_cast<int>(start[a] & 0xffffu), en = static_cast<int>(start[a + 1] & 0xffffu);
      const int wf = (st - R0) / per, wl = (en - 1 - R0) / per;
      float4 S = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kin) {
        if (wf == wl) {
          S = *reinterpret_cast<const float4*>(sbuf + static_cast<size_t>(gi) * K + k0);
        } else {
          S = *reinterpret_cast<const float4*>(part + static_cast<size_t>(2 * wf + 1) * K + k0);
          for (int w = wf + 1; w <= wl; ++w) {
            const float4 t = *reinterpret_cast<const float4*>(part + static_cast<size_t>(2 * w) * K + k0);
            S.x += t.x; S.y += t.y; S.z += t.z; S.w += t.w;
          }
        }
      }
      const float yv[4] = {y4.x, y4.y, y4.z, y4.w}, sv[4] = {S.x, S.y, S.z, S.w};
      float4 mean4 = make_float4(0.f, 0.f, 0.f, 0.f), istd4 = mean4, gam4 = mean4, bet4 = mean4;
      if (kin) {
        mean4 = *reinterpret_cast<const float4*>(bn + k0);         istd4 = *reinterpret_cast<const float4*>(bn + K + k0);
        gam4 = *reinterpret_cast<const float4*>(bn + 2 * K + k0);  bet4 = *reinterpret_cast<const float4*>(bn + 3 * K + k0);
      }
      const float mu[4] = {mean4.x, mean4.y, mean4.z, mean4.w}, is[4] = {istd4.x, istd4.y, istd4.z, istd4.w};
      const float ga[4] = {gam4.x, gam4.y, gam4.z, gam4.w}, be[4] = {bet4.x, bet4.y, bet4.z, bet4.w};
      float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float yh = (yv[jj] - mu[jj]) * is[jj];
        const bool on = !o.relu || fmaf(yh, ga[jj], be[jj]) > 0.f;
        const float v = (kin && on) ? sv[jj] : 0.f;
        ws[jj] += v;
        wq[jj] = fmaf(v, yh, wq[jj]);
        dv[jj] += v;
      }
      if (kin)
        *reinterpret_cast<float4*>(o.dz_prev + (static_cast<int64_t>(b) * N + a) * K + k0) = make_float4(dv[0], dv[1], dv[2], dv[3]);
    };
#pragma unroll
    for (int t = 0; t < kSsPre; ++t)
      if (warp + t * kSsWarps < G) do_point(warp + t * kSsWarps, yp[t], dp[t]);
    for (int gi = warp + kSsPre * kSsWarps; gi < G; gi += kSsWarps) {
      float4 y4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = y4;
      if (kin) {
        const int64_t mrow = static_cast<int64_t>(b) * N + gpoint[g0 + gi];
        y4 = *reinterpret_cast<const float4*>(o.y_prev + mrow * K + k0);
        d4 = *reinterpret_cast<const float4*>(o.dz_prev + mrow * K + k0);
      }
      do_point(gi, y4, d4);
    }
  }
  // fixed-order combination of the warps' sums, one fp64 atomic per channel and CTA
  __syncthreads();
  float* red = part;                                       // [2][16][K]
  if (kin) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { red[warp * K + k0 + jj] = ws[jj]; red[(kSsWarps + warp) * K + k0 + jj] = wq[jj]; }
  }
  __syncthreads();
  for (int k = tid; k < K; k += blockDim.x) {
    double sd = 0.0, qd = 0.0;
    for (int w = 0; w < kSsWarps; ++w) { sd += static_cast<double>(red[w * K + k]); qd += static_cast<double>(red[(kSsWarps + w) * K + k]); }
    if (sd != 0.0 || qd != 0.0) { atomicAdd(&o.sums[k], sd); atomicAdd(&o.sums[K + k], qd); }
  }
}

// dW[c,k] = sum_b coef[b,c] (a[(b,sel),k] - abar_k) - kappa_c sum_k' W[c,k'] Ghat[k',k]
// A CTA owns kDwCh channels (their rows of W staged in shared memory), a thread owns output columns k:
// one coalesced read of a row of Ghat serves kDwCh accumulators (the previous one-thread-per-output
// version walked K dependent loads per output: 17 us at C=1024, K=128).
constexpr int kDwCh = 8;
constexpr int kDwParts = 4;                    // the two contraction loops are split 4 ways inside the CTA
__global__ void __launch_bounds__(128 * kDwParts)
pool_dw_kernel(ActSrc src, const float* __restrict__ W, const float* __restrict__ coef,
               const int32_t* __restrict__ arg, const float* __restrict__ kappa,
               const float* __restrict__ abar, const float* __restrict__ Ghat, int B, int N,
               int C, int train, float* __restrict__ dW) {
  // dynamic smem: wrow[kDwCh][K] | sel_cf[B][kDwCh] | sel_m[B][kDwCh] | red[kDwParts][128][kDwCh] doubles
  extern __shared__ __align__(16) float wrow[];
  const int K = src.C;
  float* sel_cf = wrow + kDwCh * K;
  int* sel_m = reinterpret_cast<int*>(sel_cf + B * kDwCh);
  double* red = reinterpret_cast<double*>(wrow + ((kDwCh * (K + 2 * B) + 3) & ~3));
  const int c0 = blockIdx.x * kDwCh;
  const int tk = threadIdx.x & 127, part = threadIdx.x >> 7;
  for (int i = threadIdx.x; i < kDwCh * K; i += blockDim.x) {
    const int j = i / K, k2 = i - j * K;
    wrow[i] = c0 + j < C ? W[static_cast<int64_t>(c0 + j) * K + k2] : 0.f;
  }
  pdl_entry();                                   // the rows of W (parameters) were staged before the wait
  // the selected point and its coefficient of every (cloud, channel) of this CTA: one round trip here
  // instead of a dependent coef -> arg -> activation chain per term below
  for (int i = threadIdx.x; i < B * kDwCh; i += blockDim.x) {
    const int b = i / kDwCh, c = c0 + (i - b * kDwCh);
    const float cf = c < C ? coef[static_cast<int64_t>(b) * C + c] : 0.f;
    sel_cf[i] = cf;
    sel_m[i] = cf != 0.f ? b * N + arg[static_cast<int64_t>(b) * C + c] : 0;
  }
  __syncthreads();
  for (int k0 = 0; k0 < K; k0 += 128) {
    const int k = k0 + tk;
    double acc[kDwCh];                         // this part's share of  sum_b coef (a - abar) - kappa * (W Ghat)
#pragma unroll
    for (int j = 0; j < kDwCh; ++j) acc[j] = 0.0;
    if (k < K) {
      const int per = (K + kDwParts - 1) / kDwParts;
      const int k2a = part * per, k2b = min(K, k2a + per);
      // fp32 inside a part's K/4-long slice (two sub-chains), the parts are combined in double below
      float t0[kDwCh], t1[kDwCh];
#pragma unroll
      for (int j = 0; j < kDwCh; ++j) { t0[j] = 0.f; t1[j] = 0.f; }
      int k2 = k2a;
#pragma unroll 4
      for (; k2 + 1 < k2b; k2 += 2) {
        const float g0 = Ghat[static_cast<int64_t>(k2) * K + k], g1 = Ghat[static_cast<int64_t>(k2 + 1) * K + k];
#pragma unroll
        for (int j = 0; j < kDwCh; ++j) { t0[j] = fmaf(wrow[j * K + k2], g0, t0[j]); t1[j] = fmaf(wrow[j * K + k2 + 1], g1, t1[j]); }
      }
      if (k2 < k2b) {
        const float g0 = Ghat[static_cast<int64_t>(k2) * K + k];
#pragma unroll
        for (int j = 0; j < kDwCh; ++j) t0[j] = fmaf(wrow[j * K + k2], g0, t0[j]);
      }
      double t[kDwCh];
#pragma unroll
      for (int j = 0; j < kDwCh; ++j) t[j] = static_cast<double>(t0[j]) + static_cast<double>(t1[j]);
      // train: sum_b coef = M*alpha, which folds the rank-one BN term into a centring of a[sel];
      // eval: there is no BN correction at all
      const float ab = train ? abar[k] : 0.f;
      float sc = 1.f, sh = 0.f;                 // a = relu?(y * sc + sh) for a BatchNorm'd source
      if (src.y != nullptr) {
        sc = src.invstd[k] * src.gamma[k];
        sh = src.beta[k] - src.mean[k] * sc;
      }
      for (int b = part; b < B; b += kDwParts) {
        float av[kDwCh];
#pragma unroll
        for (int j = 0; j < kDwCh; ++j) {       // independent loads: kDwCh in flight
          const int m = sel_m[b * kDwCh + j];
          if (src.y != nullptr) av[j] = src.y[static_cast<int64_t>(m) * K + k];
          else av[j] = load_act1(src, m, k);
        }
#pragma unroll
        for (int j = 0; j < kDwCh; ++j) {
          float a = av[j];
          if (src.y != nullptr) { a = fmaf(a, sc, sh); a = src.relu ? fmaxf(a, 0.f) : a; }
          acc[j] += static_cast<double>(sel_cf[b * kDwCh + j]) * static_cast<double>(a - ab);
        }
      }
#pragma unroll
      for (int j = 0; j < kDwCh; ++j) {
        const int c = c0 + j;
        acc[j] -= (c < C ? static_cast<double>(kappa[c]) : 0.0) * t[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kDwCh; ++j) red[(part * 128 + tk) * kDwCh + j] = acc[j];
    __syncthreads();
    if (part == 0 && k < K) {
#pragma unroll
      for (int j = 0; j < kDwCh; ++j) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < kDwParts; ++q) v += red[(q * 128 + tk) * kDwCh + j];   // fixed order
        if (c0 + j < C) dW[static_cast<int64_t>(c0 + j) * K + k] = static_cast<float>(v);
      }
    }
  }
}

__device__ __forceinline__ void dgrad_epilogue(const DgradOut& o, float (&acc)[4][4], int b, int n0, int N,
                                               int c0, int ty, int tx, float* ssum, float* ssq) {
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
    const int64_t m = static_cast<int64_t>(b) * N + n;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      if (c >= o.Cp) continue;
      float v = acc[i][j];
      if (o.grad_x != nullptr) {
        o.grad_x[(static_cast<int64_t>(b) * o.Cp + c) * N + n] = v;
      } else {
        const float yh = (o.y_prev[m * o.Cp + c] - o.mean[c]) * o.invstd[c];
        if (o.relu && !(fmaf(yh, o.gamma[c], o.beta[c]) > 0.f)) v = 0.f;
        o.dz_prev[m * o.Cp + c] = v;
        cs[j] += v;
        cq[j] = fmaf(v, yh, cq[j]);
      }
    }
  }
  if (o.grad_x != nullptr) return;
  // fixed-order combination of the 16 row groups (see mlp_fwd_kernel): ssum / ssq are [16][TN]
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) { ssum[ty * TN + tx * 4 + j] = cs[j]; ssq[ty * TN + tx * 4 + j] = cq[j]; }
  __syncthreads();
  if (threadIdx.x < TN && c0 + threadIdx.x < o.Cp) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int g = 0; g < 16; ++g) { s += static_cast<double>(ssum[g * TN + threadIdx.x]); q += static_cast<double>(ssq[g * TN + threadIdx.x]); }
    atomicAdd(&o.sums[c0 + threadIdx.x], s);
    atomicAdd(&o.sums[o.Cp + c0 + threadIdx.x], q);
  }
}

// pooled layer dgrad: da[m,k] = sparse[m,k] - u'_k - sum_k' a[m,k'] Q[k',k]
__global__ void __launch_bounds__(kThreads)
pool_dgrad_kernel(ActSrc src, const float* __restrict__ Q, const float* __restrict__ u, const float* __restrict__ W,
                  const float* __restrict__ coef, const int32_t* __restrict__ arg, int C, int N,
                  int tiles_per_sample, DgradOut o) {
  pdl_entry();
  __shared__ Tiles t;
  __shared__ float ssum[16 * TN], ssq[16 * TN];
  __shared__ float sparse[TM][TN + 1];
  const int K = src.C;  // == o.Cp
  const int b = blockIdx.x / tiles_per_sample;
  const int n0 = (blockIdx.x - b * tiles_per_sample) * TM;
  const int c0 = blockIdx.y * TN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  for (int i = threadIdx.x; i < TM * (TN + 1); i += kThreads) (&sparse[0][0])[i] = 0.f;
  __syncthreads();
  // sparse part, deterministic: thread kt owns column kt and walks the channels in order
  if (threadIdx.x < TN && c0 + threadIdx.x < K) {
    const int k = c0 + threadIdx.x;
    for (int c = 0; c < C; ++c) {
      const int r = arg[static_cast<int64_t>(b) * C + c] - n0;
      if (r >= 0 && r < TM) {
        const float cf = coef[static_cast<int64_t>(b) * C + c];
        if (cf != 0.f) sparse[r][threadIdx.x] = fmaf(cf, W[static_cast<int64_t>(c) * K + k], sparse[r][threadIdx.x]);
      }
    }
  }
  float acc[4][4] = {};
  auto a4 = [&](int r, int k) -> float4 {
    if (n0 + r >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
    return load_act4(src, static_cast<int64_t>(b) * N + n0 + r, k);
  };
  auto q4 = [&](int kk, int c4) -> float4 {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float* pv = &v.x;
    if (kk < K) for (int i = 0; i < 4; ++i) if (c0 + c4 + i < K) pv[i] = Q[static_cast<int64_t>(kk) * K + c0 + c4 + i];
    return v;
  };
  for (int k0 = 0; k0 < K; k0 += TK) {
    __syncthreads();
    stage_rows(t.As, a4, k0);
    stage_cols(t.Bs, q4, k0);
    __syncthreads();
    mma_chunk(t, acc, ty, tx);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + tx * 4 + j;
      acc[i][j] = sparse[ty * 4 + i][tx * 4 + j] - (c < K ? u[c] : 0.f) - acc[i][j];
    }
  dgrad_epilogue(o, acc, b, n0, N, c0, ty, tx, ssum, ssq);
}

// dense layer dgrad: da_prev[m,k] = sum_c dy[m,c] W[c,k]
__global__ void __launch_bounds__(kThreads)
dense_dgrad_kernel(DySrc dys, const float* __restrict__ W, int N, int tiles_per_sample, DgradOut o) {
  pdl_entry();
  __shared__ Tiles t;
  __shared__ float ssum[16 * TN], ssq[16 * TN];
  const int C = dys.C, Kp = o.Cp;
  const int b = blockIdx.x / tiles_per_sample;
  const int n0 = (blockIdx.x - b * tiles_per_sample) * TM;
  const int c0 = blockIdx.y * TN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[4][4] = {};
  auto a4 = [&](int r, int c) -> float4 {
    if (n0 + r >= N) return make_float4(0.f, 0.f, 0.f, 0.f);
    return load_dy4(dys, static_cast<int64_t>(b) * N + n0 + r, c);
  };
  auto w4 = [&](int cc, int c4) -> float4 {  // W[cc, c0+c4 .. +3]
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    float* pv = &v.x;
    if (cc < C) for (int i = 0; i < 4; ++i) if (c0 + c4 + i < Kp) pv[i] = W[static_cast<int64_t>(cc) * Kp + c0 + c4 + i];
    return v;
  };
  for (int k0 = 0; k0 < C; k0 += TK) {
    __syncthreads();
    stage_rows(t.As, a4, k0);
    stage_cols(t.Bs, w4, k0);
    __syncthreads();
    mma_chunk(t, acc, ty, tx);
  }
  __syncthreads();
  dgrad_epilogue(o, acc, b, n0, N, c0, ty, tx, ssum, ssq);
}

// alpha_c = s_c*dbeta_c/M, kappa_c = s_c*invstd_c*dgamma_c/M (train) ; grads of gamma/beta/bias
__global__ void bn_bwd_coef_kernel(const double* __restrict__ sums, const float* __restrict__ invstd,
                                   const float* __restrict__ gamma, int C, double count, int train,
                                   float* __restrict__ alpha, float* __restrict__ kappa,
                                   float* __restrict__ grad_gamma, float* __restrict__ grad_beta,
                                   float* __restrict__ grad_bias) {
  pdl_entry();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double dbeta = sums[c], dgamma = sums[C + c];
  const double sc = static_cast<double>(gamma[c]) * static_cast<double>(invstd[c]);
  alpha[c] = train ? static_cast<float>(sc * dbeta / count) : 0.f;
  kappa[c] = train ? static_cast<float>(sc * static_cast<double>(invstd[c]) * dgamma / count) : 0.f;
  if (grad_gamma) grad_gamma[c] = static_cast<float>(dgamma);
  if (grad_beta) grad_beta[c] = static_cast<float>(dbeta);
  // train-mode BN removes any per-channel shift: d/d(bias) is exactly 0; eval: sum_m dy = s_c*dbeta
  if (grad_bias) grad_bias[c] = train ? 0.f : static_cast<float>(sc * dbeta);
}

// top of a dense (pool == 0) stack: dz[m,c] = mask * grad_out[b,c,n]; sums of dz and dz*yhat
__global__ void dense_top_kernel(const float* __restrict__ grad_out, const float* __restrict__ y,
                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                 int N, int C, float* __restrict__ dz, double* __restrict__ sums) {
  pdl_entry();
  __shared__ float tile[32][33];
  __shared__ float ssum[8 * 32], ssq[8 * 32];   // launched with blockDim (32, 8)
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && n < N) ? grad_out[(static_cast<int64_t>(b) * C + c) * N + n] : 0.f;
  }
  __syncthreads();
  float s = 0.f, q = 0.f;
  const int c = c0 + threadIdx.x;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r;
    if (c < C && n < N) {
      const int64_t m = static_cast<int64_t>(b) * N + n;
      float v = tile[threadIdx.x][r];
      const float yh = (y[m * C + c] - mean[c]) * invstd[c];
      if (relu && !(fmaf(yh, gamma[c], beta[c]) > 0.f)) v = 0.f;
      dz[m * C + c] = v;
      s += v;
      q = fmaf(v, yh, q);
    }
  }
  ssum[threadIdx.y * 32 + threadIdx.x] = s;
  ssq[threadIdx.y * 32 + threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    double sd = 0.0, qd = 0.0;
    for (int g = 0; g < static_cast<int>(blockDim.y); ++g) { sd += static_cast<double>(ssum[g * 32 + threadIdx.x]); qd += static_cast<double>(ssq[g * 32 + threadIdx.x]); }
    atomicAdd(&sums[c], sd);
    atomicAdd(&sums[C + c], qd);
  }
}

// ---- host orchestration --------------------------------------------------------------------------
inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + off);
    off = align_up(off + n * sizeof(T));
    return p;
  }
};

int splits_for(int64_t M) {
  int64_t s = (M + 127) / 128;
  const int64_t cap = 2 * static_cast<int64_t>(sm_count());
  if (s > cap) s = cap;
  return static_cast<int>(s < 1 ? 1 : s);
}

struct Shape {
  int B, N, L, pool;
  int64_t M;
  int sumC = 0, maxC = 0, Clast = 0, Kpool = 0;
  int64_t maxWW = 0;  // largest cout*cin among layers handled by the generic wgrad
  int S;              // upper bound of point-range splits of any contraction over points
};

Shape make_shape(int B, int N, int L, const pcuda_mlp_layer_t* layers, int pool) {
  Shape s{B, N, L, pool, static_cast<int64_t>(B) * N};
  for (int l = 0; l < L; ++l) {
    s.sumC += layers[l].cout;
    // dz buffers exist only for layers whose gradient is dense (not the pooled one)
    if (!(pool && l == L - 1)) s.maxC = std::max(s.maxC, layers[l].cout);
    if (!(pool && l == L - 1)) s.maxWW = std::max<int64_t>(s.maxWW, static_cast<int64_t>(layers[l].cout) * layers[l].cin);
  }
  s.Clast = layers[L - 1].cout;
  s.Kpool = layers[L - 1].cin;
  if (pool) s.maxWW = std::max<int64_t>(s.maxWW, static_cast<int64_t>(s.Kpool) * s.Kpool);
  s.S = std::max(splits_for(s.M), sm_count());
  return s;
}

// Workspace layouts are produced by ONE routine per direction, used both to size the workspace
// (base == nullptr) and to carve it.
struct FwdWs {
  double* stats; float* zeros; unsigned long long* keys; size_t zero_span; float* pivots;
  uint8_t* apack[16]; uint8_t* wpack[16];
  size_t bytes;
};

FwdWs carve_fwd(const Shape& s, const pcuda_mlp_layer_t* layers, void* base) {
  FwdWs w{};
  Carver cv(base);
  w.stats = cv.take<double>(2 * static_cast<size_t>(s.sumC));
  w.zeros = cv.take<float>(static_cast<size_t>(s.sumC) + 4);   // pivot of bias-free layers
  w.keys = s.pool ? cv.take<unsigned long long>(static_cast<size_t>(s.B) * s.Clast) : nullptr;
  w.zero_span = cv.off;                                          // stats, zeros and keys: one memset
  w.pivots = cv.take<float>(static_cast<size_t>(s.sumC) + 4);
  const tc::Tiling tl = tc::make_tiling(s.B, s.N);
  for (int l = 0; l < s.L; ++l) {
    if (!tc::supports(layers[l].cin)) continue;
    w.apack[l] = cv.take<uint8_t>(tc::act_pack_bytes(tl, layers[l].cin));
    w.wpack[l] = cv.take<uint8_t>(tc::w_pack_bytes(layers[l].cout, layers[l].cin));
  }
  w.bytes = cv.off;
  return w;
}

struct BwdWs {
  double* sums; float* alpha; float* kappa; float* partial; float* qpartial; float* upartial;
  float *coef, *gsel, *gyh, *Ghat, *Q, *abar, *u;
  double *colsum, *G, *abar_d;
  float* dzbuf[2];
  float* dxp;           // [B, C0, N] gradient w.r.t. the transformed input cloud (fused input transform, C0 <= 4)
  uint8_t* apack[17];   // apack[l + 1] = packed activation a_l; index 0 = the raw network input
  uint8_t* dypack; uint8_t* wtpack; uint8_t* qpack;
  size_t bytes;
};

BwdWs carve_bwd(const Shape& s, const pcuda_mlp_layer_t* layers, void* base) {
  BwdWs w{};
  Carver cv(base);
  w.sums = cv.take<double>(2 * static_cast<size_t>(s.sumC));
  w.alpha = cv.take<float>(s.sumC);
  w.kappa = cv.take<float>(s.sumC);
  w.partial = cv.take<float>(static_cast<size_t>(s.S) * s.maxWW);
  const tc::Tiling tl = tc::make_tiling(s.B, s.N);
  if (s.pool) {
    const size_t K = s.Kpool;
    w.coef = cv.take<float>(static_cast<size_t>(s.B) * s.Clast);
    w.gsel = cv.take<float>(static_cast<size_t>(s.B) * s.Clast);
    w.gyh = cv.take<float>(static_cast<size_t>(s.B) * s.Clast);
    w.colsum = cv.take<double>(s.S * K);
    w.G = cv.take<double>(K * K);
    w.Ghat = cv.take<float>(K * K);
    w.Q = cv.take<float>(K * K);
    w.abar = cv.take<float>(K);
    w.u = cv.take<float>(K);
    w.abar_d = cv.take<double>(K);
    w.qpartial = cv.take<float>(static_cast<size_t>(kPoolQSplits) * K * K);   // pool_q runs beside the Gram chain
    w.upartial = cv.take<float>(static_cast<size_t>(kPoolQSplits) * K);
    if (tc::supports(static_cast<int>(K))) w.qpack = cv.take<uint8_t>(tc::w_pack_bytes(static_cast<int>(K), static_cast<int>(K)));
  }
  w.dzbuf[0] = cv.take<float>(static_cast<size_t>(s.M) * s.maxC);
  w.dzbuf[1] = cv.take<float>(static_cast<size_t>(s.M) * s.maxC);
  w.dxp = layers[0].cin <= 4 ? cv.take<float>(static_cast<size_t>(s.M) * layers[0].cin) : nullptr;
  size_t dy_max = 0, wt_max = 0;
  for (int l = -1; l <= s.L - 2; ++l) {
    const int C = l < 0 ? layers[0].cin : layers[l].cout;
    if (tc::supports(C)) w.apack[l + 1] = cv.take<uint8_t>(tc::act_pack_bytes(tl, C));
  }
  for (int l = 0; l < s.L; ++l) {
    if (s.pool && l == s.L - 1) continue;
    if (tc::supports(layers[l].cout)) {
      dy_max = std::max(dy_max, tc::act_pack_bytes(tl, layers[l].cout));
      wt_max = std::max(wt_max, tc::w_pack_bytes(layers[l].cin, layers[l].cout));
    }
  }
  if (dy_max) w.dypack = cv.take<uint8_t>(dy_max);
  if (wt_max) w.wtpack = cv.take<uint8_t>(wt_max);
  w.bytes = cv.off;
  return w;
}


// ---- fork / join inside one library call -----------------------------------------------------------
// The pooled layer's backward is a chain of ~14 small launches of which several are independent (the
// Gram-matrix side: column sums, Gram, centring; the selection side: coefficients, Q = W^T K W; the weight
// gradient).  At the reference's sizes each launch is latency (~3-10 us), so the independent sides run on
// an auxiliary stream tied to the caller's stream with events — ordinary stream semantics, and under
// CUDA-graph capture the event edges become graph dependencies (the auxiliary stream joins the capture).
// One auxiliary stream + 3 events per caller stream, created on first use (never freed: a handful per
// process).  Creation is an ordinary runtime call, legal during capture.
struct Aux {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, side = nullptr, done = nullptr, wfork = nullptr, wdone = nullptr;
};
static std::mutex g_aux_mu;
static std::map<std::pair<int, cudaStream_t>, Aux> g_aux;

static std::map<int, std::vector<Aux>> g_aux_pool;     // per device: spare (stream, events) sets, created outside of any capture

static bool make_aux(Aux& a) {
  if (cudaStreamCreateWithFlags(&a.s, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return false; }
  if (cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&a.side, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&a.wfork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&a.wdone, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return false; }
  return true;
}

// The auxiliary set of caller stream `st` (nullptr: none available -> the caller stays on one stream).
// Streams are only ever CREATED while `st` is not capturing (a pool of spares is filled then); a caller
// stream first seen during a capture takes a spare.
static Aux* aux_for(cudaStream_t st) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  const bool capturing = cs != cudaStreamCaptureStatusNone;
  std::lock_guard<std::mutex> lk(g_aux_mu);
  std::vector<Aux>& pool = g_aux_pool[dev];
  if (!capturing) {
    while (pool.size() < 8) {
      Aux a;
      if (!make_aux(a)) break;
      pool.push_back(a);
    }
  }
  auto key = std::make_pair(dev, st);
  auto it = g_aux.find(key);
  if (it != g_aux.end()) return &it->second;
  if (pool.empty()) return nullptr;
  Aux a = pool.back();
  pool.pop_back();
  return &g_aux.emplace(key, a).first->second;
}

// Joins the auxiliary stream back into the caller's stream on EVERY exit path once a fork has been issued: an early
// error return would otherwise leave aux-stream kernels running on a workspace the caller is about to free (eager
// mode), or an unjoined forked stream inside a capture (which invalidates the capture with an unrelated error).
struct AuxJoin {
  Aux* ax;
  cudaStream_t st;
  bool pending = false;
  AuxJoin(Aux* a, cudaStream_t s) : ax(a), st(s) {}
  void join() {
    if (ax != nullptr && pending) {
      cudaEventRecord(ax->done, ax->s);
      cudaStreamWaitEvent(st, ax->done, 0);
    }
    pending = false;
  }
  ~AuxJoin() { join(); }
};

int validate(const char* who, int B, int N, int L, const pcuda_mlp_layer_t* layers, int pool) {
  PCUDA_REQUIRE(layers != nullptr, PCUDA_E_NULL, "%s: layers is NULL", who);
  PCUDA_REQUIRE(B >= 1 && N >= 1 && L >= 1 && L <= 16, PCUDA_E_SHAPE, "%s: bad shape B=%d N=%d L=%d", who, B, N, L);
  PCUDA_REQUIRE(static_cast<int64_t>(B) * N < (1ll << 31) / 4, PCUDA_E_UNSUPPORTED, "%s: B*N too large", who);
  for (int l = 0; l < L; ++l) {
    const pcuda_mlp_layer_t& y = layers[l];
    PCUDA_REQUIRE(y.cin >= 1 && y.cout >= 1, PCUDA_E_SHAPE, "%s: layer %d has cin=%d cout=%d", who, l, y.cin, y.cout);
    PCUDA_REQUIRE(l == 0 || y.cin == layers[l - 1].cout, PCUDA_E_SHAPE, "%s: layer %d cin=%d != previous cout=%d", who, l, y.cin, layers[l - 1].cout);
    PCUDA_REQUIRE(y.weight && y.gamma && y.beta && y.save_mean && y.save_invstd, PCUDA_E_NULL, "%s: layer %d has a NULL parameter", who, l);
    PCUDA_REQUIRE((y.cout & 3) == 0, PCUDA_E_UNSUPPORTED, "%s: layer %d cout=%d must be a multiple of 4", who, l, y.cout);
    PCUDA_REQUIRE(y.y != nullptr || (pool && l == L - 1), PCUDA_E_NULL, "%s: layer %d needs a y buffer", who, l);
    PCUDA_REQUIRE(aligned16(y.gamma) && aligned16(y.beta) && aligned16(y.save_mean) && aligned16(y.save_invstd) &&
                      aligned16(y.y) && ((y.cin & 3) != 0 || aligned16(y.weight)),
                  PCUDA_E_ALIGN, "%s: layer %d tensors must be 16-byte aligned", who, l);
  }
  return 0;
}

ActSrc input_src(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int N, int C0, const float* trans = nullptr) {
  return ActSrc{nullptr, nullptr, nullptr, nullptr, nullptr, 0, x, sxb, sxc, sxn, N, C0, trans};
}
// the input transform is applied inside the first layer's kernels: only where those are the narrow ones
int validate_trans(const char* who, const float* trans, int L, const pcuda_mlp_layer_t* layers, int pool) {
  if (trans == nullptr) return 0;
  PCUDA_REQUIRE(layers[0].cin <= 4, PCUDA_E_UNSUPPORTED, "%s: the fused input transform needs cin <= 4 (got %d); use pcuda_point_transform_fwd", who, layers[0].cin);
  PCUDA_REQUIRE(narrow_ok(layers[0].cin, layers[0].cout) && layers[0].cout <= 128 && !(pool && L == 1) && layers[0].y != nullptr, PCUDA_E_UNSUPPORTED,
                "%s: the fused input transform needs a first layer of <= 128 channels (power of two x 4) that is not the pooled one", who);
  return 0;
}
ActSrc layer_src(const pcuda_mlp_layer_t& y, int N) {
  return ActSrc{y.y, y.save_mean, y.save_invstd, y.gamma, y.beta, y.relu, nullptr, 0, 0, 0, N, y.cout, nullptr};
}

// Which pieces run on the tensor cores.  TUNE_MLP_TC_MASK bits switch single pieces back to the
// FP32 kernels (A/B runs, fault isolation): 1 forward, 2 pooled dgrad, 4 dense dgrad, 8 wgrad, 16 Gram.
enum TcPiece { TC_FWD = 1, TC_POOL_DGRAD = 2, TC_DGRAD = 4, TC_WGRAD = 8, TC_GRAM = 16 };
bool tc_on(int precision, int piece) {
  if (precision != PCUDA_MLP_BF16 || tuning(TUNE_MLP_FORCE_FP32)) return false;
  return (tuning(TUNE_MLP_TC_MASK) & piece) == 0;
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" size_t pcuda_pointmlp_ws_bytes(int B, int N, int L, const pcuda_mlp_layer_t* layers, int pool,
                                          int backward) {
  if (B < 1 || N < 1 || L < 1 || !layers) return 0;
  const Shape s = make_shape(B, N, L, layers, pool);
  return backward ? carve_bwd(s, layers, nullptr).bytes : carve_fwd(s, layers, nullptr).bytes;
}

static int pointmlp_fwd_impl(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N, int L,
                             const pcuda_mlp_layer_t* layers, int pool, int train, float momentum,
                             float eps, int precision, float* out, int32_t* pool_arg, void* ws,
                             const float* in_trans, pcuda_comm_t* sync_bn, pcuda_stream_t stream) {
  if (int rc = validate("pointmlp_fwd", B, N, L, layers, pool)) return rc;
  if (int rc = validate_trans("pointmlp_fwd", in_trans, L, layers, pool)) return rc;
  PCUDA_REQUIRE(x && out && ws, PCUDA_E_NULL, "pointmlp_fwd: NULL x/out/ws");
  PCUDA_REQUIRE(!pool || pool_arg, PCUDA_E_NULL, "pointmlp_fwd: pool needs pool_arg");
  PCUDA_REQUIRE(precision == PCUDA_MLP_FP32 || precision == PCUDA_MLP_BF16, PCUDA_E_UNSUPPORTED, "pointmlp_fwd: precision %d", precision);
  const Shape s = make_shape(B, N, L, layers, pool);
  PCUDA_REQUIRE(!train || s.M > 1, PCUDA_E_SHAPE, "pointmlp_fwd: train-mode BatchNorm needs more than 1 value per channel");
  if (!train)
    for (int l = 0; l < L; ++l)
      PCUDA_REQUIRE(layers[l].running_mean && layers[l].running_var, PCUDA_E_NULL, "pointmlp_fwd: eval mode needs running stats (layer %d)", l);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const FwdWs w = carve_fwd(s, layers, ws);
  cudaMemsetAsync(w.stats, 0, w.zero_span, st);
  // cross-rank BatchNorm statistics (train mode, more than one rank): see include/pcuda.h
  pcuda_comm_t* const sync = (train && sync_bn != nullptr && comm_world(sync_bn) > 1) ? sync_bn : nullptr;
  const double count_bn = static_cast<double>(s.M) * (sync ? comm_world(sync) : 1);

  const bool tcf = tc_on(precision, TC_FWD);
  auto layer_tc = [&](int l) { return tcf && tc::fwd_fits(layers[l].cout, layers[l].cin); };
  const tc::Tiling tl = tc::make_tiling(B, N);
  const int tps = (N + TM - 1) / TM;
  double* st_l = w.stats;
  float* piv_l = w.pivots;
  int launches = 0;
  // bf16 copies of the tensor-core layers' weights: independent of the activations, so they are packed on the
  // auxiliary stream (see Aux below) while the first layer runs; joined before the first tensor-core GEMM
  bool any_pack = false;
  for (int l = 0; l < L; ++l) any_pack = any_pack || layer_tc(l);
  Aux* ax = (any_pack && !tuning(TUNE_MLP_NO_FORK)) ? aux_for(st) : nullptr;
  AuxJoin joiner(ax, st);
  bool packs_pending = false;
  int sync_launches = 0;
  if (any_pack) {
    cudaStream_t sp = ax ? ax->s : st;
    if (ax) { cudaEventRecord(ax->fork, st); cudaStreamWaitEvent(sp, ax->fork, 0); joiner.pending = true; }
    for (int l = 0; l < L; ++l)
      if (layer_tc(l))
        if (int rc = tc::pack_w(layers[l].weight, layers[l].cout, layers[l].cin, false, w.wpack[l], sp,
                                (pool && l == L - 1) ? layers[l].gamma : nullptr)) return rc;
    if (ax) { cudaEventRecord(ax->done, sp); packs_pending = true; }
  }
  auto layer_narrow = [&](int l) {
    return !layer_tc(l) && !(pool && l == L - 1) && layers[l].y != nullptr && narrow_ok(layers[l].cin, layers[l].cout);
  };
  if (!layer_tc(0) && !layer_narrow(0)) {
    // pivot of layer 0 (no BN to finalise yet; the narrow-layer kernel publishes its own pivot)
    PCUDA_LAUNCH(bn_finalize_pivot_kernel, 1, 1024, 0, st, nullptr, nullptr, 0, 1.0, eps, momentum, train, nullptr, nullptr, nullptr, nullptr,
                                                 input_src(x, sxb, sxc, sxn, N, layers[0].cin, in_trans), layers[0].weight, layers[0].bias,
                                                 layers[0].cout, piv_l);
    launches += 1;
  }
  BnRaw raw_prev{};     // raw sums of the previous layer when its BatchNorm is finalised by this layer's operand packer
  for (int l = 0; l < L; ++l) {
    const pcuda_mlp_layer_t& y = layers[l];
    const ActSrc src = l == 0 ? input_src(x, sxb, sxc, sxn, N, y.cin, in_trans) : layer_src(layers[l - 1], N);
    const bool is_pool = pool && l == L - 1;
    const float* pivot_used = piv_l;
    if (layer_tc(l)) {
      // tensor-core layer: one pass packs a_{l-1} = relu?(bn(y_{l-1})) as bf16 slabs, the GEMM streams them.
      // Statistics are accumulated bias-free, i.e. centred on pivot = bias.
      if (int rc = tc::pack_act(src, tl, w.apack[l], st, raw_prev.stats ? &raw_prev : nullptr)) return rc;
      raw_prev = BnRaw{};
      if (packs_pending) { cudaStreamWaitEvent(st, ax->done, 0); packs_pending = false; joiner.pending = false; }
      if (int rc = tc::fwd_layer(tl, w.apack[l], w.wpack[l], y, is_pool, st_l, w.keys, st)) return rc;
      pivot_used = y.bias ? y.bias : w.zeros;
    } else if (layer_narrow(l)) {
      if (y.cin <= 4) PCUDA_LAUNCH_PDL(mlp_fwd_narrow_kernel<4>, narrow_grid(s.M, y.cout), kNarrowThreads, 0, st, src, y.weight, y.bias, y.cout, s.M, y.y, st_l, piv_l);
      else PCUDA_LAUNCH_PDL(mlp_fwd_narrow_kernel<8>, narrow_grid(s.M, y.cout), kNarrowThreads, 0, st, src, y.weight, y.bias, y.cout, s.M, y.y, st_l, piv_l);
      launches += 1;
    } else {
      const dim3 grid(B * tps, (y.cout + TN - 1) / TN);
      if (is_pool)
        PCUDA_LAUNCH(mlp_fwd_kernel<true>, grid, kThreads, 0, st, src, y.weight, y.bias, y.cout, N, tps, y.y, st_l, piv_l, y.gamma, w.keys);
      else
        PCUDA_LAUNCH(mlp_fwd_kernel<false>, grid, kThreads, 0, st, src, y.weight, y.bias, y.cout, N, tps, y.y, st_l, piv_l, y.gamma, nullptr);
      launches += 1;
    }
    // finalise this layer's BN.  When the only consumer of its output is the operand packer of a tensor-core
    // layer, or the pooled-output kernel, that kernel finalises from the raw sums itself (BnRaw): no one-block
    // launch between the GEMM and its consumer.  Otherwise: the finalise kernel; if an FP32 layer follows, it
    // computes that layer's pivot in the same launch.
    if (sync) {
      // this rank's sums, re-expressed around a pivot all ranks share (tensor-core layers already use it), summed over the ranks
      const float* common = y.bias ? y.bias : w.zeros;
      if (pivot_used != common) {
        PCUDA_LAUNCH(stats_rebase_kernel, (y.cout + 127) / 128, 128, 0, st, st_l, pivot_used, y.bias, y.cout, static_cast<double>(s.M));
        sync_launches += 1;
        pivot_used = common;
      }
      if (int rc = comm_sum_f64(sync, st_l, 2 * static_cast<int64_t>(y.cout), st)) return rc;
    }
    const bool next_piv = l + 1 < L && !layer_tc(l + 1) && !layer_narrow(l + 1);
    const bool fold = !tuning(TUNE_MLP_NO_FORK) && ((l + 1 < L && layer_tc(l + 1)) || (l == L - 1 && pool));
    if (fold) {
      raw_prev = BnRaw{st_l, pivot_used, count_bn, eps, momentum, train, y.save_mean, y.save_invstd,
                       y.running_mean, y.running_var};
    } else {
      PCUDA_LAUNCH(bn_finalize_pivot_kernel, 1, 1024, 0, st, st_l, pivot_used, y.cout, count_bn, eps, momentum, train,
                                                   y.save_mean, y.save_invstd, y.running_mean, y.running_var,
                                                   next_piv ? layer_src(y, N) : ActSrc{}, next_piv ? layers[l + 1].weight : nullptr,
                                                   next_piv ? layers[l + 1].bias : nullptr, next_piv ? layers[l + 1].cout : 0,
                                                   piv_l + y.cout);
      launches += 1;
    }
    st_l += 2 * y.cout;
    piv_l += y.cout;
  }
  const pcuda_mlp_layer_t& last = layers[L - 1];
  if (pool) {
    const int64_t n = static_cast<int64_t>(B) * s.Clast;
    PCUDA_LAUNCH(pool_finalize_kernel, static_cast<int>((n + 255) / 256), 256, 0, st, w.keys, last.save_mean, last.save_invstd, last.gamma,
                                                                            last.beta, last.relu, B, s.Clast, out, pool_arg, raw_prev);
  } else {
    const dim3 grid((N + 31) / 32, (s.Clast + 31) / 32, B);
    PCUDA_LAUNCH(dense_out_kernel, grid, dim3(32, 8), 0, st, layer_src(last, N), B, out);
  }
  count_launch(launches + 1 + sync_launches);
  return check_launch("pointmlp_fwd");
}

extern "C" int pcuda_pointmlp_fwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N, int L,
                                  const pcuda_mlp_layer_t* layers, int pool, int train, float momentum,
                                  float eps, int precision, float* out, int32_t* pool_arg, void* ws,
                                  pcuda_stream_t stream) {
  return pointmlp_fwd_impl(x, sxb, sxc, sxn, B, N, L, layers, pool, train, momentum, eps, precision, out, pool_arg, ws, nullptr, nullptr, stream);
}

extern "C" int pcuda_pointmlp_fwd_xf(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* in_trans, int B, int N, int L,
                                     const pcuda_mlp_layer_t* layers, int pool, int train, float momentum,
                                     float eps, int precision, float* out, int32_t* pool_arg, void* ws,
                                     pcuda_comm_t* sync_bn, pcuda_stream_t stream) {
  return pointmlp_fwd_impl(x, sxb, sxc, sxn, B, N, L, layers, pool, train, momentum, eps, precision, out, pool_arg, ws, in_trans, sync_bn,
                           stream);
}

static int pointmlp_bwd_impl(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N, int L,
                             const pcuda_mlp_layer_t* layers, int pool, int train, float eps,
                             int precision, const float* out, const int32_t* pool_arg,
                             const float* grad_out, float* grad_x, void* ws, const void* fwd_ws,
                             const float* in_trans, float* grad_trans, pcuda_comm_t* sync_bn, pcuda_stream_t stream) {
  (void)eps;
  if (int rc = validate("pointmlp_bwd", B, N, L, layers, pool)) return rc;
  if (int rc = validate_trans("pointmlp_bwd", in_trans, L, layers, pool)) return rc;
  PCUDA_REQUIRE(grad_trans == nullptr || in_trans != nullptr, PCUDA_E_NULL, "pointmlp_bwd: grad_trans without in_trans");
  PCUDA_REQUIRE(x && grad_out && ws, PCUDA_E_NULL, "pointmlp_bwd: NULL x/grad_out/ws");
  PCUDA_REQUIRE(!pool || (pool_arg && out), PCUDA_E_NULL, "pointmlp_bwd: pool needs out and pool_arg");
  PCUDA_REQUIRE(precision == PCUDA_MLP_FP32 || precision == PCUDA_MLP_BF16, PCUDA_E_UNSUPPORTED, "pointmlp_bwd: precision %d", precision);
  const Shape s = make_shape(B, N, L, layers, pool);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // cross-rank BatchNorm (include/pcuda.h): `count` is the number of values a statistic was taken over (all ranks),
  // count_local this rank's share
  pcuda_comm_t* const sync = (train && sync_bn != nullptr && comm_world(sync_bn) > 1) ? sync_bn : nullptr;
  const double count_local = static_cast<double>(s.M);
  const double count = count_local * (sync ? comm_world(sync) : 1);
  const int tps = (N + TM - 1) / TM;
  const tc::Tiling tl = tc::make_tiling(B, N);
  int launches = 0;

  const BwdWs w = carve_bwd(s, layers, ws);
  double* sums = w.sums;
  float *alpha = w.alpha, *kappa = w.kappa, *partial = w.partial;
  float* const* dzbuf = w.dzbuf;
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * s.sumC, st);

  std::vector<int> off(L + 1, 0);
  for (int l = 0; l < L; ++l) off[l + 1] = off[l] + layers[l].cout;
  const int S32 = splits_for(s.M);                       // FP32 contraction kernels
  const int64_t chunk32 = (s.M + S32 - 1) / S32;
  auto src_of = [&](int l) { return l < 0 ? input_src(x, sxb, sxc, sxn, N, layers[0].cin, in_trans) : layer_src(layers[l], N); };
  // with an input transform the first layer's dgrad lands in the workspace (gradient w.r.t. the TRANSFORMED cloud);
  // input_transform_bwd_kernel turns it into grad_x and grad_trans
  float* const gx_target = in_trans != nullptr ? w.dxp : grad_x;
  const bool want_gx = grad_x != nullptr || (in_trans != nullptr && grad_trans != nullptr);
  auto dgrad_out = [&](int lp, float* dzp) {  // epilogue target: previous layer lp (or input if lp < 0)
    DgradOut o{};
    if (lp < 0) {
      o.grad_x = gx_target;
      o.Cp = layers[0].cin;
    } else {
      const pcuda_mlp_layer_t& p = layers[lp];
      o = DgradOut{dzp, p.y, p.save_mean, p.save_invstd, p.gamma, p.beta, p.relu, sums + 2 * off[lp], nullptr, p.cout};
    }
    return o;
  };

  // bf16 copies of the activations the tensor-core pieces stream: a_l for l = -1 (input) .. L-2
  const bool any_tc = precision == PCUDA_MLP_BF16 && !tuning(TUNE_MLP_FORCE_FP32) && (tuning(TUNE_MLP_TC_MASK) & 30) != 30;
  // The forward call packed a_{j-1} for every layer j it ran on the tensor cores; with its workspace at hand
  // (fwd_ws, contents intact) those slabs are read in place instead of being packed a second time.
  const uint8_t* apack[17] = {};
  if (any_tc) {
    const bool tcf = tc_on(precision, TC_FWD);
    const FwdWs fw = fwd_ws != nullptr ? carve_fwd(s, layers, const_cast<void*>(fwd_ws)) : FwdWs{};
    for (int l = -1; l <= L - 2; ++l)
      if (w.apack[l + 1] != nullptr && (l >= 0 || layers[0].grad_weight != nullptr)) {
        const int j = l + 1;     // a_l is the operand of layer j
        if (fwd_ws != nullptr && tcf && fw.apack[j] != nullptr && tc::fwd_fits(layers[j].cout, layers[j].cin)) {
          apack[l + 1] = fw.apack[j];
          continue;
        }
        if (int rc = tc::pack_act(src_of(l), tl, w.apack[l + 1], st)) return rc;
        apack[l + 1] = w.apack[l + 1];
      }
  }

  // auxiliary stream of this call (see Aux above); tuning key 6 keeps everything on the caller's stream (A/B)
  // (cross-rank mode: one stream, so that the collectives of one communicator are issued in one order)
  Aux* ax = (tuning(TUNE_MLP_NO_FORK) || sync) ? nullptr : aux_for(st);
  cudaStream_t sa = ax ? ax->s : st;
  AuxJoin joiner(ax, st);  // joins on every exit path once joiner.pending is set
  bool& aux_used = joiner.pending;   // something was issued on the auxiliary stream: join before returning
  bool wgrad_pending = false;   // a weight-gradient chain of the previous layer may still be running there
  int cur = 0;      // dzbuf[cur] holds dz of layer `top`
  int top = L - 1;  // highest layer whose dz is dense and stored
  const pcuda_mlp_layer_t& last = layers[L - 1];
  if (pool) {
    const int C = last.cout, K = last.cin;
    const ActSrc src = src_of(L - 2);
    const uint8_t* a_in = apack[L - 1];
    float* al = alpha + off[L - 1];
    float* ka = kappa + off[L - 1];
    const int64_t bc = static_cast<int64_t>(B) * C;
    const int64_t kk = static_cast<int64_t>(K) * K;
    bool q_packed = false;   // Q already sits in w.qpack (tc::q_finish)
    // selection side on the caller's stream, Gram side on the auxiliary stream
    if (ax) { cudaEventRecord(ax->fork, st); cudaStreamWaitEvent(sa, ax->fork, 0); aux_used = true; }
    // -- Gram side: centred Gram matrix of the pooled layer's input (train-mode BN correction terms only)
    int Sg = S32;
    PCUDA_LAUNCH(act_colsum_kernel, dim3(S32, (K + 31) / 32), 256, 0, sa, src, s.M, chunk32, w.colsum);
    launches += 1;
    if (tc_on(precision, TC_GRAM) && a_in && tc::pt_supports(K, K, true)) {
      Sg = tc::pt_splits(tl, (K + 127) / 128);
      if (int rc = tc::gram(tl, a_in, K, Sg, partial, sa)) return rc;
    } else {
      PCUDA_LAUNCH(point_contract_kernel<1>, dim3(S32, (K + TM - 1) / TM, (K + TN - 1) / TN), kThreads, 0, sa, DySrc{}, src, src, s.M, chunk32, K, K, partial);
      launches += 1;
    }
    if (!sync) {
      PCUDA_LAUNCH(gram_finish_kernel, static_cast<int>((kk + 255) / 256), 256, sizeof(double) * K, sa, partial, Sg, w.colsum, S32, K, count,
                   w.abar, w.abar_d, w.Ghat);
    } else {
      PCUDA_LAUNCH(reduce_partials_kernel<double>, static_cast<int>((kk + 255) / 256), 256, 0, sa, partial, kk, Sg, w.G);
      // column means over ALL ranks' points; the Gram matrix stays local and is centred with the global mean
      // (sum_r [G_r - M_r abar abar^T] is the centred Gram matrix of the global batch)
      PCUDA_LAUNCH(abar_kernel, (K + 7) / 8, 256, 0, sa, w.colsum, S32, K, 1.0, w.abar, w.abar_d);
      if (int rc = comm_sum_f64(sync, w.abar_d, K, sa)) return rc;
      PCUDA_LAUNCH(abar_kernel, (K + 7) / 8, 256, 0, sa, w.abar_d, 1, K, count, w.abar, w.abar_d);
      PCUDA_LAUNCH(gram_center_kernel, static_cast<int>((kk + 255) / 256), 256, 0, sa, w.abar_d, w.G, K, count_local, w.Ghat);
      launches += 3;
    }
    if (ax) cudaEventRecord(ax->side, sa);
    // -- selection side: per-(cloud, channel) coefficients, alpha / kappa, Q = W^T diag(kappa) W
    PCUDA_LAUNCH_PDL(pool_sel_kernel, static_cast<int>((bc * 32 + 255) / 256), 256, 0, st, src, last.weight, last.bias, last.save_mean,
                                                                            last.save_invstd, last.gamma, last.relu, out, pool_arg,
                                                                            grad_out, B, N, C, w.coef, w.gsel, w.gyh);
    const bool want_last = last.grad_weight != nullptr;
    if (sync) {
      double* xs = sums + 2 * off[L - 1];      // the pooled layer's own slot of `sums` is otherwise unused
      PCUDA_LAUNCH(pool_coef_kernel, (C + 127) / 128, 128, 0, st, w.gsel, w.gyh, last.save_invstd, last.gamma, B, C, count, train, al, ka,
                   want_last ? last.grad_gamma : nullptr, want_last ? last.grad_beta : nullptr, want_last ? last.grad_bias : nullptr, xs, 1);
      if (int rc = comm_sum_f64(sync, xs, 2 * static_cast<int64_t>(C), st)) return rc;
      PCUDA_LAUNCH(pool_coef_kernel, (C + 127) / 128, 128, 0, st, w.gsel, w.gyh, last.save_invstd, last.gamma, B, C, count, train, al, ka,
                   static_cast<float*>(nullptr), static_cast<float*>(nullptr), static_cast<float*>(nullptr), xs, 2);
      launches += 1;
    } else {
      PCUDA_LAUNCH(pool_coef_kernel, (C + 127) / 128, 128, 0, st, w.gsel, w.gyh, last.save_invstd, last.gamma, B, C, count, train, al, ka,
                   want_last ? last.grad_gamma : nullptr, want_last ? last.grad_beta : nullptr, want_last ? last.grad_bias : nullptr,
                   static_cast<double*>(nullptr), 0);
    }
    // the low-rank dgrad runs on the tensor cores (decided here: it changes how Q is finished)
    const bool tc_pool = (L >= 2 || grad_x) && tc_on(precision, TC_POOL_DGRAD) && a_in && w.qpack && tc::pool_dgrad_fits(K) &&
                         pool_sparse_smem(C, N, K) <= 200 * 1024 && K <= 512 && C <= 65535;
    const bool fused_tail = tc_pool && !tuning(TUNE_MLP_NO_FORK);
    PCUDA_LAUNCH_PDL(pool_q_kernel, dim3((K + 31) / 32, (K + 31) / 32, kPoolQSplits), 256, 0, st, last.weight, ka, al, C, K, w.qpartial,
                     fused_tail ? w.upartial : static_cast<float*>(nullptr));
    if (fused_tail) {
      // -- join: u needs Q (this stream) and abar (Gram side); one launch sums the partials of Q (and of the alpha term
      // of u), packs Q as the stationary bf16 operand of the dgrad GEMM and forms u
      if (ax) cudaStreamWaitEvent(st, ax->side, 0);
      if (int rc = tc::q_finish(w.qpartial, w.upartial, kPoolQSplits, w.abar, K, w.Q, w.qpack, w.u, st)) return rc;
      q_packed = true;
      launches += 4;     // gram_finish, pool_sel, pool_coef, pool_q (tc:: calls count themselves; cross-rank: + 3 above)
    } else {
      PCUDA_LAUNCH(reduce_partials_kernel<float>, static_cast<int>((kk + 255) / 256), 256, 0, st, w.qpartial, kk, kPoolQSplits, w.Q);
      // -- join: u needs Q (this stream) and abar (Gram side)
      if (ax) cudaStreamWaitEvent(st, ax->side, 0);
      PCUDA_LAUNCH(pool_u_kernel, (K + 31) / 32, 1024, 0, st, last.weight, al, w.Q, w.abar, C, K, w.u);
      launches += 6;
    }
    if (last.grad_weight) {
      const size_t dw_smem = sizeof(float) * ((kDwCh * (K + 2 * static_cast<size_t>(B)) + 3) & ~static_cast<size_t>(3)) +
                             sizeof(double) * kDwParts * 128 * kDwCh;
      PCUDA_REQUIRE(dw_smem <= 200 * 1024, PCUDA_E_UNSUPPORTED, "pointmlp_bwd: batch %d too large for the pooled wgrad kernel", B);
      smem_optin(pool_dw_kernel, 200 * 1024);
      // the pooled layer's weight gradient reads coef / kappa (this stream) and abar / Ghat (Gram side) and writes
      // only grad_weight: it runs on the auxiliary stream beside the dgrad kernels below; joined before returning
      if (ax) { cudaEventRecord(ax->fork, st); cudaStreamWaitEvent(sa, ax->fork, 0); }
      PCUDA_LAUNCH_PDL(pool_dw_kernel, (C + kDwCh - 1) / kDwCh, 128 * kDwParts, dw_smem, sa, src, last.weight, w.coef, pool_arg, ka, w.abar, w.Ghat, B, N, C, train, last.grad_weight);
      launches += 1;
    }
    if (L >= 2 || grad_x) {
      const DgradOut o = dgrad_out(L - 2, dzbuf[cur]);
      const size_t sparse_smem = pool_sparse_smem(C, N, K);
      if (tc_on(precision, TC_POOL_DGRAD) && a_in && w.qpack && tc::pool_dgrad_fits(K) && sparse_smem <= 200 * 1024 && K <= 512 && C <= 65535) {
        smem_optin(pool_sparse_kernel<4, 12, 1>, 200 * 1024);
        smem_optin(pool_sparse_kernel<4, 4, 2>, 100 * 1024);
        smem_optin(pool_sparse_kernel<16, 2, 1>, 200 * 1024);
        // dense part on the tensor cores: dz_prev = mask * (-u - a Q) (+ its column sums) ...
        if (!q_packed)
          if (int rc = tc::pack_w(w.Q, K, K, false, w.qpack, st)) return rc;
        if (int rc = tc::pool_dgrad(tl, a_in, K, w.qpack, w.u, o, st)) return rc;
        // ... then the <= min(C, N) selected points per cloud get their sparse rows added (one warp per point).
        // Every CTA repeats the O(C + N) prologue and ends with 2K fp64 atomics on the same 2K addresses
        // (beyond ~300 CTAs those serialise: tools/micro/atomic_tail.cu), so about one wave of CTAs.
        const bool many_points = K <= 128 && N >= C && sparse_smem <= 100 * 1024;   // ~1 channel per selected point
        const int per_sm = many_points ? 2 : 1;
        const int parts = std::max(1, std::min((std::min(C, N) + 31) / 32, (per_sm * sm_count() + B - 1) / B));
        // few points, many channels each (N < C, K <= 128): the sorted, evenly cut variant (tuning key 10 < 0: off)
        const int sb_rows = (C + parts - 1) / parts + 2;
        const size_t sorted_smem = ss_layout(C, N, K, sb_rows).bytes;
        const bool rows16 = (K & 3) == 0 && ((reinterpret_cast<uintptr_t>(last.weight) | reinterpret_cast<uintptr_t>(o.y_prev) |
                                              reinterpret_cast<uintptr_t>(o.dz_prev) | reinterpret_cast<uintptr_t>(o.mean) |
                                              reinterpret_cast<uintptr_t>(o.invstd) | reinterpret_cast<uintptr_t>(o.gamma) |
                                              reinterpret_cast<uintptr_t>(o.beta)) & 15u) == 0;
        const bool sorted = !many_points && K <= 128 && N <= 1024 && o.grad_x == nullptr && rows16 &&
                            sorted_smem <= 200 * 1024 && tuning(TUNE_SPARSE_SORTED) >= 0;
        if (sorted) {
          smem_optin(pool_sparse_sorted_kernel, 200 * 1024);
          PCUDA_LAUNCH(pool_sparse_sorted_kernel, dim3(B, parts), 32 * kSsWarps, sorted_smem, st, pool_arg, w.coef, last.weight, C, N, K, sb_rows, o);
        } else if (many_points) PCUDA_LAUNCH((pool_sparse_kernel<4, 4, 2>), dim3(B, parts), 512, sparse_smem, st, pool_arg, w.coef, last.weight, C, N, K, o);
        else if (K <= 128) PCUDA_LAUNCH((pool_sparse_kernel<4, 12, 1>), dim3(B, parts), 512, sparse_smem, st, pool_arg, w.coef, last.weight, C, N, K, o);
        else PCUDA_LAUNCH((pool_sparse_kernel<16, 2, 1>), dim3(B, parts), 512, sparse_smem, st, pool_arg, w.coef, last.weight, C, N, K, o);
        launches += 1;
      } else {
        PCUDA_LAUNCH(pool_dgrad_kernel, dim3(B * tps, (K + TN - 1) / TN), kThreads, 0, st, src, w.Q, w.u, last.weight, w.coef, pool_arg, C, N, tps, o);
        launches += 1;
      }
    }
    top = L - 2;
  } else {
    const int C = last.cout;
    PCUDA_LAUNCH(dense_top_kernel, dim3((N + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, st, grad_out, last.y, last.save_mean, last.save_invstd,
                                                                                   last.gamma, last.beta, last.relu, N, C,
                                                                                   dzbuf[cur], sums + 2 * off[L - 1]);
    launches += 1;
  }

  for (int l = top; l >= 0; --l) {
    const pcuda_mlp_layer_t& y = layers[l];
    const int C = y.cout, Kp = y.cin;
    float* al = alpha + off[l];
    float* ka = kappa + off[l];
    const bool want_w = y.grad_weight != nullptr;
    if (sync) {
      // parameter gradients from this rank's sums (they are summed over the ranks by the gradient all-reduce) ...
      if (want_w) {
        PCUDA_LAUNCH(bn_bwd_coef_kernel, (C + 127) / 128, 128, 0, st, sums + 2 * off[l], y.save_invstd, y.gamma, C, count, train, al, ka,
                     y.grad_gamma, y.grad_beta, y.grad_bias);
        launches += 1;
      }
      // ... the BatchNorm backward coefficients from the sums over all ranks
      if (int rc = comm_sum_f64(sync, sums + 2 * off[l], 2 * static_cast<int64_t>(C), st)) return rc;
      PCUDA_LAUNCH(bn_bwd_coef_kernel, (C + 127) / 128, 128, 0, st, sums + 2 * off[l], y.save_invstd, y.gamma, C, count, train, al, ka,
                   static_cast<float*>(nullptr), static_cast<float*>(nullptr), static_cast<float*>(nullptr));
    } else {
      PCUDA_LAUNCH(bn_bwd_coef_kernel, (C + 127) / 128, 128, 0, st, sums + 2 * off[l], y.save_invstd, y.gamma, C, count, train, al, ka,
                   want_w ? y.grad_gamma : nullptr, want_w ? y.grad_beta : nullptr, want_w ? y.grad_bias : nullptr);
    }
    launches += 1;
    const DySrc dys{dzbuf[cur], y.y, y.save_mean, y.save_invstd, y.gamma, al, ka, C};
    const bool need_dgrad = l > 0 || want_gx;
    const bool wgrad_tc = want_w && tc_on(precision, TC_WGRAD) && w.dypack && apack[l] && tc::pt_supports(C, Kp, false);
    const bool dgrad_side = l - 1 >= 0 && (Kp % 64) == 0 && apack[l] != nullptr;
    const bool dgrad_tc = need_dgrad && tc_on(precision, TC_DGRAD) && w.dypack && w.wtpack && (Kp % 8) == 0 &&
                          tc::dgrad_fits(Kp, C, dgrad_side);
    // the previous layer's weight-gradient chain (auxiliary stream) read dypack / dzbuf[cur ^ 1]: both are
    // rewritten below, so it has to be over first — it ran beside that layer's dgrad, which took longer
    if (wgrad_pending) { cudaStreamWaitEvent(st, ax->wdone, 0); wgrad_pending = false; }
    if (wgrad_tc || dgrad_tc)
      if (int rc = tc::pack_dy(dys, tl, w.dypack, st)) return rc;      // dy_l as bf16 slabs, once for both GEMMs
    if (want_w) {
      // dW_l and da_{l-1} both start from dy_l and are independent: the weight gradient (contraction + split
      // reduction) runs on the auxiliary stream beside the dgrad GEMM
      const bool fork_w = ax != nullptr && need_dgrad;
      cudaStream_t sw = fork_w ? sa : st;
      if (fork_w) { cudaEventRecord(ax->wfork, st); cudaStreamWaitEvent(sa, ax->wfork, 0); aux_used = true; }
      const int64_t ck = static_cast<int64_t>(C) * Kp;
      int Sw = S32;
      if (wgrad_tc) {
        Sw = tc::pt_splits(tl, (C + 127) / 128);
        if (int rc = tc::wgrad_layer(tl, w.dypack, C, apack[l], Kp, Sw, partial, sw)) return rc;
      } else if (narrow_ok(Kp, C)) {
        Sw = std::min(narrow_grid(s.M, C), s.S);
        const size_t smem = sizeof(float) * 1024 * static_cast<size_t>(Kp);   // [pslots][C][Kp], pslots * C = 1024
        if (Kp <= 4) PCUDA_LAUNCH(wgrad_narrow_kernel<4>, Sw, kNarrowThreads, smem, sw, dys, src_of(l - 1), s.M, partial);
        else PCUDA_LAUNCH(wgrad_narrow_kernel<8>, Sw, kNarrowThreads, smem, sw, dys, src_of(l - 1), s.M, partial);
        launches += 1;
      } else {
        PCUDA_LAUNCH(point_contract_kernel<0>, dim3(S32, (C + TM - 1) / TM, (Kp + TN - 1) / TN), kThreads, 0, sw, dys, ActSrc{}, src_of(l - 1), s.M, chunk32, C, Kp, partial);
        launches += 1;
      }
      if (Sw >= 32 && ck <= 16384) PCUDA_LAUNCH(reduce_partials_split_kernel<float>, static_cast<int>((ck + 31) / 32), 256, 0, sw, partial, ck, Sw, y.grad_weight);
      else PCUDA_LAUNCH(reduce_partials_kernel<float>, static_cast<int>((ck + 255) / 256), 256, 0, sw, partial, ck, Sw, y.grad_weight);
      launches += 1;
      if (fork_w) { cudaEventRecord(ax->wdone, sa); wgrad_pending = true; }
    }
    if (need_dgrad) {
      const DgradOut o = dgrad_out(l - 1, dzbuf[cur ^ 1]);
      if (dgrad_tc) {
        // A[r = channel of layer l-1, k = channel of layer l] = W_l[k, r]
        if (int rc = tc::pack_w(y.weight, Kp, C, true, w.wtpack, st)) return rc;
        const uint8_t* side = dgrad_side ? apack[l] : nullptr;
        if (int rc = tc::dgrad_layer(tl, w.dypack, C, w.wtpack, side, o, st)) return rc;
      } else if (o.grad_x != nullptr && narrow_ok(Kp, C) && C <= 128) {
        if (Kp <= 4) PCUDA_LAUNCH_PDL(dgrad_input_narrow_kernel<4>, narrow_grid(s.M, C), kNarrowThreads, 0, st, dys, y.weight, Kp, N, s.M, o.grad_x);
        else PCUDA_LAUNCH_PDL(dgrad_input_narrow_kernel<8>, narrow_grid(s.M, C), kNarrowThreads, 0, st, dys, y.weight, Kp, N, s.M, o.grad_x);
        launches += 1;
      } else {
        PCUDA_LAUNCH(dense_dgrad_kernel, dim3(B * tps, (Kp + TN - 1) / TN), kThreads, 0, st, dys, y.weight, N, tps, o);
        launches += 1;
      }
      cur ^= 1;
    }
  }
  if (in_trans != nullptr && want_gx) {
    PCUDA_LAUNCH(input_transform_bwd_kernel, B, 256, 0, st, x, sxb, sxc, sxn, in_trans, w.dxp, layers[0].cin, N, grad_x, grad_trans);
    launches += 1;
  }
  joiner.join();    // the pooled weight gradient / the last weight-gradient chain
  count_launch(launches);
  return check_launch("pointmlp_bwd");
}

extern "C" int pcuda_pointmlp_bwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N, int L,
                                  const pcuda_mlp_layer_t* layers, int pool, int train, float eps,
                                  int precision, const float* out, const int32_t* pool_arg,
                                  const float* grad_out, float* grad_x, void* ws, pcuda_stream_t stream) {
  return pointmlp_bwd_impl(x, sxb, sxc, sxn, B, N, L, layers, pool, train, eps, precision, out, pool_arg, grad_out, grad_x, ws,
                           nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int pcuda_pointmlp_bwd_reuse(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N, int L,
                                        const pcuda_mlp_layer_t* layers, int pool, int train, float eps,
                                        int precision, const float* out, const int32_t* pool_arg,
                                        const float* grad_out, float* grad_x, void* ws, const void* fwd_ws,
                                        pcuda_stream_t stream) {
  return pointmlp_bwd_impl(x, sxb, sxc, sxn, B, N, L, layers, pool, train, eps, precision, out, pool_arg, grad_out, grad_x, ws,
                           fwd_ws, nullptr, nullptr, nullptr, stream);
}

extern "C" int pcuda_pointmlp_bwd_xf(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* in_trans, int B, int N, int L,
                                     const pcuda_mlp_layer_t* layers, int pool, int train, float eps,
                                     int precision, const float* out, const int32_t* pool_arg,
                                     const float* grad_out, float* grad_x, float* grad_trans, void* ws, const void* fwd_ws,
                                     pcuda_comm_t* sync_bn, pcuda_stream_t stream) {
  return pointmlp_bwd_impl(x, sxb, sxc, sxn, B, N, L, layers, pool, train, eps, precision, out, pool_arg, grad_out, grad_x, ws,
                           fwd_ws, in_trans, grad_trans, sync_bn, stream);
}
