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
#include <vector>

#include "pcuda_common.cuh"

namespace pcuda {
namespace {

constexpr int TM = 64;   // tile rows
constexpr int TN = 64;   // tile cols
constexpr int TK = 16;   // contraction chunk
constexpr int LD = 68;   // padded leading dimension of the smem tiles
constexpr int kThreads = 256;

// a_l[m, k] = relu?((y[m,k] - mean_k) * invstd_k * gamma_k + beta_k), or raw x for the network input
struct ActSrc {
  const float* y;  // [M, C] or nullptr -> raw input
  const float* mean;
  const float* invstd;
  const float* gamma;
  const float* beta;
  int relu;
  const float* x;  // raw input, x[b*sxb + c*sxc + n*sxn]
  int64_t sxb, sxc, sxn;
  int N;
  int C;
};

__device__ __forceinline__ float bn_act(float y, float mean, float invstd, float gamma, float beta, int relu) {
  const float z = fmaf(y - mean, invstd * gamma, beta);
  return relu ? fmaxf(z, 0.0f) : z;
}

// 4 consecutive channels k..k+3 of point m (zero beyond C). k % 4 == 0.
__device__ __forceinline__ float4 load_act4(const ActSrc& s, int64_t m, int k) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (s.y != nullptr) {
    if (k + 3 < s.C) {
      const float4 yy = *reinterpret_cast<const float4*>(s.y + m * s.C + k);
      const float4 mu = *reinterpret_cast<const float4*>(s.mean + k);
      const float4 is = *reinterpret_cast<const float4*>(s.invstd + k);
      const float4 ga = *reinterpret_cast<const float4*>(s.gamma + k);
      const float4 be = *reinterpret_cast<const float4*>(s.beta + k);
      v.x = bn_act(yy.x, mu.x, is.x, ga.x, be.x, s.relu);
      v.y = bn_act(yy.y, mu.y, is.y, ga.y, be.y, s.relu);
      v.z = bn_act(yy.z, mu.z, is.z, ga.z, be.z, s.relu);
      v.w = bn_act(yy.w, mu.w, is.w, ga.w, be.w, s.relu);
    } else {
      float* pv = &v.x;
      for (int i = 0; i < 4 && k + i < s.C; ++i)
        pv[i] = bn_act(s.y[m * s.C + k + i], s.mean[k + i], s.invstd[k + i], s.gamma[k + i], s.beta[k + i], s.relu);
    }
  } else {
    const int64_t b = m / s.N, n = m - b * s.N;
    const float* px = s.x + b * s.sxb + n * s.sxn;
    float* pv = &v.x;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (k + i < s.C) pv[i] = __ldg(px + (k + i) * s.sxc);
  }
  return v;
}

__device__ __forceinline__ float load_act1(const ActSrc& s, int64_t m, int k) {
  if (s.y != nullptr) return bn_act(s.y[m * s.C + k], s.mean[k], s.invstd[k], s.gamma[k], s.beta[k], s.relu);
  const int64_t b = m / s.N, n = m - b * s.N;
  return __ldg(s.x + b * s.sxb + n * s.sxn + k * s.sxc);
}

// dy_l[m, c] = s_c*dz[m,c] - alpha_c - kappa_c*(y[m,c] - mean_c)   (train-mode BN backward; in
// eval mode alpha = kappa = 0).  s_c = gamma_c * invstd_c.
struct DySrc {
  const float* dz;  // [M, C]
  const float* y;   // [M, C]
  const float* mean;
  const float* invstd;
  const float* gamma;
  const float* alpha;
  const float* kappa;
  int C;
};

__device__ __forceinline__ float4 load_dy4(const DySrc& s, int64_t m, int c) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  float* pv = &v.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cc = c + i;
    if (cc < s.C) {
      const float sc = s.gamma[cc] * s.invstd[cc];
      pv[i] = fmaf(sc, s.dz[m * s.C + cc], -s.alpha[cc]) - s.kappa[cc] * (s.y[m * s.C + cc] - s.mean[cc]);
    }
  }
  return v;
}

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

__device__ __forceinline__ unsigned long long pool_key(float v, int n) {
  return (static_cast<unsigned long long>(float_to_ordered(v)) << 32) |
         static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned int>(n));
}

// ---- forward layer -------------------------------------------------------------------------------
// grid (B * tiles_per_sample, ceil(Cout/64)).  Tiles never straddle two clouds.
template <bool POOL>
__global__ void __launch_bounds__(kThreads)
mlp_fwd_kernel(ActSrc src, const float* __restrict__ W, const float* __restrict__ bias, int Cout,
               int N, int tiles_per_sample, float* __restrict__ y_out, double* __restrict__ stats,
               const float* __restrict__ pivot, const float* __restrict__ gamma,
               unsigned long long* __restrict__ keys) {
  __shared__ Tiles t;
  __shared__ float ssum[TN], ssq[TN];
  __shared__ unsigned long long skey[TN];
  const int b = blockIdx.x / tiles_per_sample;
  const int n0 = (blockIdx.x - b * tiles_per_sample) * TM;
  const int c0 = blockIdx.y * TN;
  const int K = src.C;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (threadIdx.x < TN) { ssum[threadIdx.x] = 0.f; ssq[threadIdx.x] = 0.f; skey[threadIdx.x] = 0ull; }

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
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    atomicAdd(&ssum[tx * 4 + j], cs[j]);
    atomicAdd(&ssq[tx * 4 + j], cq[j]);
    if (POOL) atomicMax(&skey[tx * 4 + j], ck[j]);
  }
  __syncthreads();
  if (threadIdx.x < TN && c0 + threadIdx.x < Cout) {
    const int c = c0 + threadIdx.x;
    atomicAdd(&stats[c], static_cast<double>(ssum[threadIdx.x]));
    atomicAdd(&stats[Cout + c], static_cast<double>(ssq[threadIdx.x]));
    if (POOL) atomicMax(&keys[static_cast<int64_t>(b) * Cout + c], skey[threadIdx.x]);
  }
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
                                     float* __restrict__ out, int32_t* __restrict__ arg) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(B) * C) return;
  const int c = static_cast<int>(i % C);
  const unsigned long long key = keys[i];
  const float vv = ordered_to_float(static_cast<uint32_t>(key >> 32));
  const int n = static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull));
  const float g = gamma[c];
  const float y = g > 0.f ? vv : (g < 0.f ? -vv : mean[c]);
  out[i] = bn_act(y, mean[c], invstd[c], g, beta[c], relu);
  arg[i] = n;
}

// dense output (pool == 0): out[b, c, n] = act(y[(b,n), c]); 32x32 smem transpose
__global__ void dense_out_kernel(ActSrc src, int B, float* __restrict__ out) {
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
// One warp per (cloud, channel): masked upstream gradient and yhat at the selected point.
__global__ void pool_sel_kernel(ActSrc src, const float* __restrict__ W, const float* __restrict__ bias,
                                const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, int relu, const float* __restrict__ out,
                                const int32_t* __restrict__ arg, const float* __restrict__ grad_out, int B, int N,
                                int C, float* __restrict__ coef, float* __restrict__ gsel, float* __restrict__ gyh) {
  const int64_t w = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= static_cast<int64_t>(B) * C) return;
  const int c = static_cast<int>(w % C);
  const int64_t b = w / C;
  const int K = src.C;
  float g = grad_out[w];
  if (relu && !(out[w] > 0.f)) g = 0.f;
  float yhat = 0.f;
  if (g != 0.f) {
    const int64_t m = b * N + arg[w];
    float dot = 0.f;
    for (int k = lane; k < K; k += 32) dot = fmaf(W[static_cast<int64_t>(c) * K + k], load_act1(src, m, k), dot);
    dot = warp_sum(dot);
    yhat = (dot + (bias ? bias[c] : 0.f) - mean[c]) * invstd[c];
  }
  if (lane == 0) {
    coef[w] = gamma[c] * invstd[c] * g;
    gsel[w] = g;
    gyh[w] = g * yhat;
  }
}

// One thread per channel: fixed-order sums over the B clouds.
__global__ void pool_coef_kernel(const float* __restrict__ gsel, const float* __restrict__ gyh,
                                 const float* __restrict__ invstd, const float* __restrict__ gamma, int B, int C,
                                 double count, int train, float* __restrict__ alpha, float* __restrict__ kappa,
                                 float* __restrict__ grad_gamma, float* __restrict__ grad_beta,
                                 float* __restrict__ grad_bias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double dbeta = 0.0, dgamma = 0.0;
  for (int b = 0; b < B; ++b) {
    dbeta += static_cast<double>(gsel[static_cast<int64_t>(b) * C + c]);
    dgamma += static_cast<double>(gyh[static_cast<int64_t>(b) * C + c]);
  }
  const double sc = static_cast<double>(gamma[c]) * static_cast<double>(invstd[c]);
  if (grad_gamma) grad_gamma[c] = static_cast<float>(dgamma);
  if (grad_beta) grad_beta[c] = static_cast<float>(dbeta);
  if (grad_bias) grad_bias[c] = train ? 0.f : static_cast<float>(sc * dbeta);
  alpha[c] = train ? static_cast<float>(sc * dbeta / count) : 0.f;
  kappa[c] = train ? static_cast<float>(sc * static_cast<double>(invstd[c]) * dgamma / count) : 0.f;
}

// column sums of an activation: partial[s, k] over the s-th slice of points
__global__ void act_colsum_kernel(ActSrc src, int64_t M, int64_t chunk, double* __restrict__ partial) {
  const int k = blockIdx.y * blockDim.x + threadIdx.x;
  if (k >= src.C) return;
  const int64_t m0 = blockIdx.x * chunk;
  const int64_t m1 = m0 + chunk < M ? m0 + chunk : M;
  double s = 0.0;
  for (int64_t m = m0; m < m1; ++m) s += static_cast<double>(load_act1(src, m, k));
  partial[static_cast<int64_t>(blockIdx.x) * src.C + k] = s;
}

// OUT[s, r, c] = sum_{m in slice s} P[m, r] * R[m, c]   (contraction over points; 64x64 tile / CTA)
// MODE 0: P = dy (DySrc), R = act (wgrad).  MODE 1: P = R = act (Gram matrix).
template <int MODE>
__global__ void __launch_bounds__(kThreads)
point_contract_kernel(DySrc dys, ActSrc pact, ActSrc ract, int64_t M, int64_t chunk, int CR, int CC,
                      float* __restrict__ partial) {
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
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int k = 0; k < S; ++k) s += static_cast<double>(partial[static_cast<int64_t>(k) * n + i]);
  out[i] = static_cast<TOut>(s);
}

// abar[k] = (sum_s colsum_partial[s,k]) / M ;  Ghat[k,k'] = G[k,k'] - M abar_k abar_k'
__global__ void gram_center_kernel(const double* __restrict__ colsum_partial, int S, const double* __restrict__ G,
                                   int K, double count, float* __restrict__ abar, float* __restrict__ Ghat) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(K) * K) return;
  const int k = static_cast<int>(i / K), k2 = static_cast<int>(i % K);
  double a = 0.0, b = 0.0;
  for (int s = 0; s < S; ++s) { a += colsum_partial[static_cast<int64_t>(s) * K + k]; b += colsum_partial[static_cast<int64_t>(s) * K + k2]; }
  a /= count; b /= count;
  Ghat[i] = static_cast<float>(G[i] - count * a * b);
  if (k2 == 0) abar[k] = static_cast<float>(a);
}

// Q[k,k'] = sum_c kappa_c W[c,k] W[c,k']
__global__ void pool_q_kernel(const float* __restrict__ W, const float* __restrict__ kappa, int C, int K,
                              float* __restrict__ Q) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(K) * K) return;
  const int k = static_cast<int>(i / K), k2 = static_cast<int>(i % K);
  double s = 0.0;
  for (int c = 0; c < C; ++c)
    s += static_cast<double>(kappa[c] * W[static_cast<int64_t>(c) * K + k]) * static_cast<double>(W[static_cast<int64_t>(c) * K + k2]);
  Q[i] = static_cast<float>(s);
}

// u'[k] = sum_c alpha_c W[c,k] - sum_k' Q[k,k'] abar_k'
__global__ void pool_u_kernel(const float* __restrict__ W, const float* __restrict__ alpha, const float* __restrict__ Q,
                              const float* __restrict__ abar, int C, int K, float* __restrict__ u) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  double s = 0.0;
  for (int c = 0; c < C; ++c) s += static_cast<double>(alpha[c]) * static_cast<double>(W[static_cast<int64_t>(c) * K + k]);
  for (int k2 = 0; k2 < K; ++k2) s -= static_cast<double>(Q[static_cast<int64_t>(k2) * K + k]) * static_cast<double>(abar[k2]);
  u[k] = static_cast<float>(s);
}

// dW[c,k] = sum_b coef[b,c] (a[(b,sel),k] - abar_k) - kappa_c sum_k' W[c,k'] Ghat[k',k]
__global__ void pool_dw_kernel(ActSrc src, const float* __restrict__ W, const float* __restrict__ coef,
                               const int32_t* __restrict__ arg, const float* __restrict__ kappa,
                               const float* __restrict__ abar, const float* __restrict__ Ghat, int B, int N,
                               int C, int train, float* __restrict__ dW) {
  const int K = src.C;
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<int64_t>(C) * K) return;
  const int c = static_cast<int>(i / K), k = static_cast<int>(i % K);
  double s = 0.0;
  // train: sum_b coef = M*alpha, which folds the rank-one BN term into a centring of a[sel];
  // eval: there is no BN correction at all
  const float ab = train ? abar[k] : 0.f;
  for (int b = 0; b < B; ++b) {
    const float cf = coef[static_cast<int64_t>(b) * C + c];
    if (cf != 0.f) {
      const int64_t m = static_cast<int64_t>(b) * N + arg[static_cast<int64_t>(b) * C + c];
      s += static_cast<double>(cf) * static_cast<double>(load_act1(src, m, k) - ab);
    }
  }
  double t = 0.0;
  for (int k2 = 0; k2 < K; ++k2) t += static_cast<double>(W[static_cast<int64_t>(c) * K + k2]) * static_cast<double>(Ghat[static_cast<int64_t>(k2) * K + k]);
  dW[i] = static_cast<float>(s - static_cast<double>(kappa[c]) * t);
}

// Epilogue shared by the two dgrad kernels: val -> (ReLU mask of the producing layer) -> store
// dz_prev + accumulate that layer's dbeta / dgamma sums; or, for the network input, store grad_x.
struct DgradOut {
  float* dz_prev;        // [M, Cp] or nullptr
  const float* y_prev;   // [M, Cp] pre-BN of the previous layer (mask + yhat)
  const float* mean; const float* invstd; const float* gamma; const float* beta;
  int relu;
  double* sums;          // [2*Cp]: dbeta, dgamma of the previous layer
  float* grad_x;         // [B, Cp, N] when the previous "layer" is the input
  int Cp;
};

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
#pragma unroll
  for (int j = 0; j < 4; ++j) { atomicAdd(&ssum[tx * 4 + j], cs[j]); atomicAdd(&ssq[tx * 4 + j], cq[j]); }
  __syncthreads();
  if (threadIdx.x < TN && c0 + threadIdx.x < o.Cp) {
    atomicAdd(&o.sums[c0 + threadIdx.x], static_cast<double>(ssum[threadIdx.x]));
    atomicAdd(&o.sums[o.Cp + c0 + threadIdx.x], static_cast<double>(ssq[threadIdx.x]));
  }
}

// pooled layer dgrad: da[m,k] = sparse[m,k] - u'_k - sum_k' a[m,k'] Q[k',k]
__global__ void __launch_bounds__(kThreads)
pool_dgrad_kernel(ActSrc src, const float* __restrict__ Q, const float* __restrict__ u, const float* __restrict__ W,
                  const float* __restrict__ coef, const int32_t* __restrict__ arg, int C, int N,
                  int tiles_per_sample, DgradOut o) {
  __shared__ Tiles t;
  __shared__ float ssum[TN], ssq[TN];
  __shared__ float sparse[TM][TN + 1];
  const int K = src.C;  // == o.Cp
  const int b = blockIdx.x / tiles_per_sample;
  const int n0 = (blockIdx.x - b * tiles_per_sample) * TM;
  const int c0 = blockIdx.y * TN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (threadIdx.x < TN) { ssum[threadIdx.x] = 0.f; ssq[threadIdx.x] = 0.f; }
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
  __shared__ Tiles t;
  __shared__ float ssum[TN], ssq[TN];
  const int C = dys.C, Kp = o.Cp;
  const int b = blockIdx.x / tiles_per_sample;
  const int n0 = (blockIdx.x - b * tiles_per_sample) * TM;
  const int c0 = blockIdx.y * TN;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  if (threadIdx.x < TN) { ssum[threadIdx.x] = 0.f; ssq[threadIdx.x] = 0.f; }
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
  __shared__ float tile[32][33];
  __shared__ float ssum[32], ssq[32];
  const int b = blockIdx.z, n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  if (threadIdx.y == 0) { ssum[threadIdx.x] = 0.f; ssq[threadIdx.x] = 0.f; }
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
  atomicAdd(&ssum[threadIdx.x], s);
  atomicAdd(&ssq[threadIdx.x], q);
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    atomicAdd(&sums[c], static_cast<double>(ssum[threadIdx.x]));
    atomicAdd(&sums[C + c], static_cast<double>(ssq[threadIdx.x]));
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
  int64_t s = (M + 1023) / 1024;
  const int64_t cap = 2 * static_cast<int64_t>(sm_count());
  if (s > cap) s = cap;
  return static_cast<int>(s < 1 ? 1 : s);
}

struct Shape {
  int B, N, L, pool;
  int64_t M;
  int sumC = 0, maxC = 0, Clast = 0, Kpool = 0;
  int64_t maxWW = 0;  // largest cout*cin among layers handled by the generic wgrad
  int S;
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
  s.S = splits_for(s.M);
  return s;
}

size_t fwd_ws_bytes(const Shape& s) {
  size_t b = align_up(sizeof(double) * 2 * s.sumC);
  if (s.pool) b += align_up(sizeof(unsigned long long) * static_cast<size_t>(s.B) * s.Clast);
  b += align_up(sizeof(float) * (static_cast<size_t>(s.sumC) + 4));  // pivots
  return b;
}

size_t bwd_ws_bytes(const Shape& s) {
  size_t b = 0;
  b += align_up(sizeof(double) * 2 * s.sumC);                       // sums
  b += 2 * align_up(sizeof(float) * s.sumC);                        // alpha, kappa
  b += align_up(sizeof(float) * static_cast<size_t>(s.S) * s.maxWW);  // contraction partials
  if (s.pool) {
    const size_t K = s.Kpool;
    b += 3 * align_up(sizeof(float) * static_cast<size_t>(s.B) * s.Clast);  // coef, gsel, gyh
    b += align_up(sizeof(double) * s.S * K);                            // colsum partials
    b += align_up(sizeof(double) * K * K);                              // G
    b += 2 * align_up(sizeof(float) * K * K);                           // Ghat, Q
    b += 2 * align_up(sizeof(float) * K);                               // abar, u
  }
  b += 2 * align_up(sizeof(float) * static_cast<size_t>(s.M) * s.maxC);  // dz ping-pong
  return b;
}

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

ActSrc input_src(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int N, int C0) {
  return ActSrc{nullptr, nullptr, nullptr, nullptr, nullptr, 0, x, sxb, sxc, sxn, N, C0};
}
ActSrc layer_src(const pcuda_mlp_layer_t& y, int N) {
  return ActSrc{y.y, y.save_mean, y.save_invstd, y.gamma, y.beta, y.relu, nullptr, 0, 0, 0, N, y.cout};
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" size_t pcuda_pointmlp_ws_bytes(int B, int N, int L, const pcuda_mlp_layer_t* layers, int pool,
                                          int backward) {
  if (B < 1 || N < 1 || L < 1 || !layers) return 0;
  const Shape s = make_shape(B, N, L, layers, pool);
  return backward ? bwd_ws_bytes(s) : fwd_ws_bytes(s);
}

extern "C" int pcuda_pointmlp_fwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N, int L,
                                  const pcuda_mlp_layer_t* layers, int pool, int train, float momentum,
                                  float eps, int precision, float* out, int32_t* pool_arg, void* ws,
                                  pcuda_stream_t stream) {
  if (int rc = validate("pointmlp_fwd", B, N, L, layers, pool)) return rc;
  PCUDA_REQUIRE(x && out && ws, PCUDA_E_NULL, "pointmlp_fwd: NULL x/out/ws");
  PCUDA_REQUIRE(!pool || pool_arg, PCUDA_E_NULL, "pointmlp_fwd: pool needs pool_arg");
  PCUDA_REQUIRE(precision == PCUDA_MLP_FP32 || precision == PCUDA_MLP_BF16, PCUDA_E_UNSUPPORTED, "pointmlp_fwd: precision %d", precision);
  const Shape s = make_shape(B, N, L, layers, pool);
  PCUDA_REQUIRE(!train || s.M > 1, PCUDA_E_SHAPE, "pointmlp_fwd: train-mode BatchNorm needs more than 1 value per channel");
  if (!train)
    for (int l = 0; l < L; ++l)
      PCUDA_REQUIRE(layers[l].running_mean && layers[l].running_var, PCUDA_E_NULL, "pointmlp_fwd: eval mode needs running stats (layer %d)", l);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver cv(ws);
  double* stats = cv.take<double>(2 * static_cast<size_t>(s.sumC));
  unsigned long long* keys = pool ? cv.take<unsigned long long>(static_cast<size_t>(B) * s.Clast) : nullptr;
  float* pivots = cv.take<float>(static_cast<size_t>(s.sumC) + 4);
  cudaMemsetAsync(stats, 0, sizeof(double) * 2 * s.sumC, st);
  if (pool) cudaMemsetAsync(keys, 0, sizeof(unsigned long long) * static_cast<size_t>(B) * s.Clast, st);

  const int tps = (N + TM - 1) / TM;
  double* st_l = stats;
  float* piv_l = pivots;
  int launches = 0;
  // pivot of layer 0 (no BN to finalise yet)
  bn_finalize_pivot_kernel<<<1, 1024, 0, st>>>(nullptr, nullptr, 0, 1.0, eps, momentum, train, nullptr, nullptr, nullptr, nullptr,
                                               input_src(x, sxb, sxc, sxn, N, layers[0].cin), layers[0].weight, layers[0].bias,
                                               layers[0].cout, piv_l);
  launches += 1;
  for (int l = 0; l < L; ++l) {
    const pcuda_mlp_layer_t& y = layers[l];
    const ActSrc src = l == 0 ? input_src(x, sxb, sxc, sxn, N, y.cin) : layer_src(layers[l - 1], N);
    const dim3 grid(B * tps, (y.cout + TN - 1) / TN);
    const bool is_pool = pool && l == L - 1;
    if (is_pool)
      mlp_fwd_kernel<true><<<grid, kThreads, 0, st>>>(src, y.weight, y.bias, y.cout, N, tps, y.y, st_l, piv_l, y.gamma, keys);
    else
      mlp_fwd_kernel<false><<<grid, kThreads, 0, st>>>(src, y.weight, y.bias, y.cout, N, tps, y.y, st_l, piv_l, y.gamma, nullptr);
    // finalise this layer's BN; if a dense layer follows, compute its pivot in the same launch
    const bool has_next = l + 1 < L;
    bn_finalize_pivot_kernel<<<1, 1024, 0, st>>>(st_l, piv_l, y.cout, static_cast<double>(s.M), eps, momentum, train,
                                                 y.save_mean, y.save_invstd, y.running_mean, y.running_var,
                                                 has_next ? layer_src(y, N) : ActSrc{}, has_next ? layers[l + 1].weight : nullptr,
                                                 has_next ? layers[l + 1].bias : nullptr, has_next ? layers[l + 1].cout : 0,
                                                 piv_l + y.cout);
    launches += 2;
    st_l += 2 * y.cout;
    piv_l += y.cout;
  }
  const pcuda_mlp_layer_t& last = layers[L - 1];
  if (pool) {
    const int64_t n = static_cast<int64_t>(B) * s.Clast;
    pool_finalize_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(keys, last.save_mean, last.save_invstd, last.gamma,
                                                                            last.beta, last.relu, B, s.Clast, out, pool_arg);
  } else {
    const dim3 grid((N + 31) / 32, (s.Clast + 31) / 32, B);
    dense_out_kernel<<<grid, dim3(32, 8), 0, st>>>(layer_src(last, N), B, out);
  }
  count_launch(launches + 1);
  return check_launch("pointmlp_fwd");
}

extern "C" int pcuda_pointmlp_bwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N, int L,
                                  const pcuda_mlp_layer_t* layers, int pool, int train, float eps,
                                  int precision, const float* out, const int32_t* pool_arg,
                                  const float* grad_out, float* grad_x, void* ws, pcuda_stream_t stream) {
  (void)eps;
  if (int rc = validate("pointmlp_bwd", B, N, L, layers, pool)) return rc;
  PCUDA_REQUIRE(x && grad_out && ws, PCUDA_E_NULL, "pointmlp_bwd: NULL x/grad_out/ws");
  PCUDA_REQUIRE(!pool || (pool_arg && out), PCUDA_E_NULL, "pointmlp_bwd: pool needs out and pool_arg");
  PCUDA_REQUIRE(precision == PCUDA_MLP_FP32 || precision == PCUDA_MLP_BF16, PCUDA_E_UNSUPPORTED, "pointmlp_bwd: precision %d", precision);
  const Shape s = make_shape(B, N, L, layers, pool);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const double count = static_cast<double>(s.M);
  const int tps = (N + TM - 1) / TM;
  int launches = 0;

  Carver cv(ws);
  double* sums = cv.take<double>(2 * static_cast<size_t>(s.sumC));
  float* alpha = cv.take<float>(s.sumC);
  float* kappa = cv.take<float>(s.sumC);
  float* partial = cv.take<float>(static_cast<size_t>(s.S) * s.maxWW);
  float *coef = nullptr, *gsel = nullptr, *gyh = nullptr, *Ghat = nullptr, *Q = nullptr, *abar = nullptr, *u = nullptr;
  double *colsum = nullptr, *G = nullptr;
  if (pool) {
    const size_t K = s.Kpool;
    coef = cv.take<float>(static_cast<size_t>(B) * s.Clast);
    gsel = cv.take<float>(static_cast<size_t>(B) * s.Clast);
    gyh = cv.take<float>(static_cast<size_t>(B) * s.Clast);
    colsum = cv.take<double>(s.S * K);
    G = cv.take<double>(K * K);
    Ghat = cv.take<float>(K * K);
    Q = cv.take<float>(K * K);
    abar = cv.take<float>(K);
    u = cv.take<float>(K);
  }
  float* dzbuf[2] = {cv.take<float>(static_cast<size_t>(s.M) * s.maxC), cv.take<float>(static_cast<size_t>(s.M) * s.maxC)};
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * s.sumC, st);

  std::vector<int> off(L + 1, 0);
  for (int l = 0; l < L; ++l) off[l + 1] = off[l] + layers[l].cout;
  const int64_t chunk = (s.M + s.S - 1) / s.S;
  auto src_of = [&](int l) { return l < 0 ? input_src(x, sxb, sxc, sxn, N, layers[0].cin) : layer_src(layers[l], N); };
  auto dgrad_out = [&](int lp, float* dzp) {  // epilogue target: previous layer lp (or input if lp < 0)
    DgradOut o{};
    if (lp < 0) {
      o.grad_x = grad_x;
      o.Cp = layers[0].cin;
    } else {
      const pcuda_mlp_layer_t& p = layers[lp];
      o = DgradOut{dzp, p.y, p.save_mean, p.save_invstd, p.gamma, p.beta, p.relu, sums + 2 * off[lp], nullptr, p.cout};
    }
    return o;
  };

  int cur = 0;      // dzbuf[cur] holds dz of layer `top`
  int top = L - 1;  // highest layer whose dz is dense and stored
  const pcuda_mlp_layer_t& last = layers[L - 1];
  if (pool) {
    const int C = last.cout, K = last.cin;
    const ActSrc src = src_of(L - 2);
    float* al = alpha + off[L - 1];
    float* ka = kappa + off[L - 1];
    const int64_t bc = static_cast<int64_t>(B) * C;
    pool_sel_kernel<<<static_cast<int>((bc * 32 + 255) / 256), 256, 0, st>>>(src, last.weight, last.bias, last.save_mean,
                                                                            last.save_invstd, last.gamma, last.relu, out, pool_arg,
                                                                            grad_out, B, N, C, coef, gsel, gyh);
    const bool want_last = last.grad_weight != nullptr;
    pool_coef_kernel<<<(C + 127) / 128, 128, 0, st>>>(gsel, gyh, last.save_invstd, last.gamma, B, C, count, train, al, ka,
                                                      want_last ? last.grad_gamma : nullptr, want_last ? last.grad_beta : nullptr,
                                                      want_last ? last.grad_bias : nullptr);
    launches += 1;
    act_colsum_kernel<<<dim3(s.S, (K + 127) / 128), 128, 0, st>>>(src, s.M, chunk, colsum);
    point_contract_kernel<1><<<dim3(s.S, (K + TM - 1) / TM, (K + TN - 1) / TN), kThreads, 0, st>>>(DySrc{}, src, src, s.M, chunk, K, K, partial);
    const int64_t kk = static_cast<int64_t>(K) * K;
    reduce_partials_kernel<double><<<static_cast<int>((kk + 255) / 256), 256, 0, st>>>(partial, kk, s.S, G);
    gram_center_kernel<<<static_cast<int>((kk + 255) / 256), 256, 0, st>>>(colsum, s.S, G, K, count, abar, Ghat);
    pool_q_kernel<<<static_cast<int>((kk + 255) / 256), 256, 0, st>>>(last.weight, ka, C, K, Q);
    pool_u_kernel<<<(K + 127) / 128, 128, 0, st>>>(last.weight, al, Q, abar, C, K, u);
    launches += 7;
    if (last.grad_weight) {
      const int64_t ck = static_cast<int64_t>(C) * K;
      pool_dw_kernel<<<static_cast<int>((ck + 255) / 256), 256, 0, st>>>(src, last.weight, coef, pool_arg, ka, abar, Ghat, B, N, C, train, last.grad_weight);
      launches += 1;
    }
    if (L >= 2 || grad_x) {
      const DgradOut o = dgrad_out(L - 2, dzbuf[cur]);
      pool_dgrad_kernel<<<dim3(B * tps, (K + TN - 1) / TN), kThreads, 0, st>>>(src, Q, u, last.weight, coef, pool_arg, C, N, tps, o);
      launches += 1;
    }
    top = L - 2;
  } else {
    const int C = last.cout;
    dense_top_kernel<<<dim3((N + 31) / 32, (C + 31) / 32, B), dim3(32, 8), 0, st>>>(grad_out, last.y, last.save_mean, last.save_invstd,
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
    bn_bwd_coef_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums + 2 * off[l], y.save_invstd, y.gamma, C, count, train, al, ka,
                                                        want_w ? y.grad_gamma : nullptr, want_w ? y.grad_beta : nullptr,
                                                        want_w ? y.grad_bias : nullptr);
    launches += 1;
    const DySrc dys{dzbuf[cur], y.y, y.save_mean, y.save_invstd, y.gamma, al, ka, C};
    if (want_w) {
      point_contract_kernel<0><<<dim3(s.S, (C + TM - 1) / TM, (Kp + TN - 1) / TN), kThreads, 0, st>>>(dys, ActSrc{}, src_of(l - 1), s.M, chunk, C, Kp, partial);
      const int64_t ck = static_cast<int64_t>(C) * Kp;
      reduce_partials_kernel<float><<<static_cast<int>((ck + 255) / 256), 256, 0, st>>>(partial, ck, s.S, y.grad_weight);
      launches += 2;
    }
    if (l > 0 || grad_x) {
      const DgradOut o = dgrad_out(l - 1, dzbuf[cur ^ 1]);
      dense_dgrad_kernel<<<dim3(B * tps, (Kp + TN - 1) / TN), kThreads, 0, st>>>(dys, y.weight, N, tps, o);
      launches += 1;
      cur ^= 1;
    }
  }
  count_launch(launches);
  return check_launch("pointmlp_bwd");
}
