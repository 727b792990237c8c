// Per-cloud feature transform of PointNetfeat: x' = (x^T T)^T, i.e. x'[b, k, n] = sum_j T[b][j][k] x[b, j, n]
// (torch.bmm(x.transpose(2, 1), trans_feat).transpose(2, 1), networks/PointNetCls.py:147-151, with the 64 x 64
// matrix STNkd predicts; the 3 x 3 input transform of :140-142 is fused into the first shared-MLP layer's operand
// load instead, see pointmlp.cu) and its autograd backward
//   grad_x[b, j, n] = sum_k T[b][j][k] g[b, k, n],        grad_T[b][j][k] = sum_n x[b, j, n] g[b, k, n].
// FP32 CUDA cores: 2 K^2 FLOP per point next to the 2 * 139,456 of the shared MLP it feeds.  A thread owns one point
// (its K inputs stream through one register, K accumulators stay in registers), T[b] sits in shared memory and is read
// as warp-wide broadcasts; summation order is j (or k) ascending.  grad_T is a contraction over points: 64 x 64 tile
// per CTA over a slice of the cloud, partials combined in a fixed order (deterministic).
#include "pcuda_common.cuh"

namespace pcuda {
namespace {

constexpr int kMaxK = 64;
constexpr int kPtThreads = 128;

// TRANSPOSED = false: out[k] = sum_j T[j][k] in[j];   true: out[j] = sum_k T[j][k] in[k]
template <bool TRANSPOSED>
__global__ void __launch_bounds__(kPtThreads)
point_transform_kernel(const float* __restrict__ x, int64_t sxb, int64_t sxc, int64_t sxn, const float* __restrict__ T,
                       int K, int N, float* __restrict__ out) {
  pdl_entry();
  __shared__ __align__(16) float sT[kMaxK * kMaxK];          // sT[i * kMaxK + o]: input channel i -> output channel o
  const int b = blockIdx.y;
  const float* Tb = T + static_cast<int64_t>(b) * K * K;
  for (int i = threadIdx.x; i < kMaxK * kMaxK; i += kPtThreads) {
    const int in = i / kMaxK, o = i - in * kMaxK;
    float v = 0.f;
    if (in < K && o < K) v = TRANSPOSED ? Tb[o * K + in] : Tb[in * K + o];
    sT[i] = v;
  }
  __syncthreads();
  const int n = blockIdx.x * kPtThreads + threadIdx.x;
  if (n >= N) return;
  float acc[kMaxK];
#pragma unroll
  for (int o = 0; o < kMaxK; ++o) acc[o] = 0.f;
  const float* px = x + b * sxb + n * sxn;
  for (int in = 0; in < K; ++in) {
    const float v = __ldg(px + in * sxc);
    const float4* row = reinterpret_cast<const float4*>(sT + in * kMaxK);
#pragma unroll
    for (int q = 0; q < kMaxK / 4; ++q) {
      const float4 t = row[q];
      acc[4 * q + 0] = fmaf(v, t.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(v, t.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(v, t.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(v, t.w, acc[4 * q + 3]);
    }
  }
  float* po = out + static_cast<int64_t>(b) * K * N + n;
#pragma unroll
  for (int o = 0; o < kMaxK; ++o)
    if (o < K) po[static_cast<int64_t>(o) * N] = acc[o];
}

// partial[s][b][j][k] = sum_{n in slice s} x[b, j, n] g[b, k, n].  grid (S, B), 256 threads = 16 x 16, 4 x 4 outputs each
__global__ void __launch_bounds__(256)
transform_wgrad_kernel(const float* __restrict__ x, int64_t sxb, int64_t sxc, int64_t sxn, const float* __restrict__ g,
                       int K, int N, int chunk, float* __restrict__ partial) {
  pdl_entry();
  __shared__ float xs[32][kMaxK + 4], gs[32][kMaxK + 4];
  const int b = blockIdx.y, s = blockIdx.x;
  const int n0 = s * chunk, n1 = min(N, n0 + chunk);
  const int tj = threadIdx.x >> 4, tk = threadIdx.x & 15;
  float acc[4][4] = {};
  for (int nn = n0; nn < n1; nn += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * kMaxK; i += 256) {
      const int c = i >> 5, r = i & 31;                 // consecutive threads -> consecutive points: coalesced
      const int n = nn + r;
      const bool ok = n < n1 && c < K;
      xs[r][c] = ok ? __ldg(x + b * sxb + c * sxc + n * sxn) : 0.f;
      gs[r][c] = ok ? g[(static_cast<int64_t>(b) * K + c) * N + n] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&xs[r][tj * 4]);
      const float4 c = *reinterpret_cast<const float4*>(&gs[r][tk * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], cv[j], acc[i][j]);
    }
  }
  float* outp = partial + (static_cast<int64_t>(s) * gridDim.y + b) * K * K;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int jj = tj * 4 + i, kk = tk * 4 + j;
      if (jj < K && kk < K) outp[jj * K + kk] = acc[i][j];
    }
}

__global__ void transform_wgrad_reduce_kernel(const float* __restrict__ partial, int64_t n, int S, float* __restrict__ out) {
  pdl_entry();
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double t = 0.0;
  for (int s = 0; s < S; ++s) t += static_cast<double>(partial[static_cast<int64_t>(s) * n + i]);
  out[i] = static_cast<float>(t);
}

int wgrad_splits(int N) { return std::max(1, std::min(32, (N + 255) / 256)); }

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" size_t pcuda_point_transform_ws_bytes(int B, int K, int N) {
  if (B < 1 || K < 1 || N < 1) return 0;
  return sizeof(float) * static_cast<size_t>(wgrad_splits(N)) * B * K * K;
}

extern "C" int pcuda_point_transform_fwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* trans, int B, int K,
                                         int N, float* out, pcuda_stream_t stream) {
  PCUDA_REQUIRE(B >= 0 && N >= 0 && K >= 1, PCUDA_E_SHAPE, "point_transform_fwd: bad shape B=%d K=%d N=%d", B, K, N);
  PCUDA_REQUIRE(K <= kMaxK, PCUDA_E_UNSUPPORTED, "point_transform_fwd: K=%d > %d", K, kMaxK);
  if (B == 0 || N == 0) return 0;
  PCUDA_REQUIRE(x && trans && out, PCUDA_E_NULL, "point_transform_fwd: NULL argument");
  PCUDA_LAUNCH(point_transform_kernel<false>, dim3((N + kPtThreads - 1) / kPtThreads, B), kPtThreads, 0, static_cast<cudaStream_t>(stream), 
      x, sxb, sxc, sxn, trans, K, N, out);
  count_launch(1);
  return check_launch("point_transform_fwd");
}

extern "C" int pcuda_point_transform_bwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* trans,
                                         const float* grad_out, int B, int K, int N, float* grad_x, float* grad_trans, void* ws,
                                         pcuda_stream_t stream) {
  PCUDA_REQUIRE(B >= 0 && N >= 0 && K >= 1, PCUDA_E_SHAPE, "point_transform_bwd: bad shape B=%d K=%d N=%d", B, K, N);
  PCUDA_REQUIRE(K <= kMaxK, PCUDA_E_UNSUPPORTED, "point_transform_bwd: K=%d > %d", K, kMaxK);
  if (B == 0 || N == 0) return 0;
  PCUDA_REQUIRE(x && trans && grad_out, PCUDA_E_NULL, "point_transform_bwd: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int launches = 0;
  if (grad_x != nullptr) {
    // grad_out is contiguous [B, K, N]
    PCUDA_LAUNCH(point_transform_kernel<true>, dim3((N + kPtThreads - 1) / kPtThreads, B), kPtThreads, 0, st, 
        grad_out, static_cast<int64_t>(K) * N, N, 1, trans, K, N, grad_x);
    launches += 1;
  }
  if (grad_trans != nullptr) {
    PCUDA_REQUIRE(ws != nullptr, PCUDA_E_WORKSPACE, "point_transform_bwd: grad_trans needs the workspace");
    const int S = wgrad_splits(N);
    const int chunk = ((N + S - 1) / S + 31) / 32 * 32;
    float* partial = S == 1 ? grad_trans : static_cast<float*>(ws);
    PCUDA_LAUNCH(transform_wgrad_kernel, dim3(S, B), 256, 0, st, x, sxb, sxc, sxn, grad_out, K, N, chunk, partial);
    launches += 1;
    if (S > 1) {
      const int64_t n = static_cast<int64_t>(B) * K * K;
      PCUDA_LAUNCH(transform_wgrad_reduce_kernel, static_cast<int>((n + 255) / 256), 256, 0, st, partial, n, S, grad_trans);
      launches += 1;
    }
  }
  count_launch(launches);
  return check_launch("point_transform_bwd");
}
