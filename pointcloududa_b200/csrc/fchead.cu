// A whole FC head of the point-cloud discriminator — Linear(+Dropout mask)+BatchNorm1d+ReLU, twice, then a plain
// Linear (+ flattened identity): STN3d / STNkd heads 1024->512->256->k*k (networks/PointNetCls.py:46-62, :89-101) and the
// classifier head 1024->512->256->1 (:208-213) — as ONE thread-block-cluster launch per direction.
//
// fcstack.cu runs a head as one launch per layer and direction (3 forward + 3 backward + 1 input-gradient launches of
// 5-9 us each): at the reference's batch (8 clouds) that is ~85 us of dependent launches per D4 pass, on the critical
// path of a 0.48 ms step.  Here the layers of a head live in one kernel:
//   * a cluster of kCluster CTAs; CTA r owns channel slice r of EVERY layer's activation, so BatchNorm over the batch
//     (forward statistics, backward reductions) stays CTA-local exactly as in fcstack.cu: no atomics, deterministic;
//   * between layers the slices are exchanged through distributed shared memory (cluster.map_shared_rank) behind ONE
//     cluster barrier per layer (the slices ping-pong between two buffers);
//   * forward: CTA r computes y_l[:, slice r] = a_{l-1} W_l[slice r, :]^T from the gathered a_{l-1};
//     backward: CTA r gathers dy_{l+1}, pulls da_l[:, slice r] = dy_{l+1} W_{l+1}[:, slice r], runs ReLU / BatchNorm
//     backward on its channels, and writes rows slice r of dW_l, db_l, dgamma_l, dbeta_l; the input gradient is a last
//     pull through W_0 by columns.
// Fast path for small batches (B <= kHeadMaxB rows, the regime where launches dominate); anything else — larger
// batches, other layer patterns — stays on the per-layer kernels of fcstack.cu.  Same arithmetic per element as those.
#include <cooperative_groups.h>

#include <algorithm>

#include "pcuda_common.cuh"

namespace cg = cooperative_groups;

namespace pcuda {

constexpr int kHeadThreads = 512;
constexpr int kHeadWarps = kHeadThreads / 32;
constexpr int kHeadMaxB = 16;
constexpr int kHeadMaxL = 4;
constexpr int kHeadMaxHidden = 512;      // widest activation that is exchanged between CTAs (all but the last layer)
constexpr int kHeadMaxCin = 1024;

struct HeadLayer {
  const float* W; const float* bias; const float* mask; const float* gamma; const float* beta;
  float* running_mean; float* running_var; float* save_mean; float* save_invstd; float* y; float* a;
  float* gW; float* gb; float* ggamma; float* gbeta;
  int cin, cout, bn, relu;
};
struct HeadParams {
  HeadLayer l[kHeadMaxL];
  const float* x;
  const float* grad_out;
  float* grad_x;
  int B, L, train, iden_k;
  float momentum, eps;
};

namespace {

__device__ __forceinline__ int slice_per(int cout, int cs) { return (cout + cs - 1) / cs; }

// ---- forward ----------------------------------------------------------------------------------------------
// dynamic smem: xs[B][cin_max] | slice[2][B][per_hidden_max] | ys[nc_max][kHeadMaxB]
__global__ void __launch_bounds__(kHeadThreads) fc_head_fwd_kernel(const __grid_constant__ HeadParams p, int cin_max, int per_hidden_max, int nc_max) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = static_cast<int>(cluster.num_blocks()), rank = static_cast<int>(cluster.block_rank());
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;
  float* slices = xs + static_cast<size_t>(p.B) * cin_max;
  float* ys = slices + 2 * static_cast<size_t>(p.B) * per_hidden_max;
  const int B = p.B;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    const int n4 = B * p.l[0].cin / 4;
    for (int i = threadIdx.x; i < n4; i += kHeadThreads) reinterpret_cast<float4*>(xs)[i] = __ldg(reinterpret_cast<const float4*>(p.x) + i);
  }
  __syncthreads();
  for (int l = 0; l < p.L; ++l) {
    const HeadLayer& y = p.l[l];
    const int cin = y.cin, cout = y.cout, cin4 = cin >> 2;
    const int per = slice_per(cout, CS);
    const int c_begin = min(cout, rank * per), nc = min(cout, c_begin + per) - c_begin;
    float* my_slice = slices + static_cast<size_t>(l & 1) * B * per_hidden_max;
    // (1) y[b][c] for this CTA's channels: a warp takes two channels at a time, lanes split the contraction
    for (int cp = warp * 2; cp < nc; cp += 2 * kHeadWarps) {
      const bool two = cp + 1 < nc;
      const float4* w0 = reinterpret_cast<const float4*>(y.W + static_cast<int64_t>(c_begin + cp) * cin);
      const float4* w1 = reinterpret_cast<const float4*>(y.W + static_cast<int64_t>(c_begin + cp + (two ? 1 : 0)) * cin);
      float acc0[kHeadMaxB], acc1[kHeadMaxB];
#pragma unroll
      for (int b = 0; b < kHeadMaxB; ++b) { acc0[b] = 0.f; acc1[b] = 0.f; }
#pragma unroll 2
      for (int k4 = lane; k4 < cin4; k4 += 32) {
        const float4 a0 = __ldg(w0 + k4), a1 = __ldg(w1 + k4);
#pragma unroll
        for (int b = 0; b < kHeadMaxB; ++b) {
          if (b < B) {
            const float4 xv = reinterpret_cast<const float4*>(xs)[b * cin4 + k4];
            acc0[b] = fmaf(xv.x, a0.x, acc0[b]); acc0[b] = fmaf(xv.y, a0.y, acc0[b]);
            acc0[b] = fmaf(xv.z, a0.z, acc0[b]); acc0[b] = fmaf(xv.w, a0.w, acc0[b]);
            acc1[b] = fmaf(xv.x, a1.x, acc1[b]); acc1[b] = fmaf(xv.y, a1.y, acc1[b]);
            acc1[b] = fmaf(xv.z, a1.z, acc1[b]); acc1[b] = fmaf(xv.w, a1.w, acc1[b]);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < kHeadMaxB; ++b) {
        if (b < B) {
          const float s0 = warp_sum(acc0[b]), s1 = warp_sum(acc1[b]);
          if (lane == b) {
            ys[cp * kHeadMaxB + b] = s0;
            if (two) ys[(cp + 1) * kHeadMaxB + b] = s1;
          }
        }
      }
    }
    __syncthreads();
    // (2) bias, dropout mask, BatchNorm over the batch, ReLU, + identity: one thread per channel of the slice
    for (int t = threadIdx.x; t < nc; t += kHeadThreads) {
      const int c = c_begin + t;
      float v[kHeadMaxB];
      const float bj = y.bias ? y.bias[c] : 0.f;
#pragma unroll
      for (int b = 0; b < kHeadMaxB; ++b) {
        v[b] = 0.f;
        if (b < B) {
          float u = ys[t * kHeadMaxB + b] + bj;
          if (y.mask) u *= y.mask[static_cast<int64_t>(b) * cout + c];
          v[b] = u;
        }
      }
      float mean = 0.f, scale = 1.f, shift = 0.f;
      if (y.bn) {
        float invstd;
        if (p.train) {
          float s = 0.f;
#pragma unroll
          for (int b = 0; b < kHeadMaxB; ++b) if (b < B) s += v[b];
          mean = s / static_cast<float>(B);
          float q = 0.f;
#pragma unroll
          for (int b = 0; b < kHeadMaxB; ++b) if (b < B) { const float d = v[b] - mean; q = fmaf(d, d, q); }
          const float var = q / static_cast<float>(B);
          invstd = rsqrtf(var + p.eps);
          if (y.running_mean) y.running_mean[c] = (1.f - p.momentum) * y.running_mean[c] + p.momentum * mean;
          if (y.running_var) {
            const float unb = B > 1 ? var * static_cast<float>(B) / static_cast<float>(B - 1) : var;
            y.running_var[c] = (1.f - p.momentum) * y.running_var[c] + p.momentum * unb;
          }
        } else {
          mean = y.running_mean[c];
          invstd = rsqrtf(y.running_var[c] + p.eps);
        }
        y.save_mean[c] = mean;
        y.save_invstd[c] = invstd;
        scale = invstd * y.gamma[c];
        shift = y.beta[c];
      }
      const bool iden = l == p.L - 1 && p.iden_k > 0 && c % (p.iden_k + 1) == 0;
#pragma unroll
      for (int b = 0; b < kHeadMaxB; ++b) {
        if (b < B) {
          float z = fmaf(v[b] - mean, scale, shift);
          if (y.relu) z = fmaxf(z, 0.f);
          if (iden) z += 1.0f;
          const int64_t o = static_cast<int64_t>(b) * cout + c;
          if (y.y) y.y[o] = v[b];
          y.a[o] = z;
          if (l + 1 < p.L) my_slice[b * per + t] = z;
        }
      }
    }
    if (l + 1 == p.L) break;
    cluster.sync();                     // every CTA's slice of a_l is complete and visible cluster-wide
    // (3) gather a_l = input of the next layer from the peers' slices (distributed shared memory)
    for (int i = threadIdx.x; i < B * cout; i += kHeadThreads) {
      const int b = i / cout, c = i - b * cout;
      const int r = c / per;
      const float* remote = cluster.map_shared_rank(my_slice, r);
      xs[b * cout + c] = remote[b * per + (c - r * per)];
    }
    __syncthreads();
  }
  cluster.sync();                       // no CTA exits while a peer may still read its shared memory
}

// ---- backward ---------------------------------------------------------------------------------------------
// dynamic smem: xs[B][cin_max] | dfull[B][kHeadMaxHidden] | dslice[2][B][per_hidden_max]
__global__ void __launch_bounds__(kHeadThreads) fc_head_bwd_kernel(const __grid_constant__ HeadParams p, int cin_max, int per_hidden_max) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = static_cast<int>(cluster.num_blocks()), rank = static_cast<int>(cluster.block_rank());
  extern __shared__ __align__(16) float sm[];
  float* xs = sm;
  float* dfull = xs + static_cast<size_t>(p.B) * cin_max;
  float* dslices = dfull + static_cast<size_t>(p.B) * kHeadMaxHidden;
  const int B = p.B, L = p.L;
  for (int l = L - 1; l >= 0; --l) {
    const HeadLayer& y = p.l[l];
    const int cin = y.cin, cout = y.cout;
    const int per = slice_per(cout, CS);
    const int c_begin = min(cout, rank * per), nc = min(cout, c_begin + per) - c_begin;
    const bool last = l == L - 1;          // plain Linear (validated by the host): dy = grad_out, read from global
    float* my_d = dslices + static_cast<size_t>(l & 1) * B * per_hidden_max;
    if (!last) {
      const HeadLayer& nx = p.l[l + 1];
      const int cn = nx.cout;
      const float* dnext;                 // dy_{l+1} as [B][cn]
      if (l + 1 == L - 1) {
        dnext = p.grad_out;
      } else {
        const int pern = slice_per(cn, CS);
        const float* their = dslices + static_cast<size_t>((l + 1) & 1) * B * per_hidden_max;
        for (int i = threadIdx.x; i < B * cn; i += kHeadThreads) {
          const int b = i / cn, o = i - b * cn;
          const int r = o / pern;
          dfull[i] = cluster.map_shared_rank(their, r)[b * pern + (o - r * pern)];
        }
        __syncthreads();
        dnext = dfull;
      }
      // (1) da[b][t] = sum_o dy_{l+1}[b][o] W_{l+1}[o][c_begin + t]: thread = (channel t, row group g); the row
      //     segments of W_{l+1} are read coalesced across t, dy is a shared-memory / L1 broadcast
      const int groups = kHeadThreads / per;                       // per <= 64 here (hidden layers)
      const int t = threadIdx.x % per, g = threadIdx.x / per;
      const int rows = (B + groups - 1) / groups;                  // rows per group
      if (g < groups && t < nc && g * rows < B) {
        float acc[kHeadMaxB];
#pragma unroll
        for (int j = 0; j < kHeadMaxB; ++j) acc[j] = 0.f;
        const int b0 = g * rows;
        const float* wcol = nx.W + c_begin + t;
#pragma unroll 4
        for (int o = 0; o < cn; ++o) {
          const float wv = __ldg(wcol + static_cast<int64_t>(o) * cout);
#pragma unroll
          for (int j = 0; j < kHeadMaxB; ++j)
            if (j < rows && b0 + j < B) acc[j] = fmaf(dnext[(b0 + j) * cn + o], wv, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < kHeadMaxB; ++j)
          if (j < rows && b0 + j < B) my_d[(b0 + j) * per + t] = acc[j];
      }
      __syncthreads();
      // (2) ReLU mask, BatchNorm backward over the batch, dropout mask: one thread per channel
      for (int tt = threadIdx.x; tt < nc; tt += kHeadThreads) {
        const int c = c_begin + tt;
        const bool bn = y.bn != 0;
        const float mean = bn ? y.save_mean[c] : 0.f, invstd = bn ? y.save_invstd[c] : 1.f;
        float dz[kHeadMaxB], yh[kHeadMaxB];
        float sb = 0.f, sg = 0.f;
#pragma unroll
        for (int b = 0; b < kHeadMaxB; ++b) {
          dz[b] = 0.f; yh[b] = 0.f;
          if (b < B) {
            const int64_t o = static_cast<int64_t>(b) * cout + c;
            float d = my_d[b * per + tt];
            if (y.relu && !(y.a[o] > 0.f)) d = 0.f;
            dz[b] = d;
            if (bn) {
              yh[b] = (y.y[o] - mean) * invstd;
              sb += d;
              sg = fmaf(d, yh[b], sg);
            }
          }
        }
        float dbias = 0.f;
        const float gsc = bn ? y.gamma[c] * invstd : 1.f;
        const float kb = (bn && p.train) ? sb / static_cast<float>(B) : 0.f;
        const float kg = (bn && p.train) ? sg / static_cast<float>(B) : 0.f;
#pragma unroll
        for (int b = 0; b < kHeadMaxB; ++b) {
          if (b < B) {
            float d = bn ? gsc * (dz[b] - kb - yh[b] * kg) : dz[b];
            if (y.mask) d *= y.mask[static_cast<int64_t>(b) * cout + c];
            my_d[b * per + tt] = d;
            dbias += d;
          }
        }
        if (y.gW != nullptr) {
          if (bn && y.ggamma) y.ggamma[c] = sg;
          if (bn && y.gbeta) y.gbeta[c] = sb;
          if (y.gb) y.gb[c] = dbias;
        }
      }
      __syncthreads();
    }
    // (3) parameter gradients of this CTA's rows: dW[c][k] = sum_b dy[b][c] a_{l-1}[b][k]
    if (y.gW != nullptr && nc > 0) {
      const float* aprev = l == 0 ? p.x : p.l[l - 1].a;
      const int n4 = B * cin / 4;
      for (int i = threadIdx.x; i < n4; i += kHeadThreads) reinterpret_cast<float4*>(xs)[i] = __ldg(reinterpret_cast<const float4*>(aprev) + i);
      // dy of this CTA's channels as [b][stride]: the slice buffer, or (last layer: dy = grad_out) a staged copy
      const float* dsrc = my_d;
      int dstride = per;
      if (last) {
        for (int i = threadIdx.x; i < B * nc; i += kHeadThreads) {
          const int b = i / nc, tt = i - b * nc;
          dfull[b * nc + tt] = p.grad_out[static_cast<int64_t>(b) * cout + c_begin + tt];
        }
        dsrc = dfull;
        dstride = nc;
      }
      __syncthreads();
      if (last && y.gb != nullptr) {
        for (int tt = threadIdx.x; tt < nc; tt += kHeadThreads) {
          float s = 0.f;
          for (int b = 0; b < B; ++b) s += dsrc[b * dstride + tt];
          y.gb[c_begin + tt] = s;
        }
      }
      for (int t0 = 0; t0 < nc; t0 += 8) {
        for (int k = threadIdx.x; k < cin; k += kHeadThreads) {
          float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int b = 0; b < kHeadMaxB; ++b) {
            if (b < B) {
              const float xv = xs[b * cin + k];
              const float* dr = dsrc + b * dstride + t0;        // warp-wide broadcast reads
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (t0 + j < nc) acc[j] = fmaf(dr[j], xv, acc[j]);
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (t0 + j < nc) y.gW[static_cast<int64_t>(c_begin + t0 + j) * cin + k] = acc[j];
        }
      }
    }
    if (!last && l > 0) cluster.sync();    // dy_l slices complete before the peers gather them
    else __syncthreads();
  }
  // (4) input gradient: dx[b][k] = sum_c dy_0[b][c] W_0[c][k] for this CTA's columns k
  if (p.grad_x != nullptr) {
    const HeadLayer& y0 = p.l[0];
    const int cin = y0.cin, c0n = y0.cout;
    const float* d0;
    if (L == 1) {
      d0 = p.grad_out;
    } else {
      cluster.sync();                     // dy_0 slices complete
      const int per0 = slice_per(c0n, CS);
      const float* their = dslices;       // layer 0 -> buffer 0
      for (int i = threadIdx.x; i < B * c0n; i += kHeadThreads) {
        const int b = i / c0n, c = i - b * c0n;
        const int r = c / per0;
        dfull[i] = cluster.map_shared_rank(their, r)[b * per0 + (c - r * per0)];
      }
      __syncthreads();
      d0 = dfull;
    }
    const int perk = slice_per(cin, CS);
    const int k_begin = min(cin, rank * perk), nk = min(cin, k_begin + perk) - k_begin;
    const int groups = max(1, kHeadThreads / max(perk, 1));
    const int rows = (B + groups - 1) / groups;
    for (int kk = threadIdx.x % perk; kk < nk; kk += perk) {        // one pass: perk <= kHeadThreads (validated by the host)
      const int g = threadIdx.x / perk;
      const int b0 = g * rows;
      if (g >= groups || b0 >= B) continue;
      float acc[kHeadMaxB];
#pragma unroll
      for (int j = 0; j < kHeadMaxB; ++j) acc[j] = 0.f;
      const float* wcol = y0.W + k_begin + kk;
#pragma unroll 4
      for (int c = 0; c < c0n; ++c) {
        const float wv = __ldg(wcol + static_cast<int64_t>(c) * cin);
#pragma unroll
        for (int j = 0; j < kHeadMaxB; ++j)
          if (j < rows && b0 + j < B) acc[j] = fmaf(d0[(b0 + j) * c0n + c], wv, acc[j]);
      }
#pragma unroll
      for (int j = 0; j < kHeadMaxB; ++j)
        if (j < rows && b0 + j < B) p.grad_x[static_cast<int64_t>(b0 + j) * cin + k_begin + kk] = acc[j];
    }
  }
  cluster.sync();                         // no CTA exits while a peer may still read its shared memory
}

struct HeadShape {
  int cin_max = 0, per_hidden_max = 1, nc_max = 1;
};

}  // namespace

// Whether a head runs on the cluster kernels: small batch, BatchNorm / ReLU / mask only on the layers before a plain
// last Linear, hidden widths that split evenly over the cluster.
bool fc_head_supported(int B, int L, const pcuda_fc_layer_t* layers, int cluster) {
  if (tuning(TUNE_FC_NO_CLUSTER)) return false;
  if (B < 1 || B > kHeadMaxB || L < 2 || L > kHeadMaxL) return false;
  for (int l = 0; l < L; ++l) {
    const pcuda_fc_layer_t& y = layers[l];
    if ((y.cin & 3) != 0 || y.cin > kHeadMaxCin) return false;
    if (l < L - 1 && (y.cout > kHeadMaxHidden || y.cout % cluster != 0 || y.cout / cluster > 64 || kHeadThreads % (y.cout / cluster) != 0)) return false;
    if (l == L - 1 && (y.bn || y.relu || y.mask != nullptr || (y.cout + cluster - 1) / cluster > kHeadMaxHidden)) return false;
  }
  const int perk = (layers[0].cin + cluster - 1) / cluster;
  if (perk > kHeadThreads) return false;
  return true;
}

static HeadShape head_shape(int L, const pcuda_fc_layer_t* layers, int cluster) {
  HeadShape s;
  for (int l = 0; l < L; ++l) {
    s.cin_max = std::max(s.cin_max, layers[l].cin);
    const int per = (layers[l].cout + cluster - 1) / cluster;
    s.nc_max = std::max(s.nc_max, per);
    if (l < L - 1) s.per_hidden_max = std::max(s.per_hidden_max, per);
  }
  return s;
}

static void fill_params(HeadParams& p, const float* x, int B, int L, const pcuda_fc_layer_t* layers, int train, float momentum,
                        float eps, int iden_k) {
  p.x = x; p.B = B; p.L = L; p.train = train; p.iden_k = iden_k; p.momentum = momentum; p.eps = eps;
  for (int l = 0; l < L; ++l) {
    const pcuda_fc_layer_t& y = layers[l];
    HeadLayer& h = p.l[l];
    h.W = y.weight; h.bias = y.bias; h.mask = y.mask; h.gamma = y.gamma; h.beta = y.beta;
    h.running_mean = y.running_mean; h.running_var = y.running_var; h.save_mean = y.save_mean; h.save_invstd = y.save_invstd;
    h.y = y.y; h.a = y.a; h.gW = y.grad_weight; h.gb = y.grad_weight ? y.grad_bias : nullptr;
    h.ggamma = y.grad_weight ? y.grad_gamma : nullptr; h.gbeta = y.grad_weight ? y.grad_beta : nullptr;
    h.cin = y.cin; h.cout = y.cout; h.bn = y.bn; h.relu = y.relu;
  }
}

template <class K, class... Args>
static cudaError_t launch_cluster(K kernel, int cluster, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster);
  cfg.blockDim = dim3(kHeadThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}

int fc_head_fwd(const float* x, int B, int L, const pcuda_fc_layer_t* layers, int train, float momentum, float eps, int iden_k,
                int cluster, cudaStream_t st) {
  HeadParams p{};
  fill_params(p, x, B, L, layers, train, momentum, eps, iden_k);
  const HeadShape s = head_shape(L, layers, cluster);
  const size_t smem = sizeof(float) * (static_cast<size_t>(B) * s.cin_max + 2 * static_cast<size_t>(B) * s.per_hidden_max +
                                       static_cast<size_t>(s.nc_max) * kHeadMaxB);
  if (smem > 200 * 1024) return PCUDA_E_UNSUPPORTED;
  if (cudaError_t e = smem_optin(fc_head_fwd_kernel, 200 * 1024)) return fail(static_cast<int>(e), "fc_head_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const cudaError_t e = launch_cluster(fc_head_fwd_kernel, cluster, smem, st, p, s.cin_max, s.per_hidden_max, s.nc_max);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "fc_head_fwd: %s", cudaGetErrorString(e));
  count_launch(1);
  return check_launch("fc_head_fwd");
}

int fc_head_bwd(const float* x, int B, int L, const pcuda_fc_layer_t* layers, int train, const float* grad_out, float* grad_x,
                int cluster, cudaStream_t st) {
  HeadParams p{};
  fill_params(p, x, B, L, layers, train, 0.f, 0.f, 0);
  p.grad_out = grad_out; p.grad_x = grad_x;
  const HeadShape s = head_shape(L, layers, cluster);
  const size_t smem = sizeof(float) * (static_cast<size_t>(B) * s.cin_max + static_cast<size_t>(B) * kHeadMaxHidden +
                                       2 * static_cast<size_t>(B) * s.per_hidden_max);
  if (smem > 200 * 1024) return PCUDA_E_UNSUPPORTED;
  if (cudaError_t e = smem_optin(fc_head_bwd_kernel, 200 * 1024)) return fail(static_cast<int>(e), "fc_head_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const cudaError_t e = launch_cluster(fc_head_bwd_kernel, cluster, smem, st, p, s.cin_max, s.per_hidden_max);
  if (e != cudaSuccess) return fail(static_cast<int>(e), "fc_head_bwd: %s", cudaGetErrorString(e));
  count_launch(1);
  return check_launch("fc_head_bwd");
}

}  // namespace pcuda
