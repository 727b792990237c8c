// FC heads of the point-cloud discriminator: Linear (+Dropout mask) (+BatchNorm1d over the batch) (+ReLU)
// stacks on [B, C] features — STN3d / STNkd heads 1024->512->256->k*k (networks/PointNetCls.py:46-62,
// :89-101) and the classifier head 1024->512->256->1 (:208-213).
//
// These layers are 0.2 % of D4's FLOPs but, run as individual framework ops, about half of its kernel
// launches (GEMM + bias + three BatchNorm kernels + ReLU + counters, forward and backward).  Here one
// layer is ONE launch in each direction.  The trick is the partitioning: a CTA owns a few OUTPUT
// CHANNELS and all B rows, so everything BatchNorm needs (statistics over the batch, forward and
// backward) is CTA-local — no grid-wide synchronisation and no atomics, results are deterministic.
//
//   forward  (layer l):  y = (x W^T + b) * mask ; (mean, var over b) ; a = relu?(bn(y))
//   backward (layer l):  da = dy_{l+1} W_{l+1}   (pulled: this CTA's channels are columns of W_{l+1})
//                        dz = da * [a > 0] ; BN backward over b -> dy ; dy *= mask
//                        dW[c,:] = sum_b dy[b,c] x[b,:] ; db = sum_b dy ; dgamma, dbeta
//   input gradient:      dx = dy_0 W_0           (one more pull launch)
//
// Latency-bound by construction (B <= a few hundred rows): the kernels are written for few, wide,
// coalesced memory phases rather than for FLOPs.
#include <algorithm>

#include "pcuda_common.cuh"

namespace pcuda {
namespace {

constexpr int kCPB = 4;        // output channels per CTA
constexpr int kFcThreads = 256;
constexpr int kMaxCin = 4096;  // staged weight rows: kCPB * cin floats of shared memory

struct FcFwdParams {
  const float* x;        // [B, cin]
  const float* W;        // [cout, cin]
  const float* bias;     // [cout] or nullptr
  const float* mask;     // [B, cout] or nullptr
  const float* gamma; const float* beta;       // BN (nullptr: no BN)
  float* running_mean; float* running_var;
  float* save_mean; float* save_invstd;
  float* y;              // [B, cout] BN input (after mask); may be nullptr when there is no BN
  float* a;              // [B, cout] layer output
  int B, cin, cout, relu, train, iden_k;
  float momentum, eps;
};

// dynamic shared memory: Wsm[kCPB][cin] | ys[B][kCPB]
__global__ void __launch_bounds__(kFcThreads) fc_fwd_kernel(const FcFwdParams p) {
  extern __shared__ __align__(16) float sm[];
  float* Wsm = sm;
  float* ys = sm + kCPB * p.cin;
  __shared__ float s_mean[kCPB], s_scale[kCPB], s_shift[kCPB];
  const int c0 = blockIdx.x * kCPB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cin4 = p.cin >> 2;
  // the weight rows are parameters (nothing in flight writes them): staged while the previous kernel is still running
  for (int i = threadIdx.x; i < kCPB * cin4; i += kFcThreads) {
    const int cc = i / cin4, k4 = i - cc * cin4;
    const int c = c0 + cc;
    reinterpret_cast<float4*>(Wsm)[i] =
        c < p.cout ? __ldg(reinterpret_cast<const float4*>(p.W + static_cast<int64_t>(c) * p.cin) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_entry();
  __syncthreads();
  // a warp takes two rows at a time (b, b + 8): each row of x is read once for the kCPB channels and
  // 16 independent 128-bit loads are in flight per lane
  for (int b = warp; b < p.B; b += 2 * (kFcThreads / 32)) {
    const int b1 = b + kFcThreads / 32;
    const bool two = b1 < p.B;
    const float4* xr0 = reinterpret_cast<const float4*>(p.x + static_cast<int64_t>(b) * p.cin);
    const float4* xr1 = reinterpret_cast<const float4*>(p.x + static_cast<int64_t>(two ? b1 : b) * p.cin);
    float acc[2][kCPB] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll 8
    for (int k4 = lane; k4 < cin4; k4 += 32) {
      const float4 x0 = __ldg(xr0 + k4), x1 = __ldg(xr1 + k4);
#pragma unroll
      for (int cc = 0; cc < kCPB; ++cc) {
        const float4 wv = reinterpret_cast<const float4*>(Wsm)[cc * cin4 + k4];
        acc[0][cc] = fmaf(x0.x, wv.x, acc[0][cc]); acc[0][cc] = fmaf(x0.y, wv.y, acc[0][cc]);
        acc[0][cc] = fmaf(x0.z, wv.z, acc[0][cc]); acc[0][cc] = fmaf(x0.w, wv.w, acc[0][cc]);
        acc[1][cc] = fmaf(x1.x, wv.x, acc[1][cc]); acc[1][cc] = fmaf(x1.y, wv.y, acc[1][cc]);
        acc[1][cc] = fmaf(x1.z, wv.z, acc[1][cc]); acc[1][cc] = fmaf(x1.w, wv.w, acc[1][cc]);
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int cc = 0; cc < kCPB; ++cc) acc[h][cc] = warp_sum(acc[h][cc]);
    if (lane < 2 * kCPB) {
      const int h = lane / kCPB, cc = lane % kCPB, c = c0 + cc;
      const int bb = h == 0 ? b : b1;
      float v = 0.f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int j = 0; j < kCPB; ++j)
          if (hh == h && j == cc) v = acc[hh][j];
      if (h == 0 || two) {
        if (c < p.cout) {
          if (p.bias) v += p.bias[c];
          if (p.mask) v *= p.mask[static_cast<int64_t>(bb) * p.cout + c];
        }
        ys[bb * kCPB + cc] = v;
      }
    }
  }
  __syncthreads();
  // BatchNorm over the batch: warp cc owns channel c0 + cc (two-pass, fixed order: deterministic)
  if (warp < kCPB) {
    const int c = c0 + warp;
    float scale = 1.f, shift = 0.f, mean = 0.f;
    if (p.gamma != nullptr && c < p.cout) {
      float invstd;
      if (p.train) {
        float s = 0.f;
        for (int b = lane; b < p.B; b += 32) s += ys[b * kCPB + warp];
        mean = warp_sum(s) / static_cast<float>(p.B);
        float q = 0.f;
        for (int b = lane; b < p.B; b += 32) { const float d = ys[b * kCPB + warp] - mean; q = fmaf(d, d, q); }
        const float var = warp_sum(q) / static_cast<float>(p.B);
        invstd = rsqrtf(var + p.eps);
        if (lane == 0) {
          p.save_mean[c] = mean;
          p.save_invstd[c] = invstd;
          if (p.running_mean) p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
          if (p.running_var) {
            const float unb = p.B > 1 ? var * static_cast<float>(p.B) / static_cast<float>(p.B - 1) : var;
            p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * unb;
          }
        }
      } else {
        mean = p.running_mean[c];
        invstd = rsqrtf(p.running_var[c] + p.eps);
        if (lane == 0) { p.save_mean[c] = mean; p.save_invstd[c] = invstd; }
      }
      scale = invstd * p.gamma[c];
      shift = p.beta[c];
    }
    if (lane == 0) { s_mean[warp] = mean; s_scale[warp] = scale; s_shift[warp] = shift; }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < p.B * kCPB; i += kFcThreads) {
    const int b = i / kCPB, cc = i - b * kCPB, c = c0 + cc;
    if (c >= p.cout) continue;
    const float yv = ys[i];
    float z = fmaf(yv - s_mean[cc], s_scale[cc], s_shift[cc]);
    if (p.relu) z = fmaxf(z, 0.f);
    if (p.iden_k > 0 && c % (p.iden_k + 1) == 0) z += 1.0f;       // + flattened identity (STN heads)
    const int64_t o = static_cast<int64_t>(b) * p.cout + c;
    if (p.y) p.y[o] = yv;
    p.a[o] = z;
  }
}

struct FcBwdParams {
  // this layer
  const float* x;        // [B, cin] layer input
  const float* W;        // [cout, cin]   (unused here; the next launch pulls through it)
  const float* mask;     // [B, cout] or nullptr
  const float* gamma; const float* save_mean; const float* save_invstd;   // BN (gamma nullptr: none)
  const float* y;        // [B, cout] BN input
  const float* a;        // [B, cout] layer output (ReLU mask)
  // upstream: either grad_out (last layer) or the next layer's dy and weights
  const float* grad_out; // [B, cout] or nullptr
  const float* dy_next;  // [B, cout_next]
  const float* W_next;   // [cout_next, cout]
  int cout_next;
  // outputs
  float* dy;             // [B, cout] gradient w.r.t. the linear output (consumed by the next launch)
  float* grad_weight; float* grad_bias; float* grad_gamma; float* grad_beta;   // may be nullptr
  int B, cin, cout, relu, train;
};

// dynamic shared memory: dys[B][kCPB]
constexpr int kBwdPre = 16;   // rows of W_next per lane requested before the grid-dependency wait (cout_next <= 512)
__global__ void __launch_bounds__(kFcThreads) fc_bwd_kernel(const FcBwdParams p) {
  extern __shared__ __align__(16) float sm[];
  float* dys = sm;
  const int c0 = blockIdx.x * kCPB;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the next layer's weights (this CTA's kCPB columns of rows lane, lane + 32, ...) are parameters: in flight while the
  // kernel that produces dy_next is still running
  const bool vec = (p.cout & 3) == 0;
  const bool pre = p.grad_out == nullptr && vec;
  float4 wpre[kBwdPre];
#pragma unroll
  for (int j = 0; j < kBwdPre; ++j) {
    const int o = lane + 32 * j;
    wpre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pre && o < p.cout_next) wpre[j] = __ldg(reinterpret_cast<const float4*>(p.W_next + static_cast<int64_t>(o) * p.cout + c0));
  }
  pdl_entry();
  // (1) da for this CTA's channels
  if (p.grad_out != nullptr) {
    for (int i = threadIdx.x; i < p.B * kCPB; i += kFcThreads) {
      const int b = i / kCPB, c = c0 + (i - b * kCPB);
      dys[i] = c < p.cout ? p.grad_out[static_cast<int64_t>(b) * p.cout + c] : 0.f;
    }
  } else {
    for (int b = warp; b < p.B; b += 2 * (kFcThreads / 32)) {
      const int b1 = b + kFcThreads / 32;
      const bool two = b1 < p.B;
      const float* dn0 = p.dy_next + static_cast<int64_t>(b) * p.cout_next;
      const float* dn1 = p.dy_next + static_cast<int64_t>(two ? b1 : b) * p.cout_next;
      float acc[2][kCPB] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      if (pre) {
#pragma unroll
        for (int j = 0; j < kBwdPre; ++j) {
          const int o = lane + 32 * j;
          if (o < p.cout_next) {
            const float d0 = __ldg(dn0 + o), d1 = __ldg(dn1 + o);
            acc[0][0] = fmaf(d0, wpre[j].x, acc[0][0]); acc[0][1] = fmaf(d0, wpre[j].y, acc[0][1]);
            acc[0][2] = fmaf(d0, wpre[j].z, acc[0][2]); acc[0][3] = fmaf(d0, wpre[j].w, acc[0][3]);
            acc[1][0] = fmaf(d1, wpre[j].x, acc[1][0]); acc[1][1] = fmaf(d1, wpre[j].y, acc[1][1]);
            acc[1][2] = fmaf(d1, wpre[j].z, acc[1][2]); acc[1][3] = fmaf(d1, wpre[j].w, acc[1][3]);
          }
        }
      }
#pragma unroll 8
      for (int o = lane + (pre ? 32 * kBwdPre : 0); o < p.cout_next; o += 32) {
        const float d0 = __ldg(dn0 + o), d1 = __ldg(dn1 + o);
        const float* wr = p.W_next + static_cast<int64_t>(o) * p.cout + c0;
        float wv[kCPB];
        if (vec) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(wr));
          wv[0] = w4.x; wv[1] = w4.y; wv[2] = w4.z; wv[3] = w4.w;
        } else {
#pragma unroll
          for (int cc = 0; cc < kCPB; ++cc) wv[cc] = c0 + cc < p.cout ? __ldg(wr + cc) : 0.f;
        }
#pragma unroll
        for (int cc = 0; cc < kCPB; ++cc) { acc[0][cc] = fmaf(d0, wv[cc], acc[0][cc]); acc[1][cc] = fmaf(d1, wv[cc], acc[1][cc]); }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int cc = 0; cc < kCPB; ++cc) acc[h][cc] = warp_sum(acc[h][cc]);
      if (lane < 2 * kCPB) {
        const int h = lane / kCPB, cc = lane % kCPB;
        float v = 0.f;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int j = 0; j < kCPB; ++j)
            if (hh == h && j == cc) v = acc[hh][j];
        if (h == 0) dys[b * kCPB + cc] = v;
        else if (two) dys[b1 * kCPB + cc] = v;
      }
    }
  }
  __syncthreads();
  // (2) ReLU mask, (3) BatchNorm backward over the batch, (4) dropout mask: warp cc owns channel c0 + cc
  if (warp < kCPB) {
    const int c = c0 + warp;
    if (c < p.cout) {
      const bool bn = p.gamma != nullptr;
      const float mean = bn ? p.save_mean[c] : 0.f, invstd = bn ? p.save_invstd[c] : 1.f;
      float sb = 0.f, sg = 0.f;
      for (int b = lane; b < p.B; b += 32) {
        const int64_t o = static_cast<int64_t>(b) * p.cout + c;
        float dz = dys[b * kCPB + warp];
        if (p.relu && !(p.a[o] > 0.f)) dz = 0.f;
        dys[b * kCPB + warp] = dz;
        if (bn) {
          sb += dz;
          sg = fmaf(dz, (p.y[o] - mean) * invstd, sg);
        }
      }
      float dbias = 0.f;
      if (bn) {
        sb = warp_sum(sb);
        sg = warp_sum(sg);
        const float g = p.gamma[c];
        const float kb = p.train ? sb / static_cast<float>(p.B) : 0.f;
        const float kg = p.train ? sg / static_cast<float>(p.B) : 0.f;
        for (int b = lane; b < p.B; b += 32) {
          const int64_t o = static_cast<int64_t>(b) * p.cout + c;
          const float yh = (p.y[o] - mean) * invstd;
          float d = g * invstd * (dys[b * kCPB + warp] - kb - yh * kg);
          if (p.mask) d *= p.mask[o];
          dys[b * kCPB + warp] = d;
          dbias += d;
        }
        if (lane == 0) {
          if (p.grad_gamma) p.grad_gamma[c] = sg;
          if (p.grad_beta) p.grad_beta[c] = sb;
        }
      } else {
        for (int b = lane; b < p.B; b += 32) {
          float d = dys[b * kCPB + warp];
          if (p.mask) d *= p.mask[static_cast<int64_t>(b) * p.cout + c];
          dys[b * kCPB + warp] = d;
          dbias += d;
        }
      }
      dbias = warp_sum(dbias);
      if (lane == 0 && p.grad_bias) p.grad_bias[c] = dbias;
    }
  }
  __syncthreads();
  // (5) dy for the next launch
  for (int i = threadIdx.x; i < p.B * kCPB; i += kFcThreads) {
    const int b = i / kCPB, c = c0 + (i - b * kCPB);
    if (c < p.cout) p.dy[static_cast<int64_t>(b) * p.cout + c] = dys[i];
  }
  // (6) weight gradient rows of this CTA's channels
  if (p.grad_weight != nullptr) {
    for (int k = threadIdx.x; k < p.cin; k += kFcThreads) {
      float acc[kCPB] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
      for (int b = 0; b < p.B; ++b) {
        const float xv = __ldg(p.x + static_cast<int64_t>(b) * p.cin + k);
#pragma unroll
        for (int cc = 0; cc < kCPB; ++cc) acc[cc] = fmaf(dys[b * kCPB + cc], xv, acc[cc]);
      }
#pragma unroll
      for (int cc = 0; cc < kCPB; ++cc)
        if (c0 + cc < p.cout) p.grad_weight[static_cast<int64_t>(c0 + cc) * p.cin + k] = acc[cc];
    }
  }
}

// dx[b, k] = sum_c dy[b, c] W[c, k].  A CTA owns 32 input channels k (one per lane: coalesced 128-byte
// segments of the rows of W); its 8 warps split the contraction index c and are combined in a fixed
// order through shared memory; kRB batch rows per pass, their dy values staged as [c][row] so a warp
// reads them as broadcast float4.
template <int RB, int kPullThreads>
__global__ void __launch_bounds__(kPullThreads) fc_pull_kernel(const float* __restrict__ dy, const float* __restrict__ W,
                                                              int B, int cin, int cout, float* __restrict__ dx) {
  extern __shared__ __align__(16) float sm[];   // dys[cout][RB], reused as red[warps][RB][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane;
  // this thread's first kPullPre weights (rows warp, warp + nwarps, ...: parameters) are requested before the wait
  constexpr int kPullPre = 16, kWarps = kPullThreads / 32;
  float wpre[kPullPre];
#pragma unroll
  for (int j = 0; j < kPullPre; ++j) {
    const int c = warp + kWarps * j;
    wpre[j] = (c < cout && k < cin) ? __ldg(W + static_cast<int64_t>(c) * cin + k) : 0.f;
  }
  pdl_entry();
  for (int b0 = 0; b0 < B; b0 += RB) {
    const int nb = min(RB, B - b0);
    __syncthreads();
    for (int i = threadIdx.x; i < cout * RB; i += kPullThreads) {   // coalesced global reads (c fastest), transposed store
      const int r = i / cout, c = i - r * cout;
      sm[c * RB + r] = r < nb ? dy[static_cast<int64_t>(b0 + r) * cout + c] : 0.f;
    }
    __syncthreads();
    float acc[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) acc[r] = 0.f;
    auto row = [&](int c, float wv) {
      const float4* d4 = reinterpret_cast<const float4*>(sm + c * RB);
#pragma unroll
      for (int r4 = 0; r4 < RB / 4; ++r4) {
        const float4 d = d4[r4];
        acc[4 * r4 + 0] = fmaf(d.x, wv, acc[4 * r4 + 0]); acc[4 * r4 + 1] = fmaf(d.y, wv, acc[4 * r4 + 1]);
        acc[4 * r4 + 2] = fmaf(d.z, wv, acc[4 * r4 + 2]); acc[4 * r4 + 3] = fmaf(d.w, wv, acc[4 * r4 + 3]);
      }
    };
#pragma unroll
    for (int j = 0; j < kPullPre; ++j)
      if (warp + kWarps * j < cout) row(warp + kWarps * j, wpre[j]);
#pragma unroll 8
    for (int c = warp + kWarps * kPullPre; c < cout; c += kWarps)
      row(c, k < cin ? __ldg(W + static_cast<int64_t>(c) * cin + k) : 0.f);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RB; ++r) sm[(warp * RB + r) * 32 + lane] = acc[r];
    __syncthreads();
    for (int i = threadIdx.x; i < nb * 32; i += kPullThreads) {
      const int r = i >> 5, l = i & 31;
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kPullThreads / 32; ++w) t += sm[(w * RB + r) * 32 + l];
      if (blockIdx.x * 32 + l < cin) dx[static_cast<int64_t>(b0 + r) * cin + blockIdx.x * 32 + l] = t;
    }
  }
}

constexpr int kMaxBnItems = 64;
constexpr int kMaxBnGroups = 32;
constexpr int kMaxPerGroup = 8;
// items that update the same buffers form a group (applied in order, in registers); groups are independent
struct BnGroups {
  struct G {
    float* running_mean; float* running_var;
    const float* save_mean[kMaxPerGroup]; const float* save_invstd[kMaxPerGroup];
    double count[kMaxPerGroup];
    float momentum[kMaxPerGroup], eps[kMaxPerGroup];
    int C, n;
  } g[kMaxBnGroups];
};
// blockIdx.y = group, thread = channel: every load of the group is issued up front, the updates compose
// in registers exactly like sequential forward passes, one store per buffer
constexpr int kBnGroupsPerLaunch = 12;        // 12 x 280 bytes of kernel parameters
struct BnLaunch { BnGroups::G g[kBnGroupsPerLaunch]; };
__global__ void bn_running_update_group_kernel(const BnLaunch a, int ngroups) {
  pdl_entry();
  const BnGroups::G& u = a.g[blockIdx.y];
  if (static_cast<int>(blockIdx.y) >= ngroups) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= u.C) return;
  float sm[kMaxPerGroup], is[kMaxPerGroup];
#pragma unroll
  for (int i = 0; i < kMaxPerGroup; ++i) {
    sm[i] = i < u.n ? u.save_mean[i][c] : 0.f;
    is[i] = i < u.n ? u.save_invstd[i][c] : 1.f;
  }
  float rm = u.running_mean[c], rv = u.running_var[c];
#pragma unroll
  for (int i = 0; i < kMaxPerGroup; ++i) {
    if (i >= u.n) break;
    const float m = u.momentum[i];
    const double isd = static_cast<double>(is[i]);
    double var = 1.0 / (isd * isd) - static_cast<double>(u.eps[i]);
    var = var > 0.0 ? var : 0.0;
    const double unb = u.count[i] > 1.0 ? var * (u.count[i] / (u.count[i] - 1.0)) : var;
    rm = (1.f - m) * rm + m * sm[i];
    rv = (1.f - m) * rv + m * static_cast<float>(unb);
  }
  u.running_mean[c] = rm;
  u.running_var[c] = rv;
}

int validate_fc(const char* who, int B, int L, const pcuda_fc_layer_t* layers) {
  PCUDA_REQUIRE(layers != nullptr, PCUDA_E_NULL, "%s: layers is NULL", who);
  PCUDA_REQUIRE(B >= 1 && L >= 1 && L <= 8, PCUDA_E_SHAPE, "%s: bad shape B=%d L=%d", who, B, L);
  PCUDA_REQUIRE(B <= 2048, PCUDA_E_UNSUPPORTED, "%s: B=%d > 2048 rows", who, B);
  for (int l = 0; l < L; ++l) {
    const pcuda_fc_layer_t& y = layers[l];
    PCUDA_REQUIRE(y.cin >= 1 && y.cout >= 1, PCUDA_E_SHAPE, "%s: layer %d has cin=%d cout=%d", who, l, y.cin, y.cout);
    PCUDA_REQUIRE(l == 0 || y.cin == layers[l - 1].cout, PCUDA_E_SHAPE, "%s: layer %d cin=%d != previous cout=%d", who, l, y.cin, layers[l - 1].cout);
    PCUDA_REQUIRE((y.cin & 3) == 0 && y.cin <= kMaxCin, PCUDA_E_UNSUPPORTED, "%s: layer %d cin=%d must be a multiple of 4, <= %d", who, l, y.cin, kMaxCin);
    PCUDA_REQUIRE(y.weight && y.a, PCUDA_E_NULL, "%s: layer %d needs weight and a", who, l);
    PCUDA_REQUIRE(aligned16(y.weight), PCUDA_E_ALIGN, "%s: layer %d weight must be 16-byte aligned", who, l);
    if (y.bn) PCUDA_REQUIRE(y.gamma && y.beta && y.save_mean && y.save_invstd && y.y, PCUDA_E_NULL, "%s: layer %d BatchNorm tensors missing", who, l);
  }
  return 0;
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" size_t pcuda_fcstack_ws_bytes(int B, int L, const pcuda_fc_layer_t* layers, int backward) {
  if (!backward || B < 1 || L < 1 || !layers) return 16;
  size_t mx = 0;
  for (int l = 0; l < L; ++l) mx = mx > static_cast<size_t>(layers[l].cout) ? mx : static_cast<size_t>(layers[l].cout);
  return 2 * sizeof(float) * static_cast<size_t>(B) * mx + 16;   // dy ping-pong
}

extern "C" int pcuda_fcstack_fwd(const float* x, int B, int L, const pcuda_fc_layer_t* layers, int train, float momentum,
                                 float eps, int add_identity_k, pcuda_stream_t stream) {
  if (int rc = validate_fc("fcstack_fwd", B, L, layers)) return rc;
  PCUDA_REQUIRE(x != nullptr && aligned16(x), PCUDA_E_NULL, "fcstack_fwd: x is NULL or not 16-byte aligned");
  for (int l = 0; l < L; ++l) {
    PCUDA_REQUIRE(!layers[l].bn || train || (layers[l].running_mean && layers[l].running_var), PCUDA_E_NULL,
                  "fcstack_fwd: eval mode needs running stats (layer %d)", l);
    PCUDA_REQUIRE(!(layers[l].bn && train) || B > 1, PCUDA_E_SHAPE, "fcstack_fwd: train-mode BatchNorm needs more than 1 row");
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  smem_optin(fc_fwd_kernel, 160 * 1024);
  const float* in = x;
  for (int l = 0; l < L; ++l) {
    const pcuda_fc_layer_t& y = layers[l];
    FcFwdParams p{};
    p.x = in; p.W = y.weight; p.bias = y.bias; p.mask = y.mask;
    p.gamma = y.bn ? y.gamma : nullptr; p.beta = y.beta;
    p.running_mean = y.running_mean; p.running_var = y.running_var; p.save_mean = y.save_mean; p.save_invstd = y.save_invstd;
    p.y = y.y; p.a = y.a; p.B = B; p.cin = y.cin; p.cout = y.cout; p.relu = y.relu; p.train = train;
    p.iden_k = l == L - 1 ? add_identity_k : 0;
    p.momentum = momentum; p.eps = eps;
    const size_t smem = sizeof(float) * (static_cast<size_t>(kCPB) * y.cin + static_cast<size_t>(B) * kCPB);
    PCUDA_REQUIRE(smem <= 160 * 1024, PCUDA_E_UNSUPPORTED, "fcstack_fwd: layer %d does not fit shared memory", l);
    PCUDA_LAUNCH_PDL(fc_fwd_kernel, (y.cout + kCPB - 1) / kCPB, kFcThreads, smem, st, p);
    in = y.a;
  }
  count_launch(L);
  return check_launch("fcstack_fwd");
}

extern "C" int pcuda_fcstack_bwd(const float* x, int B, int L, const pcuda_fc_layer_t* layers, int train,
                                 const float* grad_out, float* grad_x, void* ws, pcuda_stream_t stream) {
  if (int rc = validate_fc("fcstack_bwd", B, L, layers)) return rc;
  PCUDA_REQUIRE(x && grad_out && ws, PCUDA_E_NULL, "fcstack_bwd: NULL x/grad_out/ws");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  size_t mx = 0;
  for (int l = 0; l < L; ++l) mx = mx > static_cast<size_t>(layers[l].cout) ? mx : static_cast<size_t>(layers[l].cout);
  float* dybuf[2] = {static_cast<float*>(ws), static_cast<float*>(ws) + static_cast<size_t>(B) * mx};
  int cur = 0, launches = 0;
  for (int l = L - 1; l >= 0; --l) {
    const pcuda_fc_layer_t& y = layers[l];
    FcBwdParams p{};
    p.x = l == 0 ? x : layers[l - 1].a; p.W = y.weight; p.mask = y.mask;
    p.gamma = y.bn ? y.gamma : nullptr; p.save_mean = y.save_mean; p.save_invstd = y.save_invstd; p.y = y.y; p.a = y.a;
    if (l == L - 1) p.grad_out = grad_out;
    else { p.dy_next = dybuf[cur ^ 1]; p.W_next = layers[l + 1].weight; p.cout_next = layers[l + 1].cout; }
    p.dy = dybuf[cur];
    p.grad_weight = y.grad_weight; p.grad_bias = y.grad_weight ? y.grad_bias : nullptr;
    p.grad_gamma = y.grad_weight ? y.grad_gamma : nullptr; p.grad_beta = y.grad_weight ? y.grad_beta : nullptr;
    p.B = B; p.cin = y.cin; p.cout = y.cout; p.relu = y.relu; p.train = train;
    PCUDA_LAUNCH_PDL(fc_bwd_kernel, (y.cout + kCPB - 1) / kCPB, kFcThreads, sizeof(float) * static_cast<size_t>(B) * kCPB, st, p);
    ++launches;
    cur ^= 1;
  }
  if (grad_x != nullptr) {
    const pcuda_fc_layer_t& y = layers[0];
    // rows per pass: 32 when the batch has more than 16 rows (one pass over W for B <= 32); the contraction
    // index is split over 32 (16) warps so that a warp walks only cout/32 (cout/16) rows of W
    const int rb = B > 16 ? 32 : 16;
    const int threads = rb == 32 ? 512 : 1024;
    const size_t smem = sizeof(float) * static_cast<size_t>(rb) * std::max(y.cout, threads);
    PCUDA_REQUIRE(smem <= 160 * 1024, PCUDA_E_UNSUPPORTED, "fcstack_bwd: first layer too wide for the input-gradient kernel");
    smem_optin(fc_pull_kernel<32, 512>, 160 * 1024);
    smem_optin(fc_pull_kernel<16, 1024>, 160 * 1024);
    if (rb == 32) PCUDA_LAUNCH_PDL((fc_pull_kernel<32, 512>), (y.cin + 31) / 32, 512, smem, st, dybuf[cur ^ 1], y.weight, B, y.cin, y.cout, grad_x);
    else PCUDA_LAUNCH_PDL((fc_pull_kernel<16, 1024>), (y.cin + 31) / 32, 1024, smem, st, dybuf[cur ^ 1], y.weight, B, y.cin, y.cout, grad_x);
    ++launches;
  }
  count_launch(launches);
  return check_launch("fcstack_bwd");
}

extern "C" int pcuda_bn_running_update(int n, const pcuda_bn_update_t* items, pcuda_stream_t stream) {
  PCUDA_REQUIRE(n >= 0 && n <= kMaxBnItems, PCUDA_E_SHAPE, "bn_running_update: n=%d not in [0, %d]", n, kMaxBnItems);
  if (n == 0) return 0;
  PCUDA_REQUIRE(items != nullptr, PCUDA_E_NULL, "bn_running_update: items is NULL");
  // group the items by target buffer, keeping their order inside a group
  static thread_local BnGroups gs;
  int ng = 0;
  for (int i = 0; i < n; ++i) {
    PCUDA_REQUIRE(items[i].running_mean && items[i].running_var && items[i].save_mean && items[i].save_invstd, PCUDA_E_NULL,
                  "bn_running_update: item %d has a NULL pointer", i);
    PCUDA_REQUIRE(items[i].C >= 1, PCUDA_E_SHAPE, "bn_running_update: item %d has C=%d", i, items[i].C);
    int g = 0;
    while (g < ng && gs.g[g].running_mean != items[i].running_mean) ++g;
    if (g == ng) {
      PCUDA_REQUIRE(ng < kMaxBnGroups, PCUDA_E_UNSUPPORTED, "bn_running_update: more than %d distinct buffers", kMaxBnGroups);
      gs.g[g].running_mean = items[i].running_mean; gs.g[g].running_var = items[i].running_var;
      gs.g[g].C = items[i].C; gs.g[g].n = 0;
      ++ng;
    }
    BnGroups::G& G = gs.g[g];
    PCUDA_REQUIRE(G.n < kMaxPerGroup && G.C == items[i].C && G.running_var == items[i].running_var, PCUDA_E_UNSUPPORTED,
                  "bn_running_update: item %d: more than %d updates of one buffer, or inconsistent buffers", i, kMaxPerGroup);
    G.save_mean[G.n] = items[i].save_mean; G.save_invstd[G.n] = items[i].save_invstd;
    G.count[G.n] = items[i].count; G.momentum[G.n] = items[i].momentum; G.eps[G.n] = items[i].eps;
    ++G.n;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int launches = 0;
  for (int g0 = 0; g0 < ng; g0 += kBnGroupsPerLaunch) {
    static thread_local BnLaunch a;
    int maxc = 0;
    const int cnt = ng - g0 < kBnGroupsPerLaunch ? ng - g0 : kBnGroupsPerLaunch;
    for (int k = 0; k < cnt; ++k) { a.g[k] = gs.g[g0 + k]; maxc = maxc > a.g[k].C ? maxc : a.g[k].C; }
    PCUDA_LAUNCH(bn_running_update_group_kernel, dim3((maxc + 255) / 256, cnt), 256, 0, st, a, cnt);
    ++launches;
  }
  count_launch(launches);
  return check_launch("bn_running_update");
}
