// Shared between pointmlp.cu (FP32 CUDA-core kernels) and pointmlp_tc.cu (tcgen05 kernels):
// how an activation / a pre-BN gradient is addressed and re-materialised on load.
#pragma once
#include "pcuda_common.cuh"

namespace pcuda {

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
  // raw input only: per-cloud input transform T[b] (row-major [C, C], C <= 4) applied on load,
  // a[m, k] = sum_j x[b, j, n] T[b][j][k]  — torch.bmm(x.transpose(2, 1), trans) of PointNetfeat.forward
  // (networks/PointNetCls.py:140-142) without materialising the transformed cloud.  nullptr: identity.
  const float* trans;
};

// the C <= 4 raw channels of point (b, n), transformed when the source carries an input transform
__device__ __forceinline__ void load_raw_point(const ActSrc& s, int64_t b, int64_t n, float (&v)[4]) {
  const float* px = s.x + b * s.sxb + n * s.sxn;
  float r[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r[j] = j < s.C ? __ldg(px + j * s.sxc) : 0.f;
  if (s.trans == nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = r[j];
    return;
  }
  const float* T = s.trans + b * s.C * s.C;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float acc = 0.f;
    if (k < s.C) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < s.C) acc = fmaf(r[j], __ldg(T + j * s.C + k), acc);     // j ascending: the order of a K = 3 dot product
    }
    v[k] = acc;
  }
}

// Finalise-on-read: the raw BatchNorm sums of the producing layer, handed to the kernel that consumes its
// output (the operand packer of the next layer, the pooled-output kernel) so that no one-block
// "finalise" launch sits between a GEMM and its consumer.  stats == nullptr: save_mean / save_invstd of the
// ActSrc are already final.  Same arithmetic as bn_finalize_pivot_kernel.
struct BnRaw {
  const double* stats;       // [2C]: sum and sum of squares of (y - pivot)
  const float* pivot;        // [C]
  double count;
  float eps, momentum;
  int train;
  float* save_mean;          // [C] written by the consumer's first CTA (the backward pass reads them)
  float* save_invstd;
  float* running_mean;       // [C] or nullptr (deferred / frozen): updated by the consumer's first CTA
  float* running_var;
};

__device__ __forceinline__ void bn_raw_channel(const BnRaw& r, int c, int C, bool writer, float& mean_f, float& invstd_f) {
  if (r.train) {
    const double d = r.stats[c] / r.count;
    const double mean = static_cast<double>(r.pivot[c]) + d;
    double var = r.stats[C + c] / r.count - d * d;
    var = var > 0.0 ? var : 0.0;
    mean_f = static_cast<float>(mean);
    invstd_f = static_cast<float>(1.0 / sqrt(var + static_cast<double>(r.eps)));
    if (writer) {
      if (r.running_mean) r.running_mean[c] = (1.0f - r.momentum) * r.running_mean[c] + r.momentum * static_cast<float>(mean);
      if (r.running_var) {
        const double unbiased = var * (r.count / (r.count - 1.0));
        r.running_var[c] = (1.0f - r.momentum) * r.running_var[c] + r.momentum * static_cast<float>(unbiased);
      }
    }
  } else {
    mean_f = r.running_mean[c];
    invstd_f = static_cast<float>(1.0 / sqrt(static_cast<double>(r.running_var[c]) + static_cast<double>(r.eps)));
  }
  if (writer) { r.save_mean[c] = mean_f; r.save_invstd[c] = invstd_f; }
}

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
    if (s.trans != nullptr) {        // C <= 4, so k == 0
      float t[4];
      load_raw_point(s, b, n, t);
      return make_float4(t[0], t[1], t[2], t[3]);
    }
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
  if (s.trans != nullptr) {
    float t[4];
    load_raw_point(s, b, n, t);
    return k == 0 ? t[0] : (k == 1 ? t[1] : (k == 2 ? t[2] : t[3]));
  }
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

// The same for a thread that keeps its 4 channels over many points (narrow-layer kernels): the per-channel
// constants live in registers, dz and y move as 128-bit loads.  dy = dz*p0 + (y*p1 + p2), p0 = gamma*invstd,
// p1 = -kappa, p2 = kappa*mean - alpha (the form pack_dy_kernel uses).  Needs C % 4 == 0 and c0 % 4 == 0.
struct DyConst4 {
  float p0[4], p1[4], p2[4];
};
__device__ __forceinline__ DyConst4 dy_const4(const DySrc& s, int c0) {
  DyConst4 k;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cc = c0 + i;
    const bool ok = cc < s.C;
    k.p0[i] = ok ? s.gamma[cc] * s.invstd[cc] : 0.f;
    k.p1[i] = ok ? -s.kappa[cc] : 0.f;
    k.p2[i] = ok ? s.kappa[cc] * s.mean[cc] - s.alpha[cc] : 0.f;
  }
  return k;
}
__device__ __forceinline__ float4 load_dy4c(const DySrc& s, const DyConst4& k, int64_t m, int c0) {
  const float4 dz = *reinterpret_cast<const float4*>(s.dz + m * s.C + c0);
  const float4 y = *reinterpret_cast<const float4*>(s.y + m * s.C + c0);
  return make_float4(fmaf(dz.x, k.p0[0], fmaf(y.x, k.p1[0], k.p2[0])), fmaf(dz.y, k.p0[1], fmaf(y.y, k.p1[1], k.p2[1])),
                     fmaf(dz.z, k.p0[2], fmaf(y.z, k.p1[2], k.p2[2])), fmaf(dz.w, k.p0[3], fmaf(y.w, k.p1[3], k.p2[3])));
}


// Target of a dgrad epilogue: val -> (ReLU mask of the producing layer) -> store
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

// Order-preserving (value, first-index) key for the fused global max-pool: larger value wins,
// then the smaller point index.
__device__ __forceinline__ unsigned long long pool_key(float v, int n) {
  return (static_cast<unsigned long long>(float_to_ordered(v)) << 32) |
         static_cast<unsigned long long>(0xFFFFFFFFu - static_cast<unsigned int>(n));
}

}  // namespace pcuda
