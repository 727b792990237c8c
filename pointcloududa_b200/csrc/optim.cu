// Parameter-gradient packing and the SGD step of the point-cloud discriminator (D4), one launch each.
//
// The reference sums D4's parameter gradients implicitly (two loss.backward() calls accumulate into
// .grad, train_mscmrseg.py:288,319) and then runs torch.optim.SGD(momentum .99, weight_decay 5e-4)
// (train_mscmrseg.py:329-330,:450-455).  The two D4 passes of this implementation run concurrently and
// return separate gradient tensors, so the sum is explicit — as framework ops it was two multi-tensor
// launches plus a two-launch fused SGD, ~65 us of serial tail behind the join of a 650 us step.  Here:
//   pcuda_grad_sum_pack      flat[off_i + j] = scale * (ga_i[j] + gb_i[j])      (the all-reduce bucket)
//   pcuda_sgd_momentum_step  g = flat[off_i + j] + wd * p;  buf = mu * buf + g;  p -= lr * buf
// The per-tensor pointers travel as kernel parameters (<= 48 tensors per launch), so both calls are
// legal under CUDA-graph capture; lr is read from device memory so a schedule needs no re-capture.
// Arithmetic order is torch's (_single_tensor_sgd / the fused functor): weight decay added to the
// gradient, buf = mu*buf + g in two roundings, p = p - lr*buf; buf starts at zero, which reproduces
// torch's first step (buf = g) exactly.
#include "pcuda_common.cuh"

namespace pcuda {
namespace {

constexpr int kSlotsPerLaunch = 48;
constexpr int kThreads = 256;
constexpr int kPerThread = 8;
constexpr int kChunk = kThreads * kPerThread;

struct Slot {
  const float* ga;
  const float* gb;
  float* param;
  long long offset;
  long long numel;
};
struct Table {
  Slot s[kSlotsPerLaunch];
  int cta_start[kSlotsPerLaunch + 1];
  int n;
};

__device__ __forceinline__ int find_slot(const Table& t, int cta) {
  int i = 0;
  while (i + 1 < t.n && cta >= t.cta_start[i + 1]) ++i;
  return i;
}

__global__ void __launch_bounds__(kThreads) grad_sum_pack_kernel(const __grid_constant__ Table t, float scale, float* __restrict__ flat) {
  pdl_entry();
  const int i = find_slot(t, blockIdx.x);
  const Slot& s = t.s[i];
  const long long base = static_cast<long long>(blockIdx.x - t.cta_start[i]) * kChunk;
#pragma unroll
  for (int u = 0; u < kPerThread; ++u) {
    const long long j = base + u * kThreads + threadIdx.x;
    if (j < s.numel) {
      float g = s.ga[j];
      if (s.gb != nullptr) g = __fadd_rn(g, s.gb[j]);
      flat[s.offset + j] = scale == 1.0f ? g : g * scale;
    }
  }
}

// The two kernels that WRITE PARAMETERS never release their successor early (pdl_wait() only, no launch_dependents):
// other kernels stage weights before their own grid-dependency wait on the grounds that nothing in flight writes them
// (pcuda_common.cuh), which must also hold for a forward pass launched right behind an optimiser step.
__global__ void __launch_bounds__(kThreads) sgd_momentum_kernel(const __grid_constant__ Table t, const float* __restrict__ flat_grad,
                                                                float* __restrict__ flat_mom, const float* __restrict__ lr_dev,
                                                                float momentum, float weight_decay) {
  pdl_wait();
  const int i = find_slot(t, blockIdx.x);
  const Slot& s = t.s[i];
  const float lr = __ldg(lr_dev);
  const long long base = static_cast<long long>(blockIdx.x - t.cta_start[i]) * kChunk;
#pragma unroll
  for (int u = 0; u < kPerThread; ++u) {
    const long long j = base + u * kThreads + threadIdx.x;
    if (j < s.numel) {
      const float p = s.param[j];
      float g = flat_grad[s.offset + j];
      if (weight_decay != 0.0f) g = __fmaf_rn(weight_decay, p, g);              // torch: grad.add(param, alpha=wd)
      const float buf = __fadd_rn(__fmul_rn(momentum, flat_mom[s.offset + j]), g);   // buf.mul_(mu).add_(g)
      flat_mom[s.offset + j] = buf;
      s.param[j] = __fmaf_rn(-lr, buf, p);                                       // param.add_(buf, alpha=-lr)
    }
  }
}

// single-process step (no exchange between the sum and the update): both of the above in one pass, the summed gradient
// still written to the bucket (flat_grad) for whoever inspects it.  Same roundings as the two-launch path.
__global__ void __launch_bounds__(kThreads) sgd_momentum_sum_kernel(const __grid_constant__ Table t, float scale, float* __restrict__ flat_grad,
                                                                    float* __restrict__ flat_mom, const float* __restrict__ lr_dev,
                                                                    float momentum, float weight_decay) {
  pdl_wait();
  const int i = find_slot(t, blockIdx.x);
  const Slot& s = t.s[i];
  const float lr = __ldg(lr_dev);
  const long long base = static_cast<long long>(blockIdx.x - t.cta_start[i]) * kChunk;
#pragma unroll
  for (int u = 0; u < kPerThread; ++u) {
    const long long j = base + u * kThreads + threadIdx.x;
    if (j < s.numel) {
      float g = s.ga[j];
      if (s.gb != nullptr) g = __fadd_rn(g, s.gb[j]);
      if (scale != 1.0f) g = __fmul_rn(g, scale);
      flat_grad[s.offset + j] = g;
      const float p = s.param[j];
      if (weight_decay != 0.0f) g = __fmaf_rn(weight_decay, p, g);
      const float buf = __fadd_rn(__fmul_rn(momentum, flat_mom[s.offset + j]), g);
      flat_mom[s.offset + j] = buf;
      s.param[j] = __fmaf_rn(-lr, buf, p);
    }
  }
}

// ---- discriminator bookkeeping: BCE-with-logits (mean), its gradient and the accuracy, one warp ----------
// loss = weight * mean_i [ (1 - t) x_i - log_sigmoid(x_i) ]   (F.binary_cross_entropy_with_logits, reduction 'mean',
// train_mscmrseg.py:233,286,316), grad_i = weight * (sigmoid(x_i) - t) / n, acc = mean_i [ (sigmoid(x_i) >= 0.5) == (t >= 0.5) ]
// (train_mscmrseg.py:290-296,:320-322).  n is the batch (<= a few hundred): one warp, fixed-order sums.
__global__ void __launch_bounds__(32) bce_logits_kernel(const float* __restrict__ x, int n, float t, float weight,
                                                        float* __restrict__ loss, float* __restrict__ grad, float* __restrict__ acc) {
  pdl_entry();
  double ls = 0.0;
  int hits = 0;
  for (int i = threadIdx.x; i < n; i += 32) {
    const float v = x[i];
    // log_sigmoid(v) = min(v, 0) - log1p(exp(-|v|))
    const float lsg = fminf(v, 0.f) - log1pf(expf(-fabsf(v)));
    ls += static_cast<double>((1.f - t) * v - lsg);
    const float sg = 1.f / (1.f + expf(-v));
    if (grad != nullptr) grad[i] = weight * (sg - t) / static_cast<float>(n);
    hits += ((sg >= 0.5f) == (t >= 0.5f)) ? 1 : 0;
  }
  ls = warp_sum(ls);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
  if (threadIdx.x == 0) {
    *loss = weight * static_cast<float>(ls / static_cast<double>(n));
    if (acc != nullptr) *acc = static_cast<float>(hits) / static_cast<float>(n);
  }
}

template <typename F>
int for_each_table(const pcuda_param_slot_t* slots, int n, const char* who, bool need_param, F&& launch) {
  for (int s0 = 0; s0 < n; s0 += kSlotsPerLaunch) {
    Table t{};
    t.n = n - s0 < kSlotsPerLaunch ? n - s0 : kSlotsPerLaunch;
    int ctas = 0;
    for (int i = 0; i < t.n; ++i) {
      const pcuda_param_slot_t& p = slots[s0 + i];
      PCUDA_REQUIRE(p.numel >= 0 && p.offset >= 0, PCUDA_E_SHAPE, "%s: slot %d has numel %lld offset %lld", who, s0 + i,
                    static_cast<long long>(p.numel), static_cast<long long>(p.offset));
      PCUDA_REQUIRE(p.numel == 0 || (need_param ? p.param != nullptr : p.grad_a != nullptr), PCUDA_E_NULL, "%s: slot %d has a NULL tensor", who, s0 + i);
      t.s[i] = Slot{p.grad_a, p.grad_b, p.param, static_cast<long long>(p.offset), static_cast<long long>(p.numel)};
      t.cta_start[i] = ctas;
      ctas += static_cast<int>((p.numel + kChunk - 1) / kChunk);
    }
    t.cta_start[t.n] = ctas;
    if (ctas == 0) continue;
    launch(t, ctas);
    count_launch(1);
    if (int rc = check_launch(who)) return rc;
  }
  return 0;
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" int pcuda_grad_sum_pack(const pcuda_param_slot_t* slots, int n, float scale, float* flat, pcuda_stream_t stream) {
  PCUDA_REQUIRE(n >= 0, PCUDA_E_SHAPE, "grad_sum_pack: n=%d", n);
  if (n == 0) return 0;
  PCUDA_REQUIRE(slots && flat, PCUDA_E_NULL, "grad_sum_pack: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return for_each_table(slots, n, "grad_sum_pack", false, [&](const Table& t, int ctas) {
    PCUDA_LAUNCH(grad_sum_pack_kernel, ctas, kThreads, 0, st, t, scale, flat);
  });
}

extern "C" int pcuda_sgd_momentum_step(const pcuda_param_slot_t* slots, int n, const float* flat_grad, float* flat_momentum,
                                       const float* lr_dev, float momentum, float weight_decay, pcuda_stream_t stream) {
  PCUDA_REQUIRE(n >= 0, PCUDA_E_SHAPE, "sgd_momentum_step: n=%d", n);
  if (n == 0) return 0;
  PCUDA_REQUIRE(slots && flat_grad && flat_momentum && lr_dev, PCUDA_E_NULL, "sgd_momentum_step: NULL argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return for_each_table(slots, n, "sgd_momentum_step", true, [&](const Table& t, int ctas) {
    PCUDA_LAUNCH(sgd_momentum_kernel, ctas, kThreads, 0, st, t, flat_grad, flat_momentum, lr_dev, momentum, weight_decay);
  });
}

extern "C" int pcuda_sgd_momentum_sum_step(const pcuda_param_slot_t* slots, int n, float scale, float* flat_grad, float* flat_momentum,
                                           const float* lr_dev, float momentum, float weight_decay, pcuda_stream_t stream) {
  PCUDA_REQUIRE(n >= 0, PCUDA_E_SHAPE, "sgd_momentum_sum_step: n=%d", n);
  if (n == 0) return 0;
  PCUDA_REQUIRE(slots && flat_grad && flat_momentum && lr_dev, PCUDA_E_NULL, "sgd_momentum_sum_step: NULL argument");
  for (int i = 0; i < n; ++i)
    PCUDA_REQUIRE(slots[i].numel == 0 || (slots[i].grad_a && slots[i].param), PCUDA_E_NULL, "sgd_momentum_sum_step: slot %d has a NULL tensor", i);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return for_each_table(slots, n, "sgd_momentum_sum_step", true, [&](const Table& t, int ctas) {
    PCUDA_LAUNCH(sgd_momentum_sum_kernel, ctas, kThreads, 0, st, t, scale, flat_grad, flat_momentum, lr_dev, momentum, weight_decay);
  });
}

extern "C" int pcuda_bce_logits(const float* logit, int n, float target, float weight, float* loss, float* grad_logit,
                                float* accuracy, pcuda_stream_t stream) {
  PCUDA_REQUIRE(n >= 1, PCUDA_E_SHAPE, "bce_logits: n=%d", n);
  PCUDA_REQUIRE(logit && loss, PCUDA_E_NULL, "bce_logits: NULL argument");
  PCUDA_LAUNCH(bce_logits_kernel, 1, 32, 0, static_cast<cudaStream_t>(stream), logit, n, target, weight, loss, grad_logit, accuracy);
  count_launch(1);
  return check_launch("bce_logits");
}
