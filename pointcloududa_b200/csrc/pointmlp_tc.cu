// PointNet shared MLP on the Blackwell tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Two kernels cover every wide (cin >= 64) contraction of the Conv1d(k=1)+BatchNorm1d(+ReLU) stacks of
// networks/PointNetCls.py:41-44,:84-87,:143-162 and of their backward:
//
//   ws_kernel  "weight-stationary":  D[r, m] = sum_k A[r, k] * Bop[m, k]
//        TMEM lane r = an output channel, TMEM column m = a point of the current 256-point tile.
//        A (weights, 128 rows per block) is converted to bf16 once per CTA and stays in shared
//        memory; Bop tiles are produced by 4 warps straight from the fp32 tensors in HBM with the
//        previous layer's BatchNorm + ReLU (or the BN-backward transform) folded into the load, so
//        no normalised / activated / bf16 copy of an activation ever exists in HBM.
//          FWD_DENSE   y = W a            epilogue: store y (pre-BN), per-channel sum / sum^2
//          FWD_POOL    y = W a            epilogue: per-(cloud,channel) max + first arg-max, sums;
//                                         the [B*N, 1024] activation is never written
//          DGRAD       da = dy W          epilogue: ReLU mask of the producing layer, store dz,
//                                         dbeta / dgamma sums of that layer
//          POOL_DGRAD  da = S - u - a Q   (low-rank backward of the pooled layer, see pointmlp.cu)
//        Because a thread owns a channel and walks over points, every per-channel reduction
//        (BN statistics, max-pool, dbeta/dgamma) is thread-local: no shuffles, no shared atomics.
//
//   pt_kernel  "point-contraction":  D[c, k] = sum_m P[m, c] * R[m, k]      (wgrad, Gram matrix)
//        both operands stream (MN-major reads of the same slab format), the accumulator stays in
//        TMEM over the CTA's whole point range and is written once.
//
// Warp roles (ws_kernel, 288 threads): warps 0-3 producers, warps 4-7 epilogue (TMEM lane
// quarter = warp % 4), warp 8 = TMEM allocation + the single MMA-issuing thread.
// Pipelines: smem ring full/empty mbarriers (producers <-> MMA), TMEM accumulator full/empty
// mbarriers (MMA <-> epilogue); two 256-column accumulators so the epilogue of one overlaps the
// MMAs of the other.
#include <algorithm>

#include "pointmlp_common.cuh"
#include "pointmlp_tc.cuh"
#include "tc_common.cuh"

namespace pcuda {
namespace tc {
namespace {

constexpr int kNT = 256;                 // points per tile (UMMA N)
constexpr int kSlabA = 128 * 128;        // bytes of one A slab: 128 rows x 64 bf16
constexpr int kSlabB = kNT * 128;        // bytes of one B slab: 256 rows x 64 bf16
constexpr int kProducers = 128;
constexpr int kWsThreads = 288;
constexpr int kMaxK = 512;
constexpr uint32_t kTmemCols = 512;

enum Mode { FWD_DENSE = 0, FWD_POOL = 1, DGRAD = 2, POOL_DGRAD = 3 };

struct WsParams {
  int B, N, tpc, n_tiles;   // tiles never straddle clouds: tile t -> cloud t / tpc, points (t % tpc) * 256 ...
  int K;                    // contraction length (multiple of 64, <= 512)
  int R;                    // output rows (TMEM lanes): Cout (forward), channels of the previous layer (dgrad)
  int CB, G, nstage;        // row blocks per CTA (1|2), row groups, ring depth
  ActSrc act;               // FWD_*, POOL_DGRAD: Bop = act(prev layer)
  DySrc dy;                 // DGRAD: Bop = dy of this layer
  const float* A;           // FWD_*: W [R, K];  POOL_DGRAD: Q [R, K] (symmetric);  DGRAD: W [K, R] (read transposed)
  // forward epilogue
  const float* bias; const float* gamma; float* y_out; double* stats; unsigned long long* keys;
  // dgrad epilogues
  DgradOut out;
  const float* u;           // POOL_DGRAD: [R]
  const float* coef;        // POOL_DGRAD: [B, Cpool]
  const int* head;          // POOL_DGRAD: [B*N] first channel whose arg-max is this point, or -1
  const int* next;          // POOL_DGRAD: [B, Cpool] next channel selecting the same point, or -1
  const float* Wpool;       // POOL_DGRAD: [Cpool, K]
  int Cpool;
};

// per-contraction-channel constants of the producer transform, in shared memory
//   ActSrc : a  = relu?(y * p0 + p1)                 p0 = invstd*gamma, p1 = beta - mean*p0
//   DySrc  : dy = dz * p0 + (y * p1 + p2)            p0 = gamma*invstd, p1 = -kappa, p2 = kappa*mean - alpha
struct ProducerConsts {
  float p0[kMaxK], p1[kMaxK], p2[kMaxK];
};

template <int MODE>
__device__ __forceinline__ void init_consts(const WsParams& p, ProducerConsts& pc) {
  for (int k = threadIdx.x; k < p.K; k += blockDim.x) {
    if (MODE == DGRAD) {
      const float sc = p.dy.gamma[k] * p.dy.invstd[k];
      pc.p0[k] = sc;
      pc.p1[k] = -p.dy.kappa[k];
      pc.p2[k] = p.dy.kappa[k] * p.dy.mean[k] - p.dy.alpha[k];
    } else if (p.act.y != nullptr) {
      const float sc = p.act.invstd[k] * p.act.gamma[k];
      pc.p0[k] = sc;
      pc.p1[k] = p.act.beta[k] - p.act.mean[k] * sc;
      pc.p2[k] = 0.f;
    } else {
      pc.p0[k] = 1.f; pc.p1[k] = 0.f; pc.p2[k] = 0.f;
    }
  }
}

__device__ __forceinline__ void ld8(const float* p, float (&v)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

// One 256-row x 64-channel slab of the streamed operand, produced by the 128 producer threads.
// Thread -> fixed 16-byte chunk column (8 channels) and 16 rows; rows past the end of the cloud are zero.
template <int MODE>
__device__ __forceinline__ void produce_slab(const WsParams& p, const ProducerConsts& pc, uint8_t* stage, int b,
                                             int n0, int slab, int ptid) {
  const int c = ptid & 7, rg = ptid >> 3;
  const int k0 = slab * 64 + c * 8;
  const int N = p.N;
  float q0[8], q1[8], q2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { q0[i] = pc.p0[k0 + i]; q1[i] = pc.p1[k0 + i]; q2[i] = pc.p2[k0 + i]; }
  const int64_t mbase = static_cast<int64_t>(b) * N + n0;
  constexpr int RB = 4;  // rows in flight per thread
#pragma unroll 1
  for (int i0 = 0; i0 < 16; i0 += RB) {
    float v[RB][8], w[RB][8];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int r = rg + 16 * (i0 + j);
      const bool ok = n0 + r < N;
      if (MODE == DGRAD) {
        if (ok) {
          ld8(p.dy.dz + (mbase + r) * p.K + k0, v[j]);
          ld8(p.dy.y + (mbase + r) * p.K + k0, w[j]);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) { v[j][e] = 0.f; w[j][e] = 0.f; }
        }
      } else if (p.act.y != nullptr) {
        if (ok) ld8(p.act.y + (mbase + r) * p.K + k0, v[j]);
        else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[j][e] = 0.f;
        }
      } else {
        // raw network input x[b, k, n] with explicit strides (the reference passes transposed views)
        const float* px = p.act.x + static_cast<int64_t>(b) * p.act.sxb + static_cast<int64_t>(n0 + r) * p.act.sxn;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[j][e] = ok ? __ldg(px + static_cast<int64_t>(k0 + e) * p.act.sxc) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) {
      const int r = rg + 16 * (i0 + j);
      const bool ok = n0 + r < N;
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float t;
        if (MODE == DGRAD) t = fmaf(v[j][e], q0[e], fmaf(w[j][e], q1[e], q2[e]));
        else {
          t = fmaf(v[j][e], q0[e], q1[e]);
          if (p.act.relu) t = fmaxf(t, 0.f);
        }
        a[e] = ok ? t : 0.f;
      }
      *reinterpret_cast<uint4*>(stage + slab_off(r, c)) = pack8(a);
    }
  }
}

// The stationary operand: 128 rows (TMEM lanes) x K, bf16, written once per CTA.
//   K-major  (FWD_*, POOL_DGRAD): slab s = rows x channels [64 s, 64 s + 64);  A[r, k] = src[(r0 + r) * K + k]
//   MN-major (DGRAD)            : group g (64 lanes) = K contraction rows x 64 lanes; A[r, k] = src[k * R + r0 + r]
template <int MODE>
__device__ __forceinline__ void load_A(const WsParams& p, uint8_t* abase, int r0) {
  const int K = p.K, R = p.R;
  if (MODE != DGRAD) {
    const int chunks = 128 * (K / 8);
    for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
      const int r = i / (K / 8), kc = i - r * (K / 8);
      const int k = kc * 8;
      float a[8];
      if (r0 + r < R) ld8(p.A + static_cast<int64_t>(r0 + r) * K + k, a);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = 0.f;
      }
      *reinterpret_cast<uint4*>(abase + (k >> 6) * kSlabA + slab_off(r, kc & 7)) = pack8(a);
    }
  } else {
    // A^T view of W [K, R]: contraction row k, lanes r0 .. r0+127 -> two 64-lane groups of K rows each
    const int chunks = K * 16;
    for (int i = threadIdx.x; i < chunks; i += blockDim.x) {
      const int k = i >> 4, lc = i & 15;  // lc: 8-lane chunk within the 128 lanes
      const int r = lc * 8;
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = (r0 + r + e < R) ? __ldg(p.A + static_cast<int64_t>(k) * R + r0 + r + e) : 0.f;
      *reinterpret_cast<uint4*>(abase + (lc >> 3) * (K * 128) + slab_off(k, lc & 7)) = pack8(a);
    }
  }
}

struct __align__(8) WsBarriers {
  uint64_t full[8], empty[8], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  uint32_t pad;
};

template <int MODE>
__global__ void __launch_bounds__(kWsThreads, 1) ws_kernel(const WsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int KS = p.K >> 6;
  uint8_t* a_smem = smem;                                   // CB * KS slabs of 16 KB
  uint8_t* b_smem = a_smem + p.CB * KS * kSlabA;            // nstage slabs of 32 KB
  ProducerConsts* pc = reinterpret_cast<ProducerConsts*>(b_smem + p.nstage * kSlabB);
  WsBarriers* bars = reinterpret_cast<WsBarriers*>(pc + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x % p.G;                           // row group of this CTA
  const int P = gridDim.x / p.G, j = blockIdx.x / p.G;      // CTAs per group, index inside the group
  const int t_begin = static_cast<int>(static_cast<int64_t>(p.n_tiles) * j / P);
  const int t_end = static_cast<int>(static_cast<int64_t>(p.n_tiles) * (j + 1) / P);
  const int row_base = g * p.CB * 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.nstage; ++s) { mbar_init(smem_u32(&bars->full[s]), kProducers); mbar_init(smem_u32(&bars->empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bars->acc_full[s]), 1); mbar_init(smem_u32(&bars->acc_empty[s]), 128); }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
  init_consts<MODE>(p, *pc);
  for (int cb = 0; cb < p.CB; ++cb) load_A<MODE>(p, a_smem + cb * KS * kSlabA, row_base + cb * 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    // ================================ producers ===================================================
    uint32_t it = 0;
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / p.tpc, n0 = (t - b * p.tpc) * kNT;
      for (int s = 0; s < KS; ++s, ++it) {
        const uint32_t stage = it % p.nstage, ph = (it / p.nstage) & 1u;
        mbar_wait(smem_u32(&bars->empty[stage]), ph ^ 1u);
        produce_slab<MODE>(p, *pc, b_smem + stage * kSlabB, b, n0, s, threadIdx.x);
        fence_proxy_async();
        mbar_arrive(smem_u32(&bars->full[stage]));
      }
    }
  } else if (warp == 8) {
    // ================================ MMA issuer ==================================================
    const uint32_t idesc = make_idesc(128, kNT, MODE == DGRAD ? 1 : 0, 0);
    uint32_t it = 0, ac = 0;
    for (int t = t_begin; t < t_end; ++t, it += KS) {
      for (int cb = 0; cb < p.CB; ++cb, ++ac) {
        const uint32_t slot = ac & 1u, aph = (ac >> 1) & 1u;
        mbar_wait(smem_u32(&bars->acc_empty[slot]), aph ^ 1u);
        tc_fence_after();
        const uint32_t a_cb = smem_u32(a_smem + cb * KS * kSlabA);
        for (int s = 0; s < KS; ++s) {
          const uint32_t seq = it + s, stage = seq % p.nstage, ph = (seq / p.nstage) & 1u;
          mbar_wait(smem_u32(&bars->full[stage]), ph);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t b_addr = smem_u32(b_smem + stage * kSlabB);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              uint64_t adesc, bdesc = make_sdesc(b_addr + kk * 32, 16, 1024);
              if (MODE == DGRAD) adesc = make_sdesc(a_cb + (s * 64 + kk * 16) * 128, p.K * 128, 1024);
              else adesc = make_sdesc(a_cb + s * kSlabA + kk * 32, 16, 1024);
              umma_bf16(tmem + slot * kNT, adesc, bdesc, idesc, (s | kk) != 0);
            }
            if (cb == p.CB - 1) umma_commit(smem_u32(&bars->empty[stage]));   // slab consumed by every row block
            if (s == KS - 1) umma_commit(smem_u32(&bars->acc_full[slot]));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================ epilogue ====================================================
    const int q = warp & 3;                                  // TMEM lane quarter this warp may access
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    uint32_t ac = 0;
    double S[2] = {0.0, 0.0}, Q[2] = {0.0, 0.0};
    float best[2] = {-INFINITY, -INFINITY};
    int besti[2] = {0, 0};
    int cur_b = -1;
    auto flush_pool = [&](int b) {
      if (MODE != FWD_POOL || b < 0) return;
      for (int cb = 0; cb < p.CB; ++cb) {
        const int r = row_base + cb * 128 + q * 32 + lane;
        if (r < p.R) {
          // the running maximum is over the bias-free accumulator times sign(gamma); the key holds
          // sign(gamma) * (y + bias), what pool_finalize_kernel expects
          const float gm = p.gamma[r];
          const float sg = gm > 0.f ? 1.f : (gm < 0.f ? -1.f : 0.f);
          const float bv = p.bias ? p.bias[r] : 0.f;
          atomicMax(&p.keys[static_cast<int64_t>(b) * p.R + r], pool_key(best[cb] + sg * bv, besti[cb]));
        }
        best[cb] = -INFINITY; besti[cb] = 0;
      }
    };
    for (int t = t_begin; t < t_end; ++t) {
      const int b = t / p.tpc, n0 = (t - b * p.tpc) * kNT;
      const int nvalid = min(kNT, p.N - n0);
      const int64_t m0 = static_cast<int64_t>(b) * p.N + n0;
      if (b != cur_b) { flush_pool(cur_b); cur_b = b; }
#pragma unroll
      for (int cb = 0; cb < 2; ++cb) {
        if (cb >= p.CB) break;
        const uint32_t slot = ac & 1u, aph = (ac >> 1) & 1u;
        ++ac;
        const int r = row_base + cb * 128 + q * 32 + lane;   // this thread's output channel
        const bool rok = r < p.R;
        mbar_wait(smem_u32(&bars->acc_full[slot]), aph);
        tc_fence_after();
        float s_t = 0.f, q_t = 0.f;
        // per-channel constants
        float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
        if (rok) {
          if (MODE == FWD_DENSE) e0 = p.bias ? p.bias[r] : 0.f;
          if (MODE == FWD_POOL) { const float gm = p.gamma[r]; e0 = gm > 0.f ? 1.f : (gm < 0.f ? -1.f : 0.f); }
          if (MODE == DGRAD || MODE == POOL_DGRAD) {
            if (p.out.grad_x == nullptr) { e0 = p.out.mean[r]; e1 = p.out.invstd[r]; e2 = p.out.gamma[r]; e3 = p.out.beta[r]; }
          }
        }
        const float uk = (MODE == POOL_DGRAD && rok) ? p.u[r] : 0.f;
        for (int ch = 0; ch * 32 < nvalid; ++ch) {
          float v[32];
          tmem_ld32(tmem + lane_addr + slot * kNT + ch * 32, v);
          const int ncol = min(32, nvalid - ch * 32);
          if (MODE == FWD_DENSE) {
            if (rok) {
              float* yp = p.y_out + (m0 + ch * 32) * p.R + r;
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (i < ncol) {
                  yp[static_cast<int64_t>(i) * p.R] = v[i] + e0;
                  s_t += v[i];
                  q_t = fmaf(v[i], v[i], q_t);
                }
              }
            }
          } else if (MODE == FWD_POOL) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < ncol) {
                s_t += v[i];
                q_t = fmaf(v[i], v[i], q_t);
                const float vv = v[i] * e0;     // sign(gamma) * y: BN is monotone per channel
                if (vv > best[cb]) { best[cb] = vv; besti[cb] = n0 + ch * 32 + i; }
              }
            }
          } else {
            // dgrad epilogues
            int hv = -1;
            if (MODE == POOL_DGRAD) hv = (lane < ncol) ? __ldg(p.head + m0 + ch * 32 + lane) : -1;
            const bool to_x = p.out.grad_x != nullptr;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (i < ncol) {
                float val = v[i];
                if (MODE == POOL_DGRAD) {
                  float sp = 0.f;
                  int c = __shfl_sync(0xffffffffu, hv, i);
                  while (c >= 0) {       // warp-uniform: every lane walks the same list
                    if (rok) sp = fmaf(__ldg(p.coef + static_cast<int64_t>(b) * p.Cpool + c), __ldg(p.Wpool + static_cast<int64_t>(c) * p.K + r), sp);
                    c = __ldg(p.next + static_cast<int64_t>(b) * p.Cpool + c);
                  }
                  val = sp - uk - val;
                }
                if (rok) {
                  const int64_t m = m0 + ch * 32 + i;
                  if (to_x) {
                    p.out.grad_x[(static_cast<int64_t>(b) * p.R + r) * p.N + n0 + ch * 32 + i] = val;
                  } else {
                    const float yh = (__ldg(p.out.y_prev + m * p.R + r) - e0) * e1;
                    if (p.out.relu && !(fmaf(yh, e2, e3) > 0.f)) val = 0.f;
                    p.out.dz_prev[m * p.R + r] = val;
                    s_t += val;
                    q_t = fmaf(val, yh, q_t);
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&bars->acc_empty[slot]));
        S[cb] += static_cast<double>(s_t);
        Q[cb] += static_cast<double>(q_t);
      }
    }
    flush_pool(cur_b);
    for (int cb = 0; cb < p.CB; ++cb) {
      const int r = row_base + cb * 128 + q * 32 + lane;
      if (r >= p.R) continue;
      if (MODE == FWD_DENSE || MODE == FWD_POOL) {
        if (p.stats) { atomicAdd(&p.stats[r], S[cb]); atomicAdd(&p.stats[p.R + r], Q[cb]); }
      } else if (p.out.grad_x == nullptr) {
        atomicAdd(&p.out.sums[r], S[cb]);
        atomicAdd(&p.out.sums[p.R + r], Q[cb]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, kTmemCols);
}

// ---- point-contraction kernel (wgrad / Gram) ----------------------------------------------------
//   D[c, k] = sum_{m in range} Pm[m, c] * Rm[m, k]     c: 128-lane block, k: Kr columns (<= 256)
// GRAM: Pm = Rm = act (CPb = Kr = K <= 128 per block);  WGRAD: Pm = dy (this layer), Rm = act (previous layer)
constexpr int kPT = 128;          // points per stage
constexpr int kPtThreads = 288;   // warps 0-7 producers, warp 8 MMA (+ all of 0-3 for the final epilogue)

struct PtParams {
  int64_t M;
  int S;                    // point-range splits
  int C, Kr;                // rows of D (channels of P), columns of D (channels of R)
  int gram;                 // 1: P == R == act
  DySrc dy;                 // WGRAD
  ActSrc act;               // R operand (and P for GRAM)
  float* partial;           // [S, C, Kr]
  double* colsum;           // GRAM: [S, Kr] column sums of the bf16-rounded activation
};

struct PtConsts {
  float d0[kMaxK], d1[kMaxK], d2[kMaxK];   // dy transform (per P channel)
  float a0[kMaxK], a1[kMaxK];              // act transform (per R channel)
};

struct __align__(8) PtBarriers {
  uint64_t full[4], empty[4], done;
  uint32_t tmem_base, pad;
};

__global__ void __launch_bounds__(kPtThreads, 1) pt_kernel(const PtParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int GR = p.Kr >> 6;                       // 64-channel groups of R
  const int GP = p.gram ? 0 : 2;                  // 64-channel groups of P per 128-lane block (GRAM reuses R)
  const int stage_bytes = (GP + GR) * kPT * 128;
  const int nstage = 2;
  PtConsts* pc = reinterpret_cast<PtConsts*>(smem + nstage * stage_bytes);
  PtBarriers* bars = reinterpret_cast<PtBarriers*>(pc + 1);
  float* csum = reinterpret_cast<float*>(bars + 1);   // [16][Kr] per-row-group column sums (GRAM)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = blockIdx.y;                    // 128-row block of D
  const int s = blockIdx.x;
  const int64_t chunk = ((p.M + p.S - 1) / p.S + kPT - 1) / kPT * kPT;
  const int64_t m_begin = s * chunk, m_end = min(p.M, m_begin + chunk);
  const int n_steps = m_end > m_begin ? static_cast<int>((m_end - m_begin + kPT - 1) / kPT) : 0;
  const uint32_t ncols = p.Kr <= 32 ? 32u : (p.Kr <= 64 ? 64u : (p.Kr <= 128 ? 128u : 256u));

  if (threadIdx.x == 0) {
    for (int i = 0; i < nstage; ++i) { mbar_init(smem_u32(&bars->full[i]), 256); mbar_init(smem_u32(&bars->empty[i]), 1); }
    mbar_init(smem_u32(&bars->done), 1);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(smem_u32(&bars->tmem_base), ncols);
  for (int k = threadIdx.x; k < p.Kr; k += blockDim.x) {
    if (p.act.y != nullptr) {
      const float sc = p.act.invstd[k] * p.act.gamma[k];
      pc->a0[k] = sc; pc->a1[k] = p.act.beta[k] - p.act.mean[k] * sc;
    } else { pc->a0[k] = 1.f; pc->a1[k] = 0.f; }
  }
  if (!p.gram)
    for (int k = threadIdx.x; k < 128; k += blockDim.x) {
      const int c = cblk * 128 + k;
      if (c < p.C) {
        pc->d0[k] = p.dy.gamma[c] * p.dy.invstd[c];
        pc->d1[k] = -p.dy.kappa[c];
        pc->d2[k] = p.dy.kappa[c] * p.dy.mean[c] - p.dy.alpha[c];
      } else { pc->d0[k] = 0.f; pc->d1[k] = 0.f; pc->d2[k] = 0.f; }
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 8) {
    // producers: 256 threads; thread -> chunk column cc (8 channels), rows rg + 32*i
    const int ptid = threadIdx.x;
    const int cc = ptid & 7, rg = ptid >> 3;       // rg in [0, 32)
    float colacc[4][8];                            // GRAM: up to 4 R groups... (Kr <= 256)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int e = 0; e < 8; ++e) colacc[a][e] = 0.f;
    for (int it = 0; it < n_steps; ++it) {
      const uint32_t stage = it & 1u, ph = (it >> 1) & 1u;
      mbar_wait(smem_u32(&bars->empty[stage]), ph ^ 1u);
      uint8_t* sb = smem + stage * stage_bytes;
      const int64_t mt = m_begin + static_cast<int64_t>(it) * kPT;
      // P groups (dy of this layer's channels cblk*128 ..)
      for (int gidx = 0; gidx < GP; ++gidx) {
        const int kl = gidx * 64 + cc * 8;          // local channel in the 128-block
        const int c0 = cblk * 128 + kl;
        float q0[8], q1[8], q2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { q0[e] = pc->d0[kl + e]; q1[e] = pc->d1[kl + e]; q2[e] = pc->d2[kl + e]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rg + 32 * i;
          const int64_t m = mt + r;
          float a[8];
          if (m < m_end && c0 < p.C) {
            float v[8], w[8];
            ld8(p.dy.dz + m * p.C + c0, v);
            ld8(p.dy.y + m * p.C + c0, w);
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = fmaf(v[e], q0[e], fmaf(w[e], q1[e], q2[e]));
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = 0.f;
          }
          *reinterpret_cast<uint4*>(sb + gidx * (kPT * 128) + slab_off(r, cc)) = pack8(a);
        }
      }
      // R groups (activation of the previous layer)
#pragma unroll
      for (int gidx = 0; gidx < 4; ++gidx) {
        if (gidx >= GR) break;
        const int k0 = gidx * 64 + cc * 8;
        float q0[8], q1[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { q0[e] = pc->a0[k0 + e]; q1[e] = pc->a1[k0 + e]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rg + 32 * i;
          const int64_t m = mt + r;
          float a[8];
          if (m < m_end) {
            if (p.act.y != nullptr) {
              float v[8];
              ld8(p.act.y + m * p.Kr + k0, v);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float t = fmaf(v[e], q0[e], q1[e]);
                a[e] = p.act.relu ? fmaxf(t, 0.f) : t;
              }
            } else {
              const int64_t b = m / p.act.N, n = m - b * p.act.N;
              const float* px = p.act.x + b * p.act.sxb + n * p.act.sxn;
#pragma unroll
              for (int e = 0; e < 8; ++e) a[e] = __ldg(px + static_cast<int64_t>(k0 + e) * p.act.sxc);
            }
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) a[e] = 0.f;
          }
          const uint4 pk = pack8(a);
          *reinterpret_cast<uint4*>(sb + (GP + gidx) * (kPT * 128) + slab_off(r, cc)) = pk;
          if (p.gram) {
            // column sums of exactly what the tensor core will see
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(h[e]);
              colacc[gidx][2 * e] += f.x;
              colacc[gidx][2 * e + 1] += f.y;
            }
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&bars->full[stage]));
    }
    if (p.gram) {
      // reduce the 32 row groups in a fixed order: csum[rg][k] -> thread k sums over rg
#pragma unroll
      for (int gidx = 0; gidx < 4; ++gidx) {
        if (gidx >= GR) break;
#pragma unroll
        for (int e = 0; e < 8; ++e) csum[rg * p.Kr + gidx * 64 + cc * 8 + e] = colacc[gidx][e];
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int k = ptid; k < p.Kr; k += 256) {
        double acc = 0.0;
        for (int r2 = 0; r2 < 32; ++r2) acc += static_cast<double>(csum[r2 * p.Kr + k]);
        p.colsum[static_cast<int64_t>(s) * p.Kr + k] = acc;
      }
    }
  } else {
    // MMA issuer
    const uint32_t idesc = make_idesc(128, p.Kr, 1, 1);
    for (int it = 0; it < n_steps; ++it) {
      const uint32_t stage = it & 1u, ph = (it >> 1) & 1u;
      mbar_wait(smem_u32(&bars->full[stage]), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sb = smem_u32(smem + stage * stage_bytes);
        const uint32_t pa = p.gram ? sb + (cblk * 2) * (kPT * 128) : sb;   // GRAM: lanes = R groups 2*cblk, 2*cblk+1
        const uint32_t ra = sb + GP * (kPT * 128);
#pragma unroll
        for (int kk = 0; kk < kPT / 16; ++kk) {
          const uint64_t adesc = make_sdesc(pa + kk * 2048, kPT * 128, 1024);
          const uint64_t bdesc = make_sdesc(ra + kk * 2048, kPT * 128, 1024);
          umma_bf16(tmem, adesc, bdesc, idesc, (it | kk) != 0);
        }
        umma_commit(smem_u32(&bars->empty[stage]));
        if (it == n_steps - 1) umma_commit(smem_u32(&bars->done));
      }
      __syncwarp();
    }
  }
  // final epilogue: warps 0-3 read the accumulator (lane quarter = warp) and write the partial
  if (warp < 4) {
    float* outp = p.partial + (static_cast<int64_t>(s) * p.C) * p.Kr;
    const int c = cblk * 128 + warp * 32 + lane;
    if (n_steps > 0) {
      mbar_wait(smem_u32(&bars->done), 0);
      tc_fence_after();
    }
    for (int ch = 0; ch * 32 < p.Kr; ++ch) {
      float v[32];
      if (n_steps > 0) tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + ch * 32, v);
      else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      }
      if (c < p.C) {
#pragma unroll
        for (int i = 0; i < 32; i += 4)
          *reinterpret_cast<float4*>(outp + static_cast<int64_t>(c) * p.Kr + ch * 32 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem, ncols);
}

// ---- host side ----------------------------------------------------------------------------------
constexpr int kSmemBudget = 227 * 1024;

struct WsPlan {
  int CB, G, nstage, grid;
  size_t smem;
  bool ok;
};

WsPlan plan_ws(int R, int K, int n_tiles) {
  WsPlan pl{};
  const int KS = K / 64;
  const int rblocks = (R + 127) / 128;
  const size_t fixed = 1024 + sizeof(ProducerConsts) + sizeof(WsBarriers) + 64;
  pl.CB = (rblocks >= 2 && (rblocks % 2) == 0 && K <= 128) ? 2 : 1;
  for (;;) {
    const size_t a = static_cast<size_t>(pl.CB) * KS * kSlabA;
    const int ns = static_cast<int>((kSmemBudget - fixed - a) / kSlabB);
    pl.nstage = std::min(8, ns);
    const int need = pl.CB == 2 ? 2 * KS : 2;   // CB == 2 keeps a whole tile resident while the next is produced
    if (pl.nstage >= need) break;
    if (pl.CB == 2) { pl.CB = 1; continue; }
    pl.ok = false;
    return pl;
  }
  pl.G = rblocks / pl.CB;
  const int P = std::max(1, std::min(n_tiles, sm_count() / pl.G));
  pl.grid = P * pl.G;
  pl.smem = fixed + static_cast<size_t>(pl.CB) * KS * kSlabA + static_cast<size_t>(pl.nstage) * kSlabB;
  pl.ok = true;
  return pl;
}

template <int MODE>
int launch_ws(WsParams& p, cudaStream_t st, const char* what) {
  const WsPlan pl = plan_ws(p.R, p.K, p.n_tiles);
  if (!pl.ok) return fail(PCUDA_E_UNSUPPORTED, "%s: K=%d does not fit the tensor-core kernel", what, p.K);
  p.CB = pl.CB; p.G = pl.G; p.nstage = pl.nstage;
  static bool attr_done[4] = {false, false, false, false};
  if (!attr_done[MODE]) {
    cudaError_t e = cudaFuncSetAttribute(ws_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    attr_done[MODE] = true;
  }
  ws_kernel<MODE><<<pl.grid, kWsThreads, pl.smem, st>>>(p);
  count_launch();
  return check_launch(what);
}

WsParams base_params(int B, int N, int K, int R) {
  WsParams p{};
  p.B = B; p.N = N; p.tpc = (N + kNT - 1) / kNT; p.n_tiles = B * p.tpc; p.K = K; p.R = R;
  return p;
}

}  // namespace

bool supports(int cin) { return cin >= 64 && cin <= kMaxK && (cin % 64) == 0; }

int fwd_layer(const ActSrc& src, const pcuda_mlp_layer_t& L, bool pool, int B, int N, double* stats,
              unsigned long long* keys, cudaStream_t st) {
  WsParams p = base_params(B, N, L.cin, L.cout);
  p.act = src; p.A = L.weight; p.bias = L.bias; p.gamma = L.gamma; p.y_out = L.y; p.stats = stats; p.keys = keys;
  return pool ? launch_ws<FWD_POOL>(p, st, "tc::fwd_layer(pool)") : launch_ws<FWD_DENSE>(p, st, "tc::fwd_layer");
}

int dgrad_layer(const DySrc& dys, const float* W, int B, int N, const DgradOut& out, cudaStream_t st) {
  WsParams p = base_params(B, N, dys.C, out.Cp);
  p.dy = dys; p.A = W; p.out = out;
  return launch_ws<DGRAD>(p, st, "tc::dgrad_layer");
}

int pool_dgrad(const ActSrc& src, const float* Q, const float* u, const float* Wpool, const float* coef,
               const int* head, const int* next, int Cpool, int B, int N, const DgradOut& out, cudaStream_t st) {
  WsParams p = base_params(B, N, src.C, src.C);
  p.act = src; p.A = Q; p.u = u; p.Wpool = Wpool; p.coef = coef; p.head = head; p.next = next; p.Cpool = Cpool;
  p.out = out;
  return launch_ws<POOL_DGRAD>(p, st, "tc::pool_dgrad");
}

int pt_splits(int64_t M, int rblocks) {
  const int64_t steps = (M + kPT - 1) / kPT;
  int64_t S = std::max<int64_t>(1, sm_count() / std::max(1, rblocks));
  if (S > steps) S = steps;
  return static_cast<int>(S);
}

bool pt_supports(int C, int Kr, bool gram) {
  if (Kr % 64 != 0 || Kr < 64 || Kr > 256) return false;
  if (gram) return C == Kr && Kr <= 256;
  return C >= 64 && (C % 8) == 0;
}

static int launch_pt(PtParams& p, cudaStream_t st, const char* what) {
  const int GR = p.Kr / 64, GP = p.gram ? 0 : 2;
  const size_t smem = 1024 + 2 * static_cast<size_t>(GP + GR) * kPT * 128 + sizeof(PtConsts) + sizeof(PtBarriers) +
                      (p.gram ? sizeof(float) * 32 * p.Kr : 0) + 64;
  if (smem > static_cast<size_t>(kSmemBudget)) return fail(PCUDA_E_UNSUPPORTED, "%s: Kr=%d does not fit", what, p.Kr);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(pt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    attr_done = true;
  }
  const dim3 grid(p.S, (p.C + 127) / 128);
  pt_kernel<<<grid, kPtThreads, smem, st>>>(p);
  count_launch();
  return check_launch(what);
}

int wgrad_layer(const DySrc& dys, const ActSrc& prev, int64_t M, int S, float* partial, cudaStream_t st) {
  PtParams p{};
  p.M = M; p.S = S; p.C = dys.C; p.Kr = prev.C; p.gram = 0; p.dy = dys; p.act = prev; p.partial = partial;
  return launch_pt(p, st, "tc::wgrad_layer");
}

int gram(const ActSrc& act, int64_t M, int S, float* partial, double* colsum, cudaStream_t st) {
  PtParams p{};
  p.M = M; p.S = S; p.C = act.C; p.Kr = act.C; p.gram = 1; p.act = act; p.partial = partial; p.colsum = colsum;
  return launch_pt(p, st, "tc::gram");
}

}  // namespace tc
}  // namespace pcuda
