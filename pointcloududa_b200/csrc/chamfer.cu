// Chamfer / nearest-neighbour loss: tiled all-pairs search (forward) and index scatter (backward).
//
// Replaces utils/loss.py:40-76 (batch_NN_loss + batch_pairwise_dist) of the reference, which
// materialises >= 8 [B,N,N] fp32 temporaries through six K=3 bmm calls.  Here nothing of size
// N x N ever exists: each thread owns R query points in registers, candidate points stream
// through shared memory as packed (x, y, z, |p|^2) float4 broadcast reads, and only
// (distance, index) per point reaches HBM.
//
// Bit-exactness contract: the arg-min the reference returns is
//     argmin_j sqrtf( fl( fl(rx_i + ry_j) - 2*zz_ij ) + 1e-5f ),   first index on ties,
// with every dot product the K=3 chain fmaf(a2,b2, fmaf(a1,b1, a0*b0)).  The kernel evaluates
// exactly that expression with explicit round-to-nearest intrinsics (no contraction, no
// reassociation).  sqrt is applied lazily: f(P) = sqrtf(P + 1e-5f) is monotone non-decreasing
// in P, so the running minimum is tracked on P and f is evaluated only when a strictly smaller P
// arrives — then the index moves only if f also strictly decreased (distinct P can round to the
// same distance, and the reference then keeps the EARLIER index).
//
// Roofline: FP32 issue rate, not HBM — algorithmic HBM bytes are 12*B*(N+M) in and
// 12*B*(N+M) out, the work is 2*B*N*M ordered pair evaluations at 5 FP32 ops + compare each.
#include <algorithm>

#include "pcuda_common.cuh"

namespace pcuda {
namespace {

constexpr int kTileJ = 1024;           // candidate points per shared-memory tile (16 KB)
constexpr float kEps = 0.00001f;       // loss.py:68,71
constexpr double kFixScale = 1099511627776.0;  // 2^40, backward fixed-point accumulators

struct FwdWs {
  unsigned int ticket;
  unsigned int pad;
  double partial[1];  // [2][B][tiles]
};

__device__ __forceinline__ float norm3(float a0, float a1, float a2) {
  return __fmaf_rn(a2, a2, __fmaf_rn(a1, a1, __fmul_rn(a0, a0)));
}

// Deterministic loss reduction shared by the three forward kernels: every CTA (active or not) leaves the fp64 sum
// of its rows' distances in ws->partial; the CTA that takes the last ticket adds all partials in a fixed order
//     loss = sum_b (sum_i d1)/N /B + sum_b (sum_j d2)/N /B      (loss.py:73-75)
// and re-arms the ticket, so the workspace is reusable without a host-side reset.
template <int THREADS>
__device__ __forceinline__ void finish_loss(double my_sum, FwdWs* __restrict__ ws, float* __restrict__ loss, int dir,
                                            int b, int B, int N) {
  __shared__ double warp_part[THREADS / 32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double w = warp_sum(my_sum);
  if (lane == 0) warp_part[wid] = w;
  __syncthreads();
  const int tiles_max = gridDim.x;
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) s += warp_part[i];
    ws->partial[(static_cast<int64_t>(dir) * B + b) * tiles_max + blockIdx.x] = s;
    __threadfence();
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    is_last = atomicAdd(&ws->ticket, 1u) == total - 1;
  }
  __syncthreads();
  if (is_last && wid == 0) {
    __threadfence();
    double acc = 0.0;
    const int64_t n_part = static_cast<int64_t>(2) * B * tiles_max;
    const volatile double* part = ws->partial;
    for (int64_t k = lane; k < n_part; k += 32) acc += part[k];
    acc = warp_sum(acc);
    if (lane == 0) {
      *loss = static_cast<float>(acc / (static_cast<double>(N) * static_cast<double>(B)));
      ws->ticket = 0u;
      __threadfence();
    }
  }
}

// One launch covers both directions: blockIdx.z == 0 searches y for every x_i (d1,i1),
// blockIdx.z == 1 searches x for every y_j (d2,i2).
template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
chamfer_nn_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M,
                  float* __restrict__ d1, int64_t* __restrict__ i1, float* __restrict__ d2,
                  int64_t* __restrict__ i2, FwdWs* __restrict__ ws, float* __restrict__ loss,
                  int tiles_x, int tiles_y, int B) {
  pdl_entry();
  __shared__ float4 tile[kTileJ];
  const int dir = blockIdx.z;
  const int b = blockIdx.y;
  const int nq = dir == 0 ? N : M;  // queries (rows)
  const int nc = dir == 0 ? M : N;  // candidates (columns)
  const int tiles_q = dir == 0 ? tiles_x : tiles_y;
  const bool active_block = static_cast<int>(blockIdx.x) < tiles_q;

  const float* __restrict__ qbase = (dir == 0 ? x : y) + static_cast<int64_t>(b) * nq * 3;
  const float* __restrict__ cbase = (dir == 0 ? y : x) + static_cast<int64_t>(b) * nc * 3;
  float* __restrict__ dout = (dir == 0 ? d1 : d2) + static_cast<int64_t>(b) * nq;
  int64_t* __restrict__ iout = (dir == 0 ? i1 : i2) + static_cast<int64_t>(b) * nq;

  double my_sum = 0.0;
  if (active_block) {
    float q0[R], q1[R], q2[R], rq[R], bestP[R], bestD[R];
    int bestI[R];
    const int row0 = blockIdx.x * (THREADS * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * THREADS;
      const int rr = row < nq ? row : nq - 1;  // clamp: out-of-range rows compute but never store
      q0[r] = __ldg(qbase + rr * 3 + 0);
      q1[r] = __ldg(qbase + rr * 3 + 1);
      q2[r] = __ldg(qbase + rr * 3 + 2);
      rq[r] = norm3(q0[r], q1[r], q2[r]);
      bestP[r] = INFINITY;
      bestD[r] = INFINITY;
      bestI[r] = 0;
    }

    for (int j0 = 0; j0 < nc; j0 += kTileJ) {
      const int tj = min(kTileJ, nc - j0);
      __syncthreads();
      for (int t = threadIdx.x; t < tj; t += THREADS) {
        const float c0 = __ldg(cbase + (j0 + t) * 3 + 0);
        const float c1 = __ldg(cbase + (j0 + t) * 3 + 1);
        const float c2 = __ldg(cbase + (j0 + t) * 3 + 2);
        tile[t] = make_float4(c0, c1, c2, norm3(c0, c1, c2));
      }
      __syncthreads();
#pragma unroll 4
      for (int t = 0; t < tj; ++t) {
        const float4 c = tile[t];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float zz = __fmaf_rn(q2[r], c.z, __fmaf_rn(q1[r], c.y, __fmul_rn(q0[r], c.x)));
          // direction 0: rx_i + ry_j ; direction 1: ry_j + rx_i — fp32 addition commutes.
          const float P = __fmaf_rn(-2.0f, zz, __fadd_rn(rq[r], c.w));
          if (P < bestP[r]) {
            bestP[r] = P;
            const float d = __fsqrt_rn(__fadd_rn(P, kEps));
            if (d < bestD[r]) {
              bestD[r] = d;
              bestI[r] = j0 + t;
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * THREADS;
      if (row < nq) {
        dout[row] = bestD[r];
        iout[row] = bestI[r];
        my_sum += static_cast<double>(bestD[r]);
      }
    }
  }

  if (loss != nullptr) finish_loss<THREADS>(my_sum, ws, loss, dir, b, B, N);
}


// ---- v2: packed FP32 (FFMA2 / FADD2 / FMUL2, new on sm_100) -----------------------------------
// Two candidate columns ride in the two halves of every 64-bit register operand, so the five
// IEEE-rounded FP32 operations of a pair evaluation cost 2.5 issue slots instead of 5; the
// per-component rounding is identical to the scalar instructions, so indices stay bit-exact.
// The (rare) "new minimum" handling is hoisted behind ONE warp-level branch per column pair.
// Shared-memory tile layout: A[jp] = (x_j, x_j+1, y_j, y_j+1), B[jp] = (z_j, z_j+1, r_j, r_j+1).
__device__ __forceinline__ void nn_update(float P, int j, float& bestP, float& bestD, int& bestI) {
  if (P < bestP) {
    bestP = P;
    const float d = __fsqrt_rn(__fadd_rn(P, kEps));
    if (d < bestD) {
      bestD = d;
      bestI = j;
    }
  }
}

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
chamfer_nn2_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M,
                   float* __restrict__ d1, int64_t* __restrict__ i1, float* __restrict__ d2,
                   int64_t* __restrict__ i2, FwdWs* __restrict__ ws, float* __restrict__ loss,
                   int tiles_x, int tiles_y, int B) {
  pdl_entry();
  __shared__ float4 tileA[kTileJ / 2];
  __shared__ float4 tileB[kTileJ / 2];
  const int dir = blockIdx.z;
  const int b = blockIdx.y;
  const int nq = dir == 0 ? N : M;
  const int nc = dir == 0 ? M : N;
  const int tiles_q = dir == 0 ? tiles_x : tiles_y;
  const bool active_block = static_cast<int>(blockIdx.x) < tiles_q;

  const float* __restrict__ qbase = (dir == 0 ? x : y) + static_cast<int64_t>(b) * nq * 3;
  const float* __restrict__ cbase = (dir == 0 ? y : x) + static_cast<int64_t>(b) * nc * 3;
  float* __restrict__ dout = (dir == 0 ? d1 : d2) + static_cast<int64_t>(b) * nq;
  int64_t* __restrict__ iout = (dir == 0 ? i1 : i2) + static_cast<int64_t>(b) * nq;

  double my_sum = 0.0;
  if (active_block) {
    float2 q0[R], q1[R], q2[R], rq[R];
    float bestP[R], bestD[R];
    int bestI[R];
    const int row0 = blockIdx.x * (THREADS * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * THREADS;
      const int rr = row < nq ? row : nq - 1;
      const float a0 = __ldg(qbase + rr * 3 + 0);
      const float a1 = __ldg(qbase + rr * 3 + 1);
      const float a2 = __ldg(qbase + rr * 3 + 2);
      const float n = norm3(a0, a1, a2);
      q0[r] = make_float2(a0, a0);
      q1[r] = make_float2(a1, a1);
      q2[r] = make_float2(a2, a2);
      rq[r] = make_float2(n, n);
      bestP[r] = INFINITY;
      bestD[r] = INFINITY;
      bestI[r] = 0;
    }
    const float2 neg2 = make_float2(-2.0f, -2.0f);

    for (int j0 = 0; j0 < nc; j0 += kTileJ) {
      const int tj = min(kTileJ, nc - j0);
      const int npair = (tj + 1) >> 1;
      __syncthreads();
      for (int t = threadIdx.x; t < npair; t += THREADS) {
        const int ja = j0 + 2 * t;
        const bool has_b = (2 * t + 1) < tj;
        const float xa = __ldg(cbase + ja * 3 + 0), ya = __ldg(cbase + ja * 3 + 1), za = __ldg(cbase + ja * 3 + 2);
        // odd tail: a sentinel column whose P is +inf can never be selected
        const float xb = has_b ? __ldg(cbase + ja * 3 + 3) : 0.0f;
        const float yb = has_b ? __ldg(cbase + ja * 3 + 4) : 0.0f;
        const float zb = has_b ? __ldg(cbase + ja * 3 + 5) : 0.0f;
        const float ra = norm3(xa, ya, za);
        const float rb = has_b ? norm3(xb, yb, zb) : INFINITY;
        tileA[t] = make_float4(xa, xb, ya, yb);
        tileB[t] = make_float4(za, zb, ra, rb);
      }
      __syncthreads();
#pragma unroll 2
      for (int t = 0; t < npair; ++t) {
        const float4 A = tileA[t];
        const float4 Bv = tileB[t];
        const float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w);
        const float2 cz = make_float2(Bv.x, Bv.y), cr = make_float2(Bv.z, Bv.w);
        float2 P[R];
        bool any = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float2 zz = __ffma2_rn(q2[r], cz, __ffma2_rn(q1[r], cy, __fmul2_rn(q0[r], cx)));
          P[r] = __ffma2_rn(neg2, zz, __fadd2_rn(rq[r], cr));
          any = any || (P[r].x < bestP[r]) || (P[r].y < bestP[r]);
        }
        if (any) {
          const int j = j0 + 2 * t;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            nn_update(P[r].x, j, bestP[r], bestD[r], bestI[r]);
            nn_update(P[r].y, j + 1, bestP[r], bestD[r], bestI[r]);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * THREADS;
      if (row < nq) {
        dout[row] = bestD[r];
        iout[row] = bestI[r];
        my_sum += static_cast<double>(bestD[r]);
      }
    }
  }

  if (loss != nullptr) finish_loss<THREADS>(my_sum, ws, loss, dir, b, B, N);
}

// ---- v3: conservative 3-FMA prefilter in front of the exact evaluation ------------------------------
// Only candidates that can still lower a row's running minimum need the reference's exact arithmetic.
// With n = -2*a (exact) the chain
//     Q = fma(n2,c2, fma(n1,c1, fma(n0,c0, r_c)))          ~  r_c - 2 a.c  =  P - r_a
// costs 3 packed FFMA2 per two candidates instead of the 5 packed operations of the exact formula, and
//     |P - (r_a + Q)| <= 13 u (r_a + r_c),  u = 2^-24
// (3 roundings in Q: 3.1u(r_a + 2 r_c); 5 in P: 6.1u(r_a + r_c); Cauchy-Schwarz for sum|a_k c_k|).  A pair
// is skipped iff  Q >= thr_a := (lim_a - r_a) + 2^-19 (r_a + max_tile r_c) + 1e-36  (rounded up), which
// implies P >= lim_a.  lim_a = min(bestP_a, cap_a):
//   * P >= bestP_a: the exact scan would not have touched its state either;
//   * P >  cap_a, where cap_a = U_a + 2^-19 (|U_a| + eps) and U_a >= min_j P_aj is an upper bound of the
//     row's final minimum taken from a strided sample of the candidates (seed pass, filter arithmetic
//     only): then sqrtf(P + eps) > sqrtf(U_a + eps) >= the final minimum distance, so the candidate is
//     neither the arg-min nor tied with it, and dropping it from the in-order scan changes no output.
//     The seed cuts the ~ln N running-minimum updates per row (each a divergent slow-path entry for the
//     whole warp) to ~ln(sample stride).
// Everything else takes the slow path: the exact expression of v2 and the same in-order update rule, so
// distances and indices are bit-identical to the scalar kernel by construction.
constexpr float kFilterMargin = 1.9073486328125e-06f;   // 2^-19  (> (13 + rounding of thr) * 2^-24)
constexpr float kFilterAbs = 1e-36f;                    // covers underflow in the products
constexpr int kSeedStride = 16;                         // seed pass: every 16th column pair
constexpr int kSeedMinCols = 2048;                      // below this the seed pass does not pay ...
constexpr int kSeedMaxCols = 8192;                      // ... nor above (measured: +4 % at 4096, -1 % at 16384 columns)

// 128-bit shared-memory load from a 32-bit shared address: the scan walks the tile with one register and
// immediate offsets (through generic pointers the compiler re-derived the shared window base — S2UR,
// ULEA, LEA: 9 extra instructions per two column pairs — in every iteration).
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// in-order update of (bestP, bestI); the distance itself is recomputed from bestP at the end.
// The index moves iff sqrtf(P + eps) < sqrtf(bestP + eps): decided without the square roots when the
// arguments differ by more than 2^-20 relative (then the rounded roots differ too).
__device__ __forceinline__ void nn_update_lazy(float P, int j, float& bestP, int& bestI) {
  if (P < bestP) {
    const float tn = __fadd_rn(P, kEps), to = __fadd_rn(bestP, kEps);
    bestP = P;
    if (tn <= __fmul_rd(to, 0.99999904632568359375f) || __fsqrt_rn(tn) < __fsqrt_rn(to)) bestI = j;
  }
}

template <int R, int THREADS, int MINB = (R == 4 ? 8 : 1)>
__global__ void __launch_bounds__(THREADS, MINB)
chamfer_nn3_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M,
                   float* __restrict__ d1, int64_t* __restrict__ i1, float* __restrict__ d2,
                   int64_t* __restrict__ i2, FwdWs* __restrict__ ws, float* __restrict__ loss,
                   int tiles_x, int tiles_y, int B, int seed_stride, int seed_forced) {
  pdl_entry();
  __shared__ float4 tileA[kTileJ / 2 + 1];   // + 1: the scan prefetches one entry ahead
  __shared__ float4 tileB[kTileJ / 2 + 1];
  __shared__ float warp_rmax[THREADS / 32];
  const int dir = blockIdx.z;
  const int b = blockIdx.y;
  const int nq = dir == 0 ? N : M;
  const int nc = dir == 0 ? M : N;
  const int tiles_q = dir == 0 ? tiles_x : tiles_y;
  const bool active_block = static_cast<int>(blockIdx.x) < tiles_q;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

  const float* __restrict__ qbase = (dir == 0 ? x : y) + static_cast<int64_t>(b) * nq * 3;
  const float* __restrict__ cbase = (dir == 0 ? y : x) + static_cast<int64_t>(b) * nc * 3;
  float* __restrict__ dout = (dir == 0 ? d1 : d2) + static_cast<int64_t>(b) * nq;
  int64_t* __restrict__ iout = (dir == 0 ? i1 : i2) + static_cast<int64_t>(b) * nq;

  double my_sum = 0.0;
  if (active_block) {
    float2 n0[R], n1[R], n2[R];          // (-2a_k, -2a_k)
    float rq[R], thr[R], cap[R], bestP[R];
    int bestI[R];
    const int row0 = blockIdx.x * (THREADS * R) + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * THREADS;
      const int rr = row < nq ? row : nq - 1;
      const float a0 = __ldg(qbase + rr * 3 + 0);
      const float a1 = __ldg(qbase + rr * 3 + 1);
      const float a2 = __ldg(qbase + rr * 3 + 2);
      rq[r] = norm3(a0, a1, a2);
      n0[r] = make_float2(-2.0f * a0, -2.0f * a0);
      n1[r] = make_float2(-2.0f * a1, -2.0f * a1);
      n2[r] = make_float2(-2.0f * a2, -2.0f * a2);
      bestP[r] = INFINITY;
      bestI[r] = 0;
      cap[r] = INFINITY;
      thr[r] = INFINITY;
    }
    const float2 neg2 = make_float2(-2.0f, -2.0f);
    const uint32_t sA = static_cast<uint32_t>(__cvta_generic_to_shared(tileA));
    const uint32_t sB = static_cast<uint32_t>(__cvta_generic_to_shared(tileB));

    // Fill the shared tile with columns [j0, j0 + tj); returns (max norm of the tile, all columns identical).
    // Layout: A[t] = (x_2t, x_2t+1, y_2t, y_2t+1), B[t] = (z_2t, z_2t+1, r_2t, r_2t+1).
    auto fill_tile = [&](int j0, int tj, float& rmax, bool& all_same) {
      const int npair = (tj + 1) >> 1;
      __syncthreads();
      float lmax = 0.0f;
      // A tile whose candidates are all bitwise identical (the reference's all-zero GT cloud of an empty
      // slice, npy2point.py:72,115) yields one P per row: only its first column can win under the
      // first-index rule, so the scan visits one column pair instead of tying tj times.
      const uint32_t f0 = __float_as_uint(__ldg(cbase + j0 * 3 + 0)), f1 = __float_as_uint(__ldg(cbase + j0 * 3 + 1)),
                     f2 = __float_as_uint(__ldg(cbase + j0 * 3 + 2));
      bool same = true;
      for (int t = threadIdx.x; t < npair; t += THREADS) {
        const int ja = j0 + 2 * t;
        const bool has_b = (2 * t + 1) < tj;
        const float xa = __ldg(cbase + ja * 3 + 0), ya = __ldg(cbase + ja * 3 + 1), za = __ldg(cbase + ja * 3 + 2);
        // odd tail: a sentinel column whose Q and P are +inf can never be selected
        const float xb = has_b ? __ldg(cbase + ja * 3 + 3) : 0.0f;
        const float yb = has_b ? __ldg(cbase + ja * 3 + 4) : 0.0f;
        const float zb = has_b ? __ldg(cbase + ja * 3 + 5) : 0.0f;
        const float ra = norm3(xa, ya, za);
        const float rb = has_b ? norm3(xb, yb, zb) : INFINITY;
        tileA[t] = make_float4(xa, xb, ya, yb);
        tileB[t] = make_float4(za, zb, ra, rb);
        // a NaN norm never enters the bound: that column's own Q is NaN and takes the slow path
        lmax = fmaxf(lmax, ra);
        if (has_b) lmax = fmaxf(lmax, rb);
        same = same && __float_as_uint(xa) == f0 && __float_as_uint(ya) == f1 && __float_as_uint(za) == f2 &&
               (!has_b || (__float_as_uint(xb) == f0 && __float_as_uint(yb) == f1 && __float_as_uint(zb) == f2));
      }
      if (threadIdx.x == 0) {                    // the prefetch slot past the last pair: a sentinel
        tileA[npair] = make_float4(0.f, 0.f, 0.f, 0.f);
        tileB[npair] = make_float4(0.f, 0.f, INFINITY, INFINITY);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, o));
      if (lane == 0) warp_rmax[wid] = lmax;
      all_same = __syncthreads_and(same) != 0;
      rmax = warp_rmax[0];
#pragma unroll
      for (int i = 1; i < THREADS / 32; ++i) rmax = fmaxf(rmax, warp_rmax[i]);
    };
    auto margin = [&](float ra, float rmax) {
      return __fadd_ru(__fmul_ru(kFilterMargin, __fadd_ru(ra, rmax)), kFilterAbs);
    };

    // ---- seed pass: an upper bound of every row's final minimum from every kSeedStride-th column pair
    if (nc >= kSeedMinCols && (nc <= kSeedMaxCols || seed_forced) && seed_stride > 0) {
      float qmin[R];
#pragma unroll
      for (int r = 0; r < R; ++r) qmin[r] = INFINITY;
      float rmax_all = 0.0f;
      for (int j0 = 0; j0 < nc; j0 += kTileJ) {
        const int tj = min(kTileJ, nc - j0);
        const int npair = (tj + 1) >> 1;
        float rmax;
        bool all_same;
        fill_tile(j0, tj, rmax, all_same);
        rmax_all = fmaxf(rmax_all, rmax);
#pragma unroll 2
        for (int t = 0; t < npair; t += seed_stride) {
          const float4 A = lds128(sA + 16 * t);
          const float4 Bv = lds128(sB + 16 * t);
          const float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w);
          const float2 cz = make_float2(Bv.x, Bv.y), cr = make_float2(Bv.z, Bv.w);
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float2 Q = __ffma2_rn(n2[r], cz, __ffma2_rn(n1[r], cy, __ffma2_rn(n0[r], cx, cr)));
            qmin[r] = fminf(qmin[r], fminf(Q.x, Q.y));      // NaN columns are ignored (they can never win)
          }
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        // P_j <= r_a + Q_j + 13u(r_a + r_j) at the sampled arg-min column j, hence >= the row's minimum P
        const float U = __fadd_ru(__fadd_ru(qmin[r], rq[r]), margin(rq[r], rmax_all));
        cap[r] = __fadd_ru(U, __fmul_ru(kFilterMargin, __fadd_ru(fabsf(U), kEps)));
      }
    }

    // ---- in-order scan
    for (int j0 = 0; j0 < nc; j0 += kTileJ) {
      const int tj = min(kTileJ, nc - j0);
      float rmax;
      bool all_same;
      fill_tile(j0, tj, rmax, all_same);
      const int scan_pairs = all_same ? 1 : (tj + 1) >> 1;
#pragma unroll
      for (int r = 0; r < R; ++r)
        thr[r] = __fadd_ru(__fsub_ru(fminf(bestP[r], cap[r]), rq[r]), margin(rq[r], rmax));
      uint32_t pa = sA, pb = sB;
      float4 A = lds128(pa), Bv = lds128(pb);
#pragma unroll 4
      for (int t = 0; t < scan_pairs; ++t, pa += 16, pb += 16) {
        const float4 An = lds128(pa + 16), Bn = lds128(pb + 16);   // next pair's operands: in flight during this one
        const float2 cx = make_float2(A.x, A.y), cy = make_float2(A.z, A.w);
        const float2 cz = make_float2(Bv.x, Bv.y), cr = make_float2(Bv.z, Bv.w);
        float2 Q[R];
        bool h[R];
        bool hit = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          Q[r] = __ffma2_rn(n2[r], cz, __ffma2_rn(n1[r], cy, __ffma2_rn(n0[r], cx, cr)));
          h[r] = !(Q[r].x >= thr[r]) | !(Q[r].y >= thr[r]);    // NaN -> slow path
          hit = hit | h[r];
        }
        if (hit) {
          const int j = j0 + 2 * t;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            if (h[r]) {
              // the reference's arithmetic, exactly as in v2 (a = -0.5 * (-2a) is exact).  The factor is
              // laundered so the compiler cannot hoist the unscaled rows out of the scan (12 more
              // registers per thread, i.e. spills or a lost CTA per SM).
              float hf = -0.5f;
              asm volatile("" : "+f"(hf));
              const float2 negh = make_float2(hf, hf);
              const float2 q0 = __fmul2_rn(n0[r], negh), q1 = __fmul2_rn(n1[r], negh), q2 = __fmul2_rn(n2[r], negh);
              const float2 zz = __ffma2_rn(q2, cz, __ffma2_rn(q1, cy, __fmul2_rn(q0, cx)));
              const float2 P = __ffma2_rn(neg2, zz, __fadd2_rn(make_float2(rq[r], rq[r]), cr));
              nn_update_lazy(P.x, j, bestP[r], bestI[r]);
              nn_update_lazy(P.y, j + 1, bestP[r], bestI[r]);
              thr[r] = __fadd_ru(__fsub_ru(fminf(bestP[r], cap[r]), rq[r]), margin(rq[r], rmax));
            }
          }
        }
        A = An;
        Bv = Bn;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int row = row0 + r * THREADS;
      if (row < nq) {
        // bestP = +inf only if no column was ever comparable (NaN input): v1/v2 then report +inf, index 0
        const float d = __fsqrt_rn(__fadd_rn(bestP[r], kEps));
        dout[row] = d;
        iout[row] = bestI[r];
        my_sum += static_cast<double>(d);
      }
    }
  }

  if (loss != nullptr) finish_loss<THREADS>(my_sum, ws, loss, dir, b, B, N);
}

// ---- backward ---------------------------------------------------------------------------------
// Pass 1 (scatter): many-to-one terms, accumulated as 64-bit fixed point so the sum is exact
// and order-independent.  Pass 2 (finalise): one-to-one term + conversion + scale.
__global__ void chamfer_bwd_scatter(const float* __restrict__ q, const float* __restrict__ c,
                                    const float* __restrict__ dq, const int64_t* __restrict__ iq,
                                    int nq, int nc, int B, long long* __restrict__ acc_c) {
  pdl_entry();
  // for every query point (b,i) with nearest candidate j: acc_c[b,j] += (c_j - q_i)/d_i
  const int64_t total = static_cast<int64_t>(B) * nq;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = g / nq;
    const int64_t j = iq[g];
    const float d = dq[g];
    const float* qp = q + g * 3;
    const float* cp = c + (b * nc + j) * 3;
    long long* a = acc_c + (b * nc + j) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float u = __fdiv_rn(__fsub_rn(cp[k], qp[k]), d);
      atomicAdd(reinterpret_cast<unsigned long long*>(a + k),
                static_cast<unsigned long long>(__double2ll_rn(static_cast<double>(u) * kFixScale)));
    }
  }
}

__global__ void chamfer_bwd_finalize(const float* __restrict__ q, const float* __restrict__ c,
                                     const float* __restrict__ dq, const int64_t* __restrict__ iq,
                                     const long long* __restrict__ acc_q,
                                     const float* __restrict__ grad_loss, int nq, int nc, int B,
                                     double inv_nb, float* __restrict__ grad_q) {
  pdl_entry();
  const int64_t total = static_cast<int64_t>(B) * nq;
  const double scale = static_cast<double>(__ldg(grad_loss)) * inv_nb;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t b = g / nq;
    const int64_t j = iq[g];
    const float d = dq[g];
    const float* qp = q + g * 3;
    const float* cp = c + (b * nc + j) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float u = __fdiv_rn(__fsub_rn(qp[k], cp[k]), d);
      const double many = static_cast<double>(acc_q[g * 3 + k]) * (1.0 / kFixScale);
      grad_q[g * 3 + k] = static_cast<float>(scale * (static_cast<double>(u) + many));
    }
  }
}

// Both of the above for a SMALL cloud in one launch (no memset, no global atomics): one CTA per cloud, the fixed-point
// accumulators of its np points in shared memory.  Same arithmetic, and integer sums do not depend on their order, so
// the result equals the two-kernel path bit for bit.  grad_p[i] = scale * ((p_i - o_{ip[i]}) / dp_i +
// sum_{j : io[j] == i} (p_i - o_j) / do_j).
constexpr int kSmallBwdPoints = 2048;          // 3 * 8 B * 2048 = 48 KB of shared memory
__global__ void __launch_bounds__(512)
chamfer_bwd_small_kernel(const float* __restrict__ p, const float* __restrict__ o, const float* __restrict__ dp,
                         const int64_t* __restrict__ ip, const float* __restrict__ d_o, const int64_t* __restrict__ io,
                         const float* __restrict__ grad_loss, int np, int no, double inv_nb, float* __restrict__ grad_p) {
  pdl_entry();
  extern __shared__ long long sacc[];          // [np][3]
  const int64_t b = blockIdx.x;
  for (int i = threadIdx.x; i < 3 * np; i += blockDim.x) sacc[i] = 0;
  __syncthreads();
  // Ground-truth clouds repeat points (and one cloud per batch is all zeros: every query's nearest neighbour is index
  // 0), so many lanes hit the same accumulator: a 64-bit shared-memory atomic is a compare-and-swap loop and 300
  // contenders on one address cost ~25 us.  The lanes of a warp that share a target first add their (integer)
  // contributions through shuffles; one lane per distinct target issues the atomic.
  const int lane = threadIdx.x & 31;
  for (int j0 = threadIdx.x - lane; j0 < no; j0 += blockDim.x) {       // warp-uniform trip count
    const int j = j0 + lane;
    const bool valid = j < no;
    const int64_t g = b * no + (valid ? j : 0);
    const int i = valid ? static_cast<int>(io[g]) : -1 - lane;          // invalid lanes: unique negative keys
    long long v[3] = {0, 0, 0};
    if (valid) {
      const float d = d_o[g];
      const float* op = o + g * 3;
      const float* pp = p + (b * np + i) * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float u = __fdiv_rn(__fsub_rn(pp[k], op[k]), d);
        v[k] = __double2ll_rn(static_cast<double>(u) * kFixScale);
      }
    }
    const unsigned mk = __match_any_sync(0xffffffffu, i);
    long long sum[3] = {0, 0, 0};
    for (int src = 0; src < 32; ++src) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const long long t = __shfl_sync(0xffffffffu, v[k], src);
        if ((mk >> src) & 1u) sum[k] += t;
      }
    }
    if (valid && lane == __ffs(mk) - 1) {
#pragma unroll
      for (int k = 0; k < 3; ++k)
        atomicAdd(reinterpret_cast<unsigned long long*>(sacc + i * 3 + k), static_cast<unsigned long long>(sum[k]));
    }
  }
  __syncthreads();
  const double scale = static_cast<double>(__ldg(grad_loss)) * inv_nb;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    const int64_t g = b * np + i;
    const int64_t j = ip[g];
    const float d = dp[g];
    const float* pp = p + g * 3;
    const float* op = o + (b * no + j) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float u = __fdiv_rn(__fsub_rn(pp[k], op[k]), d);
      const double many = static_cast<double>(sacc[i * 3 + k]) * (1.0 / kFixScale);
      grad_p[g * 3 + k] = static_cast<float>(scale * (static_cast<double>(u) + many));
    }
  }
}

struct Plan {
  int R, threads, tiles_x, tiles_y;
};

Plan make_plan(int B, int N, int M) {
  // Rows per thread / CTA size, from the measured table profiles/r2_chamfer_plan_probe.txt (22 shapes x 5 variants):
  //   * 4 rows per thread (each shared-memory broadcast amortised over 4 pair evaluations) pays once its grid —
  //     2 directions x B x ceil(rows / 512) CTAs — reaches ~512 CTAs AND the scan is long (>= 4096 columns):
  //     B32 x 4096, B16 x 8192, B8 x 16384 and up (5-10 % over the old rule at the boundary);
  //   * everything smaller runs 1 row per thread in 128-thread CTAs (most warps per SM: 11-18 % faster than the old
  //     2-row choice at B64 x 2048, B32 x 2048, B8 x 4096, B64 x 256), 64-thread CTAs while even that grid is < 1 wave;
  //   * 2 rows per thread never won a shape; 8 rows always lost.
  // pcuda_tune(1, rows * 1000 + threads) forces a variant (A/B runs, tests).
  const int forced = tuning(TUNE_CHAMFER_ROWS);
  auto plan = [&](int R, int T) { return Plan{R, T, (N + R * T - 1) / (R * T), (M + R * T - 1) / (R * T)}; };
  if (forced > 0) {
    const int cand[5][2] = {{8, 128}, {4, 128}, {2, 128}, {1, 128}, {1, 64}};
    for (int k = 0; k < 5; ++k) {
      if (forced / 100000 == 1 && cand[k][0] == 8) continue;  // the scalar kernel has no 8-row instantiation
      if (forced % 100000 == cand[k][0] * 1000 + cand[k][1]) return plan(cand[k][0], cand[k][1]);
    }
    return plan(1, 64);
  }
  const Plan p4 = plan(4, 128);
  if (static_cast<int64_t>(B) * (p4.tiles_x + p4.tiles_y) >= 512 && std::min(N, M) >= 4096) return p4;
  const Plan p1 = plan(1, 128);
  if (static_cast<int64_t>(B) * (p1.tiles_x + p1.tiles_y) >= sm_count()) return p1;
  return plan(1, 64);
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" size_t pcuda_chamfer_ws_bytes(int B, int N, int M) {
  if (B <= 0 || N <= 0 || M <= 0) return 16;
  // worst case one partial per 64 rows per direction
  const int64_t tiles = (static_cast<int64_t>(N > M ? N : M) + 63) / 64;
  return 16 + sizeof(double) * 2 * static_cast<size_t>(B) * static_cast<size_t>(tiles);
}

extern "C" int pcuda_chamfer_fwd(const float* x, const float* y, int B, int N, int M, float* d1,
                                 int64_t* i1, float* d2, int64_t* i2, float* loss, void* ws,
                                 pcuda_stream_t stream) {
  PCUDA_REQUIRE(B >= 0 && N >= 0 && M >= 0, PCUDA_E_SHAPE, "chamfer_fwd: bad shape B=%d N=%d M=%d", B, N, M);
  PCUDA_REQUIRE(B <= 65535, PCUDA_E_UNSUPPORTED, "chamfer_fwd: B=%d > 65535", B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B == 0 || N == 0 || M == 0) {
    // torch: min over an empty dimension raises; an empty batch gives 0/0 = NaN.  Surface the
    // degenerate case as an argument error instead of launching.
    PCUDA_REQUIRE(B == 0 && N > 0 && M > 0, PCUDA_E_SHAPE, "chamfer_fwd: empty point cloud (N=%d, M=%d)", N, M);
    if (loss) {
      const float nanv = __builtin_nanf("");
      cudaMemcpyAsync(loss, &nanv, sizeof(float), cudaMemcpyHostToDevice, st);
    }
    return 0;
  }
  PCUDA_REQUIRE(x && y && d1 && i1 && d2 && i2, PCUDA_E_NULL, "chamfer_fwd: NULL tensor");
  PCUDA_REQUIRE(!loss || ws, PCUDA_E_NULL, "chamfer_fwd: loss needs ws");
  const Plan p = make_plan(B, N, M);
  const dim3 grid(p.tiles_x > p.tiles_y ? p.tiles_x : p.tiles_y, B, 2);
  FwdWs* w = static_cast<FwdWs*>(ws);
#define PCUDA_LAUNCH_NN(RR, TT)                                                              \
  PCUDA_LAUNCH((chamfer_nn_kernel<RR, TT>), grid, TT, 0, st, x, y, N, M, d1, i1, d2, i2, w, loss,        \
                                                 p.tiles_x, p.tiles_y, B)
#define PCUDA_LAUNCH_NN3(RR, TT)                                                             \
  PCUDA_LAUNCH((chamfer_nn3_kernel<RR, TT>), grid, TT, 0, st, x, y, N, M, d1, i1, d2, i2, w, loss,       \
                                                  p.tiles_x, p.tiles_y, B, seed_stride, seed_tune > 0)
#define PCUDA_LAUNCH_NN3B(RR, TT, MB)                                                        \
  PCUDA_LAUNCH((chamfer_nn3_kernel<RR, TT, MB>), grid, TT, 0, st, x, y, N, M, d1, i1, d2, i2, w, loss,   \
                                                      p.tiles_x, p.tiles_y, B, seed_stride, seed_tune > 0)
#define PCUDA_LAUNCH_NN2(RR, TT)                                                             \
  PCUDA_LAUNCH((chamfer_nn2_kernel<RR, TT>), grid, TT, 0, st, x, y, N, M, d1, i1, d2, i2, w, loss,       \
                                                  p.tiles_x, p.tiles_y, B)
  // tuning value >= 200000 selects the packed exact kernel (v2), >= 100000 the scalar one (v1), for A/B runs
  const int variant = tuning(TUNE_CHAMFER_ROWS) / 100000;
  // seed pass: every seed_stride-th column pair (tuning key 5: 0 = default, < 0 = no seed pass)
  const int seed_tune = tuning(TUNE_CHAMFER_SEED);
  const int seed_stride = seed_tune == 0 ? kSeedStride : (seed_tune < 0 ? 0 : seed_tune);
  if (variant == 1) {
    if (p.R >= 4) PCUDA_LAUNCH_NN(4, 128);
    else if (p.R == 2) PCUDA_LAUNCH_NN(2, 128);
    else if (p.threads == 128) PCUDA_LAUNCH_NN(1, 128);
    else PCUDA_LAUNCH_NN(1, 64);
  } else if (variant == 2) {
    if (p.R == 8) PCUDA_LAUNCH_NN2(8, 128);
    else if (p.R == 4) PCUDA_LAUNCH_NN2(4, 128);
    else if (p.R == 2) PCUDA_LAUNCH_NN2(2, 128);
    else if (p.threads == 128) PCUDA_LAUNCH_NN2(1, 128);
    else PCUDA_LAUNCH_NN2(1, 64);
  } else if (variant == 3) {
    PCUDA_LAUNCH_NN3B(4, 128, 7);      // A/B: 72 registers (7 CTAs per SM) instead of 64 with a spilling slow path (+3 %)
  } else {
    if (p.R == 8) PCUDA_LAUNCH_NN3(8, 128);
    else if (p.R == 4) PCUDA_LAUNCH_NN3(4, 128);
    else if (p.R == 2) PCUDA_LAUNCH_NN3(2, 128);
    else if (p.threads == 128) PCUDA_LAUNCH_NN3(1, 128);
    else PCUDA_LAUNCH_NN3(1, 64);
  }
#undef PCUDA_LAUNCH_NN
#undef PCUDA_LAUNCH_NN2
#undef PCUDA_LAUNCH_NN3
#undef PCUDA_LAUNCH_NN3B
  count_launch(1);
  return check_launch("chamfer_nn_kernel");
}

extern "C" size_t pcuda_chamfer_bwd_ws_bytes(int B, int N, int M) {
  if (B <= 0) return 0;
  return sizeof(long long) * 3 * static_cast<size_t>(B) * (static_cast<size_t>(N > 0 ? N : 0) + static_cast<size_t>(M > 0 ? M : 0));
}

extern "C" int pcuda_chamfer_bwd(const float* x, const float* y, const float* d1,
                                 const int64_t* i1, const float* d2, const int64_t* i2,
                                 const float* grad_loss, int B, int N, int M, float* grad_x,
                                 float* grad_y, void* ws, pcuda_stream_t stream) {
  PCUDA_REQUIRE(B >= 0 && N >= 0 && M >= 0, PCUDA_E_SHAPE, "chamfer_bwd: bad shape B=%d N=%d M=%d", B, N, M);
  if (B == 0 || N == 0 || M == 0) return 0;
  if (!grad_x && !grad_y) return 0;
  PCUDA_REQUIRE(x && y && d1 && i1 && d2 && i2 && grad_loss && ws, PCUDA_E_NULL, "chamfer_bwd: NULL tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long* acc_x = static_cast<long long*>(ws);
  long long* acc_y = acc_x + static_cast<size_t>(B) * N * 3;
  const double inv_nb = 1.0 / (static_cast<double>(N) * static_cast<double>(B));
  const int threads = 256;
  auto blocks = [&](int64_t n) {
    int64_t g = (n + threads - 1) / threads;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
    return static_cast<int>(g < cap ? g : cap);
  };
  int launches = 0;
  const bool small_ok = !tuning(TUNE_CHAMFER_BWD_TWO_PASS);     // A/B and tests: key 11 forces memset + scatter + finalise
  if (grad_x && small_ok && N <= kSmallBwdPoints) {
    PCUDA_LAUNCH(chamfer_bwd_small_kernel, B, 512, sizeof(long long) * 3 * static_cast<size_t>(N), st, x, y, d1, i1, d2, i2, grad_loss, N, M,
                 inv_nb, grad_x);
    launches += 1;
  } else if (grad_x) {
    launches += 2;
    cudaMemsetAsync(acc_x, 0, sizeof(long long) * 3 * static_cast<size_t>(B) * N, st);
    // y_j -> nearest x_i contributes to grad_x[i2(j)]
    PCUDA_LAUNCH(chamfer_bwd_scatter, blocks(static_cast<int64_t>(B) * M), threads, 0, st, y, x, d2, i2, M, N, B, acc_x);
    PCUDA_LAUNCH(chamfer_bwd_finalize, blocks(static_cast<int64_t>(B) * N), threads, 0, st, x, y, d1, i1, acc_x, grad_loss, N, M, B, inv_nb, grad_x);
  }
  if (grad_y && small_ok && M <= kSmallBwdPoints) {
    PCUDA_LAUNCH(chamfer_bwd_small_kernel, B, 512, sizeof(long long) * 3 * static_cast<size_t>(M), st, y, x, d2, i2, d1, i1, grad_loss, M, N,
                 inv_nb, grad_y);
    launches += 1;
  } else if (grad_y) {
    launches += 2;
    cudaMemsetAsync(acc_y, 0, sizeof(long long) * 3 * static_cast<size_t>(B) * M, st);
    PCUDA_LAUNCH(chamfer_bwd_scatter, blocks(static_cast<int64_t>(B) * N), threads, 0, st, x, y, d1, i1, N, M, B, acc_y);
    PCUDA_LAUNCH(chamfer_bwd_finalize, blocks(static_cast<int64_t>(B) * M), threads, 0, st, y, x, d2, i2, acc_y, grad_loss, M, N, B, inv_nb, grad_y);
  }
  count_launch(launches);
  return check_launch("chamfer_bwd");
}
