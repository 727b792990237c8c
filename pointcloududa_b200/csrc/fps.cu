// Farthest-point sampling of the ground-truth boundary clouds:
//   utils/npy2point.py:11-18  graipher(pts, K, dim) — start at a given point, then K-1 times take the point whose
//   squared distance to the chosen set is largest (np.argmax: first index on ties) and fold its distances into the
//   running minimum (np.minimum).  Called per slice by the data generators (utils/npy2point.py:101-125,
//   data_generator_mmwhs.py:256-264) on the marching-cubes vertices of the mask.
// The reference runs it in numpy on the data thread: O(K V) float64 operations per slice.  Here one CTA owns one cloud:
// the running minima live in shared memory (float64, like numpy), every iteration is one pass over the V points plus
// a block-wide (value, first index) arg-max.  Arithmetic is the reference's: ((dx*dx + dy*dy) + dz*dz) in float64,
// so the selected indices are bit-identical for any input numpy would take.
#include "pcuda_common.cuh"

namespace pcuda {
namespace {

constexpr int kFpsThreads = 1024;

__global__ void __launch_bounds__(kFpsThreads)
fps_kernel(const double* __restrict__ pts, const int32_t* __restrict__ counts, const int32_t* __restrict__ starts, int Vmax, int K,
           int dim, double* __restrict__ out_pts, int32_t* __restrict__ out_idx) {
  pdl_entry();
  extern __shared__ double dist[];                    // [Vmax]
  __shared__ double red_v[kFpsThreads / 32];
  __shared__ int red_i[kFpsThreads / 32];
  __shared__ int s_sel;
  const int b = blockIdx.x;
  const int V = counts ? counts[b] : Vmax;
  const double* P = pts + static_cast<int64_t>(b) * Vmax * dim;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int sel = starts ? starts[b] : 0;
  if (V <= 0) {                                       // empty cloud: K zero rows (the reference's `vertices = np.zeros`)
    for (int i = threadIdx.x; i < K * dim; i += kFpsThreads) out_pts[static_cast<int64_t>(b) * K * dim + i] = 0.0;
    for (int i = threadIdx.x; i < K; i += kFpsThreads) out_idx[static_cast<int64_t>(b) * K + i] = -1;
    return;
  }
  sel = min(max(sel, 0), V - 1);
  for (int k = 0; k < K; ++k) {
    const double p0 = P[static_cast<int64_t>(sel) * dim], p1 = dim > 1 ? P[static_cast<int64_t>(sel) * dim + 1] : 0.0,
                 p2 = dim > 2 ? P[static_cast<int64_t>(sel) * dim + 2] : 0.0;
    if (threadIdx.x == 0) {
      out_idx[static_cast<int64_t>(b) * K + k] = sel;
      double* o = out_pts + (static_cast<int64_t>(b) * K + k) * dim;
      o[0] = p0;
      if (dim > 1) o[1] = p1;
      if (dim > 2) o[2] = p2;
    }
    if (k + 1 == K) break;
    // distances = (k == 0) ? d : minimum(distances, d); arg-max with the first index on ties
    double bv = -1.0;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < V; i += kFpsThreads) {
      // every product and sum rounded separately (numpy squares element-wise, then adds along the axis): no FMA contraction
      const double dx = p0 - P[static_cast<int64_t>(i) * dim];
      double d = __dmul_rn(dx, dx);
      if (dim > 1) { const double dy = p1 - P[static_cast<int64_t>(i) * dim + 1]; d = __dadd_rn(d, __dmul_rn(dy, dy)); }
      if (dim > 2) { const double dz = p2 - P[static_cast<int64_t>(i) * dim + 2]; d = __dadd_rn(d, __dmul_rn(dz, dz)); }
      const double cur = k == 0 ? d : fmin(dist[i], d);
      dist[i] = cur;
      if (cur > bv) { bv = cur; bi = i; }            // i ascends within a thread: strict > keeps the first
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      bv = lane < kFpsThreads / 32 ? red_v[lane] : -1.0;
      bi = lane < kFpsThreads / 32 ? red_i[lane] : 0x7fffffff;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
      }
      if (lane == 0) s_sel = bi;
    }
    __syncthreads();
    sel = s_sel;
  }
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

extern "C" int pcuda_fps(const double* pts, const int32_t* counts, const int32_t* starts, int B, int Vmax, int K, int dim,
                         double* out_pts, int32_t* out_idx, pcuda_stream_t stream) {
  PCUDA_REQUIRE(B >= 0 && Vmax >= 0 && K >= 1 && dim >= 1 && dim <= 3, PCUDA_E_SHAPE, "fps: bad shape B=%d V=%d K=%d dim=%d", B, Vmax, K, dim);
  if (B == 0) return 0;
  PCUDA_REQUIRE(pts && out_pts && out_idx, PCUDA_E_NULL, "fps: NULL argument");
  PCUDA_REQUIRE(Vmax >= 1, PCUDA_E_SHAPE, "fps: Vmax = 0");
  const size_t smem = sizeof(double) * static_cast<size_t>(Vmax);
  PCUDA_REQUIRE(smem <= 200 * 1024, PCUDA_E_UNSUPPORTED, "fps: %d points per cloud exceed the shared-memory distance buffer (25600)", Vmax);
  if (cudaError_t e = smem_optin(fps_kernel, 200 * 1024)) return fail(static_cast<int>(e), "fps: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  PCUDA_LAUNCH(fps_kernel, B, kFpsThreads, smem, static_cast<cudaStream_t>(stream), pts, counts, starts, Vmax, K, dim, out_pts, out_idx);
  count_launch(1);
  return check_launch("fps");
}
