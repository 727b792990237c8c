/*
 * pcuda_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference algorithms for the entropy map and the Chamfer / NN loss of
 * sulaimanvesal/PointCloudUDA.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product path (pointcloududa_b200/) never does.
 *
 * Parity status: the reference ships no golden vectors or tests for this path.
 * The oracle is therefore pinned against OUTPUTS OF THE REFERENCE ITSELF, produced by importing
 * /root/reference/src/utils/loss.py and restating the inline entropy expressions literally
 * (oracle/gen_golden.py -> tests/golden/*.npz); tests/test_oracle_golden.py checks it.
 *
 * Citations are to /root/reference/src/.
 *
 * Build: gcc -O2 -std=c11 -fPIC -shared -fopenmp -mavx2 -mfma -ffp-contract=off pcuda_oracle.c -lm
 *        (-ffp-contract=off: every rounding below is deliberate; fmaf() is the only fused op)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Entropy map.
 *   train_mscmrseg.py:222   -1.0 * torch.sigmoid(oT) * torch.log(torch.sigmoid(oT) + smooth)
 *   train_mmwhs.py:213,216  predS = F.softmax(oS, dim=1) | F.sigmoid(oS)
 *   train_mmwhs.py:224      -1.0 * predS * torch.log(predS + smooth) / math.log(c)
 *   train_mmwhs.py:225      torch.mean(torch.sum(uncertainty_mapS, dim=1))
 * activation: 0 sigmoid, 1 softmax.  norm: 0 -> no division, else divide by `norm` (= ln C).
 * z, m, p: [B, C, HW]; p and mean_out may be NULL.
 */
void oracle_entropy_fwd(const float* z, float* m, float* p, float* mean_out, int B, int C,
                        int64_t HW, int activation, float norm, float smooth) {
  double total = 0.0;
#pragma omp parallel for reduction(+ : total) schedule(static) collapse(2)
  for (int64_t b = 0; b < (int64_t)B; ++b)
  for (int64_t px = 0; px < HW; ++px) {
    const float* zp = z + b * C * HW + px;
    float mx = -INFINITY, sum = 0.0f;
    if (activation == 1) {
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, zp[c * HW]);
      for (int c = 0; c < C; ++c) sum += expf(zp[c * HW] - mx);
    }
    for (int c = 0; c < C; ++c) {
      const float zz = zp[c * HW];
      /* ATen: softmax = exp(x - max) / sum ; sigmoid = 1 / (1 + exp(-x)) */
      const float pr = activation == 1 ? expf(zz - mx) / sum : 1.0f / (1.0f + expf(-zz));
      float e = (-1.0f * pr) * logf(pr + smooth);
      if (norm != 0.0f) e = e / norm;
      const int64_t o = b * C * HW + c * HW + px;
      m[o] = e;
      if (p) p[o] = pr;
      total += (double)e;
    }
  }
  if (mean_out) *mean_out = (float)(total / ((double)B * (double)HW));
}

/* Analytic backward in double, used as the fp64 truth for gradient checks.
 * grad_m, grad_p may be NULL; grad_mean is the upstream gradient of the mean scalar. */
void oracle_entropy_bwd(const float* z, const float* grad_m, const float* grad_p, float grad_mean,
                        double* grad_z, int B, int C, int64_t HW, int activation, float norm,
                        float smooth) {
  const double k = norm != 0.0f ? 1.0 / (double)norm : 1.0;
  const double gs = (double)grad_mean / ((double)B * (double)HW);
  const double s = (double)smooth;
#pragma omp parallel for schedule(static) collapse(2)
  for (int64_t b = 0; b < (int64_t)B; ++b)
  for (int64_t px = 0; px < HW; ++px) {
    const int64_t base = b * C * HW + px;
    double pr[64], t[64];
    double mx = -INFINITY, sum = 0.0, dot = 0.0;
    if (activation == 1) {
      for (int c = 0; c < C; ++c) mx = fmax(mx, (double)z[base + c * HW]);
      for (int c = 0; c < C; ++c) sum += exp((double)z[base + c * HW] - mx);
    }
    for (int c = 0; c < C; ++c) {
      const double zz = (double)z[base + c * HW];
      pr[c] = activation == 1 ? exp(zz - mx) / sum : 1.0 / (1.0 + exp(-zz));
      const double g = gs + (grad_m ? (double)grad_m[base + c * HW] : 0.0);
      t[c] = -k * g * (log(pr[c] + s) + pr[c] / (pr[c] + s));
      if (grad_p) t[c] += (double)grad_p[base + c * HW];
      dot += t[c] * pr[c];
    }
    for (int c = 0; c < C; ++c)
      grad_z[base + c * HW] = activation == 1 ? pr[c] * (t[c] - dot) : t[c] * pr[c] * (1.0 - pr[c]);
  }
}

/* ------------------------------------------------------------------------------------------
 * Chamfer / NN loss.   utils/loss.py:40-76
 *
 *   xx = bmm(x, x^T); yy = bmm(y, y^T); zz = bmm(x, y^T)            (:56-58)
 *   rx = diag(xx), ry = diag(yy)                                    (:59-63)
 *   P  = rx^T + ry - 2*zz                                           (:64)
 *   dist = sqrt(P + 0.00001); values, indices = dist.min(dim=2)     (:68-72)
 *
 * A K=3 sgemm accumulates k = 0,1,2 in order with fused multiply-adds, i.e.
 *   dot(a,b) = fmaf(a2,b2, fmaf(a1,b1, a0*b0))
 * (verified against torch.bmm on CPU: 0 mismatches, see oracle/gen_golden.py), and
 * P = fl(fl(rx+ry) - fl(2*zz)) where 2*zz is exact.
 */
static inline float dot3(const float* a, const float* b) {
  return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
}

/* P[i*M + j] for one sample: x [N,3], y [M,3] */
void oracle_pairwise_dist(const float* x, const float* y, int N, int M, float* P) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; ++i) {
    const float rx = dot3(x + 3 * i, x + 3 * i);
    for (int j = 0; j < M; ++j) {
      const float ry = dot3(y + 3 * j, y + 3 * j);
      const float zz = dot3(x + 3 * i, y + 3 * j);
      const float t = rx + ry;
      P[(int64_t)i * M + j] = t - 2.0f * zz;
    }
  }
}

/* nearest neighbour of every q_i among c_j: torch.min(dim) returns the FIRST minimal index.
 * The distance row is evaluated into a buffer (vectorisable: fmaf and sqrtf are IEEE per lane),
 * then scanned in ascending j with a strict '<'. */
static void nn_search(const float* q, const float* c, int nq, int nc, float* d, int64_t* idx) {
  float* c0 = (float*)malloc(sizeof(float) * (size_t)nc * 5);
  float *c1 = c0 + nc, *c2 = c1 + nc, *rc = c2 + nc, *row = rc + nc;
  for (int j = 0; j < nc; ++j) {
    c0[j] = c[3 * j]; c1[j] = c[3 * j + 1]; c2[j] = c[3 * j + 2];
    rc[j] = dot3(c + 3 * j, c + 3 * j);
  }
  for (int i = 0; i < nq; ++i) {
    const float q0 = q[3 * i], q1 = q[3 * i + 1], q2 = q[3 * i + 2];
    const float rq = dot3(q + 3 * i, q + 3 * i);
#pragma omp simd
    for (int j = 0; j < nc; ++j) {
      const float zz = fmaf(q2, c2[j], fmaf(q1, c1[j], q0 * c0[j]));
      const float t = rq + rc[j];
      const float P = t - 2.0f * zz;
      row[j] = sqrtf(P + 0.00001f);
    }
    float best = INFINITY;
    int64_t bi = 0;
    for (int j = 0; j < nc; ++j) {
      if (row[j] < best) {
        best = row[j];
        bi = j;
      }
    }
    d[i] = best;
    idx[i] = bi;
  }
  free(c0);
}

/* x [B,N,3], y [B,M,3] -> d1,i1 [B,N]; d2,i2 [B,M]; loss scalar (loss.py:73-75) */
void oracle_chamfer_fwd(const float* x, const float* y, int B, int N, int M, float* d1,
                        int64_t* i1, float* d2, int64_t* i2, float* loss) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int t = 0; t < 2 * B; ++t) {
    const int b = t >> 1;
    if ((t & 1) == 0)
      nn_search(x + (int64_t)b * N * 3, y + (int64_t)b * M * 3, N, M, d1 + (int64_t)b * N, i1 + (int64_t)b * N);
    else
      nn_search(y + (int64_t)b * M * 3, x + (int64_t)b * N * 3, M, N, d2 + (int64_t)b * M, i2 + (int64_t)b * M);
  }
  if (loss) {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) {
      double s1 = 0.0, s2 = 0.0;
      for (int i = 0; i < N; ++i) s1 += (double)d1[(int64_t)b * N + i];
      for (int j = 0; j < M; ++j) s2 += (double)d2[(int64_t)b * M + j];
      acc += s1 / (double)N + s2 / (double)N; /* both divided by x.size(1) (loss.py:73-74) */
    }
    *loss = (float)(acc / (double)B);
  }
}

/* fp64 analytic gradient THROUGH GIVEN INDICES AND FORWARD DISTANCES.
 * autograd differentiates d = sqrt(P + 1e-5) with the fp32 forward value of d in the denominator
 * (d sqrt(u) = du / (2 d)) and dP/dx_i = 2 x_i - 2 y_j, so the exact derivative of what the
 * reference computed is (x_i - y_j) / d_fp32.  d1/d2 may be NULL: then d is re-evaluated in
 * double as sqrt(|x-y|^2 + 1e-5) (the derivative of the exact-arithmetic loss). */
void oracle_chamfer_bwd(const float* x, const float* y, const float* d1, const int64_t* i1,
                        const float* d2, const int64_t* i2, double grad_loss, int B, int N, int M,
                        double* grad_x, double* grad_y) {
  const double scale = grad_loss / ((double)N * (double)B);
  const double eps = (double)0.00001f;
  if (grad_x) memset(grad_x, 0, sizeof(double) * (size_t)B * N * 3);
  if (grad_y) memset(grad_y, 0, sizeof(double) * (size_t)B * M * 3);
  for (int b = 0; b < B; ++b) {
    const float* xb = x + (int64_t)b * N * 3;
    const float* yb = y + (int64_t)b * M * 3;
    for (int i = 0; i < N; ++i) {
      const int64_t j = i1[(int64_t)b * N + i];
      double df[3], sq = 0.0;
      for (int k = 0; k < 3; ++k) { df[k] = (double)xb[3 * i + k] - (double)yb[3 * j + k]; sq += df[k] * df[k]; }
      const double d = d1 ? (double)d1[(int64_t)b * N + i] : sqrt(sq + eps);
      for (int k = 0; k < 3; ++k) {
        if (grad_x) grad_x[((int64_t)b * N + i) * 3 + k] += scale * df[k] / d;
        if (grad_y) grad_y[((int64_t)b * M + j) * 3 + k] -= scale * df[k] / d;
      }
    }
    for (int j = 0; j < M; ++j) {
      const int64_t i = i2[(int64_t)b * M + j];
      double df[3], sq = 0.0;
      for (int k = 0; k < 3; ++k) { df[k] = (double)xb[3 * i + k] - (double)yb[3 * j + k]; sq += df[k] * df[k]; }
      const double d = d2 ? (double)d2[(int64_t)b * M + j] : sqrt(sq + eps);
      for (int k = 0; k < 3; ++k) {
        if (grad_x) grad_x[((int64_t)b * N + i) * 3 + k] += scale * df[k] / d;
        if (grad_y) grad_y[((int64_t)b * M + j) * 3 + k] -= scale * df[k] / d;
      }
    }
  }
}
