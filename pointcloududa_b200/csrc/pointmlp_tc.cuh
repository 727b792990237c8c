// Host-side entry points of the tcgen05 shared-MLP kernels (pointmlp_tc.cu), called by the layer
// orchestration in pointmlp.cu when precision == PCUDA_MLP_BF16.
#pragma once
#include "pointmlp_common.cuh"

namespace pcuda {
namespace tc {

// layers whose contraction length the tensor-core kernels accept (multiple of 64, 64..512)
bool supports(int cin);

// y_l = W a_{l-1} (+ bias): stores y (dense) or reduces to per-(cloud,channel) arg-max keys (pool);
// accumulates bias-free per-channel sum / sum of squares into stats[0..C), stats[C..2C).
int fwd_layer(const ActSrc& src, const pcuda_mlp_layer_t& L, bool pool, int B, int N, double* stats,
              unsigned long long* keys, cudaStream_t st);

// da_{l-1} = dy_l W_l, then the ReLU mask / BN sums of layer l-1 (or grad_x when l == 0)
int dgrad_layer(const DySrc& dys, const float* W, int B, int N, const DgradOut& out, cudaStream_t st);

// pooled layer: da = S - u - a Q with the sparse part S given as per-point channel lists
int pool_dgrad(const ActSrc& src, const float* Q, const float* u, const float* Wpool, const float* coef,
               const int* head, const int* next, int Cpool, int B, int N, const DgradOut& out, cudaStream_t st);

// contraction over points.  partial: [S, C, Kr] fp32 (one slice per point-range split)
int pt_splits(int64_t M, int rblocks);
bool pt_supports(int C, int Kr, bool gram);
int wgrad_layer(const DySrc& dys, const ActSrc& prev, int64_t M, int S, float* partial, cudaStream_t st);
int gram(const ActSrc& act, int64_t M, int S, float* partial, double* colsum, cudaStream_t st);

}  // namespace tc
}  // namespace pcuda
