// Host-side entry points of the tcgen05 shared-MLP kernels (pointmlp_tc.cu), called by the layer
// orchestration in pointmlp.cu when precision == PCUDA_MLP_BF16.
#pragma once
#include "pointmlp_common.cuh"

namespace pcuda {
namespace tc {

// 256-point tiles that never straddle a cloud: tile t -> cloud t / tpc, points (t % tpc) * 256 ...
struct Tiling {
  int B, N, tpc, n_tiles;
};
Tiling make_tiling(int B, int N);

// layers whose contraction length the tensor-core kernels accept (multiple of 64, 64..512)
bool supports(int cin);
// ... and whose operands fit the kernels' shared-memory plan (otherwise the FP32 kernels run)
bool fwd_fits(int cout, int cin);
bool dgrad_fits(int Kp, int C, bool with_side);
bool pool_dgrad_fits(int K);

// bf16 operand copies in the swizzled slab format (tc_common.cuh)
size_t act_pack_bytes(const Tiling& tl, int C);          // [n_tiles][C/64] slabs of 256 rows
size_t w_pack_bytes(int R, int K);                       // ceil(R/128) blocks of 128 rows x K
// raw != nullptr: the BatchNorm of src's layer is finalised here from its raw sums (see BnRaw)
int pack_act(const ActSrc& src, const Tiling& tl, uint8_t* out, cudaStream_t st, const BnRaw* raw = nullptr);
int pack_dy(const DySrc& dys, const Tiling& tl, uint8_t* out, cudaStream_t st);
// transposed == false: A[r,k] = W[r*K + k];  true: A[r,k] = W[k*R + r]
// sign_src (non-transposed only): rows with sign_src[r] < 0 are negated (pooled forward layer)
int pack_w(const float* W, int R, int K, bool transposed, uint8_t* out, cudaStream_t st, const float* sign_src = nullptr);
// Q = sum_z partial[z] ([S][K*K] split partials, symmetric) -> Q (fp32), qpack (= pack_w(Q, K, K, false)), and
// u[k] = sum_z upartial[z][k] - sum_k2 Q[k,k2] abar[k2]   (upartial[z][k] = sum_{c in slice z} alpha_c W[c,k]), in one launch
int q_finish(const float* partial, const float* upartial, int S, const float* abar, int K, float* Q, uint8_t* qpack, float* u,
             cudaStream_t st);

// y_l = W a_{l-1} (+ bias): stores y (dense) or reduces to per-(cloud,channel) arg-max keys (pool);
// accumulates bias-free per-channel sum / sum of squares into stats[0..C), stats[C..2C).
int fwd_layer(const Tiling& tl, const uint8_t* a_pack, const uint8_t* w_pack, const pcuda_mlp_layer_t& L, bool pool,
              double* stats, unsigned long long* keys, cudaStream_t st);

// da_{l-1} = dy_l W_l, then the ReLU mask / BN sums of layer l-1 (mask source: its packed activation,
// or y_prev in `out` when aprev_pack == nullptr), or grad_x when out.grad_x != nullptr
int dgrad_layer(const Tiling& tl, const uint8_t* dy_pack, int C, const uint8_t* wt_pack, const uint8_t* aprev_pack,
                const DgradOut& out, cudaStream_t st);

// pooled layer, dense part: da = -u - a Q (the sparse rows S are added afterwards by pool_sparse_kernel)
int pool_dgrad(const Tiling& tl, const uint8_t* a_pack, int K, const uint8_t* q_pack, const float* u, const DgradOut& out,
               cudaStream_t st);

// contraction over points.  partial: [S, C, Kr] fp32 (one slice per tile-range split)
int pt_splits(const Tiling& tl, int rblocks);
bool pt_supports(int C, int Kr, bool gram);
int wgrad_layer(const Tiling& tl, const uint8_t* dy_pack, int C, const uint8_t* aprev_pack, int Kr, int S, float* partial,
                cudaStream_t st);
int gram(const Tiling& tl, const uint8_t* a_pack, int K, int S, float* partial, cudaStream_t st);

}  // namespace tc
}  // namespace pcuda
