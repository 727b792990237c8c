/*
 * pcuda.h — C-ABI of libpcuda.so: the B200 (sm_100a) kernels behind PointCloudUDA's
 * adversarial-adaptation hot path.
 *
 * The reference (sulaimanvesal/PointCloudUDA) has no FFI of its own: the boundary is three
 * Python-level call sites.  Each entry point below names the reference code it replaces
 * (paths relative to the reference's src/):
 *
 *   pcuda_entropy_fwd / _bwd      inline expressions  train_mscmrseg.py:222,265
 *                                                     train_mmwhs.py:213-217,224-225,240-243
 *   pcuda_chamfer_fwd / _bwd      utils/loss.py:40-76  (batch_NN_loss, batch_pairwise_dist)
 *   pcuda_pointmlp_fwd / _bwd     networks/PointNetCls.py:38-44 (STN3d trunk), :84-88 (STNkd
 *                                 trunk), :143-163 (PointNetfeat trunk + global max-pool)
 *   pcuda_pointmlp_fwd_xf/_bwd_xf the same with the 3x3 input transform fused in: networks/PointNetCls.py:140-142
 *   pcuda_point_transform_fwd/_bwd the 64x64 feature transform, networks/PointNetCls.py:147-151
 *   pcuda_fcstack_fwd / _bwd      networks/PointNetCls.py:46-62 (STN3d head), :89-101 (STNkd head),
 *                                 :208-213 (classifier head): Linear [+Dropout] [+BatchNorm1d] [+ReLU]
 *   pcuda_bn_running_update       BatchNorm1d running statistics of several forward passes, in pass order
 *   pcuda_bce_logits              F.binary_cross_entropy_with_logits + accuracy, train_mscmrseg.py:233,286-296,316-322
 *   pcuda_grad_sum_pack /         .grad accumulation of the two D4 backward passes (train_mscmrseg.py:288,319) and
 *   pcuda_sgd_momentum_step /     optim_dis4.step() = torch.optim.SGD(momentum, weight_decay) (:329-330,:450-455)
 *   pcuda_sgd_momentum_sum_step
 *   pcuda_fps                     utils/npy2point.py:11-18 (graipher: farthest-point sampling of the GT boundary clouds)
 *   pcuda_comm_*                  no counterpart (the reference is single-process): sum of D4's parameter gradients over
 *                                 the batch-sharded ranks in front of optim_dis4.step()
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented "host";
 *   - the caller allocates inputs, outputs and workspaces and keeps them alive until the
 *     stream has passed the launch; the library owns no device memory except what a communicator
 *     (pcuda_comm_init ... pcuda_comm_destroy) holds: its NCCL communicator and its peer-memory region;
 *   - launches are asynchronous on `stream`; no entry point synchronises, allocates or frees,
 *     so all of them are legal under CUDA-graph stream capture (pcuda_pointmlp_fwd / _bwd
 *     fork onto an internal auxiliary stream and join back before returning; under capture the
 *     fork / join events become graph dependencies);
 *   - return value: 0 on success; a negative PCUDA_E* code for a rejected argument;
 *     a positive value is a cudaError_t from the launch.  Nothing throws or exits.
 *     pcuda_last_error_string() gives a thread-local description of the last failure.
 */
#ifndef PCUDA_H_
#define PCUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCUDA_VERSION 100 /* major*100 + minor */

typedef void* pcuda_stream_t; /* a cudaStream_t */

enum {
  PCUDA_OK = 0,
  PCUDA_E_NULL = -1,        /* required pointer is NULL */
  PCUDA_E_SHAPE = -2,       /* negative / inconsistent extent */
  PCUDA_E_UNSUPPORTED = -3, /* valid request this build has no kernel for */
  PCUDA_E_ALIGN = -4,       /* pointer not aligned as documented */
  PCUDA_E_WORKSPACE = -5    /* workspace too small */
};

enum { PCUDA_ACT_SIGMOID = 0, PCUDA_ACT_SOFTMAX = 1 };

int pcuda_version(void);
const char* pcuda_last_error_string(void);
const char* pcuda_error_name(int code);
/* SM count of the current device (cached); used by callers to size benchmarks. */
int pcuda_sm_count(void);
/* Benchmark / fault-isolation knob for A/B-ing kernel variants (key, value); not part of the
 * reference-facing contract.  Keys: 0 = entropy arithmetic (0 MUFU fast path, 1 libdevice precise),
 * 1 = chamfer rows*1000+threads (default: prefiltered packed kernel; +100000: scalar, +200000: exact packed,
 * +300000: prefiltered at 72 registers), 2 = force the FP32 MLP kernels,
 * 3 = bit mask of MLP pieces switched from tcgen05 back to FP32 (1 forward, 2 pooled dgrad,
 * 4 dense dgrad, 8 wgrad, 16 Gram), 4 = timing experiments of the pooled tensor-core layer (results invalid),
 * 5 = chamfer seed-pass stride (0 default, < 0 off), 6 = no auxiliary-stream fork / finalise-on-read in the MLP,
 * 7 = CTAs of the peer-memory all-reduce (0 default), 8 = programmatic dependent launch (0 default: every launch; 1: off;
 * 3: launches of >= 2 waves of CTAs; 4: only kernels with a prologue before their wait), 9 = small fp64 sums through NCCL instead of the peer-memory mailbox,
 * 10 = pooled dgrad, sparse rows: < 0 switches the sorted / evenly cut kernel off (one warp per selected point instead),
 * 11 = Chamfer backward always as memset + scatter + finalise (default: one launch per gradient for clouds of <= 2048 points). */
int pcuda_tune(int key, int value);
/* Number of kernels this library has launched in this process (monotone; for bench accounting). */
uint64_t pcuda_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Entropy / self-information map.
 *   p = sigmoid(z)            (PCUDA_ACT_SIGMOID, train_mscmrseg.py:222)
 *   p = softmax(z, dim=C)     (PCUDA_ACT_SOFTMAX, train_mmwhs.py:213,240)
 *   m = ((-1*p) * log(p + smooth)) * inv_norm     inv_norm = 1 or 1/ln(C) (train_mmwhs.py:224)
 *   mean_out = mean_{b,hw} sum_c m                (train_mmwhs.py:225,243), optional
 * z, m, p: contiguous [B, C, HW] fp32 (NCHW with H*W flattened).  p and mean_out may be NULL.
 * ws: PCUDA_ENTROPY_WS_BYTES bytes, zero-filled once by the caller at allocation; the kernel
 *     leaves it zeroed again.  Required only when mean_out != NULL.
 * C in [1, 16].
 */
#define PCUDA_ENTROPY_WS_BYTES 16
int pcuda_entropy_fwd(const float* z, float* m, float* p, float* mean_out, void* ws, int B,
                      int C, int64_t HW, int activation, float inv_norm, float smooth,
                      pcuda_stream_t stream);

/* Backward of the above w.r.t. z.
 *   grad_m   [B,C,HW]  upstream gradient of the map           (may be NULL = zeros)
 *   grad_p   [B,C,HW]  upstream gradient of p (D1 branch)     (may be NULL)
 *   grad_mean  device scalar, upstream gradient of mean_out   (may be NULL)
 *   grad_z   [B,C,HW]  output
 */
int pcuda_entropy_bwd(const float* z, const float* grad_m, const float* grad_p,
                      const float* grad_mean, float* grad_z, int B, int C, int64_t HW,
                      int activation, float inv_norm, float smooth, pcuda_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Chamfer / nearest-neighbour loss  (utils/loss.py:40-76).
 *   P[b,i,j] = fl(fl(|x_i|^2 + |y_j|^2) - 2 * x_i.y_j)   with every dot product evaluated as
 *              fmaf(a2,b2, fmaf(a1,b1, a0*b0))            (what bmm with K=3 computes)
 *   d        = sqrtf(P + 1e-5f)
 *   d1[b,i], i1[b,i] = min / first arg-min over j         (loss.py:68-69)
 *   d2[b,j], i2[b,j] = min / first arg-min over i         (loss.py:71-72)
 *   loss     = sum_b (sum_i d1)/N /B + sum_b (sum_j d2)/N /B      (loss.py:73-75; the
 *              reference divides both terms by x.size(1) = N)
 * x [B,N,3], y [B,M,3] contiguous fp32.  i1/i2 are int64 (torch.min index dtype).
 * ws: pcuda_chamfer_ws_bytes(B,N,M) bytes; zero-filled once by the caller at allocation, left
 *     zeroed by every call.  loss may be NULL (then ws may be NULL too).
 */
size_t pcuda_chamfer_ws_bytes(int B, int N, int M);
int pcuda_chamfer_fwd(const float* x, const float* y, int B, int N, int M, float* d1,
                      int64_t* i1, float* d2, int64_t* i2, float* loss, void* ws,
                      pcuda_stream_t stream);

/* Backward through fixed indices:
 *   grad_x[b,i] = G/(N*B) * [ (x_i - y_{i1(i)})/d1_i + sum_{j: i2(j)=i} (x_i - y_j)/d2_j ]
 *   grad_y[b,j] = G/(N*B) * [ (y_j - x_{i2(j)})/d2_j + sum_{i: i1(i)=j} (y_j - x_i)/d1_i ]
 * G = *grad_loss (device scalar).  grad_x / grad_y may each be NULL.
 * The many-to-one sums are accumulated in 64-bit fixed point (2^-40 resolution), so the result
 * is independent of scheduling order.
 * ws: pcuda_chamfer_bwd_ws_bytes(B,N,M) bytes, any contents.
 */
size_t pcuda_chamfer_bwd_ws_bytes(int B, int N, int M);
int pcuda_chamfer_bwd(const float* x, const float* y, const float* d1, const int64_t* i1,
                      const float* d2, const int64_t* i2, const float* grad_loss, int B, int N,
                      int M, float* grad_x, float* grad_y, void* ws, pcuda_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * PointNet shared MLP (+ global max-pool): a stack of L layers
 *     y_l = W_l a_{l-1} + b_l ; z_l = BatchNorm1d(y_l) ; a_l = relu_l ? max(z_l,0) : z_l
 * applied independently to each of the B*N points (Conv1d with kernel_size 1,
 * networks/PointNetCls.py:41-43, :84-86, :143-161), followed — when pool != 0 — by
 * max over the N points of each cloud (:44, :87, :162).
 *
 * BatchNorm1d semantics (train != 0): per-channel mean and biased variance over all B*N
 * columns; running_mean/var updated with `momentum` and the unbiased variance, exactly as
 * torch.nn.BatchNorm1d does.  train == 0 normalises with running_mean/var and updates nothing.
 */
typedef struct pcuda_mlp_layer {
  int32_t cin, cout;
  int32_t relu;          /* apply ReLU after BN */
  int32_t reserved;
  const float* weight;   /* [cout, cin] */
  const float* bias;     /* [cout] or NULL */
  const float* gamma;    /* [cout] (BN weight) */
  const float* beta;     /* [cout] (BN bias) */
  float* running_mean;   /* [cout] or NULL */
  float* running_var;    /* [cout] or NULL */
  float* save_mean;      /* [cout]  fwd: out, bwd: in */
  float* save_invstd;    /* [cout]  fwd: out, bwd: in */
  float* y;              /* [B*N, cout] pre-BN activations, point-major. fwd: out (saved for
                            bwd), bwd: in.  May be NULL for the LAST layer when pool != 0 —
                            the pooled layer is never materialised. */
  float* grad_weight;    /* bwd out [cout, cin]  (NULL: skip weight/bias/BN-param gradients) */
  float* grad_bias;      /* bwd out [cout] (train-mode BN cancels the bias: exact zeros) */
  float* grad_gamma;     /* bwd out [cout] */
  float* grad_beta;      /* bwd out [cout] */
} pcuda_mlp_layer_t;

enum { PCUDA_MLP_FP32 = 0, PCUDA_MLP_BF16 = 1 };

/* x: [B, C0, N] fp32 addressed as x[b*sxb + c*sxc + n*sxn] (element strides) — the reference
 *    passes a transposed view (train_mscmrseg.py:232), so strides are explicit.
 * out: pool ? [B, C_L] : [B, C_L, N] contiguous.
 * pool_arg: [B, C_L] int32, index n of the selected point (first on ties); required iff pool.
 * precision: PCUDA_MLP_FP32 (CUDA-core reference path) or PCUDA_MLP_BF16 (tcgen05 tensor-core
 *    path for layers with cin >= 64: bf16 operands, fp32 accumulate).
 * ws: pcuda_pointmlp_ws_bytes(...) bytes, any contents.
 */
size_t pcuda_pointmlp_ws_bytes(int B, int N, int L, const pcuda_mlp_layer_t* layers /*host*/,
                               int pool, int backward);
int pcuda_pointmlp_fwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N,
                       int L, const pcuda_mlp_layer_t* layers /*host*/, int pool, int train,
                       float momentum, float eps, int precision, float* out, int32_t* pool_arg,
                       void* ws, pcuda_stream_t stream);

/* grad_out: pool ? [B, C_L] : [B, C_L, N].  grad_x: [B, C0, N] contiguous or NULL.
 * Uses layers[l].y / save_mean / save_invstd written by the matching forward call. */
int pcuda_pointmlp_bwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N,
                       int L, const pcuda_mlp_layer_t* layers /*host*/, int pool, int train,
                       float eps, int precision, const float* out, const int32_t* pool_arg,
                       const float* grad_out, float* grad_x, void* ws, pcuda_stream_t stream);
/* Same, with the workspace of the matching forward call still intact (fwd_ws, same B / N / layers / pool /
 * precision): the bf16 operand slabs the forward packed are read in place instead of being packed again. */
int pcuda_pointmlp_bwd_reuse(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, int B, int N,
                             int L, const pcuda_mlp_layer_t* layers /*host*/, int pool, int train,
                             float eps, int precision, const float* out, const int32_t* pool_arg,
                             const float* grad_out, float* grad_x, void* ws, const void* fwd_ws,
                             pcuda_stream_t stream);

/* The same two calls with the per-cloud INPUT TRANSFORM of PointNetfeat fused into the first layer's operand load:
 *   a_0[b, k, n] = sum_j x[b, j, n] * in_trans[b][j][k]        (in_trans: [B, C0, C0] contiguous, C0 <= 4)
 * i.e. torch.bmm(x.transpose(2, 1), trans).transpose(2, 1) (networks/PointNetCls.py:140-142) without materialising
 * the transformed cloud.  in_trans == NULL: identical to pcuda_pointmlp_fwd / _bwd_reuse.  The first layer must be a
 * narrow one (cin <= 4, cout <= 128, not the pooled layer).  Backward: grad_x is the gradient w.r.t. the
 * UNtransformed cloud x, grad_trans [B, C0, C0] the gradient w.r.t. in_trans (either may be NULL); fwd_ws may be NULL.
 *
 * sync_bn != NULL (train mode): CROSS-RANK BatchNorm statistics.  Every per-channel sum BatchNorm takes
 * over the points — forward (sum y, sum y^2), backward (sum dz, sum dz*yhat; for the pooled layer also the column means
 * of its input) — is summed over the ranks of the communicator (ncclAllReduce of the raw fp64 sums on `stream`, between
 * the producing and the consuming kernel) and the point count becomes world * B * N, so that R ranks holding B/R clouds
 * each normalise and back-propagate exactly like one process holding all B.  The parameter gradients a rank returns are
 * its LOCAL share: summed over the ranks (the gradient all-reduce) they are the single-process gradient.  Every rank
 * must call with the same shapes. */
struct pcuda_comm;
int pcuda_pointmlp_fwd_xf(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* in_trans, int B, int N,
                          int L, const pcuda_mlp_layer_t* layers /*host*/, int pool, int train,
                          float momentum, float eps, int precision, float* out, int32_t* pool_arg,
                          void* ws, struct pcuda_comm* sync_bn, pcuda_stream_t stream);
int pcuda_pointmlp_bwd_xf(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* in_trans, int B, int N,
                          int L, const pcuda_mlp_layer_t* layers /*host*/, int pool, int train,
                          float eps, int precision, const float* out, const int32_t* pool_arg,
                          const float* grad_out, float* grad_x, float* grad_trans, void* ws, const void* fwd_ws,
                          struct pcuda_comm* sync_bn, pcuda_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Per-cloud feature transform of PointNetfeat (networks/PointNetCls.py:147-151, the 64 x 64 matrix of STNkd):
 *   out[b, k, n] = sum_j trans[b][j][k] * x[b, j, n]           x: [B, K, N] by element strides, K <= 64
 * = torch.bmm(x.transpose(2, 1), trans).transpose(2, 1); out, grad_out, grad_x: [B, K, N] contiguous;
 * trans, grad_trans: [B, K, K] contiguous.  Backward: grad_x = trans . grad_out per point, grad_trans[b] =
 * x[b] grad_out[b]^T (contraction over the N points, deterministic).  grad_x / grad_trans may be NULL.
 * ws: pcuda_point_transform_ws_bytes(B, K, N) bytes (backward with grad_trans only). */
size_t pcuda_point_transform_ws_bytes(int B, int K, int N);
int pcuda_point_transform_fwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* trans, int B, int K,
                              int N, float* out, pcuda_stream_t stream);
int pcuda_point_transform_bwd(const float* x, int64_t sxb, int64_t sxc, int64_t sxn, const float* trans,
                              const float* grad_out, int B, int K, int N, float* grad_x, float* grad_trans, void* ws,
                              pcuda_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FC head: a stack of L layers on [B, C] features (the rows are the clouds of the batch)
 *     y_l = (a_{l-1} W_l^T + b_l) * mask_l ;  z_l = bn ? BatchNorm1d(y_l) : y_l ;
 *     a_l = relu ? max(z_l, 0) : z_l
 * networks/PointNetCls.py:46-62 (STN3d: fc1-bn4-relu, fc2-bn5-relu, fc3, + identity), :89-101
 * (STNkd), :208-213 (classifier: fc1-bn1-relu, fc2-dropout-bn2-relu, fc3).  mask_l is the Dropout
 * mask (already scaled by 1/(1-p); drawn by the caller so the RNG stream stays the framework's),
 * or NULL.  BatchNorm1d semantics as for pcuda_pointmlp_* (statistics over the B rows).
 * One launch per layer and direction: a CTA owns 4 output channels and all B rows, so BatchNorm
 * (forward and backward) needs no cross-CTA reduction; results are deterministic.
 * Limits: cin % 4 == 0, cin <= 4096, B <= 2048.
 */
typedef struct pcuda_fc_layer {
  int32_t cin, cout;
  int32_t bn;            /* BatchNorm1d after the (masked) linear map */
  int32_t relu;
  const float* weight;   /* [cout, cin], 16-byte aligned */
  const float* bias;     /* [cout] or NULL */
  const float* mask;     /* [B, cout] or NULL */
  const float* gamma;    /* [cout] BN weight (bn != 0) */
  const float* beta;     /* [cout] BN bias */
  float* running_mean;   /* [cout] or NULL */
  float* running_var;    /* [cout] or NULL */
  float* save_mean;      /* [cout]  fwd: out, bwd: in (bn != 0) */
  float* save_invstd;    /* [cout]  fwd: out, bwd: in */
  float* y;              /* [B, cout] BatchNorm input; fwd: out, bwd: in (bn != 0, else may be NULL) */
  float* a;              /* [B, cout] layer output;    fwd: out, bwd: in */
  float* grad_weight;    /* bwd out [cout, cin] (NULL: skip every parameter gradient of the layer) */
  float* grad_bias;      /* bwd out [cout] or NULL */
  float* grad_gamma;     /* bwd out [cout] or NULL */
  float* grad_beta;      /* bwd out [cout] or NULL */
} pcuda_fc_layer_t;

/* x: [B, layers[0].cin] contiguous, 16-byte aligned.  The output is layers[L-1].a.
 * add_identity_k > 0: adds the flattened k x k identity to the last layer's output (the STN heads;
 * requires layers[L-1].cout == k*k). */
int pcuda_fcstack_fwd(const float* x, int B, int L, const pcuda_fc_layer_t* layers /*host*/, int train,
                      float momentum, float eps, int add_identity_k, pcuda_stream_t stream);
/* grad_out: [B, layers[L-1].cout]; grad_x: [B, layers[0].cin] or NULL.
 * ws: pcuda_fcstack_ws_bytes(B, L, layers, 1) bytes, any contents. */
size_t pcuda_fcstack_ws_bytes(int B, int L, const pcuda_fc_layer_t* layers /*host*/, int backward);
int pcuda_fcstack_bwd(const float* x, int B, int L, const pcuda_fc_layer_t* layers /*host*/, int train,
                      const float* grad_out, float* grad_x, void* ws, pcuda_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Deferred BatchNorm running-statistics update.  pcuda_pointmlp_fwd / pcuda_fcstack_fwd skip the
 * running_mean / running_var update when those pointers are NULL; a caller that runs several forward
 * passes of the same network concurrently (independent streams) applies the updates afterwards, in
 * the order the reference would have executed the passes:
 *     running_mean <- (1-momentum) running_mean + momentum * save_mean
 *     running_var  <- (1-momentum) running_var  + momentum * var * count/(count-1),
 *     var = 1/save_invstd^2 - eps
 * items[0..n) are applied in order by ONE launch (n <= 64); several items may name the same buffers.
 */
typedef struct pcuda_bn_update {
  float* running_mean;      /* [C] */
  float* running_var;       /* [C] */
  const float* save_mean;   /* [C] from the forward call */
  const float* save_invstd; /* [C] */
  int32_t C;
  int32_t reserved;
  double count;             /* values per channel the statistics were taken over (B*N, or B) */
  float momentum, eps;
} pcuda_bn_update_t;
int pcuda_bn_running_update(int n, const pcuda_bn_update_t* items /*host*/, pcuda_stream_t stream);

/* ---- D4 parameter gradients and optimiser step ---------------------------------------------------
 * Replaces the implicit .grad accumulation of the two discriminator backward passes
 * (train_mscmrseg.py:288,319) and optim_dis4.step() = torch.optim.SGD(momentum, weight_decay)
 * (train_mscmrseg.py:329-330,:450-455) by one launch each (per 48 tensors).  `slots` is a HOST array; the
 * pointers it holds are device pointers and travel as kernel parameters (CUDA-graph capturable).
 *   pcuda_grad_sum_pack:      flat[offset_i + j] = scale * (grad_a_i[j] + grad_b_i[j])   (grad_b may be NULL;
 *                             scale = 1/world folds the mean of the gradient all-reduce)
 *   pcuda_sgd_momentum_step:  g = flat_grad[offset_i + j] + weight_decay * param_i[j];
 *                             buf = momentum * flat_momentum[offset_i + j] + g;  param_i[j] -= lr * buf
 *                             (torch.optim.SGD, dampening 0, nesterov off; a zero-initialised buffer reproduces
 *                             torch's first step).  lr is read from device memory (*lr_dev).
 *   pcuda_sgd_momentum_sum_step: both in one launch for a single process (nothing to exchange in between): flat_grad
 *                             receives scale * (grad_a + grad_b) as above and the update uses it; same roundings.
 */
typedef struct pcuda_param_slot {
  const float* grad_a;   /* [numel] (pack only) */
  const float* grad_b;   /* [numel] or NULL (pack only) */
  float* param;          /* [numel] (sgd only) */
  int64_t offset;        /* element offset of this tensor inside the flat buffers */
  int64_t numel;
} pcuda_param_slot_t;
int pcuda_grad_sum_pack(const pcuda_param_slot_t* slots /*host*/, int n, float scale, float* flat, pcuda_stream_t stream);
int pcuda_sgd_momentum_step(const pcuda_param_slot_t* slots /*host*/, int n, const float* flat_grad, float* flat_momentum,
                            const float* lr_dev, float momentum, float weight_decay, pcuda_stream_t stream);
int pcuda_sgd_momentum_sum_step(const pcuda_param_slot_t* slots /*host*/, int n, float scale, float* flat_grad, float* flat_momentum,
                                const float* lr_dev, float momentum, float weight_decay, pcuda_stream_t stream);

/* ---- discriminator bookkeeping ---------------------------------------------------------------------
 * Replaces F.binary_cross_entropy_with_logits(D_out4, full_like(D_out4, label)) (reduction 'mean',
 * train_mscmrseg.py:233,286,316), its backward into the logits, and the accuracy bookkeeping
 * (train_mscmrseg.py:290-296,:320-322) by one launch:
 *   loss = weight * mean_i[(1 - target) x_i - log_sigmoid(x_i)],  grad_logit[i] = weight * (sigmoid(x_i) - target) / n
 *   accuracy = mean_i[(sigmoid(x_i) >= 0.5) == (target >= 0.5)]      (grad_logit / accuracy may be NULL)
 */
int pcuda_bce_logits(const float* logit, int n, float target, float weight, float* loss, float* grad_logit,
                     float* accuracy, pcuda_stream_t stream);

/* ---- farthest-point sampling ------------------------------------------------------
 * utils/npy2point.py:11-18 `graipher(pts, K, dim)`: the K points of a cloud chosen greedily by largest squared distance
 * to the set chosen so far, starting from point starts[b] (the reference draws it with np.random.randint), float64
 * arithmetic and np.argmax tie-breaking (first index), so the selection is bit-identical to numpy's.
 * pts: [B, Vmax, dim] float64, counts[b] (<= Vmax, NULL: all Vmax) valid points of cloud b, dim <= 3, Vmax <= 25600.
 * out_pts: [B, K, dim] float64, out_idx: [B, K] int32 (the chosen rows; an empty cloud gives zero rows / -1). */
int pcuda_fps(const double* pts, const int32_t* counts, const int32_t* starts, int B, int Vmax, int K, int dim,
              double* out_pts, int32_t* out_idx, pcuda_stream_t stream);

/* ---- multi-GPU exchange ----------------------------------------------------------
 * One process per GPU; rank r holds samples [r*B/R, (r+1)*B/R).  The only data-path exchange of the step is the sum
 * of D4's parameter gradients in front of the SGD step.  A communicator is created once per process on the CURRENT
 * device (collective call: every rank calls pcuda_comm_init with the same 128-byte id, which rank 0 obtains from
 * pcuda_comm_unique_id and distributes by any means, e.g. torch.distributed.broadcast_object_list).
 *   pcuda_comm_allreduce      in-place sum of `count` floats over the ranks: ncclAllReduce on `stream` (libnccl is
 *                             bound with dlopen on first use; error codes 1000 + ncclResult_t).
 *   pcuda_comm_allreduce_p2p  the same sum over NVLink peer memory, for buffers that live inside the communicator's
 *                             symmetric region (pcuda_comm_p2p_buffers): reads `count` floats of every rank's `in`
 *                             buffer, leaves the sum in every rank's `out` buffer — one kernel, two-shot (each rank
 *                             reduces its 1/R slice in rank order and stores it to all peers), bit-identical on all
 *                             ranks.  Needs p2p_floats >= count at init, <= 8 ranks on one node with peer access
 *                             (pcuda_comm_info tells whether it is available).  Every rank must issue the same
 *                             sequence of calls.  The producer of `in` and the consumer of `out` are ordinary
 *                             kernels before / after it on `stream`.
 * Both are asynchronous launches on `stream` and legal under CUDA-graph capture.  pcuda_comm_status returns
 * non-zero if a peer-memory wait ever gave up (a peer never arrived within 30 s).
 */
typedef struct pcuda_comm pcuda_comm_t;
int pcuda_comm_unique_id(void* id_out /*host, >= 128 bytes*/, int bytes);
int pcuda_comm_init(const void* unique_id /*host, 128 bytes*/, int rank, int world, size_t p2p_floats, pcuda_comm_t** out);
int pcuda_comm_allreduce(pcuda_comm_t* comm, float* buf, int64_t count, pcuda_stream_t stream);
/* in-place sum of `count` doubles (BatchNorm statistics in the cross-rank mode): up to 2048 doubles with peer memory
 * available go through a one-shot NVLink mailbox kernel (every rank stores its values into every peer's mailbox, one flag
 * round trip, sum in rank order: bit-identical on all ranks; one channel per caller stream), anything else through
 * ncclAllReduce.  pcuda_comm_allgather: ncclAllGather of count_per_rank floats. */
int pcuda_comm_allreduce_f64(pcuda_comm_t* comm, double* buf, int64_t count, pcuda_stream_t stream);
int pcuda_comm_allgather(pcuda_comm_t* comm, const float* send, float* recv /*[world * count_per_rank]*/, int64_t count_per_rank,
                         pcuda_stream_t stream);
int pcuda_comm_p2p_buffers(pcuda_comm_t* comm, float** in, float** out, int64_t* capacity_floats);
int pcuda_comm_allreduce_p2p(pcuda_comm_t* comm, int64_t count, pcuda_stream_t stream);
int pcuda_comm_status(pcuda_comm_t* comm);
int pcuda_comm_info(pcuda_comm_t* comm, int* rank, int* world, int* p2p_available, int* nccl_version);
int pcuda_comm_destroy(pcuda_comm_t* comm);

#ifdef __cplusplus
}
#endif
#endif /* PCUDA_H_ */
