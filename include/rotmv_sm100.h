/* librotmv_sm100.so -- C ABI of the B200-native Rot-MVGaze multi-view hot path.
 *
 * The reference (ut-vision/Rot-MVGaze) is pure PyTorch and has no FFI of its own: the drop-in
 * boundary its callers see is the nn.Module contract of models/rot_mv.py:102-269
 * (`FeatRotationSymm`), mirrored by rot-mvgaze_b200/rotmv_b200/module.py. This header is the
 * C boundary underneath that module: each entry point replaces the ATen/cuDNN/cuBLAS dispatches
 * the cited reference lines perform today. INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates or frees
 *     device memory and never synchronises the host (every call is CUDA-graph capturable);
 *   - activations are NHWC (channels innermost); "strides" are in ELEMENTS;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = OK, <0 = invalid argument / unsupported shape, >0 = cudaError_t;
 *     rmv_last_error() returns a thread-local description of the last failure;
 *   - dtype codes: RMV_DTYPE_F32 = 0 (fp32 storage, FFMA kernels), RMV_DTYPE_BF16 = 1
 *     (bf16 storage, tcgen05 tensor-core kernels with fp32 accumulation in TMEM).
 */
#ifndef ROTMV_SM100_H_
#define ROTMV_SM100_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RMV_VERSION 100
#define RMV_DTYPE_F32 0
#define RMV_DTYPE_BF16 1

#define RMV_ENGINE_AUTO 0 /* bf16 storage -> tcgen05 when the shape allows, else FFMA */
#define RMV_ENGINE_SIMT 1 /* hand-written FFMA kernels (fp32 parity mode; on-device reference) */
#define RMV_ENGINE_TC 2   /* tcgen05/TMEM/TMA implicit GEMM; error if the shape is unsupported */

int rmv_version(void);
const char* rmv_last_error(void);
/* 0 when `device` is compute capability 10.x (B200); negative otherwise. */
int rmv_device_check(int device);
/* Kernel-variant switches, normally read once from the environment (RMV_<KEY>=digit): "PDL"
 * (0/1/2 programmatic dependent launch level), "HALO" (halo-patch 3x3 kernel), "CTA2" (cta_group::2
 * pair kernel: 0 off, 1 where it applies, 2 forced for every c_out % 256 == 0 layer),
 * "WGRAD_ROWS" (tap-row weight-gradient kernel). This call overrides them at run time. */
int rmv_set_tuning(const char* key, int value);

/* BatchNorm(train) finalize parameters. Passed (non-NULL) to a statistics / reduction call, the LAST
 * thread block of that launch turns the fp64 sums into the coefficients of the apply pass, so no
 * separate rmv_bn_finalize / rmv_bn_bwd_finalize launch is needed. `ticket`: one zero-initialised
 * unsigned int in device memory (reset by the call). Forward: mean, invstd, a = gamma*invstd,
 * b = beta - mean*a per (view, channel) [views][c]; running statistics updated in view order
 * (momentum, unbiased variance), num_batches += views. Backward: dgamma, dbeta [c] and
 * k0, k1, k2 [views][c] of dz = k0*dy + k1*z + k2. Unused fields may be NULL. */
typedef struct rmv_bn_params {
  unsigned int* ticket;
  const float* gamma;
  const float* beta;
  float* running_mean;
  float* running_var;
  long long* num_batches;
  float* mean;
  float* invstd;
  float* a;
  float* b;
  float* dgamma;
  float* dbeta;
  float* k0;
  float* k1;
  float* k2;
  float eps, momentum;
} rmv_bn_params;

/* ----------------------------------------------------------------------------------------------
 * Convolution / linear layer with fused epilogue:
 *     y = act( scale[k] * conv(x, w)[..., k] + shift[k] + residual )
 * Replaces nn.Conv2d + nn.BatchNorm2d(eval) + residual add + nn.ReLU of
 * models/resnet.py:128-148,261-266 and nn.Linear(+ReLU) of models/backbones/blocks.py:41-60.
 * A Linear layer is the 1x1 case with n_img=1, in_h=1, in_w=rows, x_sw=lda, y_sw=ldc.
 * -------------------------------------------------------------------------------------------- */
typedef struct rmv_conv_args {
  int x_dtype;            /* RMV_DTYPE_*; weights have the same dtype */
  int y_dtype;            /* RMV_DTYPE_* of y */
  int engine;             /* RMV_ENGINE_* */
  int block_n;            /* 0 = auto; tcgen05 N tile (64/128/256) */
  const void* x;          /* input, element (n,h,w,c) at x[n*x_sn + h*x_sh + w*x_sw + c*x_sc] */
  long long x_sn, x_sh, x_sw, x_sc;
  int n_img, in_h, in_w, c_in;
  const void* w;          /* filters [c_out][kh][kw][c_in], contiguous */
  int c_out, kh, kw, stride, pad;
  void* y;                /* output, element (n,oh,ow,k) at y[n*y_sn + oh*y_sh + ow*y_sw + k] */
  long long y_sn, y_sh, y_sw;
  int out_h, out_w;
  const float* scale;     /* [c_out] or NULL (=1) */
  const float* shift;     /* [c_out] or NULL (=0): folded BN shift or Linear bias */
  const void* residual;   /* same dtype as y, or NULL */
  long long r_sn, r_sh, r_sw;
  int relu;
  /* Training (tcgen05 engine, bf16 output, no residual): fuse the BatchNorm batch statistics of y
   * into the epilogue -- stat_acc[stat_views][c_out][2] (fp64, caller-zeroed or accumulating) +=
   * per-(view, channel) sum and sum of squares of the bf16 values written to y; image n belongs
   * to view n % stat_views. NULL = off. Only stat_views == 2 is implemented. */
  double* stat_acc;
  int stat_views;
  const void* stat_finalize; /* const rmv_bn_params* or NULL: finalize inside this launch (forward fields) */
  /* Training, RECOMPUTED BatchNorm of the expanding 1x1 convolutions (tcgen05 engine, bf16, two
   * views: image n belongs to view n & 1). The conv output z is never written to HBM; the
   * statistics come from rmv_conv_bn_stats / rmv_conv_bn_bwd_reduce, the apply passes are epilogue
   * modes of this call with per-(view, channel) coefficient tables [2][c_out] (scale/shift unset):
   *   bn_mode 1 (forward):  y = relu?( bn_a*z + bn_b + residual );  bn_bits (optional) receives the
   *                         sign mask of the pre-ReLU value
   *   bn_mode 2 (backward): y = dz = bn_a*dy + bn_b*z + bn_c, with dy passed as `residual`
   * Packed masks hold 1 bit per element of a tensor with y's geometry: element offset
   * n*y_sn + oh*y_sh + ow*y_sw + k (+ mask_off) -> bit (off % 8) of byte off / 8.
   * mask_bits (any tcgen05 launch with bf16 output): the value about to be stored is zeroed where its
   * mask bit is 0 -- the data gradient of a block input arrives already multiplied by the ReLU
   * derivative of that tensor (models/resnet.py:146).
   *   bn_mode 4 (data gradients of the mid layers, rmv_conv2d_dgrad): y = dy * mask is stored AND reduced
   *                         for the BatchNorm backward of the layer that produced the tensor:
   *                         stat_acc += (sum dy, sum dy*xhat) per (view, channel), xhat = (z - mean)*invstd
   *                         with z passed as `residual` (NOT added), bn_a = mean, bn_b = invstd [2][c_out];
   *                         stat_finalize (backward fields) finalizes k0/k1/k2/dgamma/dbeta in the launch. */
  int bn_mode;
  const float* bn_a;
  const float* bn_b;
  const float* bn_c;
  void* bn_bits;
  const void* mask_bits;
  long long mask_off;
  /* Optional scratch for split-K (tcgen05 engine; pointwise / Linear GEMMs whose tile count is far
   * below the SM count -- the lifter / fuser / head layers of models/rot_mv.py:35-50,91-98,179-184
   * at M = B*V rows): S CTAs share one output tile along K, write fp32 partial tiles here and a
   * second kernel adds them in a fixed order (bit-reproducible). NULL or too small = no split.
   * rmv_splitk_workspace_bytes() is always enough. 16-byte aligned device memory. Launches that
   * share one workspace must be ordered by their stream (one workspace per concurrently used stream). */
  void* workspace;
  size_t workspace_bytes;
} rmv_conv_args;

int rmv_conv2d_fwd(const rmv_conv_args* args, void* stream);

/* Entry points that are NOT on the product path (kept for the tests and as building blocks; the
 * engines call the fused forms): rmv_bn_stats + rmv_bn_finalize and rmv_bn_bwd_reduce +
 * rmv_bn_bwd_finalize (the product uses rmv_bn_stats_finalize / rmv_bn_bwd_reduce_finalize or the
 * finalize-in-launch forms via rmv_bn_params), rmv_maxpool3x3s2_bwd (the product uses the index form
 * rmv_maxpool3x3s2_fwd_idx / _bwd_idx), rmv_stem_im2col and rmv_nchw_to_nhwc (layout helpers of the
 * tests; the stem kernel builds its im2col rows in shared memory). rmv_dilate2 serves the fp32 (FFMA)
 * parity engine's stride-2 data gradient only. */

/* Data gradient of y = conv(x, w, stride, pad) on the tcgen05 engine (the backward of
 * models/resnet.py:31-47 that autograd/cuDNN computes for trainer.py:142):
 *     dx[n,ih,iw,c] (+ residual) = sum_{r,s,k} dy[n,oh,ow,k] * w[k,c,r,s],   ih = oh*stride - pad + r
 * The argument struct is read "transposed": x/in_h/in_w/c_in describe dy (c_in = the forward
 * c_out), y/out_h/out_w/c_out describe dx (c_out = the forward c_in), kh/kw/stride/pad are the
 * FORWARD conv's, and w holds the reversed, transposed filters wt[c][kh-1-r][kw-1-s][k] = w[k][c][r][s]
 * (bf16). scale/shift/relu must be unset; residual (bf16, dx geometry) is added. stride 2 runs as
 * four parity-class convolutions over the undilated dy (1x1 filters: dx is zero-filled first and
 * must be dense). */
int rmv_conv2d_dgrad(const rmv_conv_args* args, void* stream);

/* BatchNorm(train) reductions over the RECOMPUTED output z = conv1x1(x, w) (stride 1 or 2, bf16,
 * c_in % 64 == 0, c_out % 128 == 0; `args` describes the forward conv, y is not touched). The GEMM
 * runs transposed on the tensor cores (channels = TMEM lanes), the accumulator is reduced on the
 * fly and z never reaches HBM. acc = fp64 [2][c_out][2] (rmv_bn_workspace_bytes), accumulated into:
 *   rmv_conv_bn_stats:       (sum z, sum z^2) per (view, channel)            -> rmv_bn_finalize
 *   rmv_conv_bn_bwd_reduce:  (sum dy, sum dy*xhat), xhat = (z - mean)*invstd -> rmv_bn_bwd_finalize
 * or, with `finalize` != NULL, finalized by the last block of the launch itself (acc is reset).
 * dy (bf16) has the geometry args->y_s* of the conv output and is expected to be already masked by
 * the ReLU that follows the BatchNorm. Replaces the statistics passes of nn.BatchNorm2d behind
 * models/resnet.py:122-123,139-146,229-230 (autograd at trainer.py:142). */
int rmv_conv_bn_stats(const rmv_conv_args* args, double* acc, const rmv_bn_params* finalize,
                      void* stream);
int rmv_conv_bn_bwd_reduce(const rmv_conv_args* args, const void* dy, const float* mean,
                           const float* invstd, double* acc, const rmv_bn_params* finalize,
                           void* stream);

/* Stem im2col: x fp32 NCHW [n,3,224,224]-like -> A[n*out_h*out_w, k_pad] (bf16 or fp32), row =
 * (kh,kw,c)-ordered 7x7x3 patch (stride 2, pad 3) zero-padded to k_pad columns. Feeds the stem
 * conv (models/resnet.py:184-186,262) to the tensor-core GEMM. */
int rmv_stem_im2col(const float* x, void* a, int n_img, int c_in, int in_h, int in_w, int kh,
                    int kw, int stride, int pad, int out_h, int out_w, int k_pad, int a_dtype,
                    void* stream);

/* Fused stem on the tensor cores: fp32 NCHW image [n,3,in_h,in_w] -> Conv2d(3,64,k7,s2,p3) ->
 * scale/shift (folded BatchNorm) -> ReLU -> bf16 NHWC [n,out_h,out_w,64]; the im2col rows are
 * built in shared memory and never touch HBM (models/resnet.py:184-188,262-264).
 * w_packed is produced once by rmv_stem_pack_weights from the fp32 [64,3,7,7] filters
 * (bf16 [64,192], k = c*56 + kh*8 + kw). */
int rmv_stem_pack_weights(const float* w_oihw, void* w_packed, void* stream);
int rmv_stem_conv_fwd(const float* x_nchw, const void* w_packed, const float* scale,
                      const float* shift, void* y_nhwc, int n_img, int in_h, int in_w, int relu,
                      void* stream);

/* Same, from uint8 HWC images [n,in_h,in_w,3] (what an image decoder delivers): the loader applies
 * ToTensor + Normalize(mean3, std3) (main.py:38-56; HOST pointers to three floats each) on the
 * fly, so the host->HBM copy and the HBM read shrink 4x. Needs 3*in_w % 16 == 0. */
int rmv_stem_conv_fwd_u8(const unsigned char* x_nhwc_u8, const float* mean3, const float* std3,
                         const void* w_packed, const float* scale, const float* shift, void* y_nhwc,
                         int n_img, int in_h, int in_w, int relu, void* stream);

/* Stem weight gradient on the tensor cores (training): dw_oihw[64,3,7,7] (fp32, =) from the fp32
 * NCHW images and dz (bf16 NHWC [n,out_h,out_w,64], gradient of the stem conv output); the im2col
 * rows are rebuilt in shared memory exactly as in the forward. scratch: fp32 [192*64] workspace. */
int rmv_stem_wgrad(const float* x_nchw, const void* dz_nhwc, float* scratch, float* dw_oihw,
                   int n_img, int in_h, int in_w, void* stream);

/* Workspace queries (boundary B2: the library never allocates device memory -- the caller owns
 * every buffer and sizes the scratch regions with these; all are pure host functions).
 *   rmv_stem_wgrad_workspace_bytes        `scratch` of rmv_stem_wgrad
 *   rmv_bn_workspace_bytes                the fp64 accumulator `acc` shared by the rmv_bn_* calls and
 *                                         the conv epilogue statistics (stat_acc), for layers of up
 *                                         to `max_channels` channels and `views` views; the `ticket`
 *                                         counter those calls take is one more zero-initialised int
 *   rmv_conv2d_wgrad_tc_workspace_bytes   the fp32 [c_out][kh][kw][c_in] accumulation buffer
 *                                         `dw_krsc` of rmv_conv2d_wgrad_tc (zeroed by the caller)
 *   rmv_splitk_workspace_bytes            upper bound of rmv_conv_args.workspace for any launch */
size_t rmv_stem_wgrad_workspace_bytes(void);
size_t rmv_splitk_workspace_bytes(void);
size_t rmv_bn_workspace_bytes(int max_channels, int views);
size_t rmv_conv2d_wgrad_tc_workspace_bytes(const rmv_conv_args* args);

/* fp32 NCHW -> NHWC (fp32 or bf16) layout change (the reference keeps NCHW, trainer.py:100-106). */
int rmv_nchw_to_nhwc(const float* x, void* y, int n_img, int c, int h, int w, int y_dtype,
                     void* stream);

/* MaxPool2d(kernel 3, stride 2, pad 1), NHWC. models/resnet.py:189,265. */
int rmv_maxpool3x3s2_fwd(const void* x, void* y, int n_img, int in_h, int in_w, int c, int dtype,
                         void* stream);

/* AdaptiveAvgPool2d(1) + Flatten: x NHWC [n, hw, c] -> y0[n, c] (row stride ld0) and, if y1 !=
 * NULL, the same values to y1 (row stride ld1). models/resnet.py:272-273, models/rot_mv.py:124-128. */
int rmv_avgpool_fwd(const void* x, int n_img, int hw, int c, int dtype, void* y0, long long ld0,
                    void* y1, long long ld1, void* stream);

/* Rotation-constrained cross-view gather (models/rot_mv.py:193-194,234,238; SURVEY D1 for V>2):
 *   dst[(b*V+v), r*nvec + k] = 1/(V-1) * sum_{u != v} sum_c rot[b,v,u,r,c] * feat[(b*V+u), c*nvec + k]
 * feat rows have stride ld_feat, dst rows ld_dst (both in elements); rot is fp32 [B,V,V,3,3] with
 * rot[b,i,j] = R_i R_j^T. apply_rot bit 0: apply the rotation (0 = ignore_rotmat=True, plain
 * mean of the partner features); bit 1: BACKWARD of the gather -- dst row (b,v) receives
 * 1/(V-1) * sum_{u != v} rot[b,u,v]^T applied to feat row (b,u) (feat = gradient w.r.t. dst). */
int rmv_rotate_gather_fwd(const void* feat, long long ld_feat, const float* rot, void* dst,
                          long long ld_dst, int batch, int views, int nvec, int dtype,
                          int apply_rot, void* stream);

/* Gaze head tail + loss (models/rot_mv.py:249-254 last Linear; utils/math.py:52-60;
 * losses/gaze_loss.py:42-52; losses/stereo_loss.py:46-54,65-84):
 *   pred[m, 0:2] = hidden[m, :] . w2[0:2, :]^T + b2           (hidden: [rows, hid] fp32/bf16)
 *   if gt != NULL: loss_out[0] += loss_scale * sum_m d_m * angular_deg(pred[m], gt[m])  (fp32 atomics)
 *                  with d_m = 1 for view 0 rows (m % views == 0) and aux_decay otherwise
 *                  (StereoL1Loss.reference_decay).
 * The caller folds StereoL1Loss.rel_weight, the IterationLoss decay of this iteration and the
 * 1/batch of the mean into loss_scale. pred is fp32 [rows, 2]; gt fp32 [rows, 2]; w2/b2 fp32. */
int rmv_head_loss_fwd(const void* hidden, long long ld_hidden, int hid_dtype, const float* w2,
                      const float* b2, int rows, int hid, float* pred, const float* gt,
                      float loss_scale, int views, float aux_decay, float* loss_out,
                      void* stream);

/* Mean angular error in degrees between pitch-yaw predictions and labels, clamped cosine
 * (utils/math.py:96-137; the on-device replacement for the per-step D2H at trainer.py:128).
 * Adds sum of per-row errors to err_sum[0] and rows to err_sum[1]. */
int rmv_angular_error_accum(const float* pred, long long ld_pred, const float* gt, long long ld_gt,
                            int rows, float* err_sum, void* stream);

/* head pose (pitch,yaw) [B,V,2] -> rotations [B,V,V,3,3], rot[b,i,j] = R_i R_j^T with
 * R = R_y(yaw) R_x(-pitch) (utils/math.py:188-219, trainer.py:110-111, models/rot_mv.py:193-194). */
int rmv_pose_to_rotations(const float* head_pose, float* rotations, int batch, int views,
                          void* stream);

/* per-view rotations rot[B,V,3,3] -> rotations[B,V,V,3,3], [b,i,j] = rot[b,i] rot[b,j]^T
 * (models/rot_mv.py:193-194: rot_10 = rot_0 rot_1^T, rot_01 = rot_1 rot_0^T). */
int rmv_relative_rotations(const float* rot, float* rotations, int batch, int views,
                           void* stream);

/* ==============================================================================================
 * Training step (trainer.py:119-123,141-143): batch-statistic BatchNorm, backward kernels, Adam.
 * Image n of a [B*V, ...] activation belongs to view n % views; BatchNorm statistics are PER VIEW
 * because the reference calls the trunk once per view (models/rot_mv.py:196-197).
 * ============================================================================================ */

/* acc[views][c][2] (fp64, zero on entry) += per-(view,channel) sum and sum of squares of z[n,pix,c]. */
int rmv_bn_stats(const void* z, int dtype, int n_img, int pix, int c, int views, double* acc,
                 void* stream);
/* mean/invstd[views][c]; fused affine a = gamma*invstd, b = beta - mean*a; running_mean/var
 * updated once per view in view order with momentum (unbiased variance), num_batches += views
 * (nn.BatchNorm2d train semantics, models/resnet.py:187 etc.); resets acc to zero. */
int rmv_bn_finalize(double* acc, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches, float* mean, float* invstd,
                    float* a, float* b, int c, int views, long long count_per_view, float eps,
                    float momentum, void* stream);
/* rmv_bn_stats + rmv_bn_finalize in ONE launch: the last block to finish (ticket counter, zero on
 * entry and on exit) turns the sums into the coefficients; count_per_view = n_img/views * pix. */
int rmv_bn_stats_finalize(const void* z, int dtype, int n_img, int pix, int c, int views,
                          double* acc, unsigned int* ticket, const float* gamma, const float* beta,
                          float* running_mean, float* running_var, long long* num_batches,
                          float* mean, float* invstd, float* a, float* b, float eps, float momentum,
                          void* stream);
/* y = relu?(a[v,c]*z + b[v,c] + residual). relu_bits (may be NULL): packed ReLU mask for the
 * backward pass, [n_img*pix*c/8] bytes, bit e of byte i = element 8*i+e was > 0 (16x less
 * backward traffic than re-reading y). */
int rmv_bn_apply(const void* z, const float* a, const float* b, const void* residual, void* y,
                 unsigned char* relu_bits, int dtype, int n_img, int pix, int c, int views,
                 int relu, void* stream);
/* acc[views][c][2] += { sum dyr, sum dyr*xhat }, dyr = dy masked by the ReLU: y_mask is NULL (no
 * ReLU), the ReLU output tensor (mask_is_bits = 0: y > 0) or rmv_bn_apply's relu_bits (= 1). */
int rmv_bn_bwd_reduce(const void* z, const void* dy, const void* y_mask, int mask_is_bits,
                      const float* mean, const float* invstd, int dtype, int n_img, int pix, int c,
                      int views, double* acc, void* stream);
/* dgamma/dbeta[c] and the per-(view,channel) coefficients of dz = k0*dyr + k1*z + k2; resets acc. */
int rmv_bn_bwd_finalize(double* acc, const float* gamma, const float* mean, const float* invstd,
                        float* dgamma, float* dbeta, float* k0, float* k1, float* k2, int c,
                        int views, long long count_per_view, void* stream);
/* rmv_bn_bwd_reduce + rmv_bn_bwd_finalize in ONE launch (same ticket protocol). */
int rmv_bn_bwd_reduce_finalize(const void* z, const void* dy, const void* y_mask, int mask_is_bits,
                               const float* mean, const float* invstd, int dtype, int n_img,
                               int pix, int c, int views, double* acc, unsigned int* ticket,
                               const float* gamma, float* dgamma, float* dbeta, float* k0,
                               float* k1, float* k2, void* stream);
/* dz = k0*dyr + k1*z + k2; optionally also writes dyr (the gradient of the residual branch). */
int rmv_bn_bwd_apply(const void* z, const void* dy, const void* y_mask, int mask_is_bits,
                     const float* k0, const float* k1, const float* k2, void* dz, void* dyr_out,
                     int dtype, int n_img, int pix, int c, int views, void* stream);
/* dst = (mask > 0 ? src : 0) + add, 2-D row-strided (mask/add may be NULL): ReLU backward,
 * gradient accumulation of the concat-free fusion buffers. */
int rmv_relu_bwd(const void* src, long long ld_src, const void* mask, long long ld_mask,
                 const void* add, long long ld_add, void* dst, long long ld_dst, int rows, int cols,
                 int dtype, void* stream);
/* out[col] += sum_rows x[row,col] (fp32): Linear bias gradients. */
int rmv_colsum(const void* x, long long ld, int rows, int cols, int dtype, float* out,
               void* stream);
/* dst[i0,i1,i2,i3] (contiguous, fp32 or bf16) = src[i0*s0 + f1(i1)*s1 + f2(i2)*s2 + i3*s3] with
 * optional flips of dims 1,2: filter layout changes (OIHW fp32 master -> KRSC, and the
 * flipped/transposed filters of the data-gradient convolution). */
int rmv_permute_cast(const float* src, void* dst, int d0, int d1, int d2, int d3, long long s0,
                     long long s1, long long s2, long long s3, int flip1, int flip2, int dst_dtype,
                     void* stream);
/* Batched rmv_permute_cast: `jobs_dev` is a DEVICE array of n_jobs jobs sorted by first_block;
 * job i owns blocks [first_block_i, first_block_{i+1}) (total_blocks in all). One launch re-derives
 * every engine-layout filter tensor of a training step. `kind` picks the access pattern (all give
 * the result the generic form describes):
 *   0 generic 4-D permute (rmv_permute_cast semantics), ceil(total/1024) blocks;
 *   1 contiguous cast, ceil(total/1024) blocks;
 *   2 dims (C,R,S,K): src [K][C][R*S] -> dst [C][R*S (reversed if flip1)][K], 64x64 tiles through
 *     shared memory, ceil(C*R*S/64)*ceil(K/64) blocks;
 *   3 dims (K,R,S,C): src [K][C][R*S] -> dst [K][R*S][C], one block per k. */
typedef struct rmv_permute_job {
  const float* src;
  void* dst;
  int d0, d1, d2, d3;
  long long s0, s1, s2, s3;
  int flip1, flip2, dst_dtype;
  unsigned first_block;
  int kind;
} rmv_permute_job;
int rmv_permute_cast_batch(const rmv_permute_job* jobs_dev, int n_jobs, unsigned total_blocks,
                           void* stream);
/* dst[n,2h,2w,:] = src[n,h,w,:], zeros elsewhere (dst [n,2H,2W,c]): stride-2 data gradient. */
int rmv_dilate2(const void* src, void* dst, int n_img, int h, int w, int c, int dtype,
                void* stream);
int rmv_maxpool3x3s2_bwd(const void* x, const void* dy, void* dx, int n_img, int in_h, int in_w,
                         int c, int dtype, void* stream);
/* Training variant of the max-pool: forward also records the winning window position (uint8,
 * 0..8 = r*3+s, first maximum in scan order like ATen) per output element; the backward routes
 * each gradient through that index. idx has the shape of y. */
int rmv_maxpool3x3s2_fwd_idx(const void* x, void* y, void* idx, int n_img, int in_h, int in_w,
                             int c, int dtype, void* stream);
int rmv_maxpool3x3s2_bwd_idx(const void* idx, const void* dy, void* dx, int n_img, int in_h,
                             int in_w, int c, int dtype, void* stream);
int rmv_avgpool_bwd(const void* dfeat, long long ld, void* dx, int n_img, int hw, int c, int dtype,
                    void* stream);
/* dst[i] = (bit i of the packed mask `bits`) ? src[i] : 0 over n elements (n % 8 == 0; dst may alias
 * src): a gradient times the derivative of the ReLU whose sign mask rmv_bn_apply / bn_mode 1 packed
 * (models/resnet.py:146 under autograd). */
int rmv_mask_bits(const void* src, const void* bits, void* dst, long long n, int dtype, void* stream);
/* Analytic gradient of the weighted angular loss w.r.t. pred (zero where the cosine saturates,
 * like F.hardtanh) + backward of the head's last Linear(512,2) and the ReLU before it:
 * dhidden = (hidden>0) * dpred w2; dw2 += dpred^T hidden; db2 += colsum(dpred).
 * gt == NULL: `dpred` [rows,2] is an INPUT -- d(loss)/d(pred) from the caller's own loss objects
 * (losses/stereo_loss.py:65-84 under torch.autograd, trainer.py:141-142); pred/loss_scale/views/
 * aux_decay are ignored and only the Linear(512,2) + ReLU backward runs. */
int rmv_head_loss_bwd(const float* pred, const float* gt, const void* hidden, long long ld_hidden,
                      int hid_dtype, const float* w2, int rows, int hid, float loss_scale,
                      int views, float aux_decay, void* dhidden, long long ld_dhidden,
                      float* dpred, float* dw2, float* db2, void* stream);
/* Filter gradient: dw[k,c,r,s] (fp32, OIHW like the PyTorch parameter) += sum over output pixels
 * of dy[p,k] * x[p shifted by (r,s), c]; `args` describes the forward convolution (x, strides,
 * shapes; args->y_s* are the strides of dy). */
int rmv_conv2d_wgrad(const rmv_conv_args* args, const void* dy, float* dw, void* stream);
/* Same gradient on the tensor cores (bf16 x and dy, c_in % 64 == 0): tcgen05.mma with MN-major
 * operands straight from the NHWC tensors (TMA boxes, tap shift + zero padding by TMA), split over
 * the pixel axis with TMA reduce-add. dw_krsc is fp32 [c_out][kh][kw][c_in] (+=, caller zeroes);
 * rmv_permute_cast brings it to the parameter layout. */
int rmv_conv2d_wgrad_tc(const rmv_conv_args* args, const void* dy, float* dw_krsc, void* stream);
/* Fused Adam over a flat fp32 buffer; hyper = double{lr, beta1, beta2, eps, weight_decay, step} in
 * DEVICE memory (step is incremented by the call). decoupled=0: torch.optim.Adam(weight_decay),
 * i.e. coupled L2 as trainer.py:54; decoupled=1: AdamW. grads are multiplied by grad_scale. */
int rmv_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                  double* hyper, long long n, int decoupled, float grad_scale, void* stream);

/* ---- re-layout kernels of the constructor variants (off the path main.py builds) ---- */
/* dst[i0,i1,i2] = (accumulate ? dst[i0,i1,i2] : 0) + src[i0,i1,i2] * (scale ? scale[i2] : 1) over an
 * n0 x n1 x n2 index space with ELEMENT strides (ss*, ds*) and independent dtypes (RMV_DTYPE_F32 /
 * RMV_DTYPE_BF16; arithmetic in fp32, one rounding on the store). Replaces the tensor copies of
 * ImageRotmatFeatFuser / RotFeatFuser: zero-padded corners of the 3593-wide layers and their
 * transposes (a source contiguous along i1 with a destination contiguous along i2 runs as 32x32
 * shared-memory tiles), the 9 rotation entries appended per row (models/rot_mv.py:53-67,225-231),
 * the cat(...,-1).flatten(-2,-1) interleave [3][2][nvec] of two [3][nvec] features
 * (models/rot_mv.py:80-84,243-248), x / (running_std + eps) of IntensityBatchNorm as `scale`
 * (:32), and the backward scatter / accumulation of all of these. src and dst must not overlap. */
int rmv_strided_copy(const void* src, int src_dtype, long long ss0, long long ss1, long long ss2,
                     void* dst, int dst_dtype, long long ds0, long long ds1, long long ds2, int n0,
                     int n1, int n2, const float* scale, int accumulate, void* stream);
/* Train-mode IntensityBatchNorm statistics of ONE call (models/rot_mv.py:13-32): feat is
 * [rows][3][nvec] with row stride ld (elements); intensity[r][j] = ||feat[r,:,j]||_2 (detached in the
 * reference), std_j = sqrt(max(var_r(intensity) [biased], eps)), running[j] = (1-momentum) *
 * running[j] + momentum * std_j (the buffer the reference calls `running_mean`), scale_out[j] =
 * 1 / (running[j] + eps): the factor this call applies (through rmv_strided_copy's `scale`). */
int rmv_intensity_bn_train(const void* feat, long long ld, int dtype, int rows, int nvec,
                           float* running, float momentum, float eps, float* scale_out,
                           void* stream);
/* Zero `bytes` bytes at dst: gradient accumulators (flat gradient buffer, split-K
 * weight-gradient scratch; the reference's optimizer.zero_grad(), trainer.py:141) and padding. */
int rmv_fill_zero(void* dst, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROTMV_SM100_H_ */
