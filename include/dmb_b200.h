/*
 * dmb_b200.h -- C ABI of the B200-native cost-volume hot path for DenseMatchingBenchmark (dmb).
 *
 * The reference has no C boundary for this path: every function below replaces a piece of
 * PyTorch/ATen code reached from `dmb.modeling.stereo.{cost_processors,disp_predictors}`
 * or the pybind11 module of `dmb.ops.spn` (dmb/ops/spn/src/gaterecurrent2dnoind_cuda.cpp:86-89,
 * whose entry points take torch::Tensor).  Here the boundary is plain C: raw DEVICE pointers,
 * explicit sizes, a dtype code, and the CUDA stream to launch on.  No torch types.
 *
 * Conventions
 *   - every entry point returns 0 on success, a negative DMB_ERR_* code otherwise and never
 *     aborts the process (contrast gaterecurrent2dnoind_kernel.cu:544-549, `exit(-1)`);
 *     `dmb_b200_last_error()` returns a thread-local message for the last failure;
 *   - kernels are launched on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) -- contrast the reference op which always uses the legacy stream
 *     (gaterecurrent2dnoind_kernel.cu:542);
 *   - tensors are dense, contiguous, in the layout stated per function; "NCDHW" etc. are the
 *     reference layouts, "blocked" = [B][C/8][D][H][W][8] bf16 is the internal layout of the tensor-core trunk;
 *   - all pointers are device pointers unless the name ends in `_host`;
 *   - the library is re-entrant per (device, stream); it keeps no global mutable state besides
 *     lazily-resolved driver entry points and per-device attribute caches.
 */
#ifndef DMB_B200_H_
#define DMB_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DMB_OK 0
#define DMB_ERR_INVALID (-1)      /* bad argument (shape, dtype, alignment, null pointer) */
#define DMB_ERR_CUDA (-2)         /* a CUDA runtime / driver call failed */
#define DMB_ERR_UNSUPPORTED (-3)  /* valid request this build has no kernel for */

#define DMB_F32 0
#define DMB_BF16 1

/* library / ABI version, bumped whenever a signature changes */
int dmb_b200_abi_version(void);
const char* dmb_b200_last_error(void);
/* number of kernels this library has launched in the calling process (all threads) */
int64_t dmb_b200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Raw cost-volume builders.  `disp_idx_host` holds the D integer disparities enumerated by the
 * caller exactly like the reference (`int(linspace(start, start+max-1, D))`, cat_fms.py:27-35).
 * Features are [B,C,H,W] float32 (NCHW).
 * ---------------------------------------------------------------------------------------- */

/* cat_fms (dmb/modeling/stereo/cost_processors/utils/cat_fms.py:7-48)
 * out: [B,2C,D,H,W] float32, every element written (zeros where the reference leaves zeros). */
int dmb_b200_cat_volume(const float* left, const float* right, float* out,
                        int B, int C, int H, int W, const int* disp_idx_host, int D, void* stream);

/* dif_fms (dmb/modeling/stereo/cost_processors/utils/dif_fms.py:7-46); out: [B,C,D,H,W] float32 */
int dmb_b200_dif_volume(const float* left, const float* right, float* out,
                        int B, int C, int H, int W, const int* disp_idx_host, int D, void* stream);

/* group-wise correlation (GwcNet eq. 3; NO reference code, see oracle/dmb_oracle.py:gwc_volume)
 * out: [B,G,D,H,W] float32; C % G == 0. */
int dmb_b200_gwc_volume(const float* left, const float* right, float* out,
                        int B, int C, int H, int W, int G, const int* disp_idx_host, int D, void* stream);

/* fast_cat_fms / fast_dif_fms (cat_fms.py:51-82, dif_fms.py:49-86) with inverse_warp_3d
 * (dmb/modeling/stereo/layers/inverse_warp_3d.py:4-52) folded in.
 * disp_sample: [B,D,H,W] float32 (per-pixel disparity samples).
 * mode 0: concat -> out [B,2C,D,H,W];  mode 1: difference -> out [B,C,D,H,W];
 * mode 2: difference then p-norm over C (normalize=True) -> out [B,D,H,W]. */
int dmb_b200_warp_volume(const float* left, const float* right, const float* disp_sample, float* out,
                         int B, int C, int H, int W, int D, int mode, float p, void* stream);

/* ------------------------------------------------------------------------------------------
 * 3-D convolution, generic direct kernel (any channel count / 3-D kernel size / stride /
 * transposed), fp32 NCDHW.  Replaces the cuDNN calls behind conv3d_bn[_relu] / deconv3d_bn
 * (dmb/modeling/stereo/layers/basic_layers.py:68-216) with BN folded by the caller:
 *     y = act( conv(x, w) + bias + residual )
 * w_packed: [KD*KH*KW][Cin][Cout] float32 (caller repacks; for a transposed conv the caller
 * passes the ConvTranspose3d weight [Cin,Cout,k,k,k] permuted the same way, un-flipped).
 * bias: [Cout] or NULL.  residual: same shape as y or NULL.  relu: 0/1 applied last.
 * dims_in = {D,H,W} of x, dims_out = {D,H,W} of y.
 * ---------------------------------------------------------------------------------------- */
int dmb_b200_conv3d_direct(const float* x, const float* w_packed, const float* bias, const float* residual,
                           float* y, int B, int Cin, int Cout, const int* dims_in, const int* dims_out,
                           const int* ksize, int stride, int pad, int transposed, int relu, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cost upsampling + disparity regression.
 * ---------------------------------------------------------------------------------------- */

/* Upsample a low-resolution 1-channel cost [B,Dl,Hl,Wl] (float32) to [B,D,H,W] and/or regress
 * the disparity map [B,1,H,W] from the upsampled values without materialising them.
 *   mode 0: trilinear, align_corners=True (F.interpolate in aggregators/PSMNet.py:75-88)
 *   mode 1: ConvTranspose3d(1,1,8,stride 4,pad 2) with `up_weight` [8*8*8] (AcfNet.py:55-57,81-83);
 *           requires D==4*Dl, H==4*Hl, W==4*Wl.
 * cost_out (nullable): [B,D,H,W] float32.  disp_out (nullable): [B,1,H,W] float32 =
 *   sum_d softmax_d(alpha*cost)[d] * (start_disp + d*disp_step)   (normalize=1)
 *   sum_d alpha*cost[d] * (...)                                     (normalize=0)
 * i.e. FasterSoftArgmin / SoftArgmin (faster_soft_argmin.py:51-75, soft_argmin.py:44-75) fused
 * onto the last aggregation write.  disp_values (nullable, [D] float32 device) overrides the
 * arithmetic ramp with explicit sample values (the frozen `disp_regression.weight`). */
int dmb_b200_upsample_regress(const float* cost_low, const float* up_weight, float* cost_out, float* disp_out,
                              int B, int Dl, int Hl, int Wl, int D, int H, int W, int mode,
                              float alpha, int normalize, float start_disp, float disp_step,
                              const float* disp_values, void* stream);

/* SoftArgmin / FasterSoftArgmin on a materialised cost [B,D,H,W] float32 -> [B,1,H,W].
 * Exactly one of disp_values ([D]) / disp_sample ([B,D,H,W]) may be non-NULL; if both are NULL
 * the ramp start_disp + d*disp_step is used. */
int dmb_b200_soft_argmin(const float* cost, float* disp_out, int B, int D, int H, int W,
                         float alpha, int normalize, float start_disp, float disp_step,
                         const float* disp_values, const float* disp_sample, void* stream);

/* LocalSoftArgmin (disp_predictors/local_soft_argmin.py:47-105) */
int dmb_b200_local_soft_argmin(const float* cost, float* disp_out, int B, int D, int H, int W,
                               int radius, int radius_dilation, float alpha, float start_disp, float dilation,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * dmb.ops: SPN 3-neighbour gated scan (dmb/ops/spn/src/gaterecurrent2dnoind_kernel.cu).
 * All tensors [N,C,H,W] float32.  One launch per call (the reference issues W or H launches).
 * ---------------------------------------------------------------------------------------- */
int dmb_b200_spn_forward(const float* X, const float* G1, const float* G2, const float* G3, float* Hout,
                         int N, int C, int H, int W, int horizontal, int reverse, void* stream);
/* grad_out is not modified (the reference accumulates into it in place); the four gradients
 * are fully written. */
int dmb_b200_spn_backward(const float* X, const float* G1, const float* G2, const float* G3,
                          const float* Hout, const float* grad_out,
                          float* gX, float* gG1, float* gG2, float* gG3,
                          int N, int C, int H, int W, int horizontal, int reverse, void* stream);

/* ------------------------------------------------------------------------------------------
 * GANet aggregation layers (NO reference code; semantics fixed by oracle/dmb_oracle.py).
 * ---------------------------------------------------------------------------------------- */
/* SGA schedule switch: 1 = two bidirectional launches per L2-sized channel group (1.15x algorithmic DRAM traffic,
   latency-bound at present), 0 = four all-channel single-direction launches (default), -1 = re-read
   DMB_B200_SGA_BIDIR.  Returns the previous setting (not a status code). */
int dmb_b200_sga_set_bidirectional(int on);
/* SGA: x [B,C,D,H,W], guidance [B,4,5,C,H,W] (un-normalised; L1-normalised over the 5 taps
 * inside), out [B,C,D,H,W] = max over the 4 scan directions. */
int dmb_b200_sga(const float* x, const float* guidance, float* out,
                 int B, int C, int D, int H, int W, void* stream);
/* LGA: x [B,D,H,W], guidance [B,3,K,K,H,W] with K=2*radius+1 (L1-normalised over all 3*K*K
 * weights inside), out [B,D,H,W]. */
int dmb_b200_lga(const float* x, const float* guidance, float* out,
                 int B, int D, int H, int W, int radius, void* stream);

/* ------------------------------------------------------------------------------------------
 * Tensor-core (tcgen05) trunk: channels-last bf16 activations, optionally as a (hi, lo) split
 * pair that carries ~16 mantissa bits through bf16 MMAs with fp32 accumulation.
 * ---------------------------------------------------------------------------------------- */

/* correlation1d_cost (cost_processors/utils/correlation1d_cost.py:7-27; the reference goes through the un-vendored
 * SpatialCorrelationSampler): out[b, j, y, x] = leaky_relu( sum_c L[b,c,y,x] * R[b,c,y,x - (max_disp-1-j)] ), zero where
 * the shifted column leaves the image; out [B, max_disp, H, W] float32.  W % 4 == 0. */
int dmb_b200_corr1d_volume(const float* left, const float* right, float* out, int B, int C, int H, int W, int max_disp,
                           float negative_slope, void* stream);

/* cat volume straight into the trunk's blocked layout: out_hi/out_lo [B][2C/8][D][H][W][8] bf16
 * from the float32 NCHW features (out_lo NULL => plain bf16, no split).  C % 8 == 0. */
int dmb_b200_cat_volume_blocked(const float* left, const float* right, void* out_hi, void* out_lo,
                                int B, int C, int H, int W, const int* disp_idx_host, int D, int fp16, void* stream);

/* 3x3x3 convolutions of the trunk on tcgen05, blocked channels-last 16-bit activations.
 *   kind 0: stride 1, pad 1, one MMA per tap     (conv3d_bn[_relu], basic_layers.py:68-177)
 *   kind 3: stride 1, pad 1, the three kw taps merged into the MMA N dimension (same result,
 *           3x fewer A-operand reads; the production kernel for stride-1 layers)
 *   kind 1: stride 2, pad 1 (even input extents)  (Hourglass conv1/conv3, utils/hourglass.py:35-48)
 *   kind 4: the same convolution on the kw-merged scheme (row-parity TMA boxes; production kernel)
 *   kind 2: transposed, stride 2, pad 1, output_padding 1 (Hourglass conv5/conv6, :53-60)
 *   kind 5: the same transposed convolution for Cin % 64 == 0: one pass = all 64 input channels x 16 output
 *           channels, so every output element is written once (kind 2 accumulates two 32-input-channel
 *           passes in place); measured slower than kind 2 (twice the MMAs), kept for A/B
 *   kind 6: the same transposed convolution for Cin == 64: one pass = all 64 input channels x 32 output channels,
 *           run as three class-group launches (every 3x3x3 tap belongs to exactly one output parity class, so
 *           the weights split by class: 12 + 12 + 3 taps); every output element is written once with kind 2's MMA
 *           count; the production kernel for the hourglass' 64-channel transposed layers
 * x_hi/x_lo: [B][Cin/8][D][H][W][8] (x_lo NULL => single plane, else the (hi,lo) split pair);
 * B,D,H,W are the INPUT extents; the output grid is the same / halved / doubled.
 * w_blob: packed by dmb_b200_conv3d_tc_pack_weights with the same split / fp16 / kind; w_scale its
 * `scale`.  bias [Cout] f32 or NULL (BatchNorm folded by the caller, as for conv3d_direct).
 * Cout % 32 == 0: y_hi/y_lo [B][Cout/8][Do][Ho][Wo][8] (+ optional residual res_hi/res_lo of the
 *                 same geometry, added before the ReLU);
 * Cout == 1     : y_f32 [B,1,Do,Ho,Wo] float32 (+ optional res_f32 of the same shape) -- the 32->1
 *                 classifier heads (aggregators/PSMNet.py:41-52).
 * Cin % 32 == 0.  Wider layers run as several passes accumulating in place.
 * fp16: 0 = bfloat16 elements, 1 = IEEE half elements (all 16-bit tensors of a call share it). */
int dmb_b200_conv3d_tc(const void* x_hi, const void* x_lo, int Cin, const void* w_blob, float w_scale,
                       const float* bias, const void* res_hi, const void* res_lo, void* y_hi, void* y_lo, int Cout,
                       float* y_f32, const float* res_f32, int B, int D, int H, int W, int kind, int relu, int fp16,
                       void* stream);
/* 2-D 3x3 / stride 1 / pad 1 convolution on tcgen05 -- the confidence heads' conv_bn_relu(192, 64) on a full-resolution
 * cost volume (dmb/modeling/stereo/cmn/cmn.py:29-32): the stride-1 kernel (kind 3) on one depth plane with only the
 * kd = 1 taps issued.  x: blocked [B][Cin/8][1][H][W][8]; w_blob: dmb_b200_conv3d_tc_pack_weights(kind 3) of the 2-D
 * weight embedded in a [27][Cin][Cout] tensor (kd = 0 / 2 taps zero); other arguments as dmb_b200_conv3d_tc. */
int dmb_b200_conv2d_tc(const void* x_hi, const void* x_lo, int Cin, const void* w_blob, float w_scale, const float* bias,
                       const void* res_hi, const void* res_lo, void* y_hi, void* y_lo, int Cout, int B, int H, int W,
                       int relu, int fp16, void* stream);
/* y = a + b on blocked 16-bit activations (n_blocks = B * C/8 * D*H*W 16-byte voxel blocks; lo planes all NULL or all
 * given): the post-ReLU skip additions of GCAggregator (aggregators/GCNet.py:108-116). */
int dmb_b200_blocked_add(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* y_hi, void* y_lo,
                         int64_t n_blocks, int fp16, void* stream);
/* y[b][s] = sum_c w[c] * x[b][c][s] over a blocked activation [B][C/8][S][8] (hi, optional lo): the 1x1
 * Conv2d(C, 1, bias=False) closing a confidence head (cmn.py:31).  y: [B][S] float32. */
int dmb_b200_blocked_dot(const void* x_hi, const void* x_lo, const float* w, float* y, int B, int C, int64_t S, int fp16,
                         void* stream);

/* Fused classifier head (replaces the nn.Sequential(conv3d_bn_relu(32,32), Conv3d(32,1,3,1,1,bias=False)) of
 * aggregators/PSMNet.py:41-52 / AcfNet.py:43-53): the 32->32 stride-1 layer (w_blob packed for kind 3, BatchNorm
 * folded, bias or NULL) runs on tcgen05; its epilogue does not store the activation a but projects it on the 27 taps
 * of the Conv3d(32,1) weight (head_w as [27][32] float32) and, marching along depth, sums the three depth taps in
 * registers: head_t holds, per batch element, the nine planes Q[kh][kw][d][h][w] followed by 2 x 9 spill planes per
 * depth segment of the launch's schedule (dmb_b200_conv3d_tc_head_floats gives the total size in floats).
 * dmb_b200_head_gather then forms y[b,d,h,w] = res? + sum_{kh,kw} Q[kh][kw][d][h+kh-1][w+kw-1] (zero padding, spills
 * added at segment borders): the 32->1 convolution as one streaming pass over 9 (not 27) values per output instead
 * of a second tensor-core launch.  Both calls must see the same B, D, H, W on the same device. */
int64_t dmb_b200_conv3d_tc_head_floats(int B, int D, int H, int W);
int dmb_b200_conv3d_tc_head(const void* x_hi, const void* x_lo, const void* w_blob, float w_scale, const float* bias,
                            const float* head_w, float* head_t, int B, int D, int H, int W, int relu, int fp16,
                            void* stream);
int dmb_b200_head_gather(const float* head_t, const float* res, float* y, int B, int D, int H, int W, void* stream);
/* ------------------------------------------------------------------------------------------
 * Stereo focal loss on a raw cost volume.  Replaces
 * StereoFocalLoss.loss_per_level (losses/stereo_focal_loss.py:63-101) + LaplaceDisp2Prob (losses/utils/
 * disp2prob.py:29-173) and their autograd: one pass over the volume per direction instead of five materialised
 * [B,D,H,W] intermediates.  cost [B,D,H,W]; gt [B,1,H,W] already at the cost's resolution (the caller rescales /
 * pools it, stereo_focal_loss.py:66-73); variance = var_map [B,1,H,W] or (NULL, var_scalar); the disparity samples
 * are either shared (disp_values [D], device) or per pixel (disp_sample [B,D,H,W]) -- exactly one of the two.
 * lower/upper: the outer validity mask lower < gt < upper (:78-81); inner_end = start + max_disp - 1 (disp2prob.py:60).
 * forward : sums[0] += sum of the per-pixel losses, sums[1] += number of valid pixels (float64, zeroed by the
 *           caller; loss = sums[0] / max(sums[1], 1)); stats [B,2,H,W] (log-partition, weight sum) for backward or NULL.
 * backward: gscale = device pointer to (upstream gradient * level weight / max(#valid, 1)); writes dcost [B,D,H,W]
 *           and, when dvar != NULL (needs var_map), d(loss)/d(variance) [B,1,H,W]. */
int dmb_b200_focal_loss_forward(const float* cost, const float* gt, const float* var_map, float var_scalar,
                                const float* disp_values, const float* disp_sample, int B, int D, int H, int W,
                                float lower, float upper, float inner_end, float coefficient, double* sums, float* stats,
                                void* stream);
int dmb_b200_focal_loss_backward(const float* cost, const float* gt, const float* var_map, float var_scalar,
                                 const float* disp_values, const float* disp_sample, const float* stats,
                                 const float* gscale, int B, int D, int H, int W, float lower, float upper,
                                 float inner_end, float coefficient, float* dcost, float* dvar, void* stream);

/* The static schedule dmb_b200_conv3d_tc would use for input extents B,D,H,W (host arithmetic only; no launch):
 * out[8] = {tiles_h, tiles_w, depth segments, planes per segment, work items, grid, tile rows, tile column step}. */
int dmb_b200_conv3d_tc_schedule(int kind, int B, int D, int H, int W, int* out);
/* Debug: device buffer of 3 x 4096 int64 that CTA 0 of every following conv3d_tc launch fills with clock64()
 * stamps of its MMA-issue, epilogue and TMA-producer roles (tools/tc_trace.py); NULL switches tracing off. */
int dmb_b200_debug_set_trace(long long* device_buffer);
/* w_packed: [27][Cin][Cout] float32 (the conv3d_direct packing; for kind 2 the ConvTranspose3d
 * weight packed the same way, un-flipped) -> w_blob (16-bit), one [27][cbk][32|64][8] block per
 * (32 out, 8*cbk in) channel pair; split=1 stores hi rows then lo rows.  `scale` (a power of two)
 * pre-multiplies the weights so that the fp16 `lo` parts stay out of the subnormal range; the
 * kernel's epilogue multiplies the accumulators by 1/w_scale (exact). */
int dmb_b200_conv3d_tc_pack_weights(const float* w_packed, void* w_blob, int Cin, int Cout, int split, int fp16,
                                    float scale, int kind, void* stream);
/* bytes of the packed blob */
int64_t dmb_b200_conv3d_tc_weight_bytes(int Cin, int Cout, int split, int kind);
/* 1 if this device/driver can run the tcgen05 path */
int dmb_b200_conv3d_tc_available(void);

/* layout helpers for the trunk boundary: [B,C,D,H,W] float32 <-> [B][C/8][D][H][W][8] 16-bit (hi[,lo]) */
int dmb_b200_ncdhw_to_blocked(const float* x, void* y_hi, void* y_lo, int B, int C, int D, int H, int W, int fp16, void* stream);
/* the same conversion with every row W-PARITY-SPLIT: [B][C/8][D][H][even | odd][W/2][8] -- the operand layout of the
   stride-2 tcgen05 weight gradient (dmb_b200_conv3d_wgrad_tc with stride 2); W even */
int dmb_b200_ncdhw_to_blocked_wsplit(const float* x, void* y_hi, void* y_lo, int B, int C, int D, int H, int W, int fp16, void* stream);
int dmb_b200_blocked_to_ncdhw(const void* x_hi, const void* x_lo, float* y, int B, int C, int D, int H, int W, int fp16, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-mode kernels (BASELINE config 5).  The reference trains through autograd over
 * cuDNN/ATen: Conv3d + BatchNorm3d(batch statistics) + ReLU (layers/basic_layers.py:68-216),
 * upsampling (aggregators/PSMNet.py:75-88, AcfNet.py:55-57,81-83), softmax regression
 * (disp_predictors/faster_soft_argmin.py:51-75) and the volume builder (utils/cat_fms.py:7-48).
 * All tensors float32 NCDHW, S = D*H*W.  conv dgrad is dmb_b200_conv3d_direct with the weight
 * roles swapped (see densematchingbenchmark_b200/ops/autograd.py).
 * ---------------------------------------------------------------------------------------- */
/* sums [2C] float64, zero-initialised by the caller: sums[c] += sum z, sums[C+c] += sum z^2.
 * Kept as raw sums so that a multi-GPU run can all-reduce them before finalize (SyncBN). */
int dmb_b200_bn_stats(const float* z, double* sums, int B, int C, long long S, void* stream);
/* mean/invstd/scale/shift [C] from the sums over `count` elements per channel; running stats
 * (nullable) updated like nn.BatchNorm3d: momentum, unbiased variance. gamma/beta nullable. */
int dmb_b200_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float eps,
                         float momentum, float* running_mean, float* running_var, float* mean, float* invstd,
                         float* scale, float* shift, int C, void* stream);
/* y = relu?( z * scale[c] + shift[c] + residual ) */
int dmb_b200_bn_apply(const float* z, const float* scale, const float* shift, const float* residual, float* y,
                      int B, int C, long long S, int relu, void* stream);
/* g = dy * (y_relu > 0) (y_relu nullable: no ReLU).  sums [2C] float64 zero-initialised:
 * sums[c] += sum g; sums[C+c] += sum g * (z - mean) * invstd (skipped when mean is NULL). */
int dmb_b200_bn_backward_reduce(const float* dy, const float* y_relu, const float* z, const float* mean,
                                const float* invstd, double* sums, int B, int C, long long S, void* stream);
/* dz = gamma*invstd*(g - sums[c]/count - xhat*sums[C+c]/count) (mean != NULL) or g; dres (nullable) = g. */
int dmb_b200_bn_backward_apply(const float* dy, const float* y_relu, const float* z, const float* mean,
                               const float* invstd, const float* gamma, const double* sums, double count,
                               float* dz, float* dres, int B, int C, long long S, void* stream);
/* 3x3x3 conv weight gradient, regular-conv geometry (in = out*stride - pad + tap), stride 1 or 2:
 * dw_packed [27][Cin][Cout] float32 (the conv3d_direct packing), zero-initialised by the caller,
 * += sum_{b,o} dz[b,co,o] * x[b,ci,in(o,tap)].  A ConvTranspose3d's gradient is obtained by swapping
 * the roles of x and dz. */
int dmb_b200_conv3d_wgrad(const float* x, const float* dz, float* dw_packed, int B, int Cin, int Cout,
                          const int* dims_in, const int* dims_out, int stride, int pad, void* stream);
/* weight gradient of AcfNet's ConvTranspose3d(1,1,8,4,2) upsampling: dw [512] zero-initialised. */
int dmb_b200_upsample_deconv_wgrad(const float* cost_low, const float* dcost, float* dw, int B, int Dl, int Hl,
                                   int Wl, int D, int H, int W, void* stream);
/* backward of the trilinear (align_corners=True) cost upsampling: dcost [B,D,H,W] -> dcost_low [B,Dl,Hl,Wl] */
int dmb_b200_upsample_trilinear_backward(const float* dcost, float* dcost_low, int B, int Dl, int Hl, int Wl,
                                         int D, int H, int W, void* stream);
/* Weight gradient of a 3x3x3 / pad 1 convolution of stride 1 or 2 on tcgen05 (csrc/wgrad_tc.cu; replaces cuDNN's
   wgrad behind nn.Conv3d / nn.ConvTranspose3d in the backward of layers/basic_layers.py:68-216):
       dw[tap][ca][cg] += sum_{b,d,h,w} a[b][ca][s*d+kd-1][s*h+kh-1][s*w+kw-1] * g[b][cg][d][h][w]        (s = stride)
   g has extents D, H, W; a has stride times those.  For a strided Conv3d a = layer input, g = gradient w.r.t. the layer
   output; for the stride-2 ConvTranspose3d the roles swap (a = output gradient, g = layer input).  Both are blocked
   16-bit split pairs [B][C/8][..][8] (hi and lo planes, dmb_b200_ncdhw_to_blocked); with stride 2, `a` must be the
   W-parity-split variant (dmb_b200_ncdhw_to_blocked_wsplit).  fp16 = 0: bfloat16 elements, 1: IEEE half.
   dw [27][Ca][Cg] fp32 is ACCUMULATED into (zero it first); Ca, Cg multiples of 32. */
int dmb_b200_conv3d_wgrad_tc(const void* a_hi, const void* a_lo, const void* g_hi, const void* g_lo, float* dw, int B,
                             int Ca, int Cg, int D, int H, int W, int stride, int fp16, void* stream);
/* backward of dmb_b200_soft_argmin (disp_values / ramp variants): dcost [B,D,H,W] fully written */
int dmb_b200_soft_argmin_backward(const float* cost, const float* grad_disp, float* dcost, int B, int D, int H, int W,
                                  float alpha, int normalize, float start_disp, float disp_step,
                                  const float* disp_values, void* stream);
/* backward of dmb_b200_cat_volume: dvol [B,2C,D,H,W] -> dleft, dright [B,C,H,W]; D <= 256 */
int dmb_b200_cat_volume_backward(const float* dvol, float* dleft, float* dright, int B, int C, int H, int W,
                                 const int* disp_idx_host, int D, void* stream);
/* backward of dmb_b200_dif_volume (dif_fms.py:7-46, out = ref - shifted tgt): dvol [B,C,D,H,W] -> dleft = +sum over
   the valid disparities, dright = -sum at the shifted columns; D <= 256 */
int dmb_b200_dif_volume_backward(const float* dvol, float* dleft, float* dright, int B, int C, int H, int W,
                                 const int* disp_idx_host, int D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small-vector exchange between the GPUs of one box over peer memory (csrc/peer_comm.cu): the per-layer
 * statistics of synchronised BatchNorm in data-parallel training (dmb/apis/train.py:95-97, apex
 * convert_syncbn_model; torch.distributed all-reduce / all-gather of 2*C numbers per BatchNorm layer and direction).
 * One kernel per exchange on the caller's stream: peer stores into every rank's receive buffer, a system-scope
 * release / acquire flag per (slot, rank), a local fixed-order reduction.  One process per GPU; buffers are shared
 * through CUDA IPC handles that the host exchanges once (utils/dist_utils.py:PeerComm).  All ranks must issue the same
 * sequence of exchanges on one stream each.  seq = 0 numbers the exchanges with a counter kept in the rank's own buffer
 * (no per-launch host state: the launch can be captured in a CUDA graph and replayed); seq = 1, 2, 3, ... passes the
 * number explicitly (do not mix the two on one set of buffers).
 * ------------------------------------------------------------------------------------------ */
int64_t dmb_b200_peer_buffer_bytes(void);
int dmb_b200_peer_alloc(void** ptr);                             /* this rank's zeroed receive buffer (cudaMalloc) */
int dmb_b200_peer_free(void* ptr);
int dmb_b200_peer_export(void* ptr, void* handle64);             /* 64-byte CUDA IPC handle */
int dmb_b200_peer_import(const void* handle64, void** ptr);      /* map a peer's buffer into this process */
int dmb_b200_peer_close(void* ptr);
/* bufs: host array of `world` device pointers (own buffer at index `rank`); src: nbytes (multiple of 4, <= 8192);
 * mode 0: gather -> dst [world][nbytes]; mode 1: float64 sum -> dst [nbytes]; mode 2: float32 sum -> dst [nbytes] */
int dmb_b200_peer_exchange(void* const* bufs, int rank, int world, long long seq, const void* src, void* dst, int nbytes,
                           int mode, void* stream);
/* float32 sums over the ranks of TWO vectors in one exchange: dst0[n0], dst1[n1] ((n0 + n1) * 4 <= 8192).  The backward
 * statistics of synchronised BatchNorm (sum(dy), sum(dy * (x - mean)): torch.batch_norm_backward_reduce -> all_reduce ->
 * batch_norm_backward_elemt in torch.nn.SyncBatchNorm / apex.parallel.SyncBatchNorm) */
int dmb_b200_peer_sum2_f32(void* const* bufs, int rank, int world, const float* src0, int n0, const float* src1, int n1,
                           float* dst0, float* dst1, void* stream);
/* forward statistics of synchronised BatchNorm, exchange FUSED with the merge: this rank's mean[C], invstd[C] (1 / sqrt(var
 * + eps), biased variance) and element count -> the joint batch's out_mean[C] / out_invstd[C], every rank's count
 * (out_counts[world] int32, may be NULL) and the running-statistics update (momentum; unbiased variance; pointers may be
 * NULL).  Replaces all_gather + batch_norm_gather_stats_with_counts of torch.nn.SyncBatchNorm (apex: welford_parallel).
 * (2 * C + 1) * 4 <= 8192 */
int dmb_b200_peer_bn_forward(void* const* bufs, int rank, int world, const float* mean, const float* invstd, float count,
                             float eps, float momentum, float* running_mean, float* running_var, float* out_mean,
                             float* out_invstd, int* out_counts, int C, void* stream);
/* exchanges completed so far on this rank's buffer in seq = 0 mode (synchronises the device) */
int dmb_b200_peer_count(const void* own_buf, long long* count);

#ifdef __cplusplus
}
#endif
#endif /* DMB_B200_H_ */
