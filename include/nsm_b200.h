/* nsm_b200.h -- C ABI of libnsm_b200.so, the B200 (sm_100a) implementation of the Neural-Shadow-Mapping U-Net
 * hot path of SDU-Gary/PCSS-Unet.
 *
 * Conventions:
 *   - plain C: raw device pointers, explicit sizes, `void* stream` is a cudaStream_t (NULL = legacy default
 *     stream); no C++ types, no exceptions cross this boundary;
 *   - every function returns 0 on success; on failure a non-zero code and a message retrievable with
 *     nsm_last_error() (thread-local).  Work is only enqueued on `stream`; nothing synchronises except the
 *     *_host entry point, which is documented to;
 *   - the caller owns every buffer (weights blob, workspace, inputs, outputs).  Sizes are queried first;
 *   - thread-safe: PyTorch calls backward from its autograd worker thread (main.py:281).
 *
 * `mode` selects the arithmetic:
 *   NSM_MODE_BF16  activations/weights bf16, fp32 accumulate, rounding points of torch.autocast(bfloat16)
 *   NSM_MODE_FP32  fp32-accurate: every tensor is held as hi+lo half-precision planes (fp16 pairs in eval, bf16 pairs
 *                  in NSM_MODE_FP32_TRAIN; 8-bit cross planes for the decoder's 3x3 layers), products
 *                  hi*hi+hi*lo+lo*hi are accumulated in fp32 on the tensor cores (network output within 1e-4 of the
 *                  fp32 reference)
 *
 * Each entry point cites the reference interface (file:line in SDU-Gary/PCSS-Unet) it stands in for.
 */
#ifndef NSM_B200_H_
#define NSM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSM_MODE_BF16 0
#define NSM_MODE_FP32 1       /* eval / forward-only: hi+lo fp16 planes (|activation| <= 65504), 22 significand bits */
#define NSM_MODE_FP32_TRAIN 2 /* training: hi+lo bf16 planes (full fp32 range, safe for tiny gradients), 16 bits     */
/* Operand storage format of the decoder's 3x3 convolutions inside NSM_MODE_FP32 (stage-level entry points only:
 * nsm_nchw_to_planes / nsm_planes_to_nchw / nsm_pack_conv_weight / nsm_conv_fwd take it as `mode`, nsm_upsample_match
 * produces it from an NSM_MODE_FP32 source): plane 0 = fp16 hi, plane 1 = per 16 channels 16+16 e4m3 bytes that one 8-bit
 * tensor-core MMA turns into both cross terms of the hi+lo product.  Results / skips / pooled tensors of such a convolution
 * are plain NSM_MODE_FP32 planes.  The whole-network calls pick it internally (NSM_NO_X8=1 in the environment disables). */
#define NSM_FMT_F16_X8 3

/* number of fp32 tensors nsm_unet_pack() consumes, in this order:
 *   for K in conv2..conv9 (Unetmodel.py:39-61):
 *     conv.0.weight, conv.0.bias, conv.1.weight, conv.1.bias, conv.1.running_mean, conv.1.running_var,
 *     conv.4.weight, conv.4.bias, conv.5.weight, conv.5.bias, conv.5.running_mean, conv.5.running_var
 *   conv10.weight, conv10.bias (Unetmodel.py:63)                                                      */
#define NSM_UNET_NUM_TENSORS 98

/* Order-independent accumulator slot.  Every cross-block reduction of this library (BatchNorm statistics and backward
 * sums, loss sums, per-channel input statistics, the gradient norm) adds its per-block fp64 partials as 128-bit
 * two's-complement fixed-point numbers (value = hi + lo * 2^-64) with integer atomics: the total is bit-identical
 * from run to run whatever order the blocks arrive in (the reference asks for deterministic kernels, main.py:81-82;
 * fp64 atomicAdd is not).  `bad` != 0: a non-finite or out-of-range (>= 2^62) partial was added, the slot reads as NaN.
 * The caller zero-fills slots before the accumulating call; nsm_acc_to_double converts n slots to fp64. */
typedef struct nsm_acc {
  unsigned long long lo, hi, bad, pad;
} nsm_acc;
int nsm_acc_to_double(const nsm_acc* acc, long long n, double* out /* device */, void* stream);

const char* nsm_last_error(void);
int nsm_version(void);
/* kernels launched by this library since load (bench.py: gpu_launches) */
long long nsm_launch_count(void);
/* TMA descriptors (CUtensorMap) are encoded once per (buffer, shape, box, swizzle) and kept in a table behind a mutex:
 * lookups that found / did not find their descriptor since load (a steady-state frame or step adds hits only) */
void nsm_tmap_cache_stats(long long* hits, long long* misses);
/* 0 if the current CUDA device is compute capability 10.x (B200), else non-zero + message */
int nsm_check_device(void);

/* ---------------------------------------------------------------------------------------------------------
 * Whole network, eval mode:  Unet.forward under model.eval()  (Unetmodel.py:90-149; call sites infer.py:65,
 * inference.py:194, main.py:611, validate_consistency.py:49,68)
 * --------------------------------------------------------------------------------------------------------- */
size_t nsm_unet_packed_bytes(int mode);
/* state_dict tensors (device pointers, reference layouts: OIHW weights, [C] vectors) -> packed blob */
int nsm_unet_pack(const float* const* tensors, int mode, void* blob, void* stream);
size_t nsm_unet_workspace_bytes(int B, int H, int W, int mode);
/* x: [B,4,H,W] fp32 NCHW (device).  y: [B,1,H-H%2,W-W%2] fp32 (device).  mean/std: optional [4] device vectors:
 * fuses MmapLiverDataset's (x-mean)/(std+1e-8) (setdata.py:316) into the first load; NULL = input already
 * standardised (what main.py feeds) or raw (what infer.py feeds). */
int nsm_unet_infer(const void* blob, int mode, const float* x, int B, int H, int W, const float* mean,
                   const float* std, float* y, void* workspace, size_t workspace_bytes, void* stream);
/* Same with HOST buffers (pinned or pageable): H2D copy, forward, D2H copy, stream synchronise. */
int nsm_unet_infer_host(const void* blob, int mode, const float* x_host, int B, int H, int W, const float* mean,
                        const float* std, float* y_host, void* workspace, size_t workspace_bytes, void* stream);
/* Output path of infer.py:79-80 fused into the last kernel: y_u8 = (uint8)(y * 255), [B,1,H',W'] bytes,
 * a quarter of the D2H traffic of the fp32 result. */
int nsm_unet_infer_u8(const void* blob, int mode, const float* x, int B, int H, int W, const float* mean,
                      const float* std, uint8_t* y_u8, void* workspace, size_t workspace_bytes, void* stream);
int nsm_unet_infer_host_u8(const void* blob, int mode, const float* x_host, int B, int H, int W, const float* mean,
                           const float* std, uint8_t* y_host, void* workspace, size_t workspace_bytes, void* stream);
/* Frame pipeline for sequences (infer.py:46-80 / inference.py:256-300 process one frame after another): two device
 * staging slots and three internal streams, so that the H2D copy of frame k+1 and the D2H copy of result k-1 overlap the
 * kernels of frame k.  workspace: caller-owned device memory of nsm_unet_pipe_workspace_bytes().  mean/std as above (they
 * must outlive the pipe).  submit() takes PINNED host buffers, exactly one of y_host (fp32) / y_host_u8, and returns once
 * the frame is queued; when submit() of frame k returns, result k-2 is complete; sync() completes all of them.  The input
 * must not be rewritten before the frame's own result is complete. */
size_t nsm_unet_pipe_workspace_bytes(int B, int H, int W, int mode);
int nsm_unet_pipe_create(const void* blob, int mode, int B, int H, int W, const float* mean, const float* std,
                         void* workspace, size_t workspace_bytes, void** pipe);
int nsm_unet_pipe_submit(void* pipe, const float* x_host, float* y_host, uint8_t* y_host_u8);
int nsm_unet_pipe_sync(void* pipe);
int nsm_unet_pipe_destroy(void* pipe);
/* Copy a named intermediate of the last nsm_unet_infer() on this workspace to NCHW fp32 (tests / debugging).
 * names: c2 p2 t3 c3 p3 t4 c4 p4 t5 c5 u6 t6 m6 u7 t7 m7 u8 t8 m8 u9 t9.  *C,*h,*w receive its shape. */
int nsm_unet_tap(const void* workspace, int B, int H, int W, int mode, const char* name, float* out, int* C,
                 int* h, int* w, void* stream);

/* Optional per-launch timing for bench.py: while enabled, every kernel launch of nsm_unet_infer is bracketed by CUDA
 * events on the launch stream.  nsm_profile_read() waits for them and writes "name,ms,flops,bytes\n" lines
 * (algorithmic FLOPs / bytes of that launch) into `out`, then clears the records. */
int nsm_profile_enable(int on);
int nsm_profile_read(char* out, size_t cap);

/* ---------------------------------------------------------------------------------------------------------
 * Stage-level entry points (per-fused-stage parity tests; SURVEY.md 8c tolerance protocol)
 * Activations are NHWC bf16 "planes": plane 0 (= the value in bf16 mode, the hi part in fp32 mode), plane 1
 * (lo part, fp32 mode only).
 * --------------------------------------------------------------------------------------------------------- */
int nsm_nchw_to_planes(const float* x, int N, int C, int H, int W, int mode, void* plane0, void* plane1,
                       void* stream);
int nsm_planes_to_nchw(const void* plane0, const void* plane1, int N, int C, int H, int W, int mode, float* y,
                       void* stream);
/* nn.Conv2d weight [Cout,Cin,k,k] -> GEMM operand [Cout][k*k][Cin] (dgrad!=0: [Cin][k*k mirrored][Cout]) */
int nsm_pack_conv_weight(const float* w, int Cout, int Cin, int ksize, int dgrad, int mode, void* plane0,
                         void* plane1, void* stream);

typedef struct nsm_conv_args {
  int N, H, W, Cin, Cout, ksize; /* ksize 1 or 3 (pad 1); Cin, Cout multiples of 64 */
  int mode;
  const void* in[2];       /* [N,H,W,Cin] planes */
  const void* weight[2];   /* packed planes */
  const float* bias;       /* [Cout] or NULL                           nn.Conv2d bias      Unetmodel.py:21,26 */
  const float* bn_scale;   /* [Cout] or NULL  gamma/sqrt(var+eps)      nn.BatchNorm2d      Unetmodel.py:22,27 */
  const float* bn_shift;   /* [Cout]          beta - mean*scale                                              */
  int lrelu;               /* 1: LeakyReLU(0.2)  Unetmodel.py:23,28;  2: ReLU (VGG19 stacks, customLoss.py:20) */
  void* out[2];            /* [N,H,W,Cout] planes or NULL */
  const void* residual[2]; /* [N,H,W,Cout] planes or NULL: skip add    Unetmodel.py:125,131,137              */
  void* pool[2];           /* [N,H/2,W/2,Cout] planes or NULL: AvgPool2d(2)               Unetmodel.py:40,43,46 */
  float* out_f32;          /* optional [N,H,W,Cout] fp32: conv + bias before BN */
  nsm_acc* stats;          /* optional [2*Cout] slots, zeroed by the caller: += per-channel sum / sum of squares of the
                              stored conv+bias output = train-mode BatchNorm statistics (Unetmodel.py:22,27) */
} nsm_conv_args;
int nsm_conv_fwd(const nsm_conv_args* a, void* stream);

/* Fused eval-mode decoder block, ONE launch (Unetmodel.py:51-60, 134-148 with DoubleConv :20-30):
 *   x = F.interpolate(nn.Upsample(x2)(src), size=(H,W))  -- composite bilinear stencil built on the 3x3 convolution's
 *       operand path in shared memory, never written to HBM
 *   t = LeakyReLU(BN(conv3x3(x)));  o = LeakyReLU(BN(conv1x1(t)))   -- t stays in tensor memory
 *   tail == 0:  out = o + residual                                                       (conv8 + skip, :137)
 *   tail == 1:  y = sigmoid(pixel_shuffle(conv10(o), 2)), optional uint8 copy            (conv9 .. output, :143-148)
 * mode NSM_MODE_BF16 or NSM_MODE_FP32 (eval); weight3 = 3x3 planes as nsm_unet_pack stores them for the decoder
 * (fp32 mode: fp16 hi + 8-bit cross plane), weight1 = 1x1 planes ([Cout][Cmid], fp32 mode: fp16 hi + lo). */
typedef struct nsm_upblock_args {
  int mode, N, Hs, Ws, H, W, Cmid, Cout;   /* Cmid 64|128, Cout 16|64 */
  const void* src[2];                      /* [N,Hs,Ws,Cmid] planes */
  const void* weight3[2];
  const void* weight1[2];
  const float *bias3, *bn_scale3, *bn_shift3;   /* [Cmid] */
  const float *bias1, *bn_scale1, *bn_shift1;   /* [Cout] */
  const void* residual[2];                 /* [N,H,W,Cout] planes or NULL */
  void* out[2];                            /* [N,H,W,Cout] planes (tail == 0) */
  int tail;
  const float *w10, *b10;                  /* conv10 [4][16], [4] fp32 (tail == 1; pre-rounded to bf16 in bf16 mode) */
  float* y;                                /* [N,1,2H,2W] */
  uint8_t* y_u8;                           /* optional */
} nsm_upblock_args;
int nsm_upblock(const nsm_upblock_args* a, void* stream);
/* 1 when nsm_unet_infer runs conv8 / conv9 as fused blocks (default), 0 with NSM_NO_FUSED=1 in the environment */
int nsm_unet_fused_decoder(void);
/* tests: 1 / 0 force the fused / stage-by-stage decoder (the latter leaves u8 t8 u9 t9 visible to nsm_unet_tap), -1 restores
 * the environment default */
int nsm_unet_set_fused_decoder(int on);
/* debugging (NSM_UB_DBG=64): cycle counters of one worker warp of the fused block kernel, read and cleared */
int nsm_upblock_prof(unsigned long long* out32);

/* nn.Upsample(scale_factor=2, bilinear, align_corners=True) then F.interpolate(size=(hd,wd)) -- Unetmodel.py:51-60,
 * 118-141 */
int nsm_upsample_match(const void* const* src, int N, int hs, int ws, int C, void* const* dst, int hd, int wd,
                       int mode, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Objective: nn.L1Loss value + gradient (customLoss.py:96,134,160; pert_loss.py:23,84-90) in one pass.
 *   acc[0] += sum|out-target|   acc[1] += sum_i sum|out-perturbed_i|   acc[2] += #{out<0 or out>1 or NaN}
 *   grad    = coef_l1*sign(out-target) + coef_pert*sum_i sign(out-perturbed_i)        (grad may be NULL)
 * target may be NULL; n_perturbed <= 4.  acc: three slots, zeroed by the caller. */
int nsm_l1_loss_fwd_bwd(const float* out, const float* target, const float* const* perturbed, int n_perturbed,
                        long long numel, float coef_l1, float coef_pert, float* grad, nsm_acc* acc, void* stream);

/* customLoss.EnhancedCustomLoss (customLoss.py:195-238; not used by main.py, which takes pert_loss.EnhancedCustomLoss):
 *   compute_perturbation_loss :226-231   out = clamp(x + eps * noise, lo, hi)      (noise: the caller's randn draw)
 *   F.mse_loss(output, perturbed_output) :238   acc[0] += sum (out - ref)^2;  diff = out - ref (may be NULL): the
 *   gradient of the mean is 2 * diff / numel.  acc: one slot, zeroed by the caller. */
int nsm_add_noise_clamp(const float* x, const float* noise, long long numel, float eps, float lo, float hi, float* out,
                        void* stream);
int nsm_mse_loss_fwd_bwd(const float* out, const float* ref, long long numel, float* diff, nsm_acc* acc, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Perceptual term: MultiLayerVGGLoss (customLoss.py:7-90) = weighted L1 between VGG19 features of output and target.
 * The VGG19 convolutions run through nsm_conv_fwd (lrelu = 2 for the fused ReLU, 0 for a tapped pre-activation
 * feature); these are the passes around them.
 * --------------------------------------------------------------------------------------------------------- */
/* customLoss.py:44-61: clamp(0,1), nan_to_num(nan=.5), grey -> 3 channels, (v-0.485)/(0.229+1e-8).  output, target:
 * [B,1,H,W] fp32; out: planes [2B,H,W,64] (images 0..B-1 = output, B..2B-1 = target; channels 3..63 zero) */
int nsm_vgg_input_prep(const float* output, const float* target, int B, int H, int W, int mode, void* out0, void* out1,
                       void* stream);
/* nn.ReLU [+ nn.MaxPool2d(2) when pool != 0]: in [N,H,W,C] -> out [N,H/2,W/2,C] or [N,H,W,C] */
int nsm_relu_maxpool(const void* in0, const void* in1, int N, int H, int W, int C, int pool, int mode, void* out0,
                     void* out1, void* stream);
/* customLoss.py:76-80: *acc += sum |nan_to_num(a) - nan_to_num(b)| where a = the first numel_half elements of the
 * planes, b = the next numel_half (nan -> 0, +inf -> 1, -inf -> -1); acc: one slot, device, zeroed by the caller */
int nsm_feature_l1(const void* f0, const void* f1, long long numel_half, int mode, nsm_acc* acc, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Input statistics / standardisation / perturbation
 * --------------------------------------------------------------------------------------------------------- */
/* calculate_dataset_stats.py:59-79: x [S,C,HW] fp32; means==NULL: sums[c] += sum x, else sums[c] += sum (x-means[c])^2 */
int nsm_channel_sums(const float* x, long long S, int C, long long HW, const double* means, nsm_acc* sums /*[C]*/,
                     void* stream);
/* setdata.py:316: y = (x - mean[c]) / (std[c] + 1e-8) */
int nsm_standardize(const float* x, float* y, long long S, int C, long long HW, const float* mean,
                    const float* std, void* stream);
/* pert_loss.py:50-57: out[i] = x + (noise[i] * stds[c]) * std_factor, i < count; out [count,B,C,HW];
 * noise [count,C,B,HW] (one contiguous block per randn_like draw of the reference) */
int nsm_perturb(const float* x, const float* noise, float* out, int count, long long B, int C, long long HW,
                const float* stds, float std_factor, void* stream);


/* ---------------------------------------------------------------------------------------------------------
 * Training path (Unet.forward under model.train() + autograd backward, main.py:263-281).  All 17 convolutions run
 * through nsm_conv_fwd (forward and dgrad = nsm_conv_fwd with weights packed with dgrad=1) and nsm_wgrad; thin layers
 * (16 / 4 channels) are zero-padded to 64 channels.  `p0/p1` arguments are the two planes of an NHWC tensor.
 * --------------------------------------------------------------------------------------------------------- */
/* nn.BatchNorm2d in training mode (Unetmodel.py:22,27): batch statistics, running-stat update, affine + LeakyReLU(0.2)
 * + Dropout2d mask (Unetmodel.py:23-24) (+ skip add :125 / AvgPool2d :40) */
int nsm_bn_stats(const void* z0, const void* z1, long long P, int C, int mode, nsm_acc* sums /*[2C], zeroed*/,
                 void* stream);
int nsm_bn_finalize(const nsm_acc* sums, long long P, int C, const float* gamma, const float* beta, float eps,
                    float momentum, int updates, float* running_mean, float* running_var, float* scale, float* shift,
                    float* save_mean, float* save_invstd, void* stream);
int nsm_bn_act(const void* z0, const void* z1, int N, int H, int W, int C, int mode, const float* scale,
               const float* shift, const float* mask /*[N][C] or NULL*/, int lrelu, const void* res0, const void* res1,
               void* out0, void* out1, void* pool0, void* pool1, void* stream);
/* backward of the above: g = dy*mask*LeakyReLU'(.), BatchNorm backward with batch statistics */
int nsm_bn_bwd(const void* dy0, const void* dy1, const void* z0, const void* z1, int N, int H, int W, int C, int mode,
               const float* scale, const float* shift, const float* mask, const float* mean, const float* invstd,
               int lrelu, nsm_acc* sums /*[3C] scratch, zeroed*/, void* dz0, void* dz1, float* dgamma, float* dbeta,
               float* dbias /*conv bias grad = sum dz, or NULL*/, void* stream);
/* adjoint of AvgPool2d(2) (+ optional second gradient `a`), plain add, adjoint of one bilinear align_corners resize */
int nsm_pool_bwd_add(const void* a0, const void* a1, const void* dp0, const void* dp1, void* out0, void* out1, int N,
                     int H, int W, int C, int mode, void* stream);
int nsm_planes_add(const void* a0, const void* a1, const void* b0, const void* b1, void* out0, void* out1,
                   long long numel, int mode, void* stream);
int nsm_bilinear_bwd(const void* dout0, const void* dout1, int N, int ho, int wo, int C, void* din0, void* din1, int hi,
                     int wi, int mode, void* stream);
/* adjoint of nsm_upsample_match in one pass (dout [N,ho,wo,C] -> din [N,hi,wi,C]) */
int nsm_upsample_match_bwd(const void* dout0, const void* dout1, int N, int ho, int wo, int C, void* din0, void* din1,
                           int hi, int wi, int mode, void* stream);
/* network input (even fix + pixel_unshuffle, Unetmodel.py:92-101; 16 -> 64 channels zero padded) and its adjoint */
int nsm_train_input_prep(const float* x, int N, int Hin, int Win, void* out0, void* out1, int mode, void* stream);
int nsm_train_input_grad(const void* d0, const void* d1, int N, int H, int W, float* dx, int mode, void* stream);
/* sigmoid(pixel_shuffle(c10)) (Unetmodel.py:147-148) and its adjoint; c10 planes have 64 channels, 4 used */
int nsm_sigmoid_shuffle_fwd(const void* c0, const void* c1, int N, int h, int w, int mode, float* y, void* stream);
int nsm_sigmoid_shuffle_bwd(const float* dy, const float* y, int N, int h, int w, int mode, void* d0, void* d1,
                            void* stream);
int nsm_pack_conv_weight_padded(const float* w, int Cout, int Cin, int ksize, int CoutP, int CinP, int dgrad, int mode,
                                void* plane0, void* plane1, void* stream);
int nsm_pad_vector(const float* src, int n, int npad, float fill, int round_bf16, float* dst, void* stream);
/* Pixel-packed thin layers (conv2, conv9 1x1, conv10: 16 / 4 channels).  Instead of zero-padding 16 channels to the 64 the
 * tensor-core kernel needs, four horizontally adjacent pixels x C channels are addressed as ONE pixel of 4C "virtual"
 * channels ([N,H,W,C] and [N,H,W/4,4C] are the same bytes; W % 4 == 0).  A 1x1 convolution becomes a 1x1 convolution
 * with a block-diagonal weight, a 3x3 convolution a 3x3 convolution over pixel groups with a banded weight
 * (nsm_pack_conv_weight_px4; CoutV / CinV = 4*Cout / 4*Cin, or 64 when that is smaller: conv10); nsm_conv_fwd and
 * nsm_wgrad then run on the virtual shapes.  nsm_px4_reduce_dw folds the virtual weight gradient [CoutV][CinV][k][k]
 * back to [Cout][Cin][k][k]; nsm_fold_channel_sums folds per-virtual-channel sums (BatchNorm statistics of the conv
 * epilogue, bias gradients) to the real channels: out[v][c] = sum_g in[v][g*C + c]; nsm_tile_vector repeats a bias. */
int nsm_pack_conv_weight_px4(const float* w, int Cout, int Cin, int ksize, int CoutV, int CinV, int dgrad, int mode,
                             void* plane0, void* plane1, void* stream);
int nsm_px4_reduce_dw(const float* dwv, int Cout, int Cin, int ksize, int CoutV, int CinV, float* dw, void* stream);
int nsm_fold_channel_sums(const nsm_acc* in, int nvec, int CV, int groups, int C, nsm_acc* out, void* stream);
int nsm_tile_vector(const float* src, int n, int rep, int npad, float fill, int round_bf16, float* dst, void* stream);
/* input / output stages with un-padded tensors: x16 planes [N,h,w,16]; c10 planes [N,h,w/4,64] (pixel po of a group
 * holds its 4 channels at [4*po, 4*po+4), channels 16..63 zero) */
int nsm_train_input_prep_c16(const float* x, int N, int Hin, int Win, void* out0, void* out1, int mode, void* stream);
int nsm_train_input_grad_c16(const void* d0, const void* d1, int N, int H, int W, float* dx, int mode, void* stream);
int nsm_sigmoid_shuffle_fwd_px4(const void* c0, const void* c1, int N, int h, int w, int mode, float* y, void* stream);
int nsm_sigmoid_shuffle_bwd_px4(const float* dy, const float* y, int N, int h, int w, int mode, void* d0, void* d1,
                                void* stream);
/* weight gradient dW = dz^T (*) x  (autograd of nn.Conv2d): tcgen05 GEMM over the pixel axis with split-K */
size_t nsm_wgrad_workspace_bytes(int N, int H, int W, int Cout, int Cin, int ksize, int mode);
int nsm_wgrad(const void* dz0, const void* dz1, const void* x0, const void* x1, int N, int H, int W, int Cout, int Cin,
              int ksize, int mode, int Cout_real, int Cin_real, void* workspace, size_t workspace_bytes, float* dw,
              void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Optimizer + gradient hygiene (main.py:295-423, :955): non-finite scan, global-norm clip
 * (torch.nn.utils.clip_grad_norm_ semantics) and AdamW over up to 128 tensors in two launches, no host sync.
 *   acc = EIGHT doubles on the device: acc[0] = sum of squared gradients (before clipping), acc[1] = number of NaN/Inf
 *   gradient elements, acc[2] = step counter (below), acc[3] unused, acc[4..7] = one nsm_acc slot of scratch (the sum of
 *   squares is accumulated order-independently there and copied to acc[0]); when
 *   acc[1] != 0 the update is skipped entirely (GradScaler.step behaviour).  `step` > 0 is the 1-based AdamW step; with
 *   `step` <= 0 the counter lives on the device: acc[2] = number of updates applied so far (kept
 *   across calls, seeded by the caller), the bias corrections use acc[2] + 1 and acc[2] advances only when the update was
 *   applied -- so a skipped step does not run the bias corrections ahead of the moments (torch AdamW under GradScaler).
 * --------------------------------------------------------------------------------------------------------- */
int nsm_adamw_clip_step(int count, float* const* params, const float* const* grads, float* const* exp_avg,
                        float* const* exp_avg_sq, const long long* numel, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float max_norm, int step, double* acc /*[8]; device, 32-byte aligned*/,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NSM_B200_H_ */
