// HBM-streaming (non tensor-core) kernels of the hot path: head / tail stages of the U-Net, bilinear up-sampling,
// layout + weight packing, loss, channel statistics, standardisation, perturbation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_gemm.cuh"

namespace nsm {

// ---- parameter packing --------------------------------------------------------------------------------------
// OIHW fp32 [Cout][Cin][k][k] -> [Cout][tap][Cin] bf16 planes (hi[, lo]).  flip_transpose: dgrad form
// [Cin][tap'][Cout] with tap' = mirrored tap.
int pack_conv_weight(const float* w, int Cout, int Cin, int ksize, int flip_transpose, int fmt, void* hi, void* lo,
                     cudaStream_t st);
// eval-mode BatchNorm as an affine: scale = gamma / sqrt(var + eps), shift = beta - mean * scale
int bn_fold_eval(const float* gamma, const float* beta, const float* mean, const float* var, int C, float eps,
                 float* scale, float* shift, cudaStream_t st);
// dst[i] = round_bf16 ? bf16_rn(src[i]) : src[i]
int copy_round(const float* src, float* dst, int n, int round_bf16, cudaStream_t st);

// ---- layout conversion (tests, debugging taps) --------------------------------------------------------------
int nchw_to_planes(const float* x, int N, int C, int H, int W, int fmt, void* hi, void* lo, cudaStream_t st);
int planes_to_nchw(const void* hi, const void* lo, int N, int C, int H, int W, int fmt, float* y,
                   cudaStream_t st);

// ---- head: [standardise] + even-size fix + pixel_unshuffle(2) + conv2 DoubleConv (16->16 3x3, 16->64 1x1) -----
struct HeadParams {
  const float* x;   // [N,4,Hin,Win] fp32 NCHW
  int N, Hin, Win;  // raw input size; the stage works on H = Hin - Hin%2, W = Win - Win%2, h = H/2, w = W/2
  const float* mean;  // [4] or nullptr: fused (x - mean) / (std + 1e-8)
  const float* std;
  const void* img;  // kHeadImageBytes: conv2's weights / bias / BN affine as built by head_pack_image()
  int fmt;          // storage format: 0 bf16 (autocast rounding points), 1 fp16 hi+lo, 2 bf16 hi+lo
  Planes c2;        // [N,h,w,64]
  Planes p2;        // [N,h/2,w/2,64]
  Planes x16;       // optional tap of the un-shuffled input [N,h,w,16] (tests) or {nullptr}
};
int head_eval(const HeadParams& p, cudaStream_t st);
// The head's operand image: both convolutions' weights split into 16-bit planes in the kernel's swizzled shared-memory
// layout, followed by the per-channel vectors.  w0 [16][16][3][3], w1 [64][16] fp32 (pre-rounded to bf16 in bf16 mode).
constexpr int kHeadImageBytes = 2 * 9 * 16 * 32 + 2 * 64 * 32 + (3 * 16 + 3 * 64) * 4;
int head_pack_image(const float* w0, const float* b0, const float* s0, const float* t0, const float* w1,
                    const float* b1, const float* s1, const float* t1, int fmt, void* img, cudaStream_t st);

// ---- tail: conv9 1x1 (64->16) + BN + LReLU + conv10 (16->4) + sigmoid + pixel_shuffle(2) ----------------------
struct TailParams {
  Planes a;         // [N,h,w,64] activated output of conv9's 3x3 stage
  int N, h, w;
  const void* img;  // kTailImageBytes: conv9's 1x1 / conv10 operand image built by tail_pack_image()
  int fmt;
  float* y;          // [N,1,2h,2w] fp32 (or nullptr)
  uint8_t* y_u8;     // optional [N,1,2h,2w] uint8 = (uint8)(y * 255), the quantisation of infer.py:79 (or nullptr)
};
int tail_eval(const TailParams& p, cudaStream_t st);
// w1 [16][64], w10 [4][16] fp32 (pre-rounded to bf16 in bf16 mode) -> mma.sync B fragments in per-lane order + vectors
constexpr int kTailImageBytes = 2 * 4 * 32 * 16 + 2 * 32 * 8 + (3 * 16 + 4) * 4;
int tail_pack_image(const float* w1, const float* b1, const float* s1, const float* t1, const float* w10,
                    const float* b10, int fmt, void* img, cudaStream_t st);

// ---- nn.Upsample(x2, bilinear, align_corners) followed by F.interpolate(size=(hd, wd)) -----------------------
int upsample_match(const Planes& src, int N, int hs, int ws, int C, const Planes& dst, int hd, int wd, int fmt,
                   cudaStream_t st);

// ---- objective ---------------------------------------------------------------------------------------------------
// acc[0] += sum|o-t|, acc[1] += sum_i sum|o-y_i|, acc[2] += #(o<0 or o>1); grad = a_l1*sign(o-t)+a_p*sum_i sign(o-y_i)
int add_noise_clamp(const float* x, const float* noise, long long numel, float eps, float lo, float hi, float* out,
                    cudaStream_t st);
int mse_loss_fwd_bwd(const float* out, const float* ref, long long numel, float* diff, Acc* acc, cudaStream_t st);
int l1_loss_fwd_bwd(const float* out, const float* target, const float* const* perturbed, int n_perturbed,
                    long long numel, float coef_l1, float coef_pert, float* grad, Acc* acc, cudaStream_t st);

// ---- statistics / standardise / perturb ------------------------------------------------------------------------------
// x: [S][C][HW] fp32.  means == nullptr: sums[c] += sum x;  else sums[c] += sum (x - means[c])^2   (fp64)
int channel_sums(const float* x, long long S, int C, long long HW, const double* means, Acc* sums,
                 cudaStream_t st);
int standardize(const float* x, float* y, long long S, int C, long long HW, const float* mean, const float* std,
                cudaStream_t st);
// out[i][b][c][hw] = x[b][c][hw] + (noise[i][c][b][hw] * stds[c]) * std_factor  (noise: [count][C][B][HW])
int perturb(const float* x, const float* noise, float* out, int count, long long B, int C, long long HW,
            const float* stds, float std_factor, cudaStream_t st);

}  // namespace nsm
