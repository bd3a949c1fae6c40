// Training-path kernels: train-mode BatchNorm (statistics, apply, backward), dropout / LeakyReLU fused into the BN passes,
// AvgPool / bilinear adjoints, network input / output stages, and the weight-gradient GEMM (wgrad_gemm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_gemm.cuh"

namespace nsm {

// ---- train-mode BatchNorm2d over NHWC planes z[P][C] ------------------------------------------------------------------
// sums[0..C) += sum_p z, sums[C..2C) += sum_p z^2   (fp64 accumulators, zeroed by the caller)
int bn_stats(const Planes& z, long long P, int C, int fmt, Acc* sums, cudaStream_t st);
// batch mean / biased var -> scale = gamma*invstd, shift = beta - mean*scale; saves mean & invstd; running stats update
// (momentum, unbiased variance) applied `updates` times (the reference's checkpoint re-runs conv5's BN, see oracle).
int bn_finalize(const Acc* sums, long long P, int C, const float* gamma, const float* beta, float eps,
                float momentum, int updates, float* running_mean, float* running_var, float* scale, float* shift,
                float* save_mean, float* save_invstd, cudaStream_t st);
// a = mask[n][c] * LeakyReLU(z*scale + shift) (+ residual); optional AvgPool2d(2) output
struct BnActParams {
  Planes z, out, residual, pool;
  int N, H, W, C, fmt;
  const float* scale;
  const float* shift;
  const float* mask;  // [N][C] Dropout2d keep/(1-p) factors or nullptr
  int lrelu;
};
int bn_act(const BnActParams& p, cudaStream_t st);

// ---- BatchNorm backward ---------------------------------------------------------------------------------------------------
// g = dy * mask[n][c] * LeakyReLU'(z*scale+shift);  sums[0..C) += sum g, sums[C..2C) += sum g * z
struct BnBwdParams {
  Planes dy, z, dz;
  int N, H, W, C, fmt;
  const float* scale;
  const float* shift;
  const float* mask;
  const float* mean;
  const float* invstd;
  int lrelu;
  Acc* sums;        // [2C] order-independent accumulators (reduce) ; apply reads them
  Acc* dbias;       // [C] sum of dz (apply), zeroed by the caller, or nullptr
};
int bn_bwd_reduce(const BnBwdParams& p, cudaStream_t st);
// dz = scale * (g - mean(g) - xhat * mean(g*xhat));  dgamma = sum g*xhat, dbeta = sum g written by bn_bwd_finalize
int bn_bwd_apply(const BnBwdParams& p, cudaStream_t st);
int bn_bwd_finalize(const Acc* sums, const Acc* dbias, const float* mean, const float* invstd, int C,
                    int round_bf16, float* dgamma, float* dbeta, float* dbias_out, cudaStream_t st);

// ---- adjoints of AvgPool2d(2) (+ skip-gradient add) and of one bilinear align_corners resize -------------------------
// out[n,y,x,c] = (a ? a : 0) + 0.25 * dpool[n,y/2,x/2,c] (inside the pooled footprint)
int pool_bwd_add(const Planes& a, const Planes& dpool, const Planes& out, int N, int H, int W, int C, int fmt,
                 cudaStream_t st);
// out = a + b  (planes), used to merge two gradient paths
int planes_add(const Planes& a, const Planes& b, const Planes& out, long long numel, int fmt, cudaStream_t st);
// din[n,hi,wi,c] = sum over outputs of the weights F.interpolate((ho,wo), bilinear, align_corners=True) used
int bilinear_bwd(const Planes& dout, int N, int ho, int wo, int C, const Planes& din, int hi, int wi, int fmt,
                 cudaStream_t st);

// adjoint of upsample_match (x2 up-sample then resize to (ho, wo)): dout [N,ho,wo,C] -> din [N,hi,wi,C], one pass
int upsample_match_bwd(const Planes& dout, int N, int ho, int wo, int C, const Planes& din, int hi, int wi, int fmt,
                       cudaStream_t st);

// ---- network input / output stages of the training path ----------------------------------------------------------------
// x [N,4,Hin,Win] fp32 -> (even-size fix) -> pixel_unshuffle(2) -> NHWC planes [N,h,w,64] (channels 16..63 zero)
int train_input_prep(const float* x, int N, int Hin, int Win, const Planes& out, int fmt, cudaStream_t st, int cpad = 64);
// adjoint for even Hin/Win: d x16 planes [N,h,w,64] -> dx [N,4,H,W] fp32
int train_input_grad(const Planes& dx16, int N, int H, int W, float* dx, int fmt, cudaStream_t st, int cpad = 64);
// c10 planes [N,h,w,64] (first 4 channels) -> sigmoid(pixel_shuffle) -> y [N,1,2h,2w] fp32
int sigmoid_shuffle_fwd(const Planes& c10, int N, int h, int w, int fmt, float* y, cudaStream_t st, int px4 = 0);
// dy [N,1,2h,2w], y -> d c10 planes [N,h,w,64] (channels 4..63 zero)
int sigmoid_shuffle_bwd(const float* dy, const float* y, int N, int h, int w, int fmt, const Planes& dc10,
                        cudaStream_t st, int px4 = 0);
// pixel-packed thin layers (train_kernels.cu): four horizontally adjacent pixels x C channels = one pixel of 4C virtual
// channels; weights [Cout][Cin][k][k] -> virtual conv [CoutV][tap][CinV] (block-diagonal / banded), and back for dW
int pack_conv_weight_px4(const float* w, int Cout, int Cin, int ksize, int CoutV, int CinV, int flip_transpose, int fmt,
                         void* hi, void* lo, cudaStream_t st);
int px4_reduce_dw(const float* dwv, int Cout, int Cin, int ksize, int CoutV, int CinV, float* dw, cudaStream_t st);
int fold_channel_sums(const Acc* in, int nvec, int CV, int groups, int C, Acc* out, cudaStream_t st);
int tile_vector(const float* src, int n, int rep, int npad, float fill, int round_bf16, float* dst, cudaStream_t st);
// zero-padded weight packing: OIHW [Cout][Cin][k][k] -> [CoutP][tap][CinP] planes (dgrad: [CinP][tap'][CoutP])
int pack_conv_weight_padded(const float* w, int Cout, int Cin, int ksize, int CoutP, int CinP, int flip_transpose,
                            int fmt, void* hi, void* lo, cudaStream_t st);
// dst[0..n) = src[0..n) (optionally bf16-rounded), dst[n..npad) = fill
int pad_vector(const float* src, int n, int npad, float fill, int round_bf16, float* dst, cudaStream_t st);

// ---- weight gradient: dW[co][tap][ci] = sum_pixels dz[p][co] * x[p + tap][ci] -----------------------------------------------
struct WgradShape {
  int N, H, W;
  int Cout, Cin;    // (padded) channel counts of dz and x, multiples of 64
  int taps;         // 1 or 9
  int fmt;
};
size_t wgrad_workspace_bytes(const WgradShape& s);
// dw: fp32 OIHW [Cout_real][Cin_real][k][k] (un-padded), optionally rounded to bf16 values
int wgrad_launch(const WgradShape& s, const Planes& dz, const Planes& x, void* workspace, size_t workspace_bytes,
                 int Cout_real, int Cin_real, int round_bf16, float* dw, cudaStream_t st);

}  // namespace nsm
