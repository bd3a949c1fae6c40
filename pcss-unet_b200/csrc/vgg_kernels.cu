// Streaming kernels of the perceptual term (MultiLayerVGGLoss, customLoss.py:7-90).  The VGG19 convolutions themselves run
// through the tcgen05 implicit-GEMM kernel (nsm_conv_fwd with a ReLU epilogue); these are the passes around them:
//   vgg_input_prep : clamp to [0,1] + NaN/Inf scrub (customLoss.py:44-52), grey -> 3 channels (:55-56), (v - 0.485) /
//                    (0.229 + 1e-8) (:59-61), NCHW fp32 -> NHWC planes zero-padded to 64 channels
//   relu_maxpool   : nn.ReLU [+ nn.MaxPool2d(2)] between a tapped (pre-activation) feature and the next convolution
//   feature_l1     : sum |nan_to_num(a) - nan_to_num(b)| of two feature tensors (customLoss.py:76-80), fp64 accumulator
// All HBM-bound: 16-byte accesses, one pass.
#include <stdio.h>

#include "../../include/nsm_b200.h"
#include "nsm_common.cuh"
#include "plane_io.cuh"

namespace nsm {

static inline int vgrid(long long work, int block, int cap = 148 * 16) {
  long long g = (work + block - 1) / block;
  return int(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float scrub01(float v) {   // torch.clamp(0,1) then nan_to_num(nan=0.5, posinf=1, neginf=0)
  if (v != v) return 0.5f;
  return fminf(fmaxf(v, 0.f), 1.f);
}

__global__ void __launch_bounds__(256) vgg_input_prep_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             long long per_tensor, Planes out, int fmt) {
  // pixel-major: thread = (pixel, group of 8 channels); only group 0 carries data (3 real channels)
  const long long total = 2 * per_tensor * 8;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long pix = i >> 3;
    const int cg = int(i & 7);
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cg == 0) {
      const float x = pix < per_tensor ? __ldg(a + pix) : __ldg(b + (pix - per_tensor));
      const float n = (scrub01(x) - 0.485f) / (0.229f + 1e-8f);
      v[0] = v[1] = v[2] = fmt == kFmtBf16 ? rbf(n) : n;
    }
    store8(out, (size_t)pix * 64 + cg * 8, fmt, v);
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) relu_maxpool_kernel(const Planes in, const Planes out, int N, int H, int W, int C,
                                                           int pool) {
  const int cgs = C / 8;
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const long long total = (long long)N * Ho * Wo * cgs;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = int(i % cgs);
    long long t = i / cgs;
    const int x = int(t % Wo); t /= Wo;
    const int y = int(t % Ho);
    const int n = int(t / Ho);
    float m[8];
    if (pool) {
      float q[4][8];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        load8(in, (((size_t)n * H + 2 * y + (k >> 1)) * W + 2 * x + (k & 1)) * C + cg * 8, FMT, q[k]);
#pragma unroll
      for (int e = 0; e < 8; ++e) m[e] = fmaxf(fmaxf(q[0][e], q[1][e]), fmaxf(q[2][e], q[3][e]));
    } else {
      load8(in, (((size_t)n * H + y) * W + x) * C + cg * 8, FMT, m);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], 0.f);   // ReLU (max-pool and ReLU commute; NaN -> 0 like fmax)
    store8(out, (((size_t)n * Ho + y) * Wo + x) * C + cg * 8, FMT, m);
  }
}

__device__ __forceinline__ float scrub_feat(float v) {   // nan_to_num(nan=0, posinf=1, neginf=-1), customLoss.py:76-77
  if (v != v) return 0.f;
  if (v - v != 0.f) return v > 0.f ? 1.f : -1.f;
  return v;
}

template <int FMT>
__global__ void __launch_bounds__(256) feature_l1_kernel(const Planes f, long long half8, Acc* acc) {
  // f holds 2B images; the first half are the features of `output`, the second half those of `target`
  float s = 0.f;
  double tot = 0.0;
  int run = 0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < half8; i += (long long)gridDim.x * 256) {
    float a[8], b[8];
    load8(f, (size_t)i * 8, FMT, a);
    load8(f, (size_t)(i + half8) * 8, FMT, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) s += fabsf(scrub_feat(a[e]) - scrub_feat(b[e]));
    if (++run == 32) { tot += double(s); s = 0.f; run = 0; }   // short fp32 runs, fp64 across them
  }
  tot += double(s);
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    acc_add(acc, t);
  }
}

}  // namespace nsm

using namespace nsm;

#define VGG_CHECK(name)                                                  \
  do {                                                                   \
    cudaError_t e__ = cudaGetLastError();                                \
    if (e__ != cudaSuccess) {                                            \
      set_error("%s launch failed: %s", name, cudaGetErrorString(e__));  \
      return 1;                                                          \
    }                                                                    \
    count_launch();                                                      \
  } while (0)

static int vgg_mode_ok(const char* what, int mode) {
  if (mode != kFmtBf16 && mode != kFmtF16x2 && mode != kFmtBf16x2) {
    set_error("%s: unsupported mode %d", what, mode);
    return 0;
  }
  return 1;
}

extern "C" int nsm_vgg_input_prep(const float* output, const float* target, int B, int H, int W, int mode, void* out0,
                                  void* out1, void* stream) {
  if (!vgg_mode_ok("nsm_vgg_input_prep", mode)) return 1;
  if (!output || !target || !out0 || B < 1 || H < 1 || W < 1) {
    set_error("nsm_vgg_input_prep: bad arguments");
    return 1;
  }
  Planes o{{out0, out1}};
  const long long per = (long long)B * H * W;
  vgg_input_prep_kernel<<<vgrid(2 * per * 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(output, target, per, o, mode);
  VGG_CHECK("vgg_input_prep");
  return 0;
}

extern "C" int nsm_relu_maxpool(const void* in0, const void* in1, int N, int H, int W, int C, int pool, int mode,
                                void* out0, void* out1, void* stream) {
  if (!vgg_mode_ok("nsm_relu_maxpool", mode)) return 1;
  if (C % 8 || N < 1 || H < (pool ? 2 : 1) || W < (pool ? 2 : 1)) {
    set_error("nsm_relu_maxpool: bad shape N=%d H=%d W=%d C=%d", N, H, W, C);
    return 1;
  }
  Planes i{{const_cast<void*>(in0), const_cast<void*>(in1)}}, o{{out0, out1}};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long work = (long long)N * (pool ? H / 2 : H) * (pool ? W / 2 : W) * (C / 8);
  if (mode == kFmtBf16) relu_maxpool_kernel<kFmtBf16><<<vgrid(work, 256), 256, 0, st>>>(i, o, N, H, W, C, pool);
  else if (mode == kFmtF16x2) relu_maxpool_kernel<kFmtF16x2><<<vgrid(work, 256), 256, 0, st>>>(i, o, N, H, W, C, pool);
  else relu_maxpool_kernel<kFmtBf16x2><<<vgrid(work, 256), 256, 0, st>>>(i, o, N, H, W, C, pool);
  VGG_CHECK("relu_maxpool");
  return 0;
}

extern "C" int nsm_feature_l1(const void* f0, const void* f1, long long numel_half, int mode, nsm_acc* acc, void* stream) {
  if (!vgg_mode_ok("nsm_feature_l1", mode)) return 1;
  if (numel_half % 8 || numel_half < 8 || !acc) {
    set_error("nsm_feature_l1: numel_half %lld must be a positive multiple of 8", numel_half);
    return 1;
  }
  Planes f{{const_cast<void*>(f0), const_cast<void*>(f1)}};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n8 = numel_half / 8;
  Acc* slot = reinterpret_cast<Acc*>(acc);
  if (mode == kFmtBf16) feature_l1_kernel<kFmtBf16><<<vgrid(n8, 256, 148 * 8), 256, 0, st>>>(f, n8, slot);
  else if (mode == kFmtF16x2) feature_l1_kernel<kFmtF16x2><<<vgrid(n8, 256, 148 * 8), 256, 0, st>>>(f, n8, slot);
  else feature_l1_kernel<kFmtBf16x2><<<vgrid(n8, 256, 148 * 8), 256, 0, st>>>(f, n8, slot);
  VGG_CHECK("feature_l1");
  return 0;
}
