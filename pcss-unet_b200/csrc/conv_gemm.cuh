// Host-visible description of one implicit-GEMM convolution launch (tcgen05 kernel in conv_gemm.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nsm {

struct Acc;   // order-independent accumulator slot (nsm_common.cuh) = nsm_acc of the C ABI

constexpr int kTileW = 16;   // spatial patch = kTileH x kTileW = 128 output pixels = UMMA M
constexpr int kTileH = 8;
constexpr int kKChunk = 64;  // bf16 elements per 128-byte swizzled smem row (one K block)

// Activation storage: NHWC "planes" of bf16.  bf16 mode: one plane.  fp32 mode: two planes (hi, lo) with
// value = hi + lo (|err| <= 2^-18 |v|); the GEMM then issues hi*hi + hi*lo + lo*hi on the bf16 tensor pipe.
struct Planes {
  void* p[2];
};

struct ConvEpilogue {
  const float* bias;     // [Cout] conv bias or nullptr
  const float* scale;    // [Cout] BatchNorm scale  (gamma * rsqrt(var + eps)) or nullptr
  const float* shift;    // [Cout] BatchNorm shift  (beta - mean * scale)
  int lrelu;             // activation after the affine: 0 none, 1 LeakyReLU(0.2), 2 ReLU
  int round_bf16;        // round to bf16 after conv, affine, activation, residual (autocast rounding points)
  Planes out;            // [N,H,W,Cout]
  Planes residual;       // [N,H,W,Cout] added after the activation (skip connection) or {nullptr}
  Planes pool;           // [N,H/2,W/2,Cout] AvgPool2d(2) of the output or {nullptr}
  float* out_f32;        // optional fp32 NHWC output (raw accumulators + bias), or nullptr
  Acc* stats;            // optional [2*Cout] accumulators: += per-channel sum and sum of squares of the stored conv+bias values
                         // (train-mode BatchNorm statistics fused into the producing convolution)
};

struct ConvShape {
  int N, H, W;     // batch and spatial size (stride 1, "same" padding => input == output size)
  int Cin, Cout;   // multiples of 64
  int taps;        // 1 (1x1) or 9 (3x3, pad 1)
  int fmt;         // storage format (nsm_common.cuh): 0 bf16, 1 fp16 hi+lo, 2 bf16 hi+lo
};

// in: NHWC planes [N,H,W,Cin]; w: packed [Cout][taps][Cin] bf16 planes.  Returns cudaError-like 0 on success.
int conv_gemm_launch(const ConvShape& s, const Planes& in, const Planes& w, const ConvEpilogue& ep,
                     cudaStream_t stream);

// resolved lazily through cudaGetDriverEntryPoint (no link-time dependency on libcuda)
int encode_tmap_tiled(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int elem_bytes, int swizzle_bytes = 128);

void set_error(const char* fmt, ...);
// number of kernels this library has launched (bench.py reports it as gpu_launches)
void count_launch(int n = 1);
long long launch_count();
void tmap_cache_stats(long long* hits, long long* misses);

}  // namespace nsm
