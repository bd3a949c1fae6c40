// Optimizer + gradient hygiene step: what main.py:295-423 does with ~66 x (5..8) tiny kernels and host
// syncs -- NaN/Inf scan, global-norm clip (torch.nn.utils.clip_grad_norm_, main.py:405) and AdamW (main.py:955) -- as two
// multi-tensor HBM-streaming launches without any host synchronisation:
//   grad_norm_kernel : acc[0] += sum g^2 (fp64), acc[1] += #non-finite gradient elements
//   adamw_kernel     : if acc[1] == 0:  g *= min(1, max_norm / (sqrt(acc[0]) + 1e-6));  decoupled weight decay; Adam
//                      moments; bias-corrected update (same formula as torch.optim.AdamW)
// Tensor pointers travel in the kernel parameter block (up to 128 tensors), so no device-side table is needed.
#include <stdio.h>

#include "../../include/nsm_b200.h"
#include "conv_gemm.cuh"
#include "nsm_common.cuh"

namespace nsm {

constexpr int kMaxTensors = 128;
constexpr int kChunk = 256 * 4 * 16;   // elements per block: 256 threads x float4 x 16 iterations

struct MultiTensor {
  float* p[kMaxTensors];
  const float* g[kMaxTensors];
  float* m[kMaxTensors];
  float* v[kMaxTensors];
  int block_start[kMaxTensors + 1];   // first block of every tensor (prefix sum of ceil(numel / kChunk))
  long long numel[kMaxTensors];
  int count;
};

__device__ __forceinline__ int find_tensor(const MultiTensor& t, int block) {
  int lo = 0, hi = t.count - 1;
  while (lo < hi) {   // last tensor with block_start <= block
    const int mid = (lo + hi + 1) >> 1;
    if (t.block_start[mid] <= block) lo = mid;
    else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) grad_norm_kernel(const __grid_constant__ MultiTensor t, double* acc) {
  const int ti = find_tensor(t, blockIdx.x);
  const long long n = t.numel[ti];
  const long long base = (long long)(blockIdx.x - t.block_start[ti]) * kChunk;
  const float* g = t.g[ti];
  float ss = 0.f, bad = 0.f;
  const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
  for (int it = 0; it < 16; ++it) {
    const long long i = base + ((long long)it * 256 + threadIdx.x) * 4;
    if (i >= n) break;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (vec && i + 3 < n) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(g + i));
      x[0] = q.x; x[1] = q.y; x[2] = q.z; x[3] = q.w;
    } else {
      for (int e = 0; e < 4 && i + e < n; ++e) x[e] = g[i + e];
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      ss = fmaf(x[e], x[e], ss);
      bad += (x[e] - x[e] != 0.f) ? 1.f : 0.f;   // NaN or +-Inf
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
    bad += __shfl_xor_sync(0xffffffffu, bad, o);
  }
  __shared__ float r0[8], r1[8];
  if ((threadIdx.x & 31) == 0) {
    r0[threadIdx.x >> 5] = ss;
    r1[threadIdx.x >> 5] = bad;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int k = 0; k < 8; ++k) {
      a += double(r0[k]);
      b += double(r1[k]);
    }
    acc_add(reinterpret_cast<Acc*>(acc + 4), a);   // order-independent: the clip coefficient is reproducible run to run
    if (b != 0.0) atomicAdd(&acc[1], b);           // a count: integer-valued doubles add exactly in any order
  }
}

struct AdamWArgs {
  float lr, beta1, beta2, eps, weight_decay, max_norm;
  float bias_c1, bias_c2_sqrt;   // 1 - beta1^step, sqrt(1 - beta2^step)
  int device_step;               // step <= 0 at the ABI: the 1-based step is acc[2] + 1 (applied updates so far)
};

__global__ void __launch_bounds__(256) adamw_kernel(const __grid_constant__ MultiTensor t, const AdamWArgs a,
                                                    const double* acc) {
  if (acc[1] != 0.0) return;   // non-finite gradients: skip the update (what GradScaler.step does, main.py:421)
  float bias_c1 = a.bias_c1, bias_c2_sqrt = a.bias_c2_sqrt;
  if (a.device_step) {         // step counter lives in acc[2] and only counts APPLIED updates (torch AdamW under GradScaler)
    const float step = float(acc[2]) + 1.f;
    bias_c1 = 1.f - powf(a.beta1, step);
    bias_c2_sqrt = sqrtf(1.f - powf(a.beta2, step));
  }
  const float norm = float(sqrt(acc_load(reinterpret_cast<const Acc*>(acc + 4))));
  const float clip = a.max_norm > 0.f ? fminf(1.f, a.max_norm / (norm + 1e-6f)) : 1.f;
  const int ti = find_tensor(t, blockIdx.x);
  const long long n = t.numel[ti];
  const long long base = (long long)(blockIdx.x - t.block_start[ti]) * kChunk;
  float* p = t.p[ti];
  const float* g = t.g[ti];
  float* m = t.m[ti];
  float* v = t.v[ti];
  const float step_size = a.lr / bias_c1;
  for (int it = 0; it < 16; ++it) {
    const long long i0 = base + ((long long)it * 256 + threadIdx.x) * 4;
    if (i0 >= n) break;
    for (int e = 0; e < 4 && i0 + e < n; ++e) {   // (scalar accesses of one thread are contiguous: 16 B per array)
      const long long i = i0 + e;
      const float gg = g[i] * clip;
      float pp = p[i] * (1.f - a.lr * a.weight_decay);
      const float mm = a.beta1 * m[i] + (1.f - a.beta1) * gg;
      const float vv = a.beta2 * v[i] + (1.f - a.beta2) * gg * gg;
      pp -= step_size * (mm / (sqrtf(vv) / bias_c2_sqrt + a.eps));
      p[i] = pp;
      m[i] = mm;
      v[i] = vv;
    }
  }
}

__global__ void adamw_finish_kernel(double* acc, int device_step) {
  acc[0] = acc_load(reinterpret_cast<const Acc*>(acc + 4));   // sum of squared gradients, for the host
  if (device_step && acc[1] == 0.0) acc[2] += 1.0;   // after every block of adamw_kernel has read acc[2] (stream order)
}

}  // namespace nsm

using namespace nsm;

extern "C" int nsm_adamw_clip_step(int count, float* const* params, const float* const* grads, float* const* exp_avg,
                                   float* const* exp_avg_sq, const long long* numel, float lr, float beta1, float beta2,
                                   float eps, float weight_decay, float max_norm, int step, double* acc, void* stream) {
  if (count < 1 || count > kMaxTensors) {
    set_error("nsm_adamw_clip_step: %d tensors (1..%d supported per call)", count, kMaxTensors);
    return 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MultiTensor t;
  int blocks = 0;
  for (int i = 0; i < count; ++i) {
    t.p[i] = params[i]; t.g[i] = grads[i]; t.m[i] = exp_avg[i]; t.v[i] = exp_avg_sq[i];
    t.numel[i] = numel[i];
    t.block_start[i] = blocks;
    blocks += int((numel[i] + kChunk - 1) / kChunk);
  }
  t.block_start[count] = blocks;
  t.count = count;
  cudaError_t e = cudaMemsetAsync(acc, 0, 2 * sizeof(double), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(acc + 4, 0, sizeof(Acc), st);
  if (e != cudaSuccess) {
    set_error("nsm_adamw_clip_step: memset: %s", cudaGetErrorString(e));
    return 1;
  }
  grad_norm_kernel<<<blocks, 256, 0, st>>>(t, acc);
  AdamWArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay; a.max_norm = max_norm;
  a.device_step = step <= 0 ? 1 : 0;
  a.bias_c1 = 1.f - powf(beta1, float(step > 0 ? step : 1));
  a.bias_c2_sqrt = sqrtf(1.f - powf(beta2, float(step > 0 ? step : 1)));
  adamw_kernel<<<blocks, 256, 0, st>>>(t, a, acc);
  adamw_finish_kernel<<<1, 1, 0, st>>>(acc, a.device_step);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("nsm_adamw_clip_step launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  count_launch(3);
  return 0;
}
