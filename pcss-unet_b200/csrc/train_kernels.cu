// Streaming kernels of the training path (see train_kernels.cuh).  NHWC planes, 8 channels (16 bytes) per thread,
// per-channel reductions: fp32 in registers over short runs -> fp64 per thread -> shared memory (fixed order) -> one add per
// channel and block into an order-independent accumulator slot (nsm_common.cuh: Acc), so results are bit-reproducible.
#include <stdio.h>
#include <stdlib.h>

#include "nsm_common.cuh"
#include "plane_io.cuh"
#include "resample.cuh"
#include "train_kernels.cuh"

namespace nsm {

#define NSM_CHECK_LAUNCH(name)                                             \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      set_error("%s launch failed: %s", name, cudaGetErrorString(e__));    \
      return 1;                                                            \
    }                                                                      \
    count_launch();                                                        \
  } while (0)

static inline int grid_for(long long work, int block, int cap = 148 * 16) {
  long long g = (work + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return int(g);
}

constexpr int kBatch = 4;   // independent pixels in flight per thread (memory-level parallelism)

// Block-level per-channel reduction of NV value sets; thread t owns channel group (t % groups), 256 threads.
template <int NV, typename T>
__device__ __forceinline__ void block_channel_reduce(const T (&acc)[NV][8], int C, Acc* out) {
  __shared__ double red[256][NV * 8 + 1];
  const int tid = threadIdx.x, groups = C / 8, lanes = 256 / groups;
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) red[tid][v * 8 + e] = double(acc[v][e]);
  __syncthreads();
  for (int i = tid; i < NV * C; i += 256) {
    const int v = i / C, c = i % C;
    double s = 0.0;
    for (int l = 0; l < lanes; ++l) s += red[l * groups + (c >> 3)][v * 8 + (c & 7)];
    if (s != 0.0) acc_add(&out[v * C + c], s);   // order-independent (nsm_common.cuh: Acc)
  }
}

// ------------------------------------------------------------------------------------------------
// BatchNorm statistics
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) bn_stats_kernel(const Planes z, long long P, int C, int fmt, Acc* sums) {
  const int groups = C / 8, lanes = 256 / groups;
  const int cg = threadIdx.x % groups, lane = threadIdx.x / groups;
  double acc[2][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[0][e] = acc[1][e] = 0.0;
  const long long stride = (long long)gridDim.x * lanes;
  long long p = (long long)blockIdx.x * lanes + lane;
  while (p < P) {
    float fs[8], fq[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) fs[e] = fq[e] = 0.f;
    for (int it = 0; it < 4 && p < P; ++it, p += kBatch * stride) {
      Raw8 r[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u)
        if (p + u * stride < P) r[u] = load_raw8(z, (size_t)(p + u * stride) * C + cg * 8, FMT);
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        if (p + u * stride < P) {
          float v[8];
          unpack8(r[u], FMT, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            fs[e] += v[e];
            fq[e] = fmaf(v[e], v[e], fq[e]);
          }
        }
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += double(fs[e]);
      acc[1][e] += double(fq[e]);
    }
  }
  block_channel_reduce<2, double>(acc, C, sums);
}

static int check_32(const char* who, long long total) {
  if (total >= (1LL << 32)) {
    set_error("%s: %lld work items exceed the 32-bit index range", who, total);
    return 1;
  }
  return 0;
}

static int check_c(const char* who, int C) {
  if (C % 8 || C < 8 || (256 % (C / 8)) || C > 2048) {
    set_error("%s: unsupported channel count %d", who, C);
    return 1;
  }
  return 0;
}

int bn_stats(const Planes& z, long long P, int C, int fmt, Acc* sums, cudaStream_t st) {
  if (check_c("bn_stats", C)) return 1;
  const int lanes = 256 / (C / 8);
  {
    if (fmt == kFmtBf16) bn_stats_kernel<kFmtBf16><<<grid_for((P + lanes * 16 - 1) / (lanes * 16), 1, 148 * 8), 256, 0, st>>>(z, P, C, fmt, sums);
    else if (fmt == kFmtF16x2) bn_stats_kernel<kFmtF16x2><<<grid_for((P + lanes * 16 - 1) / (lanes * 16), 1, 148 * 8), 256, 0, st>>>(z, P, C, fmt, sums);
    else bn_stats_kernel<kFmtBf16x2><<<grid_for((P + lanes * 16 - 1) / (lanes * 16), 1, 148 * 8), 256, 0, st>>>(z, P, C, fmt, sums);
  }
  NSM_CHECK_LAUNCH("bn_stats");
  return 0;
}

__global__ void bn_finalize_kernel(const Acc* sums, long long P, int C, const float* gamma, const float* beta,
                                   float eps, float momentum, int updates, float* running_mean, float* running_var,
                                   float* scale, float* shift, float* save_mean, float* save_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = double(P);
  const double mean = acc_load(sums + c) / n;
  double var = acc_load(sums + C + c) / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = 1.0f / sqrtf(float(var) + eps);
  const float s = gamma[c] * invstd;
  scale[c] = s;
  shift[c] = beta[c] - float(mean) * s;
  if (save_mean) save_mean[c] = float(mean);
  if (save_invstd) save_invstd[c] = invstd;
  if (running_mean) {
    const float unbiased = float(P > 1 ? var * n / (n - 1.0) : var);
    float rm = running_mean[c], rv = running_var[c];
    for (int u = 0; u < updates; ++u) {
      rm = (1.f - momentum) * rm + momentum * float(mean);
      rv = (1.f - momentum) * rv + momentum * unbiased;
    }
    running_mean[c] = rm;
    running_var[c] = rv;
  }
}

int bn_finalize(const Acc* sums, long long P, int C, const float* gamma, const float* beta, float eps,
                float momentum, int updates, float* running_mean, float* running_var, float* scale, float* shift,
                float* save_mean, float* save_invstd, cudaStream_t st) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, P, C, gamma, beta, eps, momentum, updates, running_mean,
                                                      running_var, scale, shift, save_mean, save_invstd);
  NSM_CHECK_LAUNCH("bn_finalize");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// BN apply + LeakyReLU + Dropout2d mask (+ residual, + AvgPool2d(2))
// ------------------------------------------------------------------------------------------------
struct Chan8 {
  float s[8], t[8];   // BN scale / shift of the thread's 8 channels
};
__device__ __forceinline__ Chan8 load_chan8(const float* scale, const float* shift, int c0) {
  Chan8 c;
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c0)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c0 + 4));
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(shift + c0)), t1 = __ldg(reinterpret_cast<const float4*>(shift + c0 + 4));
  c.s[0] = s0.x; c.s[1] = s0.y; c.s[2] = s0.z; c.s[3] = s0.w; c.s[4] = s1.x; c.s[5] = s1.y; c.s[6] = s1.z; c.s[7] = s1.w;
  c.t[0] = t0.x; c.t[1] = t0.y; c.t[2] = t0.z; c.t[3] = t0.w; c.t[4] = t1.x; c.t[5] = t1.y; c.t[6] = t1.z; c.t[7] = t1.w;
  return c;
}
__device__ __forceinline__ void load_mask8(const float* mask, size_t off, float* m) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(mask + off)), b = __ldg(reinterpret_cast<const float4*>(mask + off + 4));
  m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w; m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
}

__device__ __forceinline__ void bn_act8(float* v, const BnActParams& p, const Chan8& ch, int n, int c0, bool rb) {
  float m[8];
  if (p.mask) load_mask8(p.mask, (size_t)n * p.C + c0, m);
  // (bf16 roundings pairwise: one packed conversion per two values instead of one quarter-rate F2F each)
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], ch.s[e], ch.t[e]);
  if (rb) {
#pragma unroll
    for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
  }
  if (p.lrelu) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = lrelu02(v[e]);
    if (rb) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
    }
  }
  if (p.mask) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] *= m[e];
    if (rb) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(256, 3) bn_act_kernel(const BnActParams p) {
  const int cgs = p.C / 8;
  const bool rb = FMT == kFmtBf16;
  const long long total = (long long)p.N * p.H * p.W * cgs;
  const int cg_shift = __ffs(cgs) - 1;                 // channel-group counts are powers of two
  const unsigned HW = (unsigned)(p.H * p.W);
  // the grid stride (gridDim.x * 256) is a multiple of cgs, so a thread's channel group never changes
  const Chan8 ch = load_chan8(p.scale, p.shift, int((blockIdx.x * 256u + threadIdx.x) & (unsigned)(cgs - 1)) * 8);
  // kBatch independent 16-byte loads (x2 with a skip tensor) per thread and iteration: the pass is latency-bound otherwise
  // (ncu: 47-53 % of DRAM peak with one load in flight per thread)
  const bool has_res = p.residual.p[0] != nullptr;
  const long long step = (long long)gridDim.x * 256;
  for (long long i0 = blockIdx.x * 256LL + threadIdx.x; i0 < total; i0 += kBatch * step) {
    Raw8 rz[kBatch], rr[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const long long i = i0 + u * step;
      if (i < total) {
        rz[u] = load_raw8(p.z, (size_t)i * 8, FMT);    // element offset (pix * C + cg * 8) == i * 8
        if (has_res) rr[u] = load_raw8(p.residual, (size_t)i * 8, FMT);
      }
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const long long i = i0 + u * step;
      if (i >= total) break;
      const unsigned iu = (unsigned)i;                 // total < 2^32 (host check): 32-bit index math only
      const int cg = int(iu & (unsigned)(cgs - 1));
      const unsigned pix = iu >> cg_shift;
      const int n = int(pix / HW);
      float v[8];
      unpack8(rz[u], FMT, v);
      bn_act8(v, p, ch, n, cg * 8, rb);
      if (has_res) {
        float r[8];
        unpack8(rr[u], FMT, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += r[e];
        if (rb) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
        }
      }
      store8(p.out, (size_t)i * 8, FMT, v);
    }
  }
}

// quad variant: one thread = 2x2 pixels x 8 channels, writes the activation and its 2x2 average
template <int FMT>
__global__ void __launch_bounds__(256) bn_act_pool_kernel(const BnActParams p) {
  const int cgs = p.C / 8;
  const bool rb = FMT == kFmtBf16;
  const int Hq = (p.H + 1) / 2, Wq = (p.W + 1) / 2, Hp = p.H / 2, Wp = p.W / 2;
  const long long total = (long long)p.N * Hq * Wq * cgs;
  const Chan8 ch = load_chan8(p.scale, p.shift, int((blockIdx.x * 256u + threadIdx.x) % (unsigned)cgs) * 8);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const unsigned iu = (unsigned)i;
    const int cg = int(iu % (unsigned)cgs);
    unsigned t = iu / (unsigned)cgs;
    const unsigned t1 = t / (unsigned)Wq;
    const int qx = int(t - t1 * (unsigned)Wq);
    const int n = int(t1 / (unsigned)Hq);
    const int qy = int(t1 - (unsigned)n * (unsigned)Hq);
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    Raw8 rq[4];   // the quad's four loads are issued before any of them is used
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = 2 * qy + (k >> 1), x = 2 * qx + (k & 1);
      if (y < p.H && x < p.W) rq[k] = load_raw8(p.z, (((size_t)n * p.H + y) * p.W + x) * p.C + cg * 8, FMT);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int y = 2 * qy + (k >> 1), x = 2 * qx + (k & 1);
      if (y < p.H && x < p.W) {
        const size_t pix = ((size_t)n * p.H + y) * p.W + x;
        float v[8];
        unpack8(rq[k], FMT, v);
        bn_act8(v, p, ch, n, cg * 8, rb);
        store8(p.out, pix * p.C + cg * 8, FMT, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += v[e];
      }
    }
    if (qy < Hp && qx < Wp) {
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] *= 0.25f;
      if (rb) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) rbf2(s[e], s[e + 1]);
      }
      store8(p.pool, (((size_t)n * Hp + qy) * Wp + qx) * p.C + cg * 8, FMT, s);
    }
  }
}

// TMA-staged variants (defined below, next to the BatchNorm backward that introduced the ring)
template <int NPL>
static int bn_act_stream_launch(const BnActParams& p, cudaStream_t st);
static bool sb_stream_ok(int N, int H, int W, int C);

int bn_act(const BnActParams& p, cudaStream_t st) {
  if (check_c("bn_act", p.C)) return 1;
  if (p.pool.p[0]) {
    if (p.residual.p[0]) {
      set_error("bn_act: pool and residual are not combined in this network");
      return 1;
    }
    const long long total = (long long)p.N * ((p.H + 1) / 2) * ((p.W + 1) / 2) * (p.C / 8);
    {
      if (p.fmt == kFmtBf16) bn_act_pool_kernel<kFmtBf16><<<grid_for(total, 256), 256, 0, st>>>(p);
      else if (p.fmt == kFmtF16x2) bn_act_pool_kernel<kFmtF16x2><<<grid_for(total, 256), 256, 0, st>>>(p);
      else bn_act_pool_kernel<kFmtBf16x2><<<grid_for(total, 256), 256, 0, st>>>(p);
    }
  } else if (p.fmt == kFmtBf16x2 && sb_stream_ok(p.N, p.H, p.W, p.C)) {
    // hi+lo planes: TMA-staged (cfg2 sizes: 3.9 -> 6.2 TB/s).  The single-plane pass is bound by its instructions (three bf16
    // rounding points per element), not by bytes in flight: staged it runs at 4.77 TB/s, one or two CTAs per SM alike,
    // against 4.95 TB/s for the register-staged kernel below, which therefore stays (tools/bn_probe.py).
    if (bn_act_stream_launch<2>(p, st)) return 1;
  } else {
    const long long total = (long long)p.N * p.H * p.W * (p.C / 8);
    {
      if (p.fmt == kFmtBf16) bn_act_kernel<kFmtBf16><<<grid_for(total, 256), 256, 0, st>>>(p);
      else if (p.fmt == kFmtF16x2) bn_act_kernel<kFmtF16x2><<<grid_for(total, 256), 256, 0, st>>>(p);
      else bn_act_kernel<kFmtBf16x2><<<grid_for(total, 256), 256, 0, st>>>(p);
    }
  }
  NSM_CHECK_LAUNCH("bn_act");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm backward
// ------------------------------------------------------------------------------------------------
// g = dy * mask * LeakyReLU'(z*scale + shift), with the bf16 rounding points of autograd under autocast
// Lean formulation (the kernels are instruction-bound, not byte-bound):
//   g    = dy * mask * (z*s + t > 0 ? 1 : 0.2)                 (bf16 rounding points of autograd kept in bf16 mode)
//   pass 1 accumulates S1 = sum g and S2 = sum g*z              (sum g*xhat = (S2 - mean*S1) * invstd)
//   pass 2 writes dz = s*g + A*z + B with per-channel A = -s*invstd*m2, B = -s*m1 + s*invstd*m2*mean,
//          m1 = S1/P, m2 = sum g*xhat / P                       (= s * (g - m1 - xhat*m2))
// Training tensors always hold bf16 elements (fmt 0: one plane, fmt 2: hi+lo planes) -> NPL template, shift unpack.
template <int NPL>
__device__ __forceinline__ void unpack8_bf16(const Raw8& r, float* v) {
  const uint32_t hw[4] = {r.h.x, r.h.y, r.h.z, r.h.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = bf16lo_to_f32(hw[e]);
    v[2 * e + 1] = bf16hi_to_f32(hw[e]);
  }
  if (NPL == 2) {
    const uint32_t lw[4] = {r.l.x, r.l.y, r.l.z, r.l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] += bf16lo_to_f32(lw[e]);
      v[2 * e + 1] += bf16hi_to_f32(lw[e]);
    }
  }
}
template <int NPL>
__device__ __forceinline__ Raw8 load_raw8_t(const Planes& p, size_t elem) {
  Raw8 r;
  r.h = ldg16(reinterpret_cast<const uint8_t*>(p.p[0]) + elem * 2);
  if (NPL == 2) r.l = ldg16(reinterpret_cast<const uint8_t*>(p.p[1]) + elem * 2);
  return r;
}

struct BwdChan8 {
  float s[8], t[8];
};

// g = dy * mask * LeakyReLU'(z*s + t).  The products are formed in fp32 and rounded once, when dz is stored: gradients are
// held to a relative L2 bound (1e-2 in bf16 mode), not to autograd's intermediate bf16 rounding points, and every extra
// convert costs issue slots in a pass that is instruction-limited (ncu: 60 instructions per 16-byte load).
template <int NPL>
__device__ __forceinline__ void bn_bwd_g8(const BnBwdParams& p, const BwdChan8& ch, const Raw8& rdy, const Raw8& rz, int n,
                                          int c0, float* g, float* z) {
  float dy[8];
  unpack8_bf16<NPL>(rdy, dy);
  unpack8_bf16<NPL>(rz, z);
  if (p.mask) {
    float m[8];
    load_mask8(p.mask, (size_t)n * p.C + c0, m);
#pragma unroll
    for (int e = 0; e < 8; ++e) dy[e] *= m[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool pos = !p.lrelu || fmaf(z[e], ch.s[e], ch.t[e]) > 0.f;   // the sign of y survives its bf16 rounding
    g[e] = pos ? dy[e] : 0.2f * dy[e];
  }
}

// One kernel for both passes (APPLY = false: the two sums; true: dz and the conv-bias gradient).  A thread owns one
// 8-channel group and walks pixels with a grid stride; PB pixels = 2*PB independent 16-byte loads are in flight per thread
// (ncu of the first version: 128 registers, 24 % occupancy, 40-46 % of DRAM peak -- latency-bound).  Per-thread runs are
// short (P / (grid * lanes) pixels), so the per-thread accumulators are fp32; the cross-thread reduction is fp64.
template <int NPL, bool APPLY>
__global__ void __launch_bounds__(256, 2) bn_bwd_kernel(const BnBwdParams p) {
  constexpr bool rb = NPL == 1;
  constexpr int PB = NPL == 1 ? 4 : 2;
  const int groups = p.C / 8, lanes = 256 / groups;
  const int cg = threadIdx.x % groups, lane = threadIdx.x / groups;
  const long long P = (long long)p.N * p.H * p.W;
  const unsigned HW = (unsigned)(p.H * p.W);
  BwdChan8 ch;
  float A[8], B[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = cg * 8 + e;
    ch.s[e] = __ldg(p.scale + c);
    ch.t[e] = __ldg(p.shift + c);
    if (APPLY) {
      const double mu = double(__ldg(p.mean + c)), is = double(__ldg(p.invstd + c));
      const double s1 = acc_load(p.sums + c), s2 = acc_load(p.sums + p.C + c);
      const double m1 = s1 / double(P);
      const double m2 = (s2 - mu * s1) * is / double(P);
      A[e] = float(-double(ch.s[e]) * is * m2);
      B[e] = float(-double(ch.s[e]) * m1 + double(ch.s[e]) * is * m2 * mu);
    }
  }
  float acc[APPLY ? 1 : 2][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[0][e] = 0.f;
    if (!APPLY) acc[1][e] = 0.f;
  }
  const long long stride = (long long)gridDim.x * lanes;
  for (long long px = (long long)blockIdx.x * lanes + lane; px < P; px += PB * stride) {
    Raw8 rdy[PB], rz[PB];
#pragma unroll
    for (int u = 0; u < PB; ++u) {
      const long long q = px + u * stride;
      if (q < P) {
        rdy[u] = load_raw8_t<NPL>(p.dy, (size_t)q * p.C + cg * 8);
        rz[u] = load_raw8_t<NPL>(p.z, (size_t)q * p.C + cg * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < PB; ++u) {
      const long long q = px + u * stride;
      if (q >= P) break;
      float g[8], z[8];
      bn_bwd_g8<NPL>(p, ch, rdy[u], rz[u], int((unsigned)q / HW), cg * 8, g, z);
      if (APPLY) {
        float dz[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dz[e] = fmaf(ch.s[e], g[e], fmaf(A[e], z[e], B[e]));
        if (rb) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) rbf2(dz[e], dz[e + 1]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[0][e] += dz[e];
        store8(p.dz, (size_t)q * p.C + cg * 8, p.fmt, dz);
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          acc[0][e] += g[e];
          acc[1][e] = fmaf(g[e], z[e], acc[1][e]);
        }
      }
    }
  }
  if (!APPLY) block_channel_reduce<APPLY ? 1 : 2, float>(acc, p.C, p.sums);
  else if (p.dbias) block_channel_reduce<APPLY ? 1 : 2, float>(acc, p.C, p.dbias);
}

// ---- the same two passes with the operands STAGED THROUGH SHARED MEMORY BY THE TMA UNIT --------------------------------------
// ncu of the register-staged kernel above (profiles/r02_bn_stream_full_cfg2_bf16.txt): 108-122 registers, 24 % occupancy,
// 2.6-5.6 TB/s in the reduction and 2.9-4.4 TB/s in the apply pass depending on the tensor size -- bytes in flight are tied
// to registers x resident warps.  Here one persistent CTA per SM streams dy and z as contiguous 8 KB chunks (NHWC tensors are
// flat arrays of 16-byte channel groups) through a ring of kSbStages shared-memory stages filled by 1-D bulk copies
// (cp.async.bulk issued by thread 0, full / empty mbarriers): 96-192 KB per SM are in flight whatever the consumers do.
// Sixteen consumer warps take one 16-byte group per thread and chunk; a thread's channel group never changes (chunk and
// thread counts are multiples of the channel-group count), so the per-channel constants stay in registers.
constexpr int kSbConsumers = 512;                 // sixteen warps = four per scheduler at up to 128 registers; thread 0 also feeds the ring
constexpr int kSbThreads = kSbConsumers;
constexpr int kSbChunk = kSbConsumers * 16;       // bytes per plane and stage: one 16-byte group per consumer thread
constexpr int kSbStages = 6;
template <int NPL>
struct SbCfg {
  static constexpr int STAGE_BYTES = 2 * NPL * kSbChunk;   // dy planes, then z planes
  static constexpr int RED_BYTES = kSbConsumers * 17 * 8;   // final cross-thread reduction reuses the ring
  static constexpr int RING_BYTES = kSbStages * STAGE_BYTES > RED_BYTES ? kSbStages * STAGE_BYTES : RED_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + 2 * kSbStages * 8 + 128;
};

template <int NPL, bool APPLY>
__global__ void __launch_bounds__(kSbThreads, 1) bn_bwd_stream_kernel(const BnBwdParams p) {
  using Cfg = SbCfg<NPL>;
  extern __shared__ __align__(128) uint8_t sb_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sb_smem + Cfg::RING_BYTES);
  uint64_t* empty = full + kSbStages;
  const int tid = threadIdx.x, lane = tid & 31;
  const int groups = p.C / 8;
  const long long P = (long long)p.N * p.H * p.W;
  const long long plane_bytes = P * p.C * 2;
  const long long nchunks = (plane_bytes + kSbChunk - 1) / kSbChunk;
  if (tid == 0) {
    for (int s = 0; s < kSbStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kSbConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  constexpr int NV = APPLY ? 1 : 2;
  double acc64[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc64[v][e] = 0.0;

  // ring refill (thread 0): chunk number `k` of this CTA goes to stage k % kSbStages
  auto issue_chunk = [&](long long ci, int s) {
    const long long off = ci * kSbChunk;
    const uint32_t bytes = uint32_t(plane_bytes - off < kSbChunk ? plane_bytes - off : kSbChunk);
    mbar_expect_tx(&full[s], 2 * NPL * bytes);
    const uint32_t dst = smem_u32(sb_smem) + s * Cfg::STAGE_BYTES;
#pragma unroll
    for (int pl = 0; pl < NPL; ++pl) {
      bulk_load_1d(dst + pl * kSbChunk, reinterpret_cast<const uint8_t*>(p.dy.p[pl]) + off, bytes, &full[s]);
      bulk_load_1d(dst + (NPL + pl) * kSbChunk, reinterpret_cast<const uint8_t*>(p.z.p[pl]) + off, bytes, &full[s]);
    }
  };
  if (tid == 0) {
    long long ci = blockIdx.x;
    for (int s = 0; s < kSbStages && ci < nchunks; ++s, ci += gridDim.x) issue_chunk(ci, s);
  }
  {
    // ===================== consumers =====================
    // (instruction budget: the first version spent 173 / 236 instructions per 16-byte group -- 64-bit index divisions, a mask
    //  load per group, generic-format packing -- and was issue-bound at 60 % issue-active; everything per-sample is now
    //  carried incrementally and the Dropout2d mask is folded with the LeakyReLU slope into two factors per channel)
    constexpr bool rb = NPL == 1;
    const int cg = tid % groups;   // kSbConsumers and the groups per chunk are multiples of `groups`
    const int gs = __ffs(groups) - 1;
    const unsigned HW = (unsigned)(p.H * p.W);
    BwdChan8 ch;
    float A[8], B[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int c = cg * 8 + e;
      ch.s[e] = __ldg(p.scale + c);
      ch.t[e] = __ldg(p.shift + c);
      if (APPLY) {
        const double mu = double(__ldg(p.mean + c)), is = double(__ldg(p.invstd + c));
        const double s1 = acc_load(p.sums + c), s2 = acc_load(p.sums + p.C + c);
        const double m1 = s1 / double(P);
        const double m2 = (s2 - mu * s1) * is / double(P);
        A[e] = float(-double(ch.s[e]) * is * m2);
        B[e] = float(-double(ch.s[e]) * m1 + double(ch.s[e]) * is * m2 * mu);
      }
    }
    const unsigned ngrp = (unsigned)(plane_bytes >> 4);            // 16-byte groups per plane (< 2^32: host check)
    unsigned grp = blockIdx.x * (unsigned)kSbConsumers + tid;      // this thread's group in the current chunk
    const unsigned dgrp = gridDim.x * (unsigned)kSbConsumers;
    const unsigned dpix = dgrp >> gs;
    unsigned n = (grp >> gs) / HW, rem = (grp >> gs) - n * HW;     // sample index and pixel inside the sample
    // factor on dy where the activation passed / was on the leaky side: mask, 0.2 * mask (1, 0.2 without a mask)
    const float slope = p.lrelu ? 0.2f : 1.f;
    float fpos[8], fneg[8];
    auto load_factors = [&](unsigned nn) {
#pragma unroll
      for (int e = 0; e < 8; ++e) fpos[e] = 1.f;
      if (p.mask && nn < (unsigned)p.N) load_mask8(p.mask, (size_t)nn * p.C + cg * 8, fpos);
#pragma unroll
      for (int e = 0; e < 8; ++e) fneg[e] = slope * fpos[e];
    };
    load_factors(n);
    float acc[NV][8];
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[v][e] = 0.f;
    int run = 0, s = 0;
    uint32_t ph = 0;
    for (long long ci = blockIdx.x; ci < nchunks; ci += gridDim.x) {
      mbar_wait(&full[s], ph);
      if (grp < ngrp) {
        const uint32_t src = smem_u32(sb_smem) + s * Cfg::STAGE_BYTES + tid * 16;
        Raw8 rdy, rz;
        rdy.h = lds16(src);
        rz.h = lds16(src + NPL * kSbChunk);
        if (NPL == 2) {
          rdy.l = lds16(src + kSbChunk);
          rz.l = lds16(src + 3 * kSbChunk);
        }
        float g[8], z[8];
        unpack8_bf16<NPL>(rdy, g);
        unpack8_bf16<NPL>(rz, z);
#pragma unroll
        for (int e = 0; e < 8; ++e)   // (the sign of y survives its bf16 rounding)
          g[e] *= fmaf(z[e], ch.s[e], ch.t[e]) > 0.f ? fpos[e] : fneg[e];
        if (APPLY) {
          float dz[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) dz[e] = fmaf(ch.s[e], g[e], fmaf(A[e], z[e], B[e]));
          uint4 hv, lv;
          uint32_t* hw = &hv.x;
          uint32_t* lw = &lv.x;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            hw[e] = pack_bf16(dz[2 * e], dz[2 * e + 1]);
            const float h0 = bf16lo_to_f32(hw[e]), h1 = bf16hi_to_f32(hw[e]);
            if (rb) {          // bf16 mode: the bias gradient sums the STORED (rounded) values
              acc[0][2 * e] += h0;
              acc[0][2 * e + 1] += h1;
            } else {
              acc[0][2 * e] += dz[2 * e];
              acc[0][2 * e + 1] += dz[2 * e + 1];
              lw[e] = pack_bf16(dz[2 * e] - h0, dz[2 * e + 1] - h1);
            }
          }
          stg16(reinterpret_cast<uint8_t*>(p.dz.p[0]) + (size_t)grp * 16, hv);
          if (NPL == 2) stg16(reinterpret_cast<uint8_t*>(p.dz.p[1]) + (size_t)grp * 16, lv);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[0][e] += g[e];
            acc[NV - 1][e] = fmaf(g[e], z[e], acc[NV - 1][e]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);   // this warp has read its part of the stage
      if (tid == 0) {
        // refill the stage with the chunk kSbStages ahead once all sixteen warps have released it (this warp then trails the
        // others by at most one chunk; kSbStages - 1 chunks stay in flight)
        const long long nxt = ci + (long long)kSbStages * gridDim.x;
        if (nxt < nchunks) {
          mbar_wait(&empty[s], ph);
          issue_chunk(nxt, s);
        }
      }
      __syncwarp();
      if (++s == kSbStages) { s = 0; ph ^= 1; }
      grp += dgrp;
      rem += dpix;
      if (rem >= HW) {     // next sample(s): new Dropout2d mask row
        do { rem -= HW; ++n; } while (rem >= HW);
        if (p.mask) load_factors(n);
      }
      if (++run == 32) {   // short fp32 runs, fp64 across them
        run = 0;
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc64[v][e] += double(acc[v][e]);
            acc[v][e] = 0.f;
          }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) acc64[v][e] += double(acc[v][e]);
  }
  // ---- per-channel totals of the CTA: fixed summation order, then the order-independent accumulator slots
  Acc* out = APPLY ? p.dbias : p.sums;
  __syncthreads();   // every bulk copy has landed and has been read: the ring is free
  if (out == nullptr) return;
  double* red = reinterpret_cast<double*>(sb_smem);
  {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 8; ++e) red[tid * 17 + v * 8 + e] = acc64[v][e];
  }
  __syncthreads();
  const int lanes = kSbConsumers / groups;
  for (int i = tid; i < NV * p.C; i += kSbThreads) {
    const int v = i / p.C, c = i % p.C;
    double t = 0.0;
    for (int l = 0; l < lanes; ++l) t += red[(l * groups + (c >> 3)) * 17 + v * 8 + (c & 7)];
    if (t != 0.0) acc_add(&out[v * p.C + c], t);
  }
}

template <int NPL, bool APPLY>
static int bn_bwd_stream_launch(const BnBwdParams& p, cudaStream_t st) {
  using Cfg = SbCfg<NPL>;
  auto kern = bn_bwd_stream_kernel<NPL, APPLY>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("bn_bwd_stream: cudaFuncSetAttribute(%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return 1;
    }
    attr_set = true;
  }
  const long long plane_bytes = (long long)p.N * p.H * p.W * p.C * 2;
  const long long nchunks = (plane_bytes + kSbChunk - 1) / kSbChunk;
  const int grid = int(nchunks < 148 ? nchunks : 148);
  kern<<<grid, kSbThreads, Cfg::SMEM_BYTES, st>>>(p);
  return 0;
}
static bool sb_stream_ok(int N, int H, int W, int C) {
  static const bool off = getenv("NSM_BN_NO_STREAM") != nullptr;
  // (a thread's channel group must not change from chunk to chunk: the groups per chunk are a multiple of C / 8)
  return !off && kSbConsumers % (C / 8) == 0 && (long long)N * H * W * C / 8 < (1LL << 31);   // 32-bit group index
}
static bool bn_bwd_use_stream(const BnBwdParams& p) { return sb_stream_ok(p.N, p.H, p.W, p.C); }

// ---- BN apply + LeakyReLU + Dropout2d mask (+ skip add) on the same TMA-staged ring (no pooling: that variant walks 2x2 quads) ----
template <int NPL>
struct SaCfg {
  static constexpr int STAGE_BYTES = 2 * NPL * kSbChunk;   // z planes, then skip planes (unused without a skip tensor)
  static constexpr int SMEM_BYTES = kSbStages * STAGE_BYTES + 2 * kSbStages * 8 + 128;
};
template <int NPL>
__global__ void __launch_bounds__(kSbThreads, 1) bn_act_stream_kernel(const BnActParams p) {
  using Cfg = SaCfg<NPL>;
  extern __shared__ __align__(128) uint8_t sb_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sb_smem + kSbStages * Cfg::STAGE_BYTES);
  uint64_t* empty = full + kSbStages;
  const int tid = threadIdx.x, lane = tid & 31;
  const int groups = p.C / 8;
  const long long plane_bytes = (long long)p.N * p.H * p.W * p.C * 2;
  const long long nchunks = (plane_bytes + kSbChunk - 1) / kSbChunk;
  const bool has_res = p.residual.p[0] != nullptr;
  if (tid == 0) {
    for (int s = 0; s < kSbStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kSbConsumers / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  auto issue_chunk = [&](long long ci, int s) {
    const long long off = ci * kSbChunk;
    const uint32_t bytes = uint32_t(plane_bytes - off < kSbChunk ? plane_bytes - off : kSbChunk);
    mbar_expect_tx(&full[s], (has_res ? 2 : 1) * NPL * bytes);
    const uint32_t dst = smem_u32(sb_smem) + s * Cfg::STAGE_BYTES;
#pragma unroll
    for (int pl = 0; pl < NPL; ++pl) {
      bulk_load_1d(dst + pl * kSbChunk, reinterpret_cast<const uint8_t*>(p.z.p[pl]) + off, bytes, &full[s]);
      if (has_res)
        bulk_load_1d(dst + (NPL + pl) * kSbChunk, reinterpret_cast<const uint8_t*>(p.residual.p[pl]) + off, bytes, &full[s]);
    }
  };
  if (tid == 0) {
    long long ci = blockIdx.x;
    for (int s = 0; s < kSbStages && ci < nchunks; ++s, ci += gridDim.x) issue_chunk(ci, s);
  }
  constexpr bool rb = NPL == 1;
  const int cg = tid % groups;
  const int gs = __ffs(groups) - 1;
  const unsigned HW = (unsigned)(p.H * p.W);
  const Chan8 ch = load_chan8(p.scale, p.shift, cg * 8);
  const unsigned ngrp = (unsigned)(plane_bytes >> 4);
  unsigned grp = blockIdx.x * (unsigned)kSbConsumers + tid;
  const unsigned dgrp = gridDim.x * (unsigned)kSbConsumers;
  const unsigned dpix = dgrp >> gs;
  unsigned n = (grp >> gs) / HW, rem = (grp >> gs) - n * HW;
  float m[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) m[e] = 1.f;
  if (p.mask && n < (unsigned)p.N) load_mask8(p.mask, (size_t)n * p.C + cg * 8, m);
  int s = 0;
  uint32_t ph = 0;
  for (long long ci = blockIdx.x; ci < nchunks; ci += gridDim.x) {
    mbar_wait(&full[s], ph);
    if (grp < ngrp) {
      const uint32_t src = smem_u32(sb_smem) + s * Cfg::STAGE_BYTES + tid * 16;
      Raw8 rz, rr;
      rz.h = lds16(src);
      if (NPL == 2) rz.l = lds16(src + kSbChunk);
      if (has_res) {
        rr.h = lds16(src + NPL * kSbChunk);
        if (NPL == 2) rr.l = lds16(src + 3 * kSbChunk);
      }
      float v[8];
      unpack8_bf16<NPL>(rz, v);
      // same operation order and bf16 rounding points as bn_act8
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], ch.s[e], ch.t[e]);
      if (rb) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
      }
      if (p.lrelu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = lrelu02(v[e]);
        if (rb) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
        }
      }
      if (p.mask) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= m[e];
        if (rb) {
#pragma unroll
          for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
        }
      }
      if (has_res) {
        float r[8];
        unpack8_bf16<NPL>(rr, r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += r[e];
      }
      uint4 hv, lv;
      uint32_t* hw = &hv.x;
      uint32_t* lw = &lv.x;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hw[e] = pack_bf16(v[2 * e], v[2 * e + 1]);   // (bf16 mode: this IS the rounding after the skip add / last step)
        if (NPL == 2) lw[e] = pack_bf16(v[2 * e] - bf16lo_to_f32(hw[e]), v[2 * e + 1] - bf16hi_to_f32(hw[e]));
      }
      stg16(reinterpret_cast<uint8_t*>(p.out.p[0]) + (size_t)grp * 16, hv);
      if (NPL == 2) stg16(reinterpret_cast<uint8_t*>(p.out.p[1]) + (size_t)grp * 16, lv);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    if (tid == 0) {
      const long long nxt = ci + (long long)kSbStages * gridDim.x;
      if (nxt < nchunks) {
        mbar_wait(&empty[s], ph);
        issue_chunk(nxt, s);
      }
    }
    __syncwarp();
    if (++s == kSbStages) { s = 0; ph ^= 1; }
    grp += dgrp;
    rem += dpix;
    if (rem >= HW) {
      do { rem -= HW; ++n; } while (rem >= HW);
      if (p.mask && n < (unsigned)p.N) load_mask8(p.mask, (size_t)n * p.C + cg * 8, m);
    }
  }
}

template <int NPL>
static int bn_act_stream_launch(const BnActParams& p, cudaStream_t st) {
  using Cfg = SaCfg<NPL>;
  auto kern = bn_act_stream_kernel<NPL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("bn_act_stream: cudaFuncSetAttribute(%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return 1;
    }
    attr_set = true;
  }
  const long long plane_bytes = (long long)p.N * p.H * p.W * p.C * 2;
  const long long nchunks = (plane_bytes + kSbChunk - 1) / kSbChunk;
  kern<<<int(nchunks < 148 ? nchunks : 148), kSbThreads, Cfg::SMEM_BYTES, st>>>(p);
  return 0;
}

static int bn_bwd_grid(const BnBwdParams& p) {
  const long long P = (long long)p.N * p.H * p.W;
  const int lanes = 256 / (p.C / 8);
  // per-thread runs of <= 64 pixels keep the fp32 accumulators exact enough; at least 148 * 8 blocks when there is work
  long long want = (P + lanes * 64 - 1) / (lanes * 64);
  if (want < 148 * 8) want = (P + lanes * 8 - 1) / (lanes * 8) < 148 * 8 ? (P + lanes * 8 - 1) / (lanes * 8) : 148 * 8;
  if (want < 1) want = 1;
  if (want > 148 * 64) want = 148 * 64;
  return int(want);
}

int bn_bwd_reduce(const BnBwdParams& p, cudaStream_t st) {
  if (check_c("bn_bwd_reduce", p.C)) return 1;
  if (p.fmt == kFmtF16x2) {
    set_error("bn_bwd: training tensors use fmt 0 or 2");
    return 1;
  }
  if (bn_bwd_use_stream(p)) {
    if (p.fmt == kFmtBf16 ? bn_bwd_stream_launch<1, false>(p, st) : bn_bwd_stream_launch<2, false>(p, st)) return 1;
    NSM_CHECK_LAUNCH("bn_bwd_reduce");
    return 0;
  }
  const int grid = bn_bwd_grid(p);
  if (p.fmt == kFmtBf16) bn_bwd_kernel<1, false><<<grid, 256, 0, st>>>(p);
  else bn_bwd_kernel<2, false><<<grid, 256, 0, st>>>(p);
  NSM_CHECK_LAUNCH("bn_bwd_reduce");
  return 0;
}

int bn_bwd_apply(const BnBwdParams& p, cudaStream_t st) {
  if (check_c("bn_bwd_apply", p.C)) return 1;
  if (bn_bwd_use_stream(p)) {
    if (p.fmt == kFmtBf16 ? bn_bwd_stream_launch<1, true>(p, st) : bn_bwd_stream_launch<2, true>(p, st)) return 1;
    NSM_CHECK_LAUNCH("bn_bwd_apply");
    return 0;
  }
  const int grid = bn_bwd_grid(p);
  if (p.fmt == kFmtBf16) bn_bwd_kernel<1, true><<<grid, 256, 0, st>>>(p);
  else bn_bwd_kernel<2, true><<<grid, 256, 0, st>>>(p);
  NSM_CHECK_LAUNCH("bn_bwd_apply");
  return 0;
}

__global__ void bn_bwd_finalize_kernel(const Acc* sums, const Acc* dbias, const float* mean, const float* invstd,
                                       int C, int rb, float* dgamma, float* dbeta, float* dbias_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  // sums[c] = sum g, sums[C + c] = sum g*z  ->  dgamma = sum g*xhat, dbeta = sum g
  const double s1 = acc_load(sums + c), s2 = acc_load(sums + C + c);
  dgamma[c] = float((s2 - double(mean[c]) * s1) * double(invstd[c]));
  dbeta[c] = float(s1);
  if (dbias_out) {
    const float v = dbias ? float(acc_load(dbias + c)) : 0.f;
    dbias_out[c] = rb ? rbf(v) : v;
  }
}
int bn_bwd_finalize(const Acc* sums, const Acc* dbias, const float* mean, const float* invstd, int C,
                    int round_bf16, float* dgamma, float* dbeta, float* dbias_out, cudaStream_t st) {
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, dbias, mean, invstd, C, round_bf16, dgamma, dbeta,
                                                          dbias_out);
  NSM_CHECK_LAUNCH("bn_bwd_finalize");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// AvgPool2d(2) adjoint (+ skip gradient), plane add, bilinear adjoint
// ------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256) pool_bwd_add_kernel(const Planes a, const Planes dpool, const Planes out, int N,
                                                           int H, int W, int C, int fmt) {
  const int cgs = C / 8, Hp = H / 2, Wp = W / 2;
  const bool rb = FMT == kFmtBf16;
  const long long total = (long long)N * H * W * cgs;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const unsigned iu = (unsigned)i;
    const int cg = int(iu % (unsigned)cgs);
    const unsigned pixi = iu / (unsigned)cgs;
    const unsigned t1 = pixi / (unsigned)W;
    const int x = int(pixi - t1 * (unsigned)W);
    const int n = int(t1 / (unsigned)H);
    const int y = int(t1 - (unsigned)n * (unsigned)H);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (a.p[0]) load8(a, (size_t)pixi * C + cg * 8, FMT, v);
    if ((y >> 1) < Hp && (x >> 1) < Wp) {
      float d[8];
      load8(dpool, (((size_t)n * Hp + (y >> 1)) * Wp + (x >> 1)) * C + cg * 8, FMT, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = v[e] + d[e] * 0.25f;
      if (rb) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
      }
    }
    store8(out, (size_t)pixi * C + cg * 8, FMT, v);
  }
}
int pool_bwd_add(const Planes& a, const Planes& dpool, const Planes& out, int N, int H, int W, int C, int fmt,
                 cudaStream_t st) {
  if (check_c("pool_bwd_add", C)) return 1;
  {
    if (fmt == kFmtBf16) pool_bwd_add_kernel<kFmtBf16><<<grid_for((long long)N * H * W * (C / 8), 256), 256, 0, st>>>(a, dpool, out, N, H, W, C, fmt);
    else if (fmt == kFmtF16x2) pool_bwd_add_kernel<kFmtF16x2><<<grid_for((long long)N * H * W * (C / 8), 256), 256, 0, st>>>(a, dpool, out, N, H, W, C, fmt);
    else pool_bwd_add_kernel<kFmtBf16x2><<<grid_for((long long)N * H * W * (C / 8), 256), 256, 0, st>>>(a, dpool, out, N, H, W, C, fmt);
  }
  NSM_CHECK_LAUNCH("pool_bwd_add");
  return 0;
}

template <int FMT>
__global__ void __launch_bounds__(256) planes_add_kernel(const Planes a, const Planes b, const Planes out,
                                                         long long n8, int fmt) {
  const bool rb = FMT == kFmtBf16;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    float x[8], y[8];
    load8(a, (size_t)i * 8, FMT, x);
    load8(b, (size_t)i * 8, FMT, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    if (rb) {
#pragma unroll
      for (int e = 0; e < 8; e += 2) rbf2(x[e], x[e + 1]);
    }
    store8(out, (size_t)i * 8, FMT, x);
  }
}
int planes_add(const Planes& a, const Planes& b, const Planes& out, long long numel, int fmt, cudaStream_t st) {
  if (numel % 8) {
    set_error("planes_add: numel %lld not a multiple of 8", numel);
    return 1;
  }
  {
    if (fmt == kFmtBf16) planes_add_kernel<kFmtBf16><<<grid_for(numel / 8, 256), 256, 0, st>>>(a, b, out, numel / 8, fmt);
    else if (fmt == kFmtF16x2) planes_add_kernel<kFmtF16x2><<<grid_for(numel / 8, 256), 256, 0, st>>>(a, b, out, numel / 8, fmt);
    else planes_add_kernel<kFmtBf16x2><<<grid_for(numel / 8, 256), 256, 0, st>>>(a, b, out, numel / 8, fmt);
  }
  NSM_CHECK_LAUNCH("planes_add");
  return 0;
}

// forward weights of output index `dst` of an align_corners resize in_size -> out_size (same arithmetic as
// stream_kernels.cu / ATen)
struct Lerp2 {
  int i0, i1;
  float w0, w1;
};
__device__ __forceinline__ Lerp2 lerp_of(int dst, int in_size, int out_size) {
  const float scale = out_size > 1 ? float(in_size - 1) / float(out_size - 1) : 0.f;
  const float src = scale * float(dst);
  int i0 = int(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  Lerp2 l;
  l.i0 = i0;
  l.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  float lam = fminf(fmaxf(src - float(i0), 0.f), 1.f);
  l.w1 = lam;
  l.w0 = 1.f - lam;
  return l;
}
// range of output indices whose interpolation may touch input index r
__device__ __forceinline__ void touch_range(int r, int in_size, int out_size, int& lo, int& hi) {
  if (out_size <= 1 || in_size <= 1) {
    lo = 0;
    hi = out_size - 1;
    return;
  }
  const float inv = float(out_size - 1) / float(in_size - 1);
  lo = int(floorf(float(r - 1) * inv)) - 1;
  hi = int(ceilf(float(r + 1) * inv)) + 1;
  if (lo < 0) lo = 0;
  if (hi > out_size - 1) hi = out_size - 1;
}

template <int FMT>
__global__ void __launch_bounds__(256) bilinear_bwd_kernel(const Planes dout, int N, int ho, int wo, int C,
                                                           const Planes din, int hi, int wi, int fmt) {
  const int cgs = C / 8;
  const long long total = (long long)N * hi * wi * cgs;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const unsigned iu = (unsigned)i;
    const int cg = int(iu % (unsigned)cgs);
    const unsigned pixi = iu / (unsigned)cgs;
    const unsigned t1 = pixi / (unsigned)wi;
    const int q = int(pixi - t1 * (unsigned)wi);
    const int n = int(t1 / (unsigned)hi);
    const int r = int(t1 - (unsigned)n * (unsigned)hi);
    int ylo, yhi, xlo, xhi;
    touch_range(r, hi, ho, ylo, yhi);
    touch_range(q, wi, wo, xlo, xhi);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int y = ylo; y <= yhi; ++y) {
      const Lerp2 ly = lerp_of(y, hi, ho);
      const float wy = (ly.i0 == r ? ly.w0 : 0.f) + (ly.i1 == r ? ly.w1 : 0.f);
      if (wy == 0.f) continue;
      for (int x = xlo; x <= xhi; ++x) {
        const Lerp2 lx = lerp_of(x, wi, wo);
        const float wx = (lx.i0 == q ? lx.w0 : 0.f) + (lx.i1 == q ? lx.w1 : 0.f);
        if (wx == 0.f) continue;
        float d[8];
        load8(dout, (((size_t)n * ho + y) * wo + x) * C + cg * 8, FMT, d);
        const float wgt = wy * wx;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, d[e], acc[e]);
      }
    }
    store8(din, (size_t)pixi * C + cg * 8, FMT, acc);
  }
}
// Adjoint of the plain x2 align_corners up-sample (ho == 2hi, wo == 2wi): source pixel r receives from output rows
// 2r-1 (weight wo1(r-1)), 2r (we1(r)), 2r+1 (wo0(r)), 2r+2 (we0(r+1)) -- the transposed stencil of upsample2x_kernel in
// stream_kernels.cu -- and likewise for columns; separable, 16 loads of 16 B per thread, static indices.
__device__ __forceinline__ void up2x_w(int b, int in_size, float& we0, float& we1, float& wo0, float& wo1) {
  const int out_size = 2 * in_size;
  const float scale = out_size > 1 ? float(in_size - 1) / float(out_size - 1) : 0.f;
  const float se = scale * float(2 * b), so = scale * float(2 * b + 1);
  we1 = b == 0 ? 1.f : fminf(fmaxf(se - float(b - 1), 0.f), 1.f);
  we0 = 1.f - we1;
  wo1 = b >= in_size - 1 ? 0.f : fminf(fmaxf(so - float(b), 0.f), 1.f);
  wo0 = 1.f - wo1;
}
__device__ __forceinline__ void up2x_adjoint_weights(int r, int in_size, float* w) {
  float a0, a1, b0, b1;
  w[0] = w[3] = 0.f;
  if (r > 0) {
    up2x_w(r - 1, in_size, a0, a1, b0, b1);
    w[0] = b1;                       // odd row 2(r-1)+1 uses sources (r-1, r)
  }
  up2x_w(r, in_size, a0, a1, b0, b1);
  w[1] = a1;                         // even row 2r uses (r-1, r)
  w[2] = b0;                         // odd row 2r+1 uses (r, r+1)
  if (r == 0) w[1] = a0 + a1;        // row 0: both taps hit source 0
  if (r == in_size - 1) w[2] = b0 + b1;
  if (r < in_size - 1) {
    up2x_w(r + 1, in_size, a0, a1, b0, b1);
    w[3] = a0;                       // even row 2(r+1) uses (r, r+1)
  }
}

template <int FMT>
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const Planes dout, int N, int C, const Planes din, int hi,
                                                             int wi, int fmt, int cg_shift) {
  const int cgs = 1 << cg_shift;
  const int j = blockIdx.y * 256 + threadIdx.x;
  if (j >= wi * cgs) return;
  const int cg = j & (cgs - 1), q = j >> cg_shift;
  const int n = blockIdx.x / hi, r = blockIdx.x - n * hi;
  const int ho = 2 * hi, wo = 2 * wi;
  float wy[4], wx[4];
  up2x_adjoint_weights(r, hi, wy);
  up2x_adjoint_weights(q, wi, wx);
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int y = 2 * r - 1 + a;
    if (y < 0 || y >= ho || wy[a] == 0.f) continue;
    float t[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int x = 2 * q - 1 + b;
      if (x < 0 || x >= wo || wx[b] == 0.f) continue;
      float d[8];
      load8(dout, (((size_t)n * ho + y) * wo + x) * C + cg * 8, FMT, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = fmaf(wx[b], d[e], t[e]);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = fmaf(wy[a], t[e], acc[e]);
  }
  store8(din, (((size_t)n * hi + r) * wi + q) * C + cg * 8, FMT, acc);
}

// Adjoint of the composite (x2 up-sample then resize to (ho, wo)) in one pass: source pixel (r, q) gathers from every
// output whose 3-tap window (resample.cuh: composite_taps) contains it.
template <int FMT>
__global__ void __launch_bounds__(256) composite_bwd_kernel(const Planes dout, int N, int ho, int wo, int C,
                                                            const Planes din, int hi, int wi, int fmt, int cg_shift) {
  constexpr int kMaxCand = 16;
  const int cgs = 1 << cg_shift;
  const int j = blockIdx.y * 256 + threadIdx.x;
  const bool active = j < wi * cgs;
  const int cg = j & (cgs - 1), q = active ? (j >> cg_shift) : 0;
  const int n = blockIdx.x / hi, r = blockIdx.x - n * hi;
  // Block-shared weight tables: for the source row r (one per block) and for each of the <= 256/cgs + 1 source columns
  // of this block, the weight of that source index in every candidate output index (resample.cuh: composite_taps).
  __shared__ float s_wy[kMaxCand], s_wx[260][kMaxCand + 1];
  __shared__ int s_ylo, s_ny, s_xlo[260], s_nx[260];
  auto range = [](int idx, int in_size, int out_size, int& lo, int& cnt) {
    int hi_;
    if (in_size <= 1 || out_size <= 1) {
      lo = 0;
      hi_ = out_size - 1;
    } else {
      const float inv = float(out_size - 1) / float(in_size - 1);
      lo = int(floorf(float(idx - 2) * inv)) - 1;
      hi_ = int(ceilf(float(idx + 2) * inv)) + 1;
      if (lo < 0) lo = 0;
      if (hi_ > out_size - 1) hi_ = out_size - 1;
    }
    cnt = hi_ - lo + 1;
    if (cnt > kMaxCand) cnt = kMaxCand;   // cannot happen for the ratios of this network (<= 12)
  };
  auto weight_of = [](int src, int out_idx, int in_size, int out_size) {
    const Tap3 t = composite_taps(out_idx, in_size, out_size);
    const int d = src - t.rmin;
    return d == 0 ? t.w[0] : (d == 1 ? t.w[1] : (d == 2 ? t.w[2] : 0.f));
  };
  const int q_first = (blockIdx.y * 256) >> cg_shift;
  const int q_count = ((blockIdx.y * 256 + 255) >> cg_shift) - q_first + 1;
  if (threadIdx.x < 32) {   // warp 0: row table
    int lo, cnt;
    range(r, hi, ho, lo, cnt);
    if (threadIdx.x == 0) {
      s_ylo = lo;
      s_ny = cnt;
    }
    if ((int)threadIdx.x < cnt) s_wy[threadIdx.x] = weight_of(r, lo + threadIdx.x, hi, ho);
  }
  // column tables: (column, candidate) pairs spread over the block
  for (int k = threadIdx.x; k < q_count * kMaxCand; k += 256) {
    const int qi = k / kMaxCand, c = k % kMaxCand, qq = q_first + qi;
    if (qq < wi) {
      int lo, cnt;
      range(qq, wi, wo, lo, cnt);
      if (c == 0) {
        s_xlo[qi] = lo;
        s_nx[qi] = cnt;
      }
      s_wx[qi][c] = c < cnt ? weight_of(qq, lo + c, wi, wo) : 0.f;
    }
  }
  __syncthreads();
  if (!active) return;
  const int ylo = s_ylo, ny = s_ny, qi = q - q_first, xlo = s_xlo[qi], nx = s_nx[qi];
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int ky = 0; ky < ny; ++ky) {
    const float wy = s_wy[ky];
    if (wy == 0.f) continue;
    for (int kx = 0; kx < nx; ++kx) {
      const float wx = s_wx[qi][kx];
      if (wx == 0.f) continue;
      float d[8];
      load8(dout, (((size_t)n * ho + ylo + ky) * wo + xlo + kx) * C + cg * 8, FMT, d);
      const float wgt = wy * wx;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(wgt, d[e], acc[e]);
    }
  }
  store8(din, (((size_t)n * hi + r) * wi + q) * C + cg * 8, FMT, acc);
}

// Strip-walking adjoint for the two shapes that carry the bytes: the plain x2 up-sample (ho == 2hi: levels 6-8 at sizes
// divisible by 8) and the same-size composite of level 9 (x2 up, then back down: a position-dependent 3-tap blur).
//   din[r][q] = sum_y wy(y, r) * T[y][q],   T[y][q] = sum_x wx(x, q) * dout[y][x]
// The weights are read off the forward taps (resample.cuh: composite_taps), so forward and adjoint cannot drift apart.
// Source row r only receives from WIN consecutive output rows (x2: 2r-1 .. 2r+2; same size: r-2 .. r+2, the outer two with
// zero weight away from the borders), and columns likewise.  One thread owns a source column (8 channels) and walks down a
// strip of source rows: every horizontally reduced row T[y] is formed ONCE (WIN loads) and feeds all source rows that touch
// it -- 4-5 loads per source pixel where the gather kernels below issue 16-25 -- and stays in registers (static indices).
constexpr int kBwdStrip = 16;
template <int FMT, bool X2>
__global__ void __launch_bounds__(256) upsample_bwd_strip_kernel(const Planes dout, int N, int C, const Planes din, int hi,
                                                                 int wi, int cg_shift, int strips) {
  constexpr int WIN = X2 ? 4 : 5, STEP = X2 ? 2 : 1, OFF = X2 ? 1 : 2;   // window of output rows y = STEP*r - OFF + j
  const int ho = X2 ? 2 * hi : hi, wo = X2 ? 2 * wi : wi;
  const int cgs = 1 << cg_shift;
  const int j = blockIdx.y * 256 + threadIdx.x;
  const bool active = j < wi * cgs;
  const int cg = j & (cgs - 1), q = active ? (j >> cg_shift) : 0;
  const int n = blockIdx.x / strips, r0 = (blockIdx.x - n * strips) * kBwdStrip;
  const int r1 = min(r0 + kBwdStrip, hi);
  auto weight_of = [](int src, int out_idx, int in_size, int out_size) {
    if (out_idx < 0 || out_idx >= out_size) return 0.f;
    const Tap3 t = composite_taps(out_idx, in_size, out_size);
    const int d = src - t.rmin;
    return d == 0 ? t.w[0] : (d == 1 ? t.w[1] : (d == 2 ? t.w[2] : 0.f));
  };
  __shared__ float s_wy[kBwdStrip][WIN];
  for (int k = threadIdx.x; k < kBwdStrip * WIN; k += 256) {
    const int rr = r0 + k / WIN, jj = k % WIN;
    s_wy[k / WIN][jj] = rr < hi ? weight_of(rr, STEP * rr - OFF + jj, hi, ho) : 0.f;
  }
  __syncthreads();
  if (!active) return;
  float wx[WIN];
#pragma unroll
  for (int b = 0; b < WIN; ++b) wx[b] = weight_of(q, STEP * q - OFF + b, wi, wo);
  const int x0 = STEP * q - OFF;
  // T rows of the window, horizontally reduced; a row outside the image (or with all-zero weights) is zero
  auto hrow = [&](int y, float (&t)[8]) {
#pragma unroll
    for (int e = 0; e < 8; ++e) t[e] = 0.f;
    if (y < 0 || y >= ho) return;
    const size_t rbase = ((size_t)n * ho + y) * wo;
    Raw8 raw[WIN];
#pragma unroll
    for (int b = 0; b < WIN; ++b)
      if (wx[b] != 0.f) raw[b] = load_raw8(dout, (rbase + x0 + b) * C + cg * 8, FMT);   // wx != 0 implies 0 <= x < wo
#pragma unroll
    for (int b = 0; b < WIN; ++b) {
      if (wx[b] == 0.f) continue;
      float d[8];
      unpack8(raw[b], FMT, d);
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = fmaf(wx[b], d[e], t[e]);
    }
  };
  float T[WIN][8];
#pragma unroll
  for (int a = STEP; a < WIN; ++a) hrow(STEP * r0 - OFF + a - STEP, T[a]);   // rows the first iteration shifts into place
  for (int r = r0; r < r1; ++r) {
#pragma unroll
    for (int a = 0; a + STEP < WIN; ++a)
#pragma unroll
      for (int e = 0; e < 8; ++e) T[a][e] = T[a + STEP][e];
#pragma unroll
    for (int a = WIN - STEP; a < WIN; ++a) hrow(STEP * r - OFF + a, T[a]);
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int a = 0; a < WIN; ++a) {
      const float w = s_wy[r - r0][a];
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, T[a][e], acc[e]);
    }
    store8(din, (((size_t)n * hi + r) * wi + q) * C + cg * 8, FMT, acc);
  }
}

int upsample_match_bwd(const Planes& dout, int N, int ho, int wo, int C, const Planes& din, int hi, int wi, int fmt,
                       cudaStream_t st) {
  if (check_c("upsample_match_bwd", C)) return 1;
  const int cgs = C / 8;
  int shift = 0;
  while ((1 << shift) < cgs) ++shift;
  dim3 grid((unsigned)(N * hi), (unsigned)((wi * cgs + 255) / 256));
  static const bool strip_off = getenv("NSM_NO_BWD_STRIP") != nullptr;   // A/B switch: the per-pixel gather kernels
  const bool x2 = ho == 2 * hi && wo == 2 * wi, same = ho == hi && wo == wi;
  if ((x2 || same) && !strip_off && hi >= 2 && wi >= 2) {
    const int strips = (hi + kBwdStrip - 1) / kBwdStrip;
    dim3 sgrid((unsigned)(N * strips), (unsigned)((wi * cgs + 255) / 256));
#define NSM_BWD_STRIP(F)                                                                                               \
    if (x2) upsample_bwd_strip_kernel<F, true><<<sgrid, 256, 0, st>>>(dout, N, C, din, hi, wi, shift, strips);          \
    else upsample_bwd_strip_kernel<F, false><<<sgrid, 256, 0, st>>>(dout, N, C, din, hi, wi, shift, strips)
    if (fmt == kFmtBf16) { NSM_BWD_STRIP(kFmtBf16); }
    else if (fmt == kFmtF16x2) { NSM_BWD_STRIP(kFmtF16x2); }
    else { NSM_BWD_STRIP(kFmtBf16x2); }
#undef NSM_BWD_STRIP
  } else if (x2) {
    if (fmt == kFmtBf16) upsample2x_bwd_kernel<kFmtBf16><<<grid, 256, 0, st>>>(dout, N, C, din, hi, wi, fmt, shift);
    else if (fmt == kFmtF16x2) upsample2x_bwd_kernel<kFmtF16x2><<<grid, 256, 0, st>>>(dout, N, C, din, hi, wi, fmt, shift);
    else upsample2x_bwd_kernel<kFmtBf16x2><<<grid, 256, 0, st>>>(dout, N, C, din, hi, wi, fmt, shift);
  }
  else {
    if (fmt == kFmtBf16) composite_bwd_kernel<kFmtBf16><<<grid, 256, 0, st>>>(dout, N, ho, wo, C, din, hi, wi, fmt, shift);
    else if (fmt == kFmtF16x2) composite_bwd_kernel<kFmtF16x2><<<grid, 256, 0, st>>>(dout, N, ho, wo, C, din, hi, wi, fmt, shift);
    else composite_bwd_kernel<kFmtBf16x2><<<grid, 256, 0, st>>>(dout, N, ho, wo, C, din, hi, wi, fmt, shift);
  }
  NSM_CHECK_LAUNCH("upsample_match_bwd");
  return 0;
}

int bilinear_bwd(const Planes& dout, int N, int ho, int wo, int C, const Planes& din, int hi, int wi, int fmt,
                 cudaStream_t st) {
  if (check_c("bilinear_bwd", C)) return 1;
  if (ho == 2 * hi && wo == 2 * wi) {
    const int cgs = C / 8;
    int shift = 0;
    while ((1 << shift) < cgs) ++shift;
    dim3 grid((unsigned)(N * hi), (unsigned)((wi * cgs + 255) / 256));
    {
      if (fmt == kFmtBf16) upsample2x_bwd_kernel<kFmtBf16><<<grid, 256, 0, st>>>(dout, N, C, din, hi, wi, fmt, shift);
      else if (fmt == kFmtF16x2) upsample2x_bwd_kernel<kFmtF16x2><<<grid, 256, 0, st>>>(dout, N, C, din, hi, wi, fmt, shift);
      else upsample2x_bwd_kernel<kFmtBf16x2><<<grid, 256, 0, st>>>(dout, N, C, din, hi, wi, fmt, shift);
    }
    NSM_CHECK_LAUNCH("upsample2x_bwd");
    return 0;
  }
  {
    if (fmt == kFmtBf16) bilinear_bwd_kernel<kFmtBf16><<<grid_for((long long)N * hi * wi * (C / 8), 256), 256, 0, st>>>(dout, N, ho, wo, C, din, hi, wi,
                                                                                         fmt);
    else if (fmt == kFmtF16x2) bilinear_bwd_kernel<kFmtF16x2><<<grid_for((long long)N * hi * wi * (C / 8), 256), 256, 0, st>>>(dout, N, ho, wo, C, din, hi, wi,
                                                                                         fmt);
    else bilinear_bwd_kernel<kFmtBf16x2><<<grid_for((long long)N * hi * wi * (C / 8), 256), 256, 0, st>>>(dout, N, ho, wo, C, din, hi, wi,
                                                                                         fmt);
  }
  NSM_CHECK_LAUNCH("bilinear_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// network input / output stages
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) train_input_prep_kernel(const float* __restrict__ x, int N, int Hin, int Win,
                                                               const Planes out, int fmt, int cpad) {
  const int H = Hin - (Hin & 1), W = Win - (Win & 1), h = H / 2, w = W / 2;
  const bool resize = (Hin & 1) || (Win & 1);
  const bool rb = fmt == kFmtBf16;
  // one thread = one output pixel x one 8-channel group of the cpad (16, or 64 zero-padded) channels (groups >= 2 are zero)
  const int gshift = cpad == 16 ? 1 : 3;
  const long long total = ((long long)N * h * w) << gshift;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = int(i & ((1 << gshift) - 1));
    const long long pix = i >> gshift;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (cg < 2) {
      const unsigned t = (unsigned)pix / (unsigned)w;
      const int px = int((unsigned)pix - t * (unsigned)w);
      const int n = int(t / (unsigned)h);
      const int py = int(t - (unsigned)n * (unsigned)h);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int ch = cg * 8 + e;           // un-shuffled channel = c*4 + dy*2 + dx
        const int c = ch >> 2, dy = (ch >> 1) & 1, dx = ch & 1;
        const int Y = 2 * py + dy, X = 2 * px + dx;
        const float* base = x + ((size_t)n * 4 + c) * Hin * Win;
        float val;
        if (!resize) {
          val = base[(size_t)Y * Win + X];
        } else {
          const Lerp2 ly = lerp_of(Y, Hin, H), lx = lerp_of(X, Win, W);
          const float v00 = base[(size_t)ly.i0 * Win + lx.i0], v01 = base[(size_t)ly.i0 * Win + lx.i1];
          const float v10 = base[(size_t)ly.i1 * Win + lx.i0], v11 = base[(size_t)ly.i1 * Win + lx.i1];
          val = ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
        }
        v[e] = rb ? rbf(val) : val;
      }
    }
    store8(out, (size_t)pix * cpad + cg * 8, fmt, v);
  }
}
int train_input_prep(const float* x, int N, int Hin, int Win, const Planes& out, int fmt, cudaStream_t st, int cpad) {
  if (cpad != 16 && cpad != 64) {
    set_error("train_input_prep: channel pitch %d (16 or 64)", cpad);
    return 1;
  }
  const long long total = (long long)N * (Hin / 2) * (Win / 2) * (cpad / 8);
  train_input_prep_kernel<<<grid_for(total, 256), 256, 0, st>>>(x, N, Hin, Win, out, fmt, cpad);
  NSM_CHECK_LAUNCH("train_input_prep");
  return 0;
}

__global__ void __launch_bounds__(256) train_input_grad_kernel(const Planes dx16, int N, int H, int W, float* dx,
                                                               int fmt, int cpad) {
  const int h = H / 2, w = W / 2;
  const long long total = (long long)N * h * w * 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = int(i & 1);
    const long long pix = i >> 1;
    const unsigned t = (unsigned)pix / (unsigned)w;
    const int px = int((unsigned)pix - t * (unsigned)w);
    const int n = int(t / (unsigned)h);
    const int py = int(t - (unsigned)n * (unsigned)h);
    float v[8];
    load8(dx16, (size_t)pix * cpad + cg * 8, fmt, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int ch = cg * 8 + e;
      const int c = ch >> 2, dy = (ch >> 1) & 1, dxx = ch & 1;
      dx[(((size_t)n * 4 + c) * H + 2 * py + dy) * W + 2 * px + dxx] = v[e];
    }
  }
}
int train_input_grad(const Planes& dx16, int N, int H, int W, float* dx, int fmt, cudaStream_t st, int cpad) {
  if ((H & 1) || (W & 1)) {
    set_error("train_input_grad: odd input sizes are not supported for the input gradient");
    return 1;
  }
  if (cpad != 16 && cpad != 64) {
    set_error("train_input_grad: channel pitch %d (16 or 64)", cpad);
    return 1;
  }
  train_input_grad_kernel<<<grid_for((long long)N * (H / 2) * (W / 2) * 2, 256), 256, 0, st>>>(dx16, N, H, W, dx, fmt, cpad);
  NSM_CHECK_LAUNCH("train_input_grad");
  return 0;
}

__global__ void __launch_bounds__(256) sigmoid_shuffle_fwd_kernel(const Planes c10, int N, int h, int w, int fmt,
                                                                  float* y, int px4) {
  const bool rb = fmt == kFmtBf16;
  const long long total = (long long)N * h * w;
  const int W = 2 * w, H = 2 * h;
  for (long long pix = blockIdx.x * 256LL + threadIdx.x; pix < total; pix += (long long)gridDim.x * 256) {
    float v[8];
    // px4: four pixels share one 64-channel row, pixel po of the group holds its 4 channels at [4*po, 4*po+4)
    const int off = px4 ? int(pix & 3) * 4 : 0;
    load8(c10, px4 ? (size_t)(pix >> 2) * 64 + (off & 8) : (size_t)pix * 64, fmt, v);
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float s = 1.f / (1.f + expf(-((off & 4) ? v[4 + k] : v[k])));
      r[k] = rb ? rbf(s) : s;
    }
    const unsigned t = (unsigned)pix / (unsigned)w;
    const int x = int((unsigned)pix - t * (unsigned)w);
    const unsigned n = t / (unsigned)h;
    const int yy = int(t - n * (unsigned)h);
    float* o = y + ((size_t)n * H + 2 * yy) * W + 2 * x;
    *reinterpret_cast<float2*>(o) = make_float2(r[0], r[1]);
    *reinterpret_cast<float2*>(o + W) = make_float2(r[2], r[3]);
  }
}
int sigmoid_shuffle_fwd(const Planes& c10, int N, int h, int w, int fmt, float* y, cudaStream_t st, int px4) {
  if (px4 && (w & 3)) {
    set_error("sigmoid_shuffle_fwd: pixel-packed layout needs w %% 4 == 0 (w = %d)", w);
    return 1;
  }
  sigmoid_shuffle_fwd_kernel<<<grid_for((long long)N * h * w, 256), 256, 0, st>>>(c10, N, h, w, fmt, y, px4);
  NSM_CHECK_LAUNCH("sigmoid_shuffle_fwd");
  return 0;
}

// px4: one thread = one group of four pixels x one 8-channel group: groups 0 and 1 carry two pixels each, 2..7 are zero
__global__ void __launch_bounds__(256) sigmoid_shuffle_bwd_px4_kernel(const float* __restrict__ dy,
                                                                      const float* __restrict__ y, int N, int h, int w,
                                                                      int fmt, const Planes dc10) {
  const bool rb = fmt == kFmtBf16;
  const long long total = (long long)N * h * (w / 4) * 8;
  const int W = 2 * w, H = 2 * h;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = int(i & 7);
    const long long grp = i >> 3;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (cg < 2) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const long long pix = grp * 4 + cg * 2 + half;
        const unsigned t = (unsigned)pix / (unsigned)w;
        const int x = int((unsigned)pix - t * (unsigned)w);
        const unsigned n = t / (unsigned)h;
        const int yy = int(t - n * (unsigned)h);
        const size_t o = ((size_t)n * H + 2 * yy) * W + 2 * x;
        const float2 g0 = *reinterpret_cast<const float2*>(dy + o), g1 = *reinterpret_cast<const float2*>(dy + o + W);
        const float2 s0 = *reinterpret_cast<const float2*>(y + o), s1 = *reinterpret_cast<const float2*>(y + o + W);
        const float g[4] = {g0.x, g0.y, g1.x, g1.y}, s[4] = {s0.x, s0.y, s1.x, s1.y};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float gg = rb ? rbf(g[k]) : g[k];
          const float d = gg * ((1.f - s[k]) * s[k]);
          v[half * 4 + k] = rb ? rbf(d) : d;
        }
      }
    }
    store8(dc10, (size_t)grp * 64 + cg * 8, fmt, v);
  }
}

__global__ void __launch_bounds__(256) sigmoid_shuffle_bwd_kernel(const float* __restrict__ dy,
                                                                  const float* __restrict__ y, int N, int h, int w,
                                                                  int fmt, const Planes dc10) {
  const bool rb = fmt == kFmtBf16;
  const long long total = (long long)N * h * w * 8;
  const int W = 2 * w, H = 2 * h;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int cg = int(i & 7);
    const long long pix = i >> 3;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    if (cg == 0) {
      const unsigned t = (unsigned)pix / (unsigned)w;
      const int x = int((unsigned)pix - t * (unsigned)w);
      const unsigned n = t / (unsigned)h;
      const int yy = int(t - n * (unsigned)h);
      const size_t o = ((size_t)n * H + 2 * yy) * W + 2 * x;
      const float2 g0 = *reinterpret_cast<const float2*>(dy + o), g1 = *reinterpret_cast<const float2*>(dy + o + W);
      const float2 s0 = *reinterpret_cast<const float2*>(y + o), s1 = *reinterpret_cast<const float2*>(y + o + W);
      const float g[4] = {g0.x, g0.y, g1.x, g1.y}, s[4] = {s0.x, s0.y, s1.x, s1.y};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float gg = rb ? rbf(g[k]) : g[k];
        const float d = gg * ((1.f - s[k]) * s[k]);   // sigmoid_backward: grad * (1 - y) * y
        v[k] = rb ? rbf(d) : d;
      }
    }
    store8(dc10, (size_t)pix * 64 + cg * 8, fmt, v);
  }
}
int sigmoid_shuffle_bwd(const float* dy, const float* y, int N, int h, int w, int fmt, const Planes& dc10,
                        cudaStream_t st, int px4) {
  if (px4 && (w & 3)) {
    set_error("sigmoid_shuffle_bwd: pixel-packed layout needs w %% 4 == 0 (w = %d)", w);
    return 1;
  }
  if (px4) sigmoid_shuffle_bwd_px4_kernel<<<grid_for((long long)N * h * (w / 4) * 8, 256), 256, 0, st>>>(dy, y, N, h, w, fmt, dc10);
  else sigmoid_shuffle_bwd_kernel<<<grid_for((long long)N * h * w * 8, 256), 256, 0, st>>>(dy, y, N, h, w, fmt, dc10);
  NSM_CHECK_LAUNCH("sigmoid_shuffle_bwd");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// padded packing helpers
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_padded_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int CoutP,
                                               int CinP, int flip_t, int fmt, unsigned short* __restrict__ hi,
                                               unsigned short* __restrict__ lo) {
  const long long total = (long long)CoutP * CinP * taps;
  const int inner = flip_t ? CoutP : CinP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = int(i % inner);
    long long t = i / inner;
    const int tap = int(t % taps);
    const int row = int(t / taps);
    int co, ci, stap;
    if (!flip_t) {
      co = row; ci = col; stap = tap;
    } else {
      ci = row; co = col; stap = taps - 1 - tap;
    }
    const float v = (co < Cout && ci < Cin) ? w[((long long)co * Cin + ci) * taps + stap] : 0.f;
    unsigned short h, l;
    split_fmt(v, fmt, h, l);
    hi[i] = h;
    if (fmt != kFmtBf16) lo[i] = l;
  }
}
int pack_conv_weight_padded(const float* w, int Cout, int Cin, int ksize, int CoutP, int CinP, int flip_transpose,
                            int fmt, void* hi, void* lo, cudaStream_t st) {
  const int taps = ksize * ksize;
  const long long total = (long long)CoutP * CinP * taps;
  pack_conv_weight_padded_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, Cout, Cin, taps, CoutP, CinP, flip_transpose,
                                                                       fmt, (unsigned short*)hi, (unsigned short*)lo);
  NSM_CHECK_LAUNCH("pack_conv_weight_padded");
  return 0;
}

// Pixel-packed thin layers.  A convolution with few channels (16 -> 16, 16 -> 64, 64 -> 16, 16 -> 4) runs on the tcgen05
// kernel -- which wants multiples of 64 channels -- WITHOUT padding its tensors: four horizontally adjacent pixels x C
// channels are read as ONE pixel of 4C "virtual" channels ([N,H,W,C] is byte-identical to [N,H,W/4,4C]).  A 1x1 convolution
// becomes a 1x1 convolution with a block-diagonal weight, a 3x3 convolution a 3x3 convolution over pixel groups with a
// banded weight: virtual output (po, co) takes virtual input (pi, ci) of the group at horizontal offset dg with the real
// tap dx = 4*dg + pi - po when that lies in [-1, 1].  Bytes are the real ones (a quarter of the zero-padded layout).
__device__ __forceinline__ float px4_weight(const float* __restrict__ w, int Cout, int Cin, int taps, int co_v, int ci_v,
                                            int tap) {
  const int po = co_v / Cout, co = co_v - po * Cout, pi = ci_v / Cin, ci = ci_v - pi * Cin;
  if (po >= 4 || pi >= 4) return 0.f;   // zero-padded virtual channels (conv10: 4 x 4 = 16 of 64)
  if (taps == 1) return po == pi ? w[(long long)co * Cin + ci] : 0.f;
  const int dy = tap / 3, dg = tap % 3 - 1;
  const int dx = 4 * dg + pi - po;
  if (dx < -1 || dx > 1) return 0.f;
  return w[((long long)co * Cin + ci) * 9 + dy * 3 + dx + 1];
}
__global__ void pack_conv_weight_px4_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int CoutV,
                                            int CinV, int flip_t, int fmt, unsigned short* __restrict__ hi,
                                            unsigned short* __restrict__ lo) {
  const long long total = (long long)CoutV * CinV * taps;
  const int inner = flip_t ? CoutV : CinV;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = int(i % inner);
    long long t = i / inner;
    const int tap = int(t % taps);
    const int row = int(t / taps);
    int co, ci, stap;
    if (!flip_t) {
      co = row; ci = col; stap = tap;
    } else {
      ci = row; co = col; stap = taps - 1 - tap;
    }
    unsigned short h, l;
    split_fmt(px4_weight(w, Cout, Cin, taps, co, ci, stap), fmt, h, l);
    hi[i] = h;
    if (fmt != kFmtBf16) lo[i] = l;
  }
}
int pack_conv_weight_px4(const float* w, int Cout, int Cin, int ksize, int CoutV, int CinV, int flip_transpose, int fmt,
                         void* hi, void* lo, cudaStream_t st) {
  if (CoutV < 4 * Cout || CinV < 4 * Cin || (ksize != 1 && ksize != 3)) {
    set_error("pack_conv_weight_px4: %d->%d k%d does not fit %d->%d virtual channels", Cin, Cout, ksize, CinV, CoutV);
    return 1;
  }
  const int taps = ksize * ksize;
  const long long total = (long long)CoutV * CinV * taps;
  pack_conv_weight_px4_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, Cout, Cin, taps, CoutV, CinV, flip_transpose, fmt,
                                                                    (unsigned short*)hi, (unsigned short*)lo);
  NSM_CHECK_LAUNCH("pack_conv_weight_px4");
  return 0;
}
// gradient of the real weight = sum of the virtual weight's gradient over every (po, pi, dg) that maps onto the same tap
__global__ void px4_reduce_dw_kernel(const float* __restrict__ dwv, int Cout, int Cin, int taps, int CoutV, int CinV,
                                     float* __restrict__ dw) {
  const int total = Cout * Cin * taps;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int tap = i % taps, ci = (i / taps) % Cin, co = i / (taps * Cin);
    float s = 0.f;
    for (int po = 0; po < 4; ++po)
      for (int pi = 0; pi < 4; ++pi) {
        int vtap;
        if (taps == 1) {
          if (po != pi) continue;
          vtap = 0;
        } else {
          const int dx = tap % 3 - 1, d = dx - pi + po;   // 4 * dg
          if (d != -4 && d != 0 && d != 4) continue;
          vtap = (tap / 3) * 3 + d / 4 + 1;
        }
        s += dwv[((long long)(po * Cout + co) * CinV + pi * Cin + ci) * taps + vtap];
      }
    dw[i] = s;
  }
}
int px4_reduce_dw(const float* dwv, int Cout, int Cin, int ksize, int CoutV, int CinV, float* dw, cudaStream_t st) {
  const int taps = ksize * ksize;
  px4_reduce_dw_kernel<<<(Cout * Cin * taps + 255) / 256, 256, 0, st>>>(dwv, Cout, Cin, taps, CoutV, CinV, dw);
  NSM_CHECK_LAUNCH("px4_reduce_dw");
  return 0;
}
// out[v][c] = sum over the `groups` copies of in[v][g*C + c] (per-channel sums of a pixel-packed tensor -> real channels)
__global__ void fold_channel_sums_kernel(const Acc* in, int nvec, int CV, int groups, int C, Acc* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nvec * C) return;
  const int v = i / C, c = i - v * C;
  double s = 0.0;
  for (int g = 0; g < groups; ++g) s += acc_load(in + (long long)v * CV + g * C + c);   // fixed order
  acc_store(out + i, s);
}
int fold_channel_sums(const Acc* in, int nvec, int CV, int groups, int C, Acc* out, cudaStream_t st) {
  if (groups * C > CV) {
    set_error("fold_channel_sums: %d x %d channels do not fit %d", groups, C, CV);
    return 1;
  }
  fold_channel_sums_kernel<<<(nvec * C + 127) / 128, 128, 0, st>>>(in, nvec, CV, groups, C, out);
  NSM_CHECK_LAUNCH("fold_channel_sums");
  return 0;
}
// dst[i] = src[i % n] for i < rep * n, `fill` beyond (bias of the virtual channels of a pixel-packed convolution)
__global__ void tile_vector_kernel(const float* src, int n, int rep, int npad, float fill, int rb, float* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) dst[i] = i < rep * n ? (rb ? rbf(src[i % n]) : src[i % n]) : fill;
}
int tile_vector(const float* src, int n, int rep, int npad, float fill, int round_bf16, float* dst, cudaStream_t st) {
  tile_vector_kernel<<<(npad + 127) / 128, 128, 0, st>>>(src, n, rep, npad, fill, round_bf16, dst);
  NSM_CHECK_LAUNCH("tile_vector");
  return 0;
}

__global__ void pad_vector_kernel(const float* src, int n, int npad, float fill, int rb, float* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) dst[i] = i < n ? (rb ? rbf(src[i]) : src[i]) : fill;
}
int pad_vector(const float* src, int n, int npad, float fill, int round_bf16, float* dst, cudaStream_t st) {
  pad_vector_kernel<<<(npad + 127) / 128, 128, 0, st>>>(src, n, npad, fill, round_bf16, dst);
  NSM_CHECK_LAUNCH("pad_vector");
  return 0;
}

}  // namespace nsm
