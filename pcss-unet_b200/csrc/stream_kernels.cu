// HBM-streaming kernels (see stream_kernels.cuh).  All are bandwidth-bound: vectorised 16-byte accesses, one pass
// over each tensor, per-channel parameters broadcast from shared memory, warp-shuffle reductions.
#include <stdio.h>

#include "nsm_common.cuh"
#include "resample.cuh"
#include "stream_kernels.cuh"

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

// ------------------------------------------------------------------------------------------------
// parameter packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int Cout, int Cin, int taps, int flip_t,
                                        int fmt, unsigned short* __restrict__ hi, unsigned short* __restrict__ lo) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    // destination index i -> (row, tap, col)
    long long dst_rows_k = flip_t ? Cout : Cin;  // inner (K) extent per tap
    int col = int(i % dst_rows_k);
    long long t = i / dst_rows_k;
    int tap = int(t % taps);
    int row = int(t / taps);
    int co, ci, stap;
    if (!flip_t) {
      co = row; ci = col; stap = tap;
    } else {  // dgrad: rows = Cin, inner = Cout, taps mirrored
      ci = row; co = col; stap = taps - 1 - tap;
    }
    const float v = w[((long long)co * Cin + ci) * taps + stap];
    unsigned short h, l;
    split_fmt(v, fmt, h, l);
    hi[i] = h;
    if (fmt != kFmtBf16) lo[i] = l;
  }
}

int pack_conv_weight(const float* w, int Cout, int Cin, int ksize, int flip_transpose, int fmt, void* hi, void* lo,
                     cudaStream_t st) {
  const int taps = ksize * ksize;
  const long long total = (long long)Cout * Cin * taps;
  pack_conv_weight_kernel<<<grid_for(total, 256), 256, 0, st>>>(w, Cout, Cin, taps, flip_transpose, fmt,
                                                               (unsigned short*)hi, (unsigned short*)lo);
  NSM_CHECK_LAUNCH("pack_conv_weight");
  return 0;
}

__global__ void bn_fold_eval_kernel(const float* g, const float* b, const float* m, const float* v, int C, float eps,
                                    float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float inv = 1.0f / sqrtf(v[c] + eps);
    const float s = g[c] * inv;
    scale[c] = s;
    shift[c] = b[c] - m[c] * s;
  }
}
int bn_fold_eval(const float* gamma, const float* beta, const float* mean, const float* var, int C, float eps,
                 float* scale, float* shift, cudaStream_t st) {
  bn_fold_eval_kernel<<<(C + 127) / 128, 128, 0, st>>>(gamma, beta, mean, var, C, eps, scale, shift);
  NSM_CHECK_LAUNCH("bn_fold_eval");
  return 0;
}

__global__ void copy_round_kernel(const float* src, float* dst, int n, int r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = r ? rbf(src[i]) : src[i];
}
int copy_round(const float* src, float* dst, int n, int round_bf16, cudaStream_t st) {
  copy_round_kernel<<<(n + 255) / 256, 256, 0, st>>>(src, dst, n, round_bf16);
  NSM_CHECK_LAUNCH("copy_round");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// layout conversion NCHW fp32 <-> NHWC bf16 planes (smem transpose, 32 channels x 32 pixels tiles)
// ------------------------------------------------------------------------------------------------
__global__ void nchw_to_planes_kernel(const float* __restrict__ x, int C, long long HW, int fmt, unsigned short* hi,
                                      unsigned short* lo) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? x[((long long)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    if (c < C && p < HW) {
      const long long o = ((long long)n * HW + p) * C + c;
      unsigned short h, l;
      split_fmt(tile[threadIdx.x][i], fmt, h, l);
      hi[o] = h;
      if (fmt != kFmtBf16) lo[o] = l;
    }
  }
}
int nchw_to_planes(const float* x, int N, int C, int H, int W, int fmt, void* hi, void* lo, cudaStream_t st) {
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  nchw_to_planes_kernel<<<grid, block, 0, st>>>(x, C, HW, fmt, (unsigned short*)hi, (unsigned short*)lo);
  NSM_CHECK_LAUNCH("nchw_to_planes");
  return 0;
}

__global__ void planes_to_nchw_kernel(const unsigned short* __restrict__ hi, const unsigned short* __restrict__ lo,
                                      int fmt, int C, long long HW, float* __restrict__ y) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int c = c0 + threadIdx.x;
    float v = 0.f;
    if (c < C && p < HW) {
      const long long o = ((long long)n * HW + p) * C + c;
      v = join_fmt(hi[o], fmt != kFmtBf16 ? lo[o] : (unsigned short)0, fmt);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const long long p = p0 + threadIdx.x;
    if (c < C && p < HW) y[((long long)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}
int planes_to_nchw(const void* hi, const void* lo, int N, int C, int H, int W, int fmt, float* y,
                   cudaStream_t st) {
  const long long HW = (long long)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (C + 31) / 32, N), block(32, 8);
  planes_to_nchw_kernel<<<grid, block, 0, st>>>((const unsigned short*)hi, (const unsigned short*)lo, fmt, C, HW, y);
  NSM_CHECK_LAUNCH("planes_to_nchw");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// bilinear helpers (align_corners=True), mirroring ATen's area_pixel_compute_scale / source_index:
//   scale = (in-1)/(out-1) (0 if out==1); src = scale*dst; i0 = floor(src) clamped; lambda = src - i0
// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// head stage
// ------------------------------------------------------------------------------------------------
constexpr int HEAD_TH = 16, HEAD_TW = 32;  // output tile (rows x cols) at (h, w) resolution; 256 threads x 2 pixels

struct HeadSmem {
  float x16[16][HEAD_TH + 2][HEAD_TW + 4];  // un-shuffled (optionally standardised / even-fixed) input + halo (pitch 36)
  float w0[16 * 9 * 16];                    // [ci][tap][co]
  float w1[16 * 64];                        // [ci][co]
  float b0[16], s0[16], t0[16];
  float b1[64], s1[64], t1[64];
};

__device__ __forceinline__ float head_fetch(const HeadParams& p, int n, int c, int Y, int X, int H, int W,
                                            bool resize) {
  // value of the (standardised, even-fixed) full-resolution input at (Y, X), channel c
  const float* base = p.x + ((size_t)n * 4 + c) * p.Hin * p.Win;
  const float mu = p.mean ? p.mean[c] : 0.f;
  const float sd = p.mean ? (p.std[c] + 1e-8f) : 1.f;
  if (!resize) {
    float v = base[(size_t)Y * p.Win + X];
    return p.mean ? (v - mu) / sd : v;
  }
  const Lerp ly = make_lerp(Y, p.Hin, H), lx = make_lerp(X, p.Win, W);
  float v00 = base[(size_t)ly.i0 * p.Win + lx.i0], v01 = base[(size_t)ly.i0 * p.Win + lx.i1];
  float v10 = base[(size_t)ly.i1 * p.Win + lx.i0], v11 = base[(size_t)ly.i1 * p.Win + lx.i1];
  if (p.mean) {
    v00 = (v00 - mu) / sd; v01 = (v01 - mu) / sd; v10 = (v10 - mu) / sd; v11 = (v11 - mu) / sd;
  }
  return ly.w0 * (lx.w0 * v00 + lx.w1 * v01) + ly.w1 * (lx.w0 * v10 + lx.w1 * v11);
}

__device__ __forceinline__ void store_planes64(const Planes& dst, size_t elem_off, const float* v, int fmt) {
  uint8_t* o0 = reinterpret_cast<uint8_t*>(dst.p[0]) + elem_off * 2;
  uint8_t* o1 = fmt != kFmtBf16 ? reinterpret_cast<uint8_t*>(dst.p[1]) + elem_off * 2 : nullptr;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t hw[4], lw[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
      hw[e] = pack_hi(a, b, fmt);
      lw[e] = pack_lo_resid(a, b, hw[e], fmt);
    }
    stg16(o0 + 16 * j, make_uint4(hw[0], hw[1], hw[2], hw[3]));
    if (fmt != kFmtBf16) stg16(o1 + 16 * j, make_uint4(lw[0], lw[1], lw[2], lw[3]));
  }
}

// One thread computes TWO horizontally adjacent output pixels: every weight vector fetched from shared memory feeds two
// FMAs (6-8 FMAs per LDS instead of 3-4), the two 3x3 input windows share 3x4 values, and the horizontal half of the
// 2x2 average pool stays inside the thread.
__global__ void __launch_bounds__(256, 2) head_eval_kernel(const __grid_constant__ HeadParams p) {
  extern __shared__ uint8_t head_smem_raw[];
  HeadSmem& s = *reinterpret_cast<HeadSmem*>(head_smem_raw);
  const int H = p.Hin - (p.Hin & 1), W = p.Win - (p.Win & 1);
  const int h = H >> 1, w = W >> 1;
  const bool resize = (p.Hin & 1) || (p.Win & 1);
  const bool rb = p.fmt == kFmtBf16;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * HEAD_TH, x0 = blockIdx.x * HEAD_TW;
  const int tid = threadIdx.x;

  for (int i = tid; i < 16 * 9 * 16; i += 256) {
    const int co = i & 15, tap = (i >> 4) % 9, ci = i / 144;
    s.w0[i] = p.w0[(co * 16 + ci) * 9 + tap];
  }
  for (int i = tid; i < 16 * 64; i += 256) {
    const int co = i & 63, ci = i >> 6;
    s.w1[i] = p.w1[co * 16 + ci];
  }
  if (tid < 16) { s.b0[tid] = p.b0[tid]; s.s0[tid] = p.s0[tid]; s.t0[tid] = p.t0[tid]; }
  if (tid < 64) { s.b1[tid] = p.b1[tid]; s.s1[tid] = p.s1[tid]; s.t1[tid] = p.t1[tid]; }

  // input tile: full-resolution rows 2*(y0-1) .. 2*(y0+HEAD_TH+1)-1, zero outside the (even-fixed) image
  constexpr int FRH = 2 * (HEAD_TH + 2), FRW = 2 * (HEAD_TW + 2);
  // Batches of 8 independent global loads per thread, all issued before the first use (the profile of the one-load-per-
  // iteration loop showed 56 % of all stall samples on the load -> convert dependency).
  constexpr int kLoadBatch = 8;
  for (int base = tid; base < 4 * FRH * FRW; base += 256 * kLoadBatch) {
    float v[kLoadBatch];
#pragma unroll
    for (int u = 0; u < kLoadBatch; ++u) {
      const int i = base + u * 256;
      v[u] = 0.f;
      if (i < 4 * FRH * FRW) {
        const int fx = i % FRW, fy = (i / FRW) % FRH, c = i / (FRW * FRH);
        const int Y = 2 * (y0 - 1) + fy, X = 2 * (x0 - 1) + fx;
        if (Y >= 0 && Y < H && X >= 0 && X < W) v[u] = head_fetch(p, n, c, Y, X, H, W, resize);
      }
    }
#pragma unroll
    for (int u = 0; u < kLoadBatch; ++u) {
      const int i = base + u * 256;
      if (i < 4 * FRH * FRW) {
        const int fx = i % FRW, fy = (i / FRW) % FRH, c = i / (FRW * FRH);
        // autocast casts the conv input to bf16; pixel_unshuffle(2): ch = c*4 + dy*2 + dx
        s.x16[c * 4 + (fy & 1) * 2 + (fx & 1)][fy >> 1][fx >> 1] = rb ? rbf(v[u]) : v[u];
      }
    }
  }
  __syncthreads();

  // thread -> pixels (y0 + ly, x0 + 2*lxp + {0,1}); a warp covers 2 rows x 32 columns, vertical pool partner = lane^16
  const int lane = tid & 31, wrp = tid >> 5;
  const int ly = 2 * wrp + (lane >> 4), lxp = lane & 15;
  const int y = y0 + ly, xa = x0 + 2 * lxp;
  const bool valid0 = y < h && xa < w, valid1 = y < h && xa + 1 < w;
  const size_t pix0 = ((size_t)n * h + y) * w + xa;

  if (p.x16.p[0]) {  // optional tap for tests
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (!(px ? valid1 : valid0)) continue;
      float t[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) t[c] = s.x16[c][ly + 1][2 * lxp + px + 1];
      uint8_t* o0 = reinterpret_cast<uint8_t*>(p.x16.p[0]) + (pix0 + px) * 16 * 2;
      uint8_t* o1 = p.fmt != kFmtBf16 ? reinterpret_cast<uint8_t*>(p.x16.p[1]) + (pix0 + px) * 16 * 2 : nullptr;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = t[8 * j + 2 * e], b = t[8 * j + 2 * e + 1];
          hw[e] = pack_hi(a, b, p.fmt);
          lw[e] = pack_lo_resid(a, b, hw[e], p.fmt);
        }
        stg16(o0 + 16 * j, make_uint4(hw[0], hw[1], hw[2], hw[3]));
        if (o1) stg16(o1 + 16 * j, make_uint4(lw[0], lw[1], lw[2], lw[3]));
      }
    }
  }

  // conv2.conv.0 : 3x3, 16 -> 16, two pixels
  float a[2][16];
#pragma unroll
  for (int co = 0; co < 16; ++co) a[0][co] = a[1][co] = 0.f;
#pragma unroll 1
  for (int ci = 0; ci < 16; ++ci) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const float2 v01 = *reinterpret_cast<const float2*>(&s.x16[ci][ly + r][2 * lxp]);
      const float2 v23 = *reinterpret_cast<const float2*>(&s.x16[ci][ly + r][2 * lxp + 2]);
      const float xv[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4* wv = reinterpret_cast<const float4*>(&s.w0[(ci * 9 + r * 3 + kx) * 16]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = wv[q];
          a[0][4 * q] = fmaf(xv[kx], ww.x, a[0][4 * q]); a[0][4 * q + 1] = fmaf(xv[kx], ww.y, a[0][4 * q + 1]);
          a[0][4 * q + 2] = fmaf(xv[kx], ww.z, a[0][4 * q + 2]); a[0][4 * q + 3] = fmaf(xv[kx], ww.w, a[0][4 * q + 3]);
          a[1][4 * q] = fmaf(xv[kx + 1], ww.x, a[1][4 * q]); a[1][4 * q + 1] = fmaf(xv[kx + 1], ww.y, a[1][4 * q + 1]);
          a[1][4 * q + 2] = fmaf(xv[kx + 1], ww.z, a[1][4 * q + 2]); a[1][4 * q + 3] = fmaf(xv[kx + 1], ww.w, a[1][4 * q + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int px = 0; px < 2; ++px)
#pragma unroll
    for (int co = 0; co < 16; ++co) {
      float v = a[px][co] + s.b0[co];
      if (rb) v = rbf(v);
      v = fmaf(v, s.s0[co], s.t0[co]);
      if (rb) v = rbf(v);
      v = lrelu02(v);
      if (rb) v = rbf(v);
      a[px][co] = v;
    }
  // conv2.conv.4 : 1x1, 16 -> 64, in two halves of 32 output channels (register budget)
  const int hp = h >> 1, wp = w >> 1;
  const bool pool_ok = !(lane & 16) && (y >> 1) < hp && (xa >> 1) < wp;
  const size_t ppix = ((size_t)n * hp + (y >> 1)) * wp + (xa >> 1);
#pragma unroll 1
  for (int half = 0; half < 2; ++half) {
    float o[2][32];
#pragma unroll
    for (int co = 0; co < 32; ++co) o[0][co] = o[1][co] = 0.f;
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) {
      const float4* wv = reinterpret_cast<const float4*>(&s.w1[ci * 64 + half * 32]);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 ww = wv[q];
        o[0][4 * q] = fmaf(a[0][ci], ww.x, o[0][4 * q]); o[0][4 * q + 1] = fmaf(a[0][ci], ww.y, o[0][4 * q + 1]);
        o[0][4 * q + 2] = fmaf(a[0][ci], ww.z, o[0][4 * q + 2]); o[0][4 * q + 3] = fmaf(a[0][ci], ww.w, o[0][4 * q + 3]);
        o[1][4 * q] = fmaf(a[1][ci], ww.x, o[1][4 * q]); o[1][4 * q + 1] = fmaf(a[1][ci], ww.y, o[1][4 * q + 1]);
        o[1][4 * q + 2] = fmaf(a[1][ci], ww.z, o[1][4 * q + 2]); o[1][4 * q + 3] = fmaf(a[1][ci], ww.w, o[1][4 * q + 3]);
      }
    }
    float pl[32];
#pragma unroll
    for (int co = 0; co < 32; ++co) {
      const int c = half * 32 + co;
#pragma unroll
      for (int px = 0; px < 2; ++px) {
        float v = o[px][co] + s.b1[c];
        if (rb) v = rbf(v);
        v = fmaf(v, s.s1[c], s.t1[c]);
        if (rb) v = rbf(v);
        v = lrelu02(v);
        if (rb) v = rbf(v);
        o[px][co] = v;
      }
      // AvgPool2d(2): horizontal pair in-thread, vertical partner = lane^16
      float t = o[0][co] + o[1][co];
      t += __shfl_xor_sync(0xffffffffu, t, 16);
      t *= 0.25f;
      pl[co] = rb ? rbf(t) : t;
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (!(px ? valid1 : valid0)) continue;
      const size_t off = (pix0 + px) * 64 + half * 32;
      uint8_t* o0 = reinterpret_cast<uint8_t*>(p.c2.p[0]) + off * 2;
      uint8_t* o1 = p.fmt != kFmtBf16 ? reinterpret_cast<uint8_t*>(p.c2.p[1]) + off * 2 : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float va = o[px][8 * j + 2 * e], vb = o[px][8 * j + 2 * e + 1];
          hw[e] = pack_hi(va, vb, p.fmt);
          lw[e] = pack_lo_resid(va, vb, hw[e], p.fmt);
        }
        stg16(o0 + 16 * j, make_uint4(hw[0], hw[1], hw[2], hw[3]));
        if (o1) stg16(o1 + 16 * j, make_uint4(lw[0], lw[1], lw[2], lw[3]));
      }
    }
    if (pool_ok) {
      const size_t off = ppix * 64 + half * 32;
      uint8_t* o0 = reinterpret_cast<uint8_t*>(p.p2.p[0]) + off * 2;
      uint8_t* o1 = p.fmt != kFmtBf16 ? reinterpret_cast<uint8_t*>(p.p2.p[1]) + off * 2 : nullptr;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float va = pl[8 * j + 2 * e], vb = pl[8 * j + 2 * e + 1];
          hw[e] = pack_hi(va, vb, p.fmt);
          lw[e] = pack_lo_resid(va, vb, hw[e], p.fmt);
        }
        stg16(o0 + 16 * j, make_uint4(hw[0], hw[1], hw[2], hw[3]));
        if (o1) stg16(o1 + 16 * j, make_uint4(lw[0], lw[1], lw[2], lw[3]));
      }
    }
  }
}

int head_eval(const HeadParams& p, cudaStream_t st) {
  const int H = p.Hin - (p.Hin & 1), W = p.Win - (p.Win & 1);
  const int h = H / 2, w = W / 2;
  if (h < 1 || w < 1 || p.N < 1) {
    set_error("head_eval: bad shape N=%d Hin=%d Win=%d", p.N, p.Hin, p.Win);
    return 1;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(HeadSmem));
    if (e != cudaSuccess) {
      set_error("head_eval: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    attr = true;
  }
  dim3 grid((w + HEAD_TW - 1) / HEAD_TW, (h + HEAD_TH - 1) / HEAD_TH, p.N);
  head_eval_kernel<<<grid, 256, sizeof(HeadSmem), st>>>(p);
  NSM_CHECK_LAUNCH("head_eval");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// tail stage
// ------------------------------------------------------------------------------------------------
struct TailSmem {
  float w1[64 * 16];  // [ci][co]
  float w10[16 * 4];  // [ci][co]
  float b1[16], s1[16], t1[16], b10[4];
};

template <int FMT>
__global__ void __launch_bounds__(256) tail_eval_kernel(const __grid_constant__ TailParams p) {
  __shared__ TailSmem s;
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * 16; i += 256) {
    const int co = i & 15, ci = i >> 4;
    s.w1[i] = p.w1[co * 64 + ci];
  }
  if (tid < 64) {
    const int co = tid & 3, ci = tid >> 2;
    s.w10[tid] = p.w10[co * 16 + ci];
  }
  if (tid < 16) { s.b1[tid] = p.b1[tid]; s.s1[tid] = p.s1[tid]; s.t1[tid] = p.t1[tid]; }
  if (tid < 4) s.b10[tid] = p.b10[tid];
  __syncthreads();
  const bool rb = FMT == kFmtBf16;
  const long long npix = (long long)p.N * p.h * p.w;
  const int W = 2 * p.w, H = 2 * p.h;
  for (long long pix = blockIdx.x * 256LL + tid; pix < npix; pix += (long long)gridDim.x * 256) {
    float a[16];
#pragma unroll
    for (int co = 0; co < 16; ++co) a[co] = 0.f;
    const uint8_t* a0 = reinterpret_cast<const uint8_t*>(p.a.p[0]) + pix * 128;
    const uint8_t* a1 = FMT != kFmtBf16 ? reinterpret_cast<const uint8_t*>(p.a.p[1]) + pix * 128 : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 hv = ldg16(a0 + 16 * j);
      uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
      float xin[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        xin[2 * e] = hi_lo_to_f32(hw[e], FMT);
        xin[2 * e + 1] = hi_hi_to_f32(hw[e], FMT);
      }
      if (a1) {
        const uint4 lv = ldg16(a1 + 16 * j);
        uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xin[2 * e] += lo_lo_to_f32(lw[e], FMT);
          xin[2 * e + 1] += lo_hi_to_f32(lw[e], FMT);
        }
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4* wv = reinterpret_cast<const float4*>(&s.w1[(8 * j + e) * 16]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = wv[q];
          a[4 * q] = fmaf(xin[e], ww.x, a[4 * q]); a[4 * q + 1] = fmaf(xin[e], ww.y, a[4 * q + 1]);
          a[4 * q + 2] = fmaf(xin[e], ww.z, a[4 * q + 2]); a[4 * q + 3] = fmaf(xin[e], ww.w, a[4 * q + 3]);
        }
      }
    }
    float c10[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ci = 0; ci < 16; ++ci) {
      float v = a[ci] + s.b1[ci];
      if (rb) v = rbf(v);
      v = fmaf(v, s.s1[ci], s.t1[ci]);
      if (rb) v = rbf(v);
      v = lrelu02(v);
      if (rb) v = rbf(v);
      const float4 ww = *reinterpret_cast<const float4*>(&s.w10[ci * 4]);
      c10[0] = fmaf(v, ww.x, c10[0]); c10[1] = fmaf(v, ww.y, c10[1]);
      c10[2] = fmaf(v, ww.z, c10[2]); c10[3] = fmaf(v, ww.w, c10[3]);
    }
    float r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float v = c10[k] + s.b10[k];
      if (rb) v = rbf(v);
      v = 1.f / (1.f + expf(-v));
      r[k] = rb ? rbf(v) : v;
    }
    // pixel_shuffle(2): channel k = dy*2 + dx -> (2y+dy, 2x+dx)
    const unsigned pix32 = (unsigned)pix;           // npix < 2^31 (checked on the host): 32-bit div/mod
    const unsigned t = pix32 / (unsigned)p.w;
    const int x = int(pix32 - t * (unsigned)p.w);
    const unsigned n = t / (unsigned)p.h;
    const int y = int(t - n * (unsigned)p.h);
    const size_t oo = ((size_t)n * H + 2 * y) * W + 2 * x;
    if (p.y) {
      *reinterpret_cast<float2*>(p.y + oo) = make_float2(r[0], r[1]);
      *reinterpret_cast<float2*>(p.y + oo + W) = make_float2(r[2], r[3]);
    }
    if (p.y_u8) {   // (out * 255).astype(uint8): truncation toward zero of a value in [0, 255]
      *reinterpret_cast<uchar2*>(p.y_u8 + oo) = make_uchar2((unsigned char)(r[0] * 255.f), (unsigned char)(r[1] * 255.f));
      *reinterpret_cast<uchar2*>(p.y_u8 + oo + W) = make_uchar2((unsigned char)(r[2] * 255.f), (unsigned char)(r[3] * 255.f));
    }
  }
}

int tail_eval(const TailParams& p, cudaStream_t st) {
  const long long npix = (long long)p.N * p.h * p.w;
  if (npix >= (1LL << 31)) {
    set_error("tail_eval: too many pixels");
    return 1;
  }
  {
    if (p.fmt == kFmtBf16) tail_eval_kernel<kFmtBf16><<<grid_for(npix, 256, 148 * 8), 256, 0, st>>>(p);
    else if (p.fmt == kFmtF16x2) tail_eval_kernel<kFmtF16x2><<<grid_for(npix, 256, 148 * 8), 256, 0, st>>>(p);
    else tail_eval_kernel<kFmtBf16x2><<<grid_for(npix, 256, 148 * 8), 256, 0, st>>>(p);
  }
  NSM_CHECK_LAUNCH("tail_eval");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// up-sample x2 (align_corners) then resize to (hd, wd) (align_corners); NHWC planes, 8 channels / thread
// ------------------------------------------------------------------------------------------------
struct UpParams {
  const uint8_t* s0;
  const uint8_t* s1;
  uint8_t* d0;
  uint8_t* d1;
  int N, hs, ws, C, hd, wd, fmt;
};

template <int FMT>
__device__ __forceinline__ void load8(const UpParams& p, size_t elem, float* v) {
  const uint4 hv = ldg16(p.s0 + elem * 2);
  const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = hi_lo_to_f32(hw[e], FMT);
    v[2 * e + 1] = hi_hi_to_f32(hw[e], FMT);
  }
  if (FMT != kFmtBf16) {
    const uint4 lv = ldg16(p.s1 + elem * 2);
    const uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] += lo_lo_to_f32(lw[e], FMT);
      v[2 * e + 1] += lo_hi_to_f32(lw[e], FMT);
    }
  }
}

// value of the x2 up-sampled tensor at intermediate pixel (row lerp ly given, column X), 8 channels
template <int FMT>
__device__ __forceinline__ void up2_at(const UpParams& p, size_t nbase, const Lerp& ly, int X, int cg, bool rb,
                                       float* out) {
  const Lerp lx = make_lerp(X, p.ws, 2 * p.ws);
  float v00[8], v01[8], v10[8], v11[8];
  const size_t r0 = nbase + (size_t)ly.i0 * p.ws, r1 = nbase + (size_t)ly.i1 * p.ws;
  load8<FMT>(p, (r0 + lx.i0) * p.C + cg * 8, v00);
  load8<FMT>(p, (r0 + lx.i1) * p.C + cg * 8, v01);
  load8<FMT>(p, (r1 + lx.i0) * p.C + cg * 8, v10);
  load8<FMT>(p, (r1 + lx.i1) * p.C + cg * 8, v11);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float v = ly.w0 * (lx.w0 * v00[e] + lx.w1 * v01[e]) + ly.w1 * (lx.w0 * v10[e] + lx.w1 * v11[e]);
    out[e] = rb ? rbf(v) : v;
  }
}

// grid: x = output row (n * hd + y), y = 256-thread chunks of the row's (pixel, channel-group) pairs.  All index
// arithmetic is 32-bit with shifts (channel-group counts are powers of two); the row interpolation is per block.
template <int FMT>
__global__ void __launch_bounds__(256) upsample_match_kernel(const UpParams p, int cg_shift) {
  const int cgs = 1 << cg_shift;
  const int j = blockIdx.y * 256 + threadIdx.x;
  const bool active = j < p.wd * cgs;   // no early return: the block synchronises on its shared tap tables
  const int cg = j & (cgs - 1), x = active ? (j >> cg_shift) : 0;
  const int n = blockIdx.x / p.hd, y = blockIdx.x - n * p.hd;
  const bool rb = FMT == kFmtBf16;
  const bool same = (p.hd == 2 * p.hs) && (p.wd == 2 * p.ws);
  const size_t nbase = (size_t)n * p.hs * p.ws;
  // block-shared composite taps: the row's (one per block) and those of the <= 256/cgs + 1 columns this block touches
  __shared__ Tap3 s_ty, s_tx[260];
  const int x_first = (blockIdx.y * 256) >> cg_shift;
  if (!same) {
    const int x_count = ((blockIdx.y * 256 + 255) >> cg_shift) - x_first + 1;
    if (threadIdx.x == 0) s_ty = composite_taps(y, p.hs, p.hd);
    if ((int)threadIdx.x < x_count && x_first + (int)threadIdx.x < p.wd)
      s_tx[threadIdx.x] = composite_taps(x_first + threadIdx.x, p.ws, p.wd);
    __syncthreads();
  }
  if (!active) return;   // (only after the barrier; inactive threads must not index the tap tables)
  float r[8];
  if (same) {
    up2_at<FMT>(p, nbase, make_lerp(y, p.hs, 2 * p.hs), x, cg, rb, r);  // second resize has scale 1 -> exact copy
  } else {
    // composite of the two resizes = separable stencil over <= 3x3 source pixels (resample.cuh); evaluated in fp32 with
    // one final rounding (the x2 intermediate is never materialised).  Taps come from the block's shared tables.
    const Tap3 ty = s_ty, tx = s_tx[x - x_first];
#pragma unroll
    for (int e = 0; e < 8; ++e) r[e] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (ty.w[i] == 0.f) continue;
      const size_t rbase = nbase + (size_t)(ty.rmin + i) * p.ws;
      float row[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) row[e] = 0.f;
#pragma unroll
      for (int jx = 0; jx < 3; ++jx) {
        if (tx.w[jx] == 0.f) continue;
        float v[8];
        load8<FMT>(p, (rbase + tx.rmin + jx) * p.C + cg * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) row[e] = fmaf(tx.w[jx], v[e], row[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = fmaf(ty.w[i], row[e], r[e]);
    }
    if (rb) {
#pragma unroll
      for (int e = 0; e < 8; ++e) r[e] = rbf(r[e]);
    }
  }
  const size_t o = (((size_t)n * p.hd + y) * p.wd + x) * p.C + cg * 8;
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hw[e] = pack_hi(r[2 * e], r[2 * e + 1], FMT);
    lw[e] = pack_lo_resid(r[2 * e], r[2 * e + 1], hw[e], FMT);
  }
  stg16(p.d0 + o * 2, make_uint4(hw[0], hw[1], hw[2], hw[3]));
  if (FMT != kFmtBf16) stg16(p.d1 + o * 2, make_uint4(lw[0], lw[1], lw[2], lw[3]));
}

// Fast path of the plain x2 up-sample (destination exactly 2hs x 2ws): one thread produces a 2x2 output block of 8
// channels from the 3x3 source neighbourhood (rows/cols {b-1, b, b+1} clamped).  With align_corners the even output row
// 2b interpolates source rows (b-1, b) and the odd row 2b+1 rows (b, b+1) -- see make_lerp: src = dst*(hs-1)/(2hs-1) --
// so all register indices are static; horizontal interpolation is done once per source row (separable).
__device__ __forceinline__ void up2x_weights(int b, int in_size, float& we0, float& we1, float& wo0, float& wo1) {
  const int out_size = 2 * in_size;
  const float scale = out_size > 1 ? float(in_size - 1) / float(out_size - 1) : 0.f;
  const float se = scale * float(2 * b), so = scale * float(2 * b + 1);
  // even: rows (b-1, b); ATen's (i0, lambda) may be (b, 0) when se rounds to b: identical value with weights (0, 1)
  we1 = b == 0 ? 1.f : fminf(fmaxf(se - float(b - 1), 0.f), 1.f);
  we0 = 1.f - we1;
  wo1 = b >= in_size - 1 ? 0.f : fminf(fmaxf(so - float(b), 0.f), 1.f);
  wo0 = 1.f - wo1;
}

template <int FMT>
__global__ void __launch_bounds__(256) upsample2x_kernel(const UpParams p, int cg_shift) {
  const int cgs = 1 << cg_shift;
  const int j = blockIdx.y * 256 + threadIdx.x;
  if (j >= p.ws * cgs) return;
  const int cg = j & (cgs - 1), xb = j >> cg_shift;
  const int n = blockIdx.x / p.hs, yb = blockIdx.x - n * p.hs;
  const bool rb = FMT == kFmtBf16;
  float wye0, wye1, wyo0, wyo1, wxe0, wxe1, wxo0, wxo1;
  up2x_weights(yb, p.hs, wye0, wye1, wyo0, wyo1);
  up2x_weights(xb, p.ws, wxe0, wxe1, wxo0, wxo1);
  const int rows[3] = {yb > 0 ? yb - 1 : 0, yb, yb < p.hs - 1 ? yb + 1 : p.hs - 1};
  const int cols[3] = {xb > 0 ? xb - 1 : 0, xb, xb < p.ws - 1 ? xb + 1 : p.ws - 1};
  float he[3][8], ho[3][8];   // horizontally interpolated source rows for the even / odd output column
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const size_t rbase = ((size_t)n * p.hs + rows[r]) * p.ws;
    float a[8], b[8], c[8];
    load8<FMT>(p, (rbase + cols[0]) * p.C + cg * 8, a);
    load8<FMT>(p, (rbase + cols[1]) * p.C + cg * 8, b);
    load8<FMT>(p, (rbase + cols[2]) * p.C + cg * 8, c);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      he[r][e] = wxe0 * a[e] + wxe1 * b[e];
      ho[r][e] = wxo0 * b[e] + wxo1 * c[e];
    }
  }
#pragma unroll
  for (int oy = 0; oy < 2; ++oy) {
    const float w0 = oy ? wyo0 : wye0, w1 = oy ? wyo1 : wye1;
#pragma unroll
    for (int ox = 0; ox < 2; ++ox) {
      float r[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float top = ox ? ho[oy][e] : he[oy][e], bot = ox ? ho[oy + 1][e] : he[oy + 1][e];
        const float v = w0 * top + w1 * bot;
        r[e] = rb ? rbf(v) : v;
      }
      const size_t o = (((size_t)n * p.hd + 2 * yb + oy) * p.wd + 2 * xb + ox) * p.C + cg * 8;
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hw[e] = pack_hi(r[2 * e], r[2 * e + 1], FMT);
        lw[e] = pack_lo_resid(r[2 * e], r[2 * e + 1], hw[e], FMT);
      }
      stg16(p.d0 + o * 2, make_uint4(hw[0], hw[1], hw[2], hw[3]));
      if (FMT != kFmtBf16) stg16(p.d1 + o * 2, make_uint4(lw[0], lw[1], lw[2], lw[3]));
    }
  }
}

int upsample_match(const Planes& src, int N, int hs, int ws, int C, const Planes& dst, int hd, int wd, int fmt,
                   cudaStream_t st) {
  const int cgs = C / 8;
  if (C % 8 || (cgs & (cgs - 1))) {
    set_error("upsample_match: C=%d must be 8 x a power of two", C);
    return 1;
  }
  int shift = 0;
  while ((1 << shift) < cgs) ++shift;
  UpParams p;
  p.s0 = (const uint8_t*)src.p[0]; p.s1 = (const uint8_t*)src.p[1];
  p.d0 = (uint8_t*)dst.p[0]; p.d1 = (uint8_t*)dst.p[1];
  p.N = N; p.hs = hs; p.ws = ws; p.C = C; p.hd = hd; p.wd = wd; p.fmt = fmt;
  if (hd == 2 * hs && wd == 2 * ws) {
    dim3 grid((unsigned)(N * hs), (unsigned)((ws * cgs + 255) / 256));
    {
      if (fmt == kFmtBf16) upsample2x_kernel<kFmtBf16><<<grid, 256, 0, st>>>(p, shift);
      else if (fmt == kFmtF16x2) upsample2x_kernel<kFmtF16x2><<<grid, 256, 0, st>>>(p, shift);
      else upsample2x_kernel<kFmtBf16x2><<<grid, 256, 0, st>>>(p, shift);
    }
  } else {
    dim3 grid((unsigned)(N * hd), (unsigned)((wd * cgs + 255) / 256));
    {
      if (fmt == kFmtBf16) upsample_match_kernel<kFmtBf16><<<grid, 256, 0, st>>>(p, shift);
      else if (fmt == kFmtF16x2) upsample_match_kernel<kFmtF16x2><<<grid, 256, 0, st>>>(p, shift);
      else upsample_match_kernel<kFmtBf16x2><<<grid, 256, 0, st>>>(p, shift);
    }
  }
  NSM_CHECK_LAUNCH("upsample_match");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// objective: L1 (+ perturbation L1) value and gradient in one pass
// ------------------------------------------------------------------------------------------------
struct LossParams {
  const float* out;
  const float* target;
  const float* pert[4];
  int n_pert;
  long long numel;
  float coef_l1, coef_pert;
  float* grad;
  double* acc;
};
__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) l1_loss_kernel(const LossParams p) {
  float s_l1 = 0.f, s_p = 0.f, bad = 0.f;
  const long long n4 = p.numel >> 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    const float4 o = __ldg(reinterpret_cast<const float4*>(p.out) + i);
    const float ov[4] = {o.x, o.y, o.z, o.w};
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (p.target) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p.target) + i);
      const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = ov[e] - tv[e];
        s_l1 += fabsf(d);
        g[e] = p.coef_l1 * sgn(d);
      }
    }
    for (int k = 0; k < p.n_pert; ++k) {
      const float4 y = __ldg(reinterpret_cast<const float4*>(p.pert[k]) + i);
      const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d = ov[e] - yv[e];
        s_p += fabsf(d);
        g[e] += p.coef_pert * sgn(d);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) bad += (ov[e] < 0.f || ov[e] > 1.f || ov[e] != ov[e]) ? 1.f : 0.f;
    if (p.grad) reinterpret_cast<float4*>(p.grad)[i] = make_float4(g[0], g[1], g[2], g[3]);
  }
  // scalar tail
  for (long long i = (n4 << 2) + blockIdx.x * 256LL + threadIdx.x; i < p.numel; i += (long long)gridDim.x * 256) {
    const float o = p.out[i];
    float g = 0.f;
    if (p.target) {
      const float d = o - p.target[i];
      s_l1 += fabsf(d);
      g = p.coef_l1 * sgn(d);
    }
    for (int k = 0; k < p.n_pert; ++k) {
      const float d = o - p.pert[k][i];
      s_p += fabsf(d);
      g += p.coef_pert * sgn(d);
    }
    bad += (o < 0.f || o > 1.f || o != o) ? 1.f : 0.f;
    if (p.grad) p.grad[i] = g;
  }
  __shared__ float red[3][8];
  s_l1 = warp_sum(s_l1); s_p = warp_sum(s_p); bad = warp_sum(bad);
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  if (lane == 0) { red[0][wrp] = s_l1; red[1][wrp] = s_p; red[2][wrp] = bad; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += double(red[threadIdx.x][k]);
    if (t != 0.0) atomicAdd(&p.acc[threadIdx.x], t);
  }
}

int l1_loss_fwd_bwd(const float* out, const float* target, const float* const* perturbed, int n_perturbed,
                    long long numel, float coef_l1, float coef_pert, float* grad, double* acc, cudaStream_t st) {
  if (n_perturbed < 0 || n_perturbed > 4) {
    set_error("l1_loss: at most 4 perturbed outputs per launch (got %d)", n_perturbed);
    return 1;
  }
  LossParams p;
  p.out = out; p.target = target; p.n_pert = n_perturbed; p.numel = numel;
  for (int k = 0; k < 4; ++k) p.pert[k] = k < n_perturbed ? perturbed[k] : nullptr;
  p.coef_l1 = coef_l1; p.coef_pert = coef_pert; p.grad = grad; p.acc = acc;
  l1_loss_kernel<<<grid_for((numel + 3) / 4, 256, 148 * 8), 256, 0, st>>>(p);
  NSM_CHECK_LAUNCH("l1_loss");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// per-channel sums over [S][C][HW] in fp64 (two-pass statistics like calculate_dataset_stats.py)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) channel_sums_kernel(const float* __restrict__ x, long long S, int C,
                                                           long long HW, const double* __restrict__ means,
                                                           double* __restrict__ sums, int chunks) {
  // blockIdx.x = (plane index s*C + c) * chunks + chunk
  const long long plane = blockIdx.x / chunks;
  const int chunk = blockIdx.x % chunks;
  const int c = int(plane % C);
  const float* base = x + plane * HW;
  const long long per = (HW + chunks - 1) / chunks;
  const long long lo = chunk * per, hi = (lo + per < HW) ? lo + per : HW;
  const double mu = means ? means[c] : 0.0;
  double acc = 0.0;
  const float* b = base + lo;
  const long long n = hi - lo;
  const long long nv = ((reinterpret_cast<uintptr_t>(b) & 15) == 0) ? (n >> 2) : 0;  // 16-byte vector part
  for (long long k = threadIdx.x; k < nv; k += 256) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(b) + k);
    if (means) {
      const double a0 = double(v.x) - mu, a1 = double(v.y) - mu, a2 = double(v.z) - mu, a3 = double(v.w) - mu;
      acc += (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    } else {
      acc += (double(v.x) + double(v.y)) + (double(v.z) + double(v.w));
    }
  }
  for (long long j = (nv << 2) + threadIdx.x; j < n; j += 256) {  // ragged end / unaligned planes
    const double v = double(b[j]) - mu;
    acc += means ? v * v : v;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(&sums[c], t);
  }
}

int channel_sums(const float* x, long long S, int C, long long HW, const double* means, double* sums,
                 cudaStream_t st) {
  const long long planes = S * C;
  int chunks = int((148LL * 8 + planes - 1) / planes);
  if (chunks < 1) chunks = 1;
  const long long max_chunks = (HW + 4095) / 4096;
  if (chunks > max_chunks) chunks = int(max_chunks);
  if (planes * chunks > 0x7fffffffLL) {
    set_error("channel_sums: too many blocks");
    return 1;
  }
  channel_sums_kernel<<<(unsigned)(planes * chunks), 256, 0, st>>>(x, S, C, HW, means, sums, chunks);
  NSM_CHECK_LAUNCH("channel_sums");
  return 0;
}

__global__ void __launch_bounds__(256) standardize_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          long long planes, int C, long long HW,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ std) {
  const long long total = planes * HW;
  const bool vec = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
  if (vec) {
    const long long n4 = total >> 2;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
      const int c = int(((i << 2) / HW) % C);
      const float mu = mean[c], sd = std[c] + 1e-8f;
      float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
      v.x = (v.x - mu) / sd; v.y = (v.y - mu) / sd; v.z = (v.z - mu) / sd; v.w = (v.w - mu) / sd;
      reinterpret_cast<float4*>(y)[i] = v;
    }
  } else {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
      const int c = int((i / HW) % C);
      y[i] = (x[i] - mean[c]) / (std[c] + 1e-8f);
    }
  }
}
int standardize(const float* x, float* y, long long S, int C, long long HW, const float* mean, const float* std,
                cudaStream_t st) {
  standardize_kernel<<<grid_for(S * C * HW / 4 + 1, 256), 256, 0, st>>>(x, y, S * C, C, HW, mean, std);
  NSM_CHECK_LAUNCH("standardize");
  return 0;
}

__global__ void __launch_bounds__(256) perturb_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                      float* __restrict__ out, int count, long long per_copy, int C,
                                                      long long HW, const float* __restrict__ stds,
                                                      float std_factor) {
  const long long total = per_copy * count;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long j = i % per_copy;              // offset inside one [B][C][HW] copy
    const long long copy = i / per_copy;
    const long long hw = j % HW;
    const int c = int((j / HW) % C);
    const long long b = j / (HW * C);
    const long long B = per_copy / (HW * C);
    // noise is laid out [count][C][B][HW]: each (copy, channel) block is one contiguous randn_like draw
    const float nz = noise[((copy * C + c) * B + b) * HW + hw];
    out[i] = x[j] + (nz * stds[c]) * std_factor;   // pert_loss.py:54-55 evaluation order
  }
}
int perturb(const float* x, const float* noise, float* out, int count, long long B, int C, long long HW,
            const float* stds, float std_factor, cudaStream_t st) {
  const long long per = B * C * HW;
  perturb_kernel<<<grid_for(per * count, 256), 256, 0, st>>>(x, noise, out, count, per, C, HW, stds, std_factor);
  NSM_CHECK_LAUNCH("perturb");
  return 0;
}

}  // namespace nsm
