// HBM-streaming kernels (see stream_kernels.cuh): layout / packing, up-sampling, loss, statistics -- vectorised 16-byte
// accesses, one pass over each tensor -- and the 16-channel head / tail stages of the network, whose two small
// convolutions each run on the warp-level tensor path (mma.sync) inside one kernel.
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
    if (fmt == kFmtF16X8) {   // fp16 hi plane + cross plane [(w - hi) * 2^16 | w * 2^6] per 16 inner channels
      const uint32_t hw = pack_hi(v, 0.f, fmt);
      hi[i] = (unsigned short)(hw & 0xffffu);
      uint8_t* row = reinterpret_cast<uint8_t*>(lo) + (i - col) * 2;
      row[x8_byte(col)] = (uint8_t)(pack_e4m3x2((v - f16lo_to_f32(hw)) * kX8WgtLo, 0.f) & 0xffu);
      row[x8_byte(col) + 16] = (uint8_t)(pack_e4m3x2(v * kX8WgtHi, 0.f) & 0xffu);
      continue;
    }
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
      const float v = tile[threadIdx.x][i];
      if (fmt == kFmtF16X8) {   // cross plane [v * 2 | (v - hi) * 2^11] per 16 channels
        const uint32_t hw = pack_hi(v, 0.f, fmt);
        hi[o] = (unsigned short)(hw & 0xffffu);
        uint8_t* row = reinterpret_cast<uint8_t*>(lo) + (o - c) * 2;
        row[x8_byte(c)] = (uint8_t)(pack_e4m3x2(v * kX8ActHi, 0.f) & 0xffu);
        row[x8_byte(c) + 16] = (uint8_t)(pack_e4m3x2((v - f16lo_to_f32(hw)) * kX8ActLo, 0.f) & 0xffu);
        continue;
      }
      unsigned short h, l;
      split_fmt(v, fmt, h, l);
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
      if (fmt == kFmtF16X8) {   // hi + the residual byte of the cross plane
        const uint8_t* row = reinterpret_cast<const uint8_t*>(lo) + (o - c) * 2;
        v = hi_lo_to_f32(hi[o], fmt) + e4m3_to_f32(row[x8_byte(c) + 16]) * (1.f / kX8ActLo);
      } else {
        v = join_fmt(hi[o], fmt != kFmtBf16 ? lo[o] : (unsigned short)0, fmt);
      }
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
constexpr int HEAD_TH = 16, HEAD_TW = 32;             // output tile (rows x cols) at (h, w) resolution per 256-thread block
constexpr int HEAD_HW = HEAD_TW + 2, HEAD_HH = HEAD_TH + 2;
constexpr int HEAD_PIX = HEAD_HW * HEAD_HH;          // halo tile pixels

// Both convolutions of the head run on the warp-level tensor-core path (mma.sync m16n8k16, fp32 accumulate): with 16
// channels the layer is far too thin for a 128-wide tcgen05 tile, and as fp32 FMAs it was issue-bound at 13.7 k
// instructions per warp (profiles/: 0.26 ms against an HBM floor of 0.03 ms).  The hi+lo modes issue the same three
// products as the big kernel (hi*hi + hi*lo + lo*hi).
// Shared-memory operands are 16-bit, 32 bytes (16 channels) per row, the two 16-byte halves XOR-swizzled with bit 2 of the
// row index so that ldmatrix reads 8 consecutive rows without bank conflicts.
struct HeadWeights {
  uint4 w0[2][9 * 16 * 2];     // [plane][tap][co][half of ci]
  uint4 w1[2][64 * 2];         // [plane][co][half of ci]
  float b0[16], s0[16], t0[16];
  float b1[64], s1[64], t1[64];
};
struct HeadSmem {
  // per-warp staging for the TMA stores of one 16-channel slice: out hi / lo (2 rows x 32 px x 32 B each), pooled hi / lo
  // (16 px x 32 B each); every sub-tile starts on a 256-byte boundary (32-byte swizzle = the XOR of head_row_off)
  uint4 stage[8][320];
  uint4 xt[2][HEAD_PIX * 2];   // [plane][halo pixel][half]   un-shuffled (standardised, even-fixed) input
  HeadWeights wt;              // copied verbatim from the image head_pack_image() built at pack time
};
static_assert(sizeof(HeadWeights) == kHeadImageBytes, "head weight image size");
__device__ __forceinline__ uint32_t head_row_off(int row, int half) {
  return uint32_t(row) * 32u + (uint32_t((half ^ (row >> 2)) & 1) << 4);
}

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

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
template <int FMT>
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (FMT == kFmtF16x2) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
// d += a * w with the operand planes of the storage format (1 product in bf16 mode, 3 in the hi+lo modes)
template <int FMT>
__device__ __forceinline__ void head_mma(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                         uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816<FMT>(d, ah, bh0, bh1);
  if (FMT != kFmtBf16) {
    mma16816<FMT>(d, ah, bl0, bl1);
    mma16816<FMT>(d, al, bh0, bh1);
  }
}
__device__ __forceinline__ void sts8(uint32_t saddr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts2(uint32_t saddr, unsigned short v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(saddr), "h"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 = zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_saddr), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void sts4(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

template <int FMT>
__global__ void __launch_bounds__(256, 2)
head_eval_kernel(const __grid_constant__ HeadParams p, const __grid_constant__ CUtensorMap tmC0,
                 const __grid_constant__ CUtensorMap tmC1, const __grid_constant__ CUtensorMap tmP0,
                 const __grid_constant__ CUtensorMap tmP1, int has_pool) {
  extern __shared__ uint8_t head_smem_raw[];
  const uint32_t raw_addr = smem_u32(head_smem_raw);
  HeadSmem& s = *reinterpret_cast<HeadSmem*>(head_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr));
  constexpr bool rb = FMT == kFmtBf16;
  constexpr bool two = FMT != kFmtBf16;
  const int H = p.Hin - (p.Hin & 1), W = p.Win - (p.Win & 1);
  const int h = H >> 1, w = W >> 1;
  const bool resize = (p.Hin & 1) || (p.Win & 1);
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * HEAD_TH, x0 = blockIdx.x * HEAD_TW;
  const int tid = threadIdx.x;
  const uint32_t xt0 = smem_u32(&s.xt[0][0]), xt1 = smem_u32(&s.xt[1][0]);
  const uint32_t w00 = smem_u32(&s.wt.w0[0][0]), w01 = smem_u32(&s.wt.w0[1][0]);
  const uint32_t w10 = smem_u32(&s.wt.w1[0][0]), w11 = smem_u32(&s.wt.w1[1][0]);

  // ---- weights: the ready-made operand image (split, swizzled at pack time) ----------------------------------------------
  //      asynchronous copies: they fly while the input tile below is fetched and converted
  {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(p.img);
    const uint32_t dst = smem_u32(&s.wt);
    for (int i = tid; i < kHeadImageBytes / 16; i += 256) cp_async16(dst + i * 16, src + i * 16, true);
  }

  // ---- input: one item = one halo pixel x one of the 4 input channels = a 2x2 full-resolution quad = 4 of the 16
  //      un-shuffled channels (ch = c*4 + dy*2 + dx), zero outside the (even-fixed) image --------------------------------
  {
    const bool vec_ok = !resize && ((reinterpret_cast<uintptr_t>(p.x) & 7) == 0);
    const float* xn = p.x + (size_t)n * 4 * p.Hin * p.Win;
    constexpr int kItems = HEAD_PIX * 4, kBatch = 5;
    for (int base = tid; base < kItems; base += 256 * kBatch) {
      float v[kBatch][4];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int item = base + u * 256;
        v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
        if (item < kItems) {
          const int c = item & 3, pi = item >> 2;
          const int ry = pi / HEAD_HW, cx = pi - ry * HEAD_HW;
          const int y = y0 - 1 + ry, x = x0 - 1 + cx;
          if (y >= 0 && y < h && x >= 0 && x < w) {
            if (vec_ok) {
              const float* q = xn + ((size_t)c * p.Hin + 2 * y) * p.Win + 2 * x;
              const float2 r0 = __ldg(reinterpret_cast<const float2*>(q));
              const float2 r1 = __ldg(reinterpret_cast<const float2*>(q + p.Win));
              v[u][0] = r0.x; v[u][1] = r0.y; v[u][2] = r1.x; v[u][3] = r1.y;
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k) v[u][k] = head_fetch(p, n, c, 2 * y + (k >> 1), 2 * x + (k & 1), H, W, resize);
            }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int item = base + u * 256;
        if (item < kItems) {
          const int c = item & 3, pi = item >> 2;
          if (vec_ok && p.mean) {
            const int ry = pi / HEAD_HW, cx = pi - ry * HEAD_HW;
            const int y = y0 - 1 + ry, x = x0 - 1 + cx;
            if (y >= 0 && y < h && x >= 0 && x < w) {   // the zero padding is applied AFTER standardisation
              const float mu = p.mean[c], sd = p.std[c] + 1e-8f;
#pragma unroll
              for (int k = 0; k < 4; ++k) v[u][k] = (v[u][k] - mu) / sd;
            }
          }
          if (rb) {
#pragma unroll
            for (int k = 0; k < 4; ++k) v[u][k] = rbf(v[u][k]);   // autocast casts the conv input to bf16
          }
          const uint32_t h0 = pack_hi(v[u][0], v[u][1], FMT), h1 = pack_hi(v[u][2], v[u][3], FMT);
          const uint32_t off = head_row_off(pi, c >> 1) + (c & 1) * 8;
          sts8(xt0 + off, h0, h1);
          if (two) sts8(xt1 + off, pack_lo_resid(v[u][0], v[u][1], h0, FMT), pack_lo_resid(v[u][2], v[u][3], h1, FMT));
        }
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  const int lane = tid & 31, wrp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;       // mma fragment coordinates
  const int lm = lane >> 3, lr = lane & 7;     // ldmatrix: this lane addresses row lr of matrix lm

  if (p.x16.p[0]) {  // optional tap for tests: the 16-channel un-shuffled input of the tile's interior pixels
    for (int i = tid; i < HEAD_TH * HEAD_TW; i += 256) {
      const int ly = i / HEAD_TW, lx = i - ly * HEAD_TW;
      const int y = y0 + ly, x = x0 + lx;
      if (y < h && x < w) {
        const int pi = (ly + 1) * HEAD_HW + lx + 1;
        const size_t o = (((size_t)n * h + y) * w + x) * 16 * 2;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          stg16(reinterpret_cast<uint8_t*>(p.x16.p[0]) + o + 16 * hf, lds16(xt0 + head_row_off(pi, hf)));
          if (two) stg16(reinterpret_cast<uint8_t*>(p.x16.p[1]) + o + 16 * hf, lds16(xt1 + head_row_off(pi, hf)));
        }
      }
    }
  }

  // warp tile: rows 2*wrp, 2*wrp+1 of the block tile x 32 columns = 4 m-tiles of 16 consecutive pixels:
  //   m-tile mt -> row 2*wrp + (mt >> 1), columns (mt & 1)*16 .. +15
  // ---- conv2.conv.0 : 3x3, 16 -> 16 as 9 k-steps (one per tap) of K = 16 input channels --------------------------------
  float acc[4][2][4];
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  {
    const int a_col = lr + 8 * (lm & 1), a_half = lm >> 1;         // A: matrices (rows 0-7 | 8-15) x (k 0-7 | 8-15)
    const int b_co = (lm >> 1) * 8 + lr, b_half = lm & 1;          // B: matrices (n-tile 0 | 1) x (k 0-7 | 8-15)
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int dyy = tap / 3, dxx = tap - dyy * 3;
      uint32_t bh[4], bl[4] = {0, 0, 0, 0};
      const uint32_t boff = head_row_off(tap * 16 + b_co, b_half);
      ldmatrix_x4(bh, w00 + boff);
      if (two) ldmatrix_x4(bl, w01 + boff);
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        const int pi = (2 * wrp + (mt >> 1) + dyy) * HEAD_HW + (mt & 1) * 16 + a_col + dxx;
        const uint32_t aoff = head_row_off(pi, a_half);
        uint32_t ah[4], al[4] = {0, 0, 0, 0};
        ldmatrix_x4(ah, xt0 + aoff);
        if (two) ldmatrix_x4(al, xt1 + aoff);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          head_mma<FMT>(acc[mt][nt], ah, al, bh[2 * nt], bh[2 * nt + 1], bl[2 * nt], bl[2 * nt + 1]);
      }
    }
  }
  // bias + BN + LeakyReLU, then straight into the A fragments of the 1x1 convolution: the accumulator layout of two
  // 8-wide n-tiles (row g / g+8, columns 2t, 2t+1) IS the m16k16 A layout
  uint32_t a1h[4][4], a1l[4][4];
  {
    float pb[2][2], ps[2][2], pt[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int co = nt * 8 + 2 * t + k;
        pb[nt][k] = s.wt.b0[co]; ps[nt][k] = s.wt.s0[co]; pt[nt][k] = s.wt.t0[co];
      }
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = acc[mt][nt][i] + pb[nt][i & 1];
          if (rb) v = rbf(v);
          v = fmaf(v, ps[nt][i & 1], pt[nt][i & 1]);
          if (rb) v = rbf(v);
          v = lrelu02(v);
          if (rb) v = rbf(v);
          acc[mt][nt][i] = v;
        }
#pragma unroll
      for (int k = 0; k < 4; ++k) {   // a0: (g, nt0), a1: (g+8, nt0), a2: (g, nt1), a3: (g+8, nt1)
        const float e0 = acc[mt][k >> 1][(k & 1) * 2], e1 = acc[mt][k >> 1][(k & 1) * 2 + 1];
        a1h[mt][k] = pack_hi(e0, e1, FMT);
        a1l[mt][k] = two ? pack_lo_resid(e0, e1, a1h[mt][k], FMT) : 0u;
      }
    }
  }

  // ---- conv2.conv.4 : 1x1, 16 -> 64, 16 output channels (two n-tiles) at a time; every slice leaves through the warp's
  //      staging tile and TMA stores (which also clip the pixels outside the image) -----------------------------------------
  const int yrow = y0 + 2 * wrp;
  const uint32_t stg = smem_u32(&s.stage[wrp][0]);
#pragma unroll 1
  for (int np = 0; np < 4; ++np) {
    uint32_t bh[4], bl[4] = {0, 0, 0, 0};
    {
      const int co = np * 16 + (lm >> 1) * 8 + lr;
      const uint32_t boff = head_row_off(co, lm & 1);
      ldmatrix_x4(bh, w10 + boff);
      if (two) ldmatrix_x4(bl, w11 + boff);
    }
    float o[4][2][4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) o[mt][nt][i] = 0.f;
        head_mma<FMT>(o[mt][nt], a1h[mt], a1l[mt], bh[2 * nt], bh[2 * nt + 1], bl[2 * nt], bl[2 * nt + 1]);
      }
    float pb[2][2], ps[2][2], pt[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int co = np * 16 + nt * 8 + 2 * t + k;
        pb[nt][k] = s.wt.b1[co]; ps[nt][k] = s.wt.s1[co]; pt[nt][k] = s.wt.t1[co];
      }
    if (lane == 0) bulk_wait_read0();   // the previous slice has left the staging tile
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = o[mt][nt][i] + pb[nt][i & 1];
          if (rb) v = rbf(v);
          v = fmaf(v, ps[nt][i & 1], pt[nt][i & 1]);
          if (rb) v = rbf(v);
          v = lrelu02(v);
          if (rb) v = rbf(v);
          o[mt][nt][i] = v;
        }
#pragma unroll
        for (int ih = 0; ih < 2; ++ih) {   // fragment rows g and g+8 = two pixels of the m-tile
          const int row = (mt >> 1) * 32 + (mt & 1) * 16 + g + 8 * ih;
          const uint32_t off = head_row_off(row, nt) + 4 * t;
          const uint32_t hw = pack_hi(o[mt][nt][2 * ih], o[mt][nt][2 * ih + 1], FMT);
          sts4(stg + off, hw);
          if (two) sts4(stg + 2048 + off, pack_lo_resid(o[mt][nt][2 * ih], o[mt][nt][2 * ih + 1], hw, FMT));
        }
      }
    }
    // AvgPool2d(2): vertical partner = m-tile + 2 (same lane), horizontal partner = fragment row g^1 = lane ^ 4
#pragma unroll
    for (int mx = 0; mx < 2; ++mx)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float pl[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float sv = o[mx][nt][i] + o[mx + 2][nt][i];
          sv += __shfl_xor_sync(0xffffffffu, sv, 4);
          sv *= 0.25f;
          pl[i] = rb ? rbf(sv) : sv;
        }
        if ((g & 1) == 0) {
#pragma unroll
          for (int ih = 0; ih < 2; ++ih) {
            const int row = mx * 8 + (g >> 1) + 4 * ih;   // pooled pixel inside the warp's 16-pixel pooled row
            const uint32_t off = head_row_off(row, nt) + 4 * t;
            const uint32_t hw = pack_hi(pl[2 * ih], pl[2 * ih + 1], FMT);
            sts4(stg + 4096 + off, hw);
            if (two) sts4(stg + 4608 + off, pack_lo_resid(pl[2 * ih], pl[2 * ih + 1], hw, FMT));
          }
        }
      }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(&tmC0, stg, np * 16, x0, yrow, n);
      if (two) tma_store_4d(&tmC1, stg + 2048, np * 16, x0, yrow, n);
      if (has_pool) {
        tma_store_4d(&tmP0, stg + 4096, np * 16, x0 >> 1, yrow >> 1, n);
        if (two) tma_store_4d(&tmP1, stg + 4608, np * 16, x0 >> 1, yrow >> 1, n);
      }
      bulk_commit();
    }
  }
  if (lane == 0) bulk_wait0();
}

// Builds the head's shared-memory weight image once per parameter version (nsm_unet_pack).
struct HeadPackArgs {
  const float *w0, *b0, *s0, *t0, *w1, *b1, *s1, *t1;
  int fmt;
  HeadWeights* img;
};
__global__ void head_pack_kernel(const HeadPackArgs a) {
  uint8_t* w0p[2] = {reinterpret_cast<uint8_t*>(&a.img->w0[0][0]), reinterpret_cast<uint8_t*>(&a.img->w0[1][0])};
  uint8_t* w1p[2] = {reinterpret_cast<uint8_t*>(&a.img->w1[0][0]), reinterpret_cast<uint8_t*>(&a.img->w1[1][0])};
  const int tid = threadIdx.x;
  for (int i = tid; i < 9 * 16 * 16; i += blockDim.x) {
    const int ci = i & 15, co = (i >> 4) & 15, tap = i >> 8;
    unsigned short hi, lo;
    split_fmt(a.w0[(co * 16 + ci) * 9 + tap], a.fmt, hi, lo);
    const uint32_t off = head_row_off(tap * 16 + co, ci >> 3) + (ci & 7) * 2;
    *reinterpret_cast<unsigned short*>(w0p[0] + off) = hi;
    *reinterpret_cast<unsigned short*>(w0p[1] + off) = a.fmt == kFmtBf16 ? (unsigned short)0 : lo;
  }
  for (int i = tid; i < 64 * 16; i += blockDim.x) {
    const int ci = i & 15, co = i >> 4;
    unsigned short hi, lo;
    split_fmt(a.w1[co * 16 + ci], a.fmt, hi, lo);
    const uint32_t off = head_row_off(co, ci >> 3) + (ci & 7) * 2;
    *reinterpret_cast<unsigned short*>(w1p[0] + off) = hi;
    *reinterpret_cast<unsigned short*>(w1p[1] + off) = a.fmt == kFmtBf16 ? (unsigned short)0 : lo;
  }
  if (tid < 16) { a.img->b0[tid] = a.b0[tid]; a.img->s0[tid] = a.s0[tid]; a.img->t0[tid] = a.t0[tid]; }
  if (tid < 64) { a.img->b1[tid] = a.b1[tid]; a.img->s1[tid] = a.s1[tid]; a.img->t1[tid] = a.t1[tid]; }
}
int head_pack_image(const float* w0, const float* b0, const float* s0, const float* t0, const float* w1,
                    const float* b1, const float* s1, const float* t1, int fmt, void* img, cudaStream_t st) {
  HeadPackArgs a{w0, b0, s0, t0, w1, b1, s1, t1, fmt, reinterpret_cast<HeadWeights*>(img)};
  head_pack_kernel<<<1, 256, 0, st>>>(a);
  NSM_CHECK_LAUNCH("head_pack_image");
  return 0;
}

template <int FMT>
static int head_launch(const HeadParams& p, const CUtensorMap* maps, int has_pool, dim3 grid, cudaStream_t st) {
  constexpr int kSmem = (int)sizeof(HeadSmem) + 1024;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(head_eval_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) {
      set_error("head_eval: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    attr = true;
  }
  head_eval_kernel<FMT><<<grid, 256, kSmem, st>>>(p, maps[0], maps[1], maps[2], maps[3], has_pool);
  return 0;
}

int head_eval(const HeadParams& p, cudaStream_t st) {
  const int H = p.Hin - (p.Hin & 1), W = p.Win - (p.Win & 1);
  const int h = H / 2, w = W / 2;
  if (h < 1 || w < 1 || p.N < 1 || !p.img) {
    set_error("head_eval: bad shape N=%d Hin=%d Win=%d or missing weight image", p.N, p.Hin, p.Win);
    return 1;
  }
  const int planes = fmt_planes(p.fmt);
  const int hp = h / 2, wp = w / 2;
  const int has_pool = (p.p2.p[0] && hp > 0 && wp > 0) ? 1 : 0;
  CUtensorMap maps[4];   // c2 hi/lo, p2 hi/lo: boxes of 16 channels (32 bytes, 32-byte swizzle)
  memset(maps, 0, sizeof(maps));
  const uint64_t cdims[4] = {64, uint64_t(w), uint64_t(h), uint64_t(p.N)};
  const uint64_t cstr[3] = {128, uint64_t(w) * 128, uint64_t(h) * w * 128};
  const uint32_t cbox[4] = {16, HEAD_TW, 2, 1};
  const uint64_t pdims[4] = {64, uint64_t(wp > 0 ? wp : 1), uint64_t(hp > 0 ? hp : 1), uint64_t(p.N)};
  const uint64_t pstr[3] = {128, uint64_t(wp > 0 ? wp : 1) * 128, uint64_t(hp > 0 ? hp : 1) * (wp > 0 ? wp : 1) * 128};
  const uint32_t pbox[4] = {16, HEAD_TW / 2, 1, 1};
  for (int pl = 0; pl < 2; ++pl) {
    const int src = pl < planes ? pl : 0;
    if (!p.c2.p[src]) {
      set_error("head_eval: null output plane %d", src);
      return 1;
    }
    if (encode_tmap_tiled(&maps[pl], p.c2.p[src], 4, cdims, cstr, cbox, 2, 32)) return 1;
    if (has_pool) {
      if (encode_tmap_tiled(&maps[2 + pl], p.p2.p[src], 4, pdims, pstr, pbox, 2, 32)) return 1;
    } else {
      maps[2 + pl] = maps[pl];
    }
  }
  dim3 grid((w + HEAD_TW - 1) / HEAD_TW, (h + HEAD_TH - 1) / HEAD_TH, p.N);
  int rc;
  if (p.fmt == kFmtBf16) rc = head_launch<kFmtBf16>(p, maps, has_pool, grid, st);
  else if (p.fmt == kFmtF16x2) rc = head_launch<kFmtF16x2>(p, maps, has_pool, grid, st);
  else rc = head_launch<kFmtBf16x2>(p, maps, has_pool, grid, st);
  if (rc) return rc;
  NSM_CHECK_LAUNCH("head_eval");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// tail stage
// ------------------------------------------------------------------------------------------------
// conv9's 1x1 (64 -> 16) and conv10 (16 -> 4, padded to 8 outputs) on mma.sync like the head.  A warp works on 32
// consecutive pixels: their 64 channels are 4 KB contiguous per plane, copied with coalesced 16-byte cp.async into a
// warp-private swizzled tile (pixel rows of 128 B, chunk ^ (pixel & 7)) and read back as A fragments with ldmatrix.  The
// weight B fragments come from an image in per-lane register order built at pack time (tail_pack_image).
struct TailImage {
  uint4 w1f[2][4][32];   // [plane][k-step of 16 ci][lane] = {b0, b1 of n-tile 0, b0, b1 of n-tile 1}
  uint2 w10f[2][32];     // [plane][lane] = {b0, b1} of the single (zero-padded) n-tile
  float b1[16], s1[16], t1[16], b10[4];
};
static_assert(sizeof(TailImage) == kTailImageBytes, "tail weight image size");
constexpr int kTailWarps = 8;


template <int FMT>
__global__ void __launch_bounds__(32 * kTailWarps) tail_eval_kernel(const __grid_constant__ TailParams p) {
  extern __shared__ uint8_t tail_smem_raw[];   // [warp][plane][32 px][128 B]
  constexpr bool rb = FMT == kFmtBf16;
  constexpr bool two = FMT != kFmtBf16;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3, lm = lane >> 3, lr = lane & 7;
  const uint32_t raw_addr = smem_u32(tail_smem_raw);
  const uint32_t tile0 = ((raw_addr + 127u) & ~127u) + wrp * 8192, tile1 = tile0 + 4096;
  const TailImage* img = reinterpret_cast<const TailImage*>(p.img);

  uint32_t wh[4][4], wl[4][4], w10h[2], w10l[2] = {0, 0};
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    const uint4 a = __ldg(&img->w1f[0][ks][lane]);
    wh[ks][0] = a.x; wh[ks][1] = a.y; wh[ks][2] = a.z; wh[ks][3] = a.w;
    if (two) {
      const uint4 b = __ldg(&img->w1f[1][ks][lane]);
      wl[ks][0] = b.x; wl[ks][1] = b.y; wl[ks][2] = b.z; wl[ks][3] = b.w;
    } else {
      wl[ks][0] = wl[ks][1] = wl[ks][2] = wl[ks][3] = 0;
    }
  }
  {
    const uint2 a = __ldg(&img->w10f[0][lane]);
    w10h[0] = a.x; w10h[1] = a.y;
    if (two) {
      const uint2 b = __ldg(&img->w10f[1][lane]);
      w10l[0] = b.x; w10l[1] = b.y;
    }
  }
  float pb[2][2], ps[2][2], pt[2][2], pb10[2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int co = nt * 8 + 2 * t + k;
      pb[nt][k] = __ldg(&img->b1[co]); ps[nt][k] = __ldg(&img->s1[co]); pt[nt][k] = __ldg(&img->t1[co]);
    }
  pb10[0] = t < 2 ? __ldg(&img->b10[2 * t]) : 0.f;
  pb10[1] = t < 2 ? __ldg(&img->b10[2 * t + 1]) : 0.f;

  const long long npix = (long long)p.N * p.h * p.w;
  const long long units = (npix + 31) >> 5;
  const int W = 2 * p.w, H = 2 * p.h;
  const uint8_t* a0 = reinterpret_cast<const uint8_t*>(p.a.p[0]);
  const uint8_t* a1 = reinterpret_cast<const uint8_t*>(p.a.p[1]);
  for (long long unit = (long long)blockIdx.x * kTailWarps + wrp; unit < units; unit += (long long)gridDim.x * kTailWarps) {
    const long long pix0 = unit << 5;
    // 32 px x 128 B per plane: instruction i moves pixels 4i .. 4i+3 (lane -> pixel 4i + lane/8, chunk lane%8)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int px = 4 * i + (lane >> 3), ch = lane & 7;
      const bool ok = pix0 + px < npix;
      const size_t goff = ok ? (size_t)(pix0 + px) * 128 + ch * 16 : 0;
      const uint32_t soff = px * 128 + ((ch ^ (px & 7)) << 4);
      cp_async16(tile0 + soff, a0 + goff, ok);
      if (two) cp_async16(tile1 + soff, a1 + goff, ok);
    }
    cp_async_wait_all();
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float acc[2][4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;
      const int px = mt * 16 + lr + 8 * (lm & 1);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t soff = px * 128 + (((2 * ks + (lm >> 1)) ^ (px & 7)) << 4);
        uint32_t ah[4], al[4] = {0, 0, 0, 0};
        ldmatrix_x4(ah, tile0 + soff);
        if (two) ldmatrix_x4(al, tile1 + soff);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          head_mma<FMT>(acc[nt], ah, al, wh[ks][2 * nt], wh[ks][2 * nt + 1], wl[ks][2 * nt], wl[ks][2 * nt + 1]);
      }
      // bias + BN + LeakyReLU -> A fragments of conv10
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v = acc[nt][i] + pb[nt][i & 1];
          if (rb) v = rbf(v);
          v = fmaf(v, ps[nt][i & 1], pt[nt][i & 1]);
          if (rb) v = rbf(v);
          v = lrelu02(v);
          if (rb) v = rbf(v);
          acc[nt][i] = v;
        }
      uint32_t bh[4], bl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float e0 = acc[k >> 1][(k & 1) * 2], e1 = acc[k >> 1][(k & 1) * 2 + 1];
        bh[k] = pack_hi(e0, e1, FMT);
        bl[k] = two ? pack_lo_resid(e0, e1, bh[k], FMT) : 0u;
      }
      float c10[4] = {0.f, 0.f, 0.f, 0.f};
      head_mma<FMT>(c10, bh, bl, w10h[0], w10h[1], w10l[0], w10l[1]);
      if (t < 2) {   // columns 2t, 2t+1 = channels (dy = t, dx = 0 / 1) of pixel_shuffle(2)
#pragma unroll
        for (int ih = 0; ih < 2; ++ih) {
          const long long pix = pix0 + mt * 16 + g + 8 * ih;
          if (pix < npix) {
            float r[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              float v = c10[2 * ih + k] + pb10[k];
              if (rb) v = rbf(v);
              v = 1.f / (1.f + expf(-v));
              r[k] = rb ? rbf(v) : v;
            }
            const unsigned pix32 = (unsigned)pix;           // npix < 2^31 (checked on the host): 32-bit div/mod
            const unsigned q = pix32 / (unsigned)p.w;
            const int x = int(pix32 - q * (unsigned)p.w);
            const unsigned n = q / (unsigned)p.h;
            const int y = int(q - n * (unsigned)p.h);
            const size_t oo = ((size_t)n * H + 2 * y + t) * W + 2 * x;
            if (p.y) *reinterpret_cast<float2*>(p.y + oo) = make_float2(r[0], r[1]);
            if (p.y_u8)   // (out * 255).astype(uint8): truncation toward zero of a value in [0, 255]
              *reinterpret_cast<uchar2*>(p.y_u8 + oo) =
                  make_uchar2((unsigned char)(r[0] * 255.f), (unsigned char)(r[1] * 255.f));
          }
        }
      }
    }
    __syncwarp();   // all lanes are done with the tile before the next unit's copies overwrite it
  }
}

struct TailPackArgs {
  const float *w1, *b1, *s1, *t1, *w10, *b10;
  int fmt;
  TailImage* img;
};
__global__ void tail_pack_kernel(const TailPackArgs a) {
  const int lane = threadIdx.x;   // one warp
  const int g = lane >> 2, t = lane & 3;
  auto pair = [&](float v0, float v1, uint32_t& hi, uint32_t& lo) {
    hi = pack_hi(v0, v1, a.fmt);
    lo = a.fmt == kFmtBf16 ? 0u : pack_lo_resid(v0, v1, hi, a.fmt);
  };
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t h[4], l[4];
    for (int nt = 0; nt < 2; ++nt)
      for (int kh = 0; kh < 2; ++kh) {   // b0: k = 2t, 2t+1 ; b1: k + 8 ; n = nt*8 + g   (w1 is [16 co][64 ci])
        const int co = nt * 8 + g, ci = ks * 16 + kh * 8 + 2 * t;
        pair(a.w1[co * 64 + ci], a.w1[co * 64 + ci + 1], h[2 * nt + kh], l[2 * nt + kh]);
      }
    a.img->w1f[0][ks][lane] = make_uint4(h[0], h[1], h[2], h[3]);
    a.img->w1f[1][ks][lane] = make_uint4(l[0], l[1], l[2], l[3]);
  }
  {
    uint32_t h[2], l[2];
    for (int kh = 0; kh < 2; ++kh) {   // w10 is [4 co][16 ci]; output columns 4..7 are padding
      const int ci = kh * 8 + 2 * t;
      const float v0 = g < 4 ? a.w10[g * 16 + ci] : 0.f, v1 = g < 4 ? a.w10[g * 16 + ci + 1] : 0.f;
      pair(v0, v1, h[kh], l[kh]);
    }
    a.img->w10f[0][lane] = make_uint2(h[0], h[1]);
    a.img->w10f[1][lane] = make_uint2(l[0], l[1]);
  }
  if (lane < 16) { a.img->b1[lane] = a.b1[lane]; a.img->s1[lane] = a.s1[lane]; a.img->t1[lane] = a.t1[lane]; }
  if (lane < 4) a.img->b10[lane] = a.b10[lane];
}
int tail_pack_image(const float* w1, const float* b1, const float* s1, const float* t1, const float* w10,
                    const float* b10, int fmt, void* img, cudaStream_t st) {
  TailPackArgs a{w1, b1, s1, t1, w10, b10, fmt, reinterpret_cast<TailImage*>(img)};
  tail_pack_kernel<<<1, 32, 0, st>>>(a);
  NSM_CHECK_LAUNCH("tail_pack_image");
  return 0;
}

template <int FMT>
static int tail_launch(const TailParams& p, int grid, cudaStream_t st) {
  constexpr int kSmem = kTailWarps * 8192 + 128;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tail_eval_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) {
      set_error("tail_eval: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return 1;
    }
    attr = true;
  }
  tail_eval_kernel<FMT><<<grid, 32 * kTailWarps, kSmem, st>>>(p);
  return 0;
}

int tail_eval(const TailParams& p, cudaStream_t st) {
  const long long npix = (long long)p.N * p.h * p.w;
  if (npix >= (1LL << 31) || !p.img) {
    set_error("tail_eval: too many pixels or missing weight image");
    return 1;
  }
  const long long units = (npix + 31) / 32;
  long long blocks = (units + kTailWarps - 1) / kTailWarps;
  if (blocks > 148 * 3) blocks = 148 * 3;
  int rc;
  if (p.fmt == kFmtBf16) rc = tail_launch<kFmtBf16>(p, (int)blocks, st);
  else if (p.fmt == kFmtF16x2) rc = tail_launch<kFmtF16x2>(p, (int)blocks, st);
  else rc = tail_launch<kFmtBf16x2>(p, (int)blocks, st);
  if (rc) return rc;
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

// 8 channels of one output pixel -> destination planes.  o = element offset of the first channel (pixel * C + 8 * cg).
// OFMT = kFmtF16X8 writes the fp16 hi plane and the 8-bit cross plane the decoder's 3x3 convolutions consume.
template <int OFMT>
__device__ __forceinline__ void up_store8(const UpParams& p, size_t o, int cg, float (&v)[8]) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (OFMT == kFmtBf16) { v[2 * e] = rbf(v[2 * e]); v[2 * e + 1] = rbf(v[2 * e + 1]); }
    hw[e] = pack_hi(v[2 * e], v[2 * e + 1], OFMT);
    if (OFMT == kFmtF16x2 || OFMT == kFmtBf16x2) lw[e] = pack_lo_resid(v[2 * e], v[2 * e + 1], hw[e], OFMT);
  }
  stg16(p.d0 + o * 2, make_uint4(hw[0], hw[1], hw[2], hw[3]));
  if (OFMT == kFmtF16x2 || OFMT == kFmtBf16x2) stg16(p.d1 + o * 2, make_uint4(lw[0], lw[1], lw[2], lw[3]));
  if (OFMT == kFmtF16X8) {
    uint2 first, second;
    x8_act_bytes(v, hw, first, second);
    uint8_t* row = p.d1 + (o - cg * 8) * 2 + (cg >> 1) * 32 + (cg & 1) * 8;
    *reinterpret_cast<uint2*>(row) = first;
    *reinterpret_cast<uint2*>(row + 16) = second;
  }
}

// Taps of a PAIR of adjacent outputs of the composite resize (hd >= hs, wd >= ws: adjacent outputs then start at most two
// source indices apart, so the pair touches <= 5 consecutive source indices).  Taps an output does not use enter with
// weight 0, so the per-output FMA order is the same as in the one-pixel kernel above and both produce identical bits.
struct PairTaps {
  int lo;          // first source index of the union
  float w[2][5];   // weights of the two outputs on lo .. lo+4
};
__device__ __forceinline__ PairTaps pair_taps(int d0, int in_size, int out_size) {
  const Tap3 a = composite_taps(d0, in_size, out_size);
  const bool has_b = d0 + 1 < out_size;
  const Tap3 b = has_b ? composite_taps(d0 + 1, in_size, out_size) : a;
  PairTaps t;
  t.lo = a.rmin;
  const int sh = b.rmin - a.rmin;   // 0, 1 or 2
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    t.w[0][i] = i < 3 ? a.w[i] : 0.f;
    const int k = i - sh;
    t.w[1][i] = (has_b && k >= 0 && k < 3) ? (k == 0 ? b.w[0] : (k == 1 ? b.w[1] : b.w[2])) : 0.f;
  }
  return t;
}

// Strip-walking form of the composite resize: one thread owns an output column pair (8 channels) and walks down
// kMatchStrip output rows, keeping the three source rows of the current vertical stencil horizontally interpolated in
// registers; moving to the next output row shifts that window by 0, 1 or 2 source rows, so every source row is loaded and
// unpacked once per strip instead of once per output row.  Same FMA order per output as the kernels above.
constexpr int kMatchStrip = 8;

template <int FMT>
__device__ __forceinline__ void match_hrow(const UpParams& p, size_t nbase, int r, const PairTaps& tx, int cg,
                                           float (&h)[2][8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) h[0][e] = h[1][e] = 0.f;
  if (r > p.hs - 1) return;   // only reached with zero vertical weight
  const size_t rbase = nbase + (size_t)r * p.ws;
#pragma unroll
  for (int c = 0; c < 5; ++c) {
    if (tx.w[0][c] == 0.f && tx.w[1][c] == 0.f) continue;
    float v[8];
    load8<FMT>(p, (rbase + tx.lo + c) * p.C + cg * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      h[0][e] = fmaf(tx.w[0][c], v[e], h[0][e]);
      h[1][e] = fmaf(tx.w[1][c], v[e], h[1][e]);
    }
  }
}

template <int FMT, int OFMT>
__global__ void __launch_bounds__(256) upsample_match_strip_kernel(const UpParams p, int cg_shift, int strips,
                                                                   int wpairs) {
  const int cgs = 1 << cg_shift;
  const int j = blockIdx.y * 256 + threadIdx.x;
  const bool active = j < wpairs * cgs;
  const int cg = j & (cgs - 1), xp = active ? (j >> cg_shift) : 0;
  const int n = blockIdx.x / strips, y_first = (blockIdx.x - n * strips) * kMatchStrip;
  const int y_end = min(y_first + kMatchStrip, p.hd);
  __shared__ Tap3 s_ty[kMatchStrip];
  __shared__ PairTaps s_tx[257];
  const int xp_first = (blockIdx.y * 256) >> cg_shift;
  {
    const int count = ((blockIdx.y * 256 + 255) >> cg_shift) - xp_first + 1;
    if ((int)threadIdx.x < y_end - y_first) s_ty[threadIdx.x] = composite_taps(y_first + threadIdx.x, p.hs, p.hd);
    for (int i = threadIdx.x; i < count; i += 256)
      if (xp_first + i < wpairs) s_tx[i] = pair_taps(2 * (xp_first + i), p.ws, p.wd);
    __syncthreads();
  }
  if (!active) return;   // (only after the barrier; inactive threads must not index the tap tables)
  const PairTaps tx = s_tx[xp - xp_first];
  const size_t nbase = (size_t)n * p.hs * p.ws;
  float h0[2][8], h1[2][8], h2[2][8];   // source rows base, base+1, base+2, interpolated to the two output columns
  int base = s_ty[0].rmin;
  match_hrow<FMT>(p, nbase, base, tx, cg, h0);
  match_hrow<FMT>(p, nbase, base + 1, tx, cg, h1);
  match_hrow<FMT>(p, nbase, base + 2, tx, cg, h2);
  for (int y = y_first; y < y_end; ++y) {
    const Tap3 ty = s_ty[y - y_first];
    const int d = ty.rmin - base;   // block-uniform, 0..2 for hd >= hs
    if (d == 1) {
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) { h0[b][e] = h1[b][e]; h1[b][e] = h2[b][e]; }
      match_hrow<FMT>(p, nbase, base + 3, tx, cg, h2);
      base += 1;
    } else if (d >= 2) {
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) h0[b][e] = h2[b][e];
      match_hrow<FMT>(p, nbase, base + 3, tx, cg, h1);
      match_hrow<FMT>(p, nbase, base + 4, tx, cg, h2);
      base += 2;
    }
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int x = 2 * xp + b;
      if (x >= p.wd) continue;
      const size_t o = (((size_t)n * p.hd + y) * p.wd + x) * p.C + cg * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = fmaf(ty.w[0], h0[b][e], 0.f);
        a = fmaf(ty.w[1], h1[b][e], a);
        v[e] = fmaf(ty.w[2], h2[b][e], a);
      }
      up_store8<OFMT>(p, o, cg, v);
    }
  }
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

// One thread walks down a strip of kUpStrip source rows at source column xb: every source row is loaded (3 columns),
// unpacked and interpolated horizontally ONCE and then feeds the two output rows above and the two below it, so a 2x2
// output block costs 3 loads instead of 9.
constexpr int kUpStrip = 4;

template <int FMT>
__device__ __forceinline__ void up2x_hrow(const UpParams& p, size_t rbase, const int (&cols)[3], int cg, float wxe0,
                                          float wxe1, float wxo0, float wxo1, float (&he)[8], float (&ho)[8]) {
  float a[8], b[8], c[8];
  load8<FMT>(p, (rbase + cols[0]) * p.C + cg * 8, a);
  load8<FMT>(p, (rbase + cols[1]) * p.C + cg * 8, b);
  load8<FMT>(p, (rbase + cols[2]) * p.C + cg * 8, c);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    he[e] = wxe0 * a[e] + wxe1 * b[e];
    ho[e] = wxo0 * b[e] + wxo1 * c[e];
  }
}

template <int OFMT>
__device__ __forceinline__ void up2x_store(const UpParams& p, size_t o, int cg, float w0, float w1,
                                           const float (&top)[8], const float (&bot)[8]) {
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = w0 * top[e] + w1 * bot[e];
  up_store8<OFMT>(p, o, cg, v);
}

template <int FMT, int OFMT>
__global__ void __launch_bounds__(256) upsample2x_kernel(const UpParams p, int cg_shift, int strips) {
  const int cgs = 1 << cg_shift;
  const int j = blockIdx.y * 256 + threadIdx.x;
  if (j >= p.ws * cgs) return;
  const int cg = j & (cgs - 1), xb = j >> cg_shift;
  const int n = blockIdx.x / strips, y_first = (blockIdx.x - n * strips) * kUpStrip;
  const int y_end = min(y_first + kUpStrip, p.hs);
  float wxe0, wxe1, wxo0, wxo1;
  up2x_weights(xb, p.ws, wxe0, wxe1, wxo0, wxo1);
  const int cols[3] = {xb > 0 ? xb - 1 : 0, xb, xb < p.ws - 1 ? xb + 1 : p.ws - 1};
  const size_t nrow = (size_t)n * p.hs;
  // horizontally interpolated source rows (even / odd output column): previous, current, next
  float pe[8], po[8], ce[8], co[8], ne[8], no[8];
  up2x_hrow<FMT>(p, (nrow + (y_first > 0 ? y_first - 1 : 0)) * p.ws, cols, cg, wxe0, wxe1, wxo0, wxo1, pe, po);
  up2x_hrow<FMT>(p, (nrow + y_first) * p.ws, cols, cg, wxe0, wxe1, wxo0, wxo1, ce, co);
  for (int yb = y_first; yb < y_end; ++yb) {
    up2x_hrow<FMT>(p, (nrow + (yb < p.hs - 1 ? yb + 1 : p.hs - 1)) * p.ws, cols, cg, wxe0, wxe1, wxo0, wxo1, ne, no);
    float wye0, wye1, wyo0, wyo1;
    up2x_weights(yb, p.hs, wye0, wye1, wyo0, wyo1);
    const size_t o = (((size_t)n * p.hd + 2 * yb) * p.wd + 2 * xb) * p.C + cg * 8;
    const size_t down = (size_t)p.wd * p.C;
    up2x_store<OFMT>(p, o, cg, wye0, wye1, pe, ce);                  // even row: source rows (yb-1, yb)
    up2x_store<OFMT>(p, o + p.C, cg, wye0, wye1, po, co);
    up2x_store<OFMT>(p, o + down, cg, wyo0, wyo1, ce, ne);           // odd row: (yb, yb+1)
    up2x_store<OFMT>(p, o + down + p.C, cg, wyo0, wyo1, co, no);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      pe[e] = ce[e]; po[e] = co[e];
      ce[e] = ne[e]; co[e] = no[e];
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
  // fmt kFmtF16X8: source in the fp16 hi+lo format, destination as fp16 hi plane + 8-bit cross plane
  if (hd == 2 * hs && wd == 2 * ws) {
    const int strips = (hs + kUpStrip - 1) / kUpStrip;
    dim3 grid((unsigned)(N * strips), (unsigned)((ws * cgs + 255) / 256));
    if (fmt == kFmtBf16) upsample2x_kernel<kFmtBf16, kFmtBf16><<<grid, 256, 0, st>>>(p, shift, strips);
    else if (fmt == kFmtF16x2) upsample2x_kernel<kFmtF16x2, kFmtF16x2><<<grid, 256, 0, st>>>(p, shift, strips);
    else if (fmt == kFmtF16X8) upsample2x_kernel<kFmtF16x2, kFmtF16X8><<<grid, 256, 0, st>>>(p, shift, strips);
    else upsample2x_kernel<kFmtBf16x2, kFmtBf16x2><<<grid, 256, 0, st>>>(p, shift, strips);
  } else if (hd >= hs && wd >= ws) {
    const int strips = (hd + kMatchStrip - 1) / kMatchStrip, wpairs = (wd + 1) / 2;
    dim3 grid((unsigned)(N * strips), (unsigned)((wpairs * cgs + 255) / 256));
    if (fmt == kFmtBf16)
      upsample_match_strip_kernel<kFmtBf16, kFmtBf16><<<grid, 256, 0, st>>>(p, shift, strips, wpairs);
    else if (fmt == kFmtF16x2)
      upsample_match_strip_kernel<kFmtF16x2, kFmtF16x2><<<grid, 256, 0, st>>>(p, shift, strips, wpairs);
    else if (fmt == kFmtF16X8)
      upsample_match_strip_kernel<kFmtF16x2, kFmtF16X8><<<grid, 256, 0, st>>>(p, shift, strips, wpairs);
    else
      upsample_match_strip_kernel<kFmtBf16x2, kFmtBf16x2><<<grid, 256, 0, st>>>(p, shift, strips, wpairs);
  } else {
    if (fmt == kFmtF16X8) {
      set_error("upsample_match: the 8-bit cross format is only produced by the up-sampling paths (hd >= hs, wd >= ws)");
      return 1;
    }
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
  Acc* acc;
  int vec;
};
__device__ __forceinline__ float sgn(float d) { return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) l1_loss_kernel(const LossParams p) {
  float s_l1 = 0.f, s_p = 0.f, bad = 0.f;
  const long long n4 = p.vec ? p.numel >> 2 : 0;   // 16-byte path only when every pointer is 16-byte aligned
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
    if (t != 0.0) acc_add(&p.acc[threadIdx.x], t);   // order-independent (nsm_common.cuh: Acc)
  }
}

int l1_loss_fwd_bwd(const float* out, const float* target, const float* const* perturbed, int n_perturbed,
                    long long numel, float coef_l1, float coef_pert, float* grad, Acc* acc, cudaStream_t st) {
  if (n_perturbed < 0 || n_perturbed > 4) {
    set_error("l1_loss: at most 4 perturbed outputs per launch (got %d)", n_perturbed);
    return 1;
  }
  LossParams p;
  p.out = out; p.target = target; p.n_pert = n_perturbed; p.numel = numel;
  for (int k = 0; k < 4; ++k) p.pert[k] = k < n_perturbed ? perturbed[k] : nullptr;
  p.coef_l1 = coef_l1; p.coef_pert = coef_pert; p.grad = grad; p.acc = acc;
  // a contiguous view with an odd element offset (e.g. a slice of a larger buffer) is legal input: scalar loop then
  uintptr_t bits = reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(grad);
  for (int k = 0; k < n_perturbed; ++k) bits |= reinterpret_cast<uintptr_t>(perturbed[k]);
  p.vec = (bits & 15) == 0 ? 1 : 0;
  l1_loss_kernel<<<grid_for((numel + 3) / 4, 256, 148 * 8), 256, 0, st>>>(p);
  NSM_CHECK_LAUNCH("l1_loss");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// customLoss.EnhancedCustomLoss (customLoss.py:195-238): input jitter and the MSE stability term
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_noise_clamp_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                              long long numel, float eps, float lo, float hi,
                                                              float* __restrict__ out) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < numel; i += (long long)gridDim.x * 256) {
    // torch: noise * eps, then inputs + (that), then clamp -- two roundings, no fused multiply-add
    const float v = __fadd_rn(x[i], __fmul_rn(noise[i], eps));
    out[i] = fminf(fmaxf(v, lo), hi);   // (a NaN input stays NaN in torch.clamp; fmaxf/fminf would drop it)
    if (v != v) out[i] = v;
  }
}

int add_noise_clamp(const float* x, const float* noise, long long numel, float eps, float lo, float hi, float* out,
                    cudaStream_t st) {
  if (numel <= 0) return 0;
  add_noise_clamp_kernel<<<grid_for(numel, 256, 148 * 8), 256, 0, st>>>(x, noise, numel, eps, lo, hi, out);
  NSM_CHECK_LAUNCH("add_noise_clamp");
  return 0;
}

__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ out, const float* __restrict__ ref,
                                                       long long numel, float* __restrict__ diff, Acc* acc) {
  float s = 0.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < numel; i += (long long)gridDim.x * 256) {
    const float d = out[i] - ref[i];
    s = fmaf(d, d, s);
    if (diff) diff[i] = d;
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += double(red[k]);
    if (t != 0.0) acc_add(&acc[0], t);   // order-independent (nsm_common.cuh: Acc)
  }
}

int mse_loss_fwd_bwd(const float* out, const float* ref, long long numel, float* diff, Acc* acc, cudaStream_t st) {
  if (numel <= 0) return 0;
  mse_loss_kernel<<<grid_for(numel, 256, 148 * 8), 256, 0, st>>>(out, ref, numel, diff, acc);
  NSM_CHECK_LAUNCH("mse_loss");
  return 0;
}

// ------------------------------------------------------------------------------------------------
// per-channel sums over [S][C][HW] in fp64 (two-pass statistics like calculate_dataset_stats.py)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) channel_sums_kernel(const float* __restrict__ x, long long S, int C,
                                                           long long HW, const double* __restrict__ means,
                                                           Acc* __restrict__ sums, int chunks) {
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
    acc_add(&sums[c], t);
  }
}

int channel_sums(const float* x, long long S, int C, long long HW, const double* means, Acc* sums,
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
