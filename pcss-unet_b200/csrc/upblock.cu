// Fused decoder block of the eval-mode U-Net (Unetmodel.py:134-148): ONE kernel per thin DoubleConv of the decoder,
//
//   conv8:  up8 = resize(up2x(merge7)) -> 3x3 (128->128) + BN + LReLU -> 1x1 (128->64) + BN + LReLU -> + c2      = merge8
//   conv9:  up9 = resize(up2x(merge8)) -> 3x3 (64->64)   + BN + LReLU -> 1x1 (64->16)  + BN + LReLU
//           -> conv10 (16->4) + bias -> sigmoid -> pixel_shuffle(2)                                               = output
//
// instead of six launches (two up-samplers, two 3x3 GEMMs, a 1x1 GEMM, the tail) that wrote and re-read the up-sampled
// tensors u8 / u9 and the intermediates t8 / t9 (1.2 GB of HBM traffic per 1080p frame in fp32 mode).
//
// Work item = 16 x 8 output pixels (M = 128 TMEM lanes, lane = row * 8 + column).  Per item and 64-channel chunk:
//   * warp 2 fetches the low-resolution source box the tile's (16+2) x (8+2) halo depends on with ONE TMA box load;
//   * the 16 worker warps interpolate the halo (composite of the x2 bilinear up-sample and the resize to the skip's size:
//     a separable, position-dependent stencil of up to 3 x 3 source pixels, resample.cuh) and write it to shared memory
//     directly in the K-major 128-byte-swizzled UMMA operand layout, pixel rows in halo order;
//   * the nine taps of the 3x3 convolution are nine UMMA descriptors into that ONE halo tile (start shifted by
//     (dy * 10 + dx) pixel rows, 8-row groups at a stride of one halo row = 1280 B; the swizzle follows absolute
//     shared-memory address bits, so a descriptor may start at any 128-byte row) -- no im2col, no per-tap re-load;
//   * weights stream through a TMA ring; accumulators live in TMEM;
//   * the workers turn the 3x3 accumulator into the 1x1 convolution's A operand (bias, BN, LeakyReLU, hi/lo split) and
//     store it back to TENSOR MEMORY (tcgen05.st); the 1x1 GEMM reads A from TMEM (.ts form), B from the ring;
//   * final epilogue: bias, BN, LeakyReLU, then the skip add and the store (conv8), or conv10 + sigmoid + pixel_shuffle
//     in registers (conv9).
// fp32 mode: source planes fp16 hi+lo; 3x3 operands fp16 hi + 8-bit cross plane (two MMA slots per k-step, nsm_common.cuh);
// 1x1 operands fp16 hi+lo (a_hi x [w_hi | w_lo] as one wide MMA + a_lo x w_hi).  bf16 mode: one plane, autocast rounding
// points (the composite resize is rounded once, like upsample_match).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "conv_gemm.cuh"
#include "nsm_common.cuh"
#include "resample.cuh"
#include "upblock.cuh"

namespace nsm {

namespace {

constexpr int kUbTW = 8, kUbTH = 16;                  // output tile
constexpr int kHaloW = kUbTW + 2, kHaloH = kUbTH + 2;  // 10 x 18 halo pixels
constexpr int kHaloRows = kHaloW * kHaloH;             // 180 rows of 128 B
constexpr int kHaloPlaneBytes = kHaloRows * 128;       // 23040
constexpr int kUbThreads = 640;                        // warps 0-3: B producer, MMA, source producer, idle; 4-19: workers
constexpr int kWorkers = 16;

struct UbKernelParams {
  int N, Hs, Ws, H, W;
  int Cout, tail;
  int dbg;   // NSM_UB_DBG=64: cycle counters of one worker warp (nsm_upblock_prof)
  int tiles_x, tiles_y, total_tiles;
  int sbw, sbh;                       // source box in pixels
  uint32_t off_src, off_b, off_vec, off_taps, off_bar;
  uint32_t src_plane_bytes, b_stage_bytes, b_stages;
  uint32_t w_resident;                // every weight tile (9 * NCH of the 3x3, NCH of the 1x1) has its own stage: loaded once per CTA
  uint32_t halo_bufs, src_bufs;       // ring depths of the halo tiles (<= 4) and of the staged source boxes (<= 2)
  uint32_t acc1_bufs;                 // 2: the 3x3 accumulator is double-buffered in TMEM, the MMAs run one tile ahead
  uint32_t tm_acc2, tm_a2, tmem_cols;  // TMEM column offsets
  uint32_t idesc1, idesc2w, idesc2c;
  const float *bias3, *scale3, *shift3, *bias1, *scale1, *shift1, *w10, *b10;
  Planes out, residual;
  float* y;
  uint8_t* y_u8;
};

constexpr int kStrips = kHaloH / 3;   // 3-row strips of the halo
constexpr int kStripRows = 6;         // source rows one strip may touch (host-checked)
constexpr int kTabs = 4;              // interpolation tables in flight (one per tile, shared by its channel chunks)
struct JobTab {
  float colw[kHaloW][4];
  float roww[kHaloH][kStripRows];
  int colp[kHaloW];
  int strip_r0[kStrips], strip_rn[kStrips];
  int row_rmin[kHaloH];        // scratch of the preparing warp: per halo row first source row (-1: outside the image) ...
  float row_w[kHaloH][4];      // ... and its three vertical weights
};

// cycle counters of block 0's worker warp 0 (NSM_UB_DBG bit 64; read with nsm_upblock_prof): where a worker's time goes
__device__ unsigned long long g_ub_prof[32];   // [tail][worker phases 0-7 | MMA-thread phases 8-15]
#define UB_T(i) do { if (prof) { const long long c_ = clock64(); acc_[i] += c_ - t_; t_ = c_; } } while (0)

struct TileCoord {
  int n, y0, x0;
};
__device__ __forceinline__ TileCoord ub_tile(const UbKernelParams& p, int item) {
  TileCoord t;
  const int tx = item % p.tiles_x;
  int r = item / p.tiles_x;
  const int ty = r % p.tiles_y;
  t.n = r / p.tiles_y;
  t.y0 = ty * kUbTH;
  t.x0 = tx * kUbTW;
  return t;
}
// first source row / column the tile's halo touches (= origin of the TMA source box)
__device__ __forceinline__ void ub_src_origin(const UbKernelParams& p, const TileCoord& t, int& sy0, int& sx0) {
  sy0 = composite_taps(t.y0 > 0 ? t.y0 - 1 : 0, p.Hs, p.H).rmin;
  sx0 = composite_taps(t.x0 > 0 ? t.x0 - 1 : 0, p.Ws, p.W).rmin;
}

// 8 consecutive channels of one source pixel from the staged box (hi [+ lo] planes) as fp32
template <int NP>
__device__ __forceinline__ void ub_load8(uint32_t saddr, uint32_t plane_bytes, float (&v)[8]) {
  const uint4 h = lds16(saddr);
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
  constexpr int FMT = NP == 2 ? kFmtF16x2 : kFmtBf16;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = hi_lo_to_f32(hw[e], FMT);
    v[2 * e + 1] = hi_hi_to_f32(hw[e], FMT);
  }
  if (NP == 2) {
    const uint4 l = lds16(saddr + plane_bytes);
    const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] += lo_lo_to_f32(lw[e], FMT);
      v[2 * e + 1] += lo_hi_to_f32(lw[e], FMT);
    }
  }
}

template <int NP, int CMID>
__global__ void __launch_bounds__(kUbThreads, 1)
upblock_kernel(const __grid_constant__ CUtensorMap tmS0, const __grid_constant__ CUtensorMap tmS1,
               const __grid_constant__ CUtensorMap tmW3a, const __grid_constant__ CUtensorMap tmW3b,
               const __grid_constant__ CUtensorMap tmW1a, const __grid_constant__ CUtensorMap tmW1b,
               const __grid_constant__ UbKernelParams p) {
  constexpr int NCH = CMID / 64;                 // 64-channel chunks of the 3x3 convolution's K (per tap)
  constexpr int FMT = NP == 2 ? kFmtF16x2 : kFmtBf16;
  constexpr bool rb = NP == 1;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const uint32_t sbase = smem_u32(smem);
  // halo ring at offset 0: buffer b, plane pl at (b * NP + pl) * kHaloPlaneBytes
  float* vec = reinterpret_cast<float*>(smem + p.off_vec);
  JobTab* tabs = reinterpret_cast<JobTab*>(smem + p.off_taps);   // [kTabs]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* b_full = bars;                 // [b_stages]
  uint64_t* b_empty = bars + 8;            // [b_stages]
  uint64_t* halo_full = bars + 16;         // [4]
  uint64_t* halo_empty = bars + 20;        // [4]
  uint64_t* src_full = bars + 24;          // [2]
  uint64_t* src_empty = bars + 26;         // [2]
  uint64_t* acc1_full = bars + 28;         // [2]
  uint64_t* acc1_empty = bars + 30;        // [2]
  uint64_t* a2_full = bars + 32;
  uint64_t* acc2_full = bars + 33;
  uint64_t* acc2_empty = bars + 34;
  uint64_t* tab_full = bars + 35;          // [kTabs]
  uint64_t* tab_empty = bars + 39;         // [kTabs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 43);

  // warp index through a shuffle: the compiler then knows it is warp-uniform, and the MMA warp below runs its loops with
  // uniform control flow and ONE elected lane issuing (elect.sync): the tcgen05.mma / commit instructions come out as
  // straight-line UTCHMMA with descriptors advanced by one 64-bit add.  Under `if (lane == 0)` every one of them was
  // wrapped in a five-instruction ELECT / BRA.U.ANY loop behind a chain of uniform-datapath descriptor arithmetic, and the
  // short MMAs of these blocks (N = 64 / 128: 32 / 64 cycles of math) were issue-bound at ~135 cycles each.
  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int nt = (p.total_tiles - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);   // tiles of this CTA

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmS0); tma_prefetch_desc(&tmW3a); tma_prefetch_desc(&tmW1a);
    if (NP == 2) { tma_prefetch_desc(&tmS1); tma_prefetch_desc(&tmW3b); tma_prefetch_desc(&tmW1b); }
    for (uint32_t s = 0; s < p.b_stages; ++s) {
      mbar_init(&b_full[s], 1);                         // resident weights: up to 16 "full" barriers, no "empty" ones
      if (!p.w_resident) mbar_init(&b_empty[s], 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&halo_full[b], kWorkers);
      mbar_init(&halo_empty[b], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&src_full[b], 1);
      mbar_init(&src_empty[b], kWorkers);
    }
    for (int b = 0; b < kTabs; ++b) {
      mbar_init(&tab_full[b], 1);
      mbar_init(&tab_empty[b], kWorkers);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], kWorkers);
    }
    mbar_init(a2_full, kWorkers);
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, kWorkers);
    fence_barrier_init();
  }
  // per-channel vectors -> shared memory: [bias3 | scale3 | shift3](CMID) [bias1 | scale1 | shift1](Cout) [w10 64] [b10 4]
  for (int i = threadIdx.x; i < CMID; i += kUbThreads) {
    vec[i] = p.bias3[i]; vec[CMID + i] = p.scale3[i]; vec[2 * CMID + i] = p.shift3[i];
  }
  for (int i = threadIdx.x; i < p.Cout; i += kUbThreads) {
    vec[3 * CMID + i] = p.bias1[i]; vec[3 * CMID + p.Cout + i] = p.scale1[i]; vec[3 * CMID + 2 * p.Cout + i] = p.shift1[i];
  }
  if (p.tail && threadIdx.x < 68)
    vec[3 * CMID + 3 * p.Cout + threadIdx.x] = threadIdx.x < 64 ? p.w10[threadIdx.x] : p.b10[threadIdx.x - 64];
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_acc2 = tmem_base + p.tm_acc2, tm_a2 = tmem_base + p.tm_a2;
  const bool db = p.acc1_bufs == 2;
  // 3x3 accumulator of tile j (double-buffered where tensor memory allows) and the phase of its barriers
  auto tm_acc1_of = [&](int j) { return tmem_base + (db ? uint32_t(j & 1) * uint32_t(NP * CMID) : 0u); };
  auto acc1_slot = [&](int j) { return db ? (j & 1) : 0; };
  auto acc1_use = [&](int j) { return uint32_t(db ? (j >> 1) : j); };

  if (warp == 0) {
    // ===================== weight producer: 3x3 tiles (chunk, tap), then the 1x1 tiles (chunk) =====================
    // (uniform control flow for the warp, one elected lane issues the copies)
    {
      const bool leader = elect_one();
      uint32_t stage = 0, phase = 0;
      auto advance = [&]() {
        if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
      };
      auto w3_tiles = [&]() {
        for (int c = 0; c < NCH; ++c)
          for (int tap = 0; tap < 9; ++tap) {
            if (!p.w_resident) mbar_wait(&b_empty[stage], phase ^ 1);
            uint8_t* sb = smem + p.off_b + stage * p.b_stage_bytes;
            if (leader) {
              mbar_expect_tx(&b_full[stage], NP * CMID * 128);
              tma_load_2d(sb, &tmW3a, &b_full[stage], tap * CMID + c * 64, 0);
              if (NP == 2) tma_load_2d(sb + CMID * 128, &tmW3b, &b_full[stage], tap * CMID + c * 64, 0);
            }
            __syncwarp();
            advance();
          }
      };
      auto w1_tiles = [&]() {
        for (int c = 0; c < NCH; ++c) {
          if (!p.w_resident) mbar_wait(&b_empty[stage], phase ^ 1);
          uint8_t* sb = smem + p.off_b + stage * p.b_stage_bytes;
          if (leader) {
            mbar_expect_tx(&b_full[stage], NP * p.Cout * 128);
            tma_load_2d(sb, &tmW1a, &b_full[stage], c * 64, 0);
            if (NP == 2) tma_load_2d(sb + p.Cout * 128, &tmW1b, &b_full[stage], c * 64, 0);   // lo rows right after hi rows
          }
          __syncwarp();
          advance();
        }
      };
      // same order as the MMA thread consumes them: with a double-buffered 3x3 accumulator the 3x3 GEMM of tile j+1 is
      // issued before the 1x1 GEMM of tile j
      if (p.w_resident) {
        // the whole block's weights fit next to the halo ring: stage (c * 9 + tap) / (9 * NCH + c), loaded once, never
        // released -- no weight traffic (L2 -> shared memory, shared-memory writes) per tile
        w3_tiles();
        w1_tiles();
      } else {
        if (db) w3_tiles();
        for (int j = 0; j < nt; ++j) {
          if (db) {
            if (j + 1 < nt) w3_tiles();
          } else {
            w3_tiles();
          }
          w1_tiles();
        }
      }
    }
  } else if (warp == 2) {
    // ===================== job preparation (runs ahead of the workers): interpolation table + source box =====================
    // Per tile: lanes 0-9 the halo columns' horizontal taps, lanes 10-27 the halo rows' vertical taps (one composite_taps =
    // a few float divisions per lane, off the workers' critical path here), lanes 0-5 the 3-row strips' source-row windows;
    // per (tile, chunk) job: lane 0 the TMA load of the source box.  Tables ring through kTabs buffers (tab_full / tab_empty),
    // source boxes through src_bufs buffers.
    const bool leader2 = elect_one();
    for (int j = 0, q = 0; j < nt; ++j) {
      const TileCoord t = ub_tile(p, int(blockIdx.x) + j * int(gridDim.x));
      int sy0, sx0;
      ub_src_origin(p, t, sy0, sx0);
      // ---- the tile's table (one composite_taps per lane: 10 columns, 18 rows), then the strips' row windows
      JobTab* tb = tabs + (j % kTabs);
      mbar_wait(&tab_empty[j % kTabs], ((j / kTabs) & 1) ^ 1);
      if (lane < kHaloW) {
        const int gx = t.x0 - 1 + lane;
        Tap3 tp;
        tp.rmin = sx0; tp.w[0] = tp.w[1] = tp.w[2] = 0.f;
        if (gx >= 0 && gx < p.W) tp = composite_taps(gx, p.Ws, p.W);
        tb->colp[lane] = tp.rmin - sx0;
        tb->colw[lane][0] = tp.w[0]; tb->colw[lane][1] = tp.w[1]; tb->colw[lane][2] = tp.w[2];
      } else if (lane < kHaloW + kHaloH) {
        const int r = lane - kHaloW, gy = t.y0 - 1 + r;
        Tap3 tp;
        tp.rmin = -1; tp.w[0] = tp.w[1] = tp.w[2] = 0.f;
        if (gy >= 0 && gy < p.H) tp = composite_taps(gy, p.Hs, p.H);
        tb->row_rmin[r] = tp.rmin;
        tb->row_w[r][0] = tp.w[0]; tb->row_w[r][1] = tp.w[1]; tb->row_w[r][2] = tp.w[2];
      }
      __syncwarp();
      if (lane < kStrips) {
        const int s = lane;
        int r0 = -1, rn = 0;
        float wv[3][kStripRows];
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int k = 0; k < kStripRows; ++k) wv[rr][k] = 0.f;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          const int rmin = tb->row_rmin[3 * s + rr];
          if (rmin < 0) continue;
          if (r0 < 0) r0 = rmin;
#pragma unroll
          for (int i = 0; i < 3; ++i) {
            const float wi = tb->row_w[3 * s + rr][i];
            const int k = rmin - r0 + i;   // < kStripRows (checked on the host for every strip of the launch)
            if (wi != 0.f && k < kStripRows) {
#pragma unroll
              for (int kk = 0; kk < kStripRows; ++kk)
                if (kk == k) wv[rr][kk] = wi;
              if (k + 1 > rn) rn = k + 1;
            }
          }
        }
        tb->strip_r0[s] = r0 < 0 ? 0 : r0 - sy0;
        tb->strip_rn[s] = rn;
#pragma unroll
        for (int rr = 0; rr < 3; ++rr)
#pragma unroll
          for (int k = 0; k < kStripRows; ++k) tb->roww[3 * s + rr][k] = wv[rr][k];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tab_full[j % kTabs]);   // (mbarrier arrive has release semantics for the table writes above)
      // ---- the source boxes of the tile's chunks
      for (int c = 0; c < NCH; ++c, ++q) {
        const uint32_t sb = uint32_t(q) % p.src_bufs, use = uint32_t(q) / p.src_bufs;
        uint8_t* dst = smem + p.off_src + sb * NP * p.src_plane_bytes;
        mbar_wait(&src_empty[sb], (use & 1) ^ 1);
        if (leader2) {
          mbar_expect_tx(&src_full[sb], NP * p.sbw * p.sbh * 128);
          tma_load_4d(dst, &tmS0, &src_full[sb], c * 64, sx0, sy0, t.n);
          if (NP == 2) tma_load_4d(dst + p.src_plane_bytes, &tmS1, &src_full[sb], c * 64, sx0, sy0, t.n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // the whole warp walks the loops (uniform control flow, all lanes poll the barriers); the elected lane issues
    {
      const bool leader = elect_one();
      uint32_t stage = 0, phase = 0;
      auto advance = [&]() {
        if (++stage == p.b_stages) { stage = 0; phase ^= 1; }
      };
      const bool prof = (p.dbg & 64) && blockIdx.x == 0;
      long long acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_ = clock64();
      // descriptor halves that never change: LBO 16 B, SBO (one halo row for A, 1024 B for B), version 1, 128-byte swizzle
      const uint64_t desc_a_hi = make_desc_sw128(0, 16, uint32_t(kHaloW * 128));
      const uint64_t desc_b_hi = make_desc_sw128(0, 16, 1024);
      auto desc_of = [](uint64_t hi, uint32_t addr) { return hi | uint64_t((addr & 0x3FFFFu) >> 4); };
      // ---- GEMM 1: 3x3 convolution of tile j out of its halo tiles ----
      auto gemm1 = [&](int j) {
        UB_T(7);
        mbar_wait(&acc1_empty[acc1_slot(j)], (acc1_use(j) & 1) ^ 1);
        tc_fence_after();
        UB_T(0);
        const uint32_t tm_acc1 = tm_acc1_of(j);
        for (int c = 0; c < NCH; ++c) {
          const uint32_t q = uint32_t(j * NCH + c), hb = q % p.halo_bufs;
          mbar_wait(&halo_full[hb], (q / p.halo_bufs) & 1);
          tc_fence_after();
          UB_T(1);
          const uint32_t halo = sbase + uint32_t(hb * NP) * kHaloPlaneBytes;
          for (int tap = 0; tap < 9; ++tap) {
            if (p.w_resident) { stage = uint32_t(c * 9 + tap); phase = 0; }
            mbar_wait(&b_full[stage], phase);
            tc_fence_after();
            UB_T(2);
            const uint32_t a0 = halo + uint32_t((tap / 3) * kHaloW + tap % 3) * 128u;
            const uint32_t b0 = sbase + p.off_b + stage * p.b_stage_bytes;
            // K advances by 32 B per MMA = +2 in the descriptor's 16-byte address field (shared memory ends below 2^18)
            const uint64_t da = desc_of(desc_a_hi, a0), dbw = desc_of(desc_b_hi, b0);
            const uint64_t da2 = desc_of(desc_a_hi, a0 + kHaloPlaneBytes), dbw2 = desc_of(desc_b_hi, b0 + CMID * 128);
            if (leader) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t accum = (c | tap | k) != 0 ? 1u : 0u;
                umma_bf16(tm_acc1, da + 2 * k, dbw + 2 * k, p.idesc1, accum);
                if (NP == 2)   // both cross terms as one e4m3 MMA of K = 32 (8-bit cross planes)
                  umma_f8(tm_acc1 + CMID, da2 + 2 * k, dbw2 + 2 * k, p.idesc1, accum);
              }
              if (!p.w_resident) umma_commit(&b_empty[stage]);
              if (tap == 8) {
                umma_commit(&halo_empty[hb]);
                if (c == NCH - 1) umma_commit(&acc1_full[acc1_slot(j)]);
              }
            }
            __syncwarp();
            if (!p.w_resident) advance();
            UB_T(3);
          }
        }
      };
      // ---- GEMM 2: 1x1 convolution of tile j, A from tensor memory ----
      auto gemm2 = [&](int j) {
        UB_T(7);
        mbar_wait(a2_full, j & 1);
        UB_T(4);
        mbar_wait(acc2_empty, (j & 1) ^ 1);
        tc_fence_after();
        UB_T(5);
        for (int c = 0; c < NCH; ++c) {
          if (p.w_resident) { stage = uint32_t(9 * NCH + c); phase = 0; }
          mbar_wait(&b_full[stage], phase);
          tc_fence_after();
          UB_T(2);
          const uint32_t b0 = sbase + p.off_b + stage * p.b_stage_bytes;
          const uint64_t dbw = desc_of(desc_b_hi, b0);
          if (leader) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t accum = (c | k) != 0 ? 1u : 0u;
              const uint32_t a_hi = tm_a2 + uint32_t(c * 4 + k) * 8u;
              if (NP == 2) {
                umma_bf16_ts(tm_acc2, a_hi, dbw + 2 * k, p.idesc2w, accum);
                umma_bf16_ts(tm_acc2 + p.Cout, a_hi + CMID / 2, dbw + 2 * k, p.idesc2c, 1u);
              } else {
                umma_bf16_ts(tm_acc2, a_hi, dbw + 2 * k, p.idesc2c, accum);
              }
            }
            if (!p.w_resident) umma_commit(&b_empty[stage]);
            if (c == NCH - 1) umma_commit(acc2_full);
          }
          __syncwarp();
          if (!p.w_resident) advance();
          UB_T(6);
        }
      };
      // Double-buffered 3x3 accumulator: the 3x3 GEMM of tile j+1 is issued BEFORE this warp waits for the workers' 1x1
      // operand of tile j, so the tensor pipe keeps working through the workers' accumulator -> operand pass.
      if (db) gemm1(0);
      for (int j = 0; j < nt; ++j) {
        if (db) {
          if (j + 1 < nt) gemm1(j + 1);
        } else {
          gemm1(j);
        }
        gemm2(j);
      }
      if (prof && leader)
        for (int i = 0; i < 8; ++i) atomicAdd(&g_ub_prof[(p.tail ? 16 : 0) + 8 + i], (unsigned long long)acc_[i]);
    }
  } else if (warp >= 4) {
    // ===================== workers: halo interpolation, both epilogues =====================
    const int w = warp - 4;
    const int qd = w & 3;              // TMEM lane quarter (= warp id % 4)
    const int cs = w >> 2;             // column slice
    const int wt = threadIdx.x - 128;  // 0 .. 511
    const uint32_t lane_base = uint32_t(qd * 32) << 16;
    const int ly = qd * 4 + (lane >> 3), lx = lane & 7;   // this lane's pixel inside the tile
    const bool prof = (p.dbg & 64) && blockIdx.x == 0 && w == 0 && lane == 0;
    long long acc_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_ = clock64();

    // ---- one (tile, chunk) job: interpolate the halo of 64 channels into halo buffer q & 1 ----
    // Branch-free per element: a job table (built ahead of time by warp 2, double-buffered) holds per halo column the first source
    // column and three horizontal weights, per 3-row strip the first source row and row count, per halo row its weights on
    // the strip's source rows.  Rows / columns outside the image have all-zero weights = the convolution's zero padding.
    auto do_job = [&](int q) {
      const int jt = q / NCH;                      // the job's tile: its table is shared by the tile's chunks
      JobTab* tb = tabs + (jt % kTabs);
      mbar_wait(&tab_full[jt % kTabs], (jt / kTabs) & 1);
      const uint32_t hb = uint32_t(q) % p.halo_bufs, sb = uint32_t(q) % p.src_bufs;
      mbar_wait(&halo_empty[hb], ((uint32_t(q) / p.halo_bufs) & 1) ^ 1);   // the MMAs of the buffer's previous use have read it
      mbar_wait(&src_full[sb], (uint32_t(q) / p.src_bufs) & 1);
      UB_T(0);
      const int cg = wt & 7, slot = wt >> 3;
      if (slot < kStrips * kHaloW) {
        const int hx = slot % kHaloW, strip = slot / kHaloW;
        const float cw0 = tb->colw[hx][0], cw1 = tb->colw[hx][1], cw2 = tb->colw[hx][2];
        const int px0 = tb->colp[hx];
        const int px1 = px0 + 1 < p.sbw ? px0 + 1 : p.sbw - 1, px2 = px0 + 2 < p.sbw ? px0 + 2 : p.sbw - 1;
        const int r0 = tb->strip_r0[strip], rn = tb->strip_rn[strip];
        const uint32_t src0 = sbase + p.off_src + sb * NP * p.src_plane_bytes + uint32_t(cg) * 16u;
        float o0[8], o1[8], o2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o0[e] = o1[e] = o2[e] = 0.f;
#pragma unroll 1
        for (int k = 0; k < rn; ++k) {
          const int r = r0 + k < p.sbh ? r0 + k : p.sbh - 1;   // rows past the box carry zero weight only
          const uint32_t rowa = src0 + uint32_t(r * p.sbw) * 128u;
          float v0[8], v1[8], h[8];
          ub_load8<NP>(rowa + uint32_t(px0) * 128u, p.src_plane_bytes, v0);
          ub_load8<NP>(rowa + uint32_t(px1) * 128u, p.src_plane_bytes, v1);
#pragma unroll
          for (int e = 0; e < 8; ++e) h[e] = fmaf(cw1, v1[e], fmaf(cw0, v0[e], 0.f));
          if (cw2 != 0.f) {
            ub_load8<NP>(rowa + uint32_t(px2) * 128u, p.src_plane_bytes, v0);
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = fmaf(cw2, v0[e], h[e]);
          }
          const float w0 = tb->roww[3 * strip][k], w1 = tb->roww[3 * strip + 1][k], w2 = tb->roww[3 * strip + 2][k];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            o0[e] = fmaf(w0, h[e], o0[e]);
            o1[e] = fmaf(w1, h[e], o1[e]);
            o2[e] = fmaf(w2, h[e], o2[e]);
          }
        }
        const uint32_t halo = sbase + uint32_t(hb * NP) * kHaloPlaneBytes;
        auto put = [&](int rr, float (&v)[8]) {
          const uint32_t row = halo + uint32_t((strip * 3 + rr) * kHaloW + hx) * 128u;
          const uint32_t sw = (row >> 7) & 7u;
          if (rb) {
#pragma unroll
            for (int e = 0; e < 8; e += 2) rbf2(v[e], v[e + 1]);
          }
          uint32_t hw[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) hw[e] = pack_hi(v[2 * e], v[2 * e + 1], FMT);
          sts16(row + ((uint32_t(cg) ^ sw) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
          if (NP == 2) {
            // 8-bit cross plane: per 16 channels [16 B of e4m3(2 v) | 16 B of e4m3(2^11 (v - hi))]
            uint2 first, second;
            x8_act_bytes(v, hw, first, second);
            const uint32_t row2 = row + kHaloPlaneBytes;
            const uint32_t sw2 = (row2 >> 7) & 7u;
            const uint32_t ch = uint32_t(cg >> 1) * 2u, sub = uint32_t(cg & 1) * 8u;
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(row2 + ((ch ^ sw2) << 4) + sub), "r"(first.x), "r"(first.y)
                         : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(row2 + (((ch + 1u) ^ sw2) << 4) + sub), "r"(second.x),
                         "r"(second.y)
                         : "memory");
          }
        };
        put(0, o0); put(1, o1); put(2, o2);
      }
      UB_T(1);
      fence_proxy_async();   // generic-proxy writes of the halo -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&halo_full[hb]);
        mbar_arrive(&src_empty[sb]);
        if (q - jt * NCH == NCH - 1) mbar_arrive(&tab_empty[jt % kTabs]);   // the tile's last chunk is done with the table
      }
      UB_T(2);
    };

    // ---- 3x3 accumulator -> A operand of the 1x1 GEMM in tensor memory ----
    auto mid_epilogue = [&](int j) {
      mbar_wait(&acc1_full[acc1_slot(j)], acc1_use(j) & 1);
      tc_fence_after();
      const uint32_t tm_acc1 = tm_acc1_of(j);
      UB_T(3);
      constexpr int CW = CMID / 4;   // columns per warp
#pragma unroll
      for (int g = 0; g < CW / 16; ++g) {
        const int col = cs * CW + g * 16;
        uint32_t r0[16], r1[16];
        tmem_ld_32x16(tm_acc1 + lane_base + col, r0);
        if (NP == 2) tmem_ld_32x16(tm_acc1 + lane_base + CMID + col, r1);
        tmem_ld_wait();
        uint32_t hw[8], lw[8];
#pragma unroll
        for (int e4 = 0; e4 < 16; e4 += 4) {
          const float4 B = *reinterpret_cast<const float4*>(vec + col + e4);
          const float4 S = *reinterpret_cast<const float4*>(vec + CMID + col + e4);
          const float4 T = *reinterpret_cast<const float4*>(vec + 2 * CMID + col + e4);
          const float bb[4] = {B.x, B.y, B.z, B.w}, sc[4] = {S.x, S.y, S.z, S.w}, sh[4] = {T.x, T.y, T.z, T.w};
          float a[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float v = __uint_as_float(r0[e4 + k]);
            if (NP == 2) v = fmaf(__uint_as_float(r1[e4 + k]), kX8CrossScale, v);
            a[k] = v + bb[k];
          }
          if (rb) { rbf2(a[0], a[1]); rbf2(a[2], a[3]); }
#pragma unroll
          for (int k = 0; k < 4; ++k) a[k] = fmaf(a[k], sc[k], sh[k]);
          if (rb) { rbf2(a[0], a[1]); rbf2(a[2], a[3]); }
#pragma unroll
          for (int k = 0; k < 4; ++k) a[k] = lrelu02(a[k]);
          if (rb) { rbf2(a[0], a[1]); rbf2(a[2], a[3]); }
          hw[e4 / 2] = pack_hi(a[0], a[1], FMT);
          hw[e4 / 2 + 1] = pack_hi(a[2], a[3], FMT);
          if (NP == 2) {
            lw[e4 / 2] = pack_lo_resid(a[0], a[1], hw[e4 / 2], FMT);
            lw[e4 / 2 + 1] = pack_lo_resid(a[2], a[3], hw[e4 / 2 + 1], FMT);
          }
        }
        tmem_st_32x8(tm_a2 + lane_base + col / 2, hw);
        if (NP == 2) tmem_st_32x8(tm_a2 + lane_base + CMID / 2 + col / 2, lw);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&acc1_empty[acc1_slot(j)]);   // this 3x3 accumulator may be overwritten (tile j + acc1_bufs)
        mbar_arrive(a2_full);      // the 1x1 GEMM's A operand is in place
      }
      UB_T(4);
    };

    // ---- 1x1 accumulator -> output ----
    auto final_epilogue = [&](int j, const TileCoord& t) {
      const bool active = cs * 16 < p.Cout;
      const int col = cs * 16;
      const int y = t.y0 + ly, x = t.x0 + lx;
      const bool inside = active && y < p.H && x < p.W;
      const size_t off = (((size_t)t.n * p.H + y) * p.W + x) * p.Cout + col;   // element offset (conv8 form)
      // the skip values do not depend on the accumulator: request them before waiting for it
      uint4 rh[2], rl[2];
      const bool has_res = !p.tail && p.residual.p[0] != nullptr;
      if (inside && has_res) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          rh[hf] = ldg16(reinterpret_cast<const uint8_t*>(p.residual.p[0]) + (off + 8 * hf) * 2);
          if (NP == 2) rl[hf] = ldg16(reinterpret_cast<const uint8_t*>(p.residual.p[1]) + (off + 8 * hf) * 2);
        }
      }
      UB_T(5);
      mbar_wait(acc2_full, j & 1);
      tc_fence_after();
      UB_T(6);
      if (active) {
        uint32_t r0[16], r1[16];
        tmem_ld_32x16(tm_acc2 + lane_base + col, r0);
        if (NP == 2) tmem_ld_32x16(tm_acc2 + lane_base + p.Cout + col, r1);
        const float* v1 = vec + 3 * CMID;
        tmem_ld_wait();
        float a[16];
#pragma unroll
        for (int e4 = 0; e4 < 16; e4 += 4) {
          const float4 B = *reinterpret_cast<const float4*>(v1 + col + e4);
          const float4 S = *reinterpret_cast<const float4*>(v1 + p.Cout + col + e4);
          const float4 T = *reinterpret_cast<const float4*>(v1 + 2 * p.Cout + col + e4);
          const float bb[4] = {B.x, B.y, B.z, B.w}, sc[4] = {S.x, S.y, S.z, S.w}, sh[4] = {T.x, T.y, T.z, T.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float v = __uint_as_float(r0[e4 + k]);
            if (NP == 2) v += __uint_as_float(r1[e4 + k]);
            a[e4 + k] = v + bb[k];
          }
          if (rb) { rbf2(a[e4], a[e4 + 1]); rbf2(a[e4 + 2], a[e4 + 3]); }
#pragma unroll
          for (int k = 0; k < 4; ++k) a[e4 + k] = fmaf(a[e4 + k], sc[k], sh[k]);
          if (rb) { rbf2(a[e4], a[e4 + 1]); rbf2(a[e4 + 2], a[e4 + 3]); }
#pragma unroll
          for (int k = 0; k < 4; ++k) a[e4 + k] = lrelu02(a[e4 + k]);
          if (rb) { rbf2(a[e4], a[e4 + 1]); rbf2(a[e4 + 2], a[e4 + 3]); }
        }
        if (inside) {
          if (p.tail) {
            // conv10 (16 -> 4) + bias, sigmoid, pixel_shuffle(2): channel k = dy * 2 + dx -> pixel (2y + dy, 2x + dx)
            const float* w10 = v1 + 3 * p.Cout;
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              float s = 0.f;
#pragma unroll
              for (int e = 0; e < 16; e += 4) {
                const float4 w = *reinterpret_cast<const float4*>(w10 + k * 16 + e);
                s = fmaf(w.x, a[e], s); s = fmaf(w.y, a[e + 1], s); s = fmaf(w.z, a[e + 2], s); s = fmaf(w.w, a[e + 3], s);
              }
              s += w10[64 + k];
              if (rb) s = rbf(s);
              s = 1.f / (1.f + expf(-s));
              o[k] = rb ? rbf(s) : s;
            }
            const int Wo = 2 * p.W, Ho = 2 * p.H;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              const size_t oo = ((size_t)t.n * Ho + 2 * y + dy) * Wo + 2 * x;
              if (p.y) *reinterpret_cast<float2*>(p.y + oo) = make_float2(o[2 * dy], o[2 * dy + 1]);
              if (p.y_u8)   // (out * 255).astype(uint8): truncation toward zero of a value in [0, 255]
                *reinterpret_cast<uchar2*>(p.y_u8 + oo) =
                    make_uchar2((unsigned char)(o[2 * dy] * 255.f), (unsigned char)(o[2 * dy + 1] * 255.f));
            }
          } else {
            if (has_res) {
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                const uint32_t hw[4] = {rh[hf].x, rh[hf].y, rh[hf].z, rh[hf].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  a[8 * hf + 2 * e] += hi_lo_to_f32(hw[e], FMT);
                  a[8 * hf + 2 * e + 1] += hi_hi_to_f32(hw[e], FMT);
                }
                if (NP == 2) {
                  const uint32_t lw[4] = {rl[hf].x, rl[hf].y, rl[hf].z, rl[hf].w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    a[8 * hf + 2 * e] += lo_lo_to_f32(lw[e], FMT);
                    a[8 * hf + 2 * e + 1] += lo_hi_to_f32(lw[e], FMT);
                  }
                }
              }
              if (rb) {
#pragma unroll
                for (int e = 0; e < 16; e += 2) rbf2(a[e], a[e + 1]);
              }
            }
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hw[e] = pack_hi(a[8 * hf + 2 * e], a[8 * hf + 2 * e + 1], FMT);
                if (NP == 2) lw[e] = pack_lo_resid(a[8 * hf + 2 * e], a[8 * hf + 2 * e + 1], hw[e], FMT);
              }
              stg16(reinterpret_cast<uint8_t*>(p.out.p[0]) + (off + 8 * hf) * 2, make_uint4(hw[0], hw[1], hw[2], hw[3]));
              if (NP == 2)
                stg16(reinterpret_cast<uint8_t*>(p.out.p[1]) + (off + 8 * hf) * 2, make_uint4(lw[0], lw[1], lw[2], lw[3]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc2_empty);
      UB_T(7);
    };

    // Program order of a worker warp.  Single 3x3 accumulator (fp32 mode, 128 channels: tensor memory is full): the halo jobs
    // of tile j+1 are interleaved with the epilogues of tile j, and the 3x3 MMAs of tile j+1 start as soon as the accumulator
    // of tile j has been drained.  Double-buffered accumulator: halo tiles are prepared TWO tiles ahead (2 * NCH buffers), so
    // that the MMA thread can run the 3x3 GEMM of tile j+1 while the workers are still turning tile j's accumulator into the
    // 1x1 operand; the jobs of tile j+2 sit between the two epilogues of tile j (the 1x1 GEMM of tile j executes behind the 3x3
    // GEMM of tile j+1 in the tensor pipe, the workers use that time).
    int q_next = 0;
    if (db) {
      for (; q_next < 2 * NCH && q_next < nt * NCH; ++q_next) do_job(q_next);
      for (int j = 0; j < nt; ++j) {
        const TileCoord t = ub_tile(p, int(blockIdx.x) + j * int(gridDim.x));
        mid_epilogue(j);
        if (j + 2 < nt)
          for (int c = 0; c < NCH; ++c) do_job(q_next++);
        final_epilogue(j, t);
      }
    } else {
      for (; q_next < NCH && q_next < nt * NCH; ++q_next) do_job(q_next);
      for (int j = 0; j < nt; ++j) {
        const bool more = j + 1 < nt;
        const TileCoord t = ub_tile(p, int(blockIdx.x) + j * int(gridDim.x));
        if (more) do_job(q_next++);              // first chunk of the next tile: overlaps this tile's 3x3 MMAs
        mid_epilogue(j);
        if (more && NCH > 1) do_job(q_next++);   // second chunk: its halo buffer was released by this tile's MMAs
        final_epilogue(j, t);
      }
    }
    if (prof)
      for (int i = 0; i < 8; ++i) atomicAdd(&g_ub_prof[(p.tail ? 16 : 0) + i], (unsigned long long)acc_[i]);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

int g_ub_sms = 0;

// extent of the source box a tile's halo can touch along one axis (brute force over all tiles; same float arithmetic as
// the device's composite_taps, taps with zero weight do not count)
int src_box_extent(int in_size, int out_size, int tile, int halo) {
  int best = 1;
  for (int t0 = 0; t0 < out_size; t0 += tile) {
    const int first = t0 > 0 ? t0 - 1 : 0;
    const int last = t0 + halo - 2 < out_size ? t0 + halo - 2 : out_size - 1;   // t0 - 1 + halo - 1
    const int lo = composite_taps(first, in_size, out_size).rmin;
    int hi = lo;
    for (int d = first; d <= last; ++d) {
      const Tap3 tp = composite_taps(d, in_size, out_size);
      for (int k = 0; k < 3; ++k)
        if (tp.w[k] != 0.f && tp.rmin + k > hi) hi = tp.rmin + k;
    }
    if (hi - lo + 1 > best) best = hi - lo + 1;
  }
  return best;
}

// most source rows a 3-row strip of any tile's halo touches (the kernel's job table holds kStripRows of them)
int max_strip_rows(int in_size, int out_size) {
  int best = 1;
  for (int t0 = 0; t0 < out_size; t0 += kUbTH)
    for (int s = 0; s < kStrips; ++s) {
      int r0 = -1;
      for (int rr = 0; rr < 3; ++rr) {
        const int g = t0 - 1 + 3 * s + rr;
        if (g < 0 || g >= out_size) continue;
        const Tap3 tp = composite_taps(g, in_size, out_size);
        if (r0 < 0) r0 = tp.rmin;
        for (int i = 0; i < 3; ++i)
          if (tp.w[i] != 0.f && tp.rmin - r0 + i + 1 > best) best = tp.rmin - r0 + i + 1;
      }
    }
  return best;
}

template <int NP, int CMID>
int launch_ub(const CUtensorMap* maps, const UbKernelParams& kp, int grid, size_t smem_bytes, cudaStream_t st) {
  auto kern = upblock_kernel<NP, CMID>;
  static size_t attr_bytes = 0;
  if (smem_bytes > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes));
    if (e != cudaSuccess) {
      set_error("upblock<%d,%d>: cudaFuncSetAttribute(%zu B): %s", NP, CMID, smem_bytes, cudaGetErrorString(e));
      return 1;
    }
    attr_bytes = smem_bytes;
  }
  kern<<<grid, kUbThreads, smem_bytes, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("upblock<%d,%d> launch failed: %s", NP, CMID, cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  return 0;
}

}  // namespace

// debugging: read and clear the cycle counters (2 x 8 values, block with skip / block with tail: job wait / math / publish, mid wait / math, final prefetch /
// wait / math)
int upblock_prof(unsigned long long* out) {
  unsigned long long z[32] = {0};
  if (cudaMemcpyFromSymbol(out, g_ub_prof, sizeof(z)) != cudaSuccess) return 1;
  return cudaMemcpyToSymbol(g_ub_prof, z, sizeof(z)) != cudaSuccess;
}

bool upblock_supported(int Hs, int Ws, int H, int W) {
  return H >= Hs && W >= Ws && Hs >= 1 && Ws >= 1 && max_strip_rows(Hs, H) <= kStripRows;
}

int upblock_launch(const UpBlockArgs& a, cudaStream_t st) {
  const int NP = a.mode == kFmtBf16 ? 1 : 2;
  if ((a.mode != kFmtBf16 && a.mode != kFmtF16x2) || (a.Cmid != 64 && a.Cmid != 128) ||
      (a.Cout != 16 && a.Cout != 64) || a.N < 1 || a.H < 1 || a.W < 1 || a.Hs < 1 || a.Ws < 1 || a.H < a.Hs ||
      a.W < a.Ws || (a.tail && a.Cout != 16) || (!a.tail && !a.out.p[0])) {
    set_error("upblock: unsupported configuration mode=%d Cmid=%d Cout=%d src %dx%d dst %dx%d tail=%d", a.mode, a.Cmid,
              a.Cout, a.Hs, a.Ws, a.H, a.W, a.tail);
    return 1;
  }
  if (g_ub_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_ub_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_ub_sms <= 0) g_ub_sms = 148;
  }
  UbKernelParams kp;
  memset(&kp, 0, sizeof(kp));
  {
    static const int dbg = getenv("NSM_UB_DBG") ? atoi(getenv("NSM_UB_DBG")) : 0;
    kp.dbg = dbg;
  }
  kp.N = a.N; kp.Hs = a.Hs; kp.Ws = a.Ws; kp.H = a.H; kp.W = a.W; kp.Cout = a.Cout; kp.tail = a.tail;
  kp.tiles_x = (a.W + kUbTW - 1) / kUbTW;
  kp.tiles_y = (a.H + kUbTH - 1) / kUbTH;
  kp.total_tiles = a.N * kp.tiles_x * kp.tiles_y;
  kp.sbh = src_box_extent(a.Hs, a.H, kUbTH, kHaloH);
  kp.sbw = src_box_extent(a.Ws, a.W, kUbTW, kHaloW);
  if (max_strip_rows(a.Hs, a.H) > kStripRows) {
    set_error("upblock: a 3-row strip spans more than %d source rows (%d -> %d)", kStripRows, a.Hs, a.H);
    return 1;
  }
  if (kp.sbh > 256 || kp.sbw > 256) {
    set_error("upblock: source box %dx%d too large", kp.sbh, kp.sbw);
    return 1;
  }
  // shared-memory layout (offsets from the 1024-byte aligned base)
  auto up = [](size_t v, size_t al) { return (v + al - 1) / al * al; };
  // ring depths: two halo buffers + two source buffers where they fit next to >= 3 weight stages; otherwise the
  // configuration picked by measurement (NSM_UB_RINGS=<halo><src> overrides, e.g. 21, 12)
  kp.src_plane_bytes = uint32_t(up(size_t(kp.sbw) * kp.sbh * 128, 128));
  const size_t stage_bytes = size_t(NP) * a.Cmid * 128;
  const size_t fixed_tail = up(size_t(3 * a.Cmid + 3 * a.Cout + 68) * 4, 16) + kTabs * sizeof(JobTab) + 384;
  const size_t budget = size_t(227) * 1024 - 1024;
  auto fits = [&](int hb, int sb, int stages) {
    return up(size_t(hb) * NP * kHaloPlaneBytes + size_t(sb) * NP * kp.src_plane_bytes, 1024) + stages * stage_bytes +
               fixed_tail <= budget;
  };
  // 3x3 accumulator double-buffered when [2 x 3x3 | 1x1 | operand] fit the 512 TMEM columns (everything but fp32 / 128 ch)
  const int nch = a.Cmid / 64;
  const uint32_t tm_need2 = uint32_t(2 * NP * a.Cmid + NP * a.Cout + NP * a.Cmid / 2);
  static const bool db_off = getenv("NSM_UB_NO_DB") != nullptr;
  bool dbuf = tm_need2 <= 512 && !db_off;
  int hb = dbuf ? 2 * nch : 2, sb = 2;
  if (dbuf && !fits(hb, 1, 2)) { dbuf = false; hb = 2; }
  if (!fits(hb, 2, 3)) sb = 1;   // measured (profiles/README.md): one halo buffer + two source buffers is slower than 2 + 1
  {
    static const int force = getenv("NSM_UB_RINGS") ? atoi(getenv("NSM_UB_RINGS")) : 0;
    if (!dbuf && force >= 11 && force <= 22 && force % 10 >= 1 && force % 10 <= 2) { hb = force / 10; sb = force % 10; }
  }
  kp.halo_bufs = uint32_t(hb); kp.src_bufs = uint32_t(sb); kp.acc1_bufs = dbuf ? 2u : 1u;
  size_t off = size_t(hb) * NP * kHaloPlaneBytes;
  kp.off_src = uint32_t(off = up(off, 128));
  off += size_t(sb) * NP * kp.src_plane_bytes;
  kp.off_b = uint32_t(off = up(off, 1024));
  kp.b_stage_bytes = uint32_t(stage_bytes);
  if (off + fixed_tail + 2 * kp.b_stage_bytes > budget) {
    set_error("upblock: shared-memory budget exceeded (source box %dx%d)", kp.sbh, kp.sbw);
    return 1;
  }
  size_t stages = (budget - off - fixed_tail) / kp.b_stage_bytes;
  // all 9 * nch + nch weight tiles resident (conv9 in bf16 mode: 10 stages of 8 KB) where they fit: loaded once per CTA
  static const bool no_resident = getenv("NSM_UB_NO_RESIDENT") != nullptr;
  const size_t all_tiles = size_t(10) * nch;
  kp.w_resident = (!no_resident && all_tiles <= 16 && stages >= all_tiles) ? 1u : 0u;
  if (kp.w_resident) stages = all_tiles;
  if (stages > 8 && !kp.w_resident) stages = 8;
  {
    static const int force = getenv("NSM_UB_BSTAGES") ? atoi(getenv("NSM_UB_BSTAGES")) : 0;
    if (force >= 2 && size_t(force) < stages && !kp.w_resident) stages = size_t(force);
  }
  kp.b_stages = uint32_t(stages);
  off += stages * kp.b_stage_bytes;
  kp.off_vec = uint32_t(off = up(off, 16));
  off += up(size_t(3 * a.Cmid + 3 * a.Cout + 68) * 4, 16);
  kp.off_taps = uint32_t(off);
  off += kTabs * sizeof(JobTab);
  kp.off_bar = uint32_t(off = up(off, 8));
  off += 384;
  const size_t smem_bytes = off + 1024;
  // tensor memory: [3x3 main | 3x3 cross] [1x1 main | 1x1 cross] [A2 hi | A2 lo]
  kp.tm_acc2 = uint32_t(kp.acc1_bufs * NP * a.Cmid);
  kp.tm_a2 = kp.tm_acc2 + uint32_t(NP * a.Cout);
  const uint32_t need = kp.tm_a2 + uint32_t(NP * a.Cmid / 2);
  kp.tmem_cols = need <= 128 ? 128 : (need <= 256 ? 256 : 512);
  if (need > 512) {
    set_error("upblock: %u tensor-memory columns needed", need);
    return 1;
  }
  const uint32_t ef = NP == 2 ? kFmtF16 : kFmtBF16;   // fp16 / e4m3 share descriptor code 0
  kp.idesc1 = make_idesc_f16(128, a.Cmid, ef, ef, 0, 0);
  kp.idesc2c = make_idesc_f16(128, a.Cout, ef, ef, 0, 0);
  kp.idesc2w = make_idesc_f16(128, 2 * a.Cout, ef, ef, 0, 0);
  kp.bias3 = a.bias3; kp.scale3 = a.scale3; kp.shift3 = a.shift3;
  kp.bias1 = a.bias1; kp.scale1 = a.scale1; kp.shift1 = a.shift1;
  kp.w10 = a.w10; kp.b10 = a.b10;
  kp.out = a.out; kp.residual = a.residual; kp.y = a.y; kp.y_u8 = a.y_u8;
  if (a.tail && (!a.w10 || !a.b10 || (!a.y && !a.y_u8))) {
    set_error("upblock: the tail needs conv10's weights and an output");
    return 1;
  }
  CUtensorMap maps[6];
  memset(maps, 0, sizeof(maps));
  const uint64_t sdims[4] = {uint64_t(a.Cmid), uint64_t(a.Ws), uint64_t(a.Hs), uint64_t(a.N)};
  const uint64_t sstr[3] = {uint64_t(a.Cmid) * 2, uint64_t(a.Ws) * a.Cmid * 2, uint64_t(a.Hs) * a.Ws * a.Cmid * 2};
  const uint32_t sbox[4] = {64, uint32_t(kp.sbw), uint32_t(kp.sbh), 1};
  const uint64_t K3 = uint64_t(9) * a.Cmid;
  const uint64_t w3dims[2] = {K3, uint64_t(a.Cmid)};
  const uint64_t w3str[1] = {K3 * 2};
  const uint32_t w3box[2] = {64, uint32_t(a.Cmid)};
  const uint64_t w1dims[2] = {uint64_t(a.Cmid), uint64_t(a.Cout)};
  const uint64_t w1str[1] = {uint64_t(a.Cmid) * 2};
  const uint32_t w1box[2] = {64, uint32_t(a.Cout)};
  for (int pl = 0; pl < NP; ++pl) {
    if (!a.src.p[pl] || !a.w3.p[pl] || !a.w1.p[pl]) {
      set_error("upblock: null operand plane %d", pl);
      return 1;
    }
    if (encode_tmap_tiled(&maps[pl], a.src.p[pl], 4, sdims, sstr, sbox, 2, 0)) return 1;
    if (encode_tmap_tiled(&maps[2 + pl], a.w3.p[pl], 2, w3dims, w3str, w3box, 2, 128)) return 1;
    if (encode_tmap_tiled(&maps[4 + pl], a.w1.p[pl], 2, w1dims, w1str, w1box, 2, 128)) return 1;
  }
  if (NP == 1) { maps[1] = maps[0]; maps[3] = maps[2]; maps[5] = maps[4]; }
  const int grid = kp.total_tiles < g_ub_sms ? kp.total_tiles : g_ub_sms;
  if (NP == 2) {
    return a.Cmid == 128 ? launch_ub<2, 128>(maps, kp, grid, smem_bytes, st) : launch_ub<2, 64>(maps, kp, grid, smem_bytes, st);
  }
  return a.Cmid == 128 ? launch_ub<1, 128>(maps, kp, grid, smem_bytes, st) : launch_ub<1, 64>(maps, kp, grid, smem_bytes, st);
}

}  // namespace nsm
