// Weight-gradient GEMM on tcgen05:   dW[co][tap][ci] = sum_pixels dz[p][co] * x[p + tap][ci]
//
//   M = 128 output channels (TMEM lanes), N = BN input channels (TMEM columns), K = pixels.
//   Both operands are read from NHWC tensors whose CHANNEL axis is contiguous, i.e. they are "MN-major" UMMA operands:
//   one TMA box {64 ch, 16, 8, 1} = 128 pixels (K) x 64 channels (128-byte swizzled rows); 128 channels = two boxes
//   16 KB apart (descriptor LBO), consecutive 8-pixel groups are 1024 B apart (SBO).  The x box is shifted by the filter
//   tap exactly like in the forward kernel, so zero padding and ragged tiles come from the TMA's out-of-bounds fill.
//
//   BN = 256 (wide layers, single plane): 64-pixel k-blocks (box {64 ch, 16, 4, 1}) keep four pipeline stages in shared
//   memory and eight epilogue warps (two column halves) keep the chunk sums in registers; per MMA it moves 1.5 KB of
//   operands per 2 M MACs against 2 KB at BN = 128, which is what bounds this kernel (shared-memory bandwidth).
//
//   Work item = (co tile, ci tile, tap, K split).  The pixel axis is long (up to 2 M), so it is split over CTAs (split-K
//   partial results reduced by wgrad_reduce_kernel, deterministic) AND chunked inside a CTA: the tensor core's truncating
//   fp32 accumulator is drained every kWChunkPix pixels and the chunk results are added in fp32 registers with
//   round-to-nearest (see conv_gemm.cu).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "nsm_common.cuh"
#include "train_kernels.cuh"

namespace nsm {

constexpr int kWChunkPix = 1024;  // pixels per TMEM accumulation chunk

struct WgradKernelParams {
  int N, H, W, Cout, Cin, taps;
  int tiles_x, tiles_y, patches;   // patches = N * tiles_y * tiles_x   (k-blocks of KPIX pixels)
  int co_blocks, ci_blocks, splits, patches_per_split, total_items;
  uint32_t idesc;
  float* partial;                  // [splits][Cout][taps][Cin] fp32
};

// CG = 2: cta_group::2 pairs.  Two CTAs compute two output-channel tiles (M = 256) against one input-channel block of
// which each holds HALF the columns in shared memory: per 64-pixel k-block a CTA writes 32 KB and its tensor core reads
// 32 KB (16 KB of dz, 16 KB of x) where the single-CTA 256-wide tile moves 48 + 48 KB -- the kernel is bound by exactly
// that shared-memory traffic (ncu: tensor pipe 59-65 % active).  The leader issues the MMAs for both.
template <int BN, int NP, int CG = 1>
struct WgradCfg {
  static constexpr int KPIX = BN == 256 ? 64 : 128;           // pixels per k-block (patch = 16 x KPIX/16)
  static constexpr int TILE_H = KPIX / kTileW;
  static constexpr int KSTEPS = KPIX / 16;
  static constexpr int CHUNK = kWChunkPix / KPIX;             // k-blocks per accumulation chunk
  static constexpr int EPI_WARPS = BN == 256 ? 8 : 4;
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static constexpr int BOX_BYTES = KPIX * 128;                // KPIX pixels x 64 channels x 2 B
  static constexpr int A_BYTES = 2 * BOX_BYTES;               // 128 output channels
  static constexpr int B_BYTES = (BN / 64 / CG) * BOX_BYTES;   // pairs: this CTA's half of the input-channel block
  static constexpr int STAGE_BYTES = NP * (A_BYTES + B_BYTES);
  static constexpr int BUDGET = 227 * 1024 - 1024 - 256;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int ACC_COLS = NP * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "bad TMEM column count");
};

__device__ __forceinline__ void wgrad_decode(const WgradKernelParams& p, int item, int& cob, int& cib, int& tap,
                                             int& split) {
  cib = item % p.ci_blocks;
  int t = item / p.ci_blocks;
  cob = t % p.co_blocks;
  t /= p.co_blocks;
  tap = t % p.taps;
  split = t / p.taps;
}

template <int BN, int NP, int CG>
__global__ void __launch_bounds__(WgradCfg<BN, NP, CG>::THREADS, 1)
wgrad_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                  const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ WgradKernelParams p) {
  using Cfg = WgradCfg<BN, NP, CG>;
  constexpr bool PAIR = CG == 2;
  static_assert(!PAIR || NP == 1, "CTA pairs: single-plane operands");
  const int cta_rank = PAIR ? int(cluster_ctarank()) : 0;
  // work items of this CTA (pair): p.co_blocks counts tiles of 128 * CG output channels
  const int it_first = int(blockIdx.x) / CG, it_step = int(gridDim.x) / CG;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], Cfg::EPI_WARPS * CG);   // pairs: the leader's copy collects both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (uniform control flow for the warp, one elected lane issues the copies: straight-line UTMALDG)
    {
      const bool leader = elect_one();
      uint32_t stage = 0, phase = 0;
      for (int item = it_first; item < p.total_items; item += it_step) {
        int cob, cib, tap, split;
        wgrad_decode(p, item, cob, cib, tap, split);
        cob = cob * CG + cta_rank;                       // this CTA's tile of 128 output channels
        const int ci0 = cib * BN + cta_rank * (BN / CG); // ... and its share of the input-channel block
        const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
        const int dx = p.taps == 9 ? tap % 3 - 1 : 0;
        const int pt0 = split * p.patches_per_split;
        const int pt1 = pt0 + p.patches_per_split < p.patches ? pt0 + p.patches_per_split : p.patches;
        for (int pt = pt0; pt < pt1; ++pt) {
          const int tx = pt % p.tiles_x;
          const int t2 = pt / p.tiles_x;
          const int ty = t2 % p.tiles_y;
          const int n = t2 / p.tiles_y;
          const int x0 = tx * kTileW, y0 = ty * Cfg::TILE_H;
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + NP * Cfg::A_BYTES;
          if (leader) {
            if (PAIR) {   // both CTAs' boxes complete on the leader's barrier, which expects the bytes of the pair
              if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
#pragma unroll
              for (int g = 0; g < 2; ++g)
                tma_load_4d_pair(sa + g * Cfg::BOX_BYTES, &tmA0, &full_bar[stage], cob * 128 + g * 64, x0, y0, n);
#pragma unroll
              for (int g = 0; g < BN / 64 / CG; ++g)
                tma_load_4d_pair(sb + g * Cfg::BOX_BYTES, &tmB0, &full_bar[stage], ci0 + g * 64, x0 + dx, y0 + dy, n);
            } else {
              mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
              for (int pl = 0; pl < NP; ++pl) {
                const CUtensorMap* ma = pl == 0 ? &tmA0 : &tmA1;
                const CUtensorMap* mb = pl == 0 ? &tmB0 : &tmB1;
#pragma unroll
                for (int g = 0; g < 2; ++g)   // dz at the output pixel; channels beyond Cout are zero-filled
                  tma_load_4d(sa + pl * Cfg::A_BYTES + g * Cfg::BOX_BYTES, ma, &full_bar[stage], cob * 128 + g * 64, x0,
                              y0, n);
#pragma unroll
                for (int g = 0; g < BN / 64; ++g)   // x at the tap-shifted pixel
                  tma_load_4d(sb + pl * Cfg::B_BYTES + g * Cfg::BOX_BYTES, mb, &full_bar[stage], cib * BN + g * 64,
                              x0 + dx, y0 + dy, n);
              }
            }
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // uniform control flow for the warp, one elected lane issues: straight-line UTCHMMA with descriptors advanced by one
    // 64-bit add (under `if (lane == 0)` every tcgen05 instruction sat in an ELECT / BRA.U.ANY loop, see conv_gemm.cu)
    if (cta_rank == 0) {   // pairs: the leader issues for both
      const bool leader = elect_one();
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint64_t desc_hi = make_desc_sw128(0, Cfg::BOX_BYTES, 1024);
      auto desc_of = [&](uint32_t addr) { return desc_hi | uint64_t((addr & 0x3FFFFu) >> 4); };
      for (int item = it_first; item < p.total_items; item += it_step) {
        int cob, cib, tap, split;
        wgrad_decode(p, item, cob, cib, tap, split);
        const int pt0 = split * p.patches_per_split;
        const int pt1 = pt0 + p.patches_per_split < p.patches ? pt0 + p.patches_per_split : p.patches;
        for (int c0 = pt0; c0 < pt1; c0 += Cfg::CHUNK) {
          const int c1 = c0 + Cfg::CHUNK < pt1 ? c0 + Cfg::CHUNK : pt1;
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_main = tmem_base + acc * Cfg::ACC_COLS;
          const uint32_t d_cross = d_main + BN;
          for (int pt = c0; pt < c1; ++pt) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t b_hi = a_hi + NP * Cfg::A_BYTES;
            const uint64_t da_hi = desc_of(a_hi), db_hi = desc_of(b_hi);
            const uint64_t da_lo = desc_of(a_hi + Cfg::A_BYTES), db_lo = desc_of(b_hi + Cfg::B_BYTES);
            if (leader) {
#pragma unroll
              for (int k = 0; k < Cfg::KSTEPS; ++k) {   // UMMA_K = 16 pixels = 16 rows of 128 B = 2048 B per step (+128 in the address field)
                const uint32_t accum = ((pt - c0) | k) != 0 ? 1u : 0u;
                const uint64_t step = uint64_t(k) * (2048 >> 4);
                if (PAIR) umma_bf16_pair(d_main, da_hi + step, db_hi + step, p.idesc, accum);
                else umma_bf16(d_main, da_hi + step, db_hi + step, p.idesc, accum);
                if (NP == 2) {
                  umma_bf16(d_cross, da_hi + step, db_lo + step, p.idesc, accum);
                  umma_bf16(d_cross, da_lo + step, db_hi + step, p.idesc, 1u);
                }
              }
              if (PAIR) umma_commit_pair(&empty_bar[stage]);
              else umma_commit(&empty_bar[stage]);
              if (pt + 1 == c1) {
                if (PAIR) umma_commit_pair(&tfull_bar[acc]);
                else umma_commit(&tfull_bar[acc]);
              }
            }
            __syncwarp();
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..): chunk sums in registers, fp32 partial tile to global ==============
    // warp -> TMEM lane quarter (warp & 3, a hardware restriction) and, with eight warps, a column half
    constexpr int CW = BN / (Cfg::EPI_WARPS / 4);   // columns per epilogue warp
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;   // output channel inside the co tile
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16) + half * CW;
    uint32_t acc = 0, acc_phase = 0;
    for (int item = it_first; item < p.total_items; item += it_step) {
      int cob, cib, tap, split;
      wgrad_decode(p, item, cob, cib, tap, split);
      cob = cob * CG + cta_rank;
      const int pt0 = split * p.patches_per_split;
      const int pt1 = pt0 + p.patches_per_split < p.patches ? pt0 + p.patches_per_split : p.patches;
      const int nchunks = (pt1 - pt0 + Cfg::CHUNK - 1) / Cfg::CHUNK;
      float sum[CW];
#pragma unroll
      for (int j = 0; j < CW; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < CW; c0 += 32) {
          uint32_t r0[32];
          tmem_ld_32x32(lane_addr + acc * Cfg::ACC_COLS + c0, r0);
          if (NP == 2) {
            uint32_t r1[32];
            tmem_ld_32x32(lane_addr + acc * Cfg::ACC_COLS + BN + c0, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[c0 + j] += __uint_as_float(r0[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(&tempty_bar[acc]);
          else mbar_arrive(&tempty_bar[acc]);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      const int co = cob * 128 + row;
      if (co < p.Cout) {
        float* o = p.partial + (((size_t)split * p.Cout + co) * p.taps + tap) * p.Cin + cib * BN + half * CW;
#pragma unroll
        for (int j = 0; j < CW; j += 4)
          *reinterpret_cast<float4*>(o + j) = make_float4(sum[j], sum[j + 1], sum[j + 2], sum[j + 3]);
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();   // nobody leaves while the peer may still signal this CTA's barriers / read its smem
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// dw[co][ci][tap] (OIHW, un-padded) = sum_splits partial[split][co][tap][ci]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int Cout,
                                                           int Cin, int taps, int Cout_real, int Cin_real, int rb,
                                                           float* __restrict__ dw) {
  const long long total = (long long)Cout_real * Cin_real * taps;
  const size_t split_stride = (size_t)Cout * taps * Cin;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int tap = int(i % taps);
    long long t = i / taps;
    const int ci = int(t % Cin_real);
    const int co = int(t / Cin_real);
    const float* src = partial + ((size_t)co * taps + tap) * Cin + ci;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += src[k * split_stride];
    dw[i] = rb ? rbf(s) : s;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct WgradPlan {
  int BN, CG, tile_h, chunk, splits, patches_per_split, tiles_x, tiles_y, patches, co_blocks, ci_blocks, items;
};

static WgradPlan wgrad_plan(const WgradShape& s) {
  WgradPlan pl;
  // measured (profiles/): the 256-wide tile wins for Cin = 1024 (+10 %), the 128-wide one for Cin = 512 (+7 %)
  pl.BN = s.fmt != kFmtBf16 ? 64 : ((s.Cin % 256 == 0 && s.Cin >= 1024) ? 256 : (s.Cin % 128 == 0 ? 128 : 64));
  {
    static const char* force = getenv("NSM_WGRAD_BN");   // tuning / A-B runs
    const int f = force ? atoi(force) : 0;
    if (s.fmt == kFmtBf16 && (f == 64 || f == 128 || f == 256) && s.Cin % f == 0) pl.BN = f;
  }
  // CTA pairs (M = 256 x N = 256) wherever both channel counts allow it (NSM_NO_WGRAD_PAIR=1: single CTAs)
  static const bool pair_off = getenv("NSM_NO_WGRAD_PAIR") != nullptr;
  pl.CG = (!pair_off && s.fmt == kFmtBf16 && s.Cin % 256 == 0 && s.Cout % 256 == 0) ? 2 : 1;
  if (pl.CG == 2) pl.BN = 256;
  pl.tile_h = pl.BN == 256 ? 4 : kTileH;
  pl.chunk = kWChunkPix / (kTileW * pl.tile_h);
  pl.tiles_x = (s.W + kTileW - 1) / kTileW;
  pl.tiles_y = (s.H + pl.tile_h - 1) / pl.tile_h;
  pl.patches = s.N * pl.tiles_x * pl.tiles_y;
  pl.co_blocks = (s.Cout + 128 * pl.CG - 1) / (128 * pl.CG);   // tiles of 128 (256 for pairs) output channels
  pl.ci_blocks = s.Cin / pl.BN;
  const int base = pl.co_blocks * pl.ci_blocks * s.taps;
  // K splits: enough work items for every CTA (pair), chosen so that the last wave is nearly full
  const int workers = 148 / pl.CG;
  const int max_splits = (pl.patches + pl.chunk - 1) / pl.chunk;
  int s_lo = (workers + base - 1) / base;
  if (s_lo < 1) s_lo = 1;
  int splits = s_lo;
  double best = 0.0;
  for (int cand = s_lo; cand <= 3 * s_lo + 2 && cand <= max_splits; ++cand) {
    const int items = base * cand;
    const int rounds = (items + workers - 1) / workers;
    const double eff = double(items) / (double(rounds) * workers) - 0.004 * (cand - s_lo);   // mild cost per extra split
    if (eff > best) {
      best = eff;
      splits = cand;
    }
    if (eff >= 0.96) break;
  }
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  pl.patches_per_split = (pl.patches + splits - 1) / splits;
  pl.splits = (pl.patches + pl.patches_per_split - 1) / pl.patches_per_split;
  pl.items = base * pl.splits;
  return pl;
}

size_t wgrad_workspace_bytes(const WgradShape& s) {
  const WgradPlan pl = wgrad_plan(s);
  return (size_t)pl.splits * s.Cout * s.taps * s.Cin * 4;
}

template <int BN, int NP, int CG = 1>
static int wgrad_launch_t(const CUtensorMap* maps, const WgradKernelParams& kp, int grid, cudaStream_t stream) {
  using Cfg = WgradCfg<BN, NP, CG>;
  auto kern = wgrad_gemm_kernel<BN, NP, CG>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(wgrad_gemm<%d,%d>): %s", BN, NP, cudaGetErrorString(e));
      return 1;
    }
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  if (CG == 2) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], kp);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("wgrad_gemm<%d,%d> launch failed: %s", BN, NP, cudaGetErrorString(e));
    return 1;
  }
  count_launch(2);   // GEMM + split-K reduce
  return 0;
}

int wgrad_launch(const WgradShape& s, const Planes& dz, const Planes& x, void* workspace, size_t workspace_bytes,
                 int Cout_real, int Cin_real, int round_bf16, float* dw, cudaStream_t st) {
  if (s.Cin % 64 || s.Cout % 64 || (s.taps != 1 && s.taps != 9) || s.fmt == kFmtF16x2 || s.fmt < 0 || s.fmt > 2) {
    set_error("wgrad: unsupported shape Cout=%d Cin=%d taps=%d fmt=%d (training uses fmt 0 or 2)", s.Cout, s.Cin, s.taps,
              s.fmt);
    return 1;
  }
  const WgradPlan pl = wgrad_plan(s);
  if (workspace_bytes < wgrad_workspace_bytes(s)) {
    set_error("wgrad: workspace %zu B < required %zu B", workspace_bytes, wgrad_workspace_bytes(s));
    return 1;
  }
  const int planes = fmt_planes(s.fmt);
  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  const uint32_t box[4] = {64, uint32_t(kTileW), uint32_t(pl.tile_h), 1};
  const uint64_t adims[4] = {uint64_t(s.Cout), uint64_t(s.W), uint64_t(s.H), uint64_t(s.N)};
  const uint64_t astr[3] = {uint64_t(s.Cout) * 2, uint64_t(s.W) * s.Cout * 2, uint64_t(s.H) * s.W * s.Cout * 2};
  const uint64_t bdims[4] = {uint64_t(s.Cin), uint64_t(s.W), uint64_t(s.H), uint64_t(s.N)};
  const uint64_t bstr[3] = {uint64_t(s.Cin) * 2, uint64_t(s.W) * s.Cin * 2, uint64_t(s.H) * s.W * s.Cin * 2};
  for (int p = 0; p < planes; ++p) {
    if (!dz.p[p] || !x.p[p]) {
      set_error("wgrad: null operand plane %d", p);
      return 1;
    }
    if (encode_tmap_tiled(&maps[p], dz.p[p], 4, adims, astr, box, 2)) return 1;
    if (encode_tmap_tiled(&maps[2 + p], x.p[p], 4, bdims, bstr, box, 2)) return 1;
  }
  if (planes == 1) {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  WgradKernelParams kp;
  kp.N = s.N; kp.H = s.H; kp.W = s.W; kp.Cout = s.Cout; kp.Cin = s.Cin; kp.taps = s.taps;
  kp.tiles_x = pl.tiles_x; kp.tiles_y = pl.tiles_y; kp.patches = pl.patches;
  kp.co_blocks = pl.co_blocks; kp.ci_blocks = pl.ci_blocks; kp.splits = pl.splits;
  kp.patches_per_split = pl.patches_per_split; kp.total_items = pl.items;
  kp.idesc = make_idesc_f16(128 * pl.CG, pl.BN, kFmtBF16, kFmtBF16, 1, 1);   // both operands MN-major
  kp.partial = reinterpret_cast<float*>(workspace);
  int sms = 148;
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int grid = kp.total_items < sms ? kp.total_items : sms;
  if (pl.CG == 2) grid = 2 * (kp.total_items < sms / 2 ? kp.total_items : sms / 2);
  int rc;
  if (pl.CG == 2) rc = wgrad_launch_t<256, 1, 2>(maps, kp, grid, st);
  else if (planes == 1)
    rc = pl.BN == 256   ? wgrad_launch_t<256, 1>(maps, kp, grid, st)
         : pl.BN == 128 ? wgrad_launch_t<128, 1>(maps, kp, grid, st)
                        : wgrad_launch_t<64, 1>(maps, kp, grid, st);
  else rc = wgrad_launch_t<64, 2>(maps, kp, grid, st);
  if (rc) return rc;
  const long long total = (long long)Cout_real * Cin_real * s.taps;
  long long g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  wgrad_reduce_kernel<<<(unsigned)g, 256, 0, st>>>(kp.partial, pl.splits, s.Cout, s.Cin, s.taps, Cout_real, Cin_real,
                                                   round_bf16, dw);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("wgrad_reduce launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

}  // namespace nsm
