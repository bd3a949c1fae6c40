// Implicit-GEMM convolution (3x3 pad 1 / 1x1, stride 1) for NHWC bf16 "planes" on Blackwell tensor cores.
//
//   GEMM view:  D[pixel, cout] = sum_{tap, cin} A[pixel + tap, cin] * Wp[cout, tap, cin]
//   M tile  = 8 x 16 spatial patch (128 pixels)           -> TMEM lanes
//   N tile  = BN output channels (64 / 128 / 256)          -> TMEM columns (fp32 accumulators, 2 stages)
//   K block = one filter tap x 64 input channels (128-byte swizzled smem rows)
//
// A tiles are fetched by ONE 4-D tiled TMA box {64 ch, 16, 8, 1} whose (x, y) start is shifted by the tap
// offset; out-of-bounds (incl. negative) coordinates are zero-filled by the TMA unit, which is exactly the
// convolution's zero padding and also pads ragged right/bottom tiles.  The box lands in shared memory as 128
// rows of 128 B with the 128-byte swizzle = the canonical K-major UMMA operand layout, so no im2col buffer,
// no index math and no smem transform is needed.  B tiles (weights, pre-packed [Cout][tap][Cin]) are 2-D boxes.
//
// Warp roles (320 threads, 1 CTA / SM, persistent over work items):
//   warp 0    TMA producer          (smem full/empty mbarrier ring)
//   warp 1    tcgen05.mma issuer    (single thread; accumulators double-buffered in TMEM)
//   warps 2-9 epilogue              (tcgen05.ld -> bias, [BN statistics], BN affine, LeakyReLU, skip add, 2x2 avg-pool,
//                                    bf16 / hi+lo split); two warps per TMEM lane quarter, each owning half of the
//                                    tile's columns.  Results leave, and the skip tensor arrives, through per-warp
//                                    swizzled staging tiles and TMA (cp.async.bulk.tensor) -- no per-lane global
//                                    accesses, ragged tiles are clipped by the TMA unit.
//
// fp32 mode ("planes == 2"): activations and weights are stored as hi + lo planes (value = hi + lo) and the issuer
// accumulates a_hi*w_hi into a main and a_hi*w_lo + a_lo*w_hi into a cross TMEM accumulator (the dropped a_lo*w_lo is
// <= 2^-22), in chunks of K that the epilogue warps add in fp32 registers because the tensor core's accumulator
// truncates.  Three operand encodings:
//   fp16 hi + fp16 lo (eval)  : a_hi x [w_hi | w_lo] as ONE MMA of width 2*BN, then a_lo x w_hi     (3 MMA slots / k-step)
//   bf16 hi + bf16 lo (train) : same, full fp32 range
//   fp16 hi + 8-bit cross     : a_hi x w_hi, then ONE e4m3 MMA of K = 32 for both cross terms        (2 MMA slots / k-step)
// The 256-wide 8-bit-cross layers (conv6 / conv7 3x3 in fp32 mode, 79 % of the FLOPs) run conv_gemm_wide_kernel as
// cta_group::2 pairs: M = 256 (two pixel tiles) x N = 256, each CTA holds half of the weight tile.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <unordered_map>

#include "conv_gemm.cuh"
#include "nsm_common.cuh"

namespace nsm {

// ------------------------------------------------------------------------------------------------
// error string (thread local), shared by all translation units
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ------------------------------------------------------------------------------------------------
// TMA descriptor encoding through the driver entry point
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// A descriptor is a pure function of (base, rank, dims, strides, box, element size, swizzle): the same buffers come back
// every frame / training step (caller-owned workspaces, packed weights), so encoded descriptors are kept in a small table
// behind a mutex and a launch costs one lookup instead of a driver call per operand (6-10 per conv launch).
struct TmapKey {
  uint64_t w[16];
  bool operator==(const TmapKey& o) const { return memcmp(w, o.w, sizeof(w)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (uint64_t v : k.w) {
      h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
      h *= 0xFF51AFD7ED558CCDull;
    }
    return size_t(h ^ (h >> 33));
  }
};
static std::mutex g_tmap_mu;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::atomic<long long> g_tmap_hits{0}, g_tmap_misses{0};
constexpr size_t kTmapCacheMax = 65536;  // entries of 128 B + key (<= 17 MB of host memory); flushed when full
void tmap_cache_stats(long long* hits, long long* misses) {
  *hits = g_tmap_hits.load(std::memory_order_relaxed);
  *misses = g_tmap_misses.load(std::memory_order_relaxed);
}

int encode_tmap_tiled(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box, int elem_bytes, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return 1;
  }
  if (rank < 1 || rank > 5) {
    set_error("encode_tmap_tiled: rank %d", rank);
    return 1;
  }
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.w[0] = uint64_t(reinterpret_cast<uintptr_t>(base));
  key.w[1] = uint64_t(rank) | (uint64_t(elem_bytes) << 8) | (uint64_t(swizzle_bytes) << 16);
  for (int i = 0; i < rank; ++i) {
    key.w[2 + i] = dims[i];
    key.w[7 + i] = box[i];
    if (i > 0) key.w[11 + i] = strides_bytes[i - 1];
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      memcpy(map, &it->second, sizeof(CUtensorMap));
      g_tmap_hits.fetch_add(1, std::memory_order_relaxed);
      return 0;
    }
  }
  g_tmap_misses.fetch_add(1, std::memory_order_relaxed);
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle_bytes == 0    ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, dt, rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu %llu box %u %u)", int(r), rank,
              (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return 1;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
    memcpy(&g_tmap_cache[key], map, sizeof(CUtensorMap));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel
// ------------------------------------------------------------------------------------------------
struct ConvKernelParams {
  int N, H, W, Cin, Cout, taps;
  int tiles_x, tiles_y, n_blocks, total_items, kc_per_tap;
  int total_pairs;    // CTA-pair kernels: (pixel-tile pairs) x n_blocks
  int fmt;
  int chunk_kb;       // fp32 modes: k-blocks per TMEM accumulation chunk
  int x8;             // plane 1 of both operands holds 8-bit cross-term operands (kFmtF16X8, nsm_common.cuh)
  int w_evict_last;   // CTA-pair kernel: weight boxes are loaded with the L2 evict_last policy (NSM_NO_W_EVICT_LAST=1: off)
  float cross_scale;  // factor of the cross accumulator when the chunk results are summed (2^-17 with x8, else 1)
  uint32_t idesc_hi;    // M = 128 (256 for CTA pairs), N = BN
  uint32_t idesc_wide;  // M = 128, N = 2*BN: hi+lo modes, a_hi x [w_hi | w_lo] in one instruction
  ConvEpilogue ep;
};

// EW = epilogue warps: 8 (two column halves x four TMEM lane quarters), or 16 (four column quarters) for the layers with a short
// K loop.  Those are bound by their epilogue -- ~650 instructions per 32-column group in long dependent chains; with two
// epilogue warps per scheduler the issue slots are 30 % used (profiles/r02_thin_conv_stalls.txt) -- so they get four warps
// per scheduler at 96 registers (the bf16 epilogue needs no more; the hi+lo one keeps a quarter of the chunk sums).
template <int BN, int NP, int EW = 8>
struct GemmCfg {
  static constexpr int TW = kTileW, TH = kTileH;
  static constexpr int THREADS = 64 + 32 * EW;
  static constexpr int A_BYTES = 128 * 128;  // 128 pixels x 64 elements (2 B)
  static constexpr int B_BYTES = BN * 128;   // weight tile rows x 64 elements
  static constexpr int STAGE_BYTES = NP * (A_BYTES + B_BYTES);
  // 8 epilogue warps x 4 KB staging tiles for the TMA stores of the output (32 pixels x 32 channels x {hi, lo})
  static constexpr int STAGING_BYTES = EW * 4096;
  static constexpr int BUDGET = 227 * 1024 - 1024 /*align slack*/ - 512 /*barriers*/ - STAGING_BYTES;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  // fp32 mode keeps two accumulators per stage: [main = a_hi*w_hi | cross = a_hi*w_lo + a_lo*w_hi]
  static constexpr int ACC_COLS = NP * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 512;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns must be a power of two <= 512");
};

// fp32 mode: the tensor core adds into its fp32 accumulator with truncation, so the error of one accumulator grows like
// (#MMAs) * 2^-24 (measured: ~1e-4 relative after K = 9216).  The K loop is therefore cut into chunks of at most
// kChunkKB k-blocks; every chunk accumulates from zero in TMEM and the epilogue warps add the chunk results into
// fp32 registers with round-to-nearest.
// Default 24 (round 2; 16 in round 1).  Every drain of the single-buffered 256-wide accumulator stalls the MMA issuer: conv6 /
// conv7 3x3 take 0.879 / 0.917 ms at 16, 0.859 / 0.882 at 24, 0.840 / 0.885 at 32.  With the cross terms in their own
// accumulator the output error depends only weakly on the chunk length (1080p frame of the parity test: 3.6e-5 at 16 / 24 / 32,
// 5.3e-5 unchunked; a 257 x 259 frame with other BatchNorm draws: 6.1e-5 at 16, 7.4e-5 at 32) -- 24 keeps most of the speed and
// most of the margin to the 1e-4 bound.
constexpr int kChunkKBDefault = 24;
constexpr int kConvThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two column halves x four lane quarters)

template <int TW, int TH>
__device__ __forceinline__ void decode_item(const ConvKernelParams& p, int item, int& n, int& y0, int& x0,
                                            int& nb) {
  nb = item % p.n_blocks;
  int t = item / p.n_blocks;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  n = t / p.tiles_y;
  y0 = ty * TH;
  x0 = tx * TW;
}

// Where one epilogue warp's 32 pixels (2 patch rows x 16 columns) go: the output leaves through TMA stores from a 4 KB
// per-warp staging tile (one store instruction of 32 lanes x 16 B would touch 32 different cache lines and serialise in
// the load/store unit: the profile of the thin layers showed the epilogue waiting on exactly those stores).
// The skip tensor arrives the same way (TMA load into the slot its sum will leave from), issued before the warp waits for
// the accumulator so that its latency hides behind the MMAs.
struct EpiTile {
  uint32_t stg;              // shared-memory address of the warp's 4 KB staging area (1024-byte aligned)
  const CUtensorMap* out[2]; // box {32 ch, 16, 2, 1}, 64-byte swizzle
  const CUtensorMap* pool[2];// box {32 ch, 8, 1, 1}
  const CUtensorMap* res[2]; // as out
  uint64_t* rbar;            // this warp's two "residual landed" mbarriers (one per slot)
  uint32_t rphase[2];
  int x0, y0, n;             // pixel coordinates of the warp's first row
};
// Staging slots: hi+lo modes use the whole area per 32-channel group (hi at +0, lo at +2048); the single-plane mode has
// two 2 KB slots that alternate between groups.
template <int NP>
__device__ __forceinline__ uint32_t epi_slot(int g) { return NP == 2 ? 0u : uint32_t(g & 1) * 2048u; }
template <int NP>
__device__ __forceinline__ void epi_issue_residual(const EpiTile& et, int g, int cb) {  // one lane
  uint64_t* bar = et.rbar + (NP == 2 ? 0 : (g & 1));
  mbar_expect_tx(bar, NP * 2048);
  const uint32_t dst = et.stg + epi_slot<NP>(g);
  tma_load_4d_saddr(dst, et.res[0], bar, cb, et.x0, et.y0, et.n);
  if (NP == 2) tma_load_4d_saddr(dst + 2048, et.res[1], bar, cb, et.x0, et.y0, et.n);
}

// CTA-pair work item: pair index -> (column block, this CTA's pixel tile = 2 * tile-pair + rank).  A tile index past the end
// (odd tile count) decodes to n == N: its loads are zero-filled and its stores clipped by the TMA unit.
__device__ __forceinline__ void decode_pair(const ConvKernelParams& p, int pi, int rank, int& n, int& y0, int& x0,
                                            int& nb) {
  nb = pi % p.n_blocks;
  int t = 2 * (pi / p.n_blocks) + rank;
  const int tx = t % p.tiles_x;
  t /= p.tiles_x;
  const int ty = t % p.tiles_y;
  n = t / p.tiles_y;
  y0 = ty * kTileH;
  x0 = tx * kTileW;
}

// Epilogue math + stores for 32 consecutive output channels [cb, cb+32) of one pixel (one thread).
// TW = tile width: a warp's 32 pixels are 32/TW rows of TW columns (lane = row * TW + column).
template <int NP, int TW>
__device__ __forceinline__ void epilogue_cols(float (&v)[32], const ConvKernelParams& p, int cb, bool valid,
                                              size_t pix, EpiTile& et, int g, int ng, double& st1, double& st2) {
  const ConvEpilogue& ep = p.ep;
  const int lane = threadIdx.x & 31;
  const bool has_res = ep.residual.p[0] != nullptr;
  const uint32_t slot = et.stg + epi_slot<NP>(g);
  if (ep.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + cb + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (ep.out_f32 && valid) {
    float4* o = reinterpret_cast<float4*>(ep.out_f32 + pix * p.Cout + cb);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if (ep.round_bf16) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) rbf2(v[j], v[j + 1]);
  }
  if (ep.stats) {
    // Per-channel sum / sum of squares over the warp's 32 pixels by a butterfly "transpose reduction" (31 shuffles per
    // quantity): after the five halving steps lane l holds the totals of channel cb + l, which it keeps accumulating in
    // fp64 registers across the tiles of this CTA.
    float s1[32], s2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s1[j] = valid ? v[j] : 0.f;
      s2[j] = s1[j] * s1[j];
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const bool upper = (lane & off) != 0;
#pragma unroll
      for (int i = 0; i < off; ++i) {
        const float keep1 = upper ? s1[i + off] : s1[i], send1 = upper ? s1[i] : s1[i + off];
        const float keep2 = upper ? s2[i + off] : s2[i], send2 = upper ? s2[i] : s2[i + off];
        s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, off);
        s2[i] = keep2 + __shfl_xor_sync(0xffffffffu, send2, off);
      }
    }
    st1 += double(s1[0]);   // flushed into the channel's order-independent accumulator slot (nsm_common.cuh: Acc) when the CTA's column block changes / at the end
    st2 += double(s2[0]);
  }
  if (ep.scale) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(ep.scale + cb + j));
      const float4 t = __ldg(reinterpret_cast<const float4*>(ep.shift + cb + j));
      v[j] = fmaf(v[j], s.x, t.x); v[j + 1] = fmaf(v[j + 1], s.y, t.y);
      v[j + 2] = fmaf(v[j + 2], s.z, t.z); v[j + 3] = fmaf(v[j + 3], s.w, t.w);
    }
    if (ep.round_bf16) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) rbf2(v[j], v[j + 1]);
    }
  }
  if (ep.lrelu) {
    const float slope = ep.lrelu == 2 ? 0.f : 0.2f;   // 2 = ReLU (the VGG19 stacks of the perceptual term)
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : slope * v[j];
    if (ep.round_bf16) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) rbf2(v[j], v[j + 1]);
    }
  }
  if (has_res) {
    // the skip values of this group were requested earlier (tile start / after the slot's previous store)
    const int b = NP == 2 ? 0 : (g & 1);
    mbar_wait(et.rbar + b, et.rphase[b]);
    et.rphase[b] ^= 1;
    const uint32_t row = slot + lane * 64;
    const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 h = lds16(row + ((j ^ sw) << 4));
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[8 * j + 2 * e] += hi_lo_to_f32(hw[e], p.fmt);
        v[8 * j + 2 * e + 1] += hi_hi_to_f32(hw[e], p.fmt);
      }
      if (NP == 2) {
        const uint4 l = lds16(row + 2048 + ((j ^ sw) << 4));
        const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          v[8 * j + 2 * e] += lo_lo_to_f32(lw[e], p.fmt);
          v[8 * j + 2 * e + 1] += lo_hi_to_f32(lw[e], p.fmt);
        }
      }
    }
    if (ep.round_bf16) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) rbf2(v[j], v[j + 1]);
    }
  }
  if (ep.out.p[0]) {
    // registers -> swizzled staging tile (row = lane = pixel, 64 B of channels per plane) -> one TMA store per plane;
    // pixels outside the image are clipped by the TMA unit
    if (!has_res) {
      // the stores that last used this slot must have read it (with a residual the slot was already claimed for its
      // load, and every lane overwrites only the row it has just read)
      if (lane == 0) {
        if (NP == 2 || ng == 1) bulk_wait_read0(); else bulk_wait_read1();   // two alternating slots when ng is even
      }
      __syncwarp();
    }
    const uint32_t row = slot + lane * 64;
    const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hw[4], lw[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
        hw[e] = pack_hi(a, b, p.fmt);
        if (NP == 2) lw[e] = pack_lo_resid(a, b, hw[e], p.fmt);
      }
      sts16(row + ((j ^ sw) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
      if (NP == 2) sts16(row + 2048 + ((j ^ sw) << 4), make_uint4(lw[0], lw[1], lw[2], lw[3]));
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(et.out[0], slot, cb, et.x0, et.y0, et.n);
      if (NP == 2) tma_store_4d(et.out[1], slot + 2048, cb, et.x0, et.y0, et.n);
      bulk_commit();
    }
  }
  if (ep.pool.p[0]) {  // warp-uniform: AvgPool2d(2) over (lane^1, lane^TW) partners; anchors: even column of an even row
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      float s = v[j] + __shfl_xor_sync(0xffffffffu, v[j], 1);
      s += __shfl_xor_sync(0xffffffffu, s, TW);
      v[j] = s * 0.25f;
    }
    if (ep.round_bf16) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) rbf2(v[j], v[j + 1]);
    }
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if ((lane & (TW | 1)) == 0) {
      // pooled pixel index inside the warp's pooled box (TW/2 columns x 16/TW rows, x fastest)
      const int r = TW == 16 ? (lane >> 1) : (((lane >> 4) << 2) | ((lane >> 1) & 3));
      const uint32_t row = slot + r * 64;
      const uint32_t sw = (r >> 1) & 3;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = v[8 * j + 2 * e], b = v[8 * j + 2 * e + 1];
          hw[e] = pack_hi(a, b, p.fmt);
          if (NP == 2) lw[e] = pack_lo_resid(a, b, hw[e], p.fmt);
        }
        sts16(row + ((j ^ sw) << 4), make_uint4(hw[0], hw[1], hw[2], hw[3]));
        if (NP == 2) sts16(row + 512 + ((j ^ sw) << 4), make_uint4(lw[0], lw[1], lw[2], lw[3]));
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_4d(et.pool[0], slot, cb, et.x0 >> 1, et.y0 >> 1, et.n);
      if (NP == 2) tma_store_4d(et.pool[1], slot + 512, cb, et.x0 >> 1, et.y0 >> 1, et.n);
      bulk_commit();
    }
  }
  if (has_res) {
    // next group that will use this slot: request its skip values as soon as the store above has read the slot
    const int gs = NP == 2 ? 1 : 2;
    if (g + gs < ng && lane == 0) {
      bulk_wait_read0();
      epi_issue_residual<NP>(et, g + gs, cb + 32 * gs);
    }
  }
}

template <int BN, int NP, int EW>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                 const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                 const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmP1,
                 const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                 const __grid_constant__ ConvKernelParams p) {
  using Cfg = GemmCfg<BN, NP, EW>;
  constexpr int TW = Cfg::TW, TH = Cfg::TH;
  const int it_first = int(blockIdx.x), it_step = int(gridDim.x), it_count = p.total_items;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem_base = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);  // 1024-byte aligned (SWIZZLE_128B)
  uint8_t* smem = smem_base;                      // the stage ring
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;   // [EW epilogue warps][2 slots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * EW);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (NP == 2) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    if (p.ep.out.p[0]) {
      tma_prefetch_desc(&tmO0);
      if (NP == 2) tma_prefetch_desc(&tmO1);
    }
    if (p.ep.pool.p[0]) {
      tma_prefetch_desc(&tmP0);
      if (NP == 2) tma_prefetch_desc(&tmP1);
    }
    if (p.ep.residual.p[0]) {
      tma_prefetch_desc(&tmR0);
      if (NP == 2) tma_prefetch_desc(&tmR1);
    }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EW);
    }
    for (int a = 0; a < 2 * EW; ++a) mbar_init(&res_bar[a], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.taps * p.kc_per_tap;
  // accumulation chunks (fp32 mode only; bf16 mode accumulates the whole K in one TMEM accumulator)
  const int num_chunks = NP == 2 ? (num_kb + p.chunk_kb - 1) / p.chunk_kb : 1;
  const int chunk_len = (num_kb + num_chunks - 1) / num_chunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (uniform control flow for the warp, one elected lane issues the copies: straight-line UTMALDG, like the MMA warp)
    {
      const bool leader = elect_one();
      uint32_t stage = 0, phase = 0;
      for (int item = it_first; item < it_count; item += it_step) {
        int n, y0, x0, nb;
        decode_item<TW, TH>(p, item, n, y0, x0, nb);
        const int brow = nb * BN;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
          const int dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < p.kc_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + NP * Cfg::A_BYTES;
            if (leader) {
              mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              tma_load_4d(sa, &tmA0, &full_bar[stage], kc * kKChunk, x0 + dx, y0 + dy, n);
              if (NP == 2) tma_load_4d(sa + Cfg::A_BYTES, &tmA1, &full_bar[stage], kc * kKChunk, x0 + dx, y0 + dy, n);
              tma_load_2d(sb, &tmB0, &full_bar[stage], tap * p.Cin + kc * kKChunk, brow);
              if (NP == 2) tma_load_2d(sb + Cfg::B_BYTES, &tmB1, &full_bar[stage], tap * p.Cin + kc * kKChunk, brow);
            }
            __syncwarp();
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Uniform control flow for the whole warp (its index comes out of a shuffle, so the compiler knows), ONE elected lane
    // issues: the tcgen05 instructions are straight-line UTCHMMA / UTCBAR with descriptors advanced by one 64-bit add each.
    // Under `if (lane == 0)` each was wrapped in an ELECT / BRA.U.ANY loop behind a chain of uniform-datapath descriptor
    // arithmetic (~100+ cycles per MMA: more than the 32-64 cycles of math of an N <= 128 MMA).
    {
      const bool leader = elect_one();
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      const uint64_t desc_hi = make_desc_sw128(0, 16, 1024);   // LBO 16 B, SBO 1024 B, 128-byte swizzle; address added below
      auto desc_of = [&](uint32_t addr) { return desc_hi | uint64_t((addr & 0x3FFFFu) >> 4); };
      for (int item = it_first; item < it_count; item += it_step) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += chunk_len) {
          const int kb1 = kb0 + chunk_len < num_kb ? kb0 + chunk_len : num_kb;
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_main = tmem_base + acc * Cfg::ACC_COLS;
          const uint32_t d_cross = d_main + BN;  // fp32 mode only
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t b_hi = a_hi + NP * Cfg::A_BYTES;
            // K advances by 32 B per MMA = +2 in the descriptor's 16-byte address field
            const uint64_t da_hi = desc_of(a_hi), db_hi = desc_of(b_hi);
            const uint64_t da_lo = desc_of(a_hi + Cfg::A_BYTES), db_lo = desc_of(b_hi + Cfg::B_BYTES);
            if (leader) {
#pragma unroll
              for (int k = 0; k < kKChunk / 16; ++k) {
                const uint32_t accum = ((kb - kb0) | k) != 0 ? 1u : 0u;  // first MMA of a chunk overwrites
                if (NP == 2) {
                  // The weight planes lie back to back in the stage (hi rows, then lo rows) = ONE K-major tile of 2*BN rows:
                  // a_hi x [w_hi | w_lo] is a single MMA of width 2*BN that fills main (columns [0,BN)) and cross
                  // ([BN,2BN)) at once; a_lo x w_hi then accumulates into cross.  Same tensor cycles as three MMAs of
                  // width BN, but a_hi is fetched from shared memory once instead of twice.
                  if (p.x8) {
                    // 8-bit cross planes: one e4m3 MMA of K = 32 (16 channels x two halves) yields both cross terms at
                    // twice the fp16 rate -> two MMA slots per k-step instead of three
                    umma_bf16(d_main, da_hi + 2 * k, db_hi + 2 * k, p.idesc_hi, accum);
                    umma_f8(d_cross, da_lo + 2 * k, db_lo + 2 * k, p.idesc_hi, accum);
                  } else {
                    umma_bf16(d_main, da_hi + 2 * k, db_hi + 2 * k, p.idesc_wide, accum);
                    umma_bf16(d_cross, da_lo + 2 * k, db_hi + 2 * k, p.idesc_hi, 1u);
                  }
                } else {
                  umma_bf16(d_main, da_hi + 2 * k, db_hi + 2 * k, p.idesc_hi, accum);
                }
              }
              umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
              if (kb + 1 == kb1) umma_commit(&tfull_bar[acc]);  // chunk accumulator complete -> epilogue
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
    // ===================== epilogue (warps 2..9) =====================
    // warp -> TMEM lane quarter (warp & 3, a hardware restriction) and column slice ((warp - 2) >> 2: half or quarter)
    constexpr int HB = BN / (EW / 4);      // columns per epilogue warp
    static_assert(HB % 32 == 0, "an epilogue warp works on 32-column groups");
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;  // pixel index inside the patch
    const int ly = row / TW, lx = row % TW;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16) + half * HB;
    uint32_t acc = 0, acc_phase = 0;
    EpiTile et;
    et.stg = smem_u32(staging) + (warp - 2) * 4096;
    et.out[0] = &tmO0; et.out[1] = &tmO1;
    et.pool[0] = &tmP0; et.pool[1] = &tmP1;
    et.res[0] = &tmR0; et.res[1] = &tmR1;
    et.rbar = res_bar + 2 * (warp - 2);
    et.rphase[0] = et.rphase[1] = 0;
    constexpr int NG = HB / 32;   // 32-channel groups per epilogue warp and tile
    // fused BatchNorm statistics: lane l owns channels (column block base + 32*k + l), k < HB/32
    constexpr int NCH = HB / 32;
    double st1[NCH], st2[NCH];
#pragma unroll
    for (int k = 0; k < NCH; ++k) st1[k] = st2[k] = 0.0;
    int st_nb = -1;
    auto flush_stats = [&](int nb_old) {
      if (p.ep.stats == nullptr || nb_old < 0) return;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const int c = nb_old * BN + half * HB + 32 * k + lane;
        if (st1[k] != 0.0) acc_add(p.ep.stats + c, st1[k]);
        if (st2[k] != 0.0) acc_add(p.ep.stats + p.Cout + c, st2[k]);
        st1[k] = st2[k] = 0.0;
      }
    };
    for (int item = it_first; item < it_count; item += it_step) {
      int n, y0, x0, nb;
      decode_item<TW, TH>(p, item, n, y0, x0, nb);
      if (nb != st_nb) {
        flush_stats(st_nb);
        st_nb = nb;
      }
      const int y = y0 + ly, x = x0 + lx;
      const bool valid = (y < p.H) && (x < p.W) && (n < p.N);
      const size_t pix = (size_t(n) * p.H + y) * p.W + x;
      const int cbase = nb * BN + half * HB;
      et.x0 = x0; et.y0 = y0 + (32 / TW) * q; et.n = n;   // a warp's 32 pixels = 32/TW tile rows
      if (p.ep.residual.p[0]) {
        // request the skip values of the first group(s) now: they land in the staging slots while the MMAs run
        if (lane == 0) {
          bulk_wait_read0();
          epi_issue_residual<NP>(et, 0, cbase);
          if (NP == 1 && NG > 1) epi_issue_residual<NP>(et, 1, cbase + 32);
        }
        if (valid && NG > (NP == 2 ? 1 : 2)) {   // later groups: at least pull their lines into L2
#pragma unroll
          for (int pl = 0; pl < NP; ++pl)
#pragma unroll
            for (int off = 0; off < HB * 2; off += 128)
              prefetch_l2(reinterpret_cast<const uint8_t*>(p.ep.residual.p[pl]) + (pix * p.Cout + cbase) * 2 + off);
        }
      }

      if (NP == 1) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HB; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(lane_addr + acc * Cfg::ACC_COLS + c0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_cols<NP, TW>(v, p, cbase + c0, valid, pix, et, c0 / 32, NG, st1[c0 / 32], st2[c0 / 32]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        // fp32 mode: sum the chunk accumulators (main + cross) in registers with round-to-nearest adds
        float sum[HB];
#pragma unroll
        for (int j = 0; j < HB; ++j) sum[j] = 0.f;
        for (int ch = 0; ch < num_chunks; ++ch) {
          mbar_wait(&tfull_bar[acc], acc_phase);
          tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < HB; c0 += 32) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32(lane_addr + acc * Cfg::ACC_COLS + c0, r0);
            tmem_ld_32x32(lane_addr + acc * Cfg::ACC_COLS + BN + c0, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j)
              sum[c0 + j] += fmaf(__uint_as_float(r1[j]), p.cross_scale, __uint_as_float(r0[j]));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
#pragma unroll
        for (int c0 = 0; c0 < HB; c0 += 32) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = sum[c0 + j];
          epilogue_cols<NP, TW>(v, p, cbase + c0, valid, pix, et, c0 / 32, NG, st1[c0 / 32], st2[c0 / 32]);
        }
      }
    }
    flush_stats(st_nb);
    if (lane == 0) bulk_wait0();   // the staging tiles must outlive the TMA stores that read them
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// 256-wide tiles for the 8-bit-cross format (column blocks of 256 output channels).
//
// The 128-wide hi+lo kernel above is bound by shared-memory traffic: a 128-wide MMA needs 8 KB of operands for 64 cycles
// of math, and issuing the same MMAs 256 wide costs only 0.62x the time per MAC (profiles/README.md).  A 256-wide tile
// needs main + cross accumulators of 256 columns = all 512 TMEM columns, so the accumulators are NOT double-buffered here:
// the issuer waits while the epilogue warps drain a chunk (~10 % of a chunk's MMA time).  K blocks are 32 channels (64-byte
// rows, 64-byte swizzle) so that four stages of A (2 x 8 KB) + B (2 x 16 KB) fit next to the epilogue staging tiles.
// ------------------------------------------------------------------------------------------------
// (BN_ = 128 for layers with 128 output channels: same 32-channel ring, six finer stages and the weight-tile multicast;
//  its MMAs stay 128 wide.)
constexpr int kWideBK = 32;
template <int BN_, int CG = 1>
struct WideCfg {
  static constexpr int BN = BN_, BK = kWideBK;
  static constexpr int A_BYTES = 128 * 64;     // 128 pixels x 32 channels x 2 B (or 64 e4m3 bytes)
  static constexpr int B_BYTES = (BN / CG) * 64;   // CTA pairs: every CTA holds half the rows of the weight tile
  static constexpr int STAGE_BYTES = 2 * (A_BYTES + B_BYTES);
  static constexpr int STAGING_BYTES = 8 * 4096;
  static constexpr int STAGES = (227 * 1024 - 1024 - 512 - STAGING_BYTES) / STAGE_BYTES;   // 4 (BN 256) / 6 (BN 128)
  static constexpr int TMEM_COLS = 2 * BN;     // main + cross, single-buffered
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 + 512;
  static_assert(SMEM_BYTES <= 227 * 1024 && STAGES >= 3 && STAGES <= 8, "shared memory budget");
};

// MC = 2: two CTAs (a cluster) work on two pixel tiles of the same column block; each fetches HALF of the weight tile and
// multicasts it into both -- a CTA then has 32 KB instead of 48 KB of TMA requests in flight per k-block.
// CG = 2 (with MC = 1): the two CTAs are a cta_group::2 PAIR executing one MMA of M = 256 (two pixel tiles) x N = BN.  Each
// CTA keeps only ITS half of the weight rows in shared memory (no multicast, no second copy): per k-block a CTA writes
// 32 KB and its tensor core reads 32 KB instead of 48 + 48 KB -- the single-CTA kernel sits on the shared-memory bandwidth
// (96 B/clk written by TMA + 96 B/clk read by the MMAs at full tensor rate), the pair needs 64 + 64 -- and the 32 KB stages
// make the ring six deep.  The leader (rank 0) issues all MMAs; both CTAs' TMA boxes complete on the leader's barrier.
template <int BN_, int MC, int CG = 1>
__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_wide_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                      const __grid_constant__ CUtensorMap tmP0, const __grid_constant__ CUtensorMap tmP1,
                      const __grid_constant__ CUtensorMap tmR0, const __grid_constant__ CUtensorMap tmR1,
                      const __grid_constant__ ConvKernelParams p) {
  using Cfg = WideCfg<BN_, CG>;
  static_assert(!(MC > 1 && CG > 1), "multicast clusters and CTA pairs are alternatives");
  constexpr int BN = Cfg::BN, NP = 2, TW = kTileW, TH = kTileH;
  constexpr bool PAIR = CG == 2;
  constexpr int CL = (MC > 1 || PAIR) ? 2 : 1;   // CTAs per cluster
  const int cta_rank = CL > 1 ? int(cluster_ctarank()) : 0;
  const int it_first = int(blockIdx.x) / CL, it_step = int(gridDim.x) / CL;
  const int it_count = CL > 1 ? p.total_pairs : p.total_items;
  constexpr unsigned short kAllCtas = (unsigned short)((1u << MC) - 1);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::STAGES;
  uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint64_t* res_bar = tempty_bar + 1;   // [8 epilogue warps][2 slots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 16);

  const int warp = __shfl_sync(0xffffffffu, int(threadIdx.x >> 5), 0);   // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA0); tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB0); tma_prefetch_desc(&tmB1);
    if (p.ep.out.p[0]) { tma_prefetch_desc(&tmO0); tma_prefetch_desc(&tmO1); }
    if (p.ep.pool.p[0]) { tma_prefetch_desc(&tmP0); tma_prefetch_desc(&tmP1); }
    if (p.ep.residual.p[0]) { tma_prefetch_desc(&tmR0); tma_prefetch_desc(&tmR1); }
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], MC);   // free once every CTA that receives multicast slices in it has consumed it
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 8 * CG);   // pairs: the leader's copy collects the epilogue warps of both CTAs
    for (int a = 0; a < 16; ++a) mbar_init(&res_bar[a], 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
    else tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kc_per_tap = p.Cin / Cfg::BK;
  const int num_kb = p.taps * kc_per_tap;
  // accumulation chunks of the same K extent as in the 128-wide kernel (chunk_kb counts 64-channel blocks)
  const int chunk_kb = 2 * p.chunk_kb;
  const int num_chunks = (num_kb + chunk_kb - 1) / chunk_kb;
  const int chunk_len = (num_kb + num_chunks - 1) / num_chunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (uniform control flow for the warp, one elected lane issues the copies)
    {
      const bool leader = elect_one();
      const uint64_t w_policy = l2_policy_evict_last();
      uint32_t stage = 0, phase = 0;
      for (int item = it_first; item < it_count; item += it_step) {
        int n, y0, x0, nb;
        if (CL > 1) decode_pair(p, item, cta_rank, n, y0, x0, nb);
        else decode_item<TW, TH>(p, item, n, y0, x0, nb);
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0;
          const int dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kc = 0; kc < kc_per_tap; ++kc) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + 2 * Cfg::A_BYTES;
            if (leader) {
              if (PAIR) {
                // both CTAs' boxes complete on the leader's barrier, which expects the bytes of the pair
                if (cta_rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                const int brow = nb * BN + cta_rank * (BN / 2);   // this CTA's rows of the weight tile
                tma_load_4d_pair(sa, &tmA0, &full_bar[stage], kc * Cfg::BK, x0 + dx, y0 + dy, n);
                tma_load_4d_pair(sa + Cfg::A_BYTES, &tmA1, &full_bar[stage], kc * Cfg::BK, x0 + dx, y0 + dy, n);
                if (p.w_evict_last) {   // the weight matrix stays in L2 (every work item streams its column block again)
                  tma_load_2d_pair_hint(sb, &tmB0, &full_bar[stage], tap * p.Cin + kc * Cfg::BK, brow, w_policy);
                  tma_load_2d_pair_hint(sb + Cfg::B_BYTES, &tmB1, &full_bar[stage], tap * p.Cin + kc * Cfg::BK, brow, w_policy);
                } else {
                  tma_load_2d_pair(sb, &tmB0, &full_bar[stage], tap * p.Cin + kc * Cfg::BK, brow);
                  tma_load_2d_pair(sb + Cfg::B_BYTES, &tmB1, &full_bar[stage], tap * p.Cin + kc * Cfg::BK, brow);
                }
              } else {
                mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                tma_load_4d(sa, &tmA0, &full_bar[stage], kc * Cfg::BK, x0 + dx, y0 + dy, n);
                tma_load_4d(sa + Cfg::A_BYTES, &tmA1, &full_bar[stage], kc * Cfg::BK, x0 + dx, y0 + dy, n);
                if (MC > 1) {   // this CTA's slice of the weight rows, into every CTA of the cluster
                  const int rows = BN / MC;
                  uint8_t* sl = sb + cta_rank * rows * 64;
                  tma_load_2d_mc(sl, &tmB0, &full_bar[stage], tap * p.Cin + kc * Cfg::BK, nb * BN + cta_rank * rows, kAllCtas);
                  tma_load_2d_mc(sl + Cfg::B_BYTES, &tmB1, &full_bar[stage], tap * p.Cin + kc * Cfg::BK,
                                 nb * BN + cta_rank * rows, kAllCtas);
                } else {
                  tma_load_2d(sb, &tmB0, &full_bar[stage], tap * p.Cin + kc * Cfg::BK, nb * BN);
                  tma_load_2d(sb + Cfg::B_BYTES, &tmB1, &full_bar[stage], tap * p.Cin + kc * Cfg::BK, nb * BN);
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
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // uniform control flow, one elected lane issues (see conv_gemm_kernel): straight-line UTC*MMA, descriptors by 64-bit adds
    if (!PAIR || cta_rank == 0) {   // CTA pairs: the leader issues for both
      const bool leader = elect_one();
      uint32_t stage = 0, phase = 0, acc_phase = 0;
      const uint32_t d_main = tmem_base, d_cross = tmem_base + BN;
      const uint64_t desc_hi = make_desc_sw64(0);
      auto desc_of = [&](uint32_t addr) { return desc_hi | uint64_t((addr & 0x3FFFFu) >> 4); };
      for (int item = it_first; item < it_count; item += it_step) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += chunk_len) {
          const int kb1 = kb0 + chunk_len < num_kb ? kb0 + chunk_len : num_kb;
          mbar_wait(tempty_bar, acc_phase ^ 1);   // the epilogue warps have drained the previous chunk
          acc_phase ^= 1;
          tc_fence_after();
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint32_t b_hi = a_hi + 2 * Cfg::A_BYTES;
            const uint64_t da = desc_of(a_hi), dbw = desc_of(b_hi);
            const uint64_t da2 = desc_of(a_hi + Cfg::A_BYTES), dbw2 = desc_of(b_hi + Cfg::B_BYTES);
            if (leader) {
#pragma unroll
              for (int k = 0; k < Cfg::BK / 16; ++k) {
                // main: a fresh accumulation per chunk (truncating tensor-core adds).  cross: one
                // accumulation over the whole K -- its terms are 2^-12 of the result, 576 truncations cost nothing -- so the
                // per-chunk drain touches only the main half of TMEM
                const uint32_t accum = ((kb - kb0) | k) != 0 ? 1u : 0u;
                const uint32_t accum_x = (kb | k) != 0 ? 1u : 0u;
                if (PAIR) {
                  umma_bf16_pair(d_main, da + 2 * k, dbw + 2 * k, p.idesc_hi, accum);
                  umma_f8_pair(d_cross, da2 + 2 * k, dbw2 + 2 * k, p.idesc_hi, accum_x);
                } else {
                  umma_bf16(d_main, da + 2 * k, dbw + 2 * k, p.idesc_hi, accum);
                  umma_f8(d_cross, da2 + 2 * k, dbw2 + 2 * k, p.idesc_hi, accum_x);
                }
              }
              if (PAIR) umma_commit_pair(&empty_bar[stage]);
              else if (MC > 1) umma_commit_mc(&empty_bar[stage], kAllCtas);
              else umma_commit(&empty_bar[stage]);
              if (kb + 1 == kb1) {
                if (PAIR) umma_commit_pair(tfull_bar);
                else umma_commit(tfull_bar);
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
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    constexpr int HB = BN / 2;      // 128 columns per epilogue warp
    constexpr int NG = HB / 32;
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ly = row / TW, lx = row % TW;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16) + half * HB;
    uint32_t acc_phase = 0;
    EpiTile et;
    et.stg = smem_u32(staging) + (warp - 2) * 4096;
    et.out[0] = &tmO0; et.out[1] = &tmO1;
    et.pool[0] = &tmP0; et.pool[1] = &tmP1;
    et.res[0] = &tmR0; et.res[1] = &tmR1;
    et.rbar = res_bar + 2 * (warp - 2);
    et.rphase[0] = et.rphase[1] = 0;
    double st1[NG], st2[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) st1[k] = st2[k] = 0.0;
    int st_nb = -1;
    auto flush_stats = [&](int nb_old) {
      if (p.ep.stats == nullptr || nb_old < 0) return;
#pragma unroll
      for (int k = 0; k < NG; ++k) {
        const int c = nb_old * BN + half * HB + 32 * k + lane;
        if (st1[k] != 0.0) acc_add(p.ep.stats + c, st1[k]);
        if (st2[k] != 0.0) acc_add(p.ep.stats + p.Cout + c, st2[k]);
        st1[k] = st2[k] = 0.0;
      }
    };
    for (int item = it_first; item < it_count; item += it_step) {
      int n, y0, x0, nb;
      if (CL > 1) decode_pair(p, item, cta_rank, n, y0, x0, nb);
      else decode_item<TW, TH>(p, item, n, y0, x0, nb);
      if (nb != st_nb) {
        flush_stats(st_nb);
        st_nb = nb;
      }
      const int y = y0 + ly, x = x0 + lx;
      const bool valid = (y < p.H) && (x < p.W) && (n < p.N);
      const size_t pix = (size_t(n) * p.H + y) * p.W + x;
      const int cbase = nb * BN + half * HB;
      et.x0 = x0; et.y0 = y0 + 2 * q; et.n = n;
      if (p.ep.residual.p[0] && lane == 0) {
        bulk_wait_read0();
        epi_issue_residual<NP>(et, 0, cbase);
      }
      float sum[HB];
#pragma unroll
      for (int j = 0; j < HB; ++j) sum[j] = 0.f;
      for (int ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(tfull_bar, acc_phase);
        acc_phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < HB; c0 += 16) {
          uint32_t r0[16];
          tmem_ld_32x16(lane_addr + c0, r0);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) sum[c0 + j] += __uint_as_float(r0[j]);
        }
        if (ch == num_chunks - 1) {   // the cross accumulator holds the whole K by now
#pragma unroll
          for (int c0 = 0; c0 < HB; c0 += 16) {
            uint32_t r1[16];
            tmem_ld_32x16(lane_addr + BN + c0, r1);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) sum[c0 + j] = fmaf(__uint_as_float(r1[j]), p.cross_scale, sum[c0 + j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_leader(tempty_bar);
          else mbar_arrive(tempty_bar);
        }
      }
#pragma unroll
      for (int c0 = 0; c0 < HB; c0 += 32) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = sum[c0 + j];
        epilogue_cols<NP, TW>(v, p, cbase + c0, valid, pix, et, c0 / 32, NG, st1[c0 / 32], st2[c0 / 32]);
      }
    }
    flush_stats(st_nb);
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  if (CL > 1) cluster_sync_all();   // the peer may still multicast into this CTA / signal its barriers / read its smem
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host launcher
// ------------------------------------------------------------------------------------------------
template <int BN, int NP, int EW = 8>
static int launch_t(const CUtensorMap* maps, const ConvKernelParams& kp, int grid, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, NP, EW>;
  auto kern = conv_gemm_kernel<BN, NP, EW>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("cudaFuncSetAttribute(conv_gemm<%d,%d>, %d B smem): %s", BN, NP, Cfg::SMEM_BYTES, cudaGetErrorString(e));
      return 1;
    }
    attr_set = true;
  }
  kern<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7],
                                                        maps[8], maps[9], kp);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("conv_gemm<%d,%d> launch failed: %s", BN, NP, cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  return 0;
}

static int g_num_sms = 0;
static int g_force_bn = 0;
static int g_chunk_kb = kChunkKBDefault;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
    const char* f = getenv("NSM_FORCE_BN");
    if (f) g_force_bn = atoi(f);
    const char* c = getenv("NSM_CHUNK_KB");
    if (c && atoi(c) > 0) g_chunk_kb = atoi(c);
  }
  return g_num_sms;
}

int conv_gemm_pick_bn(const ConvShape& s) {
  num_sms();
  if (g_force_bn && s.Cout % g_force_bn == 0 && !(s.fmt != 0 && g_force_bn == 256)) return g_force_bn;
  if (s.fmt == 0 && s.Cout % 256 == 0) return 256;
  if (s.Cout % 128 == 0) return 128;
  return 64;
}

int conv_gemm_launch(const ConvShape& s, const Planes& in, const Planes& w, const ConvEpilogue& ep,
                     cudaStream_t stream) {
  if (s.Cin % kKChunk || s.Cout % 64 || (s.taps != 1 && s.taps != 9) || (s.fmt < 0 || s.fmt > 3) ||
      s.N <= 0 || s.H <= 0 || s.W <= 0) {
    set_error("conv_gemm: unsupported shape N=%d H=%d W=%d Cin=%d Cout=%d taps=%d fmt=%d", s.N, s.H, s.W,
              s.Cin, s.Cout, s.taps, s.fmt);
    return 1;
  }
  int BN = conv_gemm_pick_bn(s);
  // 8-bit-cross operands with column blocks of 256: conv_gemm_wide_kernel (NSM_NO_WIDE=1: the 128-wide kernel) ...
  static const bool wide_off = getenv("NSM_NO_WIDE") != nullptr;
  const bool wide = s.fmt == kFmtF16X8 && s.Cout % 256 == 0 && s.Cin % 32 == 0 && !wide_off;
  if (wide) BN = 256;
  const int wide_tiles = s.N * ((s.W + kTileW - 1) / kTileW) * ((s.H + kTileH - 1) / kTileH);
  // ... as cta_group::2 pairs that split the weight tile (NSM_NO_WIDE_PAIR=1: clusters of two CTAs that multicast halves of
  // it to each other; NSM_NO_WIDE_MC=1 on top of that: single CTAs)
  static const bool wide_pair_on = getenv("NSM_NO_WIDE_PAIR") == nullptr;
  static const bool wide_mc_on = getenv("NSM_NO_WIDE_MC") == nullptr;
  const bool wide_pair = wide && wide_pair_on && wide_tiles >= 8;
  const int wide_mc = (wide && !wide_pair && wide_mc_on && wide_tiles >= 8) ? 2 : 1;
  constexpr int TW = kTileW, TH = kTileH;
  const int m_tiles = s.N * ((s.W + TW - 1) / TW) * ((s.H + TH - 1) / TH);
  CUtensorMap maps[10];  // A hi/lo, B hi/lo, output hi/lo, pooled output hi/lo, skip hi/lo
  memset(maps, 0, sizeof(maps));
  const uint64_t adims[4] = {uint64_t(s.Cin), uint64_t(s.W), uint64_t(s.H), uint64_t(s.N)};
  const uint64_t astr[3] = {uint64_t(s.Cin) * 2, uint64_t(s.W) * s.Cin * 2, uint64_t(s.H) * s.W * s.Cin * 2};
  const uint32_t kblk = wide ? kWideBK : kKChunk;   // channels per k-block = 64- or 128-byte operand rows
  const int op_swz = wide ? 64 : 128;
  const uint32_t abox[4] = {kblk, uint32_t(TW), uint32_t(TH), 1};
  const uint64_t K = uint64_t(s.taps) * s.Cin;
  const uint64_t bdims[2] = {K, uint64_t(s.Cout)};
  const uint64_t bstr[1] = {K * 2};
  const uint32_t bbox[2] = {kblk, uint32_t(BN / (wide_mc * (wide_pair ? 2 : 1)))};
  const int planes = fmt_planes(s.fmt);
  for (int pl = 0; pl < planes; ++pl) {
    if (!in.p[pl] || !w.p[pl]) {
      set_error("conv_gemm: null operand plane %d", pl);
      return 1;
    }
    if (encode_tmap_tiled(&maps[pl], in.p[pl], 4, adims, astr, abox, 2, op_swz)) return 1;
    if (encode_tmap_tiled(&maps[2 + pl], w.p[pl], 2, bdims, bstr, bbox, 2, op_swz)) return 1;
  }
  if (planes == 1) {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  // output side: 64-byte swizzled boxes of 32 channels x (2 x 16) pixels = what one epilogue warp stages per plane
  const uint64_t odims[4] = {uint64_t(s.Cout), uint64_t(s.W), uint64_t(s.H), uint64_t(s.N)};
  const uint64_t ostr[3] = {uint64_t(s.Cout) * 2, uint64_t(s.W) * s.Cout * 2, uint64_t(s.H) * s.W * s.Cout * 2};
  const uint32_t obox[4] = {32, uint32_t(TW), uint32_t(32 / TW), 1};   // one epilogue warp = 32 pixels
  const uint64_t pdims[4] = {uint64_t(s.Cout), uint64_t(s.W / 2), uint64_t(s.H / 2), uint64_t(s.N)};
  const uint64_t pstr[3] = {uint64_t(s.Cout) * 2, uint64_t(s.W / 2) * s.Cout * 2,
                            uint64_t(s.H / 2) * (s.W / 2) * s.Cout * 2};
  const uint32_t pbox[4] = {32, uint32_t(TW / 2), uint32_t(16 / TW), 1};
  if (ep.pool.p[0] && (s.W < 2 || s.H < 2)) {
    set_error("conv_gemm: pooled output requested for a %dx%d image", s.H, s.W);
    return 1;
  }
  for (int pl = 0; pl < 2; ++pl) {
    maps[4 + pl] = maps[0];
    maps[6 + pl] = maps[0];
    maps[8 + pl] = maps[0];
    if (pl >= planes) continue;
    if (ep.residual.p[0]) {
      if (!ep.residual.p[pl] || !ep.out.p[0]) {
        set_error("conv_gemm: skip tensor needs plane %d and an output", pl);
        return 1;
      }
      if (encode_tmap_tiled(&maps[8 + pl], ep.residual.p[pl], 4, odims, ostr, obox, 2, 64)) return 1;
    }
    if (ep.out.p[0]) {
      if (!ep.out.p[pl]) {
        set_error("conv_gemm: null output plane %d", pl);
        return 1;
      }
      if (encode_tmap_tiled(&maps[4 + pl], ep.out.p[pl], 4, odims, ostr, obox, 2, 64)) return 1;
    }
    if (ep.pool.p[0]) {
      if (!ep.pool.p[pl]) {
        set_error("conv_gemm: null pooled output plane %d", pl);
        return 1;
      }
      if (encode_tmap_tiled(&maps[6 + pl], ep.pool.p[pl], 4, pdims, pstr, pbox, 2, 64)) return 1;
    }
  }
  ConvKernelParams kp;
  kp.N = s.N; kp.H = s.H; kp.W = s.W; kp.Cin = s.Cin; kp.Cout = s.Cout; kp.taps = s.taps;
  kp.tiles_x = (s.W + TW - 1) / TW;
  kp.tiles_y = (s.H + TH - 1) / TH;
  kp.n_blocks = s.Cout / BN;
  kp.total_items = s.N * kp.tiles_x * kp.tiles_y * kp.n_blocks;
  kp.total_pairs = ((m_tiles + 1) / 2) * kp.n_blocks;
  kp.kc_per_tap = s.Cin / kKChunk;
  // kFmtF16X8 describes the OPERANDS (input + weights); the output, skip and pooled tensors are plain fp16 hi+lo planes
  kp.fmt = s.fmt == kFmtF16X8 ? kFmtF16x2 : s.fmt;
  kp.chunk_kb = g_chunk_kb;
  kp.x8 = s.fmt == kFmtF16X8 ? 1 : 0;
  {
    static const bool off = getenv("NSM_NO_W_EVICT_LAST") != nullptr;
    kp.w_evict_last = off ? 0 : 1;
  }
  kp.cross_scale = kp.x8 ? kX8CrossScale : 1.f;
  const uint32_t ef = fmt_is_f16(s.fmt) ? kFmtF16 : kFmtBF16;  // fp16 / e4m3 share the descriptor code 0
  kp.idesc_hi = make_idesc_f16(wide_pair ? 256 : 128, BN, ef, ef, 0, 0);
  kp.idesc_wide = planes == 2 ? make_idesc_f16(128, 2 * BN, ef, ef, 0, 0) : kp.idesc_hi;
  kp.ep = ep;
  if (wide) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(kConvThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    if (wide_mc == 2 || wide_pair) {
      const int groups = kp.total_pairs < num_sms() / 2 ? kp.total_pairs : num_sms() / 2;
      cfg.gridDim = dim3(2 * groups);
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
    } else {
      cfg.gridDim = dim3(kp.total_items < num_sms() ? kp.total_items : num_sms());
    }
    auto launch = [&](auto kern, int smem_bytes, bool& attr_set) -> cudaError_t {
      if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
      }
      cfg.dynamicSmemBytes = smem_bytes;
      return cudaLaunchKernelEx(&cfg, kern, maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], maps[7], maps[8],
                                maps[9], kp);
    };
    static bool attr_set[3] = {false, false, false};
    cudaError_t e;
    if (wide_pair)
      e = launch(conv_gemm_wide_kernel<256, 1, 2>, WideCfg<256, 2>::SMEM_BYTES, attr_set[2]);
    else
      e = wide_mc == 2 ? launch(conv_gemm_wide_kernel<256, 2>, WideCfg<256>::SMEM_BYTES, attr_set[0])
                       : launch(conv_gemm_wide_kernel<256, 1>, WideCfg<256>::SMEM_BYTES, attr_set[1]);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) {
      set_error("conv_gemm_wide<%d,%d,%d> launch failed: %s", BN, wide_mc, wide_pair ? 2 : 1, cudaGetErrorString(e));
      return 1;
    }
    count_launch();
    return 0;
  }
  const int grid = kp.total_items < num_sms() ? kp.total_items : num_sms();
  // sixteen epilogue warps for the layers with a short K loop (<= kb16 k-blocks per output tile; NSM_EW16_KB overrides, 0 = off).
  // The hi+lo kernel at BN = 128 keeps only two 64 KB stages next to 64 KB of staging tiles: there only the shortest loops.
  static const int kb16_env = getenv("NSM_EW16_KB") ? atoi(getenv("NSM_EW16_KB")) : -1;
  const int num_kb = s.taps * (s.Cin / kKChunk);
  const int kb16 = kb16_env >= 0 ? kb16_env : (planes == 1 ? 18 : 4);
  const bool ew16 = BN >= 128 && num_kb <= kb16;
  if (planes == 1) {
    if (BN == 256) return ew16 ? launch_t<256, 1, 16>(maps, kp, grid, stream) : launch_t<256, 1>(maps, kp, grid, stream);
    if (BN == 128) return ew16 ? launch_t<128, 1, 16>(maps, kp, grid, stream) : launch_t<128, 1>(maps, kp, grid, stream);
    return launch_t<64, 1>(maps, kp, grid, stream);
  }
  if (BN == 128) return ew16 ? launch_t<128, 2, 16>(maps, kp, grid, stream) : launch_t<128, 2>(maps, kp, grid, stream);
  return launch_t<64, 2>(maps, kp, grid, stream);
}

}  // namespace nsm
