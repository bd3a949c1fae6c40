// C ABI of libnsm_b200.so (see include/nsm_b200.h) and the eval-mode whole-network orchestration.
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/nsm_b200.h"
#include "conv_gemm.cuh"
#include "nsm_common.cuh"
#include "stream_kernels.cuh"
#include "train_kernels.cuh"
#include "upblock.cuh"

namespace nsm {
const char* last_error();

// ------------------------------------------------------------------------------------------------
// network description (Unetmodel.py:36-63)
// ------------------------------------------------------------------------------------------------
struct BlockDef {
  int cin, cout;
};
static const BlockDef kBlocks[8] = {{16, 64},   {64, 128},  {128, 512}, {512, 1024},
                                    {1024, 512}, {512, 128}, {128, 64},  {64, 16}};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Packed parameter blob: for every block the two per-channel vector triples (bias, BN scale, BN shift) and the
// weights either as GEMM operand planes (tensor-core layers) or as fp32 OIHW (head: conv2, tail: conv9 1x1, conv10).
struct PackedLayout {
  size_t w3[8][2];  // 3x3 weight planes (block 0: fp32 OIHW copy in [0])
  size_t w1[8][2];  // 1x1 weight planes (blocks 0 and 7: fp32 copy in [0])
  size_t v3[8][3];  // bias, scale, shift of the 3x3 stage  [cin]
  size_t v1[8][3];  // bias, scale, shift of the 1x1 stage  [cout]
  size_t w10, b10;
  size_t w1g7[2];   // conv9's 1x1 weights as GEMM operand planes [16][64] (fused decoder block)
  size_t head_img;  // conv2's operand image for the head kernel (head_pack_image)
  size_t tail_img;  // conv9 1x1 + conv10 operand image for the tail kernel (tail_pack_image)
  size_t total;
};

// fp32 mode: the decoder's up-sampled tensors (u6..u9) and the weights of the 3x3 convolutions that consume them use the
// fp16 + 8-bit cross format (nsm_common.cuh: two MMA slots per k-step instead of three).  NSM_NO_X8=1 keeps fp16 hi+lo.
static int decoder3x3_fmt(int mode) {
  static const bool off = getenv("NSM_NO_X8") != nullptr;
  return (mode == NSM_MODE_FP32 && !off) ? kFmtF16X8 : mode;
}

// conv8 / conv9 (+ conv10, sigmoid, pixel_shuffle) of the eval forward as fused blocks (upblock.cu): no up-sample launch,
// no u8 / t8 / u9 / t9 in HBM.  NSM_NO_FUSED=1: the stage-by-stage path (every intermediate visible to nsm_unet_tap).
static std::atomic<int> g_fused_override{-1};   // nsm_unet_set_fused_decoder: -1 = environment default
static bool fused_decoder(int mode) {
  static const bool off = getenv("NSM_NO_FUSED") != nullptr;
  const int ov = g_fused_override.load(std::memory_order_relaxed);
  // (the fp32-mode blocks consume the 8-bit-cross form of the 3x3 weights: NSM_NO_X8 implies the stage-by-stage decoder)
  return (ov < 0 ? !off : ov != 0) &&
         (mode == NSM_MODE_BF16 || (mode == NSM_MODE_FP32 && decoder3x3_fmt(mode) == kFmtF16X8));
}

static PackedLayout packed_layout(int mode) {
  const int np = fmt_planes(mode);
  PackedLayout L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  for (int b = 0; b < 8; ++b) {
    const size_t cin = kBlocks[b].cin, cout = kBlocks[b].cout;
    const bool gemm3 = b != 0, gemm1 = (b != 0 && b != 7);
    for (int p = 0; p < 2; ++p) {
      if (gemm3) L.w3[b][p] = p < np ? take(cin * cin * 9 * 2) : 0;   // bf16 GEMM operand plane
      else L.w3[b][p] = p == 0 ? take(cin * cin * 9 * 4) : 0;         // fp32 OIHW for the SIMT head
      if (gemm1) L.w1[b][p] = p < np ? take(cin * cout * 2) : 0;
      else L.w1[b][p] = p == 0 ? take(cin * cout * 4) : 0;
    }
    for (int k = 0; k < 3; ++k) {
      L.v3[b][k] = take(cin * 4);
      L.v1[b][k] = take(cout * 4);
    }
  }
  L.w10 = take(4 * 16 * 4);
  L.b10 = take(4 * 4);
  for (int p = 0; p < 2; ++p) L.w1g7[p] = p < np ? take(16 * 64 * 2) : 0;
  L.head_img = take(kHeadImageBytes);
  L.tail_img = take(kTailImageBytes);
  L.total = off;
  return L;
}

// Workspace: one NHWC planes buffer per intermediate (names as in nsm_unet_tap).
struct Level {
  int h, w;
};
struct TapDef {
  const char* name;
  int level;  // 1..4 : spatial level (h1 = H/2 ... h4)
  int C;
};
static const TapDef kTaps[] = {
    {"c2", 1, 64},   {"p2", 2, 64},   {"t3", 2, 64},   {"c3", 2, 128},  {"p3", 3, 128}, {"t4", 3, 128},
    {"c4", 3, 512},  {"p4", 4, 512},  {"t5", 4, 512},  {"c5", 4, 1024}, {"u6", 3, 1024}, {"t6", 3, 1024},
    {"m6", 3, 512},  {"u7", 2, 512},  {"t7", 2, 512},  {"m7", 2, 128},  {"u8", 1, 128}, {"t8", 1, 128},
    {"m8", 1, 64},   {"u9", 1, 64},   {"t9", 1, 64},
};
constexpr int kNumTaps = sizeof(kTaps) / sizeof(kTaps[0]);

struct WorkspaceLayout {
  Level lv[5];
  size_t off[kNumTaps][2];
  size_t x_stage;  // device staging of the input for the *_host entry point
  size_t y_stage;
  size_t total;
};

static WorkspaceLayout workspace_layout(int B, int H, int W, int mode) {
  const int np = fmt_planes(mode);
  WorkspaceLayout L;
  const int He = H - (H & 1), We = W - (W & 1);
  L.lv[0] = {He, We};
  L.lv[1] = {He / 2, We / 2};
  for (int l = 2; l <= 4; ++l) L.lv[l] = {L.lv[l - 1].h / 2, L.lv[l - 1].w / 2};
  size_t off = 0;
  for (int t = 0; t < kNumTaps; ++t) {
    const Level& v = L.lv[kTaps[t].level];
    const size_t bytes = size_t(B) * v.h * v.w * kTaps[t].C * 2;
    for (int p = 0; p < 2; ++p) {
      L.off[t][p] = off;
      if (p < np) off = align_up(off + bytes, 1024);
    }
  }
  L.x_stage = off;
  off = align_up(off + size_t(B) * 4 * H * W * 4, 1024);
  L.y_stage = off;
  off = align_up(off + size_t(B) * He * We * 4, 1024);
  L.total = off;
  return L;
}

static int tap_index(const char* name) {
  for (int t = 0; t < kNumTaps; ++t)
    if (!strcmp(kTaps[t].name, name)) return t;
  return -1;
}

static Planes ws_planes(const WorkspaceLayout& L, void* ws, const char* name, int np) {
  const int t = tap_index(name);
  Planes p;
  p.p[0] = reinterpret_cast<uint8_t*>(ws) + L.off[t][0];
  p.p[1] = np == 2 ? reinterpret_cast<uint8_t*>(ws) + L.off[t][1] : nullptr;
  return p;
}

// ------------------------------------------------------------------------------------------------
// optional per-launch timing: CUDA events recorded on the launch stream around every kernel of a forward
// ------------------------------------------------------------------------------------------------
struct ProfRecord {
  std::string name;
  double flops, bytes;
  cudaEvent_t start, stop;
};
static std::mutex g_prof_mutex;
static bool g_prof_on = false;
static std::vector<ProfRecord> g_prof;

struct ProfScope {
  cudaStream_t st;
  int idx = -1;
  ProfScope(const char* name, double flops, double bytes, cudaStream_t s) : st(s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    ProfRecord r;
    r.name = name; r.flops = flops; r.bytes = bytes;
    cudaEventCreate(&r.start);
    cudaEventCreate(&r.stop);
    cudaEventRecord(r.start, st);
    g_prof.push_back(r);
    idx = int(g_prof.size()) - 1;
  }
  ~ProfScope() {
    if (idx < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    cudaEventRecord(g_prof[idx].stop, st);
  }
};

}  // namespace nsm

using namespace nsm;

static inline Planes mk(const void* a, const void* b) { return Planes{{const_cast<void*>(a), const_cast<void*>(b)}}; }
static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
// nsm_acc (C ABI) and nsm::Acc (nsm_common.cuh) are the same four 64-bit words
static_assert(sizeof(nsm_acc) == sizeof(Acc), "nsm_acc layout");
static inline Acc* A(nsm_acc* a) { return reinterpret_cast<Acc*>(a); }
static inline const Acc* A(const nsm_acc* a) { return reinterpret_cast<const Acc*>(a); }

#define NSM_TRY(expr)          \
  do {                         \
    if ((expr) != 0) return 1; \
  } while (0)

extern "C" {

const char* nsm_last_error(void) { return nsm::last_error(); }
int nsm_version(void) { return 100; }
long long nsm_launch_count(void) { return launch_count(); }
void nsm_tmap_cache_stats(long long* hits, long long* misses) { tmap_cache_stats(hits, misses); }

int nsm_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("cudaGetDevice: %s", cudaGetErrorString(e));
    return 1;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("device %d is sm_%d%d; libnsm_b200 is built for sm_100a only (no fallback path)", dev, major, minor);
    return 2;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- pack
size_t nsm_unet_packed_bytes(int mode) { return packed_layout(mode).total; }

int nsm_unet_pack(const float* const* T, int mode, void* blob, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode < 0 || mode > 2) {
    set_error("nsm_unet_pack: bad mode %d", mode);
    return 1;
  }
  const PackedLayout L = packed_layout(mode);
  const int rb = mode == NSM_MODE_BF16;
  uint8_t* base = reinterpret_cast<uint8_t*>(blob);
  for (int b = 0; b < 8; ++b) {
    const float* const* t = T + 12 * b;
    const int cin = kBlocks[b].cin, cout = kBlocks[b].cout;
    // 3x3 stage
    if (b == 0) {
      NSM_TRY(copy_round(t[0], reinterpret_cast<float*>(base + L.w3[b][0]), cin * cin * 9, rb, st));
    } else {
      NSM_TRY(pack_conv_weight(t[0], cin, cin, 3, 0, b >= 4 ? decoder3x3_fmt(mode) : mode, base + L.w3[b][0],
                               rb ? nullptr : base + L.w3[b][1], st));
    }
    NSM_TRY(copy_round(t[1], reinterpret_cast<float*>(base + L.v3[b][0]), cin, rb, st));
    NSM_TRY(bn_fold_eval(t[2], t[3], t[4], t[5], cin, 1e-5f, reinterpret_cast<float*>(base + L.v3[b][1]),
                         reinterpret_cast<float*>(base + L.v3[b][2]), st));
    // 1x1 stage
    if (b == 0 || b == 7) {
      NSM_TRY(copy_round(t[6], reinterpret_cast<float*>(base + L.w1[b][0]), cin * cout, rb, st));
    } else {
      NSM_TRY(pack_conv_weight(t[6], cout, cin, 1, 0, mode, base + L.w1[b][0], rb ? nullptr : base + L.w1[b][1], st));
    }
    NSM_TRY(copy_round(t[7], reinterpret_cast<float*>(base + L.v1[b][0]), cout, rb, st));
    NSM_TRY(bn_fold_eval(t[8], t[9], t[10], t[11], cout, 1e-5f, reinterpret_cast<float*>(base + L.v1[b][1]),
                         reinterpret_cast<float*>(base + L.v1[b][2]), st));
  }
  NSM_TRY(copy_round(T[96], reinterpret_cast<float*>(base + L.w10), 64, rb, st));
  NSM_TRY(copy_round(T[97], reinterpret_cast<float*>(base + L.b10), 4, rb, st));
  NSM_TRY(pack_conv_weight(T[12 * 7 + 6], 16, 64, 1, 0, mode, base + L.w1g7[0], rb ? nullptr : base + L.w1g7[1], st));
  auto fv = [&](size_t off) { return reinterpret_cast<const float*>(base + off); };
  NSM_TRY(head_pack_image(fv(L.w3[0][0]), fv(L.v3[0][0]), fv(L.v3[0][1]), fv(L.v3[0][2]), fv(L.w1[0][0]), fv(L.v1[0][0]),
                          fv(L.v1[0][1]), fv(L.v1[0][2]), mode, base + L.head_img, st));
  NSM_TRY(tail_pack_image(fv(L.w1[7][0]), fv(L.v1[7][0]), fv(L.v1[7][1]), fv(L.v1[7][2]), fv(L.w10), fv(L.b10), mode,
                          base + L.tail_img, st));
  return 0;
}

// ---------------------------------------------------------------------------------------------- infer
size_t nsm_unet_workspace_bytes(int B, int H, int W, int mode) {
  if (B < 1 || H < 16 || W < 16) return 0;
  return workspace_layout(B, H, W, mode).total;
}

static int infer_impl(const void* blob, int mode, const float* x, int B, int H, int W, const float* mean,
                      const float* std, float* y, uint8_t* y_u8, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode < 0 || mode > 2) {
    set_error("nsm_unet_infer: bad mode %d", mode);
    return 1;
  }
  if (B < 1 || H < 16 || W < 16) {
    set_error("nsm_unet_infer: input %dx%dx%d too small (need H,W >= 16)", B, H, W);
    return 1;
  }
  const WorkspaceLayout WL = workspace_layout(B, H, W, mode);
  if (ws_bytes < WL.total) {
    set_error("nsm_unet_infer: workspace %zu B < required %zu B", ws_bytes, WL.total);
    return 1;
  }
  const PackedLayout PL = packed_layout(mode);
  const int np = fmt_planes(mode);
  const uint8_t* pb = reinterpret_cast<const uint8_t*>(blob);
  auto fvec = [&](size_t off) { return reinterpret_cast<const float*>(pb + off); };
  auto wplanes = [&](const size_t (&o)[2]) {
    Planes p;
    p.p[0] = const_cast<uint8_t*>(pb + o[0]);
    p.p[1] = np == 2 ? const_cast<uint8_t*>(pb + o[1]) : nullptr;
    return p;
  };
  auto buf = [&](const char* name) { return ws_planes(WL, ws, name, np); };
  const Planes none = {{nullptr, nullptr}};

  // ---- head: conv2 block + pool
  {
    HeadParams hp;
    hp.x = x; hp.N = B; hp.Hin = H; hp.Win = W; hp.mean = mean; hp.std = std;
    hp.img = pb + PL.head_img;
    hp.fmt = mode; hp.c2 = buf("c2"); hp.p2 = buf("p2"); hp.x16 = none;
    const double px = double(B) * WL.lv[1].h * WL.lv[1].w;
    ProfScope ps("head(conv2)", px * 2.0 * (16 * 144 + 16 * 64),
                 double(B) * 4 * H * W * 4 + px * 64 * 2 * np * 1.25, st);
    NSM_TRY(head_eval(hp, st));
  }
  // ---- one DoubleConv on the tensor cores
  auto double_conv = [&](int b, int level, const Planes& in, const char* tname, const char* oname,
                         const char* resname, const char* poolname) -> int {
    const Level& lv = WL.lv[level];
    ConvShape s3 = {B, lv.h, lv.w, kBlocks[b].cin, kBlocks[b].cin, 9, b >= 4 ? decoder3x3_fmt(mode) : mode};
    ConvEpilogue e3;
    e3.bias = fvec(PL.v3[b][0]); e3.scale = fvec(PL.v3[b][1]); e3.shift = fvec(PL.v3[b][2]);
    e3.lrelu = 1; e3.round_bf16 = np == 1; e3.out = buf(tname); e3.residual = none; e3.pool = none;
    e3.out_f32 = nullptr;
    e3.stats = nullptr;
    const double px = double(B) * lv.h * lv.w, ci = kBlocks[b].cin, co = kBlocks[b].cout;
    char nm[32];
    {
      snprintf(nm, sizeof(nm), "conv%d.3x3", b + 2);
      ProfScope ps(nm, 2.0 * px * 9 * ci * ci, (px * ci * 2 + ci * ci * 9) * 2.0 * np, st);
      NSM_TRY(conv_gemm_launch(s3, in, wplanes(PL.w3[b]), e3, st));
    }
    if (!oname) return 0;
    ConvShape s1 = {B, lv.h, lv.w, kBlocks[b].cin, kBlocks[b].cout, 1, mode};
    ConvEpilogue e1 = e3;
    e1.bias = fvec(PL.v1[b][0]); e1.scale = fvec(PL.v1[b][1]); e1.shift = fvec(PL.v1[b][2]);
    e1.out = buf(oname);
    e1.residual = resname ? buf(resname) : none;
    e1.pool = poolname ? buf(poolname) : none;
    snprintf(nm, sizeof(nm), "conv%d.1x1", b + 2);
    ProfScope ps(nm, 2.0 * px * ci * co,
                 (px * (ci + co * (1.0 + (resname ? 1.0 : 0.0) + (poolname ? 0.25 : 0.0))) + ci * co) * 2.0 * np, st);
    NSM_TRY(conv_gemm_launch(s1, buf(tname), wplanes(PL.w1[b]), e1, st));
    return 0;
  };
  auto up = [&](const char* src, int slevel, int C, const char* dst, int dlevel) -> int {
    char unm[32];
    snprintf(unm, sizeof(unm), "upsample %s", dst);
    ProfScope ps(unm, 0.0,
                 (double(B) * WL.lv[slevel].h * WL.lv[slevel].w + double(B) * WL.lv[dlevel].h * WL.lv[dlevel].w) *
                     C * 2.0 * np, st);
    return upsample_match(buf(src), B, WL.lv[slevel].h, WL.lv[slevel].w, C, buf(dst), WL.lv[dlevel].h,
                          WL.lv[dlevel].w, decoder3x3_fmt(mode), st);
  };
  NSM_TRY(double_conv(1, 2, buf("p2"), "t3", "c3", nullptr, "p3"));   // conv3 + pool3
  NSM_TRY(double_conv(2, 3, buf("p3"), "t4", "c4", nullptr, "p4"));   // conv4 + pool4
  NSM_TRY(double_conv(3, 4, buf("p4"), "t5", "c5", nullptr, nullptr));  // conv5
  NSM_TRY(up("c5", 4, 1024, "u6", 3));
  NSM_TRY(double_conv(4, 3, buf("u6"), "t6", "m6", "c4", nullptr));   // conv6 + skip
  NSM_TRY(up("m6", 3, 512, "u7", 2));
  NSM_TRY(double_conv(5, 2, buf("u7"), "t7", "m7", "c3", nullptr));   // conv7 + skip
  if (fused_decoder(mode) && upblock_supported(WL.lv[2].h, WL.lv[2].w, WL.lv[1].h, WL.lv[1].w) &&
      upblock_supported(WL.lv[1].h, WL.lv[1].w, WL.lv[1].h, WL.lv[1].w)) {
    // ---- conv8 (+ skip) and conv9 + conv10 + sigmoid + pixel_shuffle as two fused blocks
    auto block = [&](int b, int slevel, const char* src, const char* res, const char* dst) -> int {
      UpBlockArgs a;
      memset(&a, 0, sizeof(a));
      a.mode = mode; a.N = B; a.Hs = WL.lv[slevel].h; a.Ws = WL.lv[slevel].w; a.H = WL.lv[1].h; a.W = WL.lv[1].w;
      a.Cmid = kBlocks[b].cin; a.Cout = kBlocks[b].cout;
      a.src = buf(src);
      a.w3 = wplanes(PL.w3[b]);
      a.w1 = b == 7 ? wplanes(PL.w1g7) : wplanes(PL.w1[b]);
      a.bias3 = fvec(PL.v3[b][0]); a.scale3 = fvec(PL.v3[b][1]); a.shift3 = fvec(PL.v3[b][2]);
      a.bias1 = fvec(PL.v1[b][0]); a.scale1 = fvec(PL.v1[b][1]); a.shift1 = fvec(PL.v1[b][2]);
      a.residual = res ? buf(res) : none;
      a.out = dst ? buf(dst) : none;
      a.tail = dst ? 0 : 1;
      a.w10 = fvec(PL.w10); a.b10 = fvec(PL.b10);
      a.y = y; a.y_u8 = y_u8;
      const double px = double(B) * a.H * a.W, spx = double(B) * a.Hs * a.Ws, ci = a.Cmid, co = a.Cout;
      char nm[48];
      snprintf(nm, sizeof(nm), "conv%d block (up+3x3+1x1%s)", b + 2, dst ? "+skip" : "+conv10+sigmoid");
      ProfScope ps(nm, 2.0 * px * (9 * ci * ci + ci * co + (dst ? 0 : 64)),
                   (spx * ci + (dst ? px * co * 2 : 0) + 9 * ci * ci + ci * co) * 2.0 * np + (dst ? 0 : px * 16), st);
      return upblock_launch(a, st);
    };
    NSM_TRY(block(6, 2, "m7", "c2", "m8"));
    NSM_TRY(block(7, 1, "m8", nullptr, nullptr));
    return 0;
  }
  NSM_TRY(up("m7", 2, 128, "u8", 1));
  NSM_TRY(double_conv(6, 1, buf("u8"), "t8", "m8", "c2", nullptr));   // conv8 + skip
  NSM_TRY(up("m8", 1, 64, "u9", 1));                                   // x2 up then back down to (h1, w1)
  NSM_TRY(double_conv(7, 1, buf("u9"), "t9", nullptr, nullptr, nullptr));  // conv9 3x3 stage
  // ---- tail: conv9 1x1 + conv10 + sigmoid + pixel_shuffle
  {
    TailParams tp;
    tp.a = buf("t9"); tp.N = B; tp.h = WL.lv[1].h; tp.w = WL.lv[1].w;
    tp.img = pb + PL.tail_img; tp.fmt = mode; tp.y = y; tp.y_u8 = y_u8;
    const double px = double(B) * tp.h * tp.w;
    ProfScope ps("tail(conv9.1x1+conv10)", px * 2.0 * (64 * 16 + 16 * 4), px * (64 * 2.0 * np + 16), st);
    NSM_TRY(tail_eval(tp, st));
  }
  return 0;
}

int nsm_unet_infer(const void* blob, int mode, const float* x, int B, int H, int W, const float* mean,
                   const float* std, float* y, void* ws, size_t ws_bytes, void* stream) {
  return infer_impl(blob, mode, x, B, H, W, mean, std, y, nullptr, ws, ws_bytes, stream);
}
int nsm_unet_infer_u8(const void* blob, int mode, const float* x, int B, int H, int W, const float* mean,
                      const float* std, uint8_t* y_u8, void* ws, size_t ws_bytes, void* stream) {
  return infer_impl(blob, mode, x, B, H, W, mean, std, nullptr, y_u8, ws, ws_bytes, stream);
}

int nsm_unet_infer_host(const void* blob, int mode, const float* x_host, int B, int H, int W, const float* mean,
                        const float* std, float* y_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B < 1 || H < 16 || W < 16) {
    set_error("nsm_unet_infer_host: bad shape");
    return 1;
  }
  const WorkspaceLayout WL = workspace_layout(B, H, W, mode);
  if (ws_bytes < WL.total) {
    set_error("nsm_unet_infer_host: workspace %zu B < required %zu B", ws_bytes, WL.total);
    return 1;
  }
  float* xd = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + WL.x_stage);
  float* yd = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + WL.y_stage);
  const size_t xin = size_t(B) * 4 * H * W * 4, yout = size_t(B) * WL.lv[0].h * WL.lv[0].w * 4;
  cudaError_t e = cudaMemcpyAsync(xd, x_host, xin, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    set_error("H2D copy: %s", cudaGetErrorString(e));
    return 1;
  }
  NSM_TRY(nsm_unet_infer(blob, mode, xd, B, H, W, mean, std, yd, ws, ws_bytes, stream));
  e = cudaMemcpyAsync(y_host, yd, yout, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("D2H copy / sync: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int nsm_unet_infer_host_u8(const void* blob, int mode, const float* x_host, int B, int H, int W, const float* mean,
                           const float* std, uint8_t* y_host, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (B < 1 || H < 16 || W < 16) {
    set_error("nsm_unet_infer_host_u8: bad shape");
    return 1;
  }
  const WorkspaceLayout WL = workspace_layout(B, H, W, mode);
  if (ws_bytes < WL.total) {
    set_error("nsm_unet_infer_host_u8: workspace %zu B < required %zu B", ws_bytes, WL.total);
    return 1;
  }
  float* xd = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + WL.x_stage);
  uint8_t* yd = reinterpret_cast<uint8_t*>(ws) + WL.y_stage;   // the fp32 staging area is large enough for uint8
  const size_t xin = size_t(B) * 4 * H * W * 4, yout = size_t(B) * WL.lv[0].h * WL.lv[0].w;
  cudaError_t e = cudaMemcpyAsync(xd, x_host, xin, cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    set_error("H2D copy: %s", cudaGetErrorString(e));
    return 1;
  }
  NSM_TRY(infer_impl(blob, mode, xd, B, H, W, mean, std, nullptr, yd, ws, ws_bytes, stream));
  e = cudaMemcpyAsync(y_host, yd, yout, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("D2H copy / sync: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------- frame pipeline
// infer.py / inference.py feed one frame after another.  The pipe keeps two device staging slots for inputs and results
// and three streams: the H2D copy of frame k+1 and the D2H copy of result k-1 overlap the kernels of frame k.
namespace {
struct FramePipe {
  const void* blob;
  int mode, B, H, W;
  const float *mean, *std;
  uint8_t* ws;
  size_t ws_bytes;       // the part nsm_unet_infer uses
  float* xd[2];
  uint8_t* yd[2];
  size_t x_bytes, y_elems;
  cudaStream_t s_in, s_comp, s_out;
  cudaEvent_t in_done[2], comp_done[2], out_done[2];
  unsigned long long submitted;
};
size_t pipe_slot_bytes(const WorkspaceLayout& WL, int B, int H, int W, size_t* xb, size_t* yb) {
  const size_t x = align_up(size_t(B) * 4 * H * W * 4, 256), y = align_up(size_t(B) * WL.lv[0].h * WL.lv[0].w * 4, 256);
  if (xb) *xb = x;
  if (yb) *yb = y;
  return x + y;
}
}  // namespace

size_t nsm_unet_pipe_workspace_bytes(int B, int H, int W, int mode) {
  if (B < 1 || H < 16 || W < 16 || mode < 0 || mode > 2) return 0;
  const WorkspaceLayout WL = workspace_layout(B, H, W, mode);
  return align_up(WL.total, 256) + 2 * pipe_slot_bytes(WL, B, H, W, nullptr, nullptr);
}

int nsm_unet_pipe_create(const void* blob, int mode, int B, int H, int W, const float* mean, const float* std,
                         void* ws, size_t ws_bytes, void** pipe) {
  if (!pipe || !blob || !ws) {
    set_error("nsm_unet_pipe_create: null argument");
    return 1;
  }
  const size_t need = nsm_unet_pipe_workspace_bytes(B, H, W, mode);
  if (need == 0 || ws_bytes < need) {
    set_error("nsm_unet_pipe_create: bad shape/mode or workspace %zu B < required %zu B", ws_bytes, need);
    return 1;
  }
  const WorkspaceLayout WL = workspace_layout(B, H, W, mode);
  FramePipe* fp = new FramePipe();
  fp->blob = blob; fp->mode = mode; fp->B = B; fp->H = H; fp->W = W; fp->mean = mean; fp->std = std;
  fp->ws = reinterpret_cast<uint8_t*>(ws);
  fp->ws_bytes = WL.total;
  size_t xb, yb;
  const size_t slot = pipe_slot_bytes(WL, B, H, W, &xb, &yb);
  uint8_t* base = fp->ws + align_up(WL.total, 256);
  for (int k = 0; k < 2; ++k) {
    fp->xd[k] = reinterpret_cast<float*>(base + k * slot);
    fp->yd[k] = base + k * slot + xb;
  }
  fp->x_bytes = size_t(B) * 4 * H * W * 4;
  fp->y_elems = size_t(B) * WL.lv[0].h * WL.lv[0].w;
  fp->submitted = 0;
  cudaError_t e = cudaSuccess;
  cudaStream_t* streams[3] = {&fp->s_in, &fp->s_comp, &fp->s_out};
  for (auto sp : streams)
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(sp, cudaStreamNonBlocking);
  for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
    e = cudaEventCreateWithFlags(&fp->in_done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fp->comp_done[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&fp->out_done[k], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    set_error("nsm_unet_pipe_create: %s", cudaGetErrorString(e));
    delete fp;
    return 1;
  }
  *pipe = fp;
  return 0;
}

// x_host [B,4,H,W] fp32; exactly one of y_host (fp32) / y_host_u8 receives [B,1,H',W'].  Returns once the frame is
// queued; when the call for frame k returns, the result of frame k-2 is complete in its host buffer.
int nsm_unet_pipe_submit(void* pipe, const float* x_host, float* y_host, uint8_t* y_host_u8) {
  FramePipe* fp = reinterpret_cast<FramePipe*>(pipe);
  if (!fp || !x_host || (!y_host == !y_host_u8)) {
    set_error("nsm_unet_pipe_submit: need a pipe, an input and exactly one output buffer");
    return 1;
  }
  const int k = int(fp->submitted & 1);
  cudaError_t e = cudaSuccess;
  if (fp->submitted >= 2) {
    e = cudaEventSynchronize(fp->out_done[k]);   // bounds the queue; frame k-2 is now on the host
    if (e == cudaSuccess) e = cudaStreamWaitEvent(fp->s_in, fp->comp_done[k], 0);   // its input slot is free
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(fp->xd[k], x_host, fp->x_bytes, cudaMemcpyHostToDevice, fp->s_in);
  if (e == cudaSuccess) e = cudaEventRecord(fp->in_done[k], fp->s_in);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(fp->s_comp, fp->in_done[k], 0);
  if (e != cudaSuccess) {
    set_error("nsm_unet_pipe_submit: H2D stage: %s", cudaGetErrorString(e));
    return 1;
  }
  // (the result slot is free: out_done[k] of frame k-2 was synchronised above)
  NSM_TRY(infer_impl(fp->blob, fp->mode, fp->xd[k], fp->B, fp->H, fp->W, fp->mean, fp->std,
                     y_host ? reinterpret_cast<float*>(fp->yd[k]) : nullptr, y_host ? nullptr : fp->yd[k], fp->ws,
                     fp->ws_bytes, fp->s_comp));
  e = cudaEventRecord(fp->comp_done[k], fp->s_comp);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(fp->s_out, fp->comp_done[k], 0);
  if (e == cudaSuccess)
    e = y_host ? cudaMemcpyAsync(y_host, fp->yd[k], fp->y_elems * 4, cudaMemcpyDeviceToHost, fp->s_out)
               : cudaMemcpyAsync(y_host_u8, fp->yd[k], fp->y_elems, cudaMemcpyDeviceToHost, fp->s_out);
  if (e == cudaSuccess) e = cudaEventRecord(fp->out_done[k], fp->s_out);
  if (e != cudaSuccess) {
    set_error("nsm_unet_pipe_submit: D2H stage: %s", cudaGetErrorString(e));
    return 1;
  }
  ++fp->submitted;
  return 0;
}

int nsm_unet_pipe_sync(void* pipe) {
  FramePipe* fp = reinterpret_cast<FramePipe*>(pipe);
  if (!fp) {
    set_error("nsm_unet_pipe_sync: null pipe");
    return 1;
  }
  cudaError_t e = cudaStreamSynchronize(fp->s_out);
  if (e == cudaSuccess) e = cudaStreamSynchronize(fp->s_comp);
  if (e != cudaSuccess) {
    set_error("nsm_unet_pipe_sync: %s", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

int nsm_unet_pipe_destroy(void* pipe) {
  FramePipe* fp = reinterpret_cast<FramePipe*>(pipe);
  if (!fp) return 0;
  cudaStreamSynchronize(fp->s_in);
  cudaStreamSynchronize(fp->s_comp);
  cudaStreamSynchronize(fp->s_out);
  for (int k = 0; k < 2; ++k) {
    cudaEventDestroy(fp->in_done[k]);
    cudaEventDestroy(fp->comp_done[k]);
    cudaEventDestroy(fp->out_done[k]);
  }
  cudaStreamDestroy(fp->s_in);
  cudaStreamDestroy(fp->s_comp);
  cudaStreamDestroy(fp->s_out);
  delete fp;
  return 0;
}

int nsm_unet_tap(const void* ws, int B, int H, int W, int mode, const char* name, float* out, int* C, int* h,
                 int* w, void* stream) {
  const int t = tap_index(name);
  if (t < 0) {
    set_error("nsm_unet_tap: unknown tap '%s'", name);
    return 1;
  }
  const WorkspaceLayout WL = workspace_layout(B, H, W, mode);
  const int np = fmt_planes(mode);
  const Level& lv = WL.lv[kTaps[t].level];
  if (C) *C = kTaps[t].C;
  if (h) *h = lv.h;
  if (w) *w = lv.w;
  if (!out) return 0;
  const Planes p = ws_planes(WL, const_cast<void*>(ws), name, np);
  const bool upsampled = name[0] == 'u';   // u6..u9 live in the decoder's operand format
  return planes_to_nchw(p.p[0], p.p[1], B, kTaps[t].C, lv.h, lv.w, upsampled ? decoder3x3_fmt(mode) : mode, out,
                        static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------- profiling
int nsm_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  g_prof_on = on != 0;
  for (auto& r : g_prof) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  g_prof.clear();
  return 0;
}

int nsm_profile_read(char* out, size_t cap) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  std::string txt;
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.stop) != cudaSuccess) {
      set_error("nsm_profile_read: event sync failed");
      return 1;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.start, r.stop);
    char line[160];
    snprintf(line, sizeof(line), "%s,%.6f,%.6e,%.6e\n", r.name.c_str(), ms, r.flops, r.bytes);
    txt += line;
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  g_prof.clear();
  if (txt.size() + 1 > cap) {
    set_error("nsm_profile_read: buffer too small (%zu needed)", txt.size() + 1);
    return 1;
  }
  memcpy(out, txt.c_str(), txt.size() + 1);
  return 0;
}

// ---------------------------------------------------------------------------------------------- stages
int nsm_nchw_to_planes(const float* x, int N, int C, int H, int W, int mode, void* p0, void* p1, void* stream) {
  return nchw_to_planes(x, N, C, H, W, mode, p0, p1, static_cast<cudaStream_t>(stream));
}
int nsm_planes_to_nchw(const void* p0, const void* p1, int N, int C, int H, int W, int mode, float* y,
                       void* stream) {
  return planes_to_nchw(p0, p1, N, C, H, W, mode, y, static_cast<cudaStream_t>(stream));
}
int nsm_pack_conv_weight(const float* w, int Cout, int Cin, int ksize, int dgrad, int mode, void* p0, void* p1,
                         void* stream) {
  return pack_conv_weight(w, Cout, Cin, ksize, dgrad, mode, p0, mode != NSM_MODE_BF16 ? p1 : nullptr,
                          static_cast<cudaStream_t>(stream));
}

int nsm_conv_fwd(const nsm_conv_args* a, void* stream) {
  if (!a) {
    set_error("nsm_conv_fwd: null args");
    return 1;
  }
  ConvShape s = {a->N, a->H, a->W, a->Cin, a->Cout, a->ksize * a->ksize, a->mode};
  Planes in = {{const_cast<void*>(a->in[0]), const_cast<void*>(a->in[1])}};
  Planes w = {{const_cast<void*>(a->weight[0]), const_cast<void*>(a->weight[1])}};
  ConvEpilogue e;
  e.bias = a->bias; e.scale = a->bn_scale; e.shift = a->bn_shift; e.lrelu = a->lrelu;
  e.round_bf16 = a->mode == NSM_MODE_BF16;
  e.out = {{a->out[0], a->out[1]}};
  e.residual = {{const_cast<void*>(a->residual[0]), const_cast<void*>(a->residual[1])}};
  e.pool = {{a->pool[0], a->pool[1]}};
  e.out_f32 = a->out_f32;
  e.stats = A(a->stats);
  char nm[64];
  snprintf(nm, sizeof(nm), "conv_gemm k%d %d->%d @%dx%d", a->ksize, a->Cin, a->Cout, a->H, a->W);
  const double px = double(a->N) * a->H * a->W;
  ProfScope ps(nm, 2.0 * px * s.taps * a->Cin * a->Cout,
               (px * (a->Cin + a->Cout) + double(s.taps) * a->Cin * a->Cout) * 2.0 * fmt_planes(a->mode),
               static_cast<cudaStream_t>(stream));
  return conv_gemm_launch(s, in, w, e, static_cast<cudaStream_t>(stream));
}

int nsm_unet_fused_decoder(void) { return fused_decoder(NSM_MODE_FP32) ? 1 : 0; }
int nsm_unet_set_fused_decoder(int on) {
  g_fused_override.store(on < 0 ? -1 : (on ? 1 : 0), std::memory_order_relaxed);
  return 0;
}

int nsm_upblock_prof(unsigned long long* out8) { return upblock_prof(out8); }

int nsm_upblock(const nsm_upblock_args* a, void* stream) {
  if (!a) {
    set_error("nsm_upblock: null args");
    return 1;
  }
  UpBlockArgs u;
  memset(&u, 0, sizeof(u));
  u.mode = a->mode; u.N = a->N; u.Hs = a->Hs; u.Ws = a->Ws; u.H = a->H; u.W = a->W; u.Cmid = a->Cmid; u.Cout = a->Cout;
  u.src = mk(a->src[0], a->src[1]);
  u.w3 = mk(a->weight3[0], a->weight3[1]);
  u.w1 = mk(a->weight1[0], a->weight1[1]);
  u.bias3 = a->bias3; u.scale3 = a->bn_scale3; u.shift3 = a->bn_shift3;
  u.bias1 = a->bias1; u.scale1 = a->bn_scale1; u.shift1 = a->bn_shift1;
  u.residual = mk(a->residual[0], a->residual[1]);
  u.out = mk(a->out[0], a->out[1]);
  u.tail = a->tail; u.w10 = a->w10; u.b10 = a->b10; u.y = a->y; u.y_u8 = a->y_u8;
  ProfScope ps("upblock", 2.0 * double(a->N) * a->H * a->W * (9.0 * a->Cmid * a->Cmid + double(a->Cmid) * a->Cout), 0.0,
               S(stream));
  return upblock_launch(u, S(stream));
}

int nsm_upsample_match(const void* const* src, int N, int hs, int ws, int C, void* const* dst, int hd, int wd,
                       int mode, void* stream) {
  ProfScope ps_("upsample_match", 0.0, (double(N) * hs * ws + double(N) * hd * wd) * C * 2.0 * fmt_planes(mode), S(stream));
  Planes s = {{const_cast<void*>(src[0]), const_cast<void*>(src[1])}};
  Planes d = {{dst[0], dst[1]}};
  return upsample_match(s, N, hs, ws, C, d, hd, wd, mode, static_cast<cudaStream_t>(stream));
}

int nsm_add_noise_clamp(const float* x, const float* noise, long long numel, float eps, float lo, float hi, float* out,
                        void* stream) {
  if (!x || !noise || !out) {
    set_error("nsm_add_noise_clamp: null pointer");
    return 1;
  }
  ProfScope ps_("add_noise_clamp", 0.0, double(numel) * 12.0, S(stream));
  return add_noise_clamp(x, noise, numel, eps, lo, hi, out, S(stream));
}
int nsm_mse_loss_fwd_bwd(const float* out, const float* ref, long long numel, float* diff, nsm_acc* acc, void* stream) {
  if (!out || !ref || !acc) {
    set_error("nsm_mse_loss_fwd_bwd: null pointer");
    return 1;
  }
  ProfScope ps_("mse_loss", 0.0, double(numel) * 4.0 * (diff ? 3 : 2), S(stream));
  return mse_loss_fwd_bwd(out, ref, numel, diff, A(acc), S(stream));
}
int nsm_l1_loss_fwd_bwd(const float* out, const float* target, const float* const* perturbed, int n_perturbed,
                        long long numel, float coef_l1, float coef_pert, float* grad, nsm_acc* acc, void* stream) {
  ProfScope ps_("l1_loss", 0.0, double(numel) * 4.0 * (3 + n_perturbed), S(stream));
  return l1_loss_fwd_bwd(out, target, perturbed, n_perturbed, numel, coef_l1, coef_pert, grad, A(acc),
                         static_cast<cudaStream_t>(stream));
}
__global__ void acc_to_double_kernel(const Acc* acc, long long n, double* out) {
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  if (i < n) out[i] = acc_load(acc + i);
}
int nsm_acc_to_double(const nsm_acc* acc, long long n, double* out, void* stream) {
  if (n <= 0) return 0;
  acc_to_double_kernel<<<unsigned((n + 255) / 256), 256, 0, S(stream)>>>(A(acc), n, out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("nsm_acc_to_double launch failed: %s", cudaGetErrorString(e));
    return 1;
  }
  count_launch();
  return 0;
}
int nsm_channel_sums(const float* x, long long S, int C, long long HW, const double* means, nsm_acc* sums,
                     void* stream) {
  return channel_sums(x, S, C, HW, means, A(sums), static_cast<cudaStream_t>(stream));
}
int nsm_standardize(const float* x, float* y, long long S, int C, long long HW, const float* mean,
                    const float* std, void* stream) {
  return standardize(x, y, S, C, HW, mean, std, static_cast<cudaStream_t>(stream));
}
int nsm_perturb(const float* x, const float* noise, float* out, int count, long long B, int C, long long HW,
                const float* stds, float std_factor, void* stream) {
  return perturb(x, noise, out, count, B, C, HW, stds, std_factor, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------------- training stages

int nsm_bn_stats(const void* z0, const void* z1, long long P, int C, int mode, nsm_acc* sums, void* stream) {
  ProfScope ps_("bn_stats", 0.0, double(P) * C * 2.0 * fmt_planes(mode), S(stream));
  return bn_stats(mk(z0, z1), P, C, mode, A(sums), S(stream));
}
int nsm_bn_finalize(const nsm_acc* sums, long long P, int C, const float* gamma, const float* beta, float eps,
                    float momentum, int updates, float* running_mean, float* running_var, float* scale, float* shift,
                    float* save_mean, float* save_invstd, void* stream) {
  return bn_finalize(A(sums), P, C, gamma, beta, eps, momentum, updates, running_mean, running_var, scale, shift, save_mean,
                     save_invstd, S(stream));
}
int nsm_bn_act(const void* z0, const void* z1, int N, int H, int W, int C, int mode, const float* scale,
               const float* shift, const float* mask, int lrelu, const void* res0, const void* res1, void* out0,
               void* out1, void* pool0, void* pool1, void* stream) {
  ProfScope ps_("bn_act", 0.0, double(N) * H * W * C * 4.0 * fmt_planes(mode), S(stream));
  BnActParams p;
  p.z = mk(z0, z1); p.out = mk(out0, out1); p.residual = mk(res0, res1); p.pool = mk(pool0, pool1);
  p.N = N; p.H = H; p.W = W; p.C = C; p.fmt = mode; p.scale = scale; p.shift = shift; p.mask = mask; p.lrelu = lrelu;
  return bn_act(p, S(stream));
}
int nsm_bn_bwd(const void* dy0, const void* dy1, const void* z0, const void* z1, int N, int H, int W, int C, int mode,
               const float* scale, const float* shift, const float* mask, const float* mean, const float* invstd,
               int lrelu, nsm_acc* sums, void* dz0, void* dz1, float* dgamma, float* dbeta, float* dbias, void* stream) {
  ProfScope ps_("bn_bwd", 0.0, double(N) * H * W * C * 10.0 * fmt_planes(mode), S(stream));
  BnBwdParams p;
  p.dy = mk(dy0, dy1); p.z = mk(z0, z1); p.dz = mk(dz0, dz1);
  p.N = N; p.H = H; p.W = W; p.C = C; p.fmt = mode; p.scale = scale; p.shift = shift; p.mask = mask; p.mean = mean;
  p.invstd = invstd; p.lrelu = lrelu; p.sums = A(sums); p.dbias = A(sums) + 2 * C;
  NSM_TRY(bn_bwd_reduce(p, S(stream)));
  NSM_TRY(bn_bwd_apply(p, S(stream)));
  return bn_bwd_finalize(A(sums), A(sums) + 2 * C, mean, invstd, C, mode == NSM_MODE_BF16, dgamma, dbeta, dbias, S(stream));
}
int nsm_pool_bwd_add(const void* a0, const void* a1, const void* dp0, const void* dp1, void* out0, void* out1, int N,
                     int H, int W, int C, int mode, void* stream) {
  ProfScope ps_("pool_bwd_add", 0.0, double(N) * H * W * C * 4.5 * fmt_planes(mode), S(stream));
  return pool_bwd_add(mk(a0, a1), mk(dp0, dp1), mk(out0, out1), N, H, W, C, mode, S(stream));
}
int nsm_planes_add(const void* a0, const void* a1, const void* b0, const void* b1, void* out0, void* out1,
                   long long numel, int mode, void* stream) {
  return planes_add(mk(a0, a1), mk(b0, b1), mk(out0, out1), numel, mode, S(stream));
}
int nsm_bilinear_bwd(const void* dout0, const void* dout1, int N, int ho, int wo, int C, void* din0, void* din1, int hi,
                     int wi, int mode, void* stream) {
  ProfScope ps_("bilinear_bwd", 0.0, (double(N) * ho * wo + double(N) * hi * wi) * C * 2.0 * fmt_planes(mode), S(stream));
  return bilinear_bwd(mk(dout0, dout1), N, ho, wo, C, mk(din0, din1), hi, wi, mode, S(stream));
}
int nsm_upsample_match_bwd(const void* dout0, const void* dout1, int N, int ho, int wo, int C, void* din0, void* din1,
                           int hi, int wi, int mode, void* stream) {
  ProfScope ps_("upsample_match_bwd", 0.0, (double(N) * ho * wo + double(N) * hi * wi) * C * 2.0 * fmt_planes(mode),
                S(stream));
  return upsample_match_bwd(mk(dout0, dout1), N, ho, wo, C, mk(din0, din1), hi, wi, mode, S(stream));
}
int nsm_train_input_prep(const float* x, int N, int Hin, int Win, void* out0, void* out1, int mode, void* stream) {
  ProfScope ps_("train_input_prep", 0.0, double(N) * Hin * Win * 16.0, S(stream));
  return train_input_prep(x, N, Hin, Win, mk(out0, out1), mode, S(stream));
}
int nsm_train_input_grad(const void* d0, const void* d1, int N, int H, int W, float* dx, int mode, void* stream) {
  return train_input_grad(mk(d0, d1), N, H, W, dx, mode, S(stream));
}
int nsm_sigmoid_shuffle_fwd(const void* c0, const void* c1, int N, int h, int w, int mode, float* y, void* stream) {
  ProfScope ps_("sigmoid_shuffle_fwd", 0.0, double(N) * h * w * 16.0, S(stream));
  return sigmoid_shuffle_fwd(mk(c0, c1), N, h, w, mode, y, S(stream));
}
int nsm_sigmoid_shuffle_bwd(const float* dy, const float* y, int N, int h, int w, int mode, void* d0, void* d1,
                            void* stream) {
  ProfScope ps_("sigmoid_shuffle_bwd", 0.0, double(N) * h * w * 32.0, S(stream));
  return sigmoid_shuffle_bwd(dy, y, N, h, w, mode, mk(d0, d1), S(stream));
}
int nsm_pack_conv_weight_padded(const float* w, int Cout, int Cin, int ksize, int CoutP, int CinP, int dgrad, int mode,
                                void* plane0, void* plane1, void* stream) {
  return pack_conv_weight_padded(w, Cout, Cin, ksize, CoutP, CinP, dgrad, mode, plane0, plane1, S(stream));
}
int nsm_train_input_prep_c16(const float* x, int N, int Hin, int Win, void* out0, void* out1, int mode, void* stream) {
  ProfScope ps_("train_input_prep", 0.0, double(N) * Hin * Win * (4.0 + 2.0 * fmt_planes(mode)), S(stream));
  return train_input_prep(x, N, Hin, Win, mk(out0, out1), mode, S(stream), 16);
}
int nsm_train_input_grad_c16(const void* d0, const void* d1, int N, int H, int W, float* dx, int mode, void* stream) {
  return train_input_grad(mk(d0, d1), N, H, W, dx, mode, S(stream), 16);
}
int nsm_sigmoid_shuffle_fwd_px4(const void* c0, const void* c1, int N, int h, int w, int mode, float* y, void* stream) {
  ProfScope ps_("sigmoid_shuffle_fwd", 0.0, double(N) * h * w * (16.0 + 32.0 * fmt_planes(mode)), S(stream));
  return sigmoid_shuffle_fwd(mk(c0, c1), N, h, w, mode, y, S(stream), 1);
}
int nsm_sigmoid_shuffle_bwd_px4(const float* dy, const float* y, int N, int h, int w, int mode, void* d0, void* d1,
                                void* stream) {
  ProfScope ps_("sigmoid_shuffle_bwd", 0.0, double(N) * h * w * (32.0 + 32.0 * fmt_planes(mode)), S(stream));
  return sigmoid_shuffle_bwd(dy, y, N, h, w, mode, mk(d0, d1), S(stream), 1);
}
int nsm_pack_conv_weight_px4(const float* w, int Cout, int Cin, int ksize, int CoutV, int CinV, int dgrad, int mode,
                             void* plane0, void* plane1, void* stream) {
  return pack_conv_weight_px4(w, Cout, Cin, ksize, CoutV, CinV, dgrad, mode, plane0, plane1, S(stream));
}
int nsm_px4_reduce_dw(const float* dwv, int Cout, int Cin, int ksize, int CoutV, int CinV, float* dw, void* stream) {
  return px4_reduce_dw(dwv, Cout, Cin, ksize, CoutV, CinV, dw, S(stream));
}
int nsm_fold_channel_sums(const nsm_acc* in, int nvec, int CV, int groups, int C, nsm_acc* out, void* stream) {
  return fold_channel_sums(A(in), nvec, CV, groups, C, A(out), S(stream));
}
int nsm_tile_vector(const float* src, int n, int rep, int npad, float fill, int round_bf16, float* dst, void* stream) {
  return tile_vector(src, n, rep, npad, fill, round_bf16, dst, S(stream));
}
int nsm_pad_vector(const float* src, int n, int npad, float fill, int round_bf16, float* dst, void* stream) {
  return pad_vector(src, n, npad, fill, round_bf16, dst, S(stream));
}
size_t nsm_wgrad_workspace_bytes(int N, int H, int W, int Cout, int Cin, int ksize, int mode) {
  WgradShape s = {N, H, W, Cout, Cin, ksize * ksize, mode};
  return wgrad_workspace_bytes(s);
}
int nsm_wgrad(const void* dz0, const void* dz1, const void* x0, const void* x1, int N, int H, int W, int Cout, int Cin,
              int ksize, int mode, int Cout_real, int Cin_real, void* workspace, size_t workspace_bytes, float* dw,
              void* stream) {
  WgradShape s = {N, H, W, Cout, Cin, ksize * ksize, mode};
  char nm[64];
  snprintf(nm, sizeof(nm), "wgrad_gemm k%d %d->%d @%dx%d", ksize, Cin, Cout, H, W);
  const double px = double(N) * H * W;
  ProfScope ps(nm, 2.0 * px * ksize * ksize * Cin * Cout, px * (Cin + Cout) * 2.0 * fmt_planes(mode), S(stream));
  return wgrad_launch(s, mk(dz0, dz1), mk(x0, x1), workspace, workspace_bytes, Cout_real, Cin_real,
                      mode == NSM_MODE_BF16, dw, S(stream));
}

}  // extern "C"
