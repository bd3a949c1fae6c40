// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM wrappers, bf16 helpers.
// Everything here is inline PTX for Blackwell (compile with -gencode arch=compute_100a,code=sm_100a).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nsm {

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// Order-independent accumulator (= nsm_acc of the C ABI).  Cross-block reductions (BatchNorm statistics, BatchNorm-backward
// sums, loss sums, the gradient norm) add their per-block fp64 partials as 128-bit two's-complement FIXED-POINT numbers
// (LSB 2^-64, |partial| < 2^62) with two integer atomics.  Integer addition is associative, so the total is bit-identical
// whatever order the blocks arrive in -- an fp64 atomicAdd is not, and made training runs differ from run to run where
// the reference asks for deterministic kernels (main.py:81-82).  A partial that is NaN / Inf / out of range sets `bad`
// and the slot reads back as NaN, so non-finite values still surface in the loss and the statistics.
// ---------------------------------------------------------------------------------------------
struct Acc {
  unsigned long long lo, hi, bad, pad;
};
static_assert(sizeof(Acc) == 32, "nsm_acc is four 64-bit words");

__device__ __forceinline__ void acc_split(double v, unsigned long long& lo, unsigned long long& hi) {
  const double f = floor(v);   // v - f is exact in [0, 1] (1 only when v is a tiny negative number: saturates below)
  lo = __double2ull_rz((v - f) * 18446744073709551616.0);
  hi = (unsigned long long)__double2ll_rz(f);
}
__device__ __forceinline__ void acc_add(Acc* a, double v) {
  if (!(fabs(v) < 4611686018427387904.0)) {   // NaN, Inf, |v| >= 2^62
    atomicOr(&a->bad, 1ull);
    return;
  }
  unsigned long long lo, hi;
  acc_split(v, lo, hi);
  const unsigned long long old = atomicAdd(&a->lo, lo);
  hi += (old + lo < lo) ? 1ull : 0ull;   // carry out of the low word: their number depends only on the total
  if (hi) atomicAdd(&a->hi, hi);
}
__device__ __forceinline__ double acc_load(const Acc* a) {
  if (a->bad) return __longlong_as_double(0x7ff8000000000000LL);
  return double((long long)a->hi) + double(a->lo) * 5.42101086242752217e-20;
}
// plain (non-atomic) store of a value into a slot nobody else is writing
__device__ __forceinline__ void acc_store(Acc* a, double v) {
  Acc r{0ull, 0ull, 0ull, 0ull};
  if (!(fabs(v) < 4611686018427387904.0)) r.bad = 1ull;
  else acc_split(v, r.lo, r.hi);
  *a = r;
}

// round-to-nearest-even fp32 -> bf16 -> fp32 (the rounding points of torch.autocast(bfloat16))
__device__ __forceinline__ float rbf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
// the same for two values with ONE packed conversion (F2FP.BF16.PACK_AB on the ALU pipe) + two unpacking logic ops; the
// single conversion above is an F2F on the quarter-rate conversion pipe, which bounded the bf16-mode epilogues (three to
// four roundings per output element)
__device__ __forceinline__ void rbf2(float& a, float& b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  const uint32_t r = *reinterpret_cast<const uint32_t*>(&t);
  a = __uint_as_float(r << 16);
  b = __uint_as_float(r & 0xffff0000u);
}

// fp32-mode split of activations and weights: v ~= hi (fp16, saturated to +-65504) + lo (fp16):
// |v - hi - lo| <= max(2^-23 |v|, 2^-25) for |v| <= 65504; the dropped lo*lo product of the 3-term GEMM is <= 2^-22
// relative.  (kind::f16 MMAs must not mix fp16 and bf16 operands -- the hardware raises an illegal instruction -- so
// both planes of both operands are fp16 in this mode; bf16 mode uses one bf16 plane.)
__device__ __forceinline__ float sat_f16(float v) { return fminf(fmaxf(v, -65504.f), 65504.f); }
__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(sat_f16(v));
  lo = __float2half_rn(sat_f16(v - __half2float(hi)));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

__device__ __forceinline__ uint32_t pack_f16(float a, float b) {
  __half2 t = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float f16lo_to_f32(uint32_t w) {
  return __low2float(*reinterpret_cast<const __half2*>(&w));
}
__device__ __forceinline__ float f16hi_to_f32(uint32_t w) {
  return __high2float(*reinterpret_cast<const __half2*>(&w));
}

// Activation / weight storage formats ("fmt", equal to the C ABI's NSM_MODE_* values):
//   0  one bf16 plane                      (bf16 mode, autocast rounding points)
//   1  hi + lo fp16 planes, |v| <= 65504   (fp32 mode, eval: 22 significand bits)
//   2  hi + lo bf16 planes, full range     (fp32 mode, training: 16 significand bits, safe for tiny gradients)
//   3  fp16 hi plane + an 8-bit "cross" plane (fp32 mode, eval, operands of the decoder's 3x3 convolutions only):
//      per 16 channels the cross plane holds 16 bytes + 16 bytes of e4m3 values that ONE 8-bit MMA of K = 32 multiplies
//      pairwise: activations [v * 2 | (v - hi) * 2^11], weights [(w - hi) * 2^16 | w * 2^6], so that
//          sum_32 a8 * w8 = 2^17 * (v~ * w_lo + v_lo * w~)   = the two cross terms of the hi+lo product, at twice the
//      tensor rate of the fp16 pipe.  The 3-bit mantissas cost 2^-4 relative on terms that are 2^-12 of the result
//      (CPU emulation: 4e-6 rms per layer against 1.2e-6 for fp16 hi+lo and 2e-6 for bf16 hi+lo).
constexpr int kFmtBf16 = 0, kFmtF16x2 = 1, kFmtBf16x2 = 2, kFmtF16X8 = 3;
constexpr float kX8ActHi = 2.f, kX8ActLo = 2048.f, kX8WgtLo = 65536.f, kX8WgtHi = 64.f;
constexpr float kX8CrossScale = 1.f / 131072.f;   // 2^-17 = 1 / (kX8ActHi * kX8WgtLo) = 1 / (kX8ActLo * kX8WgtHi)
__host__ __device__ __forceinline__ int fmt_planes(int fmt) { return fmt == 0 ? 1 : 2; }
__host__ __device__ __forceinline__ bool fmt_is_f16(int fmt) { return fmt == kFmtF16x2 || fmt == kFmtF16X8; }

// two floats -> packed fp16 pair, saturating to +-65504 in ONE instruction (low half = a)
__device__ __forceinline__ uint32_t pack_f16_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ uint32_t pack_hi(float a, float b, int fmt) {
  return fmt_is_f16(fmt) ? pack_f16_sat(a, b) : pack_bf16(a, b);
}
__device__ __forceinline__ float hi_lo_to_f32(uint32_t w, int fmt) {
  return fmt_is_f16(fmt) ? f16lo_to_f32(w) : bf16lo_to_f32(w);
}
__device__ __forceinline__ float hi_hi_to_f32(uint32_t w, int fmt) {
  return fmt_is_f16(fmt) ? f16hi_to_f32(w) : bf16hi_to_f32(w);
}
// two floats -> two e4m3 bytes (low byte = a), round to nearest, saturating to +-448
__device__ __forceinline__ uint32_t pack_e4m3x2(float a, float b) {
  unsigned short r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float e4m3_to_f32(uint32_t byte) {
  uint32_t h2;
  asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(h2) : "h"((unsigned short)(byte & 0xffu)));
  return f16lo_to_f32(h2);
}
// kFmtF16X8 activations: 8 consecutive channels -> 8 bytes for the first half (v * 2) and 8 for the second ((v - hi) * 2^11)
// of their 16-channel group; hw[] are the already packed fp16 hi words of the same 8 values.
__device__ __forceinline__ void x8_act_bytes(const float (&v)[8], const uint32_t (&hw)[4], uint2& first, uint2& second) {
  uint32_t a[4], b[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float v0 = v[2 * e], v1 = v[2 * e + 1];
    a[e] = pack_e4m3x2(v0 * kX8ActHi, v1 * kX8ActHi);
    b[e] = pack_e4m3x2((v0 - f16lo_to_f32(hw[e])) * kX8ActLo, (v1 - f16hi_to_f32(hw[e])) * kX8ActLo);
  }
  first = make_uint2(a[0] | (a[1] << 16), a[2] | (a[3] << 16));
  second = make_uint2(b[0] | (b[1] << 16), b[2] | (b[3] << 16));
}
// byte offset, inside one pixel's (or weight row's) 2*C-byte cross-plane row, of channel c's byte in the first half of its
// 16-channel group (the second-half byte is 16 further)
__host__ __device__ __forceinline__ int x8_byte(int c) { return (c >> 4) * 32 + (c & 15); }
// lo plane word of a value pair given the packed hi word
__device__ __forceinline__ uint32_t pack_lo_resid(float a, float b, uint32_t hw, int fmt) {
  // |residual| <= half an ulp of hi, so it cannot overflow unless hi itself saturated (then satfinite clamps it)
  return fmt == kFmtF16x2 ? pack_f16_sat(a - f16lo_to_f32(hw), b - f16hi_to_f32(hw))
                          : pack_bf16(a - bf16lo_to_f32(hw), b - bf16hi_to_f32(hw));
}
__device__ __forceinline__ float lo_lo_to_f32(uint32_t w, int fmt) { return hi_lo_to_f32(w, fmt); }
__device__ __forceinline__ float lo_hi_to_f32(uint32_t w, int fmt) { return hi_hi_to_f32(w, fmt); }
// scalar split used by the (non-vectorised) packing kernels: returns raw 16-bit patterns
__device__ __forceinline__ void split_fmt(float v, int fmt, unsigned short& hi, unsigned short& lo) {
  const uint32_t hw = pack_hi(v, 0.f, fmt);
  hi = (unsigned short)(hw & 0xffffu);
  lo = (unsigned short)(pack_lo_resid(v, 0.f, hw, fmt) & 0xffffu);
}
// (formats 0..2; the cross plane of kFmtF16X8 is not element-addressable as 16-bit words)
__device__ __forceinline__ float join_fmt(unsigned short hi, unsigned short lo, int fmt) {
  float v = hi_lo_to_f32(hi, fmt);
  if (fmt != kFmtBf16) v += lo_lo_to_f32(lo, fmt);
  return v;
}

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails loudly) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    // ~60 s at 2 GHz: far beyond any legitimate wait even when the context is time-sliced or preempted (a trap is sticky and
    // kills the CUDA context), still short enough that a protocol bug ends the launch instead of hanging the GPU box
    if (clock64() - t0 > 120000000000LL) {
      printf("nsm: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tiled mode, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// 1-D bulk copy global -> shared (no tensor map; 16-byte aligned addresses, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_saddr, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// TMA store (shared -> global, tiled, bulk-group completion).  The issuing thread must make the generic-proxy writes to
// shared memory visible first (fence_proxy_async after the writes, then a warp/CTA barrier).
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src_saddr, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_saddr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_4d_saddr(uint32_t dst_saddr, const CUtensorMap* m, uint64_t* bar, int c0,
                                                  int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_saddr), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source (it may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... all but the most recent one
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// all bulk groups of this thread are complete (writes performed)
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint4 lds16(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts16(uint32_t saddr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A * B with 8-bit floating-point operands (e4m3 / e5m2 per the instruction descriptor), K = 32 per instruction
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile with 64-byte rows (32 bf16 / 64 e4m3 of K per row) and the 64-byte swizzle: 8-row atoms of 512 B
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(1) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) |
         (uint64_t(4) << 61);
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// A operand from tensor memory (".ts" form): lane = M row, every 32-bit column holds two consecutive 16-bit K elements (low
// half = even k); one instruction consumes K = 16 elements = 8 columns.  B from shared memory as usual.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: 32 lanes x 8 consecutive 32-bit columns; thread i of the warp writes lane (base + i)
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of one cluster (same TPC) execute one MMA of M = 256; each holds its own 128 rows of
// A, HALF of the B tile and its own 128 accumulator lanes.  Only the leader (cluster rank 0) issues MMAs and commits; the
// TMA loads of both CTAs complete on the LEADER's mbarrier (shared::cluster address with the peer bit 24 cleared).
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {  // same warp in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// L2 eviction-priority policies for TMA copies (createpolicy): evict_last for operands every work item streams again (the
// weight matrix of a big convolution), so that the activations and outputs passing through L2 do not push them out
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_2d_pair_hint(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                      uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier at this shared-memory offset in BOTH CTAs of the pair arrives once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((unsigned short)3)
      : "memory");
}
// Multicast within a cluster (single-CTA MMAs): the box lands at the same shared-memory offset in every CTA of ctaMask and
// completes on each of those CTAs' own copy of the mbarrier; the commit arrives on every masked CTA's copy of the barrier.
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               unsigned short mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, unsigned short mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// arrive on the LEADER CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask)
               : "memory");
}

// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::f16, fp32 accumulate.
//   bits [4,6) c_format=1(F32)  [7,10) a_format (0=F16, 1=BF16)  [10,13) b_format (0=F16, 1=BF16)
//   bit 15 a_major (0=K, 1=MN)  bit 16 b_major  [17,23) N>>3  [24,29) M>>4
constexpr uint32_t kFmtF16 = 0, kFmtBF16 = 1;
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, uint32_t a_fmt, uint32_t b_fmt, int a_mn_major,
                                                      int b_mn_major) {
  return (1u << 4) | (a_fmt << 7) | (b_fmt << 10) | (uint32_t(a_mn_major) << 15) | (uint32_t(b_mn_major) << 16) |
         (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), 128-byte swizzle:
//   [0,14) start>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [61,64) layout=2 (SWIZZLE_128B)
// K-major operand tile (rows = M/N index, 128 B = 64 bf16 of K per row): 8-row swizzle atoms of 1024 B,
// SBO = 1024 B between atoms, LBO unused (1).  Advancing K by 16 elements = +32 B on the start address.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((smem_addr & 0x3FFFFu) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) |
         (uint64_t(sbo_bytes >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61);
}

// TMEM -> registers: 32 lanes x 32 consecutive fp32 columns; thread i of the warp gets lane (base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16-byte global access helpers
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stg16(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

}  // namespace nsm
