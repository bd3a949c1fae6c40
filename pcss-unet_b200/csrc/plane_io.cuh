// 8-channel (16-byte) access to NHWC activation planes in any storage format (nsm_common.cuh), shared by the streaming
// kernels of the training path and of the perceptual (VGG) term.
#pragma once
#include "conv_gemm.cuh"
#include "nsm_common.cuh"

namespace nsm {

// ------------------------------------------------------------------------------------------------
// 8-channel plane access
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8(const Planes& p, size_t elem, int fmt, float* v) {
  const uint4 hv = ldg16(reinterpret_cast<const uint8_t*>(p.p[0]) + elem * 2);
  const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = hi_lo_to_f32(hw[e], fmt);
    v[2 * e + 1] = hi_hi_to_f32(hw[e], fmt);
  }
  if (fmt != kFmtBf16) {
    const uint4 lv = ldg16(reinterpret_cast<const uint8_t*>(p.p[1]) + elem * 2);
    const uint32_t lw[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] += lo_lo_to_f32(lw[e], fmt);
      v[2 * e + 1] += lo_hi_to_f32(lw[e], fmt);
    }
  }
}
__device__ __forceinline__ void store8(const Planes& p, size_t elem, int fmt, const float* v) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    hw[e] = pack_hi(v[2 * e], v[2 * e + 1], fmt);
    lw[e] = pack_lo_resid(v[2 * e], v[2 * e + 1], hw[e], fmt);
  }
  stg16(reinterpret_cast<uint8_t*>(p.p[0]) + elem * 2, make_uint4(hw[0], hw[1], hw[2], hw[3]));
  if (fmt != kFmtBf16) stg16(reinterpret_cast<uint8_t*>(p.p[1]) + elem * 2, make_uint4(lw[0], lw[1], lw[2], lw[3]));
}

// raw 16-byte words of 8 channels (both planes): lets a loop issue several independent loads before converting
struct Raw8 {
  uint4 h, l;
};
__device__ __forceinline__ Raw8 load_raw8(const Planes& p, size_t elem, int fmt) {
  Raw8 r;
  r.h = ldg16(reinterpret_cast<const uint8_t*>(p.p[0]) + elem * 2);
  if (fmt != kFmtBf16) r.l = ldg16(reinterpret_cast<const uint8_t*>(p.p[1]) + elem * 2);
  else r.l = make_uint4(0, 0, 0, 0);
  return r;
}
__device__ __forceinline__ void unpack8(const Raw8& r, int fmt, float* v) {
  const uint32_t hw[4] = {r.h.x, r.h.y, r.h.z, r.h.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[2 * e] = hi_lo_to_f32(hw[e], fmt);
    v[2 * e + 1] = hi_hi_to_f32(hw[e], fmt);
  }
  if (fmt != kFmtBf16) {
    const uint32_t lw[4] = {r.l.x, r.l.y, r.l.z, r.l.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] += lo_lo_to_f32(lw[e], fmt);
      v[2 * e + 1] += lo_hi_to_f32(lw[e], fmt);
    }
  }
}

}  // namespace nsm
