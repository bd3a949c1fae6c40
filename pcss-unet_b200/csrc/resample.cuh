// Bilinear align_corners=True resampling arithmetic shared by the forward and adjoint kernels.
// Mirrors ATen: scale = (in-1)/(out-1) (0 if out == 1), src = scale*dst, i0 = floor(src) clamped, lambda = src - i0.
#pragma once
#include <cuda_runtime.h>

namespace nsm {

struct Lerp {
  int i0, i1;
  float w0, w1;
};
__host__ __device__ __forceinline__ Lerp make_lerp(int dst, int in_size, int out_size) {
  const float scale = out_size > 1 ? float(in_size - 1) / float(out_size - 1) : 0.f;
  const float src = scale * float(dst);
  int i0 = int(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  Lerp l;
  l.i0 = i0;
  l.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  float lam = src - float(i0);
  lam = fminf(fmaxf(lam, 0.f), 1.f);
  l.w1 = lam;
  l.w0 = 1.f - lam;
  return l;
}

// Composite of nn.Upsample(x2) followed by F.interpolate(size=out): output index `dst` (0..out) depends on at most
// three consecutive source indices rmin, rmin+1, rmin+2 (clamped to in-1) with weights w[0..2]:
//   out(dst) = sum_k a_k * up(Y_k),  up(Y) = sum_j b_j(Y) * src(R_j(Y)),  Y_1 = Y_0 (+1)  =>  R in [R_0(Y_0), R_0(Y_0)+2]
struct Tap3 {
  int rmin;
  float w[3];
};
__host__ __device__ __forceinline__ Tap3 composite_taps(int dst, int in_size, int out_size) {
  const int mid = 2 * in_size;
  const Lerp m = make_lerp(dst, mid, out_size);
  const Lerp a = make_lerp(m.i0, in_size, mid), b = make_lerp(m.i1, in_size, mid);
  Tap3 t;
  t.rmin = a.i0;
  t.w[0] = t.w[1] = t.w[2] = 0.f;
  const int idx[4] = {a.i0, a.i1, b.i0, b.i1};
  const float wv[4] = {m.w0 * a.w0, m.w0 * a.w1, m.w1 * b.w0, m.w1 * b.w1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = idx[k] - t.rmin;
    t.w[0] += d == 0 ? wv[k] : 0.f;
    t.w[1] += d == 1 ? wv[k] : 0.f;
    t.w[2] += d == 2 ? wv[k] : 0.f;
  }
  return t;
}

}  // namespace nsm
