// Fused decoder block (upblock.cu): composite bilinear up-sample on the operand path -> 3x3 conv + BN + LReLU -> 1x1 conv
// + BN + LReLU -> skip add | conv10 + sigmoid + pixel_shuffle, one launch (Unetmodel.py:134-148, eval mode).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "conv_gemm.cuh"

namespace nsm {

struct UpBlockArgs {
  int mode;            // 0 bf16 | 1 fp32 eval (source fp16 hi+lo planes)
  int N, Hs, Ws;       // source (low-resolution) tensor [N,Hs,Ws,Cmid]
  int H, W;            // output resolution of the block (the skip tensor's size)
  int Cmid, Cout;      // 3x3: Cmid -> Cmid, 1x1: Cmid -> Cout   (128/64: conv8, 64/16: conv9)
  Planes src;          // planes of the source
  Planes w3;           // 3x3 weights [Cmid][9][Cmid]: bf16 | fp16 hi + 8-bit cross plane (kFmtF16X8)
  Planes w1;           // 1x1 weights [Cout][Cmid]:    bf16 | fp16 hi + lo
  const float *bias3, *scale3, *shift3;   // [Cmid] conv bias, eval-BN scale / shift of the 3x3 stage
  const float *bias1, *scale1, *shift1;   // [Cout] same for the 1x1 stage
  Planes residual;     // [N,H,W,Cout] skip tensor added after the block (or nullptr)
  Planes out;          // [N,H,W,Cout] (tail == 0)
  int tail;            // 1: Cout == 16, followed by conv10 (w10 [4][16], b10 [4]) + sigmoid + pixel_shuffle(2)
  const float *w10, *b10;
  float* y;            // [N,1,2H,2W] fp32 (tail)
  uint8_t* y_u8;       // optional uint8 form of y
};
int upblock_launch(const UpBlockArgs& a, cudaStream_t st);
// false: the resize needs a wider vertical stencil than the kernel's job table holds (use the stage-by-stage path)
bool upblock_supported(int Hs, int Ws, int H, int W);
int upblock_prof(unsigned long long* out);

}  // namespace nsm
