// Direct fp32 kernels of the 3 -> 32 channel 3x3 image conv1 (first_conv.cu).
#pragma once
#include "common.cuh"

namespace var {
struct FirstConvArgs {
  const void* x;                 // [N, 3, H, W] via element strides (uint8 or fp32)
  long long sN, sH, sW, sC;
  float scale;                   // 1/255 for uint8 frames
  int N, H, W, P, Q, stride;     // pad 1, 3x3
  const float* w;                // packed [32][kpad], k = (r*3 + s)*3 + c
  int kpad;
  const float* bias;
  float* y;                      // fwd: [N, P, Q, 32]
  int relu, round_out;
  const float* dy;               // wgrad: [N, P, Q, 32]
  float* dw;                     // wgrad: packed [32][kpad] (+=)
  float* db;                     // wgrad: [32] (+=, nullable)
  int rows;                      // set by the launcher
};
int first_conv_fwd(FirstConvArgs a, int u8, cudaStream_t st);
int first_conv_wgrad(FirstConvArgs a, int u8, cudaStream_t st);
}  // namespace var
