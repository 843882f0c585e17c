// Persistent tcgen05 kernels of the image conv1 of both encoders (3 -> 32 channels, 3x3, pad 1,
// stride 1 or 2, uint8 NCHW frames; models/pretext/arm_pretext_model.py:11,
// ai2thor_pretext_model.py:15 with the /255 of dataset.py:67-68 folded in) -- cin3_conv.cu.
#pragma once
#include "common.cuh"

namespace var {
struct Cin3Args {
  const unsigned char* x;  // [N, 3, H, W] uint8, contiguous, W % 4 == 0
  float scale;             // 1/255
  int N, H, W, P, Q, stride;
  const float* w;          // fwd: packed tf32 weights [32][32], k = (r*3 + s)*3 + c
  const float* bias;       // fwd: [32] (nullable)
  float* y;                // fwd: [N, P, Q, 32]
  int relu, round_out;
  const float* dy;         // wgrad: [N, P, Q, 32] (tf32-rounded values)
  float* dw;               // wgrad: packed [32][32] (+=)
  float* db;               // wgrad: [32] (+=, nullable)
};
bool cin3_conv_match(int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw, int P, int Q,
                     long long sN, long long sH, long long sW, long long sC, const void* x);
int cin3_conv_fwd(const Cin3Args& a, cudaStream_t st);
int cin3_conv_wgrad(const Cin3Args& a, cudaStream_t st);
}  // namespace var
