// Persistent tcgen05 kernels of the iTHOR sound conv1 (1 -> 64 channels, 11x11, stride 2, pad 5
// over [N, H, 40] MFCC maps; models/pretext/ai2thor_pretext_model.py:27) -- cin1_conv.cu.
#pragma once
#include "common.cuh"

namespace var {
struct Cin1Args {
  const float* x;      // [N, H, 40] fp32, contiguous
  int N, H, P;         // P = output rows (Q = 20)
  const float* w;      // fwd: packed tf32 weights [64][128], k = r*11 + s
  const float* bias;   // fwd: [64] fp32 (nullable)
  float* y;            // fwd: [N, P, 20, 64] fp32, or IEEE f16 when out_f16 != 0 (the 16-bit conv region reads it)
  int relu, round_out, out_f16;
  const float* dy;     // wgrad: [N, P, 20, 64] (tf32-rounded values)
  float* dw;           // wgrad: packed [64][128] (+=)
  float* db;           // wgrad: [64] (+=, nullable)
};
// true when (geometry, layout) is the one these kernels are specialised for
bool cin1_conv_match(int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw,
                     long long sN, long long sH, long long sW, float scale, const void* x);
int cin1_conv_fwd(const Cin1Args& a, cudaStream_t st);
int cin1_conv_wgrad(const Cin1Args& a, cudaStream_t st);
}  // namespace var
