// Fused fp32 forward of the Kuka sound branch (kuka_sound.cu).
#pragma once
#include "common.cuh"

namespace var {
struct KukaSoundArgs {
  const float* x;                       // [N, 100, 40] MFCC features
  const float *w1, *b1;                 // packed [32][224], [32]   (soundCNN.0)
  const float *w2, *b2, *w3, *b3, *w4, *b4;  // packed [32][96], [32] (soundCNN.2/4/6)
  const float *wl, *bl;                 // packed [128][160], [128] (soundTriplet.0)
  float *act1, *act2, *act3;            // [N,48,32] [N,23,32] [N,11,32] tf32-rounded (nullable: inference)
  float* act4;                          // [N,5,32]  tf32-rounded (raw feature, always written)
  float* hidden;                        // [N,128]   fp32, post-ReLU (input of the fused tail)
};
int kuka_sound_fwd(const KukaSoundArgs& a, int N, cudaStream_t st);
// weight + bias gradient of a full-width first conv (Cin = 1, S == W, Q == 1, Cout = 32, contiguous fp32 input):
// dw[32][kpad] += dY^T x-windows, db[32] += colsum(dY)
bool fullw_conv_match(int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw, int P, int Q,
                      long long sN, long long sH, long long sW, float scale, const void* x);
int fullw_conv_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int H, int W, int R, int sh, int P,
                     int kpad, cudaStream_t st);
}  // namespace var
