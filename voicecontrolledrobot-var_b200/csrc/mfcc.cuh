// Fused MFCC front-end launchers (mfcc.cu).
#pragma once
#include "common.cuh"

namespace var {
struct MfccPlan;
int mfcc_plan_create(int flavour, int fs, int n_fft, int win_length, int hop, MfccPlan** out);
void mfcc_plan_destroy(MfccPlan* p);
int mfcc_num_frames(const MfccPlan* p, int n_samples);
int mfcc_fwd(const MfccPlan* p, const int16_t* wav, const long long* offsets, const int* lengths,
             int B, int F, float* out, cudaStream_t st);
}  // namespace var
