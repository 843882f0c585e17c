// Launchers of the HBM-bound helper kernels (aux_kernels.cu).
#pragma once
#include "common.cuh"

namespace var {

int pack_weight(const float* ref, float* master, float* mma, int O, int I, int R, int S, int kpad,
                cudaStream_t st);
int unpack_weight(const float* packed, float* ref, int O, int I, int R, int S, int kpad,
                  cudaStream_t st);
int round_copy(const float* src, float* dst, long long n, cudaStream_t st);

// 16-bit operand region helpers (aux_kernels.cu)
int cvt_f16(const float* src, void* dst, long long n, cudaStream_t st);
int grad_to_f16_scaled(const float* src, void* dst, long long n, float* scale, unsigned int* amax_bits, cudaStream_t st);
// last BPTT step: one gradient scale for both directions (amax over the dgi slices), dgh[d] ([rows, cols] dense) and
// the dgi[d] slices (row pitches ldgi -> ldgi_h) stored as f16 * S
int gru_last_to_f16(const float* const dgh[2], void* const dgh_h[2], const float* const dgi[2], long long ldgi,
                    void* const dgi_h[2], long long ldgi_h, int rows, int cols, float* scale, unsigned int* amax_bits,
                    cudaStream_t st);
// dst[c][col0 + r] = f16(src[r][c]) (fp32 [R, C] -> f16 rows of pitch dst_ld)
int cvt_f16_transpose(const float* src, void* dst, int R, int C, long long dst_ld, int dst_col0, cudaStream_t st);

int maxpool_fwd(const float* x, float* y, int N, int H, int W, int C, cudaStream_t st);
int maxpool_bwd(const float* x, const float* dy, float* dx, int N, int H, int W, int C,
                cudaStream_t st);

struct GruBwdArgs {
  const float* dh;       // [B, H] grad wrt h_t
  const float* gates;    // [B, 3H] saved r, z, n
  const float* hn_save;  // [B, H] saved W_hn h + b_hn
  const float* hprev;    // [B, H] h_{t-1}
  float* dgi;            // [B, ldgi] slot for this step (batch-major [B, T, 3H])
  long long ldgi;
  float* dgh;            // [B, 3H]
  float* dhd;            // [B, H] dh * z
};
int gru_cell_bwd(const GruBwdArgs& a0, const GruBwdArgs& a1, int ndir, int B, int Hd,
                 cudaStream_t st);

int adam_step(float* p, const float* g, float* m, float* v, float* p_mma, long long n, float lr,
              float beta1, float beta2, float eps, float wd, long long step, float gscale,
              cudaStream_t st);

int reward_normalize(const float* rew, const unsigned char* done, int N, double* ret, double* rms,
                     double gamma, double eps, double cliprew, int update_rms, float* orig, float* out,
                     cudaStream_t st);
int nhwc_to_nchw(const float* x, float* y, int B, int HW, int C, cudaStream_t st);
int concat2(const float* a, const float* c, float* out, float* out_r, int B, int Hd,
            cudaStream_t st);
int split2(const float* x, float* a, float* c, int B, int Hd, cudaStream_t st);

}  // namespace var
