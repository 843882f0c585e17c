// Host side of the tcgen05 engine: TMA descriptor cache and kernel launchers.
#pragma once
#include "tc_engine.cuh"

namespace var {

// MN-major (transposed) tf32 operands use the 32-byte-granule 128B swizzle.
struct MnCfg {
  int lbo = 4096, sbo = 512, type = 1, swz32 = 1;
  int tma_swizzle = (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
};
MnCfg& mn_cfg();  // process-wide (selftest may override to probe the hardware)

// 2-D fp32 tensor map over a row-major [rows, cols] matrix with row pitch
// `pitch` elements; box = {32 cols, box_rows}.  Cached by value of all args.
int get_tmap_2d(const float* ptr, int rows, int cols, long long pitch, int box_rows, int swizzle,
                CUtensorMap* out);
// esize-aware forms (esize 4: fp32 boxes of 32 elements; esize 2: f16 / bf16 boxes of 64 elements)
int get_tmap_2d_e(const void* ptr, int esize, int rows, int cols, long long pitch, int box_rows, int swizzle,
                  CUtensorMap* out);
int get_tmap_im2col_e(const void* ptr, int esize, int N, int H, int W, int C, int low_w, int low_h, int up_w,
                      int up_h, int stride_w, int stride_h, int pixels, int swizzle, CUtensorMap* out);
int get_tmap_im2col(const float* ptr, int N, int H, int W, int C, int low_w, int low_h, int up_w,
                    int up_h, int stride_w, int stride_h, int pixels, int swizzle, CUtensorMap* out);
// 3-D fp32 tensor map over a dense [d2][d1][d0] array, box {b0, b1, 1}, no swizzle (not cached).
int get_tmap_3d(const float* ptr, int d0, int d1, int d2, int b0, int b1, CUtensorMap* out);
int gather_mode();  // 1: TMA-fed A operands (default), 0: cp.async gather (VAR_GATHER=cp_async)

struct ConvShape {
  int N, H, W, Cin;       // input  (H, W spatial)
  int Cout, R, S;
  int sh, sw, ph, pw;
  int P, Q;               // output spatial
};

enum SrcKind : int { SRC_NHWC_F32 = 0, SRC_STRIDED_F32 = 1, SRC_STRIDED_U8 = 2 };

struct SrcLayout {       // element strides for strided sources
  long long sN, sH, sW, sC;
  float scale;
};

inline int round_up32(int k) { return (k + 31) & ~31; }

// y[N,P,Q,Cout] = act(conv(x, w) + b).  w packed [Cout, Kpad], k = (r*S+s)*Cin + c.
// out_kind 1 stores y as IEEE f16 (only the first-layer kernel feeding the 16-bit region supports it).
int conv_fwd(const ConvShape& cs, const void* x, int src_kind, const SrcLayout* sl, const float* w,
             const float* bias, float* y, int relu, int round_out, cudaStream_t st, int out_kind = 0);

// dx[N,H,W,Cin] = conv_transpose(dy, w) [* (mask > 0)] ; mask = forward output of the
// previous layer (ReLU backward fused), may be null.
// `addsrc` ([rows, Cin], nullable) is added before the mask.
int conv_dgrad(const ConvShape& cs, const float* dy, const float* w, float* dx, const float* mask,
               const float* addsrc, int round_out, cudaStream_t st);
int linear_dgrad2(int ndir, int M, int Cin, int Cout, const float* const dy[2],
                  const float* const w[2], float* const dx[2], const float* const addsrc[2],
                  int round_out, cudaStream_t st);
int gru_step_fwd(int ndir, int B, int Hd, const float* const hprev_r[2], const float* const whh[2],
                 const GruEpiParams q[2], cudaStream_t st);
int gru_step_bwd(int ndir, int B, int Hd, const float* const dgh[2], const float* const whh[2],
                 const GruBwdEpiParams q[2], cudaStream_t st);
// whh16 / h_h (nullable): f16 copies of W_hh and a [(T+1), B, H] f16 hidden-state buffer (slot 0 zeroed) -> the
// recurrent GEMM runs on 16-bit operands (half the per-step operand stream).
int gru_persist_fwd(int B, int Hd, int T, const float* const xproj[2], long long ldx, const float* const whh[2],
                    const float* const bhh[2], float* const h32[2][2], float* const h_r[2],
                    float* const gates[2], float* const hn_save[2], unsigned int* counters, cudaStream_t st,
                    const void* const whh16[2] = nullptr, void* const h_h[2] = nullptr, long long x_tstride = 0);
struct GruBwdExtra {
  const void* whh16[2];    // f16 W_hh copies
  void* dgh_h[2];          // [T][B, 3H] f16 scaled gate gradients; slot T-1 filled by the caller (grad_to_f16_scaled)
  const float* gscale[2];  // device {S, 1/S} per direction
  float* db_ih[2];         // bias gradients accumulated inside the BPTT kernel (nullable)
  float* db_hh[2];
  int* bias_done;          // out: 1 when the kernel produced the bias gradients (else the caller sums columns)
  // 16-bit input projection: scaled f16 dgi, element (b, t, c) at b * dgi_h_ld + t * dgi_h_ts + c (nullable); when the
  // kernel wrote it (*dgi_h_done = 1) the fp32 dgi / dgh of steps < T-1 were NOT stored
  void* dgi_h[2];
  long long dgi_h_ld, dgi_h_ts;
  int* dgi_h_done;
};
int gru_persist_bwd(int B, int Hd, int T, const float* const whh[2], const float* const gates[2],
                    const float* const hn_save[2], const float* const h_r[2], float* const dgh[2],
                    float* const dgi[2], float* const dhd[2][2], unsigned int* counters, cudaStream_t st,
                    const GruBwdExtra* ex = nullptr);
bool gru_h16_enabled();
bool gru_x16_enabled();
int colsum(const float* dy, long long M, long long ld, int C, float* db, cudaStream_t st);

// dw[Cout, Kpad] += im2col(x)^T dy ; db[Cout] += colsum(dy) (db may be null).
int conv_wgrad(const ConvShape& cs, const void* x, int src_kind, const SrcLayout* sl,
               const float* dy, float* dw, float* db, cudaStream_t st);

// ---- 16-bit operand region (f16 NHWC activations, f16 packed weights, f16 scaled gradients) ----
bool conv_h16_ok(const ConvShape& cs);
int conv_fwd_h16(const ConvShape& cs, const void* x, const void* w, const float* bias, void* y, int out_kind, int relu,
                 int round_out, cudaStream_t st);
int conv_dgrad_h16(const ConvShape& cs, const void* dy, const void* w, void* dx, int out_kind, const void* mask,
                   int mask_kind, const float* out_scale, int round_out, cudaStream_t st);
int conv_wgrad_h16(const ConvShape& cs, const void* x, const void* dy, float* dw, float* db, const float* inv_scale,
                   int* db_done, cudaStream_t st);
int linear_h16(int M, int K, int N, const void* a, long long lda, const void* w, const float* bias, float* out,
               long long ldo, const void* mask, int mask_kind, long long ldm, const float* out_scale, int round_out,
               int is_dgrad, cudaStream_t st);
int linear_wgrad_h16(int M, int K, int N, const void* x, long long ldx, const void* dy, long long ldy, float* dw,
                     int kpad, const float* inv_scale, cudaStream_t st);

}  // namespace var
