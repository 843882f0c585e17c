// Kuka sound branch forward in full fp32 (models/pretext/arm_pretext_model.py:21-34,51-55):
//   conv(5x40, s(2,1)) 1->32, 3 x conv(3x1, s(2,1)) 32->32, ReLU each, flatten, Linear 160->128, ReLU.
// The branch is 0.43 M MAC per clip on 16 KB of input with M = 48/23/11/5 rows per layer: far
// too small for 128-row MMA tiles, and its first layer multiplies MFCC values of magnitude
// 1e1-1e2 where tf32 rounding alone costs ~1e-3 of the embedding.  One CTA per clip keeps the
// clip, every activation and the conv weights in shared memory and uses plain FFMA, so the
// forward embedding is exact to fp32 round-off.  Activations are also written to HBM (NHWC,
// tf32-rounded) because the backward pass reuses the tcgen05 engine on them.
#include "kuka_sound.cuh"

namespace var {

namespace {
constexpr int kF = 100, kMelW = 40, kC = 32;
constexpr int kP1 = 48, kP2 = 23, kP3 = 11, kP4 = 5;
constexpr int kK1 = 200, kK1Pad = 224, kK2 = 96, kHid = 128, kFlat = 160;
constexpr int kThreads = 256;

// out[p][c] = relu(b[c] + sum_{r<3, ci<32} in[(2p+r)][ci] * wt[(r*32+ci)][c]); lane = c
template <int PIN, int POUT>
__device__ __forceinline__ void conv3x1(const float* __restrict__ in, const float* __restrict__ wt,
                                        const float* __restrict__ bias, float* __restrict__ out,
                                        int warp, int lane) {
  constexpr int kWarps = kThreads / 32;
  constexpr int J = (POUT + kWarps - 1) / kWarps;
  float acc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = bias[lane];
#pragma unroll 4
  for (int k = 0; k < kK2; ++k) {
    const float w = wt[k * kC + lane];
    const int r = k >> 5, ci = k & 31;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int p = warp + kWarps * j;
      if (p < POUT) acc[j] = fmaf(in[(2 * p + r) * kC + ci], w, acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int p = warp + kWarps * j;
    if (p < POUT) out[p * kC + lane] = fmaxf(acc[j], 0.f);
  }
}

__device__ __forceinline__ void store_rounded(const float* __restrict__ s, float* __restrict__ g, int n) {
  if (!g) return;
  for (int i = threadIdx.x; i < n; i += kThreads) g[i] = round_tf32(s[i]);
}
}  // namespace

// Persistent CTAs: the transposed conv weights (62 KB) are staged ONCE per CTA, clips stream through a cp.async
// double buffer (restaging the weights per clip cost more than the clip's arithmetic).
__global__ void __launch_bounds__(kThreads, 2) kuka_sound_fwd_kernel(KukaSoundArgs a, int N) {
  extern __shared__ __align__(16) float sm[];
  float* sx0 = sm;                      // 2 x [100*40]
  float* w1t = sx0 + 2 * kF * kMelW;    // [200][32]
  float* w2t = w1t + kK1 * kC;          // 3 x [96][32]
  float* a1 = w2t + 3 * kK2 * kC;       // [48*32]
  float* a2 = a1 + kP1 * kC;            // [23*32]
  float* a3 = a2 + kP2 * kC;            // [11*32]
  float* a4 = a3 + kP3 * kC;            // [5*32]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  auto load_clip = [&](int n, int b) {
    const float* src = a.x + (long long)n * kF * kMelW;
    const uint32_t dst = smem_u32(sx0 + b * kF * kMelW);
    for (int i = tid; i < kF * kMelW / 4; i += kThreads) cp_async_16(dst + (uint32_t)i * 16u, src + 4 * i, 16u);
  };
  int n = blockIdx.x, buf = 0;
  if (n < N) load_clip(n, 0);
  cp_async_commit();
  {  // transposed conv weights
    for (int i = tid; i < kK1 * kC; i += kThreads) {
      const int c = i / kK1, k = i - c * kK1;  // coalesced read of packed [32][224]
      w1t[k * kC + c] = a.w1[c * kK1Pad + k];
    }
    for (int l = 0; l < 3; ++l) {
      const float* w = l == 0 ? a.w2 : (l == 1 ? a.w3 : a.w4);
      for (int i = tid; i < kK2 * kC; i += kThreads) {
        const int c = i / kK2, k = i - c * kK2;
        w2t[(l * kK2 + k) * kC + c] = w[c * kK2 + k];
      }
    }
  }
  const float b1 = a.b1[lane];
  for (; n < N; n += gridDim.x, buf ^= 1) {
    cp_async_wait<0>();
    __syncthreads();  // clip n landed; every read of the activation buffers by the previous clip is done
    if (n + (int)gridDim.x < N) load_clip(n + gridDim.x, buf ^ 1);
    cp_async_commit();
    const float* sx = sx0 + buf * kF * kMelW;

    // conv1: out[p][c] = relu(b + sum_{k<200} x[2p*40 + k] * w1t[k][c])   (5x40 window is contiguous);
    // lane = c, 6 rows per warp, x read as broadcast 128-bit loads (10 shared loads per 24 FMAs)
    {
      constexpr int J = kP1 / 8;
      float acc[J];
#pragma unroll
      for (int j = 0; j < J; ++j) acc[j] = b1;
#pragma unroll 2
      for (int k = 0; k < kK1; k += 4) {
        const float w0 = w1t[k * kC + lane], w1 = w1t[(k + 1) * kC + lane], w2 = w1t[(k + 2) * kC + lane],
                    w3 = w1t[(k + 3) * kC + lane];
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float4 xv = *reinterpret_cast<const float4*>(sx + (2 * (warp + 8 * j)) * kMelW + k);
          acc[j] = fmaf(xv.x, w0, acc[j]); acc[j] = fmaf(xv.y, w1, acc[j]);
          acc[j] = fmaf(xv.z, w2, acc[j]); acc[j] = fmaf(xv.w, w3, acc[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < J; ++j) a1[(warp + 8 * j) * kC + lane] = fmaxf(acc[j], 0.f);
    }
    __syncthreads();
    conv3x1<kP1, kP2>(a1, w2t, a.b2, a2, warp, lane);
    __syncthreads();
    conv3x1<kP2, kP3>(a2, w2t + kK2 * kC, a.b3, a3, warp, lane);
    __syncthreads();
    conv3x1<kP3, kP4>(a3, w2t + 2 * kK2 * kC, a.b4, a4, warp, lane);
    __syncthreads();

    store_rounded(a1, a.act1 ? a.act1 + (long long)n * kP1 * kC : nullptr, kP1 * kC);
    store_rounded(a2, a.act2 ? a.act2 + (long long)n * kP2 * kC : nullptr, kP2 * kC);
    store_rounded(a3, a.act3 ? a.act3 + (long long)n * kP3 * kC : nullptr, kP3 * kC);
    store_rounded(a4, a.act4 + (long long)n * kP4 * kC, kP4 * kC);

    // Linear 160 -> 128 + ReLU: 16 outputs per warp, lanes split k (weights stream from L1 / L2)
    for (int o = warp * 16; o < warp * 16 + 16; ++o) {
      const float* w = a.wl + (long long)o * kFlat;
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < kFlat / 32; ++u) acc = fmaf(a4[lane + 32 * u], w[lane + 32 * u], acc);
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
      if (lane == 0) a.hidden[(long long)n * kHid + o] = fmaxf(acc + a.bl[o], 0.f);
    }
  }
}

int kuka_sound_fwd(const KukaSoundArgs& a, int N, cudaStream_t st) {
  if (N <= 0) return VAR_OK;
  const size_t smem = (size_t)(2 * kF * kMelW + kK1 * kC + 3 * kK2 * kC + (kP1 + kP2 + kP3 + kP4) * kC) * 4;
  VAR_ENSURE_SMEM(kuka_sound_fwd_kernel, smem);
  const int grid = N < 2 * kNumSMs ? N : 2 * kNumSMs;
  LaunchScope sc(T_MISC, 2.0 * N * 447872.0, st);
  kuka_sound_fwd_kernel<<<grid, kThreads, smem, st>>>(a, N);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}


// ---------------------------------------------------------------------------
// Weight + bias gradient of a full-width first conv (Cin = 1, S == W, pad 0, Q == 1, Cout = 32): Kuka soundCNN.0,
// models/pretext/arm_pretext_model.py (5 x 40 kernel over [N, 100, 40] MFCC maps, stride (2, 1)).
// The im2col row of output pixel (n, p) is the CONTIGUOUS run x[n][p*sh .. p*sh + R - 1][0 .. W - 1] of K = R*W
// floats, so dW[o][k] = sum_{n,p} dY[n,p,o] * x[n][p*sh*W + k] needs no gather at all.  The GEMM is tiny (10 GFLOP at
// N = 16384) and HBM bound (x + dY = 360 MB); the generic first-layer tensor path spent 3.4 ms on it (per-tile set-up
// for 48-row tiles).  Here: persistent CTAs, one clip per iteration double-buffered through cp.async, plain fp32 FMAs
// with a 4 (k) x 8 (o) register tile per thread, partial sums kept in registers across ALL clips of the CTA and
// reduced with one atomic per element at the end.
// ---------------------------------------------------------------------------
struct FullwArgs {
  const float* x; const float* dy; float* dw; float* db;
  int N, HW, K, P, step, kpad;  // HW floats per clip, K = R*W, step = sh*W
};
constexpr int kFwThreads = 256, kFwCout = 32;
__global__ void __launch_bounds__(kFwThreads)
fullw_wgrad_kernel(const FullwArgs a) {
  extern __shared__ __align__(16) uint8_t fw_smem[];
  const int tid = threadIdx.x;
  const int xs_floats = (a.HW + 3) & ~3, ds_floats = a.P * kFwCout;
  const int buf_floats = xs_floats + ds_floats;
  float* bufs = reinterpret_cast<float*>(fw_smem);
  const int groups = a.K >> 2;                 // 4-k groups; 4 * groups <= 224 compute threads
  const bool compute = tid < 4 * groups;
  const int og = compute ? tid / groups : 0, k4 = compute ? tid - og * groups : 0;
  const bool biasthr = tid >= kFwThreads - kFwCout;
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
  float bacc = 0.f;
  auto load = [&](int n, int b) {
    const uint32_t dst = smem_u32(bufs + (size_t)b * buf_floats);
    const float* xg = a.x + (long long)n * a.HW;
    const float* dg = a.dy + (long long)n * ds_floats;
    for (int i = tid; i < a.HW / 4; i += kFwThreads) cp_async_16(dst + (uint32_t)i * 16u, xg + 4 * i, 16u);
    for (int i = tid; i < ds_floats / 4; i += kFwThreads)
      cp_async_16(dst + (uint32_t)(xs_floats + 4 * i) * 4u, dg + 4 * i, 16u);
  };
  int n = blockIdx.x, b = 0;
  if (n < a.N) load(n, 0);
  cp_async_commit();
  for (; n < a.N; n += gridDim.x, b ^= 1) {
    const int nn = n + gridDim.x;
    if (nn < a.N) load(nn, b ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* xs = bufs + (size_t)b * buf_floats;
    const float* ds = xs + xs_floats;
    if (compute) {
      const float* xp = xs + 4 * k4;
      const float* dp = ds + 8 * og;
#pragma unroll 4
      for (int p = 0; p < a.P; ++p) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + p * a.step);
        const float4 d0 = *reinterpret_cast<const float4*>(dp + p * kFwCout);
        const float4 d1 = *reinterpret_cast<const float4*>(dp + p * kFwCout + 4);
        const float xc[4] = {xv.x, xv.y, xv.z, xv.w};
        const float dj[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c][j] = fmaf(xc[c], dj[j], acc[c][j]);
      }
    } else if (biasthr) {
      const int o = tid - (kFwThreads - kFwCout);
      for (int p = 0; p < a.P; ++p) bacc += ds[p * kFwCout + o];
    }
    __syncthreads();
  }
  if (compute) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) atomicAdd(a.dw + (long long)(8 * og + j) * a.kpad + 4 * k4 + c, acc[c][j]);
  } else if (biasthr && a.db) {
    atomicAdd(a.db + tid - (kFwThreads - kFwCout), bacc);
  }
}

bool fullw_conv_match(int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw, int P, int Q,
                      long long sN, long long sH, long long sW, float scale, const void* x) {
  const int K = R * W;
  return Cin == 1 && Cout == kFwCout && S == W && sw == 1 && ph == 0 && pw == 0 && Q == 1 && P >= 1 && (K & 3) == 0 &&
         K <= 224 && ((H * W) & 3) == 0 && ((sh * W) & 3) == 0 && sW == 1 && sH == W && sN == (long long)H * W &&
         scale == 1.f && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
         (size_t)2 * (((H * W + 3) & ~3) + P * kFwCout) * 4 <= 200 * 1024;
}

int fullw_conv_wgrad(const float* x, const float* dy, float* dw, float* db, int N, int H, int W, int R, int sh, int P,
                     int kpad, cudaStream_t st) {
  if (N <= 0) return VAR_OK;
  FullwArgs a;
  a.x = x; a.dy = dy; a.dw = dw; a.db = db;
  a.N = N; a.HW = H * W; a.K = R * W; a.P = P; a.step = sh * W; a.kpad = kpad;
  const size_t smem = (size_t)2 * (((a.HW + 3) & ~3) + P * kFwCout) * 4;
  VAR_ENSURE_SMEM(fullw_wgrad_kernel, smem);
  int grid = kNumSMs * (smem * 4 + 4096 <= 227 * 1024 ? 4 : (smem * 2 + 2048 <= 227 * 1024 ? 2 : 1));
  if (grid > N) grid = N;
  LaunchScope sc(T_WGRAD, 2.0 * N * (double)P * a.K * kFwCout, st);
  fullw_wgrad_kernel<<<grid, kFwThreads, smem, st>>>(a);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var
