// Fused MFCC front-end: int16 wav -> reflect pad -> framing ->
// periodic-Hamming window (centred in n_fft) -> real FFT -> |.|^2 -> HTK mel
// filterbank -> log(x + 1e-6) -> ortho DCT-II -> [B, F, 40] crop / zero-pad.
// Replaces torchaudio.transforms.MFCC as configured at Envs/audioLoader.py:147-157
// plus processSoundFeat (Envs/audioLoader.py:241-252).
//
// One CTA per (clip, chunk of <= kFramesPerCta frames), 8 warps, one frame per warp at a time.  Interior frames read
// their sample pairs as 32-bit words straight from global memory (frames overlap, so the re-reads hit L1/L2; no
// staging phase, no staging shared memory -> 4 CTAs = 32 warps per SM); frames touching a clip end take a scalar path
// with the reflect / zero padding.  The packed-real frame goes through a Stockham FFT of n_fft/2 complex points,
// radix 8 x 8 x 4 (or 8 x 8 x 8), with the two exchanges through a per-warp shared buffer of float2 values in
// layouts that make every 64- / 128-bit access conflict free.  The last pass gives lane l the butterfly columns l
// and 64 - l, so both members Z[m], Z[N - m] of every real-FFT pair sit in the same lane: the split needs no shuffles
// and one pair yields two power-spectrum bins.  The 40 mel filters are dot products over aligned 4-bin groups (lane =
// filter, the 8 widest filters split over 4 lanes), then log; once all frames of the chunk are done the CTA applies
// the 40 x 40 DCT with 4 x 4 register tiles and writes `[F, 40]` rows with 128-bit stores.  Pure fp32 CUDA-core
// work, bound by shared-memory wavefronts and instruction issue (profiles/r02_mfcc_ncu.md), not a tensor-core shape.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cstring>

#include "mfcc.cuh"

namespace var {

constexpr int kMel = 40;
constexpr int kFramesPerCta = 52;
constexpr int kMfccThreads = 256;
constexpr float kLogEps = 1e-6f;
constexpr int kLmLd = 44;  // log-mel row pitch in shared memory ([frame][filter], 16-byte aligned rows)

struct MfccTables {       // device pointers
  const float* window;    // [n_fft] window centred/padded
  const float2* tw;       // [n_fft] exp(-2 pi i q / n_fft), q in [0, n_fft)
  const int* fstart;      // [40] first bin of filter
  const int* fcount;      // [40] number of bins
  const int* foff;        // [40] offset into fweights
  const float* fweights;  // [nnz]
  const float* dct;       // [40 f][40 k]
  const float* lifter;    // [40] cepstral lifter (flavour 1)
  const float2* tw2;      // [7][8]   pass-2 twiddles exp(-2 pi i t k / 64), t = 1..7
  const float2* tw3;      // [R3-1][64] pass-3 twiddles exp(-2 pi i t j / N)
  // mel filters as aligned 4-bin groups: lane l < 32 owns filter l (groups lo_base4[l] .. + lo_n4[l], weights
  // w4lo[i][l]); filters 32..39 are split over 4 lanes each (lane l: filter 32 + l/4, groups hi_base4[l] + 4 q,
  // q < hi_nq[l], weights w4hi[q][l]); bins outside a filter carry weight 0
  const int* lo_base4; const int* lo_n4; const int* hi_base4; const int* hi_nq;
  const float4* w4lo; const float4* w4hi;
  int nnz, maxcnt, maxn4lo, maxqhi;
};

struct MfccPlan {
  int flavour;            // 0 torchaudio, 1 python_speech_features
  int fs, n_fft, win_length, hop;
  MfccTables t;
  void* dev_blob;
  int max_filter_bins;
};

struct MfccArgs {
  const int16_t* wav;
  const long long* offsets;  // [B] sample offset of clip b in wav, < 0 => all-zero feature
  const int* lengths;        // [B] samples
  int B, F, hop;
  int flavour, win_length;
  float* out;                // [B, F, 40]
  MfccTables t;
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// In-place radix-R DFT (forward, e^{-i...}) on v[0..R)
__device__ __forceinline__ void dft2(float2& a, float2& b) {
  const float2 t = a;
  a = make_float2(t.x + b.x, t.y + b.y);
  b = make_float2(t.x - b.x, t.y - b.y);
}
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)
__device__ __forceinline__ void dft4(float2* v) {
  // outputs in natural order
  dft2(v[0], v[2]);
  dft2(v[1], v[3]);
  v[3] = mul_mi(v[3]);
  dft2(v[0], v[1]);
  dft2(v[2], v[3]);
  const float2 t = v[1];
  v[1] = v[2];
  v[2] = t;
}
__device__ __forceinline__ void dft8(float2* v) {
  const float h = 0.70710678118654752440f;
  dft2(v[0], v[4]); dft2(v[1], v[5]); dft2(v[2], v[6]); dft2(v[3], v[7]);
  v[5] = make_float2((v[5].x + v[5].y) * h, (v[5].y - v[5].x) * h);   // * e^{-i pi/4}
  v[6] = mul_mi(v[6]);
  v[7] = make_float2((v[7].y - v[7].x) * h, -(v[7].x + v[7].y) * h);  // * e^{-3i pi/4}
  dft2(v[0], v[2]); dft2(v[1], v[3]);
  v[3] = mul_mi(v[3]);
  dft2(v[4], v[6]); dft2(v[5], v[7]);
  v[7] = mul_mi(v[7]);
  dft2(v[0], v[1]); dft2(v[2], v[3]); dft2(v[4], v[5]); dft2(v[6], v[7]);
  // bit-reversed -> natural
  float2 t;
  t = v[1]; v[1] = v[4]; v[4] = t;
  t = v[3]; v[3] = v[6]; v[6] = t;
}

// Per-warp exchange buffer of the three FFT passes: complex values as float2, every access 64 or 128 bits wide
// and bank-conflict free.
//   X1 (pass 1 -> 2): rows of 8 values (64 B); row r starts at 64 r + 16 (r >> 1).  Pass 1 writes row j with four
//       128-bit stores (8 consecutive rows of a quarter warp land on 8 different 16-byte slots mod 128); pass 2
//       reads column j & 7 of rows (j >> 3) + (N/64) t: a half warp covers an even/odd row pair = 128 contiguous B.
//   X2 (pass 2 -> 3): value i at 8 i + 64 (i >> 6); pass 2 writes i = 64 (j >> 3) + (j & 7) + 8 t, pass 3 reads
//       i = j + 64 t.
// Both images take 9 N bytes; the power spectrum (N + 1 floats) later reuses the same bytes.
template <int NFFT, bool PSF>
__global__ void __launch_bounds__(kMfccThreads, NFFT == 512 ? 4 : 3)  // 32 / 24 warps per SM measured best
mfcc_kernel(const __grid_constant__ MfccArgs a) {
  constexpr int N = NFFT / 2;        // complex FFT length
  constexpr int BPL = N / 8 / 32;    // radix-8 butterflies per lane per pass (1 or 2)
  constexpr int R3 = N / 64;         // radix of the last pass (4 or 8)
  constexpr int XB = 9 * N;          // exchange bytes per warp
  constexpr int kWarps = kMfccThreads / 32;

  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kFramesPerCta;
  const int F = a.F;
  if (f0 >= F) return;
  const int f1 = min(f0 + kFramesPerCta, F);
  const int nfr = f1 - f0;
  float* outp = a.out + ((long long)b * F + f0) * kMel;
  const long long off = a.offsets[b];
  const int S = a.lengths[b];
  // valid frames: torchaudio centre-pads (1 + S/hop); python_speech_features frames from sample 0
  // and zero-pads the tail (1 + ceil((S - win)/hop))
  constexpr bool psf = PSF;
  const int T = off < 0 ? 0
                        : (!psf ? 1 + S / a.hop
                                : (S <= a.win_length ? 1 : 1 + (S - a.win_length + a.hop - 1) / a.hop));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int nvalid = max(0, min(nfr, T - f0));
  if (nvalid < nfr) {  // zero-fill padded frames (processSoundFeat): rows are 160 B, the tail is 16-byte aligned
    float4* z4 = reinterpret_cast<float4*>(outp + nvalid * kMel);
    const int n4 = (nfr - nvalid) * (kMel / 4);
    for (int i = tid; i < n4; i += kMfccThreads) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (nvalid == 0) return;

  // ---- shared layout (tables + per-warp exchange buffers; samples are read straight from global memory)
  size_t o = 0;
  float* s_win = reinterpret_cast<float*>(sm_raw + o); o += NFFT * 4;
  float2* s_tw = reinterpret_cast<float2*>(sm_raw + o); o += N * 8;
  float4* s_w4lo = reinterpret_cast<float4*>(sm_raw + o); o += (size_t)a.t.maxn4lo * 32 * 16;
  float4* s_w4hi = reinterpret_cast<float4*>(sm_raw + o); o += (size_t)a.t.maxqhi * 32 * 16;
  float2* s_tw2 = reinterpret_cast<float2*>(sm_raw + o); o += 7 * 8 * 8;
  float2* s_tw3 = reinterpret_cast<float2*>(sm_raw + o); o += (R3 - 1) * 64 * 8;
  float* s_lm = reinterpret_cast<float*>(sm_raw + o); o += kLmLd * (kFramesPerCta + 4) * 4;  // [frame][f]
  float* s_le = reinterpret_cast<float*>(sm_raw + o); o += (kFramesPerCta + 4) * 4;          // log frame energy
  uint8_t* s_x = sm_raw + o;  // kWarps exchange buffers; the DCT matrix is staged here once the FFTs are done
  uint8_t* xb = s_x + warp * XB;

  {  // ---- tables
    for (int i = tid; i < NFFT; i += kMfccThreads) s_win[i] = a.t.window[i];
    for (int i = tid; i < N; i += kMfccThreads) s_tw[i] = a.t.tw[i];
    for (int i = tid; i < a.t.maxn4lo * 32; i += kMfccThreads) s_w4lo[i] = a.t.w4lo[i];
    for (int i = tid; i < a.t.maxqhi * 32; i += kMfccThreads) s_w4hi[i] = a.t.w4hi[i];
    for (int i = tid; i < 7 * 8; i += kMfccThreads) s_tw2[i] = a.t.tw2[i];
    for (int i = tid; i < (R3 - 1) * 64; i += kMfccThreads) s_tw3[i] = a.t.tw3[i];
  }
  __syncthreads();

  // mel filters as 4-bin groups: lane f owns filter f (< 32); the 8 widest filters (32..39) are split over
  // 4 lanes each (lane handles groups part, part + 4, ...)
  const int lo_base = a.t.lo_base4[lane], lo_n4 = a.t.lo_n4[lane];
  const int hi_base = a.t.hi_base4[lane], hi_nq = a.t.hi_nq[lane];
  const int maxn4lo = a.t.maxn4lo, maxqhi = a.t.maxqhi;
  const int fhi = 32 + (lane >> 2), part = lane & 3;
  const int j0 = lane, j1 = lane ? 64 - lane : 32;  // last-pass butterflies of this lane: spectrum pairs stay in-lane

  for (int fr = warp; fr < nvalid; fr += kWarps) {
    // frame element e is clip sample g0 + e (flavour 0: centred frames with reflect padding at the clip ends;
    // flavour 1: frames start at sample f * hop, zero padded behind the clip).  Interior frames read sample
    // pairs as 32-bit words (clip offsets and hops are even); frames touching a clip end take the scalar path.
    const int g0 = (f0 + fr) * a.hop - (psf ? 0 : NFFT / 2);
    const int16_t* wclip = a.wav + off;
    const bool interior = !(off & 1) && g0 >= (psf ? 2 : 0) && g0 + NFFT <= S;
    const uint32_t* xw = reinterpret_cast<const uint32_t*>(wclip + g0);
    const int lim = S - g0;  // clip samples left from the frame start (flavour 1)
    // ---------------- pass 1: radix 8, Ns = 1 (reads windowed samples) -------------
#pragma unroll
    for (int h = 0; h < BPL; ++h) {
      const int j = lane + 32 * h;
      float2 v[8];
      if (interior) {
        uint32_t cur[8], prev[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          cur[t] = __ldg(xw + j + (N / 8) * t);
          if (psf) prev[t] = __ldg(xw + j + (N / 8) * t - 1);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int e = 2 * (j + (N / 8) * t);
          const float2 w2 = *reinterpret_cast<const float2*>(s_win + e);  // flavour 0: window / 32768
          const float x0 = (float)(int16_t)(cur[t] & 0xffffu), x1 = (float)(int16_t)(cur[t] >> 16);
          if (!psf) {
            v[t] = make_float2(x0 * w2.x, x1 * w2.y);
          } else {  // pre-emphasis 0.97 on the raw int16 scale (sigproc.preemphasis)
            const float xm1 = (float)(int16_t)(prev[t] >> 16);
            v[t] = make_float2((x0 - 0.97f * xm1) * w2.x, (x1 - 0.97f * x0) * w2.y);
          }
        }
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int e = 2 * (j + (N / 8) * t);
          const float2 w2 = *reinterpret_cast<const float2*>(s_win + e);
          float xs[3];  // samples e - 1, e, e + 1
#pragma unroll
          for (int q = 0; q < 3; ++q) {
            int idx = g0 + e - 1 + q;
            if (!psf) {  // reflect padding at the clip ends
              if (idx < 0) idx = -idx;
              if (idx >= S) idx = 2 * (S - 1) - idx;
            }
            xs[q] = (idx >= 0 && idx < S) ? (float)wclip[idx] : 0.f;
          }
          if (!psf) {
            v[t] = make_float2(xs[1] * w2.x, xs[2] * w2.y);
          } else {  // the zero padding of the last frame is appended AFTER the pre-emphasis: samples >= S are 0
            const float y0 = e < lim ? xs[1] - 0.97f * xs[0] : 0.f;
            const float y1 = e + 1 < lim ? xs[2] - 0.97f * xs[1] : 0.f;
            v[t] = make_float2(y0 * w2.x, y1 * w2.y);
          }
        }
      }
      dft8(v);
      uint8_t* row = xb + 64 * j + 16 * (j >> 1);
#pragma unroll
      for (int t = 0; t < 4; ++t)
        *reinterpret_cast<float4*>(row + 16 * t) = make_float4(v[2 * t].x, v[2 * t].y, v[2 * t + 1].x, v[2 * t + 1].y);
    }
    __syncwarp();
    // ---------------- pass 2: radix 8, Ns = 8 ----------------
    {
      float2 v2[BPL][8];
#pragma unroll
      for (int h = 0; h < BPL; ++h) {
        const int j = lane + 32 * h;
        const int k = j & 7;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int r = (j >> 3) + (N / 64) * t;
          float2 u = *reinterpret_cast<const float2*>(xb + 64 * r + 16 * (r >> 1) + 8 * k);
          if (t) u = cmul(u, s_tw2[(t - 1) * 8 + k]);
          v2[h][t] = u;
        }
        dft8(v2[h]);
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < BPL; ++h) {
        const int j = lane + 32 * h;
        uint8_t* dst = xb + 576 * (j >> 3) + 8 * (j & 7);
#pragma unroll
        for (int t = 0; t < 8; ++t) *reinterpret_cast<float2*>(dst + 64 * t) = v2[h][t];
      }
    }
    __syncwarp();
    // ---------------- pass 3: radix N/64 (4 or 8), Ns = 64: zA = Z[j0 + 64 t], zB = Z[j1 + 64 t] ----------------
    float2 zA[R3], zB[R3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = h ? j1 : j0;
      float2 v[R3];
#pragma unroll
      for (int t = 0; t < R3; ++t) {
        float2 u = *reinterpret_cast<const float2*>(xb + 8 * j + 576 * t);
        if (t) u = cmul(u, s_tw3[(t - 1) * 64 + j]);
        v[t] = u;
      }
      if constexpr (R3 == 4) dft4(v); else dft8(v);
#pragma unroll
      for (int t = 0; t < R3; ++t) { if (h) zB[t] = v[t]; else zA[t] = v[t]; }
    }
    __syncwarp();
    // ---------------- real-FFT split, power spectrum ----------------
    // X[m] = E + W^m O and X[N - m] = conj(E - W^m O) come from the same pair (Z[m], Z[N - m]); with the lane's
    // two butterfly columns j0 and 64 - j0 both members of every pair are already in this lane's registers.
    float* spw = reinterpret_cast<float*>(xb);
    {
      const float pscale = psf ? 1.f / (float)NFFT : 1.f;
      const float pq = 0.25f * pscale;
      float esum = 0.f;
#pragma unroll
      for (int s = 0; s < R3; ++s) {
        // lanes >= 1: (Z[lane + 64 s], Z[(64 - lane) + 64 (R3 - 1 - s)]).  Lane 0 holds columns 0 and 32, which pair
        // with themselves: slots 0 .. R3/2-1 = (zA[s + 1], zA[R3 - 1 - s]), the rest = (zB[u], zB[R3 - 1 - u]).
        constexpr int HALF = R3 / 2;
        const float2 a0 = s < HALF ? zA[(s + 1) % R3] : zB[(s - HALF + R3) % R3];
        const float2 b0 = s < HALF ? zA[(R3 - 1 - s) % R3] : zB[(2 * R3 - 1 - (s - HALF)) % R3];
        const int m0 = s < HALF ? 64 * (s + 1) : 32 + 64 * (s - HALF);
        const bool l0 = lane == 0;
        const float2 za = l0 ? a0 : zA[s];
        const float2 zb = l0 ? b0 : zB[R3 - 1 - s];
        const int m = l0 ? m0 : lane + 64 * s;
        const float sx = za.x + zb.x, sy = za.y - zb.y;  // 2 E
        const float ox = za.y + zb.y, oy = zb.x - za.x;  // 2 O
        const float2 w = s_tw[m];                        // exp(-2 pi i m / NFFT)
        const float tx = ox * w.x - oy * w.y, ty = ox * w.y + oy * w.x;
        const float ar = sx + tx, ai = sy + ty, br = sx - tx, bi = sy - ty;
        const float pa = (ar * ar + ai * ai) * pq;
        const float pb = (br * br + bi * bi) * pq;
        spw[m] = pa;
        spw[N - m] = pb;
        esum += pa + ((l0 && s == HALF - 1) ? 0.f : pb);  // lane 0, m = N/2 pairs with itself
      }
      if (lane == 0) {
        const float dc = zA[0].x + zA[0].y, ny = zA[0].x - zA[0].y;
        spw[0] = dc * dc * pscale;
        spw[N] = ny * ny * pscale;
        esum += dc * dc * pscale + ny * ny * pscale;
      }
      if (psf) {  // frame energy -> c0 (base.mfcc appendEnergy=True)
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) esum += __shfl_xor_sync(0xffffffffu, esum, sft);
        if (lane == 0) s_le[fr] = logf(esum == 0.f ? 2.220446049250313e-16f : esum);
      }
    }
    __syncwarp();
    // ---------------- mel filterbank + log ----------------
    {
      const float4* spw4 = reinterpret_cast<const float4*>(xb);
      float acc = 0.f;
      for (int i = 0; i < maxn4lo; ++i)
        if (i < lo_n4) {
          const float4 p = spw4[lo_base + i], w = s_w4lo[i * 32 + lane];
          acc = fmaf(p.x, w.x, acc); acc = fmaf(p.y, w.y, acc); acc = fmaf(p.z, w.z, acc); acc = fmaf(p.w, w.w, acc);
        }
      s_lm[fr * kLmLd + lane] =
          psf ? logf(acc == 0.f ? 2.220446049250313e-16f : acc) : logf(acc + kLogEps);
      float acc1 = 0.f;
      for (int q = 0; q < maxqhi; ++q)
        if (q < hi_nq) {
          const float4 p = spw4[hi_base + 4 * q], w = s_w4hi[q * 32 + lane];
          acc1 = fmaf(p.x, w.x, acc1); acc1 = fmaf(p.y, w.y, acc1); acc1 = fmaf(p.z, w.z, acc1); acc1 = fmaf(p.w, w.w, acc1);
        }
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
      if (part == 0)
        s_lm[fr * kLmLd + fhi] =
            psf ? logf(acc1 == 0.f ? 2.220446049250313e-16f : acc1) : logf(acc1 + kLogEps);
    }
    __syncwarp();
  }
  __syncthreads();
  float* s_dct = reinterpret_cast<float*>(s_x);
  for (int i = tid; i < kMel * kMel; i += kMfccThreads) s_dct[i] = a.t.dct[i];
  __syncthreads();

  // ---- DCT: out[frame][k] = sum_f lm[frame][f] * dct[f][k]; thread = (4 outputs k, 4 frames): 8 128-bit shared
  // loads per 64 FMAs; rows of lm beyond nvalid hold stale but finite values and are not stored
  {
    const int ngroups = (nvalid + 3) / 4;
    for (int item = tid; item < ngroups * (kMel / 4); item += kMfccThreads) {
      const int kk = item % (kMel / 4), g = item / (kMel / 4);
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
      const float* lrow = s_lm + 4 * g * kLmLd;
#pragma unroll 2
      for (int f4 = 0; f4 < kMel / 4; ++f4) {
        float4 l[4], d[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) l[i] = *reinterpret_cast<const float4*>(lrow + i * kLmLd + 4 * f4);
#pragma unroll
        for (int c = 0; c < 4; ++c) d[c] = *reinterpret_cast<const float4*>(s_dct + (4 * f4 + c) * kMel + 4 * kk);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float lv[4] = {l[i].x, l[i].y, l[i].z, l[i].w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            acc[i][0] = fmaf(lv[c], d[c].x, acc[i][0]); acc[i][1] = fmaf(lv[c], d[c].y, acc[i][1]);
            acc[i][2] = fmaf(lv[c], d[c].z, acc[i][2]); acc[i][3] = fmaf(lv[c], d[c].w, acc[i][3]);
          }
        }
      }
      float4 lf = make_float4(1.f, 1.f, 1.f, 1.f);
      if (psf) lf = *reinterpret_cast<const float4*>(a.t.lifter + 4 * kk);  // lifter, then c0 := log frame energy
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int fr = 4 * g + i;
        if (fr < nvalid) {
          float4 r = make_float4(acc[i][0] * lf.x, acc[i][1] * lf.y, acc[i][2] * lf.z, acc[i][3] * lf.w);
          if (psf && kk == 0) r.x = s_le[fr];
          *reinterpret_cast<float4*>(outp + fr * kMel + 4 * kk) = r;
        }
      }
    }
  }
}

template <int NFFT>
static size_t mfcc_smem_bytes(int hop, int maxn4lo, int maxqhi) {
  constexpr int N = NFFT / 2;
  (void)hop;
  size_t o = 0;
  o += NFFT * 4 + N * 8 + (size_t)(maxn4lo + maxqhi) * 32 * 16 + 7 * 8 * 8 + (N / 64 - 1) * 64 * 8 +
       kLmLd * (kFramesPerCta + 4) * 4;
  o += (kFramesPerCta + 4) * 4;
  size_t x = (size_t)(kMfccThreads / 32) * 9 * N;
  if (x < (size_t)kMel * kMel * 4) x = (size_t)kMel * kMel * 4;
  return o + x + 16;
}

// ---------------------------------------------------------------------------
// Host: table construction (float32 arithmetic mirrors torchaudio's)
// ---------------------------------------------------------------------------
static void linspace_f32(float start, float end, int steps, std::vector<float>& out) {
  out.resize(steps);
  const float step = (end - start) / (float)(steps - 1);
  const int half = steps / 2;
  for (int i = 0; i < steps; ++i)
    out[i] = i < half ? start + step * (float)i : end - step * (float)(steps - 1 - i);
}

int mfcc_num_frames(const MfccPlan* p, int n_samples) {
  if (p->flavour == 0) return 1 + n_samples / p->hop;
  if (n_samples <= p->win_length) return 1;
  return 1 + (n_samples - p->win_length + p->hop - 1) / p->hop;
}

int mfcc_plan_create(int flavour, int fs, int n_fft, int win_length, int hop, MfccPlan** out) {
  if (flavour != 0 && flavour != 1) return VAR_ERR_UNSUPPORTED;
  if ((n_fft != 512 && n_fft != 1024) || win_length > n_fft || hop <= 0 || (hop & 1)) return VAR_ERR_UNSUPPORTED;
  const int nfreq = n_fft / 2 + 1;
  std::vector<float> window(n_fft, 0.f);
  if (flavour == 0) {
    // torch.hamming_window (periodic), centred in the n_fft frame by torch.stft
    const int left = (n_fft - win_length) / 2;
    for (int k = 0; k < win_length; ++k)
      window[left + k] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * (double)k / (double)win_length)) *
                         (1.f / 32768.f);  // int16 -> [-1, 1) scale folded in (a power of two: exact)
  } else {
    // np.hamming (symmetric) on the first win_length samples; the rFFT zero-pads the tail
    for (int k = 0; k < win_length; ++k)
      window[k] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * (double)k / (double)(win_length - 1)));
  }
  std::vector<float2> tw(n_fft);
  for (int q = 0; q < n_fft; ++q) {
    const double ang = -2.0 * M_PI * (double)q / (double)n_fft;
    tw[q] = make_float2((float)cos(ang), (float)sin(ang));
  }
  std::vector<int> fstart(kMel), fcount(kMel), foff(kMel);
  std::vector<float> fw;
  int maxbins = 0;
  if (flavour == 0) {
    // melscale_fbanks(n_freqs, 0, fs/2, 40, fs, norm=None, 'htk'), float32 like torch
    std::vector<float> all_freqs, m_pts;
    linspace_f32(0.f, (float)(fs / 2), nfreq, all_freqs);
    const float m_min = 2595.0f * log10f(1.0f + 0.f / 700.0f);
    const float m_max = (float)(2595.0 * log10(1.0 + (double)(fs / 2) / 700.0));
    linspace_f32(m_min, m_max, kMel + 2, m_pts);
    std::vector<float> f_pts(kMel + 2);
    for (int i = 0; i < kMel + 2; ++i) f_pts[i] = 700.0f * (powf(10.0f, m_pts[i] / 2595.0f) - 1.0f);
    for (int f = 0; f < kMel; ++f) {
      int first = -1, last = -1;
      std::vector<float> col(nfreq);
      for (int m = 0; m < nfreq; ++m) {
        const float down = -(f_pts[f] - all_freqs[m]) / (f_pts[f + 1] - f_pts[f]);
        const float up = (f_pts[f + 2] - all_freqs[m]) / (f_pts[f + 2] - f_pts[f + 1]);
        const float v = fmaxf(0.f, fminf(down, up));
        col[m] = v;
        if (v > 0.f) { if (first < 0) first = m; last = m; }
      }
      if (first < 0) { first = 0; last = -1; }
      fstart[f] = first; fcount[f] = last - first + 1; foff[f] = (int)fw.size();
      for (int m = first; m <= last; ++m) fw.push_back(col[m]);
      if (fcount[f] > maxbins) maxbins = fcount[f];
    }
  } else {
    // python_speech_features.base.get_filterbanks: triangles on FFT-bin indices (float64)
    const double lowmel = 2595.0 * log10(1.0 + 0.0 / 700.0);
    const double highmel = 2595.0 * log10(1.0 + (fs / 2.0) / 700.0);
    std::vector<double> bins(kMel + 2);
    for (int i = 0; i < kMel + 2; ++i) {
      const double mel = lowmel + (highmel - lowmel) * (double)i / (double)(kMel + 1);
      bins[i] = floor((n_fft + 1) * (700.0 * (pow(10.0, mel / 2595.0) - 1.0)) / fs);
    }
    for (int f = 0; f < kMel; ++f) {
      std::vector<double> col(nfreq, 0.0);
      for (int i = (int)bins[f]; i < (int)bins[f + 1]; ++i) col[i] = (i - bins[f]) / (bins[f + 1] - bins[f]);
      for (int i = (int)bins[f + 1]; i < (int)bins[f + 2]; ++i)
        col[i] = (bins[f + 2] - i) / (bins[f + 2] - bins[f + 1]);
      int first = -1, last = -1;
      for (int m = 0; m < nfreq; ++m)
        if (col[m] != 0.0) { if (first < 0) first = m; last = m; }
      if (first < 0) { first = 0; last = -1; }
      fstart[f] = first; fcount[f] = last - first + 1; foff[f] = (int)fw.size();
      for (int m = first; m <= last; ++m) fw.push_back((float)col[m]);
      if (fcount[f] > maxbins) maxbins = fcount[f];
    }
  }
  // create_dct(40, 40, 'ortho') -> [n_mels f][n_mfcc k]
  std::vector<float> dct(kMel * kMel);
  for (int k = 0; k < kMel; ++k)
    for (int f = 0; f < kMel; ++f) {
      float v = cosf((float)(M_PI / kMel) * ((float)f + 0.5f) * (float)k);
      if (k == 0) v *= (float)(1.0 / sqrt(2.0));
      v *= (float)sqrt(2.0 / kMel);
      dct[f * kMel + k] = v;
    }
  // compact, conflict-free twiddle tables of FFT passes 2 and 3, transposed filter weights
  const int Nc = n_fft / 2, R3 = Nc / 64;
  std::vector<float2> tw2(7 * 8), tw3((size_t)7 * 64, make_float2(1.f, 0.f));
  for (int t = 1; t < 8; ++t)
    for (int k = 0; k < 8; ++k) {
      const double ang = -2.0 * M_PI * (double)(t * k) / 64.0;
      tw2[(t - 1) * 8 + k] = make_float2((float)cos(ang), (float)sin(ang));
    }
  for (int t = 1; t < R3; ++t)
    for (int j = 0; j < 64; ++j) {
      const double ang = -2.0 * M_PI * (double)(t * j) / (double)Nc;
      tw3[(t - 1) * 64 + j] = make_float2((float)cos(ang), (float)sin(ang));
    }
  if (maxbins < 1) maxbins = 1;
  auto weight = [&](int f, int bin) { return bin >= fstart[f] && bin < fstart[f] + fcount[f] ? fw[foff[f] + bin - fstart[f]] : 0.f; };
  auto n4_of = [&](int f) { return fcount[f] ? (fstart[f] + fcount[f] + 3) / 4 - fstart[f] / 4 : 0; };
  std::vector<int> lo_base4(32), lo_n4(32), hi_base4(32), hi_nq(32);
  int maxn4lo = 1, maxqhi = 1;
  for (int l = 0; l < 32; ++l) {
    lo_base4[l] = fstart[l] / 4; lo_n4[l] = n4_of(l);
    if (lo_n4[l] > maxn4lo) maxn4lo = lo_n4[l];
    const int f = 32 + (l >> 2), part = l & 3, n4 = n4_of(f);
    hi_base4[l] = fstart[f] / 4 + part; hi_nq[l] = n4 > part ? (n4 - part + 3) / 4 : 0;
    if (hi_nq[l] > maxqhi) maxqhi = hi_nq[l];
  }
  std::vector<float> w4lo((size_t)maxn4lo * 32 * 4, 0.f), w4hi((size_t)maxqhi * 32 * 4, 0.f);
  for (int l = 0; l < 32; ++l) {
    for (int i = 0; i < lo_n4[l]; ++i)
      for (int c = 0; c < 4; ++c) w4lo[((size_t)i * 32 + l) * 4 + c] = weight(l, 4 * (lo_base4[l] + i) + c);
    const int f = 32 + (l >> 2);
    for (int q = 0; q < hi_nq[l]; ++q)
      for (int c = 0; c < 4; ++c) w4hi[((size_t)q * 32 + l) * 4 + c] = weight(f, 4 * (hi_base4[l] + 4 * q) + c);
  }
  std::vector<float> lifter(kMel, 1.f);
  if (flavour == 1)
    for (int n = 0; n < kMel; ++n) lifter[n] = (float)(1.0 + 11.0 * sin(M_PI * (double)n / 22.0));
  // pack into one device blob
  size_t o_win = 0, o_tw = o_win + window.size() * 4, o_fs = o_tw + tw.size() * 8,
         o_fc = o_fs + kMel * 4, o_fo = o_fc + kMel * 4, o_fw = o_fo + kMel * 4,
         o_dct = o_fw + ((fw.size() + 3) & ~(size_t)3) * 4, o_lf = o_dct + dct.size() * 4,
         o_tw2 = o_lf + kMel * 4, o_tw3 = o_tw2 + tw2.size() * 8, o_lb = o_tw3 + tw3.size() * 8,
         o_ln = o_lb + 32 * 4, o_hb = o_ln + 32 * 4, o_hn = o_hb + 32 * 4, o_w4lo = o_hn + 32 * 4,
         o_w4hi = o_w4lo + w4lo.size() * 4, total = o_w4hi + w4hi.size() * 4;
  std::vector<uint8_t> blob(total, 0);
  memcpy(&blob[o_win], window.data(), window.size() * 4);
  memcpy(&blob[o_tw], tw.data(), tw.size() * 8);
  memcpy(&blob[o_fs], fstart.data(), kMel * 4);
  memcpy(&blob[o_fc], fcount.data(), kMel * 4);
  memcpy(&blob[o_fo], foff.data(), kMel * 4);
  memcpy(&blob[o_fw], fw.data(), fw.size() * 4);
  memcpy(&blob[o_dct], dct.data(), dct.size() * 4);
  memcpy(&blob[o_lf], lifter.data(), kMel * 4);
  memcpy(&blob[o_tw2], tw2.data(), tw2.size() * 8);
  memcpy(&blob[o_tw3], tw3.data(), tw3.size() * 8);
  memcpy(&blob[o_lb], lo_base4.data(), 32 * 4);
  memcpy(&blob[o_ln], lo_n4.data(), 32 * 4);
  memcpy(&blob[o_hb], hi_base4.data(), 32 * 4);
  memcpy(&blob[o_hn], hi_nq.data(), 32 * 4);
  memcpy(&blob[o_w4lo], w4lo.data(), w4lo.size() * 4);
  memcpy(&blob[o_w4hi], w4hi.data(), w4hi.size() * 4);
  uint8_t* dev = nullptr;
  VAR_CUDA_CHECK(cudaMalloc(&dev, total));
  VAR_CUDA_CHECK(cudaMemcpy(dev, blob.data(), total, cudaMemcpyHostToDevice));
  MfccPlan* p = new MfccPlan();
  p->flavour = flavour;
  p->fs = fs; p->n_fft = n_fft; p->win_length = win_length; p->hop = hop;
  p->dev_blob = dev; p->max_filter_bins = maxbins;
  p->t.window = reinterpret_cast<float*>(dev + o_win);
  p->t.tw = reinterpret_cast<float2*>(dev + o_tw);
  p->t.fstart = reinterpret_cast<int*>(dev + o_fs);
  p->t.fcount = reinterpret_cast<int*>(dev + o_fc);
  p->t.foff = reinterpret_cast<int*>(dev + o_fo);
  p->t.fweights = reinterpret_cast<float*>(dev + o_fw);
  p->t.dct = reinterpret_cast<float*>(dev + o_dct);
  p->t.lifter = reinterpret_cast<float*>(dev + o_lf);
  p->t.tw2 = reinterpret_cast<float2*>(dev + o_tw2);
  p->t.tw3 = reinterpret_cast<float2*>(dev + o_tw3);
  p->t.lo_base4 = reinterpret_cast<int*>(dev + o_lb);
  p->t.lo_n4 = reinterpret_cast<int*>(dev + o_ln);
  p->t.hi_base4 = reinterpret_cast<int*>(dev + o_hb);
  p->t.hi_nq = reinterpret_cast<int*>(dev + o_hn);
  p->t.w4lo = reinterpret_cast<float4*>(dev + o_w4lo);
  p->t.w4hi = reinterpret_cast<float4*>(dev + o_w4hi);
  p->t.nnz = (int)fw.size();
  p->t.maxcnt = maxbins;
  p->t.maxn4lo = maxn4lo;
  p->t.maxqhi = maxqhi;
  *out = p;
  return VAR_OK;
}

void mfcc_plan_destroy(MfccPlan* p) {
  if (!p) return;
  cudaFree(p->dev_blob);
  delete p;
}

int mfcc_fwd(const MfccPlan* p, const int16_t* wav, const long long* offsets, const int* lengths,
             int B, int F, float* out, cudaStream_t st) {
  if (B <= 0 || F <= 0) return VAR_OK;
  MfccArgs a;
  a.wav = wav; a.offsets = offsets; a.lengths = lengths; a.B = B; a.F = F; a.hop = p->hop;
  a.flavour = p->flavour; a.win_length = p->win_length;
  a.out = out; a.t = p->t;
  dim3 grid((F + kFramesPerCta - 1) / kFramesPerCta, B);
  LaunchScope sc(T_MFCC, 0, st);
  const bool psf = p->flavour == 1;
#define VAR_MFCC_LAUNCH(NF, PS)                                   \
  do {                                                            \
    VAR_ENSURE_SMEM((mfcc_kernel<NF, PS>), smem);                 \
    mfcc_kernel<NF, PS><<<grid, kMfccThreads, smem, st>>>(a);     \
  } while (0)
  if (p->n_fft == 512) {
    const size_t smem = mfcc_smem_bytes<512>(p->hop, p->t.maxn4lo, p->t.maxqhi);  // (table sizes differ per flavour)
    if (psf) VAR_MFCC_LAUNCH(512, true); else VAR_MFCC_LAUNCH(512, false);
  } else {
    const size_t smem = mfcc_smem_bytes<1024>(p->hop, p->t.maxn4lo, p->t.maxqhi);
    if (psf) VAR_MFCC_LAUNCH(1024, true); else VAR_MFCC_LAUNCH(1024, false);
  }
#undef VAR_MFCC_LAUNCH
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var
