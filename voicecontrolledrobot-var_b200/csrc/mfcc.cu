// Fused MFCC front-end: int16 wav -> reflect pad -> framing ->
// periodic-Hamming window (centred in n_fft) -> real FFT -> |.|^2 -> HTK mel
// filterbank -> log(x + 1e-6) -> ortho DCT-II -> [B, F, 40] crop / zero-pad.
// Replaces torchaudio.transforms.MFCC as configured at Envs/audioLoader.py:147-157
// plus processSoundFeat (Envs/audioLoader.py:241-252).
//
// One CTA per (clip, chunk of <= kFramesPerCta frames).  The chunk's samples are
// staged once in shared memory; each warp then runs a register/shared Stockham
// FFT per frame (n_fft/2-point complex FFT of the packed real frame, radix 8x8x4
// or 8x8x8, bank-conflict-free exchanges), the power spectrum is reduced to 40
// log-mel values, and the whole CTA finishes with a register-tiled DCT whose
// output rows are written coalesced.  Pure fp32 CUDA-core work: the transform is
// FP32-issue bound, not a tensor-core shape.
#include <cmath>
#include <cstdio>
#include <vector>

#include <cstring>

#include "mfcc.cuh"

namespace var {

constexpr int kMel = 40;
constexpr int kFramesPerCta = 52;
constexpr int kMfccThreads = 256;
constexpr float kLogEps = 1e-6f;

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
  const float* fwt;       // [maxcnt][40] filter weights, i-th weight of filter f at [i*40 + f]
  int nnz, maxcnt, maxcnt_lo;  // maxcnt_lo = longest of the first 32 filters
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

// exchange paddings (every shared access of the FFT is conflict free)
__device__ __forceinline__ int padA(int i) { return i + (i >> 5); }
__device__ __forceinline__ int padB(int i) { return i + ((i >> 6) << 3); }

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

template <int NFFT>
__global__ void __launch_bounds__(kMfccThreads)
mfcc_kernel(const __grid_constant__ MfccArgs a) {
  constexpr int N = NFFT / 2;        // complex FFT length
  constexpr int NB = N / 32;         // spectrum values per lane
  constexpr int BPL = N / 8 / 32;    // radix-8 butterflies per lane per pass (1 or 2)
  constexpr int NBIN = N + 1;
  constexpr int kWarps = kMfccThreads / 32;
  constexpr int SCR = N + (N >> 5) + 40;  // per-warp scratch floats per component (>= padB max)

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
  const bool psf = a.flavour == 1;
  const int T = off < 0 ? 0
                        : (!psf ? 1 + S / a.hop
                                : (S <= a.win_length ? 1 : 1 + (S - a.win_length + a.hop - 1) / a.hop));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  const int nvalid = max(0, min(nfr, T - f0));
  if (nvalid < nfr) {  // zero-fill padded frames (processSoundFeat)
    for (int i = nvalid * kMel + tid; i < nfr * kMel; i += kMfccThreads) outp[i] = 0.f;
  }
  if (nvalid == 0) return;

  // ---- shared layout
  const int span = (nvalid - 1) * a.hop + NFFT;  // samples needed by this chunk
  const int span2 = (span + 1) & ~1;
  int16_t* s_wav = reinterpret_cast<int16_t*>(sm_raw);
  size_t o = ((size_t)(span2 + 2) * 2 + 15) & ~(size_t)15;
  float* s_win = reinterpret_cast<float*>(sm_raw + o); o += NFFT * 4;
  float2* s_tw = reinterpret_cast<float2*>(sm_raw + o); o += NFFT * 8;
  float* s_fwt = reinterpret_cast<float*>(sm_raw + o); o += (size_t)a.t.maxcnt * kMel * 4;
  float2* s_tw2 = reinterpret_cast<float2*>(sm_raw + o); o += 7 * 8 * 8;
  float2* s_tw3 = reinterpret_cast<float2*>(sm_raw + o); o += 7 * 64 * 8;
  float* s_dct = reinterpret_cast<float*>(sm_raw + o); o += kMel * kMel * 4;
  float* s_lm = reinterpret_cast<float*>(sm_raw + o); o += kMel * (kFramesPerCta + 4) * 4;  // [f][frame]
  float* s_le = reinterpret_cast<float*>(sm_raw + o); o += (kFramesPerCta + 4) * 4;          // log frame energy
  float* s_scr = reinterpret_cast<float*>(sm_raw + o);  // per warp: re[SCR], im[SCR], pw[NBIN+..]
  constexpr int WSCR = 2 * SCR + ((NBIN + 7) & ~3);
  float* sre = s_scr + warp * WSCR;
  float* sim = sre + SCR;
  float* spw = sim + SCR;

  // ---- stage samples (reflect padding at the clip ends), tables
  {
    const int16_t* w = a.wav + off;
    // logical sample index of s_wav[0]; flavour 1 keeps one extra leading sample for the pre-emphasis
    const int base = psf ? f0 * a.hop - 2 : f0 * a.hop - NFFT / 2;
    // two samples per 32-bit access (clip offsets and hops are even); reflection / zero fill at
    // the clip ends goes through the scalar path
    const int npairs = (span2 + (psf ? 2 : 0) + 1) >> 1;
    for (int pi = tid; pi < npairs; pi += kMfccThreads) {
      const int i = pi * 2;
      const int idx0 = base + i;
      if (idx0 >= 0 && idx0 + 1 < S && !(off & 1)) {
        *reinterpret_cast<uint32_t*>(s_wav + i) = *reinterpret_cast<const uint32_t*>(w + idx0);
      } else {
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          int idx = idx0 + e2;
          if (!psf) {  // reflect padding at the clip ends
            if (idx < 0) idx = -idx;
            if (idx >= S) idx = 2 * (S - 1) - idx;
          }
          int16_t v = 0;
          if (idx >= 0 && idx < S) v = w[idx];
          s_wav[i + e2] = v;
        }
      }
    }
    for (int i = tid; i < NFFT; i += kMfccThreads) { s_win[i] = a.t.window[i]; s_tw[i] = a.t.tw[i]; }
    for (int i = tid; i < a.t.maxcnt * kMel; i += kMfccThreads) s_fwt[i] = a.t.fwt[i];
    for (int i = tid; i < 7 * 8; i += kMfccThreads) s_tw2[i] = a.t.tw2[i];
    for (int i = tid; i < (N / 64 - 1) * 64; i += kMfccThreads) s_tw3[i] = a.t.tw3[i];
    for (int i = tid; i < kMel * kMel; i += kMfccThreads) s_dct[i] = a.t.dct[i];
  }
  __syncthreads();

  // mel filters: lane f owns filter f (< 32); the 8 widest filters (32..39) are split over
  // 4 lanes each so that both loops have short, balanced trip counts
  const int fst0 = a.t.fstart[lane], fcn0 = a.t.fcount[lane];
  const int fhi = 32 + (lane >> 2), part = lane & 3;
  const int fst1 = a.t.fstart[fhi], fcn1 = a.t.fcount[fhi];
  const int cnt_lo = a.t.maxcnt_lo, cnt_hi = (a.t.maxcnt + 3) >> 2;

  for (int fr = warp; fr < nvalid; fr += kWarps) {
    const int16_t* x = s_wav + fr * a.hop + (psf ? 2 : 0);  // frame element e -> x[e]
    const int lim = S - (f0 + fr) * a.hop;                   // clip samples left from the frame start
    float2 z[NB];
    // ---------------- pass 1: radix 8, Ns = 1 (reads windowed samples) -------------
#pragma unroll
    for (int h = 0; h < BPL; ++h) {
      const int j = lane + 32 * h;
      float2 v[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int e = 2 * (j + (N / 8) * t);
        const short2 s2 = *reinterpret_cast<const short2*>(x + e);
        const float2 w2 = *reinterpret_cast<const float2*>(s_win + e);
        if (!psf) {
          v[t] = make_float2((float)s2.x * (1.f / 32768.f) * w2.x, (float)s2.y * (1.f / 32768.f) * w2.y);
        } else {  // pre-emphasis 0.97 on the raw int16 scale (sigproc.preemphasis); the zero
          // padding of the last frame is appended AFTER the pre-emphasis, so samples >= S are 0
          const float xm1 = (float)x[e - 1];
          const float y0 = e < lim ? (float)s2.x - 0.97f * xm1 : 0.f;
          const float y1 = e + 1 < lim ? (float)s2.y - 0.97f * (float)s2.x : 0.f;
          v[t] = make_float2(y0 * w2.x, y1 * w2.y);
        }
      }
      dft8(v);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int p = padA(8 * j + t);
        sre[p] = v[t].x;
        sim[p] = v[t].y;
      }
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
          const int p = padA(j + (N / 8) * t);
          float2 u = make_float2(sre[p], sim[p]);
          if (t) u = cmul(u, s_tw2[(t - 1) * 8 + k]);
          v2[h][t] = u;
        }
        dft8(v2[h]);
      }
      __syncwarp();
#pragma unroll
      for (int h = 0; h < BPL; ++h) {
        const int j = lane + 32 * h;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int p = padB(((j >> 3) << 6) + (j & 7) + 8 * t);
          sre[p] = v2[h][t].x;
          sim[p] = v2[h][t].y;
        }
      }
    }
    __syncwarp();
    // ---------------- pass 3: radix N/64 (4 or 8), Ns = 64 ----------------
    {
      constexpr int R3 = N / 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;  // k = j
        float2 v[R3];
#pragma unroll
        for (int t = 0; t < R3; ++t) {
          const int p = padB(j + 64 * t);
          float2 u = make_float2(sre[p], sim[p]);
          if (t) u = cmul(u, s_tw3[(t - 1) * 64 + j]);
          v[t] = u;
        }
        if constexpr (R3 == 4) dft4(v); else dft8(v);
#pragma unroll
        for (int t = 0; t < R3; ++t) z[h + 2 * t] = v[t];  // m = lane + 32*(h + 2t)
      }
    }
    __syncwarp();
    // ---------------- real-FFT split, power spectrum ----------------
    {
      const float pscale = psf ? 1.f / (float)NFFT : 1.f;
      float esum = 0.f;
      const int src = (32 - lane) & 31;
#pragma unroll
      for (int u = 0; u < NB; ++u) {
        // partner Z[N - m]: lane' = 32 - lane, u' = NB-1-u (lane != 0); lane 0: u' = NB - u
        float2 pz;
        pz.x = __shfl_sync(0xffffffffu, z[NB - 1 - u].x, src);
        pz.y = __shfl_sync(0xffffffffu, z[NB - 1 - u].y, src);
        if (lane == 0) pz = z[(NB - u) % NB];
        const float2 zm = z[u];
        const float er = 0.5f * (zm.x + pz.x), ei = 0.5f * (zm.y - pz.y);
        const float orr = 0.5f * (zm.y + pz.y), oi = -0.5f * (zm.x - pz.x);
        const float2 w = s_tw[lane + 32 * u];  // exp(-2 pi i m / NFFT)
        const float xr = er + (orr * w.x - oi * w.y);
        const float xi = ei + (orr * w.y + oi * w.x);
        const float pw = (xr * xr + xi * xi) * pscale;
        spw[lane + 32 * u] = pw;
        esum += pw;
      }
      if (lane == 0) {
        const float ny = z[0].x - z[0].y;
        spw[N] = ny * ny * pscale;
        esum += ny * ny * pscale;
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
      float acc = 0.f;
      for (int i = 0; i < cnt_lo; ++i)
        if (i < fcn0) acc = fmaf(spw[fst0 + i], s_fwt[i * kMel + lane], acc);
      s_lm[lane * (kFramesPerCta + 4) + fr] =
          psf ? logf(acc == 0.f ? 2.220446049250313e-16f : acc) : logf(acc + kLogEps);
      float acc1 = 0.f;
      for (int q4 = 0; q4 < cnt_hi; ++q4) {
        const int i = q4 * 4 + part;
        if (i < fcn1) acc1 = fmaf(spw[fst1 + i], s_fwt[i * kMel + fhi], acc1);
      }
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
      if (part == 0)
        s_lm[fhi * (kFramesPerCta + 4) + fr] =
            psf ? logf(acc1 == 0.f ? 2.220446049250313e-16f : acc1) : logf(acc1 + kLogEps);
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- DCT: out[frame][k] = sum_f lm[f][frame] * dct[f][k]; thread = (k, 4 frames)
  {
    constexpr int LD = kFramesPerCta + 4;
    const int ngroups = (nvalid + 3) / 4;
    for (int item = tid; item < ngroups * kMel; item += kMfccThreads) {
      const int k = item % kMel, g = item / kMel;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
      for (int f = 0; f < kMel; ++f) {
        const float4 l4 = *reinterpret_cast<const float4*>(s_lm + f * LD + 4 * g);
        const float d = s_dct[f * kMel + k];
        acc.x = fmaf(l4.x, d, acc.x); acc.y = fmaf(l4.y, d, acc.y);
        acc.z = fmaf(l4.z, d, acc.z); acc.w = fmaf(l4.w, d, acc.w);
      }
      const int fr = 4 * g;
      if (psf) {  // lifter, then c0 := log frame energy
        const float lf = a.t.lifter[k];
        acc.x *= lf; acc.y *= lf; acc.z *= lf; acc.w *= lf;
        if (k == 0) {
          acc.x = s_le[fr]; acc.y = s_le[fr + 1]; acc.z = s_le[fr + 2]; acc.w = s_le[fr + 3];
        }
      }
      if (fr < nvalid) outp[(fr)*kMel + k] = acc.x;
      if (fr + 1 < nvalid) outp[(fr + 1) * kMel + k] = acc.y;
      if (fr + 2 < nvalid) outp[(fr + 2) * kMel + k] = acc.z;
      if (fr + 3 < nvalid) outp[(fr + 3) * kMel + k] = acc.w;
    }
  }
}

template <int NFFT>
static size_t mfcc_smem_bytes(int hop, int maxcnt) {
  constexpr int N = NFFT / 2;
  constexpr int SCR = N + (N >> 5) + 40;
  constexpr int NBIN = N + 1;
  constexpr int WSCR = 2 * SCR + ((NBIN + 7) & ~3);
  size_t span = (size_t)(kFramesPerCta - 1) * hop + NFFT + 4;
  size_t o = (span * 2 + 15) & ~(size_t)15;
  o += NFFT * 4 + NFFT * 8 + (size_t)maxcnt * kMel * 4 + 7 * 8 * 8 + 7 * 64 * 8 + kMel * kMel * 4 +
       kMel * (kFramesPerCta + 4) * 4;
  o += (kFramesPerCta + 4) * 4;
  o += (size_t)(kMfccThreads / 32) * WSCR * 4;
  return o + 16;
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
      window[left + k] = (float)(0.54 - 0.46 * cos(2.0 * M_PI * (double)k / (double)win_length));
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
  int maxcnt_lo = 1;
  for (int f = 0; f < 32; ++f) if (fcount[f] > maxcnt_lo) maxcnt_lo = fcount[f];
  if (maxbins < 1) maxbins = 1;
  std::vector<float> fwt((size_t)maxbins * kMel, 0.f);
  for (int f = 0; f < kMel; ++f)
    for (int i = 0; i < fcount[f]; ++i) fwt[(size_t)i * kMel + f] = fw[foff[f] + i];
  std::vector<float> lifter(kMel, 1.f);
  if (flavour == 1)
    for (int n = 0; n < kMel; ++n) lifter[n] = (float)(1.0 + 11.0 * sin(M_PI * (double)n / 22.0));
  // pack into one device blob
  size_t o_win = 0, o_tw = o_win + window.size() * 4, o_fs = o_tw + tw.size() * 8,
         o_fc = o_fs + kMel * 4, o_fo = o_fc + kMel * 4, o_fw = o_fo + kMel * 4,
         o_dct = o_fw + ((fw.size() + 3) & ~(size_t)3) * 4, o_lf = o_dct + dct.size() * 4,
         o_tw2 = o_lf + kMel * 4, o_tw3 = o_tw2 + tw2.size() * 8, o_fwt = o_tw3 + tw3.size() * 8,
         total = o_fwt + fwt.size() * 4;
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
  memcpy(&blob[o_fwt], fwt.data(), fwt.size() * 4);
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
  p->t.fwt = reinterpret_cast<float*>(dev + o_fwt);
  p->t.nnz = (int)fw.size();
  p->t.maxcnt = maxbins;
  p->t.maxcnt_lo = maxcnt_lo;
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
  if (p->n_fft == 512) {
    const size_t smem = mfcc_smem_bytes<512>(p->hop, p->t.maxcnt);
    VAR_ENSURE_SMEM(mfcc_kernel<512>, smem);  // (maxcnt differs between the two filterbank flavours)
    mfcc_kernel<512><<<grid, kMfccThreads, smem, st>>>(a);
  } else {
    const size_t smem = mfcc_smem_bytes<1024>(p->hop, p->t.maxcnt);
    VAR_ENSURE_SMEM(mfcc_kernel<1024>, smem);
    mfcc_kernel<1024><<<grid, kMfccThreads, smem, st>>>(a);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var
