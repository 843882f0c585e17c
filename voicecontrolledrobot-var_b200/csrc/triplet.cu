// Triplet sampling and the fused warp-level head / triplet-margin kernel.
//
//  * var::sampler_* : bit-exact device restatement of the integer sampling the
//    reference performs on torch's CPU mt19937 stream: DataLoader(shuffle=True)
//    epoch permutation (torch.randperm from a RandomSampler-private generator),
//    the negative-class draw of dataset.py:72-78 and the clip draws of
//    Envs/audioLoader.py:166-177, in the per-item order of dataset.py:34-62.
//    The generator state lives in HBM and advances exactly like torch's.
//  * var::tail_* : one warp per triplet: last Linear of each head -> F.normalize
//    (models/pretext/pretext_base.py:18,23) -> TripletMarginLoss(margin, p=2)
//    (VAR/pretext_VAR.py:38,64) -> gradients back through the normalisation and
//    the last Linear (dW, db, and the ReLU-masked grad of its input), or for the
//    reward query: embeddings -> image . goal-sound dot product (+ env reward)
//    (Envs/vec_env/vec_pretext_normalize.py:96-101).
#include "triplet.cuh"

namespace var {

// ===========================================================================
// mt19937
// ===========================================================================
constexpr int MT_N = 624, MT_M = 397;

__device__ __forceinline__ uint32_t mt_mix(uint32_t a, uint32_t b) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7FFFFFFFu);
  return (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9D2C5680u;
  y ^= (y << 15) & 0xEFC60000u;
  y ^= y >> 18;
  return y;
}
// Whole-CTA twist of s[0..624) in shared memory (three dependent phases).
__device__ void mt_twist(uint32_t* s) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // phase 1: k in [0, 227): new[k] = old[k+397] ^ mix(old[k], old[k+1])
  uint32_t v[3];
  int cnt = 0;
  for (int k = tid; k < MT_N - MT_M; k += nt) v[cnt++] = s[k + MT_M] ^ mt_mix(s[k], s[k + 1]);
  __syncthreads();
  cnt = 0;
  for (int k = tid; k < MT_N - MT_M; k += nt) s[k] = v[cnt++];
  __syncthreads();
  // phase 2: k in [227, 454): new[k] = new[k-227] ^ mix(old[k], old[k+1])
  cnt = 0;
  for (int k = MT_N - MT_M + tid; k < 2 * (MT_N - MT_M); k += nt)
    v[cnt++] = s[k - (MT_N - MT_M)] ^ mt_mix(s[k], s[k + 1]);
  __syncthreads();
  cnt = 0;
  for (int k = MT_N - MT_M + tid; k < 2 * (MT_N - MT_M); k += nt) s[k] = v[cnt++];
  __syncthreads();
  // phase 3: k in [454, 624): new[k] = new[k-227] ^ mix(old[k], old[k+1]) (old[624] := new[0])
  cnt = 0;
  for (int k = 2 * (MT_N - MT_M) + tid; k < MT_N; k += nt)
    v[cnt++] = s[k - (MT_N - MT_M)] ^ mt_mix(s[k], k + 1 < MT_N ? s[k + 1] : s[0]);
  __syncthreads();
  cnt = 0;
  for (int k = 2 * (MT_N - MT_M) + tid; k < MT_N; k += nt) s[k] = v[cnt++];
  __syncthreads();
}
// init_genrand(seed) by thread 0 (sequential recurrence, 623 steps).
__device__ void mt_seed(uint32_t* s, uint32_t seed) {
  if (threadIdx.x == 0) {
    s[0] = seed;
    for (int i = 1; i < MT_N; ++i) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
  }
  __syncthreads();
}
// Fill out[0..count) with the next `count` tempered outputs starting at (s, pos);
// returns the new pos.  Whole CTA; out may be shared or global.
__device__ int mt_fill(uint32_t* s, int pos, uint32_t* out, int count) {
  int done = 0;
  while (done < count) {
    if (pos >= MT_N) {
      mt_twist(s);
      pos = 0;
    }
    const int take = min(MT_N - pos, count - done);
    for (int i = threadIdx.x; i < take; i += blockDim.x) out[done + i] = mt_temper(s[pos + i]);
    __syncthreads();
    pos += take;
    done += take;
  }
  return pos;
}
// Advance (s, pos) by `count` outputs without materialising them.
__device__ int mt_skip(uint32_t* s, int pos, long long count) {
  while (count > 0) {
    if (pos >= MT_N) {
      mt_twist(s);
      pos = 0;
    }
    const int take = (int)min((long long)(MT_N - pos), count);
    pos += take;
    count -= take;
  }
  return pos;
}

__global__ void sampler_seed_kernel(uint32_t* state, uint32_t seed) {
  __shared__ uint32_t s[MT_N];
  mt_seed(s, seed);
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) state[i] = s[i];
  if (threadIdx.x == 0) state[MT_N] = MT_N;  // pos: twist before the first draw
}

// DataLoader epoch start (oracle/sampler.py::epoch_batches): two random_() int64
// draws from the global generator (each = hi32, lo32; value mod 2^63), the
// second seeds a private mt19937 whose Fisher-Yates randperm(n) is the epoch
// order.  perm lives in global memory; thread 0 performs the dependent swaps.
__global__ void sampler_epoch_kernel(uint32_t* state, int n, int* perm) {
  extern __shared__ uint16_t s_perm[];  // [n] when n <= 65536 (the dependent swaps then stay on chip)
  __shared__ uint32_t s[MT_N];
  __shared__ uint32_t d[4];
  __shared__ uint32_t buf[MT_N];
  const bool in_smem = n <= 65536;
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) s[i] = state[i];
  __syncthreads();
  int pos = (int)state[MT_N];
  pos = mt_fill(s, pos, d, 4);
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) state[i] = s[i];
  if (threadIdx.x == 0) state[MT_N] = (uint32_t)pos;
  __syncthreads();
  // sampler seed = ((d[2] << 32) | d[3]) mod 2^63 ; at::mt19937 seeds with the low 32 bits
  const uint32_t seed = d[3];
  mt_seed(s, seed);
  if (in_smem) for (int i = threadIdx.x; i < n; i += blockDim.x) s_perm[i] = (uint16_t)i;
  else for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = i;
  __syncthreads();
  int ppos = MT_N;
  for (int base = 0; base < n - 1; base += MT_N) {
    const int cnt = min(MT_N, n - 1 - base);
    ppos = mt_fill(s, ppos, buf, cnt);
    // the variable-divisor modulo of every swap is independent of the swaps: all threads reduce
    // their draws first, the single dependent thread then only moves data
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) buf[i] = buf[i] % (uint32_t)(n - (base + i));
    __syncthreads();
    if (threadIdx.x == 0) {
      if (in_smem) {
        for (int i = 0; i < cnt; ++i) {
          const int ii = base + i;
          const int z = (int)buf[i];
          const uint16_t t = s_perm[ii];
          s_perm[ii] = s_perm[z + ii];
          s_perm[z + ii] = t;
        }
      } else {
        for (int i = 0; i < cnt; ++i) {
          const int ii = base + i;
          const int z = (int)buf[i];
          const int t = perm[ii];
          perm[ii] = perm[z + ii];
          perm[z + ii] = t;
        }
      }
    }
    __syncthreads();
  }
  if (in_smem) for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = (int)s_perm[i];
}

// One batch of triplets.  item i = items[i] (index into gt / stored_sn).
// Draw order per item (dataset.py:64-89, :34-62; audioLoader.py:171-177):
//   [sn draw unless stored]  then  positive sound unless gt == taskNum,
//   then negative sound unless sn == taskNum.
// A sound costs dps draws: 2 for the pybullet tables (dataset, clip; audioLoader.py:174-176) and
// 3 for the iTHOR task tables (location synonym, object synonym, clip; audioLoader.py:223-237,
// :208-209), where nds / nds2 hold the two synonym counts of a task and the clip list of the
// resolved (location, object, action) sits at [task * max_ds + li * max_ds2 + oi].
__global__ void sampler_batch_kernel(SamplerArgs a) {
  extern __shared__ uint32_t sm[];
  uint32_t* s = sm;            // [624]
  uint32_t* draws = sm + MT_N; // [(1 + 2 dps) * chunk]
  __shared__ int s_off_next;
  const int T = a.task_num;
  const int dps = a.nds2 ? 3 : 2;
  const int dpi = 1 + 2 * dps;  // most draws one item can consume
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) s[i] = a.state[i];
  int pos = (int)a.state[MT_N];  // uniform across the CTA
  __syncthreads();
  for (int c0 = 0; c0 < a.B; c0 += a.chunk) {
    const int nb = min(a.chunk, a.B - c0);
    // speculatively materialise the maximum this chunk can consume, from a scratch copy
    // of the state (the authoritative state is advanced by the consumed count below)
    uint32_t* s2 = draws + dpi * a.chunk;  // [624] scratch state
    for (int i = threadIdx.x; i < MT_N; i += blockDim.x) s2[i] = s[i];
    __syncthreads();
    mt_fill(s2, pos, draws, dpi * nb);
    // stage the chunk's labels in shared memory so the chain below never waits on HBM
    int16_t* s_lab = reinterpret_cast<int16_t*>(s2 + MT_N);  // [chunk] gt | (stored sn + 1) << 8
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      const int item = a.items ? a.items[c0 + i] : c0 + i;
      const int sn1 = a.stored_sn ? a.stored_sn[item] + 1 : 0;
      s_lab[i] = (int16_t)(a.gt[item] | (sn1 << 8));
    }
    __syncthreads();
    // sequential chain: where does each item's draw window start?  Everything the loop touches
    // is in shared memory and the modulo is a multiply-high (T is tiny), so one iteration is a
    // dependent LDS + ~10 ALU ops (~25 ns) instead of a global store and a 32-bit division.
    uint16_t* s_off = reinterpret_cast<uint16_t*>(s_lab + a.chunk);  // [chunk], offsets < dpi*chunk <= 40960
    if (threadIdx.x == 0) {
      const uint32_t Tm = (uint32_t)((0x100000000ull + (uint32_t)T - 1) / (uint32_t)T);
      int off = 0;
      for (int i = 0; i < nb; ++i) {
        s_off[i] = (uint16_t)off;
        const int lab = s_lab[i];
        const int gt = lab & 0xFF;
        int sn = (lab >> 8) - 1;
        int used = 0;
        if (sn < 0) {
          const uint32_t x = draws[off];
          int r = (int)(x - __umulhi(x, Tm) * (uint32_t)T);  // x % T, exact after one correction
          if (r < 0) r += T;
          if (T == 1) r = 0;
          sn = r == gt ? T : r;
          used = 1;
        }
        if (gt == T) used += dps;                    // negative only
        else used += dps + (sn == T ? 0 : dps);
        off += used;
      }
      s_off_next = off;
    }
    __syncthreads();
    // parallel decode
    for (int i = threadIdx.x; i < nb; i += blockDim.x) {
      int off = s_off[i];
      const int item = a.items ? a.items[c0 + i] : c0 + i;
      const int gt = a.gt[item];
      int sn;
      if (a.stored_sn) sn = a.stored_sn[item];
      else {
        sn = (int)(draws[off++] % (uint32_t)T);
        if (sn == gt) sn = T;
      }
      int rec[6] = {-1, -1, -1, -1, -1, -1};
      long long poff = -1, noff = -1;
      int plen = 0, nlen = 0;
      auto draw_clip = [&](int intent, int* r, long long* o, int* l) {
        if (intent > T - 1) intent = T - 1;
        const int nds = a.nds[intent];
        int ds = (int)(draws[off++] % (uint32_t)nds);
        if (a.nds2) ds = ds * a.max_ds2 + (int)(draws[off++] % (uint32_t)a.nds2[intent]);
        const int cnt = a.nclips[intent * a.max_ds + ds];
        const int clip = (int)(draws[off++] % (uint32_t)cnt);
        const int cid = a.clip_base[intent * a.max_ds + ds] + clip;
        r[0] = intent; r[1] = ds; r[2] = clip;
        *o = a.clip_off[cid];
        *l = a.clip_len[cid];
      };
      if (gt == T) {
        draw_clip(sn, rec + 3, &noff, &nlen);
      } else {
        draw_clip(gt, rec, &poff, &plen);
        if (sn != T) draw_clip(sn, rec + 3, &noff, &nlen);
      }
      const int o = c0 + i;
      a.out_sn[o] = sn;
      a.out_gt[o] = gt;
      a.out_item[o] = item;
#pragma unroll
      for (int j = 0; j < 6; ++j) a.out_rec[o * 6 + j] = rec[j];
      a.out_off[o] = poff; a.out_len[o] = plen;
      a.out_off[a.B + o] = noff; a.out_len[a.B + o] = nlen;
    }
    __syncthreads();
    const int consumed = s_off_next;
    __syncthreads();
    pos = mt_skip(s, pos, consumed);
    __syncthreads();
  }
  for (int i = threadIdx.x; i < MT_N; i += blockDim.x) a.state[i] = s[i];
  if (threadIdx.x == 0) a.state[MT_N] = (uint32_t)pos;
}

int sampler_seed(uint32_t* state, unsigned long long seed, cudaStream_t st) {
  LaunchScope sc(T_SAMPLER, 0, st);
  sampler_seed_kernel<<<1, 256, 0, st>>>(state, (uint32_t)(seed & 0xFFFFFFFFull));
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}
// Adopt a host generator state (torch.get_rng_state(): 624 words + read position) so the device
// stream continues exactly where torch's global CPU generator stands.
int sampler_set_state(uint32_t* state, const uint32_t* host_words, int pos, cudaStream_t st) {
  uint32_t tmp[kSamplerStateWords];
  for (int i = 0; i < MT_N; ++i) tmp[i] = host_words[i];
  tmp[MT_N] = (uint32_t)pos;
  // pageable source: the runtime stages the 2.5 KB before returning, so the stack buffer may die
  VAR_CUDA_CHECK(cudaMemcpyAsync(state, tmp, sizeof(tmp), cudaMemcpyHostToDevice, st));
  return VAR_OK;
}
int sampler_epoch(uint32_t* state, int n, int* perm, cudaStream_t st) {
  if (n <= 0) return VAR_ERR_ARG;
  const size_t smem = n <= 65536 ? (size_t)n * 2 + 16 : 0;
  VAR_ENSURE_SMEM(sampler_epoch_kernel, smem);
  LaunchScope sc(T_SAMPLER, 0, st);
  sampler_epoch_kernel<<<1, 256, smem, st>>>(state, n, perm);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}
int sampler_batch(const SamplerArgs& a_in, cudaStream_t st) {
  SamplerArgs a = a_in;
  if (a.B <= 0 || a.task_num <= 0 || a.task_num > 126) return VAR_ERR_ARG;
  if (a.nds2 && (a.max_ds2 <= 0 || a.max_ds % a.max_ds2 != 0)) return VAR_ERR_ARG;
  const int dpi = a.nds2 ? 7 : 5, cap = a.nds2 ? 4096 : 8192;  // keeps the draw window under the smem limit
  a.chunk = a.B < cap ? a.B : cap;
  const size_t smem = (size_t)(MT_N + dpi * a.chunk + MT_N) * 4 + (size_t)a.chunk * 4 + 16;
  VAR_ENSURE_SMEM(sampler_batch_kernel, smem);
  LaunchScope sc(T_SAMPLER, 0, st);
  sampler_batch_kernel<<<1, 256, smem, st>>>(a);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// ===========================================================================
// Fused head / triplet kernel
// ===========================================================================
constexpr int kMaxD = 8;
constexpr int kMaxKPerLane = 4;  // Kh <= 128

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct RoleOut {
  float e[kMaxD];
  float feat[kMaxD];
  float norm;
};

__device__ __forceinline__ void head_forward(const float* __restrict__ h, const float* sW,
                                             const float* sb, int Kh, int D, int lane,
                                             float* hreg, RoleOut& r) {
  const int kpl = Kh >> 5;
#pragma unroll
  for (int u = 0; u < kMaxKPerLane; ++u) hreg[u] = u < kpl ? h[lane + 32 * u] : 0.f;
  float nn = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxD; ++j) {
    if (j < D) {
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < kMaxKPerLane; ++u)
        if (u < kpl) acc = fmaf(sW[j * Kh + lane + 32 * u], hreg[u], acc);
      acc = warp_sum(acc) + sb[j];
      r.e[j] = acc;
      nn = fmaf(acc, acc, nn);
    } else {
      r.e[j] = 0.f;
    }
  }
  r.norm = sqrtf(nn);
  const float inv = 1.f / fmaxf(r.norm, 1e-12f);
#pragma unroll
  for (int j = 0; j < kMaxD; ++j) r.feat[j] = r.e[j] * inv;
}

// Backward of feat = e / max(|e|, eps), then of e = W h + b.  Accumulates dW/db in
// per-lane registers, writes the ReLU-masked, tf32-rounded grad of h.
__device__ __forceinline__ void head_backward(const RoleOut& r, const float* dfeat, const float* sW,
                                              int Kh, int D, int lane, const float* hreg,
                                              float* __restrict__ dh, float* accW, float* accb) {
  const int kpl = Kh >> 5;
  float dot = 0.f;
#pragma unroll
  for (int j = 0; j < kMaxD; ++j)
    if (j < D) dot = fmaf(r.feat[j], dfeat[j], dot);
  const bool clamped = !(r.norm > 1e-12f);
  const float inv = 1.f / fmaxf(r.norm, 1e-12f);
  float de[kMaxD];
#pragma unroll
  for (int j = 0; j < kMaxD; ++j)
    de[j] = j < D ? (clamped ? dfeat[j] * inv : (dfeat[j] - r.feat[j] * dot) * inv) : 0.f;
#pragma unroll
  for (int u = 0; u < kMaxKPerLane; ++u) {
    if (u < kpl) {
      float g = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxD; ++j)
        if (j < D) {
          g = fmaf(de[j], sW[j * Kh + lane + 32 * u], g);
          accW[j * kMaxKPerLane + u] = fmaf(de[j], hreg[u], accW[j * kMaxKPerLane + u]);
        }
      if (dh) dh[lane + 32 * u] = hreg[u] > 0.f ? round_tf32(g) : 0.f;
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < kMaxD; ++j)
      if (j < D) accb[j] += de[j];
  }
}

__global__ void __launch_bounds__(256) tail_kernel(TailArgs a) {
  extern __shared__ float smf[];
  const int D = a.D, Ki = a.Kh_img, Ks = a.Kh_snd;
  float* sWi = smf;                 // [D*Ki]
  float* sbi = sWi + D * Ki;        // [D]
  float* sWs = sbi + kMaxD;         // [D*Ks]
  float* sbs = sWs + D * Ks;        // [D]
  float* sdWi = sbs + kMaxD;        // [D*Ki] grads
  float* sdbi = sdWi + D * Ki;
  float* sdWs = sdbi + kMaxD;
  float* sdbs = sdWs + D * Ks;
  float* sloss = sdbs + kMaxD;      // [1]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int i = tid; i < D * Ki; i += blockDim.x) { sWi[i] = a.W_img ? a.W_img[i] : 0.f; sdWi[i] = 0.f; }
  for (int i = tid; i < D * Ks; i += blockDim.x) { sWs[i] = a.W_snd ? a.W_snd[i] : 0.f; sdWs[i] = 0.f; }
  if (tid < kMaxD) {
    sbi[tid] = (a.b_img && tid < D) ? a.b_img[tid] : 0.f;
    sbs[tid] = (a.b_snd && tid < D) ? a.b_snd[tid] : 0.f;
    sdbi[tid] = 0.f; sdbs[tid] = 0.f;
  }
  if (tid == 0) *sloss = 0.f;
  __syncthreads();

  float accWi[kMaxD * kMaxKPerLane], accWs[kMaxD * kMaxKPerLane], accbi[kMaxD], accbs[kMaxD];
#pragma unroll
  for (int i = 0; i < kMaxD * kMaxKPerLane; ++i) { accWi[i] = 0.f; accWs[i] = 0.f; }
#pragma unroll
  for (int i = 0; i < kMaxD; ++i) { accbi[i] = 0.f; accbs[i] = 0.f; }
  float loss_acc = 0.f;

  const bool do_bwd = a.mode == TAIL_TRIPLET || a.mode == TAIL_BWD;
  for (int row = blockIdx.x * nw + warp; row < a.B; row += gridDim.x * nw) {
    RoleOut ri, rp, rn;
    float hi[kMaxKPerLane], hp[kMaxKPerLane], hn[kMaxKPerLane];
    const bool has_i = a.h_img != nullptr, has_p = a.h_pos != nullptr, has_n = a.h_neg != nullptr;
    if (has_i) head_forward(a.h_img + (long long)row * Ki, sWi, sbi, Ki, D, lane, hi, ri);
    if (has_p) head_forward(a.h_pos + (long long)row * Ks, sWs, sbs, Ks, D, lane, hp, rp);
    if (has_n) head_forward(a.h_neg + (long long)row * Ks, sWs, sbs, Ks, D, lane, hn, rn);
    if (lane < D) {
      if (has_i && a.feat_img) a.feat_img[(long long)row * D + lane] = ri.feat[lane];
      if (has_p && a.feat_pos) a.feat_pos[(long long)row * D + lane] = rp.feat[lane];
      if (has_n && a.feat_neg) a.feat_neg[(long long)row * D + lane] = rn.feat[lane];
    }
    if (a.mode == TAIL_REWARD) {
      // reward = sum_d image_feat[:, :D] * goal_sound_feat + env_reward
      float gs[kMaxD];
#pragma unroll
      for (int j = 0; j < kMaxD; ++j)
        gs[j] = j < D ? (has_p ? rp.feat[j] : a.goal_feat_in[(long long)row * D + j]) : 0.f;
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxD; ++j)
        if (j < D) dot = fmaf(ri.feat[j], gs[j], dot);
      if (lane == 0) {
        if (a.dot_out) a.dot_out[row] = dot;
        if (a.reward_out) a.reward_out[row] = dot + (a.env_reward ? a.env_reward[row] : 0.f);
      }
      continue;
    }
    if (!do_bwd) continue;
    float dfi[kMaxD], dfp[kMaxD], dfn[kMaxD];
    if (a.mode == TAIL_TRIPLET) {
      // TripletMarginLoss(margin, p=2, eps=1e-6): d(x, y) = || x - y + eps ||
      float vp[kMaxD], vn[kMaxD], dp2 = 0.f, dn2 = 0.f;
#pragma unroll
      for (int j = 0; j < kMaxD; ++j) {
        vp[j] = j < D ? ri.feat[j] - rp.feat[j] + 1e-6f : 0.f;
        vn[j] = j < D ? ri.feat[j] - rn.feat[j] + 1e-6f : 0.f;
        dp2 = fmaf(vp[j], vp[j], dp2);
        dn2 = fmaf(vn[j], vn[j], dn2);
      }
      const float dp = sqrtf(dp2), dn = sqrtf(dn2);
      const float l = dp - dn + a.margin;
      const bool active = l > 0.f;
      if (lane == 0 && active) loss_acc += l;
      if (lane == 0 && a.loss_rows) a.loss_rows[row] = active ? l : 0.f;
      const float sc = active ? a.grad_scale : 0.f;
      const float ip = dp > 0.f ? sc / dp : 0.f, in = dn > 0.f ? sc / dn : 0.f;
#pragma unroll
      for (int j = 0; j < kMaxD; ++j) {
        dfp[j] = -vp[j] * ip;
        dfn[j] = vn[j] * in;
        dfi[j] = vp[j] * ip - vn[j] * in;
      }
    } else {
#pragma unroll
      for (int j = 0; j < kMaxD; ++j) {
        dfi[j] = (j < D && a.dfeat_img) ? a.dfeat_img[(long long)row * D + j] : 0.f;
        dfp[j] = (j < D && a.dfeat_pos) ? a.dfeat_pos[(long long)row * D + j] : 0.f;
        dfn[j] = (j < D && a.dfeat_neg) ? a.dfeat_neg[(long long)row * D + j] : 0.f;
      }
    }
    if (has_i && (a.mode == TAIL_TRIPLET || a.dfeat_img))
      head_backward(ri, dfi, sWi, Ki, D, lane, hi, a.dh_img ? a.dh_img + (long long)row * Ki : nullptr,
                    accWi, accbi);
    if (has_p && (a.mode == TAIL_TRIPLET || a.dfeat_pos))
      head_backward(rp, dfp, sWs, Ks, D, lane, hp, a.dh_pos ? a.dh_pos + (long long)row * Ks : nullptr,
                    accWs, accbs);
    if (has_n && (a.mode == TAIL_TRIPLET || a.dfeat_neg))
      head_backward(rn, dfn, sWs, Ks, D, lane, hn, a.dh_neg ? a.dh_neg + (long long)row * Ks : nullptr,
                    accWs, accbs);
  }
  if (!do_bwd) return;
  // CTA reduction of the weight gradients, then one atomic per element
  const int kpi = Ki >> 5, kps = Ks >> 5;
#pragma unroll
  for (int j = 0; j < kMaxD; ++j) {
    if (j < D) {
#pragma unroll
      for (int u = 0; u < kMaxKPerLane; ++u) {
        if (u < kpi) atomicAdd(&sdWi[j * Ki + lane + 32 * u], accWi[j * kMaxKPerLane + u]);
        if (u < kps) atomicAdd(&sdWs[j * Ks + lane + 32 * u], accWs[j * kMaxKPerLane + u]);
      }
      if (lane == 0) { atomicAdd(&sdbi[j], accbi[j]); atomicAdd(&sdbs[j], accbs[j]); }
    }
  }
  if (lane == 0 && a.mode == TAIL_TRIPLET) atomicAdd(sloss, loss_acc);
  __syncthreads();
  if (a.dW_img) for (int i = tid; i < D * Ki; i += blockDim.x) atomicAdd(a.dW_img + i, sdWi[i]);
  if (a.dW_snd) for (int i = tid; i < D * Ks; i += blockDim.x) atomicAdd(a.dW_snd + i, sdWs[i]);
  if (tid < D) {
    if (a.db_img) atomicAdd(a.db_img + tid, sdbi[tid]);
    if (a.db_snd) atomicAdd(a.db_snd + tid, sdbs[tid]);
  }
  if (tid == 0 && a.mode == TAIL_TRIPLET && a.loss) atomicAdd(a.loss, *sloss * a.loss_scale);
}

int tail_launch(const TailArgs& a, cudaStream_t st) {
  if (a.B <= 0) return VAR_OK;
  if (a.D < 1 || a.D > kMaxD) return VAR_ERR_UNSUPPORTED;
  if ((a.Kh_img % 32) || (a.Kh_snd % 32) || a.Kh_img > 32 * kMaxKPerLane ||
      a.Kh_snd > 32 * kMaxKPerLane || a.Kh_img <= 0 || a.Kh_snd <= 0)
    return VAR_ERR_UNSUPPORTED;
  if (a.mode == TAIL_REWARD && (!a.h_img || (!a.h_pos && !a.goal_feat_in))) return VAR_ERR_ARG;
  if (a.mode == TAIL_TRIPLET && (!a.h_img || !a.h_pos || !a.h_neg)) return VAR_ERR_ARG;
  // layout: 2 x (D*Ki + 8 + D*Ks + 8) + 1 floats
  const size_t smem2 = (size_t)(2 * (a.D * a.Kh_img + kMaxD + a.D * a.Kh_snd + kMaxD) + 4) * 4;
  const int warps_per_cta = 8;
  int grid = (a.B + warps_per_cta * 4 - 1) / (warps_per_cta * 4);  // >= 4 rows per warp
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;
  if (grid < 1) grid = 1;
  LaunchScope sc(T_TAIL, 0, st);
  tail_kernel<<<grid, warps_per_cta * 32, smem2, st>>>(a);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var
