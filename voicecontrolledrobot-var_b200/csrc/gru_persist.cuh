// Persistent GRU kernels: all time steps of torch.nn.GRU(448 -> 512, bidirectional)
// (models/pretext/ai2thor_pretext_model.py:6,33-38) in ONE cooperative launch per pass.
//
// A step-per-launch GRU spends more on launch gaps, barrier / TMEM set-up and cold TMA round
// trips than on its 1.6 GFLOP.  Here every CTA keeps its role for the whole sequence: tile =
// (128 batch rows) x (a slice of hidden units, all three gates) x (direction); barriers, TMEM and
// tensor-map prefetches are set up once; per step the CTA streams h_{s} (A, K-major, tiled TMA)
// and its W_hh slice (B) through the tcgen05 pipeline, runs the GRU cell (or the fused BPTT
// cell backward) in the epilogue and publishes its slice of h_{s+1}.  The only cross-CTA
// dependency is "all hidden slices of my (row tile, direction) are written": a monotonic global
// counter per group, release = CTA barrier + __threadfence + ONE atomicAdd per CTA and step,
// acquire = ld.acquire spin by the A-operand producer followed by a proxy fence before the TMA
// reads.  A pipeline stage holds two k-blocks (fewer barrier round trips per byte).
// Launched with cudaLaunchCooperativeKernel so all CTAs are co-resident (spin-waits are safe).
// The backward pass normally runs as gru_bwd_ksplit_kernel (gru_ksplit.cuh); the BWD
// instantiation here is its fallback when the 2-CTA clusters cannot all be resident.
#pragma once
#include "tc_engine.cuh"

namespace var {

struct GruPersistParams {
  int B, Hd, T;
  int bn;             // fwd: 3 * jb ; bwd: hidden units per CTA
  int num_kb;         // k-blocks per step: fwd Hd/32, bwd 3*Hd/32
  int stages;
  int kps;            // k-blocks (32 columns) per pipeline stage: 1 or 2
  int a_split;        // 1: the A tile arrives as four 32-row boxes issued by four threads
  int arrivals;       // CTA arrivals per group per step = gridDim.y
  int rt0;            // first 128-row tile of this launch (batches beyond the co-resident grid run in row chunks)
  unsigned int* counters;  // [gridDim.x * gridDim.z], zero before launch
  int mn_lbo, mn_sbo, mn_type;
  // forward
  const float* xproj[2];  // [B, T, 3H] incl. b_ih: element (b, t, c) at b * ldx + t * xts + c
  long long ldx, xts;
  const float* bhh[2];
  float* h32[2][2];       // fp32 hidden state ping-pong [B, H]
  float* h_r[2];          // [(T+1), B, H] tf32-rounded hidden states (slot 0 = zeros): A operand
  float* gates[2];        // [T][B, 3H] (nullptr = inference)
  float* hn_save[2];      // [T][B, H]
  // backward (step s = T-1 .. 1 computes the cell backward of step s-1)
  const float* gates_c[2];
  const float* hn_save_c[2];
  const float* h_r_c[2];
  float* dgh[2];          // [T][B, 3H]: A operand (slot s), written for slot s-1
  float* dgi[2];          // [B, T, 3H] batch-major
  float* dhd[2][2];       // dh * z ping-pong [B, H]
  long long* trace;       // optional [steps][8] clock64 samples of CTA (0,0,0) (VAR_GRU_TRACE=1)
  // 16-bit operand variants (H16): the recurrent GEMM operands are f16 copies -- half the per-step operand
  // stream that bounds both kernels -- while the cell state, gates and every saved tensor stay fp32.
  uint16_t* h_h[2];         // fwd: [(T+1), B, H] f16 hidden states (slot 0 = zeros): A operand
  uint16_t* dgh_h[2];       // bwd: [T][B, 3H] f16 gate gradients times gscale[z][0]: A operand
  const float* gscale[2];   // bwd: device {S, 1/S} per direction (power of two, from the last step's gradient)
  float* db_ih[2];          // bwd: bias gradients accumulated in the epilogue (+=, nullable): colsum(dgi) / colsum(dgh)
  float* db_hh[2];
  // bwd, 16-bit input projection: scaled f16 copy of dgi, element (b, t, c) at b * dgi_h_ld + t * dgi_h_ts + c
  // (both directions interleaved in one [B*T, 6H] matrix); skip_f32 drops the fp32 dgi / dgh stores (their only
  // readers are then the f16 weight-gradient / dgrad GEMMs)
  uint16_t* dgi_h[2];
  long long dgi_h_ld, dgi_h_ts;
  int skip_f32;
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}

// grid = (row tiles, hidden tiles, directions), block = kGruThreads.
//   warps 0-3   TMA issue (lane 0) and TMEM -> smem transposition of their 32-row quarter
//   warps 4-15  epilogue helpers: quarter q is finished by 4 warps (the TMEM warp + 3 helpers),
//               8 rows each, lane = hidden unit (128-byte coalesced global rows)
//   warp 4      additionally issues the MMAs of the step between its prefetch and its rows
// Every epilogue warp issues the global loads of its cell inputs (x-projection and h_{s-1}, or
// the saved gates for BPTT) at the TOP of the step, so they fly while the operands stream in
// and the MMAs run; after the accumulator is ready only TMEM -> smem, the gate maths and the
// stores are left.  (One epilogue warp per scheduler with loads after the MMA cost 12.8 of
// the 19.9 us per forward step.)
constexpr int kGruThreads = 512;
constexpr int kGruEpiWarps = 16;
__host__ __device__ inline size_t gru_scr_bytes(int bwd) { return (size_t)4 * (bwd ? 32 * 33 : 3 * 32 * 33) * 4; }

__device__ __forceinline__ void quarter_sync(int q) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
}

// f32 -> f16 with saturation to +-65504 instead of overflow to inf
__device__ __forceinline__ uint16_t f16_sat_bits(float x) {
  x = fminf(fmaxf(x, -65504.f), 65504.f);
  uint16_t h;
  asm("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}

template <int BWD, bool H16 = false>
__global__ void __launch_bounds__(kGruThreads, 1)
gru_persist_kernel(const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                   const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                   const __grid_constant__ GruPersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int z = blockIdx.z;
  const CUtensorMap* tmB = z == 0 ? &tmB0 : &tmB1;
  const CUtensorMap* tmA = z == 0 ? &tmA0 : &tmA1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, bn = p.bn, num_kb = p.num_kb;
  const int Hd = p.Hd, B = p.B, T = p.T;
  const uint32_t tileB_bytes = (uint32_t)bn * 128u;
  const int kps = p.kps;
  const uint32_t stageA = (uint32_t)kps * kTileABytes, stageB = (uint32_t)kps * tileB_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)stages * stageA;
  const uint32_t bars = sB + (uint32_t)stages * stageB;
  auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  const uint32_t tfull_bar = bars + (uint32_t)(2 * stages) * 8u;
  const uint32_t tslot = tfull_bar + 8u;
  float* scr_base = reinterpret_cast<float*>(smem_raw + (((tslot + 8u + 15u) & ~15u) - smem_u32(smem_raw)));

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(tmB);
    tma_prefetch_desc(tmA);
  }
  const uint32_t ncols = (uint32_t)tmem_cols_for(bn);
  if (warp == 4) tmem_alloc(tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  const int m0 = (p.rt0 + (int)blockIdx.x) * kTileM;
  const int ntile = blockIdx.y;
  unsigned int* counter = p.counters + (blockIdx.z * gridDim.x + blockIdx.x);
  const int nsteps = BWD ? T - 1 : T;
  const int jb = BWD ? bn : bn / 3;  // == 32: one 32-column chunk per gate
  const int nb_boxes = BWD ? (bn >> 5) : 3;
  const bool multi = nb_boxes >= 3;

  {
    const int quarter = warp & 3;
    const int sub = warp >> 2;
    static_assert(!(BWD && H16), "the 16-bit BPTT kernel is gru_bwd_ksplit_kernel<true>");
    constexpr int KE = H16 ? 64 : 32;  // elements per 128-byte k-block row
    const uint32_t idesc = H16 ? make_idesc_h16(bn, 0, 0, 0, 0) : make_idesc_tf32(bn, 0, BWD ? 1 : 0);
    const uint64_t adesc0 = make_smem_desc(sA, 16u, 1024u);
    const uint64_t bdesc0 = BWD ? make_smem_desc(sB, (uint32_t)p.mn_lbo, (uint32_t)p.mn_sbo, (uint32_t)p.mn_type)
                                : make_smem_desc(sB, 16u, 1024u);
    int mst = 0, mph = 0;  // MMA ring position (warp 4)
    constexpr int RB = 8;                         // rows per epilogue warp
    const int mrow0 = m0 + quarter * 32 + sub * RB;
    const int j = ntile * jb + lane;
    float* scr = scr_base + quarter * (BWD ? 32 * 33 : 3 * 32 * 33);
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int st = 0, ph = 0;  // producer ring position (kept by every issuing lane)
    const bool issuer = warp < 4 && lane == 0 && (p.a_split || warp == 0 || (multi && warp - 1 < nb_boxes));
    const bool tracer = p.trace && tid == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
    float br = 0.f, bz = 0.f, bq = 0.f;
    if constexpr (!BWD) { br = __ldg(p.bhh[z] + j); bz = __ldg(p.bhh[z] + Hd + j); bq = __ldg(p.bhh[z] + 2 * Hd + j); }
    // ---------------------------------------- cell inputs of a step: issued at the top of the step
    // (after the previous release: a fence behind outstanding loads would wait for them) so the
    // loads fly during the wait for peers, the operand stream and the MMAs
    float in0[RB], in1[RB], in2[RB], in3[RB], in4[RB], in5[RB];
    auto prefetch = [&](int it) {
      const int s = BWD ? T - 1 - it : it;
      if constexpr (!BWD) {
        const int t = z == 0 ? s : T - 1 - s;
        const float* __restrict__ xproj = p.xproj[z] + (long long)t * p.xts;
        const float* hprev = p.h32[z][s & 1];  // rows of this warp: written by this very thread last step
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int mr = mrow0 + u;
          const bool ok = mr < B;
          const float* xp = xproj + (long long)(ok ? mr : 0) * p.ldx;
          in0[u] = ok ? __ldg(xp + j) : 0.f;
          in1[u] = ok ? __ldg(xp + Hd + j) : 0.f;
          in2[u] = ok ? __ldg(xp + 2 * Hd + j) : 0.f;
          in3[u] = ok ? hprev[(long long)mr * Hd + j] : 0.f;
          in4[u] = 0.f; in5[u] = 0.f;
        }
      } else {
        const int sp = s - 1;
        const float* dhd_in = p.dhd[z][it & 1];  // written by this very thread last step
        const float* __restrict__ gates = p.gates_c[z] + (long long)sp * B * 3 * Hd;
        const float* __restrict__ hn_save = p.hn_save_c[z] + (long long)sp * B * Hd;
        const float* __restrict__ hprev = p.h_r_c[z] + (long long)sp * B * Hd;
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int mr = mrow0 + u;
          const bool ok = mr < B;
          const long long hoff = (long long)(ok ? mr : 0) * Hd + j;
          const float* gt = gates + (long long)(ok ? mr : 0) * 3 * Hd + j;
          in0[u] = ok ? dhd_in[hoff] : 0.f;
          in1[u] = ok ? __ldg(gt) : 0.f;
          in2[u] = ok ? __ldg(gt + Hd) : 0.f;
          in3[u] = ok ? __ldg(gt + 2 * Hd) : 0.f;
          in4[u] = ok ? __ldg(hn_save + hoff) : 0.f;
          in5[u] = ok ? __ldg(hprev + hoff) : 0.f;
        }
      }
    };
    for (int it_s = 0; it_s < nsteps; ++it_s) {
      const int s = BWD ? T - 1 - it_s : it_s;  // fwd: step index; bwd: slot whose dgh is the A operand
      if (tracer) p.trace[it_s * 8 + 0] = clock64();
      prefetch(it_s);
      if (warp < 4) {
        // ---------------------------------------------------------- producers
        if (issuer) {
          if (warp == 0 && it_s > 0) {
            const unsigned int target = (unsigned int)(p.arrivals * it_s);
            while (ld_acquire_gpu(counter) < target) {
            }
            fence_proxy_async_all();  // the generic-proxy writes just acquired are read by TMA below
          }
          if (tracer) p.trace[it_s * 8 + 1] = clock64();
          const int arow = s * B + m0;  // row of the A tile in the [slots * B, K] matrix
          for (int kb0 = 0; kb0 < num_kb; kb0 += kps) {
            mbar_wait(empty_bar(st), (uint32_t)(ph ^ 1));
            if (warp == 0) mbar_arrive_expect_tx(full_bar(st), stageA + stageB);
            for (int sub = 0; sub < kps; ++sub) {
              const int kb = kb0 + sub;
              const uint32_t dstA = sA + (uint32_t)st * stageA + (uint32_t)sub * kTileABytes;
              const uint32_t dstB = sB + (uint32_t)st * stageB + (uint32_t)sub * tileB_bytes;
              if (p.a_split) tma_load_2d(dstA + (uint32_t)warp * 4096u, tmA, full_bar(st), kb * KE, arow + warp * 32);
              else if (warp == 0) tma_load_2d(dstA, tmA, full_bar(st), kb * KE, arow);
              if constexpr (!BWD) {
                for (int b = 0; b < 3; ++b)
                  if (warp == 1 + b)
                    tma_load_2d(dstB + (uint32_t)(b * jb) * 128u, tmB, full_bar(st), kb * KE, b * Hd + ntile * jb);
              } else {
                for (int gidx = 0; gidx < (bn >> 5); ++gidx)
                  if (warp == ((multi || p.a_split) ? 1 + (gidx % 3) : 0))
                    tma_load_2d(dstB + (uint32_t)gidx * 4096u, tmB, full_bar(st), ntile * bn + gidx * 32, kb * 32);
              }
            }
            if (++st == stages) { st = 0; ph ^= 1; }
          }
        }
        __syncwarp();
        if (tracer) p.trace[it_s * 8 + 2] = clock64();
        // -------------------------------------- accumulator quarter -> shared memory (transposed use)
        mbar_wait(tfull_bar, (uint32_t)(it_s & 1));
        tc_fence_after();
        if (tracer) p.trace[it_s * 8 + 3] = clock64();
        float v[32];
#pragma unroll
        for (int gq = 0; gq < (BWD ? 1 : 3); ++gq) {
          tmem_ld32(trow + (uint32_t)(gq * jb), v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) scr[(gq * 32 + lane) * 33 + c] = v[c];
        }
        tc_fence_before();
      } else if (warp == 4) {
        // ---------------------------------------------------------- MMA issue for this step
        for (int kb0 = 0; kb0 < num_kb; kb0 += kps) {
          mbar_wait(full_bar(mst), (uint32_t)mph);
          tc_fence_after();
          if (elect_one_sync()) {  // one lane of the converged warp: operands stay in uniform registers
            for (int sub = 0; sub < kps; ++sub) {  // base descriptor + start-address offset (bytes >> 4)
              const uint64_t ad0 = adesc0 + (uint64_t)(((uint32_t)mst * stageA + (uint32_t)sub * kTileABytes) >> 4);
              const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)mst * stageB + (uint32_t)sub * tileB_bytes) >> 4);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                if constexpr (H16)
                  umma_f16(tmem_base, ad0 + (uint64_t)(jj * 2), bd0 + (uint64_t)(jj * 2), idesc,
                           (uint32_t)((kb0 | sub | jj) != 0));
                else
                  umma_tf32(tmem_base, ad0 + (uint64_t)(jj * 2), bd0 + (uint64_t)(jj * (BWD ? 64 : 2)), idesc,
                            (uint32_t)((kb0 | sub | jj) != 0));
              }
            }
            umma_commit(empty_bar(mst));
            if (kb0 + kps >= num_kb) umma_commit(tfull_bar);
          }
          __syncwarp();
          if (++mst == stages) { mst = 0; mph ^= 1; }
        }
      }
      quarter_sync(quarter);
      // --------------------------------------------------------------- cell maths, 8 rows
      if constexpr (!BWD) {
        float* hnew = p.h32[z][(s + 1) & 1];
        float* hnew_r = p.h_r[z] + (long long)(s + 1) * B * Hd;
        float* gates = p.gates[z] ? p.gates[z] + (long long)s * B * 3 * Hd : nullptr;
        float* hn_save = p.hn_save[z] ? p.hn_save[z] + (long long)s * B * Hd : nullptr;
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int rr = sub * RB + u, mr = mrow0 + u;
          if (mr < B) {
            const float vr = scr[(0 * 32 + rr) * 33 + lane], vz = scr[(1 * 32 + rr) * 33 + lane],
                        vn = scr[(2 * 32 + rr) * 33 + lane];
            const float r_ = sigmoidf_(in0[u] + vr + br);
            const float z_ = sigmoidf_(in1[u] + vz + bz);
            const float hnv = vn + bq;
            const float n_ = tanhf_(in2[u] + r_ * hnv);
            const float h_ = (1.f - z_) * n_ + z_ * in3[u];
            const long long ho = (long long)mr * Hd + j;
            hnew[ho] = h_;
            hnew_r[ho] = round_tf32(h_);
            if constexpr (H16) p.h_h[z][(long long)(s + 1) * B * Hd + ho] = f16_sat_bits(h_);  // |h| < 1
            if (gates) {
              float* gs = gates + (long long)mr * 3 * Hd + j;
              gs[0] = r_; gs[Hd] = z_; gs[2 * Hd] = n_;
              hn_save[ho] = hnv;
            }
          }
        }
      } else {
        // acc = dgh_s . W_hh ; dh_{s-1} = acc + dh_s * z_s ; then the cell backward of step s-1
        const int sp = s - 1;
        const int t = z == 0 ? sp : T - 1 - sp;
        float* dhd_out = p.dhd[z][(it_s + 1) & 1];
        float* dgi = p.dgi[z] + (long long)t * 3 * Hd;
        const long long ldgi = (long long)T * 3 * Hd;
        float* dgh = p.dgh[z] + (long long)sp * B * 3 * Hd;
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int rr = sub * RB + u, mr = mrow0 + u;
          if (mr < B) {
            const long long hoff = (long long)mr * Hd + j;
            const float dh = scr[rr * 33 + lane] + in0[u];
            const float r_ = in1[u], z_ = in2[u], n_ = in3[u];
            const float dnn = dh * (1.f - z_);
            const float dzz = dh * (in5[u] - n_);
            const float dnp = dnn * (1.f - n_ * n_);
            const float dzp = dzz * z_ * (1.f - z_);
            const float drp = dnp * in4[u] * r_ * (1.f - r_);
            const float dr = round_tf32(drp), dz = round_tf32(dzp), dn = round_tf32(dnp);
            float* gi = dgi + (long long)mr * ldgi + j;
            gi[0] = dr; gi[Hd] = dz; gi[2 * Hd] = dn;
            float* gh = dgh + (long long)mr * 3 * Hd + j;
            gh[0] = dr; gh[Hd] = dz; gh[2 * Hd] = round_tf32(dnp * r_);
            dhd_out[hoff] = dh * z_;
          }
        }
      }
      if (tracer) p.trace[it_s * 8 + 4] = clock64();
      // publish the CTA's slice of the new state: ONE release per CTA per step (256 same-address
      // atomics per group and step -- one per epilogue warp -- serialise in L2 for ~3 us).  All
      // warps pass the barrier after their stores; thread 0's fence is cumulative over them.
      // The next write of a quarter's scratch comes after the next accumulator is ready, i.e.
      // after every CTA of the group has arrived here.
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
      }
      if (tracer) p.trace[it_s * 8 + 5] = clock64();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

}  // namespace var
