// BPTT steps of the bidirectional GRU with the reduction dimension split over a 2-CTA cluster.
//
// gru_persist_kernel<1> gives every CTA a (128 rows x 32 hidden units) tile, so each of the 16
// CTAs of a (row tile, direction) group streams the whole dgh tile (128 x 1536 fp32 = 768 KB) plus
// its W_hh slice (192 KB) per step -- and the step is bound by that 960 KB operand stream per SM.
// Here a tile is 128 rows x 64 hidden units owned by a CLUSTER of two CTAs: CTA r multiplies
// columns [768 r, 768 r + 768) of dgh with the matching rows of W_hh into its own TMEM accumulator
// (384 KB + 192 KB per step), then the two exchange halves of their partial sums through
// distributed shared memory (CTA r keeps hidden units [32 r, 32 r + 32) of the tile and receives
// the partner's 128 x 32 partials: 16 KB) and each runs the fused cell backward for its 32
// units.  Everything else -- 16 epilogue warps with prefetched cell inputs, one release per CTA
// and step on the group counter -- is as in gru_persist.cuh.
#pragma once
#include "gru_persist.cuh"

namespace var {

__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_shared_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// grid = (2 = k half / cluster rank, hidden tiles of 64, row tiles * 2 directions), cluster (2,1,1),
// block = kGruThreads.  p.bn = 64, p.num_kb = k-blocks per CTA and step (24), p.kps | p.num_kb.
// H16: dgh (A) and W_hh (B, read transposed) are f16; dgh_h carries the gradients times the power-of-two
// scale gscale[z][0], the epilogue multiplies the partial sums by 1/S.  A k-block is 64 columns of dgh.
template <bool H16>
__global__ void __launch_bounds__(kGruThreads, 1)
gru_bwd_ksplit_kernel(const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
                      const __grid_constant__ GruPersistParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int rank = blockIdx.x;            // == %cluster_ctarank (cluster spans grid.x)
  const int ntile = blockIdx.y;           // 64 hidden units
  const int rt = blockIdx.z >> 1, z = blockIdx.z & 1;
  const CUtensorMap* tmB = z == 0 ? &tmB0 : &tmB1;
  const CUtensorMap* tmA = z == 0 ? &tmA0 : &tmA1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, num_kb = p.num_kb, kps = p.kps;
  const int Hd = p.Hd, B = p.B, T = p.T;
  constexpr int bn = 64;
  constexpr uint32_t tileB_bytes = bn * 128u;
  const uint32_t stageA = (uint32_t)kps * kTileABytes, stageB = (uint32_t)kps * tileB_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)stages * stageA;
  const uint32_t bars = sB + (uint32_t)stages * stageB;
  auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  const uint32_t tfull_bar = bars + (uint32_t)(2 * stages) * 8u;
  const uint32_t tslot = tfull_bar + 8u;
  const uint32_t scr_u32 = (tslot + 8u + 15u) & ~15u;
  float* scr_base = reinterpret_cast<float*>(smem_raw + (scr_u32 - smem_u32(smem_raw)));
  constexpr int kQuarterFloats = 32 * 33;
  float* xch_base = scr_base + 4 * kQuarterFloats;                      // partner's partial sums land here
  const uint32_t xch_u32 = scr_u32 + 4u * kQuarterFloats * 4u;

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
  if (warp == 4) tmem_alloc(tslot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));
  cluster_sync_all();  // the partner's shared memory exists before anybody writes into it

  const int m0 = (p.rt0 + rt) * kTileM;
  unsigned int* counter = p.counters + (z * (int)(gridDim.z >> 1) + rt);
  const int nsteps = T - 1;
  const int kb_base = rank * num_kb;
  const int quarter = warp & 3, sub = warp >> 2;
  constexpr int KE = H16 ? 64 : 32;  // dgh columns per k-block
  const uint32_t idesc = H16 ? make_idesc_h16(bn, 0, 0, 0, 1) : make_idesc_tf32(bn, 0, 1);
  const uint64_t adesc0 = make_smem_desc(sA, 16u, 1024u);
  const uint64_t bdesc0 = H16 ? make_smem_desc(sB, 8192u, 1024u, 2)
                              : make_smem_desc(sB, (uint32_t)p.mn_lbo, (uint32_t)p.mn_sbo, (uint32_t)p.mn_type);
  float gS = 1.f, gInv = 1.f;
  if constexpr (H16) { gS = __ldg(p.gscale[z]); gInv = __ldg(p.gscale[z] + 1); }
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};  // column sums of dr, dz, dn, dn*r over this thread's rows and all steps
  int mst = 0, mph = 0;  // MMA ring position (warp 4)
  int st = 0, ph = 0;    // producer ring position (warp 0 lane 0)
  constexpr int RB = 8;
  const int mrow0 = m0 + quarter * 32 + sub * RB;
  const int j = ntile * 64 + rank * 32 + lane;  // hidden unit of this lane
  float* scr = scr_base + quarter * kQuarterFloats;
  const float* xch = xch_base + quarter * kQuarterFloats;
  const uint32_t xch_remote = mapa_shared(xch_u32 + (uint32_t)quarter * kQuarterFloats * 4u, (uint32_t)(rank ^ 1));
  const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16);
  const unsigned int arrivals = gridDim.x * gridDim.y;

  for (int it_s = 0; it_s < nsteps; ++it_s) {
    const int s = T - 1 - it_s;  // slot whose dgh is the A operand
    // ---- cell inputs of step s-1, in flight during the operand stream
    float in0[RB], in1[RB], in2[RB], in3[RB], in4[RB], in5[RB];
    {
      const int sp = s - 1;
      const float* dhd_in = p.dhd[z][it_s & 1];  // written by this very thread last step
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
    if (warp == 0) {
      // ---- producer: this CTA's half of the k-blocks
      if (lane == 0) {
        if (it_s > 0) {
          const unsigned int target = arrivals * (unsigned int)it_s;
          while (ld_acquire_gpu(counter) < target) {
          }
          fence_proxy_async_all();
        }
        const int arow = s * B + m0;
        for (int kb0 = 0; kb0 < num_kb; kb0 += kps) {
          mbar_wait(empty_bar(st), (uint32_t)(ph ^ 1));
          mbar_arrive_expect_tx(full_bar(st), stageA + stageB);
          for (int sb = 0; sb < kps; ++sb) {
            const int kb = kb_base + kb0 + sb;
            const uint32_t dstA = sA + (uint32_t)st * stageA + (uint32_t)sb * kTileABytes;
            const uint32_t dstB = sB + (uint32_t)st * stageB + (uint32_t)sb * tileB_bytes;
            tma_load_2d(dstA, tmA, full_bar(st), kb * KE, arow);
            if constexpr (H16) {
              tma_load_2d(dstB, tmB, full_bar(st), ntile * 64, kb * KE);  // one {64 n, 64 k} box
            } else {
              tma_load_2d(dstB, tmB, full_bar(st), ntile * 64, kb * 32);
              tma_load_2d(dstB + 4096u, tmB, full_bar(st), ntile * 64 + 32, kb * 32);
            }
          }
          if (++st == stages) { st = 0; ph ^= 1; }
        }
      }
      __syncwarp();
    } else if (warp == 4) {
      // ---- MMA issue
      if (lane == 0) {
        for (int kb0 = 0; kb0 < num_kb; kb0 += kps) {
          mbar_wait(full_bar(mst), (uint32_t)mph);
          tc_fence_after();
          for (int sb = 0; sb < kps; ++sb) {
            const uint64_t ad0 = adesc0 + (uint64_t)(((uint32_t)mst * stageA + (uint32_t)sb * kTileABytes) >> 4);
            const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)mst * stageB + (uint32_t)sb * tileB_bytes) >> 4);
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              if constexpr (H16)
                umma_f16(tmem_base, ad0 + (uint64_t)(jj * 2), bd0 + (uint64_t)(jj * 128), idesc,
                         (uint32_t)((kb0 | sb | jj) != 0));
              else
                umma_tf32(tmem_base, ad0 + (uint64_t)(jj * 2), bd0 + (uint64_t)(jj * 64), idesc,
                          (uint32_t)((kb0 | sb | jj) != 0));
            }
          }
          umma_commit(empty_bar(mst));
          if (kb0 + kps >= num_kb) umma_commit(tfull_bar);
          if (++mst == stages) { mst = 0; mph ^= 1; }
        }
      }
      __syncwarp();
    }
    if (warp < 4) {
      // ---- partial sums: keep my 32 hidden units (-> scr), send the other 32 to the partner
      mbar_wait(tfull_bar, (uint32_t)(it_s & 1));
      tc_fence_after();
      float v[32];
      tmem_ld32(trow + (uint32_t)(rank * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) scr[lane * 33 + c] = v[c];
      tmem_ld32(trow + (uint32_t)((rank ^ 1) * 32), v);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) st_shared_cluster_f32(xch_remote + (uint32_t)(lane * 33 + c) * 4u, v[c]);
      tc_fence_before();
    }
    cluster_sync_all();  // both CTAs: own partials in scr, partner's in xch
    // ---- dh_{s-1} = acc + dh_s * z_s ; then the cell backward of step s-1 (8 rows per warp)
    {
      const int sp = s - 1;
      const int t = z == 0 ? sp : T - 1 - sp;
      float* dhd_out = p.dhd[z][(it_s + 1) & 1];
      float* dgi = p.dgi[z] + (long long)t * 3 * Hd;
      const long long ldgi = (long long)T * 3 * Hd;
      float* dgh = p.dgh[z] + (long long)sp * B * 3 * Hd;
      if (it_s == 0 && p.db_ih[z]) {
        // the cell backward of the LAST step ran in gru_cell_bwd: fold its gate gradients into the bias sums
        const float* gl = p.dgh[z] + (long long)s * B * 3 * Hd;
        const float* gil = p.dgi[z] + (long long)(z == 0 ? s : T - 1 - s) * 3 * Hd;
#pragma unroll
        for (int u = 0; u < RB; ++u) {
          const int mr = mrow0 + u;
          if (mr < B) {
            const float* gh = gl + (long long)mr * 3 * Hd + j;
            bsum[0] += gh[0]; bsum[1] += gh[Hd]; bsum[3] += gh[2 * Hd];
            bsum[2] += gil[(long long)mr * ldgi + 2 * Hd + j];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int rr = sub * RB + u, mr = mrow0 + u;
        if (mr < B) {
          const long long hoff = (long long)mr * Hd + j;
          const float dh = (scr[rr * 33 + lane] + xch[rr * 33 + lane]) * gInv + in0[u];
          const float r_ = in1[u], z_ = in2[u], n_ = in3[u];
          const float dnn = dh * (1.f - z_);
          const float dzz = dh * (in5[u] - n_);
          const float dnp = dnn * (1.f - n_ * n_);
          const float dzp = dzz * z_ * (1.f - z_);
          const float drp = dnp * in4[u] * r_ * (1.f - r_);
          const float dr = round_tf32(drp), dz = round_tf32(dzp), dn = round_tf32(dnp);
          const float dnr = round_tf32(dnp * r_);
          if (!p.skip_f32) {
            float* gi = dgi + (long long)mr * ldgi + j;
            gi[0] = dr; gi[Hd] = dz; gi[2 * Hd] = dn;
            float* gh = dgh + (long long)mr * 3 * Hd + j;
            gh[0] = dr; gh[Hd] = dz; gh[2 * Hd] = dnr;
          }
          if constexpr (H16) {
            uint16_t* ghh = p.dgh_h[z] + (long long)sp * B * 3 * Hd + (long long)mr * 3 * Hd + j;
            const uint16_t hr = f16_sat_bits(dr * gS), hz = f16_sat_bits(dz * gS);
            ghh[0] = hr; ghh[Hd] = hz; ghh[2 * Hd] = f16_sat_bits(dnr * gS);
            if (p.dgi_h[z]) {
              uint16_t* gih = p.dgi_h[z] + (long long)mr * p.dgi_h_ld + (long long)t * p.dgi_h_ts + j;
              gih[0] = hr; gih[Hd] = hz; gih[2 * Hd] = f16_sat_bits(dn * gS);
            }
          }
          bsum[0] += dr; bsum[1] += dz; bsum[2] += dn; bsum[3] += dnr;
          dhd_out[hoff] = dh * z_;
        }
      }
    }
    // ---- publish: one release per CTA and step.  (scr / xch are rewritten only after the next
    // accumulators are ready, i.e. after every CTA of the group -- the partner included -- got here.)
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
    }
  }
  if (p.db_ih[z]) {  // bias gradients: b_ih gets (dr, dz, dn), b_hh gets (dr, dz, dn * r)
    atomicAdd(p.db_ih[z] + j, bsum[0]); atomicAdd(p.db_ih[z] + Hd + j, bsum[1]); atomicAdd(p.db_ih[z] + 2 * Hd + j, bsum[2]);
    atomicAdd(p.db_hh[z] + j, bsum[0]); atomicAdd(p.db_hh[z] + Hd + j, bsum[1]); atomicAdd(p.db_hh[z] + 2 * Hd + j, bsum[3]);
  }
  tc_fence_before();
  cluster_sync_all();  // nobody leaves while the partner could still address its shared memory
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

}  // namespace var
