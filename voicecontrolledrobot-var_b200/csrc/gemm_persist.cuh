// Persistent form of the TMA-fed implicit GEMM (tc_engine.cuh) for the standard epilogue.
//
// One launch of tc_gemm_kernel pays, per 128-row tile, a CTA launch, barrier initialisation, a
// TMEM allocation, a cold pipeline fill and a fully exposed epilogue -- about as long as 45
// k-blocks of MMA work, i.e. more than the main loop of every dgrad sub-problem (20-36
// k-blocks) and of the 3x3 image convs (9-18).  Here a CTA (192 threads) is resident for the
// whole GEMM and walks tiles blockIdx.x, +gridDim.x, ...:
//   warp 5      TMA producer (lane 0; lanes 1-3 help when a k-block needs >= 4 instructions)
//   warp 4      MMA issuer, alternating between two TMEM accumulators
//   warps 0-3   epilogue of tile i (tcgen05.ld -> bias / add / ReLU / mask / round -> HBM) while
//               the producer and the tensor core already work on tile i+1
// Barriers: full/empty per smem stage (ring shared by all tiles), tfull/tempty per accumulator.
// A stage holds p.kps (1-2) k-blocks; the producer and the MMA issuer are single threads whose
// instruction streams pace the CTA (4 MMAs of a 64-column k-block are 128 tensor-pipe cycles), so
// they carry no divisions and no descriptor rebuilds in their loops.  Wide short-K tiles use a
// coalescing epilogue (p.epi_coalesce): accumulator chunks transposed through a swizzled smem tile.
#pragma once
#include "tc_engine.cuh"

namespace var {

// H16: operands are 16-bit (f16 activations / weights, bf16 gradients; kind::f16 MMAs with K = 16): a
// k-block is still one 128-byte row per pixel / weight row, i.e. 64 elements instead of 32 -- half the
// shared-memory fill per MAC, which is what bounds the N = 64 convs.  MN-major B
// boxes are {64 n, 64 k} (8 KB) in the plain 128B swizzle (16-byte granules).  The epilogue can store
// f16 / bf16 (e.out_kind) and read an f16 ReLU mask (e.mask_kind); no addsrc / coalescing path.
template <int GMODE, bool H16 = false>
__global__ void __launch_bounds__(192, 2)
tc_gemm_persist_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmA,
                       const __grid_constant__ CUtensorMap tmC, const __grid_constant__ GemmParams p, int m_tiles,
                       int n_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const GatherGeom& g = p.g[0];
  const EpiParams& e = p.e[0];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, bn = p.bn, num_kb = p.num_kb;
  const uint32_t tileB_bytes = (uint32_t)bn * 128u;
  const int kps = p.kps;  // k-blocks per pipeline stage: fewer barrier round trips per byte
  const uint32_t stageA = (uint32_t)kps * kTileABytes, stageB = (uint32_t)kps * tileB_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)stages * stageA;
  const uint32_t bars = sB + (uint32_t)stages * stageB;
  auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  auto tfull_bar = [&](int a) { return bars + (uint32_t)(2 * stages + a) * 8u; };
  auto tempty_bar = [&](int a) { return bars + (uint32_t)(2 * stages + 2 + a) * 8u; };
  const uint32_t tslot = bars + (uint32_t)(2 * stages + 4) * 8u;
  const uint32_t stage_base = bars + 128u;  // epi_coalesce: 4 x (32 rows x 128 B + 32 row offsets)
  const uint32_t tma_stage = (bars + 256u + 1023u) & ~1023u;  // epi_tma: two 128-row x 128-byte swizzled boxes

  const uint32_t acc_cols = (uint32_t)tmem_cols_for(bn);
  const int nacc = acc_cols * 2 <= 256 ? 2 : 1;  // 2 CTAs per SM share the 512 TMEM columns
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    mbar_fence_init();
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmA);
  }
  if (warp == 4) tmem_alloc(tslot, acc_cols * (uint32_t)nacc);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  const int total = m_tiles * n_tiles;
  if (warp == 5) {
    // ===================== TMA producer =====================
    constexpr int KSH = H16 ? 6 : 5;  // log2(elements per 128-byte k-block row)
    const int nb_boxes = p.b_mn_major ? (bn >> KSH) : p.nbox;
    {  // the converged warp runs the loop, one elected lane issues every box of a stage from uniform registers
      int st = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int ntile = tile % n_tiles, mtile = tile / n_tiles;
        const int m0 = mtile * kTileM;
        int w0 = 0, h0 = 0, n0 = 0;
        if constexpr (GMODE == G_TMA_IM2COL) {
          const int pq0 = g.P * g.Q;
          n0 = m0 / pq0;
          const int rem0 = m0 - n0 * pq0;
          const int p0 = rem0 / g.Q;
          w0 = (rem0 - p0 * g.Q) * p.step_w + p.base_w;
          h0 = p0 * p.step_h + p.base_h;
        }
        // (tap, channel block) = divmod(k-block, blocks per tap), kept incrementally: this thread's
        // instruction stream is on the critical path
        const int period = GMODE == G_TMA_IM2COL ? p.cpb : (p.b_mn_major ? p.kb_per_rs : 1 << 30);
        int tap = 0, cb = 0;
        for (int it0 = 0; it0 < num_kb; it0 += kps) {
          const int nsub = min(kps, num_kb - it0);
          mbar_wait(empty_bar(st), (uint32_t)(ph ^ 1));
          const bool leader = elect_one_sync();
          if (leader) mbar_arrive_expect_tx(full_bar(st), (uint32_t)nsub * ((uint32_t)kTileABytes + tileB_bytes));
          for (int sub = 0; sub < nsub; ++sub) {
            const int it = it0 + sub;
            const uint32_t dstA = sA + (uint32_t)st * stageA + (uint32_t)sub * kTileABytes;
            const uint32_t dstB = sB + (uint32_t)st * stageB + (uint32_t)sub * tileB_bytes;
            const int c0 = GMODE == G_TMA_IM2COL ? (cb << KSH) : (it << KSH);
            if (leader) {
              if constexpr (GMODE == G_TMA_IM2COL)
                tma_load_im2col_4d(dstA, &tmA, full_bar(st), c0, w0, h0, n0, p.tap_w[tap], p.tap_h[tap]);
              else
                tma_load_2d(dstA, &tmA, full_bar(st), it << KSH, m0);
              if (!p.b_mn_major) {
                for (int b = 0; b < p.nbox; ++b)
                  tma_load_2d(dstB + (uint32_t)(b * p.box_rows) * 128u, &tmB, full_bar(st), it << KSH,
                              p.boxbase[b] + ntile * p.box_rows);
              } else {
                const int rs = GMODE == G_TMA_IM2COL ? (int)p.tap_id[tap] : tap;
                const int k0 = cb << KSH;
                // one box = {128 bytes of n, one k-block of k rows}: 4 KB (tf32: 32 x 32) or 8 KB (16-bit: 64 x 64)
                for (int gidx = 0; gidx < (bn >> KSH); ++gidx)
                  tma_load_2d(dstB + (uint32_t)gidx * (H16 ? 8192u : 4096u), &tmB, full_bar(st),
                              rs * p.cin_total + ntile * bn + (gidx << KSH), k0);
              }
            }
            if (++cb == period) { cb = 0; ++tap; }
          }
          __syncwarp();
          if (++st == stages) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = H16 ? make_idesc_h16(bn, p.a_fmt, p.b_fmt, 0, p.b_mn_major) : make_idesc_tf32(bn, 0, p.b_mn_major);
    const uint64_t adesc0 = make_smem_desc(sA, 16u, 1024u);
    // MN-major B: tf32 = 32-byte-granule swizzle (mn_cfg), 16-bit = plain 128B swizzle with 8-k-row atoms
    // (SBO 1024) and 8 KB between 64-wide n groups (LBO)
    const uint64_t bdesc0 = p.b_mn_major ? (H16 ? make_smem_desc(sB, 8192u, 1024u, 2)
                                                : make_smem_desc(sB, (uint32_t)p.mn_lbo, (uint32_t)p.mn_sbo, (uint32_t)p.mn_type))
                                         : make_smem_desc(sB, 16u, 1024u);
    // start-address step per MMA (>> 4): K-major 32 B; MN-major 8 k-rows (tf32, 1024 B) / 16 k-rows (16-bit, 2048 B)
    const int bstep = p.b_mn_major ? (H16 ? 128 : 64) : 2;
    int st = 0, ph = 0, i = 0;
    // the converged warp runs the loop and one elected lane issues: descriptor arithmetic stays on the uniform datapath (a
    // `lane == 0` loop wraps every UTCHMMA in an ELECT / BRA.U.ANY loop and moves each operand through R2UR)
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int acc = nacc == 2 ? (i & 1) : 0;
      const int use = nacc == 2 ? (i >> 1) : i;            // how often this accumulator was used
      mbar_wait(tempty_bar(acc), (uint32_t)((use & 1) ^ 1));  // epilogue drained it
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
      for (int kb0 = 0; kb0 < num_kb; kb0 += kps) {
        const int nsub = min(kps, num_kb - kb0);
        mbar_wait(full_bar(st), (uint32_t)ph);
        tc_fence_after();
        if (elect_one_sync()) {
          // descriptors differ only in their 14-bit start-address field (bytes >> 4): add offsets
          // to a base descriptor instead of rebuilding both per MMA (the issuing thread is on the
          // critical path: 4 MMAs of a 64-column k-block are only 128 tensor-pipe cycles)
          for (int sub = 0; sub < nsub; ++sub) {
            const uint64_t ad0 = adesc0 + (uint64_t)(((uint32_t)st * stageA + (uint32_t)sub * kTileABytes) >> 4);
            const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)st * stageB + (uint32_t)sub * tileB_bytes) >> 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if constexpr (H16)
                umma_f16(d_tmem, ad0 + (uint64_t)(j * 2), bd0 + (uint64_t)(j * bstep), idesc,
                         (uint32_t)((kb0 | sub | j) != 0));
              else
                umma_tf32(d_tmem, ad0 + (uint64_t)(j * 2), bd0 + (uint64_t)(j * bstep), idesc,
                          (uint32_t)((kb0 | sub | j) != 0));
            }
          }
          umma_commit(empty_bar(st));
          if (kb0 + kps >= num_kb) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    tc_fence_before();
  } else {
    // ===================== epilogue warps 0-3 =====================
    int i = 0;
    uint32_t nbox_out = 0;  // epi_tma: boxes stored so far (staging buffer = parity)
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int ntile = tile % n_tiles, mtile = tile / n_tiles;
      const int acc = nacc == 2 ? (i & 1) : 0;
      const int use = nacc == 2 ? (i >> 1) : i;
      const int m = mtile * kTileM + warp * 32 + lane;
      long long orow = m;
      if (e.map.on) {
        const int pq2 = e.map.P2 * e.map.Q2;
        const int n_ = m / pq2, rem_ = m - n_ * pq2;
        const int h2 = rem_ / e.map.Q2, w2 = rem_ - h2 * e.map.Q2;
        orow = ((long long)n_ * e.map.H + h2 * e.map.sh + e.map.oh) * e.map.W + w2 * e.map.sw + e.map.ow;
      }
      // ReLU-backward mask of this thread's output row, fetched BEFORE waiting for the accumulator and packed to one
      // bit per column: short-K dgrad tiles (1-4 k-blocks for stride-2 3x3 convs) are epilogue bound, and the mask
      // read was a dependent global round trip between tcgen05.ld and the store.
      uint32_t mbits[8];
      const bool mask_pre = e.mask != nullptr && !p.epi_coalesce && bn <= 256;
      if (mask_pre) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          mbits[q] = 0xFFFFFFFFu;
          const int col0 = ntile * bn + q * 32;
          if (q * 32 < bn && m < g.M && col0 < e.ncols) {
            uint32_t bits = 0u;
            if (H16 && e.mask_kind == 1) {
              const uint4* mk = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(e.mask) + orow * e.ldm + col0);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const uint4 w4 = __ldg(mk + q4);
                const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  // f16 > 0  <=>  sign bit clear and magnitude bits non-zero
                  const uint32_t lo = ww[u] & 0xFFFFu, hi = ww[u] >> 16;
                  bits |= (uint32_t)(lo != 0u && lo < 0x8000u) << (q4 * 8 + u * 2);
                  bits |= (uint32_t)(hi != 0u && hi < 0x8000u) << (q4 * 8 + u * 2 + 1);
                }
              }
            } else {
              const float4* mk = reinterpret_cast<const float4*>(e.mask + orow * e.ldm + col0);
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 k4 = __ldg(mk + j);
                bits |= (uint32_t)(k4.x > 0.f) << (4 * j) | (uint32_t)(k4.y > 0.f) << (4 * j + 1) |
                        (uint32_t)(k4.z > 0.f) << (4 * j + 2) | (uint32_t)(k4.w > 0.f) << (4 * j + 3);
              }
            }
            mbits[q] = bits;
          }
        }
      }
      mbar_wait(tfull_bar(acc), (uint32_t)(use & 1));
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)acc * acc_cols + ((uint32_t)(warp * 32) << 16);
      if (p.epi_tma) {
        // Plain fp32 tiles of wide, short-K GEMMs (the GRU input projection writes 460 MB): a thread owns a row after
        // tcgen05.ld, so the four epilogue warps write each 128 x 32 chunk (+ bias / ReLU / rounding) into a
        // 128B-swizzled shared box and ONE thread hands it to the TMA store engine -- full-line asynchronous writes,
        // no per-row address arithmetic, two boxes in flight.
        for (int c = 0; c < bn; c += 32, ++nbox_out) {
          const uint32_t buf = tma_stage + (nbox_out & 1u) * 16384u;
          if (tid == 0) bulk_wait_group_read<1>();  // the store issued two chunks ago has drained this buffer
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float v[32];
          tmem_ld32(trow + (uint32_t)c, v);
          tmem_ld_wait();
          const int col0 = ntile * bn + c;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 r4 = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (e.bias) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.bias + col0) + j);
              r4.x += b4.x; r4.y += b4.y; r4.z += b4.z; r4.w += b4.w;
            }
            if (e.relu) {
              r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f);
              r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f);
            }
            if (e.round_out) {
              r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
              r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
            }
            st_shared_v4(buf + swz128((uint32_t)(warp * 32 + lane), (uint32_t)j), r4.x, r4.y, r4.z, r4.w);
          }
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (tid == 0) {
            tma_store_2d(&tmC, buf, col0, mtile * kTileM);
            bulk_commit_group();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
        continue;
      }
      if (p.epi_coalesce) {
        // thread = row after tcgen05.ld; a direct store would touch 32 different 128-byte lines per
        // instruction.  The 32 x 32 chunk goes through a swizzled smem tile so that 8 lanes cover
        // one 128-byte row segment: 4 lines per store instruction, same for the mask / add loads.
        const uint32_t stg = stage_base + (uint32_t)warp * 4352u;
        const uint32_t srow = stg + 4096u;
        asm volatile("st.shared.b64 [%0], %1;" ::"r"(srow + (uint32_t)lane * 8u), "l"(m < g.M ? orow : -1LL) : "memory");
        const int chunk = lane & 7, rsub = lane >> 3;
        for (int c = 0; c < bn; c += 32) {
          float v[32];
          tmem_ld32(trow + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            st_shared_v4(stg + swz128((uint32_t)lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          const int col = ntile * bn + c + chunk * 4;
          if (col < e.ncols) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e.bias) b4 = *reinterpret_cast<const float4*>(e.bias + col);
#pragma unroll
            for (int i2 = 0; i2 < 8; ++i2) {
              const int row = i2 * 4 + rsub;
              long long orw;
              asm volatile("ld.shared.b64 %0, [%1];" : "=l"(orw) : "r"(srow + (uint32_t)row * 8u));
              if (orw < 0) continue;
              float4 r4 = ld_shared_v4(stg + swz128((uint32_t)row, chunk));
              r4.x += b4.x; r4.y += b4.y; r4.z += b4.z; r4.w += b4.w;
              if (e.addsrc) {
                const float4 a4 = *reinterpret_cast<const float4*>(e.addsrc + orw * e.lda + col);
                r4.x += a4.x; r4.y += a4.y; r4.z += a4.z; r4.w += a4.w;
              }
              if (e.relu) {
                r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f);
                r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f);
              }
              if (e.mask) {
                const float4 k4 = *reinterpret_cast<const float4*>(e.mask + orw * e.ldm + col);
                r4.x = k4.x > 0.f ? r4.x : 0.f; r4.y = k4.y > 0.f ? r4.y : 0.f;
                r4.z = k4.z > 0.f ? r4.z : 0.f; r4.w = k4.w > 0.f ? r4.w : 0.f;
              }
              if (e.round_out) {
                r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
                r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
              }
              *reinterpret_cast<float4*>(e.out + orw * e.ldo + col) = r4;
            }
          }
          __syncwarp();
        }
      } else if (H16 && (e.out_kind != 0 || e.mask_kind != 0 || e.out_scale != nullptr)) {
        // 16-bit storage: a thread owns 32 consecutive columns of its row = 64 bytes = four 16-byte stores
        const float osc = e.out_scale ? __ldg(e.out_scale) : 1.f;
        for (int c = 0; c < bn; c += 32) {
          float v[32];
          tmem_ld32(trow + (uint32_t)c, v);
          tmem_ld_wait();
          const int col0 = ntile * bn + c;
          if (m < g.M && col0 < e.ncols) {
            if (e.out_scale) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= osc;
            }
            if (e.bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b4 = *reinterpret_cast<const float4*>(e.bias + col0 + j);
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
            if (e.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (e.mask) {
              const uint32_t bits = mbits[c >> 5];
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (!((bits >> j) & 1u)) v[j] = 0.f;
            }
            if (e.out_kind == 0) {
              float* o = e.out + orow * e.ldo + col0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float4 r4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                if (e.round_out) {
                  r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
                  r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
                }
                *reinterpret_cast<float4*>(o + j) = r4;
              }
            } else {
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(e.out) + orow * e.ldo + col0);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                uint4 w4;
                if (e.out_kind == 1) {
                  w4.x = pack_f16x2(v[q4 * 8], v[q4 * 8 + 1]); w4.y = pack_f16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                  w4.z = pack_f16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); w4.w = pack_f16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                } else {
                  w4.x = pack_bf16x2(v[q4 * 8], v[q4 * 8 + 1]); w4.y = pack_bf16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                  w4.z = pack_bf16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); w4.w = pack_bf16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                }
                o[q4] = w4;
              }
            }
          }
        }
      } else
      for (int c = 0; c < bn; c += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)c, v);
        tmem_ld_wait();
        const int col0 = ntile * bn + c;
        if (m < g.M && col0 < e.ncols) {
          float* o = e.out + orow * e.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 r4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (e.bias) {
              const float4 b4 = *reinterpret_cast<const float4*>(e.bias + col0 + j);
              r4.x += b4.x; r4.y += b4.y; r4.z += b4.z; r4.w += b4.w;
            }
            if (e.addsrc) {
              const float4 a4 = *reinterpret_cast<const float4*>(e.addsrc + orow * e.lda + col0 + j);
              r4.x += a4.x; r4.y += a4.y; r4.z += a4.z; r4.w += a4.w;
            }
            if (e.relu) {
              r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f);
              r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f);
            }
            if (e.mask) {
              const uint32_t bits = mbits[c >> 5] >> j;
              r4.x = (bits & 1u) ? r4.x : 0.f; r4.y = (bits & 2u) ? r4.y : 0.f;
              r4.z = (bits & 4u) ? r4.z : 0.f; r4.w = (bits & 8u) ? r4.w : 0.f;
            }
            if (e.round_out) {
              r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
              r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
            }
            *reinterpret_cast<float4*>(o + j) = r4;
          }
        }
      }
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  if (p.epi_tma && tid == 0) bulk_wait_group<0>();  // every output box has left shared memory and is written
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, acc_cols * (uint32_t)nacc);
  }
}

}  // namespace var
