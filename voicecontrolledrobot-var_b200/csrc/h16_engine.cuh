// 16-bit-operand weight gradient for the f16 conv region (see gemm_persist.cuh, H16, for the forward /
// dgrad side): dW[o, k] += sum_pix X[pix@tap, c] * dY[pix, o] with X (f16 NHWC activations) and dY (f16,
// scaled by the region's power-of-two gradient scale) both MN-major operands of kind::f16 MMAs
// (M = 128 k-rows = two taps x 64 channels, N = Cout, K = 16 pixels per instruction).  A pipeline stage
// is one block of p.pb (32 ... 128) pixels: two im2col-mode TMA boxes of X^T (64 channels x pb pixels,
// one per tap group) + one box of dY per 64 output channels -- half the bytes of the tf32 kernel
// (tc_wgrad_tma_kernel) for the same MACs, and with 128-pixel boxes a quarter of its TMA instructions
// (32-pixel boxes reached only ~10 TB/s of L2 -> smem fill; the forward kernel's 16 KB boxes reach 18).
// Bias gradient: when K is not a multiple of 128 the last k tile has an unused 64-row group; it is
// filled with ones once, so row K of the accumulator is colsum(dY) -- no separate column-sum pass.
#pragma once
#include "tc_engine.cuh"

namespace var {

struct WgradH16Params {
  int M, K, cout, kpad;      // cout: output channels of ONE CTA (slab blockIdx.z covers channels [z * cout, (z+1) * cout));
                             // kpad: row pitch (floats) of the fp32 gradient dw[.][kpad]
  int a_tiled;               // 1: X is a plain [M, K] matrix (2-D tensor map, box {64 channels, pb rows}), not im2col
  float* dw;
  float* db;                 // nullable; only written when ones_ktile >= 0
  const float* inv_scale;    // device scalar: results are multiplied by it (nullptr = 1)
  int pix_per_cta;           // multiple of pb
  int stages, pb;            // pipeline stages; pixels per stage (multiple of 16, <= 256 by the TMA box limit)
  int ones_ktile;            // k tile whose second 64-row group is free (-1: none)
  int P, Q, cpb;             // output extents; 64-channel chunks per tap
  int base_w, base_h, step_w, step_h;
  uint8_t tap_w[kMaxTaps], tap_h[kMaxTaps];
};

__host__ __device__ inline size_t wgrad_h16_smem_bytes(int cout, int stages, int pb) {
  return (size_t)stages * (size_t)pb * 128 * (2 + (size_t)(cout / 64)) + 1024 + 256;
}

__global__ void __launch_bounds__(160)
tc_wgrad_h16_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                    const __grid_constant__ WgradH16Params p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, pb = p.pb;
  const int bgroups = p.cout >> 6;
  const uint32_t grp = (uint32_t)pb * 128u;  // bytes of one 64-wide MN group of a stage
  const uint32_t stageA = 2u * grp, stageB = (uint32_t)bgroups * grp;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)stages * stageA;
  const uint32_t bars = sB + (uint32_t)stages * stageB;
  auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  const uint32_t tfull_bar = bars + (uint32_t)(2 * stages) * 8u;
  const uint32_t tslot = tfull_bar + 8u;

  const int ktile = blockIdx.x;
  const int col0 = blockIdx.z * p.cout;  // first output channel of this CTA's slab
  const int pix0 = blockIdx.y * p.pix_per_cta;
  const int pix1 = min(pix0 + p.pix_per_cta, p.M);
  const int num_kb = (pix1 - pix0 + pb - 1) / pb;
  const int kgroups = min(2, (p.K - ktile * 128 + 63) / 64);
  const bool ones = p.db != nullptr && ktile == p.ones_ktile && kgroups == 1;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  if (ones) {
    // second 64-row group of every X^T block := 1.0 (f16 0x3C00); TMA never writes it in this k tile
    for (uint32_t blk = 0; blk < (uint32_t)stages; ++blk)
      for (uint32_t i = tid; i < grp / 16u; i += 160u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + blk * stageA + grp + i * 16u), "r"(0x3C003C00u) : "memory");
    fence_proxy_async_smem();
  }
  const uint32_t ncols = (uint32_t)tmem_cols_for(p.cout);
  if (warp == 4) tmem_alloc(tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  if (num_kb > 0) {
    if (warp < 4) {
      // lane 0 of warp w loads X^T group w (w < kgroups) and dY groups w, w+4, ...
      if (lane == 0 && (warp == 0 || warp < kgroups || warp < bgroups)) {
        const int gq = warp;
        const int kbg = ktile * 2 + gq;
        const int tap = kbg / p.cpb;
        const int c0 = (kbg - tap * p.cpb) << 6;
        const int pq = p.P * p.Q;
        int n = pix0 / pq;
        const int rem = pix0 - n * pq;
        int pp = rem / p.Q, qq = rem - pp * p.Q;
        int st = 0, ph = 0;
        for (int it = 0; it < num_kb; ++it) {
          mbar_wait(empty_bar(st), (uint32_t)(ph ^ 1));
          if (warp == 0)
            mbar_arrive_expect_tx(full_bar(st), (uint32_t)kgroups * grp + stageB);
          {
            const int m = pix0 + it * pb;
            const uint32_t dA = sA + (uint32_t)st * stageA;
            const uint32_t dB = sB + (uint32_t)st * stageB;
            if (gq < kgroups) {
              if (p.a_tiled)
                tma_load_2d(dA + (uint32_t)gq * grp, &tmX, full_bar(st), kbg << 6, m);
              else
                tma_load_im2col_4d(dA + (uint32_t)gq * grp, &tmX, full_bar(st), c0, qq * p.step_w + p.base_w,
                                   pp * p.step_h + p.base_h, n, p.tap_w[tap], p.tap_h[tap]);
            }
            if (!p.a_tiled) {
              qq += pb;
              while (qq >= p.Q) { qq -= p.Q; ++pp; }
              while (pp >= p.P) { pp -= p.P; ++n; }
            }
            for (int bg = warp; bg < bgroups; bg += 4)
              tma_load_2d(dB + (uint32_t)bg * grp, &tmDY, full_bar(st), col0 + bg * 64, m);
          }
          if (++st == stages) { st = 0; ph ^= 1; }
        }
      }
      __syncwarp();
      // ---- epilogue: row = k index, columns = output channel
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
      const int k = ktile * 128 + warp * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c = 0; c < p.cout; c += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)c, v);
        tmem_ld_wait();
        if (k < p.K) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(p.dw + (long long)(col0 + c + j) * p.kpad + k, v[j] * inv);
        } else if (ones && k == p.K) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(p.db + col0 + c + j, v[j] * inv);
        }
      }
      tc_fence_before();
    } else {
      const uint32_t idesc = make_idesc_h16(p.cout, 0, 0, 1, 1);
      // MN-major, plain 128B swizzle: 8-pixel atoms 1024 B apart (SBO), 64-wide MN groups one group apart (LBO)
      const uint64_t adesc0 = make_smem_desc(sA, grp, 1024u, 2), bdesc0 = make_smem_desc(sB, grp, 1024u, 2);
      const int nmma = pb >> 4;  // 16 pixels (2048 B) per MMA
      int st = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {  // converged warp, one elected lane issues
        mbar_wait(full_bar(st), (uint32_t)ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t ad0 = adesc0 + (uint64_t)(((uint32_t)st * stageA) >> 4);
          const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)st * stageB) >> 4);
          for (int j = 0; j < nmma; ++j)
            umma_f16(tmem_base, ad0 + (uint64_t)(j * 128), bd0 + (uint64_t)(j * 128), idesc, (uint32_t)((kb | j) != 0));
          umma_commit(empty_bar(st));
          if (kb == num_kb - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
        if (++st == stages) { st = 0; ph ^= 1; }
      }
      __syncwarp();
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

}  // namespace var
