// Im2col-free ("halo") forward convolution for the strided f16 convs of the 16-bit region (64 -> 64 channels,
// stride 2: iTHOR cnn.2 11x5 and cnn.4 7x3; models/pretext/ai2thor_pretext_model.py:28-31).
//
// The im2col kernel (gemm_persist.cuh) re-fetches a 16 KB A tile for every one of the R*S taps: 55 x 16 KB + 55 x 8 KB of
// weights per 128 output pixels, which pins it to the L2 -> SM fill rate (17.8 TB/s chip-wide, 38 % tensor pipe).  Here
// the input footprint of a tile of TP output rows is staged ONCE, as four parity planes
//     plane(a, b)[i][j] = x[2 i + a][2 j + b]          (one strided 4-D TILED TMA load each; OOB = zero padding)
// of dense 128-byte rows (64 channels) in the 128B swizzle.  With output pixels numbered m = (p - p0) * WP + q over a
// row pitch WP = Q + (column taps' plane span) -- q >= Q are phantom pixels whose results are dropped -- the A operand of
// tap (r, s) is the SAME plane image read from a shifted start row:
//     A_tap[m] = plane(rp, sp)[(p + rj) * WP + q + sj] = plane rows  m + shift(tap),
// and tcgen05.mma accepts any 128-byte-aligned start inside a swizzled image (csrc/halo_probe.cu).  Fill per tile drops
// from 1.3 MB to 92 KB of input + 440 KB of weights, at the price of (WP - Q) / WP wasted MMA rows.
// An M = 128 operand that starts `shift` rows into a plane reads up to 128 - TP * WP + (column span) < 66 rows (8.3 KB) past the
// plane's last row: phantom output rows only, and always inside the CTA's allocation (the next plane slot or the weight ring,
// which is at least 16 KB).
// CTA = 224 threads, persistent over tiles: warp 6 streams the parity planes through a ring of slots, warp 5 the weight
// ring, warp 4 issues the MMAs (two TMEM accumulators), warps 0-3 are the epilogue (bias / ReLU / rounding, f16 or fp32 rows).
#pragma once
#include "tc_engine.cuh"

namespace var {

// plane slots (p.slots): ring of staged parity planes (a tile uses 4 in a row; the next tile's planes stream in while the
// last planes of this tile are still being multiplied); 3 slots + 4 weight stages = 2 CTAs/SM, 2 + 3 = 3 CTAs/SM

struct HaloParams {
  const float* bias;      // [64] or nullptr
  void* out;              // [N, P, Q, 64] f16 (out_kind 1) or fp32 (0)
  int out_kind, relu, round_out;
  int N, P, Q;            // output extents
  int TP, WP, TPI;        // output rows per tile, row pitch of the M numbering, tiles per image
  int ntaps, stages;      // taps; weight ring depth (w_resident: == ntaps, every tap's box is loaded once per CTA)
  int nplanes, step_h;    // 4 parity planes / tile row step 2 (stride-2 convs) or 1 plane / step 1 (stride-1 convs)
  int bn, kelems;         // output channels (MMA N); elements per 128-byte operand row (64 f16 / 32 tf32)
  int w_resident;
  int slots;              // plane ring depth
  int h_start, w_start;   // tensor coordinates of plane (0, 0) element [0][0] for tile row p0 = 0: 2 * rjmin, 2 * sjmin
  int box_rows, box_cols; // plane rows / cols staged per tile (in plane units); box_cols == WP
  uint32_t plane_stride;  // bytes of one plane slot (multiple of 1024)
  // taps are processed plane by plane (plane = rp * 2 + sp): taps [plane_begin[pl], plane_begin[pl + 1]) read plane pl
  int plane_begin[5];
  uint16_t tap_shift[kMaxTaps];  // start row of the tap's A operand inside its plane: (rj - rjmin) * WP + (sj - sjmin)
  uint16_t tap_wcol[kMaxTaps];   // 64-column block of the packed weights [64][kpad] (= r * S + s)
};

__host__ __device__ inline size_t halo_smem_bytes(uint32_t plane_stride, int stages, int slots, int bn = 64) {
  return (size_t)slots * plane_stride + (size_t)stages * (size_t)bn * 128 + 1024 /*align*/ + 512 /*barriers*/;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <bool TF32>
__global__ void __launch_bounds__(224, 3)
halo_conv_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, nslots = p.slots;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sH = base;                                          // nslots plane slots
  const uint32_t wbox = (uint32_t)p.bn * 128u;                   // one tap's weight box: bn rows x 128 B
  const uint32_t sB = base + (uint32_t)nslots * p.plane_stride;  // weight ring: stages x wbox
  const uint32_t bars = sB + (uint32_t)stages * wbox;
  auto bfull = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto bempty = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  auto pfull = [&](int b) { return bars + (uint32_t)(2 * stages + b) * 8u; };
  auto pempty = [&](int b) { return bars + (uint32_t)(2 * stages + nslots + b) * 8u; };
  auto tfull = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * nslots + a) * 8u; };
  auto tempty = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * nslots + 2 + a) * 8u; };
  const uint32_t tslot = bars + (uint32_t)(2 * stages + 2 * nslots + 4) * 8u;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int b = 0; b < nslots; ++b) { mbar_init(pfull(b), 1); mbar_init(pempty(b), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  const uint32_t tcols = p.bn <= 16 ? 32u : 2u * (uint32_t)p.bn;  // two bn-column fp32 accumulators (bn = 32, 64, 128)
  if (warp == 4) tmem_alloc(tslot, tcols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  const int total = p.N * p.TPI;
  if (warp == 6) {
    // ===================== TMA producer of the parity planes (its own warp: the two producers wait on barriers with
    // very different rhythms, and two spinning lanes of one warp starve each other) =====================
    {  // converged warp, one elected lane issues (see elect_one_sync)
      const uint32_t plane_bytes = (uint32_t)p.box_rows * (uint32_t)p.box_cols * 128u;
      int slot = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / p.TPI, p0 = (tile - n * p.TPI) * p.TP;
        for (int pl = 0; pl < p.nplanes; ++pl) {  // plane = rp * 2 + sp
          mbar_wait(pempty(slot), (uint32_t)(ph ^ 1));
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(pfull(slot), plane_bytes);
            tma_load_4d(sH + (uint32_t)slot * p.plane_stride, &tmX, pfull(slot), 0, p.w_start + (pl & 1),
                        p.h_start + p.step_h * p0 + (pl >> 1), n);
          }
          __syncwarp();
          if (++slot == nslots) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== TMA producer of the weight ring =====================
    {
      int st = 0, ph = 0;
      if (p.w_resident) {
        if ((int)blockIdx.x < total && elect_one_sync())
          for (int t = 0; t < p.ntaps; ++t) {
            mbar_arrive_expect_tx(bfull(t), wbox);
            tma_load_2d(sB + (uint32_t)t * wbox, &tmW, bfull(t), (int)p.tap_wcol[t] * p.kelems, 0);
          }
        __syncwarp();
      } else
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x)
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(bempty(st), (uint32_t)(ph ^ 1));
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(bfull(st), wbox);
            tma_load_2d(sB + (uint32_t)st * wbox, &tmW, bfull(st), (int)p.tap_wcol[t] * p.kelems, 0);
          }
          __syncwarp();
          if (++st == stages) { st = 0; ph ^= 1; }
        }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer: the whole warp runs the loop, one elected lane issues =====================
    {
      const uint32_t idesc = TF32 ? make_idesc_tf32(p.bn, 0, 0) : make_idesc_h16(p.bn, 0, 0, 0, 0);
      const uint64_t bdesc0 = make_smem_desc(sB, 16u, 1024u);
      const uint64_t adesc0 = make_smem_desc(sH, 16u, 1024u);
      const bool resident = p.w_resident != 0;
      int st = 0, ph = 0, slot = 0, sph = 0, i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int acc = i & 1, use = i >> 1;
        mbar_wait(tempty(acc), (uint32_t)((use & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * (uint32_t)p.bn;
        int t = 0;
        for (int pl = 0; pl < p.nplanes; ++pl) {
          mbar_wait(pfull(slot), (uint32_t)sph);
          tc_fence_after();
          const uint64_t aplane = adesc0 + (uint64_t)(((uint32_t)slot * p.plane_stride) >> 4);
          for (; t < p.plane_begin[pl + 1]; ++t) {
            if (resident) { st = t; ph = 0; }  // phase 0 of a resident box completes once and stays complete
            mbar_wait(bfull(st), (uint32_t)ph);
            tc_fence_after();
            const uint64_t ad0 = aplane + (uint64_t)((uint32_t)p.tap_shift[t] * 8u);  // rows x 128 B >> 4
            const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)st * wbox) >> 4);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {  // 32 bytes of K per MMA: 16 f16 or 8 tf32
                if constexpr (TF32) umma_tf32(d_tmem, ad0 + (uint64_t)(j * 2), bd0 + (uint64_t)(j * 2), idesc, (uint32_t)((t | j) != 0));
                else umma_f16(d_tmem, ad0 + (uint64_t)(j * 2), bd0 + (uint64_t)(j * 2), idesc, (uint32_t)((t | j) != 0));
              }
              if (!resident) umma_commit(bempty(st));
            }
            __syncwarp();
            if (!resident && ++st == stages) { st = 0; ph ^= 1; }
          }
          if (elect_one_sync()) umma_commit(pempty(slot));  // every MMA reading this plane was issued before this commit
          __syncwarp();
          if (++slot == nslots) { slot = 0; sph ^= 1; }
        }
        if (elect_one_sync()) umma_commit(tfull(acc));
        __syncwarp();
      }
    }
    tc_fence_before();
  } else {
    // ===================== epilogue warps 0-3 =====================
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
      const int acc = i & 1, use = i >> 1;
      const int n = tile / p.TPI, p0 = (tile - n * p.TPI) * p.TP;
      const int m = warp * 32 + lane;
      const int pp = m / p.WP, q = m - pp * p.WP;
      const bool valid = pp < p.TP && q < p.Q && p0 + pp < p.P;
      const long long orow = ((long long)n * p.P + p0 + pp) * p.Q + q;
      mbar_wait(tfull(acc), (uint32_t)(use & 1));
      tc_fence_after();
      const uint32_t trow = tmem_base + (uint32_t)acc * (uint32_t)p.bn + ((uint32_t)(warp * 32) << 16);
      for (int c = 0; c < p.bn; c += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)c, v);
        tmem_ld_wait();
        if (valid) {
          if (p.bias) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.out_kind == 0) {
            float* o = reinterpret_cast<float*>(p.out) + orow * p.bn + c;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 r4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              if (p.round_out) {
                r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
                r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
              }
              *reinterpret_cast<float4*>(o + j) = r4;
            }
          } else {
            uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + orow * p.bn + c);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              uint4 w4;
              w4.x = pack_f16x2(v[q4 * 8], v[q4 * 8 + 1]); w4.y = pack_f16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
              w4.z = pack_f16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); w4.w = pack_f16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
              o[q4] = w4;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(acc));
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tcols);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Im2col-free data gradient of the same convs.  dx[2 i + a][2 j + b] = sum over the taps of stride-parity class (a, b) of
// dy[i + dr][j + ds] . W[r, s]^T with dr = (a + ph - r) / 2, ds = (b + pw - s) / 2: every class is a stride-1 conv over
// the SAME dy rows, so ONE unstrided dy tile (TP + span rows x WP cols of 128-byte pixels, zero-filled outside the image)
// is staged per tile and serves all R * S taps of the four classes (the im2col kernel re-fetches a 16 KB A tile per tap
// and runs one launch per class).  Output pixels of a class are numbered m = ii * WP + j (j >= W2 are phantoms); the A
// operand of a tap starts (dr - drmin) * WP + (ds - dsmin) rows into the tile.  B is the packed forward weight box of the
// tap read MN-major (n = ci contiguous, k = co rows).  Accumulator use = (tile, class), two TMEM accumulators in turn.

struct HaloDgradParams {
  void* out;               // dx [N, H, W, 64] fp32 (out_kind 0) or f16 (1)
  const void* mask;        // forward activation of the previous layer (ReLU backward) or nullptr
  const float* out_scale;  // device scalar multiplied in (1 / gradient scale) or nullptr
  int out_kind, mask_kind, round_out;  // mask_kind 0 fp32, 1 f16
  int N, H, W;             // dx extents
  int TP, WP, TPI;         // half-resolution rows per tile, row pitch of the M numbering, tiles per image
  int stages, slots;       // weight ring depth, dy tile ring depth
  int h_start, w_start;    // dy coordinates of tile element [0][0] for tile row i0 = 0: drmin, dsmin
  int box_rows, box_cols;
  uint32_t plane_stride;
  int class_begin[5];      // taps [class_begin[c], class_begin[c + 1]) belong to class c = a * 2 + b
  uint16_t tap_shift[kMaxTaps];
  uint16_t tap_wcol[kMaxTaps];
};

__host__ __device__ inline size_t halo_dgrad_smem_bytes(uint32_t plane_stride, int stages, int slots) {
  return (size_t)slots * plane_stride + (size_t)stages * 8192 + 1024 /*align*/ + 512 /*barriers*/;
}

__global__ void __launch_bounds__(224, 3)
halo_conv_dgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ HaloDgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, nslots = p.slots;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sH = base;
  const uint32_t sB = base + (uint32_t)nslots * p.plane_stride;
  const uint32_t bars = sB + (uint32_t)stages * 8192u;
  auto bfull = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto bempty = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  auto pfull = [&](int b) { return bars + (uint32_t)(2 * stages + b) * 8u; };
  auto pempty = [&](int b) { return bars + (uint32_t)(2 * stages + nslots + b) * 8u; };
  auto tfull = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * nslots + a) * 8u; };
  auto tempty = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * nslots + 2 + a) * 8u; };
  const uint32_t tslot = bars + (uint32_t)(2 * stages + 2 * nslots + 4) * 8u;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int b = 0; b < nslots; ++b) { mbar_init(pfull(b), 1); mbar_init(pempty(b), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 4) tmem_alloc(tslot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  const int total = p.N * p.TPI;
  const int ntaps = p.class_begin[4];
  if (warp == 6) {
    // ===================== TMA producer of the dy tiles =====================
    {
      const uint32_t plane_bytes = (uint32_t)p.box_rows * (uint32_t)p.box_cols * 128u;
      int slot = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / p.TPI, i0 = (tile - n * p.TPI) * p.TP;
        mbar_wait(pempty(slot), (uint32_t)(ph ^ 1));
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(pfull(slot), plane_bytes);
          tma_load_4d(sH + (uint32_t)slot * p.plane_stride, &tmDY, pfull(slot), 0, p.w_start, p.h_start + i0, n);
        }
        __syncwarp();
        if (++slot == nslots) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ===================== TMA producer of the weight ring =====================
    {
      int st = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x)
        for (int t = 0; t < ntaps; ++t) {
          mbar_wait(bempty(st), (uint32_t)(ph ^ 1));
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(bfull(st), 8192u);
            tma_load_2d(sB + (uint32_t)st * 8192u, &tmW, bfull(st), (int)p.tap_wcol[t] * 64, 0);
          }
          __syncwarp();
          if (++st == stages) { st = 0; ph ^= 1; }
        }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer: converged warp, one elected lane issues =====================
    {
      const uint32_t idesc = make_idesc_h16(64, 0, 0, 0, 1);
      // B = the packed forward weight box read MN-major: 8-k-row atoms 1024 B apart, 16 k rows (2048 B) per MMA
      const uint64_t bdesc0 = make_smem_desc(sB, 8192u, 1024u, 2);
      constexpr int bstep = 128;
      const uint64_t adesc0 = make_smem_desc(sH, 16u, 1024u);
      int st = 0, ph = 0, slot = 0, sph = 0, i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        mbar_wait(pfull(slot), (uint32_t)sph);
        tc_fence_after();
        const uint64_t aplane = adesc0 + (uint64_t)(((uint32_t)slot * p.plane_stride) >> 4);
        int t = 0;
        for (int cls = 0; cls < 4; ++cls, ++i) {
          const int acc = i & 1, use = i >> 1;
          mbar_wait(tempty(acc), (uint32_t)((use & 1) ^ 1));
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 64u;
          const int t0 = t;
          for (; t < p.class_begin[cls + 1]; ++t) {
            mbar_wait(bfull(st), (uint32_t)ph);
            tc_fence_after();
            const uint64_t ad0 = aplane + (uint64_t)((uint32_t)p.tap_shift[t] * 8u);
            const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)st * 8192u) >> 4);
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 4; ++j)
                umma_f16(d_tmem, ad0 + (uint64_t)(j * 2), bd0 + (uint64_t)(j * bstep), idesc, (uint32_t)((t != t0) | (j != 0)));
              umma_commit(bempty(st));
            }
            __syncwarp();
            if (++st == stages) { st = 0; ph ^= 1; }
          }
          if (elect_one_sync()) umma_commit(tfull(acc));
          __syncwarp();
        }
        if (elect_one_sync()) umma_commit(pempty(slot));
        __syncwarp();
        if (++slot == nslots) { slot = 0; sph ^= 1; }
      }
    }
    tc_fence_before();
  } else {
    // ===================== epilogue warps 0-3 =====================
    const float osc = p.out_scale ? __ldg(p.out_scale) : 1.f;
    const int m = warp * 32 + lane;
    const int ii = m / p.WP, j = m - ii * p.WP;
    int i = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int n = tile / p.TPI, i0 = (tile - n * p.TPI) * p.TP;
      for (int cls = 0; cls < 4; ++cls, ++i) {
        const int acc = i & 1, use = i >> 1;
        const int h = 2 * (i0 + ii) + (cls >> 1), w = 2 * j + (cls & 1);
        const bool valid = ii < p.TP && h < p.H && w < p.W;
        const long long orow = ((long long)n * p.H + h) * p.W + w;
        // ReLU-backward mask bits of this thread's row, fetched before the accumulator wait
        uint32_t mbits[2] = {0xFFFFFFFFu, 0xFFFFFFFFu};
        if (valid && p.mask) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t bits = 0u;
            if (p.mask_kind == 1) {
              const uint4* mk = reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(p.mask) + orow * 64 + q * 32);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const uint4 w4 = __ldg(mk + q4);
                const uint32_t ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const uint32_t lo = ww[u] & 0xFFFFu, hi = ww[u] >> 16;
                  bits |= (uint32_t)(lo != 0u && lo < 0x8000u) << (q4 * 8 + u * 2);
                  bits |= (uint32_t)(hi != 0u && hi < 0x8000u) << (q4 * 8 + u * 2 + 1);
                }
              }
            } else {
              const float4* mk = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.mask) + orow * 64 + q * 32);
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                const float4 k4 = __ldg(mk + jj);
                bits |= (uint32_t)(k4.x > 0.f) << (4 * jj) | (uint32_t)(k4.y > 0.f) << (4 * jj + 1) |
                        (uint32_t)(k4.z > 0.f) << (4 * jj + 2) | (uint32_t)(k4.w > 0.f) << (4 * jj + 3);
              }
            }
            mbits[q] = bits;
          }
        }
        mbar_wait(tfull(acc), (uint32_t)(use & 1));
        tc_fence_after();
        const uint32_t trow = tmem_base + (uint32_t)acc * 64u + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          float v[32];
          tmem_ld32(trow + (uint32_t)c, v);
          tmem_ld_wait();
          if (valid) {
            const uint32_t bits = mbits[c >> 5];
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = ((bits >> jj) & 1u) ? v[jj] * osc : 0.f;
            if (p.out_kind == 0) {
              float* o = reinterpret_cast<float*>(p.out) + orow * 64 + c;
#pragma unroll
              for (int jj = 0; jj < 32; jj += 4) {
                float4 r4 = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
                if (p.round_out) {
                  r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
                  r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
                }
                *reinterpret_cast<float4*>(o + jj) = r4;
              }
            } else {
              uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.out) + orow * 64 + c);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                uint4 w4;
                w4.x = pack_f16x2(v[q4 * 8], v[q4 * 8 + 1]); w4.y = pack_f16x2(v[q4 * 8 + 2], v[q4 * 8 + 3]);
                w4.z = pack_f16x2(v[q4 * 8 + 4], v[q4 * 8 + 5]); w4.w = pack_f16x2(v[q4 * 8 + 6], v[q4 * 8 + 7]);
                o[q4] = w4;
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(acc));
      }
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Im2col-free weight gradient of the same convs: dW[co][tap][ci] += sum_m X_tap[m][ci] * dY[m][co] over the pixel
// numbering m = pp * WP + q of the forward kernel above.  Both operands are MN-major views of staged images: B = the dY
// tile (rows = m, 128 bytes = 64 co; phantom columns q >= Q and rows past the image are zero-filled by TMA, so phantom
// pixels contribute nothing), A = the parity plane of a tap PAIR read from a shifted start row -- the two 64-channel row
// groups of an M = 128 instruction are the two taps of the pair, LBO = (shift1 - shift0) * 128 bytes.  K = 16 pixels per
// MMA.  A CTA owns a group of <= 8 pairs of ONE plane (64 TMEM columns each) over a range of tiles; per tile it stages that
// plane + the dY tile once (the im2col kernel re-fetches 24 KB per 64 pixels per tap pair) and issues, for every 16-pixel
// step, one MMA per pair -- consecutive MMAs go to different accumulators.  fp32 partial sums are added to dw with
// red.global at the end.  Shared memory is zeroed once: reads past a plane (start shift + rounding of the pixel count to
// 16) must meet finite values, and the dY rows behind TP * WP must stay zero.
struct HaloWgradParams {
  float* dw;               // [64][kpad] fp32, accumulated
  float* db;               // [64] bias gradient (colsum of dY), accumulated by the CTAs of group 0; nullable
  const float* inv_scale;  // device scalar or nullptr
  int kpad;
  int N, P, Q;
  int TP, WP, TPI;         // output rows per tile, pitch, tiles per image
  int nk;                  // 16-pixel MMA steps per tile = ceil(TP * WP / 16)
  int slots;
  int h_start, w_start, box_rows, box_cols;
  uint32_t plane_bytes, dy_bytes;       // TMA transaction sizes
  uint32_t plane_region, slot_stride;   // dY tile offset inside a slot; bytes per slot (multiples of 1024)
  int ngroups;
  int g_cta_begin[17];     // CTAs [g_cta_begin[g], g_cta_begin[g + 1]) work on group g
  uint8_t g_plane[16];     // parity plane (rp * 2 + sp) of the group
  uint8_t g_pair_begin[17];
  uint16_t pair_shift[32]; // start row of the pair's first tap
  uint16_t pair_lbo[32];   // rows between the two taps (0: single tap, second row group ignored)
  uint8_t pair_tap0[32], pair_tap1[32];  // packed-weight tap index r * S + s (tap1 = 255: none)
};

__global__ void __launch_bounds__(192)
halo_conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                       const __grid_constant__ HaloWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + (uint32_t)p.slots * p.slot_stride;
  auto full = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty = [&](int s) { return bars + (uint32_t)(p.slots + s) * 8u; };
  const uint32_t tfull = bars + (uint32_t)(2 * p.slots) * 8u;
  const uint32_t tslot = tfull + 8u;

  int g = 0;
  while (g + 1 < p.ngroups && (int)blockIdx.x >= p.g_cta_begin[g + 1]) ++g;
  const int ncta = p.g_cta_begin[g + 1] - p.g_cta_begin[g], cidx = (int)blockIdx.x - p.g_cta_begin[g];
  const int total = p.N * p.TPI;
  const int tile0 = (int)(((long long)total * cidx) / ncta), tile1 = (int)(((long long)total * (cidx + 1)) / ncta);
  const int pr0 = p.g_pair_begin[g], npair = p.g_pair_begin[g + 1] - pr0;
  const int plane = p.g_plane[g];
  const bool do_db = p.db != nullptr && g == 0;  // group 0 also column-sums the dY tiles it stages

  for (uint32_t i = tid; i < ((uint32_t)p.slots * p.slot_stride) / 16u; i += 192u)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16u), "r"(0u) : "memory");
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    for (int s = 0; s < p.slots; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), do_db ? 2 : 1); }
    mbar_init(tfull, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDY);
  }
  const uint32_t ncols = npair <= 1 ? 64u : (npair <= 2 ? 128u : (npair <= 4 ? 256u : 512u));
  if (warp == 4) tmem_alloc(tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  if (tile1 > tile0) {
    if (warp == 5) {
      // ===================== TMA producer: plane + dY tile per slot =====================
      if (lane == 0) {
        int slot = 0, ph = 0;
        for (int tile = tile0; tile < tile1; ++tile) {
          const int n = tile / p.TPI, p0 = (tile - n * p.TPI) * p.TP;
          mbar_wait(empty(slot), (uint32_t)(ph ^ 1));
          mbar_arrive_expect_tx(full(slot), p.plane_bytes + p.dy_bytes);
          const uint32_t dst = base + (uint32_t)slot * p.slot_stride;
          tma_load_4d(dst, &tmX, full(slot), 0, p.w_start + (plane & 1), p.h_start + 2 * p0 + (plane >> 1), n);
          tma_load_4d(dst + p.plane_region, &tmDY, full(slot), 0, 0, p0, n);
          if (++slot == p.slots) { slot = 0; ph ^= 1; }
        }
      }
    } else if (warp == 4) {
      // ===================== MMA issuer: converged warp, one elected lane issues =====================
      {
        const uint32_t idesc = make_idesc_h16(64, 0, 0, 1, 1);
        uint64_t adesc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int pr = pr0 + (q < npair ? q : 0);
          adesc[q] = make_smem_desc(base + (uint32_t)p.pair_shift[pr] * 128u, (uint32_t)p.pair_lbo[pr] * 128u, 1024u, 2);
        }
        const uint64_t bdesc0 = make_smem_desc(base + p.plane_region, 1024u, 1024u, 2);
        int slot = 0, ph = 0;
        for (int tile = tile0; tile < tile1; ++tile) {
          mbar_wait(full(slot), (uint32_t)ph);
          tc_fence_after();
          const uint64_t soff = (uint64_t)(((uint32_t)slot * p.slot_stride) >> 4);
          if (elect_one_sync()) {
            for (int j = 0; j < p.nk; ++j) {
              const uint64_t koff = soff + (uint64_t)(j * 128);  // 16 pixel rows = 2048 B
              const uint32_t accum = (uint32_t)((tile != tile0) | (j != 0));
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (q < npair) umma_f16(tmem_base + (uint32_t)q * 64u, adesc[q] + koff, bdesc0 + koff, idesc, accum);
            }
            umma_commit(empty(slot));
          }
          __syncwarp();
          if (++slot == p.slots) { slot = 0; ph ^= 1; }
        }
        if (elect_one_sync()) umma_commit(tfull);
        __syncwarp();
      }
      tc_fence_before();
    } else {
      // ===================== epilogue warps 0-3: rows = (tap of the pair, ci), columns = co =====================
      const float inv = p.inv_scale ? __ldg(p.inv_scale) : 1.f;
      if (do_db) {
        // bias gradient: thread = (8-channel chunk, row phase); 128-bit reads of the swizzled dY rows while the MMAs run
        const uint32_t chunk = (uint32_t)tid & 7u, rsub = (uint32_t)tid >> 3;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const uint32_t rows = (uint32_t)(p.TP * p.WP);
        int slot = 0, ph = 0;
        for (int tile = tile0; tile < tile1; ++tile) {
          mbar_wait(full(slot), (uint32_t)ph);
          const uint32_t dyb = base + (uint32_t)slot * p.slot_stride + p.plane_region;
          for (uint32_t m = rsub; m < rows; m += 16u) {
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                         : "r"(dyb + m * 128u + ((chunk ^ (m & 7u)) << 4)));
            const uint32_t ww[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              acc[2 * u] += f16_bits_to_f32((uint16_t)(ww[u] & 0xFFFFu));
              acc[2 * u + 1] += f16_bits_to_f32((uint16_t)(ww[u] >> 16));
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (tid == 0) mbar_arrive(empty(slot));
          if (++slot == p.slots) { slot = 0; ph ^= 1; }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // lanes l, l + 8, l + 16, l + 24 hold the same chunk
          float a = acc[u];
          a += __shfl_xor_sync(0xFFFFFFFFu, a, 8);
          a += __shfl_xor_sync(0xFFFFFFFFu, a, 16);
          if (lane < 8) atomicAdd(p.db + chunk * 8u + (uint32_t)u, a * inv);
        }
      }
      mbar_wait(tfull, 0);
      tc_fence_after();
      const int row = warp * 32 + lane;
      for (int q = 0; q < npair; ++q) {
        const int tap = row < 64 ? (int)p.pair_tap0[pr0 + q] : (int)p.pair_tap1[pr0 + q];
        const uint32_t trow = tmem_base + (uint32_t)q * 64u + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
          float v[32];
          tmem_ld32(trow + (uint32_t)c, v);
          tmem_ld_wait();
          if (tap != 255) {
            float* o = p.dw + (long long)c * p.kpad + tap * 64 + (row & 63);
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(o + (long long)j * p.kpad, v[j] * inv);
          }
        }
      }
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
