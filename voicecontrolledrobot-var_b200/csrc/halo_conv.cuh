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
// CTA = 224 threads, persistent over tiles: warp 6 streams the parity planes through a ring of slots, warp 5 the weight
// ring, warp 4 issues the MMAs (two TMEM accumulators), warps 0-3 are the epilogue (bias / ReLU / rounding, f16 or fp32 rows).
#pragma once
#include "tc_engine.cuh"

namespace var {

constexpr int kHaloSlots = 3;  // ring of staged parity planes (a tile uses 4 in a row; the next tile's planes stream
                               // in while the last planes of this tile are still being multiplied)

struct HaloParams {
  const float* bias;      // [64] or nullptr
  void* out;              // [N, P, Q, 64] f16 (out_kind 1) or fp32 (0)
  int out_kind, relu, round_out;
  int N, P, Q;            // output extents
  int TP, WP, TPI;        // output rows per tile, row pitch of the M numbering, tiles per image
  int ntaps, stages;      // taps; weight ring depth
  int h_start, w_start;   // tensor coordinates of plane (0, 0) element [0][0] for tile row p0 = 0: 2 * rjmin, 2 * sjmin
  int box_rows, box_cols; // plane rows / cols staged per tile (in plane units); box_cols == WP
  uint32_t plane_stride;  // bytes of one plane slot (multiple of 1024)
  // taps are processed plane by plane (plane = rp * 2 + sp): taps [plane_begin[pl], plane_begin[pl + 1]) read plane pl
  int plane_begin[5];
  uint16_t tap_shift[kMaxTaps];  // start row of the tap's A operand inside its plane: (rj - rjmin) * WP + (sj - sjmin)
  uint16_t tap_wcol[kMaxTaps];   // 64-column block of the packed weights [64][kpad] (= r * S + s)
};

__host__ __device__ inline size_t halo_smem_bytes(uint32_t plane_stride, int stages) {
  return (size_t)kHaloSlots * plane_stride + (size_t)stages * 8192 + 1024 /*align*/ + 512 /*barriers*/;
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__global__ void __launch_bounds__(224, 2)
halo_conv_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
                     const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sH = base;                                          // kHaloSlots plane slots
  const uint32_t sB = base + (uint32_t)kHaloSlots * p.plane_stride;  // weight ring: stages x 8 KB
  const uint32_t bars = sB + (uint32_t)stages * 8192u;
  auto bfull = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto bempty = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  auto pfull = [&](int b) { return bars + (uint32_t)(2 * stages + b) * 8u; };
  auto pempty = [&](int b) { return bars + (uint32_t)(2 * stages + kHaloSlots + b) * 8u; };
  auto tfull = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * kHaloSlots + a) * 8u; };
  auto tempty = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * kHaloSlots + 2 + a) * 8u; };
  const uint32_t tslot = bars + (uint32_t)(2 * stages + 2 * kHaloSlots + 4) * 8u;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int b = 0; b < kHaloSlots; ++b) { mbar_init(pfull(b), 1); mbar_init(pempty(b), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull(a), 1); mbar_init(tempty(a), 4); }
    mbar_fence_init();
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
  }
  if (warp == 4) tmem_alloc(tslot, 128);  // two 64-column fp32 accumulators
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  const int total = p.N * p.TPI;
  if (warp == 6) {
    // ===================== TMA producer of the parity planes (its own warp: the two producers wait on barriers with
    // very different rhythms, and two spinning lanes of one warp starve each other) =====================
    if (lane == 0) {
      const uint32_t plane_bytes = (uint32_t)p.box_rows * (uint32_t)p.box_cols * 128u;
      int slot = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / p.TPI, p0 = (tile - n * p.TPI) * p.TP;
        for (int pl = 0; pl < 4; ++pl) {  // plane = rp * 2 + sp
          mbar_wait(pempty(slot), (uint32_t)(ph ^ 1));
          mbar_arrive_expect_tx(pfull(slot), plane_bytes);
          tma_load_4d(sH + (uint32_t)slot * p.plane_stride, &tmX, pfull(slot), 0, p.w_start + (pl & 1),
                      p.h_start + 2 * p0 + (pl >> 1), n);
          if (++slot == kHaloSlots) { slot = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    // ===================== TMA producer of the weight ring =====================
    if (lane == 0) {
      int st = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x)
        for (int t = 0; t < p.ntaps; ++t) {
          mbar_wait(bempty(st), (uint32_t)(ph ^ 1));
          mbar_arrive_expect_tx(bfull(st), 8192u);
          tma_load_2d(sB + (uint32_t)st * 8192u, &tmW, bfull(st), (int)p.tap_wcol[t] * 64, 0);
          if (++st == stages) { st = 0; ph ^= 1; }
        }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_h16(64, 0, 0, 0, 0);
      const uint64_t bdesc0 = make_smem_desc(sB, 16u, 1024u);
      const uint64_t adesc0 = make_smem_desc(sH, 16u, 1024u);
      int st = 0, ph = 0, slot = 0, sph = 0, i = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++i) {
        const int acc = i & 1, use = i >> 1;
        mbar_wait(tempty(acc), (uint32_t)((use & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 64u;
        int t = 0;
        for (int pl = 0; pl < 4; ++pl) {
          mbar_wait(pfull(slot), (uint32_t)sph);
          tc_fence_after();
          const uint64_t aplane = adesc0 + (uint64_t)(((uint32_t)slot * p.plane_stride) >> 4);
          for (; t < p.plane_begin[pl + 1]; ++t) {
            mbar_wait(bfull(st), (uint32_t)ph);
            tc_fence_after();
            const uint64_t ad0 = aplane + (uint64_t)((uint32_t)p.tap_shift[t] * 8u);  // rows x 128 B >> 4
            const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)st * 8192u) >> 4);
#pragma unroll
            for (int j = 0; j < 4; ++j) umma_f16(d_tmem, ad0 + (uint64_t)(j * 2), bd0 + (uint64_t)(j * 2), idesc, (uint32_t)((t | j) != 0));
            umma_commit(bempty(st));
            if (++st == stages) { st = 0; ph ^= 1; }
          }
          umma_commit(pempty(slot));  // every MMA reading this plane was issued before this commit
          if (++slot == kHaloSlots) { slot = 0; sph ^= 1; }
        }
        umma_commit(tfull(acc));
      }
    }
    __syncwarp();
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
      const uint32_t trow = tmem_base + (uint32_t)acc * 64u + ((uint32_t)(warp * 32) << 16);
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
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
            float* o = reinterpret_cast<float*>(p.out) + orow * 64 + c;
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
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
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
constexpr int kHaloDgSlots = 2;

struct HaloDgradParams {
  void* out;               // dx [N, H, W, 64] fp32 (out_kind 0) or f16 (1)
  const void* mask;        // forward activation of the previous layer (ReLU backward) or nullptr
  const float* out_scale;  // device scalar multiplied in (1 / gradient scale) or nullptr
  int out_kind, mask_kind, round_out;  // mask_kind 0 fp32, 1 f16
  int N, H, W;             // dx extents
  int TP, WP, TPI;         // half-resolution rows per tile, row pitch of the M numbering, tiles per image
  int stages;              // weight ring depth
  int h_start, w_start;    // dy coordinates of tile element [0][0] for tile row i0 = 0: drmin, dsmin
  int box_rows, box_cols;
  uint32_t plane_stride;
  int class_begin[5];      // taps [class_begin[c], class_begin[c + 1]) belong to class c = a * 2 + b
  uint16_t tap_shift[kMaxTaps];
  uint16_t tap_wcol[kMaxTaps];
};

__host__ __device__ inline size_t halo_dgrad_smem_bytes(uint32_t plane_stride, int stages) {
  return (size_t)kHaloDgSlots * plane_stride + (size_t)stages * 8192 + 1024 /*align*/ + 512 /*barriers*/;
}

__global__ void __launch_bounds__(224, 2)
halo_conv_dgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ HaloDgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sH = base;
  const uint32_t sB = base + (uint32_t)kHaloDgSlots * p.plane_stride;
  const uint32_t bars = sB + (uint32_t)stages * 8192u;
  auto bfull = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto bempty = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  auto pfull = [&](int b) { return bars + (uint32_t)(2 * stages + b) * 8u; };
  auto pempty = [&](int b) { return bars + (uint32_t)(2 * stages + kHaloDgSlots + b) * 8u; };
  auto tfull = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * kHaloDgSlots + a) * 8u; };
  auto tempty = [&](int a) { return bars + (uint32_t)(2 * stages + 2 * kHaloDgSlots + 2 + a) * 8u; };
  const uint32_t tslot = bars + (uint32_t)(2 * stages + 2 * kHaloDgSlots + 4) * 8u;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
    for (int b = 0; b < kHaloDgSlots; ++b) { mbar_init(pfull(b), 1); mbar_init(pempty(b), 1); }
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
    if (lane == 0) {
      const uint32_t plane_bytes = (uint32_t)p.box_rows * (uint32_t)p.box_cols * 128u;
      int slot = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int n = tile / p.TPI, i0 = (tile - n * p.TPI) * p.TP;
        mbar_wait(pempty(slot), (uint32_t)(ph ^ 1));
        mbar_arrive_expect_tx(pfull(slot), plane_bytes);
        tma_load_4d(sH + (uint32_t)slot * p.plane_stride, &tmDY, pfull(slot), 0, p.w_start, p.h_start + i0, n);
        if (++slot == kHaloDgSlots) { slot = 0; ph ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ===================== TMA producer of the weight ring =====================
    if (lane == 0) {
      int st = 0, ph = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x)
        for (int t = 0; t < ntaps; ++t) {
          mbar_wait(bempty(st), (uint32_t)(ph ^ 1));
          mbar_arrive_expect_tx(bfull(st), 8192u);
          tma_load_2d(sB + (uint32_t)st * 8192u, &tmW, bfull(st), (int)p.tap_wcol[t] * 64, 0);
          if (++st == stages) { st = 0; ph ^= 1; }
        }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_h16(64, 0, 0, 0, 1);
      const uint64_t bdesc0 = make_smem_desc(sB, 8192u, 1024u, 2);  // MN-major: 8-k-row atoms 1024 B apart
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
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_f16(d_tmem, ad0 + (uint64_t)(j * 2), bd0 + (uint64_t)(j * 128), idesc, (uint32_t)((t != t0) | (j != 0)));
            umma_commit(bempty(st));
            if (++st == stages) { st = 0; ph ^= 1; }
          }
          umma_commit(tfull(acc));
        }
        umma_commit(pempty(slot));
        if (++slot == kHaloDgSlots) { slot = 0; sph ^= 1; }
      }
    }
    __syncwarp();
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

}  // namespace var
