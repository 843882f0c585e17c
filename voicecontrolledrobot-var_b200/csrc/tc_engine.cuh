// tcgen05 implicit-GEMM engine for the VAR encoders (sm_100a).
//
// Two kernels cover every dense contraction on the hot path:
//
//  tc_gemm_kernel  : D[m, n] = sum_k A(m, k) * B(n, k)
//      A (K-major) is *gathered* by 4 producer warps straight from the NHWC
//      activation tensor (implicit im2col; cp.async 16 B with zero fill for
//      padding) into 128B-swizzled shared tiles.  B (the weights) arrives by
//      TMA, either K-major (forward / linear / GRU step) or MN-major (all
//      dgrads: the forward-packed weight matrix is read transposed, no second
//      copy).  One thread issues tcgen05.mma kind::tf32 (M=128, N<=256, K=8),
//      accumulators live in TMEM, the producer warps double as the epilogue
//      warps (tcgen05.ld -> bias / ReLU / mask / GRU cell -> HBM).
//
//  tc_wgrad_kernel : dW[k, o] = sum_pixels A'(pix, k) * dY(pix, o)
//      both operands MN-major (reduction runs over pixels, the slow dimension
//      of both NHWC tensors), split over pixel ranges, red.global.add to dW.
//
// Replaces cuDNN convolution / cuBLAS calls made by torch for
// models/pretext/arm_pretext_model.py:9-34 and ai2thor_pretext_model.py:5-38.
#pragma once
#include "common.cuh"

namespace var {

enum GatherMode : int {
  G_VEC_FWD = 0,    // NHWC source, C % 32 == 0, forward im2col
  G_VEC_DGRAD = 1,  // NHWC dY source, rows are input pixels (transposed conv)
  G_SCALAR_F32 = 2, // generic strides, float source (C in {1,3}), forward im2col
  G_SCALAR_U8 = 3,  // generic strides, uint8 source scaled by `scale`
  G_TMA_IM2COL = 4, // NHWC source, C % 32 == 0: A tiles by im2col-mode TMA (one instruction / k-block)
  G_TMA_TILED = 5   // plain [M, K] row-major source (Linear / GRU step): A tiles by tiled TMA
};

constexpr int kMaxTaps = 64;

struct GatherGeom {
  const void* src;
  int M;              // number of rows (pixels)
  int P, Q;           // row m -> (n, p, q): m = (n*P + p)*Q + q
  int H, W, C;        // source spatial extent / channels
  int R, S;           // filter taps
  int sh, sw, ph, pw; // stride / padding
  long long sN, sH, sW, sC;  // element strides of src
  int K;              // R*S*C
  float scale;        // scalar path multiplier (1/255 for u8 images)
};

enum EpiKind : int { EPI_STD = 0, EPI_GRU_FWD = 1, EPI_GRU_BWD = 2 };

// Output-row remapping of the stride-2 dgrad sub-problems: GEMM row m enumerates the pixels
// (n, h2, w2) of one parity class, stored at pixel (h2*sh + oh, w2*sw + ow) of the full map.
struct RowMap {
  int on;
  int P2, Q2;   // sub-grid extents
  int H, W;     // full extents
  int sh, sw, oh, ow;
};

struct EpiParams {
  float* out;            // [M, ldo]
  long long ldo;
  const float* bias;     // per output column or nullptr
  const float* mask;     // ReLU-backward mask source ([M, ldm], >0 keeps) or nullptr
  long long ldm;
  const float* addsrc;   // out = acc*... + addsrc[m, n] (nullptr = none)
  long long lda;
  int ncols;             // total valid output columns
  int relu;
  int round_out;         // store tf32-rounded values
  RowMap map;
  // 16-bit engine (gemm_persist.cuh, H16): storage type of `out` (0 fp32, 1 f16, 2 bf16; ldo in elements)
  // and of the ReLU mask source (0 fp32, 1 f16)
  int out_kind, mask_kind;
  const float* out_scale;  // device scalar multiplied into the result before mask / store (nullptr = 1): undoes the
                           // f16 gradient scale where a gradient leaves the 16-bit region
};

// GRU cell epilogue (forward): columns of the tile are [r | z | n] blocks of
// `jb` hidden units.  Follows torch.nn.GRU gate order (r, z, n).
struct GruEpiParams {
  const float* xproj;   // [B, T, 3H] (includes b_ih), row pitch ldx, offset for step t pre-applied
  long long ldx;
  const float* bhh;     // [3H]
  const float* hprev;   // [B, H]
  float* hnew;          // [B, H]
  float* hnew_r;        // [B, H] tf32-rounded copy (next step's MMA operand; nullptr = skip)
  float* gates;         // [B, 3H] saved r, z, n for backward (nullptr = inference)
  float* hn_save;       // [B, H] saved (W_hn h + b_hn) for backward (nullptr = inference)
  int Hdim;
};

// GRU BPTT epilogue: the tile holds dgh_s . W_hh for a slice of hidden units; adding dh_s * z_s
// gives dh_{s-1}, and the cell backward of step s-1 follows in registers (elementwise per unit).
struct GruBwdEpiParams {
  const float* dhd_in;   // [B, H]  dh_s * z_s
  const float* gates;    // [B, 3H] saved r, z, n of step s-1
  const float* hn_save;  // [B, H]
  const float* hprev;    // [B, H]  h before step s-1
  float* dgi;            // [B, ldgi] slot of step s-1 in the batch-major [B, T, 3H] buffer
  long long ldgi;
  float* dgh;            // [B, 3H] of step s-1
  float* dhd_out;        // [B, H]  dh_{s-1} * z_{s-1}
  int Hdim;
};

struct GemmParams {
  GatherGeom g[2];
  EpiParams e[2];
  GruEpiParams gru[2];
  GruBwdEpiParams grub[2];
  int bn;          // UMMA N / tile columns
  int b_mn_major;  // 0: B K-major boxes {32k, box_rows}; 1: MN-major boxes {32n, 32k}
  int nbox;        // K-major: number of row boxes (1 or 3)
  int box_rows;    // K-major: rows per box
  int boxbase[3];  // K-major: first global row of each box (tile n adds n*box_rows)
  int num_kb;      // K blocks of 32
  int kb_per_rs;   // dgrad MN-major: C_out/32 (k-blocks per filter tap)
  int cin_total;   // dgrad MN-major: C_in (columns per tap in packed weight)
  int stages;
  int lookahead;
  int mn_lbo, mn_sbo, mn_type;  // MN-major descriptor fields (bytes, bytes, layout type)
  // TMA im2col A operand: k-block -> (tap, channel chunk); per tap the im2col offsets and the
  // tap id (r*S + s) that addresses the packed weights
  int ntaps, cpb;              // taps, 32-channel chunks per tap
  int base_w, base_h;          // base-pixel coordinate of output (p, q) = (0, 0)
  int step_w, step_h;          // base-pixel step per output pixel (= traversal stride)
  uint8_t tap_w[kMaxTaps], tap_h[kMaxTaps], tap_id[kMaxTaps];
  // scalar (first-layer) mode: M tiles are whole output rows of one image; the input patch of the
  // tile is staged once in shared memory
  int sc_rpt, sc_tpi;          // output rows per tile, tiles per image
  int sc_nh, sc_wpad;          // staged rows / padded row width
  int epi_coalesce;            // persistent kernel: transpose the accumulator through smem (512 B per store instr.)
  int epi_tma;                 // persistent kernel: plain fp32 output tiles leave through swizzled smem boxes + TMA stores
  int kps;                     // persistent kernel: k-blocks per pipeline stage (1 or 2)
  int a_fmt, b_fmt;            // 16-bit engine: operand formats of the kind::f16 MMA (0 f16, 1 bf16)
  int prof_dgrad;              // profiler tag only: count this K-major-B GEMM with the dgrad family
};

constexpr int kTileM = 128;
constexpr int kTileABytes = kTileM * 128;  // 128 rows x 32 tf32

__host__ __device__ inline int tmem_cols_for(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}
__host__ __device__ inline size_t gemm_smem_bytes(int bn, int stages) {
  return (size_t)stages * (kTileABytes + (size_t)bn * 128) + 1024 /*align*/ + 256 /*barriers*/;
}

// ---------------------------------------------------------------------------
// A-row gather for one K block.  Thread `row` owns one 128-byte row.
// ---------------------------------------------------------------------------
struct RowCtx {
  bool valid;
  int n, p, q;
  long long base;  // n*sN
};

template <int GMODE>
__device__ __forceinline__ void gather_row(const GatherGeom& g, const RowCtx& rc, int kb,
                                           uint32_t tile, int row, const int* lut_off,
                                           const int* lut_rs) {
  if constexpr (GMODE == G_VEC_FWD || GMODE == G_VEC_DGRAD) {
    const int cpb = g.C >> 5;
    const int rs = kb / cpb;
    const int c0 = (kb - rs * cpb) << 5;
    const int r = rs / g.S, s = rs - r * g.S;
    bool ok = rc.valid;
    int h, w;
    if constexpr (GMODE == G_VEC_FWD) {
      h = rc.p * g.sh - g.ph + r;
      w = rc.q * g.sw - g.pw + s;
    } else {
      int th = rc.p + g.ph - r, tw = rc.q + g.pw - s;
      ok = ok && th >= 0 && tw >= 0;
      h = th / g.sh;
      w = tw / g.sw;
      ok = ok && (h * g.sh == th) && (w * g.sw == tw);
    }
    ok = ok && h >= 0 && h < g.H && w >= 0 && w < g.W;
    const float* src = reinterpret_cast<const float*>(g.src);
    if (ok) src += rc.base + (long long)h * g.sH + (long long)w * g.sW + c0;
    const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
    for (int c = 0; c < 8; ++c) cp_async_16(tile + swz128(row, c), src + (ok ? c * 4 : 0), nb);
  } else {
    // scalar im2col from the staged patch: lut_off[k] = (c*nh + r)*wpad + s (or -1 beyond K);
    // rc.base = offset of the row's window origin in the patch (or -1: row not in this tile)
    const float* sin = reinterpret_cast<const float*>(lut_rs);
    const int b0 = (int)rc.base;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int o = lut_off[kb * 32 + c * 4 + j];
        v[j] = (b0 >= 0 && o >= 0) ? sin[b0 + o] : 0.f;
      }
      st_shared_v4(tile + swz128(row, c), v[0], v[1], v[2], v[3]);
    }
  }
}

__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: cp_async_wait<0>(); break;
    case 1: cp_async_wait<1>(); break;
    case 2: cp_async_wait<2>(); break;
    default: cp_async_wait<3>(); break;
  }
}

// Gate non-linearities of the GRU epilogues: the cell maths of a step is issue bound (128 rows x 32 units x 3 gates per
// CTA and step), so both go through ex2.approx + rcp.approx (absolute error ~1e-7, two orders below the operand
// rounding) instead of the IEEE division / libm tanhf sequences.
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanhf_(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// ---------------------------------------------------------------------------
// Main GEMM kernel.  grid = (M tiles, N tiles, instances<=2), block = 160.
// ---------------------------------------------------------------------------
template <int GMODE, int EPI>
__global__ void __launch_bounds__(160)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmB0, const __grid_constant__ CUtensorMap tmB1,
               const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  griddep_launch();  // PDL: the next GRU step may run its prologue under this grid
  if constexpr (GMODE == G_SCALAR_F32 || GMODE == G_SCALAR_U8) griddep_wait();  // stages inputs first
  const int z = blockIdx.z;
  const CUtensorMap* tmB = z == 0 ? &tmB0 : &tmB1;
  const CUtensorMap* tmA = z == 0 ? &tmA0 : &tmA1;
  constexpr bool kTmaA = (GMODE == G_TMA_IM2COL || GMODE == G_TMA_TILED);
  const GatherGeom& g = p.g[z];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages, bn = p.bn, num_kb = p.num_kb;
  const uint32_t tileB_bytes = (uint32_t)bn * 128u;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)stages * kTileABytes;
  const uint32_t bars = sB + (uint32_t)stages * tileB_bytes;  // full[s], empty[s], tmem_full, slot
  auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  const uint32_t tfull_bar = bars + (uint32_t)(2 * stages) * 8u;
  const uint32_t tslot = tfull_bar + 8u;

  constexpr bool kScalar = (GMODE == G_SCALAR_F32 || GMODE == G_SCALAR_U8);
  __shared__ int lut_off[kScalar ? 256 : 1];
  const float* s_patch = reinterpret_cast<const float*>(smem_raw + (((tslot + 23u) & ~15u) - smem_u32(smem_raw)));
  int sc_n = 0, sc_p0 = 0, sc_rows = 0;
  if constexpr (kScalar) {
    // tile = up to sc_rpt output rows of image sc_n; stage their input rows (zero padded borders,
    // scale and tf32 rounding applied once per element)
    sc_n = blockIdx.x / p.sc_tpi;
    sc_p0 = (blockIdx.x - sc_n * p.sc_tpi) * p.sc_rpt;
    sc_rows = min(p.sc_rpt, g.P - sc_p0);
    const int kpad = num_kb * 32;
    for (int k = tid; k < kpad; k += blockDim.x) {
      int o = -1;
      if (k < g.K) {
        const int c = k % g.C, rs = k / g.C, s = rs % g.S, r = rs / g.S;
        o = (c * p.sc_nh + r) * p.sc_wpad + s;
      }
      lut_off[k] = o;
    }
    float* patch = const_cast<float*>(s_patch);
    const int h_lo = sc_p0 * g.sh - g.ph;
    const int nh = (sc_rows - 1) * g.sh + g.R;
    const int total = g.C * nh * p.sc_wpad;
    for (int i = tid; i < total; i += blockDim.x) {
      const int ww = i % p.sc_wpad;
      const int t = i / p.sc_wpad;
      const int hh = t % nh, c = t / nh;
      const int h = h_lo + hh, w = ww - g.pw;
      float x = 0.f;
      if (h >= 0 && h < g.H && w >= 0 && w < g.W) {
        const long long idx = (long long)sc_n * g.sN + (long long)h * g.sH + (long long)w * g.sW +
                              (long long)c * g.sC;
        if constexpr (GMODE == G_SCALAR_U8)
          x = (float)reinterpret_cast<const uint8_t*>(g.src)[idx] * g.scale;
        else
          x = reinterpret_cast<const float*>(g.src)[idx] * g.scale;
      }
      patch[(c * p.sc_nh + hh) * p.sc_wpad + ww] = round_tf32(x);
    }
  }

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), kTmaA ? 1 : 128 + 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(tmB);
    if constexpr (kTmaA) tma_prefetch_desc(tmA);
  }
  const uint32_t ncols = (uint32_t)tmem_cols_for(bn);
  if (warp == 4) tmem_alloc(tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));
  griddep_wait();  // everything above overlapped the previous kernel; its results are read below

  const int m0 = kScalar ? (sc_n * g.P + sc_p0) * g.Q : blockIdx.x * kTileM;
  const int m_end = kScalar ? m0 + sc_rows * g.Q : g.M;
  const int ntile = blockIdx.y;

  if (warp < 4) {
    // ===================== producers (then epilogue) =====================
    const int row = tid;
    RowCtx rc;
    if constexpr (kScalar) {
      const int pr = row / g.Q, q = row - pr * g.Q;
      rc.valid = pr < sc_rows;
      rc.n = sc_n; rc.p = sc_p0 + pr; rc.q = q;
      rc.base = rc.valid ? (long long)(pr * g.sh) * p.sc_wpad + q * g.sw : -1;
    } else {
      const int m = m0 + row;
      rc.valid = m < g.M;
      const int pq = g.P * g.Q;
      const int mm = rc.valid ? m : 0;
      rc.n = mm / pq;
      const int rem = mm - rc.n * pq;
      rc.p = rem / g.Q;
      rc.q = rem - rc.p * g.Q;
      rc.base = (long long)rc.n * g.sN;
    }
    if constexpr (kTmaA) {
      // Warp 0 runs the producer loop CONVERGED and one elected lane issues every box of the k-block from uniform
      // registers.  (Until the end of round 2 lane 0 of all four epilogue warps shared the issue, because one `lane == 0`
      // thread was the bottleneck beyond two TMA instructions per k-block: that cost was the ELECT / R2UR wrapping nvcc
      // puts around tensor-map instructions in divergent code, not the TMA unit.)  The other warps stay out of the loop:
      // an idle waiter can fall two phases behind an mbarrier and then never sees its parity flip.
      if (warp == 0) {
        int st = 0, ph = 0;
        int w0 = 0, h0 = 0, n0 = 0;
        if constexpr (GMODE == G_TMA_IM2COL) {
          const int pq0 = g.P * g.Q;
          const int mm0 = m0 < g.M ? m0 : 0;
          n0 = mm0 / pq0;
          const int rem0 = mm0 - n0 * pq0;
          const int p0 = rem0 / g.Q;
          w0 = (rem0 - p0 * g.Q) * p.step_w + p.base_w;
          h0 = p0 * p.step_h + p.base_h;
        }
        // (tap, channel block) = divmod(k-block, blocks per tap), kept incrementally
        const int period = GMODE == G_TMA_IM2COL ? p.cpb : (p.b_mn_major ? p.kb_per_rs : 1 << 30);
        int tap = 0, cb = 0;
        for (int it = 0; it < num_kb; ++it) {
          mbar_wait(empty_bar(st), (uint32_t)(ph ^ 1));
          const uint32_t dstA = sA + (uint32_t)st * kTileABytes;
          const uint32_t dstB = sB + (uint32_t)st * tileB_bytes;
          const int c0 = GMODE == G_TMA_IM2COL ? (cb << 5) : (it << 5);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(full_bar(st), (uint32_t)kTileABytes + tileB_bytes);
            if constexpr (GMODE == G_TMA_IM2COL)
              tma_load_im2col_4d(dstA, tmA, full_bar(st), c0, w0, h0, n0, p.tap_w[tap], p.tap_h[tap]);
            else
              tma_load_2d(dstA, tmA, full_bar(st), it * 32, m0);
            if (!p.b_mn_major) {
              for (int b = 0; b < p.nbox; ++b)
                tma_load_2d(dstB + (uint32_t)(b * p.box_rows) * 128u, tmB, full_bar(st), it * 32,
                            p.boxbase[b] + ntile * p.box_rows);
            } else {
              const int rs = GMODE == G_TMA_IM2COL ? (int)p.tap_id[tap] : tap;
              const int k0 = cb << 5;
              for (int gidx = 0; gidx < (bn >> 5); ++gidx)
                tma_load_2d(dstB + (uint32_t)gidx * 4096u, tmB, full_bar(st),
                            rs * p.cin_total + ntile * bn + gidx * 32, k0);
            }
          }
          __syncwarp();
          if (++cb == period) { cb = 0; ++tap; }
          if (++st == stages) { st = 0; ph ^= 1; }
        }
      }
      __syncwarp();
    } else {
    const int la = p.lookahead;
      int st_issue = 0, ph_issue = 0, st_arr = 0;
      for (int it = 0; it < num_kb + la; ++it) {
        if (it < num_kb) {
          mbar_wait(empty_bar(st_issue), (uint32_t)(ph_issue ^ 1));
          if (tid == 0) {
            mbar_arrive_expect_tx(full_bar(st_issue), tileB_bytes);
            const uint32_t dstB = sB + (uint32_t)st_issue * tileB_bytes;
            if (!p.b_mn_major) {
              for (int b = 0; b < p.nbox; ++b)
                tma_load_2d(dstB + (uint32_t)(b * p.box_rows) * 128u, tmB, full_bar(st_issue),
                            it * 32, p.boxbase[b] + ntile * p.box_rows);
            } else {
              // MN-major: k-block `it` -> tap rs and k0; box = {32 n, 32 k}
              const int rs = it / p.kb_per_rs;
              const int k0 = (it - rs * p.kb_per_rs) << 5;
              for (int gidx = 0; gidx < (bn >> 5); ++gidx)
                tma_load_2d(dstB + (uint32_t)gidx * 4096u, tmB, full_bar(st_issue),
                            rs * p.cin_total + ntile * bn + gidx * 32, k0);
            }
          }
          gather_row<GMODE>(g, rc, it, sA + (uint32_t)st_issue * kTileABytes, row, lut_off,
                            reinterpret_cast<const int*>(s_patch));
          if (++st_issue == stages) { st_issue = 0; ph_issue ^= 1; }
        }
        cp_async_commit();
        if (it >= la) {
          cp_async_wait_dyn(la);
          fence_proxy_async_smem();
          mbar_arrive(full_bar(st_arr));
          if (++st_arr == stages) st_arr = 0;
        }
      }
    }

    // ===================== epilogue =====================
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    if constexpr (EPI == EPI_STD) {
      const EpiParams& e = p.e[z];
      for (int c = 0; c < bn; c += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)c, v);
        tmem_ld_wait();
        const int col0 = ntile * bn + c;
        if (m < m_end && col0 < e.ncols) {
          long long orow = m;
          if (e.map.on) {
            const int pq2 = e.map.P2 * e.map.Q2;
            const int n_ = m / pq2, rem_ = m - n_ * pq2;
            const int h2 = rem_ / e.map.Q2, w2 = rem_ - h2 * e.map.Q2;
            orow = ((long long)n_ * e.map.H + h2 * e.map.sh + e.map.oh) * e.map.W + w2 * e.map.sw +
                   e.map.ow;
          }
          float* o = e.out + orow * e.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 r4 = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            if (e.bias) {
              const float4 b4 = *reinterpret_cast<const float4*>(e.bias + col0 + j);
              r4.x += b4.x; r4.y += b4.y; r4.z += b4.z; r4.w += b4.w;
            }
            if (e.addsrc) {
              const float4 a4 =
                  *reinterpret_cast<const float4*>(e.addsrc + orow * e.lda + col0 + j);
              r4.x += a4.x; r4.y += a4.y; r4.z += a4.z; r4.w += a4.w;
            }
            if (e.relu) {
              r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f);
              r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f);
            }
            if (e.mask) {
              const float4 k4 =
                  *reinterpret_cast<const float4*>(e.mask + orow * e.ldm + col0 + j);
              r4.x = k4.x > 0.f ? r4.x : 0.f; r4.y = k4.y > 0.f ? r4.y : 0.f;
              r4.z = k4.z > 0.f ? r4.z : 0.f; r4.w = k4.w > 0.f ? r4.w : 0.f;
            }
            if (e.round_out) {
              r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
              r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
            }
            *reinterpret_cast<float4*>(o + j) = r4;
          }
        }
      }
    } else if constexpr (EPI == EPI_GRU_FWD) {
      // GRU forward cell: tile columns = [r(jb) | z(jb) | n(jb)], jb = bn/3.  The accumulator
      // chunk is transposed through shared memory (the operand stages are free once tfull_bar
      // fired) so that each warp-level global access covers 32 consecutive hidden units of one
      // row: 128-byte coalesced instead of 32 rows x 16 bytes.
      const GruEpiParams& q = p.gru[z];
      const int jb = bn / 3;
      const int Hd = q.Hdim;
      float* scr = reinterpret_cast<float*>(smem_raw + (sA - smem_u32(smem_raw))) + warp * (3 * 32 * 33);
      for (int c = 0; c < jb; c += 32) {
        {
          float v[32];
#pragma unroll
          for (int gq = 0; gq < 3; ++gq) {
            tmem_ld32(trow + (uint32_t)(gq * jb + c), v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) scr[(gq * 32 + lane) * 33 + j] = v[j];
          }
        }
        __syncwarp();
        const int j = ntile * jb + c + lane;  // this lane's hidden unit
        const float br = __ldg(q.bhh + j), bz = __ldg(q.bhh + Hd + j), bq = __ldg(q.bhh + 2 * Hd + j);
        const int mrow0 = m0 + warp * 32;
        const float* __restrict__ xproj = q.xproj;
        const float* __restrict__ hprev = q.hprev;
        constexpr int RB = 8;  // rows in flight: all loads of a batch are issued before any store
        for (int r0 = 0; r0 < 32; r0 += RB) {
          float xr[RB], xz[RB], xn[RB], hp[RB];
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int mr = mrow0 + r0 + u;
            const bool ok = mr < g.M;
            const float* xp = xproj + (long long)(ok ? mr : 0) * q.ldx;
            xr[u] = ok ? __ldg(xp + j) : 0.f;
            xz[u] = ok ? __ldg(xp + Hd + j) : 0.f;
            xn[u] = ok ? __ldg(xp + 2 * Hd + j) : 0.f;
            hp[u] = ok ? __ldg(hprev + (long long)mr * Hd + j) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int rr = r0 + u;
            const int mr = mrow0 + rr;
            if (mr < g.M) {
              const float vr = scr[(0 * 32 + rr) * 33 + lane], vz = scr[(1 * 32 + rr) * 33 + lane],
                          vn = scr[(2 * 32 + rr) * 33 + lane];
              const float r_ = sigmoidf_(xr[u] + vr + br);
              const float z_ = sigmoidf_(xz[u] + vz + bz);
              const float hnv = vn + bq;
              const float n_ = tanhf_(xn[u] + r_ * hnv);
              const float h_ = (1.f - z_) * n_ + z_ * hp[u];
              q.hnew[(long long)mr * Hd + j] = h_;
              if (q.hnew_r) q.hnew_r[(long long)mr * Hd + j] = round_tf32(h_);
              if (q.gates) {
                float* gs = q.gates + (long long)mr * 3 * Hd + j;
                gs[0] = r_; gs[Hd] = z_; gs[2 * Hd] = n_;
                q.hn_save[(long long)mr * Hd + j] = hnv;
              }
            }
          }
        }
        __syncwarp();
      }
    } else {
      // GRU BPTT: dh_{s-1} = acc + dh_s * z_s, then the cell backward of step s-1 (same maths as
      // gru_cell_bwd_kernel), transposed through shared memory like the forward epilogue.
      const GruBwdEpiParams& q = p.grub[z];
      const int Hd = q.Hdim;
      float* scr = reinterpret_cast<float*>(smem_raw + (sA - smem_u32(smem_raw))) + warp * (32 * 33);
      for (int c = 0; c < bn; c += 32) {
        {
          float v[32];
          tmem_ld32(trow + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) scr[lane * 33 + j] = v[j];
        }
        __syncwarp();
        const int j = ntile * bn + c + lane;
        const int mrow0 = m0 + warp * 32;
        const float* __restrict__ dhd_in = q.dhd_in;
        const float* __restrict__ gates = q.gates;
        const float* __restrict__ hn_save = q.hn_save;
        const float* __restrict__ hprev = q.hprev;
        constexpr int RB = 8;
        for (int r0 = 0; r0 < 32; r0 += RB) {
          float din[RB], gr_[RB], gz_[RB], gn_[RB], hnv[RB], hp[RB];
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int mr = mrow0 + r0 + u;
            const bool ok = mr < g.M;
            const long long hoff = (long long)(ok ? mr : 0) * Hd + j;
            const float* gt = gates + (long long)(ok ? mr : 0) * 3 * Hd + j;
            din[u] = ok ? __ldg(dhd_in + hoff) : 0.f;
            gr_[u] = ok ? __ldg(gt) : 0.f;
            gz_[u] = ok ? __ldg(gt + Hd) : 0.f;
            gn_[u] = ok ? __ldg(gt + 2 * Hd) : 0.f;
            hnv[u] = ok ? __ldg(hn_save + hoff) : 0.f;
            hp[u] = ok ? __ldg(hprev + hoff) : 0.f;
          }
#pragma unroll
          for (int u = 0; u < RB; ++u) {
            const int rr = r0 + u;
            const int mr = mrow0 + rr;
            if (mr < g.M) {
              const long long hoff = (long long)mr * Hd + j;
              const float dh = scr[rr * 33 + lane] + din[u];
              const float r_ = gr_[u], z_ = gz_[u], n_ = gn_[u];
              const float dnn = dh * (1.f - z_);
              const float dzz = dh * (hp[u] - n_);
              const float dnp = dnn * (1.f - n_ * n_);
              const float dzp = dzz * z_ * (1.f - z_);
              const float drp = dnp * hnv[u] * r_ * (1.f - r_);
              const float dr = round_tf32(drp), dz = round_tf32(dzp), dn = round_tf32(dnp);
              float* gi = q.dgi + (long long)mr * q.ldgi + j;
              gi[0] = dr; gi[Hd] = dz; gi[2 * Hd] = dn;
              float* gh = q.dgh + (long long)mr * 3 * Hd + j;
              gh[0] = dr; gh[Hd] = dz; gh[2 * Hd] = round_tf32(dnp * r_);
              q.dhd_out[hoff] = dh * z_;
            }
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  } else {
    // ===================== MMA issuer (warp 4) =====================
    const uint32_t idesc = make_idesc_tf32(bn, 0, p.b_mn_major);
    int st = 0, ph = 0;
    for (int kb = 0; kb < num_kb; ++kb) {
      mbar_wait(full_bar(st), (uint32_t)ph);
      tc_fence_after();
      if (elect_one_sync()) {  // one lane of the converged warp: operands stay in uniform registers
        const uint32_t a0 = sA + (uint32_t)st * kTileABytes;
        const uint32_t b0 = sB + (uint32_t)st * tileB_bytes;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t ad = make_smem_desc(a0 + (uint32_t)j * 32u, 16u, 1024u);
          const uint64_t bd = p.b_mn_major ? make_smem_desc(b0 + (uint32_t)j * 1024u, (uint32_t)p.mn_lbo,
                                                            (uint32_t)p.mn_sbo, (uint32_t)p.mn_type)
                                           : make_smem_desc(b0 + (uint32_t)j * 32u, 16u, 1024u);
          umma_tf32(tmem_base, ad, bd, idesc, (uint32_t)((kb | j) != 0));
        }
        umma_commit(empty_bar(st));
        if (kb == num_kb - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++st == stages) { st = 0; ph ^= 1; }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ncols);
  }
}

// ---------------------------------------------------------------------------
// Weight-gradient kernel: dW[o, k] += sum_{pix} dY[pix, o] * im2col(X)[pix, k]
// D tile = [128 k-rows x Cout columns]; both operands MN-major.
// grid = (ceil(Kfwd/128), splits), block = 160.
// ---------------------------------------------------------------------------
struct WgradParams {
  GatherGeom g;       // forward im2col geometry of X (rows = output pixels)
  const float* dy;    // [M, ldy] (a slab of `cout` columns)
  long long ldy;
  int cout;           // multiple of 16, <= 256
  float* dw;          // [cout, kpad]
  int kpad;
  int pix_per_cta;    // multiple of 32
  int stages;
  int lookahead;
  int mn_lbo, mn_sbo, mn_type;  // MN-major descriptor fields
  int mn_swz32;       // 1: 32-byte-granule swizzle (BASE32B), 0: 16-byte (plain 128B)
  // scalar (first-layer) mode: a CTA owns `sc_rpt` output rows of one image and stages their
  // input rows in shared memory once (blockIdx.y = image * sc_tpi + row tile)
  int sc_rpt, sc_tpi, sc_nh, sc_wpad;
};

__host__ __device__ inline size_t wgrad_smem_bytes(int cout, int stages) {
  const int bgroups = (cout + 31) / 32;
  return (size_t)stages * (4 * 4096 + (size_t)bgroups * 4096) + 1024 + 256;
}

template <int GMODE>
__global__ void __launch_bounds__(160)
tc_wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  const GatherGeom& g = p.g;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const int bgroups = (p.cout + 31) >> 5;
  const uint32_t tileA_bytes = 4u * 4096u;
  const uint32_t tileB_bytes = (uint32_t)bgroups * 4096u;

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)stages * tileA_bytes;
  const uint32_t bars = sB + (uint32_t)stages * tileB_bytes;
  auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  const uint32_t tfull_bar = bars + (uint32_t)(2 * stages) * 8u;
  const uint32_t tslot = tfull_bar + 8u;

  constexpr bool kScalar = (GMODE == G_SCALAR_F32 || GMODE == G_SCALAR_U8);
  __shared__ int lut_off[kScalar ? 256 : 1];
  const float* s_patch = reinterpret_cast<const float*>(smem_raw + (((tslot + 23u) & ~15u) - smem_u32(smem_raw)));
  int sc_n = 0, sc_p0 = 0, sc_rows = 0;
  if constexpr (kScalar) {
    sc_n = blockIdx.y / p.sc_tpi;
    sc_p0 = (blockIdx.y - sc_n * p.sc_tpi) * p.sc_rpt;
    sc_rows = min(p.sc_rpt, g.P - sc_p0);
    for (int k = tid; k < 256; k += blockDim.x) {
      int o = -1;
      if (k < g.K) {
        const int c = k % g.C, rs = k / g.C, s_ = rs % g.S, r = rs / g.S;
        o = (c * p.sc_nh + r) * p.sc_wpad + s_;
      }
      lut_off[k] = o;
    }
    float* patch = const_cast<float*>(s_patch);
    const int h_lo = sc_p0 * g.sh - g.ph;
    const int nh = (sc_rows - 1) * g.sh + g.R;
    const int total = g.C * nh * p.sc_wpad;
    for (int i = tid; i < total; i += blockDim.x) {
      const int ww = i % p.sc_wpad;
      const int t = i / p.sc_wpad;
      const int hh = t % nh, c = t / nh;
      const int h = h_lo + hh, w = ww - g.pw;
      float x = 0.f;
      if (h >= 0 && h < g.H && w >= 0 && w < g.W) {
        const long long idx = (long long)sc_n * g.sN + (long long)h * g.sH + (long long)w * g.sW +
                              (long long)c * g.sC;
        if constexpr (GMODE == G_SCALAR_U8)
          x = (float)reinterpret_cast<const uint8_t*>(g.src)[idx] * g.scale;
        else
          x = reinterpret_cast<const float*>(g.src)[idx] * g.scale;
      }
      patch[(c * p.sc_nh + hh) * p.sc_wpad + ww] = round_tf32(x);
    }
  }
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_fence_init();
  }
  const uint32_t ncols = (uint32_t)tmem_cols_for(p.cout);
  if (warp == 4) tmem_alloc(tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  const int ktile = blockIdx.x;  // 128 k-rows
  const int pix0 = kScalar ? (sc_n * g.P + sc_p0) * g.Q : blockIdx.y * p.pix_per_cta;
  const int pix1 = kScalar ? pix0 + sc_rows * g.Q : min(pix0 + p.pix_per_cta, g.M);
  const int num_kb = (pix1 - pix0 + 31) / 32;  // may be <= 0 for trailing CTAs

  if (num_kb > 0) {
    if (warp < 4) {
      const int grp = tid >> 5, row = tid & 31;
      const int la = p.lookahead;
      auto swz = [&](uint32_t r, uint32_t c) { return p.mn_swz32 ? swz128_32(r, c) : swz128(r, c); };
      const int pq = g.P * g.Q;
      int st_issue = 0, ph_issue = 0, st_arr = 0;
      for (int it = 0; it < num_kb + la; ++it) {
        if (it < num_kb) {
          mbar_wait(empty_bar(st_issue), (uint32_t)(ph_issue ^ 1));
          const int m = pix0 + it * 32 + row;
          const bool mv = m < pix1;
          const uint32_t tA = sA + (uint32_t)st_issue * tileA_bytes + (uint32_t)grp * 4096u;
          // ---- A' group `grp`: k index range [k0, k0+32)
          const int k0 = ktile * 128 + grp * 32;
          const int mm = mv ? m : 0;
          const int n = mm / pq;
          const int rem = mm - n * pq;
          const int pp = rem / g.Q, qq = rem - pp * g.Q;
          if constexpr (GMODE == G_VEC_FWD) {
            bool ok = mv && k0 < g.K;
            const float* src = reinterpret_cast<const float*>(g.src);
            if (ok) {
              const int cpb = g.C >> 5;
              const int kb = k0 >> 5;
              const int rs = kb / cpb;
              const int c0 = (kb - rs * cpb) << 5;
              const int r = rs / g.S, s = rs - r * g.S;
              const int h = pp * g.sh - g.ph + r, w = qq * g.sw - g.pw + s;
              ok = h >= 0 && h < g.H && w >= 0 && w < g.W;
              if (ok) src += (long long)n * g.sN + (long long)h * g.sH + (long long)w * g.sW + c0;
            }
            const uint32_t nb = ok ? 16u : 0u;
#pragma unroll
            for (int c = 0; c < 8; ++c) cp_async_16(tA + swz(row, c), src + (ok ? c * 4 : 0), nb);
          } else {
            // from the staged patch: pixel i of this CTA -> window origin (pr*sh, q*sw)
            const int i = it * 32 + row;
            const int pr = i / g.Q, q = i - pr * g.Q;
            const int b0 = mv ? (pr * g.sh) * p.sc_wpad + q * g.sw : -1;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float v[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int k = k0 + c * 4 + j;
                const int o = k < 256 ? lut_off[k] : -1;
                v[j] = (b0 >= 0 && o >= 0) ? s_patch[b0 + o] : 0.f;
              }
              st_shared_v4(tA + swz(row, c), v[0], v[1], v[2], v[3]);
            }
          }
          // ---- B' (dY) groups: 32 output channels each, 4 producer warps share them
          for (int bg = grp; bg < bgroups; bg += 4) {
            const uint32_t tB = sB + (uint32_t)st_issue * tileB_bytes + (uint32_t)bg * 4096u;
            const float* src = p.dy;
            const int cvalid = p.cout - bg * 32;  // may be < 32 (cout = 16 mult)
            if (mv) src += (long long)m * p.ldy + bg * 32;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const bool ok = mv && (c * 4 < cvalid);
              cp_async_16(tB + swz(row, c), ok ? src + c * 4 : p.dy, ok ? 16u : 0u);
            }
          }
          if (++st_issue == stages) { st_issue = 0; ph_issue ^= 1; }
        }
        cp_async_commit();
        if (it >= la) {
          cp_async_wait_dyn(la);
          fence_proxy_async_smem();
          mbar_arrive(full_bar(st_arr));
          if (++st_arr == stages) st_arr = 0;
        }
      }
      // ---- epilogue: row = k index, columns = output channel
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const int k = ktile * 128 + warp * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c = 0; c < p.cout; c += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)c, v);
        tmem_ld_wait();
        if (k < g.K) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c + j < p.cout) atomicAdd(p.dw + (long long)(c + j) * p.kpad + k, v[j]);
        }
      }
      tc_fence_before();
    } else {
      const uint32_t idesc = make_idesc_tf32(p.cout, 1, 1);
      const uint32_t lbo = (uint32_t)p.mn_lbo, sbo = (uint32_t)p.mn_sbo, lt = (uint32_t)p.mn_type;
      int st = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar(st), (uint32_t)ph);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a0 = sA + (uint32_t)st * tileA_bytes;
          const uint32_t b0 = sB + (uint32_t)st * tileB_bytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint64_t ad = make_smem_desc(a0 + (uint32_t)j * 1024u, lbo, sbo, lt);
            const uint64_t bd = make_smem_desc(b0 + (uint32_t)j * 1024u, lbo, sbo, lt);
            umma_tf32(tmem_base, ad, bd, idesc, (uint32_t)((kb | j) != 0));
          }
          umma_commit(empty_bar(st));
          if (kb == num_kb - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
        if (++st == stages) { st = 0; ph ^= 1; }
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

// ---------------------------------------------------------------------------
// TMA-fed weight gradient: same D tile / split as tc_wgrad_kernel, but both MN-major
// operands arrive by TMA issued from one thread -- X^T groups by im2col-mode loads of
// 32 pixels x 32 channels (or tiled loads when X is a plain matrix), dY groups by tiled
// loads -- in the 32-byte-atom 128B swizzle the MN-major UMMA descriptors expect.  Four threads
// share the TMA issue; a pipeline stage holds p.kps 32-pixel blocks.
// grid = (ceil(K/128), pixel splits, slabs of p.cout output channels), block = 160.
// ---------------------------------------------------------------------------
struct WgradTmaParams {
  int M, K, cout, kpad;
  float* dw;          // [cout, kpad]
  int pix_per_cta;    // multiple of 32
  int stages;
  int mn_lbo, mn_sbo, mn_type;
  int a_tiled;        // 1: X is a plain [M, K] matrix (Linear)
  int kps;            // 32-pixel blocks per pipeline stage (1 or 2)
  int P, Q;           // output extents: pixel m -> (n, p, q)
  int cpb, base_w, base_h, step_w, step_h;
  uint8_t tap_w[kMaxTaps], tap_h[kMaxTaps];
  // bias gradient without a column-sum pass: when K % 128 != 0 the last k tile has unused 32-row groups; the
  // first of them is filled with ones once, so accumulator row K is colsum(dY) (db += it; nullptr = off)
  float* db;
  int ones_ktile;
};

template <int kVariant>
__global__ void __launch_bounds__(160)
tc_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY,
                    const __grid_constant__ WgradTmaParams p) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int stages = p.stages;
  const int bgroups = p.cout >> 5;
  const int kps = p.kps;  // 32-pixel blocks per pipeline stage
  const uint32_t tileA_bytes = 4u * 4096u;
  const uint32_t tileB_bytes = (uint32_t)bgroups * 4096u;
  const uint32_t stageA = (uint32_t)kps * tileA_bytes, stageB = (uint32_t)kps * tileB_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;
  const uint32_t sB = base + (uint32_t)stages * stageA;
  const uint32_t bars = sB + (uint32_t)stages * stageB;
  auto full_bar = [&](int s) { return bars + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return bars + (uint32_t)(stages + s) * 8u; };
  const uint32_t tfull_bar = bars + (uint32_t)(2 * stages) * 8u;
  const uint32_t tslot = tfull_bar + 8u;

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
  const uint32_t ncols = (uint32_t)tmem_cols_for(p.cout);
  if (warp == 4) tmem_alloc(tslot, ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  const int ktile = blockIdx.x;
  const int pix0 = blockIdx.y * p.pix_per_cta;
  const int pix1 = min(pix0 + p.pix_per_cta, p.M);
  const int num_kb = ((pix1 - pix0 + 31) / 32 + kps - 1) / kps;  // pipeline stages of kps pixel blocks
  const int kgroups = min(4, (p.K - ktile * 128 + 31) / 32);  // valid 32-row groups of this k tile
  const bool ones = p.db != nullptr && ktile == p.ones_ktile && kgroups < 4 && blockIdx.z * p.cout < 1 << 30;
  if (ones) {  // group `kgroups` of every X^T block := 1.0f (TMA never writes it in this k tile)
    for (uint32_t blk = 0; blk < (uint32_t)(stages * kps); ++blk)
      for (uint32_t i = tid; i < 4096u / 16u; i += 160u)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sA + blk * tileA_bytes + (uint32_t)kgroups * 4096u + i * 16u),
                     "r"(0x3F800000u) : "memory");
    fence_proxy_async_smem();
    __syncthreads();
  }

  if (num_kb > 0) {
    if (warp < 4) {
      // Four producer threads (lane 0 of warps 0-3) share the TMA issue work of a k-block: warp w
      // loads X^T group w and dY groups w, w+4, ...  One thread alone spends about as long issuing
      // the 6+ small boxes as the tensor core needs for the k-block.
      // (a warp with nothing to load stays out: an idle waiter could fall two phases behind)
      if (lane == 0 && (warp == 0 || warp < kgroups || warp < bgroups)) {
        // per-lane constants and an incrementally tracked base pixel (n, pp, qq): this thread's
        // instruction stream is on the critical path, so no divisions inside the loop
        const int gq = warp;
        const int kbg = ktile * 4 + gq;
        const int tap = p.a_tiled ? 0 : kbg / p.cpb;
        const int c0 = p.a_tiled ? 0 : (kbg - tap * p.cpb) << 5;
        int n = 0, pp = 0, qq = 0;
        if (!p.a_tiled) {
          const int pq = p.P * p.Q;
          n = pix0 / pq;
          const int rem = pix0 - n * pq;
          pp = rem / p.Q;
          qq = rem - pp * p.Q;
        }
        int st = 0, ph = 0;
        for (int it = 0; it < num_kb; ++it) {
          mbar_wait(empty_bar(st), (uint32_t)(ph ^ 1));
          if (warp == 0)
            mbar_arrive_expect_tx(full_bar(st), (uint32_t)kps * ((uint32_t)kgroups * 4096u + tileB_bytes));
          for (int sub = 0; sub < kps; ++sub) {
            // (blocks past the pixel range are loaded too: rows beyond M are zero filled by TMA and
            // rows of the next CTA's range cannot occur -- ranges are multiples of 32 * kps)
            const int m = pix0 + (it * kps + sub) * 32;
            const uint32_t dA = sA + (uint32_t)st * stageA + (uint32_t)sub * tileA_bytes;
            const uint32_t dB = sB + (uint32_t)st * stageB + (uint32_t)sub * tileB_bytes;
            if (gq < kgroups) {
              if (p.a_tiled) {
                tma_load_2d(dA + (uint32_t)gq * 4096u, &tmX, full_bar(st), ktile * 128 + gq * 32, m);
              } else {
                const int w0 = qq * p.step_w + p.base_w, h0 = pp * p.step_h + p.base_h;
                tma_load_im2col_4d(dA + (uint32_t)gq * 4096u, &tmX, full_bar(st), c0, w0, h0, n,
                                   p.tap_w[tap], p.tap_h[tap]);
              }
            }
            if (!p.a_tiled) {  // base pixel of the next 32-pixel block
              qq += 32;
              while (qq >= p.Q) { qq -= p.Q; ++pp; }
              while (pp >= p.P) { pp -= p.P; ++n; }
            }
            for (int bg = warp; bg < bgroups; bg += 4)
              tma_load_2d(dB + (uint32_t)bg * 4096u, &tmDY, full_bar(st), (int)blockIdx.z * p.cout + bg * 32, m);
          }
          if (++st == stages) { st = 0; ph ^= 1; }
        }
      }
      __syncwarp();
      // ---- epilogue: row = k index, columns = output channel
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const int k = ktile * 128 + warp * 32 + lane;
      float* dw_slab = p.dw + (long long)blockIdx.z * p.cout * p.kpad;  // grid.z = slab of cout columns
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
      for (int c = 0; c < p.cout; c += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)c, v);
        tmem_ld_wait();
        if (k < p.K) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dw_slab + (long long)(c + j) * p.kpad + k, v[j]);
        } else if (ones && k == p.K) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(p.db + (long long)blockIdx.z * p.cout + c + j, v[j]);
        }
      }
      tc_fence_before();
    } else {
      const uint32_t idesc = make_idesc_tf32(p.cout, 1, 1);
      const uint32_t lbo = (uint32_t)p.mn_lbo, sbo = (uint32_t)p.mn_sbo, lt = (uint32_t)p.mn_type;
      const uint64_t adesc0 = make_smem_desc(sA, lbo, sbo, lt), bdesc0 = make_smem_desc(sB, lbo, sbo, lt);
      int st = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {  // converged warp, one elected lane issues
        mbar_wait(full_bar(st), (uint32_t)ph);
        tc_fence_after();
        if (elect_one_sync()) {
          for (int sub = 0; sub < kps; ++sub) {  // base descriptor + start-address offset (bytes >> 4)
            const uint64_t ad0 = adesc0 + (uint64_t)(((uint32_t)st * stageA + (uint32_t)sub * tileA_bytes) >> 4);
            const uint64_t bd0 = bdesc0 + (uint64_t)(((uint32_t)st * stageB + (uint32_t)sub * tileB_bytes) >> 4);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              umma_tf32(tmem_base, ad0 + (uint64_t)(j * 64), bd0 + (uint64_t)(j * 64), idesc,
                        (uint32_t)((kb | sub | j) != 0));
          }
          umma_commit(empty_bar(st));
          if (kb == num_kb - 1) umma_commit(tfull_bar);
        }
        __syncwarp();
        if (++st == stages) { st = 0; ph ^= 1; }
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
