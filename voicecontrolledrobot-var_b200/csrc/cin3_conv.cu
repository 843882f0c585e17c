// Image conv1 of both encoders (3 -> 32 channels, 3x3, pad 1, stride 1 or 2 over uint8 frames)
// as persistent tcgen05 kernels -- the same scheme as cin1_conv.cu.
//
// The direct CUDA-core form (first_conv.cu) spends 27 FMAs per lane per pixel: 32 lanes x 27
// issue slots for ONE pixel, i.e. it is bound by instruction issue at ~20 % of the fp32 peak
// while the layer only has to write 128 B per pixel.  On the tensor core the whole K = 27 is one
// 32-wide k-block: a thread builds the operand row of its pixel with 27 shared-memory loads and
// 8 vector stores, four MMAs finish 128 pixels x 32 channels, and the kernel is bound by the
// output write (forward) or the dY read (weight gradient).
//
// Exactness: the MMA operand is the uint8 value itself (exact in tf32); the 1/255 of
// dataset.py:67-68 is applied to the fp32 accumulator in the epilogue, so -- unlike a tf32
// rounding of x/255 -- no input precision is lost against the fp32 reference.
//
//   tile      = 128 / Q output rows of one image (96 pixels for the 96 x 96 frames)
//   warps 0-3 / 4-7   two builder groups, one tile each in flight; the uint8 patch of the
//               group's next tile is prefetched as 32-bit words into registers
//   warp 8    lane 0: TMA issue (weights / dY) + the MMAs
//   wgrad     D[tap, cout] accumulates in one TMEM tile over all tiles of the CTA; tap row 27 is
//             all ones, so row 27 of D is the bias gradient
#include "cin3_conv.cuh"
#include "engine_host.cuh"

#include <cstdlib>

namespace var {

namespace {
constexpr int kCout = 32, kTaps = 27;
constexpr int kMaxNH = 9;                       // (4 - 1) * 2 + 3 input rows per tile at most
constexpr int kMaxPitch = 132;
constexpr uint32_t kPatchSlot = 3 * kMaxNH * kMaxPitch * 4 + 128 - (3 * kMaxNH * kMaxPitch * 4) % 128;
constexpr uint32_t kABytes = 16384;             // 128 pixels x 32 taps
constexpr uint32_t kWBytes = kCout * 32 * 4;
constexpr uint32_t kDyBytes = 4 * 4096;         // 4 blocks of 32 pixels x 32 channels
constexpr int kThreads = 288;
constexpr int kPW = 7;                          // prefetch words per thread: 3 * 9 * 32 / 128
constexpr size_t kSmemFwd = 1024 + 2 * kABytes + kWBytes + 2 * kPatchSlot + 128;
constexpr size_t kSmemWgrad = 1024 + 2 * kABytes + 2 * kDyBytes + 2 * kPatchSlot + 128;

struct Geo {
  int rt, tpi, nh, pitch, w4, words;  // rows per tile, tiles per image, patch rows / pitch / words
};

__device__ __forceinline__ void group_sync(int g) {
  // literal ids keep the kernel at 3 named barriers (2 CTAs share an SM)
  if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
  else asm volatile("bar.sync 2, 128;" ::: "memory");
}

// word e of the tile patch = 4 consecutive bytes of input row h = p0*ST - 1 + hh, channel c
template <int ST>
__device__ __forceinline__ void patch_prefetch(const Cin3Args& a, const Geo& ge, int n, int p0, int r,
                                               uint32_t (&pre)[kPW]) {
#pragma unroll
  for (int u = 0; u < kPW; ++u) {
    const int e = r + u * 128;
    uint32_t v = 0;
    if (e < ge.words) {
      const int row = e / ge.w4, w4 = e - row * ge.w4;
      const int c = row / ge.nh, hh = row - c * ge.nh;
      const int h = p0 * ST - 1 + hh;
      if (h >= 0 && h < a.H)
        v = __ldg(reinterpret_cast<const uint32_t*>(a.x + (((long long)n * 3 + c) * a.H + h) * a.W) + w4);
    }
    pre[u] = v;
  }
}
__device__ __forceinline__ void patch_store(const Geo& ge, float* patch, int r, const uint32_t (&pre)[kPW]) {
#pragma unroll
  for (int u = 0; u < kPW; ++u) {
    const int e = r + u * 128;
    if (e < ge.words) {
      const int row = e / ge.w4, w4 = e - row * ge.w4;
      float* d = patch + row * ge.pitch + 1 + w4 * 4;  // patch column = input column + 1
      const uint32_t v = pre[u];
      d[0] = (float)(v & 255u); d[1] = (float)((v >> 8) & 255u);
      d[2] = (float)((v >> 16) & 255u); d[3] = (float)(v >> 24);
    }
  }
}

template <int ST, bool WG>
__device__ __forceinline__ void build_row(const float* patch, const Geo& ge, uint32_t sA, int r, int pr, int q,
                                          bool valid) {
  float v[32];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) {
      const float* src = patch + (c * ge.nh + pr * ST + rr) * ge.pitch + q * ST;
#pragma unroll
      for (int s = 0; s < 3; ++s) v[(rr * 3 + s) * 3 + c] = valid ? src[s] : 0.f;
    }
  v[27] = (WG && valid) ? 1.f : 0.f;
  v[28] = v[29] = v[30] = v[31] = 0.f;
#pragma unroll
  for (int cc = 0; cc < 8; ++cc) {
    const uint32_t dst = WG ? sA + (uint32_t)(r >> 5) * 4096u + swz128_32((uint32_t)(r & 31), cc)
                            : sA + swz128((uint32_t)r, cc);
    st_shared_v4(dst, v[cc * 4], v[cc * 4 + 1], v[cc * 4 + 2], v[cc * 4 + 3]);
  }
}

struct TilePos { int n, p0; };
__device__ __forceinline__ TilePos tile_pos(int i, const Geo& ge) {
  const int t = blockIdx.x + i * gridDim.x;
  TilePos tp;
  tp.n = t / ge.tpi;
  tp.p0 = (t - tp.n * ge.tpi) * ge.rt;
  return tp;
}

template <int ST>
__global__ void __launch_bounds__(kThreads, 2)
cin3_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const Cin3Args a, const Geo ge, int tiles) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sW = sA + 2 * kABytes, sP = sW + kWBytes;
  const uint32_t bars = sP + 2 * kPatchSlot;
  auto a_full = [&](int g) { return bars + (uint32_t)g * 8u; };
  auto acc_full = [&](int g) { return bars + 16u + (uint32_t)g * 8u; };
  const uint32_t w_full = bars + 32u, tslot = bars + 40u;
  const int nt = (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;

  if (tid == 0) {
    for (int g = 0; g < 2; ++g) {
      mbar_init(a_full(g), 128);
      mbar_init(acc_full(g), 1);
    }
    mbar_init(w_full, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmW);
  }
  {  // halo columns and unused rows of both patches stay zero for the whole kernel
    float* pz = reinterpret_cast<float*>(gbase + 2 * kABytes + kWBytes);
    for (int i = tid; i < (int)(2 * kPatchSlot / 4); i += kThreads) pz[i] = 0.f;
  }
  if (warp == 8) tmem_alloc(tslot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  if (warp == 8) {
    // converged warp, one elected lane issues (operands stay in uniform registers; see elect_one_sync)
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_full, kWBytes);
      tma_load_2d(sW, &tmW, w_full, 0, 0);
    }
    __syncwarp();
    mbar_wait(w_full, 0);
    const uint32_t idesc = make_idesc_tf32(kCout, 0, 0);
    for (int i = 0; i < nt; ++i) {
      const int g = i & 1;
      mbar_wait(a_full(g), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t a0 = sA + (uint32_t)g * kABytes;
      if (elect_one_sync()) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint64_t ad = make_smem_desc(a0 + (uint32_t)j * 32u, 16u, 1024u);
          const uint64_t bd = make_smem_desc(sW + (uint32_t)j * 32u, 16u, 1024u);
          umma_tf32(tmem_base + (uint32_t)g * 32u, ad, bd, idesc, (uint32_t)(j != 0));
        }
        umma_commit(acc_full(g));
      }
      __syncwarp();
    }
  } else {
    const int g = warp >> 2, r = tid & 127, wq = warp & 3;
    const int pr = r / a.Q, q = r - pr * a.Q;
    float* patch = reinterpret_cast<float*>(gbase + 2 * kABytes + kWBytes + (uint32_t)g * kPatchSlot);
    const uint32_t a0 = sA + (uint32_t)g * kABytes;
    uint32_t pre[kPW];
    if (g < nt) { const TilePos t0 = tile_pos(g, ge); patch_prefetch<ST>(a, ge, t0.n, t0.p0, r, pre); }
    for (int i = g; i < nt; i += 2) {
      const uint32_t ph = (uint32_t)((i >> 1) & 1);
      const TilePos tp = tile_pos(i, ge);
      // (every thread of the group passed acc_full of the previous tile: nobody reads the patch)
      patch_store(ge, patch, r, pre);
      group_sync(g);
      const bool valid = pr < ge.rt && tp.p0 + pr < a.P;
      build_row<ST, false>(patch, ge, a0, r, pr, q, valid);
      fence_proxy_async_smem();
      mbar_arrive(a_full(g));
      if (i + 2 < nt) { const TilePos t2 = tile_pos(i + 2, ge); patch_prefetch<ST>(a, ge, t2.n, t2.p0, r, pre); }
      // ---- epilogue: scale (1/255), bias, ReLU, tf32 rounding, 128 B per pixel
      mbar_wait(acc_full(g), ph);
      tc_fence_after();
      // thread = pixel row (128 B) after tcgen05.ld; the rows go through the group's idle operand
      // buffer so that 8 lanes cover one row and a store instruction writes 512 contiguous bytes
      const uint32_t stg = a0 + (uint32_t)(wq * 32) * 128u;
      {
        float v[32];
        tmem_ld32(tmem_base + (uint32_t)g * 32u + ((uint32_t)(wq * 32) << 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          st_shared_v4(stg + swz128((uint32_t)lane, j), v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
      __syncwarp();
      {
        const int chunk = lane & 7, rsub = lane >> 3;
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias) b4 = __ldg(reinterpret_cast<const float4*>(a.bias + chunk * 4));
        float* out = a.y + ((long long)(tp.n * a.P + tp.p0) * a.Q + wq * 32) * kCout + chunk * 4;
        const int npix = min(ge.rt, a.P - tp.p0) * a.Q;  // valid pixels of this tile
#pragma unroll
        for (int i2 = 0; i2 < 8; ++i2) {
          const int row = i2 * 4 + rsub;
          if (wq * 32 + row < npix) {
            float4 r4 = ld_shared_v4(stg + swz128((uint32_t)row, chunk));
            r4.x = r4.x * a.scale + b4.x; r4.y = r4.y * a.scale + b4.y;
            r4.z = r4.z * a.scale + b4.z; r4.w = r4.w * a.scale + b4.w;
            if (a.relu) {
              r4.x = fmaxf(r4.x, 0.f); r4.y = fmaxf(r4.y, 0.f);
              r4.z = fmaxf(r4.z, 0.f); r4.w = fmaxf(r4.w, 0.f);
            }
            if (a.round_out) {
              r4.x = round_tf32(r4.x); r4.y = round_tf32(r4.y);
              r4.z = round_tf32(r4.z); r4.w = round_tf32(r4.w);
            }
            *reinterpret_cast<float4*>(out + (long long)row * kCout) = r4;
          }
        }
      }
      tc_fence_before();  // orders the TMEM reads before the next a_full arrival
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

template <int ST>
__global__ void __launch_bounds__(kThreads, 2)
cin3_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const Cin3Args a, const Geo ge, int tiles, int mn_sbo,
                  int mn_type) {
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sA = base, sB = sA + 2 * kABytes, sP = sB + 2 * kDyBytes;
  const uint32_t bars = sP + 2 * kPatchSlot;
  auto a_full = [&](int g) { return bars + (uint32_t)g * 8u; };
  auto dy_full = [&](int g) { return bars + 16u + (uint32_t)g * 8u; };
  auto mma_done = [&](int g) { return bars + 32u + (uint32_t)g * 8u; };
  const uint32_t final_bar = bars + 48u, tslot = bars + 56u;
  const int nt = (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;

  if (tid == 0) {
    for (int g = 0; g < 2; ++g) {
      mbar_init(a_full(g), 128);
      mbar_init(dy_full(g), 1);
      mbar_init(mma_done(g), 1);
    }
    mbar_init(final_bar, 1);
    mbar_fence_init();
    tma_prefetch_desc(&tmDY);
  }
  {
    float* pz = reinterpret_cast<float*>(gbase + 2 * kABytes + 2 * kDyBytes);
    for (int i = tid; i < (int)(2 * kPatchSlot / 4); i += kThreads) pz[i] = 0.f;
  }
  if (warp == 8) tmem_alloc(tslot, 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tslot));

  if (warp == 8) {
    {  // converged warp; issue_dy and the MMA bursts run on one elected lane
      auto issue_dy = [&](int i) {
        const TilePos tp = tile_pos(i, ge);
        const int g = i & 1;
        const int m0 = (tp.n * a.P + tp.p0) * a.Q;
        mbar_arrive_expect_tx(dy_full(g), kDyBytes);
        for (int pb = 0; pb < 4; ++pb)
          tma_load_2d(sB + (uint32_t)g * kDyBytes + (uint32_t)pb * 4096u, &tmDY, dy_full(g), 0, m0 + pb * 32);
      };
      if (elect_one_sync())
        for (int i = 0; i < 2 && i < nt; ++i) issue_dy(i);
      __syncwarp();
      const uint32_t idesc = make_idesc_tf32(kCout, 1, 1);
      const uint32_t sbo = (uint32_t)mn_sbo, lt = (uint32_t)mn_type;
      for (int i = 0; i < nt; ++i) {
        const int g = i & 1;
        const uint32_t ph = (uint32_t)((i >> 1) & 1);
        mbar_wait(a_full(g), ph);
        mbar_wait(dy_full(g), ph);
        tc_fence_after();
        const uint32_t a0 = sA + (uint32_t)g * kABytes, b0 = sB + (uint32_t)g * kDyBytes;
        if (elect_one_sync()) {
#pragma unroll
        for (int pb = 0; pb < 4; ++pb)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // the A' operand has ONE 32-tap group: group stride 0 lets rows 32-127 of the M = 128
            // tile re-read it (their D rows are never used)
            const uint64_t ad = make_smem_desc(a0 + (uint32_t)pb * 4096u + (uint32_t)j * 1024u, 0u, sbo, lt);
            const uint64_t bd = make_smem_desc(b0 + (uint32_t)pb * 4096u + (uint32_t)j * 1024u, 4096u, sbo, lt);
            umma_tf32(tmem_base, ad, bd, idesc, (uint32_t)((i | pb | j) != 0));
          }
        umma_commit(mma_done(g));
        }
        __syncwarp();
        if (i >= 1 && i + 1 < nt) {  // the other group's dY buffer is free once tile i-1 retired
          mbar_wait(mma_done(g ^ 1), (uint32_t)(((i - 1) >> 1) & 1));
          if (elect_one_sync()) issue_dy(i + 1);
          __syncwarp();
        }
      }
      if (elect_one_sync()) umma_commit(final_bar);
    }
    __syncwarp();
  } else {
    const int g = warp >> 2, r = tid & 127;
    const int pr = r / a.Q, q = r - pr * a.Q;
    float* patch = reinterpret_cast<float*>(gbase + 2 * kABytes + 2 * kDyBytes + (uint32_t)g * kPatchSlot);
    const uint32_t a0 = sA + (uint32_t)g * kABytes;
    uint32_t pre[kPW];
    if (g < nt) { const TilePos t0 = tile_pos(g, ge); patch_prefetch<ST>(a, ge, t0.n, t0.p0, r, pre); }
    for (int i = g; i < nt; i += 2) {
      const TilePos tp = tile_pos(i, ge);
      // tile i-2 retired: nobody still reads the patch and the operand buffer is free
      if (i >= 2) mbar_wait(mma_done(g), (uint32_t)(((i >> 1) - 1) & 1));
      patch_store(ge, patch, r, pre);
      group_sync(g);
      if (i + 2 < nt) { const TilePos t2 = tile_pos(i + 2, ge); patch_prefetch<ST>(a, ge, t2.n, t2.p0, r, pre); }
      build_row<ST, true>(patch, ge, a0, r, pr, q, pr < ge.rt && tp.p0 + pr < a.P);  // zero rows beyond the tile
      fence_proxy_async_smem();
      mbar_arrive(a_full(g));
    }
    if (warp == 0) {
      // ---- D rows 0-26 = taps, row 27 = bias gradient; columns = output channels
      mbar_wait(final_bar, 0);
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem_base, v);
      tmem_ld_wait();
      if (lane < kTaps) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(a.dw + j * 32 + lane, v[j] * a.scale);
      } else if (lane == kTaps && a.db) {
#pragma unroll
        for (int j = 0; j < 32; ++j) atomicAdd(a.db + j, v[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

bool enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VAR_CIN3"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

Geo make_geo(const Cin3Args& a) {
  Geo ge;
  ge.rt = 128 / a.Q;
  if (ge.rt > 4) ge.rt = 4;
  if (ge.rt > a.P) ge.rt = a.P;
  ge.tpi = (a.P + ge.rt - 1) / ge.rt;
  ge.nh = (ge.rt - 1) * a.stride + 3;
  ge.pitch = a.W + 4;
  ge.w4 = a.W / 4;
  ge.words = 3 * ge.nh * ge.w4;
  return ge;
}
}  // namespace

bool cin3_conv_match(int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph, int pw, int P, int Q,
                     long long sN, long long sH, long long sW, long long sC, const void* x) {
  return enabled() && Cin == 3 && Cout == kCout && R == 3 && S == 3 && ph == 1 && pw == 1 && sh == sw &&
         (sh == 1 || sh == 2) && (W & 3) == 0 && W + 4 <= kMaxPitch && Q >= 32 && Q <= 128 && P >= 1 && sW == 1 &&
         sH == W && sC == (long long)H * W && sN == 3LL * H * W && (reinterpret_cast<uintptr_t>(x) & 3) == 0 &&
         (Q - 1) * sh + 3 <= W + 4;
}

int cin3_conv_fwd(const Cin3Args& a, cudaStream_t st) {
  const Geo ge = make_geo(a);
  const int tiles = a.N * ge.tpi;
  if (tiles <= 0) return VAR_OK;
  CUtensorMap tw;
  int rc = get_tmap_2d(a.w, kCout, 32, 32, kCout, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tw);
  if (rc) return rc;
  VAR_ENSURE_SMEM(cin3_fwd_kernel<1>, kSmemFwd);
  VAR_ENSURE_SMEM(cin3_fwd_kernel<2>, kSmemFwd);
  const int grid = tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs;
  LaunchScope sc(T_GEMM_SCALAR, 2.0 * a.N * a.P * a.Q * kCout * (double)kTaps, st);
  if (a.stride == 1) cin3_fwd_kernel<1><<<grid, kThreads, kSmemFwd, st>>>(tw, a, ge, tiles);
  else cin3_fwd_kernel<2><<<grid, kThreads, kSmemFwd, st>>>(tw, a, ge, tiles);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

int cin3_conv_wgrad(const Cin3Args& a, cudaStream_t st) {
  const Geo ge = make_geo(a);
  const int tiles = a.N * ge.tpi;
  if (tiles <= 0) return VAR_OK;
  CUtensorMap tdy;
  int rc = get_tmap_2d(a.dy, a.N * a.P * a.Q, kCout, kCout, 32, mn_cfg().tma_swizzle, &tdy);
  if (rc) return rc;
  VAR_ENSURE_SMEM(cin3_wgrad_kernel<1>, kSmemWgrad);
  VAR_ENSURE_SMEM(cin3_wgrad_kernel<2>, kSmemWgrad);
  const int grid = tiles < 2 * kNumSMs ? tiles : 2 * kNumSMs;
  LaunchScope sc(T_WGRAD, 2.0 * a.N * a.P * a.Q * kCout * (double)kTaps, st);
  if (a.stride == 1) cin3_wgrad_kernel<1><<<grid, kThreads, kSmemWgrad, st>>>(tdy, a, ge, tiles, mn_cfg().sbo, mn_cfg().type);
  else cin3_wgrad_kernel<2><<<grid, kThreads, kSmemWgrad, st>>>(tdy, a, ge, tiles, mn_cfg().sbo, mn_cfg().type);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var
