// Image conv1 of both encoders (3 -> 32 channels, 3x3, pad 1, stride 1 or 2;
// models/pretext/arm_pretext_model.py:11, ai2thor_pretext_model.py:15) as plain fp32 CUDA-core
// kernels.  K = 27 fills one tensor-core k-block at 84 % and each 128-pixel MMA tile would pay
// a TMEM allocation, barrier set-up and an epilogue for 7 kFLOP per pixel: the layer is bound by
// its 128-byte-per-pixel output write, so the direct form (lane = output channel, weights in
// registers, input rows staged once in shared memory) is both simpler and several times
// faster.  The uint8 -> float /255 conversion of dataset.py:67-68 happens while staging.
#include "first_conv.cuh"

#include <cstdlib>

namespace var {

namespace {
constexpr int kCout = 32, kK = 27, kThreads = 256, kWarps = 8;

template <bool U8>
__device__ __forceinline__ float load_px(const void* src, long long idx, float scale) {
  if constexpr (U8) return (float)reinterpret_cast<const uint8_t*>(src)[idx] * scale;
  return reinterpret_cast<const float*>(src)[idx] * scale;
}

// Stages input rows [h_lo, h_lo + nh) of image n (3 channels, zero padded by one column on each
// side) as patch[(c*nh + hh)*wpad + (w + 1)].
template <bool U8>
__device__ __forceinline__ void stage_rows(const FirstConvArgs& a, int n, int h_lo, int nh, float* patch) {
  const int wpad = a.W + 2;
  const int total = 3 * nh * wpad;
  for (int i = threadIdx.x; i < total; i += kThreads) {
    const int ww = i % wpad;
    const int t = i / wpad;
    const int hh = t % nh, c = t / nh;
    const int h = h_lo + hh, w = ww - 1;
    float x = 0.f;
    if (h >= 0 && h < a.H && w >= 0 && w < a.W)
      x = load_px<U8>(a.x, (long long)n * a.sN + (long long)c * a.sC + (long long)h * a.sH + (long long)w * a.sW,
                      a.scale);
    patch[i] = x;
  }
}
}  // namespace

// grid = N * ceil(P / rows) ; CTA = `rows` output rows of one image.
template <bool U8>
__global__ void __launch_bounds__(kThreads) first_conv_fwd_kernel(FirstConvArgs a) {
  extern __shared__ float patch[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = (a.P + a.rows - 1) / a.rows;
  const int n = blockIdx.x / tpi;
  const int p0 = (blockIdx.x - n * tpi) * a.rows;
  const int nrows = min(a.rows, a.P - p0);
  const int nh = (nrows - 1) * a.stride + 3, wpad = a.W + 2;
  stage_rows<U8>(a, n, p0 * a.stride - 1, nh, patch);
  float w[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) w[k] = a.w[lane * a.kpad + k];  // k = (r*3 + s)*3 + c
  const float b = a.bias ? a.bias[lane] : 0.f;
  __syncthreads();
  const int npix = nrows * a.Q;
  float* out = a.y + ((long long)(n * a.P + p0) * a.Q) * kCout;
  for (int i = warp; i < npix; i += kWarps) {
    const int pr = i / a.Q, q = i - pr * a.Q;
    const float* x0 = patch + (pr * a.stride) * wpad + q * a.stride;
    float acc = b;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s)
#pragma unroll
        for (int c = 0; c < 3; ++c) acc = fmaf(x0[(c * nh + r) * wpad + s], w[(r * 3 + s) * 3 + c], acc);
    if (a.relu) acc = fmaxf(acc, 0.f);
    out[(long long)i * kCout + lane] = a.round_out ? round_tf32(acc) : acc;
  }
}

// Persistent grid; every CTA walks (image, row block) items, accumulating dW[c][27] and db[c] in
// registers (lane = c), then one shared-memory reduction over the 8 warps and 28*32 atomics.
template <bool U8>
__global__ void __launch_bounds__(kThreads) first_conv_wgrad_kernel(FirstConvArgs a, int items) {
  extern __shared__ float patch[];
  __shared__ float red[kWarps][28][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = (a.P + a.rows - 1) / a.rows;
  const int wpad = a.W + 2;
  float acc[kK], accb = 0.f;
#pragma unroll
  for (int k = 0; k < kK; ++k) acc[k] = 0.f;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = item / tpi;
    const int p0 = (item - n * tpi) * a.rows;
    const int nrows = min(a.rows, a.P - p0);
    const int nh = (nrows - 1) * a.stride + 3;
    __syncthreads();
    stage_rows<U8>(a, n, p0 * a.stride - 1, nh, patch);
    __syncthreads();
    const int npix = nrows * a.Q;
    const float* dy = a.dy + ((long long)(n * a.P + p0) * a.Q) * kCout;
    for (int i = warp; i < npix; i += kWarps) {
      const int pr = i / a.Q, q = i - pr * a.Q;
      const float* x0 = patch + (pr * a.stride) * wpad + q * a.stride;
      const float g = dy[(long long)i * kCout + lane];
      accb += g;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s)
#pragma unroll
          for (int c = 0; c < 3; ++c)
            acc[(r * 3 + s) * 3 + c] = fmaf(x0[(c * nh + r) * wpad + s], g, acc[(r * 3 + s) * 3 + c]);
    }
  }
#pragma unroll
  for (int k = 0; k < kK; ++k) red[warp][k][lane] = acc[k];
  red[warp][27][lane] = accb;
  __syncthreads();
  for (int i = threadIdx.x; i < 28 * 32; i += kThreads) {
    const int k = i >> 5, c = i & 31;
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < kWarps; ++wq) s += red[wq][k][c];
    if (k < kK) atomicAdd(a.dw + c * a.kpad + k, s);
    else if (a.db) atomicAdd(a.db + c, s);
  }
}

// ---------------------------------------------------------------------------------------
// 4-pixel register tiles (Q % 4 == 0): a warp owns 4 adjacent output pixels at a time, lane =
// output channel.  The 3 x 3 x 3 window of the 4 pixels is 6 (stride 1) or 9 (stride 2)
// consecutive patch columns per (channel, row): two / three broadcast vector loads instead of
// 12 scalar ones, so the kernels are FMA-bound rather than shared-memory-issue-bound.
// Patch rows are pitched to a multiple of 4 floats; patch column = input column + 1.
template <bool U8>
__device__ __forceinline__ void stage_rows4(const FirstConvArgs& a, int n, int h_lo, int nh, int wp, float* patch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < 3 * nh; row += kWarps) {  // one (channel, input row) per warp pass
    const int c = row / nh, hh = row - c * nh;
    const int h = h_lo + hh;
    const bool hok = h >= 0 && h < a.H;
    const long long base = (long long)n * a.sN + (long long)c * a.sC + (long long)h * a.sH;
    for (int ww = lane; ww < wp; ww += 32) {
      const int w = ww - 1;
      float x = 0.f;
      if (hok && w >= 0 && w < a.W) x = load_px<U8>(a.x, base + (long long)w * a.sW, a.scale);
      patch[row * wp + ww] = x;
    }
  }
}

template <int STRIDE>
__device__ __forceinline__ void load_window(const float* row, float (&xv)[3 * STRIDE + 3]) {
  // row is 16-byte aligned: STRIDE 1 -> 6 values, STRIDE 2 -> 9 values
  const float4 v0 = *reinterpret_cast<const float4*>(row);
  xv[0] = v0.x; xv[1] = v0.y; xv[2] = v0.z; xv[3] = v0.w;
  if constexpr (STRIDE == 1) {
    const float2 v1 = *reinterpret_cast<const float2*>(row + 4);
    xv[4] = v1.x; xv[5] = v1.y;
  } else {
    const float4 v1 = *reinterpret_cast<const float4*>(row + 4);
    xv[4] = v1.x; xv[5] = v1.y; xv[6] = v1.z; xv[7] = v1.w;
    xv[8] = row[8];
  }
}

template <bool U8, int STRIDE>
__global__ void __launch_bounds__(kThreads) first_conv_fwd4_kernel(FirstConvArgs a, int items, int wp) {
  extern __shared__ __align__(16) float patch[];
  __shared__ float wsm[kCout][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = (a.P + a.rows - 1) / a.rows;
  // weights: one coalesced read of the packed [32][kpad] block per CTA, transposed through smem
  for (int i = threadIdx.x; i < kCout * 32; i += kThreads) wsm[i >> 5][i & 31] = (i & 31) < kK ? a.w[(i >> 5) * a.kpad + (i & 31)] : 0.f;
  __syncthreads();
  float w[kK];
#pragma unroll
  for (int k = 0; k < kK; ++k) w[k] = wsm[lane][k];  // k = (r*3 + s)*3 + c
  const float b = a.bias ? a.bias[lane] : 0.f;
  const int qg = a.Q >> 2;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = item / tpi;
    const int p0 = (item - n * tpi) * a.rows;
    const int nrows = min(a.rows, a.P - p0);
    const int nh = (nrows - 1) * STRIDE + 3;
    __syncthreads();
    stage_rows4<U8>(a, n, p0 * STRIDE - 1, nh, wp, patch);
    __syncthreads();
    const int nitems = nrows * qg;
    float* out = a.y + ((long long)(n * a.P + p0) * a.Q) * kCout;
    for (int i = warp; i < nitems; i += kWarps) {
      const int pr = i / qg, q0 = (i - pr * qg) << 2;
      float acc[4] = {b, b, b, b};
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          float xv[3 * STRIDE + 3];
          load_window<STRIDE>(patch + (c * nh + pr * STRIDE + r) * wp + q0 * STRIDE, xv);
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(xv[j * STRIDE + s], w[(r * 3 + s) * 3 + c], acc[j]);
        }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[j];
        if (a.relu) v = fmaxf(v, 0.f);
        out[(long long)(pr * a.Q + q0 + j) * kCout + lane] = a.round_out ? round_tf32(v) : v;
      }
    }
  }
}

template <bool U8, int STRIDE>
__global__ void __launch_bounds__(kThreads) first_conv_wgrad4_kernel(FirstConvArgs a, int items, int wp) {
  extern __shared__ __align__(16) float patch[];
  __shared__ float red[kWarps][28][33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = (a.P + a.rows - 1) / a.rows;
  const int qg = a.Q >> 2;
  float acc[kK], accb = 0.f;
#pragma unroll
  for (int k = 0; k < kK; ++k) acc[k] = 0.f;
  for (int item = blockIdx.x; item < items; item += gridDim.x) {
    const int n = item / tpi;
    const int p0 = (item - n * tpi) * a.rows;
    const int nrows = min(a.rows, a.P - p0);
    const int nh = (nrows - 1) * STRIDE + 3;
    __syncthreads();
    stage_rows4<U8>(a, n, p0 * STRIDE - 1, nh, wp, patch);
    __syncthreads();
    const float* dy = a.dy + ((long long)(n * a.P + p0) * a.Q) * kCout;
    for (int i = warp; i < nrows * qg; i += kWarps) {
      const int pr = i / qg, q0 = (i - pr * qg) << 2;
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = dy[(long long)(pr * a.Q + q0 + j) * kCout + lane];
      accb += (g[0] + g[1]) + (g[2] + g[3]);
#pragma unroll
      for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          float xv[3 * STRIDE + 3];
          load_window<STRIDE>(patch + (c * nh + pr * STRIDE + r) * wp + q0 * STRIDE, xv);
#pragma unroll
          for (int s = 0; s < 3; ++s)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc[(r * 3 + s) * 3 + c] = fmaf(xv[j * STRIDE + s], g[j], acc[(r * 3 + s) * 3 + c]);
        }
    }
  }
#pragma unroll
  for (int k = 0; k < kK; ++k) red[warp][k][lane] = acc[k];
  red[warp][27][lane] = accb;
  __syncthreads();
  for (int i = threadIdx.x; i < 28 * 32; i += kThreads) {
    const int k = i >> 5, c = i & 31;
    float s = 0.f;
#pragma unroll
    for (int wq = 0; wq < kWarps; ++wq) s += red[wq][k][c];
    if (k < kK) atomicAdd(a.dw + c * a.kpad + k, s);
    else if (a.db) atomicAdd(a.db + c, s);
  }
}

static bool tile4_ok(const FirstConvArgs& a) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VAR_FIRST_CONV4"); on = (e && e[0] == '0') ? 0 : 1; }
  return on && (a.Q & 3) == 0 && (a.stride == 1 || a.stride == 2);
}
// patch pitch: covers patch columns 0 .. (Q-4)*stride + 3*stride + 2, rounded up to 4 floats
static int pitch4(const FirstConvArgs& a) {
  int need = (a.Q - 4) * a.stride + 3 * a.stride + 3;
  if (need < a.W + 2) need = a.W + 2;
  return (need + 3) & ~3;
}

static int rows_for(const FirstConvArgs& a) {
  // ~256-400 output pixels per CTA keeps the patch a few KB and the grid large
  int rows = 384 / a.Q;
  if (rows < 1) rows = 1;
  if (rows > a.P) rows = a.P;
  return rows;
}

int first_conv_fwd(FirstConvArgs a, int u8, cudaStream_t st) {
  a.rows = rows_for(a);
  const int tpi = (a.P + a.rows - 1) / a.rows;
  const size_t smem = (size_t)3 * ((a.rows - 1) * a.stride + 3) * (a.W + 2) * 4;
  LaunchScope sc(T_GEMM_SCALAR, 2.0 * a.N * a.P * a.Q * kCout * (double)kK, st);
  if (tile4_ok(a)) {
    const int wp = pitch4(a);
    const size_t smem4 = (size_t)3 * ((a.rows - 1) * a.stride + 3) * wp * 4;
    const int items = a.N * tpi;
    const int grid = items < 8 * kNumSMs ? items : 8 * kNumSMs;
    if (u8 && a.stride == 1) first_conv_fwd4_kernel<true, 1><<<grid, kThreads, smem4, st>>>(a, items, wp);
    else if (u8) first_conv_fwd4_kernel<true, 2><<<grid, kThreads, smem4, st>>>(a, items, wp);
    else if (a.stride == 1) first_conv_fwd4_kernel<false, 1><<<grid, kThreads, smem4, st>>>(a, items, wp);
    else first_conv_fwd4_kernel<false, 2><<<grid, kThreads, smem4, st>>>(a, items, wp);
    VAR_CUDA_CHECK(cudaGetLastError());
    return VAR_OK;
  }
  if (u8) first_conv_fwd_kernel<true><<<a.N * tpi, kThreads, smem, st>>>(a);
  else first_conv_fwd_kernel<false><<<a.N * tpi, kThreads, smem, st>>>(a);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

int first_conv_wgrad(FirstConvArgs a, int u8, cudaStream_t st) {
  a.rows = rows_for(a);
  const int tpi = (a.P + a.rows - 1) / a.rows;
  const int items = a.N * tpi;
  const size_t smem = (size_t)3 * ((a.rows - 1) * a.stride + 3) * (a.W + 2) * 4;
  int grid = 4 * kNumSMs;
  if (grid > items) grid = items;
  // static reduction buffer (29.6 KB) + patch can pass the 48 KB default
  VAR_ENSURE_SMEM(first_conv_wgrad_kernel<true>, 64 * 1024);
  VAR_ENSURE_SMEM(first_conv_wgrad_kernel<false>, 64 * 1024);
  LaunchScope sc(T_WGRAD, 2.0 * a.N * a.P * a.Q * kCout * (double)kK, st);
  if (tile4_ok(a)) {
    const int wp = pitch4(a);
    const size_t smem4 = (size_t)3 * ((a.rows - 1) * a.stride + 3) * wp * 4;
    if (smem4 + sizeof(float) * kWarps * 28 * 33 <= 48 * 1024) {
      if (u8 && a.stride == 1) first_conv_wgrad4_kernel<true, 1><<<grid, kThreads, smem4, st>>>(a, items, wp);
      else if (u8) first_conv_wgrad4_kernel<true, 2><<<grid, kThreads, smem4, st>>>(a, items, wp);
      else if (a.stride == 1) first_conv_wgrad4_kernel<false, 1><<<grid, kThreads, smem4, st>>>(a, items, wp);
      else first_conv_wgrad4_kernel<false, 2><<<grid, kThreads, smem4, st>>>(a, items, wp);
      VAR_CUDA_CHECK(cudaGetLastError());
      return VAR_OK;
    }
  }
  if (u8) first_conv_wgrad_kernel<true><<<grid, kThreads, smem, st>>>(a, items);
  else first_conv_wgrad_kernel<false><<<grid, kThreads, smem, st>>>(a, items);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var
