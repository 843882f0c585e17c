// HBM-bound helper kernels around the tcgen05 engine: weight (un)packing between
// the reference state_dict layout and the engine's K-major packed layout, NHWC
// 2x2 max pooling (forward / backward with the ReLU mask folded in), the GRU
// cell backward, fused Adam(+L2) that also refreshes the tf32 operand copy, and
// small utilities.  All kernels are grid-stride, 128-bit vectorised where the
// layout allows, and stream ordered.
#include "aux_kernels.cuh"

namespace var {

static inline int grid_for(long long work, int threads, int max_waves = 8) {
  long long g = (work + threads - 1) / threads;
  const long long cap = (long long)kNumSMs * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---------------------------------------------------------------------------
// Weight packing.  Reference tensor [O][I][R][S] (OIHW; a Linear that follows a
// NCHW flatten is [O][C][H][W]; a plain Linear is R=S=1) <-> packed
// [O][kpad], k = (r*S + s)*I + c.  Columns k >= K are zero.
// ---------------------------------------------------------------------------
__global__ void pack_w_kernel(const float* __restrict__ ref, float* __restrict__ master,
                              float* __restrict__ mma, int O, int I, int R, int S, int kpad) {
  const long long total = (long long)O * kpad;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % kpad);
    const int o = (int)(idx / kpad);
    float v = 0.f;
    if (k < R * S * I) {
      const int c = k % I, rs = k / I, s = rs % S, r = rs / S;
      v = ref[(((long long)o * I + c) * R + r) * S + s];
    }
    master[idx] = v;
    if (mma) mma[idx] = round_tf32(v);
  }
}

__global__ void unpack_w_kernel(const float* __restrict__ packed, float* __restrict__ ref, int O,
                                int I, int R, int S, int kpad) {
  const long long total = (long long)O * I * R * S;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(idx % S);
    long long t = idx / S;
    const int r = (int)(t % R); t /= R;
    const int c = (int)(t % I);
    const int o = (int)(t / I);
    ref[idx] = packed[(long long)o * kpad + (r * S + s) * I + c];
  }
}

int pack_weight(const float* ref, float* master, float* mma, int O, int I, int R, int S, int kpad,
                cudaStream_t st) {
  const long long total = (long long)O * kpad;
  LaunchScope sc(T_MISC, 0, st);
  pack_w_kernel<<<grid_for(total, 256), 256, 0, st>>>(ref, master, mma, O, I, R, S, kpad);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}
int unpack_weight(const float* packed, float* ref, int O, int I, int R, int S, int kpad,
                  cudaStream_t st) {
  const long long total = (long long)O * I * R * S;
  LaunchScope sc(T_MISC, 0, st);
  unpack_w_kernel<<<grid_for(total, 256), 256, 0, st>>>(packed, ref, O, I, R, S, kpad);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

__global__ void round_copy_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                  long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = round_tf32(src[i]);
}
int round_copy(const float* src, float* dst, long long n, cudaStream_t st) {
  LaunchScope sc(T_MISC, 0, st);
  round_copy_kernel<<<grid_for(n, 256), 256, 0, st>>>(src, dst, n);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// ---------------------------------------------------------------------------
// 16-bit operand helpers (the f16 conv region):
//  * cvt_f16: fp32 -> f16 copy (weights; n % 4 == 0)
//  * grad_scale_prepare + cvt_f16_scaled: the gradient entering the region is stored as f16 times a
//    power-of-two scale chosen on the device from its largest magnitude (so that the largest element
//    sits at ~2^12 of f16's 65504 range and ~2^-36 of that magnitude is still representable); the
//    region's kernels undo the scale where gradients leave it (weight gradients, the fp32 dX).
//    scale[0] = S, scale[1] = 1/S.
// ---------------------------------------------------------------------------
__global__ void cvt_f16_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
  }
}
int cvt_f16(const float* src, void* dst, long long n, cudaStream_t st) {
  if (n & 3) return VAR_ERR_ARG;
  LaunchScope sc(T_MISC, 0, st);
  cvt_f16_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src), reinterpret_cast<uint2*>(dst), n / 4);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}
__global__ void amax_kernel(const float4* __restrict__ src, long long n4, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));  // non-negative floats order like uints
}
__global__ void cvt_f16_scaled_kernel(const float4* __restrict__ src, uint2* __restrict__ dst, long long n4,
                                      const unsigned int* __restrict__ amax_bits, float* __restrict__ scale) {
  // S = 2^(12 - ceil(log2(amax))): exact power of two, identical in every thread
  const float amax = __uint_as_float(*amax_bits);
  float S = 1.f;
  if (amax > 0.f && amax < 3.0e38f) {
    int e;
    frexpf(amax, &e);           // amax = f * 2^e, f in [0.5, 1)
    e = 12 - e;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    S = ldexpf(1.f, e);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { scale[0] = S; scale[1] = 1.f / S; }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = src[i];
    dst[i] = make_uint2(pack_f16x2(v.x * S, v.y * S), pack_f16x2(v.z * S, v.w * S));
  }
}
int grad_to_f16_scaled(const float* src, void* dst, long long n, float* scale, unsigned int* amax_bits, cudaStream_t st) {
  if (n & 3) return VAR_ERR_ARG;
  VAR_CUDA_CHECK(cudaMemsetAsync(amax_bits, 0, sizeof(unsigned int), st));
  {
    LaunchScope sc(T_MISC, 0, st);
    amax_kernel<<<grid_for(n / 4, 256, 4), 256, 0, st>>>(reinterpret_cast<const float4*>(src), n / 4, amax_bits);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  {
    LaunchScope sc(T_MISC, 0, st);
    cvt_f16_scaled_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(src),
                                                                  reinterpret_cast<uint2*>(dst), n / 4, amax_bits, scale);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// Last BPTT step -> 16-bit recurrence / input-projection operands with ONE gradient scale for both directions:
// amax over the dgi slices (|dgi| >= |dgh| element-wise: dgh_n = dgi_n * r), then dgh[d] ([rows, cols] dense) and the
// dgi[d] slices ([rows] x cols, row pitches ldgi / ldgi_h) are stored as f16 * S.
__global__ void amax_rows_kernel(const float* __restrict__ s0, const float* __restrict__ s1, int rows, int cols4,
                                 long long ld, unsigned int* __restrict__ amax_bits) {
  float m = 0.f;
  const long long total = 2LL * rows * cols4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cols4);
    const long long r = i / cols4;
    const float* src = r < rows ? s0 : s1;
    const float4 v = *reinterpret_cast<const float4*>(src + (r < rows ? r : r - rows) * ld + 4 * c);
    m = fmaxf(fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(amax_bits, __float_as_uint(m));
}
struct GruLastArgs {
  const float* dgh[2]; uint16_t* dgh_h[2];
  const float* dgi[2]; uint16_t* dgi_h[2];
  long long ldgi, ldgi_h;
  int rows, cols4;
};
__global__ void gru_last_cvt_kernel(const GruLastArgs a, const unsigned int* __restrict__ amax_bits,
                                    float* __restrict__ scale) {
  const float amax = __uint_as_float(*amax_bits);
  float S = 1.f;
  if (amax > 0.f && amax < 3.0e38f) {
    int e;
    frexpf(amax, &e);
    e = 12 - e;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
    S = ldexpf(1.f, e);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { scale[0] = S; scale[1] = 1.f / S; }
  const long long per = (long long)a.rows * a.cols4, total = 4 * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int which = (int)(i / per);      // 0, 1: dgh of direction 0, 1; 2, 3: dgi
    const long long rem = i - which * per;
    const int c = (int)(rem % a.cols4);
    const long long r = rem / a.cols4;
    const int d = which & 1;
    const float* src = which < 2 ? a.dgh[d] + r * (4LL * a.cols4) : a.dgi[d] + r * a.ldgi;
    uint16_t* dst = which < 2 ? a.dgh_h[d] + r * (4LL * a.cols4) : a.dgi_h[d] + r * a.ldgi_h;
    const float4 v = *reinterpret_cast<const float4*>(src + 4 * c);
    *reinterpret_cast<uint2*>(dst + 4 * c) = make_uint2(pack_f16x2(v.x * S, v.y * S), pack_f16x2(v.z * S, v.w * S));
  }
}
int gru_last_to_f16(const float* const dgh[2], void* const dgh_h[2], const float* const dgi[2], long long ldgi,
                    void* const dgi_h[2], long long ldgi_h, int rows, int cols, float* scale, unsigned int* amax_bits,
                    cudaStream_t st) {
  if (cols & 3) return VAR_ERR_ARG;
  VAR_CUDA_CHECK(cudaMemsetAsync(amax_bits, 0, sizeof(unsigned int), st));
  {
    LaunchScope sc(T_MISC, 0, st);
    amax_rows_kernel<<<grid_for(2LL * rows * (cols / 4), 256, 4), 256, 0, st>>>(dgi[0], dgi[1], rows, cols / 4, ldgi, amax_bits);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  GruLastArgs a;
  for (int d = 0; d < 2; ++d) {
    a.dgh[d] = dgh[d]; a.dgh_h[d] = reinterpret_cast<uint16_t*>(dgh_h[d]);
    a.dgi[d] = dgi[d]; a.dgi_h[d] = reinterpret_cast<uint16_t*>(dgi_h[d]);
  }
  a.ldgi = ldgi; a.ldgi_h = ldgi_h; a.rows = rows; a.cols4 = cols / 4;
  {
    LaunchScope sc(T_MISC, 0, st);
    gru_last_cvt_kernel<<<grid_for(4LL * rows * (cols / 4), 256), 256, 0, st>>>(a, amax_bits, scale);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// dst[c][col0 + r] = f16(src[r][c]): fp32 [R, C] -> f16 [C, dst_ld] (transposed, at column offset col0);
// 32 x 32 tiles through shared memory
__global__ void cvt_f16_transpose_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, int R, int C,
                                         long long dst_ld, int col0) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < C) ? src[(long long)r * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < C) dst[(long long)c * dst_ld + col0 + r] = (uint16_t)(pack_f16x2(tile[threadIdx.x][i], 0.f) & 0xFFFFu);
  }
}
int cvt_f16_transpose(const float* src, void* dst, int R, int C, long long dst_ld, int dst_col0, cudaStream_t st) {
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  LaunchScope sc(T_MISC, 0, st);
  cvt_f16_transpose_kernel<<<grid, block, 0, st>>>(src, reinterpret_cast<uint16_t*>(dst), R, C, dst_ld, dst_col0);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// ---------------------------------------------------------------------------
// 2x2/2 max pooling, NHWC, C % 4 == 0 (nn.MaxPool2d(2, 2) of
// models/pretext/ai2thor_pretext_model.py:9-11).
// ---------------------------------------------------------------------------
__global__ void pool_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, int N, int H,
                                int W, int C4) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * C4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4);
    long long t = idx / C4;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const long long base = (((long long)n * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const float4 a = x[base], b = x[base + C4], d = x[base + (long long)W * C4],
                 e = x[base + (long long)W * C4 + C4];
    float4 m;
    m.x = fmaxf(fmaxf(a.x, b.x), fmaxf(d.x, e.x));
    m.y = fmaxf(fmaxf(a.y, b.y), fmaxf(d.y, e.y));
    m.z = fmaxf(fmaxf(a.z, b.z), fmaxf(d.z, e.z));
    m.w = fmaxf(fmaxf(a.w, b.w), fmaxf(d.w, e.w));
    y[idx] = m;
  }
}

// dx = dy routed to the first arg-max of each window, and zero where x <= 0
// (x is a post-ReLU activation, so this also applies the ReLU backward).
__device__ __forceinline__ void route(float a, float b, float d, float e, float g, float& oa,
                                      float& ob, float& od, float& oe) {
  const float m = fmaxf(fmaxf(a, b), fmaxf(d, e));
  oa = ob = od = oe = 0.f;
  if (!(m > 0.f)) return;
  if (a == m) oa = g;
  else if (b == m) ob = g;
  else if (d == m) od = g;
  else oe = g;
}
__global__ void pool_bwd_kernel(const float4* __restrict__ x, const float4* __restrict__ dy,
                                float4* __restrict__ dx, int N, int H, int W, int C4) {
  const int Ho = H >> 1, Wo = W >> 1;
  const long long total = (long long)N * Ho * Wo * C4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % C4);
    long long t = idx / C4;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const long long base = (((long long)n * H + 2 * ho) * W + 2 * wo) * C4 + c;
    const long long o1 = base + C4, o2 = base + (long long)W * C4, o3 = o2 + C4;
    const float4 a = x[base], b = x[o1], d = x[o2], e = x[o3], g = dy[idx];
    float4 ra, rb, rd, re;
    route(a.x, b.x, d.x, e.x, g.x, ra.x, rb.x, rd.x, re.x);
    route(a.y, b.y, d.y, e.y, g.y, ra.y, rb.y, rd.y, re.y);
    route(a.z, b.z, d.z, e.z, g.z, ra.z, rb.z, rd.z, re.z);
    route(a.w, b.w, d.w, e.w, g.w, ra.w, rb.w, rd.w, re.w);
    dx[base] = ra; dx[o1] = rb; dx[o2] = rd; dx[o3] = re;
  }
}

int maxpool_fwd(const float* x, float* y, int N, int H, int W, int C, cudaStream_t st) {
  if ((C & 3) || (H & 1) || (W & 1)) return VAR_ERR_UNSUPPORTED;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
  LaunchScope sc(T_POOL, 0, st);
  pool_fwd_kernel<<<grid_for(total, 256, 16), 256, 0, st>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), N, H, W, C / 4);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}
int maxpool_bwd(const float* x, const float* dy, float* dx, int N, int H, int W, int C,
                cudaStream_t st) {
  if ((C & 3) || (H & 1) || (W & 1)) return VAR_ERR_UNSUPPORTED;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 4);
  LaunchScope sc(T_POOL, 0, st);
  pool_bwd_kernel<<<grid_for(total, 256, 16), 256, 0, st>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy),
      reinterpret_cast<float4*>(dx), N, H, W, C / 4);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// ---------------------------------------------------------------------------
// GRU cell backward (torch.nn.GRU semantics, gate order r, z, n):
//   h' = (1-z) n + z h ;  n = tanh(x_n + r * hn) ; hn = W_hn h + b_hn
// Given dh' it emits, per direction (blockIdx.y),
//   dgi = [d r_pre, d z_pre, d n_pre]      (grad wrt W_ih x + b_ih)   -> [B, ldgi] slot
//   dgh = [d r_pre, d z_pre, d n_pre * r]  (grad wrt W_hh h + b_hh)   -> [B, 3H]
//   dhd = dh' * z                          (direct path to h)
// Both gate gradients are stored tf32-rounded: they are MMA operands next.
// ---------------------------------------------------------------------------
__global__ void gru_cell_bwd_kernel(GruBwdArgs a0, GruBwdArgs a1, int B, int Hd) {
  const GruBwdArgs& a = blockIdx.y == 0 ? a0 : a1;
  const int H4 = Hd >> 2;
  const long long total = (long long)B * H4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % H4) * 4;
    const long long b = idx / H4;
    const float4 dh = *reinterpret_cast<const float4*>(a.dh + b * Hd + j);
    const float* gt = a.gates + b * 3 * Hd + j;
    const float4 r = *reinterpret_cast<const float4*>(gt);
    const float4 z = *reinterpret_cast<const float4*>(gt + Hd);
    const float4 n = *reinterpret_cast<const float4*>(gt + 2 * Hd);
    const float4 hn = *reinterpret_cast<const float4*>(a.hn_save + b * Hd + j);
    const float4 hp = *reinterpret_cast<const float4*>(a.hprev + b * Hd + j);
    float dr[4], dz[4], dn[4], dnr[4], dd[4];
    const float dh_[4] = {dh.x, dh.y, dh.z, dh.w}, r_[4] = {r.x, r.y, r.z, r.w},
                z_[4] = {z.x, z.y, z.z, z.w}, n_[4] = {n.x, n.y, n.z, n.w},
                hn_[4] = {hn.x, hn.y, hn.z, hn.w}, hp_[4] = {hp.x, hp.y, hp.z, hp.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float dnn = dh_[u] * (1.f - z_[u]);
      const float dzz = dh_[u] * (hp_[u] - n_[u]);
      const float dnp = dnn * (1.f - n_[u] * n_[u]);
      const float dzp = dzz * z_[u] * (1.f - z_[u]);
      const float drp = dnp * hn_[u] * r_[u] * (1.f - r_[u]);
      dr[u] = round_tf32(drp);
      dz[u] = round_tf32(dzp);
      dn[u] = round_tf32(dnp);
      dnr[u] = round_tf32(dnp * r_[u]);
      dd[u] = dh_[u] * z_[u];
    }
    float* gi = a.dgi + b * a.ldgi + j;
    *reinterpret_cast<float4*>(gi) = make_float4(dr[0], dr[1], dr[2], dr[3]);
    *reinterpret_cast<float4*>(gi + Hd) = make_float4(dz[0], dz[1], dz[2], dz[3]);
    *reinterpret_cast<float4*>(gi + 2 * Hd) = make_float4(dn[0], dn[1], dn[2], dn[3]);
    float* gh = a.dgh + b * 3 * Hd + j;
    *reinterpret_cast<float4*>(gh) = make_float4(dr[0], dr[1], dr[2], dr[3]);
    *reinterpret_cast<float4*>(gh + Hd) = make_float4(dz[0], dz[1], dz[2], dz[3]);
    *reinterpret_cast<float4*>(gh + 2 * Hd) = make_float4(dnr[0], dnr[1], dnr[2], dnr[3]);
    *reinterpret_cast<float4*>(a.dhd + b * Hd + j) = make_float4(dd[0], dd[1], dd[2], dd[3]);
  }
}

int gru_cell_bwd(const GruBwdArgs& a0, const GruBwdArgs& a1, int ndir, int B, int Hd,
                 cudaStream_t st) {
  if (Hd & 3) return VAR_ERR_UNSUPPORTED;
  dim3 grid(grid_for((long long)B * (Hd / 4), 256), ndir);
  LaunchScope sc(T_GRU_CELL_BWD, 0, st);
  gru_cell_bwd_kernel<<<grid, 256, 0, st>>>(a0, a1, B, Hd);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// ---------------------------------------------------------------------------
// Adam with L2 folded into the gradient (torch.optim.Adam(weight_decay) as used
// at VAR/pretext_VAR.py:33-35), on the flat packed parameter buffer.  Operation
// order follows torch's single-tensor path: lerp, mul+addcmul, sqrt/bc2_sqrt+eps,
// addcdiv.  Also writes the tf32-rounded operand copy consumed by the MMAs.
// ---------------------------------------------------------------------------
__global__ void adam_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                            float4* __restrict__ m, float4* __restrict__ v,
                            float4* __restrict__ pr, long long n4, float beta1, float beta2,
                            float eps, float wd, float step_size, float bc2_sqrt, float gscale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 P = p[i], G = g[i], M = m[i], V = v[i], R;
    float* pp = &P.x; const float* gg = &G.x; float* mm = &M.x; float* vv = &V.x; float* rr = &R.x;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float grad = gg[u] * gscale + wd * pp[u];
      mm[u] = mm[u] + (grad - mm[u]) * (1.f - beta1);
      vv[u] = vv[u] * beta2 + (1.f - beta2) * grad * grad;
      const float denom = sqrtf(vv[u]) / bc2_sqrt + eps;
      pp[u] = pp[u] - step_size * (mm[u] / denom);
      rr[u] = round_tf32(pp[u]);
    }
    p[i] = P; m[i] = M; v[i] = V;
    if (pr) pr[i] = R;
  }
}

int adam_step(float* p, const float* g, float* m, float* v, float* p_mma, long long n, float lr,
              float beta1, float beta2, float eps, float wd, long long step, float gscale,
              cudaStream_t st) {
  if (n & 3) return VAR_ERR_ARG;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  LaunchScope sc(T_ADAM, 0, st);
  adam_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(
      reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g), reinterpret_cast<float4*>(m),
      reinterpret_cast<float4*>(v), reinterpret_cast<float4*>(p_mma), n / 4, beta1, beta2, eps, wd,
      step_size, bc2_sqrt, gscale);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// ---------------------------------------------------------------------------
// NHWC [B, HW, C] -> reference NCHW-flatten order [B, C*HW] (image_feat_raw of
// models/pretext/pretext_base.py:40); tiny, one thread per element.
// ---------------------------------------------------------------------------
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int B,
                                    int HW, int C) {
  const long long total = (long long)B * HW * C;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int hw = (int)(idx % HW);
    long long t = idx / HW;
    const int c = (int)(t % C);
    const long long b = t / C;
    y[idx] = x[(b * HW + hw) * C + c];
  }
}
int nhwc_to_nchw(const float* x, float* y, int B, int HW, int C, cudaStream_t st) {
  LaunchScope sc(T_MISC, 0, st);
  nhwc_to_nchw_kernel<<<grid_for((long long)B * HW * C, 256), 256, 0, st>>>(x, y, B, HW, C);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// out[b, :] = cat(a[b, :H], c[b, :H])  (GRU final hidden states of both directions)
__global__ void concat2_kernel(const float* __restrict__ a, const float* __restrict__ c,
                               float* __restrict__ out, float* __restrict__ out_r, int B, int Hd) {
  const long long total = (long long)B * 2 * Hd;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % (2 * Hd));
    const long long b = idx / (2 * Hd);
    const float v = j < Hd ? a[b * Hd + j] : c[b * Hd + j - Hd];
    out[idx] = v;
    if (out_r) out_r[idx] = round_tf32(v);
  }
}
int concat2(const float* a, const float* c, float* out, float* out_r, int B, int Hd,
            cudaStream_t st) {
  LaunchScope sc(T_MISC, 0, st);
  concat2_kernel<<<grid_for((long long)B * 2 * Hd, 256), 256, 0, st>>>(a, c, out, out_r, B, Hd);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// a[b, :] = x[b, :H] ; c[b, :] = x[b, H:]
__global__ void split2_kernel(const float* __restrict__ x, float* __restrict__ a,
                              float* __restrict__ c, int B, int Hd) {
  const long long total = (long long)B * 2 * Hd;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % (2 * Hd));
    const long long b = idx / (2 * Hd);
    if (j < Hd) a[b * Hd + j] = x[idx];
    else c[b * Hd + j - Hd] = x[idx];
  }
}
int split2(const float* x, float* a, float* c, int B, int Hd, cudaStream_t st) {
  LaunchScope sc(T_MISC, 0, st);
  split2_kernel<<<grid_for((long long)B * 2 * Hd, 256), 256, 0, st>>>(x, a, c, B, Hd);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// ---------------------------------------------------------------------------
// Reward post-processing of VecPretextNormalize.step_wait (vec_pretext_normalize.py:53-59) on
// the device, float64 like the reference: origStepReward copy, discounted return, RunningMeanStd
// parallel-variance update (running_mean_std.py:16-35), rews / sqrt(var + eps) clipped, and
// ret[done] = 0.  One CTA (N = number of envs, <= a few thousand); state = {mean, var, count}.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(256)
reward_norm_kernel(const float* __restrict__ rew, const unsigned char* __restrict__ done, int N,
                   double* __restrict__ ret, double* __restrict__ rms, double gamma, double eps,
                   double cliprew, int update_rms, float* __restrict__ orig, float* __restrict__ out) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const double r = ret[i] * gamma + (double)rew[i];
    ret[i] = r;
    s += r;
    if (orig) orig[i] = rew[i];
  }
  double var = rms[1];
  if (update_rms) {
    const double bmean = block_sum(s, red) / N;
    double q = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const double d = ret[i] - bmean;
      q += d * d;
    }
    const double bvar = block_sum(q, red) / N;
    const double mean = rms[0], cnt = rms[2];
    const double delta = bmean - mean, tot = cnt + N;
    const double m2 = var * cnt + bvar * N + delta * delta * cnt * N / tot;
    var = m2 / tot;
    __syncthreads();
    if (threadIdx.x == 0) { rms[0] = mean + delta * N / tot; rms[1] = var; rms[2] = tot; }
  }
  const double inv = 1.0 / sqrt(var + eps);
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    double v = update_rms ? (double)rew[i] * inv : (double)rew[i];
    if (update_rms) v = fmin(fmax(v, -cliprew), cliprew);
    out[i] = (float)v;
    if (done[i]) ret[i] = 0.0;
  }
}

int reward_normalize(const float* rew, const unsigned char* done, int N, double* ret, double* rms,
                     double gamma, double eps, double cliprew, int update_rms, float* orig, float* out,
                     cudaStream_t st) {
  if (N <= 0) return VAR_OK;
  LaunchScope sc(T_MISC, 0, st);
  reward_norm_kernel<<<1, 256, 0, st>>>(rew, done, N, ret, rms, gamma, eps, cliprew, update_rms, orig, out);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var
