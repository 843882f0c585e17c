// Standalone GPU self test of the tcgen05 engine against naive CUDA-core
// reference kernels.  Integer-valued operands make every result exact, so any
// layout / descriptor mistake shows up as a gross mismatch.
//   ./selftest kmajor          forward-style GEMMs (K-major operands, TMA weights)
//   ./selftest mn <hyp>        dgrad + wgrad (MN-major operands) under descriptor hypothesis
//   ./selftest perf            rough timing of a few conv shapes
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine_host.cuh"

using namespace var;
extern "C" const char* var_last_error(void);

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e_ = (x);                                                      \
    if (e_ != cudaSuccess) {                                                   \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                 \
    }                                                                          \
  } while (0)

struct Src {
  int kind;       // SrcKind
  SrcLayout sl;
};

__device__ __forceinline__ float load_src(const void* x, int kind, long long idx, float scale) {
  if (kind == SRC_STRIDED_U8) return (float)reinterpret_cast<const uint8_t*>(x)[idx] * scale;
  return reinterpret_cast<const float*>(x)[idx] * scale;
}

__global__ void ref_fwd(ConvShape cs, const void* x, int kind, SrcLayout sl, const float* w,
                        int kpad, const float* bias, float* y, int relu) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)cs.N * cs.P * cs.Q * cs.Cout;
  if (idx >= total) return;
  int o = idx % cs.Cout;
  long long m = idx / cs.Cout;
  int q = m % cs.Q, p = (m / cs.Q) % cs.P, n = m / ((long long)cs.P * cs.Q);
  float acc = bias ? bias[o] : 0.f;
  for (int r = 0; r < cs.R; ++r)
    for (int s = 0; s < cs.S; ++s) {
      int h = p * cs.sh - cs.ph + r, ww = q * cs.sw - cs.pw + s;
      if (h < 0 || h >= cs.H || ww < 0 || ww >= cs.W) continue;
      for (int c = 0; c < cs.Cin; ++c)
        acc += load_src(x, kind, n * sl.sN + h * sl.sH + ww * sl.sW + c * sl.sC, sl.scale) *
               w[(long long)o * kpad + (r * cs.S + s) * cs.Cin + c];
    }
  if (relu) acc = fmaxf(acc, 0.f);
  y[idx] = acc;
}

__global__ void ref_dgrad(ConvShape cs, const float* dy, const float* w, int kpad, float* dx,
                          const float* mask) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  long long total = (long long)cs.N * cs.H * cs.W * cs.Cin;
  if (idx >= total) return;
  int c = idx % cs.Cin;
  long long m = idx / cs.Cin;
  int ww = m % cs.W, h = (m / cs.W) % cs.H, n = m / ((long long)cs.H * cs.W);
  float acc = 0.f;
  for (int r = 0; r < cs.R; ++r)
    for (int s = 0; s < cs.S; ++s) {
      int th = h + cs.ph - r, tw = ww + cs.pw - s;
      if (th < 0 || tw < 0 || th % cs.sh || tw % cs.sw) continue;
      int p = th / cs.sh, q = tw / cs.sw;
      if (p >= cs.P || q >= cs.Q) continue;
      for (int o = 0; o < cs.Cout; ++o)
        acc += dy[(((long long)n * cs.P + p) * cs.Q + q) * cs.Cout + o] *
               w[(long long)o * kpad + (r * cs.S + s) * cs.Cin + c];
    }
  if (mask && !(mask[idx] > 0.f)) acc = 0.f;
  dx[idx] = acc;
}

__global__ void ref_wgrad(ConvShape cs, const void* x, int kind, SrcLayout sl, const float* dy,
                          int kpad, float* dw) {
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int K = cs.R * cs.S * cs.Cin;
  long long total = (long long)cs.Cout * K;
  if (idx >= total) return;
  int k = idx % K, o = idx / K;
  int c = k % cs.Cin, rs = k / cs.Cin, s = rs % cs.S, r = rs / cs.S;
  float acc = 0.f;
  for (int n = 0; n < cs.N; ++n)
    for (int p = 0; p < cs.P; ++p)
      for (int q = 0; q < cs.Q; ++q) {
        int h = p * cs.sh - cs.ph + r, ww = q * cs.sw - cs.pw + s;
        if (h < 0 || h >= cs.H || ww < 0 || ww >= cs.W) continue;
        acc += load_src(x, kind, n * sl.sN + h * sl.sH + ww * sl.sW + c * sl.sC, sl.scale) *
               dy[(((long long)n * cs.P + p) * cs.Q + q) * cs.Cout + o];
      }
  dw[(long long)o * kpad + k] = acc;
}

static uint32_t rng_state = 12345;
static inline uint32_t rnd() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return rng_state >> 8;
}
static std::vector<float> rand_ints(size_t n, int lo, int hi) {
  std::vector<float> v(n);
  for (auto& x : v) x = (float)(lo + (int)(rnd() % (uint32_t)(hi - lo + 1)));
  return v;
}
template <class T>
static T* to_dev(const std::vector<T>& v) {
  T* d;
  CK(cudaMalloc(&d, v.size() * sizeof(T) + 16));
  CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
static double max_abs_diff(const float* da, const float* db, size_t n, double* maxref) {
  std::vector<float> a(n), b(n);
  CK(cudaMemcpy(a.data(), da, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), db, n * 4, cudaMemcpyDeviceToHost));
  double m = 0, r = 0;
  for (size_t i = 0; i < n; ++i) {
    double d = fabs((double)a[i] - (double)b[i]);
    if (!(d <= m)) m = d;  // catches NaN
    if (fabs(b[i]) > r) r = fabs(b[i]);
  }
  *maxref = r;
  return m;
}

static ConvShape mk(int N, int H, int W, int Cin, int Cout, int R, int S, int sh, int sw, int ph,
                    int pw) {
  ConvShape c{N, H, W, Cin, Cout, R, S, sh, sw, ph, pw, 0, 0};
  c.P = (H + 2 * ph - R) / sh + 1;
  c.Q = (W + 2 * pw - S) / sw + 1;
  return c;
}

struct Case {
  const char* name;
  ConvShape cs;
  int kind;   // SrcKind
  int nchw;   // strided layout is NCHW (else NHWC)
};

static int nfail = 0;
static void report(const char* what, const char* name, double err, double ref) {
  bool ok = (err == 0.0) && ref > 0;
  printf("  [%s] %-28s %-8s max|diff|=%g max|ref|=%g\n", ok ? "PASS" : "FAIL", name, what, err, ref);
  if (!ok) ++nfail;
  fflush(stdout);
}

static void run_case(const Case& c, bool do_fwd, bool do_dgrad, bool do_wgrad) {
  const ConvShape& cs = c.cs;
  const int K = cs.R * cs.S * cs.Cin, kpad = round_up32(K);
  const size_t nx = (size_t)cs.N * cs.H * cs.W * cs.Cin;
  const size_t ny = (size_t)cs.N * cs.P * cs.Q * cs.Cout;
  SrcLayout sl;
  if (c.kind == SRC_NHWC_F32 || !c.nchw) {
    sl.sC = 1; sl.sW = cs.Cin; sl.sH = (long long)cs.W * cs.Cin; sl.sN = (long long)cs.H * cs.W * cs.Cin;
  } else {
    sl.sW = 1; sl.sH = cs.W; sl.sC = (long long)cs.H * cs.W; sl.sN = (long long)cs.Cin * cs.H * cs.W;
  }
  sl.scale = 1.f;
  void* dx_in;
  if (c.kind == SRC_STRIDED_U8) {
    std::vector<uint8_t> xb(nx);
    for (auto& v : xb) v = (uint8_t)(rnd() % 256);
    dx_in = to_dev(xb);
  } else {
    dx_in = to_dev(rand_ints(nx, -2, 2));
  }
  std::vector<float> wv((size_t)cs.Cout * kpad, 0.f);
  for (int o = 0; o < cs.Cout; ++o)
    for (int k = 0; k < K; ++k) wv[(size_t)o * kpad + k] = (float)((int)(rnd() % 5) - 2);
  float* dw = to_dev(wv);
  float* dbias = to_dev(rand_ints(cs.Cout, -3, 3));
  float* dy_ref; float* dy_out;
  CK(cudaMalloc(&dy_ref, ny * 4)); CK(cudaMalloc(&dy_out, ny * 4));
  CK(cudaMemset(dy_out, 0xFF, ny * 4));
  double ref;
  if (do_fwd) {
    ref_fwd<<<(unsigned)((ny + 255) / 256), 256>>>(cs, dx_in, c.kind, sl, dw, kpad, dbias, dy_ref, 1);
    int rc = conv_fwd(cs, dx_in, c.kind, &sl, dw, dbias, dy_out, 1, 0, 0);
    if (rc) { printf("  conv_fwd rc=%d %s\n", rc, var_last_error()); ++nfail; }
    CK(cudaDeviceSynchronize());
    double e = max_abs_diff(dy_out, dy_ref, ny, &ref);
    report("fwd", c.name, e, ref);
  }
  float* dgy = to_dev(rand_ints(ny, -2, 2));
  if (do_dgrad && c.kind == SRC_NHWC_F32) {
    float *gx_ref, *gx_out;
    CK(cudaMalloc(&gx_ref, nx * 4)); CK(cudaMalloc(&gx_out, nx * 4));
    CK(cudaMemset(gx_out, 0xFF, nx * 4));
    const float* mask = reinterpret_cast<const float*>(dx_in);
    ref_dgrad<<<(unsigned)((nx + 255) / 256), 256>>>(cs, dgy, dw, kpad, gx_ref, mask);
    int rc = conv_dgrad(cs, dgy, dw, gx_out, mask, nullptr, 0, 0);
    if (rc) { printf("  conv_dgrad rc=%d %s\n", rc, var_last_error()); ++nfail; }
    CK(cudaDeviceSynchronize());
    double e = max_abs_diff(gx_out, gx_ref, nx, &ref);
    report("dgrad", c.name, e, ref);
    cudaFree(gx_ref); cudaFree(gx_out);
  }
  if (do_wgrad) {
    float *gw_ref, *gw_out, *gb;
    const size_t nw = (size_t)cs.Cout * kpad;
    CK(cudaMalloc(&gw_ref, nw * 4)); CK(cudaMalloc(&gw_out, nw * 4)); CK(cudaMalloc(&gb, cs.Cout * 4));
    CK(cudaMemset(gw_ref, 0, nw * 4)); CK(cudaMemset(gw_out, 0, nw * 4)); CK(cudaMemset(gb, 0, cs.Cout * 4));
    ref_wgrad<<<(unsigned)(((size_t)cs.Cout * K + 255) / 256), 256>>>(cs, dx_in, c.kind, sl, dgy, kpad, gw_ref);
    int rc = conv_wgrad(cs, dx_in, c.kind, &sl, dgy, gw_out, gb, 0);
    if (rc) { printf("  conv_wgrad rc=%d %s\n", rc, var_last_error()); ++nfail; }
    CK(cudaDeviceSynchronize());
    double e = max_abs_diff(gw_out, gw_ref, nw, &ref);
    report("wgrad", c.name, e, ref);
    // bias grad check on host
    std::vector<float> hy(ny), hb(cs.Cout);
    CK(cudaMemcpy(hy.data(), dgy, ny * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), gb, cs.Cout * 4, cudaMemcpyDeviceToHost));
    double eb = 0;
    for (int o = 0; o < cs.Cout; ++o) {
      double s = 0;
      for (size_t m = 0; m < ny / cs.Cout; ++m) s += hy[m * cs.Cout + o];
      eb = fmax(eb, fabs(s - hb[o]));
    }
    report("dbias", c.name, eb, 1.0);
    cudaFree(gw_ref); cudaFree(gw_out); cudaFree(gb);
  }
  cudaFree(dx_in); cudaFree(dw); cudaFree(dbias); cudaFree(dy_ref); cudaFree(dy_out); cudaFree(dgy);
}

static std::vector<Case> cases() {
  std::vector<Case> v;
  v.push_back({"linear 300x64->32", mk(300, 1, 1, 64, 32, 1, 1, 1, 1, 0, 0), SRC_NHWC_F32, 0});
  v.push_back({"linear 1000x1152->128", mk(1000, 1, 1, 1152, 128, 1, 1, 1, 1, 0, 0), SRC_NHWC_F32, 0});
  v.push_back({"linear 256x448->1536", mk(256, 1, 1, 448, 1536, 1, 1, 1, 1, 0, 0), SRC_NHWC_F32, 0});
  v.push_back({"conv3x3 s1 p1 32->64", mk(2, 12, 12, 32, 64, 3, 3, 1, 1, 1, 1), SRC_NHWC_F32, 0});
  v.push_back({"conv3x3 s2 p1 64->64", mk(3, 24, 24, 64, 64, 3, 3, 2, 2, 1, 1), SRC_NHWC_F32, 0});
  v.push_back({"conv3x3 s1 p1 64->128", mk(2, 12, 12, 64, 128, 3, 3, 1, 1, 1, 1), SRC_NHWC_F32, 0});
  v.push_back({"conv11x5 s2 p5 64->64", mk(1, 30, 20, 64, 64, 11, 5, 2, 2, 5, 5), SRC_NHWC_F32, 0});
  v.push_back({"conv3x1 s(2,1) 32->32", mk(4, 48, 1, 32, 32, 3, 1, 2, 1, 0, 0), SRC_NHWC_F32, 0});
  v.push_back({"linear 4x576->128 (tiny M)", mk(4, 1, 1, 576, 128, 1, 1, 1, 1, 0, 0), SRC_NHWC_F32, 0});
  v.push_back({"conv7x3 s2 p1 odd 64->64", mk(2, 31, 13, 64, 64, 7, 3, 2, 2, 1, 1), SRC_NHWC_F32, 0});
  v.push_back({"conv3x3 s2 p1 128->128 6x6", mk(5, 6, 6, 128, 128, 3, 3, 2, 2, 1, 1), SRC_NHWC_F32, 0});
  v.push_back({"conv3x3 s1 p1 32->32 20x20", mk(3, 20, 20, 32, 32, 3, 3, 1, 1, 1, 1), SRC_NHWC_F32, 0});
  v.push_back({"scalar f32 nchw c3 s2", mk(2, 16, 16, 3, 32, 3, 3, 2, 2, 1, 1), SRC_STRIDED_F32, 1});
  v.push_back({"scalar u8 nchw c3 s1", mk(2, 16, 16, 3, 32, 3, 3, 1, 1, 1, 1), SRC_STRIDED_U8, 1});
  v.push_back({"scalar f32 c1 5x40", mk(2, 100, 40, 1, 32, 5, 40, 2, 1, 0, 0), SRC_STRIDED_F32, 1});
  v.push_back({"scalar f32 c1 11x11 s2", mk(1, 60, 40, 1, 64, 11, 11, 2, 2, 5, 5), SRC_STRIDED_F32, 1});
  return v;
}

int main(int argc, char** argv) {
  const char* mode = argc > 1 ? argv[1] : "kmajor";
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s sm_%d%d, %d SMs\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  if (!strcmp(mode, "kmajor")) {
    for (auto& c : cases()) run_case(c, true, false, false);
  } else if (!strcmp(mode, "mn")) {
    int hyp = argc > 2 ? atoi(argv[2]) : 1;
    MnCfg& m = mn_cfg();
    switch (hyp) {
      case 1: m = MnCfg{4096, 512, 1, 1, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B}; break;
      case 2: m = MnCfg{512, 4096, 1, 1, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B}; break;
      case 3: m = MnCfg{4096, 1024, 2, 0, (int)CU_TENSOR_MAP_SWIZZLE_128B}; break;
      case 4: m = MnCfg{1024, 4096, 2, 0, (int)CU_TENSOR_MAP_SWIZZLE_128B}; break;
      case 5: m = MnCfg{4096, 1024, 1, 1, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B}; break;
      default: break;
    }
    printf("MN hypothesis %d: lbo=%d sbo=%d type=%d swz32=%d tma_swz=%d\n", hyp, m.lbo, m.sbo, m.type,
           m.swz32, m.tma_swizzle);
    for (auto& c : cases()) run_case(c, false, true, true);
  } else if (!strcmp(mode, "perf")) {
    struct { const char* name; ConvShape cs; } shapes[] = {
        {"thor.img.conv2 B=64", mk(64, 96, 96, 32, 32, 3, 3, 1, 1, 1, 1)},
        {"thor.img.conv3 B=64", mk(64, 48, 48, 32, 64, 3, 3, 1, 1, 1, 1)},
        {"thor.snd.conv2 B=64", mk(64, 300, 20, 64, 64, 11, 5, 2, 2, 5, 5)},
        {"kuka.img.conv2 B=1024", mk(1024, 48, 48, 32, 32, 3, 3, 2, 2, 1, 1)},
    };
    for (auto& s : shapes) {
      const ConvShape& cs = s.cs;
      const int K = cs.R * cs.S * cs.Cin, kpad = round_up32(K);
      size_t nx = (size_t)cs.N * cs.H * cs.W * cs.Cin, ny = (size_t)cs.N * cs.P * cs.Q * cs.Cout;
      float *x, *w, *y, *b, *gx, *gw;
      CK(cudaMalloc(&x, nx * 4)); CK(cudaMalloc(&w, (size_t)cs.Cout * kpad * 4));
      CK(cudaMalloc(&y, ny * 4)); CK(cudaMalloc(&b, cs.Cout * 4));
      CK(cudaMalloc(&gx, nx * 4)); CK(cudaMalloc(&gw, (size_t)cs.Cout * kpad * 4));
      CK(cudaMemset(x, 0, nx * 4)); CK(cudaMemset(w, 0, (size_t)cs.Cout * kpad * 4));
      CK(cudaMemset(b, 0, cs.Cout * 4)); CK(cudaMemset(y, 0, ny * 4));
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const double flop = 2.0 * cs.N * cs.P * cs.Q * (double)cs.Cout * K;
      for (int pass = 0; pass < 3; ++pass) {
        const char* nm = pass == 0 ? "fwd" : pass == 1 ? "dgrad" : "wgrad";
        for (int i = 0; i < 3; ++i) {
          if (pass == 0) conv_fwd(cs, x, SRC_NHWC_F32, nullptr, w, b, y, 1, 1, 0);
          else if (pass == 1) conv_dgrad(cs, y, w, gx, x, nullptr, 1, 0);
          else conv_wgrad(cs, x, SRC_NHWC_F32, nullptr, y, gw, nullptr, 0);
        }
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int iters = 10;
        for (int i = 0; i < iters; ++i) {
          if (pass == 0) conv_fwd(cs, x, SRC_NHWC_F32, nullptr, w, b, y, 1, 1, 0);
          else if (pass == 1) conv_dgrad(cs, y, w, gx, x, nullptr, 1, 0);
          else conv_wgrad(cs, x, SRC_NHWC_F32, nullptr, y, gw, nullptr, 0);
        }
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= iters;
        printf("  %-24s %-6s %8.3f ms  %7.1f TFLOP/s  act %.1f GB/s\n", s.name, nm, ms,
               flop / ms * 1e-9, (nx + ny) * 4.0 / ms * 1e-6);
      }
      cudaFree(x); cudaFree(w); cudaFree(y); cudaFree(b); cudaFree(gx); cudaFree(gw);
    }
  }
  printf("selftest %s: %s (%d failures)\n", mode, nfail ? "FAILED" : "OK", nfail);
  return nfail ? 1 : 0;
}
