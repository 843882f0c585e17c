// Host launchers for the tcgen05 engine (see tc_engine.cuh).
#include "engine_host.cuh"
#include "first_conv.cuh"
#include "cin1_conv.cuh"
#include "kuka_sound.cuh"
#include "cin3_conv.cuh"
#include "gemm_persist.cuh"
#include "gru_persist.cuh"
#include "gru_ksplit.cuh"
#include "h16_engine.cuh"
#include "halo_conv.cuh"

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

static thread_local char g_last_error[512] = "";
void var_set_last_error(const char* msg, const char* file, int line) {
  snprintf(g_last_error, sizeof(g_last_error), "%s (%s:%d)", msg, file, line);
}
extern "C" const char* var_last_error(void) { return g_last_error; }

cudaError_t var_ensure_dyn_smem(const void* kernel, size_t bytes) {
  struct Key { const void* f; int dev; bool operator==(const Key& o) const { return f == o.f && dev == o.dev; } };
  struct Hash { size_t operator()(const Key& k) const { return std::hash<const void*>()(k.f) * 31u + (size_t)k.dev; } };
  static std::unordered_map<Key, size_t, Hash> done;
  static std::mutex mu;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = done[Key{kernel, dev}];
  if (bytes <= cur) return cudaSuccess;
  if (bytes > 48 * 1024 || cur > 0) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
  }
  cur = bytes;
  return cudaSuccess;
}

namespace var {

// ---------------------------------------------------------------------------
// launch accounting / per-kernel event profiler
// ---------------------------------------------------------------------------
static thread_local char g_prof_note[48] = "";
void prof_note(const char* fmt, int a, int b, int c, int d, int e) {
  snprintf(g_prof_note, sizeof(g_prof_note), fmt, a, b, c, d, e);
}
namespace {
struct ProfRec { int tag; double flops; cudaEvent_t e0, e1; char note[48]; };
struct Prof {
  std::mutex mu;
  bool on = false;
  std::vector<ProfRec> recs;
  std::vector<cudaEvent_t> pool;
  long long launches = 0;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
};
Prof& prof() { static Prof p; return p; }
}  // namespace

LaunchScope::LaunchScope(int tag, double flops, cudaStream_t s) : idx(-1), st(s) {
  Prof& p = prof();
  std::lock_guard<std::mutex> lk(p.mu);
  ++p.launches;
  if (!p.on) return;
  ProfRec r{tag, flops, p.get(), p.get(), {0}};
  snprintf(r.note, sizeof(r.note), "%s", g_prof_note);
  cudaEventRecord(r.e0, st);
  idx = (int)p.recs.size();
  p.recs.push_back(r);
}
LaunchScope::~LaunchScope() {
  if (idx < 0) return;
  Prof& p = prof();
  std::lock_guard<std::mutex> lk(p.mu);
  cudaEventRecord(p.recs[idx].e1, st);
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

MnCfg& mn_cfg() {
  static MnCfg c;
  return c;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int rows, cols, box_rows, swz;
  long long pitch;
  int esize;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && box_rows == o.box_rows &&
           swz == o.swz && pitch == o.pitch && esize == o.esize;
  }
};
struct TmapHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1315423911u + (size_t)k.rows;
    h = h * 1315423911u + (size_t)k.cols;
    h = h * 1315423911u + (size_t)k.box_rows;
    h = h * 1315423911u + (size_t)k.swz;
    h = h * 1315423911u + (size_t)k.pitch;
    h = h * 1315423911u + (size_t)k.esize;
    return h;
  }
};

int get_tmap_2d(const float* ptr, int rows, int cols, long long pitch, int box_rows, int swizzle,
                CUtensorMap* out) {
  return get_tmap_2d_e(ptr, 4, rows, cols, pitch, box_rows, swizzle, out);
}

// esize = 4 (fp32 / tf32) or 2 (f16 / bf16): the box is always 128 bytes wide (32 or 64 elements).
int get_tmap_2d_e(const void* ptr, int esize, int rows, int cols, long long pitch, int box_rows, int swizzle,
                  CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapHash> cache;
  static std::mutex mu;
  TmapKey key{ptr, rows, cols, box_rows, swizzle, pitch, esize};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return VAR_OK;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    var_set_last_error("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((pitch * esize) & 15) || (esize != 4 && esize != 2)) {
    var_set_last_error("tensor map needs 16-byte aligned base and pitch", __FILE__, __LINE__);
    return VAR_ERR_ARG;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch * (cuuint64_t)esize};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esize), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUtensorMap m;
  // 16-bit payloads are moved as opaque 16-bit words (f16 and bf16 alike)
  CUresult r = fn(&m, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                  const_cast<void*>(ptr), gdim, gstr, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[128];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed: %d (rows=%d cols=%d box_rows=%d)",
             (int)r, rows, cols, box_rows);
    var_set_last_error(buf, __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = m;
  }
  *out = m;
  return VAR_OK;
}

int get_tmap_3d(const float* ptr, int d0, int d1, int d2, int b0, int b1, CUtensorMap* out) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    var_set_last_error("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((d0 * 4) & 15) || ((b0 * 4) & 15)) return VAR_ERR_ARG;
  cuuint64_t gdim[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t gstr[2] = {(cuuint64_t)d0 * 4, (cuuint64_t)d0 * d1 * 4};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled(3d) failed: %d (dims %d %d %d box %d %d)", (int)r, d0, d1,
             d2, b0, b1);
    var_set_last_error(buf, __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  return VAR_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeIm2colFn encode_im2col_fn() {
  static EncodeIm2colFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeIm2colFn>(p);
  });
  return fn;
}

// im2col-mode map over an NHWC fp32 tensor [N, H, W, C]: base pixels walk the box
// [low, extent-1+up] with `stride`; each load covers `pixels` base pixels x 32 channels.
int get_tmap_im2col(const float* ptr, int N, int H, int W, int C, int low_w, int low_h, int up_w,
                    int up_h, int stride_w, int stride_h, int pixels, int swizzle, CUtensorMap* out) {
  return get_tmap_im2col_e(ptr, 4, N, H, W, C, low_w, low_h, up_w, up_h, stride_w, stride_h, pixels, swizzle, out);
}

int get_tmap_im2col_e(const void* ptr, int esize, int N, int H, int W, int C, int low_w, int low_h, int up_w,
                      int up_h, int stride_w, int stride_h, int pixels, int swizzle, CUtensorMap* out) {
  EncodeIm2colFn fn = encode_im2col_fn();
  if (!fn) {
    var_set_last_error("cuTensorMapEncodeIm2col entry point unavailable", __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((C * esize) & 15) || (esize != 4 && esize != 2)) return VAR_ERR_ARG;
  const cuuint64_t es = (cuuint64_t)esize;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
  int lower[2] = {low_w, low_h};
  int upper[2] = {up_w, up_h};
  cuuint32_t estr[4] = {1u, (cuuint32_t)stride_w, (cuuint32_t)stride_h, 1u};
  CUresult r = fn(out, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 4,
                  const_cast<void*>(ptr), gdim, gstr, lower,
                  upper, (cuuint32_t)(128 / esize), (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[200];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeIm2col failed: %d (N=%d H=%d W=%d C=%d low=%d,%d up=%d,%d stride=%d,%d)",
             (int)r, N, H, W, C, low_w, low_h, up_w, up_h, stride_w, stride_h);
    var_set_last_error(buf, __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  return VAR_OK;
}

// VAR_PDL=1 launches the GRU steps with programmatic dependent launch.  Measured on B200: no gain
// (16.8 vs 16.4 ms/step) -- a step is bound by its TMA round trips and epilogue, not by launch or
// prologue latency -- so ordinary stream-ordered launches stay the default.
static bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VAR_PDL"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}

int gather_mode() {  // VAR_GATHER=cp_async keeps the LSU gather kernels (A/B testing)
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("VAR_GATHER");
    mode = (e && !strcmp(e, "cp_async")) ? 0 : 1;
  }
  return mode;
}

// ---------------------------------------------------------------------------
template <int GMODE, int EPI>
static int launch_gemm_t(const CUtensorMap& t0, const CUtensorMap& t1, const CUtensorMap& a0,
                         const CUtensorMap& a1, const GemmParams& p, dim3 grid, cudaStream_t st) {
  size_t smem = gemm_smem_bytes(p.bn, p.stages);
  if (GMODE == G_SCALAR_F32 || GMODE == G_SCALAR_U8)
    smem += (size_t)p.g[0].C * p.sc_nh * p.sc_wpad * 4 + 16;  // staged input patch
  VAR_ENSURE_SMEM((tc_gemm_kernel<GMODE, EPI>), smem);
  {
    double flops = 0;
    for (unsigned z = 0; z < grid.z; ++z)
      flops += 2.0 * p.g[z].M * (double)(EPI != EPI_STD ? p.bn * grid.y : p.e[z].ncols) * p.g[z].K;
    int tag = T_GEMM_SCALAR;
    if (EPI != EPI_STD) tag = T_GRU_STEP;
    else if (p.b_mn_major) tag = T_GEMM_DGRAD;
    else if (GMODE == G_VEC_FWD || GMODE == G_TMA_IM2COL || GMODE == G_TMA_TILED) tag = T_GEMM_FWD;
    LaunchScope sc(tag, flops, st);
    if (EPI != EPI_STD && pdl_enabled()) {
      // GRU time steps: 145 small dependent launches per training step.  Programmatic dependent
      // launch lets step s+1 set up (barriers, TMEM, descriptors) while step s is still running.
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = grid; cfg.blockDim = dim3(160, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      VAR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, tc_gemm_kernel<GMODE, EPI>, t0, t1, a0, a1, p));
    } else {
      tc_gemm_kernel<GMODE, EPI><<<grid, 160, smem, st>>>(t0, t1, a0, a1, p);
    }
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

static bool gemm_persist_enabled() {  // VAR_GEMM_PERSIST=0 keeps one CTA per tile
  static int on = -1;
  if (on < 0) { const char* e = getenv("VAR_GEMM_PERSIST"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

template <int GMODE, bool H16 = false>
static int launch_gemm_persist_t(const CUtensorMap& tb, const CUtensorMap& ta, const GemmParams& p_in, int m_tiles,
                                 int n_tiles, cudaStream_t st) {
  GemmParams p = p_in;
  p.kps = env_int("VAR_GEMM_KPS", 2);
  if (p.kps < 1 || p.stages % p.kps) p.kps = 1;
  p.stages /= p.kps;
  size_t smem = gemm_smem_bytes(p.bn, p.stages * p.kps);
  // Wide short-K tiles (one CTA per SM) are epilogue bound: after tcgen05.ld a thread owns a row, so
  // direct stores touch 32 different 128-byte lines per instruction (~32 * bn cycles per tile).  When
  // that is at least half of the tile's operand stream (~50 B/clk) the accumulator is transposed
  // through a swizzled smem tile instead.  (Measured: GRU input projection 198 -> 272 TF/s; for the
  // two-CTA-per-SM configurations -- masked dgrads of the small convs -- it costs 10-25 %.)
  static int coal = -1;
  if (coal < 0) { const char* e = getenv("VAR_EPI_COALESCE"); coal = (e && e[0] == '0') ? 0 : 1; }
  const long long stream_clk = (long long)p.num_kb * (kTileABytes + p.bn * 128) / 50;
  const EpiParams& e0 = p.e[0];
  const bool plain_f32_epi = !H16 || (e0.out_kind == 0 && e0.mask_kind == 0 && e0.out_scale == nullptr);
  CUtensorMap tc = tb;
  static int epi_tma = -1;
  if (epi_tma < 0) epi_tma = env_int("VAR_EPI_TMA", 1);
  if (plain_f32_epi && coal && epi_tma && smem * 2 + 4096 > 227 * 1024 && 64LL * p.bn > stream_clk && !e0.mask && !e0.addsrc &&
      !e0.map.on && p.bn % 32 == 0 && e0.ncols % 32 == 0 && (e0.ldo % 4) == 0) {
    // epilogue-bound wide tiles with a plain fp32 output: TMA-store epilogue (two 16 KB staging boxes); a shallower
    // operand ring makes room for them (the main loop of these tiles is far from being the limiter)
    p.kps = 1; p.stages = 3;
    smem = gemm_smem_bytes(p.bn, p.stages) + 1024 + 2 * 16384;
    if (smem <= 227 * 1024 &&
        get_tmap_2d(e0.out, p.g[0].M, e0.ncols, e0.ldo, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tc) == VAR_OK)
      p.epi_tma = 1;
    else { p = p_in; p.kps = env_int("VAR_GEMM_KPS", 2); if (p.kps < 1 || p.stages % p.kps) p.kps = 1; p.stages /= p.kps;
           smem = gemm_smem_bytes(p.bn, p.stages * p.kps); }
  }
  if (!p.epi_tma && plain_f32_epi && coal && smem * 2 + 4096 > 227 * 1024 && smem + 4 * 4352 <= 227 * 1024 && 64LL * p.bn > stream_clk) {
    p.epi_coalesce = 1;
    smem += 4 * 4352;
  }
  VAR_ENSURE_SMEM((tc_gemm_persist_kernel<GMODE, H16>), smem);
  const long long total = (long long)m_tiles * n_tiles;
  const int per_sm = smem * 2 + 4096 <= 227 * 1024 ? 2 : 1;
  int grid = kNumSMs * per_sm;
  if (grid > total) grid = (int)total;
  const double flops = 2.0 * p.g[0].M * (double)p.e[0].ncols * p.g[0].K;
  const bool is_dgrad = p.b_mn_major || p.prof_dgrad;
  LaunchScope sc(H16 ? (is_dgrad ? T_GEMM_DGRAD16 : T_GEMM_FWD16) : (is_dgrad ? T_GEMM_DGRAD : T_GEMM_FWD), flops, st);
  tc_gemm_persist_kernel<GMODE, H16><<<grid, 192, smem, st>>>(tb, ta, tc, p, m_tiles, n_tiles);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

int launch_gemm(int gmode, int epi, const CUtensorMap& t0, const CUtensorMap& t1,
                const CUtensorMap& a0, const CUtensorMap& a1, const GemmParams& p, dim3 grid,
                cudaStream_t st) {
  if (epi == EPI_STD && grid.z == 1 && gemm_persist_enabled() && (gmode == G_TMA_IM2COL || gmode == G_TMA_TILED) &&
      (long long)grid.x * grid.y > 2 * kNumSMs) {
    // more tiles than CTA slots: resident CTAs walking the tile list beat one CTA per tile
    if (gmode == G_TMA_IM2COL) return launch_gemm_persist_t<G_TMA_IM2COL>(t0, a0, p, (int)grid.x, (int)grid.y, st);
    return launch_gemm_persist_t<G_TMA_TILED>(t0, a0, p, (int)grid.x, (int)grid.y, st);
  }
  if (epi == EPI_GRU_FWD) {
    if (gmode == G_VEC_FWD) return launch_gemm_t<G_VEC_FWD, EPI_GRU_FWD>(t0, t1, a0, a1, p, grid, st);
    if (gmode == G_TMA_TILED) return launch_gemm_t<G_TMA_TILED, EPI_GRU_FWD>(t0, t1, a0, a1, p, grid, st);
    return VAR_ERR_UNSUPPORTED;
  }
  if (epi == EPI_GRU_BWD) {
    if (gmode == G_TMA_TILED) return launch_gemm_t<G_TMA_TILED, EPI_GRU_BWD>(t0, t1, a0, a1, p, grid, st);
    return VAR_ERR_UNSUPPORTED;
  }
  switch (gmode) {
    case G_VEC_FWD: return launch_gemm_t<G_VEC_FWD, EPI_STD>(t0, t1, a0, a1, p, grid, st);
    case G_VEC_DGRAD: return launch_gemm_t<G_VEC_DGRAD, EPI_STD>(t0, t1, a0, a1, p, grid, st);
    case G_SCALAR_F32: return launch_gemm_t<G_SCALAR_F32, EPI_STD>(t0, t1, a0, a1, p, grid, st);
    case G_SCALAR_U8: return launch_gemm_t<G_SCALAR_U8, EPI_STD>(t0, t1, a0, a1, p, grid, st);
    case G_TMA_IM2COL: return launch_gemm_t<G_TMA_IM2COL, EPI_STD>(t0, t1, a0, a1, p, grid, st);
    case G_TMA_TILED: return launch_gemm_t<G_TMA_TILED, EPI_STD>(t0, t1, a0, a1, p, grid, st);
  }
  return VAR_ERR_UNSUPPORTED;
}

template <int GMODE>
static int launch_wgrad_t(const WgradParams& p, dim3 grid, cudaStream_t st) {
  size_t smem = wgrad_smem_bytes(p.cout, p.stages);
  if (GMODE == G_SCALAR_F32 || GMODE == G_SCALAR_U8)
    smem += (size_t)p.g.C * p.sc_nh * p.sc_wpad * 4 + 16;  // staged input patch
  VAR_ENSURE_SMEM(tc_wgrad_kernel<GMODE>, smem);
  {
    LaunchScope sc(T_WGRAD, 2.0 * p.g.M * (double)p.cout * p.g.K, st);
    tc_wgrad_kernel<GMODE><<<grid, 160, smem, st>>>(p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

static void pick_pipeline(int bn, int* stages, int* la) {
  if (bn <= 64) { *stages = 4; *la = 2; }
  else if (bn <= 128) { *stages = 3; *la = 1; }
  else { *stages = 4; *la = 2; }
}

static int pick_bn(int n) {
  if (n <= 256) return n;
  for (int bn = 256; bn >= 32; bn -= 32)
    if (n % bn == 0) return bn;
  return 0;
}
// Narrow the N tile while the grid would leave most SMs idle (small-M GEMMs: Linear heads,
// GRU recurrences).  `mult` keeps the tile a multiple of the operand's box granule.
static int widen_grid_bn(int bn, int n, int m_tiles, int instances, int mult) {
  while (bn % 2 == 0 && (bn / 2) % mult == 0 && n % (bn / 2) == 0 &&
         (long long)m_tiles * (n / bn) * instances < kNumSMs)
    bn /= 2;
  return bn;
}

static int fill_fwd_geom(GatherGeom* g, const ConvShape& cs, const void* x, int src_kind,
                         const SrcLayout* sl) {
  g->src = x;
  g->M = cs.N * cs.P * cs.Q;
  g->P = cs.P; g->Q = cs.Q;
  g->H = cs.H; g->W = cs.W; g->C = cs.Cin;
  g->R = cs.R; g->S = cs.S;
  g->sh = cs.sh; g->sw = cs.sw; g->ph = cs.ph; g->pw = cs.pw;
  g->K = cs.R * cs.S * cs.Cin;
  if (src_kind == SRC_NHWC_F32) {
    if (cs.Cin % 32) return VAR_ERR_UNSUPPORTED;
    g->sC = 1; g->sW = cs.Cin; g->sH = (long long)cs.W * cs.Cin;
    g->sN = (long long)cs.H * cs.W * cs.Cin;
    g->scale = 1.f;
  } else {
    if (!sl || g->K > 256) return VAR_ERR_UNSUPPORTED;
    g->sN = sl->sN; g->sH = sl->sH; g->sW = sl->sW; g->sC = sl->sC;
    g->scale = sl->scale;
  }
  return VAR_OK;
}

// image conv1 of both encoders: 3 -> 32 channels, 3x3, pad 1, stride 1 or 2, strided source
static bool is_first_conv(const ConvShape& cs, int src_kind) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("VAR_FIRST_CONV"); on = (e && e[0] == '0') ? 0 : 1; }
  return on && src_kind != SRC_NHWC_F32 && cs.Cin == 3 && cs.Cout == 32 && cs.R == 3 && cs.S == 3 &&
         cs.ph == 1 && cs.pw == 1 && cs.sh == cs.sw && (cs.sh == 1 || cs.sh == 2);
}

static int gmode_of(int src_kind) {
  return src_kind == SRC_NHWC_F32 ? G_VEC_FWD
                                  : (src_kind == SRC_STRIDED_F32 ? G_SCALAR_F32 : G_SCALAR_U8);
}

static bool conv_halo32_ok(const ConvShape& cs);
static int conv_fwd_halo_tf32(const ConvShape& cs, const float* x, const float* w, const float* bias, float* y, int relu,
                              int round_out, cudaStream_t st);

int conv_fwd(const ConvShape& cs, const void* x, int src_kind, const SrcLayout* sl, const float* w,
             const float* bias, float* y, int relu, int round_out, cudaStream_t st, int out_kind) {
  prof_note("fwd N%d H%d Ci%d Co%d R%d", cs.N, cs.H, cs.Cin, cs.Cout, cs.R);
  if (out_kind != 0 && !(src_kind == SRC_STRIDED_F32 && sl &&
                         cin1_conv_match(cs.H, cs.W, cs.Cin, cs.Cout, cs.R, cs.S, cs.sh, cs.sw, cs.ph, cs.pw, sl->sN,
                                         sl->sH, sl->sW, sl->scale, x)))
    return VAR_ERR_UNSUPPORTED;
  if (src_kind == SRC_STRIDED_U8 && sl &&
      cin3_conv_match(cs.H, cs.W, cs.Cin, cs.Cout, cs.R, cs.S, cs.sh, cs.sw, cs.ph, cs.pw, cs.P, cs.Q, sl->sN, sl->sH,
                      sl->sW, sl->sC, x)) {
    Cin3Args a;
    memset(&a, 0, sizeof(a));
    a.x = reinterpret_cast<const unsigned char*>(x); a.scale = sl->scale;
    a.N = cs.N; a.H = cs.H; a.W = cs.W; a.P = cs.P; a.Q = cs.Q; a.stride = cs.sh;
    a.w = w; a.bias = bias; a.y = y; a.relu = relu; a.round_out = round_out;
    return cin3_conv_fwd(a, st);
  }
  if (is_first_conv(cs, src_kind) && sl) {
    FirstConvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.sN = sl->sN; a.sH = sl->sH; a.sW = sl->sW; a.sC = sl->sC; a.scale = sl->scale;
    a.N = cs.N; a.H = cs.H; a.W = cs.W; a.P = cs.P; a.Q = cs.Q; a.stride = cs.sh;
    a.w = w; a.kpad = round_up32(27); a.bias = bias; a.y = y; a.relu = relu; a.round_out = round_out;
    return first_conv_fwd(a, src_kind == SRC_STRIDED_U8, st);
  }
  if (src_kind == SRC_STRIDED_F32 && sl &&
      cin1_conv_match(cs.H, cs.W, cs.Cin, cs.Cout, cs.R, cs.S, cs.sh, cs.sw, cs.ph, cs.pw, sl->sN, sl->sH, sl->sW,
                      sl->scale, x)) {
    Cin1Args a;
    memset(&a, 0, sizeof(a));
    a.x = reinterpret_cast<const float*>(x); a.N = cs.N; a.H = cs.H; a.P = cs.P;
    a.w = w; a.bias = bias; a.y = y; a.relu = relu; a.round_out = round_out; a.out_f16 = out_kind == 1;
    return cin1_conv_fwd(a, st);
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_fwd_geom(&p.g[0], cs, x, src_kind, sl);
  if (rc) return rc;
  const int K = p.g[0].K, kpad = round_up32(K);
  int bn = pick_bn(cs.Cout);
  if (bn == 0 || bn % 16) return VAR_ERR_UNSUPPORTED;
  bn = widen_grid_bn(bn, cs.Cout, (p.g[0].M + 127) / 128, 1, 32);  // epilogue works in 32-column chunks
  p.bn = bn;
  p.nbox = 1; p.box_rows = bn; p.boxbase[0] = 0;
  p.num_kb = kpad / 32;
  pick_pipeline(bn, &p.stages, &p.lookahead);
  EpiParams& e = p.e[0];
  e.out = y; e.ldo = cs.Cout; e.bias = bias; e.ncols = cs.Cout; e.relu = relu;
  e.round_out = round_out;
  CUtensorMap tm, ta;
  rc = get_tmap_2d(w, cs.Cout, kpad, kpad, bn, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm);
  if (rc) return rc;
  ta = tm;
  int gmode = gmode_of(src_kind);
  const int M = p.g[0].M;
  if (gmode == G_VEC_FWD && gather_mode() == 1) {
    if (cs.R == 1 && cs.S == 1 && cs.sh == 1 && cs.sw == 1 && cs.ph == 0 && cs.pw == 0) {
      // 1x1 / Linear: the activation is a plain [M, K] matrix
      rc = get_tmap_2d(reinterpret_cast<const float*>(x), M, cs.Cin, cs.Cin, 128,
                       (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta);
      if (rc) return rc;
      gmode = G_TMA_TILED;
    } else if (cs.R * cs.S <= kMaxTaps) {
      if (conv_halo32_ok(cs)) {
        rc = conv_fwd_halo_tf32(cs, reinterpret_cast<const float*>(x), w, bias, y, relu, round_out, st);
        if (rc != VAR_ERR_UNSUPPORTED) return rc;
      }
      rc = get_tmap_im2col(reinterpret_cast<const float*>(x), cs.N, cs.H, cs.W, cs.Cin, -cs.pw, -cs.ph,
                           cs.pw - (cs.S - 1), cs.ph - (cs.R - 1), cs.sw, cs.sh, 128,
                           (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta);
      if (rc) return rc;
      gmode = G_TMA_IM2COL;
      p.ntaps = cs.R * cs.S; p.cpb = cs.Cin / 32;
      p.base_w = -cs.pw; p.base_h = -cs.ph; p.step_w = cs.sw; p.step_h = cs.sh;
      for (int r = 0; r < cs.R; ++r)
        for (int s_ = 0; s_ < cs.S; ++s_) {
          const int t = r * cs.S + s_;
          p.tap_w[t] = (uint8_t)s_; p.tap_h[t] = (uint8_t)r; p.tap_id[t] = (uint8_t)t;
        }
    }
  }
  dim3 grid((M + 127) / 128, cs.Cout / bn, 1);
  if (gmode == G_SCALAR_F32 || gmode == G_SCALAR_U8) {
    // tiles of whole output rows of one image (<= 128 pixels) over a smem-staged input patch
    if (cs.Q > 128) return VAR_ERR_UNSUPPORTED;
    p.sc_rpt = 128 / cs.Q < cs.P ? 128 / cs.Q : cs.P;
    p.sc_tpi = (cs.P + p.sc_rpt - 1) / p.sc_rpt;
    p.sc_nh = (p.sc_rpt - 1) * cs.sh + cs.R;
    p.sc_wpad = cs.W + 2 * cs.pw;
    grid.x = (unsigned)(cs.N * p.sc_tpi);
    // few k-blocks per tile, no global latency to hide: a shallow pipeline keeps the CTA small
    // so several tiles overlap their stage / build / MMA / epilogue phases on one SM
    p.stages = p.num_kb < 2 ? 1 : 2;
    p.lookahead = 0;
  }
  return launch_gemm(gmode, EPI_STD, tm, tm, ta, ta, p, grid, st);
}

// dgrad geometry shared by all paths: rows of the GEMM are input pixels, K runs over
// (tap, cout), the packed forward weights are read transposed (MN-major B operand).
static int fill_dgrad(GemmParams& p, int z, const ConvShape& cs, const float* dy, float* dx,
                      const float* mask, const float* addsrc, int round_out) {
  if (cs.Cout % 32 || cs.Cin % 32) return VAR_ERR_UNSUPPORTED;
  GatherGeom& g = p.g[z];
  g.src = dy;
  g.M = cs.N * cs.H * cs.W;         // rows are input pixels
  g.P = cs.H; g.Q = cs.W;
  g.H = cs.P; g.W = cs.Q; g.C = cs.Cout;  // source = dY
  g.R = cs.R; g.S = cs.S;
  g.sh = cs.sh; g.sw = cs.sw; g.ph = cs.ph; g.pw = cs.pw;
  g.sC = 1; g.sW = cs.Cout; g.sH = (long long)cs.Q * cs.Cout;
  g.sN = (long long)cs.P * cs.Q * cs.Cout;
  g.K = cs.R * cs.S * cs.Cout;
  g.scale = 1.f;
  int bn = pick_bn(cs.Cin);
  if (bn == 0) return VAR_ERR_UNSUPPORTED;
  bn = widen_grid_bn(bn, cs.Cin, (g.M + 127) / 128, 2, 32);
  p.bn = bn;
  p.b_mn_major = 1;
  p.num_kb = g.K / 32;
  p.kb_per_rs = cs.Cout / 32;
  p.cin_total = cs.Cin;
  p.mn_lbo = mn_cfg().lbo; p.mn_sbo = mn_cfg().sbo; p.mn_type = mn_cfg().type;
  pick_pipeline(bn, &p.stages, &p.lookahead);
  EpiParams& e = p.e[z];
  e.out = dx; e.ldo = cs.Cin; e.ncols = cs.Cin; e.mask = mask; e.ldm = cs.Cin;
  e.addsrc = addsrc; e.lda = cs.Cin;
  e.round_out = round_out;
  return VAR_OK;
}

static bool is_linear(const ConvShape& cs) {
  return cs.R == 1 && cs.S == 1 && cs.sh == 1 && cs.sw == 1 && cs.ph == 0 && cs.pw == 0;
}

int conv_dgrad(const ConvShape& cs, const float* dy, const float* w, float* dx, const float* mask,
               const float* addsrc, int round_out, cudaStream_t st) {
  prof_note("dgrad N%d H%d Ci%d Co%d R%d", cs.N, cs.H, cs.Cin, cs.Cout, cs.R);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_dgrad(p, 0, cs, dy, dx, mask, addsrc, round_out);
  if (rc) return rc;
  const int kfwd = cs.R * cs.S * cs.Cin, kpad = round_up32(kfwd);
  CUtensorMap tm, ta;
  rc = get_tmap_2d(w, cs.Cout, kpad, kpad, 32, mn_cfg().tma_swizzle, &tm);
  if (rc) return rc;
  if (gather_mode() == 0 || cs.R * cs.S > kMaxTaps) {
    dim3 grid((p.g[0].M + 127) / 128, cs.Cin / p.bn, 1);
    return launch_gemm(G_VEC_DGRAD, EPI_STD, tm, tm, tm, tm, p, grid, st);
  }
  if (is_linear(cs)) {
    rc = get_tmap_2d(dy, p.g[0].M, cs.Cout, cs.Cout, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta);
    if (rc) return rc;
    dim3 grid((p.g[0].M + 127) / 128, cs.Cin / p.bn, 1);
    return launch_gemm(G_TMA_TILED, EPI_STD, tm, tm, ta, ta, p, grid, st);
  }
  // One sub-problem per output-pixel parity class (hp, wp): pixels h = h2*sh + hp only meet
  // taps r = r0 + sh*j, so each class is a stride-1 correlation over dY with ~1/(sh*sw) of the
  // taps -- no multiply-by-zero work for strided convolutions.
  if (cs.R < cs.sh || cs.S < cs.sw) return VAR_ERR_UNSUPPORTED;
  p.cpb = cs.Cout / 32;
  p.step_w = 1; p.step_h = 1;
  for (int hp = 0; hp < cs.sh; ++hp)
    for (int wp = 0; wp < cs.sw; ++wp) {
      const int H2 = (cs.H - hp + cs.sh - 1) / cs.sh, W2 = (cs.W - wp + cs.sw - 1) / cs.sw;
      if (H2 <= 0 || W2 <= 0) continue;
      const int r0 = (hp + cs.ph) % cs.sh, s0 = (wp + cs.pw) % cs.sw;
      const int J = (cs.R - r0 + cs.sh - 1) / cs.sh, I = (cs.S - s0 + cs.sw - 1) / cs.sw;
      const int a_h = (hp + cs.ph - r0) / cs.sh, a_w = (wp + cs.pw - s0) / cs.sw;
      GemmParams q = p;
      q.base_h = a_h - (J - 1); q.base_w = a_w - (I - 1);
      q.ntaps = J * I;
      for (int j = 0; j < J; ++j)
        for (int i = 0; i < I; ++i) {
          const int t = j * I + i;
          q.tap_h[t] = (uint8_t)(J - 1 - j); q.tap_w[t] = (uint8_t)(I - 1 - i);
          q.tap_id[t] = (uint8_t)((r0 + cs.sh * j) * cs.S + (s0 + cs.sw * i));
        }
      q.num_kb = q.ntaps * q.cpb;
      GatherGeom& g = q.g[0];
      g.M = cs.N * H2 * W2; g.P = H2; g.Q = W2;
      g.K = q.num_kb * 32;
      EpiParams& e = q.e[0];
      if (cs.sh > 1 || cs.sw > 1) {
        e.map.on = 1; e.map.P2 = H2; e.map.Q2 = W2; e.map.H = cs.H; e.map.W = cs.W;
        e.map.sh = cs.sh; e.map.sw = cs.sw; e.map.oh = hp; e.map.ow = wp;
      }
      rc = get_tmap_im2col(dy, cs.N, cs.P, cs.Q, cs.Cout, q.base_w, q.base_h, W2 - cs.Q + q.base_w,
                           H2 - cs.P + q.base_h, 1, 1, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta);
      if (rc) return rc;
      dim3 grid((g.M + 127) / 128, cs.Cin / q.bn, 1);
      rc = launch_gemm(G_TMA_IM2COL, EPI_STD, tm, tm, ta, ta, q, grid, st);
      if (rc) return rc;
    }
  return VAR_OK;
}

// Two independent linear dgrads of identical shape in one launch (grid.z = 2):
// dx[z] = dy[z] @ W[z] + addsrc[z]   (GRU backward recurrence, both directions)
int linear_dgrad2(int ndir, int M, int Cin, int Cout, const float* const dy[2],
                  const float* const w[2], float* const dx[2], const float* const addsrc[2],
                  int round_out, cudaStream_t st) {
  prof_note("gru_dgrad2 M%d Ci%d Co%d %d%d", M, Cin, Cout, 0, 0);
  ConvShape cs{M, 1, 1, Cin, Cout, 1, 1, 1, 1, 0, 0, 1, 1};
  GemmParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tm[2], ta[2];
  const int kpad = round_up32(Cin);
  const bool tma = gather_mode() == 1;
  for (int z = 0; z < ndir; ++z) {
    int rc = fill_dgrad(p, z, cs, dy[z], dx[z], nullptr, addsrc[z], round_out);
    if (rc) return rc;
    rc = get_tmap_2d(w[z], Cout, kpad, kpad, 32, mn_cfg().tma_swizzle, &tm[z]);
    if (rc) return rc;
    ta[z] = tm[z];
    if (tma) {
      rc = get_tmap_2d(dy[z], M, Cout, Cout, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta[z]);
      if (rc) return rc;
    }
  }
  if (ndir == 1) { tm[1] = tm[0]; ta[1] = ta[0]; }
  dim3 grid((M + 127) / 128, Cin / p.bn, ndir);
  return launch_gemm(tma ? G_TMA_TILED : G_VEC_DGRAD, EPI_STD, tm[0], tm[1], ta[0], ta[1], p, grid, st);
}

// One BPTT step for up to two directions: dh_{s-1} = dgh_s @ W_hh + dh_s * z_s with the cell
// backward of step s-1 in the epilogue (EPI_GRU_BWD).  TMA-fed operands only.
int gru_step_bwd(int ndir, int B, int Hd, const float* const dgh[2], const float* const whh[2],
                 const GruBwdEpiParams q[2], cudaStream_t st) {
  if (gather_mode() != 1) return VAR_ERR_UNSUPPORTED;
  prof_note("gru_bwd_step M%d H%d %d%d%d", B, Hd, 0, 0, 0);
  ConvShape cs{B, 1, 1, Hd, 3 * Hd, 1, 1, 1, 1, 0, 0, 1, 1};
  GemmParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tm[2], ta[2];
  for (int z = 0; z < ndir; ++z) {
    int rc = fill_dgrad(p, z, cs, dgh[z], nullptr, nullptr, nullptr, 0);
    if (rc) return rc;
    p.grub[z] = q[z];
    p.grub[z].Hdim = Hd;
    rc = get_tmap_2d(whh[z], 3 * Hd, Hd, Hd, 32, mn_cfg().tma_swizzle, &tm[z]);
    if (rc) return rc;
    rc = get_tmap_2d(dgh[z], B, 3 * Hd, 3 * Hd, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta[z]);
    if (rc) return rc;
  }
  if (ndir == 1) { tm[1] = tm[0]; ta[1] = ta[0]; }
  dim3 grid((B + 127) / 128, Hd / p.bn, ndir);
  return launch_gemm(G_TMA_TILED, EPI_GRU_BWD, tm[0], tm[1], ta[0], ta[1], p, grid, st);
}

bool gru_persist_enabled() {  // VAR_GRU_PERSIST=0 falls back to one launch per time step
  static int on = -1;
  if (on < 0) { const char* e = getenv("VAR_GRU_PERSIST"); on = (e && e[0] == '0') ? 0 : 1; }
  return on == 1;
}

template <int BWD, bool H16 = false>
static int launch_gru_persist(const CUtensorMap tm[4], GruPersistParams& p, dim3 grid, cudaStream_t st) {
  const size_t smem = gemm_smem_bytes(p.bn, p.stages * p.kps) + gru_scr_bytes(BWD) + 16;
  VAR_ENSURE_SMEM((gru_persist_kernel<BWD, H16>), smem);
  int max_ctas = 0;
  {  // per call and per device: one process may drive several GPUs
    int per_sm = 0, dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gru_persist_kernel<BWD, H16>, kGruThreads, smem);
    max_ctas = per_sm * sms;
  }
  // Batches whose grid cannot be co-resident run as consecutive launches over chunks of row tiles (rows are
  // independent; every chunk uses the same kernel, so results do not depend on the batch size).
  const int nrt = (int)grid.x;
  const int chunk = max_ctas / (int)(grid.y * grid.z);
  if (env_int("VAR_DEBUG", 0))
    fprintf(stderr, "[var] gru_persist<%d,%d>: B %d nrt %d max_ctas %d chunk %d\n", BWD, (int)H16, p.B, nrt, max_ctas, chunk);
  if (chunk < 1) return VAR_ERR_UNSUPPORTED;
  const int nsteps = BWD ? p.T - 1 : p.T;
  static int trace_on = -1;
  if (trace_on < 0) { const char* e = getenv("VAR_GRU_TRACE"); trace_on = (e && e[0] == '1') ? 1 : 0; }
  static long long* d_trace = nullptr;
  for (int rt0 = 0; rt0 < nrt; rt0 += chunk) {
    p.rt0 = rt0;
    grid.x = (unsigned)(nrt - rt0 < chunk ? nrt - rt0 : chunk);
    VAR_CUDA_CHECK(cudaMemsetAsync(p.counters, 0, sizeof(unsigned int) * grid.x * grid.z, st));
    void* args[] = {(void*)&tm[0], (void*)&tm[1], (void*)&tm[2], (void*)&tm[3], (void*)&p};
    const int rows = (int)grid.x * 128 < p.B - rt0 * 128 ? (int)grid.x * 128 : p.B - rt0 * 128;
    LaunchScope sc(BWD ? T_GRU_BWD : T_GRU_STEP, 2.0 * rows * (double)(p.bn * grid.y) * (p.num_kb * (H16 ? 64.0 : 32.0)) * grid.z * nsteps, st);
    if (trace_on) {  // debugging aid: per-step clock64 samples of CTA (0,0,0), dumped after a sync
      if (!d_trace) VAR_CUDA_CHECK(cudaMalloc(&d_trace, sizeof(long long) * 8 * 128));
      VAR_CUDA_CHECK(cudaMemsetAsync(d_trace, 0, sizeof(long long) * 8 * 128, st));
      p.trace = d_trace;
    }
    VAR_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)gru_persist_kernel<BWD, H16>, grid, dim3(kGruThreads, 1, 1), args, smem, st));
    if (trace_on) {
      std::vector<long long> h(8 * 128);
      VAR_CUDA_CHECK(cudaStreamSynchronize(st));
      VAR_CUDA_CHECK(cudaMemcpy(h.data(), d_trace, sizeof(long long) * 8 * 128, cudaMemcpyDeviceToHost));
      FILE* f = fopen(BWD ? "gpurun_out/gru_trace_bwd.csv" : "gpurun_out/gru_trace_fwd.csv", "w");
      if (f) {
        fprintf(f, "step,start,acquired,issued,tfull,epi_done,released\n");
        for (int i = 0; i < nsteps && i < 128; ++i)
          fprintf(f, "%d,%lld,%lld,%lld,%lld,%lld,%lld\n", i, h[i * 8] - h[0], h[i * 8 + 1] - h[0], h[i * 8 + 2] - h[0],
                  h[i * 8 + 3] - h[0], h[i * 8 + 4] - h[0], h[i * 8 + 5] - h[0]);
        fclose(f);
      }
    }
  }
  return VAR_OK;
}

// All T forward steps of both directions in one cooperative launch (gru_persist.cuh).
// h_r: [(T+1), B, H] with slot 0 zeroed; h32[d][0] zeroed.  Returns VAR_ERR_UNSUPPORTED when the
// grid cannot be co-resident (caller falls back to per-step launches).
bool gru_x16_enabled() {  // VAR_GRU_X16=0 keeps the tf32 GEMMs for the GRU input projection and its backward
  static int on = -1;
  if (on < 0) on = env_int("VAR_GRU_X16", 1);
  return on && gru_h16_enabled() && gather_mode() == 1;
}
bool gru_h16_enabled() {  // VAR_GRU_H16=0 keeps tf32 operands in the recurrent kernels
  static int on = -1;
  if (on < 0) on = env_int("VAR_GRU_H16", 1) && env_int("VAR_H16", 1);
  return on == 1;
}

int gru_persist_fwd(int B, int Hd, int T, const float* const xproj[2], long long ldx, const float* const whh[2],
                    const float* const bhh[2], float* const h32[2][2], float* const h_r[2],
                    float* const gates[2], float* const hn_save[2], unsigned int* counters, cudaStream_t st,
                    const void* const whh16[2], void* const h_h[2], long long x_tstride) {
  if (!gru_persist_enabled() || gather_mode() != 1 || Hd % 64) return VAR_ERR_UNSUPPORTED;
  const bool h16 = gru_h16_enabled() && whh16 && h_h && whh16[0] && h_h[0];
  prof_note(h16 ? "gru_persist_fwd16 M%d H%d T%d %d%d" : "gru_persist_fwd M%d H%d T%d %d%d", B, Hd, T, 0, 0);
  GruPersistParams p;
  memset(&p, 0, sizeof(p));
  const int jb = 32;
  const int ke = h16 ? 64 : 32;
  p.B = B; p.Hd = Hd; p.T = T; p.bn = 3 * jb; p.num_kb = Hd / ke; p.stages = env_int("VAR_GRU_STAGES_FWD", 2);
  p.kps = env_int("VAR_GRU_KPS_FWD", 2);
  p.a_split = 0;  // (the split-A-box experiment is retired: its extra issuers read h without acquiring the group counter)
  p.counters = counters; p.ldx = ldx; p.xts = x_tstride ? x_tstride : 3LL * Hd;
  CUtensorMap tm[4];
  for (int d = 0; d < 2; ++d) {
    p.xproj[d] = xproj[d]; p.bhh[d] = bhh[d];
    p.h32[d][0] = h32[d][0]; p.h32[d][1] = h32[d][1];
    p.h_r[d] = h_r[d]; p.gates[d] = gates[d]; p.hn_save[d] = hn_save[d];
    int rc;
    if (h16) {
      p.h_h[d] = reinterpret_cast<uint16_t*>(h_h[d]);
      rc = get_tmap_2d_e(whh16[d], 2, 3 * Hd, Hd, Hd, jb, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm[d]);
      if (rc) return rc;
      rc = get_tmap_2d_e(h_h[d], 2, (T + 1) * B, Hd, Hd, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm[2 + d]);
    } else {
      rc = get_tmap_2d(whh[d], 3 * Hd, Hd, Hd, jb, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm[d]);
      if (rc) return rc;
      rc = get_tmap_2d(h_r[d], (T + 1) * B, Hd, Hd, p.a_split ? 32 : 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm[2 + d]);
    }
    if (rc) return rc;
  }
  if (h16) p.a_split = 0;
  dim3 grid((B + 127) / 128, Hd / jb, 2);
  p.arrivals = (int)grid.y;
  if (h16) return launch_gru_persist<0, true>(tm, p, grid, st);
  return launch_gru_persist<0>(tm, p, grid, st);
}

// K-split variant of the BPTT kernel (gru_ksplit.cuh): 2-CTA clusters, each CTA streams half of the
// reduction dimension.  Returns VAR_ERR_UNSUPPORTED when the clusters cannot all be resident.
template <bool H16>
static int launch_gru_bwd_ksplit(const CUtensorMap tm[4], GruPersistParams& p, int nrt, cudaStream_t st) {
  p.bn = 64; p.num_kb = (3 * p.Hd / (H16 ? 64 : 32)) / 2; p.kps = 2; p.stages = 3;
  if (p.num_kb % p.kps) return VAR_ERR_UNSUPPORTED;
  const size_t smem = gemm_smem_bytes(p.bn, p.stages * p.kps) + 2 * gru_scr_bytes(1) + 32;
  // (cudaFuncSetAttribute is per device: set it every call -- it is a cheap host-side call)
  VAR_CUDA_CHECK(cudaFuncSetAttribute(gru_bwd_ksplit_kernel<H16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(2, p.Hd / 64, 2);  // .z = row tiles of the chunk * 2 directions, set below
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = dim3(kGruThreads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  // Cluster + cooperative attributes: the kernel spin-waits on peer CTAs, so the driver must guarantee
  // that the whole grid is co-resident (the image branch runs concurrently on another stream).  ncu
  // cannot replay a cooperative cluster launch (it aborts the application with LaunchFailed), so
  // VAR_GRU_NO_COOP=1 drops the cooperative attribute for profiling runs only; co-residency is then
  // merely checked against cudaOccupancyMaxActiveClusters.  All caches are per device.
  constexpr int kMaxDev = 64;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= kMaxDev) return VAR_ERR_UNSUPPORTED;
  static int no_coop = -1;
  if (no_coop < 0) no_coop = env_int("VAR_GRU_NO_COOP", 0);
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 1;  // the occupancy query below takes the cluster shape only
  static int max_clusters[kMaxDev];
  static bool have_max[kMaxDev], refused[kMaxDev], coop_refused[kMaxDev];
  if (!have_max[dev]) {
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gru_bwd_ksplit_kernel<H16>, &cfg) != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
    max_clusters[dev] = n;
    have_max[dev] = true;
  }
  // refused: a launch was rejected once on this device (e.g. under a profiler that cannot replay it)
  const int chunk = max_clusters[dev] / (int)(grid.y * 2);  // row tiles whose clusters are co-resident
  if (env_int("VAR_DEBUG", 0))
    fprintf(stderr, "[var] gru_bwd_ksplit<%d>: B %d nrt %d max_clusters %d chunk %d refused %d coop_refused %d\n", (int)H16, p.B,
            nrt, max_clusters[dev], chunk, (int)refused[dev], (int)coop_refused[dev]);
  if (refused[dev] || chunk < 1) return VAR_ERR_UNSUPPORTED;
  for (int rt0 = 0; rt0 < nrt; rt0 += chunk) {  // larger batches: consecutive launches over row-tile chunks
    const int n = nrt - rt0 < chunk ? nrt - rt0 : chunk;
    p.rt0 = rt0;
    grid.z = (unsigned)(n * 2);
    cfg.gridDim = grid;
    VAR_CUDA_CHECK(cudaMemsetAsync(p.counters, 0, sizeof(unsigned int) * n * 2, st));
    void* args[] = {(void*)&tm[0], (void*)&tm[1], (void*)&tm[2], (void*)&tm[3], (void*)&p};
    const int rows = n * 128 < p.B - rt0 * 128 ? n * 128 : p.B - rt0 * 128;
    LaunchScope sc(T_GRU_BWD, 2.0 * rows * (double)p.Hd * (3.0 * p.Hd) * 2 * (p.T - 1), st);
    bool done = false;
    if (!no_coop && !coop_refused[dev]) {
      cfg.numAttrs = 2;
      if (cudaLaunchKernelExC(&cfg, (const void*)gru_bwd_ksplit_kernel<H16>, args) == cudaSuccess) done = true;
      else {
        (void)cudaGetLastError();  // not sticky
        coop_refused[dev] = true;
        static bool said = false;
        if (!said) { fprintf(stderr, "[var] cooperative cluster launch of the K-split BPTT kernel refused; using the occupancy check\n"); said = true; }
      }
    }
    if (!done) {
      cfg.numAttrs = 1;
      if (cudaLaunchKernelExC(&cfg, (const void*)gru_bwd_ksplit_kernel<H16>, args) != cudaSuccess) {
        (void)cudaGetLastError();  // not sticky: fall back to the one-CTA-per-tile kernel from now on
        refused[dev] = true;
        if (rt0 > 0) { var_set_last_error("BPTT chunk launch refused after earlier chunks ran", __FILE__, __LINE__); return VAR_ERR_CUDA; }
        return VAR_ERR_UNSUPPORTED;
      }
    }
  }
  return VAR_OK;
}

// BPTT steps T-1 .. 1 of both directions in one cooperative launch: dgh[T-1] and dhd[d][0] must hold
// the cell backward of the last step (gru_cell_bwd).
int gru_persist_bwd(int B, int Hd, int T, const float* const whh[2], const float* const gates[2],
                    const float* const hn_save[2], const float* const h_r[2], float* const dgh[2],
                    float* const dgi[2], float* const dhd[2][2], unsigned int* counters, cudaStream_t st,
                    const GruBwdExtra* ex) {
  if (!gru_persist_enabled() || gather_mode() != 1 || Hd % 32) return VAR_ERR_UNSUPPORTED;
  const bool h16 = ex && gru_h16_enabled() && ex->whh16[0] && ex->dgh_h[0] && ex->gscale[0] && Hd % 64 == 0 &&
                   env_int("VAR_GRU_KSPLIT", 1);
  prof_note("gru_persist_bwd M%d H%d T%d %d%d", B, Hd, T, 0, 0);
  GruPersistParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.Hd = Hd; p.T = T; p.bn = 32; p.num_kb = 3 * Hd / 32; p.stages = env_int("VAR_GRU_STAGES_BWD", 3);
  p.kps = env_int("VAR_GRU_KPS_BWD", 2);
  p.a_split = 0;  // (the split-A-box experiment is retired: its extra issuers read h without acquiring the group counter)
  p.counters = counters;
  p.mn_lbo = mn_cfg().lbo; p.mn_sbo = mn_cfg().sbo; p.mn_type = mn_cfg().type;
  CUtensorMap tm[4];
  for (int d = 0; d < 2; ++d) {
    p.gates_c[d] = gates[d]; p.hn_save_c[d] = hn_save[d]; p.h_r_c[d] = h_r[d];
    p.dgh[d] = dgh[d]; p.dgi[d] = dgi[d]; p.dhd[d][0] = dhd[d][0]; p.dhd[d][1] = dhd[d][1];
    int rc = get_tmap_2d(whh[d], 3 * Hd, Hd, Hd, 32, mn_cfg().tma_swizzle, &tm[d]);
    if (rc) return rc;
    rc = get_tmap_2d(dgh[d], T * B, 3 * Hd, 3 * Hd, p.a_split ? 32 : 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm[2 + d]);
    if (rc) return rc;
  }
  if (ex) {
    for (int d = 0; d < 2; ++d) { p.db_ih[d] = ex->db_ih[d]; p.db_hh[d] = ex->db_hh[d]; }
    if (ex->bias_done) *ex->bias_done = 0;
    if (ex->dgi_h_done) *ex->dgi_h_done = 0;
  }
  if (h16) {
    CUtensorMap th[4];
    int rc = VAR_OK;
    for (int d = 0; d < 2 && !rc; ++d) {
      p.dgh_h[d] = reinterpret_cast<uint16_t*>(ex->dgh_h[d]); p.gscale[d] = ex->gscale[d];
      rc = get_tmap_2d_e(ex->whh16[d], 2, 3 * Hd, Hd, Hd, 64, (int)CU_TENSOR_MAP_SWIZZLE_128B, &th[d]);
      if (!rc) rc = get_tmap_2d_e(ex->dgh_h[d], 2, T * B, 3 * Hd, 3 * Hd, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &th[2 + d]);
    }
    if (rc) return rc;
    prof_note("gru_ksplit_bwd16 M%d H%d T%d %d%d", B, Hd, T, 0, 0);
    if (ex->dgi_h[0]) {
      for (int d = 0; d < 2; ++d) p.dgi_h[d] = reinterpret_cast<uint16_t*>(ex->dgi_h[d]);
      p.dgi_h_ld = ex->dgi_h_ld; p.dgi_h_ts = ex->dgi_h_ts; p.skip_f32 = 1;
    }
    rc = launch_gru_bwd_ksplit<true>(th, p, (B + 127) / 128, st);
    if (rc == VAR_OK && ex->bias_done) *ex->bias_done = ex->db_ih[0] != nullptr;
    if (rc == VAR_OK && ex->dgi_h_done) *ex->dgi_h_done = ex->dgi_h[0] != nullptr;
    if (rc != VAR_ERR_UNSUPPORTED) return rc;
    p.dgi_h[0] = p.dgi_h[1] = nullptr; p.skip_f32 = 0;
  }
  if (env_int("VAR_GRU_KSPLIT", 1) && Hd % 64 == 0 && !p.a_split) {
    prof_note("gru_ksplit_bwd M%d H%d T%d %d%d", B, Hd, T, 0, 0);
    const int rc = launch_gru_bwd_ksplit<false>(tm, p, (B + 127) / 128, st);
    if (rc == VAR_OK && ex && ex->bias_done) *ex->bias_done = ex->db_ih[0] != nullptr;
    if (rc != VAR_ERR_UNSUPPORTED) return rc;
    p.bn = 32; p.num_kb = 3 * Hd / 32;
    p.stages = env_int("VAR_GRU_STAGES_BWD", 3); p.kps = env_int("VAR_GRU_KPS_BWD", 2);
  }
  for (int d = 0; d < 2; ++d) { p.db_ih[d] = nullptr; p.db_hh[d] = nullptr; }  // fallback kernel: caller sums the columns
  dim3 grid((B + 127) / 128, Hd / p.bn, 2);
  p.arrivals = (int)grid.y;
  return launch_gru_persist<1>(tm, p, grid, st);
}

// One GRU time step for up to two directions: gates = hprev @ W_hh^T fused with the
// cell update in the epilogue (EPI_GRU_FWD).  W_hh is [3H, H] K-major, rows r|z|n.
int gru_step_fwd(int ndir, int B, int Hd, const float* const hprev_r[2], const float* const whh[2],
                 const GruEpiParams q[2], cudaStream_t st) {
  if (Hd % 64) return VAR_ERR_UNSUPPORTED;
  // hidden units per CTA: 32 (N = 96) doubles the grid when the batch gives few M tiles
  const int jb = ((B + 127) / 128) * (Hd / 64) * ndir < kNumSMs ? 32 : 64;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  CUtensorMap tm[2], ta[2];
  const bool tma = gather_mode() == 1;
  for (int z = 0; z < ndir; ++z) {
    GatherGeom& g = p.g[z];
    g.src = hprev_r[z];
    g.M = B; g.P = 1; g.Q = 1; g.H = 1; g.W = 1; g.C = Hd; g.R = 1; g.S = 1;
    g.sh = g.sw = 1; g.ph = g.pw = 0;
    g.sC = 1; g.sW = Hd; g.sH = Hd; g.sN = Hd; g.K = Hd; g.scale = 1.f;
    p.gru[z] = q[z];
    p.gru[z].Hdim = Hd;
    int rc = get_tmap_2d(whh[z], 3 * Hd, Hd, Hd, jb, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm[z]);
    if (rc) return rc;
    ta[z] = tm[z];
    if (tma) {
      rc = get_tmap_2d(hprev_r[z], B, Hd, Hd, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta[z]);
      if (rc) return rc;
    }
  }
  if (ndir == 1) { tm[1] = tm[0]; ta[1] = ta[0]; }
  p.bn = 3 * jb;
  p.nbox = 3; p.box_rows = jb;
  p.boxbase[0] = 0; p.boxbase[1] = Hd; p.boxbase[2] = 2 * Hd;
  p.num_kb = Hd / 32;
  p.stages = 4; p.lookahead = 2;
  dim3 grid((B + 127) / 128, Hd / jb, ndir);
  return launch_gemm(tma ? G_TMA_TILED : G_VEC_FWD, EPI_GRU_FWD, tm[0], tm[1], ta[0], ta[1], p, grid, st);
}

// ---- bias gradient: column sums of dY [M, ld].  CTA = 8 row lanes x 32 float4 columns
// (a 128-column slab), grid = (row chunks, slabs); 128-bit coalesced loads, one atomic per
// column per CTA.
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dy, long long M, long long ld, int C, float* __restrict__ db,
              int rows_per_cta) {
  __shared__ float4 red[256];
  const int c4 = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int col = blockIdx.y * 128 + c4 * 4;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(r0 + (long long)rows_per_cta, M);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < C) {
    for (long long r = r0 + rl; r < r1; r += 8) {
      const float4 v = *reinterpret_cast<const float4*>(dy + r * ld + col);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0 && col < C) {
#pragma unroll
    for (int l = 1; l < 8; ++l) {
      const float4 v = red[l * 32 + c4];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    atomicAdd(db + col, acc.x); atomicAdd(db + col + 1, acc.y);
    atomicAdd(db + col + 2, acc.z); atomicAdd(db + col + 3, acc.w);
  }
}

// Narrow dense matrices (ld == C, C | 1024): the slab kernel would leave 3/4 (C = 32) of its lanes
// idle.  Here dY is one flat float4 stream; with a grid stride that is a multiple of C/4 every
// thread stays on its own 4 columns, keeps 4 independent loads in flight, and the CTA folds its
// 256 partial sums per column group through shared memory.
__global__ void __launch_bounds__(256)
colsum_flat_kernel(const float4* __restrict__ dy, long long total4, int C4, float* __restrict__ db) {
  __shared__ float4 red[256];
  const long long stride = (long long)gridDim.x * 256;
  long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
  for (; i + 3 * stride < total4; i += 4 * stride) {
    const float4 v0 = dy[i], v1 = dy[i + stride], v2 = dy[i + 2 * stride], v3 = dy[i + 3 * stride];
    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
    a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
    a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
    a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
  }
  for (; i < total4; i += stride) {
    const float4 v0 = dy[i];
    a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
  }
  a0.x += a1.x + (a2.x + a3.x); a0.y += a1.y + (a2.y + a3.y);
  a0.z += a1.z + (a2.z + a3.z); a0.w += a1.w + (a2.w + a3.w);
  red[threadIdx.x] = a0;
  __syncthreads();
  if (threadIdx.x < C4) {  // 256 % C4 == 0: thread t and t + k*C4 share their 4 columns
    float4 acc = red[threadIdx.x];
    for (int l = threadIdx.x + C4; l < 256; l += C4) {
      const float4 v = red[l];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int col = threadIdx.x * 4;
    atomicAdd(db + col, acc.x); atomicAdd(db + col + 1, acc.y);
    atomicAdd(db + col + 2, acc.z); atomicAdd(db + col + 3, acc.w);
  }
}

int colsum(const float* dy, long long M, long long ld, int C, float* db, cudaStream_t st) {
  if ((C & 3) || (ld & 3)) return VAR_ERR_UNSUPPORTED;
  if (ld == C && C < 128 && 1024 % C == 0 && (reinterpret_cast<uintptr_t>(dy) & 15) == 0) {
    const long long total4 = M * C / 4;
    long long ctas = (total4 + 256 * 8 - 1) / (256 * 8);
    if (ctas > 8 * kNumSMs) ctas = 8 * kNumSMs;
    if (ctas < 1) ctas = 1;
    LaunchScope sc(T_COLSUM, 0, st);
    colsum_flat_kernel<<<(unsigned)ctas, 256, 0, st>>>(reinterpret_cast<const float4*>(dy), total4, C / 4, db);
    VAR_CUDA_CHECK(cudaGetLastError());
    return VAR_OK;
  }
  const int slabs = (C + 127) / 128;
  long long chunks = (4 * kNumSMs + slabs - 1) / slabs;
  long long rp = (M + chunks - 1) / chunks;
  if (rp < 64) rp = 64;
  dim3 grid((unsigned)((M + rp - 1) / rp), slabs);
  LaunchScope sc(T_COLSUM, 0, st);
  colsum_kernel<<<grid, 256, 0, st>>>(dy, M, ld, C, db, (int)rp);
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

static int conv_wgrad_tma(const ConvShape& cs, const float* x, const float* dy, float* dw, float* db, int* db_done,
                          cudaStream_t st) {
  WgradTmaParams p;
  memset(&p, 0, sizeof(p));
  p.M = cs.N * cs.P * cs.Q;
  p.K = cs.R * cs.S * cs.Cin;
  p.kpad = round_up32(p.K);
  // two 32-pixel blocks per pipeline stage halve the barrier round trips: +12 % for the narrow convs
  // (Cout <= 64); wide slabs (GRU, 256 columns) lose 15 % with the coarser stages
  p.kps = env_int("VAR_WGRAD_KPS", cs.Cout <= 64 ? 2 : 1);
  p.stages = 4 / p.kps;
  p.mn_lbo = mn_cfg().lbo; p.mn_sbo = mn_cfg().sbo; p.mn_type = mn_cfg().type;
  p.P = cs.P; p.Q = cs.Q;
  p.a_tiled = is_linear(cs) ? 1 : 0;
  CUtensorMap tx, tdy;
  int rc;
  if (p.a_tiled) {
    rc = get_tmap_2d(x, p.M, cs.Cin, cs.Cin, 32, mn_cfg().tma_swizzle, &tx);
  } else {
    p.cpb = cs.Cin / 32;
    p.base_w = -cs.pw; p.base_h = -cs.ph; p.step_w = cs.sw; p.step_h = cs.sh;
    for (int r = 0; r < cs.R; ++r)
      for (int s_ = 0; s_ < cs.S; ++s_) { p.tap_w[r * cs.S + s_] = (uint8_t)s_; p.tap_h[r * cs.S + s_] = (uint8_t)r; }
    rc = get_tmap_im2col(x, cs.N, cs.H, cs.W, cs.Cin, -cs.pw, -cs.ph, cs.pw - (cs.S - 1),
                         cs.ph - (cs.R - 1), cs.sw, cs.sh, 32, mn_cfg().tma_swizzle, &tx);
  }
  if (rc) return rc;
  const int ktiles = (p.K + 127) / 128;
  const int slab = pick_bn(cs.Cout);
  if (slab == 0 || slab % 32) return VAR_ERR_UNSUPPORTED;
  const int nslab = cs.Cout / slab;
  // bias gradient from an all-ones row group when the last k tile has a free one (VAR_WGRAD_ONES=0: column-sum pass)
  p.ones_ktile = (p.K % 128 != 0 && env_int("VAR_WGRAD_ONES", 1)) ? ktiles - 1 : -1;
  p.db = (db && p.ones_ktile >= 0) ? db : nullptr;
  *db_done = p.db != nullptr;
  // 2 CTAs fit per SM: size the pixel split so the grid is a whole number of 296-CTA waves
  // (a 616-CTA grid runs three waves for 2.08 waves of work)
  const int per_split = ktiles * nslab;
  int splits = (2 * 2 * kNumSMs) / per_split;
  if (splits < 1) splits = 1;
  int ppc = (p.M + splits - 1) / splits;
  ppc = ((ppc + 32 * p.kps - 1) / (32 * p.kps)) * (32 * p.kps);
  if (ppc < 256) ppc = 256;
  splits = (p.M + ppc - 1) / ppc;
  p.pix_per_cta = ppc;
  p.cout = slab;
  const size_t smem = wgrad_smem_bytes(slab, p.stages * p.kps);
  VAR_ENSURE_SMEM(tc_wgrad_tma_kernel<0>, smem);
  // all slabs of output channels in one launch (grid.z): one slab alone is ktiles * splits CTAs and
  // would leave a third of the SMs idle for wide layers (GRU: 6 slabs of 96 CTAs)
  dim3 grid(ktiles, splits, nslab);
  rc = get_tmap_2d(dy, p.M, cs.Cout, cs.Cout, 32, mn_cfg().tma_swizzle, &tdy);
  if (rc) return rc;
  p.dw = dw;
  {
    LaunchScope sc(T_WGRAD, 2.0 * p.M * (double)cs.Cout * p.K, st);
    tc_wgrad_tma_kernel<0><<<grid, 160, smem, st>>>(tx, tdy, p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

int conv_wgrad(const ConvShape& cs, const void* x, int src_kind, const SrcLayout* sl,
               const float* dy, float* dw, float* db, cudaStream_t st) {
  prof_note("wgrad N%d H%d Ci%d Co%d R%d", cs.N, cs.H, cs.Cin, cs.Cout, cs.R);
  if (src_kind == SRC_STRIDED_U8 && sl &&
      cin3_conv_match(cs.H, cs.W, cs.Cin, cs.Cout, cs.R, cs.S, cs.sh, cs.sw, cs.ph, cs.pw, cs.P, cs.Q, sl->sN, sl->sH,
                      sl->sW, sl->sC, x)) {
    Cin3Args a;
    memset(&a, 0, sizeof(a));
    a.x = reinterpret_cast<const unsigned char*>(x); a.scale = sl->scale;
    a.N = cs.N; a.H = cs.H; a.W = cs.W; a.P = cs.P; a.Q = cs.Q; a.stride = cs.sh;
    a.dy = dy; a.dw = dw; a.db = db;
    return cin3_conv_wgrad(a, st);
  }
  if (is_first_conv(cs, src_kind) && sl) {
    FirstConvArgs a;
    memset(&a, 0, sizeof(a));
    a.x = x; a.sN = sl->sN; a.sH = sl->sH; a.sW = sl->sW; a.sC = sl->sC; a.scale = sl->scale;
    a.N = cs.N; a.H = cs.H; a.W = cs.W; a.P = cs.P; a.Q = cs.Q; a.stride = cs.sh;
    a.kpad = round_up32(27); a.dy = dy; a.dw = dw; a.db = db;
    return first_conv_wgrad(a, src_kind == SRC_STRIDED_U8, st);
  }
  if (src_kind == SRC_STRIDED_F32 && sl && env_int("VAR_FULLW", 1) &&
      fullw_conv_match(cs.H, cs.W, cs.Cin, cs.Cout, cs.R, cs.S, cs.sh, cs.sw, cs.ph, cs.pw, cs.P, cs.Q, sl->sN, sl->sH,
                       sl->sW, sl->scale, x))
    return fullw_conv_wgrad(reinterpret_cast<const float*>(x), dy, dw, db, cs.N, cs.H, cs.W, cs.R, cs.sh, cs.P,
                            round_up32(cs.R * cs.S * cs.Cin), st);
  if (src_kind == SRC_STRIDED_F32 && sl &&
      cin1_conv_match(cs.H, cs.W, cs.Cin, cs.Cout, cs.R, cs.S, cs.sh, cs.sw, cs.ph, cs.pw, sl->sN, sl->sH, sl->sW,
                      sl->scale, x)) {
    Cin1Args a;  // bias gradient comes out of the same kernel (all-ones tap row)
    memset(&a, 0, sizeof(a));
    a.x = reinterpret_cast<const float*>(x); a.N = cs.N; a.H = cs.H; a.P = cs.P;
    a.dy = dy; a.dw = dw; a.db = db;
    return cin1_conv_wgrad(a, st);
  }
  if (src_kind == SRC_NHWC_F32 && gather_mode() == 1 && cs.Cin % 32 == 0 && cs.Cout % 32 == 0 &&
      cs.R * cs.S <= kMaxTaps) {
    int db_done = 0;
    const int rc = conv_wgrad_tma(cs, reinterpret_cast<const float*>(x), dy, dw, db, &db_done, st);
    if (rc) return rc;
    if (db && !db_done) return colsum(dy, (long long)cs.N * cs.P * cs.Q, cs.Cout, cs.Cout, db, st);
    return VAR_OK;
  }
  WgradParams p;
  memset(&p, 0, sizeof(p));
  int rc = fill_fwd_geom(&p.g, cs, x, src_kind, sl);
  if (rc) return rc;
  if (cs.Cout % 16) return VAR_ERR_UNSUPPORTED;
  const int K = p.g.K;
  p.kpad = round_up32(K);
  p.ldy = cs.Cout;
  p.stages = 4; p.lookahead = 2;
  p.mn_lbo = mn_cfg().lbo; p.mn_sbo = mn_cfg().sbo; p.mn_type = mn_cfg().type;
  p.mn_swz32 = mn_cfg().swz32;
  const int ktiles = (K + 127) / 128;
  const int M = p.g.M;
  const int slab = pick_bn(cs.Cout);
  if (slab == 0) return VAR_ERR_UNSUPPORTED;
  dim3 grid(ktiles, 1, 1);
  if (src_kind != SRC_NHWC_F32) {
    // first layers: a CTA owns whole output rows of one image; the largest row tile whose staged
    // input patch stays under 64 KB (2 CTAs per SM with a 2-stage operand pipeline)
    p.sc_wpad = cs.W + 2 * cs.pw;
    int rpt = cs.P;
    while (rpt > 1 && (size_t)cs.Cin * ((rpt - 1) * cs.sh + cs.R) * p.sc_wpad * 4 > 64 * 1024) rpt = (rpt + 1) / 2;
    p.sc_rpt = rpt;
    p.sc_tpi = (cs.P + rpt - 1) / rpt;
    p.sc_nh = (rpt - 1) * cs.sh + cs.R;
    p.stages = 2; p.lookahead = 1;
    p.pix_per_cta = rpt * cs.Q;
    grid.y = (unsigned)(cs.N * p.sc_tpi);
  } else {
    const int nslab = cs.Cout / slab;
    int splits = (4 * kNumSMs + ktiles * nslab - 1) / (ktiles * nslab);
    int ppc = (M + splits - 1) / splits;
    ppc = ((ppc + 31) / 32) * 32;
    if (ppc < 256) ppc = 256;
    splits = (M + ppc - 1) / ppc;
    p.pix_per_cta = ppc;
    grid.y = (unsigned)splits;
  }
  const int gm = gmode_of(src_kind);
  for (int c0 = 0; c0 < cs.Cout; c0 += slab) {
    p.dy = dy + c0; p.cout = slab; p.dw = dw + (long long)c0 * p.kpad;
    if (gm == G_VEC_FWD) rc = launch_wgrad_t<G_VEC_FWD>(p, grid, st);
    else if (gm == G_SCALAR_F32) rc = launch_wgrad_t<G_SCALAR_F32>(p, grid, st);
    else rc = launch_wgrad_t<G_SCALAR_U8>(p, grid, st);
    if (rc) return rc;
  }
  if (db) return colsum(dy, M, cs.Cout, cs.Cout, db, st);
  return VAR_OK;
}

// ===========================================================================
// 16-bit operand region (f16 activations / weights / scaled gradients; kind::f16 MMAs).  Used for convs
// with Cin % 64 == 0 whose N = 32 / 64 tiles are bound by the shared-memory fill per MAC in tf32.
// ===========================================================================
static void fill_fwd_taps(GemmParams& p, const ConvShape& cs) {
  p.ntaps = cs.R * cs.S;
  p.base_w = -cs.pw; p.base_h = -cs.ph; p.step_w = cs.sw; p.step_h = cs.sh;
  for (int r = 0; r < cs.R; ++r)
    for (int s_ = 0; s_ < cs.S; ++s_) {
      const int t = r * cs.S + s_;
      p.tap_w[t] = (uint8_t)s_; p.tap_h[t] = (uint8_t)r; p.tap_id[t] = (uint8_t)t;
    }
}

bool conv_h16_ok(const ConvShape& cs) {
  static int on = -1;
  if (on < 0) on = env_int("VAR_H16", 1);
  return on && gather_mode() == 1 && cs.Cin % 64 == 0 && cs.Cout % 64 == 0 && cs.Cout <= 256 && cs.R * cs.S <= kMaxTaps &&
         cs.R >= cs.sh && cs.S >= cs.sw && !is_linear(cs);
}

// ---- im2col-free forward for the strided 64 -> 64 f16 convs (halo_conv.cuh) ----
static int get_tmap_nhwc_strided(const void* x, int N, int H, int W, int C, int box_w, int box_h, int sw, int sh,
                                 CUtensorMap* out, int esize = 2) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    var_set_last_error("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(x) & 15) || C * esize != 128) return VAR_ERR_ARG;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gstr[3] = {(cuuint64_t)C * esize, (cuuint64_t)W * C * esize, (cuuint64_t)H * W * C * esize};
  // boxDim counts tensor elements spanned: ceil(boxDim / elementStride) elements land in shared memory
  cuuint32_t box[4] = {(cuuint32_t)C, (cuuint32_t)(box_w * sw), (cuuint32_t)(box_h * sh), 1u};
  cuuint32_t estr[4] = {1u, (cuuint32_t)sw, (cuuint32_t)sh, 1u};
  if (box[1] > 256 || box[2] > 256) return VAR_ERR_UNSUPPORTED;
  CUresult r = fn(out, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(x), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (strided NHWC) failed: %d (N=%d H=%d W=%d box=%dx%d)", (int)r, N, H, W,
             box_w, box_h);
    var_set_last_error(buf, __FILE__, __LINE__);
    return VAR_ERR_CUDA;
  }
  return VAR_OK;
}

static bool conv_halo_ok(const ConvShape& cs) {
  const int on = env_int("VAR_HALO", 1);
  return on && cs.Cin == 64 && cs.Cout == 64 && cs.sh == 2 && cs.sw == 2 && cs.R * cs.S <= kMaxTaps && cs.R >= 2 && cs.S >= 2;
}

static int conv_fwd_halo_h16(const ConvShape& cs, const void* x, const void* w, const float* bias, void* y, int out_kind,
                             int relu, int round_out, cudaStream_t st) {
  auto fdiv2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };  // floor(v / 2)
  int rjmin = 1 << 30, rjmax = -(1 << 30), sjmin = 1 << 30, sjmax = -(1 << 30);
  for (int r = 0; r < cs.R; ++r) { const int j = fdiv2(r - cs.ph); rjmin = j < rjmin ? j : rjmin; rjmax = j > rjmax ? j : rjmax; }
  for (int s_ = 0; s_ < cs.S; ++s_) { const int j = fdiv2(s_ - cs.pw); sjmin = j < sjmin ? j : sjmin; sjmax = j > sjmax ? j : sjmax; }
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.WP = cs.Q + (sjmax - sjmin);
  if (p.WP > 128) return VAR_ERR_UNSUPPORTED;
  p.TP = 128 / p.WP;
  if (p.TP > cs.P) p.TP = cs.P;
  p.TPI = (cs.P + p.TP - 1) / p.TP;
  p.box_rows = p.TP + (rjmax - rjmin); p.box_cols = p.WP;
  p.plane_stride = (uint32_t)(((size_t)p.box_rows * p.box_cols * 128 + 1023) / 1024 * 1024);
  p.h_start = 2 * rjmin; p.w_start = 2 * sjmin;
  p.N = cs.N; p.P = cs.P; p.Q = cs.Q;
  p.bias = bias; p.out = y; p.out_kind = out_kind; p.relu = relu; p.round_out = round_out;
  p.ntaps = cs.R * cs.S;
  p.nplanes = 4; p.step_h = 2; p.bn = 64; p.kelems = 64; p.w_resident = 0;
  int t = 0;
  for (int pl = 0; pl < 4; ++pl) {  // plane = rp * 2 + sp
    p.plane_begin[pl] = t;
    for (int r = 0; r < cs.R; ++r)
      for (int s_ = 0; s_ < cs.S; ++s_) {
        const int rj = fdiv2(r - cs.ph), sj = fdiv2(s_ - cs.pw);
        const int rp = (r - cs.ph) - 2 * rj, sp = (s_ - cs.pw) - 2 * sj;
        if (rp * 2 + sp != pl) continue;
        p.tap_shift[t] = (uint16_t)((rj - rjmin) * p.WP + (sj - sjmin));
        p.tap_wcol[t] = (uint16_t)(r * cs.S + s_);
        ++t;
      }
  }
  p.plane_begin[4] = t;
  // several CTAs per SM: one MMA issuer sustains only ~1 MMA (128 x 64 x 16) per 127 clk, independent streams overlap
  const int cps = env_int("VAR_HALO_CPS", 3);
  p.slots = env_int("VAR_HALO_SLOTS", cps >= 3 ? 2 : 3);
  p.stages = env_int("VAR_HALO_STAGES", cps >= 3 ? 3 : 4);
  size_t smem = halo_smem_bytes(p.plane_stride, p.stages, p.slots);
  const size_t per_sm = 227 * 1024 - (size_t)cps * 1024;  // 1 KB reserved per resident CTA
  while (smem * cps > per_sm && p.stages > 2) { --p.stages; smem = halo_smem_bytes(p.plane_stride, p.stages, p.slots); }
  while (smem * cps > per_sm && p.slots > 2) { --p.slots; smem = halo_smem_bytes(p.plane_stride, p.stages, p.slots); }
  if (smem * cps > per_sm) return VAR_ERR_UNSUPPORTED;
  const int kpad = round_up32(cs.R * cs.S * cs.Cin);
  CUtensorMap tx, tw;
  int rc = get_tmap_nhwc_strided(x, cs.N, cs.H, cs.W, cs.Cin, p.box_cols, p.box_rows, 2, 2, &tx);
  if (rc) return rc;
  rc = get_tmap_2d_e(w, 2, cs.Cout, kpad, kpad, 64, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tw);
  if (rc) return rc;
  VAR_ENSURE_SMEM(halo_conv_fwd_kernel<false>, smem);
  const int total = cs.N * p.TPI;
  const int grid = total < cps * kNumSMs ? total : cps * kNumSMs;
  {
    LaunchScope sc(T_GEMM_FWD16, 2.0 * cs.N * cs.P * cs.Q * 64.0 * (double)(cs.R * cs.S * 64), st);
    halo_conv_fwd_kernel<false><<<grid, 224, smem, st>>>(tx, tw, p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// im2col-free forward for the stride-1 tf32 convs with 32 input channels (128-byte fp32 pixels: iTHOR imgBranch.2 / .5): ONE
// plane per tile (rows p0 - ph ... , cols -pw ... of the image, zero-filled outside), tap (r, s) starts r * WP + s rows in;
// the weights of all taps stay resident in shared memory when two CTAs per SM still fit, else they stream through the ring.
static bool conv_halo32_ok(const ConvShape& cs) {
  return env_int("VAR_HALO32", 1) && cs.Cin == 32 && (cs.Cout == 32 || cs.Cout == 64) && cs.sh == 1 && cs.sw == 1 &&
         cs.R * cs.S <= kMaxTaps && cs.R * cs.S > 1 && cs.Q + cs.S - 1 <= 128;
}

static int conv_fwd_halo_tf32(const ConvShape& cs, const float* x, const float* w, const float* bias, float* y, int relu,
                              int round_out, cudaStream_t st) {
  HaloParams p;
  memset(&p, 0, sizeof(p));
  p.WP = cs.Q + cs.S - 1;
  p.TP = 128 / p.WP;
  if (p.TP > cs.P) p.TP = cs.P;
  p.TPI = (cs.P + p.TP - 1) / p.TP;
  p.box_rows = p.TP + cs.R - 1; p.box_cols = p.WP;
  p.plane_stride = (uint32_t)(((size_t)p.box_rows * p.box_cols * 128 + 1023) / 1024 * 1024);
  p.h_start = -cs.ph; p.w_start = -cs.pw;
  p.N = cs.N; p.P = cs.P; p.Q = cs.Q;
  p.bias = bias; p.out = y; p.out_kind = 0; p.relu = relu; p.round_out = round_out;
  p.ntaps = cs.R * cs.S;
  p.nplanes = 1; p.step_h = 1; p.bn = cs.Cout; p.kelems = 32;
  for (int r = 0; r < cs.R; ++r)
    for (int s_ = 0; s_ < cs.S; ++s_) {
      const int t = r * cs.S + s_;
      p.tap_shift[t] = (uint16_t)(r * p.WP + s_);
      p.tap_wcol[t] = (uint16_t)t;
    }
  p.plane_begin[0] = 0;
  for (int pl = 1; pl <= 4; ++pl) p.plane_begin[pl] = p.ntaps;
  const int cps = env_int("VAR_HALO32_CPS", 2);
  p.slots = env_int("VAR_HALO32_SLOTS", 2);
  const size_t per_sm = 227 * 1024 - (size_t)cps * 1024;
  p.w_resident = env_int("VAR_HALO32_RESIDENT", 1);
  p.stages = p.ntaps;
  size_t smem = halo_smem_bytes(p.plane_stride, p.stages, p.slots, p.bn);
  if (!p.w_resident || smem * cps > per_sm) {
    p.w_resident = 0;
    p.stages = env_int("VAR_HALO32_STAGES", 4);
    smem = halo_smem_bytes(p.plane_stride, p.stages, p.slots, p.bn);
    while (smem * cps > per_sm && p.stages > 2) { --p.stages; smem = halo_smem_bytes(p.plane_stride, p.stages, p.slots, p.bn); }
    if (smem * cps > per_sm) return VAR_ERR_UNSUPPORTED;
  }
  const int kpad = round_up32(cs.R * cs.S * cs.Cin);
  CUtensorMap tx, tw;
  int rc = get_tmap_nhwc_strided(x, cs.N, cs.H, cs.W, cs.Cin, p.box_cols, p.box_rows, 1, 1, &tx, 4);
  if (rc) return rc;
  rc = get_tmap_2d(w, cs.Cout, kpad, kpad, cs.Cout, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tw);
  if (rc) return rc;
  VAR_ENSURE_SMEM(halo_conv_fwd_kernel<true>, smem);
  const int total = cs.N * p.TPI;
  const int grid = total < cps * kNumSMs ? total : cps * kNumSMs;
  {
    LaunchScope sc(T_GEMM_FWD, 2.0 * cs.N * cs.P * cs.Q * (double)cs.Cout * (double)(cs.R * cs.S * cs.Cin), st);
    halo_conv_fwd_kernel<true><<<grid, 224, smem, st>>>(tx, tw, p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// im2col-free data gradient (halo_conv_dgrad_kernel): one launch for the four stride-parity classes
static int conv_dgrad_halo_h16(const ConvShape& cs, const void* dy, const void* w, void* dx, int out_kind, const void* mask,
                               int mask_kind, const float* out_scale, int round_out, cudaStream_t st) {
  int drmin = 1 << 30, drmax = -(1 << 30), dsmin = 1 << 30, dsmax = -(1 << 30);
  for (int a = 0; a < 2; ++a)
    for (int r = 0; r < cs.R; ++r) {
      if ((a + cs.ph - r) & 1) continue;
      const int d = (a + cs.ph - r) / 2;  // exact
      drmin = d < drmin ? d : drmin; drmax = d > drmax ? d : drmax;
    }
  for (int b = 0; b < 2; ++b)
    for (int s_ = 0; s_ < cs.S; ++s_) {
      if ((b + cs.pw - s_) & 1) continue;
      const int d = (b + cs.pw - s_) / 2;
      dsmin = d < dsmin ? d : dsmin; dsmax = d > dsmax ? d : dsmax;
    }
  HaloDgradParams p;
  memset(&p, 0, sizeof(p));
  const int H2 = (cs.H + 1) / 2, W2 = (cs.W + 1) / 2;
  p.WP = W2 + (dsmax - dsmin);
  if (p.WP > 128) return VAR_ERR_UNSUPPORTED;
  p.TP = 128 / p.WP;
  if (p.TP > H2) p.TP = H2;
  p.TPI = (H2 + p.TP - 1) / p.TP;
  p.box_rows = p.TP + (drmax - drmin); p.box_cols = p.WP;
  p.plane_stride = (uint32_t)(((size_t)p.box_rows * p.box_cols * 128 + 1023) / 1024 * 1024);
  p.h_start = drmin; p.w_start = dsmin;
  p.N = cs.N; p.H = cs.H; p.W = cs.W;
  p.out = dx; p.mask = mask; p.out_scale = out_scale; p.out_kind = out_kind; p.mask_kind = mask_kind; p.round_out = round_out;
  int t = 0;
  for (int cls = 0; cls < 4; ++cls) {  // class = a * 2 + b
    const int a = cls >> 1, b = cls & 1;
    p.class_begin[cls] = t;
    for (int r = 0; r < cs.R; ++r)
      for (int s_ = 0; s_ < cs.S; ++s_) {
        if (((a + cs.ph - r) & 1) || ((b + cs.pw - s_) & 1)) continue;
        const int dr = (a + cs.ph - r) / 2, ds = (b + cs.pw - s_) / 2;
        p.tap_shift[t] = (uint16_t)((dr - drmin) * p.WP + (ds - dsmin));
        p.tap_wcol[t] = (uint16_t)(r * cs.S + s_);
        ++t;
      }
    if (t == p.class_begin[cls]) return VAR_ERR_UNSUPPORTED;  // every class needs a tap (its accumulator is overwritten by the first)
  }
  p.class_begin[4] = t;
  const int cps = env_int("VAR_HALO_DG_CPS", 2);  // 3 CTAs per SM: MMA side 20 % faster, but the fp32 row stores then bound the kernel (0.53 -> 0.67 ms)
  p.slots = env_int("VAR_HALO_DG_SLOTS", 2);
  p.stages = env_int("VAR_HALO_DG_STAGES", 4);  // trimmed below to what cps CTAs per SM leave room for (2 at cps = 3)
  size_t smem = halo_dgrad_smem_bytes(p.plane_stride, p.stages, p.slots);
  const size_t per_sm = 227 * 1024 - (size_t)cps * 1024;
  while (smem * cps > per_sm && p.stages > 2) { --p.stages; smem = halo_dgrad_smem_bytes(p.plane_stride, p.stages, p.slots); }
  if (smem * cps > per_sm) return VAR_ERR_UNSUPPORTED;
  const int kpad = round_up32(cs.R * cs.S * cs.Cin);
  CUtensorMap ty, tw;
  int rc = get_tmap_nhwc_strided(dy, cs.N, cs.P, cs.Q, cs.Cout, p.box_cols, p.box_rows, 1, 1, &ty);
  if (rc) return rc;
  rc = get_tmap_2d_e(w, 2, cs.Cout, kpad, kpad, 64, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tw);
  if (rc) return rc;
  VAR_ENSURE_SMEM(halo_conv_dgrad_kernel, smem);
  const int total = cs.N * p.TPI;
  const int grid = total < cps * kNumSMs ? total : cps * kNumSMs;
  {
    double flops = 0.0;  // same convention as the per-class im2col launches: input pixels of the class x its taps
    for (int cls = 0; cls < 4; ++cls)
      flops += 2.0 * cs.N * (double)((cs.H - (cls >> 1) + 1) / 2) * (double)((cs.W - (cls & 1) + 1) / 2) * 64.0 *
               (double)((p.class_begin[cls + 1] - p.class_begin[cls]) * 64);
    LaunchScope sc(T_GEMM_DGRAD16, flops, st);
    halo_conv_dgrad_kernel<<<grid, 224, smem, st>>>(ty, tw, p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// y = act(conv(x, w) + b): x f16 NHWC, w f16 packed [Cout][kpad] (same indexing as the tf32 copy, K % 64 == 0),
// y stored as fp32 (out_kind 0, optionally tf32-rounded) or f16 (out_kind 1).
int conv_fwd_h16(const ConvShape& cs, const void* x, const void* w, const float* bias, void* y, int out_kind, int relu,
                 int round_out, cudaStream_t st) {
  prof_note("fwd16 N%d H%d Ci%d Co%d R%d", cs.N, cs.H, cs.Cin, cs.Cout, cs.R);
  if (!conv_h16_ok(cs)) return VAR_ERR_UNSUPPORTED;
  if (conv_halo_ok(cs)) {
    const int rc = conv_fwd_halo_h16(cs, x, w, bias, y, out_kind, relu, round_out, st);
    if (rc != VAR_ERR_UNSUPPORTED) return rc;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  GatherGeom& g = p.g[0];
  g.src = x; g.M = cs.N * cs.P * cs.Q; g.P = cs.P; g.Q = cs.Q; g.H = cs.H; g.W = cs.W; g.C = cs.Cin;
  g.R = cs.R; g.S = cs.S; g.sh = cs.sh; g.sw = cs.sw; g.ph = cs.ph; g.pw = cs.pw;
  g.K = cs.R * cs.S * cs.Cin; g.scale = 1.f;
  const int kpad = round_up32(g.K);
  p.bn = cs.Cout;
  p.nbox = 1; p.box_rows = p.bn; p.boxbase[0] = 0;
  p.num_kb = g.K / 64; p.cpb = cs.Cin / 64;
  pick_pipeline(p.bn, &p.stages, &p.lookahead);
  fill_fwd_taps(p, cs);
  EpiParams& e = p.e[0];
  e.out = reinterpret_cast<float*>(y); e.ldo = cs.Cout; e.bias = bias; e.ncols = cs.Cout; e.relu = relu;
  e.round_out = round_out; e.out_kind = out_kind;
  CUtensorMap tm, ta;
  int rc = get_tmap_2d_e(w, 2, cs.Cout, kpad, kpad, p.bn, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm);
  if (rc) return rc;
  rc = get_tmap_im2col_e(x, 2, cs.N, cs.H, cs.W, cs.Cin, -cs.pw, -cs.ph, cs.pw - (cs.S - 1), cs.ph - (cs.R - 1), cs.sw,
                         cs.sh, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta);
  if (rc) return rc;
  return launch_gemm_persist_t<G_TMA_IM2COL, true>(tm, ta, p, (g.M + 127) / 128, 1, st);
}

// dx = conv_transpose(dy, w) [* out_scale] [* (mask > 0)]: dy f16 (scaled), w f16 packed (read transposed, MN-major),
// dx stored fp32 (out_kind 0) or f16 (1); mask = forward activation of the previous layer (mask_kind 0 fp32, 1 f16).
int conv_dgrad_h16(const ConvShape& cs, const void* dy, const void* w, void* dx, int out_kind, const void* mask,
                   int mask_kind, const float* out_scale, int round_out, cudaStream_t st) {
  prof_note("dgrad16 N%d H%d Ci%d Co%d R%d", cs.N, cs.H, cs.Cin, cs.Cout, cs.R);
  if (!conv_h16_ok(cs)) return VAR_ERR_UNSUPPORTED;
  if (conv_halo_ok(cs) && env_int("VAR_HALO_DGRAD", 1)) {
    const int rc = conv_dgrad_halo_h16(cs, dy, w, dx, out_kind, mask, mask_kind, out_scale, round_out, st);
    if (rc != VAR_ERR_UNSUPPORTED) return rc;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  const int kfwd = cs.R * cs.S * cs.Cin, kpad = round_up32(kfwd);
  p.bn = cs.Cin;
  p.b_mn_major = 1;
  p.kb_per_rs = cs.Cout / 64; p.cpb = cs.Cout / 64;
  p.cin_total = cs.Cin;
  pick_pipeline(p.bn, &p.stages, &p.lookahead);
  p.step_w = 1; p.step_h = 1;
  EpiParams& e0 = p.e[0];
  e0.out = reinterpret_cast<float*>(dx); e0.ldo = cs.Cin; e0.ncols = cs.Cin;
  e0.mask = reinterpret_cast<const float*>(mask); e0.ldm = cs.Cin; e0.mask_kind = mask_kind;
  e0.round_out = round_out; e0.out_kind = out_kind; e0.out_scale = out_scale;
  CUtensorMap tm, ta;
  int rc = get_tmap_2d_e(w, 2, cs.Cout, kpad, kpad, 64, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tm);
  if (rc) return rc;
  for (int hp = 0; hp < cs.sh; ++hp)
    for (int wp = 0; wp < cs.sw; ++wp) {
      const int H2 = (cs.H - hp + cs.sh - 1) / cs.sh, W2 = (cs.W - wp + cs.sw - 1) / cs.sw;
      if (H2 <= 0 || W2 <= 0) continue;
      const int r0 = (hp + cs.ph) % cs.sh, s0 = (wp + cs.pw) % cs.sw;
      const int J = (cs.R - r0 + cs.sh - 1) / cs.sh, I = (cs.S - s0 + cs.sw - 1) / cs.sw;
      const int a_h = (hp + cs.ph - r0) / cs.sh, a_w = (wp + cs.pw - s0) / cs.sw;
      GemmParams q = p;
      q.base_h = a_h - (J - 1); q.base_w = a_w - (I - 1);
      q.ntaps = J * I;
      for (int j = 0; j < J; ++j)
        for (int i = 0; i < I; ++i) {
          const int t = j * I + i;
          q.tap_h[t] = (uint8_t)(J - 1 - j); q.tap_w[t] = (uint8_t)(I - 1 - i);
          q.tap_id[t] = (uint8_t)((r0 + cs.sh * j) * cs.S + (s0 + cs.sw * i));
        }
      q.num_kb = q.ntaps * q.cpb;
      GatherGeom& g = q.g[0];
      g.src = dy; g.M = cs.N * H2 * W2; g.P = H2; g.Q = W2;
      g.H = cs.P; g.W = cs.Q; g.C = cs.Cout; g.R = cs.R; g.S = cs.S;
      g.sh = cs.sh; g.sw = cs.sw; g.ph = cs.ph; g.pw = cs.pw;
      g.K = q.num_kb * 64; g.scale = 1.f;
      EpiParams& e = q.e[0];
      if (cs.sh > 1 || cs.sw > 1) {
        e.map.on = 1; e.map.P2 = H2; e.map.Q2 = W2; e.map.H = cs.H; e.map.W = cs.W;
        e.map.sh = cs.sh; e.map.sw = cs.sw; e.map.oh = hp; e.map.ow = wp;
      }
      rc = get_tmap_im2col_e(dy, 2, cs.N, cs.P, cs.Q, cs.Cout, q.base_w, q.base_h, W2 - cs.Q + q.base_w,
                             H2 - cs.P + q.base_h, 1, 1, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta);
      if (rc) return rc;
      rc = launch_gemm_persist_t<G_TMA_IM2COL, true>(tm, ta, q, (g.M + 127) / 128, cs.Cin / q.bn, st);
      if (rc) return rc;
    }
  return VAR_OK;
}

// im2col-free weight gradient (halo_conv_wgrad_kernel): tap pairs of one parity plane per CTA group
static int conv_wgrad_halo_h16(const ConvShape& cs, const void* x, const void* dy, float* dw, float* db, const float* inv_scale,
                               cudaStream_t st) {
  auto fdiv2 = [](int v) { return v >= 0 ? v / 2 : -((-v + 1) / 2); };
  int rjmin = 1 << 30, rjmax = -(1 << 30), sjmin = 1 << 30, sjmax = -(1 << 30);
  for (int r = 0; r < cs.R; ++r) { const int j = fdiv2(r - cs.ph); rjmin = j < rjmin ? j : rjmin; rjmax = j > rjmax ? j : rjmax; }
  for (int s_ = 0; s_ < cs.S; ++s_) { const int j = fdiv2(s_ - cs.pw); sjmin = j < sjmin ? j : sjmin; sjmax = j > sjmax ? j : sjmax; }
  HaloWgradParams p;
  memset(&p, 0, sizeof(p));
  p.WP = cs.Q + (sjmax - sjmin);
  if (p.WP > 128) return VAR_ERR_UNSUPPORTED;
  p.TP = env_int("VAR_HALO_WG_TP", 8);
  if (p.TP > cs.P) p.TP = cs.P;
  p.TPI = (cs.P + p.TP - 1) / p.TP;
  p.nk = (p.TP * p.WP + 15) / 16;
  p.box_rows = p.TP + (rjmax - rjmin); p.box_cols = p.WP;
  if (p.box_rows > 128 || p.TP * 2 > 256) return VAR_ERR_UNSUPPORTED;
  const int shift_max = (rjmax - rjmin) * p.WP + (sjmax - sjmin);
  int plane_rows = p.box_rows * p.box_cols;
  if (plane_rows < shift_max + 16 * p.nk) plane_rows = shift_max + 16 * p.nk;
  p.plane_bytes = (uint32_t)(p.box_rows * p.box_cols * 128);
  p.dy_bytes = (uint32_t)(p.TP * p.WP * 128);
  p.plane_region = (uint32_t)(((size_t)plane_rows * 128 + 1023) / 1024 * 1024);
  p.slot_stride = p.plane_region + (uint32_t)(((size_t)p.nk * 16 * 128 + 1023) / 1024 * 1024);
  p.h_start = 2 * rjmin; p.w_start = 2 * sjmin;
  p.N = cs.N; p.P = cs.P; p.Q = cs.Q;
  p.dw = dw; p.db = db; p.inv_scale = inv_scale; p.kpad = round_up32(cs.R * cs.S * cs.Cin);
  int mp = env_int("VAR_HALO_WG_PAIRS", 4);  // pairs per CTA: 64 TMEM columns each
  if (mp < 1) mp = 1;
  if (mp > 8) mp = 8;
  // pairs of each parity plane, in order of their start rows
  int npairs = 0, ng = 0;
  for (int pl = 0; pl < 4; ++pl) {
    int sh[kMaxTaps], id[kMaxTaps], n = 0;
    for (int r = 0; r < cs.R; ++r)
      for (int s_ = 0; s_ < cs.S; ++s_) {
        const int rj = fdiv2(r - cs.ph), sj = fdiv2(s_ - cs.pw);
        if (((r - cs.ph) - 2 * rj) * 2 + ((s_ - cs.pw) - 2 * sj) != pl) continue;
        sh[n] = (rj - rjmin) * p.WP + (sj - sjmin); id[n] = r * cs.S + s_; ++n;
      }
    for (int i = 1; i < n; ++i)  // insertion sort by start row
      for (int j = i; j > 0 && sh[j] < sh[j - 1]; --j) { int t = sh[j]; sh[j] = sh[j - 1]; sh[j - 1] = t; t = id[j]; id[j] = id[j - 1]; id[j - 1] = t; }
    const int np = (n + 1) / 2;
    if (np == 0) continue;
    if (npairs + np > 32) return VAR_ERR_UNSUPPORTED;
    const int first = npairs;
    for (int i = 0; i < n; i += 2, ++npairs) {
      p.pair_shift[npairs] = (uint16_t)sh[i]; p.pair_tap0[npairs] = (uint8_t)id[i];
      if (i + 1 < n) { p.pair_lbo[npairs] = (uint16_t)(sh[i + 1] - sh[i]); p.pair_tap1[npairs] = (uint8_t)id[i + 1]; }
      else { p.pair_lbo[npairs] = 0; p.pair_tap1[npairs] = 255; }
    }
    const int ngr = (np + mp - 1) / mp;
    if (ng + ngr > 16) return VAR_ERR_UNSUPPORTED;
    for (int gi = 0; gi < ngr; ++gi, ++ng) {
      p.g_plane[ng] = (uint8_t)pl;
      p.g_pair_begin[ng] = (uint8_t)(first + (np * gi) / ngr);
    }
  }
  p.g_pair_begin[ng] = (uint8_t)npairs;
  p.ngroups = ng;
  p.slots = env_int("VAR_HALO_WG_SLOTS", 2);
  const size_t smem = (size_t)p.slots * p.slot_stride + 1024 + 256;
  if (smem > 227 * 1024 - 1024) return VAR_ERR_UNSUPPORTED;
  int maxp = 0;
  for (int g = 0; g < ng; ++g) { const int c = p.g_pair_begin[g + 1] - p.g_pair_begin[g]; maxp = c > maxp ? c : maxp; }
  const int tm_cols = maxp <= 1 ? 64 : (maxp <= 2 ? 128 : (maxp <= 4 ? 256 : 512));
  int cps = (int)((227 * 1024) / (smem + 1024));
  if (cps > 512 / tm_cols) cps = 512 / tm_cols;  // co-resident CTAs must all get their TMEM columns
  if (cps > 4) cps = 4;
  if (cps < 1) return VAR_ERR_UNSUPPORTED;
  // CTAs per group in proportion to its MMA count + the fixed per-tile staging cost (~0.8 of a pair)
  const int total_tiles = cs.N * p.TPI;
  const int want = cps * kNumSMs;
  double wsum = 0.0;
  for (int g = 0; g < ng; ++g) wsum += (p.g_pair_begin[g + 1] - p.g_pair_begin[g]) + 0.8;
  int acc = 0;
  for (int g = 0; g < ng; ++g) {
    p.g_cta_begin[g] = acc;
    int c = (int)(want * ((p.g_pair_begin[g + 1] - p.g_pair_begin[g]) + 0.8) / wsum);
    if (c < 1) c = 1;
    if (c > total_tiles) c = total_tiles;
    acc += c;
  }
  p.g_cta_begin[ng] = acc;
  CUtensorMap tx, tdy;
  int rc = get_tmap_nhwc_strided(x, cs.N, cs.H, cs.W, cs.Cin, p.box_cols, p.box_rows, 2, 2, &tx);
  if (rc) return rc;
  rc = get_tmap_nhwc_strided(dy, cs.N, cs.P, cs.Q, cs.Cout, p.WP, p.TP, 1, 1, &tdy);
  if (rc) return rc;
  VAR_ENSURE_SMEM(halo_conv_wgrad_kernel, smem);
  {
    LaunchScope sc(T_WGRAD16, 2.0 * cs.N * cs.P * cs.Q * (double)cs.Cout * (double)(cs.R * cs.S * cs.Cin), st);
    halo_conv_wgrad_kernel<<<acc, 192, smem, st>>>(tx, tdy, p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

// Fallback bias gradient for f16 dY when the weight-gradient kernel has no free row group: db += inv * colsum(dy).
__global__ void __launch_bounds__(256)
colsum_f16_kernel(const uint16_t* __restrict__ dy, long long M, int C, float* __restrict__ db,
                  const float* __restrict__ inv_scale, int rows_per_cta) {
  const float inv = inv_scale ? __ldg(inv_scale) : 1.f;
  const long long r0 = (long long)blockIdx.x * rows_per_cta, r1 = min(r0 + (long long)rows_per_cta, M);
  for (int c = threadIdx.x; c < C; c += 256) {
    float acc = 0.f;
    for (long long r = r0; r < r1; ++r) acc += f16_bits_to_f32(dy[r * C + c]);
    atomicAdd(db + c, acc * inv);
  }
}

// dw[Cout][kpad] += inv_scale * im2col(x)^T dy ; db[Cout] += inv_scale * colsum(dy): x f16 NHWC, dy f16 (scaled).
// The bias gradient comes out of the same kernel when the last k tile has a free 64-row group (K % 128 == 64);
// otherwise *db_done = 0 and the caller must produce it.
int conv_wgrad_h16(const ConvShape& cs, const void* x, const void* dy, float* dw, float* db, const float* inv_scale,
                   int* db_done, cudaStream_t st) {
  prof_note("wgrad16 N%d H%d Ci%d Co%d R%d", cs.N, cs.H, cs.Cin, cs.Cout, cs.R);
  if (!conv_h16_ok(cs)) return VAR_ERR_UNSUPPORTED;
  if (conv_halo_ok(cs) && env_int("VAR_HALO_WGRAD", 0)) {
    const int rc = conv_wgrad_halo_h16(cs, x, dy, dw, db, inv_scale, st);
    if (rc != VAR_ERR_UNSUPPORTED) {
      if (db_done) *db_done = db != nullptr;
      return rc;
    }
  }
  WgradH16Params p;
  memset(&p, 0, sizeof(p));
  p.M = cs.N * cs.P * cs.Q;
  p.K = cs.R * cs.S * cs.Cin;
  p.kpad = round_up32(p.K);
  p.cout = cs.Cout;
  p.dw = dw; p.inv_scale = inv_scale;
  p.pb = env_int("VAR_WGRAD16_PB", 64);       // pixels per TMA box / pipeline stage (64 x 4 stages: 2 CTAs per SM, measured best)
  p.stages = env_int("VAR_WGRAD16_STAGES", 4);
  if (p.pb < 16 || p.pb > 256 || p.pb % 16) return VAR_ERR_ARG;
  p.P = cs.P; p.Q = cs.Q; p.cpb = cs.Cin / 64;
  p.base_w = -cs.pw; p.base_h = -cs.ph; p.step_w = cs.sw; p.step_h = cs.sh;
  for (int r = 0; r < cs.R; ++r)
    for (int s_ = 0; s_ < cs.S; ++s_) { p.tap_w[r * cs.S + s_] = (uint8_t)s_; p.tap_h[r * cs.S + s_] = (uint8_t)r; }
  const int ktiles = (p.K + 127) / 128;
  p.ones_ktile = (p.K % 128 == 64) ? ktiles - 1 : -1;
  p.db = (db && p.ones_ktile >= 0) ? db : nullptr;
  if (db_done) *db_done = p.db != nullptr;
  CUtensorMap tx, tdy;
  int rc = get_tmap_im2col_e(x, 2, cs.N, cs.H, cs.W, cs.Cin, -cs.pw, -cs.ph, cs.pw - (cs.S - 1), cs.ph - (cs.R - 1), cs.sw,
                             cs.sh, p.pb, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tx);
  if (rc) return rc;
  rc = get_tmap_2d_e(dy, 2, p.M, cs.Cout, cs.Cout, p.pb, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tdy);
  if (rc) return rc;
  size_t smem = wgrad_h16_smem_bytes(cs.Cout, p.stages, p.pb);
  while (smem > 227 * 1024 - 2048 && p.stages > 2) { --p.stages; smem = wgrad_h16_smem_bytes(cs.Cout, p.stages, p.pb); }
  const int per_sm = smem * 3 + 6144 <= 227 * 1024 ? 3 : (smem * 2 + 4096 <= 227 * 1024 ? 2 : 1);
  // one wave of resident CTAs (waves == 2 for small layers would only add partial-sum atomics)
  int splits = (per_sm * kNumSMs) / ktiles;
  if (splits < 1) splits = 1;
  int ppc = (p.M + splits - 1) / splits;
  ppc = ((ppc + p.pb - 1) / p.pb) * p.pb;
  if (ppc < 2 * p.pb) ppc = 2 * p.pb;
  splits = (p.M + ppc - 1) / ppc;
  p.pix_per_cta = ppc;
  VAR_ENSURE_SMEM(tc_wgrad_h16_kernel, smem);
  dim3 grid(ktiles, splits, 1);
  {
    LaunchScope sc(T_WGRAD16, 2.0 * p.M * (double)cs.Cout * p.K, st);
    tc_wgrad_h16_kernel<<<grid, 160, smem, st>>>(tx, tdy, p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  if (db && !p.db) {  // no free row group in the last k tile: separate (slow, rarely needed) column sum
    const int rp = 512;
    LaunchScope sc(T_COLSUM, 0, st);
    colsum_f16_kernel<<<(unsigned)((p.M + rp - 1) / rp), 256, 0, st>>>(reinterpret_cast<const uint16_t*>(dy), p.M, cs.Cout, db,
                                                                       inv_scale, rp);
    VAR_CUDA_CHECK(cudaGetLastError());
    if (db_done) *db_done = 1;
  }
  return VAR_OK;
}

// ---- plain [M, K] x [N, K]^T GEMMs on f16 operands (GRU input projection and its backward) ----
// out[M, N] (fp32, row pitch ldo) = A (f16, [M, K], pitch lda) x W^T (f16, [N, K] K-major) (+ bias) (* *out_scale)
// (* (mask > 0)); K % 64 == 0, N split into column tiles of the largest divisor <= 256 that is a multiple of 32.
int linear_h16(int M, int K, int N, const void* a, long long lda, const void* w, const float* bias, float* out,
               long long ldo, const void* mask, int mask_kind, long long ldm, const float* out_scale, int round_out,
               int is_dgrad, cudaStream_t st) {
  prof_note(is_dgrad ? "lin_dgrad16 M%d K%d N%d %d%d" : "lin_fwd16 M%d K%d N%d %d%d", M, K, N, 0, 0);
  if (!env_int("VAR_H16", 1) || gather_mode() != 1 || K % 64 || M <= 0) return VAR_ERR_UNSUPPORTED;
  const int bn = pick_bn(N);
  if (bn == 0 || bn % 32) return VAR_ERR_UNSUPPORTED;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.g[0].M = M; p.g[0].K = K; p.g[0].scale = 1.f;
  p.bn = bn; p.nbox = 1; p.box_rows = bn; p.boxbase[0] = 0;
  p.num_kb = K / 64;
  p.prof_dgrad = is_dgrad;
  pick_pipeline(bn, &p.stages, &p.lookahead);
  EpiParams& e = p.e[0];
  e.out = out; e.ldo = ldo; e.bias = bias; e.ncols = N; e.round_out = round_out;
  e.mask = reinterpret_cast<const float*>(mask); e.ldm = ldm; e.mask_kind = mask_kind; e.out_scale = out_scale;
  CUtensorMap tb, ta;
  int rc = get_tmap_2d_e(w, 2, N, K, K, bn, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tb);
  if (rc) return rc;
  rc = get_tmap_2d_e(a, 2, M, K, lda, 128, (int)CU_TENSOR_MAP_SWIZZLE_128B, &ta);
  if (rc) return rc;
  return launch_gemm_persist_t<G_TMA_TILED, true>(tb, ta, p, (M + 127) / 128, N / bn, st);
}

// dw[N][kpad] += *inv_scale * X^T dY over M rows: X f16 [M, K] (pitch ldx), dY f16 [M, N] (pitch ldy, scaled);
// output columns are processed in slabs of <= 256 (grid.z), K in tiles of 128 (grid.x), rows split over grid.y.
int linear_wgrad_h16(int M, int K, int N, const void* x, long long ldx, const void* dy, long long ldy, float* dw,
                     int kpad, const float* inv_scale, cudaStream_t st) {
  prof_note("lin_wgrad16 M%d K%d N%d %d%d", M, K, N, 0, 0);
  if (!env_int("VAR_H16", 1) || gather_mode() != 1 || K % 64 || M <= 0) return VAR_ERR_UNSUPPORTED;
  const int slab = N <= 256 ? N : 256;
  if (N % slab || slab % 64) return VAR_ERR_UNSUPPORTED;
  WgradH16Params p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.K = K; p.kpad = kpad; p.cout = slab;
  p.dw = dw; p.inv_scale = inv_scale;
  p.pb = env_int("VAR_WGRAD16_PB", 64);
  p.stages = env_int("VAR_WGRAD16_STAGES", 4);
  if (p.pb < 16 || p.pb > 256 || p.pb % 16) return VAR_ERR_ARG;
  p.a_tiled = 1;
  p.P = 1; p.Q = 1; p.cpb = K / 64;
  p.ones_ktile = -1;
  const int ktiles = (K + 127) / 128, nslab = N / slab;
  CUtensorMap tx, tdy;
  int rc = get_tmap_2d_e(x, 2, M, K, ldx, p.pb, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tx);
  if (rc) return rc;
  rc = get_tmap_2d_e(dy, 2, M, N, ldy, p.pb, (int)CU_TENSOR_MAP_SWIZZLE_128B, &tdy);
  if (rc) return rc;
  size_t smem = wgrad_h16_smem_bytes(slab, p.stages, p.pb);
  while (smem > 227 * 1024 - 2048 && p.stages > 2) { --p.stages; smem = wgrad_h16_smem_bytes(slab, p.stages, p.pb); }
  if (smem > 227 * 1024 - 2048) return VAR_ERR_UNSUPPORTED;
  const int per_sm = smem * 3 + 6144 <= 227 * 1024 ? 3 : (smem * 2 + 4096 <= 227 * 1024 ? 2 : 1);
  int splits = (per_sm * kNumSMs) / (ktiles * nslab);
  if (splits < 1) splits = 1;
  int ppc = (M + splits - 1) / splits;
  ppc = ((ppc + p.pb - 1) / p.pb) * p.pb;
  if (ppc < 2 * p.pb) ppc = 2 * p.pb;
  splits = (M + ppc - 1) / ppc;
  p.pix_per_cta = ppc;
  VAR_ENSURE_SMEM(tc_wgrad_h16_kernel, smem);
  dim3 grid(ktiles, splits, nslab);
  {
    LaunchScope sc(T_WGRAD16, 2.0 * M * (double)N * K, st);
    tc_wgrad_h16_kernel<<<grid, 160, smem, st>>>(tx, tdy, p);
  }
  VAR_CUDA_CHECK(cudaGetLastError());
  return VAR_OK;
}

}  // namespace var

extern "C" {
long long var_launch_count(void) {
  auto& p = var::prof();
  std::lock_guard<std::mutex> lk(p.mu);
  return p.launches;
}
int var_launch_count_add(long long n) {
  if (n < 0) return VAR_ERR_ARG;
  auto& p = var::prof();
  std::lock_guard<std::mutex> lk(p.mu);
  p.launches += n;
  return VAR_OK;
}
int var_prof_begin(void) {
  auto& p = var::prof();
  std::lock_guard<std::mutex> lk(p.mu);
  for (auto& r : p.recs) { p.pool.push_back(r.e0); p.pool.push_back(r.e1); }
  p.recs.clear();
  p.on = true;
  return VAR_OK;
}
// Stops profiling, waits for the device and accumulates per-tag totals (arrays of T_NUM_TAGS).
int var_prof_end(double* ms, double* flops, long long* count, int ntags) {
  auto& p = var::prof();
  std::lock_guard<std::mutex> lk(p.mu);
  p.on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return VAR_ERR_CUDA;
  for (int i = 0; i < ntags; ++i) { ms[i] = 0; flops[i] = 0; count[i] = 0; }
  for (auto& r : p.recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    if (r.tag < ntags) { ms[r.tag] += t; flops[r.tag] += r.flops; ++count[r.tag]; }
    p.pool.push_back(r.e0); p.pool.push_back(r.e1);
  }
  p.recs.clear();
  return VAR_OK;
}
int var_prof_num_tags(void) { return var::T_NUM_TAGS; }
// bit 0: the 16-bit conv region is enabled, bit 1: the recurrent kernels run f16 operands
int var_h16_flags(void) {
  var::ConvShape probe{1, 8, 8, 64, 64, 3, 3, 1, 1, 1, 1, 8, 8};
  return (var::conv_h16_ok(probe) ? 1 : 0) | (var::gru_h16_enabled() ? 2 : 0);
}
// Like var_prof_end but writes one CSV line per launch (tag, ms, flops, note) to `path`.
int var_prof_dump(const char* path) {
  auto& p = var::prof();
  std::lock_guard<std::mutex> lk(p.mu);
  p.on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return VAR_ERR_CUDA;
  FILE* f = fopen(path, "w");
  if (!f) return VAR_ERR_ARG;
  fprintf(f, "tag,ms,flops,note\n");
  for (auto& r : p.recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    fprintf(f, "%d,%.5f,%.0f,%s\n", r.tag, t, r.flops, r.note);
    p.pool.push_back(r.e0); p.pool.push_back(r.e1);
  }
  p.recs.clear();
  fclose(f);
  return VAR_OK;
}
}
