// Common device helpers for the VAR hot path on sm_100a: PTX wrappers for
// mbarrier, cp.async, TMA (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM), and
// tf32 rounding.  Everything here is hand-written inline PTX; no CUTLASS.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>

#define VAR_OK 0
#define VAR_ERR_ARG -1
#define VAR_ERR_CUDA -2
#define VAR_ERR_UNSUPPORTED -3
#define VAR_ERR_WORKSPACE -4

#define VAR_CUDA_CHECK(x)                                   \
  do {                                                      \
    cudaError_t _e = (x);                                   \
    if (_e != cudaSuccess) {                                \
      var_set_last_error(cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VAR_ERR_CUDA;                                  \
    }                                                       \
  } while (0)

void var_set_last_error(const char* msg, const char* file, int line);
// Raises a kernel's dynamic shared memory limit to at least `bytes` on the CURRENT device (remembered
// per (kernel, device), so repeated calls cost a hash lookup).  Returns a cudaError_t.
cudaError_t var_ensure_dyn_smem(const void* kernel, size_t bytes);
#define VAR_ENSURE_SMEM(kernel, bytes) VAR_CUDA_CHECK(var_ensure_dyn_smem((const void*)(kernel), (size_t)(bytes)))

namespace var {

constexpr int kNumSMs = 148;

// Launch accounting: every kernel launch of the library goes through a LaunchScope.  It
// counts launches (var_launch_count) and, while profiling is on (var_prof_begin/end),
// brackets the launch with CUDA events on the launching stream.
enum LaunchTag : int {
  T_GEMM_FWD = 0, T_GEMM_DGRAD, T_GEMM_SCALAR, T_GRU_STEP, T_WGRAD, T_COLSUM, T_MFCC, T_TAIL,
  T_POOL, T_ADAM, T_GRU_CELL_BWD, T_SAMPLER, T_MISC,
  T_GEMM_FWD16, T_GEMM_DGRAD16, T_WGRAD16,  // the 16-bit operand (kind::f16) conv kernels
  T_GRU_BWD,                                // the persistent BPTT kernels (T_GRU_STEP: forward + per-step fallbacks)
  T_NUM_TAGS
};
struct LaunchScope {
  LaunchScope(int tag, double flops, cudaStream_t st);
  ~LaunchScope();
  int idx;
  cudaStream_t st;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- tf32 rounding (round-to-nearest, ties away): values stored in HBM are
// pre-rounded so that the tensor core's truncating read is exact.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- mbarrier -------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// One lane of the (converged) warp: lets the compiler keep the surrounding code on the uniform datapath -- descriptors and
// addresses of tcgen05 / TMA instructions stay in uniform registers.  A plain `lane == 0` branch makes the region divergent:
// every operand then goes through R2UR and every UTCHMMA is wrapped in an ELECT / BRA.U.ANY loop (~100 issue cycles per MMA).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xFFFFFFFF;\n\tselp.b32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- proxies / cp.async ---------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// 16-byte async copy global->shared; src_bytes in {0,16}: 0 zero-fills.
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t dst, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ float4 ld_shared_v4(uint32_t src) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(src));
  return v;
}

// ---- TMA ------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: box lands in shared memory, completion bytes on the mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// 3-D tiled load (coordinates innermost first); out-of-tensor elements are zero filled.
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 4-D im2col load (NHWC tensor map from cuTensorMapEncodeIm2col): `pixelsPerColumn` base
// pixels starting at (w, h, n), walked in W-then-H-then-N order inside the map's bounding
// box with the map's traversal strides; each loads `channelsPerPixel` channels from c of the
// pixel at base + (woff, hoff).  Out-of-tensor pixels are zero filled (= conv padding).
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar,
                                                   int c, int w, int h, int n, uint16_t woff,
                                                   uint16_t hoff) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(woff),
      "h"(hoff)
      : "memory");
}

// ---- programmatic dependent launch --------------------------------------------
// launch_dependents: lets the next kernel of the stream (if it was launched with the
// programmatic-stream-serialization attribute) start its prologue while this grid runs;
// wait: blocks until the prerequisite grid has completed and its writes are visible.
// Both are no-ops for ordinary launches.
__device__ __forceinline__ void griddep_launch() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void griddep_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ---- tcgen05 / TMEM ---------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// Whole-warp: allocate ncols (power of two >= 32) TMEM columns; base address
// is written to *smem_dst.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, f32 accumulate.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A * B, 16-bit inputs (f16 / bf16 per the instruction descriptor), f32 accumulate, K = 16.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread retire.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- UMMA descriptors -------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1.
//  K-major : rows of 128 B (32 tf32), 8-row groups SBO bytes apart; LBO unused.
//  MN-major (tf32 => SWIZZLE_128B_BASE32B, layout_type 1): 128 B of MN per
//            k-row, 4 k-rows per 512 B atom, 32-byte granules XOR-ed with
//            (k-row % 4); LBO = byte distance between consecutive 32-element
//            MN groups, SBO = distance between consecutive 4-k groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes,
                                                   uint32_t sbo_bytes, uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)(layout_type & 7) << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// Instruction descriptor for kind::tf32, f32 accumulate, M=128.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (2u << 7)                     // a_format = TF32
         | (2u << 10)                    // b_format = TF32
         | ((uint32_t)a_mn_major << 15)  // a_major
         | ((uint32_t)b_mn_major << 16)  // b_major
         | ((uint32_t)(n >> 3) << 17)    // N
         | ((uint32_t)(128 >> 4) << 24); // M = 128
}

// Instruction descriptor for kind::f16 (a_fmt / b_fmt: 0 = f16, 1 = bf16), f32 accumulate, M=128.
__host__ __device__ constexpr uint32_t make_idesc_h16(int n, int a_fmt, int b_fmt, int a_mn_major, int b_mn_major) {
  return (1u << 4) | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// fp32 -> packed 16-bit pairs (round to nearest even)
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float f16_bits_to_f32(uint16_t h) {
  float f;
  asm("{\n\t.reg .b16 t;\n\tmov.b16 t, %1;\n\tcvt.f32.f16 %0, t;\n\t}" : "=f"(f) : "h"(h));
  return f;
}

// TMA store of one 2-D box from shared memory (bulk async group; the box is clipped at the tensor bounds)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src_smem),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {  // at most N groups still READING their shared source
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Byte offset of 16-byte chunk `chunk` (0..7) in row `row` of a 128B-swizzled
// tile whose rows are 128 bytes and whose base is 1024-byte aligned.
__device__ __forceinline__ uint32_t swz128(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

// Same for the 32-byte-granule swizzle (SWIZZLE_128B_BASE32B / TMA
// SWIZZLE_128B_ATOM_32B): granule (chunk>>1) XOR (row % 4), low chunk bit kept.
__device__ __forceinline__ uint32_t swz128_32(uint32_t row, uint32_t chunk) {
  return row * 128u + (((((chunk >> 1) ^ (row & 3u)) << 1) | (chunk & 1u)) << 4);
}

}  // namespace var
