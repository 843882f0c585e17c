// Hardware probe for an im2col-free ("halo") convolution operand (run on the B200 box: ./halo_probe):
//   P1  a K-major, 128B-swizzled f16 A operand whose start address is shifted by k rows (k not a multiple of 8)
//       inside one large swizzled shared-memory image: does tcgen05.mma read rows k .. k+127, and does the
//       descriptor's matrix-base-offset field (bits 49-51) have to carry (start >> 7) & 7?
//   P2  a 4-D TILED tensor map over an NHWC f16 tensor with element strides (1, 2, 2, 1) and a box that starts at
//       negative coordinates and reaches past the tensor: does the load work, zero-fill, and land as dense
//       [h][w][64 ch] 128-byte rows in the 128B swizzle?
// Together they would let every tap of a strided convolution read its A operand as a shifted window of ONE staged
// input tile (parity planes), instead of re-fetching an im2col tile per tap.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>
#include <vector>

#include "common.cuh"

using namespace var;

// ------------------------------------------------------------------------------------------------ P1
constexpr int kRows = 160, kN = 64, kK = 64;

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t sbo, uint32_t bo) {
  uint64_t d = make_smem_desc(saddr, 16u, sbo, 2);
  d |= (uint64_t)(bo & 7u) << 49;
  return d;
}

__global__ void __launch_bounds__(128) p1_kernel(const uint16_t* A, const uint16_t* B, float* D, int shift, int use_bo) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  uint8_t* sA = sm;                    // kRows x 128 B
  uint8_t* sB = sm + kRows * 128;      // 64 x 128 B (kRows * 128 is a multiple of 1024)
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  for (int i = tid; i < kRows * kK; i += 128) {
    const int r = i / kK, k = i % kK;
    const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
    *reinterpret_cast<uint16_t*>(sA + off) = A[i];
  }
  for (int i = tid; i < kN * kK; i += 128) {
    const int n = i / kK, k = i % kK;
    const uint32_t off = (uint32_t)n * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
    *reinterpret_cast<uint16_t*>(sB + off) = B[i];
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(smem_u32(&tslot), 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_h16(kN, 0, 0, 0, 0);
    const uint32_t a0 = smem_u32(sA) + (uint32_t)shift * 128u;
    for (int j = 0; j < kK / 16; ++j) {
      const uint32_t aj = a0 + (uint32_t)j * 32u;
      const uint64_t ad = desc_bo(aj, 1024u, use_bo ? (a0 >> 7) : 0u);
      const uint64_t bd = make_smem_desc(smem_u32(sB) + (uint32_t)j * 32u, 16u, 1024u, 2);
      umma_f16(tmem, ad, bd, idesc, j != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < kN; c += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * kN + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

static uint16_t f2h(float x) { __half h = __float2half(x); return *reinterpret_cast<uint16_t*>(&h); }

static int run_p1(int shift, int use_bo) {
  std::vector<float> a(kRows * kK), b(kN * kK);
  std::vector<uint16_t> ha(kRows * kK), hb(kN * kK);
  uint32_t s = 11;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 10) % 9) - 4; };
  for (size_t i = 0; i < a.size(); ++i) { a[i] = (float)rnd(); ha[i] = f2h(a[i]); }
  for (size_t i = 0; i < b.size(); ++i) { b[i] = (float)rnd(); hb[i] = f2h(b[i]); }
  uint16_t *dA, *dB; float* dD;
  cudaMalloc(&dA, ha.size() * 2); cudaMalloc(&dB, hb.size() * 2); cudaMalloc(&dD, 128 * kN * 4);
  cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * kN * 4);
  const int smem = kRows * 128 + kN * 128 + 2048;
  cudaFuncSetAttribute(p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  p1_kernel<<<1, 128, smem>>>(dA, dB, dD, shift, use_bo);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<float> d(128 * kN);
  cudaMemcpy(d.data(), dD, d.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  int bad_rows = 0;
  for (int m = 0; m < 128; ++m) {
    double wr = 0;
    for (int n = 0; n < kN; ++n) {
      float ref = 0;
      for (int k = 0; k < kK; ++k) ref += a[(m + shift) * kK + k] * b[n * kK + k];
      wr = fmax(wr, fabs((double)d[m * kN + n] - ref));
    }
    if (wr != 0) ++bad_rows;
    worst = fmax(worst, wr);
  }
  printf("P1 shift=%3d base_offset=%s cuda=%s max|diff|=%g bad_rows=%d  %s\n", shift, use_bo ? "(a>>7)&7" : "0",
         cudaGetErrorString(e), worst, bad_rows, (e == cudaSuccess && worst == 0) ? "OK" : "FAIL");
  if (e != cudaSuccess) cudaDeviceReset();
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return (e == cudaSuccess && worst == 0) ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ P2
constexpr int pN = 3, pH = 24, pW = 20, pC = 64;     // NHWC f16 tensor
constexpr int bW = 13, bH = 8;                       // box: 64 ch x 13 cols x 8 rows x 1 image, element strides 2

__global__ void __launch_bounds__(128) p2_kernel(const __grid_constant__ CUtensorMap tm, uint16_t* out, int w0, int h0, int n0) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  __shared__ uint64_t bar;
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
  for (int i = tid; i < bW * bH * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x7C007C00u;  // +inf halves: "not written"
  fence_proxy_async_smem();
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(smem_u32(&bar), (uint32_t)(bW * bH * 128));
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
            smem_u32(sm)),
        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(smem_u32(&bar)), "r"(0), "r"(w0), "r"(h0), "r"(n0)
        : "memory");
  }
  {  // bounded wait: a byte-count mismatch must not hang the box
    long long spins = 0;
    while (!mbar_try_wait(smem_u32(&bar), 0) && ++spins < 2000000) {
    }
    if (spins >= 2000000) { if (tid == 0) printf("P2: TMA never completed the expected %d bytes\n", bW * bH * 128); return; }
  }
  // un-swizzle into a dense [bH][bW][64] image
  for (int i = tid; i < bW * bH * 64; i += 128) {
    const int c = i % 64, row = i / 64;
    const uint32_t off = (uint32_t)row * 128u + ((((uint32_t)c >> 3) ^ ((uint32_t)row & 7u)) << 4) + ((uint32_t)c & 7u) * 2u;
    out[i] = *reinterpret_cast<uint16_t*>(sm + off);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int run_p2(int w0, int h0) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) { printf("P2 no encode fn\n"); return 1; }
  EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(p);
  std::vector<uint16_t> hx((size_t)pN * pH * pW * pC);
  for (int n = 0; n < pN; ++n)
    for (int h = 0; h < pH; ++h)
      for (int w = 0; w < pW; ++w)
        for (int c = 0; c < pC; ++c) hx[(((size_t)n * pH + h) * pW + w) * pC + c] = f2h((float)(1 + ((n * 31 + h * 7 + w * 3 + c) % 200)));
  uint16_t *dx, *dout;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dout, (size_t)bW * bH * 64 * 2);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cuuint64_t gdim[4] = {pC, pW, pH, pN};
  cuuint64_t gstr[3] = {(cuuint64_t)pC * 2, (cuuint64_t)pW * pC * 2, (cuuint64_t)pH * pW * pC * 2};
  // boxDim counts TENSOR elements spanned; ceil(boxDim / elementStride) elements land in shared memory
  cuuint32_t box[4] = {64, 2 * bW, 2 * bH, 1};
  cuuint32_t estr[4] = {1, 2, 2, 1};
  CUtensorMap tm;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, dx, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("P2 encode failed: %d\n", (int)r); return 1; }
  const int n0 = 1;
  const int smem = bW * bH * 128 + 2048;
  cudaFuncSetAttribute(p2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  p2_kernel<<<1, 128, smem>>>(tm, dout, w0, h0, n0);
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<uint16_t> ho((size_t)bW * bH * 64);
  cudaMemcpy(ho.data(), dout, ho.size() * 2, cudaMemcpyDeviceToHost);
  // hypotheses for what box element (hh, ww) holds: A: x[h0 + 2 hh][w0 + 2 ww] (stride applied to the traversal),
  // B: x[h0 + hh][w0 + ww] restricted to every 2nd (box shrinks)
  int badA = 0;
  for (int hh = 0; hh < bH; ++hh)
    for (int ww = 0; ww < bW; ++ww)
      for (int c = 0; c < 64; ++c) {
        const int h = h0 + 2 * hh, w = w0 + 2 * ww;
        const uint16_t exp = (h >= 0 && h < pH && w >= 0 && w < pW) ? hx[(((size_t)n0 * pH + h) * pW + w) * pC + c] : 0;
        if (ho[((size_t)hh * bW + ww) * 64 + c] != exp) ++badA;
      }
  printf("P2 w0=%d h0=%d cuda=%s  mismatches vs x[h0+2hh][w0+2ww] (dense [hh][ww] rows): %d  %s\n", w0, h0, cudaGetErrorString(e),
         badA, (e == cudaSuccess && badA == 0) ? "OK" : "FAIL");
  if (badA && e == cudaSuccess) {
    for (int hh = 0; hh < 2; ++hh) {
      printf("   row hh=%d first channel of each ww:", hh);
      for (int ww = 0; ww < bW; ++ww) printf(" %g", __half2float(*reinterpret_cast<__half*>(&ho[((size_t)hh * bW + ww) * 64])));
      printf("\n");
    }
  }
  if (e != cudaSuccess) cudaDeviceReset();
  cudaFree(dx); cudaFree(dout);
  return (e == cudaSuccess && badA == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "p1";
  if (which[1] == '1') {
    const int shift = argc > 2 ? atoi(argv[2]) : 0, bo = argc > 3 ? atoi(argv[3]) : 0;
    return run_p1(shift, bo);
  }
  const int w0 = argc > 2 ? atoi(argv[2]) : 0, h0 = argc > 3 ? atoi(argv[3]) : 0;
  return run_p2(w0, h0);
}
