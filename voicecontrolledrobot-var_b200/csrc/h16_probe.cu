// Hardware probe for the 16-bit operand path (run on the B200 box: ./h16_probe).
// One CTA, D[128 x 64] = A[128 x 64] * B[64 x 64]^T-style products through tcgen05.mma kind::f16 with
//   case 0: A f16  K-major, B f16  K-major            (forward convs / Linear)
//   case 1: A f16  K-major, B bf16 K-major            (mixed operand formats in one MMA)
//   case 2: A bf16 K-major, B f16  MN-major           (dgrad: dY x W^T with the packed forward weights)
//   case 3: A f16  MN-major, B bf16 MN-major          (wgrad: X^T x dY, both reduced over pixels)
// Operands are small integers (exact in f16 / bf16 / f32), so a layout or descriptor mistake is a
// gross mismatch.  Shared-memory images are written by plain stores in the canonical 128B-swizzled
// layouts, i.e. exactly what the TMA SWIZZLE_128B boxes of the engine produce.
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <vector>

#include "common.cuh"

using namespace var;

constexpr int M = 128, N = 64, K = 64;

// A: logical [M][K], B: logical [N][K] (both "K-major" logical indexing); *_mn selects the smem image.
__global__ void __launch_bounds__(128) probe(const uint16_t* A, const uint16_t* B, float* D, int a_fmt, int b_fmt,
                                            int a_mn, int b_mn, int lbo_a) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* sm = raw + (base - smem_u32(raw));
  uint8_t* sA = sm;            // 16 KB
  uint8_t* sB = sm + 16384;    // 8 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x;
  for (int i = tid; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    uint32_t off;
    if (!a_mn) off = (uint32_t)m * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)m & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
    else {  // MN-major: atoms of 64 m; row = k (128 B = 64 m), chunk = (m % 64) / 8
      const uint32_t atom = (uint32_t)m >> 6, mm = (uint32_t)m & 63u;
      off = atom * (uint32_t)lbo_a + (uint32_t)k * 128u + (((mm >> 3) ^ ((uint32_t)k & 7u)) << 4) + (mm & 7u) * 2u;
    }
    *reinterpret_cast<uint16_t*>(sA + off) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    uint32_t off;
    if (!b_mn) off = (uint32_t)n * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)k & 7u) * 2u;
    else off = (uint32_t)k * 128u + ((((uint32_t)n >> 3) ^ ((uint32_t)k & 7u)) << 4) + ((uint32_t)n & 7u) * 2u;
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
    const uint32_t idesc = make_idesc_h16(N, a_fmt, b_fmt, a_mn, b_mn);
    for (int j = 0; j < K / 16; ++j) {
      // K-major: +32 B per 16 elements inside the 128 B row; MN-major: +16 k-rows = 2048 B
      const uint64_t ad = a_mn ? make_smem_desc(smem_u32(sA) + (uint32_t)j * 2048u, (uint32_t)lbo_a, 1024u, 2)
                               : make_smem_desc(smem_u32(sA) + (uint32_t)j * 32u, 16u, 1024u, 2);
      const uint64_t bd = b_mn ? make_smem_desc(smem_u32(sB) + (uint32_t)j * 2048u, 8192u, 1024u, 2)
                               : make_smem_desc(smem_u32(sB) + (uint32_t)j * 32u, 16u, 1024u, 2);
      umma_f16(tmem, ad, bd, idesc, j != 0);
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  const int warp = tid >> 5, lane = tid & 31;
  for (int c = 0; c < N; c += 32) {
    float v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 64);
}

static uint16_t enc(float x, int fmt) {
  if (fmt == 0) { __half h = __float2half(x); return *reinterpret_cast<uint16_t*>(&h); }
  __nv_bfloat16 b = __float2bfloat16(x);
  return *reinterpret_cast<uint16_t*>(&b);
}

int main(int argc, char** argv) {
  const int only = argc > 1 ? atoi(argv[1]) : -1;  // one case per process: a faulting MMA poisons the context
  struct Case { const char* name; int a_fmt, b_fmt, a_mn, b_mn; };
  const Case cases[] = {{"A f16 K / B f16 K", 0, 0, 0, 0}, {"A f16 K / B bf16 K (mixed)", 0, 1, 0, 0},
                        {"A bf16 K / B f16 MN (dgrad)", 1, 0, 0, 1}, {"A f16 MN / B bf16 MN (wgrad)", 0, 1, 1, 1},
                        {"A bf16 K / B bf16 K", 1, 1, 0, 0}, {"A f16 K / B f16 MN (dgrad)", 0, 0, 0, 1},
                        {"A f16 MN / B f16 MN (wgrad)", 0, 0, 1, 1}, {"A bf16 MN / B bf16 MN", 1, 1, 1, 1}};
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  int fails = 0;
  int ci = -1;
  for (const Case& c : cases) {
    ++ci;
    if (only >= 0 && ci != only) continue;
    std::vector<float> a(M * K), b(N * K);
    std::vector<uint16_t> ha(M * K), hb(N * K);
    uint32_t s = 7;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (int)((s >> 10) % 9) - 4; };
    for (int i = 0; i < M * K; ++i) { a[i] = (float)rnd(); ha[i] = enc(a[i], c.a_fmt); }
    for (int i = 0; i < N * K; ++i) { b[i] = (float)rnd(); hb[i] = enc(b[i], c.b_fmt); }
    uint16_t *dA, *dB; float* dD;
    cudaMalloc(&dA, ha.size() * 2); cudaMalloc(&dB, hb.size() * 2); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, M * N * 4);
    probe<<<1, 128, 32768>>>(dA, dB, dD, c.a_fmt, c.b_fmt, c.a_mn, c.b_mn, 8192);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> d(M * N);
    cudaMemcpy(d.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < K; ++k) ref += a[m * K + k] * b[n * K + k];
        worst = fmax(worst, fabs((double)d[m * N + n] - ref));
      }
    printf("%-34s cuda=%s  max|diff|=%g  %s\n", c.name, cudaGetErrorString(e), worst, (e == cudaSuccess && worst == 0) ? "OK" : "FAIL");
    if (e != cudaSuccess || worst != 0) ++fails;
    if (e != cudaSuccess) { cudaDeviceReset(); }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  printf("h16_probe: %d failing case(s)\n", fails);
  return fails ? 1 : 0;
}
