// Bring-up test #2 (not product): K-major SWIZZLE_128B tf32 operands, 4 k-steps of 8 inside one 128-byte swizzle row
// (descriptor start address advanced by 32 B per k-step), accumulate flag, M=128 N=64.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int N = 64, KT = 32;
__global__ void k(const float* A /*[128][32]*/, const float* B /*[N][32]*/, float* D, uint32_t lbo, int adv_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t base_s;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sa = (float*)sm; float* sb = (float*)(sm + 128 * 128);
  for (int i = threadIdx.x; i < 128 * KT; i += blockDim.x) { int m = i / KT, kk = i % KT; sa[m * 32 + (((kk / 4) ^ (m % 8)) * 4) + kk % 4] = A[i]; }
  for (int i = threadIdx.x; i < N * KT; i += blockDim.x) { int n = i / KT, kk = i % KT; sb[n * 32 + (((kk / 4) ^ (n % 8)) * 4) + kk % 4] = B[i]; }
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_s)), "r"(N)); asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_s;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto mk = [&](uint32_t addr) { uint64_t d = 0; d |= (uint64_t)((addr & 0x3FFFF) >> 4); d |= (uint64_t)(lbo >> 4) << 16; d |= (uint64_t)(1024 >> 4) << 32; d |= (uint64_t)1 << 46; d |= (uint64_t)2 << 61; return d; };
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t da = mk(smem_u32(sa) + ks * adv_bytes), db = mk(smem_u32(sb) + ks * adv_bytes);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(base), "l"(da), "l"(db), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t addr = base + ((uint32_t)(warp * 32) << 16);
  for (int cb = 0; cb < N; cb += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr + cb));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) D[(warp * 32 + lane) * N + cb + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(N));
}
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
int main() {
  std::vector<float> A(128 * KT), B(N * KT), Dref(128 * N), Dh(128 * N);
  srand(2);
  auto tf = [](float v) { uint32_t b; memcpy(&b, &v, 4); b &= 0xFFFFE000u; memcpy(&v, &b, 4); return v; };
  for (auto& v : A) v = tf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = tf((rand() % 2001 - 1000) / 1000.f);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int kk = 0; kk < KT; ++kk) s += (double)A[m * KT + kk] * B[n * KT + kk]; Dref[m * N + n] = (float)s; }
  float *dA, *dB, *dD; CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, Dh.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  struct V { const char* name; uint32_t lbo; int adv; } vars[] = {{"K-major SW128, lbo=0, advance 32B", 0, 32}, {"K-major SW128, lbo=16, advance 32B", 16, 32}};
  for (auto& v : vars) {
    CK(cudaMemset(dD, 0xff, Dh.size() * 4));
    k<<<1, 128, 65536>>>(dA, dB, dD, v.lbo, v.adv);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0;
    for (size_t i = 0; i < Dh.size(); ++i) { maxerr = fmax(maxerr, fabs((double)Dh[i] - Dref[i])); maxref = fmax(maxref, fabs((double)Dref[i])); }
    printf("%-40s: max err %.3e (max ref %.3f)\n", v.name, maxerr, maxref);
  }
  return 0;
}
