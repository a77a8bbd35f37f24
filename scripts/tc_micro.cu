// Standalone tcgen05 bring-up test (not part of the product): TMEM st/ld round trip, then one
// M128 x N x K8 kind::tf32 MMA with MN-major SW128 operands written by hand, for several descriptor variants.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void tmem_roundtrip(float* out) {
  __shared__ uint32_t base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_s)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_s;
  const uint32_t addr = base + ((uint32_t)(warp * 32) << 16);
  uint32_t v[8];
  for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(threadIdx.x * 100 + i));
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out[threadIdx.x * 8 + i] = __uint_as_float(r[i]);
  if (threadIdx.x == 0) out[128 * 8] = __uint_as_float(base);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64));
}

// One MMA: D[128 x N] = A[128 x 8] * B[N x 8]^T, operands MN-major SW128 in smem.
// smem layout assumption: chunk c (32 MN elements) at c*chunk_stride; within a chunk row k (0..7) at k*128 B,
// 16-byte units XOR-swizzled with k.
template <int N>
__global__ void mma_one(const float* A /*[8][128] k-major rows of 128 m*/, const float* B /*[8][N]*/, float* D /*[128][N]*/,
                        uint32_t lbo, uint32_t sbo, uint32_t idesc, int swizzle_mode, int use_mask) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint32_t base_s;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sa = (float*)sm;                 // 4 chunks x 1024 B
  float* sb = (float*)(sm + 8192);        // N/32 chunks x 1024 B
  const int chunk_floats = 256;           // 8 rows x 32 floats
  for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) {
    const int k = i / 128, m = i % 128;
    if (swizzle_mode == 2) {   // K-major, no swizzle: core matrix 8 rows x 16 B
      sa[(m / 8) * (sbo / 4) + (k / 4) * (lbo / 4) + (m % 8) * 4 + (k % 4)] = A[i];
    } else {
      const int c = m / 32, mm = m % 32;
      if (swizzle_mode == 3) {
        const int u32 = (mm / 8) ^ (k % 4);
        sa[c * chunk_floats + k * 32 + u32 * 8 + mm % 8] = A[i];
      } else {
        int unit = mm / 4, within = mm % 4;
        if (swizzle_mode) unit ^= k;
        sa[c * chunk_floats + k * 32 + unit * 4 + within] = A[i];
      }
    }
  }
  for (int i = threadIdx.x; i < 8 * N; i += blockDim.x) {
    const int k = i / N, n = i % N;
    if (swizzle_mode == 2) {
      sb[(n / 8) * (sbo / 4) + (k / 4) * (lbo / 4) + (n % 8) * 4 + (k % 4)] = B[i];
    } else {
      const int c = n / 32, nn = n % 32;
      if (swizzle_mode == 3) {
        const int u32 = (nn / 8) ^ (k % 4);
        sb[c * chunk_floats + k * 32 + u32 * 8 + nn % 8] = B[i];
      } else {
        int unit = nn / 4, within = nn % 4;
        if (swizzle_mode) unit ^= k;
        sb[c * chunk_floats + k * 32 + unit * 4 + within] = B[i];
      }
    }
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_s)), "r"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy (UMMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_s;
  if (threadIdx.x == 0) {
    auto mk = [&](uint32_t addr) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFF) >> 4);
      d |= (uint64_t)(lbo >> 4) << 16;
      d |= (uint64_t)(sbo >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)(swizzle_mode == 1 ? 2 : swizzle_mode == 3 ? 1 : 0) << 61;
      return d;
    };
    const uint64_t da = mk(smem_u32(sa)), db = mk(smem_u32(sb));
    if (use_mask) {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                   ::"r"(base), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(0u) : "memory");
    } else {
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(base), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everyone waits for the MMA
  asm volatile("{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}"
               ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t addr = base + ((uint32_t)(warp * 32) << 16);
  for (int cb = 0; cb < N; cb += 8) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr + cb));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) D[(warp * 32 + lane) * N + cb + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(N < 32 ? 32 : N));
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

int main() {
  float* out; CK(cudaMalloc(&out, 4096 * 4));
  CK(cudaMemset(out, 0, 4096 * 4));
  tmem_roundtrip<<<1, 128>>>(out);
  CK(cudaDeviceSynchronize());
  std::vector<float> h(1025);
  CK(cudaMemcpy(h.data(), out, 1025 * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int t = 0; t < 128; ++t) for (int i = 0; i < 8; ++i) if (h[t * 8 + i] != (float)(t * 100 + i)) ++bad;
  printf("tmem roundtrip: %d mismatches, base=0x%x, sample %g %g\n", bad, *(uint32_t*)&h[1024], h[0], h[8 * 37 + 3]);

  constexpr int N = 64;
  std::vector<float> A(8 * 128), B(8 * N), Dref(128 * N), Dh(128 * N);
  srand(1);
  auto tf = [](float v) { uint32_t b; memcpy(&b, &v, 4); b &= 0xFFFFE000u; memcpy(&v, &b, 4); return v; };
  for (auto& v : A) v = tf((rand() % 2001 - 1000) / 1000.f);
  for (auto& v : B) v = tf((rand() % 2001 - 1000) / 1000.f);
  for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { double s = 0; for (int k = 0; k < 8; ++k) s += (double)A[k * 128 + m] * B[k * N + n]; Dref[m * N + n] = (float)s; }
  float *dA, *dB, *dD; CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, Dh.size() * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(mma_one<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  const uint32_t idesc_mn = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc_k = idesc_mn & ~((1u << 15) | (1u << 16));
  struct V { const char* name; uint32_t lbo, sbo, idesc; int sw, mask; } vars[] = {
      {"MN sw128 lbo=1024 sbo=1024 nomask", 1024, 1024, idesc_mn, 1, 0},
      {"MN sw128 lbo=1024 sbo=1024 mask", 1024, 1024, idesc_mn, 1, 1},
      {"MN sw128 lbo=1024 sbo=2048", 1024, 2048, idesc_mn, 1, 0},
      {"MN sw128 lbo=2048 sbo=1024", 2048, 1024, idesc_mn, 1, 0},
      {"MN nosw  lbo=1024 sbo=1024", 1024, 1024, idesc_mn, 0, 0},
      {"K-major nosw lbo=128 sbo=256", 128, 256, idesc_k, 2, 0},
      {"MN sw128_base32B lbo=1024 sbo=512", 1024, 512, idesc_mn, 3, 0},
      {"MN sw128_base32B lbo=512 sbo=1024", 512, 1024, idesc_mn, 3, 0},
      {"K-major nosw lbo=128 sbo=256 mask", 128, 256, idesc_k, 2, 1},
  };
  for (auto& v : vars) {
    CK(cudaMemset(dD, 0xff, Dh.size() * 4));
    mma_one<N><<<1, 128, 32768>>>(dA, dB, dD, v.lbo, v.sbo, v.idesc, v.sw, v.mask);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", v.name, cudaGetErrorString(e)); return 1; }
    CK(cudaMemcpy(Dh.data(), dD, Dh.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; int nz = 0;
    for (size_t i = 0; i < Dh.size(); ++i) { maxerr = fmax(maxerr, fabs((double)Dh[i] - Dref[i])); maxref = fmax(maxref, fabs((double)Dref[i])); nz += Dh[i] != 0; }
    printf("%-40s: max err %.3e (max ref %.3f) nonzero %d  D[0][0..3] = %g %g %g %g  ref %g %g %g %g\n", v.name, maxerr, maxref, nz,
           Dh[0], Dh[1], Dh[2], Dh[3], Dref[0], Dref[1], Dref[2], Dref[3]);
  }
  return 0;
}
