// atr_simt.cu -- G += A_src^T R, the reduction GEMM of the weight gradient (admm.py:308-311):
//
//   G_acc[g][k][j] += sum_{t in chunk} sum_n A_src[t-1][k][n] * R[(g,j)][t][n]
//
// Both operands are contiguous along the reduction index (sample n), which is what the
// feature-major layout buys: the tile loads below are plain coalesced row reads.  fp32 CUDA-core
// version (general shapes); partial tiles are merged with fp64 atomics so that the result does not
// depend on the split.
#include "common.cuh"
#include "gate_gemm.h"

namespace admm {
namespace {

constexpr int BR = 16;          // reduction (sample) slab per stage
constexpr int NT = 256;

template <int KT, int CT, int MK, int MC>
__global__ void __launch_bounds__(NT) atr_simt_kernel(const AtrArgs p, int chunks_per_cta, int n_chunks) {
  constexpr int TK = KT / MK, TCC = CT / MC;
  static_assert(TK * TCC == NT, "thread grid");
  constexpr int VK = MK / 4, VC = MC / 4;
  __shared__ __align__(16) float as[2][BR][KT];
  __shared__ __align__(16) float bs[2][BR][CT];

  const int tid = threadIdx.x;
  const int tk = tid / TCC, tcx = tid % TCC;
  const int k0 = blockIdx.x * KT, c0 = blockIdx.y * CT;
  const int K = p.K, C = p.rows ? p.rows : 4 * p.H;
  const int rpg = p.rpg ? p.rpg : p.H;
  const int64_t ldn = p.ldn;
  const int chunks_per_t = (int)(ldn / BR);
  const int ch_begin = blockIdx.z * chunks_per_cta;
  const int ch_end = min(ch_begin + chunks_per_cta, n_chunks);
  if (ch_begin >= ch_end) return;

  constexpr int A_IT = (KT * BR / 4 + NT - 1) / NT;
  constexpr int B_IT = (CT * BR / 4 + NT - 1) / NT;
  float4 areg[A_IT], breg[B_IT];

  auto gload = [&](int ch) {
    const int tl = ch / chunks_per_t;
    const int64_t r0 = (int64_t)(ch % chunks_per_t) * BR;
    const float* ab = p.a_src + (int64_t)tl * p.a_tstride + r0;
    const float* bb = p.scratch + (int64_t)tl * ldn + r0;
#pragma unroll
    for (int i = 0; i < A_IT; ++i) {
      const int idx = tid + i * NT;
      const int kr = idx % KT, c4 = idx / KT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < BR / 4 && k0 + kr < K) v = *reinterpret_cast<const float4*>(ab + (int64_t)(k0 + kr) * ldn + c4 * 4);
      areg[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      const int idx = tid + i * NT;
      const int cr = idx % CT, c4 = idx / CT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c4 < BR / 4 && c0 + cr < C)
        v = *reinterpret_cast<const float4*>(bb + (int64_t)(c0 + cr) * p.tc * ldn + c4 * 4);
      breg[i] = v;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_IT; ++i) {
      const int idx = tid + i * NT;
      const int kr = idx % KT, c4 = idx / KT;
      if (c4 < BR / 4) {
        as[buf][c4 * 4 + 0][kr] = areg[i].x; as[buf][c4 * 4 + 1][kr] = areg[i].y;
        as[buf][c4 * 4 + 2][kr] = areg[i].z; as[buf][c4 * 4 + 3][kr] = areg[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      const int idx = tid + i * NT;
      const int cr = idx % CT, c4 = idx / CT;
      if (c4 < BR / 4) {
        bs[buf][c4 * 4 + 0][cr] = breg[i].x; bs[buf][c4 * 4 + 1][cr] = breg[i].y;
        bs[buf][c4 * 4 + 2][cr] = breg[i].z; bs[buf][c4 * 4 + 3][cr] = breg[i].w;
      }
    }
  };

  float acc[MK][MC];
#pragma unroll
  for (int a = 0; a < MK; ++a)
#pragma unroll
    for (int b = 0; b < MC; ++b) acc[a][b] = 0.f;

  gload(ch_begin);
  sstore(0);
  __syncthreads();
  for (int ch = ch_begin; ch < ch_end; ++ch) {
    const int buf = (ch - ch_begin) & 1;
    const bool more = ch + 1 < ch_end;
    if (more) gload(ch + 1);
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      float av[MK], bv[MC];
#pragma unroll
      for (int v = 0; v < VK; ++v) {
        const float4 t4 = *reinterpret_cast<const float4*>(&as[buf][r][v * (KT / VK) + tk * 4]);
        av[v * 4 + 0] = t4.x; av[v * 4 + 1] = t4.y; av[v * 4 + 2] = t4.z; av[v * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int v = 0; v < VC; ++v) {
        const float4 t4 = *reinterpret_cast<const float4*>(&bs[buf][r][v * (CT / VC) + tcx * 4]);
        bv[v * 4 + 0] = t4.x; bv[v * 4 + 1] = t4.y; bv[v * 4 + 2] = t4.z; bv[v * 4 + 3] = t4.w;
      }
#pragma unroll
      for (int a = 0; a < MK; ++a)
#pragma unroll
        for (int b = 0; b < MC; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int a = 0; a < MK; ++a) {
    const int k = k0 + (a / 4) * (KT / VK) + tk * 4 + (a % 4);
    if (k >= K) continue;
#pragma unroll
    for (int b = 0; b < MC; ++b) {
      const int c = c0 + (b / 4) * (CT / VC) + tcx * 4 + (b % 4);
      if (c >= C) continue;
      const int g = c / rpg, j = c % rpg;
      atomicAdd(p.g_acc + ((int64_t)g * K + k) * rpg + j, (double)acc[a][b]);
    }
  }
}

template <int KT, int CT, int MK, int MC>
int launch(const AtrArgs& a, cudaStream_t st) {
  const int n_chunks = (int)(a.tc * (a.ldn / BR));
  const int rows = a.rows ? a.rows : 4 * a.H;
  const int tiles = ((a.K + KT - 1) / KT) * ((rows + CT - 1) / CT);
  int splits = (2 * 148 + tiles - 1) / tiles;
  splits = max(1, min(splits, n_chunks));
  const int cpc = (n_chunks + splits - 1) / splits;
  splits = (n_chunks + cpc - 1) / cpc;
  dim3 grid((a.K + KT - 1) / KT, (rows + CT - 1) / CT, splits);
  KernelScope ks_("atr_simt_kernel", st);
  atr_simt_kernel<KT, CT, MK, MC><<<grid, NT, 0, st>>>(a, cpc, n_chunks);
  count_launch();
  return check_launch("atr_simt");
}

}  // namespace

int atr_simt(const AtrArgs& a, cudaStream_t st) {
  if (a.K <= 16) return launch<16, 256, 4, 4>(a, st);
  return launch<128, 128, 8, 8>(a, st);
}

}  // namespace admm
