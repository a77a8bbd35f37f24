// gate_gemm_tc.cu -- tcgen05 / TMEM / TMA version of the "gate GEMM + fused epilogue" family (sm_100a).
//
//   Z[n, (g,j)] = sum_k x_t[k,n] W_g[k,j] + sum_k h_{t-1}[k,n] U_g[k,j]
//
// computed on the 5th-generation tensor cores as an fp32-accurate split product of fp16 pairs
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo      (a_hi = fp16(a 2^sa), a_lo = fp16(a 2^sa - a_hi): 22 significand
//                                                  bits per operand, the same as a TF32 pair, at twice the MMA rate
//                                                  and half the operand bytes; power-of-two scales sa, sb keep both
//                                                  halves in fp16's normal range and are undone in the epilogue)
// because a plain TF32 / BF16 / FP16 product breaks the 1e-4 parity bar.
//
// Both operands are consumed in their native feature-major layout, i.e. MN-major for UMMA:
//   A tile = 128 samples x BK features   (sample index contiguous in HBM)
//   B tile = (4 gates x JC units) x BK   (unit index contiguous in HBM)
// TMA (128B swizzle, boxes of 64 halfs x BK rows) lands them in the canonical MN-major layout, so there is no
// transposed copy of the state or of the weights anywhere.
//
// Persistent kernel, one CTA per SM (576 threads): warp 0 = TMA producer (4-stage ring), warp 1 = TMEM allocator + MMA
// issuer (one lane), warps 2..17 = epilogue (thread = one sample row x 16 units; TMEM lane == sample, TMEM column ==
// (gate, unit)).  The accumulator is double-buffered in TMEM (2 x 256 columns), so the epilogue of tile i overlaps the
// MMAs of tile i+1; it never leaves the SM: the epilogue reads it with tcgen05.ld and applies the same closed forms as
// the CUDA-core path (admm_math.cuh); all its global accesses are 128-byte warp rows.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "gate_gemm.h"
#include "tc_path.h"

namespace admm {
namespace {

constexpr int BM = 128;        // samples per tile (UMMA M)
constexpr int BK = 32;         // features per pipeline stage (2 UMMA k-steps of 16)
constexpr int CHUNK_BYTES = 64 * BK * 2;     // one MN chunk: 64 halfs (128 B) x BK rows
// h operand scale 2^11.  h_t = (rho_h o tanh(c) - lambda_h)/rho_h with o an unconstrained ADMM primal, so |h| < 1 is typical
// (o stays near a sigmoid value) but NOT guaranteed; the fp16 pair represents |h| < 2^16 / 2^11 = 32 and saturates above.
// A clamped value raises the sticky device flag TcMeta::h_overflow (admm_tc_overflow in the C ABI, opt.metrics()).
constexpr int SCALE_H = 11;

struct TcMaps {
  CUtensorMap x, x_lo, h, h_lo;         // fp16 hi / lo halves of x 2^sX and h 2^SCALE_H; dims (ldn, K, slabs)
  CUtensorMap wx_hi, wx_lo, wh_hi, wh_lo;   // dims (H, K, 4)
  CUtensorMap gx_hi, gx_lo;             // gradient of the probed weight, dims (H, Ksrc, 4)
};

// fp32 tensor maps of the state streams of the staged epilogue: a state tensor [T+1][H][ldn] viewed as
// (sample, unit-in-group 16, unit group H/16, slab) with box (128, 1, 4, 1): ONE TMA instruction fetches, for one
// unit-step, the 128-sample rows of the four unit groups the sixteen epilogue warps work on (4 x 512 B, dense in shared
// memory as [group][sample]).  zstore [4][H][zT][ldn] is 5-D, its box takes all four gates at once.
struct StateMaps {
  CUtensorMap gate[6], dual[5], dual_h, zstore;
};

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// the same with an L2 eviction-priority hint (createpolicy encodings: evict_last 0x14F0000000000000, evict_first 0x12F0...)
__device__ __forceinline__ void tma_load_4d_hint(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols));
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// MN-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) | version=1 [46,48) | layout [61,64)
// 16-bit MN-major operands use plain SWIZZLE_128B (=2): rows of 128 B (64 samples/units) per k, 16-byte units
// XOR-swizzled with (k mod 8) -- what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B.
// LBO = byte stride between 64-element chunks along MN, SBO = byte stride between 8-row groups along K.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((uint32_t)(CHUNK_BYTES) >> 4) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=F32 [4,6)=1, A=F16 [7,10)=0, B=F16 [10,13)=0, A MN-major [15],
// B MN-major [16], N>>3 [17,23), M>>4 [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct Cfg {
  static constexpr int JC = 64;                       // hidden units per tile
  static constexpr int NCOL = 4 * JC;                 // accumulator columns: (gate, unit)
  static constexpr int TMEM_COLS = 256;
  static constexpr int A_BYTES = BM * BK * 2;         // 8 KB  (one half of the pair)
  static constexpr int B_BYTES = NCOL * BK * 2;       // 16 KB
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
};

// k-block range [kb_begin, kb_end) over the concatenated feature axis (x blocks first, then h blocks) and the
// choice of B operand: the gate weights (Z = x W + h U) or the probed weight's gradient (Q = A_src G).
struct TcRange {
  int kb_begin, kb_end;
  int b_is_grad;      // 0: B = (wx | wh) hi/lo ; 1: B = gradient hi/lo, k relative to kb_begin
};

// Largest power-of-two exponent c with m 2^c < 2^13 (0 for m = 0 / non-finite): both halves of the pair then sit high
// in fp16's normal range and 2^13 * 2^13 * K stays far from fp32 overflow in the accumulator.
__device__ __forceinline__ int cap_exp(unsigned max_bits) {
  const float m = __uint_as_float(max_bits);
  if (!(m > 0.f) || !isfinite(m)) return 0;
  int e;
  frexpf(m, &e);            // m = f 2^e, f in [0.5, 1)
  return 13 - e;
}

// fp16 pair of a state value v: hi = fp16(v 2^SCALE_H), lo = fp16(v 2^SCALE_H - hi)  (saturating, never inf)
__device__ __forceinline__ void split_f16(float vs, __half* hi, __half* lo) {
  const float c = fminf(fmaxf(vs, -65504.0f), 65504.0f);
  const __half h = __float2half_rn(c);
  *hi = h;
  *lo = __float2half_rn(c - __half2float(h));
}
__device__ __forceinline__ void store_h16(__half* hi, __half* lo, float h, unsigned* ovf) {
  const float vs = h * (float)(1 << SCALE_H);
  if (!(fabsf(vs) <= 65504.0f)) *ovf = 1u;        // clamped (or NaN): the operand no longer represents h
  split_f16(vs, hi, lo);
}

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// L2 prefetch of the twelve state entries the SWEEP epilogue loads for the unit at 32-bit offset o (lane == sample: one
// 128-byte line per warp and stream).  A prefetch holds no register, so it adds the memory-level parallelism the register
// file cannot: issued one batch of units ahead, the loads of the next batch find their lines in L2 (sweep kernel
// 0.61 -> 0.52 ms per cfg3 launch).  Not used by the GRAD epilogue, where it measured slower.
__device__ __forceinline__ void sweep_prefetch(const GateGemmArgs& p, uint32_t o) {
#pragma unroll
  for (int q = 0; q < 6; ++q) prefetch_l2(p.gate[q] + o);
  prefetch_l2(p.c_prev + o);
#pragma unroll
  for (int q = 0; q < 5; ++q) prefetch_l2(p.dual[q] + o);
}

// Epilogue of one tile for the units [u_begin, u_end) (multiples of 8) of the calling warp's 32 sample rows: reads the
// accumulator columns from TMEM (t_row = TMEM address of the warp's lane quarter, accumulator buffer included) and
// applies the closed forms of admm_math.cuh with coalesced (lane == sample) global accesses.
template <int MODE>
__device__ __forceinline__ void epilogue_units(const GateGemmArgs& p, uint32_t t_row, int u_begin, int u_end, int j0,
                                               int64_t n, bool ok, int tl, float (&msum)[5]) {
  constexpr int JC = Cfg::JC;
  const int H = p.H;
  const int64_t ldn = p.ldn;
  const Rho rho = p.rho;
  // Addressing: every access is base pointer + 32-bit element offset (one IMAD.WIDE.U32 instead of 64-bit index
  // arithmetic per access; the epilogue is bound by instruction issue, and integer index arithmetic was 40 % of it).
  // The offsets are relative to the first slab of the launch and stay below 2^32 for any tensor that fits the GPU
  // (checked on the host, gate_gemm_tc): row0 = this thread's sample in unit j0 of the tile's timestep, then one
  // 32-bit step per unit.  zstore / scratch are [4][H][zT or tc][ldn]: the gate stride goes into the (uniform) base.
  const uint32_t ldn32 = (uint32_t)ldn;
  const uint32_t row0 = (uint32_t)((int64_t)tl * p.s_tstride + (int64_t)j0 * ldn + n);
  const uint32_t dh0 = (uint32_t)((int64_t)j0 * ldn + n);                                 // lambda_h has the t = T slab only
  const uint32_t zrow0 = (uint32_t)(((int64_t)j0 * p.zT + p.zt0 + tl) * ldn + n);
  const uint32_t zstep = (uint32_t)((int64_t)p.zT * ldn);
  const int64_t zgate = (int64_t)H * p.zT * ldn;
  const uint32_t srow0 = (uint32_t)(((int64_t)j0 * p.tc + tl) * ldn + n);
  const uint32_t sstep = (uint32_t)((int64_t)p.tc * ldn);
  const int64_t sgate = (int64_t)H * p.tc * ldn;
    // The epilogue is latency-bound if each unit waits for its own loads (measured: ~3x the MMA time of a tile):
    // every batch of EB units first issues ALL its global loads, then computes, then stores, so ~13*EB loads are
    // in flight per thread.  Loads use the streaming path (each state entry is touched once per launch).
    constexpr int EB = 2;
    const float rho_g[4] = {rho.i, rho.f, rho.g, rho.o};
    const float inv_rho_g[4] = {1.0f / rho.i, 1.0f / rho.f, 1.0f / rho.g, 1.0f / rho.o};
    const float r_scale = (MODE == GG_GRAD && p.r16_hi) ? ldexpf(1.0f, cap_exp(*p.r_bound)) : 1.0f;
    const float acc_scale = *p.acc_scale;            // 2^-(sa + sb) of the operand pair of this launch
    for (int jb = u_begin; jb < u_end; jb += 8) {
      float z[4][8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        tmem_ld8(t_row + g * JC + jb, z[g]);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[g][i] *= acc_scale;
      }
#pragma unroll
      for (int eb = 0; eb < 8; eb += EB) {
        uint32_t off[EB], zoff[EB], soffs[EB];
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          const uint32_t ju = (uint32_t)(jb + eb + e);          // unit within the tile
          off[e] = row0 + ju * ldn32;
          zoff[e] = zrow0 + ju * zstep;
          soffs[e] = srow0 + ju * sstep;
        }

        if (MODE == GG_RAWZ) {
#pragma unroll
          for (int e = 0; e < EB; ++e)
#pragma unroll
            for (int g = 0; g < 4; ++g)
              __stcs(p.scratch + g * sgate + soffs[e], z[g][eb + e]);
        }
        if (MODE == GG_FORWARD) {
          float cp[EB];
#pragma unroll
          for (int e = 0; e < EB; ++e) cp[e] = __ldcs(p.c_prev + off[e]);
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            const ForwardResult r = forward_point<FastMath>(z[0][eb + e], z[1][eb + e], z[2][eb + e], z[3][eb + e], cp[e]);
            if (p.zstore) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                __stcs(p.zstore + g * zgate + zoff[e], z[g][eb + e]);
            }
            if (p.gate[0]) __stcs(p.gate[0] + off[e], r.i);
            if (p.gate[1]) __stcs(p.gate[1] + off[e], r.f);
            if (p.gate[2]) __stcs(p.gate[2] + off[e], r.g);
            if (p.gate[3]) __stcs(p.gate[3] + off[e], r.o);
            __stcs(p.gate[4] + off[e], r.c);
            p.gate[5][off[e]] = r.h;
            store_h16(p.h16_hi + off[e], p.h16_lo + off[e], r.h, p.h_ovf);   // fp16 pair: the next timestep's A operand
          }
        }
        if (MODE == GG_SWEEP) {
          float in[13][EB];
          if (p.last < 0) {            // never taken (see note below)
#pragma unroll
            for (int e = 0; e < EB; ++e)
#pragma unroll
              for (int q = 0; q < 13; ++q) in[q][e] = 0.25f + 0.01f * q;
          } else {
#pragma unroll
            for (int e = 0; e < EB; ++e) {
#pragma unroll
              for (int q = 0; q < 6; ++q) in[q][e] = __ldcs(p.gate[q] + off[e]);
              in[6][e] = __ldcs(p.c_prev + off[e]);
#pragma unroll
              for (int q = 0; q < 5; ++q) in[7 + q][e] = __ldcs(p.dual[q] + off[e]);
              in[12][e] = p.last ? __ldcs(p.dual_h + (dh0 + (uint32_t)(jb + eb + e) * ldn32)) : 0.f;
            }
            if (p.epi_prefetch && jb + eb + EB < u_end) {
#pragma unroll
              for (int e = 0; e < EB; ++e) sweep_prefetch(p, off[e] + EB * ldn32);
            }
          }
          // NOTE on the never-taken branches (p.last is 0 or 1): they only shape ptxas' schedule.  Putting the loads of a
          // batch in their own basic block keeps them together ahead of the arithmetic, and the guarded `continue`
          // keeps each unit's stores behind its arithmetic; without them ptxas interleaves load / use / store per value
          // and the kernel takes 1.21 ms instead of 0.78 ms per launch (cfg3 shape; measured with scripts/gemm_bench.py).
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            SweepPoint s;
            s.zi = z[0][eb + e]; s.zf = z[1][eb + e]; s.zg = z[2][eb + e]; s.zo = z[3][eb + e];
            s.i = in[0][e]; s.f = in[1][e]; s.g = in[2][e]; s.o = in[3][e]; s.c = in[4][e]; s.h = in[5][e];
            s.c_prev = in[6][e];
            s.li = in[7][e]; s.lf = in[8][e]; s.lg = in[9][e]; s.lo = in[10][e]; s.lc = in[11][e]; s.lh = in[12][e];
            const SweepResult r = sweep_point<FastMath>(s, rho, p.last != 0);
            if (p.last < 0) {
              if (r.i == 123.456f) p.gate[0][off[e]] = r.i + r.f + r.g + r.o + r.c + r.h + r.li + r.lf + r.lg + r.lo + r.lc;
              continue;
            }
            __stcs(p.gate[0] + off[e], r.i); __stcs(p.gate[1] + off[e], r.f); __stcs(p.gate[2] + off[e], r.g);
            __stcs(p.gate[3] + off[e], r.o);
            p.gate[4][off[e]] = r.c;                         // c_t is read again by the next timestep
            if (!p.last) {
              p.gate[5][off[e]] = r.h;
              store_h16(p.h16_hi + off[e], p.h16_lo + off[e], r.h, p.h_ovf);
            }
            __stcs(p.dual[0] + off[e], r.li); __stcs(p.dual[1] + off[e], r.lf); __stcs(p.dual[2] + off[e], r.lg);
            __stcs(p.dual[3] + off[e], r.lo); __stcs(p.dual[4] + off[e], r.lc);
            if (p.zstore) {
#pragma unroll
              for (int g = 0; g < 4; ++g)
                __stcs(p.zstore + g * zgate + zoff[e], z[g][eb + e]);
            }
            if (ok) { msum[0] += r.prim_sq; msum[1] += r.dual_sq; msum[2] += r.penalty; }
            if (p.xbound_track)
              msum[4] = fmaxf(fmaxf(fmaxf(msum[4], fmaf(fabsf(r.li), inv_rho_g[0], fabsf(r.i))),
                                    fmaxf(fmaf(fabsf(r.lf), inv_rho_g[1], fabsf(r.f)), fmaf(fabsf(r.lg), inv_rho_g[2], fabsf(r.g)))),
                              fmaf(fabsf(r.lo), inv_rho_g[3], fabsf(r.o)));
          }
        }
        if (MODE == GG_GRAD) {
          float lam[4][EB], gv[4][EB], zold[4][EB];
          if (p.tc < 0) {              // never taken: same schedule-shaping device as in SWEEP above
#pragma unroll
            for (int e = 0; e < EB; ++e)
#pragma unroll
              for (int g = 0; g < 4; ++g) { lam[g][e] = 0.1f * g; gv[g][e] = 0.2f; zold[g][e] = 0.3f; }
          } else {
#pragma unroll
            for (int e = 0; e < EB; ++e)
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                lam[g][e] = __ldcs(p.dual[g] + off[e]);
                gv[g][e] = __ldcs(p.gate[g] + off[e]);
                zold[g][e] = p.z_accumulate
                    ? __ldcs(p.zstore + g * zgate + zoff[e]) : 0.f;
              }
          }
#pragma unroll
          for (int e = 0; e < EB; ++e) {
            float rv[4], zz[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float u;
              zz[g] = z[g][eb + e] + zold[g][e];
              const float rr = grad_point<FastMath>(zz[g], lam[g][e], gv[g][e], rho_g[g], g == 2, &u);
              rv[g] = ok ? rr : 0.f;
              if (ok) msum[g] += u * u;
              if (p.bound_track) msum[4] = fmaxf(msum[4], 1.0f + fabsf(lam[g][e]) * inv_rho_g[g] + fabsf(gv[g][e]));
            }
            if (p.tc < 0) {            // never taken
              if (rv[0] == 123.456f) p.scratch[off[e]] = rv[0] + rv[1] + rv[2] + rv[3] + zz[0] + zz[1] + zz[2] + zz[3];
              continue;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (p.zstore) __stcs(p.zstore + g * zgate + zoff[e], zz[g]);
              if (p.r16_hi) {                                // h-phase: fp16 pair for the fp16 A^T R GEMM
                split_f16(rv[g] * r_scale, p.r16_hi + g * sgate + soffs[e], p.r16_lo + g * sgate + soffs[e]);
              } else {
                *(p.scratch + g * sgate + soffs[e]) = rv[g];   // read right back by the A^T R GEMM: keep in L2
                if (p.scratch_q) *(p.scratch_q + g * sgate + soffs[e]) = tf32_lo(rv[g]);
              }
            }
          }
        }
      }
    }
}


// ------------------------------------------------------------------------------------------------------------
// Staged epilogue inputs (SWEEP and the h-phase GRAD refresh).  In the epilogue above every state value is a global load
// issued from the epilogue warps' own registers: with 16 warps per SM the loads in flight are bounded by the register
// file, the warps sit in long_scoreboard stalls (round 1: issue 46 %, DRAM 40 %) and an L2 prefetch was the only lever.
// Here a dedicated LOADER warp streams the state rows of the tile into a shared-memory ring with TMA box copies
// (cp.async.bulk.tensor ... mbarrier::complete_tx), up to S_STAGES unit-steps ahead of the epilogue and independent of it, so the
// memory pipeline stays full while the epilogue warps only read shared memory (~30 cycles), compute and store.
//   stage = one unit per unit group (4 groups of 16 units: what the 16 epilogue warps consume at the same time)
//           x 128 samples x S_STREAMS state streams; one TMA box (StateMaps) per stream and stage; 16 stages per tile.
// (A first version issued the 48 rows of a stage as 1-D bulk copies with per-lane addresses: ptxas serialises those
// through the uniform datapath -- ELECT / R2UR / UBLKCP per lane -- and the loader, not memory, bounded the kernel:
// ncu showed the epilogue warps waiting on the stage barrier, 0.65 ms per cfg3 sweep launch against 0.49 ms unstaged.)
constexpr int S_STREAMS = 13;                          // SWEEP: i f g o c h | c_{t-1} | lambda_i f g o c | lambda_h (t = T only)
constexpr int S_ROW_BYTES = BM * 4;                    // 128 samples of one unit
constexpr int S_STAGE_BYTES = S_STREAMS * 4 * S_ROW_BYTES;     // 26 KB
constexpr int S_STAGES = 3;

// number of streams of a launch
template <int MODE>
__device__ __forceinline__ int staged_streams(const GateGemmArgs& p) {
  if (MODE == GG_SWEEP) return p.last ? 13 : 12;
  if (MODE == GG_MOMENTS) return 12;                   // lambda_g, gate_g, stored z_g
  return p.z_accumulate ? 12 : 8;                      // GRAD: lambda_g, gate_g (, stored z_g)
}
// One tile of the staged epilogue for the calling warp (lane quarter `quarter`, unit group `ugrp`): 16 unit-steps, each
// waits for its stage, copies its 8..13 inputs to registers, releases the stage and then computes / stores exactly like
// epilogue_units (same closed forms, same store policy).  `sit` = running stage counter of this CTA.
template <int MODE>
__device__ __forceinline__ void epilogue_staged(const GateGemmArgs& p, uint32_t t_row, int ugrp, int quarter, int lane, int j0,
                                                int64_t n, bool ok, int tl, float (&msum)[5], const uint8_t* sring,
                                                uint64_t* sfull, uint64_t* sempty, uint32_t& sit) {
  constexpr int JC = Cfg::JC;
  const int H = p.H;
  const int64_t ldn = p.ldn;
  const Rho rho = p.rho;
  const uint32_t ldn32 = (uint32_t)ldn;
  const int u0 = ugrp * (JC / 4);
  const uint32_t row0 = (uint32_t)((int64_t)tl * p.s_tstride + (int64_t)(j0 + u0) * ldn + n);
  const uint32_t zrow0 = (uint32_t)(((int64_t)(j0 + u0) * p.zT + p.zt0 + tl) * ldn + n);
  const uint32_t zstep = (uint32_t)((int64_t)p.zT * ldn);
  const int64_t zgate = (int64_t)H * p.zT * ldn;
  const uint32_t srow0 = (uint32_t)(((int64_t)(j0 + u0) * p.tc + tl) * ldn + n);
  const uint32_t sstep = (uint32_t)((int64_t)p.tc * ldn);
  const int64_t sgate = (int64_t)H * p.tc * ldn;
  const float rho_g[4] = {rho.i, rho.f, rho.g, rho.o};
  const float inv_rho_g[4] = {1.0f / rho.i, 1.0f / rho.f, 1.0f / rho.g, 1.0f / rho.o};
  const float r_scale = (MODE == GG_GRAD && p.r16_hi) ? ldexpf(1.0f, cap_exp(*p.r_bound)) : 1.0f;
  const float acc_scale = *p.acc_scale;
  const uint32_t my = (uint32_t)(ugrp * BM + quarter * 32 + lane);       // float index of this thread inside a stream block
#pragma unroll 1
  for (int jb = 0; jb < JC / 4; jb += 8) {
    float z[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      tmem_ld8(t_row + g * JC + u0 + jb, z[g]);
#pragma unroll
      for (int i = 0; i < 8; ++i) z[g][i] *= acc_scale;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e, ++sit) {
      const uint32_t st = sit % S_STAGES;
      mbar_wait(&sfull[st], (sit / S_STAGES) & 1);
      const float* sp = reinterpret_cast<const float*>(sring + st * S_STAGE_BYTES) + my;
      const uint32_t ju = (uint32_t)(jb + e);
      const uint32_t off = row0 + ju * ldn32, zoff = zrow0 + ju * zstep, soff = srow0 + ju * sstep;
      if (MODE == GG_SWEEP) {
        float in[13];
#pragma unroll
        for (int q = 0; q < 12; ++q) in[q] = sp[q * (4 * BM)];
        in[12] = p.last ? sp[12 * (4 * BM)] : 0.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty[st]);            // values are in registers: the loader may refill the stage
        SweepPoint sp_;
        sp_.zi = z[0][e]; sp_.zf = z[1][e]; sp_.zg = z[2][e]; sp_.zo = z[3][e];
        sp_.i = in[0]; sp_.f = in[1]; sp_.g = in[2]; sp_.o = in[3]; sp_.c = in[4]; sp_.h = in[5];
        sp_.c_prev = in[6];
        sp_.li = in[7]; sp_.lf = in[8]; sp_.lg = in[9]; sp_.lo = in[10]; sp_.lc = in[11]; sp_.lh = in[12];
        const SweepResult r = sweep_point<FastMath>(sp_, rho, p.last != 0);
        __stcs(p.gate[0] + off, r.i); __stcs(p.gate[1] + off, r.f); __stcs(p.gate[2] + off, r.g);
        __stcs(p.gate[3] + off, r.o);
        p.gate[4][off] = r.c;                                // c_t is read again by the next timestep
        if (!p.last) {
          p.gate[5][off] = r.h;
          store_h16(p.h16_hi + off, p.h16_lo + off, r.h, p.h_ovf);
        }
        __stcs(p.dual[0] + off, r.li); __stcs(p.dual[1] + off, r.lf); __stcs(p.dual[2] + off, r.lg);
        __stcs(p.dual[3] + off, r.lo); __stcs(p.dual[4] + off, r.lc);
        if (p.zstore) {
#pragma unroll
          for (int g = 0; g < 4; ++g) __stcs(p.zstore + g * zgate + zoff, z[g][e]);
        }
        if (ok) { msum[0] += r.prim_sq; msum[1] += r.dual_sq; msum[2] += r.penalty; }
        if (p.xbound_track)
          msum[4] = fmaxf(fmaxf(fmaxf(msum[4], fmaf(fabsf(r.li), inv_rho_g[0], fabsf(r.i))),
                                fmaxf(fmaf(fabsf(r.lf), inv_rho_g[1], fabsf(r.f)), fmaf(fabsf(r.lg), inv_rho_g[2], fabsf(r.g)))),
                          fmaf(fabsf(r.lo), inv_rho_g[3], fabsf(r.o)));
      } else {                                               // GG_GRAD
        float lam[4], gv[4], zold[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          lam[g] = sp[g * (4 * BM)];
          gv[g] = sp[(4 + g) * (4 * BM)];
          zold[g] = p.z_accumulate ? sp[(8 + g) * (4 * BM)] : 0.f;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sempty[st]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float u;
          const float zz = z[g][e] + zold[g];
          const float rr = grad_point<FastMath>(zz, lam[g], gv[g], rho_g[g], g == 2, &u);
          const float rv = ok ? rr : 0.f;
          if (ok) msum[g] += u * u;
          if (p.bound_track) msum[4] = fmaxf(msum[4], 1.0f + fabsf(lam[g]) * inv_rho_g[g] + fabsf(gv[g]));
          if (p.zstore) __stcs(p.zstore + g * zgate + zoff, zz);
          if (p.r16_hi) {
            split_f16(rv * r_scale, p.r16_hi + g * sgate + soff, p.r16_lo + g * sgate + soff);
          } else {
            *(p.scratch + g * sgate + soff) = rv;
            if (p.scratch_q) *(p.scratch_q + g * sgate + soff) = tf32_lo(rv);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// MOMENTS epilogue (GateGemmArgs::mom_*): the moment pass of the backtracking probes fused behind the Q = A_src G GEMM.
// Inputs come through the staged ring exactly like the GRAD refresh (lambda_g, gate_g, stored z_g); Q is read from TMEM.
// Per thread 28 fp32 partial sums (4 gates x F0, B1..B6) live for ONE tile (16 unit-steps); after each tile a transposed
// warp reduction leaves sum i on lane i, which keeps its own fp64 total for the whole launch (two registers instead of
// 28 doubles), added to fk_acc once at the end.
__device__ __forceinline__ float tmem_ld1(uint32_t addr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(addr));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// v[i] (i < 32) summed over the 32 lanes; the total of index i ends up in v[0] of lane i
__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const bool upper = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = upper ? v[i] : v[i + w];
      const float keep = upper ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
    }
  }
}

// exact residual of one candidate: the activations of probe_eval.cu (two MUFU ops, <= 2 ulp)
__device__ __forceinline__ float moments_exact_u(bool is_g, float z, float lr, float gv) {
  const float a = is_g ? FastMath::tanh(z) : FastMath::sigmoid(z);
  return (a - lr) - gv;
}

template <int ORDER>
__device__ __forceinline__ void epilogue_moments(const GateGemmArgs& p, uint32_t t_row, int ugrp, int quarter, int lane, bool ok,
                                                 float (&qm)[4], double& macc, double (&pacc)[2], const uint8_t* sring,
                                                 uint64_t* sfull, uint64_t* sempty, uint32_t& sit) {
  constexpr int JC = Cfg::JC;
  const int u0 = ugrp * (JC / 4);
  const float acc_scale = *p.acc_scale;
  const uint32_t my = (uint32_t)(ugrp * BM + quarter * 32 + lane);
  const float rho_g[4] = {p.rho.i, p.rho.f, p.rho.g, p.rho.o};
  // per-gate constants of the tile, hoisted (the kernel is bound by its instruction count): lambda / rho is the correctly
  // rounded quotient in three instructions (div_rn below) from the correctly rounded reciprocal; Q -> t = Q 2^-k0 with the
  // accumulator scale and the ghost-row mask folded in
  float inv_rho[4], qscale[4], tscale[4];
  const float okf = ok ? 1.0f : 0.0f;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    inv_rho[g] = __fdiv_rn(1.0f, rho_g[g]);
    qscale[g] = acc_scale * okf;                                             // ghost rows: Q = 0 (no moment terms, no max)
    tscale[g] = __int_as_float((127 - p.mom_k0[g]) << 23);                    // 2^-k0
  }
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll 1
  for (int jj = 0; jj < JC / 4; ++jj, ++sit) {
    float q[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) q[g] = tmem_ld1(t_row + g * JC + u0 + jj);
    const uint32_t st = sit % S_STAGES;
    mbar_wait(&sfull[st], (sit / S_STAGES) & 1);
    const float* sp = reinterpret_cast<const float*>(sring + st * S_STAGE_BYTES) + my;
    float lam[4], gv[4], z0[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      lam[g] = sp[g * (4 * BM)];
      gv[g] = sp[(4 + g) * (4 * BM)];
      z0[g] = sp[(8 + g) * (4 * BM)];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&sempty[st]);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float lr = div_rn(lam[g], rho_g[g], inv_rho[g]);
      const float qv = q[g] * qscale[g];
      const float s = (g == 2) ? FastMath::tanh(z0[g]) : FastMath::sigmoid(z0[g]);
      const float u = (s - lr) - gv[g];
      const float t = qv * tscale[g];                    // Q 2^-k0
      float* a = acc + g * 8;
      a[0] = fmaf(u * okf, u, a[0]);
      if (ORDER == 4) {
        moment_accum4(g == 2, s, u, t, a);
      } else {
        const float t2 = t * t, t3 = t2 * t;
        float c[6];
        moment_terms(g == 2, s, u, c);
        a[1] = fmaf(c[0], t, a[1]);
        a[2] = fmaf(c[1], t2, a[2]);
        a[3] = fmaf(c[2], t3, a[3]);
        a[4] = fmaf(c[3] * t2, t2, a[4]);
        a[5] = fmaf(c[4] * t2, t3, a[5]);
        a[6] = fmaf(c[5] * t3, t3, a[6]);
      }
      qm[g] = fmaxf(qm[g], fabsf(qv));
    }
    if ((jj & 7) == 0) {
      // Lower-bound proofs below the expansion (any subset sum of squares is a lower bound of f(beta_k)): the candidates
      // k < mom_pc[g] (<= 16) evaluated exactly.  Every candidate below k0 on units 0 and 8 of every unit group = 1/8 of the
      // units, the density of the unfused path -- the margin f(beta_k) / est_k does NOT grow below the expansion (f saturates,
      // est keeps doubling; a 1/64 subset left gate o of the H = 256 workload undecided).  The two insurance candidates at
      // and above k0 (needed only if max|Q| grew since the hint) on one unit per tile.  Eight candidates x four gates per
      // transposed reduction; the second round only when a gate has more than eight.
      const bool sparse_too = (ugrp == 0 && jj == 0);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (p.mom_pc[0] <= 8 * half && p.mom_pc[1] <= 8 * half && p.mom_pc[2] <= 8 * half && p.mom_pc[3] <= 8 * half) continue;
        float v[32];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float lr = div_rn(lam[g], rho_g[g], inv_rho[g]);
          const float qv = q[g] * acc_scale;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int k = 8 * half + kk;
            float r = 0.f;
            if (k < p.mom_pc[g] && (k < p.mom_k0[g] || sparse_too)) {
              const float uu = moments_exact_u(g == 2, fmaf(qv, __int_as_float((127 - k) << 23), z0[g]), lr, gv[g]);
              r = ok ? uu * uu : 0.f;
            }
            v[g * 8 + kk] = r;
          }
        }
        warp_transpose_sum(v, lane);
        pacc[half] += (double)v[0];
      }
    }
  }
  warp_transpose_sum(acc, lane);
  macc += (double)acc[0];
}

// ------------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks the tiles of the launch (unit tile fastest, so that the CTAs running at
// the same time share A tiles in L2).  The accumulator is double-buffered in TMEM (2 x 256 columns): while the 16
// epilogue warps (512 threads, 16 units each) drain tile i, the MMA warp already accumulates tile i+1 from a 4-stage
// TMA ring, so the epilogue of a tile is hidden behind the next tile's MMAs instead of behind a second CTA.
// STAGED (SWEEP, h-phase GRAD): a third service warp (the state loader above) and a 3-stage state ring next to a 3-stage
// operand ring (3 x 48 KB + 3 x 26 KB = 222 KB of the 227 KB); otherwise a 4-stage operand ring and no loader.
constexpr int P_EPI_WARPS = 16;
template <bool STAGED> struct PCfg {
  static constexpr int STAGES = STAGED ? 3 : 4;
  static constexpr int EPI_W0 = STAGED ? 3 : 2;                          // first epilogue warp
  static constexpr int THREADS = (EPI_W0 + P_EPI_WARPS) * 32;            // 608 / 576
  static constexpr int SMEM_BYTES = STAGES * Cfg::STAGE_BYTES + (STAGED ? S_STAGES * S_STAGE_BYTES : 0) + 1024;
};

template <int MODE, bool STAGED>
__global__ void __launch_bounds__(PCfg<STAGED>::THREADS, 1)
gate_gemm_tc_persistent(const GateGemmArgs p, const __grid_constant__ TcMaps maps, const __grid_constant__ StateMaps smaps,
                        int slab0, const TcRange rng, int n_jt, int n_nt, int n_tiles) {
  using C = Cfg;
  constexpr int JC = C::JC;
  constexpr int P_STAGES = PCfg<STAGED>::STAGES;
  constexpr int EPI_W0 = PCfg<STAGED>::EPI_W0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sring = smem + P_STAGES * C::STAGE_BYTES;                      // state ring (STAGED only)
  __shared__ __align__(8) uint64_t full_bar[P_STAGES], empty_bar[P_STAGES], tfull_bar[2], tempty_bar[2];
  __shared__ __align__(8) uint64_t sfull_bar[S_STAGES], sempty_bar[S_STAGES];
  __shared__ uint32_t tmem_base_s;
  __shared__ float red[4 * P_EPI_WARPS];

  // warp index through a shuffle: tells ptxas it is warp-uniform, so the role branches below are uniform branches and
  // the epilogue keeps its pointers / memory descriptors on the uniform datapath (no R2UR round trips per load)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int D = p.D, H = p.H;
  const int nkx = (D + BK - 1) / BK;
  const int nkb = rng.kb_end - rng.kb_begin;
  (void)H;

  if (p.run_if && *p.run_if == 0) return;
  if ((MODE == GG_RAWZ || MODE == GG_MOMENTS) && p.done) {
    if (p.done[0] && p.done[1] && p.done[2] && p.done[3]) return;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < P_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], P_EPI_WARPS); }
    for (int s = 0; s < S_STAGES; ++s) { mbar_init(&sfull_bar[s], 1); mbar_init(&sempty_bar[s], P_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;                                   // running k-block counter over all tiles of this CTA
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int j0 = (tile % n_jt) * JC;
        const int n0 = ((tile / n_jt) % n_nt) * BM;
        const int tl = tile / (n_jt * n_nt);
        const int slab = slab0 + tl;
        for (int li = 0; li < nkb; ++li, ++it) {
          const int kb = rng.kb_begin + li;
          const int s = it % P_STAGES;
          const uint32_t round = it / P_STAGES;
          if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
          const bool is_x = kb < nkx;
          const int k0 = (is_x ? kb : kb - nkx) * BK;
          uint8_t* st = smem + s * C::STAGE_BYTES;
          mbar_expect_tx(&full_bar[s], 2 * C::A_BYTES + 2 * C::B_BYTES);
          const CUtensorMap* ma = is_x ? &maps.x : &maps.h;
          const CUtensorMap* ml = is_x ? &maps.x_lo : &maps.h_lo;
          tma_load_4d(st, ma, &full_bar[s], 0, k0, n0 / 64, slab);
          tma_load_4d(st + C::A_BYTES, ml, &full_bar[s], 0, k0, n0 / 64, slab);
          const CUtensorMap* bh = rng.b_is_grad ? &maps.gx_hi : (is_x ? &maps.wx_hi : &maps.wh_hi);
          const CUtensorMap* bl = rng.b_is_grad ? &maps.gx_lo : (is_x ? &maps.wx_lo : &maps.wh_lo);
          const int kb0 = rng.b_is_grad ? li * BK : k0;
          uint8_t* sb = st + 2 * C::A_BYTES;
          if (p.tma_hint) {
            // the weight operand (18 MB of fp16 pairs at H = 1024) is re-read by every sample tile: keep it in L2 against the
            // state stream (1.8 GB per launch) flowing through
            tma_load_4d_hint(sb, bh, &full_bar[s], 0, kb0, j0 / 64, 0, 0x14F0000000000000ull);
            tma_load_4d_hint(sb + C::B_BYTES, bl, &full_bar[s], 0, kb0, j0 / 64, 0, 0x14F0000000000000ull);
          } else {
            tma_load_4d(sb, bh, &full_bar[s], 0, kb0, j0 / 64, 0);
            tma_load_4d(sb + C::B_BYTES, bl, &full_bar[s], 0, kb0, j0 / 64, 0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, C::NCOL);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const uint32_t buf = ti & 1, use = ti >> 1;      // `use`-th time this accumulator buffer is filled
        if (use > 0) {                                   // wait until the epilogue drained its previous contents
          mbar_wait(&tempty_bar[buf], (use - 1) & 1);
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + buf * C::NCOL;
        for (int li = 0; li < nkb; ++li, ++it) {
          const int s = it % P_STAGES;
          mbar_wait(&full_bar[s], (it / P_STAGES) & 1);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + C::A_BYTES;
          const uint32_t b_hi = st + 2 * C::A_BYTES, b_lo = b_hi + C::B_BYTES;
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks) {
            const uint32_t off = ks * 2048;                // 16 features = two 8-row swizzle groups
            umma_f16(d_tmem, make_desc(a_hi + off), make_desc(b_hi + off), idesc, (li > 0 || ks > 0) ? 1u : 0u);
            umma_f16(d_tmem, make_desc(a_lo + off), make_desc(b_hi + off), idesc, 1u);
            umma_f16(d_tmem, make_desc(a_hi + off), make_desc(b_lo + off), idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
      }
    }
  } else if (STAGED && warp == 2) {
    // ------------------------------------------------------------------ state loader (staged epilogue inputs)
    const int ns = staged_streams<MODE>(p);
    uint32_t sit = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int j0 = (tile % n_jt) * JC;
      const int64_t n0 = (int64_t)((tile / n_jt) % n_nt) * BM;
      const int tl = tile / (n_jt * n_nt);
      for (int jj = 0; jj < JC / 4; ++jj, ++sit) {
        const uint32_t st = sit % S_STAGES, round = sit / S_STAGES;
        if (round > 0) mbar_wait(&sempty_bar[st], (round - 1) & 1);
        uint8_t* dst = sring + st * S_STAGE_BYTES;
        if (lane == 0) {
          constexpr int SB = 4 * S_ROW_BYTES;               // bytes of one stream block: [4 unit groups][128 samples]
          uint64_t* bar = &sfull_bar[st];
          const int n0i = (int)n0, jg = j0 / (JC / 4), slab = slab0 + 1 + tl;
          mbar_expect_tx(bar, (uint32_t)(ns * SB));
          if (MODE == GG_SWEEP) {
#pragma unroll
            for (int q = 0; q < 6; ++q) tma_load_4d(dst + q * SB, &smaps.gate[q], bar, n0i, jj, jg, slab);
            tma_load_4d(dst + 6 * SB, &smaps.gate[4], bar, n0i, jj, jg, slab - 1);           // c_{t-1}
#pragma unroll
            for (int q = 0; q < 5; ++q) tma_load_4d(dst + (7 + q) * SB, &smaps.dual[q], bar, n0i, jj, jg, slab);
            if (ns > 12) tma_load_3d(dst + 12 * SB, &smaps.dual_h, bar, n0i, jj, jg);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) tma_load_4d(dst + q * SB, &smaps.dual[q], bar, n0i, jj, jg, slab);
#pragma unroll
            for (int q = 0; q < 4; ++q) tma_load_4d(dst + (4 + q) * SB, &smaps.gate[q], bar, n0i, jj, jg, slab);
            if (ns > 8) tma_load_5d(dst + 8 * SB, &smaps.zstore, bar, n0i, p.zt0 + tl, jj, jg, 0);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (16 warps)
    const int ew = warp - EPI_W0;
    const int quarter = warp & 3;                      // TMEM lanes 32*quarter .. +31
    const int ugrp = ew >> 2;                          // 4 warps per lane quarter: 16 of the tile's 64 units each
    const int row = quarter * 32 + lane;
    float msum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};      // [4] = running max of the |R| bound (GRAD, x-phase)
    float qm[4] = {0.f, 0.f, 0.f, 0.f};             // MOMENTS: max |Q| per gate
    double macc = 0.0, pacc[2] = {0.0, 0.0};         // MOMENTS: this lane's moment / proof totals (warp_transpose_sum)
    uint32_t ti = 0, sit = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int j0 = (tile % n_jt) * JC;
      const int n0 = ((tile / n_jt) % n_nt) * BM;
      const int tl = tile / (n_jt * n_nt);
      const uint32_t buf = ti & 1, use = ti >> 1;
      const int64_t n = (int64_t)n0 + row;
      mbar_wait(&tfull_bar[buf], use & 1);
      tc_fence_after();
      const uint32_t t_row = tmem_base + buf * C::NCOL + ((uint32_t)(quarter * 32) << 16);
      if (MODE == GG_MOMENTS) {
        if (p.mom_order == 4) epilogue_moments<4>(p, t_row, ugrp, quarter, lane, n < p.n, qm, macc, pacc, sring, sfull_bar, sempty_bar, sit);
        else epilogue_moments<6>(p, t_row, ugrp, quarter, lane, n < p.n, qm, macc, pacc, sring, sfull_bar, sempty_bar, sit);
      }
      else if (STAGED) epilogue_staged<MODE>(p, t_row, ugrp, quarter, lane, j0, n, n < p.n, tl, msum, sring, sfull_bar, sempty_bar, sit);
      else epilogue_units<MODE>(p, t_row, ugrp * (JC / 4), (ugrp + 1) * (JC / 4), j0, n, n < p.n, tl, msum);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tempty_bar[buf])) : "memory");
      }
    }
    if (MODE == GG_GRAD && p.bound_track) {
      float b = msum[4];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
      if (lane == 0) atomicMax(p.bound_track, __float_as_uint(b));
    }
    if (MODE == GG_MOMENTS) {
      // lane l of every warp holds the fp64 total of sum index l: (gate l / 8, moment or proof candidate l % 8)
      const int g = lane >> 3, k = lane & 7;
      if (k < 7) atomicAdd(p.fk_acc + g * ADMM_FK_SLOTS + ADMM_FK_MOMENTS + k, ldexp(macc, k * p.mom_k0[g]));   // B_k in units of Q^k
#pragma unroll
      for (int half = 0; half < 2; ++half) {           // proofs: lane l holds candidate 8 half + l % 8 of gate l / 8
        const int pg = lane >> 3, pk = 8 * half + (lane & 7);
        if (pk < p.mom_pc[pg]) atomicAdd(p.fk_acc + pg * ADMM_FK_SLOTS + ADMM_MAX_CAND + 1 + pk, pacc[half]);
      }
#pragma unroll
      for (int gg = 0; gg < 4; ++gg) {
        float m = qm[gg];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(p.qmax + gg), __float_as_uint(m));
      }
    }
    if (MODE == GG_SWEEP && p.xbound_track) {
      float b = msum[4];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
      if (lane == 0) {                                   // 1 + |lambda/rho| + |gate| >= |u| >= |R|
        atomicMax(p.xbound_track, __float_as_uint(1.0f + b));          // the bound in force: only ever raised here
        atomicMax(p.xbound_track + 1, __float_as_uint(1.0f + b));      // this sweep's own maximum (replaces [0] at t = T)
      }
    }
    if (MODE == GG_SWEEP || MODE == GG_GRAD) {
      double* dst = (MODE == GG_SWEEP) ? p.metrics : p.fw_acc;
      constexpr int NOUT = (MODE == GG_SWEEP) ? 3 : 4;
      if (dst) {
#pragma unroll
        for (int k = 0; k < NOUT; ++k) {
          const float sm_ = warp_sum(msum[k]);
          if (lane == 0) red[k * P_EPI_WARPS + ew] = sm_;
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const int et = threadIdx.x - EPI_W0 * 32;
        if (et < NOUT) {
          double acc = 0.0;
#pragma unroll
          for (int w = 0; w < P_EPI_WARPS; ++w) acc += (double)red[et * P_EPI_WARPS + w];
          atomicAdd(dst + et, acc);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeFn)ptr;
  }
  return fn;
}

// 4-D view of an fp16 tensor [outer][rows][cols] (cols contiguous) for MN-major operand tiles:
// dims (64 cols-in-chunk, rows, cols/64 chunks, outer), box (64, BK, box_chunks, box_outer), SWIZZLE_128B, zero fill
// out of bounds.  One box lands in shared memory as [outer][chunk][row][64], the canonical MN-major layout.
int make_map(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t outer, uint64_t row_stride_elems,
             uint64_t outer_stride_elems, uint32_t box_chunks, uint32_t box_outer) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ADMM_ECUDA; }
  cuuint64_t dims[4] = {64, rows, (cols + 63) / 64, outer};
  cuuint64_t strides[3] = {row_stride_elems * 2, 128, outer_stride_elems * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)BK, box_chunks, box_outer};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return ADMM_ECUDA; }
  return ADMM_OK;
}

// Scales and maxima kept on the device, so that preparing an operand never synchronises with the host.
struct TcMeta {
  unsigned max_x, max_wx, max_wh, max_g, max_d;   // bit patterns of max |.| (non-negative floats order like unsigned)
  int s_x;                                        // x is stored as fp16 pairs of x 2^s_x
  float scale_z, scale_q, scale_d;                // 2^-(sa+sb): accumulator -> z (weights), Q (gradient), x dW (refresh)
  unsigned r_bound;                               // bound on |R| of the A^T R operand (bit pattern), see GateGemmArgs
  // the same bound for the NEXT iteration, measured by the sweep over what it writes (xbound_track): [0] in force (raised
  // as the sweep goes, so a partial sweep stays covered), [1] the running maximum of the current sweep, copied to [0] at t = T
  unsigned x_bound[2];
  unsigned h_overflow;                            // sticky: an h value did not fit its fp16 pair (store_h16)
  int z_dirty;                                    // inputs were replaced by different values since the z store was written
  int pad_[2];
};

// workspace layout in floats (fp16 buffers take half a float per element):
// [x_lo] tf32 low part of x for the x-phase A^T R GEMM | fp16 pairs of x, h, wx, wh, g | TcMeta
struct WsLayout {
  int64_t x_lo, x16_hi, x16_lo, h16_hi, h16_lo, wx_hi, wx_lo, wh_hi, wh_lo, g_hi, g_lo, meta, total;
};
WsLayout ws_layout(const admm_problem* p) {
  WsLayout w;
  const int64_t T = p->T, D = p->D, H = p->H, ldn = p->ldn;
  const int64_t kmax = D > H ? D : H;
  int64_t o = 0;
  auto take = [&](int64_t n) { const int64_t r = o; o += (n + 255) / 256 * 256; return r; };
  auto take16 = [&](int64_t n) { return take((n + 1) / 2); };
  w.x_lo = take(T * D * ldn);
  w.x16_hi = take16(T * D * ldn); w.x16_lo = take16(T * D * ldn);
  w.h16_hi = take16((T + 1) * H * ldn); w.h16_lo = take16((T + 1) * H * ldn);
  w.wx_hi = take16(4 * D * H); w.wx_lo = take16(4 * D * H);
  w.wh_hi = take16(4 * H * H); w.wh_lo = take16(4 * H * H);
  w.g_hi = take16(4 * kmax * H); w.g_lo = take16(4 * kmax * H);
  w.meta = take(64);
  w.total = o;
  return w;
}
inline __half* ws_half(const admm_problem* p, int64_t off) { return reinterpret_cast<__half*>((float*)p->tc_ws + off); }
inline TcMeta* ws_meta(const admm_problem* p) { return reinterpret_cast<TcMeta*>((float*)p->tc_ws + ws_layout(p).meta); }

__global__ void absmax_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, unsigned* out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(b ? a[i] - b[i] : a[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

__global__ void split_trunc_kernel(const float* __restrict__ src, float* __restrict__ lo, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    float4 r;
    r.x = tf32_lo(v.x); r.y = tf32_lo(v.y); r.z = tf32_lo(v.z); r.w = tf32_lo(v.w);
    *reinterpret_cast<float4*>(lo + i) = r;
  } else {
    for (int64_t k = i; k < n; ++k) lo[k] = tf32_lo(src[k]);
  }
}

// What a conversion launch prepares.  The exponents are recomputed by every thread from the device-side maxima.
enum PrepKind { PREP_X = 0, PREP_H = 1, PREP_WX = 2, PREP_WH = 3, PREP_GRAD_X = 4, PREP_GRAD_H = 5, PREP_DELTA = 6 };

__global__ void prep_f16_kernel(int kind, const float* __restrict__ a, const float* __restrict__ b, __half* __restrict__ hi,
                                __half* __restrict__ lo, int64_t n, TcMeta* meta) {
  int ex = 0;            // element scale exponent
  if (kind == PREP_X) {
    ex = cap_exp(meta->max_x);
    if (blockIdx.x == 0 && threadIdx.x == 0) meta->s_x = ex;
  } else if (kind == PREP_H) {
    ex = SCALE_H;
  } else if (kind == PREP_WX || kind == PREP_WH) {
    // one accumulator holds x W + h U: s_x + s_wx == SCALE_H + s_wh == S
    const int sx = cap_exp(meta->max_x);
    const int S = min(sx + cap_exp(meta->max_wx), SCALE_H + cap_exp(meta->max_wh));
    ex = (kind == PREP_WX) ? S - sx : S - SCALE_H;
    if (blockIdx.x == 0 && threadIdx.x == 0) meta->scale_z = ldexpf(1.0f, -S);
  } else if (kind == PREP_GRAD_X || kind == PREP_GRAD_H) {
    ex = cap_exp(meta->max_g);
    const int sa = (kind == PREP_GRAD_X) ? cap_exp(meta->max_x) : SCALE_H;
    if (blockIdx.x == 0 && threadIdx.x == 0) meta->scale_q = ldexpf(1.0f, -(sa + ex));
  } else {               // PREP_DELTA: x (W_new - W_old)
    ex = cap_exp(meta->max_d);
    if (blockIdx.x == 0 && threadIdx.x == 0) meta->scale_d = ldexpf(1.0f, -(cap_exp(meta->max_x) + ex));
  }
  const float sc = ldexpf(1.0f, ex);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = (b ? a[i] - b[i] : a[i]) * sc;
    if (kind == PREP_H && !(fabsf(v) <= 65504.0f)) meta->h_overflow = 1u;
    split_f16(v, hi + i, lo + i);
  }
}

unsigned prep_grid(int64_t n) {
  const int64_t b = (n + 255) / 256;
  return (unsigned)(b < 148 * 8 ? (b < 1 ? 1 : b) : 148 * 8);
}

struct MapKey {
  const void *x, *h, *ws;
  int64_t ldn;
  int T, D, H;
  bool operator<(const MapKey& o) const {
    return std::tie(x, h, ws, ldn, T, D, H) < std::tie(o.x, o.h, o.ws, o.ldn, o.T, o.D, o.H);
  }
};
std::mutex g_map_mu;
std::map<MapKey, TcMaps> g_maps;

int get_maps(const admm_problem* p, int grad_src, TcMaps* out) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  const MapKey key{p->x, p->gate[5], p->tc_ws, p->ldn, p->T, p->D, p->H};
  auto it = g_maps.find(key);
  if (it == g_maps.end()) {
    if (g_maps.size() > 64) g_maps.clear();
    TcMaps m;
    const WsLayout w = ws_layout(p);
    const uint64_t ldn = p->ldn, T = p->T, D = p->D, H = p->H;
    int rc = 0;
    rc |= make_map(&m.x, ws_half(p, w.x16_hi), ldn, D, T, ldn, D * ldn, BM / 64, 1);
    rc |= make_map(&m.x_lo, ws_half(p, w.x16_lo), ldn, D, T, ldn, D * ldn, BM / 64, 1);
    rc |= make_map(&m.h, ws_half(p, w.h16_hi), ldn, H, T + 1, ldn, H * ldn, BM / 64, 1);
    rc |= make_map(&m.h_lo, ws_half(p, w.h16_lo), ldn, H, T + 1, ldn, H * ldn, BM / 64, 1);
    rc |= make_map(&m.wx_hi, ws_half(p, w.wx_hi), H, D, 4, H, D * H, Cfg::JC / 64, 4);
    rc |= make_map(&m.wx_lo, ws_half(p, w.wx_lo), H, D, 4, H, D * H, Cfg::JC / 64, 4);
    rc |= make_map(&m.wh_hi, ws_half(p, w.wh_hi), H, H, 4, H, H * H, Cfg::JC / 64, 4);
    rc |= make_map(&m.wh_lo, ws_half(p, w.wh_lo), H, H, 4, H, H * H, Cfg::JC / 64, 4);
    if (rc) return ADMM_ECUDA;
    it = g_maps.emplace(key, m).first;
  }
  *out = it->second;
  // the gradient maps depend on src (K = D or H); they are cheap to encode per call
  const WsLayout w = ws_layout(p);
  const uint64_t K = (grad_src == ADMM_SRC_X) ? p->D : p->H, H = p->H;
  int rc = make_map(&out->gx_hi, ws_half(p, w.g_hi), H, K, 4, H, K * H, Cfg::JC / 64, 4);
  rc |= make_map(&out->gx_lo, ws_half(p, w.g_lo), H, K, 4, H, K * H, Cfg::JC / 64, 4);
  return rc ? ADMM_ECUDA : ADMM_OK;
}

// fp32 state maps (see StateMaps).  dims (ldn, 16, H/16, slabs), box (128, 1, 4, 1), no swizzle.
int make_state_map(CUtensorMap* m, const float* base, uint64_t ldn, uint64_t H, uint64_t slabs) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ADMM_ECUDA; }
  cuuint64_t dims[4] = {ldn, 16, H / 16, slabs};
  cuuint64_t strides[3] = {ldn * 4, 16 * ldn * 4, H * ldn * 4};
  cuuint32_t box[4] = {(cuuint32_t)BM, 1, 4, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, slabs > 1 ? 4 : 3, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (state) failed (%d)", (int)r); return ADMM_ECUDA; }
  return ADMM_OK;
}
int make_zstore_map(CUtensorMap* m, const float* base, uint64_t ldn, uint64_t H, uint64_t zT) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ADMM_ECUDA; }
  cuuint64_t dims[5] = {ldn, zT, 16, H / 16, 4};
  cuuint64_t strides[4] = {ldn * 4, zT * ldn * 4, 16 * zT * ldn * 4, H * zT * ldn * 4};
  cuuint32_t box[5] = {(cuuint32_t)BM, 1, 1, 4, 4};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (zstore) failed (%d)", (int)r); return ADMM_ECUDA; }
  return ADMM_OK;
}

struct StateKey {
  const void *g[6], *d[5], *dh, *z;
  int64_t ldn;
  int T, H;
  bool operator<(const StateKey& o) const {
    return std::tie(g[0], g[1], g[2], g[3], g[4], g[5], d[0], d[1], d[2], d[3], d[4], dh, z, ldn, T, H) <
           std::tie(o.g[0], o.g[1], o.g[2], o.g[3], o.g[4], o.g[5], o.d[0], o.d[1], o.d[2], o.d[3], o.d[4], o.dh, o.z, o.ldn, o.T, o.H);
  }
};
std::map<StateKey, StateMaps> g_state_maps;

// State maps of a problem (cached).  The admm_problem pointers are the tensors' bases; a launch addresses slabs through
// the TMA coordinates.
int get_state_maps(const admm_problem* p, StateMaps* out) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  StateKey key;
  for (int q = 0; q < 6; ++q) key.g[q] = p->gate[q];
  for (int q = 0; q < 5; ++q) key.d[q] = p->dual[q];
  key.dh = p->dual_h; key.z = p->zstore; key.ldn = p->ldn; key.T = p->T; key.H = p->H;
  auto it = g_state_maps.find(key);
  if (it == g_state_maps.end()) {
    if (g_state_maps.size() > 64) g_state_maps.clear();
    StateMaps m;
    memset(&m, 0, sizeof(m));
    int rc = 0;
    for (int q = 0; q < 6; ++q) rc |= make_state_map(&m.gate[q], p->gate[q], p->ldn, p->H, p->T + 1);
    for (int q = 0; q < 5; ++q) rc |= make_state_map(&m.dual[q], p->dual[q], p->ldn, p->H, p->T + 1);
    if (p->dual_h) rc |= make_state_map(&m.dual_h, p->dual_h, p->ldn, p->H, 1);
    if (p->zstore) rc |= make_zstore_map(&m.zstore, p->zstore, p->ldn, p->H, p->T);
    if (rc) return ADMM_ECUDA;
    it = g_state_maps.emplace(key, m).first;
  }
  *out = it->second;
  return ADMM_OK;
}

template <int MODE, bool STAGED = false>
int launch_tc(const admm_problem* p, const GateGemmArgs& a, const TcMaps& maps, int slab0, int tc, const TcRange& rng,
              cudaStream_t st, const char* label) {
  using C = Cfg;
  KernelScope ks_(label, st);
  static bool configured_p = false;
  if (!configured_p) {
    cudaFuncSetAttribute(gate_gemm_tc_persistent<MODE, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         PCfg<STAGED>::SMEM_BYTES);
    configured_p = true;
  }
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  }
  const int n_jt = p->H / C::JC, n_nt = (int)(p->ldn / BM);
  const int n_tiles = n_jt * n_nt * tc;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  StateMaps smaps;
  if (STAGED) {
    const int rc = get_state_maps(p, &smaps);
    if (rc) return rc;
  } else {
    memset(&smaps, 0, sizeof(smaps));
  }
  gate_gemm_tc_persistent<MODE, STAGED><<<grid, PCfg<STAGED>::THREADS, PCfg<STAGED>::SMEM_BYTES, st>>>(a, maps, smaps, slab0, rng,
                                                                                                     n_jt, n_nt, n_tiles);
  count_launch();
  return check_launch("gate_gemm_tc_persistent");
}

}  // namespace

namespace {
// shape rules of the tensor-core path; the last one keeps the epilogue's 32-bit element offsets valid (a state tensor
// of 2^32 floats is 17 GB, eleven of them exceed the GPU's memory, so this never binds on a B200)
bool tc_shape_ok(const admm_problem* p) {
  return p->H % 64 == 0 && p->H >= 64 && p->ldn % 128 == 0 && (int64_t)(p->T + 1) * p->H * p->ldn < ((int64_t)1 << 32);
}
}  // namespace

bool tc_eligible(const admm_problem* p) { return tc_shape_ok(p) && get_encode() != nullptr; }

int64_t tc_workspace_bytes(const admm_problem* p) {
  if (!tc_shape_ok(p)) return 0;
  return ws_layout(p).total * 4;
}

namespace {
// max |a - b| (b may be null) -> *slot, then the fp16 pair of the scaled values
int prep_operand(int kind, const float* a, const float* b, __half* hi, __half* lo, int64_t n, TcMeta* meta, unsigned* slot,
                 cudaStream_t st) {
  KernelScope ks_("tc_prep_operand", st);
  if (slot) {
    if (cudaMemsetAsync(slot, 0, sizeof(unsigned), st) != cudaSuccess) return check_launch("prep memset");
    absmax_kernel<<<prep_grid(n), 256, 0, st>>>(a, b, n, slot);
    count_launch();
  }
  prep_f16_kernel<<<prep_grid(n), 256, 0, st>>>(kind, a, b, hi, lo, n, meta);
  count_launch();
  return check_launch("prep_f16");
}
}  // namespace

int tc_refresh_weights(const admm_problem* p, cudaStream_t st) {
  const WsLayout w = ws_layout(p);
  TcMeta* meta = ws_meta(p);
  const int64_t nx = 4LL * p->D * p->H, nh = 4LL * p->H * p->H;
  KernelScope ks_("tc_refresh_weights", st);
  // both maxima first: the common accumulator scale depends on both
  if (cudaMemsetAsync(&meta->max_wx, 0, 2 * sizeof(unsigned), st) != cudaSuccess) return check_launch("prep memset");
  absmax_kernel<<<prep_grid(nx), 256, 0, st>>>(p->wx, nullptr, nx, &meta->max_wx);
  absmax_kernel<<<prep_grid(nh), 256, 0, st>>>(p->wh, nullptr, nh, &meta->max_wh);
  count_launch(2);
  int rc = prep_operand(PREP_WX, p->wx, nullptr, ws_half(p, w.wx_hi), ws_half(p, w.wx_lo), nx, meta, nullptr, st);
  if (rc) return rc;
  return prep_operand(PREP_WH, p->wh, nullptr, ws_half(p, w.wh_hi), ws_half(p, w.wh_lo), nh, meta, nullptr, st);
}

int tc_refresh_inputs(const admm_problem* p, cudaStream_t st) {
  const WsLayout w = ws_layout(p);
  float* ws = (float*)p->tc_ws;
  TcMeta* meta = ws_meta(p);
  const int64_t nx = (int64_t)p->T * p->D * p->ldn;
  KernelScope ks_("tc_refresh_inputs", st);
  split_trunc_kernel<<<(unsigned)((nx / 4 + 255) / 256 + 1), 256, 0, st>>>(p->x, ws + w.x_lo, nx);
  count_launch();
  int rc = prep_operand(PREP_X, p->x, nullptr, ws_half(p, w.x16_hi), ws_half(p, w.x16_lo), nx, meta, &meta->max_x, st);
  if (rc) return rc;
  return tc_refresh_weights(p, st);        // the weights' scales are tied to the scale of x
}

int tc_refresh_state(const admm_problem* p, cudaStream_t st) {
  const WsLayout w = ws_layout(p);
  const int64_t nh = (int64_t)(p->T + 1) * p->H * p->ldn;
  return prep_operand(PREP_H, p->gate[5], nullptr, ws_half(p, w.h16_hi), ws_half(p, w.h16_lo), nh, ws_meta(p), nullptr, st);
}

// operand slot <- fp16 pair of (x2g - wx_prev): what the h-phase adds to the stored pre-activations
int tc_refresh_wx_delta(const admm_problem* p, cudaStream_t st) {
  const WsLayout w = ws_layout(p);
  TcMeta* meta = ws_meta(p);
  const int64_t n = 4LL * p->D * p->H;
  return prep_operand(PREP_DELTA, p->wx, p->wx_prev, ws_half(p, w.g_hi), ws_half(p, w.g_lo), n, meta, &meta->max_d, st);
}

int tc_refresh_grad(const admm_problem* p, int src, const float* grad, cudaStream_t st) {
  const WsLayout w = ws_layout(p);
  TcMeta* meta = ws_meta(p);
  const int64_t n = 4LL * (src == ADMM_SRC_X ? p->D : p->H) * p->H;
  return prep_operand(src == ADMM_SRC_X ? PREP_GRAD_X : PREP_GRAD_H, grad, nullptr, ws_half(p, w.g_hi), ws_half(p, w.g_lo), n,
                      meta, &meta->max_g, st);
}

unsigned* tc_r_bound(const admm_problem* p) { return &ws_meta(p)->r_bound; }
unsigned* tc_x_bound(const admm_problem* p) { return ws_meta(p)->x_bound; }
unsigned* tc_h_overflow(const admm_problem* p) { return &ws_meta(p)->h_overflow; }
int32_t* tc_z_dirty(const admm_problem* p) { return &ws_meta(p)->z_dirty; }

namespace {
// Install new inputs: src [n][cols] (sample-major, as the reference holds x [N, T, D] and y [N, O]) -> dst [cols][ldn]
// (feature-major), tiled transpose through shared memory; *changed is set when any value differs bitwise from what dst held.
__global__ void load_inputs_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n, int64_t cols, int64_t ldn,
                                   int32_t* changed) {
  __shared__ float tile[32][33];
  const int64_t n0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  bool diff = false;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t nn = n0 + r, cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (nn < n && cc < cols) ? src[nn * cols + cc] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t cc = c0 + r, nn = n0 + threadIdx.x;
    if (cc < cols && nn < n) {
      const float v = tile[threadIdx.x][r];
      float* d = dst + cc * ldn + nn;
      if (__float_as_uint(*d) != __float_as_uint(v)) { diff = true; *d = v; }
    }
  }
  if (__any_sync(0xffffffffu, diff) && threadIdx.x == 0 && changed) *changed = 1;
}
}  // namespace

int tc_load_inputs(float* dst, const float* src, int64_t n, int64_t cols, int64_t ldn, int32_t* changed, cudaStream_t st) {
  KernelScope ks_("load_inputs_kernel", st);
  dim3 grid((unsigned)((n + 31) / 32), (unsigned)((cols + 31) / 32));
  load_inputs_kernel<<<grid, dim3(32, 8), 0, st>>>(src, dst, n, cols, ldn, changed);
  count_launch();
  return check_launch("load_inputs");
}

namespace {
// max over t = 1..T, units and samples of 1 + |lambda_g/rho_g| + |gate_g| (g = i,f,g,o) -> out (bit pattern, atomicMax)
__global__ void state_bound_kernel(const float* __restrict__ gate, const float* __restrict__ dual, float inv_rho, int64_t n,
                                   unsigned* out) {
  float m = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fmaf(fabsf(dual[i]), inv_rho, fabsf(gate[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(1.0f + m));
}
__global__ void set_u32_kernel(unsigned* dst, unsigned v0, unsigned v1) { dst[0] = v0; dst[1] = v1; }
}  // namespace

// x_bound from the state as it is (after the caller wrote gates / duals directly: admm_tc_refresh(ADMM_TC_STATE))
int tc_refresh_bound(const admm_problem* p, cudaStream_t st) {
  KernelScope ks_("tc_refresh_bound", st);
  unsigned* xb = tc_x_bound(p);
  if (!p->dual[0] || !p->gate[0]) return ADMM_OK;          // ADMM-LSTM-L problems carry no lambda_g here and track their own bound
  if (cudaMemsetAsync(xb, 0, 2 * sizeof(unsigned), st) != cudaSuccess) return check_launch("x_bound memset");
  const int64_t slab = (int64_t)p->H * p->ldn, n = (int64_t)p->T * slab;
  for (int g = 0; g < 4; ++g) {
    state_bound_kernel<<<prep_grid(n), 256, 0, st>>>(p->gate[g] + slab, p->dual[g] + slab, 1.0f / p->hp.rho[g], n, xb);
    count_launch();
  }
  return check_launch("state_bound");
}
// x_bound of the forward initialisation: lambda = 0 and |gate| < 1  ->  2
int tc_set_bound(const admm_problem* p, float v, cudaStream_t st) {
  unsigned bits;
  memcpy(&bits, &v, sizeof(bits));
  set_u32_kernel<<<1, 1, 0, st>>>(tc_x_bound(p), bits, bits);
  count_launch();
  return check_launch("set_bound");
}

void tc_h16(const admm_problem* p, __half** hi, __half** lo) {
  const WsLayout w = ws_layout(p);
  *hi = ws_half(p, w.h16_hi);
  *lo = ws_half(p, w.h16_lo);
}

int gate_gemm_tc(int mode, const admm_problem* p, const GateGemmArgs& a_in, int tc, cudaStream_t st) {
  GateGemmArgs a = a_in;
  TcMaps maps;
  // the epilogue addresses state, zstore and scratch with 32-bit element offsets from the launch's first slab
  if ((int64_t)(p->T + 1) * p->H * p->ldn >= ((int64_t)1 << 32) || (int64_t)tc * p->H * p->ldn >= ((int64_t)1 << 32)) {
    set_error("gate_gemm_tc: a state tensor of %lld elements exceeds the 32-bit offsets of the epilogue",
              (long long)((int64_t)(p->T + 1) * p->H * p->ldn));
    return ADMM_EINVAL;
  }
  // the "gradient" operand slot holds G of the probed weight, or, for the zstore refresh of the h-phase, W_new - W_old of x2g
  const bool z_refresh = (mode == GG_GRAD && a.z_accumulate);
  int rc = get_maps(p, z_refresh ? ADMM_SRC_X : a.src, &maps);
  if (rc) return rc;
  // slab index of h_{t-1} / x_t of the first timestep in the launch
  const int64_t slab_elems = (int64_t)p->H * p->ldn;
  const int slab0 = (int)((a.h_prev - p->gate[5]) / slab_elems);
  {
    __half *hi, *lo;
    tc_h16(p, &hi, &lo);
    a.h16_hi = hi + (a.gate[5] - p->gate[5]);
    a.h16_lo = lo + (a.gate[5] - p->gate[5]);
  }
  const TcMeta* meta = ws_meta(p);
  a.acc_scale = z_refresh ? &meta->scale_d : &meta->scale_z;
  a.h_ovf = &ws_meta(p)->h_overflow;
  static const int tma_hint = [] {              // L2 evict_last hint on the weight-operand TMA loads (A/B switch)
    const char* e = getenv("ADMM_TMA_HINT");
    return e ? atoi(e) : 1;
  }();
  a.tma_hint = tma_hint;
  static const int epi_prefetch = [] {          // on unless ADMM_EPI_PREFETCH=0 (A/B switch for measurements)
    const char* e = getenv("ADMM_EPI_PREFETCH");
    return e ? atoi(e) : 1;
  }();
  a.epi_prefetch = epi_prefetch;
  // Staged epilogue inputs (A/B switches for measurements).  GRAD (K = D: the epilogue IS the kernel): on, 41.9 -> 29.7 ms per
  // cfg3 step, 89 % of the copy bandwidth.  SWEEP at H = 1024: off -- there the full-K operand stream (3.3 GB per launch from
  // L2) plus the state traffic sits at the L2 -> SM delivery limit either way (0.49 ms staged with a 3-stage operand ring,
  // 0.49 ms unstaged with 4 stages; profiles/r02_ncu_sweep_staged_vs_unstaged.txt), so the deeper operand ring stays.
  static const int epi_staged = [] {
    const char* e = getenv("ADMM_EPI_STAGED");
    return e ? atoi(e) : 1;
  }();
  static const int epi_staged_sweep_env = [] {       // -1: by shape
    const char* e = getenv("ADMM_EPI_STAGED_SWEEP");
    return e ? atoi(e) : -1;
  }();
  // by shape: with D + H <= 768 the operand stream is small next to the state traffic and the staged epilogue wins
  // (cfg2 0.894 -> 0.774 ms per launch = 0.69 of HBM, cfg4 0.483 -> 0.409 ms = 0.66; profiles/r02_p_staged_sweep_by_shape.txt)
  const int epi_staged_sweep = epi_staged_sweep_env >= 0 ? epi_staged_sweep_env : (p->D + p->H <= 768);
  const int nkx = (p->D + BK - 1) / BK, nkh = (p->H + BK - 1) / BK;
  const TcRange full{0, nkx + nkh, 0};
  switch (mode) {
    case GG_FORWARD: return launch_tc<GG_FORWARD>(p, a, maps, slab0, tc, full, st, "gate_gemm_tc<FORWARD>");
    case GG_SWEEP:
      if (epi_staged_sweep) return launch_tc<GG_SWEEP, true>(p, a, maps, slab0, tc, full, st, "gate_gemm_tc<SWEEP>");
      return launch_tc<GG_SWEEP>(p, a, maps, slab0, tc, full, st, "gate_gemm_tc<SWEEP>");
    case GG_GRAD:
      if (z_refresh) {
        if (epi_staged) return launch_tc<GG_GRAD, true>(p, a, maps, slab0, tc, TcRange{0, nkx, 1}, st, "gate_gemm_tc<GRAD:z+=x*dW>");
        return launch_tc<GG_GRAD>(p, a, maps, slab0, tc, TcRange{0, nkx, 1}, st, "gate_gemm_tc<GRAD:z+=x*dW>");
      }
      if (epi_staged) return launch_tc<GG_GRAD, true>(p, a, maps, slab0, tc, full, st, "gate_gemm_tc<GRAD:full>");
      return launch_tc<GG_GRAD>(p, a, maps, slab0, tc, full, st, "gate_gemm_tc<GRAD:full>");
    case GG_RAWZ: return launch_tc<GG_RAWZ>(p, a, maps, slab0, tc, full, st, "gate_gemm_tc<RAWZ:z>");
    case GG_MOMENTS: {
      // Q = A_src G in TMEM, moment sums in the epilogue (needs the z store as Z0)
      GateGemmArgs q = a;
      q.acc_scale = &meta->scale_q;
      const TcRange qr = (a.src == ADMM_SRC_X) ? TcRange{0, nkx, 1} : TcRange{nkx, nkx + nkh, 1};
      return launch_tc<GG_MOMENTS, true>(p, q, maps, slab0, tc, qr, st,
                                         a.src == ADMM_SRC_X ? "gate_gemm_tc<MOMENTS:Q=x*G>" : "gate_gemm_tc<MOMENTS:Q=h*G>");
    }
    case GG_PROBE: {
      // Z0 = x W + h U, then Q = A_src G: two launches of the same kernel (each keeps the 64-unit tile and
      // two resident CTAs per SM; a fused Z0|Q tile needs all 512 TMEM columns and halves the tile width)
      if (!a.zstore) {
        rc = launch_tc<GG_RAWZ>(p, a, maps, slab0, tc, full, st, "gate_gemm_tc<RAWZ:z>");
        if (rc) return rc;
      }
      GateGemmArgs q = a;
      q.scratch = a.scratch_q;
      q.acc_scale = &meta->scale_q;
      const TcRange qr = (a.src == ADMM_SRC_X) ? TcRange{0, nkx, 1} : TcRange{nkx, nkx + nkh, 1};
      return launch_tc<GG_RAWZ>(p, q, maps, slab0, tc, qr, st, a.src == ADMM_SRC_X ? "gate_gemm_tc<RAWZ:Q=x*G>" : "gate_gemm_tc<RAWZ:Q=h*G>");
    }
  }
  set_error("gate_gemm_tc: bad mode %d", mode);
  return ADMM_EINVAL;
}

namespace {
// ------------------------------------------------------------------------------------------------------------
// G^T tile on the tensor cores (admm.py:308-311):   G_acc[g][k][j] += sum_{t,n} R[(g,j)][t][n] * h_{t-1}[k][n]
// Both operands are contiguous along the reduction index n, i.e. K-major for UMMA: plain SWIZZLE_128B boxes of
// 32 samples.  D rows (TMEM lanes) = 128 consecutive (gate, unit) columns of G, D columns = NT rows k of G, so the
// epilogue's accumulation into the fp64 G buffer is coalesced along j.  3xTF32: R_hi*h_hi + R_lo*h_hi + R_hi*h_lo.
constexpr int ATR_BKN = 32;           // tf32: samples per pipeline stage (one 128-byte swizzle row); fp16 pairs: 64
constexpr int ATR_STAGES = 2;
constexpr int ATR_THREADS = 192;       // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue

struct AtrMaps { CUtensorMap r, r_lo, h, h_lo; };

__device__ __forceinline__ uint64_t make_desc_k128(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_kmajor(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__host__ __device__ constexpr uint32_t make_idesc_kmajor_f16(int m, int n) {      // kind::f16, A = B = F16, D = F32
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// F16: operands are fp16 pairs (R 2^sR from the GRAD epilogue / packing kernel, x 2^sx or h 2^11 from the workspace),
// 64 samples per 128-byte row, kind::f16 MMAs with K = 16; the accumulator is rescaled by 2^-(sR + sa) in the epilogue.
// Same bytes per stage as the tf32 variant for twice the samples: the kernel is bound by operand delivery.
template <int NT_, bool F16>
__global__ void __launch_bounds__(ATR_THREADS, 1)
atr_tc_kernel(const __grid_constant__ AtrMaps maps, double* g_acc, int K, int H, int tc, int64_t ldn, int slab0,
              int chunks_per_cta, int n_chunks, int use_atomics, const unsigned* r_bound, const unsigned* a_max) {
  // H = rows per group of R (AtrArgs::rpg); a_max = max|x| bits when A_src = x (scale 2^cap), nullptr for h (2^SCALE_H)
  constexpr int BKN = F16 ? 64 : ATR_BKN;
  constexpr int A_BYTES = 128 * 128;                // R tile   128 rows x 128 B
  constexpr int B_BYTES = NT_ * 128;                // h tile   NT rows x 128 B
  constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[ATR_STAGES], empty_bar[ATR_STAGES], acc_bar;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = blockIdx.x * 128;          // first (gate, unit) column of G handled here
  const int k0 = blockIdx.y * NT_;          // first row k of G
  const int ch_begin = blockIdx.z * chunks_per_cta;
  const int ch_end = min(ch_begin + chunks_per_cta, n_chunks);
  if (ch_begin >= ch_end) return;
  const int chunks_per_t = (int)(ldn / BKN);
  constexpr int TCOLS = NT_ < 32 ? 32 : NT_;

  if (threadIdx.x == 0) {
    for (int s = 0; s < ATR_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, TCOLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int li = ch - ch_begin, s = li % ATR_STAGES, it = li / ATR_STAGES;
        if (it > 0) mbar_wait(&empty_bar[s], (it - 1) & 1);
        const int tl = ch / chunks_per_t, n0 = (ch % chunks_per_t) * BKN;
        uint8_t* st = smem + s * STAGE_BYTES;
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_3d(st, &maps.r, &full_bar[s], n0, tl, c0);
        tma_load_3d(st + A_BYTES, &maps.r_lo, &full_bar[s], n0, tl, c0);
        tma_load_3d(st + 2 * A_BYTES, &maps.h, &full_bar[s], n0, k0, slab0 + tl);
        tma_load_3d(st + 2 * A_BYTES + B_BYTES, &maps.h_lo, &full_bar[s], n0, k0, slab0 + tl);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = F16 ? make_idesc_kmajor_f16(128, NT_) : make_idesc_kmajor(128, NT_);
      for (int ch = ch_begin; ch < ch_end; ++ch) {
        const int li = ch - ch_begin, s = li % ATR_STAGES, it = li / ATR_STAGES;
        mbar_wait(&full_bar[s], it & 1);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t r_hi = st, r_lo = st + A_BYTES, h_hi = st + 2 * A_BYTES, h_lo = h_hi + B_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t off = ks * 32;                 // 8 floats / 16 halfs inside the 128-byte swizzle row
          const uint32_t acc0 = (li > 0 || ks > 0) ? 1u : 0u;
          if (F16) {
            umma_f16(tmem_base, make_desc_k128(r_hi + off), make_desc_k128(h_hi + off), idesc, acc0);
            umma_f16(tmem_base, make_desc_k128(r_lo + off), make_desc_k128(h_hi + off), idesc, 1u);
            umma_f16(tmem_base, make_desc_k128(r_hi + off), make_desc_k128(h_lo + off), idesc, 1u);
          } else {
            umma_tf32(tmem_base, make_desc_k128(r_hi + off), make_desc_k128(h_hi + off), idesc, acc0);
            umma_tf32(tmem_base, make_desc_k128(r_lo + off), make_desc_k128(h_hi + off), idesc, 1u);
            umma_tf32(tmem_base, make_desc_k128(r_hi + off), make_desc_k128(h_lo + off), idesc, 1u);
          }
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else {
    const int quarter = warp & 3;
    const int c = c0 + quarter * 32 + lane;            // (gate, unit) column of G owned by this thread
    const int g = c / H, j = c % H;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    double out_scale = 1.0;
    if (F16) out_scale = (double)ldexpf(1.0f, -(cap_exp(*r_bound) + (a_max ? cap_exp(*a_max) : SCALE_H)));
    mbar_wait(&acc_bar, 0);
    tc_fence_after();
    for (int kb = 0; kb < NT_; kb += 8) {
      float v[8];
      tmem_ld8(t_row + kb, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = k0 + kb + e;
        if (k < K) {
          double* dst = g_acc + ((int64_t)g * K + k) * H + j;
          const double val = (double)v[e] * out_scale;
          if (use_atomics) atomicAdd(dst, val);
          else *dst += val;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TCOLS);
  }
}

int make_map_box(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_elems,
                 uint64_t stride2_elems, uint32_t b0, uint32_t b1, uint32_t b2, bool f16 = false) {
  EncodeFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return ADMM_ECUDA; }
  const uint64_t eb = f16 ? 2 : 4;
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_elems * eb, stride2_elems * eb};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base),
                         dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (atr) failed (%d)", (int)r); return ADMM_ECUDA; }
  return ADMM_OK;
}

template <int NT_, bool F16>
int launch_atr(const admm_problem* p, const AtrArgs& a, int slab0, bool src_is_x, cudaStream_t st) {
  constexpr int SMEM = ATR_STAGES * (2 * 128 * 128 + 2 * NT_ * 128) + 1024;
  constexpr int BKN = F16 ? 64 : ATR_BKN;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(atr_tc_kernel<NT_, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    cudaFuncSetAttribute(atr_tc_kernel<NT_, F16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    configured = true;
  }
  AtrMaps m;
  const WsLayout w = ws_layout(p);
  const uint64_t ldn = a.ldn, H = a.H, tc = a.tc, T1 = p->T + 1;
  const uint64_t rows = a.rows ? a.rows : 4 * H;
  const int rpg = a.rpg ? a.rpg : a.H;
  int rc = 0;
  rc |= make_map_box(&m.r, F16 ? (const void*)a.r16_hi : (const void*)a.scratch, ldn, tc, rows, ldn, tc * ldn, BKN, 1, 128, F16);
  rc |= make_map_box(&m.r_lo, F16 ? (const void*)a.r16_lo : (const void*)a.scratch_lo, ldn, tc, rows, ldn, tc * ldn, BKN, 1, 128, F16);
  if (src_is_x) {     // A_src = x: [T][D][ldn]; rows k >= D are zero-filled by TMA and masked in the epilogue
    const uint64_t D = p->D;
    rc |= make_map_box(&m.h, F16 ? (const void*)ws_half(p, w.x16_hi) : (const void*)p->x, ldn, D, p->T, ldn, D * ldn, BKN, NT_, 1, F16);
    rc |= make_map_box(&m.h_lo, F16 ? (const void*)ws_half(p, w.x16_lo) : (const void*)((float*)p->tc_ws + w.x_lo), ldn, D, p->T,
                       ldn, D * ldn, BKN, NT_, 1, F16);
  } else {
    rc |= make_map_box(&m.h, F16 ? (const void*)ws_half(p, w.h16_hi) : (const void*)p->gate[5], ldn, H, T1, ldn, H * ldn, BKN, NT_, 1, F16);
    rc |= make_map_box(&m.h_lo, F16 ? (const void*)ws_half(p, w.h16_lo) : (const void*)nullptr, ldn, H, T1, ldn, H * ldn, BKN,
                       NT_, 1, F16);
  }
  if (rc) return ADMM_ECUDA;
  const int n_chunks = (int)(tc * (ldn / BKN));
  const int tiles = (int)((rows / 128) * ((a.K + NT_ - 1) / NT_));
  // one wave: 148 SMs x the CTAs that fit next to each other (NT = 64: two 97 KB rings per SM -- the K = D pass is bound by
  // HBM, and 32 tiles x 4 splits left 20 SMs without a CTA and 96 KB per SM in flight)
  constexpr int CTAS_PER_SM = (2 * SMEM <= 227 * 1024) ? 2 : 1;
  constexpr int WAVE = 148 * CTAS_PER_SM;
  int splits = (WAVE + tiles - 1) / tiles;
  if (tiles * splits > WAVE && splits > 1) --splits;
  splits = max(1, min(splits, n_chunks));
  const int cpc = (n_chunks + splits - 1) / splits;
  splits = (n_chunks + cpc - 1) / cpc;
  dim3 grid((unsigned)(rows / 128), (unsigned)((a.K + NT_ - 1) / NT_), (unsigned)splits);
  const unsigned* a_max = src_is_x ? &ws_meta(p)->max_x : nullptr;
  KernelScope ks_(F16 ? (NT_ == 256 ? "atr_tc<256,f16>" : NT_ == 128 ? "atr_tc<128,f16>" : "atr_tc<64,f16>")
                      : (NT_ == 256 ? "atr_tc<256,tf32>" : NT_ == 128 ? "atr_tc<128,tf32>" : "atr_tc<64,tf32>"), st);
  atr_tc_kernel<NT_, F16><<<grid, ATR_THREADS, SMEM, st>>>(m, a.g_acc, a.K, rpg, a.tc, a.ldn, slab0, cpc, n_chunks, splits > 1,
                                                           a.r_bound, a_max);
  count_launch();
  return check_launch("atr_tc");
}

template <int NT_>
int launch_atr_any(const admm_problem* p, const AtrArgs& a, int slab0, bool src_is_x, cudaStream_t st) {
  if (a.r16_hi) return launch_atr<NT_, true>(p, a, slab0, src_is_x, st);
  return launch_atr<NT_, false>(p, a, slab0, src_is_x, st);
}

}  // namespace

// Tensor-core G += A_src^T R.  `a.a_src` must be a slab of p->gate[5] (src = h, K = H) or of p->x (src = x, K = D).
int atr_tc(const admm_problem* p, const AtrArgs& a, cudaStream_t st) {
  if ((!a.scratch_lo && !a.r16_hi) || ((a.rows ? a.rows : 4 * a.H) % 128) != 0) return atr_simt(a, st);
  const bool src_is_x = (a.a_src >= p->x && a.a_src < p->x + (int64_t)p->T * p->D * p->ldn);
  if (src_is_x) {
    if (p->D < 8) return atr_simt(a, st);               // a handful of rows: not worth a 64-row MMA tile
    const int slab0 = (int)((a.a_src - p->x) / ((int64_t)p->D * p->ldn));
    if (a.K > 128) return launch_atr_any<256>(p, a, slab0, true, st);
    if (a.K > 64) return launch_atr_any<128>(p, a, slab0, true, st);
    return launch_atr_any<64>(p, a, slab0, true, st);
  }
  if (!a.r16_hi) return atr_simt(a, st);       // A_src = h exists on the tensor-core path only as fp16 pairs
  const int slab0 = (int)((a.a_src - p->gate[5]) / ((int64_t)p->H * p->ldn));
  if (p->H >= 256) return launch_atr_any<256>(p, a, slab0, false, st);
  if (p->H >= 128) return launch_atr_any<128>(p, a, slab0, false, st);
  return launch_atr_any<64>(p, a, slab0, false, st);
}

}  // namespace admm
