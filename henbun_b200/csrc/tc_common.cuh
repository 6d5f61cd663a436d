// Shared pieces of the tcgen05 engines (gemm_tc2.cu: in-kernel hi/lo split of fp32 operands; gemm_h2.cu: pre-split fp16
// hi/lo operands): mbarrier / TMA / tcgen05 PTX wrappers, UMMA shared-memory descriptors, the epilogue tile store.
#pragma once
#include "gemm.cuh"
#include "kernels.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace hb {
namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NSTAGE = 4;
constexpr int CHK = 8;                      // k-blocks per tensor-core accumulation chunk (K = 128)
constexpr int NTHREADS = 512;
constexpr int A_TILE = BM * BK * 4;         // 8 KB

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

// Shared-memory matrix descriptors (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version = 1 | [61,64) layout (2 = SW128, 4 = SW64)
// K-major, SWIZZLE_64B : rows of 16 floats (64 B); 8-row atoms of 512 B -> SBO = 512; LBO unused.
// MN-major tf32 has exactly one legal swizzled layout, SWIZZLE_128B_BASE32B (layout type 1; TMA name
// SWIZZLE_128B_ATOM_32B): rows of 32 m (128 B), atoms of 4 k-rows (512 B) in which the 32-byte granules of a row
// are XORed with (k row & 3).  Boxes of 32 m x 16 k -> SBO = 512 (next 4 k), LBO = 2048 (next 32 m = next box).
template <bool KMAJOR>
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  if (KMAJOR) {
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)4 << 61;
  } else {
    d |= (uint64_t)(2048 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 61;
  }
  d |= (uint64_t)1 << 46;
  return d;
}

// bf16 operand tiles of the cross products.  K-major: SWIZZLE_32B (layout 6), 8-row atoms of 256 B.
// MN-major: SWIZZLE_64B (layout 4), atoms of 8 k-rows x 64 B = 512 B (SBO), 32-row groups 1024 B apart (LBO).
template <bool KMAJOR>
__device__ __forceinline__ uint64_t make_desc_bf16(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  if (KMAJOR) {
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)6 << 61;
  } else {
    d |= (uint64_t)(1024 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)4 << 61;
  }
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}

struct Tc2Params {
  float* C;
  long long ldc;
  int M, N, K;
  float alpha, beta;
  int c_tri;
  int a_mode, b_mode;   // triangular structure in (row, k) space: 1 k<=r, 2 k>=r, 3 k>r, 4 k<r
  int tiles_m, tiles_n;
  int vecC;
  int ksplit;             // k-blocks per split (blockIdx.y = split index); 0 = no split
  long long csplit;       // element stride between the partial outputs of consecutive splits
  const float* bias;      // optional [N], added before the activation (MatBias: Henbun/nn.py:31-32)
  int act, clip;
  float clip_lo, clip_hi;
};

__device__ __forceinline__ float tc2_act(float x, int act, int clip, float lo, float hi) {
  if (clip) x = fminf(fmaxf(x, lo), hi);
  switch (act) {
    case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

// Epilogue of one output tile.  Each epilogue thread holds one accumulator ROW (TMEM lane) -- storing from that
// layout writes 16 bytes per row per instruction, 32 rows (ldc apart) per warp: half-used sectors and no DRAM page
// locality, which dominated short-K products (M=65280, N=K=256: 153 us against ~30 us of HBM time).  So the tile is
// staged through the (now idle) pipeline shared memory and written with one 512-byte contiguous row segment per
// warp instruction; beta*C is read with the same pattern.  ew = epilogue warp 0..7, q = ew & 3, half = ew >> 2.
template <int BN>
__device__ __forceinline__ void store_tile(const float (&acc)[BN / 2], const Tc2Params& p, float alpha, float* Cbase,
                                           uint32_t sm_tile, int m0, int n0, int ew, int lane) {
  constexpr int HALF = BN / 2;
  constexpr int LDT = BN + 4;                       // padded row (floats)
  const int q = ew & 3, half = ew >> 2;
  {
    const uint32_t row = sm_tile + (uint32_t)((q * 32 + lane) * LDT + half * HALF) * 4u;
#pragma unroll
    for (int v = 0; v < HALF / 4; ++v)
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row + v * 16), "f"(acc[4 * v]), "f"(acc[4 * v + 1]),
                   "f"(acc[4 * v + 2]), "f"(acc[4 * v + 3]) : "memory");
  }
  asm volatile("bar.sync 1, 256;" ::: "memory");     // the 8 epilogue warps
  const bool post = (p.bias != nullptr) || p.act != ACT_NONE || p.clip;
  // rows ew, ew+8, ...: RB rows per trip so that RB * (BN/128) reads of the old C are in flight per lane (one read per
  // trip made the beta = 1 epilogue latency-bound: 16 dependent DRAM round trips per warp)
  constexpr int RB = 4, SEGS = (BN + 127) / 128;       // BN = 64: one segment, lanes 16..31 idle
#pragma unroll 1
  for (int r0 = ew; r0 < BM; r0 += 8 * RB) {
    float4 oldv[RB][SEGS];
    const bool fastC = p.vecC && p.beta != 0.f;
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int gi = m0 + r0 + 8 * b;
#pragma unroll
      for (int seg = 0; seg < SEGS; ++seg) {
        const int gj = n0 + seg * 128 + lane * 4;
        oldv[b][seg] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (fastC && seg * 128 + lane * 4 < BN && r0 + 8 * b < BM && gi < p.M && gj + 3 < p.N && !(p.c_tri == 1 && gj + 3 > gi))
          oldv[b][seg] = *reinterpret_cast<const float4*>(Cbase + (long long)gi * p.ldc + gj);
      }
    }
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int r = r0 + 8 * b;
      const int gi = m0 + r;
      if (r >= BM || gi >= p.M) continue;
      float* crow = Cbase + (long long)gi * p.ldc;
#pragma unroll
      for (int seg = 0; seg < SEGS; ++seg) {
        const int cj = seg * 128 + lane * 4;
        const int gj = n0 + cj;
        if (cj >= BN || gj >= p.N) continue;
        float o[4];
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
                     : "r"(sm_tile + (uint32_t)(r * LDT + cj) * 4u));
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] *= alpha;
        const bool full = (gj + 3 < p.N) && !(p.c_tri == 1 && gj + 3 > gi);
        if (full && p.vecC) {
          if (p.beta != 0.f) {
            const float4 old = oldv[b][seg];
            o[0] = fmaf(p.beta, old.x, o[0]); o[1] = fmaf(p.beta, old.y, o[1]);
            o[2] = fmaf(p.beta, old.z, o[2]); o[3] = fmaf(p.beta, old.w, o[3]);
          }
          if (post) {
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = tc2_act(o[e] + (p.bias ? __ldg(p.bias + gj + e) : 0.f), p.act, p.clip, p.clip_lo, p.clip_hi);
          }
          *reinterpret_cast<float4*>(crow + gj) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (gj + e < p.N && !(p.c_tri == 1 && gj + e > gi)) {
              float x = o[e];
              if (p.beta != 0.f) x = fmaf(p.beta, crow[gj + e], x);
              if (post) x = tc2_act(x + (p.bias ? __ldg(p.bias + gj + e) : 0.f), p.act, p.clip, p.clip_lo, p.clip_hi);
              crow[gj + e] = x;
            }
          }
        }
      }
    }
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `rank` of the cluster.  Default (.release.cta)
// semantics on purpose: a .release.cluster arrive compiles to MEMBAR.ALL.GPU + CCTL.IVALL per call, which made the
// converter warps the bottleneck (131 vs 200 TFLOP/s).  The data it publishes is shared memory written before a
// fence.proxy.async by the same warp, or TMEM reads completed by tcgen05.wait::ld.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode2() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}


inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace
}  // namespace hb
