// tcgen05 GEMM engine for PRE-SPLIT operands:  C[MxN] = alpha * sA * sB * op(A) op(B) + beta * C   (fp32 out)
//
// gemm_tc2.cu splits every fp32 operand tile into hi/lo parts inside the kernel (converter warps).  ncu named that
// the limiter of the factorisations' big products: 80 KB of shared-memory traffic per 16-wide k-block against 64 KB/512 clk
// of tensor work (LSU shared wavefronts 46 %, tensor pipe 65 %; profiles/README.md).  In a blocked Cholesky and its
// reverse mode every operand of a big product is a FINISHED panel of L or of K-bar that is read by many later
// products, so the split is done ONCE per element when the panel is finished (split_f16_kernel, 8 bytes of traffic per
// element) into a "shadow" matrix of fp16 pairs
//        x * s = hi + lo + O(2^-23 |x s|),   hi = rn_f16(x s),  lo = rn_f16(x s - hi),   s a power of two,
// and this kernel streams the shadows with TMA straight into UMMA operand layout: no converter warps, no generic-proxy
// shared-memory traffic at all, and three kind::f16 MMAs per 16-wide k-step (hi*hi + lo*hi + hi*lo; the dropped lo*lo
// is 2^-22 relative) where the tf32 + bf16 formulation needs four tensor slots.  22 mantissa bits per operand instead
// of 19 (tf32 hi + bf16 lo), so the products are also MORE accurate than gemm_tc2's.
// fp16's narrow exponent is handled by the scales: s is chosen from the exact absolute maximum of the panel (or an
// a-priori bound for L: |L_ij| <= sqrt(max K_ii)) so that |x s| <= 2^7; elements more than 2^21 below the maximum keep
// an ABSOLUTE error of 2^-25 / s, i.e. 2^-32 of the panel maximum -- below fp32 rounding of any sum they enter.
// Scales may differ per 128-column block of the A operand: along K they are applied by the epilogue warps when they
// fold a finished K = 128 accumulation chunk into the fp32 registers (the two-level accumulation the tensor core's
// truncating fp32 adder needs anyway), along M once per CTA.
//
// Structure: CTA pairs (cta_group::2), 256 x 256 output tile per pair, BK = 64 (128-byte rows, SWIZZLE_128B),
// 3-stage ring of {A hi, A lo, B-half hi, B-half lo} x 16 KB per CTA, both CTAs' TMA loads signal the LEADER's full
// barrier (cp.async.bulk.tensor ... cta_group::2), one elected thread issues 12 MMAs per stage.
// Warp roles: 0 TMA producer | 1 TMEM alloc + MMA issuer (leader CTA) | 8..15 epilogue.
#include "tc_common.cuh"
#include "gemm_h2.cuh"

namespace hb {

namespace {

constexpr int H_BK = 64;
constexpr int H_TILE = 128 * H_BK * 2;          // 16 KB: 128 rows x 64 k of fp16
// per CTA and stage: A hi | A lo (128 rows each) | B hi | B lo (BN / 2 rows each).  BN = 256: 64 KB x 3 stages;
// BN = 64 (few samples against a big operator: HBM-bound, deeper ring): 40 KB x 5 stages
template <int BN> struct H2Geom {
  static constexpr int B_TILE = (BN / 2) * H_BK * 2;
  static constexpr int STAGE = 2 * H_TILE + 2 * B_TILE;
  static constexpr int NSTAGE = (BN == 256) ? 3 : 5;
};
constexpr int H_CHS = 2;                        // stages per accumulation chunk (K = 128)
constexpr int H_OOB = 1 << 30;                  // a coordinate outside every tensor: TMA fills the box with zeros

// cp.async.bulk.tensor whose completion is signalled on a barrier that may live in the PEER CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// 16-bit operand tiles, SWIZZLE_128B (layout type 2), 8-row / 8-k atoms of 1024 B.
//   K-major : rows of 64 k (128 B); SBO = 1024 (next 8 rows); a K = 16 step advances the start address by 32 B.
//   MN-major: TMA boxes of {64 mn, 64 k}: 64 k-rows of 128 B; SBO = 1024 (next 8 k), LBO = 8192 (next 64 mn = next box);
//             a K = 16 step advances by two atoms = 2048 B.
template <bool KMAJOR>
__device__ __forceinline__ uint64_t make_desc_h(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)(KMAJOR ? 1 : (H_BK * 128 >> 4)) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct H2Params {
  Tc2Params t;
  int a_bmode;
  int group;               // pair-tile rows per raster group (tiles of a group share their B panel in L2)
  const float* a_inv;
  const float* a_kinv;
  const float* a_dinv;
  const float* a_minv;
  const float* b_inv;
};

template <int BN, bool AKM, bool BKM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1)
gemm_h2_pair_kernel(const __grid_constant__ CUtensorMap tmAh, const __grid_constant__ CUtensorMap tmAl,
                    const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmBl, const H2Params hp) {
  constexpr int PM = 256, HALF = BN / 2;
  constexpr int H_NSTAGE = H2Geom<BN>::NSTAGE, H_STAGE = H2Geom<BN>::STAGE, B_TILE = H2Geom<BN>::B_TILE;
  static_assert(BKM || BN == 256, "MN-major B operands need the 256-wide tile (64-element TMA boxes)");
  const Tc2Params& p = hp.t;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + H_NSTAGE * H_STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };                        // used in the leader: bytes of BOTH CTAs
  auto empty_bar = [&](int s) { return bars + 8u * (H_NSTAGE + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * H_NSTAGE + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * H_NSTAGE + 2 + b); }; // used in the leader: 8 local + 8 remote warps
  const uint32_t tmem_ptr_addr = bars + 8u * (2 * H_NSTAGE + 4);

  const uint32_t rank = cluster_ctarank();
  const int GROUP = hp.group;
  const int bid = blockIdx.x >> 1;
  const int per_group = GROUP * p.tiles_n;
  const int first_m = (bid / per_group) * GROUP;
  const int gsz = min(p.tiles_m - first_m, GROUP);
  const int tm = first_m + (bid % per_group) % gsz;
  const int tn = (bid % per_group) / gsz;
  const int pm0 = tm * PM, n0 = tn * BN;
  if (p.c_tri == 1 && n0 > pm0 + PM - 1) return;      // same decision in both CTAs
  const int m0 = pm0 + 128 * (int)rank;
  const int nb0 = n0 + (BN / 2) * (int)rank;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kb_lo = 0, kb_hi = (p.K + H_BK - 1) / H_BK;
  const int t0 = pm0 >> 7;                            // first 128-row block of the pair tile
  if (hp.a_bmode == 1) kb_hi = min(kb_hi, H_CHS * (t0 + 2));          // k-blocks <= the pair tile's second row block
  else if (hp.a_bmode == 2) kb_lo = min(kb_hi, H_CHS * (t0 + 1));     // k-blocks > the pair tile's first row block
  if (p.ksplit > 0) {                                 // split-K: blockIdx.y owns k-blocks [y * ksplit, (y + 1) * ksplit)
    kb_lo = max(kb_lo, (int)blockIdx.y * p.ksplit);
    kb_hi = min(kb_hi, ((int)blockIdx.y + 1) * p.ksplit);
  }
  const int num_k = max(kb_hi - kb_lo, 0);
  const int num_c = (num_k + H_CHS - 1) / H_CHS;
  float* const Cout = p.C + (long long)blockIdx.y * p.csplit;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < H_NSTAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"((uint32_t)(2 * BN)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp < 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0 && lane == 0) {          // ---------------- TMA producer: own A rows, own half of B ----------------
      const int rb = m0 >> 7;              // this CTA's 128-row block (block masks)
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = base + s * H_STAGE;
        const uint32_t fb = map_to_rank(full_bar(s), 0);
        if (rank == 0) mbar_expect_tx(full_bar(s), 2 * H_STAGE);
        const int k0 = (kb_lo + kb) * H_BK;
        const int kc = k0 >> 7;
        const bool zeroA = (hp.a_bmode == 1 && kc > rb) || (hp.a_bmode == 2 && !(kc > rb));
        if (AKM) {
          const int r = zeroA ? H_OOB : m0;
          tma_load_2d_pair(st, &tmAh, fb, k0, r);
          tma_load_2d_pair(st + H_TILE, &tmAl, fb, k0, r);
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = zeroA ? H_OOB : m0 + 64 * i;
            tma_load_2d_pair(st + i * (H_TILE / 2), &tmAh, fb, r, k0);
            tma_load_2d_pair(st + H_TILE + i * (H_TILE / 2), &tmAl, fb, r, k0);
          }
        }
        const uint32_t sb = st + 2 * H_TILE;
        if (BKM) {
          tma_load_2d_pair(sb, &tmBh, fb, k0, nb0);
          tma_load_2d_pair(sb + B_TILE, &tmBl, fb, k0, nb0);
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            tma_load_2d_pair(sb + i * (B_TILE / 2), &tmBh, fb, nb0 + 64 * i, k0);
            tma_load_2d_pair(sb + B_TILE + i * (B_TILE / 2), &tmBl, fb, nb0 + 64 * i, k0);
          }
        }
        if (++s == H_NSTAGE) { s = 0; ph ^= 1u; }
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {   // ---------------- MMA issuer (leader CTA only) ----------------
      // kind::f16, fp16 operands (format 0), fp32 accumulate: D fmt 1 << 4, major bits 15/16, N >> 3 at 17, M >> 4 at 24
      const uint32_t idesc = (1u << 4) | (AKM ? 0u : (1u << 15)) | (BKM ? 0u : (1u << 16)) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(PM >> 4) << 24);
      constexpr uint32_t ADV_A = AKM ? (32u >> 4) : (2048u >> 4);     // one K = 16 step, in 16-byte units
      constexpr uint32_t ADV_B = BKM ? (32u >> 4) : (2048u >> 4);
      int s = 0; uint32_t ph = 0;
      for (int c = 0; c < num_c; ++c) {
        const int buf = c & 1;
        mbar_wait(tempty_bar(buf), (uint32_t)(((c >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        const int kb_end = min(num_k, (c + 1) * H_CHS);
        for (int kb = c * H_CHS; kb < kb_end; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t st = base + s * H_STAGE;
          const uint64_t a_hi = make_desc_h<AKM>(st), a_lo = make_desc_h<AKM>(st + H_TILE);
          const uint64_t b_hi = make_desc_h<BKM>(st + 2 * H_TILE), b_lo = make_desc_h<BKM>(st + 2 * H_TILE + B_TILE);
#pragma unroll
          for (int k2 = 0; k2 < H_BK / 16; ++k2) {
            const uint64_t da = (uint64_t)(ADV_A * k2), db = (uint64_t)(ADV_B * k2);
            tc_mma_bf16_pair(d_tmem, a_lo + da, b_hi + db, idesc, (kb == c * H_CHS && k2 == 0) ? 0u : 1u);   // small terms first
            tc_mma_bf16_pair(d_tmem, a_hi + da, b_lo + db, idesc, 1u);
            tc_mma_bf16_pair(d_tmem, a_hi + da, b_hi + db, idesc, 1u);
          }
          tc_commit_pair(empty_bar(s));
          if (++s == H_NSTAGE) { s = 0; ph ^= 1u; }
        }
        tc_commit_pair(tfull_bar(buf));
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ---------------- epilogue: this CTA's 128 accumulator rows ----------------
    const int q = warp & 3;
    const int half = (warp - 8) >> 2;
    float acc[HALF];
#pragma unroll
    for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
    const int c_first = kb_lo / H_CHS;                 // 128-chunk index of the first chunk (kb_lo is chunk aligned)
    const int rbk = m0 >> 7;
    for (int c = 0; c < num_c; ++c) {
      const int buf = c & 1;
      float cs = 1.f;
      if (hp.a_kinv) cs = __ldg(((hp.a_bmode == 1 && c_first + c == rbk) ? hp.a_dinv : hp.a_kinv) + c_first + c);
      mbar_wait(tfull_bar(buf), (uint32_t)((c >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < HALF / 32; ++i) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * HALF + i * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
              "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
              "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
              "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[i * 32 + j] = fmaf(__uint_as_float(r[j]), cs, acc[i * 32 + j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_bar(buf), 0);
    }
    float alpha = p.alpha;
    if (hp.a_inv) alpha *= __ldg(hp.a_inv);
    if (hp.b_inv) alpha *= __ldg(hp.b_inv);
    if (hp.a_minv) alpha *= __ldg(hp.a_minv + (m0 >> 7));
    store_tile<BN>(acc, p, alpha, Cout, base, m0, n0, warp - 8, lane);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)));
  }
}

// Tensor map of one fp16 shadow operand.  K-major: stored [rows x K], box {64 k, 128 rows}; MN-major: stored [K x rows],
// box {64 rows, 64 k}.  SWIZZLE_128B both ways.
int make_map_h(CUtensorMap* map, const __half* ptr, long long rows, long long K, long long ld, bool kmajor, int box_rows = 128) {
  auto enc = get_encode2();
  if (!enc) return HB_ERR_CUDA;
  cuuint64_t gdim[2], gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2], estr[2] = {1, 1};
  if (kmajor) { gdim[0] = (cuuint64_t)K; gdim[1] = (cuuint64_t)rows; box[0] = H_BK; box[1] = (cuuint32_t)box_rows; }
  else { gdim[0] = (cuuint64_t)rows; gdim[1] = (cuuint64_t)K; box[0] = 64; box[1] = H_BK; }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HB_OK : HB_ERR_CUDA;
}

template <int BN, bool AKM, bool BKM>
int launch_h2(const CUtensorMap* m, H2Params hp, cudaStream_t st, int nsplit = 1) {
  constexpr int SMEM = H2Geom<BN>::NSTAGE * H2Geom<BN>::STAGE + 1024 + 256;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(gemm_h2_pair_kernel<BN, AKM, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess)
      return HB_ERR_CUDA;
    attr_done = true;
  }
  hp.t.tiles_m = cdiv(hp.t.M, 256);
  hp.t.tiles_n = cdiv(hp.t.N, BN);
  gemm_h2_pair_kernel<BN, AKM, BKM><<<dim3(2 * hp.t.tiles_m * hp.t.tiles_n, nsplit), NTHREADS, SMEM, st>>>(m[0], m[1], m[2], m[3], hp);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// operand preparation: absolute maximum, power-of-two scale, fp16 hi/lo split
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned short f2h_sat(float x) {
  unsigned short h;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
  return h;
}
__device__ __forceinline__ float h2f(unsigned short h) {
  float f;
  asm("cvt.f32.f16 %0, %1;" : "=f"(f) : "h"(h));
  return f;
}

// max |A[i][j]| over the block (lower_only: j <= i + diag_off) -> atomicMax on the float's bit pattern (values >= 0)
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ A, long long ld, long long rows, int cols,
                                                     int lower_only, long long diag_off, unsigned* out_bits) {
  const int c4 = cols >> 2;                       // cols % 4 == 0
  const long long total = rows * c4;
  float m = 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / c4; const int j = (int)(e % c4) * 4;
    if (lower_only && j > i + diag_off) continue;
    const float4 v = __ldg(reinterpret_cast<const float4*>(A + i * ld + j));
    float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) if (!lower_only || j + q <= i + diag_off) m = fmaxf(m, fabsf(t[q]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}

__global__ void diag_absmax_kernel(const float* __restrict__ A, long long ld, int n, unsigned* out_bits) {
  float m = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = fmaxf(m, fabsf(A[(long long)i * ld + i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out_bits, __float_as_uint(m));
}

// power-of-two scale that brings `maxv` to [2^(H2_TARGET_EXP-1), 2^H2_TARGET_EXP]
__device__ __forceinline__ float scale_for(float maxv) {
  if (!(maxv > 0.f) || !isfinite(maxv)) return 1.f;
  int e;
  frexpf(maxv, &e);                               // maxv = f * 2^e, f in [0.5, 1)
  int se = H2_TARGET_EXP - e;
  se = max(-100, min(100, se));
  return ldexpf(1.f, se);
}

// scale_io[0] = s, scale_io[1] = 1/s derived from *max_bits (if sqrt_of_max: from sqrt(max): the bound of a Cholesky factor)
__global__ void scale_from_max_kernel(const unsigned* max_bits, int sqrt_of_max, float* scale2) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float m = __uint_as_float(*max_bits);
    if (sqrt_of_max) m = sqrtf(m) * 1.0001f;
    const float s = scale_for(m);
    scale2[0] = s; scale2[1] = 1.f / s;
  }
}

// hi/lo split of a [rows x cols] fp32 block (cols % 8 == 0, 16-byte aligned rows) into the shadows at the same (i, j).
// The scale is either read from scale2[0] (fixed) or derived from *max_bits, in which case thread 0 also publishes
// inv_out[0] = 1/s.  lower_only skips j > i + diag_off (never read by any product).
__global__ void __launch_bounds__(256) split_f16_kernel(const float* __restrict__ A, long long ld, long long rows, int cols,
                                                        const float* scale2, const unsigned* max_bits, float* inv_out,
                                                        int lower_only, long long diag_off, __half* hi, __half* lo, long long ldh) {
  float s;
  if (max_bits) {
    s = scale_for(__uint_as_float(*max_bits));
    if (inv_out && blockIdx.x == 0 && threadIdx.x == 0) *inv_out = 1.f / s;
  } else {
    s = scale2[0];
  }
  const int c8 = cols >> 3;
  const long long total = rows * c8;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / c8; const int j = (int)(e % c8) * 8;
    if (lower_only && j > i + diag_off) continue;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(A + i * ld + j));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(A + i * ld + j + 4));
    const float t[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    unsigned short h[8], l[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float x = t[q] * s;
      h[q] = f2h_sat(x);
      l[q] = f2h_sat(x - h2f(h[q]));
    }
    uint4 ph, pl;
    ph.x = h[0] | ((unsigned)h[1] << 16); ph.y = h[2] | ((unsigned)h[3] << 16);
    ph.z = h[4] | ((unsigned)h[5] << 16); ph.w = h[6] | ((unsigned)h[7] << 16);
    pl.x = l[0] | ((unsigned)l[1] << 16); pl.y = l[2] | ((unsigned)l[3] << 16);
    pl.z = l[4] | ((unsigned)l[5] << 16); pl.w = l[6] | ((unsigned)l[7] << 16);
    *reinterpret_cast<uint4*>(hi + i * ldh + j) = ph;
    *reinterpret_cast<uint4*>(lo + i * ldh + j) = pl;
  }
}

// hi/lo split of src [rows x cols] into TRANSPOSED shadows [cols x rows] (small operands only: the sample-minor
// residual R^T [M, S] of the linear-operator step has to become the K-major operand R [S, M]).  32 x 32 tiles via smem.
__global__ void __launch_bounds__(256) split_f16_transpose_kernel(const float* __restrict__ A, long long ld, int rows, int cols,
                                                                  const unsigned* max_bits, float* inv_out, __half* hi, __half* lo,
                                                                  long long ldh) {
  __shared__ float tile[32][33];
  const float s = scale_for(__uint_as_float(*max_bits));
  if (inv_out && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *inv_out = 1.f / s;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? A[(long long)r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;                               // output row = source column
    if (c < cols && r < rows) {
      const float x = tile[tx][i] * s;
      const unsigned short h = f2h_sat(x);
      const unsigned short l = f2h_sat(x - h2f(h));
      reinterpret_cast<unsigned short*>(hi)[(long long)c * ldh + r] = h;
      reinterpret_cast<unsigned short*>(lo)[(long long)c * ldh + r] = l;
    }
  }
}

inline int grid_rows(long long work, int threads) {
  long long b = (work + threads - 1) / threads;
  if (b > 148 * 16) b = 148 * 16;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

// Split count for a product with `tiles` pair tiles and `kblocks` 64-wide k-blocks: the one with the best wave efficiency
// (74 clusters per wave) among those that leave every split at least K = 1024.  0 = do not split.
static int h2_split_count(long long tiles, int kblocks, size_t part_bytes, size_t ws_bytes) {
  int best_ns = 0; double best = (double)tiles / (74.0 * (double)((tiles + 73) / 74));
  for (int ns = 2; ns <= 32 && kblocks / ns >= 16; ++ns) {
    if ((size_t)ns * part_bytes + 256 > ws_bytes) break;
    const long long cl = tiles * ns;
    const double eff = (double)cl / (74.0 * (double)((cl + 73) / 74));
    if (eff > best + 0.05) { best = eff; best_ns = ns; }
  }
  return best_ns;
}
bool gemm_h2_splitk_eligible(int M, int N, int K, size_t ws_bytes) {
  if (!(M > 128 && N > 128 && K >= 4096)) return false;
  const long long tiles = (long long)cdiv(M, 256) * cdiv(N, 256);
  if (tiles >= 32) return false;                      // the plain launch fills the GPU
  return h2_split_count(tiles, cdiv(K, H_BK), (size_t)M * N * sizeof(float), ws_bytes) >= 2;
}

bool gemm_h2_eligible(int M, int N, int K) {
  // CTA-pair tiles of 256 x 256; worth it from 32 pair tiles on (fewer leave most of the 74 cluster slots empty) and
  // K >= 256 (prologue + epilogue of a tile cost about one K = 256 main loop)
  return M > 128 && N > 128 && K >= 256 && (long long)cdiv(M, 256) * cdiv(N, 256) >= 32;
}

int gemm_h2(const H2Gemm& g, cudaStream_t st) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return HB_OK;
  if (!g.a_hi || !g.a_lo || !g.b_hi || !g.b_lo || !g.C) return HB_ERR_ARG;
  if ((g.lda & 7) || (g.ldb & 7) || !aligned16(g.a_hi) || !aligned16(g.a_lo) || !aligned16(g.b_hi) || !aligned16(g.b_lo))
    return HB_ERR_ARG;
  if (g.a_bmode && (g.M != g.K)) return HB_ERR_ARG;             // block masks are defined on a square op(A)
  const bool narrow = g.N <= 64;                 // 256 x 64 pair tiles: few samples against a big operator
  if (narrow && !g.b_kmajor) return HB_ERR_ARG;
  CUtensorMap m[4];
  HB_TRY(make_map_h(&m[0], g.a_hi, g.M, g.K, g.lda, g.a_kmajor != 0));
  HB_TRY(make_map_h(&m[1], g.a_lo, g.M, g.K, g.lda, g.a_kmajor != 0));
  HB_TRY(make_map_h(&m[2], g.b_hi, g.N, g.K, g.ldb, g.b_kmajor != 0, narrow ? 32 : 128));
  HB_TRY(make_map_h(&m[3], g.b_lo, g.N, g.K, g.ldb, g.b_kmajor != 0, narrow ? 32 : 128));
  H2Params hp;
  Tc2Params& tp = hp.t;
  tp.C = g.C; tp.ldc = g.ldc; tp.M = g.M; tp.N = g.N; tp.K = g.K; tp.alpha = g.alpha; tp.beta = g.beta;
  tp.c_tri = g.c_tri; tp.a_mode = 0; tp.b_mode = 0; tp.tiles_m = 0; tp.tiles_n = 0;
  tp.vecC = aligned16(g.C) && (g.ldc % 4 == 0);
  tp.ksplit = 0; tp.csplit = 0; tp.bias = nullptr; tp.act = ACT_NONE; tp.clip = 0; tp.clip_lo = 0.f; tp.clip_hi = 0.f;
  hp.group = (get_tc_option() >> 8) & 0xff;      // A/B experiments: hb_options.tc_option bits 8..15 override the raster group
  if (hp.group <= 0) hp.group = 4;                  // measured (tools/h2_group_probe.py): 4 beats 8 by 5-8 % at the 32768-wide top-level products, equal at 16384
  hp.a_bmode = g.a_bmode; hp.a_inv = g.a_inv; hp.a_kinv = g.a_kinv; hp.a_dinv = g.a_dinv ? g.a_dinv : g.a_kinv;
  hp.a_minv = g.a_minv; hp.b_inv = g.b_inv;
  if (narrow) return g.a_kmajor ? launch_h2<64, true, true>(m, hp, st) : launch_h2<64, false, true>(m, hp, st);
  // long K, few output tiles (the tall reductions of the reverse mode inside a block): split along K into the caller's
  // scratch, one deterministic reduction pass
  int nsplit = 1;
  float* part = nullptr;
  const long long tiles = (long long)cdiv(g.M, 256) * cdiv(g.N, 256);
  if (g.ws && !g.a_bmode && tiles < 32 && g.K >= 4096) {
    const int kblocks = cdiv(g.K, H_BK);
    const int ns = h2_split_count(tiles, kblocks, (size_t)g.M * g.N * sizeof(float), g.ws_bytes);
    if (ns >= 2) {
      int per = cdiv(kblocks, ns);
      per = (per + H_CHS - 1) / H_CHS * H_CHS;          // whole K = 128 accumulation chunks (the per-chunk scales)
      nsplit = cdiv(kblocks, per);
      if (nsplit >= 2) {
        part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(g.ws) + 255) & ~uintptr_t(255));
        tp.C = part; tp.ldc = g.N; tp.beta = 0.f; tp.ksplit = per; tp.csplit = (long long)g.M * g.N;
        tp.vecC = (g.N % 4 == 0);
      } else nsplit = 1;
    }
  }
  int rc;
  if (g.a_kmajor) rc = g.b_kmajor ? launch_h2<256, true, true>(m, hp, st, nsplit) : launch_h2<256, true, false>(m, hp, st, nsplit);
  else rc = g.b_kmajor ? launch_h2<256, false, true>(m, hp, st, nsplit) : launch_h2<256, false, false>(m, hp, st, nsplit);
  if (rc != HB_OK || nsplit == 1) return rc;
  return splitk_reduce(g.C, g.ldc, part, (long long)g.M * g.N, g.M, g.N, nsplit, g.beta, g.c_tri, st);
}

int h2_absmax(const float* A, long long ld, long long rows, int cols, int lower_only, long long diag_off, unsigned* out_bits,
              cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return HB_OK;
  if ((cols & 3) || (ld & 3) || !aligned16(A)) return HB_ERR_ARG;
  absmax_kernel<<<grid_rows(rows * (cols >> 2), 256), 256, 0, st>>>(A, ld, rows, cols, lower_only, diag_off, out_bits);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int h2_diag_absmax(const float* A, long long ld, int n, unsigned* out_bits, cudaStream_t st) {
  if (n <= 0) return HB_OK;
  diag_absmax_kernel<<<cdiv(n, 256) > 148 ? 148 : cdiv(n, 256), 256, 0, st>>>(A, ld, n, out_bits);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int h2_scale_from_max(const unsigned* max_bits, int sqrt_of_max, float* scale2, cudaStream_t st) {
  scale_from_max_kernel<<<1, 32, 0, st>>>(max_bits, sqrt_of_max, scale2);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int h2_split(const float* A, long long ld, long long rows, int cols, const float* scale2, const unsigned* max_bits,
             float* inv_out, int lower_only, long long diag_off, __half* hi, __half* lo, long long ldh, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return HB_OK;
  if ((cols & 7) || (ld & 3) || (ldh & 7) || !aligned16(A) || !aligned16(hi) || !aligned16(lo)) return HB_ERR_ARG;
  if (!scale2 && !max_bits) return HB_ERR_ARG;
  split_f16_kernel<<<grid_rows(rows * (cols >> 3), 256), 256, 0, st>>>(A, ld, rows, cols, scale2, max_bits, inv_out, lower_only,
                                                                     diag_off, hi, lo, ldh);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int h2_split_transpose(const float* A, long long ld, int rows, int cols, const unsigned* max_bits, float* inv_out, __half* hi,
                       __half* lo, long long ldh, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return HB_OK;
  if (!max_bits || (ldh & 7) || !aligned16(hi) || !aligned16(lo)) return HB_ERR_ARG;
  dim3 grid(cdiv(cols, 32), cdiv(rows, 32));
  if (grid.y > 65535) return HB_ERR_ARG;
  split_f16_transpose_kernel<<<grid, 256, 0, st>>>(A, ld, rows, cols, max_bits, inv_out, hi, lo, ldh);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

}  // namespace hb
