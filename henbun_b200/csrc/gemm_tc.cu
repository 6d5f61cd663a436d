// tcgen05 3xTF32 GEMM engine for sm_100a:  C[MxN] = alpha * A[MxK] * B[NxK]^T + beta * C   (fp32 in/out)
//
// fp32-grade accuracy on the 5th-gen tensor cores: every operand is split a = hi + lo with
// hi = a & 0xFFFFE000 (exact in TF32) and lo = a - hi, and three MMAs hi*hi + hi*lo + lo*hi accumulate
// into one fp32 TMEM accumulator (the dropped lo*lo term and the truncation of lo are ~2^-22 relative).
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0      : TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B, 4 operand tiles per stage, mbarrier tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer (kind::tf32, M=128, N=BN, K=8)
//   warps 2..5  : epilogue: tcgen05.ld 32x32b.x32 -> alpha*acc + beta*C -> global (masked tails / lower-tri)
// Operands are K-major ("TN"): rows of A and B are contiguous in k; BK = 32 floats = one 128-byte swizzle row.
#include "gemm.cuh"
#include "kernels.cuh"
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cstdlib>

namespace hb {

namespace {

constexpr int BM = 128;
constexpr int BK = 32;                 // floats per k-block: 128 bytes = SW128 atom width
constexpr int A_TILE_BYTES = BM * BK * 4;
constexpr int TC_THREADS = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (8 rows * 128 B = 1024)
//   [46,48) version=1 | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct TcParams {
  float* C;
  long long ldc;
  int M, N, K;
  float alpha, beta;
  int c_tri;
  int a_mode, b_mode;   // triangular structure in (row, k) space: 1 k<=r, 2 k>=r, 3 k>r, 4 k<r (elements are zeroed
                        // by the split pass; here the modes only trim the k-range of each tile)
  int tiles_m, tiles_n;
  int vecC;
};

// k-block range [lo, hi) that can hold non-zeros for a tile whose rows start at r0 (extent ext)
__device__ __forceinline__ void trim_range(int mode, int r0, int ext, int& lo, int& hi) {
  if (mode == 1) hi = min(hi, (r0 + ext + BK - 1) / BK);            // k <= r  -> k < r0+ext
  else if (mode == 2) lo = max(lo, r0 / BK);                          // k >= r
  else if (mode == 3) lo = max(lo, (r0 + 1) / BK);                    // k >  r
  else if (mode == 4) hi = min(hi, (r0 + ext - 1 + BK - 1) / BK);   // k <  r  -> k < r0+ext-1
}

template <int BN, int NSTAGE>
struct Smem {
  static constexpr int B_TILE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int TOTAL = NSTAGE * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Two-level accumulation: the tensor core accumulates CH k-blocks (K = 32*CH) into one of two TMEM
// buffers, the epilogue warps then add that partial tile into fp32 registers with round-to-nearest
// while the MMAs of the next chunk run into the other buffer.  The tensor-core accumulator truncates
// (measured: relative error grows linearly in K, 1.2e-4 at K=8192, without this promotion).
constexpr int CH = 4;
constexpr int EPI_WARPS = 8;
constexpr int TC_THREADS2 = 64 + 32 * EPI_WARPS;

template <int NSTAGE>
__global__ void __launch_bounds__(TC_THREADS2, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmAhi, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmBhi, const __grid_constant__ CUtensorMap tmBlo, const TcParams p) {
  constexpr int BN = 256;
  using S = Smem<BN, NSTAGE>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;           // SW128 needs 1024-byte alignment
  const uint32_t bars = base + NSTAGE * S::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * NSTAGE + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * NSTAGE + 2 + b); };
  const uint32_t tmem_ptr_addr = bars + 8u * (2 * NSTAGE + 4);

  // tile coordinates (grouped rasterisation for L2 reuse of the B panel)
  constexpr int GROUP = 8;
  const int bid = blockIdx.x;
  const int per_group = GROUP * p.tiles_n;
  const int first_m = (bid / per_group) * GROUP;
  const int gsz = min(p.tiles_m - first_m, GROUP);
  const int tm = first_m + (bid % per_group) % gsz;
  const int tn = (bid % per_group) / gsz;
  const int m0 = tm * BM, n0 = tn * BN;
  if (p.c_tri == 1 && n0 > m0 + BM - 1) return;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int kb_lo = 0, kb_hi = (p.K + BK - 1) / BK;
  trim_range(p.a_mode, m0, BM, kb_lo, kb_hi);
  trim_range(p.b_mode, n0, BN, kb_lo, kb_hi);
  const int num_k = max(kb_hi - kb_lo, 0);
  const int num_c = (num_k + CH - 1) / CH;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: two 256-column fp32 accumulator buffers = all 512 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_ptr_addr), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    if (lane == 0) {   // ---------------- TMA producer ----------------
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < num_k; ++kb) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = base + s * S::STAGE_BYTES;
        mbar_expect_tx(full_bar(s), S::STAGE_BYTES);
        const int k0 = (kb_lo + kb) * BK;
        tma_load_2d(st, &tmAhi, full_bar(s), k0, m0);
        tma_load_2d(st + A_TILE_BYTES, &tmAlo, full_bar(s), k0, m0);
        tma_load_2d(st + 2 * A_TILE_BYTES, &tmBhi, full_bar(s), k0, n0);
        tma_load_2d(st + 2 * A_TILE_BYTES + S::B_TILE_BYTES, &tmBlo, full_bar(s), k0, n0);
        if (++s == NSTAGE) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---------------- MMA issuer ----------------
      // instruction descriptor: D=F32 (1<<4), A=B=TF32 (2<<7, 2<<10), K-major both, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0; uint32_t ph = 0;
      for (int c = 0; c < num_c; ++c) {
        const int buf = c & 1;
        mbar_wait(tempty_bar(buf), (uint32_t)(((c >> 1) & 1) ^ 1));     // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        const int kb_end = min(num_k, (c + 1) * CH);
        for (int kb = c * CH; kb < kb_end; ++kb) {
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t st = base + s * S::STAGE_BYTES;
          const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + A_TILE_BYTES);
          const uint64_t b_hi = make_desc(st + 2 * A_TILE_BYTES), b_lo = make_desc(st + 2 * A_TILE_BYTES + S::B_TILE_BYTES);
          const bool first_kb = (kb == c * CH);
#pragma unroll
          for (int k4 = 0; k4 < BK / 8; ++k4) {       // K = 8 per tf32 MMA = 32 bytes along the swizzled row
            const uint64_t adv = (uint64_t)((k4 * 32) >> 4);
            tc_mma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (first_kb && k4 == 0) ? 0u : 1u);   // small terms first
            tc_mma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
            tc_mma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
          }
          tc_commit(empty_bar(s));                     // frees the smem stage once these MMAs have read it
          if (++s == NSTAGE) { s = 0; ph ^= 1u; }
        }
        tc_commit(tfull_bar(buf));                     // this chunk's partial tile is complete
      }
    }
  } else {
    // ---------------- epilogue: 8 warps; warp (w%4) owns TMEM lanes [32(w%4), +32), column half (w-2)/4 ----------------
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0.f;
    for (int c = 0; c < num_c; ++c) {
      const int buf = c & 1;
      mbar_wait(tfull_bar(buf), (uint32_t)((c >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * 128 + i * 32);
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
        for (int j = 0; j < 32; ++j) acc[i * 32 + j] += __uint_as_float(r[j]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
    }
    const int gi = m0 + q * 32 + lane;
    if (gi < p.M) {
      float* crow = p.C + (long long)gi * p.ldc;
      const int gj0 = n0 + half * 128;
#pragma unroll
      for (int v = 0; v < 32; ++v) {
        const int gj = gj0 + v * 4;
        if (gj < p.N) {
          float o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] = p.alpha * acc[v * 4 + e];
          const bool full = (gj + 3 < p.N) && !(p.c_tri == 1 && gj + 3 > gi);
          if (full && p.vecC) {
            if (p.beta != 0.f) {
              const float4 old = *reinterpret_cast<const float4*>(crow + gj);
              o[0] = fmaf(p.beta, old.x, o[0]); o[1] = fmaf(p.beta, old.y, o[1]);
              o[2] = fmaf(p.beta, old.z, o[2]); o[3] = fmaf(p.beta, old.w, o[3]);
            }
            *reinterpret_cast<float4*>(crow + gj) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              if (gj + e < p.N && !(p.c_tri == 1 && gj + e > gi)) {
                float x = o[e];
                if (p.beta != 0.f) x = fmaf(p.beta, crow[gj + e], x);
                crow[gj + e] = x;
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// ---- operand preparation: hi = a & 0xFFFFE000 (what the tf32 MMA reads from a raw fp32 word), lo = rna_tf32(a - hi) ----
__device__ __forceinline__ bool tri_keep(int mode, long long r, long long k) {
  return mode == 0 || (mode == 1 && k <= r) || (mode == 2 && k >= r) || (mode == 3 && k > r) || (mode == 4 && k < r);
}
__device__ __forceinline__ void split1(float v, float& h, float& l) {
  h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
  float d = v - h;                                   // exact
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(d));
  l = __uint_as_float(t);
}

// source stored [rows x K] (K-major): element (r,k) = src[r*lds + k]
__global__ void split_kmajor_kernel(const float* __restrict__ src, long long lds, int rows, int K, float* hi, float* lo,
                                    long long ldw, int write_hi, int mode) {
  const int c4 = (K + 3) / 4;
  const long long total = (long long)rows * c4;
  const bool vec = ((lds & 3) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / c4; const int c = (int)(e % c4) * 4;
    float v[4], h[4], l[4];
    if (vec && c + 3 < K) {
      const float4 t = *reinterpret_cast<const float4*>(src + r * lds + c);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = (c + k < K) ? src[r * lds + c + k] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (!tri_keep(mode, r, c + k)) v[k] = 0.f;
      split1(v[k], h[k], l[k]);
    }
    if (write_hi) *reinterpret_cast<float4*>(hi + r * ldw + c) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo + r * ldw + c) = make_float4(l[0], l[1], l[2], l[3]);
  }
}

// source stored [K x rows] (row-index-major): element (r,k) = src[k*lds + r]; outputs K-major.  32x32 smem transpose.
__global__ void split_transpose_kernel(const float* __restrict__ src, long long lds, int rows, int K, float* hi, float* lo,
                                       long long ldw, int mode) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int k = k0 + j, r = r0 + threadIdx.x;
    t[j][threadIdx.x] = (k < K && r < rows) ? src[(long long)k * lds + r] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, k = k0 + threadIdx.x;
    if (r < rows && k < (int)ldw) {
      float v = (k < K && tri_keep(mode, r, k)) ? t[threadIdx.x][j] : 0.f;
      float h, l;
      split1(v, h, l);
      hi[(long long)r * ldw + k] = h;
      lo[(long long)r * ldw + k] = l;
    }
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
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

// 2-D fp32 row-major [rows x cols] with leading dimension ld (elements); box = [BK x box_rows], SW128
int make_map(CUtensorMap* map, const float* ptr, long long rows, long long cols, long long ld, int box_rows) {
  auto enc = get_encode();
  if (!enc) return HB_ERR_CUDA;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HB_OK : HB_ERR_CUDA;
}

int g_tc_option = 0;   // bit0: always write an explicit "hi" copy instead of feeding the raw fp32 operand (measured: the
                       // tf32 MMA ignores the 13 low mantissa bits, results are bit-identical)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int NSTAGE>
int launch_tc(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl, TcParams tp,
              cudaStream_t st) {
  using S = Smem<256, NSTAGE>;
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(gemm_tc_kernel<NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL) != cudaSuccess)
      return HB_ERR_CUDA;
    attr_done = true;
  }
  tp.tiles_m = cdiv(tp.M, BM);
  tp.tiles_n = cdiv(tp.N, 256);
  gemm_tc_kernel<NSTAGE><<<tp.tiles_m * tp.tiles_n, TC_THREADS2, S::TOTAL, st>>>(ah, al, bh, bl, tp);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

// One operand -> (hi, lo) K-major views.  `kmajor`: stored [rows x K]; otherwise stored [K x rows].
struct Prepared { const float* hi; long long ld_hi; const float* lo; long long ld_lo; };

int prepare_operand(const float* src, long long ld, int rows, int K, bool kmajor, int mode, float*& w, Prepared& out,
                    cudaStream_t st) {
  const long long kp = ((long long)K + 3) / 4 * 4;
  const bool raw_hi = kmajor && mode == 0 && !(g_tc_option & 1) && aligned16(src) && (ld % 4 == 0);
  float* hi = nullptr;
  if (!raw_hi) { hi = w; w += (long long)rows * kp; }
  float* lo = w; w += (long long)rows * kp;
  if (kmajor) {
    const long long tot = (long long)rows * (kp / 4);
    int nb = (int)((tot + 255) / 256); if (nb > 148 * 8) nb = 148 * 8; if (nb < 1) nb = 1;
    split_kmajor_kernel<<<nb, 256, 0, st>>>(src, ld, rows, K, hi, lo, kp, !raw_hi, mode);
  } else {
    dim3 grid((unsigned)cdiv(rows, 32), (unsigned)cdiv(kp, 32));
    split_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(src, ld, rows, K, hi, lo, kp, mode);
  }
  HB_CHECK_LAUNCH();
  out.hi = raw_hi ? src : hi; out.ld_hi = raw_hi ? ld : kp; out.lo = lo; out.ld_lo = kp;
  return HB_OK;
}

}  // namespace

void set_tc_option(int v) { g_tc_option = v; }
int get_tc_option() {
  static bool env_read = false;
  if (!env_read) {                      // HB_TC_OPTION=<bits> overrides the default once, at first use (A/B runs)
    env_read = true;
    if (const char* e = getenv("HB_TC_OPTION")) g_tc_option = atoi(e);
  }
  return g_tc_option;
}

size_t gemm_tc_workspace_bytes(int M, int N, int K) {
  const long long kp = ((long long)K + 3) / 4 * 4;
  return (size_t)(2 * ((long long)M + N) * kp) * sizeof(float) + 512;
}

// Shapes worth a tensor-core launch (plus its operand-preparation passes); everything else stays on the SIMT engine.
bool gemm_tc_eligible(const GemmParams& p) {
  if (get_gemm_engine() == 1) return false;
  if (p.batch != 1 || p.bias || p.act != ACT_NONE || p.clip) return false;
  if (p.M < 128 || p.N < 128 || p.K < 32) return false;
  if (get_gemm_engine() != 2 && (double)p.M * p.N * p.K < 256.0 * 256.0 * 256.0) return false;
  if (p.C == p.A || p.C == p.B) return false;
  if (!p.ws || p.ws_bytes < gemm_tc_workspace_bytes(p.M, p.N, p.K)) return false;
  return true;
}

int gemm_tc(const GemmParams& p, cudaStream_t st) {
  float* w = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p.ws) + 255) & ~uintptr_t(255));
  static const int b2rk[5] = {0, 2, 1, 4, 3};      // mask of op(B)[k][n] expressed in (n, k) space
  const int a_mode = p.a_tri, b_mode = b2rk[p.b_tri];
  const bool a_kmajor = (p.transA == 0), b_kmajor = (p.transB == 1);
  const bool same = (p.A == p.B && p.lda == p.ldb && p.M == p.N && a_kmajor == b_kmajor && a_mode == b_mode);
  Prepared a, b;
  HB_TRY(prepare_operand(p.A, p.lda, p.M, p.K, a_kmajor, a_mode, w, a, st));
  if (same) b = a;
  else HB_TRY(prepare_operand(p.B, p.ldb, p.N, p.K, b_kmajor, b_mode, w, b, st));
  const int BN = 256;
  CUtensorMap ah, al, bh, bl;
  HB_TRY(make_map(&ah, a.hi, p.M, p.K, a.ld_hi, BM));
  HB_TRY(make_map(&al, a.lo, p.M, p.K, a.ld_lo, BM));
  HB_TRY(make_map(&bh, b.hi, p.N, p.K, b.ld_hi, BN));
  HB_TRY(make_map(&bl, b.lo, p.N, p.K, b.ld_lo, BN));
  TcParams tp;
  tp.C = p.C; tp.ldc = p.ldc; tp.M = p.M; tp.N = p.N; tp.K = p.K; tp.alpha = p.alpha; tp.beta = p.beta;
  tp.c_tri = p.c_tri; tp.a_mode = a_mode; tp.b_mode = b_mode; tp.tiles_m = 0; tp.tiles_n = 0;
  tp.vecC = aligned16(p.C) && (p.ldc % 4 == 0);
  return launch_tc<2>(ah, al, bh, bl, tp, st);
}

}  // namespace hb
