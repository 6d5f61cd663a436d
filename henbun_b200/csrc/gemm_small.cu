// Latency-oriented fp32 GEMM for the short-K products of the blocked Cholesky / reverse mode (K <= 256):
// the whole K extent of a (16*TM) x 128 output tile is staged into shared memory with ONE round trip to global
// memory (all loads of a thread are in flight together), then the tile is computed out of shared memory.
// The k-looped engine (gemm_simt.cu) pays a global-memory latency per 8-wide k step, which dominates when
// K is 128 and the grid is a handful of CTAs (measured 36 us for 128^3; this kernel: a few us).
//
// One CTA owns ALL columns of its rows when N <= 128, so C may alias A (in-place triangular-solve leaves,
// X <- X * Dinv^T): every element of A the CTA needs is in shared memory before its first store.
#include "gemm.cuh"

namespace hb {

namespace {

constexpr int SNT = 256;
constexpr int SBN = 128;

__device__ __forceinline__ bool keep_rk(int mode, int r, int k) {
  return mode == 0 || (mode == 1 && k <= r) || (mode == 2 && k >= r) || (mode == 3 && k > r) || (mode == 4 && k < r);
}

struct SParams {
  const float* A; const float* B; float* C;
  long long lda, ldb, ldc, sA, sB, sC;
  int M, N, K, Kp;          // Kp = K rounded up to 4
  float alpha, beta;
  int a_mode, b_mode, c_tri;
  int a_kmajor, b_kmajor;   // element (r,k) at P[r*ld+k] (kmajor) or P[k*ld+r]
  int vecA, vecB, vecC;
};

// Stage op(X)[r0 .. r0+ROWS) x [0,K) into S[k][r] (row stride ROWS+4), zero-filled outside, masked.
// Loads are issued in batches of UB float4 per thread BEFORE any shared-memory store, so a tile costs one or two
// global-memory round trips instead of one per loop iteration.  The code is kept deliberately compact (one path,
// modest unrolling): these kernels run once per launch and a 100 KB instruction stream was fetch-bound
// (ncu: no_instruction the top stall at IPC 0.8).  Requires 16-byte aligned rows (checked by the dispatcher).
template <int ROWS, bool KMAJOR>
__device__ __forceinline__ void stage(float* S, const float* __restrict__ P, long long ld, int r0, int R, int K,
                                      int mode) {
  constexpr int LD = ROWS + 4;
  constexpr int UB = 4;
  const int tid = threadIdx.x;
  const int kq = K / 4;                            // K % 4 == 0
  // KMAJOR: a warp reads 8 rows x 4 float4 (64 contiguous bytes per row); stores have <= 2-way bank conflicts
  // else  : a warp reads 32 consecutive float4 of one k row and stores them as float4
  const int total = KMAJOR ? (ROWS / 8) * ((kq + 3) / 4) * 32 : K * (ROWS / 4);
  for (int g0 = tid; g0 < total; g0 += SNT * UB) {
    float4 v[UB];
    int rk[UB];                                    // r | k << 16, or -1
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      const int g = g0 + u * SNT;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      rk[u] = -1;
      int r, k;
      if (KMAJOR) {
        const int un = g >> 5, l = g & 31;
        r = (un % (ROWS / 8)) * 8 + (l & 7);
        k = ((un / (ROWS / 8)) * 4 + (l >> 3)) * 4;
      } else {
        k = g / (ROWS / 4); r = (g % (ROWS / 4)) * 4;
      }
      if (g >= total || k >= K) continue;
      rk[u] = r | (k << 16);
      const int gr = r0 + r;
      if (KMAJOR) {
        if (gr < R) v[u] = __ldg(reinterpret_cast<const float4*>(P + (long long)gr * ld + k));
      } else {
        const float* src = P + (long long)k * ld + gr;
        if (gr + 3 < R) v[u] = __ldg(reinterpret_cast<const float4*>(src));
        else if (gr < R) {                         // ragged right edge of a row
          float t[4] = {0.f, 0.f, 0.f, 0.f};
          for (int c = 0; c < R - gr; ++c) t[c] = __ldg(src + c);
          v[u] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UB; ++u) {
      if (rk[u] < 0) continue;
      const int r = rk[u] & 0xffff, k = rk[u] >> 16, gr = r0 + r;
      float t[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
      if (mode) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (!(KMAJOR ? keep_rk(mode, gr, k + c) : keep_rk(mode, gr + c, k))) t[c] = 0.f;
      }
      if (KMAJOR) {
#pragma unroll
        for (int c = 0; c < 4; ++c) S[(k + c) * LD + r] = t[c];
      } else {
        *reinterpret_cast<float4*>(&S[k * LD + r]) = make_float4(t[0], t[1], t[2], t[3]);
      }
    }
  }
}

// BM = 16*TM rows x 128 columns per CTA; thread (ty = tid/16, tx = tid%16) owns rows ty*TM.. and columns
// tx*4..+3 and 64+tx*4..+3.
template <int TM, bool AKM, bool BKM>
__global__ void __launch_bounds__(SNT) gemm_small_kernel(const SParams p) {
  constexpr int BM = 16 * TM;
  constexpr int LDA_S = BM + 4, LDB_S = SBN + 4;
  extern __shared__ __align__(16) float sm[];
  float* As = sm;
  float* Bs = sm + (size_t)p.Kp * LDA_S;
  const int tiles_m = (p.M + BM - 1) / BM;
  const int tm = blockIdx.x % tiles_m, tn = blockIdx.x / tiles_m;
  const int m0 = tm * BM, n0 = tn * SBN;
  if (p.c_tri == 1 && n0 > m0 + BM - 1) return;
  const int bz = blockIdx.y;
  const float* A = p.A + (long long)bz * p.sA;
  const float* B = p.B + (long long)bz * p.sB;
  float* C = p.C + (long long)bz * p.sC;

  stage<BM, AKM>(As, A, p.lda, m0, p.M, p.K, p.a_mode);
  stage<SBN, BKM>(Bs, B, p.ldb, n0, p.N, p.K, p.b_mode);
  __syncthreads();

  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  float acc[TM][8];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // k-range of this tile under the triangular masks
  int k_lo = 0, k_hi = p.Kp;
  if (p.a_mode == 1) k_hi = min(k_hi, m0 + BM); else if (p.a_mode == 2) k_lo = max(k_lo, m0);
  else if (p.a_mode == 3) k_lo = max(k_lo, m0 + 1); else if (p.a_mode == 4) k_hi = min(k_hi, m0 + BM - 1);
  if (p.b_mode == 1) k_hi = min(k_hi, n0 + SBN); else if (p.b_mode == 2) k_lo = max(k_lo, n0);
  else if (p.b_mode == 3) k_lo = max(k_lo, n0 + 1); else if (p.b_mode == 4) k_hi = min(k_hi, n0 + SBN - 1);

#pragma unroll 2
  for (int k = k_lo; k < k_hi; ++k) {
    float a[TM], b[8];
    if (TM == 4) {
      const float4 v = *reinterpret_cast<const float4*>(&As[k * LDA_S + ty * 4]);
      a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[k * LDA_S + ty * TM + i];
    }
    const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k * LDB_S + tx * 4]);
    const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k * LDB_S + 64 + tx * 4]);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gi = m0 + ty * TM + i;
    if (gi >= p.M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int gj = n0 + h * 64 + tx * 4;
      if (gj >= p.N) continue;
      float* cp = C + (long long)gi * p.ldc + gj;
      float o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) o[c] = p.alpha * acc[i][h * 4 + c];
      const bool full = (gj + 3 < p.N) && !(p.c_tri == 1 && gj + 3 > gi);
      if (full && p.vecC) {
        if (p.beta != 0.f) {
          const float4 old = *reinterpret_cast<const float4*>(cp);
          o[0] = fmaf(p.beta, old.x, o[0]); o[1] = fmaf(p.beta, old.y, o[1]);
          o[2] = fmaf(p.beta, old.z, o[2]); o[3] = fmaf(p.beta, old.w, o[3]);
        }
        *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (gj + c >= p.N || (p.c_tri == 1 && gj + c > gi)) continue;
          float x = o[c];
          if (p.beta != 0.f) x = fmaf(p.beta, cp[c], x);
          cp[c] = x;
        }
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int TM, bool AKM, bool BKM>
int launch_small2(const SParams& sp, int batch, cudaStream_t st) {
  constexpr int BM = 16 * TM;
  const size_t smem = (size_t)sp.Kp * ((BM + 4) + (SBN + 4)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(gemm_small_kernel<TM, AKM, BKM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(208 * 1024)) != cudaSuccess)
      return HB_ERR_CUDA;
    attr = true;
  }
  dim3 grid((unsigned)(cdiv(sp.M, BM) * cdiv(sp.N, SBN)), (unsigned)batch, 1);
  gemm_small_kernel<TM, AKM, BKM><<<grid, SNT, smem, st>>>(sp);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

template <int TM>
int launch_small(const SParams& sp, int batch, cudaStream_t st) {
  if (sp.a_kmajor) return sp.b_kmajor ? launch_small2<TM, true, true>(sp, batch, st) : launch_small2<TM, true, false>(sp, batch, st);
  return sp.b_kmajor ? launch_small2<TM, false, true>(sp, batch, st) : launch_small2<TM, false, false>(sp, batch, st);
}

}  // namespace

bool gemm_small_eligible(const GemmParams& p) {
  if (p.bias || p.act != ACT_NONE || p.clip) return false;
  if (p.K < 4 || p.K > 256 || (p.K & 3) || p.batch > 65535) return false;
  if (p.M < 1 || p.N < 1) return false;
  // 16-byte aligned float4 rows for both operands (anything else goes to the k-looped kernel)
  if (!aligned16(p.A) || !aligned16(p.B) || (p.lda & 3) || (p.ldb & 3) || (p.sA & 3) || (p.sB & 3)) return false;
  return true;
}

// C may alias A only when a single CTA column covers N (N <= 128) -- checked here.
int gemm_small(const GemmParams& p, cudaStream_t st) {
  if (!gemm_small_eligible(p)) return HB_ERR_ARG;
  if (p.C == p.A && p.N > SBN) return HB_ERR_ARG;
  SParams sp;
  sp.A = p.A; sp.B = p.B; sp.C = p.C; sp.lda = p.lda; sp.ldb = p.ldb; sp.ldc = p.ldc;
  sp.sA = p.sA; sp.sB = p.sB; sp.sC = p.sC;
  sp.M = p.M; sp.N = p.N; sp.K = p.K; sp.Kp = (p.K + 3) / 4 * 4;
  sp.alpha = p.alpha; sp.beta = p.beta;
  static const int b2rk[5] = {0, 2, 1, 4, 3};
  sp.a_mode = p.a_tri; sp.b_mode = b2rk[p.b_tri]; sp.c_tri = p.c_tri;
  sp.a_kmajor = (p.transA == 0); sp.b_kmajor = (p.transB == 1);
  sp.vecA = aligned16(p.A) && (p.lda % 4 == 0) && (p.sA % 4 == 0);
  sp.vecB = aligned16(p.B) && (p.ldb % 4 == 0) && (p.sB % 4 == 0);
  sp.vecC = aligned16(p.C) && (p.ldc % 4 == 0) && (p.sC % 4 == 0);
  // few CTAs -> smaller row tiles for more parallelism
  const long long ctas64 = (long long)cdiv(p.M, 64) * cdiv(p.N, SBN) * p.batch;
  if (ctas64 < 148) return launch_small<2>(sp, p.batch, st);
  return launch_small<4>(sp, p.batch, st);
}

}  // namespace hb
