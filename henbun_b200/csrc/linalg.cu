// Blocked (recursive, all level-3) Cholesky, triangular solves and reverse-mode Cholesky.
//
//   reference ops replaced:  tf.cholesky                      (Henbun/gp/kernels.py:100-101)
//                            tf.matrix_triangular_solve        (Henbun/gp/gp.py:162,169)
//                            TF's _CholeskyGrad via minimize() (Henbun/model.py:220)
//
// Layout: row-major fp32, full-square storage, only the lower triangle is read/written.
// The recursion splits on multiples of NB; leaves (<= NB) run in one CTA out of shared memory and
// also emit the inverse of the diagonal block, so every triangular solve above the leaves is a GEMM.
#include "gemm.cuh"
#include "gemm_h2.cuh"
#include "kernels.cuh"
#include "comm.cuh"
#include "leaf_smem.cuh"
#include <cstdlib>
#include <vector>

namespace hb {

namespace {

using namespace leaf;   // NB, LDS, LEAF_THREADS, potrf_smem, trinv_smem, mm_smem, load_lower, load_full (leaf_smem.cuh)

// ------------------------------------------------------------------------------------------
// leaf kernels (one CTA per matrix; blockIdx.x = batch index)
// ------------------------------------------------------------------------------------------

// Factor in place (lower), optionally zero the strict upper triangle, optionally emit Dinv = L^{-1}.
__global__ void __launch_bounds__(LEAF_THREADS) potrf_leaf_kernel(float* A, long long lda, long long sA, int n,
                                                                  float* Dinv, long long sD, int zero_upper,
                                                                  int* err_flag, int err_base) {
  extern __shared__ float sm[];
  float* S = sm;
  float* X = sm + NB * LDS;
  float* T = sm + 2 * NB * LDS;
  float* Ab = A + (long long)blockIdx.x * sA;
  load_lower(S, Ab, lda, n, 1.f);
  potrf_smem(S, n, err_flag, err_base);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    if (j <= i) Ab[(long long)i * lda + j] = S[i * LDS + j];
    else if (zero_upper) Ab[(long long)i * lda + j] = 0.f;
  }
  if (Dinv) {
    trinv_smem(S, X, T);
    float* Db = Dinv + (long long)blockIdx.x * sD;
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Db[e] = X[(e / NB) * LDS + (e % NB)];
  }
}

// Dinv = L^{-1} for diagonal blocks of an already factored matrix; block b starts at (b*NB, b*NB).
__global__ void __launch_bounds__(LEAF_THREADS) trinv_blocks_kernel(const float* L, long long ldl, int n,
                                                                    float* Dinv) {
  extern __shared__ float sm[];
  float* S = sm;
  float* X = sm + NB * LDS;
  float* T = sm + 2 * NB * LDS;
  const int b = blockIdx.x;
  const int nb = min(NB, n - b * NB);
  load_lower(S, L + (long long)b * NB * ldl + (long long)b * NB, ldl, nb, 1.f);
  __syncthreads();
  trinv_smem(S, X, T);
  float* Db = Dinv + (long long)b * NB * NB;
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) Db[e] = X[(e / NB) * LDS + (e % NB)];
}

// Leaf of the reverse-mode recursion: G <- sym(L^{-T} Phi(L^T tril(G)) L^{-1}), written to both
// triangles of the diagonal block.  Dinv = L^{-1} of this block (NB x NB, identity padded) or null.
__global__ void __launch_bounds__(LEAF_THREADS) chol_rev_leaf_kernel(const float* L, long long ldl, long long sL,
                                                                     float* G, long long ldg, long long sG, int n,
                                                                     const float* Dinv, long long sD) {
  extern __shared__ float sm[];
  float* B0 = sm;
  float* B1 = sm + NB * LDS;
  float* B2 = sm + 2 * NB * LDS;
  const float* Lb = L + (long long)blockIdx.x * sL;
  float* Gb = G + (long long)blockIdx.x * sG;
  load_lower(B0, Lb, ldl, n, 1.f);   // L
  if (Dinv) {
    load_full(B1, Dinv + (long long)blockIdx.x * sD);
  } else {
    __syncthreads();
    trinv_smem(B0, B1, B2);          // B1 = L^{-1}, B2 scratch
  }
  load_lower(B2, Gb, ldg, n, 0.f);   // tril(Lbar)
  __syncthreads();
  // P = Phi(L^T Lbar): (L^T)[i][k] = L[k][i], nonzero for k >= i.  Result overwrites Lbar (mm_smem syncs before storing).
  mm_smem<true, false, 1>(B2, B0, B2);
  __syncthreads();
  for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    float v = B2[i * LDS + j];
    if (j > i) v = 0.f;
    else if (j == i) v *= 0.5f;
    B2[i * LDS + j] = v;
  }
  __syncthreads();
  // M1 = P * Dinv (lower x lower): k <= i   (L is dead: B0 receives M1)
  mm_smem<false, false, 2>(B0, B2, B1);
  __syncthreads();
  // S = Dinv^T * M1: (Dinv^T)[i][k] = Dinv[k][i], nonzero for k >= i
  mm_smem<true, false, 1>(B2, B1, B0);
  __syncthreads();
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e % n;
    Gb[(long long)i * ldg + j] = 0.5f * (B2[i * LDS + j] + B2[j * LDS + i]);
  }
}


// ------------------------------------------------------------------------------------------
// Panel solve by substitution: Bp (m x w, in place) <- alpha * Bp * D^{-T} (TRANS) or alpha * Bp * D^{-1}, D = w x w
// lower-triangular block (w <= NB) read where it lies (its strict upper triangle in memory is ignored).
// One thread per row, 256 rows per CTA, the row lives in shared memory as a column of Xs[k][r] (conflict-free), D in
// shared memory in "coefficient-major" form Dm[k][j] = coefficient of x_k in the equation of x_j, so that one float4
// feeds two packed FMAs (FFMA2).  Blocked by 16 columns: previous blocks are applied as 16x16 products, the diagonal
// block is solved in registers.  Backward stable (no explicit inverse), 2 w^2 flop per row like the product with the
// inverse it replaces, and no second pass over the panel.
// ------------------------------------------------------------------------------------------
constexpr int PT_ROWS = 256;
constexpr int PT_LDX = PT_ROWS + 1;
constexpr int PT_LDD = NB + 4;
constexpr size_t kPanelSmem = (size_t)(NB * PT_LDD + NB * PT_LDX + NB) * sizeof(float);

__device__ __forceinline__ void ffma2_acc(unsigned long long& acc, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <bool TRANS>
__global__ void __launch_bounds__(PT_ROWS) panel_trsm_kernel(const float* __restrict__ D, long long ldd, float* Bp, long long ldb,
                                                             int m, int w, float alpha) {
  extern __shared__ __align__(16) float psm[];
  float* Dm = psm;                         // [NB][PT_LDD]
  float* Xs = Dm + NB * PT_LDD;            // [NB][PT_LDX]
  float* rdiag = Xs + NB * PT_LDX;         // [NB]
  const int tid = threadIdx.x;
  const long long r0 = (long long)blockIdx.x * PT_ROWS;
  for (int idx = tid; idx < NB * NB; idx += PT_ROWS) {
    const int a = idx / NB, b = idx % NB;  // element D[a][b]
    float v = 0.f;
    if (a < w && b < w) { if (b <= a) v = __ldg(D + (long long)a * ldd + b); }
    else if (a == b) v = 1.f;
    if (TRANS) Dm[b * PT_LDD + a] = v;     // x_j eq.: sum_{k<=j} x_k D[j][k]  -> coefficient of x_k (k = b) for j = a
    else Dm[a * PT_LDD + b] = v;           // x_j eq.: sum_{k>=j} x_k D[k][j]  -> coefficient of x_k (k = a) for j = b
    if (a == b) rdiag[a] = 1.f / v;
  }
  for (int idx = tid; idx < PT_ROWS * NB; idx += PT_ROWS) {
    const int rr = idx / NB, k = idx % NB;
    float v = 0.f;
    if (r0 + rr < m && k < w) v = Bp[(r0 + rr) * ldb + k];
    Xs[k * PT_LDX + rr] = v;
  }
  __syncthreads();
  const int r = tid;
  const int nblk = (w + 15) / 16;
  if (r0 + r < m) {
    for (int step = 0; step < nblk; ++step) {
      const int jb = TRANS ? step : nblk - 1 - step;
      const int j0 = jb * 16;
      unsigned long long acc2[8];
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        const float lo = Xs[(j0 + 2 * p) * PT_LDX + r], hi = Xs[(j0 + 2 * p + 1) * PT_LDX + r];
        asm("mov.b64 %0, {%1, %2};" : "=l"(acc2[p]) : "f"(lo), "f"(hi));
      }
      for (int s2 = 0; s2 < step; ++s2) {
        const int kb = TRANS ? s2 : nblk - 1 - s2;
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
          const int k = kb * 16 + kk;
          const float nx = -Xs[k * PT_LDX + r];
          unsigned long long nx2;
          asm("mov.b64 %0, {%1, %1};" : "=l"(nx2) : "f"(nx));
          const ulonglong2* drow = reinterpret_cast<const ulonglong2*>(Dm + k * PT_LDD + j0);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const ulonglong2 d = drow[q];
            ffma2_acc(acc2[2 * q], nx2, d.x);
            ffma2_acc(acc2[2 * q + 1], nx2, d.y);
          }
        }
      }
      float acc[16];
#pragma unroll
      for (int p = 0; p < 8; ++p) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * p]), "=f"(acc[2 * p + 1]) : "l"(acc2[p]));
      if (TRANS) {
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float x = acc[jj] * rdiag[j0 + jj];
          acc[jj] = x;
#pragma unroll
          for (int j2 = jj + 1; j2 < 16; ++j2) acc[j2] = fmaf(-x, Dm[(j0 + jj) * PT_LDD + j0 + j2], acc[j2]);
        }
      } else {
#pragma unroll
        for (int jj = 15; jj >= 0; --jj) {
          const float x = acc[jj] * rdiag[j0 + jj];
          acc[jj] = x;
#pragma unroll
          for (int j2 = 0; j2 < jj; ++j2) acc[j2] = fmaf(-x, Dm[(j0 + jj) * PT_LDD + j0 + j2], acc[j2]);
        }
      }
#pragma unroll
      for (int jj = 0; jj < 16; ++jj) Xs[(j0 + jj) * PT_LDX + r] = acc[jj];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < PT_ROWS * NB; idx += PT_ROWS) {
    const int rr = idx / NB, k = idx % NB;
    if (r0 + rr < m && k < w) Bp[(r0 + rr) * ldb + k] = alpha * Xs[k * PT_LDX + rr];
  }
}

constexpr size_t kLeafSmem2 = 2 * NB * LDS * sizeof(float);
constexpr size_t kLeafSmem3 = 3 * NB * LDS * sizeof(float);

int ensure_attrs() {
  static bool done = false;
  if (done) return HB_OK;
  if (cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLeafSmem3) != cudaSuccess) return HB_ERR_CUDA;
  if (cudaFuncSetAttribute(trinv_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLeafSmem3) != cudaSuccess) return HB_ERR_CUDA;
  if (cudaFuncSetAttribute(chol_rev_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLeafSmem3) != cudaSuccess) return HB_ERR_CUDA;
  if (cudaFuncSetAttribute(panel_trsm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelSmem) != cudaSuccess) return HB_ERR_CUDA;
  if (cudaFuncSetAttribute(panel_trsm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPanelSmem) != cudaSuccess) return HB_ERR_CUDA;
  done = true;
  return HB_OK;
}

// split point: a multiple of NB, roughly half
inline int split_point(int n) {
  const int nblk = (n + NB - 1) / NB;
  return ((nblk + 1) / 2) * NB;
}

struct Ctx {
  cudaStream_t st;
  float* dinv;        // [nblk][NB][NB]
  float* tmp;         // [rows][NB] scratch of the refined panel solves
  int n_total = 0;    // order of the matrix being factored (selects the refinement mode)
  long long ldt;      // = NB
  int* err;
  void* tcws;         // scratch of the tensor-core GEMM engine (split-K partial tiles)
  size_t tcws_bytes;
  // look-ahead: the single-CTA leaf kernels run on a high-priority side stream while the main stream continues with
  // products that do not depend on them (null side stream = everything on the main stream)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_main = nullptr, ev_side = nullptr;
  mutable bool side_pending = false;     // a leaf is in flight on the side stream and the main stream has not joined it
  // fp16 hi/lo shadows of the finished panels (gemm_h2.cu): L in both directions, K-bar in the reverse mode; null = off
  __half *lh = nullptr, *ll = nullptr, *gh = nullptr, *gl = nullptr;
  long long ldh = 0;
  unsigned* lmax = nullptr;   // bit pattern of max |diag K| (forward) / max |L| (stand-alone reverse)
  float* lscale = nullptr;    // {s_L, 1 / s_L}
  unsigned* gmax = nullptr;   // [2 nblk]: panel maxima, then diagonal-block maxima
  float* ginv = nullptr;      // [2 nblk]: inverse scales of the K-bar panels, then of its diagonal blocks
  // block-first reverse mode (rev_block): the rows of a column block BELOW its W-wide block are finished (and split) before
  // the rows inside it, so the two parts carry separate maxima / scales; gmax / ginv then describe the in-block rows
  unsigned* gpmax = nullptr;  // [nblk]
  float* gpinv = nullptr;     // [nblk]
  int nblk = 0;
};

// Orders from which the big products run on the pre-split engine.  Below, no product has enough 256 x 256 tiles.
constexpr int H2_MIN_N = 4096;
inline bool h2_on(int n) { return opt_presplit_engine() && n >= H2_MIN_N && (n % 8) == 0; }
inline long long h2_ld(int n) { return ((long long)n + 63) / 64 * 64; }

static double h2_flops(const H2Gemm& h) {
  double f = 2.0 * (double)h.M * (double)h.N * (double)h.K;
  if (h.c_tri) {
    const double M = h.M, N = h.N;
    f *= ((M >= N) ? M * N - 0.5 * N * (N - 1.0) : 0.5 * M * (M + 1.0)) / (M * N);
  }
  if (h.a_bmode) f *= 0.5;
  return f;
}
static int run_h2(const Ctx& c, const H2Gemm& h) {
  const int slot = gemm_prof_begin(h2_flops(h), h.M, h.N, h.K, 3, c.st);
  const int rc = gemm_h2(h, c.st);
  gemm_prof_end(slot, c.st);
  return rc;
}
// split the finished panel L[r0 .. r0+rows, c0 .. c0+w) into the L shadow
static int split_L(const Ctx& c, const float* L, long long ldl, int r0, int c0, int rows, int w) {
  if (!c.lh || rows <= 0 || (w & 7)) return HB_OK;
  return h2_split(L + (long long)r0 * ldl + c0, ldl, rows, w, c.lscale, nullptr, nullptr, 0, 0,
                  c.lh + (long long)r0 * c.ldh + c0, c.ll + (long long)r0 * c.ldh + c0, c.ldh, c.st);
}

// Side stream and its two events, created once per device: a resource cache, not configuration (hb_options.lookahead = 0
// keeps a call off it).
struct SideState { cudaStream_t st = nullptr; cudaEvent_t a = nullptr, b = nullptr; bool tried = false; };
static SideState g_side[64];
static void attach_side(Ctx& c) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  SideState& s = g_side[dev];
  if (!s.tried) {
    s.tried = true;
    {
      int lo = 0, hi = 0;
      cudaDeviceGetStreamPriorityRange(&lo, &hi);
      if (cudaStreamCreateWithPriority(&s.st, cudaStreamNonBlocking, hi) != cudaSuccess) s.st = nullptr;
      if (s.st && (cudaEventCreateWithFlags(&s.a, cudaEventDisableTiming) != cudaSuccess ||
                   cudaEventCreateWithFlags(&s.b, cudaEventDisableTiming) != cudaSuccess)) s.st = nullptr;
    }
  }
  if (!opt_lookahead()) return;                 // hb_options.lookahead = 0: everything on the caller's stream
  c.side = s.st; c.ev_main = s.a; c.ev_side = s.b;
}
// side stream picks up everything issued on the main stream so far
static inline void fork_side(const Ctx& c) {
  cudaEventRecord(c.ev_main, c.st);
  cudaStreamWaitEvent(c.side, c.ev_main, 0);
}
// main stream waits for the leaf in flight on the side stream (no-op if none)
static inline void join_side(const Ctx& c) {
  if (!c.side_pending) return;
  cudaStreamWaitEvent(c.st, c.ev_side, 0);
  c.side_pending = false;
}

// Factorisations of order <= 2048 are latency-bound (2048^3 flop = 0.2 ms even on the fp32 SIMT kernels), so their
// products run in exact fp32: on the ill-conditioned notebook-size problems (config 1 / 2, cond 1e6) the split product's
// 1.5e-6 per-product error was visible next to LAPACK's (tests/probe_config2_errors.py).
// Price at n = 2000: 3 factorisations + reverse modes 17.6 -> 29.5 ms (the SIMT kernels fill 10-30 of 148 SMs on these
// shapes); gain: gradient error vs fp64 3-5x LAPACK's -> 1-3x.  hb_set_exact_below(0) turns it off.
inline int gemm_ws(const Ctx& c, GemmParams& g) {
  g.force_simt = (c.n_total > 0 && c.n_total <= opt_exact_below()) ? 1 : 0;
  g.ws = c.tcws; g.ws_bytes = c.tcws_bytes;
  return gemm(g, c.st);
}

inline float* dinv_slot(const Ctx& c, int off) { return c.dinv + (long long)(off / NB) * NB * NB; }

// Panel solves multiply by the explicit inverse of the 128 x 128 diagonal block (one in-place GEMM on the tensor-core
// engine).  That is not backward stable: on ill-conditioned matrices (the 1-D notebook inputs, cond 1e6) the factor is
// 2-3x further from the fp64 one than LAPACK's substitution-based spotrf (measured).  One step of
// iterative refinement against the triangular block itself recovers it:
//     T = B Dinv';  B <- B - T D';  T <- T + B Dinv';  B <- alpha T        (D' = D^T or D)
// at the price of two more short-K products and a copy per panel.  Mode 0 = explicit inverse only, 1 = refined,
// 2 (default) = refined for n <= 8192 (where the panels are a negligible share of the work), 3 = no inverse at all:
// panel_trsm_kernel solves the panel by substitution against the triangular block (backward stable, one pass over the
// panel).  Mode 3 measured: same accuracy as mode 1, but +27 ms on the N=65536 step (1319 vs 1292 ms: 8 warps per SM
// on a shared-memory-bound substitution lose to the tensor-core product with the inverse) -- opt-in only.
static inline bool refine_on(int n) { const int m = opt_panel_refinement(); return m == 1 || (m == 2 && n <= 8192); }

// Bp (m x w, in place) <- alpha * Bp * D^{-T} (trans) or alpha * Bp * D^{-1}, D = w x w lower block at global offset `off`
// of the factor (its strict upper triangle in memory is NOT assumed zero).
static int panel_solve(const Ctx& c, const float* D, long long ldd, int off, float* Bp, long long ldb, int m, int w,
                       bool trans, float alpha, bool refine) {
  if (opt_panel_refinement() == 3) {
    if (m <= 0 || w <= 0) return HB_OK;
    const int nb = cdiv(m, PT_ROWS);
    if (trans) panel_trsm_kernel<true><<<nb, PT_ROWS, kPanelSmem, c.st>>>(D, ldd, Bp, ldb, m, w, alpha);
    else panel_trsm_kernel<false><<<nb, PT_ROWS, kPanelSmem, c.st>>>(D, ldd, Bp, ldb, m, w, alpha);
    HB_CHECK_LAUNCH();
    return HB_OK;
  }
  GemmParams g;
  g.A = Bp; g.lda = ldb; g.B = dinv_slot(c, off); g.ldb = NB; g.transB = trans ? 1 : 0;
  g.M = m; g.N = w; g.K = w;
  if (!refine || !c.tmp) {
    g.C = Bp; g.ldc = ldb; g.alpha = alpha; g.beta = 0.f;                  // in place
    return gemm_ws(c, g);
  }
  float* T = c.tmp;
  g.C = T; g.ldc = NB; g.alpha = 1.f; g.beta = 0.f;
  HB_TRY(gemm_ws(c, g));                                                   // T = B Dinv'
  GemmParams r;
  r.A = T; r.lda = NB; r.B = D; r.ldb = ldd; r.transB = trans ? 1 : 0; r.b_tri = trans ? 2 : 1;
  r.C = Bp; r.ldc = ldb; r.M = m; r.N = w; r.K = w; r.alpha = -1.f; r.beta = 1.f;
  HB_TRY(gemm_ws(c, r));                                                   // B <- B - T D'   (the residual)
  g.beta = 1.f;
  HB_TRY(gemm_ws(c, g));                                                   // T <- T + R Dinv'
  return copy2d(Bp, ldb, T, NB, m, w, alpha, c.st);
}

// B (m x k at Bp) <- B * L^{-T},  L = k x k lower block whose diagonal starts at global offset `off`
int trsm_rlt(const Ctx& c, const float* L, long long ldl, int off, float* Bp, long long ldb, int m, int k) {
  if (m <= 0 || k <= 0) return HB_OK;
  if (k <= NB) return panel_solve(c, L, ldl, off, Bp, ldb, m, k, true, 1.f, refine_on(c.n_total));
  const int k1 = split_point(k), k2 = k - k1;
  HB_TRY(trsm_rlt(c, L, ldl, off, Bp, ldb, m, k1));
  GemmParams g;   // B2 -= B1 * L21^T
  g.A = Bp; g.lda = ldb; g.B = L + (long long)k1 * ldl; g.ldb = ldl; g.transB = 1;
  g.C = Bp + k1; g.ldc = ldb; g.M = m; g.N = k2; g.K = k1; g.alpha = -1.f; g.beta = 1.f;
  HB_TRY(gemm_ws(c, g));
  return trsm_rlt(c, L + (long long)k1 * ldl + k1, ldl, off + k1, Bp + k1, ldb, m, k2);
}

// B <- B * L^{-1}
int trsm_rln(const Ctx& c, const float* L, long long ldl, int off, float* Bp, long long ldb, int m, int k) {
  if (m <= 0 || k <= 0) return HB_OK;
  if (k <= NB) return panel_solve(c, L, ldl, off, Bp, ldb, m, k, false, 1.f, refine_on(c.n_total));
  const int k1 = split_point(k), k2 = k - k1;
  HB_TRY(trsm_rln(c, L + (long long)k1 * ldl + k1, ldl, off + k1, Bp + k1, ldb, m, k2));
  GemmParams g;   // B1 -= B2 * L21
  g.A = Bp + k1; g.lda = ldb; g.B = L + (long long)k1 * ldl; g.ldb = ldl; g.transB = 0;
  g.C = Bp; g.ldc = ldb; g.M = m; g.N = k1; g.K = k2; g.alpha = -1.f; g.beta = 1.f;
  HB_TRY(gemm_ws(c, g));
  return trsm_rln(c, L, ldl, off, Bp, ldb, m, k1);
}

// Column-recursive blocked Cholesky: factor columns [c0, c0+w) of the n x n matrix (rows c0..n-1), all updates
// from columns < c0 already applied.  Every step works on ALL rows below the diagonal block at once, so a level-w
// update is one tall GEMM and a leaf is one potrf kernel plus one in-place solve of the whole panel
// (2 n/NB - 1 GEMM launches in total; the square-recursive form needed (n/NB) log2(n/NB)).
int potrf_cols(const Ctx& c, float* A, long long lda, int c0, int w, int n, bool leaf_done = false) {
  float* D = A + (long long)c0 * lda + c0;            // diagonal block of this column range
  if (w <= NB) {
    if (!leaf_done) {
      potrf_leaf_kernel<<<1, LEAF_THREADS, kLeafSmem3, c.st>>>(D, lda, 0, w, dinv_slot(c, c0), 0, 0, c.err, c0);
      HB_CHECK_LAUNCH();
    }
    join_side(c);                                      // the look-ahead leaf (if any) must be done before the solve
    const int below = n - (c0 + w);
    if (below <= 0) return HB_OK;
    // panel <- panel * D^{-T} (in place; refined against D itself when the mode says so)
    HB_TRY(panel_solve(c, D, lda, c0, D + (long long)w * lda, lda, below, w, true, 1.f, refine_on(n)));
    return split_L(c, A, lda, c0 + w, c0, below, w);      // the panel is final: its fp16 hi/lo shadow feeds every later update
  }
  const int w1 = split_point(w), w2 = w - w1;
  HB_TRY(potrf_cols(c, A, lda, c0, w1, n, leaf_done));
  // A[c0+w1 .. n, c0+w1 .. c0+w) -= L[c0+w1 .. n, c0 .. c0+w1) * L[c0+w1 .. c0+w, c0 .. c0+w1)^T  (lower trapezoid)
  float* P = A + (long long)(c0 + w1) * lda + c0;
  const int M = n - (c0 + w1);
  GemmParams g;
  g.A = P; g.lda = lda; g.B = P; g.ldb = lda; g.transB = 1;
  g.C = P + w1; g.ldc = lda; g.M = M; g.N = w2; g.K = w1; g.alpha = -1.f; g.beta = 1.f; g.c_tri = 1;
  auto update = [&](const GemmParams& q, int row_off, int col_off) -> int {
    // q updates C = P[row_off.., w1 + col_off ..] with A = P[row_off.., 0 .. w1), B = P[col_off .., 0 .. w1)
    if (c.lh && gemm_h2_eligible(q.M, q.N, q.K)) {
      H2Gemm h;
      const long long ra = c0 + w1 + row_off, rb = c0 + w1 + col_off;
      h.a_hi = c.lh + ra * c.ldh + c0; h.a_lo = c.ll + ra * c.ldh + c0; h.lda = c.ldh; h.a_kmajor = 1;
      h.b_hi = c.lh + rb * c.ldh + c0; h.b_lo = c.ll + rb * c.ldh + c0; h.ldb = c.ldh; h.b_kmajor = 1;
      h.C = q.C; h.ldc = q.ldc; h.M = q.M; h.N = q.N; h.K = q.K; h.alpha = q.alpha; h.beta = q.beta; h.c_tri = q.c_tri;
      h.a_inv = c.lscale + 1; h.b_inv = c.lscale + 1;
      return run_h2(c, h);
    }
    GemmParams qq = q;
    return gemm_ws(c, qq);
  };
  if (!c.side || w1 > 1024) {   // splitting a long-K update costs more than hiding one 76 us leaf behind it
    HB_TRY(update(g, 0, 0));
    return potrf_cols(c, A, lda, c0 + w1, w2, n);
  }
  // look-ahead: update the next diagonal block first, factor it on the side stream, update the rest meanwhile
  const int nb = min(NB, w2);
  GemmParams d = g;                                    // rows [0, nb) x cols [0, nb), lower triangle
  d.M = nb; d.N = nb;
  HB_TRY(gemm_ws(c, d));
  fork_side(c);
  potrf_leaf_kernel<<<1, LEAF_THREADS, kLeafSmem3, c.side>>>(P + w1, lda, 0, nb, dinv_slot(c, c0 + w1), 0, 0, c.err, c0 + w1);
  HB_CHECK_LAUNCH();
  cudaEventRecord(c.ev_side, c.side);
  c.side_pending = true;
  if (M > nb) {
    GemmParams r = g;                                  // rows [nb, M) x cols [0, nb): full block
    r.A = P + (long long)nb * lda; r.C = P + w1 + (long long)nb * lda; r.M = M - nb; r.N = nb; r.c_tri = 0;
    HB_TRY(gemm_ws(c, r));
    if (w2 > nb) {
      GemmParams t = g;                                // rows [nb, M) x cols [nb, w2): lower trapezoid of its own
      t.A = P + (long long)nb * lda; t.B = P + (long long)nb * lda;
      t.C = P + w1 + (long long)nb * lda + nb; t.M = M - nb; t.N = w2 - nb; t.c_tri = 1;
      HB_TRY(update(t, nb, nb));
    }
  }
  return potrf_cols(c, A, lda, c0 + w1, w2, n, /*leaf_done=*/true);
}

int potrf_rec(const Ctx& c, float* A, long long lda, int off, int n) {
  (void)off;
  return potrf_cols(c, A, lda, 0, n, n);
}

// Column-recursive reverse mode, the mirror image of potrf_cols.  rev_cols(c0, w) owns G[c0 .. n, c0 .. c0+w):
// on entry it holds dObj/dL for these columns with every contribution of the columns to the right already applied,
// on exit dObj/dK in the full-symmetric convention (an off-diagonal entry holds half of the lower-triangle gradient).
//   w <= NB :  S  = G_below * L_DD^{-1} / 2              (in place, all rows below the diagonal block)
//              G_DD -= 2 tril(S^T L_below)               (tall reduction, K = rows below)
//              G_DD  = leaf reverse of chol(L_DD)
//   else    :  rev_cols(right half);  with R = rows below this column range, T = the right half's diagonal rows:
//              G[R, left] -= 2 G[R, right] L[T, left]              (reverse of the tall update, rows below)
//              G[T, left] -= 2 G[R, right]^T L[R, left]            (tall reduction)
//              G[T, left] -= 2 sym(G[T, right]) L[T, left]
//              rev_cols(left half)
// Launch count ~ 4 n/NB (the square recursion needed ~ 3 (n/NB) log2(n/NB)) and every GEMM spans all rows below.
// the (full, symmetric) diagonal block of K-bar the leaf just wrote gets its own scale: its entries can be orders of
// magnitude above the panel below it (short lengthscales: K ~ I)
static int split_G_diag(const Ctx& c, const float* G, long long ldg, int c0, int w, cudaStream_t st) {
  if (!c.gh || (w & 7)) return HB_OK;
  const int b = c0 / NB;
  const float* GD = G + (long long)c0 * ldg + c0;
  const long long o = (long long)c0 * c.ldh + c0;
  HB_TRY(h2_absmax(GD, ldg, w, w, 0, 0, c.gmax + c.nblk + b, st));
  return h2_split(GD, ldg, w, w, nullptr, c.gmax + c.nblk + b, c.ginv + c.nblk + b, 0, 0, c.gh + o, c.gl + o, c.ldh, st);
}

int chol_rev_cols(const Ctx& c, const float* L, long long ldl, float* G, long long ldg, int c0, int w, int n) {
  float* GD = G + (long long)c0 * ldg + c0;
  const float* LD = L + (long long)c0 * ldl + c0;
  if (w <= NB) {
    const int below = n - (c0 + w);
    if (below > 0) {
      float* Gp = GD + (long long)w * ldg;
      const float* Lp = LD + (long long)w * ldl;
      HB_TRY(panel_solve(c, LD, ldl, c0, Gp, ldg, below, w, false, 0.5f, refine_on(c.n_total)));
      if (c.gh && !(w & 7)) {      // the panel of K-bar is final: scale by its own maximum, split into the shadow
        const int b = c0 / NB;
        HB_TRY(h2_absmax(Gp, ldg, below, w, 0, 0, c.gmax + b, c.st));
        const long long o = (long long)(c0 + w) * c.ldh + c0;
        HB_TRY(h2_split(Gp, ldg, below, w, nullptr, c.gmax + b, c.ginv + b, 0, 0, c.gh + o, c.gl + o, c.ldh, c.st));
      }
      GemmParams h;
      h.A = Gp; h.lda = ldg; h.transA = 1; h.B = Lp; h.ldb = ldl; h.transB = 0;
      h.C = GD; h.ldc = ldg; h.M = w; h.N = w; h.K = below; h.alpha = -2.f; h.beta = 1.f; h.c_tri = 1;
      HB_TRY(gemm_ws(c, h));
    }
    // the leaf result (the diagonal block of K-bar) is only read by sym(G[T, right]) products further up and by the
    // Gram backward: run it on the side stream and join right before those
    if (c.side) {
      join_side(c);                                    // one leaf in flight at a time (a single pair of events)
      fork_side(c);
      chol_rev_leaf_kernel<<<1, LEAF_THREADS, kLeafSmem3, c.side>>>(LD, ldl, 0, GD, ldg, 0, w, dinv_slot(c, c0), 0);
      HB_CHECK_LAUNCH();
      HB_TRY(split_G_diag(c, G, ldg, c0, w, c.side));
      cudaEventRecord(c.ev_side, c.side);
      c.side_pending = true;
      return HB_OK;
    }
    chol_rev_leaf_kernel<<<1, LEAF_THREADS, kLeafSmem3, c.st>>>(LD, ldl, 0, GD, ldg, 0, w, dinv_slot(c, c0), 0);
    HB_CHECK_LAUNCH();
    return split_G_diag(c, G, ldg, c0, w, c.st);
  }
  const int w1 = split_point(w), w2 = w - w1;
  const int r1 = c0 + w1, rb = c0 + w, nb = n - rb;
  HB_TRY(chol_rev_cols(c, L, ldl, G, ldg, r1, w2, n));
  float* G_T_left = G + (long long)r1 * ldg + c0;           // [w2 x w1]
  const float* L_T_left = L + (long long)r1 * ldl + c0;     // [w2 x w1]
  if (nb > 0) {
    float* G_R_right = G + (long long)rb * ldg + r1;        // [nb x w2]
    float* G_R_left = G + (long long)rb * ldg + c0;         // [nb x w1]
    const float* L_R_left = L + (long long)rb * ldl + c0;   // [nb x w1]
    const bool h2 = c.gh && c.lh && !(w2 % NB);     // column blocks of G[., right] are whole 128-blocks
    if (h2 && gemm_h2_eligible(nb, w1, w2)) {
      H2Gemm h;   // G[R, left] -= 2 G[R, right] L[T, left]: A K-major with one scale per 128 columns, B = L MN-major
      const long long oa = (long long)rb * c.ldh + r1, ob = (long long)r1 * c.ldh + c0;
      h.a_hi = c.gh + oa; h.a_lo = c.gl + oa; h.lda = c.ldh; h.a_kmajor = 1; h.a_kinv = c.ginv + r1 / NB;
      h.b_hi = c.lh + ob; h.b_lo = c.ll + ob; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
      h.C = G_R_left; h.ldc = ldg; h.M = nb; h.N = w1; h.K = w2; h.alpha = -2.f; h.beta = 1.f;
      HB_TRY(run_h2(c, h));
    } else {
      GemmParams a;
      a.A = G_R_right; a.lda = ldg; a.B = L_T_left; a.ldb = ldl; a.transB = 0;
      a.C = G_R_left; a.ldc = ldg; a.M = nb; a.N = w1; a.K = w2; a.alpha = -2.f; a.beta = 1.f;
      HB_TRY(gemm_ws(c, a));
    }
    if (h2 && (gemm_h2_eligible(w2, w1, nb) || gemm_h2_splitk_eligible(w2, w1, nb, c.tcws_bytes))) {
      H2Gemm h;   // G[T, left] -= 2 G[R, right]^T L[R, left]: A MN-major (scales along M), B = L MN-major
      h.ws = c.tcws; h.ws_bytes = c.tcws_bytes;             // small output, long K: split along K
      const long long oa = (long long)rb * c.ldh + r1, ob = (long long)rb * c.ldh + c0;
      h.a_hi = c.gh + oa; h.a_lo = c.gl + oa; h.lda = c.ldh; h.a_kmajor = 0; h.a_minv = c.ginv + r1 / NB;
      h.b_hi = c.lh + ob; h.b_lo = c.ll + ob; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
      h.C = G_T_left; h.ldc = ldg; h.M = w2; h.N = w1; h.K = nb; h.alpha = -2.f; h.beta = 1.f;
      HB_TRY(run_h2(c, h));
    } else {
      GemmParams b;
      b.A = G_R_right; b.lda = ldg; b.transA = 1; b.B = L_R_left; b.ldb = ldl; b.transB = 0;
      b.C = G_T_left; b.ldc = ldg; b.M = w2; b.N = w1; b.K = nb; b.alpha = -2.f; b.beta = 1.f;
      HB_TRY(gemm_ws(c, b));
    }
  }
  join_side(c);     // the next products read the diagonal blocks of G[T, right]
  if (c.gh && c.lh && !(w2 % NB) && gemm_h2_eligible(w2, w1, w2)) {
    // G[T, left] -= 2 sym(G[T, right]) L[T, left] from the shadows: block-lower part (diagonal blocks included, they are
    // stored full and carry their own scales) + transpose of the strictly block-lower part
    H2Gemm h;
    const long long oa = (long long)r1 * c.ldh + r1, ob = (long long)r1 * c.ldh + c0;
    h.a_hi = c.gh + oa; h.a_lo = c.gl + oa; h.lda = c.ldh; h.a_kmajor = 1; h.a_bmode = 1;
    h.a_kinv = c.ginv + r1 / NB; h.a_dinv = c.ginv + c.nblk + r1 / NB;
    h.b_hi = c.lh + ob; h.b_lo = c.ll + ob; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
    h.C = G_T_left; h.ldc = ldg; h.M = w2; h.N = w1; h.K = w2; h.alpha = -2.f; h.beta = 1.f;
    HB_TRY(run_h2(c, h));
    h.a_kmajor = 0; h.a_bmode = 2; h.a_kinv = nullptr; h.a_dinv = nullptr; h.a_minv = c.ginv + r1 / NB;
    HB_TRY(run_h2(c, h));
  } else {
    GemmParams g;   // G[T, left] -= 2 sym(G[T, right]) L[T, left], sym from the lower triangle
    g.A = G + (long long)r1 * ldg + r1; g.lda = ldg; g.a_tri = 1; g.B = L_T_left; g.ldb = ldl; g.transB = 0;
    g.C = G_T_left; g.ldc = ldg; g.M = w2; g.N = w1; g.K = w2; g.alpha = -2.f; g.beta = 1.f;
    HB_TRY(gemm_ws(c, g));
    g.transA = 1; g.a_tri = 3;
    HB_TRY(gemm_ws(c, g));
  }
  return chol_rev_cols(c, L, ldl, G, ldg, c0, w1, n);
}

int chol_rev_rec(const Ctx& c, const float* L, long long ldl, float* G, long long ldg, int off, int n) {
  (void)off;
  return chol_rev_cols(c, L, ldl, G, ldg, 0, n, n);
}

// ------------------------------------------------------------------------------------------------------------------
// Right-looking ("flat") schedule over column blocks of W columns, one GPU or a column-block-cyclic group of GPUs.
//
// The column recursion above runs its narrow steps (leaf kernels, panel solves, K <= 128 updates: ~150 us per 128 columns)
// strictly between its big products, so on one GPU they are exposed time and on G GPUs they would be ALL of the time.
// Here the same pieces are scheduled as two concurrent streams:
//   chain (high priority):  F(p) = potrf_cols / chol_rev_cols of block p -- the existing recursion restricted to W columns --
//                           preceded by the one update block p still lacks, U(p-1, p);
//   main:                   U(p, j) for the blocks j after p+1 (before p-1 in the reverse mode) this rank owns, the block
//                           the chain needs next first (event), then the far ones.
// With G ranks block b belongs to rank b % G; the owner packs the finished panel (all rows below, W columns), ncclBroadcast
// on a third stream carries it to the other ranks, which unpack it into their own copy of the matrix and rebuild its fp16
// hi/lo shadow (and, for K-bar, the per-128-column scales) exactly as the owner did -- every rank ends with the complete
// factor / gradient, bit-identical, so the rest of the step needs no further exchange.
// ------------------------------------------------------------------------------------------------------------------

// A[cj.., cj .. cj+wj) -= L[cj.., p0 .. p0+pw) L[cj .. cj+wj, p0 .. p0+pw)^T   (rows cj .. n-1, lower trapezoid)
static int fwd_update(const Ctx& c, float* A, long long lda, int cj, int wj, int p0, int pw, int n) {
  const int M = n - cj;
  if (M <= 0 || wj <= 0 || pw <= 0) return HB_OK;
  float* Cp = A + (long long)cj * lda + cj;
  if (c.lh && gemm_h2_eligible(M, wj, pw)) {
    H2Gemm h;
    const long long o = (long long)cj * c.ldh + p0;
    h.a_hi = c.lh + o; h.a_lo = c.ll + o; h.lda = c.ldh; h.a_kmajor = 1;
    h.b_hi = c.lh + o; h.b_lo = c.ll + o; h.ldb = c.ldh; h.b_kmajor = 1;
    h.C = Cp; h.ldc = lda; h.M = M; h.N = wj; h.K = pw; h.alpha = -1.f; h.beta = 1.f; h.c_tri = 1;
    h.a_inv = c.lscale + 1; h.b_inv = c.lscale + 1;
    return run_h2(c, h);
  }
  GemmParams g;
  g.A = A + (long long)cj * lda + p0; g.lda = lda; g.B = g.A; g.ldb = lda; g.transB = 1;
  g.C = Cp; g.ldc = lda; g.M = M; g.N = wj; g.K = pw; g.alpha = -1.f; g.beta = 1.f; g.c_tri = 1;
  return gemm_ws(c, g);
}

// The three products that carry a finished K-bar panel `right` = columns/rows [r1, r1+w2) into the columns
// left = [cj, cj+wj), cj + wj <= r1 (chol_rev_cols' inner node with a left part that need not be adjacent):
//   G[R, left] -= 2 G[R, right] L[T, left];  G[T, left] -= 2 G[R, right]^T L[R, left];  G[T, left] -= 2 sym(G[T, right]) L[T, left]
static int rev_update(const Ctx& c, const float* L, long long ldl, float* G, long long ldg, int cj, int wj, int r1, int w2,
                      int n, bool panel_scales = false) {
  if (wj <= 0 || w2 <= 0) return HB_OK;
  const float* rinv = panel_scales ? c.gpinv : c.ginv;     // scales of the rows BELOW the right-hand range
  const int rb = r1 + w2, nb = n - rb;
  float* G_T_left = G + (long long)r1 * ldg + cj;
  const float* L_T_left = L + (long long)r1 * ldl + cj;
  const bool h2 = c.gh && c.lh && !(w2 % NB) && !(wj & 7) && !(cj & 7);
  if (nb > 0) {
    float* G_R_right = G + (long long)rb * ldg + r1;
    float* G_R_left = G + (long long)rb * ldg + cj;
    const float* L_R_left = L + (long long)rb * ldl + cj;
    if (h2 && gemm_h2_eligible(nb, wj, w2)) {
      H2Gemm h;
      const long long oa = (long long)rb * c.ldh + r1, ob = (long long)r1 * c.ldh + cj;
      h.a_hi = c.gh + oa; h.a_lo = c.gl + oa; h.lda = c.ldh; h.a_kmajor = 1; h.a_kinv = rinv + r1 / NB;
      h.b_hi = c.lh + ob; h.b_lo = c.ll + ob; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
      h.C = G_R_left; h.ldc = ldg; h.M = nb; h.N = wj; h.K = w2; h.alpha = -2.f; h.beta = 1.f;
      HB_TRY(run_h2(c, h));
    } else {
      GemmParams a;
      a.A = G_R_right; a.lda = ldg; a.B = L_T_left; a.ldb = ldl; a.transB = 0;
      a.C = G_R_left; a.ldc = ldg; a.M = nb; a.N = wj; a.K = w2; a.alpha = -2.f; a.beta = 1.f;
      HB_TRY(gemm_ws(c, a));
    }
    if (h2 && (gemm_h2_eligible(w2, wj, nb) || gemm_h2_splitk_eligible(w2, wj, nb, c.tcws_bytes))) {
      H2Gemm h;
      h.ws = c.tcws; h.ws_bytes = c.tcws_bytes;
      const long long oa = (long long)rb * c.ldh + r1, ob = (long long)rb * c.ldh + cj;
      h.a_hi = c.gh + oa; h.a_lo = c.gl + oa; h.lda = c.ldh; h.a_kmajor = 0; h.a_minv = rinv + r1 / NB;
      h.b_hi = c.lh + ob; h.b_lo = c.ll + ob; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
      h.C = G_T_left; h.ldc = ldg; h.M = w2; h.N = wj; h.K = nb; h.alpha = -2.f; h.beta = 1.f;
      HB_TRY(run_h2(c, h));
    } else {
      GemmParams b;
      b.A = G_R_right; b.lda = ldg; b.transA = 1; b.B = L_R_left; b.ldb = ldl; b.transB = 0;
      b.C = G_T_left; b.ldc = ldg; b.M = w2; b.N = wj; b.K = nb; b.alpha = -2.f; b.beta = 1.f;
      b.hint_split_waves = 1;
      HB_TRY(gemm_ws(c, b));
    }
  }
  if (h2 && gemm_h2_eligible(w2, wj, w2)) {
    H2Gemm h;
    const long long oa = (long long)r1 * c.ldh + r1, ob = (long long)r1 * c.ldh + cj;
    h.a_hi = c.gh + oa; h.a_lo = c.gl + oa; h.lda = c.ldh; h.a_kmajor = 1; h.a_bmode = 1;
    h.a_kinv = c.ginv + r1 / NB; h.a_dinv = c.ginv + c.nblk + r1 / NB;
    h.b_hi = c.lh + ob; h.b_lo = c.ll + ob; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
    h.C = G_T_left; h.ldc = ldg; h.M = w2; h.N = wj; h.K = w2; h.alpha = -2.f; h.beta = 1.f;
    HB_TRY(run_h2(c, h));
    h.a_kmajor = 0; h.a_bmode = 2; h.a_kinv = nullptr; h.a_dinv = nullptr; h.a_minv = c.ginv + r1 / NB;
    return run_h2(c, h);
  }
  GemmParams g;
  g.A = G + (long long)r1 * ldg + r1; g.lda = ldg; g.a_tri = 1; g.B = L_T_left; g.ldb = ldl; g.transB = 0;
  g.C = G_T_left; g.ldc = ldg; g.M = w2; g.N = wj; g.K = w2; g.alpha = -2.f; g.beta = 1.f;
  HB_TRY(gemm_ws(c, g));
  g.transA = 1; g.a_tri = 3;
  return gemm_ws(c, g);
}

// A received K-bar panel (columns [c0, c0+w), rows c0 .. n-1, fp32 already in place): rebuild what the owner's
// chol_rev_cols left behind for the products above -- per 128-column block the maximum and the fp16 hi/lo shadow of the
// rows below its diagonal block, and the diagonal block's own maximum and shadow.  Same data, same kernels: same bits.
static int adopt_G_panel(const Ctx& c, const float* G, long long ldg, int c0, int w, int n) {
  if (!c.gh) return HB_OK;
  for (int s = c0; s < c0 + w; s += NB) {
    const int ws = min(NB, c0 + w - s), below = n - (s + ws);
    if (ws & 7) continue;
    if (below > 0) {
      const float* Gp = G + (long long)(s + ws) * ldg + s;
      const int b = s / NB;
      HB_TRY(h2_absmax(Gp, ldg, below, ws, 0, 0, c.gmax + b, c.st));
      const long long o = (long long)(s + ws) * c.ldh + s;
      HB_TRY(h2_split(Gp, ldg, below, ws, nullptr, c.gmax + b, c.ginv + b, 0, 0, c.gh + o, c.gl + o, c.ldh, c.st));
    }
    HB_TRY(split_G_diag(c, G, ldg, s, ws, c.st));
  }
  return HB_OK;
}

// Block-first reverse mode of one column block [c0, c0+W) with nb = n - (c0+W) rows below it.  chol_rev_cols walks the
// block's 128-column leaves with ALL rows below in every step (a panel solve, an absmax, a split and a 128 x 128 x nb
// reduction per leaf); here the rows below the block are finished first, as the reverse of  X = A_R L_DD^{-T}:
//     Y = G_R L_DD^{-1} / 2          right-to-left blocked solve (rev_panel): leaf solves + products on the pre-split engine
//     G_DD -= 2 tril(Y^T X)          ONE W x W x nb product
// and only then the W x W diagonal block itself (chol_rev_cols with no rows below): its leaves touch <= W rows.
static int rev_panel(const Ctx& c, const float* L, long long ldl, float* G, long long ldg, int s, int w, int rb, int n) {
  const int nb = n - rb;
  if (w <= NB) {
    float* Gp = G + (long long)rb * ldg + s;
    HB_TRY(panel_solve(c, L + (long long)s * ldl + s, ldl, s, Gp, ldg, nb, w, false, 0.5f, refine_on(c.n_total)));
    if (c.gh && !(w & 7)) {
      const int b = s / NB;
      HB_TRY(h2_absmax(Gp, ldg, nb, w, 0, 0, c.gpmax + b, c.st));
      const long long o = (long long)rb * c.ldh + s;
      HB_TRY(h2_split(Gp, ldg, nb, w, nullptr, c.gpmax + b, c.gpinv + b, 0, 0, c.gh + o, c.gl + o, c.ldh, c.st));
    }
    return HB_OK;
  }
  const int w1 = split_point(w), w2 = w - w1, r1 = s + w1;
  HB_TRY(rev_panel(c, L, ldl, G, ldg, r1, w2, rb, n));
  // G[R, left] -= 2 Y[R, right] L[right rows, left]   (Y is final: half of the solved right part)
  if (c.gh && c.lh && !(w2 % NB) && !(w1 & 7) && gemm_h2_eligible(nb, w1, w2)) {
    H2Gemm h;
    const long long oa = (long long)rb * c.ldh + r1, ob = (long long)r1 * c.ldh + s;
    h.a_hi = c.gh + oa; h.a_lo = c.gl + oa; h.lda = c.ldh; h.a_kmajor = 1; h.a_kinv = c.gpinv + r1 / NB;
    h.b_hi = c.lh + ob; h.b_lo = c.ll + ob; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
    h.C = G + (long long)rb * ldg + s; h.ldc = ldg; h.M = nb; h.N = w1; h.K = w2; h.alpha = -2.f; h.beta = 1.f;
    HB_TRY(run_h2(c, h));
  } else {
    GemmParams a;
    a.A = G + (long long)rb * ldg + r1; a.lda = ldg; a.B = L + (long long)r1 * ldl + s; a.ldb = ldl; a.transB = 0;
    a.C = G + (long long)rb * ldg + s; a.ldc = ldg; a.M = nb; a.N = w1; a.K = w2; a.alpha = -2.f; a.beta = 1.f;
    HB_TRY(gemm_ws(c, a));
  }
  return rev_panel(c, L, ldl, G, ldg, s, w1, rb, n);
}

static int rev_block(const Ctx& c, const float* L, long long ldl, float* G, long long ldg, int c0, int W, int n) {
  const int rb = c0 + W, nb = n - rb;
  if (nb <= 0) return chol_rev_cols(c, L, ldl, G, ldg, c0, W, n);
  HB_TRY(rev_panel(c, L, ldl, G, ldg, c0, W, rb, n));
  float* GD = G + (long long)c0 * ldg + c0;
  if (c.gh && c.lh && !(W % NB) &&
      (gemm_h2_eligible(W, W, nb) || gemm_h2_splitk_eligible(W, W, nb, c.tcws_bytes))) {
    H2Gemm h;                                          // A = Y MN-major (one scale per 128 columns = per row block of op(A))
    const long long o = (long long)rb * c.ldh + c0;
    h.a_hi = c.gh + o; h.a_lo = c.gl + o; h.lda = c.ldh; h.a_kmajor = 0; h.a_minv = c.gpinv + c0 / NB;
    h.b_hi = c.lh + o; h.b_lo = c.ll + o; h.ldb = c.ldh; h.b_kmajor = 0; h.b_inv = c.lscale + 1;
    h.C = GD; h.ldc = ldg; h.M = W; h.N = W; h.K = nb; h.alpha = -2.f; h.beta = 1.f; h.c_tri = 1;
    h.ws = c.tcws; h.ws_bytes = c.tcws_bytes;
    HB_TRY(run_h2(c, h));
  } else {
    GemmParams b;
    b.A = G + (long long)rb * ldg + c0; b.lda = ldg; b.transA = 1; b.B = L + (long long)rb * ldl + c0; b.ldb = ldl; b.transB = 0;
    b.C = GD; b.ldc = ldg; b.M = W; b.N = W; b.K = nb; b.alpha = -2.f; b.beta = 1.f; b.c_tri = 1; b.hint_split_waves = 1;
    HB_TRY(gemm_ws(c, b));
  }
  return chol_rev_cols(c, L, ldl, G, ldg, c0, W, rb);   // the diagonal block: no rows below
}

// Streams and events of the flat schedule, created once per device (a resource cache like g_side).
struct FlatRes {
  cudaStream_t chain = nullptr, comm = nullptr;
  cudaEvent_t fork = nullptr, panel = nullptr, near_ = nullptr, packed = nullptr, arrived = nullptr, join = nullptr, join2 = nullptr;
  cudaEvent_t stage_free[2] = {nullptr, nullptr};
  bool tried = false, ok = false;
};
static FlatRes g_flat[64];
static FlatRes* flat_res() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  FlatRes& f = g_flat[dev];
  if (!f.tried) {
    f.tried = true;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    bool ok = cudaStreamCreateWithPriority(&f.chain, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&f.comm, cudaStreamNonBlocking, hi) == cudaSuccess;
    cudaEvent_t* evs[] = {&f.fork, &f.panel, &f.near_, &f.packed, &f.arrived, &f.join, &f.join2, &f.stage_free[0], &f.stage_free[1]};
    for (cudaEvent_t* e : evs) ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    f.ok = ok;
  }
  return f.ok ? &f : nullptr;
}

// Optional timeline of the schedule (hb_flat_trace_begin / _end): timing events at the stations of every block --
// tag 0 chain: F(p) starts | 1 chain: F(p) done | 2 chain: panel p available (sent / received + unpacked) |
// 3 chain: its own update U(p, p+-1) done | 4 main: U(p, .) may start | 5 main: U(p, .) all issued work done.
struct TraceRow { int tag, p; cudaEvent_t ev; };
static std::vector<TraceRow> g_trace;
static bool g_trace_on = false;
static inline void trace(int tag, int p, cudaStream_t st) {
  if (!g_trace_on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  g_trace.push_back({tag, p, e});
}

static inline int bcast_wait(cudaStream_t waiter, cudaEvent_t ev, cudaStream_t recorder) {
  if (cudaEventRecord(ev, recorder) != cudaSuccess || cudaStreamWaitEvent(waiter, ev, 0) != cudaSuccess) return HB_ERR_CUDA;
  return HB_OK;
}

struct FlatPlan {
  int n, W, P;
  DistEnv d;
  FlatRes* r;
  float* stage[2];
  int owner(int b) const { return (b / d.turn) % d.world; }          // `turn` consecutive blocks per rank, then the next rank
  bool mine(int b) const { return b >= 0 && b < P && owner(b) == d.rank; }
  int c0(int b) const { return b * W; }
  int bw(int b) const { return min(W, n - b * W); }
};

// owner: pack the finished panel and send it; others: receive, unpack (chain stream), `adopt` rebuilds the shadows
template <class Adopt>
static int exchange_panel(const FlatPlan& f, const Ctx& cc, float* M, long long ldm, int p, Adopt adopt) {
  if (f.d.world <= 1) return HB_OK;
  const int c0 = f.c0(p), wp = f.bw(p), rows = f.n - c0, s = p & 1;
  float* stage = f.stage[s];
  float* panel = M + (long long)c0 * ldm + c0;
  const size_t count = (size_t)rows * wp;
  if (f.mine(p)) {
    if (cudaStreamWaitEvent(cc.st, f.r->stage_free[s], 0) != cudaSuccess) return HB_ERR_CUDA;   // the broadcast of p - 2 left this buffer
    HB_TRY(copy2d(stage, wp, panel, ldm, rows, wp, 1.f, cc.st));
    HB_TRY(bcast_wait(f.r->comm, f.r->packed, cc.st));
    HB_TRY(comm_bcast_f32(f.d.comm, stage, count, f.owner(p), f.r->comm));
    if (cudaEventRecord(f.r->stage_free[s], f.r->comm) != cudaSuccess) return HB_ERR_CUDA;
    return HB_OK;
  }
  if (cudaStreamWaitEvent(f.r->comm, f.r->stage_free[s], 0) != cudaSuccess) return HB_ERR_CUDA;   // panel p - 2 has been unpacked
  HB_TRY(comm_bcast_f32(f.d.comm, stage, count, f.owner(p), f.r->comm));
  HB_TRY(bcast_wait(cc.st, f.r->arrived, f.r->comm));
  HB_TRY(copy2d(panel, ldm, stage, wp, rows, wp, 1.f, cc.st));
  if (cudaEventRecord(f.r->stage_free[s], cc.st) != cudaSuccess) return HB_ERR_CUDA;
  return adopt(c0, wp);
}

static int potrf_flat(const Ctx& base, const FlatPlan& f, float* A, long long lda, void* chain_tcws) {
  const int n = f.n, P = f.P;
  // hb_options.lookahead = 0: the same launches, all on the caller's stream (per-launch timing of the step's own kernels)
  Ctx cc = base;  cc.st = opt_lookahead() ? f.r->chain : base.st; cc.tcws = chain_tcws; cc.side_pending = false;     // panel chain (keeps the leaf look-ahead stream)
  Ctx bc = base;  bc.side = nullptr; bc.side_pending = false;                            // trailing updates on the caller's stream
  HB_TRY(bcast_wait(cc.st, f.r->fork, base.st));
  if (f.d.world > 1 && cudaStreamWaitEvent(f.r->comm, f.r->fork, 0) != cudaSuccess) return HB_ERR_CUDA;
  int pend = 0;                                            // first panel the far blocks have not received yet
  for (int p = 0; p < P; ++p) {
    const int c0 = f.c0(p), wp = f.bw(p);
    if (f.mine(p)) {
      trace(0, p, cc.st);
      HB_TRY(potrf_cols(cc, A, lda, c0, wp, n));
      join_side(cc);
      trace(1, p, cc.st);
    }
    HB_TRY(exchange_panel(f, cc, A, lda, p, [&](int q0, int w) { return split_L(cc, A, lda, q0, q0, n - q0, w); }));
    trace(2, p, cc.st);
    if (p + 1 >= P) break;
    if (cudaEventRecord(f.r->panel, cc.st) != cudaSuccess) return HB_ERR_CUDA;
    if (f.mine(p + 1)) {                                 // the chain's own update: block p+1 lacks only panel p
      if (p >= 1 && cudaStreamWaitEvent(cc.st, f.r->near_, 0) != cudaSuccess) return HB_ERR_CUDA;   // U(p-1, p+1) ran on main
      HB_TRY(fwd_update(cc, A, lda, f.c0(p + 1), f.bw(p + 1), c0, wp, n));
      trace(3, p, cc.st);
    }
    if (p + 2 >= P) continue;
    if (cudaStreamWaitEvent(bc.st, f.r->panel, 0) != cudaSuccess) return HB_ERR_CUDA;
    trace(4, p, bc.st);
    // Blocks from p+3 on are "far": they have every panel before `pend` applied and take the pending ones [pend, p] as ONE
    // product with K = all their columns, `batch` panels at a time; block p+2 leaves the far set now and is brought up to date.
    const int k0 = f.c0(pend), kw = c0 + wp - k0;
    if (f.mine(p + 2)) {
      HB_TRY(fwd_update(bc, A, lda, f.c0(p + 2), f.bw(p + 2), k0, kw, n));
      if (cudaEventRecord(f.r->near_, bc.st) != cudaSuccess) return HB_ERR_CUDA;
    }
    if (p + 3 < P && (p - pend + 1 >= f.d.batch || p + 4 >= P)) {
      if (f.d.world == 1) {
        HB_TRY(fwd_update(bc, A, lda, f.c0(p + 3), n - f.c0(p + 3), k0, kw, n));                 // everything beyond, one trapezoid
      } else {
        for (int j = p + 3; j < P; ++j)
          if (f.mine(j)) HB_TRY(fwd_update(bc, A, lda, f.c0(j), f.bw(j), k0, kw, n));
      }
      pend = p + 1;
    }
    trace(5, p, bc.st);
  }
  HB_TRY(bcast_wait(base.st, f.r->join, cc.st));
  if (f.d.world > 1) HB_TRY(bcast_wait(base.st, f.r->join2, f.r->comm));
  return HB_OK;
}

static int chol_rev_flat(const Ctx& base, const FlatPlan& f, const float* L, long long ldl, float* G, long long ldg,
                         void* chain_tcws) {
  const int n = f.n, P = f.P;
  Ctx cc = base;  cc.st = opt_lookahead() ? f.r->chain : base.st; cc.tcws = chain_tcws; cc.side_pending = false;
  Ctx bc = base;  bc.side = nullptr; bc.side_pending = false;
  HB_TRY(bcast_wait(cc.st, f.r->fork, base.st));
  if (f.d.world > 1 && cudaStreamWaitEvent(f.r->comm, f.r->fork, 0) != cudaSuccess) return HB_ERR_CUDA;
  // Block-first (rev_block) on one GPU, where the step is bound by GPU work and the single W x W x n product replaces 2 W / 128
  // narrow reductions (-10 ms at N=65536).  Over several ranks the block chain is the critical path and rev_block's is the
  // longer one (solve of all rows, then the big product, then the diagonal block: measured +20 ms at 8 GPUs), so the owner
  // runs the column recursion on its block and the other ranks rebuild its single set of scales (adopt_G_panel).
  const bool bf = f.d.world == 1;
  for (int p = P - 1; p >= 0; --p) {
    const int c0 = f.c0(p), wp = f.bw(p);
    if (f.mine(p)) {
      trace(0, p, cc.st);
      if (bf) HB_TRY(rev_block(cc, L, ldl, G, ldg, c0, wp, n));
      else HB_TRY(chol_rev_cols(cc, L, ldl, G, ldg, c0, wp, n));
      join_side(cc);
      trace(1, p, cc.st);
    }
    HB_TRY(exchange_panel(f, cc, G, ldg, p, [&](int q0, int w) { return adopt_G_panel(cc, G, ldg, q0, w, n); }));
    trace(2, p, cc.st);
    if (p == 0) break;
    if (cudaEventRecord(f.r->panel, cc.st) != cudaSuccess) return HB_ERR_CUDA;
    if (f.mine(p - 1)) {
      if (p + 1 < P && cudaStreamWaitEvent(cc.st, f.r->near_, 0) != cudaSuccess) return HB_ERR_CUDA;   // V(p+1, p-1) ran on main
      HB_TRY(rev_update(cc, L, ldl, G, ldg, f.c0(p - 1), f.bw(p - 1), c0, wp, n, bf));
      trace(3, p, cc.st);
    }
    if (p < 2) continue;
    if (cudaStreamWaitEvent(bc.st, f.r->panel, 0) != cudaSuccess) return HB_ERR_CUDA;
    trace(4, p, bc.st);
    // Every finished panel is applied on its own (hb_dist.batch concerns the forward pass only): the rows of a K-bar panel
    // below its block and the rows inside it carry different scales (rev_block), which a right-hand range of several blocks
    // would mix.
    const int r1 = c0, w2 = wp;
    if (f.mine(p - 2)) {
      HB_TRY(rev_update(bc, L, ldl, G, ldg, f.c0(p - 2), f.bw(p - 2), r1, w2, n, bf));
      if (cudaEventRecord(f.r->near_, bc.st) != cudaSuccess) return HB_ERR_CUDA;
    }
    if (p >= 3) {
      if (f.d.world == 1) {
        HB_TRY(rev_update(bc, L, ldl, G, ldg, 0, f.c0(p - 2), r1, w2, n, bf));                   // all columns further left at once
      } else {
        for (int j = p - 3; j >= 0; --j)
          if (f.mine(j)) HB_TRY(rev_update(bc, L, ldl, G, ldg, f.c0(j), f.bw(j), r1, w2, n, bf));
      }
    }
    trace(5, p, bc.st);
  }
  HB_TRY(bcast_wait(base.st, f.r->join, cc.st));
  if (f.d.world > 1) HB_TRY(bcast_wait(base.st, f.r->join2, f.r->comm));
  return HB_OK;
}

}  // namespace

// The generation-2 tensor-core engine consumes operands in place, so the factorisations need no GEMM scratch
// (generation 1 needed hi/lo copies of the largest operand pair: n^2 floats).
// What remains is the split-K scratch of the tall reductions in the reverse mode (partial output tiles).
static size_t tc_bytes_for(int, int n) { return n >= 8 * NB ? (size_t)40 << 20 : 0; }

static size_t base_bytes(long long rows, int n) {
  const long long nblk = (n + NB - 1) / NB;
  return ((size_t)(nblk * NB * NB + rows * NB) * sizeof(float) + 256 + 255) / 256 * 256;
}

// fp16 hi/lo shadows of L and K-bar + their scales (orders >= H2_MIN_N only)
static size_t h2_scale_bytes(int n) { return ((size_t)(16 + 6 * ((n + NB - 1) / NB)) * 4 + 255) / 256 * 256; }
static size_t h2_shadow_bytes(int n) { return ((size_t)n * h2_ld(n) * 2 + 255) / 256 * 256; }
static size_t h2_bytes_for(int n) { return (n >= H2_MIN_N && n % 8 == 0) ? h2_scale_bytes(n) + 4 * h2_shadow_bytes(n) : 0; }

static size_t core_bytes(int n) { return base_bytes(n, n) + tc_bytes_for(-1, n) + h2_bytes_for(n); }
// Orders from which the right-looking two-stream schedule is available inside potrf_lower / potrf_lower_bwd: its chain
// stream needs a split-K scratch of its own.
constexpr int FLAT_MIN_N = 8192;
static size_t chain_bytes(int n) { return n >= FLAT_MIN_N ? tc_bytes_for(-1, n) + 256 : 0; }
size_t potrf_workspace_bytes(int n) { return core_bytes(n) + chain_bytes(n); }
// block width of the schedule for one GPU: 0 = column recursion
static int auto_block(int n) {
  const int s = opt_schedule();
  if (s == 1 || n < FLAT_MIN_N) return 0;
  if (s >= NB) return s / NB * NB;
  if (n < 32768) return 0;
  return max(2048, n / 16 / NB * NB);      // measured at n = 65536: 4096-column blocks 738 ms (potrf + reverse), 8192: 755, 2048: 770
}


static int make_ctx(Ctx& c, int n, void* ws, size_t ws_bytes, int* err, cudaStream_t st) {
  if (ws_bytes < core_bytes(n) || !ws) return HB_ERR_WORKSPACE;
  const long long nblk = (n + NB - 1) / NB;
  c.st = st;
  c.dinv = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  c.tmp = c.dinv + nblk * NB * NB;
  c.ldt = NB;
  c.err = err;
  c.tcws = reinterpret_cast<char*>(c.dinv) + base_bytes(n, n) - 256;
  c.tcws_bytes = tc_bytes_for(-1, n);
  c.n_total = n;
  if (h2_on(n) && h2_bytes_for(n)) {
    char* hb = reinterpret_cast<char*>(c.tcws) + c.tcws_bytes;
    c.nblk = (int)nblk;
    c.lmax = reinterpret_cast<unsigned*>(hb);
    c.lscale = reinterpret_cast<float*>(hb) + 4;
    c.gmax = reinterpret_cast<unsigned*>(hb) + 8;
    c.ginv = reinterpret_cast<float*>(hb) + 8 + 2 * nblk;
    c.gpmax = reinterpret_cast<unsigned*>(hb) + 8 + 4 * nblk;
    c.gpinv = reinterpret_cast<float*>(hb) + 8 + 5 * nblk;
    hb += h2_scale_bytes(n);
    const size_t sb = h2_shadow_bytes(n);
    c.lh = reinterpret_cast<__half*>(hb); c.ll = reinterpret_cast<__half*>(hb + sb);
    c.gh = reinterpret_cast<__half*>(hb + 2 * sb); c.gl = reinterpret_cast<__half*>(hb + 3 * sb);
    c.ldh = h2_ld(n);
  }
  return HB_OK;
}

// In-place lower Cholesky of `batch` n x n matrices.  Only the lower triangle is read.
int potrf_lower(float* A, long long lda, long long strideA, int n, int batch, int zero_upper, void* ws,
                size_t ws_bytes, int* err_flag, cudaStream_t st) {
  if (n < 0 || batch < 0 || (n > 0 && (!A || lda < n))) return HB_ERR_ARG;
  if (n == 0 || batch == 0) return HB_OK;
  HB_TRY(ensure_attrs());
  if (n <= NB) {
    potrf_leaf_kernel<<<batch, LEAF_THREADS, kLeafSmem3, st>>>(A, lda, strideA, n, nullptr, 0, zero_upper, err_flag, 0);
    HB_CHECK_LAUNCH();
    return HB_OK;
  }
  if (batch == 1 && auto_block(n) && ws_bytes >= potrf_workspace_bytes(n)) {
    DistEnv d; d.block = auto_block(n);
    HB_TRY(potrf_lower_dist(A, lda, n, d, ws, ws_bytes, err_flag, st));
    return zero_upper ? zero_strict_upper(A, lda, n, st) : HB_OK;
  }
  Ctx c;
  HB_TRY(make_ctx(c, n, ws, ws_bytes, err_flag, st));
  attach_side(c);
  for (int b = 0; b < batch; ++b) {
    float* Ab = A + (long long)b * strideA;
    if (c.lh) {     // |L_ij| <= sqrt(max_i K_ii): the scale of L's shadow is known before the first panel exists
      if (cudaMemsetAsync(c.lmax, 0, 16, st) != cudaSuccess) return HB_ERR_CUDA;
      HB_TRY(h2_diag_absmax(Ab, lda, n, c.lmax, st));
      HB_TRY(h2_scale_from_max(c.lmax, 1, c.lscale, st));
    }
    HB_TRY(potrf_rec(c, Ab, lda, 0, n));
    join_side(c);
    if (zero_upper) HB_TRY(zero_strict_upper(Ab, lda, n, st));
  }
  return HB_OK;
}

// Reverse mode: on entry G's lower triangle holds dObj/dL, on exit dObj/dK (full-symmetric
// convention, lower triangle valid).  L is the factor from potrf_lower (lower triangle read).
int potrf_lower_bwd(const float* L, long long ldl, long long strideL, float* G, long long ldg, long long strideG,
                    int n, int batch, void* ws, size_t ws_bytes, cudaStream_t st, int l_shadow_valid) {
  if (n < 0 || batch < 0 || (n > 0 && (!L || !G || ldl < n || ldg < n))) return HB_ERR_ARG;
  if (n == 0 || batch == 0) return HB_OK;
  HB_TRY(ensure_attrs());
  if (n <= NB) {
    chol_rev_leaf_kernel<<<batch, LEAF_THREADS, kLeafSmem3, st>>>(L, ldl, strideL, G, ldg, strideG, n, nullptr, 0);
    HB_CHECK_LAUNCH();
    return HB_OK;
  }
  if (batch == 1 && auto_block(n) && ws_bytes >= potrf_workspace_bytes(n)) {
    DistEnv d; d.block = auto_block(n);
    return potrf_lower_bwd_dist(L, ldl, G, ldg, n, d, ws, ws_bytes, st, l_shadow_valid);
  }
  Ctx c;
  HB_TRY(make_ctx(c, n, ws, ws_bytes, nullptr, st));
  attach_side(c);
  const int nblk = (n + NB - 1) / NB;
  for (int b = 0; b < batch; ++b) {
    const float* Lb = L + (long long)b * strideL;
    if (c.gh) {
      if (cudaMemsetAsync(c.gmax, 0, (size_t)2 * c.nblk * 4, st) != cudaSuccess) return HB_ERR_CUDA;
      if (!(l_shadow_valid && batch == 1)) {     // stand-alone call: L's shadow has to be built from L itself
        if (cudaMemsetAsync(c.lmax, 0, 16, st) != cudaSuccess) return HB_ERR_CUDA;
        HB_TRY(h2_absmax(Lb, ldl, n, n, 1, 0, c.lmax, st));
        HB_TRY(h2_scale_from_max(c.lmax, 0, c.lscale, st));
        HB_TRY(h2_split(Lb, ldl, n, n, c.lscale, nullptr, nullptr, 1, 0, c.lh, c.ll, c.ldh, st));
      }
    }
    trinv_blocks_kernel<<<nblk, LEAF_THREADS, kLeafSmem3, st>>>(Lb, ldl, n, c.dinv);
    HB_CHECK_LAUNCH();
    HB_TRY(chol_rev_rec(c, Lb, ldl, G + (long long)b * strideG, ldg, 0, n));
    join_side(c);
  }
  return HB_OK;
}

// X <- X * L^{-T} (trans=1) or X * L^{-1} (trans=0), L n x n lower from potrf_lower, X m x n in place.
int trsm_right_lower(const float* L, long long ldl, float* X, long long ldx, int m, int n, int trans, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
  if (m < 0 || n < 0 || (n > 0 && m > 0 && (!L || !X || ldl < n || ldx < n))) return HB_ERR_ARG;
  if (m == 0 || n == 0) return HB_OK;
  HB_TRY(ensure_attrs());
  const long long nblk = (n + NB - 1) / NB;
  if (!ws || ws_bytes < trsm_workspace_bytes(m, n)) return HB_ERR_WORKSPACE;
  Ctx c;
  c.st = st;
  c.dinv = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  c.tmp = c.dinv + nblk * NB * NB;
  c.ldt = NB;
  c.err = nullptr;
  c.tcws = reinterpret_cast<char*>(c.dinv) + base_bytes(m, n) - 256;
  c.tcws_bytes = tc_bytes_for(m, n);
  c.n_total = n;
  trinv_blocks_kernel<<<(int)nblk, LEAF_THREADS, kLeafSmem3, st>>>(L, ldl, n, c.dinv);
  HB_CHECK_LAUNCH();
  return trans ? trsm_rlt(c, L, ldl, 0, X, ldx, m, n) : trsm_rln(c, L, ldl, 0, X, ldx, m, n);
}

size_t trsm_workspace_bytes(int m, int n) { return base_bytes(m, n) + tc_bytes_for(m, n); }

int flat_trace_begin() {
  for (auto& r : g_trace) cudaEventDestroy(r.ev);
  g_trace.clear();
  g_trace_on = true;
  return HB_OK;
}
// rows of {tag, block, ms since the first station}; returns the number of stations recorded (synchronises)
int flat_trace_end(double* out3, int capacity) {
  g_trace_on = false;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  const int nrow = (int)g_trace.size();
  for (int i = 0; i < nrow && i < capacity; ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, g_trace[0].ev, g_trace[i].ev);
    out3[3 * i] = g_trace[i].tag; out3[3 * i + 1] = g_trace[i].p; out3[3 * i + 2] = ms;
  }
  for (auto& r : g_trace) cudaEventDestroy(r.ev);
  g_trace.clear();
  return nrow;
}

// ---- flat / column-block-cyclic entry points ------------------------------------------------------------------------
// workspace = [ potrf_workspace_bytes(n) | split-K scratch of the chain stream | 2 panel staging buffers (world > 1) ]
static size_t stage_bytes(int n, int block) { return ((size_t)n * block * sizeof(float) + 255) / 256 * 256; }
size_t potrf_dist_workspace_bytes(int n, const DistEnv& d) {
  return core_bytes(n) + tc_bytes_for(-1, n) + 256 + (d.world > 1 ? 2 * stage_bytes(n, d.block) : 0);
}

static int make_flat(FlatPlan& f, void*& chain_tcws, int n, const DistEnv& d, void* ws, size_t ws_bytes) {
  if (d.world < 1 || d.rank < 0 || d.rank >= d.world || d.block < NB || (d.block % NB) || d.batch < 1 || d.turn < 1 || (d.world > 1 && !d.comm))
    return HB_ERR_ARG;
  if (!ws || ws_bytes < potrf_dist_workspace_bytes(n, d)) return HB_ERR_WORKSPACE;
  f.n = n; f.W = d.block; f.P = (n + d.block - 1) / d.block; f.d = d;
  f.r = flat_res();
  if (!f.r) return HB_ERR_CUDA;
  char* p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + core_bytes(n)) & ~uintptr_t(255));
  chain_tcws = p; p += tc_bytes_for(-1, n);
  f.stage[0] = reinterpret_cast<float*>(p);
  f.stage[1] = reinterpret_cast<float*>(p + stage_bytes(n, d.block));
  return HB_OK;
}

int potrf_lower_dist(float* A, long long lda, int n, const DistEnv& d, void* ws, size_t ws_bytes, int* err_flag,
                     cudaStream_t st) {
  if (n <= 0 || !A || lda < n) return HB_ERR_ARG;
  HB_TRY(ensure_attrs());
  FlatPlan f; void* ctws = nullptr;
  HB_TRY(make_flat(f, ctws, n, d, ws, ws_bytes));
  Ctx c;
  HB_TRY(make_ctx(c, n, ws, core_bytes(n), err_flag, st));
  attach_side(c);
  if (c.lh) {
    if (cudaMemsetAsync(c.lmax, 0, 16, st) != cudaSuccess) return HB_ERR_CUDA;
    HB_TRY(h2_diag_absmax(A, lda, n, c.lmax, st));
    HB_TRY(h2_scale_from_max(c.lmax, 1, c.lscale, st));
  }
  return potrf_flat(c, f, A, lda, ctws);
}

int potrf_lower_bwd_dist(const float* L, long long ldl, float* G, long long ldg, int n, const DistEnv& d, void* ws,
                         size_t ws_bytes, cudaStream_t st, int l_shadow_valid) {
  if (n <= 0 || !L || !G || ldl < n || ldg < n) return HB_ERR_ARG;
  HB_TRY(ensure_attrs());
  FlatPlan f; void* ctws = nullptr;
  HB_TRY(make_flat(f, ctws, n, d, ws, ws_bytes));
  Ctx c;
  HB_TRY(make_ctx(c, n, ws, core_bytes(n), nullptr, st));
  attach_side(c);
  const int nblk = (n + NB - 1) / NB;
  if (c.gh) {
    if (cudaMemsetAsync(c.gmax, 0, (size_t)2 * c.nblk * 4, st) != cudaSuccess) return HB_ERR_CUDA;
    if (cudaMemsetAsync(c.gpmax, 0, (size_t)c.nblk * 4, st) != cudaSuccess) return HB_ERR_CUDA;
    if (!l_shadow_valid) {
      if (cudaMemsetAsync(c.lmax, 0, 16, st) != cudaSuccess) return HB_ERR_CUDA;
      HB_TRY(h2_absmax(L, ldl, n, n, 1, 0, c.lmax, st));
      HB_TRY(h2_scale_from_max(c.lmax, 0, c.lscale, st));
      HB_TRY(h2_split(L, ldl, n, n, c.lscale, nullptr, nullptr, 1, 0, c.lh, c.ll, c.ldh, st));
    }
  }
  trinv_blocks_kernel<<<nblk, LEAF_THREADS, kLeafSmem3, st>>>(L, ldl, n, c.dinv);
  HB_CHECK_LAUNCH();
  return chol_rev_flat(c, f, L, ldl, G, ldg, ctws);
}

}  // namespace hb
