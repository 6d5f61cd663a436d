// GEMM engine interface: C[MxN] = epi(alpha * op(A)[MxK] * op(B)[KxN] + beta * C), row-major fp32.
#pragma once
#include "common.cuh"

struct hb_options;   // include/henbun_b200.h

namespace hb {

// Triangular masks are expressed in the index space of op(A) (m,k) and op(B) (k,n):
//   A: 1 keep k<=m (lower)   2 keep k>=m (upper)   3 keep k>m (strict upper)  4 keep k<m (strict lower)
//   B: 1 keep n<=k (lower)   2 keep n>=k (upper)   3 keep n>k (strict upper)  4 keep n<k (strict lower)
//   C: 1 compute/write only j<=i (tiles entirely above the diagonal are skipped)
enum { ACT_NONE = 0, ACT_SIGMOID = 1, ACT_RELU = 2, ACT_TANH = 3 };

struct GemmParams {
  const float* A = nullptr;
  const float* B = nullptr;
  float* C = nullptr;
  long long lda = 0, ldb = 0, ldc = 0;
  long long sA = 0, sB = 0, sC = 0;  // batch strides (elements); 0 broadcasts
  int batch = 1;
  int M = 0, N = 0, K = 0;
  float alpha = 1.f, beta = 0.f;
  int transA = 0;  // 0: A stored [M x K]; 1: A stored [K x M]
  int transB = 0;  // 0: B stored [K x N]; 1: B stored [N x K]
  int a_tri = 0, b_tri = 0, c_tri = 0;
  const float* bias = nullptr;  // optional [N] (+ batch stride sBias), added before the activation
  long long sBias = 0;
  int act = ACT_NONE;
  int clip = 0;
  float clip_lo = -50.f, clip_hi = 50.f;
  int force_simt = 0;      // exact fp32 products (SIMT kernels) whatever the global engine mode says
  int hint_split_waves = 0;  // tcgen05 engine: consider split-K up to 147 tiles (default: < 74), split count by wave efficiency
  int hint_bn128 = 0;      // tcgen05 engine: 128-wide output tiles even when N > 128 (more CTAs for short-M products
                           // that cannot use split-K: triangular masks / bias epilogue)
  void* ws = nullptr;      // optional scratch for the tcgen05 engine (hi/lo operand copies)
  size_t ws_bytes = 0;
};

// precision/engine selection: 0 = auto (tensor cores when the shape qualifies), 1 = force SIMT fp32,
// 2 = force tcgen05 3xTF32 (error if the shape does not qualify)
int gemm(const GemmParams& p, cudaStream_t stream);
int gemm_simt(const GemmParams& p, cudaStream_t stream);
int gemm_small(const GemmParams& p, cudaStream_t stream);      // K <= 256, whole-K staging (latency-oriented); C may alias A if N <= 128
bool gemm_small_eligible(const GemmParams& p);
// Options of the C call in flight on this thread (include/henbun_b200.h: hb_options; defaults outside any call).  The
// extern "C" entry points install them for their own duration (capi.cu: OptScope); the library keeps no mutable configuration.
int opt_gemm_engine();
int opt_exact_below();
int opt_panel_refinement();
int opt_presplit_engine();
int opt_small_gp_kernel();
int opt_tc_option();
int opt_lookahead();
int opt_schedule();
struct OptScope {     // installs the caller's options for the duration of one extern "C" call (capi.cu)
  const ::hb_options* prev;
  explicit OptScope(const ::hb_options* o);
  ~OptScope();
};
inline int get_tc_option() { return opt_tc_option(); }
inline int get_gemm_engine() { return opt_gemm_engine(); }
// per-launch timing hooks for launches that bypass hb::gemm (the pre-split engine): slot = prof_begin(...); launch; prof_end(slot)
int gemm_prof_begin(double useful_flops, int M, int N, int K, int kind, cudaStream_t st);   // -1 when profiling is off
void gemm_prof_end(int slot, cudaStream_t st);

}  // namespace hb
