// Interface of the pre-split fp16 hi/lo tensor-core engine (gemm_h2.cu) and its operand-preparation kernels.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace hb {

constexpr int H2_TARGET_EXP = 7;     // operands are scaled so that their absolute maximum lies in [2^6, 2^7]

// C[M x N] = alpha * sa * sb * op(A) op(B) + beta * C, fp32 C.  op(A) [M x K] and op(B)^T [N x K] are given as fp16
// hi/lo shadow pairs holding x * s (s a power of two), either K-major (stored [rows x K]) or MN-major (stored [K x rows]).
// Inverse scales (device memory): a_inv / b_inv scalars; a_kinv[k / 128] per 128-block along K; a_minv[m / 128] per
// 128-row block of M (each may be null = 1).
// a_bmode (square op(A), 128-blocks): 0 all | 1 keep k-block <= row-block, the diagonal blocks carrying their own
// inverse scales a_dinv[k / 128] (null: a_kinv) | 2 keep k-block > row-block.
struct H2Gemm {
  const __half *a_hi = nullptr, *a_lo = nullptr; long long lda = 0; int a_kmajor = 1;
  const __half *b_hi = nullptr, *b_lo = nullptr; long long ldb = 0; int b_kmajor = 1;
  float* C = nullptr; long long ldc = 0;
  int M = 0, N = 0, K = 0;
  float alpha = 1.f, beta = 0.f;
  int c_tri = 0, a_bmode = 0;
  const float *a_inv = nullptr, *a_kinv = nullptr, *a_dinv = nullptr, *a_minv = nullptr, *b_inv = nullptr;
  void* ws = nullptr;        // optional scratch: long-K products with few output tiles are split along K (partial tiles
  size_t ws_bytes = 0;       // here, then one deterministic reduction pass)
};
bool gemm_h2_eligible(int M, int N, int K);
// few 256 x 256 tiles but a long K: worth the pre-split engine when the caller can lend H2Gemm::ws (split-K)
bool gemm_h2_splitk_eligible(int M, int N, int K, size_t ws_bytes);
int splitk_reduce(float* C, long long ldc, const float* part, long long pstride, int M, int N, int nsplit, float beta, int c_tri,
                  cudaStream_t st);
int gemm_h2(const H2Gemm& g, cudaStream_t st);

// atomicMax of |A| over a [rows x cols] block into *out_bits (bit pattern of a non-negative float; zero it first)
int h2_absmax(const float* A, long long ld, long long rows, int cols, int lower_only, long long diag_off, unsigned* out_bits,
              cudaStream_t st);
int h2_diag_absmax(const float* A, long long ld, int n, unsigned* out_bits, cudaStream_t st);
// scale2 = {s, 1/s}, s the power of two that brings max (or sqrt(max)) to 2^H2_TARGET_EXP
int h2_scale_from_max(const unsigned* max_bits, int sqrt_of_max, float* scale2, cudaStream_t st);
// hi/lo fp16 split of A * s into the shadows (same indices, leading dimension ldh).  s = scale2[0], or derived from
// *max_bits (then *inv_out = 1/s is published).
int h2_split(const float* A, long long ld, long long rows, int cols, const float* scale2, const unsigned* max_bits,
             float* inv_out, int lower_only, long long diag_off, __half* hi, __half* lo, long long ldh, cudaStream_t st);

// the same split into TRANSPOSED shadows [cols x rows] (scale from *max_bits; small operands)
int h2_split_transpose(const float* A, long long ld, int rows, int cols, const unsigned* max_bits, float* inv_out, __half* hi,
                       __half* lo, long long ldh, cudaStream_t st);

}  // namespace hb
