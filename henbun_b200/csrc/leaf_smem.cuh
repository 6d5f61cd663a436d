// Shared-memory building blocks of the 128 x 128 leaf problems: blocked Cholesky, blocked triangular inverse, masked
// 128^3 products, coalesced loads.  Used by the leaf kernels of the blocked factorisations (linalg.cu) and by the one-CTA
// whole-step kernel of notebook-sized models (gp_small.cu).  All functions expect blockDim.x == LEAF_THREADS and matrices
// of NB x NB floats with leading dimension LDS, identity-padded beyond the live order n.
#pragma once
#include "common.cuh"

namespace hb {
namespace leaf {

constexpr int NB = 128;          // leaf size
constexpr int LDS = NB + 1;      // padded shared-memory leading dimension
constexpr int LEAF_THREADS = 512;

// ------------------------------------------------------------------------------------------
// in-CTA helpers (matrices are NB x NB in shared memory, ld = LDS)
// ------------------------------------------------------------------------------------------

// Column steps of the in-register 16x16 Cholesky (lane l holds row l); compile-time recursion keeps every array
// index static so the rows stay in registers.
template <int J>
__device__ __forceinline__ void diag_steps(float (&a)[16], int l, int& bad) {
  if constexpr (J < 16) {
    const float d = __shfl_sync(0xffffffffu, a[J], J);
    if (!(d > 0.f) && bad == 0) bad = J + 1;
    const float sd = sqrtf(d), rs = 1.f / sd;
    if (l == J) a[J] = sd;
    else if (l > J) a[J] *= rs;
#pragma unroll
    for (int k = J + 1; k < 16; ++k) {
      const float akj = __shfl_sync(0xffffffffu, a[J], k);
      if (l >= k) a[k] = fmaf(-a[J], akj, a[k]);
    }
    diag_steps<J + 1>(a, l, bad);
  }
}

// Blocked right-looking Cholesky of the NB x NB lower triangle held in S (rows/cols >= n are identity, so the
// full padded matrix is factored).  16-wide panels: (1) warp 0 factors the 16x16 diagonal block in registers with
// shuffles, (2) one thread per row solves its 16 panel entries, (3) all threads apply the rank-16 update out of a
// transposed copy of the panel (conflict-free).  3 barriers per panel instead of one per column.
constexpr int PB = 16;
__device__ inline void potrf_smem(float* S, int n, int* err_flag, int err_base) {
  __shared__ float rinv[PB];
  __shared__ float Pt[PB][NB + 1];      // Pt[j][i] = L[i][j0 + j] for the rows below the diagonal block
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  for (int j0 = 0; j0 < NB; j0 += PB) {
    if (warp == 0) {
      float a[PB];
      const int row = j0 + (lane & (PB - 1));
#pragma unroll
      for (int k = 0; k < PB; ++k) a[k] = S[row * LDS + j0 + k];
      int bad = 0;
      diag_steps<0>(a, lane & (PB - 1), bad);
      if (bad) bad += j0;
      if (lane < PB) {
#pragma unroll
        for (int k = 0; k < PB; ++k) if (k <= lane) S[row * LDS + j0 + k] = a[k];
        rinv[lane] = 1.f / a[lane];
      }
      if (lane == 0 && bad != 0 && bad <= n && err_flag) atomicCAS(err_flag, 0, err_base + bad);
    }
    __syncthreads();
    const int base = j0 + PB;
    if (base >= NB) break;
    // (2) panel rows: x L11^T = a  (forward substitution, L11 broadcast from shared memory)
    for (int i = base + tid; i < NB; i += blockDim.x) {
      float x[PB];
#pragma unroll
      for (int k = 0; k < PB; ++k) x[k] = S[i * LDS + j0 + k];
#pragma unroll
      for (int j = 0; j < PB; ++j) {
        float s = x[j];
#pragma unroll
        for (int k = 0; k < j; ++k) s = fmaf(-x[k], S[(j0 + j) * LDS + j0 + k], s);
        x[j] = s * rinv[j];
      }
#pragma unroll
      for (int k = 0; k < PB; ++k) { S[i * LDS + j0 + k] = x[k]; Pt[k][i] = x[k]; }
    }
    __syncthreads();
    // (3) trailing update S[i][k] -= sum_j Pt[j][i] Pt[j][k]  (k <= i), rows ty + 16 r, columns tx + 32 c
    {
      const int ty = tid >> 5, tx = tid & 31;          // 16 x 32 thread grid (512 threads)
      constexpr int R = (NB - PB) / 16, Cc = NB / 32;
      float acc[R][Cc];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < Cc; ++c) acc[r][c] = 0.f;
#pragma unroll 4
      for (int j = 0; j < PB; ++j) {
        float av[R], bv[Cc];
#pragma unroll
        for (int r = 0; r < R; ++r) { const int i = base + ty + 16 * r; av[r] = (i < NB) ? Pt[j][i] : 0.f; }
#pragma unroll
        for (int c = 0; c < Cc; ++c) { const int k = base + tx + 32 * c; bv[c] = (k < NB) ? Pt[j][k] : 0.f; }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < Cc; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const int i = base + ty + 16 * r;
#pragma unroll
        for (int c = 0; c < Cc; ++c) {
          const int k = base + tx + 32 * c;
          if (i < NB && k <= i) S[i * LDS + k] -= acc[r][c];
        }
      }
    }
    __syncthreads();
  }
  __syncthreads();
}

// One doubling level of the blocked triangular inverse: for every aligned 2S x 2S diagonal block whose two S x S
// diagonal blocks are already inverted in X, X21 = -X22 * (L21 * X11).  T is scratch (same geometry as X).
template <int S>
__device__ __forceinline__ void trinv_level(const float* L, float* X, float* T) {
  constexpr int NO = S / 8;                       // outputs per thread: rows rg + 8 r, one column j
  const int tid = threadIdx.x;
  const int pair = tid / (8 * S), j = tid % S, rg = (tid / S) % 8;
  const int r0 = 2 * S * pair;
  float acc[NO];
#pragma unroll
  for (int r = 0; r < NO; ++r) acc[r] = 0.f;
  for (int k = j; k < S; ++k) {                   // X11 is lower triangular: X11[k][j] = 0 for k < j
    const float b = X[(r0 + k) * LDS + r0 + j];
#pragma unroll
    for (int r = 0; r < NO; ++r) acc[r] = fmaf(L[(r0 + S + rg + 8 * r) * LDS + r0 + k], b, acc[r]);
  }
#pragma unroll
  for (int r = 0; r < NO; ++r) T[(r0 + S + rg + 8 * r) * LDS + r0 + j] = acc[r];
  __syncthreads();
#pragma unroll
  for (int r = 0; r < NO; ++r) acc[r] = 0.f;
  for (int k = 0; k < S; ++k) {                   // X22[i][k] = 0 for k > i
    const float b = T[(r0 + S + k) * LDS + r0 + j];
#pragma unroll
    for (int r = 0; r < NO; ++r) {
      const int i = rg + 8 * r;
      const float a = (k <= i) ? X[(r0 + S + i) * LDS + r0 + S + k] : 0.f;
      acc[r] = fmaf(a, b, acc[r]);
    }
  }
#pragma unroll
  for (int r = 0; r < NO; ++r) X[(r0 + S + rg + 8 * r) * LDS + r0 + j] = -acc[r];
  __syncthreads();
}

// X = L^{-1} for the lower-triangular L (NB x NB, identity padded), blocked: 16x16 diagonal blocks by forward
// substitution in registers (one lane per column), then three doubling levels (16 -> 32 -> 64 -> 128).
// blockDim.x must be 512.  T is an NB x LDS scratch buffer.
__device__ inline void trinv_smem(const float* L, float* X, float* T) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < NB * LDS; e += blockDim.x) X[e] = 0.f;
  __syncthreads();
  if (warp < NB / 16 && lane < 16) {
    const int b0 = 16 * warp, c = lane;
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float sacc = (i == c) ? 1.f : 0.f;
#pragma unroll
      for (int k = 0; k < i; ++k) sacc = fmaf(-L[(b0 + i) * LDS + b0 + k], x[k], sacc);
      x[i] = sacc / L[(b0 + i) * LDS + b0 + i];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) X[(b0 + i) * LDS + b0 + c] = x[i];
  }
  __syncthreads();
  trinv_level<16>(L, X, T);
  trinv_level<32>(L, X, T);
  trinv_level<64>(L, X, T);
}

// C = op(A) * op(B) on NB x NB shared-memory matrices; k restricted to [klo(i), khi(i)] by row.
// KMODE 0: all k; 1: k >= i; 2: k <= i.
template <bool TA, bool TB, int KMODE>
__device__ inline void mm_smem(float* C, const float* A, const float* B) {
  const int tid = threadIdx.x;
  const int tr = tid >> 4, tc = tid & 15;   // 32 x 16 thread grid, 4 rows x 8 interleaved cols each
  float acc[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
  const int i0 = tr * 4;
  const int klo = (KMODE == 1) ? i0 : 0;
  const int khi = (KMODE == 2) ? i0 + 3 : NB - 1;
  for (int k = klo; k <= khi; ++k) {
    float a[4], b[8];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float v = TA ? A[k * LDS + i0 + r] : A[(i0 + r) * LDS + k];
      if (KMODE == 1 && k < i0 + r) v = 0.f;
      if (KMODE == 2 && k > i0 + r) v = 0.f;
      a[r] = v;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) b[c] = TB ? B[(tc + 16 * c) * LDS + k] : B[k * LDS + tc + 16 * c];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
  }
  __syncthreads();   // C may alias an input
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 8; ++c) C[(i0 + r) * LDS + tc + 16 * c] = acc[r][c];
}

// load the n x n lower triangle of a global block into S; pad with `diag_pad` on the padded diagonal.
// All global loads of a thread are issued before the first shared-memory store (one memory round trip).
__device__ inline void load_lower(float* S, const float* G, long long ld, int n, float diag_pad) {
  const int tid = threadIdx.x;
  const bool vec = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0) && blockDim.x == 512;
  if (vec) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int q = tid + 512 * u, i = q >> 5, j4 = (q & 31) * 4;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n && j4 <= i) {
        if (j4 + 3 < n) v[u] = *reinterpret_cast<const float4*>(G + (long long)i * ld + j4);
        else {
          const float* g = G + (long long)i * ld + j4;
          v[u].x = g[0];
          if (j4 + 1 < n) v[u].y = g[1];
          if (j4 + 2 < n) v[u].z = g[2];
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int q = tid + 512 * u, i = q >> 5, j4 = (q & 31) * 4;
      const float t[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int jj = j4 + c;
        float x = (i < n && jj <= i) ? t[c] : 0.f;
        if (i >= n && jj == i) x = diag_pad;
        S[i * LDS + jj] = x;
      }
    }
    return;
  }
  for (int e = tid; e < NB * NB; e += blockDim.x) {
    const int i = e / NB, j = e % NB;
    float v = 0.f;
    if (i < n && j <= i) v = G[(long long)i * ld + j];
    else if (i == j) v = diag_pad;
    S[i * LDS + j] = v;
  }
}

// load a dense NB x NB block (ld = NB, 16-byte aligned) with one memory round trip
__device__ inline void load_full(float* S, const float* G) {
  const int tid = threadIdx.x;
  if (blockDim.x == 512 && (reinterpret_cast<uintptr_t>(G) & 15) == 0) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = *reinterpret_cast<const float4*>(G + 4 * (tid + 512 * u));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int q = tid + 512 * u, i = q >> 5, j4 = (q & 31) * 4;
      S[i * LDS + j4] = v[u].x; S[i * LDS + j4 + 1] = v[u].y; S[i * LDS + j4 + 2] = v[u].z; S[i * LDS + j4 + 3] = v[u].w;
    }
    return;
  }
  for (int e = tid; e < NB * NB; e += blockDim.x) S[(e / NB) * LDS + (e % NB)] = G[e];
}

}  // namespace leaf
}  // namespace hb
