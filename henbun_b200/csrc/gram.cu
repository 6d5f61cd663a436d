// UnitRBF / UnitCsymRBF Gram matrix and its lengthscale gradient.
//
//   forward : K_ij = exp(-0.5 * sum_d ((x_id - x2_jd)/ell_d)^2)  [+ jitter on the diagonal]
//             reference: UnitStationary.square_dist (Henbun/gp/kernels.py:54-84), UnitRBF.K (:110-111),
//             UnitCsymRBF.K (:122-126), the jitter of UnitStationary.Cholesky (:100-101).
//             The reference expands r^2 = -2 x.x' + |x|^2 + |x'|^2; we sum squared differences, which is
//             the same quantity with less cancellation (parity is against the fp64 oracle).
//   backward: g_ell_d = sum_ij G_ij K_ij (x_id - x2_jd)^2 / ell_d^3
//
// HBM-bound: the forward writes n*n2 floats once (lower-triangle tiles only on request), the
// backward reads G once and recomputes K from X (D is small), so K is never re-read.
#include "kernels.cuh"

namespace hb {

namespace {

constexpr int TILE = 64;
constexpr int DCH = 32;   // feature chunk staged in shared memory

__device__ __forceinline__ float ell_at(const float* ell, int n_ell, int d) { return n_ell == 1 ? ell[0] : ell[d]; }

template <bool CSYM>
__global__ void __launch_bounds__(256) rbf_gram_fwd_kernel(const float* __restrict__ X, const float* __restrict__ X2,
                                                           int n, int n2, int D, long long sX, long long sX2,
                                                           const float* __restrict__ ell, int n_ell, float* K,
                                                           long long ldk, long long sK, float jitter, int lower_only,
                                                           int self) {
  __shared__ float xs[TILE][DCH + 1];
  __shared__ float ys[TILE][DCH + 1];
  const int i0 = blockIdx.y * TILE, j0 = blockIdx.x * TILE;
  if (lower_only && j0 > i0 + TILE - 1) return;
  const int bz = blockIdx.z;
  const float* Xb = X + (long long)bz * sX;
  const float* Yb = X2 + (long long)bz * sX2;
  float* Kb = K + (long long)bz * sK;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float r2[4][4], r2b[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) { r2[a][b] = 0.f; r2b[a][b] = 0.f; }
  for (int d0 = 0; d0 < D; d0 += DCH) {
    const int dc = min(DCH, D - d0);
    __syncthreads();
    for (int e = threadIdx.x; e < TILE * dc; e += 256) {
      const int r = e / dc, d = e % dc;
      const float il = 1.f / ell_at(ell, n_ell, d0 + d);
      xs[r][d] = (i0 + r < n) ? Xb[(long long)(i0 + r) * D + d0 + d] * il : 0.f;
      ys[r][d] = (j0 + r < n2) ? Yb[(long long)(j0 + r) * D + d0 + d] * il : 0.f;
    }
    __syncthreads();
    for (int d = 0; d < dc; ++d) {
      float xv[4], yv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xv[a] = xs[ty + 16 * a][d];
#pragma unroll
      for (int b = 0; b < 4; ++b) yv[b] = ys[tx * 4 + b][d];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float df = xv[a] - yv[b];
          r2[a][b] = fmaf(df, df, r2[a][b]);
          if (CSYM) {
            const float sf = xv[a] + yv[b];
            r2b[a][b] = fmaf(sf, sf, r2b[a][b]);
          }
        }
    }
  }
  const bool vec = ((ldk & 3) == 0) && ((sK & 3) == 0) && ((reinterpret_cast<uintptr_t>(K) & 15) == 0);
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty + 16 * a;
    if (i >= n) continue;
    const int j = j0 + tx * 4;
    float o[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      float k = expf(-0.5f * r2[a][b]);
      if (CSYM) k += expf(-0.5f * r2b[a][b]);
      if (self && i == j + b) k += jitter;
      o[b] = k;
    }
    float* dst = Kb + (long long)i * ldk + j;
    if (vec && j + 3 < n2) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b)
        if (j + b < n2) dst[b] = o[b];
    }
  }
}

// Persistent over tiles; per-thread double accumulators per feature; one partial row per CTA.
// ISO: one shared lengthscale (n_ell == 1) -- the gradient needs only sum_ij g K r^2, one FMA per element instead of D
template <int DMAX, bool ISO = false>
__global__ void __launch_bounds__(256) rbf_gram_bwd_kernel(const float* __restrict__ G, long long ldg, long long sG,
                                                           const float* __restrict__ X, const float* __restrict__ X2,
                                                           int n, int n2, int D, long long sX, long long sX2,
                                                           const float* __restrict__ ell, int n_ell, int batch,
                                                           int sym_lower, int csym, double* partials) {
  __shared__ float xs[TILE][DMAX + 1];
  __shared__ float ys[TILE][DMAX + 1];
  __shared__ double red[32 * 8];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int tiles_i = (n + TILE - 1) / TILE, tiles_j = (n2 + TILE - 1) / TILE;
  const long long ntiles = (long long)batch * tiles_i * tiles_j;
  double dacc[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) dacc[d] = 0.0;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int bz = (int)(t / ((long long)tiles_i * tiles_j));
    const int rem = (int)(t % ((long long)tiles_i * tiles_j));
    const int i0 = (rem / tiles_j) * TILE, j0 = (rem % tiles_j) * TILE;
    if (sym_lower && j0 > i0 + TILE - 1) continue;
    const float* Xb = X + (long long)bz * sX;
    const float* Yb = X2 + (long long)bz * sX2;
    const float* Gb = G + (long long)bz * sG;
    // this thread's four 16-byte pieces of G go in flight BEFORE the tile's X rows are staged: with 119 registers two CTAs
    // share an SM, and a load issued right before its use left the 16 resident warps waiting on HBM most of the time
    const bool vec = (DMAX <= 8) && (j0 + tx * 4 + 3 < n2 && (ldg & 3) == 0 && (reinterpret_cast<uintptr_t>(Gb) & 15) == 0);
    float4 gpre[4];
    if (DMAX <= 8) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty + 16 * a;
        const bool need = vec && i < n && !(sym_lower && j0 + tx * 4 > i);
        gpre[a] = need ? __ldg(reinterpret_cast<const float4*>(Gb + (long long)i * ldg + j0 + tx * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < TILE * D; e += 256) {
      const int r = e / D, d = e % D;
      const float il = 1.f / ell_at(ell, n_ell, d);
      xs[r][d] = (i0 + r < n) ? Xb[(long long)(i0 + r) * D + d] * il : 0.f;
      ys[r][d] = (j0 + r < n2) ? Yb[(long long)(j0 + r) * D + d] * il : 0.f;
    }
    __syncthreads();
    float facc[DMAX];
#pragma unroll
    for (int d = 0; d < DMAX; ++d) facc[d] = 0.f;
    if (DMAX <= 8) {
      // small feature count: the thread's 4 x-rows and 4 y-rows live in registers (the shared-memory version
      // issued 4 LDS per element and feature and was LDS-bound: 18 ms at n = 65536, D = 8)
      float xr[4][DMAX], yr[4][DMAX];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int d = 0; d < DMAX; ++d) { xr[a][d] = (d < D) ? xs[ty + 16 * a][d] : 0.f; yr[a][d] = (d < D) ? ys[tx * 4 + a][d] : 0.f; }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = i0 + ty + 16 * a;
        if (i >= n) continue;
        const float4 g4 = gpre[a];
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int j = j0 + tx * 4 + b;
          if (j >= n2) continue;
          float w = 1.f;
          if (sym_lower) {
            if (j > i) continue;
            w = (j == i) ? 1.f : 2.f;
          }
          float df[DMAX], r2 = 0.f, r2b = 0.f;
#pragma unroll
          for (int d = 0; d < DMAX; ++d) { df[d] = xr[a][d] - yr[b][d]; r2 = fmaf(df[d], df[d], r2); }
          const float g = w * (vec ? gv[b] : __ldg(Gb + (long long)i * ldg + j));
          const float gk = g * expf(-0.5f * r2);
          if (ISO) facc[0] = fmaf(gk, r2, facc[0]);
          else {
#pragma unroll
            for (int d = 0; d < DMAX; ++d) facc[d] = fmaf(gk * df[d], df[d], facc[d]);
          }
          if (csym) {
            float sf[DMAX];
#pragma unroll
            for (int d = 0; d < DMAX; ++d) { sf[d] = xr[a][d] + yr[b][d]; r2b = fmaf(sf[d], sf[d], r2b); }
            const float gkb = g * expf(-0.5f * r2b);
            if (ISO) facc[0] = fmaf(gkb, r2b, facc[0]);
            else {
#pragma unroll
              for (int d = 0; d < DMAX; ++d) facc[d] = fmaf(gkb * sf[d], sf[d], facc[d]);
            }
          }
        }
      }
    } else {
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int i = i0 + ty + 16 * a;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = j0 + tx * 4 + b;
        if (i >= n || j >= n2) continue;
        float w = 1.f;
        if (sym_lower) {
          if (j > i) continue;
          w = (j == i) ? 1.f : 2.f;   // (the RBF diagonal has zero lengthscale derivative; Csym's does not)
        }
        float r2 = 0.f, r2b = 0.f;
#pragma unroll
        for (int d = 0; d < DMAX; ++d) {
          if (d < D) {
            const float df = xs[ty + 16 * a][d] - ys[tx * 4 + b][d];
            r2 = fmaf(df, df, r2);
            const float sf = xs[ty + 16 * a][d] + ys[tx * 4 + b][d];
            r2b = fmaf(sf, sf, r2b);
          }
        }
        const float g = w * __ldg(Gb + (long long)i * ldg + j);
        const float gk = g * expf(-0.5f * r2);
        const float gkb = csym ? g * expf(-0.5f * r2b) : 0.f;
#pragma unroll
        for (int d = 0; d < DMAX; ++d) {
          if (d < D) {
            const float df = xs[ty + 16 * a][d] - ys[tx * 4 + b][d];
            facc[d] = fmaf(gk * df, df, facc[d]);
            if (csym) {
              const float sf = xs[ty + 16 * a][d] + ys[tx * 4 + b][d];
              facc[d] = fmaf(gkb * sf, sf, facc[d]);
            }
          }
        }
      }
    }
    }
    if (ISO) dacc[0] += (double)facc[0];
    else {
#pragma unroll
      for (int d = 0; d < DMAX; ++d) dacc[d] += (double)facc[d];
    }
  }
  // block reduce, 8 features at a time
  for (int d0 = 0; d0 < DMAX; d0 += 8) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (d0 + k < DMAX) ? dacc[d0 + k] : 0.0;
    block_sum<8>(v, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (d0 + k < D) partials[(long long)blockIdx.x * DMAX + d0 + k] = v[k];
    }
  }
}

__global__ void gram_bwd_finalize_kernel(const double* __restrict__ partials, int nblocks, int D, int dmax,
                                         const float* __restrict__ ell, int n_ell,
                                         const float* __restrict__ out_scale, float* g_ell) {
  // one warp; hat-space sums -> divide by ell_d (x was pre-divided by ell, so diff^2/ell^2 is already in)
  double tot = 0.0;
  const double osc = out_scale ? (double)*out_scale : 1.0;
  for (int d = 0; d < D; ++d) {
    double s = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 32) s += partials[(long long)b * dmax + d];
    s = warp_sum(s);
    const double l = (double)ell_at(ell, n_ell, d);
    if (n_ell == 1) tot += s / l;
    else if (threadIdx.x == 0) g_ell[d] = (float)(osc * s / l);
  }
  if (n_ell == 1 && threadIdx.x == 0) g_ell[0] = (float)(osc * tot);
}


// Gradient w.r.t. the kernel inputs ("next" row SparseGP: inducing points are trainable, Henbun/gp/gp.py:95-98):
//   dX2[j,d] = sum_i Geff[i,j] K(x_i, y_j) (x_id - y_jd) / ell_d^2
// with Geff = G, or (sym_lower) the symmetric matrix whose lower triangle G holds.  One block owns 32 points y_j and
// streams G in 32-row tiles (coalesced along j); K is recomputed from X, X2.  D <= DMAX features.
template <int DMAX>
__global__ void __launch_bounds__(256) rbf_gram_bwd_x2_kernel(const float* __restrict__ G, long long ldg, long long sG,
                                                              const float* __restrict__ X, const float* __restrict__ X2,
                                                              int n, int n2, int D, long long sX, long long sX2,
                                                              const float* __restrict__ ell, int n_ell, int sym_lower,
                                                              float scale, float* dX2, long long sD) {
  __shared__ float xs[32][DMAX + 1];
  __shared__ float red[8][32][DMAX + 1];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int bz = blockIdx.y;
  const float* Gb = G + (long long)bz * sG;
  const float* Xb = X + (long long)bz * sX;
  const float* Yb = X2 + (long long)bz * sX2;
  const int j = blockIdx.x * 32 + tx;
  float y[DMAX], acc[DMAX];
#pragma unroll
  for (int d = 0; d < DMAX; ++d) {
    y[d] = (d < D && j < n2) ? Yb[(long long)j * D + d] / ell_at(ell, n_ell, d) : 0.f;
    acc[d] = 0.f;
  }
  for (int i0 = 0; i0 < n; i0 += 32) {
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * D; e += 256) {
      const int r = e / D, d = e % D;
      xs[r][d] = (i0 + r < n) ? Xb[(long long)(i0 + r) * D + d] / ell_at(ell, n_ell, d) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int r = ty + 8 * k, i = i0 + r;
      if (i >= n || j >= n2) continue;
      const float g = (sym_lower && j > i) ? __ldg(Gb + (long long)j * ldg + i) : __ldg(Gb + (long long)i * ldg + j);
      float df[DMAX], r2 = 0.f;
#pragma unroll
      for (int d = 0; d < DMAX; ++d) { df[d] = (d < D) ? xs[r][d] - y[d] : 0.f; r2 = fmaf(df[d], df[d], r2); }
      const float gk = g * expf(-0.5f * r2);
#pragma unroll
      for (int d = 0; d < DMAX; ++d) acc[d] = fmaf(gk, df[d], acc[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < DMAX; ++d) red[ty][tx][d] = acc[d];
  __syncthreads();
  for (int e = threadIdx.x; e < 32 * D; e += 256) {
    const int c = e / D, d = e % D;
    const int jj = blockIdx.x * 32 + c;
    if (jj >= n2) continue;
    float sacc = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) sacc += red[t][c][d];
    dX2[(long long)bz * sD + (long long)jj * D + d] = scale * sacc / ell_at(ell, n_ell, d);
  }
}

}  // namespace

int rbf_gram_fwd(const float* X, const float* X2, int n, int n2, int D, long long sX, long long sX2,
                 const float* ell, int n_ell, float* K, long long ldk, long long sK, int batch, float jitter,
                 int lower_only, int csym, cudaStream_t st) {
  if (n < 0 || n2 < 0 || D <= 0 || batch < 0) return HB_ERR_ARG;
  if (n == 0 || n2 == 0 || batch == 0) return HB_OK;
  if (!X || !ell || !K || ldk < n2 || (n_ell != 1 && n_ell != D)) return HB_ERR_ARG;
  const int self = (X2 == nullptr);
  if (self && n != n2) return HB_ERR_ARG;
  if (!self && lower_only) return HB_ERR_ARG;
  const float* Y = self ? X : X2;
  const long long sY = self ? sX : sX2;
  dim3 grid(cdiv(n2, TILE), cdiv(n, TILE), batch);
  if (grid.y > 65535 || grid.z > 65535) return HB_ERR_ARG;
  if (csym)
    rbf_gram_fwd_kernel<true><<<grid, 256, 0, st>>>(X, Y, n, n2, D, sX, sY, ell, n_ell, K, ldk, sK, jitter, lower_only, self);
  else
    rbf_gram_fwd_kernel<false><<<grid, 256, 0, st>>>(X, Y, n, n2, D, sX, sY, ell, n_ell, K, ldk, sK, jitter, lower_only, self);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int rbf_gram_bwd(const float* G, long long ldg, long long sG, const float* X, const float* X2, int n, int n2,
                 int D, long long sX, long long sX2, const float* ell, int n_ell, int batch, int sym_lower,
                 int csym, const float* out_scale, float* g_ell, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n < 0 || n2 < 0 || D <= 0 || D > 32 || batch < 0) return HB_ERR_ARG;
  if (!G || !X || !ell || !g_ell || ldg < n2 || (n_ell != 1 && n_ell != D)) return HB_ERR_ARG;
  if (!ws || ws_bytes < kReduceWsBytes) return HB_ERR_WORKSPACE;
  const int self = (X2 == nullptr);
  if (self && n != n2) return HB_ERR_ARG;
  if (sym_lower && !self) return HB_ERR_ARG;
  const float* Y = self ? X : X2;
  const long long sY = self ? sX : sX2;
  const long long ntiles = (long long)batch * cdiv(n, TILE) * cdiv(n2, TILE);
  int nb = (int)(ntiles < kReduceBlocks ? (ntiles > 0 ? ntiles : 1) : kReduceBlocks);
  // grid: a multiple of the SM count when the problem is large
  if (nb == kReduceBlocks) nb = 148 * 6;
  double* partials = reinterpret_cast<double*>(ws);
  const int dmax = D <= 8 ? 8 : 32;
  if (dmax == 8 && n_ell == 1)
    rbf_gram_bwd_kernel<8, true><<<nb, 256, 0, st>>>(G, ldg, sG, X, Y, n, n2, D, sX, sY, ell, n_ell, batch, sym_lower, csym, partials);
  else if (dmax == 8)
    rbf_gram_bwd_kernel<8><<<nb, 256, 0, st>>>(G, ldg, sG, X, Y, n, n2, D, sX, sY, ell, n_ell, batch, sym_lower, csym, partials);
  else
    rbf_gram_bwd_kernel<32><<<nb, 256, 0, st>>>(G, ldg, sG, X, Y, n, n2, D, sX, sY, ell, n_ell, batch, sym_lower, csym, partials);
  HB_CHECK_LAUNCH();
  gram_bwd_finalize_kernel<<<1, 32, 0, st>>>(partials, nb, D, dmax, ell, n_ell, out_scale, g_ell);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

// dX2 = scale * d/dX2 of sum_ij Geff_ij K(X, X2)_ij  (see rbf_gram_bwd_x2_kernel); D <= 32.
int rbf_gram_bwd_x2(const float* G, long long ldg, long long sG, const float* X, const float* X2, int n, int n2, int D,
                    long long sX, long long sX2, const float* ell, int n_ell, int batch, int sym_lower, float scale,
                    float* dX2, cudaStream_t st) {
  if (n < 0 || n2 < 0 || D <= 0 || D > 32 || batch < 0 || (n_ell != 1 && n_ell != D)) return HB_ERR_ARG;
  if (n2 == 0 || batch == 0) return HB_OK;
  if (!G || !X || !X2 || !ell || !dX2 || ldg < n2 || batch > 65535) return HB_ERR_ARG;
  if (sym_lower && n != n2) return HB_ERR_ARG;
  dim3 grid((unsigned)cdiv(n2, 32), (unsigned)batch);
  const long long sD = (long long)n2 * D;
  if (D <= 8)
    rbf_gram_bwd_x2_kernel<8><<<grid, 256, 0, st>>>(G, ldg, sG, X, X2, n, n2, D, sX, sX2, ell, n_ell, sym_lower, scale, dX2, sD);
  else
    rbf_gram_bwd_x2_kernel<32><<<grid, 256, 0, st>>>(G, ldg, sG, X, X2, n, n2, D, sX, sX2, ell, n_ell, sym_lower, scale, dX2, sD);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

}  // namespace hb
