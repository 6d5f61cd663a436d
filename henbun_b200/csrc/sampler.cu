// Reparameterised sampler of Variational + one-sample Monte-Carlo KL, fused.
//
//   forward : z = mu + exp(omega) * eps                         (Henbun/variationals.py:138-142)
//             KL = -0.5 * sum(2*omega + eps^2 - z^2)            (variationals.py:183-184, 225-230)
//   backward: given zbar = d obj / d z (likelihood path) and obj containing -kl_coef*KL,
//             gmu = sum_s (zbar - c z),  gomega = sum_s (zbar - c z) exp(omega) eps + c*S
//
// eps is either read from a caller buffer [S, rows, cols] (parity mode, the reference's
// feed_dict={variational.u: eps}) or regenerated from Philox(seed, offset) so it never touches HBM.
// mu/omega are addressed as [rows, cols] with a row stride: for LOCAL variationals they are the
// two halves of each encoder output row (Henbun/param.py:529-537), no slice copy is made.
#include "kernels.cuh"

namespace hb {

namespace {

__global__ void finalize_scaled_kernel(const double* __restrict__ partials, int nblocks, int nv, float* out, double scale) {
  // one warp per value
  const int v = blockIdx.x;
  double s = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) s += partials[(long long)b * nv + v];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[v] = (float)(scale * s);
}

// Each thread handles groups of 4 consecutive flat elements of the [S, P] sample tensor.
__global__ void __launch_bounds__(256) sample_diag_fwd_kernel(const float* __restrict__ mu, long long ld_mu,
                                                              const float* __restrict__ om, long long ld_om, int cols,
                                                              long long P, const float* __restrict__ eps,
                                                              unsigned long long seed, unsigned long long offset,
                                                              long long total, float* __restrict__ z, double* partials) {
  __shared__ double red[32];
  double acc[1] = {0.0};
  const long long ngroups = (total + 3) / 4;
  const bool vec = ((P & 3) == 0) && (ld_mu == cols) && (ld_om == cols) &&
                   ((reinterpret_cast<uintptr_t>(mu) & 15) == 0) && ((reinterpret_cast<uintptr_t>(om) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(z) & 15) == 0) && (!eps || (reinterpret_cast<uintptr_t>(eps) & 15) == 0);
  for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < ngroups; g += (long long)gridDim.x * blockDim.x) {
    const long long e0 = 4 * g;
    float e[4];
    if (eps) {
      if (vec) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(eps + e0));
        e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) e[c] = (e0 + c < total) ? eps[e0 + c] : 0.f;
      }
    } else {
      philox_normal4(seed, (offset >> 2) + (unsigned long long)g, e);
    }
    float zz[4];
    if (vec) {
      const long long p0 = e0 % P;
      const float4 m4 = __ldg(reinterpret_cast<const float4*>(mu + p0));
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(om + p0));
      const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, oo[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        zz[c] = fmaf(expf(oo[c]), e[c], mm[c]);
        acc[0] += (double)(2.f * oo[c] + e[c] * e[c] - zz[c] * zz[c]);
      }
      *reinterpret_cast<float4*>(z + e0) = make_float4(zz[0], zz[1], zz[2], zz[3]);
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (e0 + c >= total) continue;
        const long long p = (e0 + c) % P;
        const long long r = p / cols; const int cc = (int)(p % cols);
        const float m = mu[r * ld_mu + cc], o = om[r * ld_om + cc];
        zz[c] = fmaf(expf(o), e[c], m);
        acc[0] += (double)(2.f * o + e[c] * e[c] - zz[c] * zz[c]);
        z[e0 + c] = zz[c];
      }
    }
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}

// One thread per 4 consecutive parameter positions p; loops over the S samples (coalesced over p).
__global__ void __launch_bounds__(128) sample_diag_bwd_kernel(const float* __restrict__ mu, long long ld_mu,
                                                              const float* __restrict__ om, long long ld_om, int cols,
                                                              long long P, const float* __restrict__ eps,
                                                              unsigned long long seed, unsigned long long offset, int S,
                                                              const float* __restrict__ zbar,
                                                              const float* __restrict__ zbar_scale, float c_host,
                                                              const float* __restrict__ c_dev, float* gmu,
                                                              long long ld_gmu, float* gom, long long ld_gom, float beta) {
  const long long p0 = 4 * (blockIdx.x * (long long)blockDim.x + threadIdx.x);
  if (p0 >= P) return;
  float m[4], so[4], am[4] = {0, 0, 0, 0}, ao[4] = {0, 0, 0, 0};
  long long im[4], io[4];
  int nv = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long p = p0 + k;
    if (p < P) {
      nv = k + 1;
      const long long r = p / cols; const int cc = (int)(p % cols);
      m[k] = mu[r * ld_mu + cc]; so[k] = expf(om[r * ld_om + cc]);
      im[k] = r * ld_gmu + cc; io[k] = r * ld_gom + cc;
    } else { m[k] = 0.f; so[k] = 0.f; im[k] = 0; io[k] = 0; }
  }
  const bool grp = ((P & 3) == 0);   // Philox groups line up with p0
  const float zsc = zbar_scale ? __ldg(zbar_scale) : 1.f;
  const float c = c_dev ? c_host * __ldg(c_dev) : c_host;
  for (int s = 0; s < S; ++s) {
    const long long base = (long long)s * P + p0;
    float e[4];
    if (eps) {
#pragma unroll
      for (int k = 0; k < 4; ++k) e[k] = (k < nv) ? __ldg(eps + base + k) : 0.f;
    } else if (grp) {
      philox_normal4(seed, (offset + (unsigned long long)base) >> 2, e);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) e[k] = (k < nv) ? philox_normal_at(seed, offset, (unsigned long long)(base + k)) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k >= nv) continue;
      const float zz = fmaf(so[k], e[k], m[k]);
      const float zb = zbar ? zsc * __ldg(zbar + base + k) : 0.f;
      const float zt = zb - c * zz;
      am[k] += zt;
      ao[k] = fmaf(zt * so[k], e[k], ao[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= nv) continue;
    const float g1 = am[k], g2 = ao[k] + c * (float)S;
    gmu[im[k]] = (beta != 0.f) ? fmaf(beta, gmu[im[k]], g1) : g1;
    gom[io[k]] = (beta != 0.f) ? fmaf(beta, gom[io[k]], g2) : g2;
  }
}

// KL of a full-rank Normal: -0.5*( S_per_batch * sum_i log(Lq_ii^2) + sum(eps^2 - z^2) )
//   (variationals.py:185-186, 225-230).  Lq batched [batch, n, n]; eps, z flat with `count` elements.
__global__ void __launch_bounds__(256) tril_kl_kernel(const float* __restrict__ Lq, long long ld, long long stride, int n,
                                                      int batch, const float* __restrict__ eps,
                                                      const float* __restrict__ z, long long count, int S,
                                                      double* partials) {
  __shared__ double red[32];
  double acc[1] = {0.0};
  const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  for (long long e = gid; e < count; e += gsz) {
    const float a = eps[e], b = z[e];
    acc[0] += (double)(a * a - b * b);
  }
  for (long long e = gid; e < (long long)batch * n; e += gsz) {
    const long long b = e / n; const int i = (int)(e % n);
    const float d = Lq[b * stride + (long long)i * ld + i];
    acc[0] += (double)S * (double)logf(d * d);
  }
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}

__global__ void tril_diag_grad_kernel(float* gL, long long ld, long long stride, const float* __restrict__ Lq,
                                      long long ldl, long long strideL, int n, int batch, float coef) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)batch * n) return;
  const long long b = e / n; const int i = (int)(e % n);
  gL[b * stride + (long long)i * ld + i] += coef / Lq[b * strideL + (long long)i * ldl + i];
}

}  // namespace

int sample_diag_fwd(const float* mu, long long ld_mu, const float* omega, long long ld_om, int rows, int cols,
                    const float* eps, unsigned long long seed, unsigned long long offset, int S, float* z,
                    float* kl_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (rows < 0 || cols < 0 || S < 0) return HB_ERR_ARG;
  const long long P = (long long)rows * cols, total = P * S;
  if (total == 0) {
    if (kl_out) return fill_f32(kl_out, 1, 0.f, st);
    return HB_OK;
  }
  if (!mu || !omega || !z || ld_mu < cols || ld_om < cols) return HB_ERR_ARG;
  if (!eps && (offset & 3ull)) return HB_ERR_ARG;
  if (!ws || ws_bytes < kReduceWsBytes) return HB_ERR_WORKSPACE;
  long long nb = ((total + 3) / 4 + 255) / 256;
  if (nb > kReduceBlocks) nb = kReduceBlocks;
  double* partials = reinterpret_cast<double*>(ws);
  sample_diag_fwd_kernel<<<(int)nb, 256, 0, st>>>(mu, ld_mu, omega, ld_om, cols, P, eps, seed, offset, total, z, partials);
  HB_CHECK_LAUNCH();
  if (kl_out) {
    finalize_scaled_kernel<<<1, 32, 0, st>>>(partials, (int)nb, 1, kl_out, -0.5);
    HB_CHECK_LAUNCH();
  }
  return HB_OK;
}

int sample_diag_bwd(const float* mu, long long ld_mu, const float* omega, long long ld_om, int rows, int cols,
                    const float* eps, unsigned long long seed, unsigned long long offset, int S, const float* zbar,
                    const float* zbar_scale, float kl_coef, const float* kl_coef_dev, float* gmu, long long ld_gmu, float* gom,
                    long long ld_gom,
                    float beta, cudaStream_t st) {
  if (rows < 0 || cols < 0 || S < 0) return HB_ERR_ARG;
  const long long P = (long long)rows * cols;
  if (P == 0) return HB_OK;
  if (!mu || !omega || !gmu || !gom || ld_mu < cols || ld_om < cols || ld_gmu < cols || ld_gom < cols) return HB_ERR_ARG;
  if (!eps && (offset & 3ull)) return HB_ERR_ARG;
  const long long nthreads = (P + 3) / 4;
  sample_diag_bwd_kernel<<<(int)((nthreads + 127) / 128), 128, 0, st>>>(mu, ld_mu, omega, ld_om, cols, P, eps, seed,
                                                                       offset, S, zbar, zbar_scale, kl_coef, kl_coef_dev, gmu, ld_gmu, gom,
                                                                       ld_gom, beta);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int tril_logdet_kl(const float* Lq, long long ld, long long stride, int n, int batch, const float* eps,
                   const float* z, long long count, int S, float* kl_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n < 0 || batch < 0 || count < 0 || !kl_out) return HB_ERR_ARG;
  if (!ws || ws_bytes < kReduceWsBytes) return HB_ERR_WORKSPACE;
  long long work = count > (long long)batch * n ? count : (long long)batch * n;
  long long nb = (work + 255) / 256;
  if (nb < 1) nb = 1;
  if (nb > kReduceBlocks) nb = kReduceBlocks;
  double* partials = reinterpret_cast<double*>(ws);
  tril_kl_kernel<<<(int)nb, 256, 0, st>>>(Lq, ld, stride, n, batch, eps, z, count, S, partials);
  HB_CHECK_LAUNCH();
  finalize_scaled_kernel<<<1, 32, 0, st>>>(partials, (int)nb, 1, kl_out, -0.5);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int tril_diag_grad(float* gL, long long ld, long long stride, const float* Lq, long long ldl, long long strideL,
                   int n, int batch, float coef, cudaStream_t st) {
  const long long tot = (long long)batch * n;
  if (tot <= 0) return HB_OK;
  tril_diag_grad_kernel<<<(int)((tot + 255) / 256), 256, 0, st>>>(gL, ld, stride, Lq, ldl, strideL, n, batch, coef);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

}  // namespace hb
