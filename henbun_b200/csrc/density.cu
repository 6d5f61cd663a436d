// densities.gaussian (Henbun/densities.py:25-27) and its reduction.
//
// gauss_loglik_fwd fuses, in one pass over F: the summed log-likelihood
//   ll = sum(-0.5 log 2pi - 0.5 log var - 0.5 (F - Y)^2 / var)     (the caller's tf.reduce_sum,
//                                                                    notebooks/GaussianProcess.ipynb:147)
// the residual needed by the backward, R = -rcoef * (F - Y) / var  (= d(rcoef*ll)/dF), and the two
// sums the scalar gradients need: sum E^2 (for d/dvar) and sum E*F (for d/dk_var).
#include "kernels.cuh"

namespace hb {

namespace {

__global__ void __launch_bounds__(256) gauss_fwd_kernel(const float* __restrict__ f, const float* __restrict__ f_scale,
                                                        const float* __restrict__ y,
                                                        long long total, long long yp, long long ydiv,
                                                        const float* __restrict__ var,
                                                        float rcoef, float* __restrict__ resid, double* partials) {
  __shared__ double red[64];
  double acc[2] = {0.0, 0.0};
  const float iv = 1.f / __ldg(var);
  const float rs = -rcoef * iv;
  const float fs = f_scale ? __ldg(f_scale) : 1.f;
  // ydiv > 0: sample-minor layout f[r * ydiv + s], y read at e / ydiv (one y per row of ydiv samples)
  const bool vec = ((total & 3) == 0) && (ydiv > 0 ? ((ydiv & 3) == 0) : (((yp & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0))) &&
                   ((reinterpret_cast<uintptr_t>(f) & 15) == 0) && (!resid || (reinterpret_cast<uintptr_t>(resid) & 15) == 0);
  const long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x, gsz = (long long)gridDim.x * blockDim.x;
  if (vec) {
    for (long long g = gid; g < total / 4; g += gsz) {
      const long long e0 = 4 * g;
      float4 ff = __ldg(reinterpret_cast<const float4*>(f + e0));
      ff.x *= fs; ff.y *= fs; ff.z *= fs; ff.w *= fs;
      float4 yy;
      if (ydiv > 0) { const float y1 = __ldg(y + e0 / ydiv); yy = make_float4(y1, y1, y1, y1); }
      else yy = __ldg(reinterpret_cast<const float4*>(y + (e0 % yp)));
      const float e[4] = {ff.x - yy.x, ff.y - yy.y, ff.z - yy.z, ff.w - yy.w};
      const float fv[4] = {ff.x, ff.y, ff.z, ff.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        acc[0] += (double)(e[c] * e[c]);
        acc[1] += (double)(e[c] * fv[c]);
      }
      if (resid) *reinterpret_cast<float4*>(resid + e0) = make_float4(rs * e[0], rs * e[1], rs * e[2], rs * e[3]);
    }
  } else {
    for (long long e0 = gid; e0 < total; e0 += gsz) {
      const float ff = fs * f[e0];
      const float e = ff - (ydiv > 0 ? y[e0 / ydiv] : y[e0 % yp]);
      acc[0] += (double)(e * e);
      acc[1] += (double)(e * ff);
      if (resid) resid[e0] = rs * e;
    }
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    partials[2 * blockIdx.x] = acc[0];
    partials[2 * blockIdx.x + 1] = acc[1];
  }
}

__global__ void gauss_finalize_kernel(const double* __restrict__ partials, int nblocks, long long total,
                                      const float* __restrict__ var, float* out3) {
  double s0 = 0.0, s1 = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) {
    s0 += partials[2 * b];
    s1 += partials[2 * b + 1];
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (threadIdx.x == 0) {
    const double v = (double)*var;
    out3[0] = (float)(-0.5 * (double)total * (1.8378770664093453 + log(v)) - 0.5 * s0 / v);
    out3[1] = (float)s0;
    out3[2] = (float)s1;
  }
}

__global__ void gaussian_logpdf_kernel(const float* __restrict__ x, long long xp, const float* __restrict__ mu,
                                       long long mp, const float* __restrict__ var, long long vp, long long total,
                                       float* out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const float v = var[e % vp];
    const float d = mu[e % mp] - x[e % xp];
    out[e] = -0.9189385332046727f - 0.5f * logf(v) - 0.5f * d * d / v;
  }
}

// d out / d mu and d out / d var of the elementwise log-density, times the incoming gradient g.
__global__ void gaussian_logpdf_bwd_kernel(const float* __restrict__ x, long long xp, const float* __restrict__ mu,
                                           long long mp, const float* __restrict__ var, long long vp, long long total,
                                           const float* __restrict__ g, float* dmu, float* dvar) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const float v = var[e % vp];
    const float d = x[e % xp] - mu[e % mp];
    const float iv = 1.f / v;
    const float gg = g[e];
    if (dmu) dmu[e] = gg * d * iv;
    if (dvar) dvar[e] = gg * (-0.5f * iv + 0.5f * d * d * iv * iv);
  }
}

}  // namespace

int gaussian_logpdf_bwd(const float* x, long long x_period, const float* mu, long long mu_period, const float* var,
                        long long var_period, long long total, const float* g, float* dmu, float* dvar,
                        cudaStream_t st) {
  if (total < 0) return HB_ERR_ARG;
  if (total == 0) return HB_OK;
  if (!x || !mu || !var || !g || x_period <= 0 || mu_period <= 0 || var_period <= 0) return HB_ERR_ARG;
  long long nb = (total + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  gaussian_logpdf_bwd_kernel<<<(int)nb, 256, 0, st>>>(x, x_period, mu, mu_period, var, var_period, total, g, dmu, dvar);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int gauss_loglik_fwd(const float* f, const float* f_scale, const float* y, long long total, long long y_period,
                     const float* var, float rcoef, float* resid, float* out3, void* ws, size_t ws_bytes, cudaStream_t st) {
  return gauss_loglik_fwd_ex(f, f_scale, y, total, y_period, 0, var, rcoef, resid, out3, ws, ws_bytes, st);
}

// y_div > 0: f is [rows, y_div] (sample-minor) and y has one entry per row; otherwise y is read at e % y_period.
int gauss_loglik_fwd_ex(const float* f, const float* f_scale, const float* y, long long total, long long y_period,
                        long long y_div, const float* var, float rcoef, float* resid, float* out3, void* ws, size_t ws_bytes,
                        cudaStream_t st) {
  if (total < 0 || y_period <= 0 || y_div < 0 || !var || !out3) return HB_ERR_ARG;
  if (total > 0 && (!f || !y)) return HB_ERR_ARG;
  if (!ws || ws_bytes < kReduceWsBytes) return HB_ERR_WORKSPACE;
  long long nb = (total / 4 + 255) / 256;
  if (nb < 1) nb = 1;
  if (nb > kReduceBlocks) nb = kReduceBlocks;
  double* partials = reinterpret_cast<double*>(ws);
  gauss_fwd_kernel<<<(int)nb, 256, 0, st>>>(f, f_scale, y, total, y_period, y_div, var, rcoef, resid, partials);
  HB_CHECK_LAUNCH();
  gauss_finalize_kernel<<<1, 32, 0, st>>>(partials, (int)nb, total, var, out3);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int gaussian_logpdf(const float* x, long long x_period, const float* mu, long long mu_period, const float* var,
                    long long var_period, long long total, float* out, cudaStream_t st) {
  if (total < 0) return HB_ERR_ARG;
  if (total == 0) return HB_OK;
  if (!x || !mu || !var || !out || x_period <= 0 || mu_period <= 0 || var_period <= 0) return HB_ERR_ARG;
  long long nb = (total + 255) / 256;
  if (nb > 148 * 16) nb = 148 * 16;
  gaussian_logpdf_kernel<<<(int)nb, 256, 0, st>>>(x, x_period, mu, mu_period, var, var_period, total, out);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

}  // namespace hb
