// The elementwise log-densities of Henbun/densities.py:25-103 as one kernel family, forward and backward.
//
// The reference emits 5-15 TF elementwise ops (plus broadcasts) per density; here one pass reads every operand once and
// writes the log-density (forward) or g * d logp / d operand (backward).  Operands broadcast "modularly"
// (operand i is read at flat index e % period_i: a suffix-shaped operand, a scalar, or a full tensor), which covers
// every call in the reference's notebooks and tests; other broadcasts are expanded by the host.
//
// HBM-bound: 4 B per full-size operand read + 4 B per full-size output written.  No 64-bit division in the inner
// loop (the periodic index advances incrementally); float4 path when every operand allows it.  A scalar operand's
// gradient (period 1, e.g. the noise variance `var` of every objective in the notebooks) is reduced in-kernel
// (fp64 per-thread accumulation, deterministic two-stage sum) instead of materialising a full-size array.
//
// Operand order = the reference function's argument order:
//   0 gaussian(x, mu, var)            densities.py:25-27      5 gamma(shape, scale, x)               :49-51
//   1 lognormal(x, mu, var)           :30-32                  6 student_t(x, mean, scale, deg_free)  :54-61
//   2 bernoulli(p, y)                 :35-36                  7 beta(alpha, beta, y)                 :64-70
//   3 poisson(lamb, y)                :39-40                  8 laplace(mu, sigma, y)                :73-74
//   4 exponential(lamb, y)            :43-44                  9 bimixture(fraction, logp0, logp1)    :95-103
#include "kernels.cuh"

namespace hb {

namespace {

constexpr int kMaxArgs = 4;
constexpr float kHalfLog2Pi = 0.9189385332046727f;

struct DensityArgs {
  const float* p[kMaxArgs];
  long long period[kMaxArgs];
};
struct DensityGrads {
  float* d[kMaxArgs];      // NULL = not wanted
  int reduce[kMaxArgs];    // 1 = operand is a scalar (period 1): d[i] is a partial-sum slot, not an array
};

__host__ __device__ constexpr int density_nargs(int kind) {
  return kind == 0 ? 3 : kind == 1 ? 3 : kind == 2 ? 2 : kind == 3 ? 2 : kind == 4 ? 2 : kind == 5 ? 3 : kind == 6 ? 4
       : kind == 7 ? 3 : kind == 8 ? 3 : 3;
}

// psi(x), x > 0, in double: recurrence up to x >= 6, then the asymptotic series (|err| < 1e-12).
__device__ __forceinline__ double digamma_d(double x) {
  if (!(x > 0.0)) return nan("");
  double r = 0.0;
  while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
  const double f = 1.0 / (x * x);
  return r + log(x) - 0.5 / x -
         f * (1.0 / 12 - f * (1.0 / 120 - f * (1.0 / 252 - f * (1.0 / 240 - f * (1.0 / 132 - f * (691.0 / 32760))))));
}

// log-density of one element.  v[] = operands in the reference's argument order.
template <int KIND>
__device__ __forceinline__ float density_eval(const float (&v)[kMaxArgs]) {
  if (KIND == 0) {                                   // gaussian(x, mu, var)
    const float d = v[1] - v[0];
    return -kHalfLog2Pi - 0.5f * logf(v[2]) - 0.5f * d * d / v[2];
  } else if (KIND == 1) {                            // lognormal(x, mu, var) = gaussian(log x, mu, var) - log x
    const float l = logf(v[0]);
    const float d = v[1] - l;
    return -kHalfLog2Pi - 0.5f * logf(v[2]) - 0.5f * d * d / v[2] - l;
  } else if (KIND == 2) {                            // bernoulli(p, y)
    return logf(v[1] == 1.f ? v[0] : 1.f - v[0]);
  } else if (KIND == 3) {                            // poisson(lamb, y)
    return v[1] * logf(v[0]) - v[0] - lgammaf(v[1] + 1.f);
  } else if (KIND == 4) {                            // exponential(lamb, y)
    return -v[1] / v[0] - logf(v[0]);
  } else if (KIND == 5) {                            // gamma(shape, scale, x)
    return -v[0] * logf(v[1]) - lgammaf(v[0]) + (v[0] - 1.f) * logf(v[2]) - v[2] / v[1];
  } else if (KIND == 6) {                            // student_t(x, mean, scale, deg_free)
    const double nu = (double)v[3];
    // the lgamma difference cancels (both terms ~ nu log nu): evaluated in double
    const double c = lgamma(0.5 * (nu + 1.0)) - lgamma(0.5 * nu) -
                     0.5 * (log((double)v[2] * (double)v[2]) + log(nu) + 1.1447298858494002);
    const float t = (v[0] - v[1]) / v[2];
    return (float)(c - 0.5 * (nu + 1.0) * (double)log1pf(t * t / v[3]));
  } else if (KIND == 7) {                            // beta(alpha, beta, y), y clipped to [1e-6, 1-1e-6]
    // the clip bound 1 - 1e-6 is not an fp32 number (nearest: 1 - 1.013e-6, a 1.3 % error on 1 - y): clip and take
    // the logs in double
    const double y = fmin(fmax((double)v[2], 1e-6), 1.0 - 1e-6);
    const double a = (double)v[0], b = (double)v[1];
    const double c = lgamma(a + b) - lgamma(a) - lgamma(b);
    return (float)((a - 1.0) * log(y) + (b - 1.0) * log1p(-y) + c);
  } else if (KIND == 8) {                            // laplace(mu, sigma, y)
    return -fabsf(v[0] - v[2]) / v[1] - logf(2.f * v[1]);
  } else {                                           // bimixture(fraction, logp0, logp1)
    const float t0 = v[1] + logf(v[0]), t1 = v[2] + logf(1.f - v[0]);
    const float m = fmaxf(t0, t1);
    return m + logf(expf(t0 - m) + expf(t1 - m));
  }
}

// d logp / d operand, all operands.
template <int KIND>
__device__ __forceinline__ void density_grad(const float (&v)[kMaxArgs], float (&dv)[kMaxArgs]) {
  dv[0] = dv[1] = dv[2] = dv[3] = 0.f;
  if (KIND == 0) {
    const float iv = 1.f / v[2], d = v[0] - v[1];
    dv[0] = -d * iv; dv[1] = d * iv; dv[2] = -0.5f * iv + 0.5f * d * d * iv * iv;
  } else if (KIND == 1) {
    const float l = logf(v[0]), iv = 1.f / v[2], d = v[1] - l;
    dv[0] = (d * iv - 1.f) / v[0]; dv[1] = -d * iv; dv[2] = -0.5f * iv + 0.5f * d * d * iv * iv;
  } else if (KIND == 2) {
    dv[0] = v[1] == 1.f ? 1.f / v[0] : -1.f / (1.f - v[0]);
  } else if (KIND == 3) {
    dv[0] = v[1] / v[0] - 1.f;
    dv[1] = logf(v[0]) - (float)digamma_d((double)v[1] + 1.0);
  } else if (KIND == 4) {
    const float il = 1.f / v[0];
    dv[0] = v[1] * il * il - il; dv[1] = -il;
  } else if (KIND == 5) {
    const float is = 1.f / v[1];
    dv[0] = -logf(v[1]) - (float)digamma_d((double)v[0]) + logf(v[2]);
    dv[1] = -v[0] * is + v[2] * is * is;
    dv[2] = (v[0] - 1.f) / v[2] - is;
  } else if (KIND == 6) {
    const float nu = v[3], is = 1.f / v[2];
    const float t = (v[0] - v[1]) * is;
    const float q = 1.f + t * t / nu;
    const float w = (nu + 1.f) * t / (nu * q);                 // (nu+1) t / (nu q)
    dv[0] = -w * is; dv[1] = w * is; dv[2] = -is + w * t * is;
    const double dn = 0.5 * (digamma_d(0.5 * ((double)nu + 1.0)) - digamma_d(0.5 * (double)nu)) - 0.5 / (double)nu -
                      0.5 * (double)log1pf(t * t / nu) + 0.5 * (double)w * (double)t / (double)nu;
    dv[3] = (float)dn;
  } else if (KIND == 7) {
    const bool inside = (double)v[2] >= 1e-6 && (double)v[2] <= 1.0 - 1e-6;
    const double y = fmin(fmax((double)v[2], 1e-6), 1.0 - 1e-6);
    const double a = (double)v[0], b = (double)v[1];
    const double pab = digamma_d(a + b);
    dv[0] = (float)(log(y) + pab - digamma_d(a));
    dv[1] = (float)(log1p(-y) + pab - digamma_d(b));
    dv[2] = inside ? (float)((a - 1.0) / y - (b - 1.0) / (1.0 - y)) : 0.f;
  } else if (KIND == 8) {
    const float is = 1.f / v[1], d = v[0] - v[2];
    const float sg = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
    dv[0] = -sg * is; dv[1] = fabsf(d) * is * is - is; dv[2] = sg * is;
  } else {
    const float t0 = v[1] + logf(v[0]), t1 = v[2] + logf(1.f - v[0]);
    const float m = fmaxf(t0, t1);
    const float e0 = expf(t0 - m), e1 = expf(t1 - m);
    const float w0 = e0 / (e0 + e1), w1 = e1 / (e0 + e1);
    dv[0] = w0 / v[0] - w1 / (1.f - v[0]); dv[1] = w0; dv[2] = w1;
  }
}

// Incremental periodic index: idx = e % period without a division per element.
struct PeriodicIdx {
  long long idx, step, period;
  __device__ __forceinline__ void init(long long e0, long long stride, long long p) {
    period = p;
    idx = (p == 1) ? 0 : e0 % p;
    step = (p == 1) ? 0 : stride % p;
  }
  __device__ __forceinline__ void next() {
    idx += step;
    if (idx >= period) idx -= period;
  }
};

template <int KIND, bool VEC>
__global__ void __launch_bounds__(256) density_fwd_kernel(DensityArgs a, long long total, float* __restrict__ out) {
  constexpr int NA = density_nargs(KIND);
  constexpr int W = VEC ? 4 : 1;
  const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * W;
  const long long gsz = (long long)gridDim.x * blockDim.x * W;
  PeriodicIdx ix[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) ix[i].init(gid, gsz, a.period[i]);
  // densities.gaussian with a scalar variance (every objective of the notebooks): log and reciprocal once per thread
  // (the first ncu capture of this kernel showed it issue-bound at 80 %, not HBM-bound)
  const bool gfast = (KIND == 0) && a.period[2] == 1;
  float g_c0 = 0.f, g_iv = 0.f;
  if (gfast) { const float var = __ldg(a.p[2]); g_iv = 1.f / var; g_c0 = -kHalfLog2Pi - 0.5f * logf(var); }
  for (long long e = gid; e < total; e += gsz) {
    float v[W][kMaxArgs];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (VEC && a.period[i] != 1) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.p[i] + ix[i].idx));
        v[0][i] = q.x; v[W > 1 ? 1 : 0][i] = q.y; v[W > 1 ? 2 : 0][i] = q.z; v[W > 1 ? 3 : 0][i] = q.w;
      } else {
        const float s = __ldg(a.p[i] + ix[i].idx);
#pragma unroll
        for (int c = 0; c < W; ++c) v[c][i] = s;
      }
      ix[i].next();
    }
    float r[W];
#pragma unroll
    for (int c = 0; c < W; ++c) {
      if (gfast) { const float d = v[c][1] - v[c][0]; r[c] = g_c0 - 0.5f * d * d * g_iv; }
      else r[c] = density_eval<KIND>(v[c]);
    }
    if (VEC) *reinterpret_cast<float4*>(out + e) = make_float4(r[0], r[W > 1 ? 1 : 0], r[W > 1 ? 2 : 0], r[W > 1 ? 3 : 0]);
    else out[e] = r[0];
  }
}

template <int KIND, bool VEC>
__global__ void __launch_bounds__(256) density_bwd_kernel(DensityArgs a, long long total, const float* __restrict__ g,
                                                          long long g_period, DensityGrads o, double* partials) {
  constexpr int NA = density_nargs(KIND);
  constexpr int W = VEC ? 4 : 1;
  __shared__ double red[32 * kMaxArgs];
  const long long gid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * W;
  const long long gsz = (long long)gridDim.x * blockDim.x * W;
  PeriodicIdx ix[NA], ig;
#pragma unroll
  for (int i = 0; i < NA; ++i) ix[i].init(gid, gsz, a.period[i]);
  ig.init(gid, gsz, g_period);
  const bool gfast = (KIND == 0) && a.period[2] == 1;
  const float g_iv = gfast ? 1.f / __ldg(a.p[2]) : 0.f;
  double acc[kMaxArgs] = {0.0, 0.0, 0.0, 0.0};
  for (long long e = gid; e < total; e += gsz) {
    float v[W][kMaxArgs], gg[W];
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (VEC && a.period[i] != 1) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(a.p[i] + ix[i].idx));
        v[0][i] = q.x; v[W > 1 ? 1 : 0][i] = q.y; v[W > 1 ? 2 : 0][i] = q.z; v[W > 1 ? 3 : 0][i] = q.w;
      } else {
        const float s = __ldg(a.p[i] + ix[i].idx);
#pragma unroll
        for (int c = 0; c < W; ++c) v[c][i] = s;
      }
      ix[i].next();
    }
    if (VEC && g_period != 1) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(g + ig.idx));
      gg[0] = q.x; gg[W > 1 ? 1 : 0] = q.y; gg[W > 1 ? 2 : 0] = q.z; gg[W > 1 ? 3 : 0] = q.w;
    } else {
      const float s = __ldg(g + ig.idx);
#pragma unroll
      for (int c = 0; c < W; ++c) gg[c] = s;
    }
    ig.next();
    float dv[W][kMaxArgs];
#pragma unroll
    for (int c = 0; c < W; ++c) {
      if (gfast) {
        const float d = v[c][0] - v[c][1];
        dv[c][0] = -d * g_iv; dv[c][1] = d * g_iv; dv[c][2] = -0.5f * g_iv + 0.5f * d * d * g_iv * g_iv; dv[c][3] = 0.f;
      } else {
        density_grad<KIND>(v[c], dv[c]);
      }
#pragma unroll
      for (int i = 0; i < NA; ++i) dv[c][i] *= gg[c];
    }
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (!o.d[i]) continue;
      if (o.reduce[i]) {
#pragma unroll
        for (int c = 0; c < W; ++c) acc[i] += (double)dv[c][i];
      } else if (VEC) {
        *reinterpret_cast<float4*>(o.d[i] + e) = make_float4(dv[0][i], dv[W > 1 ? 1 : 0][i], dv[W > 1 ? 2 : 0][i], dv[W > 1 ? 3 : 0][i]);
      } else {
        o.d[i][e] = dv[0][i];
      }
    }
  }
  if (partials) {
    block_sum<kMaxArgs>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
      for (int i = 0; i < kMaxArgs; ++i) partials[kMaxArgs * blockIdx.x + i] = acc[i];
    }
  }
}

__global__ void density_bwd_finalize_kernel(const double* __restrict__ partials, int nblocks, DensityGrads o) {
  double s[kMaxArgs] = {0.0, 0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < nblocks; b += 32) {
#pragma unroll
    for (int i = 0; i < kMaxArgs; ++i) s[i] += partials[kMaxArgs * b + i];
  }
#pragma unroll
  for (int i = 0; i < kMaxArgs; ++i) s[i] = warp_sum(s[i]);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < kMaxArgs; ++i)
      if (o.d[i] && o.reduce[i]) o.d[i][0] = (float)s[i];
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int KIND>
int launch_fwd(const DensityArgs& a, long long total, float* out, bool vec, cudaStream_t st) {
  const long long per = vec ? 4 : 1;
  long long nb = (total / per + 255) / 256;
  if (nb < 1) nb = 1;
  if (nb > 148 * 8) nb = 148 * 8;
  if (vec) density_fwd_kernel<KIND, true><<<(int)nb, 256, 0, st>>>(a, total, out);
  else density_fwd_kernel<KIND, false><<<(int)nb, 256, 0, st>>>(a, total, out);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

template <int KIND>
int launch_bwd(const DensityArgs& a, long long total, const float* g, long long gp, const DensityGrads& o, bool vec,
               bool any_reduce, double* partials, cudaStream_t st) {
  const long long per = vec ? 4 : 1;
  long long nb = (total / per + 255) / 256;
  if (nb < 1) nb = 1;
  const long long cap = any_reduce ? kReduceBlocks : 148 * 8;
  if (nb > cap) nb = cap;
  if (vec) density_bwd_kernel<KIND, true><<<(int)nb, 256, 0, st>>>(a, total, g, gp, o, any_reduce ? partials : nullptr);
  else density_bwd_kernel<KIND, false><<<(int)nb, 256, 0, st>>>(a, total, g, gp, o, any_reduce ? partials : nullptr);
  HB_CHECK_LAUNCH();
  if (any_reduce) {
    density_bwd_finalize_kernel<<<1, 32, 0, st>>>(partials, (int)nb, o);
    HB_CHECK_LAUNCH();
  }
  return HB_OK;
}

}  // namespace

int density_nargs_host(int kind) { return (kind < 0 || kind > 9) ? -1 : density_nargs(kind); }

int density_logpdf(int kind, const float* const* args, const long long* periods, long long total, float* out,
                   cudaStream_t st) {
  const int na = density_nargs_host(kind);
  if (na < 0 || total < 0 || !args || !periods) return HB_ERR_ARG;
  if (total == 0) return HB_OK;
  if (!out) return HB_ERR_ARG;
  DensityArgs a{};
  bool vec = (total % 4 == 0) && aligned16(out);
  for (int i = 0; i < na; ++i) {
    if (!args[i] || periods[i] <= 0 || periods[i] > total) return HB_ERR_ARG;
    a.p[i] = args[i]; a.period[i] = periods[i];
    if (periods[i] != 1 && (periods[i] % 4 != 0 || !aligned16(args[i]))) vec = false;
  }
  switch (kind) {
    case 0: return launch_fwd<0>(a, total, out, vec, st);
    case 1: return launch_fwd<1>(a, total, out, vec, st);
    case 2: return launch_fwd<2>(a, total, out, vec, st);
    case 3: return launch_fwd<3>(a, total, out, vec, st);
    case 4: return launch_fwd<4>(a, total, out, vec, st);
    case 5: return launch_fwd<5>(a, total, out, vec, st);
    case 6: return launch_fwd<6>(a, total, out, vec, st);
    case 7: return launch_fwd<7>(a, total, out, vec, st);
    case 8: return launch_fwd<8>(a, total, out, vec, st);
    default: return launch_fwd<9>(a, total, out, vec, st);
  }
}

int density_logpdf_bwd(int kind, const float* const* args, const long long* periods, long long total, const float* g,
                       long long g_period, float* const* dargs, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int na = density_nargs_host(kind);
  if (na < 0 || total < 0 || !args || !periods || !dargs) return HB_ERR_ARG;
  if (total == 0) return HB_OK;
  if (!g || g_period <= 0 || g_period > total) return HB_ERR_ARG;
  DensityArgs a{};
  DensityGrads o{};
  bool vec = (total % 4 == 0) && (g_period == 1 || (g_period % 4 == 0 && aligned16(g)));
  bool any_reduce = false;
  for (int i = 0; i < na; ++i) {
    if (!args[i] || periods[i] <= 0 || periods[i] > total) return HB_ERR_ARG;
    a.p[i] = args[i]; a.period[i] = periods[i];
    if (periods[i] != 1 && (periods[i] % 4 != 0 || !aligned16(args[i]))) vec = false;
    o.d[i] = dargs[i];
    o.reduce[i] = (dargs[i] != nullptr && periods[i] == 1 && total > 1) ? 1 : 0;
    if (o.reduce[i]) any_reduce = true;
    else if (dargs[i] && !aligned16(dargs[i])) vec = false;
  }
  if (any_reduce && (!ws || ws_bytes < kReduceWsBytes)) return HB_ERR_WORKSPACE;
  double* partials = reinterpret_cast<double*>(ws);
  switch (kind) {
    case 0: return launch_bwd<0>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 1: return launch_bwd<1>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 2: return launch_bwd<2>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 3: return launch_bwd<3>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 4: return launch_bwd<4>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 5: return launch_bwd<5>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 6: return launch_bwd<6>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 7: return launch_bwd<7>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    case 8: return launch_bwd<8>(a, total, g, g_period, o, vec, any_reduce, partials, st);
    default: return launch_bwd<9>(a, total, g, g_period, o, vec, any_reduce, partials, st);
  }
}

}  // namespace hb
