// Constrained-parameter transforms of Henbun/transforms.py as kernels: forward y = T(x), its log-Jacobian sum, and both
// backward passes.  Used for hyper-parameters (O(1) elements) and for whole sample tensors of transformed variationals
// (variationals.py:110,129,204-208: [S, n] elements, HBM-bound: 8 B per element forward, 12 B backward).
//   kind 1 Exp      (transforms.py:90-107)   y = exp(x) + lower                 log|J| = sum x
//   kind 2 Log1pe   (transforms.py:110-143)  y = log(1 + exp(x)) + lower        log|J| = -sum log(1 + exp(-x))
//   kind 3 Logistic (transforms.py:146-180)  y = a + (b - a) / (1 + exp(-x))    log|J| = sum x - 2 log(exp(x) + 1) + log(b - a)
// (kind 0 Identity needs no kernel.)  p0 = lower or a, p1 = b.
#include "kernels.cuh"

namespace hb {

namespace {

template <int KIND>
__device__ __forceinline__ float tr_fwd(float x, float p0, float p1) {
  if (KIND == 1) return expf(x) + p0;
  if (KIND == 2) return softplus_f(x) + p0;
  return p0 + (p1 - p0) * sigmoid_f(x);
}
template <int KIND>
__device__ __forceinline__ float tr_dfwd(float x, float p0, float p1) {
  if (KIND == 1) return expf(x);
  if (KIND == 2) return sigmoid_f(x);
  const float s = sigmoid_f(x);
  return (p1 - p0) * s * (1.f - s);
}
template <int KIND>
__device__ __forceinline__ float tr_lj(float x, float p0, float p1) {
  if (KIND == 1) return x;
  if (KIND == 2) return -softplus_f(-x);
  return x - 2.f * softplus_f(x) + logf(p1 - p0);
}
template <int KIND>
__device__ __forceinline__ float tr_dlj(float x) {
  if (KIND == 1) return 1.f;
  if (KIND == 2) return sigmoid_f(-x);
  return 1.f - 2.f * sigmoid_f(x);
}

// MODE 0: y = T(x);  1: gx = gy * T'(x);  3: gx = g1 * d logjac / dx
template <int KIND, int MODE>
__global__ void __launch_bounds__(256) transform_map_kernel(const float* __restrict__ x, long long total, float p0, float p1,
                                                            const float* __restrict__ g, float* __restrict__ out) {
  const float g1 = (MODE == 3) ? __ldg(g) : 0.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const float xv = __ldg(x + e);
    if (MODE == 0) out[e] = tr_fwd<KIND>(xv, p0, p1);
    else if (MODE == 1) out[e] = __ldg(g + e) * tr_dfwd<KIND>(xv, p0, p1);
    else out[e] = g1 * tr_dlj<KIND>(xv);
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) transform_logjac_kernel(const float* __restrict__ x, long long total, float p0, float p1,
                                                               double* partials) {
  __shared__ double red[32];
  double acc[1] = {0.0};
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x)
    acc[0] += (double)tr_lj<KIND>(__ldg(x + e), p0, p1);
  block_sum<1>(acc, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}

__global__ void transform_logjac_finalize_kernel(const double* __restrict__ partials, int nblocks, float* out1) {
  double s = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) s += partials[b];
  s = warp_sum(s);
  if (threadIdx.x == 0) *out1 = (float)s;
}

inline int blocks_for(long long total, int cap) {
  long long nb = (total + 255) / 256;
  if (nb < 1) nb = 1;
  if (nb > cap) nb = cap;
  return (int)nb;
}

template <int MODE>
int launch_map(int kind, const float* x, long long total, float p0, float p1, const float* g, float* out, cudaStream_t st) {
  const int nb = blocks_for(total, 148 * 8);
  switch (kind) {
    case 1: transform_map_kernel<1, MODE><<<nb, 256, 0, st>>>(x, total, p0, p1, g, out); break;
    case 2: transform_map_kernel<2, MODE><<<nb, 256, 0, st>>>(x, total, p0, p1, g, out); break;
    case 3: transform_map_kernel<3, MODE><<<nb, 256, 0, st>>>(x, total, p0, p1, g, out); break;
    default: return HB_ERR_ARG;
  }
  HB_CHECK_LAUNCH();
  return HB_OK;
}

inline bool args_ok(int kind, float p0, float p1) { return kind >= 1 && kind <= 3 && (kind != 3 || p1 > p0); }

}  // namespace

int transform_fwd(int kind, const float* x, long long total, float p0, float p1, float* y, cudaStream_t st) {
  if (!args_ok(kind, p0, p1) || total < 0) return HB_ERR_ARG;
  if (total == 0) return HB_OK;
  if (!x || !y) return HB_ERR_ARG;
  return launch_map<0>(kind, x, total, p0, p1, nullptr, y, st);
}

int transform_bwd(int kind, const float* x, long long total, float p0, float p1, const float* gy, float* gx, cudaStream_t st) {
  if (!args_ok(kind, p0, p1) || total < 0) return HB_ERR_ARG;
  if (total == 0) return HB_OK;
  if (!x || !gy || !gx) return HB_ERR_ARG;
  return launch_map<1>(kind, x, total, p0, p1, gy, gx, st);
}

int transform_logjac(int kind, const float* x, long long total, float p0, float p1, float* out1, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  if (!args_ok(kind, p0, p1) || total < 0 || !out1) return HB_ERR_ARG;
  if (total == 0) return fill_f32(out1, 1, 0.f, st);
  if (!x) return HB_ERR_ARG;
  if (!ws || ws_bytes < kReduceWsBytes) return HB_ERR_WORKSPACE;
  const int nb = blocks_for(total, kReduceBlocks);
  double* partials = reinterpret_cast<double*>(ws);
  switch (kind) {
    case 1: transform_logjac_kernel<1><<<nb, 256, 0, st>>>(x, total, p0, p1, partials); break;
    case 2: transform_logjac_kernel<2><<<nb, 256, 0, st>>>(x, total, p0, p1, partials); break;
    default: transform_logjac_kernel<3><<<nb, 256, 0, st>>>(x, total, p0, p1, partials); break;
  }
  HB_CHECK_LAUNCH();
  transform_logjac_finalize_kernel<<<1, 32, 0, st>>>(partials, nb, out1);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

int transform_logjac_bwd(int kind, const float* x, long long total, float p0, float p1, const float* g1, float* gx,
                         cudaStream_t st) {
  if (!args_ok(kind, p0, p1) || total < 0) return HB_ERR_ARG;
  if (total == 0) return HB_OK;
  if (!x || !g1 || !gx) return HB_ERR_ARG;
  return launch_map<3>(kind, x, total, p0, p1, g1, gx, st);
}

}  // namespace hb
