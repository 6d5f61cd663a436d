// Shared device/host helpers for the henbun_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

// ---- error codes (mirrored in include/henbun_b200.h) ----
#define HB_OK 0
#define HB_ERR_ARG 1      // bad shape / pointer / flag (reference: ValueError / AssertionError)
#define HB_ERR_CUDA 2     // a CUDA runtime call or launch failed
#define HB_ERR_WORKSPACE 3  // workspace too small

namespace hb {

extern unsigned long long g_launches;   // number of kernels this library launched (host counter)

#define HB_CHECK_LAUNCH()                                   \
  do {                                                      \
    ++hb::g_launches;                                       \
    if (cudaGetLastError() != cudaSuccess) return HB_ERR_CUDA; \
  } while (0)

#define HB_TRY(expr)                 \
  do {                               \
    int _rc = (expr);                \
    if (_rc != HB_OK) return _rc;    \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---------------- reductions ----------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  blockDim.x multiple of 32, <= 1024.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /* >= 32*NV */) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[w * NV + i] = v[i];
  }
  __syncthreads();
  if (w == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double x = (lane < nw) ? smem[lane * NV + i] : 0.0;
      v[i] = warp_sum(x);
    }
  }
  __syncthreads();
}

// ---------------- math ----------------
__device__ __forceinline__ float softplus_f(float x) {
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------- Philox-4x32-10 counter RNG ----------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox-4x32-10 of Salmon et al. (Random123): 128-bit counter c, 64-bit key (k0, k1); c is replaced by the output.
__device__ __forceinline__ void philox4x32_10_core(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// The library's streams use counter = (ctr, 0, 0) and key = seed.
__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  philox4x32_10_core(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// Four standard normals for counter block `ctr` (element indices 4*ctr .. 4*ctr+3). Box-Muller.
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint64_t ctr, float (&z)[4]) {
  uint32_t r[4];
  philox4x32_10(seed, ctr, r);
  const float k = 2.3283064365386963e-10f;  // 2^-32
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float u1 = ((float)r[2 * i] + 0.5f) * k;       // (0,1]
    float u2 = ((float)r[2 * i + 1] + 0.5f) * k;
    u1 = fminf(fmaxf(u1, 1e-12f), 1.0f);
    float rad = sqrtf(-2.f * logf(u1));
    float s, c;
    sincospif(2.f * u2, &s, &c);
    z[2 * i] = rad * c;
    z[2 * i + 1] = rad * s;
  }
}

// The normal for flat element index idx of stream (seed, offset): used identically by the
// standalone generator and by kernels that regenerate eps instead of reading it.
__device__ __forceinline__ float philox_normal_at(uint64_t seed, uint64_t offset, uint64_t idx) {
  uint64_t g = offset + idx;
  float z[4];
  philox_normal4(seed, g >> 2, z);
  return z[g & 3];
}

}  // namespace hb
