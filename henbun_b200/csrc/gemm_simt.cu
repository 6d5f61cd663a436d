// fp32 SIMT GEMM for sm_100a: register-tiled 128x128x8 (or 64x128x8) CTA tiles, double-buffered
// shared memory, float4 global/shared accesses, triangular operand masks with k-range trimming,
// fused bias/activation/clip epilogue, batched.  This is the exact-fp32 engine; the tcgen05
// 3xTF32 engine (gemm_tc.cu) takes over for large aligned shapes.
#include "gemm.cuh"

namespace hb {

namespace {

constexpr int BK = 8;
constexpr int NT = 256;

__device__ __forceinline__ bool keep_rk(int mode, int r, int k) {
  switch (mode) {
    case 0: return true;
    case 1: return k <= r;
    case 2: return k >= r;
    case 3: return k > r;
    default: return k < r;
  }
}

// does the rk-mask zero anything inside the tile rows [r0,r0+ROWS) x k [k0,k0+BK) ?
__device__ __forceinline__ bool mask_crosses(int mode, int r0, int rows, int k0) {
  switch (mode) {
    case 0: return false;
    case 1: return k0 + BK - 1 > r0;
    case 2: return k0 < r0 + rows - 1;
    case 3: return k0 <= r0 + rows - 1;
    default: return k0 + BK - 1 >= r0;
  }
}

template <int ROWS, bool KMAJOR>
struct Tile {
  static constexpr int NG = ROWS * BK / 4;
  static constexpr int NV = (NG + NT - 1) / NT;

  // KMAJOR: element (r,k) at P[r*ld + k];  else at P[k*ld + r]
  __device__ static __forceinline__ void load(float4 (&v)[NV], const float* __restrict__ P, long long ld,
                                              int r0, int k0, int R, int Kend, int mode, bool vec_ok,
                                              int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int g = tid + i * NT;
      if (NG % NT != 0 && g >= NG) continue;
      if (KMAJOR) {
        const int gr = r0 + (g >> 1), gk = k0 + (g & 1) * 4;
        const float* src = P + (long long)gr * ld + gk;
        if (vec_ok && mode == 0 && gr < R && gk + 3 < Kend) {
          v[i] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          float t[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            t[c] = (gr < R && gk + c < Kend && keep_rk(mode, gr, gk + c)) ? __ldg(src + c) : 0.f;
          v[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      } else {
        const int gk = k0 + g / (ROWS / 4), gr = r0 + (g % (ROWS / 4)) * 4;
        const float* src = P + (long long)gk * ld + gr;
        if (vec_ok && mode == 0 && gk < Kend && gr + 3 < R) {
          v[i] = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          float t[4];
#pragma unroll
          for (int c = 0; c < 4; ++c)
            t[c] = (gk < Kend && gr + c < R && keep_rk(mode, gr + c, gk)) ? __ldg(src + c) : 0.f;
          v[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
  }

  __device__ static __forceinline__ void store(const float4 (&v)[NV], float (*S)[ROWS + 4], int tid) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int g = tid + i * NT;
      if (NG % NT != 0 && g >= NG) continue;
      if (KMAJOR) {
        const int r = g >> 1, kq = (g & 1) * 4;
        S[kq + 0][r] = v[i].x; S[kq + 1][r] = v[i].y; S[kq + 2][r] = v[i].z; S[kq + 3][r] = v[i].w;
      } else {
        const int k = g / (ROWS / 4), rq = (g % (ROWS / 4)) * 4;
        *reinterpret_cast<float4*>(&S[k][rq]) = v[i];
      }
    }
  }
};

__device__ __forceinline__ float epilogue_act(float x, int act, int clip, float lo, float hi) {
  if (clip) x = fminf(fmaxf(x, lo), hi);
  switch (act) {
    case ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_TANH: return tanhf(x);
    default: return x;
  }
}

struct KParams {
  const float* A; const float* B; float* C; const float* bias;
  long long lda, ldb, ldc, sA, sB, sC, sBias;
  int M, N, K;
  float alpha, beta;
  int a_mode, b_mode, c_tri;   // a_mode/b_mode are unified (r,k) modes
  int act, clip; float clip_lo, clip_hi;
  int tiles_m, tiles_n;
  int vecA, vecB, vecC;
};

template <int BM, int TM, bool TA, bool TB>
__global__ void __launch_bounds__(NT, 2) gemm_simt_kernel(const KParams p) {
  constexpr int BN = 128, TN = 8;
  using TileA = Tile<BM, !TA>;   // transA=0 -> stored [M x K] -> K-major
  using TileB = Tile<BN, TB>;    // transB=1 -> stored [N x K] -> K-major
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  // grouped rasterisation (8 row-tiles per group) for L2 reuse of the B panels
  constexpr int GROUP = 8;
  const int bid = blockIdx.x;
  const int per_group = GROUP * p.tiles_n;
  const int first_m = (bid / per_group) * GROUP;
  const int gsz = min(p.tiles_m - first_m, GROUP);
  const int tm = first_m + (bid % per_group) % gsz;
  const int tn = (bid % per_group) / gsz;
  const int m0 = tm * BM, n0 = tn * BN;
  if (p.c_tri == 1 && n0 > m0 + BM - 1) return;

  const int bz = blockIdx.y;
  const float* A = p.A + (long long)bz * p.sA;
  const float* B = p.B + (long long)bz * p.sB;
  float* C = p.C + (long long)bz * p.sC;

  // k range trimmed by the triangular masks
  int kb = 0, ke = p.K;
  switch (p.a_mode) {
    case 1: ke = min(ke, m0 + BM); break;
    case 2: kb = max(kb, m0); break;
    case 3: kb = max(kb, m0 + 1); break;
    case 4: ke = min(ke, m0 + BM - 1); break;
    default: break;
  }
  switch (p.b_mode) {
    case 1: ke = min(ke, n0 + BN); break;
    case 2: kb = max(kb, n0); break;
    case 3: kb = max(kb, n0 + 1); break;
    case 4: ke = min(ke, n0 + BN - 1); break;
    default: break;
  }
  kb = (kb / BK) * BK;
  const int nk = ke > kb ? (ke - kb + BK - 1) / BK : 0;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[TileA::NV], rb[TileB::NV];
  if (nk > 0) {
    TileA::load(ra, A, p.lda, m0, kb, p.M, ke, mask_crosses(p.a_mode, m0, BM, kb) ? p.a_mode : 0, p.vecA, tid);
    TileB::load(rb, B, p.ldb, n0, kb, p.N, ke, mask_crosses(p.b_mode, n0, BN, kb) ? p.b_mode : 0, p.vecB, tid);
    TileA::store(ra, As[0], tid);
    TileB::store(rb, Bs[0], tid);
  }
  __syncthreads();

  for (int t = 0; t < nk; ++t) {
    const int cur = t & 1;
    if (t + 1 < nk) {
      const int k0 = kb + (t + 1) * BK;
      TileA::load(ra, A, p.lda, m0, k0, p.M, ke, mask_crosses(p.a_mode, m0, BM, k0) ? p.a_mode : 0, p.vecA, tid);
      TileB::load(rb, B, p.ldb, n0, k0, p.N, ke, mask_crosses(p.b_mode, n0, BN, k0) ? p.b_mode : 0, p.vecB, tid);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
      {
        const float4 v = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
        if (TM == 8) {
          const float4 w = *reinterpret_cast<const float4*>(&As[cur][kk][BM / 2 + ty * 4]);
          a[TM - 4] = w.x; a[TM - 3] = w.y; a[TM - 2] = w.z; a[TM - 1] = w.w;
        }
      }
      {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
        const float4 w = *reinterpret_cast<const float4*>(&Bs[cur][kk][BN / 2 + tx * 4]);
        b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        b[4] = w.x; b[5] = w.y; b[6] = w.z; b[7] = w.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < nk) {
      TileA::store(ra, As[cur ^ 1], tid);
      TileB::store(rb, Bs[cur ^ 1], tid);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const float* bias = p.bias ? p.bias + (long long)bz * p.sBias : nullptr;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int li = (TM == 8 && i >= 4) ? (BM / 2 + ty * 4 + (i - 4)) : (ty * 4 + i);
    const int gi = m0 + li;
    if (gi >= p.M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int gj = n0 + jh * (BN / 2) + tx * 4;
      if (gj >= p.N) continue;
      float* cp = C + (long long)gi * p.ldc + gj;
      const bool full = (gj + 3 < p.N) && !(p.c_tri == 1 && gj + 3 > gi);
      float o[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) o[c] = p.alpha * acc[i][jh * 4 + c];
      if (full && p.vecC) {
        if (p.beta != 0.f) {
          const float4 old = *reinterpret_cast<const float4*>(cp);
          o[0] = fmaf(p.beta, old.x, o[0]); o[1] = fmaf(p.beta, old.y, o[1]);
          o[2] = fmaf(p.beta, old.z, o[2]); o[3] = fmaf(p.beta, old.w, o[3]);
        }
        if (bias) {
#pragma unroll
          for (int c = 0; c < 4; ++c) o[c] += __ldg(bias + gj + c);
        }
        if (p.act != ACT_NONE || p.clip) {
#pragma unroll
          for (int c = 0; c < 4; ++c) o[c] = epilogue_act(o[c], p.act, p.clip, p.clip_lo, p.clip_hi);
        }
        *reinterpret_cast<float4*>(cp) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (gj + c >= p.N) continue;
          if (p.c_tri == 1 && gj + c > gi) continue;
          float x = o[c];
          if (p.beta != 0.f) x = fmaf(p.beta, cp[c], x);
          if (bias) x += __ldg(bias + gj + c);
          if (p.act != ACT_NONE || p.clip) x = epilogue_act(x, p.act, p.clip, p.clip_lo, p.clip_hi);
          cp[c] = x;
        }
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BM, int TM>
int launch(const KParams& kp, int transA, int transB, int batch, cudaStream_t st) {
  dim3 grid((unsigned)(kp.tiles_m * kp.tiles_n), (unsigned)batch, 1);
  if (!transA && !transB) gemm_simt_kernel<BM, TM, false, false><<<grid, NT, 0, st>>>(kp);
  else if (!transA && transB) gemm_simt_kernel<BM, TM, false, true><<<grid, NT, 0, st>>>(kp);
  else if (transA && !transB) gemm_simt_kernel<BM, TM, true, false><<<grid, NT, 0, st>>>(kp);
  else gemm_simt_kernel<BM, TM, true, true><<<grid, NT, 0, st>>>(kp);
  HB_CHECK_LAUNCH();
  return HB_OK;
}

}  // namespace

int gemm_simt(const GemmParams& p, cudaStream_t st) {
  if (p.M < 0 || p.N < 0 || p.K < 0 || p.batch < 0) return HB_ERR_ARG;
  if (p.M == 0 || p.N == 0 || p.batch == 0) return HB_OK;
  if (!p.C || (p.K > 0 && (!p.A || !p.B))) return HB_ERR_ARG;
  if (p.batch > 65535) return HB_ERR_ARG;
  if (p.a_tri < 0 || p.a_tri > 4 || p.b_tri < 0 || p.b_tri > 4 || p.c_tri < 0 || p.c_tri > 1) return HB_ERR_ARG;
  KParams kp;
  kp.A = p.A; kp.B = p.B; kp.C = p.C; kp.bias = p.bias;
  kp.lda = p.lda; kp.ldb = p.ldb; kp.ldc = p.ldc;
  kp.sA = p.sA; kp.sB = p.sB; kp.sC = p.sC; kp.sBias = p.sBias;
  kp.M = p.M; kp.N = p.N; kp.K = p.K;
  kp.alpha = p.alpha; kp.beta = p.beta;
  kp.a_mode = p.a_tri;
  static const int b2rk[5] = {0, 2, 1, 4, 3};
  kp.b_mode = b2rk[p.b_tri];
  kp.c_tri = p.c_tri;
  kp.act = p.act; kp.clip = p.clip; kp.clip_lo = p.clip_lo; kp.clip_hi = p.clip_hi;
  kp.vecA = aligned16(p.A) && (p.lda % 4 == 0) && (p.sA % 4 == 0);
  kp.vecB = aligned16(p.B) && (p.ldb % 4 == 0) && (p.sB % 4 == 0);
  kp.vecC = aligned16(p.C) && (p.ldc % 4 == 0) && (p.sC % 4 == 0);
  // 64-row tiles when the row count is small or when 128-row tiles would leave most of the 148 SMs idle (the exact-fp32
  // factorisations of order <= 2048 live on such shapes)
  const bool small_m = p.M <= 64 || (long long)cdiv(p.M, 128) * cdiv(p.N, 128) * p.batch < 148;
  const int BM = small_m ? 64 : 128;
  kp.tiles_m = cdiv(p.M, BM);
  kp.tiles_n = cdiv(p.N, 128);
  if (small_m) return launch<64, 4>(kp, p.transA, p.transB, p.batch, st);
  return launch<128, 8>(kp, p.transA, p.transB, p.batch, st);
}

}  // namespace hb
