// gemm.cuh -- fp32 SIMT tile GEMMs with pluggable operand loaders (exact-fp32 parity path).
//
//   gemm_nn_kernel : C[M][N] = A[M][K] * Bm[K][N]         A rows via ALoad (plain rows, or the
//                    implicit-im2col rows of a channels-last 3x3 convolution)
//   gemm_tn_kernel : C[m*rs + n*cs] += sum_k A[k][m] * Bm[k][n]    split-K over CTAs, atomics
//                    (weight gradients: the reduction runs over pixels / time steps)
//
// Tiles: 256 threads, BK = 16, cp.async multi-stage ring, register tile TM x TN per thread.
#pragma once
#include "common.cuh"

namespace bsed {

constexpr int GEMM_BK = 16;
constexpr int GEMM_APAD = 20;  // A tile row stride in floats: 16 + 4 -> conflict-free LDS.128 along k

// ---------------------------------------------------------------------------------------------
// A-operand loaders for gemm_nn: give the address of 16 consecutive k-values of row m
// ---------------------------------------------------------------------------------------------
struct PlainRows {
  const float* A;
  int lda;
  int M;
  struct Row {
    const float* p;
  };
  __device__ __forceinline__ Row row(int m) const {
    Row r;
    r.p = (m < M) ? A + (size_t)m * lda : nullptr;
    return r;
  }
  __device__ __forceinline__ const float* ptr(const Row& r, int kc) const {
    return r.p ? r.p + kc * GEMM_BK : nullptr;
  }
};

// rows = output pixels of a 3x3/s1/p1 convolution over channels-last X [B][T][F][Cin];
// k-chunk kc -> (tap = kc / (Cin/16), 16 input channels)
struct ConvRows {
  const float* X;
  int T, F, Cin, M;
  int cpt;  // chunks per tap = Cin / 16
  struct Row {
    const float* p;
    int t, f;
  };
  __device__ __forceinline__ Row row(int m) const {
    Row r;
    if (m < M) {
      r.f = m % F;
      r.t = (m / F) % T;
      r.p = X + (size_t)m * Cin;
    } else {
      r.p = nullptr;
      r.t = 0;
      r.f = 0;
    }
    return r;
  }
  __device__ __forceinline__ const float* ptr(const Row& r, int kc) const {
    int tap = kc / cpt;
    int c0 = (kc - tap * cpt) * GEMM_BK;
    int dt = tap / 3 - 1, df = tap % 3 - 1;
    int tt = r.t + dt, ff = r.f + df;
    bool ok = r.p && tt >= 0 && tt < T && ff >= 0 && ff < F;
    return ok ? r.p + ((long long)dt * F + df) * Cin + c0 : nullptr;
  }
};

struct NNEpilogue {
  float* C;
  int ldc;
  const float* bias;  // [N] or null
  int accumulate;     // C += result
};

template <int BM, int BN, int TM, int TN, int STAGES, class ALoad>
__global__ void __launch_bounds__(256) gemm_nn_kernel(ALoad aload, const float* __restrict__ Bm, int ldb,
                                                      int M, int N, int K, NNEpilogue epi) {
  constexpr int NCG = BN / TN;           // column groups
  constexpr int NPG = BM / TM;           // row groups
  static_assert(NCG * NPG == 256, "tile/threads mismatch");
  static_assert(TM == 8, "TM must be 8");
  static_assert(TN == 8 || TN == 4 || TN == 2, "TN");
  constexpr int A_STAGE = BM * GEMM_APAD;
  constexpr int B_STAGE = GEMM_BK * BN;
  constexpr int A_LOADS = BM * 4 / 256;  // 16B chunks per thread per stage
  constexpr int B_CHUNKS = GEMM_BK * BN / 4;

  extern __shared__ __align__(16) float smem[];
  float* As = smem;
  float* Bs = smem + STAGES * A_STAGE;

  const int tid = threadIdx.x;
  const int cg = tid % NCG;
  const int pg = tid / NCG;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nk = K / GEMM_BK;

  typename ALoad::Row arow[A_LOADS];
#pragma unroll
  for (int i = 0; i < A_LOADS; ++i) arow[i] = aload.row(m0 + (tid + 256 * i) / 4);

  auto load_stage = [&](int stage, int kc) {
    float* as = As + stage * A_STAGE;
    float* bs = Bs + stage * B_STAGE;
#pragma unroll
    for (int i = 0; i < A_LOADS; ++i) {
      int id = tid + 256 * i;
      int r = id / 4, q = id % 4;
      const float* src = aload.ptr(arow[i], kc);
      cp_async16(as + r * GEMM_APAD + q * 4, src ? src + q * 4 : (const float*)Bm, src != nullptr);
    }
#pragma unroll
    for (int id = tid; id < B_CHUNKS; id += 256) {
      int k = id / (BN / 4), c4 = id % (BN / 4);
      cp_async16(bs + k * BN + c4 * 4, Bm + (size_t)(kc * GEMM_BK + k) * ldb + n0 + c4 * 4, true);
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nxt = kc + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const float* as = As + (kc % STAGES) * A_STAGE + (pg * TM) * GEMM_APAD;
    const float* bs = Bs + (kc % STAGES) * B_STAGE;
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4) {
      float4 a[TM];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(as + i * GEMM_APAD + k4 * 4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float b[TN];
        const float* brow = bs + (k4 * 4 + kk) * BN;
        if constexpr (TN == 8) {
          float4 b0 = *reinterpret_cast<const float4*>(brow + cg * 4);
          float4 b1 = *reinterpret_cast<const float4*>(brow + BN / 2 + cg * 4);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
          b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
        } else if constexpr (TN == 4) {
          float4 b0 = *reinterpret_cast<const float4*>(brow + cg * 4);
          b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
        } else {
          float2 b0 = *reinterpret_cast<const float2*>(brow + cg * 2);
          b[0] = b0.x; b[1] = b0.y;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, b[j], acc[i][j]);
        }
      }
    }
  }
  cp_async_wait<0>();

  // epilogue
  int cols[TN];
  if constexpr (TN == 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      cols[j] = n0 + cg * 4 + j;
      cols[4 + j] = n0 + BN / 2 + cg * 4 + j;
    }
  } else {
#pragma unroll
    for (int j = 0; j < TN; ++j) cols[j] = n0 + cg * TN + j;
  }
  float bv[TN];
#pragma unroll
  for (int j = 0; j < TN; ++j) bv[j] = epi.bias ? epi.bias[cols[j]] : 0.f;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + pg * TM + i;
    if (m >= M) continue;
    float* crow = epi.C + (size_t)m * epi.ldc;
    constexpr int VW = TN == 2 ? 2 : 4;
#pragma unroll
    for (int j0 = 0; j0 < TN; j0 += VW) {
      float v[VW];
#pragma unroll
      for (int j = 0; j < VW; ++j) v[j] = acc[i][j0 + j] + bv[j0 + j];
      float* dst = crow + cols[j0];
      if constexpr (VW == 4) {
        float4 o = make_float4(v[0], v[1], v[2], v[3]);
        if (epi.accumulate) {
          float4 c = *reinterpret_cast<float4*>(dst);
          o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
        }
        *reinterpret_cast<float4*>(dst) = o;
      } else {
        float2 o = make_float2(v[0], v[1]);
        if (epi.accumulate) {
          float2 c = *reinterpret_cast<float2*>(dst);
          o.x += c.x; o.y += c.y;
        }
        *reinterpret_cast<float2*>(dst) = o;
      }
    }
  }
}

template <int BM, int BN, int TM, int TN, int STAGES>
constexpr size_t gemm_nn_smem() {
  return (size_t)STAGES * (BM * GEMM_APAD + GEMM_BK * BN) * sizeof(float);
}

// ---------------------------------------------------------------------------------------------
// row loaders for gemm_tn: address of row k (M or N contiguous values), or null -> zeros
// ---------------------------------------------------------------------------------------------
struct PlainK {
  const float* P;
  int ld;
  __device__ __forceinline__ const float* row(long long k) const { return P + (size_t)k * ld; }
};

// rows of channels-last X shifted by a 3x3 tap (weight gradient of the convolution)
struct ShiftedPixelK {
  const float* X;
  int T, F, C;
  int dt, df;
  __device__ __forceinline__ const float* row(long long k) const {
    int f = (int)(k % F);
    int t = (int)((k / F) % T);
    int tt = t + dt, ff = f + df;
    if (tt < 0 || tt >= T || ff < 0 || ff >= F) return nullptr;
    return X + ((size_t)k + (long long)dt * F + df) * C;
  }
};

// rows of a (B, T, ld) sequence shifted by one time step (h_{t-1} of a GRU direction)
struct ShiftedTimeK {
  const float* H;
  int T, ld;
  int dt;  // -1: forward direction, +1: reverse direction
  __device__ __forceinline__ const float* row(long long k) const {
    int t = (int)(k % T);
    int tt = t + dt;
    if (tt < 0 || tt >= T) return nullptr;
    return H + ((size_t)k + dt) * ld;
  }
};

template <int BM, int BN, class ALoad, class BLoad>
__global__ void __launch_bounds__(256) gemm_tn_kernel(ALoad aload, BLoad bload, float* __restrict__ C,
                                                      long long rs, long long cs, int M, int N,
                                                      long long K, long long k_per_cta) {
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int A_CH = GEMM_BK * BM / 4, B_CH = GEMM_BK * BN / 4;
  __shared__ __align__(16) float As[2][GEMM_BK * BM];
  __shared__ __align__(16) float Bs[2][GEMM_BK * BN];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const long long kbeg = (long long)blockIdx.z * k_per_cta;
  long long kend = kbeg + k_per_cta;
  if (kend > K) kend = K;
  if (kbeg >= kend) return;
  const int nk = (int)((kend - kbeg + GEMM_BK - 1) / GEMM_BK);

  auto load_stage = [&](int stage, int it) {
    long long kb = kbeg + (long long)it * GEMM_BK;
    for (int id = tid; id < A_CH; id += 256) {
      int k = id / (BM / 4), c4 = id % (BM / 4);
      long long kk = kb + k;
      const float* r = (kk < kend) ? aload.row(kk) : nullptr;
      cp_async16(&As[stage][k * BM + c4 * 4], r ? r + m0 + c4 * 4 : (const float*)C, r != nullptr);
    }
    for (int id = tid; id < B_CH; id += 256) {
      int k = id / (BN / 4), c4 = id % (BN / 4);
      long long kk = kb + k;
      const float* r = (kk < kend) ? bload.row(kk) : nullptr;
      cp_async16(&Bs[stage][k * BN + c4 * 4], r ? r + n0 + c4 * 4 : (const float*)C, r != nullptr);
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  load_stage(0, 0);
  cp_async_commit();
  for (int it = 0; it < nk; ++it) {
    if (it + 1 < nk) load_stage((it + 1) & 1, it + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* as = As[it & 1];
    const float* bs = Bs[it & 1];
#pragma unroll
    for (int k = 0; k < GEMM_BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = as[k * BM + ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = bs[k * BN + tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n < N) atomicAdd(C + m * rs + n * cs, acc[i][j]);
    }
  }
}

}  // namespace bsed
