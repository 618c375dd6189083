// gemm.cu -- launchers for the fp32 SIMT GEMM family (plain, implicit-GEMM conv, split-K reductions).
#include "gemm.cuh"
#include "launch.h"

namespace bsed {

template <int BM, int BN, int TM, int TN, class ALoad>
static int launch_nn_t(ALoad aload, const float* Bm, int ldb, int M, int N, int K, NNEpilogue epi,
                       cudaStream_t st) {
  constexpr int STAGES = 3;
  constexpr size_t smem = gemm_nn_smem<BM, BN, TM, TN, STAGES>();
  auto kern = gemm_nn_kernel<BM, BN, TM, TN, STAGES, ALoad>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    BSED_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid(ceil_div(M, BM), N / BN);
  kern<<<grid, 256, smem, st>>>(aload, Bm, ldb, M, N, K, epi);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

template <class ALoad>
static int launch_nn(ALoad aload, const float* Bm, int ldb, int M, int N, int K, NNEpilogue epi,
                     cudaStream_t st) {
  BSED_REQUIRE(K % GEMM_BK == 0 && K > 0, "gemm_nn: K=%d must be a positive multiple of 16", K);
  BSED_REQUIRE(N % 16 == 0 && N > 0, "gemm_nn: N=%d must be a positive multiple of 16", N);
  BSED_REQUIRE(M > 0, "gemm_nn: M=%d", M);
  if (N % 128 == 0) return launch_nn_t<128, 128, 8, 8>(aload, Bm, ldb, M, N, K, epi, st);
  if (N % 64 == 0) return launch_nn_t<128, 64, 8, 4>(aload, Bm, ldb, M, N, K, epi, st);
  if (N % 32 == 0) return launch_nn_t<256, 32, 8, 4>(aload, Bm, ldb, M, N, K, epi, st);
  return launch_nn_t<256, 16, 8, 2>(aload, Bm, ldb, M, N, K, epi, st);
}

int gemm_nn(const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N, int K,
            const float* bias, int accumulate, cudaStream_t st) {
  BSED_REQUIRE(lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0, "gemm_nn: leading dims must be multiples of 4");
  PlainRows a{A, lda, M};
  NNEpilogue epi{C, ldc, bias, accumulate};
  ProfScope prof(PROF_GEMM, 2.0 * M * N * K, 4.0 * ((double)M * K + (double)K * N + (double)M * N), st);
  return launch_nn(a, Bm, ldb, M, N, K, epi, st);
}

int conv3x3_nn(const float* X, const float* Wp, float* Y, int B, int T, int F, int Cin, int Cout,
               const float* bias, int accumulate, cudaStream_t st) {
  BSED_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0, "conv3x3: Cin=%d Cout=%d must be multiples of 16", Cin, Cout);
  long long M = (long long)B * T * F;
  BSED_REQUIRE(M < (1ll << 31), "conv3x3: too many pixels");
  ConvRows a{X, T, F, Cin, (int)M, Cin / GEMM_BK};
  NNEpilogue epi{Y, Cout, bias, accumulate};
  ProfScope prof(PROF_CONV, 2.0 * M * Cout * 9.0 * Cin, 4.0 * ((double)M * Cin + (double)M * Cout + 9.0 * Cin * Cout), st);
  return launch_nn(a, Wp, Cout, (int)M, Cout, 9 * Cin, epi, st);
}

template <int BM, int BN, class ALoad, class BLoad>
static int launch_tn_t(ALoad a, BLoad b, float* C, long long rs, long long cs, int M, int N, long long K,
                       int target_ctas, cudaStream_t st) {
  int gx = M / BM, gy = N / BN;
  long long chunks = (K + GEMM_BK - 1) / GEMM_BK;
  int splits = target_ctas / (gx * gy);
  if (splits < 1) splits = 1;
  if (splits > chunks) splits = (int)chunks;
  long long k_per = ((chunks + splits - 1) / splits) * GEMM_BK;
  splits = (int)((K + k_per - 1) / k_per);
  dim3 grid(gx, gy, splits);
  gemm_tn_kernel<BM, BN, ALoad, BLoad><<<grid, 256, 0, st>>>(a, b, C, rs, cs, M, N, K, k_per);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

static inline int pick_tile(int n) { return n % 128 == 0 ? 128 : n % 64 == 0 ? 64 : n % 32 == 0 ? 32 : 16; }

template <class ALoad, class BLoad>
static int launch_tn(ALoad a, BLoad b, float* C, long long rs, long long cs, int M, int N, long long K,
                     int target_ctas, cudaStream_t st) {
  BSED_REQUIRE(M % 16 == 0 && N % 16 == 0 && M > 0 && N > 0 && K > 0, "gemm_tn: M=%d N=%d K=%lld", M, N, K);
  int bm = pick_tile(M), bn = pick_tile(N);
#define BSED_TN_CASE(BM_, BN_) \
  if (bm == BM_ && bn == BN_) return launch_tn_t<BM_, BN_>(a, b, C, rs, cs, M, N, K, target_ctas, st);
  BSED_TN_CASE(128, 128) BSED_TN_CASE(128, 64) BSED_TN_CASE(128, 32) BSED_TN_CASE(128, 16)
  BSED_TN_CASE(64, 128) BSED_TN_CASE(64, 64) BSED_TN_CASE(64, 32) BSED_TN_CASE(64, 16)
  BSED_TN_CASE(32, 128) BSED_TN_CASE(32, 64) BSED_TN_CASE(32, 32) BSED_TN_CASE(32, 16)
  BSED_TN_CASE(16, 128) BSED_TN_CASE(16, 64) BSED_TN_CASE(16, 32) BSED_TN_CASE(16, 16)
#undef BSED_TN_CASE
  bsed_set_error("gemm_tn: no tile for M=%d N=%d", M, N);
  return BSED_E_INVALID;
}

int gemm_tn(const float* A, int lda, const float* Bm, int ldb, float* C, long long rs, long long cs, int M,
            int N, long long K, int target_ctas, cudaStream_t st) {
  PlainK a{A, lda}, b{Bm, ldb};
  ProfScope prof(PROF_GEMM_TN, 2.0 * M * N * K, 4.0 * ((double)K * M + (double)K * N + (double)M * N), st);
  return launch_tn(a, b, C, rs, cs, M, N, K, target_ctas, st);
}

// dW[co][ci][tap] += sum_pixels dY[p][co] * X[p + tap][ci]   (dW in the reference's OIHW layout)
int conv3x3_wgrad(const float* X, const float* dY, float* dW, int B, int T, int F, int Cin, int Cout,
                  int target_ctas, cudaStream_t st) {
  long long K = (long long)B * T * F;
  ProfScope prof(PROF_WGRAD, 2.0 * K * Cout * 9.0 * Cin, 4.0 * ((double)K * Cin + (double)K * Cout + 9.0 * Cin * Cout), st);
  for (int tap = 0; tap < 9; ++tap) {
    PlainK a{dY, Cout};
    ShiftedPixelK b{X, T, F, Cin, tap / 3 - 1, tap % 3 - 1};
    BSED_TRY(launch_tn(a, b, dW + tap, (long long)Cin * 9, 9, Cout, Cin, K, target_ctas / 9 + 1, st));
  }
  return BSED_OK;
}

// dWhh[j][i] += sum_{b,t} dG[b][t][j] * h_{t-1}[b][t][i]  for one direction
int gru_whh_grad(const float* dG, int ldg, const float* H, int ldh, int dt, float* dW, int T, long long BT,
                 int target_ctas, cudaStream_t st) {
  PlainK a{dG, ldg};
  ShiftedTimeK b{H, T, ldh, dt};
  return launch_tn(a, b, dW, 128, 1, 384, 128, BT, target_ctas, st);
}

}  // namespace bsed
