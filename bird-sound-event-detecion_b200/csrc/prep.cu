// prep.cu -- one launch that turns the reference-layout parameters of one parameter set into the
// operand layouts the kernels consume:
//   conv  W[co][ci][tap]      -> Wp[tap][ci][co]   (forward implicit GEMM, B operand K-major)
//                             -> Wd[tap][co][ci] = W[co][ci][8-tap]   (data gradient = conv with flipped taps)
//   GLU   Wg[c'][c], BN g/b   -> WgT'[c][c'] = g[c] * Wg[c'][c],  b'[c'] = bg[c'] + sum_c Wg[c'][c] * b[c]
//                                (lin = Wg (g*xhat + b) + bg  evaluated on xhat directly)
//   GRU / Predictor matrices  -> transposed, concatenated copies
// All ops of a table run concurrently in one launch: no two ops may write the same element.
#include "launch.h"

namespace bsed {

__global__ void __launch_bounds__(256) prep_kernel(const __grid_constant__ PrepTable table) {
  const PrepOp& op = table.ops[blockIdx.x];
  const int tid = blockIdx.y * blockDim.x + threadIdx.x;
  const int nth = gridDim.y * blockDim.x;
  switch (op.type) {
    case PREP_CONV_PACK: {
      int Cout = op.d0, Cin = op.d1;
      int n = Cout * Cin * 9;
      for (int i = tid; i < n; i += nth) {  // i indexes dst [tap][ci][co]
        int co = i % Cout, ci = (i / Cout) % Cin, tap = i / (Cout * Cin);
        op.dst[i] = op.src[((size_t)co * Cin + ci) * 9 + tap];
      }
    } break;
    case PREP_CONV_PACK_FLIP: {
      int Cout = op.d0, Cin = op.d1;
      int n = Cout * Cin * 9;
      for (int i = tid; i < n; i += nth) {  // dst [tap][co][ci]
        int ci = i % Cin, co = (i / Cin) % Cout, tap = i / (Cout * Cin);
        op.dst[i] = op.src[((size_t)co * Cin + ci) * 9 + (8 - tap)];
      }
    } break;
    case PREP_CONV_KMAJOR: {  // dst [co][tap][ci]  (tensor-core forward: B operand [N = co][K = tap*Cin + ci])
      int Cout = op.d0, Cin = op.d1;
      int n = Cout * Cin * 9;
      for (int i = tid; i < n; i += nth) {
        int ci = i % Cin, tap = (i / Cin) % 9, co = i / (Cin * 9);
        op.dst[i] = op.src[((size_t)co * Cin + ci) * 9 + tap];
      }
    } break;
    case PREP_CONV_KMAJOR_FLIP: {  // dst [ci][tap][co] = W[co][ci][8-tap]  (data gradient: N = ci, K = tap*Cout + co)
      int Cout = op.d0, Cin = op.d1;
      int n = Cout * Cin * 9;
      for (int i = tid; i < n; i += nth) {
        int co = i % Cout, tap = (i / Cout) % 9, ci = i / (Cout * 9);
        op.dst[i] = op.src[((size_t)co * Cin + ci) * 9 + (8 - tap)];
      }
    } break;
    case PREP_GLU_FOLD: {
      int C = op.d0;
      if (op.d1) {
        // K-major folded matrix for the tensor-core GEMM, block-diagonal over `pack` pixels per row:
        // dst[(p*C + c')][(q*C + c)] = (p == q) * gamma[c] * Wg[c'][c]
        const int pack = op.d2 > 0 ? op.d2 : 1, CP = C * pack;
        for (int i = tid; i < CP * CP; i += nth) {
          int col = i % CP, rowi = i / CP;
          int p = rowi / C, cp = rowi % C, q = col / C, c = col % C;
          op.dst[i] = p == q ? op.aux0[c] * op.src[(size_t)cp * C + c] : 0.f;
        }
        for (int j = tid; j < CP; j += nth) {
          int cp = j % C;
          float a = op.aux2[cp];
          for (int c = 0; c < C; ++c) a = fmaf(op.src[(size_t)cp * C + c], op.aux1[c], a);
          op.dst2[j] = a;
        }
      } else {
        for (int i = tid; i < C * C; i += nth) {  // dst [c][c']
          int cp = i % C, c = i / C;
          op.dst[i] = op.aux0[c] * op.src[(size_t)cp * C + c];
        }
        for (int cp = tid; cp < C; cp += nth) {
          float a = op.aux2[cp];
          for (int c = 0; c < C; ++c) a = fmaf(op.src[(size_t)cp * C + c], op.aux1[c], a);
          op.dst2[cp] = a;
        }
      }
    } break;
    case PREP_CONV_PAIR: {
      // 16-input-channel block viewed two pixels per row: K-major weights of the equivalent 3x3 convolution over
      // pixel PAIRS, dst [(par, co)][tap' = (dt, dg)][(p', ci)] = W[co][ci][dt][df], df = 2*dg + p' - par (0 if |df| > 1);
      // dst2 [(par, co)] = bias[co]
      const int Cout = op.d0, Cin = op.d1;
      const int K = 9 * 2 * Cin, n = 2 * Cout * K;
      for (int i = tid; i < n; i += nth) {
        const int k = i % K, rowi = i / K;
        const int par = rowi / Cout, co = rowi % Cout;
        const int tap = k / (2 * Cin), pc = k % (2 * Cin);
        const int pp = pc / Cin, ci = pc % Cin;
        const int dt = tap / 3, dg = tap % 3 - 1;
        const int df = 2 * dg + pp - par;
        op.dst[i] = (df >= -1 && df <= 1) ? op.src[((size_t)co * Cin + ci) * 9 + dt * 3 + (df + 1)] : 0.f;
      }
      for (int j = tid; j < 2 * Cout; j += nth) op.dst2[j] = op.aux0[j % Cout];
    } break;
    case PREP_GATE_TAB: {  // dst [gamma x pack | beta x pack], C = d0, pack = d1
      const int C = op.d0, CP = C * op.d1;
      for (int j = tid; j < CP; j += nth) {
        op.dst[j] = op.src[j % C];
        op.dst[CP + j] = op.aux0[j % C];
      }
    } break;
    case PREP_TRANSPOSE_BD: {  // dst[(p*C + c)][(q*C + c')] = (p == q) * src[c'][c]   (src [C][C], pack = d1)
      const int C = op.d0, pack = op.d1, CP = C * pack;
      for (int i = tid; i < CP * CP; i += nth) {
        int col = i % CP, rowi = i / CP;
        int p = rowi / C, c = rowi % C, q = col / C, cp = col % C;
        op.dst[i] = p == q ? op.src[(size_t)cp * C + c] : 0.f;
      }
    } break;
    case PREP_TRANSPOSE: {
      int R = op.d0, Cc = op.d1, ldd = op.d2, off = op.d3;
      for (int i = tid; i < R * Cc; i += nth) {  // src [r][c] -> dst[c*ldd + off + r]
        int r = i % R, c = i / R;
        op.dst[(size_t)c * ldd + off + r] = op.src[(size_t)r * Cc + c];
      }
    } break;
    case PREP_COPY: {
      for (int i = tid; i < op.d0; i += nth) op.dst[i] = op.src[i];
    } break;
    case PREP_ZERO: {
      for (int i = tid; i < op.d0; i += nth) op.dst[i] = 0.f;
    } break;
    case PREP_ZERO_COLS: {  // dst[r * ld + col0 + c] = 0 for r < rows, c < ncols
      int rows = op.d0, ld = op.d1, col0 = op.d2, nc = op.d3;
      for (int i = tid; i < rows * nc; i += nth) op.dst[(size_t)(i / nc) * ld + col0 + i % nc] = 0.f;
    } break;
  }
}

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__global__ void __launch_bounds__(256) split_kernel(const __grid_constant__ SplitTable table, float* __restrict__ packed,
                                                    long long lo_offset) {
  float* w = packed + table.off[blockIdx.x];
  const long long n = table.len[blockIdx.x];
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.y * blockDim.x) {
    const float x = w[i], hi = tf32_rna(x);
    w[i] = hi;
    w[i + lo_offset] = tf32_rna(x - hi);   // 13 significant bits -> nearest tf32 (the tensor core would truncate)
  }
}

int run_split(const SplitTable& table, float* packed, long long lo_offset, cudaStream_t st) {
  if (table.n == 0) return BSED_OK;
  dim3 grid(table.n, 32);   // the big ranges (GRU / conv matrices of 10^5 floats) need more than a handful of CTAs
  split_kernel<<<grid, 256, 0, st>>>(table, packed, lo_offset);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

int split_hi_lo(float* w, float* lo, long long n, cudaStream_t st) {
  SplitTable tb;
  tb.n = 1;
  tb.off[0] = 0;
  tb.len[0] = n;
  return run_split(tb, w, lo - w, st);
}

int run_prep(const PrepTable& table, cudaStream_t st) {
  if (table.n == 0) return BSED_OK;
  dim3 grid(table.n, 32);   // the big ranges (GRU / conv matrices of 10^5 floats) need more than a handful of CTAs
  prep_kernel<<<grid, 256, 0, st>>>(table);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace bsed
