// cnn_ops.cu -- channels-last kernels of the gated CNN block around the GEMMs:
//   conv0 (Cin = 1) forward / weight gradient, BatchNorm statistics / normalise / backward,
//   GLU gate + dropout + average pool forward / backward.
// Reference arithmetic: src/models/CNN.py:5-16 (GLU), :43-67 (block order), BatchNorm2d(eps=1e-3,
// momentum=.99).  BN is applied as xhat = (y - mean) * rstd with gamma/beta folded into the GLU
// linear (prep.cu) and into the sigmoid gate here.
#include "launch.h"

namespace bsed {

__device__ __forceinline__ int group_of(const Groups& g, int clip) {
  int r = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroups; ++i)
    if (i < g.n && clip >= g.first[i]) r = i;
  return r;
}

// ---------------------------------------------------------------------------------------------
// conv0: x [B][T][F] -> y [B][T][F][Cout], weights in the reference's (Cout,1,3,3) layout.
// A thread owns one (column, channel quad) and walks down a strip of rows with the 3 x 3 input window sliding through
// registers: three new loads per output pixel instead of nine, no per-pixel index arithmetic, the next row's loads
// issued before the current row's FMAs.  Lanes = 8 consecutive columns x 4 quads (Cout = 16), so a warp writes 512
// contiguous bytes per row.  With stats != nullptr the per-channel sum / sum of squares of the output (BatchNorm batch
// statistics) are accumulated on the way: stats[(group * Cout + c) * 2 + {0, 1}].
// grid (column blocks x row strips, clip); block = 256 threads = (256 / nq) columns x nq quads.
// ---------------------------------------------------------------------------------------------
constexpr int kConv0Rows = 64;   // rows per strip

__global__ void __launch_bounds__(256) conv0_fwd_kernel(const float* __restrict__ x, Groups g, FloatPtrs w,
                                                        FloatPtrs bias, float* __restrict__ y, int T, int F,
                                                        int Cout, int col_blocks, double* __restrict__ stats) {
  const int clip = blockIdx.y;
  const int grp = group_of(g, clip);
  const int nq = Cout / 4;
  const int q = threadIdx.x % nq;
  const int cols_per_cta = 256 / nq;
  const int cb = blockIdx.x % col_blocks, strip = blockIdx.x / col_blocks;
  const int col = cb * cols_per_cta + threadIdx.x / nq;
  const int t0 = strip * kConv0Rows;
  const int t1 = min(T, t0 + kConv0Rows);
  float wr[4][9], br[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    br[j] = bias.p[grp][q * 4 + j];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) wr[j][tap] = w.p[grp][(q * 4 + j) * 9 + tap];
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < F) {
    const float* xc = x + (size_t)clip * T * F;
    const bool has_l = col > 0, has_r = col + 1 < F;
    auto load_row = [&](int t, float (&r)[3]) {
      if (t >= 0 && t < T) {
        const float* p = xc + (size_t)t * F + col;
        r[0] = has_l ? __ldg(p - 1) : 0.f;
        r[1] = __ldg(p);
        r[2] = has_r ? __ldg(p + 1) : 0.f;
      } else {
        r[0] = r[1] = r[2] = 0.f;
      }
    };
    // rows t-1 .. t+4 of the window: the loads of four output rows are issued together, ahead of their FMAs
    float rw[6][3];
    load_row(t0 - 1, rw[0]);
    load_row(t0, rw[1]);
    float* yp = y + (((size_t)clip * T + t0) * F + col) * Cout + q * 4;
    for (int t = t0; t < t1; t += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) load_row(t + 1 + u, rw[2 + u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (t + u < t1) {
          float o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float a = br[j];
#pragma unroll
            for (int k = 0; k < 3; ++k) a = fmaf(rw[u][k], wr[j][k], a);        // taps in the reference's (kh, kw) order
#pragma unroll
            for (int k = 0; k < 3; ++k) a = fmaf(rw[u + 1][k], wr[j][3 + k], a);
#pragma unroll
            for (int k = 0; k < 3; ++k) a = fmaf(rw[u + 2][k], wr[j][6 + k], a);
            o[j] = a;
            s1[j] += a;
            s2[j] = fmaf(a, a, s2[j]);
          }
          *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
          yp += (size_t)F * Cout;
        }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        rw[0][k] = rw[4][k];
        rw[1][k] = rw[5][k];
      }
    }
  }
  if (!stats) return;
  // threads with equal q inside a warp (lanes l, l + nq, ...) first, then the 8 warps through shared memory: the serial
  // 64-column sum this replaces was a third of the CTA's time
  __shared__ float red[2][8][128];   // [sum | sum of squares][warp][channel], Cout <= 128
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    for (int o = 16; o >= nq; o >>= 1) {
      s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
      s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], o);
    }
  }
  if (lane < nq) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[0][warp][lane * 4 + j] = s1[j];
      red[1][warp][lane * 4 + j] = s2[j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * Cout; i += 256) {
    const int k = i / Cout, c = i % Cout;
    float tsum = 0.f;
#pragma unroll
    for (int wdx = 0; wdx < 8; ++wdx) tsum += red[k][wdx][c];
    atomicAdd(stats + ((size_t)grp * Cout + c) * 2 + k, (double)tsum);
  }
}

int conv0_fwd(const float* x, const Groups& g, const FloatPtrs& w, const FloatPtrs& bias, float* y, int T,
              int F, int Cout, double* stats, int num_sms, cudaStream_t st) {
  BSED_REQUIRE(Cout % 4 == 0 && Cout <= 128 && 32 % (Cout / 4) == 0, "conv0: Cout=%d", Cout);
  (void)num_sms;
  int B = g.first[g.n - 1] + g.count[g.n - 1];
  const int cols_per_cta = 256 / (Cout / 4);
  const int col_blocks = ceil_div(F, cols_per_cta);
  dim3 grid(col_blocks * ceil_div(T, kConv0Rows), B);
  conv0_fwd_kernel<<<grid, 256, 0, st>>>(x, g, w, bias, y, T, F, Cout, col_blocks, stats);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// dW[co][tap] += sum_p dY[p][co] * x[p + tap]: same strip walk as the forward (thread = column x channel quad, sliding
// 3 x 3 window of x), 36 accumulators per thread, reduced over the CTA's columns through shared memory.
__global__ void __launch_bounds__(256) conv0_wgrad_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ dY, float* dW, int first_clip,
                                                          int T, int F, int Cout, int col_blocks) {
  const int clip = first_clip + blockIdx.y;
  const int nq = Cout / 4;
  const int q = threadIdx.x % nq;
  const int cols_per_cta = 256 / nq;
  const int cb = blockIdx.x % col_blocks, strip = blockIdx.x / col_blocks;
  const int col = cb * cols_per_cta + threadIdx.x / nq;
  const int t0 = strip * kConv0Rows;
  const int t1 = min(T, t0 + kConv0Rows);
  float acc[4][9];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) acc[j][tap] = 0.f;
  if (col < F) {
    const float* xc = x + (size_t)clip * T * F;
    const bool has_l = col > 0, has_r = col + 1 < F;
    auto load_row = [&](int t, float (&r)[3]) {
      if (t >= 0 && t < T) {
        const float* p = xc + (size_t)t * F + col;
        r[0] = has_l ? __ldg(p - 1) : 0.f;
        r[1] = __ldg(p);
        r[2] = has_r ? __ldg(p + 1) : 0.f;
      } else {
        r[0] = r[1] = r[2] = 0.f;
      }
    };
    constexpr int R = 8;   // rows of gradient and of input in flight before the FMAs (the kernel streams dY: bytes in flight)
    float rw[R + 2][3];
    load_row(t0 - 1, rw[0]);
    load_row(t0, rw[1]);
    const float* dp = dY + (((size_t)clip * T + t0) * F + col) * Cout + q * 4;
    for (int t = t0; t < t1; t += R) {
      float4 d[R];
#pragma unroll
      for (int u = 0; u < R; ++u) {
        load_row(t + 1 + u, rw[2 + u]);
        d[u] = t + u < t1 ? *reinterpret_cast<const float4*>(dp + (size_t)u * F * Cout) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      dp += (size_t)R * F * Cout;
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            acc[j][k] = fmaf(dv[j], rw[u][k], acc[j][k]);
            acc[j][3 + k] = fmaf(dv[j], rw[u + 1][k], acc[j][3 + k]);
            acc[j][6 + k] = fmaf(dv[j], rw[u + 2][k], acc[j][6 + k]);
          }
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        rw[0][k] = rw[R][k];
        rw[1][k] = rw[R + 1][k];
      }
    }
  }
  // reduce over the threads with equal q: lanes l, l + nq, ... inside the warp, then the 8 warps through smem
  __shared__ float red[8][32 * 9];   // [warp][(q * 4 + j) * 9 + tap], Cout <= 32
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      float v = acc[j][tap];
      for (int o = 16; o >= nq; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[j][tap] = v;
    }
  if (lane < nq) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) red[warp][(lane * 4 + j) * 9 + tap] = acc[j][tap];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * 9; i += blockDim.x) {
    float s = 0.f;
    for (int wdx = 0; wdx < 8; ++wdx) s += red[wdx][i];
    atomicAdd(dW + i, s);
  }
}

int conv0_wgrad(const float* x, const float* dY, float* dW, int first_clip, int n_clips, int T, int F,
                int Cout, int num_sms, cudaStream_t st) {
  BSED_REQUIRE(Cout % 4 == 0 && Cout <= 32 && (32 % (Cout / 4)) == 0 && (256 % (Cout / 4)) == 0,
               "conv0_wgrad: Cout=%d unsupported", Cout);
  (void)num_sms;
  const int cols_per_cta = 256 / (Cout / 4);
  const int col_blocks = ceil_div(F, cols_per_cta);
  dim3 grid(col_blocks * ceil_div(T, kConv0Rows), n_clips);
  conv0_wgrad_kernel<<<grid, 256, 0, st>>>(x, dY, dW, first_clip, T, F, Cout, col_blocks);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// per-channel column reductions over the rows of each clip, accumulated per group in double
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        Groups g, long long rows_per_clip, int C,
                                                        long long rows_per_cta, double* out) {
  const int clip_idx = blockIdx.y;  // index among the clips covered by the groups
  // map blockIdx.y -> absolute clip: groups are contiguous, first[0] may be > 0
  int clip = -1, grp = 0;
  {
    int rem = clip_idx;
    for (int i = 0; i < g.n; ++i) {
      if (rem < g.count[i]) {
        clip = g.first[i] + rem;
        grp = i;
        break;
      }
      rem -= g.count[i];
    }
  }
  if (clip < 0) return;
  const int nq = C / 4;
  const int q = threadIdx.x % nq;
  const int r0 = threadIdx.x / nq;
  const int rstep = blockDim.x / nq;
  long long rbeg = (long long)blockIdx.x * rows_per_cta;
  long long rend = rbeg + rows_per_cta;
  if (rend > rows_per_clip) rend = rows_per_clip;
  float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
  if (r0 < rstep) {
    constexpr int U = 4;   // rows in flight per thread
    const float* abase = a + (size_t)clip * rows_per_clip * C + q * 4;
    const float* bbase = MODE == 1 ? b + (size_t)clip * rows_per_clip * C + q * 4 : nullptr;
    for (long long r = rbeg + r0; r < rend; r += (long long)U * rstep) {
      float4 av[U], bv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        long long rr = r + (long long)u * rstep;
        av[u] = rr < rend ? *reinterpret_cast<const float4*>(abase + (size_t)rr * C) : make_float4(0, 0, 0, 0);
        if (MODE == 1) bv[u] = rr < rend ? *reinterpret_cast<const float4*>(bbase + (size_t)rr * C) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        s1.x += av[u].x; s1.y += av[u].y; s1.z += av[u].z; s1.w += av[u].w;
        if (MODE == 0) {
          s2.x = fmaf(av[u].x, av[u].x, s2.x); s2.y = fmaf(av[u].y, av[u].y, s2.y);
          s2.z = fmaf(av[u].z, av[u].z, s2.z); s2.w = fmaf(av[u].w, av[u].w, s2.w);
        } else if (MODE == 1) {
          s2.x = fmaf(av[u].x, bv[u].x, s2.x); s2.y = fmaf(av[u].y, bv[u].y, s2.y);
          s2.z = fmaf(av[u].z, bv[u].z, s2.z); s2.w = fmaf(av[u].w, bv[u].w, s2.w);
        }
      }
    }
  }
  __shared__ float red[2][256][4];
  red[0][threadIdx.x][0] = s1.x; red[0][threadIdx.x][1] = s1.y; red[0][threadIdx.x][2] = s1.z; red[0][threadIdx.x][3] = s1.w;
  red[1][threadIdx.x][0] = s2.x; red[1][threadIdx.x][1] = s2.y; red[1][threadIdx.x][2] = s2.z; red[1][threadIdx.x][3] = s2.w;
  __syncthreads();
  // thread c (< C) sums the partials of its channel
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    int cq = c / 4, cj = c % 4;
    float t1 = 0.f, t2 = 0.f;
    for (int r = 0; r < rstep; ++r) {
      t1 += red[0][r * nq + cq][cj];
      t2 += red[1][r * nq + cq][cj];
    }
    atomicAdd(out + ((size_t)grp * C + c) * 2 + 0, (double)t1);
    if (MODE != 2) atomicAdd(out + ((size_t)grp * C + c) * 2 + 1, (double)t2);
  }
}

int col_stats(const float* a, const float* b, int mode, const Groups& g, long long rows_per_clip, int C,
              double* out, int num_sms, cudaStream_t st) {
  BSED_REQUIRE(C % 4 == 0 && C / 4 <= 256, "col_stats: C=%d", C);
  int nclips = 0;
  for (int i = 0; i < g.n; ++i) nclips += g.count[i];
  if (nclips == 0) return BSED_OK;
  // aim for ~8 CTAs per SM overall, at least 16 rows per CTA (the short, wide matrices of the GRU backward -- 7512 x 768 --
  // are latency-bound: with 64 rows per CTA every thread walked 32 dependent load rounds, 21 us per launch)
  long long want = (long long)num_sms * 8 / nclips + 1;
  long long rows_per_cta = (rows_per_clip + want - 1) / want;
  if (rows_per_cta < 16) rows_per_cta = 16;
  dim3 grid(ceil_div(rows_per_clip, rows_per_cta), nclips);
  int threads = (C / 4) * (256 / (C / 4));  // a multiple of the quads per row, <= 256
  if (mode == 0) col_stats_kernel<0><<<grid, threads, 0, st>>>(a, b, g, rows_per_clip, C, rows_per_cta, out);
  else if (mode == 1) col_stats_kernel<1><<<grid, threads, 0, st>>>(a, b, g, rows_per_clip, C, rows_per_cta, out);
  else col_stats_kernel<2><<<grid, threads, 0, st>>>(a, b, g, rows_per_clip, C, rows_per_cta, out);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

__global__ void add_double_to_float_kernel(const double* src, int stride, float* dst, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += (float)src[(size_t)i * stride];
}
int add_double_to_float(const double* src, int stride, float* dst, int n, cudaStream_t st) {
  add_double_to_float_kernel<<<ceil_div(n, 256), 256, 0, st>>>(src, stride, dst, n);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

int col_sum_to(const float* a, long long rows, int C, float* out, double* scratch, int num_sms,
               cudaStream_t st) {
  // treat the whole matrix as one "clip" of one group
  Groups g;
  g.n = 1;
  g.first[0] = 0;
  g.count[0] = 1;
  BSED_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
  BSED_TRY(col_stats(a, nullptr, 2, g, rows, C, scratch, num_sms, st));
  return add_double_to_float(scratch, 2, out, C, st);
}

// ---------------------------------------------------------------------------------------------
// BatchNorm finalise / eval prepare / normalise
// ---------------------------------------------------------------------------------------------
struct RunPtrs {
  float* mean[kMaxGroups];
  float* var[kMaxGroups];
  int64_t* nbt[kMaxGroups];
};

__global__ void bn_finalize_train_kernel(const double* __restrict__ stats, Groups g, long long rows_per_clip,
                                         int C, float eps, float momentum, BNPtrs bn, RunPtrs run) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int i = 0; i < g.n; ++i) {
    double n = (double)g.count[i] * (double)rows_per_clip;
    double s = stats[((size_t)i * C + c) * 2], ss = stats[((size_t)i * C + c) * 2 + 1];
    double mean = s / n;
    double var = ss / n - mean * mean;
    if (var < 0) var = 0;
    bn.mean[i][c] = (float)mean;
    bn.rstd[i][c] = (float)(1.0 / sqrt(var + (double)eps));
    if (run.mean[i]) {
      double unb = n > 1 ? var * n / (n - 1) : var;
      run.mean[i][c] = (1.f - momentum) * run.mean[i][c] + momentum * (float)mean;
      run.var[i][c] = (1.f - momentum) * run.var[i][c] + momentum * (float)unb;
    }
    if (c == 0 && run.nbt[i]) *run.nbt[i] += 1;
  }
}

int bn_finalize_train(const double* stats, const Groups& g, long long rows_per_clip, int C, float eps,
                      float momentum, const BNPtrs& bn, float* const* run_mean, float* const* run_var,
                      int64_t* const* nbt, cudaStream_t st) {
  RunPtrs run;
  for (int i = 0; i < kMaxGroups; ++i) {
    run.mean[i] = i < g.n ? run_mean[i] : nullptr;
    run.var[i] = i < g.n ? run_var[i] : nullptr;
    run.nbt[i] = i < g.n ? nbt[i] : nullptr;
  }
  bn_finalize_train_kernel<<<ceil_div(C, 128), 128, 0, st>>>(stats, g, rows_per_clip, C, eps, momentum, bn, run);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// Running-statistics update of a BatchNorm that one forward applies twice (bn_fcn of CNN_FPN, src/models/CNN_FPN.py:86-96):
// per group (= reference model call) the first application's batch statistics are blended in, then the second's, in the
// reference's order; num_batches_tracked += 2.  statsA / statsB: double [group][C][2] = (sum, sum of squares).
__global__ void bn_running_update2_kernel(const double* __restrict__ statsA, long long rowsA,
                                          const double* __restrict__ statsB, long long rowsB, Groups g, int C,
                                          float momentum, RunPtrs run) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int i = 0; i < g.n; ++i) {
    if (!run.mean[i]) continue;
    for (int a = 0; a < 2; ++a) {
      const double* stats = a == 0 ? statsA : statsB;
      const double n = (double)g.count[i] * (double)(a == 0 ? rowsA : rowsB);
      const double s = stats[((size_t)i * C + c) * 2], ss = stats[((size_t)i * C + c) * 2 + 1];
      const double mean = s / n;
      double var = ss / n - mean * mean;
      if (var < 0) var = 0;
      const double unb = n > 1 ? var * n / (n - 1) : var;
      run.mean[i][c] = (1.f - momentum) * run.mean[i][c] + momentum * (float)mean;
      run.var[i][c] = (1.f - momentum) * run.var[i][c] + momentum * (float)unb;
    }
    if (c == 0 && run.nbt[i]) *run.nbt[i] += 2;
  }
}

int bn_running_update2(const double* statsA, long long rowsA, const double* statsB, long long rowsB, const Groups& g, int C,
                       float momentum, float* const* run_mean, float* const* run_var, int64_t* const* nbt,
                       cudaStream_t st) {
  RunPtrs run;
  for (int i = 0; i < kMaxGroups; ++i) {
    run.mean[i] = i < g.n ? run_mean[i] : nullptr;
    run.var[i] = i < g.n ? run_var[i] : nullptr;
    run.nbt[i] = i < g.n ? nbt[i] : nullptr;
  }
  bn_running_update2_kernel<<<ceil_div(C, 128), 128, 0, st>>>(statsA, rowsA, statsB, rowsB, g, C, momentum, run);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

__global__ void bn_prepare_eval_kernel(Groups g, int C, float eps, BNPtrs bn, RunPtrs run) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  for (int i = 0; i < g.n; ++i) {
    bn.mean[i][c] = run.mean[i][c];
    bn.rstd[i][c] = 1.0f / sqrtf(run.var[i][c] + eps);
  }
}

int bn_prepare_eval(const Groups& g, int C, float eps, const BNPtrs& bn, float* const* run_mean,
                    float* const* run_var, cudaStream_t st) {
  RunPtrs run;
  for (int i = 0; i < kMaxGroups; ++i) {
    run.mean[i] = i < g.n ? run_mean[i] : nullptr;
    run.var[i] = i < g.n ? run_var[i] : nullptr;
    run.nbt[i] = nullptr;
  }
  bn_prepare_eval_kernel<<<ceil_div(C, 128), 128, 0, st>>>(g, C, eps, bn, run);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// in place y -> xhat = (y - mean) * rstd ; grid (chunks, clip); kEltU float4 per thread, loads issued first
constexpr int kEltU = 4;
__global__ void __launch_bounds__(256) bn_normalize_kernel(float* __restrict__ y, Groups g, long long elems_per_clip,
                                                           int C, BNPtrs bn) {
  const int clip = g.first[0] + blockIdx.y;
  const int grp = group_of(g, clip);
  const long long n4 = elems_per_clip / 4;
  const long long base = (long long)blockIdx.x * (256 * kEltU) + threadIdx.x;
  float4* p = reinterpret_cast<float4*>(y + (size_t)clip * elems_per_clip);
  float4 v[kEltU];
#pragma unroll
  for (int u = 0; u < kEltU; ++u) {
    long long i4 = base + u * 256;
    if (i4 < n4) v[u] = p[i4];
  }
#pragma unroll
  for (int u = 0; u < kEltU; ++u) {
    long long i4 = base + u * 256;
    if (i4 >= n4) continue;
    int c = (int)((i4 * 4) % C);
    const float4 mu = *reinterpret_cast<const float4*>(bn.mean[grp] + c);
    const float4 rs = *reinterpret_cast<const float4*>(bn.rstd[grp] + c);
    v[u].x = (v[u].x - mu.x) * rs.x;
    v[u].y = (v[u].y - mu.y) * rs.y;
    v[u].z = (v[u].z - mu.z) * rs.z;
    v[u].w = (v[u].w - mu.w) * rs.w;
    p[i4] = v[u];
  }
}

static int total_clips(const Groups& g) {
  int n = 0;
  for (int i = 0; i < g.n; ++i) n += g.count[i];
  return n;
}

int bn_normalize(float* y, const Groups& g, long long rows_per_clip, int C, const BNPtrs& bn,
                 cudaStream_t st) {
  long long elems = rows_per_clip * C;
  dim3 grid(ceil_div(elems / 4, 256 * kEltU), total_clips(g));
  bn_normalize_kernel<<<grid, 256, 0, st>>>(y, g, elems, C, bn);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// GLU gate + dropout + average pool
//   out = lin * sigmoid(gamma * xhat + beta) * keep / (1 - p), averaged over the pt x pf window
// ---------------------------------------------------------------------------------------------
template <int PT, int PF>
__global__ void __launch_bounds__(256) glu_gate_pool_fwd_kernel(const float* __restrict__ xhat,
                                                                const float* __restrict__ lin,
                                                                float* __restrict__ pooled, Groups g, BNPtrs bn,
                                                                int T, int F, int C, int pt_rt, int pf_rt, int To, int Fo,
                                                                DropKey dkey, uint32_t thresh, float inv_keep) {
  const uint32_t key = dkey.get();
  const int pt = PT ? PT : pt_rt, pf = PF ? PF : pf_rt;
  const int clip = g.first[0] + blockIdx.y;
  const int grp = group_of(g, clip);
  const int nq = C / 4;
  long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= (long long)To * Fo * nq) return;
  int q = (int)(id % nq);
  int opix = (int)(id / nq);
  int fo = opix % Fo, to = opix / Fo;
  int c = q * 4;
  const float4 ga4 = *reinterpret_cast<const float4*>(bn.gamma[grp] + c);
  const float4 be4 = *reinterpret_cast<const float4*>(bn.beta[grp] + c);
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  auto body = [&](size_t e, const float4& xv, const float4& lv) {
    float xs[4] = {xv.x, xv.y, xv.z, xv.w};
    float ls[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = sigmoidf_(fmaf(ga[j], xs[j], be[j]));
      float o = ls[j] * s;
      if (thresh) o = bsed_keep((uint32_t)(e + j), key, thresh) ? o * inv_keep : 0.f;
      acc[j] += o;
    }
  };
  if (PT && PF) {
    float4 xv[PT ? PT * PF : 1], lv[PT ? PT * PF : 1];
#pragma unroll
    for (int w = 0; w < PT * PF; ++w) {
      size_t e = (((size_t)clip * T + to * PT + w / (PF ? PF : 1)) * F + fo * PF + w % (PF ? PF : 1)) * C + c;
      xv[w] = *reinterpret_cast<const float4*>(xhat + e);
      lv[w] = *reinterpret_cast<const float4*>(lin + e);
    }
#pragma unroll
    for (int w = 0; w < PT * PF; ++w) {
      size_t e = (((size_t)clip * T + to * PT + w / (PF ? PF : 1)) * F + fo * PF + w % (PF ? PF : 1)) * C + c;
      body(e, xv[w], lv[w]);
    }
  } else {
    for (int dt = 0; dt < pt; ++dt)
      for (int df = 0; df < pf; ++df) {
        size_t e = (((size_t)clip * T + to * pt + dt) * F + fo * pf + df) * C + c;
        body(e, *reinterpret_cast<const float4*>(xhat + e), *reinterpret_cast<const float4*>(lin + e));
      }
  }
  float inv = 1.0f / (float)(pt * pf);
  float4 o = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  *reinterpret_cast<float4*>(pooled + (((size_t)clip * To + to) * Fo + fo) * C + c) = o;
}

int glu_gate_pool_fwd(const float* xhat, const float* lin, float* pooled, const Groups& g, const BNPtrs& bn,
                      int T, int F, int C, int pt, int pf, DropKey key, uint32_t thresh, float inv_keep,
                      cudaStream_t st) {
  int To = T / pt, Fo = F / pf;
  long long work = (long long)To * Fo * (C / 4);
  dim3 grid(ceil_div(work, 256), total_clips(g));
  if (pt == 2 && pf == 2)
    glu_gate_pool_fwd_kernel<2, 2><<<grid, 256, 0, st>>>(xhat, lin, pooled, g, bn, T, F, C, pt, pf, To, Fo, key, thresh, inv_keep);
  else if (pt == 1 && pf == 2)
    glu_gate_pool_fwd_kernel<1, 2><<<grid, 256, 0, st>>>(xhat, lin, pooled, g, bn, T, F, C, pt, pf, To, Fo, key, thresh, inv_keep);
  else
    glu_gate_pool_fwd_kernel<0, 0><<<grid, 256, 0, st>>>(xhat, lin, pooled, g, bn, T, F, C, pt, pf, To, Fo, key, thresh, inv_keep);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// backward: per full-resolution element
//   g      = dpooled[window] / (pt*pf) * keep / (1-p)        (0 outside the pooled region)
//   d_lin  = g * s                      -> overwrites lin
//   dxn    = g * lin * s * (1 - s)      -> direct path of the gate (GLU linear path added by GEMM)
__global__ void __launch_bounds__(256) glu_gate_pool_bwd_kernel(const float* __restrict__ xhat,
                                                                float* __restrict__ lin_dlin,
                                                                const float* __restrict__ dpooled,
                                                                float* __restrict__ dxn, Groups g, BNPtrs bn, int T,
                                                                int F, int C, int pt, int pf, int To, int Fo,
                                                                DropKey dkey, uint32_t thresh, float inv_keep) {
  const uint32_t key = dkey.get();
  constexpr int U = 2;   // quads per thread, loads issued first
  const int clip = g.first[0] + blockIdx.y;
  const int grp = group_of(g, clip);
  const int nq = C / 4;
  const long long total = (long long)T * F * nq;
  const long long base = (long long)blockIdx.x * (256 * U) + threadIdx.x;
  const float inv = 1.0f / (float)(pt * pf);
  float4 gv[U], xv[U], lv[U];
  size_t e[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    long long id = base + u * 256;
    if (id >= total) continue;
    int q = (int)(id % nq);
    int pix = (int)(id / nq);
    int f = pix % F, t = pix / F;
    e[u] = (((size_t)clip * T + t) * F + f) * C + q * 4;
    int to = t / pt, fo = f / pf;
    gv[u] = make_float4(0, 0, 0, 0);
    if (to < To && fo < Fo) gv[u] = *reinterpret_cast<const float4*>(dpooled + (((size_t)clip * To + to) * Fo + fo) * C + q * 4);
    xv[u] = *reinterpret_cast<const float4*>(xhat + e[u]);
    lv[u] = *reinterpret_cast<const float4*>(lin_dlin + e[u]);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    long long id = base + u * 256;
    if (id >= total) continue;
    const int c = (int)(id % nq) * 4;
    const float4 ga4 = *reinterpret_cast<const float4*>(bn.gamma[grp] + c);
    const float4 be4 = *reinterpret_cast<const float4*>(bn.beta[grp] + c);
    const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
    float gs[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
    float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
    float ls[4] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w};
    float dl[4], dx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float gg = gs[j] * inv;
      if (thresh) gg = bsed_keep((uint32_t)(e[u] + j), key, thresh) ? gg * inv_keep : 0.f;
      float s = sigmoidf_(fmaf(ga[j], xs[j], be[j]));
      dl[j] = gg * s;
      dx[j] = gg * ls[j] * s * (1.f - s);
    }
    *reinterpret_cast<float4*>(lin_dlin + e[u]) = make_float4(dl[0], dl[1], dl[2], dl[3]);
    *reinterpret_cast<float4*>(dxn + e[u]) = make_float4(dx[0], dx[1], dx[2], dx[3]);
  }
}

int glu_gate_pool_bwd(const float* xhat, float* lin_dlin, const float* dpooled, float* dxn, const Groups& g,
                      const BNPtrs& bn, int T, int F, int C, int pt, int pf, DropKey key, uint32_t thresh,
                      float inv_keep, cudaStream_t st) {
  int To = T / pt, Fo = F / pf;
  long long work = (long long)T * F * (C / 4);
  dim3 grid(ceil_div(work, 256 * 2), total_clips(g));
  glu_gate_pool_bwd_kernel<<<grid, 256, 0, st>>>(xhat, lin_dlin, dpooled, dxn, g, bn, T, F, C, pt, pf, To, Fo,
                                                 key, thresh, inv_keep);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused variant used by the tensor-core path: the gate / dropout / pool backward of glu_gate_pool_bwd plus, per
// group and channel, the three column sums every later BatchNorm-backward statistic can be derived from:
//   A1[c] = sum d_lin[p][c]      A2[c] = sum dxd[p][c]      A3[c] = sum dxd[p][c] * xhat[p][c]
// (dxd = the direct gate path of dxn).  With G = d_lin^T xhat of the same group (one weight-gradient GEMM),
//   s1[c] = sum_p dxn = A2[c] + sum_c' A1[c'] Wg[c'][c]
//   s2[c] = sum_p dxn * xhat = A3[c] + sum_c' Wg[c'][c] G[c'][c]
// so no pass over d_lin / dxn is needed for them.  Each CTA owns a row range of one clip (few, long-lived CTAs:
// one fp64 atomic per channel and sum per CTA).  sums: double [group][C][4].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) glu_gate_pool_bwd_sums_kernel(const float* __restrict__ xhat,
                                                                     float* __restrict__ lin_dlin,
                                                                     const float* __restrict__ dpooled,
                                                                     float* __restrict__ dxn, Groups g, BNPtrs bn,
                                                                     int T, int F, int C, int pt, int pf, int To,
                                                                     int Fo, DropKey dkey, uint32_t thresh,
                                                                     float inv_keep, long long rows_per_cta,
                                                                     double* __restrict__ sums) {
  const uint32_t key = dkey.get();
  constexpr int U = 2;
  const int clip = g.first[0] + blockIdx.y;
  const int grp = group_of(g, clip);
  const int nq = C / 4;
  const int q = threadIdx.x % nq;
  const int r0 = threadIdx.x / nq;
  const int rstep = 256 / nq;
  const long long rows = (long long)T * F;
  const long long rbeg = (long long)blockIdx.x * rows_per_cta;
  long long rend = rbeg + rows_per_cta;
  if (rend > rows) rend = rows;
  const int c = q * 4;
  const float4 ga4 = *reinterpret_cast<const float4*>(bn.gamma[grp] + c);
  const float4 be4 = *reinterpret_cast<const float4*>(bn.beta[grp] + c);
  const float ga[4] = {ga4.x, ga4.y, ga4.z, ga4.w}, be[4] = {be4.x, be4.y, be4.z, be4.w};
  const float inv = 1.0f / (float)(pt * pf);
  float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0}, a3[4] = {0, 0, 0, 0};
  for (long long r = rbeg + r0; r < rend; r += (long long)U * rstep) {
    float4 gv[U], xv[U], lv[U];
    size_t e[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long rr = r + (long long)u * rstep;
      ok[u] = rr < rend;
      if (!ok[u]) continue;
      const int f = (int)(rr % F), t = (int)(rr / F);
      e[u] = (((size_t)clip * T + t) * F + f) * C + c;
      const int to = t / pt, fo = f / pf;
      gv[u] = make_float4(0, 0, 0, 0);
      if (to < To && fo < Fo) gv[u] = *reinterpret_cast<const float4*>(dpooled + (((size_t)clip * To + to) * Fo + fo) * C + c);
      xv[u] = *reinterpret_cast<const float4*>(xhat + e[u]);
      lv[u] = *reinterpret_cast<const float4*>(lin_dlin + e[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      const float gs[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
      const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      const float ls[4] = {lv[u].x, lv[u].y, lv[u].z, lv[u].w};
      float dl[4], dx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float gg = gs[j] * inv;
        if (thresh) gg = bsed_keep((uint32_t)(e[u] + j), key, thresh) ? gg * inv_keep : 0.f;
        const float sg = sigmoidf_(fmaf(ga[j], xs[j], be[j]));
        dl[j] = gg * sg;
        dx[j] = gg * ls[j] * sg * (1.f - sg);
        a1[j] += dl[j];
        a2[j] += dx[j];
        a3[j] = fmaf(dx[j], xs[j], a3[j]);
      }
      *reinterpret_cast<float4*>(lin_dlin + e[u]) = make_float4(dl[0], dl[1], dl[2], dl[3]);
      *reinterpret_cast<float4*>(dxn + e[u]) = make_float4(dx[0], dx[1], dx[2], dx[3]);
    }
  }
  __shared__ float red[3][256][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[0][threadIdx.x][j] = a1[j];
    red[1][threadIdx.x][j] = a2[j];
    red[2][threadIdx.x][j] = a3[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * C; i += 256) {
    const int k = i / C, ch = i % C;
    const int cq = ch / 4, cj = ch % 4;
    float t = 0.f;
    for (int r = 0; r < rstep; ++r) t += red[k][r * nq + cq][cj];
    atomicAdd(sums + ((size_t)grp * C + ch) * 4 + k, (double)t);
  }
}

int glu_gate_pool_bwd_sums(const float* xhat, float* lin_dlin, const float* dpooled, float* dxn, const Groups& g,
                           const BNPtrs& bn, int T, int F, int C, int pt, int pf, DropKey key, uint32_t thresh,
                           float inv_keep, double* sums, int num_sms, cudaStream_t st) {
  BSED_REQUIRE(C % 4 == 0 && 256 % (C / 4) == 0, "glu_gate_pool_bwd_sums: C=%d", C);
  const int To = T / pt, Fo = F / pf;
  const int nclips = total_clips(g);
  const long long rows = (long long)T * F;
  long long want = (long long)num_sms * 8 / nclips + 1;       // ~8 CTAs per SM overall
  long long rows_per_cta = (rows + want - 1) / want;
  const long long gran = 2LL * (256 / (C / 4));                 // rows one CTA pass covers
  rows_per_cta = (rows_per_cta + gran - 1) / gran * gran;
  dim3 grid(ceil_div(rows, rows_per_cta), nclips);
  glu_gate_pool_bwd_sums_kernel<<<grid, 256, 0, st>>>(xhat, lin_dlin, dpooled, dxn, g, bn, T, F, C, pt, pf, To, Fo, key,
                                                      thresh, inv_keep, rows_per_cta, sums);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// BatchNorm / GLU backward bookkeeping of the tensor-core path, one launch per block (C <= 128 threads of work):
//   s1, s2 (see above) per group -> table for the fused GEMM epilogue, PACK times replicated:
//       tab[g][0][j] = gamma*rstd, tab[g][1][j] = s1/n, tab[g][2][j] = s2/n      (j = p*C + c)
//   parameter gradients: d_beta += sum_g s1, d_gamma += sum_g s2, d_bg += sum_g A1,
//                        d_Wg[c'][c] += gamma[c] * sum_g G[g][c'][c] + beta[c] * sum_g A1[g][c']
constexpr int kPrepParts = 4;   // block 0: 512 threads = 128 channels x 4 slices of the c' range
__global__ void __launch_bounds__(128 * kPrepParts) bn_bwd_prepare_kernel(const double* __restrict__ sums, const float* __restrict__ G,
                                                             int n_groups, int C, Groups g, long long rows_per_clip,
                                                             BNPtrs bn, const float* __restrict__ wg,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int pack,
                                                             float* __restrict__ tab, float* d_gamma, float* d_beta,
                                                             float* d_wg, float* d_bg) {
  if (blockIdx.x > 0) {
    // d_Wg[c'][c] += gamma[c] * sum_g G[g][c'][c] + beta[c] * sum_g A1[g][c']     (one element per thread)
    const int i = (blockIdx.x - 1) * blockDim.x + threadIdx.x;
    if (i >= C * C) return;
    const int cp = i / C, col = i % C;
    float gsum = 0.f;
    double a1 = 0.0;
    for (int gi = 0; gi < n_groups; ++gi) {
      gsum += G[(size_t)gi * C * C + i];
      a1 += sums[((size_t)gi * C + cp) * 4];
    }
    d_wg[i] += gamma[col] * gsum + beta[col] * (float)a1;
    return;
  }
  // block 0: per-channel statistics.  Thread (c, part) sums one slice of the c' range (this launch sits on the critical
  // path of every block's backward: the loop is latency-bound, so it is cut four ways with its loads in flight together);
  // combined through smem in a fixed order.
  __shared__ double ps1[kMaxGroups][kPrepParts][128], ps2[kMaxGroups][kPrepParts][128];
  const int c = threadIdx.x % 128, part = threadIdx.x / 128;
  if (c < C) {
    for (int gi = 0; gi < n_groups; ++gi) {
      const double* sg = sums + (size_t)gi * C * 4;
      const float* Gg = G + (size_t)gi * C * C;
      double s1 = 0.0, s2 = 0.0;
#pragma unroll 4
      for (int cp = part; cp < C; cp += kPrepParts) {
        const double w = (double)wg[(size_t)cp * C + c];
        s1 += sg[cp * 4 + 0] * w;
        s2 += w * (double)Gg[(size_t)cp * C + c];
      }
      ps1[gi][part][c] = s1;
      ps2[gi][part][c] = s2;
    }
  }
  __syncthreads();
  if (part != 0 || c >= C) return;
  double s1_all = 0.0, s2_all = 0.0, a1_all = 0.0;
  for (int gi = 0; gi < n_groups; ++gi) {
    const double* sg = sums + (size_t)gi * C * 4;
    double s1 = sg[c * 4 + 1], s2 = sg[c * 4 + 2];
#pragma unroll
    for (int p = 0; p < kPrepParts; ++p) {
      s1 += ps1[gi][p][c];
      s2 += ps2[gi][p][c];
    }
    const double n = (double)g.count[gi] * (double)rows_per_clip;
    const float k = bn.gamma[gi][c] * bn.rstd[gi][c];
    for (int p = 0; p < pack; ++p) {
      float* t = tab + (size_t)gi * 3 * C * pack + p * C + c;
      t[0] = k;
      t[(size_t)C * pack] = (float)(s1 / n);
      t[(size_t)2 * C * pack] = (float)(s2 / n);
    }
    s1_all += s1;
    s2_all += s2;
    a1_all += sg[c * 4 + 0];
  }
  d_beta[c] += (float)s1_all;
  d_gamma[c] += (float)s2_all;
  d_bg[c] += (float)a1_all;
}

int bn_bwd_prepare(const double* sums, const float* G, int n_groups, int C, const Groups& g, long long rows_per_clip,
                   const BNPtrs& bn, const float* wg, const float* gamma, const float* beta, int pack, float* tab,
                   float* d_gamma, float* d_beta, float* d_wg, float* d_bg, cudaStream_t st) {
  constexpr int threads = 128 * kPrepParts;
  bn_bwd_prepare_kernel<<<1 + ceil_div(C * C, threads), threads, 0, st>>>(sums, G, n_groups, C, g, rows_per_clip, bn, wg, gamma,
                                                                          beta, pack, tab, d_gamma, d_beta, d_wg, d_bg);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// dY = gamma * rstd * (dxn - s1/n - xhat * s2/n), in place on dxn
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* __restrict__ dxn, const float* __restrict__ xhat,
                                                           const double* __restrict__ stats2, Groups g,
                                                           long long elems_per_clip, long long rows_per_clip, int C,
                                                           BNPtrs bn) {
  const int clip = g.first[0] + blockIdx.y;
  const int grp = group_of(g, clip);
  const long long n4 = elems_per_clip / 4;
  const long long base = (long long)blockIdx.x * (256 * kEltU) + threadIdx.x;
  const float invn = 1.0f / ((float)g.count[grp] * (float)rows_per_clip);
  float4* pd = reinterpret_cast<float4*>(dxn + (size_t)clip * elems_per_clip);
  const float4* px = reinterpret_cast<const float4*>(xhat + (size_t)clip * elems_per_clip);
  float4 d[kEltU], xh[kEltU];
#pragma unroll
  for (int u = 0; u < kEltU; ++u) {
    long long i4 = base + u * 256;
    if (i4 < n4) {
      d[u] = pd[i4];
      xh[u] = px[i4];
    }
  }
#pragma unroll
  for (int u = 0; u < kEltU; ++u) {
    long long i4 = base + u * 256;
    if (i4 >= n4) continue;
    int c = (int)((i4 * 4) % C);
    float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
    float xs[4] = {xh[u].x, xh[u].y, xh[u].z, xh[u].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s1 = (float)stats2[((size_t)grp * C + c + j) * 2] * invn;
      float s2 = (float)stats2[((size_t)grp * C + c + j) * 2 + 1] * invn;
      float k = bn.gamma[grp][c + j] * bn.rstd[grp][c + j];
      dv[j] = k * (dv[j] - s1 - xs[j] * s2);
    }
    pd[i4] = make_float4(dv[0], dv[1], dv[2], dv[3]);
  }
}

int bn_bwd_apply(float* dxn_dy, const float* xhat, const double* stats2, const Groups& g,
                 long long rows_per_clip, int C, const BNPtrs& bn, cudaStream_t st) {
  long long elems = rows_per_clip * C;
  dim3 grid(ceil_div(elems / 4, 256 * kEltU), total_clips(g));
  bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(dxn_dy, xhat, stats2, g, elems, rows_per_clip, C, bn);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// parameter gradients of BN (gamma, beta) and the un-folding of the GLU linear gradient:
//   lin = Wg (gamma*xhat + beta) + bg  =>  dWg[c'][c] = gamma[c] * G[c'][c] + beta[c] * dbg[c'],
//   G = d_lin^T xhat, dbg = colsum(d_lin)
__global__ void bn_glu_param_grads_kernel(const double* __restrict__ stats2, int n_groups, int C,
                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                          const float* __restrict__ G, const double* __restrict__ dbg,
                                          float* d_gamma, float* d_beta, float* d_wg, float* d_bg) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    double s1 = 0, s2 = 0;
    for (int gI = 0; gI < n_groups; ++gI) {
      s1 += stats2[((size_t)gI * C + i) * 2];
      s2 += stats2[((size_t)gI * C + i) * 2 + 1];
    }
    d_beta[i] += (float)s1;
    d_gamma[i] += (float)s2;
    d_bg[i] += (float)dbg[(size_t)i * 2];
  }
  if (i < C * C) {
    int cp = i / C, c = i % C;
    d_wg[i] += gamma[c] * G[i] + beta[c] * (float)dbg[(size_t)cp * 2];
  }
}

int bn_glu_param_grads(const double* stats2, int n_groups, int C, const float* gamma, const float* beta,
                       const float* G, const double* dbg, float* d_gamma, float* d_beta, float* d_wg,
                       float* d_bg, cudaStream_t st) {
  bn_glu_param_grads_kernel<<<ceil_div(C * C, 256), 256, 0, st>>>(stats2, n_groups, C, gamma, beta, G, dbg,
                                                                  d_gamma, d_beta, d_wg, d_bg);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace bsed
