// resnet.cu -- channels-last building blocks of the ResNet-18 weak tagger's INFERENCE path.
//
// Reference: Net_resnet (src/audio_tagging_system_cnn.py:50-64) = torchvision resnet18 with conv1 replaced by
// Conv2d(1, 64, 7, stride 2, padding 3, bias=False) and fc by Linear(512, 20), followed by a sigmoid; used in eval mode by
// src/audio_tagging_inference.py:123-133, 289-316 to write the pseudo-label TSV.  In eval mode every BatchNorm is an
// affine map that the host folds into the preceding convolution, so a stage is
//     im2col (any kernel / stride / padding)  ->  GEMM + bias (bsed_gemm_nt_tc / bsed_gemm_nn)  ->  [+ residual] -> ReLU
// plus the stem's 3x3 / stride-2 max-pool, the global average pool and the sigmoid.
#include "launch.h"

namespace bsed {

// col[(b*Ho + ho)*Wo + wo][(ky*kw + kx)*Cin + ci] = x[b][ho*sh - ph + ky][wo*sw - pw + kx][ci]  (0 outside), columns
// [kh*kw*Cin, Kpad) zero.  One thread per 4 consecutive columns when Cin % 4 == 0, else scalar.
template <bool VEC>
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const float* __restrict__ x, float* __restrict__ col, int H, int W,
                                                          int Cin, int kh, int kw, int sh, int sw, int ph, int pw, int Ho,
                                                          int Wo, int Kpad, long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  constexpr int V = VEC ? 4 : 1;
  const int kq = Kpad / V;
  const int k = (int)(id % kq) * V;
  const long long row = id / kq;
  const int wo = (int)(row % Wo);
  const int ho = (int)((row / Wo) % Ho);
  const long long b = row / ((long long)Wo * Ho);
  const int K = kh * kw * Cin;
  float v[V];
#pragma unroll
  for (int j = 0; j < V; ++j) v[j] = 0.f;
  if (k < K) {
    const int ci = k % Cin, tap = k / Cin;
    const int ky = tap / kw, kx = tap % kw;
    const int hi = ho * sh - ph + ky, wi = wo * sw - pw + kx;
    if (hi >= 0 && hi < H && wi >= 0 && wi < W) {
      const float* src = x + ((b * H + hi) * W + wi) * Cin + ci;
      if (VEC) {
        const float4 t = *reinterpret_cast<const float4*>(src);
        v[0] = t.x;
        v[V > 1 ? 1 : 0] = t.y;
        v[V > 2 ? 2 : 0] = t.z;
        v[V > 3 ? 3 : 0] = t.w;
      } else {
        v[0] = *src;
      }
    }
  }
  if (VEC)
    *reinterpret_cast<float4*>(col + row * Kpad + k) = make_float4(v[0], v[V > 1 ? 1 : 0], v[V > 2 ? 2 : 0], v[V > 3 ? 3 : 0]);
  else
    col[row * Kpad + k] = v[0];
}

int im2col_nhwc(const float* x, float* col, int B, int H, int W, int Cin, int kh, int kw, int sh, int sw, int ph, int pw,
                int Ho, int Wo, int Kpad, cudaStream_t st) {
  BSED_REQUIRE(B >= 1 && H >= 1 && W >= 1 && Cin >= 1 && kh >= 1 && kw >= 1 && sh >= 1 && sw >= 1, "im2col: bad geometry");
  BSED_REQUIRE(Ho == (H + 2 * ph - kh) / sh + 1 && Wo == (W + 2 * pw - kw) / sw + 1, "im2col: output size %dx%d", Ho, Wo);
  BSED_REQUIRE(Kpad >= kh * kw * Cin && Kpad % 4 == 0, "im2col: Kpad=%d", Kpad);
  const bool vec = Cin % 4 == 0;
  const long long total = (long long)B * Ho * Wo * (vec ? Kpad / 4 : Kpad);
  if (vec)
    im2col_nhwc_kernel<true><<<ceil_div(total, 256), 256, 0, st>>>(x, col, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo, Kpad, total);
  else
    im2col_nhwc_kernel<false><<<ceil_div(total, 256), 256, 0, st>>>(x, col, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo, Kpad, total);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// y = relu(y + residual) (residual may be NULL), float4
__global__ void __launch_bounds__(256) add_relu_kernel(float* __restrict__ y, const float* __restrict__ res, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<float4*>(y)[i];
  if (res) {
    const float4 r = reinterpret_cast<const float4*>(res)[i];
    v.x += r.x;
    v.y += r.y;
    v.z += r.z;
    v.w += r.w;
  }
  reinterpret_cast<float4*>(y)[i] = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}

int add_relu(float* y, const float* res, long long n, cudaStream_t st) {
  BSED_REQUIRE(n > 0 && n % 4 == 0, "add_relu: n=%lld must be a positive multiple of 4", n);
  add_relu_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(y, res, n / 4);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// nn.MaxPool2d(k, stride s, padding p) on channels-last x [B][H][W][C] -> y [B][Ho][Wo][C] (padding = -inf)
__global__ void __launch_bounds__(256) maxpool_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W,
                                                           int C, int k, int s, int p, int Ho, int Wo, long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of y
  if (id >= total) return;
  const int cq = C / 4;
  const int c = (int)(id % cq) * 4;
  const long long pix = id / cq;
  const int wo = (int)(pix % Wo), ho = (int)((pix / Wo) % Ho);
  const long long b = pix / ((long long)Wo * Ho);
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int dy = 0; dy < k; ++dy) {
    const int hi = ho * s - p + dy;
    if (hi < 0 || hi >= H) continue;
    for (int dx = 0; dx < k; ++dx) {
      const int wi = wo * s - p + dx;
      if (wi < 0 || wi >= W) continue;
      const float4 v = *reinterpret_cast<const float4*>(x + ((b * H + hi) * W + wi) * C + c);
      m.x = fmaxf(m.x, v.x);
      m.y = fmaxf(m.y, v.y);
      m.z = fmaxf(m.z, v.z);
      m.w = fmaxf(m.w, v.w);
    }
  }
  *reinterpret_cast<float4*>(y + pix * C + c) = m;
}

int maxpool_nhwc(const float* x, float* y, int B, int H, int W, int C, int k, int s, int p, int Ho, int Wo, cudaStream_t st) {
  BSED_REQUIRE(C % 4 == 0 && Ho == (H + 2 * p - k) / s + 1 && Wo == (W + 2 * p - k) / s + 1, "maxpool: C=%d out %dx%d", C, Ho, Wo);
  const long long total = (long long)B * Ho * Wo * (C / 4);
  maxpool_nhwc_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x, y, H, W, C, k, s, p, Ho, Wo, total);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// AdaptiveAvgPool2d(1): y[b][c] = mean over the HW pixels of x [B][HW][C]; one CTA per (clip, 32-channel slab)
__global__ void __launch_bounds__(256) avgpool_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int HW, int C) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int p = r; p < HW; p += 8) s += x[((size_t)b * HW + p) * C + c];
  __shared__ float red[8][32];
  red[r][threadIdx.x & 31] = s;
  __syncthreads();
  if (r == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    y[(size_t)b * C + c] = t / (float)HW;
  }
}

int avgpool_nhwc(const float* x, float* y, int B, int HW, int C, cudaStream_t st) {
  BSED_REQUIRE(B >= 1 && HW >= 1 && C >= 1, "avgpool: bad shape");
  dim3 grid(ceil_div(C, 32), B);
  avgpool_nhwc_kernel<<<grid, 256, 0, st>>>(x, y, HW, C);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// out[r][c] = sigmoid(logits[r * ld + c]) for c < C
__global__ void sigmoid_rows_kernel(const float* __restrict__ logits, int ld, float* __restrict__ out, int C, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / C;
  const int c = (int)(i % C);
  out[i] = 1.0f / (1.0f + expf(-logits[r * ld + c]));
}

int sigmoid_rows(const float* logits, int ld, float* out, int rows, int C, cudaStream_t st) {
  BSED_REQUIRE(rows >= 1 && C >= 1 && ld >= C, "sigmoid_rows: rows=%d C=%d ld=%d", rows, C, ld);
  const long long n = (long long)rows * C;
  sigmoid_rows_kernel<<<ceil_div(n, 256), 256, 0, st>>>(logits, ld, out, C, n);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// Training path of the tagger (src/audio_tagging_system_cnn.py:199-416: two model calls, BCE on the weak outputs, Adam).
// Every convolution is bias-free and followed by a train-mode BatchNorm2d (eps 1e-5, momentum 0.1) on rows [M][C]:
//   forward : statistics (col_stats, fp64) -> finalize (+ running statistics, unbiased variance) ->
//             xhat = (x - mean) * rstd in place, y = gamma * xhat + beta [+ residual] [ReLU]
//   backward: dz = dy * (y > 0) ; dgamma += sum dz * xhat ; dbeta += sum dz ;
//             dx = gamma * rstd * (dz - sum(dz) / M - xhat * sum(dz * xhat) / M)
// ---------------------------------------------------------------------------------------------
__global__ void bn_rows_finalize_kernel(const double* __restrict__ sums, long long M, int C, float eps, float momentum,
                                        float* __restrict__ mean_rstd, float* run_mean, float* run_var, int64_t* nbt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double n = (double)M;
  const double mean = sums[c * 2] / n;
  double var = sums[c * 2 + 1] / n - mean * mean;
  if (var < 0) var = 0;
  mean_rstd[c] = (float)mean;
  mean_rstd[C + c] = (float)(1.0 / sqrt(var + (double)eps));
  if (run_mean) {
    const double unb = n > 1 ? var * n / (n - 1) : var;
    run_mean[c] = (1.f - momentum) * run_mean[c] + momentum * (float)mean;
    run_var[c] = (1.f - momentum) * run_var[c] + momentum * (float)unb;
  }
  if (c == 0 && nbt) *nbt += 1;
}

__global__ void __launch_bounds__(256) bn_rows_apply_kernel(float* __restrict__ x, const float* __restrict__ mean_rstd,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ residual, int relu,
                                                            float* __restrict__ y, int C, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)((i * 4) % C);
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  const float4 mu = *reinterpret_cast<const float4*>(mean_rstd + c), rs = *reinterpret_cast<const float4*>(mean_rstd + C + c);
  const float4 ga = *reinterpret_cast<const float4*>(gamma + c), be = *reinterpret_cast<const float4*>(beta + c);
  const float4 xh = make_float4((v.x - mu.x) * rs.x, (v.y - mu.y) * rs.y, (v.z - mu.z) * rs.z, (v.w - mu.w) * rs.w);
  reinterpret_cast<float4*>(x)[i] = xh;
  float4 o = make_float4(fmaf(ga.x, xh.x, be.x), fmaf(ga.y, xh.y, be.y), fmaf(ga.z, xh.z, be.z), fmaf(ga.w, xh.w, be.w));
  if (residual) {
    const float4 r = reinterpret_cast<const float4*>(residual)[i];
    o.x += r.x;
    o.y += r.y;
    o.z += r.z;
    o.w += r.w;
  }
  if (relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
  reinterpret_cast<float4*>(y)[i] = o;
}

int bn_rows_train(float* x, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                  float* run_mean, float* run_var, int64_t* nbt, const float* residual, int relu, float* y, float* mean_rstd,
                  double* ws, int num_sms, cudaStream_t st) {
  BSED_REQUIRE(M >= 1 && C % 4 == 0 && C <= 1024, "bn_rows_train: M=%lld C=%d", M, C);
  Groups one;
  one.n = 1;
  for (int k = 0; k < kMaxGroups; ++k) one.first[k] = 0, one.count[k] = k == 0 ? 1 : 0;
  BSED_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  BSED_TRY(col_stats(x, nullptr, 0, one, M, C, ws, num_sms, st));
  bn_rows_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, M, C, eps, momentum, mean_rstd, run_mean, run_var, nbt);
  BSED_CHECK_LAUNCH();
  const long long n4 = M * C / 4;
  bn_rows_apply_kernel<<<ceil_div(n4, 256), 256, 0, st>>>(x, mean_rstd, gamma, beta, residual, relu, y, C, n4);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// dz = dy * (y > 0) in place (y == NULL: no ReLU); optionally copied to d_residual
__global__ void __launch_bounds__(256) relu_mask_kernel(float* __restrict__ dy, const float* __restrict__ y,
                                                        float* __restrict__ d_res, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = reinterpret_cast<float4*>(dy)[i];
  if (y) {
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    g = make_float4(v.x > 0.f ? g.x : 0.f, v.y > 0.f ? g.y : 0.f, v.z > 0.f ? g.z : 0.f, v.w > 0.f ? g.w : 0.f);
    reinterpret_cast<float4*>(dy)[i] = g;
  }
  if (d_res) reinterpret_cast<float4*>(d_res)[i] = g;
}

__global__ void bn_rows_param_grads_kernel(const double* __restrict__ sums, int C, float* d_gamma, float* d_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  d_beta[c] += (float)sums[c * 2];
  d_gamma[c] += (float)sums[c * 2 + 1];
}

__global__ void __launch_bounds__(256) bn_rows_bwd_apply_kernel(float* __restrict__ dz, const float* __restrict__ xhat,
                                                                const double* __restrict__ sums,
                                                                const float* __restrict__ mean_rstd,
                                                                const float* __restrict__ gamma, int C, double inv_m,
                                                                long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c = (int)((i * 4) % C);
  const float4 g = reinterpret_cast<const float4*>(dz)[i], xh = reinterpret_cast<const float4*>(xhat)[i];
  const float gs[4] = {g.x, g.y, g.z, g.w}, xs[4] = {xh.x, xh.y, xh.z, xh.w};
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float k = gamma[c + j] * mean_rstd[C + c + j];
    const float m1 = (float)(sums[(c + j) * 2] * inv_m), m2 = (float)(sums[(c + j) * 2 + 1] * inv_m);
    o[j] = k * (gs[j] - m1 - xs[j] * m2);
  }
  reinterpret_cast<float4*>(dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
}

int bn_rows_backward(float* dy, const float* y, const float* xhat, long long M, int C, const float* gamma,
                     const float* mean_rstd, float* d_gamma, float* d_beta, float* d_residual, double* ws, int num_sms,
                     cudaStream_t st) {
  BSED_REQUIRE(M >= 1 && C % 4 == 0 && C <= 1024, "bn_rows_backward: M=%lld C=%d", M, C);
  const long long n4 = M * C / 4;
  if (y || d_residual) {
    relu_mask_kernel<<<ceil_div(n4, 256), 256, 0, st>>>(dy, y, d_residual, n4);
    BSED_CHECK_LAUNCH();
  }
  Groups one;
  one.n = 1;
  for (int k = 0; k < kMaxGroups; ++k) one.first[k] = 0, one.count[k] = k == 0 ? 1 : 0;
  BSED_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * C, st));
  BSED_TRY(col_stats(dy, xhat, 1, one, M, C, ws, num_sms, st));          // (sum dz, sum dz * xhat)
  bn_rows_param_grads_kernel<<<ceil_div(C, 128), 128, 0, st>>>(ws, C, d_gamma, d_beta);
  BSED_CHECK_LAUNCH();
  bn_rows_bwd_apply_kernel<<<ceil_div(n4, 256), 256, 0, st>>>(dy, xhat, ws, mean_rstd, gamma, C, 1.0 / (double)M, n4);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// transpose of im2col: dx[b][hi][wi][ci] (+)= sum over the windows (ho, wo, ky, kx) that read this pixel of
// dcol[(b*Ho+ho)*Wo+wo][(ky*kw+kx)*Cin+ci]  -- a gather, so no atomics and a fixed summation order
template <bool VEC>
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int H, int W,
                                                          int Cin, int kh, int kw, int sh, int sw, int ph, int pw, int Ho,
                                                          int Wo, int Kpad, int accumulate, long long total) {
  // one thread per input element, or (VEC: Cin % 4 == 0) per four consecutive channels
  constexpr int V = VEC ? 4 : 1;
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const int cq = Cin / V;
  const int ci = (int)(id % cq) * V;
  const long long pix = id / cq;
  const int wi = (int)(pix % W), hi = (int)((pix / W) % H);
  const long long b = pix / ((long long)W * H);
  float acc[V];
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  for (int ky = 0; ky < kh; ++ky) {
    const int t = hi + ph - ky;
    if (t < 0 || t % sh != 0) continue;
    const int ho = t / sh;
    if (ho >= Ho) continue;
    for (int kx = 0; kx < kw; ++kx) {
      const int u = wi + pw - kx;
      if (u < 0 || u % sw != 0) continue;
      const int wo = u / sw;
      if (wo >= Wo) continue;
      const float* src = dcol + ((b * Ho + ho) * Wo + wo) * Kpad + (ky * kw + kx) * Cin + ci;
      if (VEC) {
        const float4 v = *reinterpret_cast<const float4*>(src);
        acc[0] += v.x;
        acc[V > 1 ? 1 : 0] += v.y;
        acc[V > 2 ? 2 : 0] += v.z;
        acc[V > 3 ? 3 : 0] += v.w;
      } else {
        acc[0] += *src;
      }
    }
  }
  float* dst = dx + pix * Cin + ci;
  if (VEC) {
    float4 o = make_float4(acc[0], acc[V > 1 ? 1 : 0], acc[V > 2 ? 2 : 0], acc[V > 3 ? 3 : 0]);
    if (accumulate) {
      const float4 p = *reinterpret_cast<const float4*>(dst);
      o = make_float4(o.x + p.x, o.y + p.y, o.z + p.z, o.w + p.w);
    }
    *reinterpret_cast<float4*>(dst) = o;
  } else {
    if (accumulate) *dst += acc[0];
    else *dst = acc[0];
  }
}

int col2im_nhwc(const float* dcol, float* dx, int B, int H, int W, int Cin, int kh, int kw, int sh, int sw, int ph, int pw,
                int Ho, int Wo, int Kpad, int accumulate, cudaStream_t st) {
  BSED_REQUIRE(Ho == (H + 2 * ph - kh) / sh + 1 && Wo == (W + 2 * pw - kw) / sw + 1 && Kpad >= kh * kw * Cin, "col2im: geometry");
  const bool vec = Cin % 4 == 0 && Kpad % 4 == 0;
  const long long total = (long long)B * H * W * (vec ? Cin / 4 : Cin);
  if (vec)
    col2im_nhwc_kernel<true><<<ceil_div(total, 256), 256, 0, st>>>(dcol, dx, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo, Kpad,
                                                                   accumulate, total);
  else
    col2im_nhwc_kernel<false><<<ceil_div(total, 256), 256, 0, st>>>(dcol, dx, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo, Kpad,
                                                                    accumulate, total);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// MaxPool2d backward: the gradient of a window goes to its first maximum in (dy, dx) scan order (ATen's rule); gather
__global__ void __launch_bounds__(256) maxpool_nhwc_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                               float* __restrict__ dx, int H, int W, int C, int k, int s,
                                                               int p, int Ho, int Wo, long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const int c = (int)(id % C);
  const long long pix = id / C;
  const int wi = (int)(pix % W), hi = (int)((pix / W) % H);
  const long long b = pix / ((long long)W * H);
  const float me = x[id];
  float acc = 0.f;
  for (int ho = max(0, (hi + p - k + s) / s); ho < Ho && ho * s - p <= hi; ++ho)
    for (int wo = max(0, (wi + p - k + s) / s); wo < Wo && wo * s - p <= wi; ++wo) {
      // is (hi, wi) the first maximum of window (ho, wo)?
      bool first = true;
      for (int dy_ = 0; dy_ < k && first; ++dy_) {
        const int h2 = ho * s - p + dy_;
        if (h2 < 0 || h2 >= H) continue;
        for (int dx_ = 0; dx_ < k; ++dx_) {
          const int w2 = wo * s - p + dx_;
          if (w2 < 0 || w2 >= W) continue;
          const float v = x[((b * H + h2) * W + w2) * C + c];
          const bool before = h2 < hi || (h2 == hi && w2 < wi);
          if (v > me || (before && v == me)) {
            first = false;
            break;
          }
        }
      }
      if (first) acc += dy[((b * Ho + ho) * Wo + wo) * C + c];
    }
  dx[id] = acc;
}

int maxpool_nhwc_backward(const float* x, const float* dy, float* dx, int B, int H, int W, int C, int k, int s, int p, int Ho,
                          int Wo, cudaStream_t st) {
  BSED_REQUIRE(Ho == (H + 2 * p - k) / s + 1 && Wo == (W + 2 * p - k) / s + 1, "maxpool_backward: geometry");
  const long long total = (long long)B * H * W * C;
  maxpool_nhwc_bwd_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x, dy, dx, H, W, C, k, s, p, Ho, Wo, total);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// dx[b][p][c] = dy[b][c] / HW
__global__ void avgpool_nhwc_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int HW, int C, long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  const int c = (int)(id % C);
  const long long b = id / ((long long)HW * C);
  dx[id] = dy[b * C + c] / (float)HW;
}

int avgpool_nhwc_backward(const float* dy, float* dx, int B, int HW, int C, cudaStream_t st) {
  const long long total = (long long)B * HW * C;
  avgpool_nhwc_bwd_kernel<<<ceil_div(total, 256), 256, 0, st>>>(dy, dx, HW, C, total);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// d_logits[r][c] = d_p[r][c] * p (1 - p) for c < C, 0 for the padding columns up to ld
__global__ void sigmoid_rows_bwd_kernel(const float* __restrict__ p, const float* __restrict__ dp, float* __restrict__ dl,
                                        int ld, int C, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long r = i / ld;
  const int c = (int)(i % ld);
  float v = 0.f;
  if (c < C) {
    const float q = p[r * C + c];
    v = dp[r * C + c] * q * (1.f - q);
  }
  dl[i] = v;
}

int sigmoid_rows_backward(const float* p, const float* dp, float* dl, int rows, int C, int ld, cudaStream_t st) {
  BSED_REQUIRE(rows >= 1 && C >= 1 && ld >= C, "sigmoid_rows_backward: rows=%d C=%d ld=%d", rows, C, ld);
  const long long n = (long long)rows * ld;
  sigmoid_rows_bwd_kernel<<<ceil_div(n, 256), 256, 0, st>>>(p, dp, dl, ld, C, n);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace bsed

using namespace bsed;

extern "C" int bsed_im2col_nhwc(bsed_handle h, const float* x, float* col, int B, int H, int W, int Cin, int kh, int kw,
                                int sh, int sw, int ph, int pw, int Ho, int Wo, int Kpad, void* stream) {
  BSED_REQUIRE(h && x && col, "bsed_im2col_nhwc: null argument");
  return im2col_nhwc(x, col, B, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo, Kpad, as_stream(stream));
}
extern "C" int bsed_add_relu(bsed_handle h, float* y, const float* residual, int64_t n, void* stream) {
  BSED_REQUIRE(h && y, "bsed_add_relu: null argument");
  return add_relu(y, residual, n, as_stream(stream));
}
extern "C" int bsed_maxpool_nhwc(bsed_handle h, const float* x, float* y, int B, int H, int W, int C, int k, int s, int p,
                                 int Ho, int Wo, void* stream) {
  BSED_REQUIRE(h && x && y, "bsed_maxpool_nhwc: null argument");
  return maxpool_nhwc(x, y, B, H, W, C, k, s, p, Ho, Wo, as_stream(stream));
}
extern "C" int bsed_avgpool_nhwc(bsed_handle h, const float* x, float* y, int B, int HW, int C, void* stream) {
  BSED_REQUIRE(h && x && y, "bsed_avgpool_nhwc: null argument");
  return avgpool_nhwc(x, y, B, HW, C, as_stream(stream));
}
extern "C" int bsed_sigmoid_rows(bsed_handle h, const float* logits, int ld, float* out, int rows, int C, void* stream) {
  BSED_REQUIRE(h && logits && out, "bsed_sigmoid_rows: null argument");
  return sigmoid_rows(logits, ld, out, rows, C, as_stream(stream));
}

extern "C" size_t bsed_bn_rows_workspace_bytes(int C) { return sizeof(double) * 2 * (size_t)(C > 0 ? C : 0); }
extern "C" int bsed_bn_rows_train(bsed_handle h, float* x, int64_t M, int C, const float* gamma, const float* beta, float eps,
                                  float momentum, float* run_mean, float* run_var, int64_t* nbt, const float* residual,
                                  int relu, float* y, float* mean_rstd, void* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(h && x && gamma && beta && y && mean_rstd && workspace, "bsed_bn_rows_train: null argument");
  if (workspace_bytes < bsed_bn_rows_workspace_bytes(C)) {
    bsed_set_error("bsed_bn_rows_train: workspace %zu < %zu", workspace_bytes, bsed_bn_rows_workspace_bytes(C));
    return BSED_E_WORKSPACE;
  }
  return bn_rows_train(x, M, C, gamma, beta, eps, momentum, run_mean, run_var, nbt, residual, relu, y, mean_rstd,
                       (double*)workspace, h->num_sms, as_stream(stream));
}
extern "C" int bsed_bn_rows_backward(bsed_handle h, float* dy, const float* y, const float* xhat, int64_t M, int C,
                                     const float* gamma, const float* mean_rstd, float* d_gamma, float* d_beta,
                                     float* d_residual, void* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(h && dy && xhat && gamma && mean_rstd && d_gamma && d_beta && workspace, "bsed_bn_rows_backward: null argument");
  if (workspace_bytes < bsed_bn_rows_workspace_bytes(C)) {
    bsed_set_error("bsed_bn_rows_backward: workspace %zu < %zu", workspace_bytes, bsed_bn_rows_workspace_bytes(C));
    return BSED_E_WORKSPACE;
  }
  return bn_rows_backward(dy, y, xhat, M, C, gamma, mean_rstd, d_gamma, d_beta, d_residual, (double*)workspace, h->num_sms,
                          as_stream(stream));
}
extern "C" int bsed_col2im_nhwc(bsed_handle h, const float* dcol, float* dx, int B, int H, int W, int Cin, int kh, int kw,
                                int sh, int sw, int ph, int pw, int Ho, int Wo, int Kpad, int accumulate, void* stream) {
  BSED_REQUIRE(h && dcol && dx, "bsed_col2im_nhwc: null argument");
  return col2im_nhwc(dcol, dx, B, H, W, Cin, kh, kw, sh, sw, ph, pw, Ho, Wo, Kpad, accumulate, as_stream(stream));
}
extern "C" int bsed_maxpool_nhwc_backward(bsed_handle h, const float* x, const float* dy, float* dx, int B, int H, int W, int C,
                                          int k, int s, int p, int Ho, int Wo, void* stream) {
  BSED_REQUIRE(h && x && dy && dx, "bsed_maxpool_nhwc_backward: null argument");
  return maxpool_nhwc_backward(x, dy, dx, B, H, W, C, k, s, p, Ho, Wo, as_stream(stream));
}
extern "C" int bsed_avgpool_nhwc_backward(bsed_handle h, const float* dy, float* dx, int B, int HW, int C, void* stream) {
  BSED_REQUIRE(h && dy && dx, "bsed_avgpool_nhwc_backward: null argument");
  return avgpool_nhwc_backward(dy, dx, B, HW, C, as_stream(stream));
}
extern "C" int bsed_sigmoid_rows_backward(bsed_handle h, const float* p, const float* dp, float* d_logits, int rows, int C, int ld,
                                          void* stream) {
  BSED_REQUIRE(h && p && dp && d_logits, "bsed_sigmoid_rows_backward: null argument");
  return sigmoid_rows_backward(p, dp, d_logits, rows, C, ld, as_stream(stream));
}
