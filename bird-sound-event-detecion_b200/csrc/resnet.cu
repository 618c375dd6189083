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
