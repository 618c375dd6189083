// disc.cu -- Clip_Discriminator of the adversarial domain-adaptation branch.
//
// Reference: src/models/CRNN_GRL.py:16-52
//   x (B,313,256) -> permute(0,2,1).unsqueeze(1) = (B,1,256,313)
//   5 x [Conv2d(3x3, stride 2, no padding) -> BatchNorm2d -> LeakyReLU(0.2)], channels 128, 64, 32, 16, 8
//   (256,313) -> (127,156) -> (63,77) -> (31,38) -> (15,18) -> (7,8)
//   AdaptiveAvgPool2d((2,1)) -> flatten (c*2 + i) -> Linear(16,1) -> sigmoid                      => (B,1)
// and the clip-level domain loss of src/DA/cdan_frame.py:89-119 (plain BCE of D(GRL(f)) against 1 = source, 0 = target).
//
// Activations are channels-last [B][H][W][C].  A strided convolution is im2col (k = tap*Cin + ci, K padded to a multiple
// of 16) followed by a GEMM of this library; its data gradient is the transposed GEMM followed by a col2im scatter, its
// weight gradient the split-K reduction GEMM over the same im2col matrix.  BatchNorm reuses the column-statistics /
// finalise / backward kernels of the CNN blocks.  Parameters arrive as ONE flat fp32 buffer in the reference's
// named_parameters() order: conv_1..5 (weight, bias), dense_d (weight, bias), bn_1..5 (weight, bias).
#include "launch.h"

namespace bsed {
namespace {

constexpr int kDL = 5;
constexpr int kDH0 = 256, kDW0 = 313;
const int kDC[kDL + 1] = {1, 128, 64, 32, 16, 8};

struct DiscGeom {
  int H[kDL + 1], W[kDL + 1];        // spatial size of the input of layer l (index kDL: final)
  int Kp[kDL], Np[kDL];              // GEMM K (9*Cin) and N (Cout) padded to multiples of 16
  long long w_off[kDL], b_off[kDL], dense_w, dense_b, bnw_off[kDL], bnb_off[kDL], n_params;
  long long rm_off[kDL], rv_off[kDL], n_bn;
};

DiscGeom disc_geom() {
  DiscGeom g;
  g.H[0] = kDH0;
  g.W[0] = kDW0;
  long long o = 0;
  for (int l = 0; l < kDL; ++l) {
    g.H[l + 1] = (g.H[l] - 3) / 2 + 1;
    g.W[l + 1] = (g.W[l] - 3) / 2 + 1;
    g.Kp[l] = (9 * kDC[l] + 15) / 16 * 16;
    g.Np[l] = (kDC[l + 1] + 15) / 16 * 16;
    g.w_off[l] = o;
    o += 9LL * kDC[l] * kDC[l + 1];
    g.b_off[l] = o;
    o += kDC[l + 1];
  }
  g.dense_w = o;
  o += 16;
  g.dense_b = o;
  o += 1;
  long long bo = 0;
  for (int l = 0; l < kDL; ++l) {
    g.bnw_off[l] = o;
    o += kDC[l + 1];
    g.bnb_off[l] = o;
    o += kDC[l + 1];
    g.rm_off[l] = bo;
    bo += kDC[l + 1];
    g.rv_off[l] = bo;
    bo += kDC[l + 1];
  }
  g.n_params = o;
  g.n_bn = bo;
  return g;
}

size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

struct DiscWs {
  size_t col[kDL], y[kDL], act[kDL], wt[kDL], wk[kDL], bp[kDL], dwt[kDL], ypad, mr, stats, stats2, feat, dscr, wgpart;
  size_t wt_lo[kDL], wk_lo[kDL];   // 3xTF32: fp32 remainders of the tf32-rounded weight operands
  size_t wgpart_bytes, total;
};

DiscWs disc_ws(const DiscGeom& g, int B) {
  DiscWs w;
  size_t o = 0;
  auto take = [&](size_t bytes) {
    size_t r = o;
    o = align_up(o + bytes);
    return r;
  };
  size_t ypad = 0;
  for (int l = 0; l < kDL; ++l) {
    const size_t M = (size_t)B * g.H[l + 1] * g.W[l + 1];
    w.col[l] = take(M * g.Kp[l] * 4);
    w.y[l] = take(M * kDC[l + 1] * 4);     // conv output, normalised in place (xhat)
    w.act[l] = take(M * kDC[l + 1] * 4);   // LeakyReLU output = input of the next layer; reused for its gradient
    w.wt[l] = take((size_t)g.Kp[l] * g.Np[l] * 4);
    w.wk[l] = take((size_t)g.Np[l] * g.Kp[l] * 4);
    w.bp[l] = take((size_t)g.Np[l] * 4);
    w.dwt[l] = take((size_t)g.Kp[l] * g.Np[l] * 4);
    w.wt_lo[l] = take((size_t)g.Kp[l] * g.Np[l] * 4);
    w.wk_lo[l] = take((size_t)g.Np[l] * g.Kp[l] * 4);
    if (g.Np[l] != kDC[l + 1] && M * g.Np[l] * 4 > ypad) ypad = M * g.Np[l] * 4;
  }
  w.ypad = take(ypad ? ypad : 16);
  w.mr = take(sizeof(float) * kDL * 2 * 128);
  w.stats = take(sizeof(double) * 128 * 2);
  w.stats2 = take(sizeof(double) * 128 * 2);
  w.feat = take(sizeof(float) * (size_t)B * 16);
  w.dscr = take(sizeof(double) * 64);
  w.wgpart_bytes = tc_wgrad_workspace_bytes(148);
  w.wgpart = take(w.wgpart_bytes);
  w.total = o;
  return w;
}

template <class T>
T* wsp(void* ws, size_t off) {
  return reinterpret_cast<T*>(reinterpret_cast<unsigned char*>(ws) + off);
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// x [B][H][W][C] (or, first == 1, d_input [B][W][H] with C = 1) -> col [B*Ho*Wo][Kp], k = tap*C + c, zero padded
__global__ void __launch_bounds__(256) disc_im2col_kernel(const float* __restrict__ x, float* __restrict__ col, int B, int H,
                                                          int W, int C, int Ho, int Wo, int Kp, int first) {
  const long long total = (long long)B * Ho * Wo * Kp;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int k = (int)(i % Kp);
    const long long m = i / Kp;
    float v = 0.f;
    if (k < 9 * C) {
      const int c = k % C, tap = k / C;
      const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho), b = (int)(m / ((long long)Wo * Ho));
      const int h = 2 * ho + tap / 3, w = 2 * wo + tap % 3;
      v = first ? x[((size_t)b * W + w) * H + h] : x[(((size_t)b * H + h) * W + w) * C + c];
    }
    col[i] = v;
  }
}

// dx (zeroed by the caller) += scatter of dcol; layouts as in disc_im2col_kernel
__global__ void __launch_bounds__(256) disc_col2im_kernel(const float* __restrict__ dcol, float* __restrict__ dx, int B, int H,
                                                          int W, int C, int Ho, int Wo, int Kp, int first) {
  const long long total = (long long)B * Ho * Wo * 9 * C;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int k = (int)(i % (9 * C));
    const long long m = i / (9 * C);
    const int c = k % C, tap = k / C;
    const int wo = (int)(m % Wo), ho = (int)((m / Wo) % Ho), b = (int)(m / ((long long)Wo * Ho));
    const int h = 2 * ho + tap / 3, w = 2 * wo + tap % 3;
    const float v = dcol[m * Kp + k];
    float* dst = first ? dx + ((size_t)b * W + w) * H + h : dx + (((size_t)b * H + h) * W + w) * C + c;
    atomicAdd(dst, v);
  }
}

// W (Cout,Cin,3,3) -> Wt [Kp][Np] (k = tap*Cin + ci) and Wk [Np][Kp]; bias -> bp [Np]; padding zero
__global__ void disc_prep_kernel(const float* __restrict__ w, const float* __restrict__ bias, float* wt, float* wk, float* bp,
                                 int Cin, int Cout, int Kp, int Np) {
  const int n = Kp * Np;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int co = i % Np, k = i / Np;
    float v = 0.f;
    if (co < Cout && k < 9 * Cin) v = w[((size_t)co * Cin + k % Cin) * 9 + k / Cin];
    wt[(size_t)k * Np + co] = v;
    wk[(size_t)co * Kp + k] = v;
    if (k == 0) bp[co] = co < Cout ? bias[co] : 0.f;
  }
}

// dWt [Kp][Np] -> dW (Cout,Cin,3,3) +=
__global__ void disc_unpack_dw_kernel(const float* __restrict__ dwt, float* dw, int Cin, int Cout, int Np) {
  const int n = Cout * Cin * 9;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int tap = i % 9, ci = (i / 9) % Cin, co = i / (9 * Cin);
    dw[i] += dwt[(size_t)(tap * Cin + ci) * Np + co];
  }
}

// [M][Np] -> [M][C] (drop the padding columns) / the reverse with zeros
__global__ void disc_compact_kernel(const float* __restrict__ src, float* __restrict__ dst, long long M, int C, int Np, int expand) {
  const long long total = M * (expand ? Np : C);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    if (expand) {
      const int c = (int)(i % Np);
      dst[i] = c < C ? src[(i / Np) * C + c] : 0.f;
    } else {
      dst[i] = src[(i / C) * Np + i % C];
    }
  }
}

// y -> xhat in place; act = leaky(gamma * xhat + beta)
__global__ void __launch_bounds__(256) disc_bn_lrelu_fwd_kernel(float* __restrict__ y, float* __restrict__ act, long long n, int C,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta) {
  // float4 per thread (C % 4 == 0 for every layer: 128 ... 8 channels)
  float4* y4 = reinterpret_cast<float4*>(y);
  float4* a4 = reinterpret_cast<float4*>(act);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n / 4; i += (long long)gridDim.x * 256) {
    const int c = (int)((i * 4) % C);
    const float4 v = y4[i];
    // per-channel values by scalar loads: the BatchNorm parameters sit behind the 17 dense-layer floats in the flat
    // parameter buffer, i.e. not on a 16-byte boundary
    const float4 m = make_float4(mean[c], mean[c + 1], mean[c + 2], mean[c + 3]);
    const float4 r = make_float4(rstd[c], rstd[c + 1], rstd[c + 2], rstd[c + 3]);
    const float4 ga = make_float4(gamma[c], gamma[c + 1], gamma[c + 2], gamma[c + 3]);
    const float4 be = make_float4(beta[c], beta[c + 1], beta[c + 2], beta[c + 3]);
    const float4 xh = make_float4((v.x - m.x) * r.x, (v.y - m.y) * r.y, (v.z - m.z) * r.z, (v.w - m.w) * r.w);
    const float4 z = make_float4(fmaf(ga.x, xh.x, be.x), fmaf(ga.y, xh.y, be.y), fmaf(ga.z, xh.z, be.z), fmaf(ga.w, xh.w, be.w));
    y4[i] = xh;
    a4[i] = make_float4(z.x > 0.f ? z.x : 0.2f * z.x, z.y > 0.f ? z.y : 0.2f * z.y, z.z > 0.f ? z.z : 0.2f * z.z,
                        z.w > 0.f ? z.w : 0.2f * z.w);
  }
}

// dact -> dz = dact * leaky'(gamma * xhat + beta), in place
__global__ void __launch_bounds__(256) disc_lrelu_bwd_kernel(float* __restrict__ dact, const float* __restrict__ xhat, long long n,
                                                             int C, const float* __restrict__ gamma, const float* __restrict__ beta) {
  float4* d4 = reinterpret_cast<float4*>(dact);
  const float4* x4 = reinterpret_cast<const float4*>(xhat);
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n / 4; i += (long long)gridDim.x * 256) {
    const int c = (int)((i * 4) % C);
    const float4 xh = x4[i];
    const float4 ga = make_float4(gamma[c], gamma[c + 1], gamma[c + 2], gamma[c + 3]);
    const float4 be = make_float4(beta[c], beta[c + 1], beta[c + 2], beta[c + 3]);
    float4 d = d4[i];
    if (!(fmaf(ga.x, xh.x, be.x) > 0.f)) d.x *= 0.2f;
    if (!(fmaf(ga.y, xh.y, be.y) > 0.f)) d.y *= 0.2f;
    if (!(fmaf(ga.z, xh.z, be.z) > 0.f)) d.z *= 0.2f;
    if (!(fmaf(ga.w, xh.w, be.w) > 0.f)) d.w *= 0.2f;
    d4[i] = d;
  }
}

// act5 [B][7][8][8] -> feat [B][16] (c*2 + i; rows [0,4) and [3,7), all 8 columns) -> prob = sigmoid(w . feat + b)
__global__ void disc_tail_fwd_kernel(const float* __restrict__ act, const float* __restrict__ dw, const float* __restrict__ db,
                                     float* __restrict__ feat, float* __restrict__ prob, int B, int H, int W) {
  const int b = blockIdx.x, j = threadIdx.x;   // 16 threads
  __shared__ float f[16];
  const int c = j / 2, i = j % 2;
  const int h0 = (i * H) / 2, h1 = ((i + 1) * H + 1) / 2;
  float s = 0.f;
  for (int h = h0; h < h1; ++h)
    for (int w = 0; w < W; ++w) s += act[(((size_t)b * H + h) * W + w) * 8 + c];
  s /= (float)((h1 - h0) * W);
  f[j] = s;
  feat[b * 16 + j] = s;
  __syncthreads();
  if (j == 0) {
    float z = db[0];
    for (int k = 0; k < 16; ++k) z = fmaf(dw[k], f[k], z);
    prob[b] = 1.0f / (1.0f + expf(-z));
  }
}

// d_prob -> gradients of dense_d and d(act5)
__global__ void disc_tail_bwd_kernel(const float* __restrict__ dprob, const float* __restrict__ prob, const float* __restrict__ feat,
                                     const float* __restrict__ dw, float* g_dw, float* g_db, float* __restrict__ dact, int B,
                                     int H, int W) {
  const int b = blockIdx.x, j = threadIdx.x;   // 16 threads
  const float p = prob[b];
  const float dz = dprob[b] * p * (1.f - p);
  atomicAdd(g_dw + j, dz * feat[b * 16 + j]);
  if (j == 0) atomicAdd(g_db, dz);
  const int c = j / 2;
  // thread pair (c, i) writes the rows owned by bin i; row 3 of a 7-row map belongs to both bins: handled by c-major loop
  if (j % 2 == 0) {
    for (int h = 0; h < H; ++h) {
      float g = 0.f;
      for (int i = 0; i < 2; ++i) {
        const int h0 = (i * H) / 2, h1 = ((i + 1) * H + 1) / 2;
        if (h >= h0 && h < h1) g += dz * dw[c * 2 + i] / (float)((h1 - h0) * W);
      }
      for (int w = 0; w < W; ++w) dact[(((size_t)b * H + h) * W + w) * 8 + c] = g;
    }
  }
}

// BCE(prob, label) mean over B (log clamped at -100 as torch does) and d(loss)/d(prob)
__global__ void disc_bce_kernel(const float* __restrict__ prob, const float* __restrict__ label, int B, float* loss, float* dprob) {
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float p = prob[b], y = label[b];
    const float lp = fmaxf(logf(p), -100.f), l1p = fmaxf(logf(1.f - p), -100.f);
    acc -= y * lp + (1.f - y) * l1p;
    // derivative of the clamped logs: 0 where the clamp is active
    const float dlp = logf(p) > -100.f ? 1.f / p : 0.f, dl1p = logf(1.f - p) > -100.f ? -1.f / (1.f - p) : 0.f;
    dprob[b] = -(y * dlp + (1.f - y) * dl1p) / (float)B;
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < (int)blockDim.x / 32; ++k) t += red[k];
    *loss = t / (float)B;
  }
}

int grid_for(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g > 148 * 16 ? 148 * 16 : (g < 1 ? 1 : g));
}

}  // namespace
}  // namespace bsed

using namespace bsed;

extern "C" int bsed_disc_set_precision(bsed_handle h, int precision) {
  BSED_REQUIRE(h && (precision == BSED_PRECISION_FP32 || precision == BSED_PRECISION_TF32 || precision == BSED_PRECISION_TF32X3),
               "disc_set_precision: bad argument");
  h->disc_precision = precision;
  return BSED_OK;
}
extern "C" int64_t bsed_disc_param_count(void) { return disc_geom().n_params; }
extern "C" int64_t bsed_disc_bn_buffer_count(void) { return disc_geom().n_bn; }
extern "C" size_t bsed_disc_workspace_bytes(int B) { return B > 0 ? disc_ws(disc_geom(), B).total : 0; }

// prob [B] = D(d_input [B][313][256]).  train != 0: BatchNorm batch statistics (+ running-stat update, momentum 0.1)
// and everything backward needs is kept in the workspace.
extern "C" int bsed_disc_forward(bsed_handle h, const float* params, float* bn_buffers, int64_t* nbt, const float* d_input,
                                 int B, int train, float* prob, void* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(h && params && bn_buffers && d_input && prob && workspace, "disc_forward: null argument");
  BSED_REQUIRE(B >= 1, "disc_forward: B=%d", B);
  const DiscGeom g = disc_geom();
  const DiscWs w = disc_ws(g, B);
  if (workspace_bytes < w.total) {
    bsed_set_error("disc_forward: workspace %zu < %zu", workspace_bytes, w.total);
    return BSED_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  void* ws = workspace;
  const float* x = d_input;
  for (int l = 0; l < kDL; ++l) {
    const int Cin = kDC[l], Cout = kDC[l + 1], Ho = g.H[l + 1], Wo = g.W[l + 1];
    const long long M = (long long)B * Ho * Wo;
    float* col = wsp<float>(ws, w.col[l]);
    float* y = wsp<float>(ws, w.y[l]);
    float* act = wsp<float>(ws, w.act[l]);
    disc_prep_kernel<<<grid_for((long long)g.Kp[l] * g.Np[l]), 256, 0, st>>>(params + g.w_off[l], params + g.b_off[l],
                                                                             wsp<float>(ws, w.wt[l]), wsp<float>(ws, w.wk[l]),
                                                                             wsp<float>(ws, w.bp[l]), Cin, Cout, g.Kp[l], g.Np[l]);
    BSED_CHECK_LAUNCH();
    if (l == 0) {   // d_input arrives [B][W][H] with one channel
      disc_im2col_kernel<<<grid_for(M * g.Kp[l]), 256, 0, st>>>(x, col, B, g.H[l], g.W[l], Cin, Ho, Wo, g.Kp[l], 1);
      BSED_CHECK_LAUNCH();
    } else {        // channels-last: the float4 im2col shared with the ResNet tagger (resnet.cu)
      BSED_TRY(im2col_nhwc(x, col, B, g.H[l], g.W[l], Cin, 3, 3, 2, 2, 0, 0, Ho, Wo, g.Kp[l], st));
    }
    BSED_REQUIRE(M < (1ll << 31), "disc_forward: too many rows");
    const bool tc = h->disc_precision != BSED_PRECISION_FP32;
    const bool x3 = h->disc_precision == BSED_PRECISION_TF32X3;
    if (g.Np[l] == Cout) {
      if (tc) {   // y = col * Wk^T on tcgen05 (Wk [Np][Kp] is the K-major B operand)
        if (x3) {   // error-compensated: the weight operands of this layer (forward and data gradient) become (hi, lo) pairs
          BSED_TRY(split_hi_lo(wsp<float>(ws, w.wk[l]), wsp<float>(ws, w.wk_lo[l]), (long long)g.Np[l] * g.Kp[l], st));
          BSED_TRY(split_hi_lo(wsp<float>(ws, w.wt[l]), wsp<float>(ws, w.wt_lo[l]), (long long)g.Kp[l] * g.Np[l], st));
        }
        BSED_TRY(tc_gemm_nt(col, g.Kp[l], wsp<float>(ws, w.wk[l]), x3 ? wsp<float>(ws, w.wk_lo[l]) : nullptr, g.Kp[l], y, Cout, M,
                            Cout, g.Kp[l], wsp<float>(ws, w.bp[l]), 0, h->num_sms, st));
      } else
        BSED_TRY(gemm_nn(col, g.Kp[l], wsp<float>(ws, w.wt[l]), g.Np[l], y, Cout, (int)M, Cout, g.Kp[l], wsp<float>(ws, w.bp[l]), 0, st));
    } else {
      float* ypad = wsp<float>(ws, w.ypad);
      BSED_TRY(gemm_nn(col, g.Kp[l], wsp<float>(ws, w.wt[l]), g.Np[l], ypad, g.Np[l], (int)M, g.Np[l], g.Kp[l],
                       wsp<float>(ws, w.bp[l]), 0, st));
      disc_compact_kernel<<<grid_for(M * Cout), 256, 0, st>>>(ypad, y, M, Cout, g.Np[l], 0);
      BSED_CHECK_LAUNCH();
    }
    // BatchNorm (eps 1e-5, momentum 0.1: nn.BatchNorm2d defaults) + LeakyReLU
    Groups one;
    one.n = 1;
    for (int k = 0; k < kMaxGroups; ++k) one.first[k] = one.count[k] = 0;
    one.count[0] = 1;
    BNPtrs bn;
    float* mr = wsp<float>(ws, w.mr) + (size_t)l * 2 * 128;
    float* rmean[kMaxGroups];
    float* rvar[kMaxGroups];
    int64_t* nb[kMaxGroups];
    for (int k = 0; k < kMaxGroups; ++k) {
      bn.gamma[k] = params + g.bnw_off[l];
      bn.beta[k] = params + g.bnb_off[l];
      bn.mean[k] = mr;
      bn.rstd[k] = mr + 128;
      rmean[k] = bn_buffers + g.rm_off[l];
      rvar[k] = bn_buffers + g.rv_off[l];
      nb[k] = nbt ? nbt + l : nullptr;
    }
    if (train) {
      double* stats = wsp<double>(ws, w.stats);
      BSED_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 128 * 2, st));
      BSED_TRY(col_stats(y, nullptr, 0, one, M, Cout, stats, h->num_sms, st));
      BSED_TRY(bn_finalize_train(stats, one, M, Cout, 1e-5f, 0.1f, bn, rmean, rvar, nb, st));
    } else {
      BSED_TRY(bn_prepare_eval(one, Cout, 1e-5f, bn, rmean, rvar, st));
    }
    disc_bn_lrelu_fwd_kernel<<<grid_for(M * Cout), 256, 0, st>>>(y, act, M * Cout, Cout, bn.mean[0], bn.rstd[0], bn.gamma[0],
                                                                 bn.beta[0]);
    BSED_CHECK_LAUNCH();
    x = act;
  }
  disc_tail_fwd_kernel<<<B, 16, 0, st>>>(x, params + g.dense_w, params + g.dense_b, wsp<float>(ws, w.feat), prob, B, g.H[kDL],
                                         g.W[kDL]);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// Backward of the last train-mode bsed_disc_forward on the same workspace.  d_prob [B]: gradient w.r.t. the output
// probabilities; grads: flat buffer like params (overwritten unless accumulate); d_dinput [B][313][256] (may be NULL).
extern "C" int bsed_disc_backward(bsed_handle h, const float* params, const float* prob, const float* d_prob, int B,
                                  float* grads, int accumulate, float* d_dinput, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  BSED_REQUIRE(h && params && prob && d_prob && grads && workspace, "disc_backward: null argument");
  const DiscGeom g = disc_geom();
  const DiscWs w = disc_ws(g, B);
  if (workspace_bytes < w.total) {
    bsed_set_error("disc_backward: workspace %zu < %zu", workspace_bytes, w.total);
    return BSED_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  void* ws = workspace;
  if (!accumulate) BSED_CHECK_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * g.n_params, st));
  // tail: act[4] becomes d(act5)
  float* dact = wsp<float>(ws, w.act[kDL - 1]);
  disc_tail_bwd_kernel<<<B, 16, 0, st>>>(d_prob, prob, wsp<float>(ws, w.feat), params + g.dense_w, grads + g.dense_w,
                                         grads + g.dense_b, dact, B, g.H[kDL], g.W[kDL]);
  BSED_CHECK_LAUNCH();
  Groups one;
  one.n = 1;
  for (int k = 0; k < kMaxGroups; ++k) one.first[k] = one.count[k] = 0;
  one.count[0] = 1;
  for (int l = kDL - 1; l >= 0; --l) {
    const int Cin = kDC[l], Cout = kDC[l + 1], Ho = g.H[l + 1], Wo = g.W[l + 1];
    const long long M = (long long)B * Ho * Wo;
    float* col = wsp<float>(ws, w.col[l]);
    float* xhat = wsp<float>(ws, w.y[l]);
    float* dz = wsp<float>(ws, w.act[l]);   // holds d(act_l) on entry
    BNPtrs bn;
    float* mr = wsp<float>(ws, w.mr) + (size_t)l * 2 * 128;
    for (int k = 0; k < kMaxGroups; ++k) {
      bn.gamma[k] = params + g.bnw_off[l];
      bn.beta[k] = params + g.bnb_off[l];
      bn.mean[k] = mr;
      bn.rstd[k] = mr + 128;
    }
    disc_lrelu_bwd_kernel<<<grid_for(M * Cout), 256, 0, st>>>(dz, xhat, M * Cout, Cout, bn.gamma[0], bn.beta[0]);
    BSED_CHECK_LAUNCH();
    double* stats2 = wsp<double>(ws, w.stats2);
    BSED_CHECK_CUDA(cudaMemsetAsync(stats2, 0, sizeof(double) * 128 * 2, st));
    BSED_TRY(col_stats(dz, xhat, 1, one, M, Cout, stats2, h->num_sms, st));
    BSED_TRY(add_double_to_float(stats2, 2, grads + g.bnb_off[l], Cout, st));        // d_beta  = sum dz
    BSED_TRY(add_double_to_float(stats2 + 1, 2, grads + g.bnw_off[l], Cout, st));    // d_gamma = sum dz * xhat
    BSED_TRY(bn_bwd_apply(dz, xhat, stats2, one, M, Cout, bn, st));                  // dz -> dy (in place)
    // the conv bias gradient is identically zero (a train-mode BatchNorm follows every conv): grads[b_off] stays as is
    const float* dy = dz;
    int ldy = Cout;
    if (g.Np[l] != Cout) {
      float* ypad = wsp<float>(ws, w.ypad);
      disc_compact_kernel<<<grid_for(M * g.Np[l]), 256, 0, st>>>(dz, ypad, M, Cout, g.Np[l], 1);
      BSED_CHECK_LAUNCH();
      dy = ypad;
      ldy = g.Np[l];
    }
    // weight gradient: dWt [Kp][Np] = col^T dy
    float* dwt = wsp<float>(ws, w.dwt[l]);
    BSED_CHECK_CUDA(cudaMemsetAsync(dwt, 0, sizeof(float) * g.Kp[l] * g.Np[l], st));
    const bool tc = h->disc_precision != BSED_PRECISION_FP32;
    const bool x3 = h->disc_precision == BSED_PRECISION_TF32X3;
    if (tc && g.Kp[l] % 32 == 0 && g.Np[l] % 32 == 0 && g.Np[l] == Cout) {
      TcOperand Aop{col, g.Kp[l], 0, g.Kp[l]}, Bop{dy, ldy, 0, g.Np[l]};   // rows = im2col rows, as a (1, M, 1) grid
      BSED_TRY(tc_wgrad_ex(Aop, Bop, 1, (int)M, 1, 1, 0, dwt, g.Np[l], 1, 0, wsp<float>(ws, w.wgpart), w.wgpart_bytes,
                           h->num_sms, st));
    } else {
      BSED_TRY(gemm_tn(col, g.Kp[l], dy, ldy, dwt, g.Np[l], 1, g.Kp[l], g.Np[l], M, h->num_sms * 4, st));
    }
    disc_unpack_dw_kernel<<<grid_for(9LL * Cin * Cout), 256, 0, st>>>(dwt, grads + g.w_off[l], Cin, Cout, g.Np[l]);
    BSED_CHECK_LAUNCH();
    // data gradient: dcol = dy Wk (overwrites col), then scatter
    float* dx = l > 0 ? wsp<float>(ws, w.act[l - 1]) : d_dinput;
    if (dx) {
      if (tc && g.Np[l] == Cout) {
        // dcol = dy * Wk on tcgen05: B operand [N = k][K = co] = Wt, in column slices of <= 128
        for (int n0 = 0; n0 < g.Kp[l];) {
          const int rem = g.Kp[l] - n0;
          const int nw = rem >= 128 ? 128 : rem >= 64 ? 64 : rem >= 32 ? 32 : 16;
          BSED_TRY(tc_gemm_nt(dy, ldy, wsp<float>(ws, w.wt[l]) + (size_t)n0 * g.Np[l],
                              x3 ? wsp<float>(ws, w.wt_lo[l]) + (size_t)n0 * g.Np[l] : nullptr, g.Np[l], col + n0, g.Kp[l], M, nw,
                              g.Np[l], nullptr, 0, h->num_sms, st));
          n0 += nw;
        }
      } else {
        BSED_TRY(gemm_nn(dy, ldy, wsp<float>(ws, w.wk[l]), g.Kp[l], col, g.Kp[l], (int)M, g.Kp[l], g.Np[l], nullptr, 0, st));
      }
      if (l == 0) {
        const long long nin = (long long)B * g.H[l] * g.W[l] * Cin;
        BSED_CHECK_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * nin, st));
        disc_col2im_kernel<<<grid_for(M * 9 * Cin), 256, 0, st>>>(col, dx, B, g.H[l], g.W[l], Cin, Ho, Wo, g.Kp[l], 1);
        BSED_CHECK_LAUNCH();
      } else {   // gather form (every input pixel sums the <= 4 windows that read it): no atomics, no memset, fixed order
        BSED_TRY(col2im_nhwc(col, dx, B, g.H[l], g.W[l], Cin, 3, 3, 2, 2, 0, 0, Ho, Wo, g.Kp[l], 0, st));
      }
    }
  }
  return BSED_OK;
}

// mean BCE of prob [B] against label [B] (src/DA/cdan_frame.py:119 with weight 1) and its gradient
extern "C" int bsed_disc_bce(bsed_handle h, const float* prob, const float* label, int B, float* loss, float* d_prob,
                             void* stream) {
  BSED_REQUIRE(h && prob && label && loss && d_prob && B >= 1, "disc_bce: bad argument");
  disc_bce_kernel<<<1, 256, 0, as_stream(stream)>>>(prob, label, B, loss, d_prob);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}
