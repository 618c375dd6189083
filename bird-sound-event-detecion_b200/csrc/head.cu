// head.cu -- Predictor head (attention pooling), mean-teacher losses, optimiser + EMA, and the
// dropout backward mask.
//
// Predictor.forward                                                src/models/CRNN.py:559-577
//   strong = sigmoid(x W1^T + b1)
//   sof    = clamp(softmax_over_classes(x W2^T + b2), 1e-7, 1)
//   weak   = sum_t strong * sof / sum_t sof
// The two linears are one GEMM producing logits [B][T][ldl] (cols 0..C-1 dense, C..2C-1
// dense_softmax); this file holds everything after it.
#include <stdlib.h>

#include "launch.h"

namespace bsed {

constexpr int kMaxC = 20;

__device__ __forceinline__ void softmax_row(const float* l2, int C, float* sof_raw) {
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, l2[c]);
  float sum = 0.f;
  for (int c = 0; c < C; ++c) {
    sof_raw[c] = expf(l2[c] - mx);
    sum += sof_raw[c];
  }
  float inv = 1.0f / sum;
  for (int c = 0; c < C; ++c) sof_raw[c] *= inv;
}

// grid B, 256 threads; thread -> frames t = tid, tid + 256, ...
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ logits, float* __restrict__ strong,
                                                       float* __restrict__ weak, int T, int C, int ldl,
                                                       int inference) {
  const int b = blockIdx.x;
  float num[kMaxC], den[kMaxC];
  for (int c = 0; c < kMaxC; ++c) num[c] = den[c] = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float* l = logits + ((size_t)b * T + t) * ldl;
    float l2[kMaxC], sof[kMaxC];
    for (int c = 0; c < C; ++c) l2[c] = l[C + c];
    softmax_row(l2, C, sof);
    for (int c = 0; c < C; ++c) {
      float s = 1.0f / (1.0f + expf(-l[c]));
      float a = fminf(fmaxf(sof[c], 1e-7f), 1.0f);
      strong[((size_t)b * T + t) * C + c] = s;
      num[c] += s * a;
      den[c] += a;
    }
  }
  __shared__ float red[2][8][kMaxC];
  __shared__ float weak_s[kMaxC];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  for (int c = 0; c < C; ++c) {
    float n = warp_sum(num[c]), d = warp_sum(den[c]);
    if (lane == 0) {
      red[0][warp][c] = n;
      red[1][warp][c] = d;
    }
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float n = 0.f, d = 0.f;
    for (int w = 0; w < 8; ++w) {
      n += red[0][w][threadIdx.x];
      d += red[1][w][threadIdx.x];
    }
    float wk = n / d;
    weak_s[threadIdx.x] = wk;
    if (weak) weak[(size_t)b * C + threadIdx.x] = wk;
  }
  if (inference) {
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += blockDim.x)
      for (int c = 0; c < C; ++c)
        if (!(weak_s[c] > 0.5f)) strong[((size_t)b * T + t) * C + c] = 0.f;
  }
}

int head_forward(const float* logits, float* strong, float* weak, int B, int T, int C, int ldl, int inference,
                 cudaStream_t st) {
  BSED_REQUIRE(C <= kMaxC && ldl >= 2 * C, "head: C=%d ldl=%d", C, ldl);
  head_fwd_kernel<<<B, 256, 0, st>>>(logits, strong, weak, T, C, ldl, inference);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// d_logits from d_strong / d_weak
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ logits,
                                                       const float* __restrict__ strong,
                                                       const float* __restrict__ weak,
                                                       const float* __restrict__ d_strong,
                                                       const float* __restrict__ d_weak,
                                                       float* __restrict__ d_logits, int first_clip, int T, int C,
                                                       int ldl) {
  const int b = first_clip + blockIdx.x;
  // pass 1: den[c] = sum_t clamp(sof)
  float den[kMaxC];
  for (int c = 0; c < kMaxC; ++c) den[c] = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float* l = logits + ((size_t)b * T + t) * ldl;
    float l2[kMaxC], sof[kMaxC];
    for (int c = 0; c < C; ++c) l2[c] = l[C + c];
    softmax_row(l2, C, sof);
    for (int c = 0; c < C; ++c) den[c] += fminf(fmaxf(sof[c], 1e-7f), 1.0f);
  }
  __shared__ float red[8][kMaxC];
  __shared__ float den_s[kMaxC], dw_s[kMaxC], wk_s[kMaxC];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  for (int c = 0; c < C; ++c) {
    float d = warp_sum(den[c]);
    if (lane == 0) red[warp][c] = d;
  }
  __syncthreads();
  if (threadIdx.x < C) {
    float d = 0.f;
    for (int w = 0; w < 8; ++w) d += red[w][threadIdx.x];
    den_s[threadIdx.x] = d;
    dw_s[threadIdx.x] = d_weak ? d_weak[(size_t)b * C + threadIdx.x] : 0.f;
    wk_s[threadIdx.x] = weak[(size_t)b * C + threadIdx.x];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const size_t rt = (size_t)b * T + t;
    const float* l = logits + rt * ldl;
    float* dl = d_logits + rt * ldl;
    float l2[kMaxC], sofr[kMaxC], dsr[kMaxC];
    for (int c = 0; c < C; ++c) l2[c] = l[C + c];
    softmax_row(l2, C, sofr);
    float dot = 0.f;
    for (int c = 0; c < C; ++c) {
      float s = strong[rt * C + c];
      float a = fminf(fmaxf(sofr[c], 1e-7f), 1.0f);
      float inv_den = 1.0f / den_s[c];
      float ds = (d_strong ? d_strong[rt * C + c] : 0.f) + dw_s[c] * a * inv_den;
      dl[c] = ds * s * (1.f - s);
      float dsof = dw_s[c] * (s - wk_s[c]) * inv_den;
      bool pass = sofr[c] >= 1e-7f && sofr[c] <= 1.0f;
      dsr[c] = pass ? dsof : 0.f;
      dot += dsr[c] * sofr[c];
    }
    for (int c = 0; c < C; ++c) dl[C + c] = sofr[c] * (dsr[c] - dot);
    for (int c = 2 * C; c < ldl; ++c) dl[c] = 0.f;
  }
}

int head_backward(const float* logits, const float* strong, const float* weak, const float* d_strong,
                  const float* d_weak, float* d_logits, int first_clip, int n_clips, int T, int C, int ldl,
                  cudaStream_t st) {
  BSED_REQUIRE(C <= kMaxC && ldl >= 2 * C, "head: C=%d ldl=%d", C, ldl);
  head_bwd_kernel<<<n_clips, 256, 0, st>>>(logits, strong, weak, d_strong, d_weak, d_logits, first_clip, T, C, ldl);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// d[i] = (d[i] + extra[i]) * keep(first_elem + i) / (1 - p)
__global__ void dropout_bwd_mask_kernel(float* d, const float* extra, long long first_elem, long long n,
                                        DropKey dkey, uint32_t thresh, float inv_keep) {
  const uint32_t key = dkey.get();
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = d[i] + (extra ? extra[i] : 0.f);
  if (thresh) v = bsed_keep((uint32_t)(first_elem + i), key, thresh) ? v * inv_keep : 0.f;
  d[i] = v;
}

__global__ void add_f32_kernel(float* dst, const float* src, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}
int add_f32(float* dst, const float* src, long long n, cudaStream_t st) {
  if (n <= 0) return BSED_OK;
  add_f32_kernel<<<ceil_div(n, 256), 256, 0, st>>>(dst, src, n);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

__global__ void scale_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n, float alpha) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = alpha * src[i];
}
int scale_f32(float* dst, const float* src, long long n, float alpha, cudaStream_t st) {
  if (n <= 0) return BSED_OK;
  scale_f32_kernel<<<ceil_div(n, 256), 256, 0, st>>>(dst, src, n, alpha);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

int dropout_bwd_mask(float* d, const float* extra, long long first_elem, long long n, DropKey key,
                     uint32_t thresh, float inv_keep, cudaStream_t st) {
  if (n <= 0) return BSED_OK;
  dropout_bwd_mask_kernel<<<ceil_div(n, 256), 256, 0, st>>>(d, extra, first_elem, n, key, thresh, inv_keep);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// feature-pyramid merge of CRNN_fpn (src/models/CRNN.py:323-328):  torch.cat((x, nn.Upsample((Ta,1), mode='bilinear',
// align_corners=True)(y)), 1) on (B, 256, T, 1) tensors, here time-major (B, T, 256).  Along time (the width axis
// has one element): src = t * (Tb-1)/(Ta-1), i0 = floor(src), i1 = min(i0+1, Tb-1), w1 = src - i0, all in fp32 as
// ATen's upsample_bilinear2d does for float inputs.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fpn_src(int t, float scale, int Tb, int* i0, int* i1, float* w1) {
  const float src = scale * (float)t;
  const int k = (int)src;
  *i0 = k;
  *i1 = k + (k < Tb - 1 ? 1 : 0);
  *w1 = src - (float)k;
}

__global__ void __launch_bounds__(256) fpn_cat_upsample_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                   float* __restrict__ cat, int Ta, int Tb, float scale,
                                                                   long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4 of cat
  if (id >= total) return;
  const int q = (int)(id % 128);
  const long long row = id / 128;
  const int t = (int)(row % Ta);
  const long long clip = row / Ta;
  float4 v;
  if (q < 64) {
    v = *reinterpret_cast<const float4*>(a + row * 256 + q * 4);
  } else {
    int i0, i1;
    float w1;
    fpn_src(t, scale, Tb, &i0, &i1, &w1);
    const float w0 = 1.f - w1;
    const float4 x0 = *reinterpret_cast<const float4*>(b + (clip * Tb + i0) * 256 + (q - 64) * 4);
    const float4 x1 = *reinterpret_cast<const float4*>(b + (clip * Tb + i1) * 256 + (q - 64) * 4);
    v = make_float4(w0 * x0.x + w1 * x1.x, w0 * x0.y + w1 * x1.y, w0 * x0.z + w1 * x1.z, w0 * x0.w + w1 * x1.w);
  }
  *reinterpret_cast<float4*>(cat + row * 512 + q * 4) = v;
}

int fpn_cat_upsample_fwd(const float* a, const float* b, float* cat, int B, int Ta, int Tb, cudaStream_t st) {
  BSED_REQUIRE(Ta >= 2 && Tb >= 1, "fpn_cat_upsample: Ta=%d Tb=%d", Ta, Tb);
  const long long total = (long long)B * Ta * 128;
  const float scale = (float)(Tb - 1) / (float)(Ta - 1);
  fpn_cat_upsample_fwd_kernel<<<ceil_div(total, 256), 256, 0, st>>>(a, b, cat, Ta, Tb, scale, total);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// da[b][t][c] = dcat[b][t][c] ; db[b][u][c] = sum_t w(t -> u) * dcat[b][t][256 + c]   (gather over the few t that read u)
__global__ void __launch_bounds__(256) fpn_cat_upsample_bwd_kernel(const float* __restrict__ dcat, float* __restrict__ da,
                                                                   float* __restrict__ db, int Ta, int Tb, float scale,
                                                                   long long total_a, long long total) {
  const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= total) return;
  if (id < total_a) {   // float4 of da
    const int q = (int)(id % 64);
    const long long row = id / 64;
    *reinterpret_cast<float4*>(da + row * 256 + q * 4) = *reinterpret_cast<const float4*>(dcat + row * 512 + q * 4);
    return;
  }
  const long long jd = id - total_a;   // float4 of db
  const int q = (int)(jd % 64);
  const long long row = jd / 64;
  const int u = (int)(row % Tb);
  const long long clip = row / Tb;
  // candidates: every t whose source interval [i0, i1] can contain u, i.e. src in (u - 1, u + 1)
  int t_lo = 0, t_hi = Ta - 1;
  if (scale > 0.f) {
    t_lo = (int)floorf((float)(u - 1) / scale) - 1;
    t_hi = (int)ceilf((float)(u + 1) / scale) + 1;
    if (t_lo < 0) t_lo = 0;
    if (t_hi > Ta - 1) t_hi = Ta - 1;
  }
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = t_lo; t <= t_hi; ++t) {
    int i0, i1;
    float w1;
    fpn_src(t, scale, Tb, &i0, &i1, &w1);
    float w = 0.f;
    if (i0 == u) w += 1.f - w1;
    if (i1 == u) w += w1;
    if (w != 0.f) {
      const float4 g = *reinterpret_cast<const float4*>(dcat + (clip * Ta + t) * 512 + 256 + q * 4);
      acc.x += w * g.x;
      acc.y += w * g.y;
      acc.z += w * g.z;
      acc.w += w * g.w;
    }
  }
  *reinterpret_cast<float4*>(db + row * 256 + q * 4) = acc;
}

int fpn_cat_upsample_bwd(const float* dcat, float* da, float* db, int B, int Ta, int Tb, cudaStream_t st) {
  BSED_REQUIRE(Ta >= 2 && Tb >= 1, "fpn_cat_upsample: Ta=%d Tb=%d", Ta, Tb);
  const long long total_a = (long long)B * Ta * 64, total_b = (long long)B * Tb * 64;
  const float scale = (float)(Tb - 1) / (float)(Ta - 1);
  fpn_cat_upsample_bwd_kernel<<<ceil_div(total_a + total_b, 256), 256, 0, st>>>(dcat, da, db, Ta, Tb, scale, total_a,
                                                                                total_a + total_b);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// mean-teacher losses (src/main.py:376,405,434,439-449) and their gradients
//   BCELoss: -(y log x + (1-y) log(1-x)), logs clamped at -100, mean; grad (x-y)/max(x(1-x),1e-12)/N
//   MSELoss: mean (a-b)^2; grad 2(a-b)/N
// grid: B clips; losses[4] must be zero on entry.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float bce_term(float x, float y) {
  float lx = fmaxf(logf(x), -100.f), l1x = fmaxf(logf(1.f - x), -100.f);
  return -(y * lx + (1.f - y) * l1x);
}
__device__ __forceinline__ float bce_grad(float x, float y) { return (x - y) / fmaxf(x * (1.f - x), 1e-12f); }

__global__ void __launch_bounds__(256) mt_loss_kernel(const float* __restrict__ strong, const float* __restrict__ weak,
                                                      int T, int C, int syn_first, int syn_n,
                                                      const float* __restrict__ syn_target, int real_first, int real_n,
                                                      const float* __restrict__ strong_ema,
                                                      const float* __restrict__ weak_ema, float cons_w_arg,
                                                      const bsed_step_state* __restrict__ ss, float* losses,
                                                      float* __restrict__ d_strong, float* __restrict__ d_weak) {
  const float cons_w = ss ? ss->cons_w : cons_w_arg;
  const int b = blockIdx.x;
  const bool is_syn = b >= syn_first && b < syn_first + syn_n;
  const bool is_real = b >= real_first && b < real_first + real_n;
  const int TC = T * C;
  float acc_s = 0.f, acc_w = 0.f;
  __shared__ float tmax[kMaxC];
  __shared__ float red[8];
  if (threadIdx.x < kMaxC) tmax[threadIdx.x] = 0.f;
  __syncthreads();
  if (is_syn) {
    const float* tg = syn_target + (size_t)(b - syn_first) * TC;
    const float inv_n = 1.0f / ((float)syn_n * (float)TC);
    for (int i = threadIdx.x; i < TC; i += blockDim.x) {
      float x = strong[(size_t)b * TC + i], y = tg[i];
      acc_s += bce_term(x, y);
      d_strong[(size_t)b * TC + i] = bce_grad(x, y) * inv_n;
      if (y > 0.f) atomicMax(reinterpret_cast<int*>(&tmax[i % C]), __float_as_int(y));
    }
    __syncthreads();
    if (threadIdx.x < C) {
      float x = weak[(size_t)b * C + threadIdx.x], y = tmax[threadIdx.x];
      const float inv_nw = 1.0f / ((float)syn_n * (float)C);
      acc_w = bce_term(x, y);
      d_weak[(size_t)b * C + threadIdx.x] = bce_grad(x, y) * inv_nw;
    }
  } else if (is_real) {
    const float* se = strong_ema + (size_t)(b - real_first) * TC;
    const float inv_n = 1.0f / ((float)real_n * (float)TC);
    for (int i = threadIdx.x; i < TC; i += blockDim.x) {
      float d = strong[(size_t)b * TC + i] - se[i];
      acc_s += d * d;
      d_strong[(size_t)b * TC + i] = 2.f * cons_w * d * inv_n;
    }
    if (threadIdx.x < C) {
      const float inv_nw = 1.0f / ((float)real_n * (float)C);
      float d = weak[(size_t)b * C + threadIdx.x] - weak_ema[(size_t)(b - real_first) * C + threadIdx.x];
      acc_w = d * d;
      d_weak[(size_t)b * C + threadIdx.x] = 2.f * cons_w * d * inv_nw;
    }
  } else {
    for (int i = threadIdx.x; i < TC; i += blockDim.x) d_strong[(size_t)b * TC + i] = 0.f;
    if (threadIdx.x < C) d_weak[(size_t)b * C + threadIdx.x] = 0.f;
    return;
  }
  // block reduce the two partial sums
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  float s = warp_sum(acc_s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot_s = 0.f;
  if (threadIdx.x == 0)
    for (int w = 0; w < 8; ++w) tot_s += red[w];
  __syncthreads();
  float wv = warp_sum(acc_w);
  if (lane == 0) red[warp] = wv;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot_w = 0.f;
    for (int w = 0; w < 8; ++w) tot_w += red[w];
    if (is_syn) {
      atomicAdd(&losses[0], tot_s / ((float)syn_n * (float)TC));
      atomicAdd(&losses[1], tot_w / ((float)syn_n * (float)C));
    } else {
      atomicAdd(&losses[2], cons_w * tot_s / ((float)real_n * (float)TC));
      atomicAdd(&losses[3], cons_w * tot_w / ((float)real_n * (float)C));
    }
  }
}

int mt_loss(const float* strong, const float* weak, int B, int T, int C, int syn_first, int syn_n,
            const float* syn_target, int real_first, int real_n, const float* strong_ema, const float* weak_ema,
            float cons_w, const bsed_step_state* ss, float* losses, float* d_strong, float* d_weak, cudaStream_t st) {
  BSED_REQUIRE(C <= kMaxC, "mt_loss: C=%d", C);
  BSED_CHECK_CUDA(cudaMemsetAsync(losses, 0, 4 * sizeof(float), st));
  mt_loss_kernel<<<B, 256, 0, st>>>(strong, weak, T, C, syn_first, syn_n, syn_target, real_first, real_n,
                                    strong_ema, weak_ema, cons_w, ss, losses, d_strong, d_weak);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// generic loss terms (shift-consistency training, src/main_baseline.py:372-529): one launch per term, so the
// gradient accumulation order is fixed.  grid = clips of the term.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) loss_term_kernel(const float* __restrict__ strong, const float* __restrict__ weak,
                                                        int T, int C, bsed_loss_term tm, float* losses,
                                                        float* __restrict__ d_strong, float* __restrict__ d_weak) {
  const int b = blockIdx.x;
  const int pb = tm.pred_first + b;
  const int TC = T * C;
  const bool is_bce = tm.kind == BSED_LOSS_BCE_STRONG || tm.kind == BSED_LOSS_BCE_WEAK;
  float acc = 0.f;
  __shared__ float tmax[kMaxC];
  __shared__ float red[8];
  if (tm.kind == BSED_LOSS_BCE_STRONG || tm.kind == BSED_LOSS_MSE_STRONG) {
    const int roll = tm.roll ? tm.roll[b] : 0;
    const float* ref = tm.ref + (size_t)b * TC;
    const float gscale = tm.grad_weight / ((float)tm.n_clips * (float)TC);
    for (int i = threadIdx.x; i < TC; i += blockDim.x) {
      const int t = i / C, c = i - t * C;
      int ts = (t - roll) % T;                 // torch.roll(ref, roll, 0)[t] = ref[(t - roll) mod T]
      if (ts < 0) ts += T;
      const float x = strong[(size_t)pb * TC + i], y = ref[(size_t)ts * C + c];
      float g;
      if (is_bce) {
        acc += bce_term(x, y);
        g = bce_grad(x, y);
      } else {
        const float d = x - y;
        acc += d * d;
        g = 2.f * d;
      }
      if (tm.grad_weight != 0.f) d_strong[(size_t)pb * TC + i] += g * gscale;
    }
  } else {
    if (tm.ref_is_strong) {                    // weak target = max over time of a strong target (syn_target.max(-2)[0])
      if (threadIdx.x < kMaxC) tmax[threadIdx.x] = 0.f;
      __syncthreads();
      const float* ref = tm.ref + (size_t)b * TC;
      for (int i = threadIdx.x; i < TC; i += blockDim.x) {
        const float y = ref[i];
        if (y > 0.f) atomicMax(reinterpret_cast<int*>(&tmax[i % C]), __float_as_int(y));
      }
      __syncthreads();
    }
    if (threadIdx.x < C) {
      const float x = weak[(size_t)pb * C + threadIdx.x];
      const float y = tm.ref_is_strong ? tmax[threadIdx.x] : tm.ref[(size_t)b * C + threadIdx.x];
      const float gscale = tm.grad_weight / ((float)tm.n_clips * (float)C);
      float g;
      if (is_bce) {
        acc = bce_term(x, y);
        g = bce_grad(x, y);
      } else {
        const float d = x - y;
        acc = d * d;
        g = 2.f * d;
      }
      if (tm.grad_weight != 0.f) d_weak[(size_t)pb * C + threadIdx.x] += g * gscale;
    }
  }
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
  const float s = warp_sum(acc);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    const float n = (float)tm.n_clips * (float)((tm.kind == BSED_LOSS_BCE_STRONG || tm.kind == BSED_LOSS_MSE_STRONG) ? TC : C);
    atomicAdd(&losses[tm.slot], tm.weight * tot / n);
  }
}

int loss_terms(const float* strong, const float* weak, int B, int T, int C, const bsed_loss_term* terms, int n_terms,
               float* losses, int n_slots, float* d_strong, float* d_weak, cudaStream_t st) {
  BSED_REQUIRE(C <= kMaxC, "loss_terms: C=%d", C);
  BSED_CHECK_CUDA(cudaMemsetAsync(losses, 0, sizeof(float) * n_slots, st));
  BSED_CHECK_CUDA(cudaMemsetAsync(d_strong, 0, sizeof(float) * (size_t)B * T * C, st));
  BSED_CHECK_CUDA(cudaMemsetAsync(d_weak, 0, sizeof(float) * (size_t)B * C, st));
  for (int i = 0; i < n_terms; ++i) {
    const bsed_loss_term& tm = terms[i];
    BSED_REQUIRE(tm.kind >= 0 && tm.kind <= 3 && tm.ref && tm.n_clips >= 1 && tm.pred_first >= 0 &&
                     tm.pred_first + tm.n_clips <= B && tm.slot >= 0 && tm.slot < n_slots,
                 "loss_terms: term %d is malformed", i);
    loss_term_kernel<<<tm.n_clips, 256, 0, st>>>(strong, weak, T, C, tm, losses, d_strong, d_weak);
    BSED_CHECK_LAUNCH();
  }
  return BSED_OK;
}

// out[b][t][f] = x[b][(t - shift_t[b]) mod T][(f - shift_f[b]) mod F]   (torch.roll along time, then frequency)
__global__ void __launch_bounds__(256) roll_clips_kernel(const float* __restrict__ x, const int* __restrict__ shift_t,
                                                         const int* __restrict__ shift_f, float* __restrict__ out, int T,
                                                         int F) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)T * F) return;
  const int t = (int)(i / F), f = (int)(i - (long long)t * F);
  int ts = (t - (shift_t ? shift_t[b] : 0)) % T, fs = (f - (shift_f ? shift_f[b] : 0)) % F;
  if (ts < 0) ts += T;
  if (fs < 0) fs += F;
  out[(size_t)b * T * F + i] = x[((size_t)b * T + ts) * F + fs];
}

int roll_clips(const float* x, const int* shift_t, const int* shift_f, float* out, int B, int T, int F, cudaStream_t st) {
  dim3 grid(ceil_div((long long)T * F, 256), B);
  roll_clips_kernel<<<grid, 256, 0, st>>>(x, shift_t, shift_f, out, T, F);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// optimiser + EMA over flat buffers
// ---------------------------------------------------------------------------------------------
struct OptScalars {
  int kind;
  float lr, beta1, beta2, eps, wd, momentum, grad_scale;
  float step_size;   // lr / (1 - beta1^t)
  float bc2_sqrt;    // sqrt(1 - beta2^t)
  float ema_a;       // min(1 - 1/(ema_step+1), alpha)
  float ema_b;       // (float)(1 - a), as python evaluates (1. - alpha) in double
  int first_step;    // SGD: momentum buffer initialised with the gradient
  int has_ema;
};

// per-iteration scalars from the device-resident step state (bsed_set_step_state), when one is installed
__device__ __forceinline__ void opt_from_state(OptScalars& o, const bsed_step_state* ss) {
  if (!ss) return;
  o.lr = ss->lr;
  o.step_size = ss->step_size;
  o.bc2_sqrt = ss->bc2_sqrt;
  o.ema_a = ss->ema_a;
  o.ema_b = ss->ema_b;
  o.first_step = ss->first_step;
}

__device__ __forceinline__ void opt_update(long long i, float grad, float* __restrict__ p, float* __restrict__ m,
                                           float* __restrict__ v, float* __restrict__ ema, const OptScalars& o) {
  float w = p[i];
  if (o.kind == 0) {
    if (o.wd != 0.f) grad = fmaf(o.wd, w, grad);
    float mi = m[i], vi = v[i];
    mi = mi + (grad - mi) * (1.f - o.beta1);            // exp_avg.lerp_(grad, 1 - beta1)
    vi = vi * o.beta2 + (1.f - o.beta2) * grad * grad;  // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    float denom = sqrtf(vi) / o.bc2_sqrt + o.eps;
    w = w - o.step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  } else {
    if (o.wd != 0.f) grad = fmaf(o.wd, w, grad);
    float buf = o.first_step ? grad : o.momentum * m[i] + grad;
    m[i] = buf;
    float upd = grad + o.momentum * buf;  // nesterov
    w = w - o.lr * upd;
  }
  p[i] = w;
  if (o.has_ema) ema[i] = __fadd_rn(__fmul_rn(ema[i], o.ema_a), __fmul_rn(w, o.ema_b));
}

__global__ void __launch_bounds__(256) opt_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                      float* __restrict__ m, float* __restrict__ v,
                                                      float* __restrict__ ema, long long n, OptScalars o,
                                                      const bsed_step_state* __restrict__ ss) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  opt_from_state(o, ss);
  opt_update(i, g[i] * o.grad_scale, p, m, v, ema, o);
}

static int make_opt_scalars(const bsed_opt_cfg* cfg, bool has_ema, OptScalars* out) {
  BSED_REQUIRE(cfg && (cfg->kind == 0 || cfg->kind == 1), "opt: bad cfg");
  BSED_REQUIRE(cfg->step >= 1, "opt: step must be >= 1");
  OptScalars o;
  o.kind = cfg->kind;
  o.lr = cfg->lr;
  o.beta1 = cfg->beta1;
  o.beta2 = cfg->beta2;
  o.eps = cfg->eps;
  o.wd = cfg->weight_decay;
  o.momentum = cfg->momentum;
  o.grad_scale = cfg->grad_scale;
  double bc1 = 1.0 - pow((double)cfg->beta1, (double)cfg->step);
  double bc2 = 1.0 - pow((double)cfg->beta2, (double)cfg->step);
  o.step_size = (float)((double)cfg->lr / bc1);
  o.bc2_sqrt = (float)sqrt(bc2);
  double a = 1.0 - 1.0 / ((double)cfg->ema_step + 1.0);
  if (a > (double)cfg->ema_alpha) a = (double)cfg->ema_alpha;
  o.ema_a = (float)a;
  o.ema_b = (float)(1.0 - a);
  o.first_step = cfg->step == 1;
  o.has_ema = has_ema;
  *out = o;
  return BSED_OK;
}

// One thread: next iteration's counters, dropout keys and derived scalars -- the host formulas of bsed_mix_key,
// make_opt_scalars and utilities/ramps.py:exp_rampup in double precision.
__global__ void step_state_advance_kernel(bsed_step_state* s, bsed_step_cfg c) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const long long g = s->global_step + 1;
  const long long t = s->opt_step + 1;
  s->global_step = g;
  s->opt_step = t;
  s->dp_epoch += 1;
  const unsigned long long dstep = (unsigned long long)(c.key_mul * g + c.key_add);
  for (int stream = 0; stream < 16; ++stream) {
    unsigned long long z = c.dropout_seed * 0x9E3779B97F4A7C15ull + dstep * 0xD1B54A32D192ED03ull +
                           (unsigned long long)stream * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    s->keys[stream] = (uint32_t)(z >> 32);
  }
  double ramp = 1.0;
  if (c.rampup_length > 0) {
    double cur = (double)g;
    if (cur < 0) cur = 0;
    if (cur > (double)c.rampup_length) cur = (double)c.rampup_length;
    const double phase = 1.0 - cur / (double)c.rampup_length;
    ramp = exp(-5.0 * phase * phase);
  }
  s->cons_w = (float)((double)c.max_consistency_cost * ramp);
  const double bc1 = 1.0 - pow((double)c.beta1, (double)t);
  const double bc2 = 1.0 - pow((double)c.beta2, (double)t);
  s->step_size = (float)((double)s->lr / bc1);
  s->bc2_sqrt = (float)sqrt(bc2);
  double a = 1.0 - 1.0 / ((double)(g + 1) + 1.0);
  if (a > (double)c.ema_alpha) a = (double)c.ema_alpha;
  s->ema_a = (float)a;
  s->ema_b = (float)(1.0 - a);
  s->first_step = t == 1;
}

int step_state_advance(bsed_step_state* state, const bsed_step_cfg* cfg, cudaStream_t st) {
  step_state_advance_kernel<<<1, 32, 0, st>>>(state, *cfg);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

int opt_ema_step(float* params, const float* grads, float* m, float* v, float* ema, long long n,
                 const bsed_opt_cfg* cfg, const bsed_step_state* ss, cudaStream_t st) {
  OptScalars o;
  BSED_TRY(make_opt_scalars(cfg, ema != nullptr, &o));
  opt_ema_kernel<<<ceil_div(n, 256), 256, 0, st>>>(params, grads, m, v, ema, n, o, ss);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// Data-parallel step in ONE kernel: gradient reduce-scatter over NVLink peer memory + optimiser + EMA + all-gather.
//
// Every rank maps every peer's flat gradient, parameter and EMA buffers and flag block (CUDA IPC).  Per step (`epoch`
// increases by one), rank r owns the slice [r * chunk, (r + 1) * chunk) of the flat buffers:
//   1. arrive: block 0 stores `epoch` into every peer's arrive[my_rank] with release.sys semantics -- the gradient
//      kernels of this rank ran earlier on the same stream, so its buffer is complete and nothing on this rank reads
//      the parameters any more; every CTA then spins on its OWN arrive[] flags (local memory) until all peers arrived;
//   2. reduce + update + publish: each thread sums element i of its rank's slice over the peers IN RANK ORDER (L1-bypassing
//      loads from peer memory), applies Adam / SGD-Nesterov and the EMA to the local copy, and stores the new parameter
//      and EMA value into every peer's buffers -- replicas are bit-identical by construction, the optimiser state
//      (m, v) is only ever touched inside the owner's slice, and each rank moves n gradient floats in and
//      2 n (world - 1) / world parameter floats out over NVLink instead of the world * n of a one-shot all-reduce;
//   3. depart: the last CTA to finish (device-scope counter) stores `epoch` into every peer's done[my_rank] and waits
//      until all peers have done the same: all pushes have landed and nobody still reads this rank's gradients.
// A spin that waits longer than the timeout (BSED_DP_TIMEOUT_S, default 60 s) is FATAL: it raises the sticky error flag
// flags[33], the rank applies nothing further and never signals depart (so its peers run into the same timeout instead of
// reducing gradients this rank may be overwriting), and the host raises on its next health check
// (utilities/shard.py: FusedDataParallel.check).
// flag block (int32[64], zero-initialised by the host): arrive[0..15], done[16..31], counter [32], error [33].
// ---------------------------------------------------------------------------------------------
constexpr int kDpMaxWorld = 8;
struct DpPeers {
  const float* grads[kDpMaxWorld];
  float* params[kDpMaxWorld];
  float* ema[kDpMaxWorld];
  int* flags[kDpMaxWorld];
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {   // peer memory: never from a stale L1 line
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until flags[idx] >= epoch for idx in [base, base + world); false on timeout or when the error flag is up
__device__ bool dp_wait(const int* flags, int base, int world, int epoch, unsigned long long timeout_ns) {
  const unsigned long long t0 = global_ns();
  for (int r = 0; r < world; ++r) {
    while (ld_acquire_sys(flags + base + r) < epoch) {
      if (global_ns() - t0 > timeout_ns || ld_acquire_sys(flags + 33) != 0) return false;
      __nanosleep(100);
    }
  }
  return true;
}

__global__ void __launch_bounds__(256) dp_opt_ema_kernel(DpPeers peers, int rank, int world, int epoch, float* __restrict__ p,
                                                         float* __restrict__ m, float* __restrict__ v,
                                                         float* __restrict__ ema, long long lo, long long hi, OptScalars o,
                                                         unsigned long long timeout_ns, const bsed_step_state* __restrict__ ss) {
  opt_from_state(o, ss);
  if (ss) epoch = (int)ss->dp_epoch;
  int* my = peers.flags[rank];
  __shared__ int ok_s;
  if (threadIdx.x == 0) {
    const bool healthy = ld_acquire_sys(my + 33) == 0;      // a rank that timed out once stays out
    if (blockIdx.x == 0 && healthy) {
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(peers.flags[r] + rank, epoch);       // arrive
    }
    ok_s = healthy && dp_wait(my, 0, world, epoch, timeout_ns) ? 1 : 0;
    if (!ok_s) atomicExch(my + 33, 1);
  }
  __syncthreads();
  if (ok_s) {
    for (long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
      float gr[kDpMaxWorld];
#pragma unroll
      for (int r = 0; r < kDpMaxWorld; ++r)
        if (r < world) gr[r] = r == rank ? peers.grads[r][i] : ld_peer(peers.grads[r] + i);
      float g = 0.f;
#pragma unroll
      for (int r = 0; r < kDpMaxWorld; ++r)
        if (r < world) g += gr[r];
      opt_update(i, g * o.grad_scale, p, m, v, ema, o);
      const float pw = p[i], ew = o.has_ema ? ema[i] : 0.f;
#pragma unroll
      for (int r = 0; r < kDpMaxWorld; ++r)
        if (r < world && r != rank) {
          st_peer(peers.params[r] + i, pw);
          if (o.has_ema) st_peer(peers.ema[r] + i, ew);
        }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                    // this CTA's pushes are ordered before the counter increment
    const int done = atomicAdd(my + 32, 1);
    if (done == (int)gridDim.x - 1) {          // last CTA of this rank: every push is out, every peer read is complete
      my[32] = 0;
      __threadfence_system();
      if (ld_acquire_sys(my + 33) == 0) {      // no depart after a timeout anywhere on this rank
        for (int r = 0; r < world; ++r) st_release_sys(peers.flags[r] + 16 + rank, epoch);   // depart
        if (!dp_wait(my, 16, world, epoch, timeout_ns)) atomicExch(my + 33, 1);
      }
    }
  }
}

int dp_opt_ema_step(int rank, int world, const float* const* peer_grads, float* const* peer_params, float* const* peer_ema,
                    int* const* peer_flags, long long epoch, float* m, float* v, long long n, const bsed_opt_cfg* cfg,
                    const bsed_step_state* ss, int num_sms, cudaStream_t st) {
  BSED_REQUIRE(world >= 1 && world <= kDpMaxWorld && rank >= 0 && rank < world, "dp_opt: rank %d of %d", rank, world);
  BSED_REQUIRE(epoch >= 1 && epoch < (1ll << 31), "dp_opt: epoch %lld", epoch);
  const bool has_ema = peer_ema != nullptr && peer_ema[rank] != nullptr;
  OptScalars o;
  BSED_TRY(make_opt_scalars(cfg, has_ema, &o));
  DpPeers peers;
  for (int r = 0; r < kDpMaxWorld; ++r) {
    peers.grads[r] = r < world ? peer_grads[r] : nullptr;
    peers.params[r] = r < world ? peer_params[r] : nullptr;
    peers.ema[r] = r < world && has_ema ? peer_ema[r] : nullptr;
    peers.flags[r] = r < world ? peer_flags[r] : nullptr;
    BSED_REQUIRE(r >= world || (peers.grads[r] && peers.params[r] && peers.flags[r] && (!has_ema || peers.ema[r])),
                 "dp_opt: peer %d not mapped", r);
  }
  long long chunk = (n + world - 1) / world;
  chunk = (chunk + 3) / 4 * 4;
  long long lo = (long long)rank * chunk, hi = lo + chunk;
  if (lo > n) lo = n;
  if (hi > n) hi = n;
  int grid = num_sms * 2;
  const long long need = (hi - lo + 255) / 256;
  if (grid > need) grid = (int)(need > 0 ? need : 1);
  static double timeout_s = -1.0;
  if (timeout_s < 0) {
    const char* e = getenv("BSED_DP_TIMEOUT_S");
    timeout_s = e && atof(e) > 0 ? atof(e) : 60.0;
  }
  dp_opt_ema_kernel<<<grid, 256, 0, st>>>(peers, rank, world, (int)epoch, peers.params[rank], m, v, peers.ema[rank], lo, hi, o,
                                          (unsigned long long)(timeout_s * 1e9), ss);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

__global__ void ema_buffers_kernel(const float* __restrict__ src, float* __restrict__ ema, long long n,
                                   const int64_t* nbt, int64_t* ema_nbt, int n_nbt, float a, float bcoef,
                                   const bsed_step_state* __restrict__ ss) {
  if (ss) {
    a = ss->ema_a;
    bcoef = ss->ema_b;
  }
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) ema[i] = __fadd_rn(__fmul_rn(ema[i], a), __fmul_rn(src[i], bcoef));
  if (i < n_nbt) {
    float e = __fadd_rn(__fmul_rn((float)ema_nbt[i], a), __fmul_rn((float)nbt[i], bcoef));
    ema_nbt[i] = (int64_t)e;  // load_state_dict copies the float back into the int64 buffer (truncation)
  }
}

int ema_buffers(const float* bn_buffers, float* ema_bn_buffers, long long n, const int64_t* nbt, int64_t* ema_nbt,
                int n_nbt, float ema_alpha, int64_t ema_step, const bsed_step_state* ss, cudaStream_t st) {
  double a = 1.0 - 1.0 / ((double)ema_step + 1.0);
  if (a > (double)ema_alpha) a = (double)ema_alpha;
  long long work = n > n_nbt ? n : n_nbt;
  if (work <= 0) return BSED_OK;
  ema_buffers_kernel<<<ceil_div(work, 256), 256, 0, st>>>(bn_buffers, ema_bn_buffers, n, nbt, ema_nbt,
                                                          nbt && ema_nbt ? n_nbt : 0, (float)a, (float)(1.0 - a), ss);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace bsed
