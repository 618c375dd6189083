// gru.cu -- bidirectional GRU recurrence with register-resident recurrent weights.
//
// Reference: nn.GRU(128 -> 128, num_layers=2, bidirectional, batch_first)   src/models/RNN.py:7-16
//   r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)
//   z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh  (W_in x + b_in + r * (W_hn h + b_hn))
//   h' = (1 - z) * n + z * h                      gate order (r, z, n), h_0 = 0
// The input projections xg = W_i* x + b_i* come from one GEMM per layer (both directions,
// [B][T][768]); this kernel runs the 313 dependent steps.  One CTA per (clip, direction),
// 384 threads: thread j owns row j of W_hh (128 registers) and produces gate pre-activation j;
// the 128 hidden values are exchanged through shared memory.
#include "launch.h"

namespace bsed {

constexpr int kH = 128;
constexpr int kG = 384;

__device__ __forceinline__ int gru_group_of(const Groups& g, int clip) {
  int r = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroups; ++i)
    if (i < g.n && clip >= g.first[i]) r = i;
  return r;
}

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kG, 1) gru_fwd_kernel(const float* __restrict__ xg, Groups g, FloatPtrs whhT,
                                                        FloatPtrs bhh, float* __restrict__ out,
                                                        float* __restrict__ enc, float* __restrict__ saved, int T,
                                                        uint32_t key, uint32_t thresh, float inv_keep) {
  const int clip = g.first[0] + blockIdx.x;
  const int dir = blockIdx.y;
  const int j = threadIdx.x;
  const int grp = gru_group_of(g, clip);
  const float* WT = whhT.p[grp] + (size_t)dir * kH * kG;  // [k][j]
  float w[kH];
#pragma unroll
  for (int k = 0; k < kH; ++k) w[k] = WT[(size_t)k * kG + j];
  const float bj = bhh.p[grp][dir * kG + j];

  __shared__ __align__(16) float h_s[2][kH];   // double-buffered hidden state: two barriers per step
  __shared__ float gates_s[kG];
  if (j < kH) h_s[0][j] = 0.f;
  __syncthreads();

  // the input projections of step s + 1 are fetched while step s computes (their latency would otherwise sit on
  // the 313-step critical path)
  auto xrow = [&](int step) { return (size_t)clip * T + (dir == 0 ? step : T - 1 - step); };
  float nxr = 0.f, nxz = 0.f, nxn = 0.f;
  if (j < kH) {
    const float* xb = xg + xrow(0) * (2 * kG) + dir * kG;
    nxr = xb[j];
    nxz = xb[kH + j];
    nxn = xb[2 * kH + j];
  }
  int buf = 0;
  for (int step = 0; step < T; ++step) {
    const size_t row = xrow(step);
    const float xr = nxr, xz = nxz, xn = nxn;
    if (j < kH && step + 1 < T) {
      const float* xb = xg + xrow(step + 1) * (2 * kG) + dir * kG;
      nxr = xb[j];
      nxz = xb[kH + j];
      nxn = xb[2 * kH + j];
    }
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float4* h4 = reinterpret_cast<const float4*>(h_s[buf]);
#pragma unroll
    for (int k4 = 0; k4 < kH / 4; ++k4) {
      float4 hv = h4[k4];
      a0 = fmaf(w[4 * k4 + 0], hv.x, a0);
      a1 = fmaf(w[4 * k4 + 1], hv.y, a1);
      a2 = fmaf(w[4 * k4 + 2], hv.z, a2);
      a3 = fmaf(w[4 * k4 + 3], hv.w, a3);
    }
    gates_s[j] = (a0 + a1) + (a2 + a3) + bj;
    __syncthreads();
    if (j < kH) {
      float r = sigmoid_acc(xr + gates_s[j]);
      float z = sigmoid_acc(xz + gates_s[kH + j]);
      float hn = gates_s[2 * kH + j];
      float n = tanhf(fmaf(r, hn, xn));
      float hold = h_s[buf][j];
      float hnew = (1.f - z) * n + z * hold;
      h_s[buf ^ 1][j] = hnew;
      size_t o = row * (2 * kH) + dir * kH + j;
      out[o] = hnew;
      if (enc) {
        float e = hnew;
        if (thresh) e = bsed_keep((uint32_t)o, key, thresh) ? e * inv_keep : 0.f;
        enc[o] = e;
      }
      if (saved) {
        float* sv = saved + (row * 2 + dir) * (4 * kH);
        sv[j] = r;
        sv[kH + j] = z;
        sv[2 * kH + j] = n;
        sv[3 * kH + j] = hn;
      }
    }
    __syncthreads();
    buf ^= 1;
  }
}

int gru_forward(const float* xg, const Groups& g, const FloatPtrs& whhT, const FloatPtrs& bhh, float* out,
                float* enc, float* saved, int T, uint32_t key, uint32_t thresh, float inv_keep,
                cudaStream_t st) {
  int B = 0;
  for (int i = 0; i < g.n; ++i) B += g.count[i];
  dim3 grid(B, 2);
  ProfScope prof(PROF_GRU, 2.0 * B * 2 * T * kG * kH, 4.0 * B * T * (2.0 * kG + 2 * kH), st);
  gru_fwd_kernel<<<grid, kG, 0, st>>>(xg, g, whhT, bhh, out, enc, saved, T, key, thresh, inv_keep);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// backward through time.  thread (i = tid % 128, gs = tid / 128) owns W_hh[gs*128 + k][i], k = 0..127
// and produces the gs-th partial of dh_{t-1}[i] = sum_j W_hh[j][i] * dg_h[j].
//   dxg : gradient w.r.t. the input projections  (dr_pre, dz_pre, dn_pre)
//   dgh : gradient w.r.t. W_hh h + b_hh          (dr_pre, dz_pre, dn_pre * r)
__global__ void __launch_bounds__(kG, 1) gru_bwd_kernel(const float* __restrict__ dout,
                                                        const float* __restrict__ saved,
                                                        const float* __restrict__ out,
                                                        const float* __restrict__ whh, float* __restrict__ dxg,
                                                        float* __restrict__ dgh, int T, int first_clip) {
  const int clip = first_clip + blockIdx.x;
  const int dir = blockIdx.y;
  const int i = threadIdx.x % kH;
  const int gs = threadIdx.x / kH;
  const float* W = whh + (size_t)dir * kG * kH;  // [j][i]
  float w[kH];
#pragma unroll
  for (int k = 0; k < kH; ++k) w[k] = W[(size_t)(gs * kH + k) * kH + i];

  __shared__ __align__(16) float dg_s[kG];
  __shared__ float part_s[3][kH];
  float dh_carry = 0.f;

  // operands of the next step are fetched one step ahead (see gru_fwd_kernel)
  struct StepIn {
    float dout, r, z, n, hn, hprev;
  };
  auto fetch = [&](int step) {
    StepIn v;
    const int s = T - 1 - step;
    const int t = dir == 0 ? s : T - 1 - s;
    const int tp = dir == 0 ? t - 1 : t + 1;
    const size_t row = (size_t)clip * T + t;
    v.dout = dout[row * (2 * kH) + dir * kH + i];
    const float* sv = saved + (row * 2 + dir) * (4 * kH);
    v.r = sv[i];
    v.z = sv[kH + i];
    v.n = sv[2 * kH + i];
    v.hn = sv[3 * kH + i];
    v.hprev = s > 0 ? out[((size_t)clip * T + tp) * (2 * kH) + dir * kH + i] : 0.f;
    return v;
  };
  StepIn nxt = {};
  if (gs == 0) nxt = fetch(0);

  for (int step = 0; step < T; ++step) {
    const int s = T - 1 - step;                 // forward step being undone
    const int t = dir == 0 ? s : T - 1 - s;
    const size_t row = (size_t)clip * T + t;
    float dh_z = 0.f;
    if (gs == 0) {
      const StepIn c = nxt;
      if (step + 1 < T) nxt = fetch(step + 1);
      float dh = dh_carry + c.dout;
      float r = c.r, z = c.z, n = c.n, hn = c.hn;
      float dn = dh * (1.f - z);
      float dz = dh * (c.hprev - n);
      float dnp = dn * (1.f - n * n);
      float drp = dnp * hn * r * (1.f - r);
      float dzp = dz * z * (1.f - z);
      float dhn = dnp * r;
      float* gx = dxg + row * (2 * kG) + dir * kG;
      gx[i] = drp;
      gx[kH + i] = dzp;
      gx[2 * kH + i] = dnp;
      float* gh = dgh + row * (2 * kG) + dir * kG;
      gh[i] = drp;
      gh[kH + i] = dzp;
      gh[2 * kH + i] = dhn;
      dg_s[i] = drp;
      dg_s[kH + i] = dzp;
      dg_s[2 * kH + i] = dhn;
      dh_z = dh * z;
    }
    __syncthreads();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float4* d4 = reinterpret_cast<const float4*>(dg_s + gs * kH);
#pragma unroll
    for (int k4 = 0; k4 < kH / 4; ++k4) {
      float4 dv = d4[k4];
      a0 = fmaf(w[4 * k4 + 0], dv.x, a0);
      a1 = fmaf(w[4 * k4 + 1], dv.y, a1);
      a2 = fmaf(w[4 * k4 + 2], dv.z, a2);
      a3 = fmaf(w[4 * k4 + 3], dv.w, a3);
    }
    part_s[gs][i] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (gs == 0) dh_carry = dh_z + part_s[0][i] + part_s[1][i] + part_s[2][i];
  }
}

int gru_backward(const float* dout, const float* saved, const float* out, const float* whh, float* dxg,
                 float* dgh, int T, int first_clip, int n_clips, cudaStream_t st) {
  dim3 grid(n_clips, 2);
  ProfScope prof(PROF_GRU, 2.0 * n_clips * 2 * T * kG * kH, 4.0 * n_clips * T * (4.0 * kG + 8 * kH), st);
  gru_bwd_kernel<<<grid, kG, 0, st>>>(dout, saved, out, whh, dxg, dgh, T, first_clip);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace bsed
