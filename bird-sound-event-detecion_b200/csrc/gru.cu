// gru.cu -- bidirectional GRU recurrence with register-resident recurrent weights.
//
// Reference: nn.GRU(128 -> 128, num_layers=2, bidirectional, batch_first)   src/models/RNN.py:7-16
//   r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)
//   z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh  (W_in x + b_in + r * (W_hn h + b_hn))
//   h' = (1 - z) * n + z * h                      gate order (r, z, n), h_0 = 0
// The input projections xg = W_i* x + b_i* come from one GEMM per layer (both directions,
// [B][T][768]); this kernel runs the 313 dependent steps.  One CTA per (clip, direction), 384 threads, W_hh resident in
// registers (128 per thread), the 128 hidden values exchanged through shared memory.  The recurrent dot products are
// split four ways along k inside each quad of lanes: a thread multiplies a quarter of h (8 broadcast LDS.128 instead of
// 32 -- the shared-memory loads were what the step stalled on) into 4 gate rows with packed FMAs, and a 3-shuffle
// transpose-reduce leaves lane q of the quad with the complete pre-activation of row 4 * quad + q = its thread index.
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

// The two nonlinearities sit on the 313-step critical path.  ex2.approx (2 ulp) + an approximate division keep the
// relative error near 3e-7 -- far inside the 1e-3 parity bound on the GRU outputs -- at a third of the instructions of
// expf / tanhf / an IEEE division.
__device__ __forceinline__ float sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_acc(float x) {
  const float t = __expf(-2.0f * fabsf(x));
  return copysignf(__fdividef(1.0f - t, 1.0f + t), x);
}

// shared-memory position of element k of a 128-vector read by quad_dot: quarter k / 32 starts at float 36 * (k / 32)
constexpr int kQuadVec = 4 * 36;
__device__ __forceinline__ int quad_pos(int k) { return (k >> 5) * 36 + (k & 31); }

// Four dot products of length 128 shared by a quad of lanes.  Lane q holds w[r][.] = rows r = 0..3 restricted to
// k in [32 q, 32 q + 32) and reads that quarter of the vector (v4: its eight float4); returns the complete dot product of
// row q.  Partial sums: two packed accumulators per row (even / odd k), then a transpose-reduce over the quad.
__device__ __forceinline__ float quad_dot(const float2 (&w)[4][kH / 8], const float4* __restrict__ v4, int q) {
  float2 a[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int k4 = 0; k4 < 8; ++k4) {
    const float4 hv = v4[k4];
    const float2 lo = make_float2(hv.x, hv.y), hi = make_float2(hv.z, hv.w);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      ffma2(a[r], w[r][2 * k4], lo);
      ffma2(a[r], w[r][2 * k4 + 1], hi);
    }
  }
  float p0 = a[0].x + a[0].y, p1 = a[1].x + a[1].y, p2 = a[2].x + a[2].y, p3 = a[3].x + a[3].y;
  // lanes with q & 1 keep rows 1, 3 and hand over rows 0, 2 (and vice versa); then q & 2 splits {0, 1} from {2, 3}
  const bool odd = q & 1;
  const float s0 = odd ? p0 : p1, s1 = odd ? p2 : p3;
  const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  const float u0 = (odd ? p1 : p0) + r0;      // row (q & 1)
  const float u1 = (odd ? p3 : p2) + r1;      // row 2 + (q & 1)
  const bool up = q & 2;
  const float r2 = __shfl_xor_sync(0xffffffffu, up ? u0 : u1, 2);
  return (up ? u1 : u0) + r2;                 // row q
}

__global__ void __launch_bounds__(kG, 1) gru_fwd_kernel(const float* __restrict__ xg, Groups g, FloatPtrs whhT,
                                                        FloatPtrs bhh, float* __restrict__ out,
                                                        float* __restrict__ enc, float* __restrict__ saved, int T,
                                                        DropKey dkey, uint32_t thresh, float inv_keep) {
  const uint32_t key = dkey.get();
  const int clip = g.first[0] + blockIdx.x;
  const int dir = blockIdx.y;
  const int j = threadIdx.x;
  const int grp = gru_group_of(g, clip);
  const float* WT = whhT.p[grp] + (size_t)dir * kH * kG;  // [k][j]
  const int q = j & 3, row0 = j & ~3;   // k quarter of this lane, first of the quad's four gate rows
  float2 w[4][kH / 8];                  // rows row0 .. row0 + 3, k in [32 q, 32 q + 32), as (even, odd) pairs
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int k = 0; k < kH / 8; ++k)
      w[r][k] = make_float2(WT[(size_t)(32 * q + 2 * k) * kG + row0 + r], WT[(size_t)(32 * q + 2 * k + 1) * kG + row0 + r]);
  const float bj = bhh.p[grp][dir * kG + j];

  // double-buffered hidden state (two barriers per step), stored as four quarters of 32 values 36 floats apart: the four
  // lanes of a quad read four different quarters in one LDS.128, and a 32-float stride would put them on the same banks
  __shared__ __align__(16) float h_s[2][kQuadVec];
  __shared__ float gates_s[kG];                // r, z (after the sigmoid) and W_hn h + b_hn
  if (j < kH) h_s[0][quad_pos(j)] = 0.f;
  __syncthreads();

  // Thread j owns gate row j end to end: its input projection x_j (fetched one step ahead, so the load latency is
  // off the 313-step critical path), its recurrent dot product, and -- for the r and z rows -- the sigmoid, so that
  // after the barrier only tanh and the blend remain (done by the 128 threads of the n rows).
  auto xrow = [&](int step) { return (size_t)clip * T + (dir == 0 ? step : T - 1 - step); };
  float nx = xg[xrow(0) * (2 * kG) + dir * kG + j];
  const int u = j - 2 * kH;                    // hidden unit of an n-row thread
  int buf = 0;
  for (int step = 0; step < T; ++step) {
    const size_t row = xrow(step);
    const float x = nx;
    if (step + 1 < T) nx = xg[xrow(step + 1) * (2 * kG) + dir * kG + j];
    const float acc = quad_dot(w, reinterpret_cast<const float4*>(h_s[buf]) + 9 * q, q) + bj;
    float* sv = saved ? saved + (row * 2 + dir) * (4 * kH) : nullptr;
    if (j < 2 * kH) {
      const float sg = sigmoid_acc(x + acc);
      gates_s[j] = sg;
      if (sv) sv[j] = sg;                      // r at [0,128), z at [128,256)
    }
    __syncthreads();
    if (j >= 2 * kH) {
      const float r = gates_s[u], z = gates_s[kH + u];
      const float n = tanh_acc(fmaf(r, acc, x));
      const float hold = h_s[buf][quad_pos(u)];
      const float hnew = (1.f - z) * n + z * hold;
      h_s[buf ^ 1][quad_pos(u)] = hnew;
      size_t o = row * (2 * kH) + dir * kH + u;
      out[o] = hnew;
      if (enc) {
        float e = hnew;
        if (thresh) e = bsed_keep((uint32_t)o, key, thresh) ? e * inv_keep : 0.f;
        enc[o] = e;
      }
      if (sv) {
        sv[2 * kH + u] = n;
        sv[3 * kH + u] = acc;
      }
    }
    __syncthreads();
    buf ^= 1;
  }
}

int gru_forward(const float* xg, const Groups& g, const FloatPtrs& whhT, const FloatPtrs& bhh, float* out,
                float* enc, float* saved, int T, DropKey key, uint32_t thresh, float inv_keep,
                cudaStream_t st) {
  int B = 0;
  for (int i = 0; i < g.n; ++i) B += g.count[i];
  dim3 grid(B, 2);
  ProfScope prof(PROF_GRU, 2.0 * B * 2 * T * kG * kH, 4.0 * B * T * (2.0 * kG + 2 * kH), st);
  gru_fwd_kernel<<<grid, kG, 0, st>>>(xg, g, whhT, bhh, out, enc, saved, T, key, thresh, inv_keep);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// backward through time.  thread (i = tid % 128, gs = tid / 128) produces the gs-th partial of
// dh_{t-1}[i] = sum_j W_hh[j][i] * dg_h[j] (the quad it belongs to shares the four outputs i & ~3 .. and splits k).
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
  // same quad scheme as the forward kernel: the quad of this thread owns outputs i0 .. i0 + 3 of gate group gs, lane q
  // the quarter k in [32 q, 32 q + 32) of the reduction over the group's 128 gate rows
  const int q = threadIdx.x & 3, i0 = i & ~3;
  float2 w[4][kH / 8];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int k = 0; k < kH / 8; ++k)
      w[r][k] = make_float2(W[(size_t)(gs * kH + 32 * q + 2 * k) * kH + i0 + r], W[(size_t)(gs * kH + 32 * q + 2 * k + 1) * kH + i0 + r]);

  __shared__ __align__(16) float dg_s[3][kQuadVec];   // per gate group, in quad_dot's padded layout
  __shared__ float part_s[3][kH];
  // operands of the coming steps (r, z, n, W_hn h + b_hn, d_out, h_prev: 6 x 128 floats = 192 16-byte chunks per
  // step) are staged through shared memory kRing - 1 steps ahead with cp.async: the DRAM latency of these cold
  // tensors (~1 us, longer than one step) stays off the critical path without spending registers
  constexpr int kRing = 8;
  __shared__ __align__(16) float ring[kRing][6 * kH];
  float dh_carry = 0.f;
  const int tid = threadIdx.x;

  auto issue = [&](int step) {
    if (step < T && tid < 192) {
      const int s = T - 1 - step;
      const int t = dir == 0 ? s : T - 1 - s;
      const int tp = dir == 0 ? t - 1 : t + 1;
      const size_t row = (size_t)clip * T + t;
      const float* src;
      bool valid = true;
      if (tid < 128) {
        src = saved + (row * 2 + dir) * (4 * kH) + tid * 4;
      } else if (tid < 160) {
        src = dout + row * (2 * kH) + dir * kH + (tid - 128) * 4;
      } else {
        valid = s > 0;
        src = valid ? out + ((size_t)clip * T + tp) * (2 * kH) + dir * kH + (tid - 160) * 4 : out;
      }
      cp_async16(&ring[step % kRing][tid * 4], src, valid);
    }
    cp_async_commit();
  };
  for (int p = 0; p < kRing - 1; ++p) issue(p);
  cp_async_wait<kRing - 2>();
  __syncthreads();

  for (int step = 0; step < T; ++step) {
    const int s = T - 1 - step;                 // forward step being undone
    const int t = dir == 0 ? s : T - 1 - s;
    const size_t row = (size_t)clip * T + t;
    float dh_z = 0.f;
    if (gs == 0) {
      const float* c = ring[step % kRing];
      float dh = dh_carry + c[4 * kH + i];
      float r = c[i], z = c[kH + i], n = c[2 * kH + i], hn = c[3 * kH + i];
      float dn = dh * (1.f - z);
      float dz = dh * (c[5 * kH + i] - n);
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
      dg_s[0][quad_pos(i)] = drp;
      dg_s[1][quad_pos(i)] = dzp;
      dg_s[2][quad_pos(i)] = dhn;
      dh_z = dh * z;
    }
    __syncthreads();
    issue(step + kRing - 1);                    // refills the slot consumed in the previous step
    part_s[gs][i] = quad_dot(w, reinterpret_cast<const float4*>(dg_s[gs]) + 9 * q, q);
    cp_async_wait<kRing - 2>();                 // the group of step + 1 has landed (for this thread) ...
    __syncthreads();                            // ... and for every thread
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
