// frontend.cu -- framed STFT -> |X| -> Slaney mel projection (K1 + K2), and the dB / noise /
// pad / normalise transform.
//
// Reference: src/data/preprocess.py:18-45 (librosa.stft n_fft=2048 hop=255 symmetric Hamming,
// center=True reflect; melspectrogram(S=|X|, 128 Slaney bands, norm=None)), and
// src/data/Transforms.py:74-139,155-197,304-322 (AugmentGaussianNoise, ApplyLog, PadOrTrunc,
// Normalize).
//
// K1/K2 design: persistent CTAs (one per SM, 16 warps = 4 per scheduler; below that the kernel is latency-bound).  A
// tile is 16 consecutive frames of one clip; its audio (2048 + 15*255 samples, reflect padding resolved per element)
// is staged in shared memory with cp.async while the previous tile is being transformed (double buffer), so the 8x
// frame overlap is served from SMEM, not HBM, and the staging latency is hidden.  Each warp owns one frame of the
// tile: the 2048-point real FFT is a 1024-point complex FFT (even/odd packing) done as 32 x 32 -- two in-register
// 32-point FFTs per lane with one padded shared-memory transpose in between -- followed by the real-FFT untangling,
// the magnitude (written back compactly over the spectrum), and the sparse mel projection (2016 non-zero weights,
// <= 60 bins per band), all without leaving the SM.  Only the 128 mel values per frame go back to HBM.
#include "launch.h"

namespace bsed {

constexpr int FE_FPC = 16;                                  // frames per tile (one per warp)
constexpr int FE_WARPS = 16;
constexpr int FE_AUD = ((kNFFT + (FE_FPC - 1) * kHop) + 15) / 16 * 16;  // 5888 staged samples
constexpr int FE_BUF = 33 * 32;                             // padded transpose buffer (float2)
constexpr int FE_IVW = 1028;                                // float2 entries of the interval filterbank
constexpr int FE_NIV = kNMels + 1;                          // intervals between the 130 band edges

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// e^{-2 pi i idx / 32}, idx = 0..15 (compile-time idx after unrolling)
__device__ __forceinline__ float2 tw32(int idx) {
  switch (idx) {
    case 0: return make_float2(1.0f, 0.0f);
    case 1: return make_float2(0.98078528040323043f, -0.19509032201612825f);
    case 2: return make_float2(0.92387953251128674f, -0.38268343236508978f);
    case 3: return make_float2(0.83146961230254524f, -0.55557023301960218f);
    case 4: return make_float2(0.70710678118654757f, -0.70710678118654757f);
    case 5: return make_float2(0.55557023301960229f, -0.83146961230254524f);
    case 6: return make_float2(0.38268343236508984f, -0.92387953251128674f);
    case 7: return make_float2(0.19509032201612833f, -0.98078528040323043f);
    case 8: return make_float2(0.0f, -1.0f);
    case 9: return make_float2(-0.19509032201612819f, -0.98078528040323043f);
    case 10: return make_float2(-0.38268343236508973f, -0.92387953251128674f);
    case 11: return make_float2(-0.55557023301960196f, -0.83146961230254546f);
    case 12: return make_float2(-0.70710678118654746f, -0.70710678118654757f);
    case 13: return make_float2(-0.83146961230254535f, -0.55557023301960218f);
    case 14: return make_float2(-0.92387953251128674f, -0.38268343236508989f);
    default: return make_float2(-0.98078528040323043f, -0.19509032201612861f);
  }
}

__device__ __forceinline__ constexpr int bitrev5(int i) {
  return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// in-register radix-2 DIF FFT of 32 complex points; result X[bitrev5(i)] = v[i]
__device__ __forceinline__ void fft32(float2 (&v)[32]) {
#pragma unroll
  for (int half = 16; half >= 1; half >>= 1) {
#pragma unroll
    for (int base = 0; base < 32; base += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; ++j) {
        float2 a = v[base + j], b = v[base + j + half];
        v[base + j] = make_float2(a.x + b.x, a.y + b.y);
        float2 d = make_float2(a.x - b.x, a.y - b.y);
        const int ti = j * (16 / half);
        if (ti == 0) v[base + j + half] = d;
        else if (ti == 8) v[base + j + half] = make_float2(d.y, -d.x);
        else v[base + j + half] = cmul(d, tw32(ti));
      }
    }
  }
}

struct FrontendTables {
  const float* window;
  const float2* tw1024;
  const float2* tw2048;
  const float2* iv_w;     // interval form of the mel filterbank (api.cu: build_tables)
  const int* iv_start;
  const int* iv_len;
  int iv_bin0, iv_n;
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(dst), "l"(gmem_src));
}

__global__ void __launch_bounds__(FE_WARPS * 32, 1) melspec_kernel(const float* __restrict__ audio, int n_samples,
                                                                   int n_frames, int tiles_per_clip, int n_tiles,
                                                                   float* __restrict__ mel, FrontendTables tb,
                                                                   double* __restrict__ clip_max_ws) {
  extern __shared__ __align__(16) unsigned char fe_smem[];
  float* audio_s = reinterpret_cast<float*>(fe_smem);          // 2 x FE_AUD (double buffer)
  float* win_s = audio_s + 2 * FE_AUD;
  float2* tw1024_s = reinterpret_cast<float2*>(win_s + kNFFT);  // [c][lane] = e^{-2 pi i lane c / 1024}
  float2* tw2048_s = tw1024_s + 1024;
  float2* ivw_s = tw2048_s + 516;
  int* ivs_s = reinterpret_cast<int*>(ivw_s + FE_IVW);
  int* ivl_s = ivs_s + FE_NIV + 3;
  float2* bufs = reinterpret_cast<float2*>(ivl_s + FE_NIV + 3);

  const int tid = threadIdx.x;
  const int lane = tid % 32, warp = tid / 32;
  const long long padded = (long long)n_samples + kNFFT;

  // stage the audio of one tile (reflect padding: ypad[p] = y[reflect(p - 1024)])
  auto stage = [&](int tile, float* dst) {
    const int b = tile / tiles_per_clip;
    const long long p0 = (long long)(tile % tiles_per_clip) * FE_FPC * kHop;
    const float* y = audio + (size_t)b * n_samples;
    for (int i = tid; i < FE_AUD; i += FE_WARPS * 32) {
      const long long p = p0 + i;
      if (p < padded) {
        long long j = p - kNFFT / 2;
        if (j < 0) j = -j;
        if (j >= n_samples) j = 2LL * (n_samples - 1) - j;
        cp_async4(dst + i, y + j);
      } else {
        dst[i] = 0.f;
      }
    }
  };

  int tile = blockIdx.x;
  if (tile < n_tiles) stage(tile, audio_s);
  cp_async_commit();
  for (int i = tid; i < kNFFT; i += blockDim.x) win_s[i] = tb.window[i];
  for (int i = tid; i < 1024; i += blockDim.x) tw1024_s[i] = tb.tw1024[i];
  for (int i = tid; i < 513; i += blockDim.x) tw2048_s[i] = tb.tw2048[i];
  for (int i = tid; i < tb.iv_n; i += blockDim.x) ivw_s[i] = tb.iv_w[i];
  for (int i = tid; i < FE_NIV; i += blockDim.x) {
    ivs_s[i] = tb.iv_start[i];
    ivl_s[i] = tb.iv_len[i];
  }

  float2* buf = bufs + warp * FE_BUF;
  // magnitudes overwrite the spectrum once every lane holds its share in registers
  float* mag = reinterpret_cast<float*>(buf);

  int cur = 0;
  for (; tile < n_tiles; tile += gridDim.x, cur ^= 1) {
    cp_async_wait<0>();
    __syncthreads();   // this tile's audio has landed; every warp is done with the other buffer
    const int next = tile + gridDim.x;
    if (next < n_tiles) stage(next, audio_s + (cur ^ 1) * FE_AUD);
    cp_async_commit();

    const int b = tile / tiles_per_clip;
    const int t = (tile % tiles_per_clip) * FE_FPC + warp;
    if (t >= n_frames) continue;
    const float* fr = audio_s + cur * FE_AUD + warp * kHop;
    float2 v[32];
    // pass 1: lane = b0; z[32 a + b0] = (x[64a + 2b0] w, x[64a + 2b0 + 1] w)
#pragma unroll
    for (int a = 0; a < 32; ++a) {
      int n0 = 64 * a + 2 * lane;
      const float2 w2 = *reinterpret_cast<const float2*>(win_s + n0);
      v[a] = make_float2(fr[n0] * w2.x, fr[n0 + 1] * w2.y);
    }
    fft32(v);
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      float2 yv = v[bitrev5(c)];
      if (c > 0) yv = cmul(yv, tw1024_s[c * 32 + lane]);
      buf[c * 33 + lane] = yv;
    }
    __syncwarp();
    // pass 2: lane = c; FFT over b0
#pragma unroll
    for (int bb = 0; bb < 32; ++bb) v[bb] = buf[lane * 33 + bb];
    fft32(v);
    __syncwarp();
#pragma unroll
    for (int d = 0; d < 32; ++d) buf[lane + 32 * d] = v[bitrev5(d)];
    __syncwarp();
    // real-FFT untangling + magnitude: lane holds |X[k]|, |X[1024-k]| for k = lane + 32 jj
    float mlo[16], mhi[16], m512 = 0.f;
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      int k = lane + 32 * jj;
      if (k == 0) {
        float2 z0 = buf[0];
        float2 zh = buf[512];
        mlo[jj] = fabsf(z0.x + z0.y);
        mhi[jj] = fabsf(z0.x - z0.y);
        m512 = sqrtf(zh.x * zh.x + zh.y * zh.y);
      } else {
        float2 zk = buf[k], zm = buf[1024 - k];
        float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
        float2 O = make_float2(0.5f * (zk.x - zm.x), 0.5f * (zk.y + zm.y));
        float2 P = cmul(tw2048_s[k], O);
        float xr = E.x + P.y, xi = E.y - P.x;
        float yr = E.x - P.y, yi = -E.y - P.x;
        mlo[jj] = sqrtf(xr * xr + xi * xi);
        mhi[jj] = sqrtf(yr * yr + yi * yi);
      }
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      int k = lane + 32 * jj;
      mag[k] = mlo[jj];
      mag[1024 - k] = mhi[jj];
    }
    if (lane == 0) mag[512] = m512;
    __syncwarp();
    // sparse mel projection by intervals between band edges: lane -> intervals lane, lane+32, lane+64, lane+96; one
    // magnitude load feeds the rising slope of band j (U) and the falling slope of band j-1 (D); mel[m] = U_m + D_(m+1)
    float* out = mel + ((size_t)b * n_frames + t) * kNMels;
    const float* magb = mag + tb.iv_bin0;
    float U[4], D[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int j = lane + 32 * r;
      const int s = ivs_s[j], len = ivl_s[j];
      float u = 0.f, d = 0.f;
      for (int i = 0; i < len; ++i) {
        const float2 w2 = ivw_s[s + i];
        const float a = magb[s + i];
        u = fmaf(w2.x, a, u);
        d = fmaf(w2.y, a, d);
      }
      U[r] = u;
      D[r] = d;
    }
    // top interval (above the centre of band 127, <= 32 bins): one bin per lane
    float dtop = 0.f;
    {
      const int s = ivs_s[kNMels], len = ivl_s[kNMels];
      if (lane < len) dtop = ivw_s[s + lane].y * magb[s + lane];
      dtop = warp_sum(dtop);
    }
    float fmx = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float dn = __shfl_down_sync(0xffffffffu, D[r], 1);                       // D of interval j + 1 (same pass)
      const float dnext = r < 3 ? __shfl_sync(0xffffffffu, D[r < 3 ? r + 1 : r], 0) : dtop;   // lane 31: first interval of the next pass
      if (lane == 31) dn = dnext;
      const float mv = U[r] + dn;
      out[lane + 32 * r] = mv;
      fmx = fmaxf(fmx, fabsf(mv));
    }
    if (clip_max_ws) {   // fused log-mel: the clip maximum the dB clamp needs (slot kNMels of the clip's amp_to_db workspace)
      fmx = warp_max(fmx);
      if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(clip_max_ws + (size_t)b * (kNMels + 1) + kNMels), __float_as_uint(fmx));
    }
    __syncwarp();
  }
  cp_async_wait<0>();
}

constexpr size_t FE_SMEM_BYTES = sizeof(float) * (2 * FE_AUD + kNFFT) + sizeof(float2) * (1024 + 516 + FE_IVW) +
                                 sizeof(int) * 2 * (FE_NIV + 3) + sizeof(float2) * FE_WARPS * FE_BUF;

static int melspec_impl(bsed_context* h, const float* audio, int B, int n_samples, float* mel, double* clip_max_ws, cudaStream_t st);

int melspec(bsed_context* h, const float* audio, int B, int n_samples, float* mel, cudaStream_t st) {
  return melspec_impl(h, audio, B, n_samples, mel, nullptr, st);
}

static int melspec_impl(bsed_context* h, const float* audio, int B, int n_samples, float* mel, double* clip_max_ws, cudaStream_t st) {
  BSED_REQUIRE(n_samples >= kNFFT / 2 + 1, "melspec: n_samples=%d < 1025 (reflect padding)", n_samples);
  BSED_REQUIRE(B > 0, "melspec: B=%d", B);
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    BSED_CHECK_CUDA(cudaFuncSetAttribute(melspec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FE_SMEM_BYTES));
  }
  int n_frames = 1 + n_samples / kHop;
  FrontendTables tb{h->window, h->tw1024, h->tw2048, h->mel_iv_w, h->mel_iv_start, h->mel_iv_len, h->mel_iv_bin0, h->mel_iv_n};
  const int tiles_per_clip = ceil_div(n_frames, FE_FPC);
  const long long n_tiles = (long long)tiles_per_clip * B;
  BSED_REQUIRE(n_tiles < (1ll << 31), "melspec: too many frames");
  const int grid = (int)(n_tiles < h->num_sms ? n_tiles : h->num_sms);
  // algorithmic bytes: audio in + mel out; flops: rFFT-2048 + magnitude + sparse mel
  ProfScope prof(PROF_MELSPEC, (double)B * n_frames * 70000.0, 4.0 * B * ((double)n_samples + (double)n_frames * kNMels), st);
  melspec_kernel<<<grid, FE_WARPS * 32, FE_SMEM_BYTES, st>>>(audio, n_samples, n_frames, tiles_per_clip, (int)n_tiles, mel, tb,
                                                             clip_max_ws);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// amplitude_to_db (+ noise, pad/trunc, normalise): HBM-bound, float4 accesses, grid = (row chunks, clips).
//   1. (noise only) per clip and mel bin: sum_t mel^2  -> std_m = sqrt(mean_t(mel^2) * 10^(-snr/10))   [fp64 atomics]
//   2. per clip: max |x|, x = mel (+ std_m * noise)                                  [atomicMax on the float bits]
//   3. elementwise: max(20 log10(max(1e-5,|x|)), maxdb - 80), rows >= t_in -> 0, (v - mean_m)/std_m
// The noisy sample x is formed in double (the reference evaluates that branch in float64; a float add would lose the
// small |x| left after cancellation); the logarithm is fp32, as numpy's is on the reference's float32 path.
// workspace per clip: 128 doubles (sum of squares) + 1 slot (max bits)
// ---------------------------------------------------------------------------------------------
constexpr int DB_CHUNKS = 16;   // row chunks per clip in the reduction passes

__global__ void __launch_bounds__(256) db_sumsq_kernel(const float* __restrict__ mel, int t_in, double* __restrict__ ws) {
  const int b = blockIdx.y;
  const int q = threadIdx.x % 32, r = threadIdx.x / 32;   // 4 mel bins per thread, 8 row lanes
  const int rows = (t_in + DB_CHUNKS - 1) / DB_CHUNKS;
  const int t0 = blockIdx.x * rows;
  const int t1 = min(t_in, t0 + rows);
  const float* x = mel + (size_t)b * t_in * kNMels;
  double s[4] = {0, 0, 0, 0};
  for (int t = t0 + r; t < t1; t += 8) {
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)t * kNMels + q * 4);
    s[0] += (double)v.x * (double)v.x;
    s[1] += (double)v.y * (double)v.y;
    s[2] += (double)v.z * (double)v.z;
    s[3] += (double)v.w * (double)v.w;
  }
  __shared__ double red[8][kNMels];
#pragma unroll
  for (int j = 0; j < 4; ++j) red[r][q * 4 + j] = s[j];
  __syncthreads();
  if (threadIdx.x < kNMels) {
    double tot = 0;
    for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
    atomicAdd(ws + (size_t)b * (kNMels + 1) + threadIdx.x, tot);
  }
}

__device__ __forceinline__ float db_sample(float m, float nz, double std_m) { return (float)fabs((double)m + std_m * (double)nz); }

template <bool NOISE>
__global__ void __launch_bounds__(256) db_max_kernel(const float* __restrict__ mel, const float* __restrict__ noise,
                                                     double snr_scale, int t_in, double* __restrict__ ws) {
  const int b = blockIdx.y;
  const int q = threadIdx.x % 32, r = threadIdx.x / 32;
  const int rows = (t_in + DB_CHUNKS - 1) / DB_CHUNKS;
  const int t0 = blockIdx.x * rows;
  const int t1 = min(t_in, t0 + rows);
  const float* x = mel + (size_t)b * t_in * kNMels;
  const float* nz = NOISE ? noise + (size_t)b * t_in * kNMels : nullptr;
  double* w = ws + (size_t)b * (kNMels + 1);
  double sd[4] = {0, 0, 0, 0};
  if (NOISE) {
#pragma unroll
    for (int j = 0; j < 4; ++j) sd[j] = sqrt(w[q * 4 + j] * snr_scale / (double)t_in);
  }
  float mx = 0.f;
  for (int t = t0 + r; t < t1; t += 8) {
    const float4 v = *reinterpret_cast<const float4*>(x + (size_t)t * kNMels + q * 4);
    if (NOISE) {
      const float4 n = *reinterpret_cast<const float4*>(nz + (size_t)t * kNMels + q * 4);
      mx = fmaxf(mx, fmaxf(fmaxf(db_sample(v.x, n.x, sd[0]), db_sample(v.y, n.y, sd[1])),
                           fmaxf(db_sample(v.z, n.z, sd[2]), db_sample(v.w, n.w, sd[3]))));
    } else {
      mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
  }
  mx = warp_max(mx);
  __shared__ float red[8];
  if (q == 0) red[r] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    atomicMax(reinterpret_cast<unsigned int*>(w + kNMels), __float_as_uint(mx));   // mx >= 0: bit order == value order
  }
}

template <bool NOISE>
__global__ void __launch_bounds__(256) db_apply_kernel(const float* __restrict__ mel, const float* __restrict__ noise,
                                                       double snr_scale, int t_in, int frames,
                                                       const float* __restrict__ sc_mean, const float* __restrict__ sc_std,
                                                       const double* __restrict__ ws, float* __restrict__ out) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one float4
  if (i >= (long long)frames * (kNMels / 4)) return;
  const int q = (int)(i % (kNMels / 4)), t = (int)(i / (kNMels / 4));
  const double* w = ws + (size_t)b * (kNMels + 1);
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  if (t < t_in) {
    const size_t e = ((size_t)b * t_in + t) * kNMels + q * 4;
    const float4 v = *reinterpret_cast<const float4*>(mel + e);
    float xs[4] = {fabsf(v.x), fabsf(v.y), fabsf(v.z), fabsf(v.w)};
    if (NOISE) {
      const float4 n = *reinterpret_cast<const float4*>(noise + e);
      const float ms[4] = {v.x, v.y, v.z, v.w}, ns[4] = {n.x, n.y, n.z, n.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) xs[j] = db_sample(ms[j], ns[j], sqrt(w[q * 4 + j] * snr_scale / (double)t_in));
    }
    const float mxv = __uint_as_float(*reinterpret_cast<const unsigned int*>(w + kNMels));
    // 20 log10(x) = 20 log10(2) * log2(x); __log2f (MUFU.LG2, <= 2 ulp: < 3e-5 dB over [1e-5, 1e4]) keeps the pass
    // HBM-bound -- log10f made it issue-bound (ncu: 78 % issue slots, 52 % DRAM)
    constexpr float kDbPerLog2 = 6.02059991327962390f;
    const float floor_db = kDbPerLog2 * __log2f(fmaxf(1e-5f, mxv)) - 80.0f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = fmaxf(kDbPerLog2 * __log2f(fmaxf(1e-5f, xs[j])), floor_db);
  }
  // ApplyLog / PadOrTrunc produce float32 (ToTensor().float()) before Normalize
  if (sc_mean) {
    const float4 mu = *reinterpret_cast<const float4*>(sc_mean + q * 4), sg = *reinterpret_cast<const float4*>(sc_std + q * 4);
    o[0] = (o[0] - mu.x) / sg.x;
    o[1] = (o[1] - mu.y) / sg.y;
    o[2] = (o[2] - mu.z) / sg.z;
    o[3] = (o[3] - mu.w) / sg.w;
  }
  *reinterpret_cast<float4*>(out + ((size_t)b * frames + t) * kNMels + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

int amp_to_db(const float* mel, const float* noise, float snr_db, int B, int t_in, int frames,
              const float* sc_mean, const float* sc_std, float* out, void* ws, size_t ws_bytes,
              cudaStream_t st) {
  BSED_REQUIRE(B > 0 && t_in > 0 && frames > 0, "amp_to_db: B=%d t_in=%d frames=%d", B, t_in, frames);
  BSED_REQUIRE((sc_mean == nullptr) == (sc_std == nullptr), "amp_to_db: scaler mean/std must come together");
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  BSED_REQUIRE(aligned16(mel) && aligned16(noise) && aligned16(out) && aligned16(sc_mean) && aligned16(sc_std) && aligned16(ws),
               "amp_to_db: buffers must be 16-byte aligned (float4 accesses)");
  const size_t need = sizeof(double) * (size_t)B * (kNMels + 1);
  if (ws_bytes < need) {
    bsed_set_error("amp_to_db: workspace %zu < %zu", ws_bytes, need);
    return BSED_E_WORKSPACE;
  }
  const double snr_scale = pow(10.0, -(double)snr_db / 10.0);
  double* w = (double*)ws;
  // algorithmic bytes: mel (and noise) read once + log-mel written once; the reductions re-read mel (and noise)
  ProfScope prof(PROF_ELEMENTWISE, 0.0, 4.0 * B * kNMels * ((noise ? 2.0 : 1.0) * t_in + (double)frames), st);
  BSED_CHECK_CUDA(cudaMemsetAsync(w, 0, need, st));
  dim3 rgrid(DB_CHUNKS, B);
  if (noise) {
    db_sumsq_kernel<<<rgrid, 256, 0, st>>>(mel, t_in, w);
    BSED_CHECK_LAUNCH();
    db_max_kernel<true><<<rgrid, 256, 0, st>>>(mel, noise, snr_scale, t_in, w);
  } else {
    db_max_kernel<false><<<rgrid, 256, 0, st>>>(mel, nullptr, snr_scale, t_in, w);
  }
  BSED_CHECK_LAUNCH();
  dim3 grid(ceil_div((long long)frames * (kNMels / 4), 256), B);
  if (noise)
    db_apply_kernel<true><<<grid, 256, 0, st>>>(mel, noise, snr_scale, t_in, frames, sc_mean, sc_std, w, out);
  else
    db_apply_kernel<false><<<grid, 256, 0, st>>>(mel, nullptr, snr_scale, t_in, frames, sc_mean, sc_std, w, out);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// preprocess(audio, compute_log=True) / the load-time ApplyLog -> PadOrTrunc -> [Normalize] of a clean clip in one call:
// the STFT + mel kernel also takes the per-clip maximum (atomicMax on the float bits), so the dB pass reads the
// amplitude-mel once instead of twice.  Results are bit-identical to melspec followed by amp_to_db without noise.
int logmel(bsed_context* h, const float* audio, int B, int n_samples, int frames, const float* sc_mean, const float* sc_std,
           float* mel, float* out, void* ws, size_t ws_bytes, cudaStream_t st) {
  BSED_REQUIRE(frames > 0, "logmel: frames=%d", frames);
  BSED_REQUIRE((sc_mean == nullptr) == (sc_std == nullptr), "logmel: scaler mean/std must come together");
  auto aligned16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  BSED_REQUIRE(aligned16(mel) && aligned16(out) && aligned16(sc_mean) && aligned16(sc_std) && aligned16(ws),
               "logmel: buffers must be 16-byte aligned (float4 accesses)");
  const size_t need = sizeof(double) * (size_t)B * (kNMels + 1);
  if (ws_bytes < need) {
    bsed_set_error("logmel: workspace %zu < %zu", ws_bytes, need);
    return BSED_E_WORKSPACE;
  }
  BSED_CHECK_CUDA(cudaMemsetAsync(ws, 0, need, st));
  BSED_TRY(melspec_impl(h, audio, B, n_samples, mel, (double*)ws, st));
  const int t_in = 1 + n_samples / kHop;
  ProfScope prof(PROF_ELEMENTWISE, 0.0, 4.0 * B * kNMels * ((double)t_in + (double)frames), st);
  dim3 grid(ceil_div((long long)frames * (kNMels / 4), 256), B);
  db_apply_kernel<false><<<grid, 256, 0, st>>>(mel, nullptr, 1.0, t_in, frames, sc_mean, sc_std, (const double*)ws, out);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// threshold -> binary median over time -> contiguous regions        (Appendix B of SURVEY.md)
// one CTA per clip; T <= 1024, C <= 32
// ---------------------------------------------------------------------------------------------
constexpr int PP_MAXT = 1024;
constexpr int PP_MAXC = 32;

__global__ void __launch_bounds__(256) median_decode_kernel(const float* __restrict__ strong, int T, int C,
                                                            float threshold, int win, int32_t* __restrict__ events,
                                                            int max_events, int32_t* __restrict__ n_events) {
  extern __shared__ unsigned char pp_smem[];
  unsigned char* bin = pp_smem;                       // [C][T]
  unsigned char* med = bin + PP_MAXC * PP_MAXT;       // [C][T]
  __shared__ int counts[PP_MAXC];
  __shared__ int offsets[PP_MAXC + 1];
  const int b = blockIdx.x;
  const float* p = strong + (size_t)b * T * C;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    int t = i / C, c = i % C;
    bin[c * PP_MAXT + t] = p[i] >= threshold ? 1 : 0;
  }
  __syncthreads();
  const int left = win / 2, right = win - 1 - left, need = win - win / 2;
  for (int i = threadIdx.x; i < T * C; i += blockDim.x) {
    int c = i / T, t = i % T;
    int ones = 0;
    for (int u = t - left; u <= t + right; ++u) {
      int r = u % (2 * T);
      if (r < 0) r += 2 * T;
      if (r >= T) r = 2 * T - 1 - r;
      ones += bin[c * PP_MAXT + r];
    }
    med[c * PP_MAXT + t] = ones >= need ? 1 : 0;
  }
  __syncthreads();
  // pass A: count runs per class
  if (threadIdx.x < C) {
    int c = threadIdx.x, n = 0;
    unsigned char prev = 0;
    for (int t = 0; t < T; ++t) {
      unsigned char cur = med[c * PP_MAXT + t];
      if (cur && !prev) ++n;
      prev = cur;
    }
    counts[c] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < C; ++c) {
      offsets[c] = acc;
      acc += counts[c];
    }
    offsets[C] = acc;
    n_events[b] = acc;
  }
  __syncthreads();
  // pass B: emit (class, on, off), class-major then time
  if (threadIdx.x < C) {
    int c = threadIdx.x, k = offsets[c], on = -1;
    int32_t* ev = events + (size_t)b * max_events * 3;
    for (int t = 0; t <= T; ++t) {
      unsigned char cur = t < T ? med[c * PP_MAXT + t] : 0;
      if (cur && on < 0) on = t;
      if (!cur && on >= 0) {
        if (k < max_events) {
          ev[k * 3 + 0] = c;
          ev[k * 3 + 1] = on;
          ev[k * 3 + 2] = t;
        }
        ++k;
        on = -1;
      }
    }
  }
}

int median_decode(const float* strong, int B, int T, int C, float threshold, int win, int32_t* events,
                  int max_events, int32_t* n_events, cudaStream_t st) {
  BSED_REQUIRE(T > 0 && T <= PP_MAXT && C > 0 && C <= PP_MAXC, "median_decode: T=%d C=%d (max 1024 x 32)", T, C);
  BSED_REQUIRE(win >= 1 && win <= 65536, "median_decode: win=%d", win);
  BSED_REQUIRE(B > 0 && max_events >= 0, "median_decode: B=%d", B);
  static bool configured[kMaxDevices] = {};
  size_t smem = 2 * PP_MAXC * PP_MAXT;
  if (first_use_on_device(configured)) {
    BSED_CHECK_CUDA(cudaFuncSetAttribute(median_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  median_decode_kernel<<<B, 256, smem, st>>>(strong, T, C, threshold, win, events, max_events, n_events);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace bsed
