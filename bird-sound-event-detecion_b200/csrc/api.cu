// api.cu -- C ABI surface of libbsed.so (see include/bsed.h): context, error plumbing, and thin
// argument-checking wrappers around the launchers.
#include <math.h>
#include <stdarg.h>

#include <vector>

#include "launch.h"

using namespace bsed;

static thread_local char g_err[512] = "";

void bsed_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ---------------------------------------------------------------------------------------------
// launch counter and per-class event timing (bench.py: gpu_launches, roofline.achieved)
// ---------------------------------------------------------------------------------------------
#include <atomic>
static std::atomic<unsigned long long> g_launches{0};
void bsed_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {
struct ProfState {
  int cls = PROF_NONE;
  std::vector<cudaEvent_t> ev;   // pairs
  size_t used = 0;
  double flops = 0, bytes = 0;
  int launches = 0;
  bool open = false;
} g_prof;
}  // namespace

void bsed_prof_begin(int cls, double flops, double bytes, cudaStream_t st) {
  if (cls != g_prof.cls || g_prof.cls == PROF_NONE) return;
  if (g_prof.used + 2 > g_prof.ev.size()) {
    size_t old = g_prof.ev.size();
    g_prof.ev.resize(old + 512);
    for (size_t i = old; i < g_prof.ev.size(); ++i) cudaEventCreate(&g_prof.ev[i]);
  }
  cudaEventRecord(g_prof.ev[g_prof.used], st);
  g_prof.flops += flops;
  g_prof.bytes += bytes;
  g_prof.launches += 1;
  g_prof.open = true;
}
void bsed_prof_end(int cls, cudaStream_t st) {
  if (cls != g_prof.cls || !g_prof.open) return;
  cudaEventRecord(g_prof.ev[g_prof.used + 1], st);
  g_prof.used += 2;
  g_prof.open = false;
}

extern "C" uint64_t bsed_launch_count(void) { return g_launches.load(); }

extern "C" int bsed_profile_begin(int kernel_class) {
  g_prof.cls = kernel_class;
  g_prof.used = 0;
  g_prof.flops = g_prof.bytes = 0;
  g_prof.launches = 0;
  g_prof.open = false;
  return BSED_OK;
}

extern "C" int bsed_profile_end(double* total_ms, double* total_flops, double* total_bytes, int* n_launches) {
  double ms = 0;
  for (size_t i = 0; i + 1 < g_prof.used; i += 2) {
    cudaError_t e = cudaEventSynchronize(g_prof.ev[i + 1]);
    if (e != cudaSuccess) {
      bsed_set_error("profile_end: %s", cudaGetErrorString(e));
      return BSED_E_CUDA;
    }
    float t = 0;
    cudaEventElapsedTime(&t, g_prof.ev[i], g_prof.ev[i + 1]);
    ms += t;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = g_prof.flops;
  if (total_bytes) *total_bytes = g_prof.bytes;
  if (n_launches) *n_launches = g_prof.launches;
  g_prof.cls = PROF_NONE;
  return BSED_OK;
}

extern "C" int bsed_version(void) { return BSED_ABI_VERSION; }
extern "C" const char* bsed_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// constant tables (float64 on the host, rounded once to fp32)
// ---------------------------------------------------------------------------------------------
namespace {

double hz_to_mel(double f) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp;
}
double mel_to_hz(double m) {
  const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
  return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : f_sp * m;
}

template <class T>
int upload(T** dst, const std::vector<T>& v) {
  BSED_CHECK_CUDA(cudaMalloc((void**)dst, sizeof(T) * v.size()));
  BSED_CHECK_CUDA(cudaMemcpy(*dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice));
  return BSED_OK;
}

int build_tables(bsed_context* h) {
  const double PI = 3.14159265358979323846;
  std::vector<float> win(kNFFT);
  for (int n = 0; n < kNFFT; ++n) win[n] = (float)(0.54 - 0.46 * cos(2.0 * PI * n / (kNFFT - 1)));
  std::vector<float2> t1(1024), t2(513);
  // inter-pass twiddles of the 32 x 32 decomposition, [c][lane] = e^{-2 pi i lane c / 1024} (conflict-free per c)
  for (int c = 0; c < 32; ++c)
    for (int l = 0; l < 32; ++l) {
      const int j = l * c;
      t1[c * 32 + l] = make_float2((float)cos(2.0 * PI * j / 1024), (float)(-sin(2.0 * PI * j / 1024)));
    }
  for (int k = 0; k <= 512; ++k) t2[k] = make_float2((float)cos(2.0 * PI * k / 2048), (float)(-sin(2.0 * PI * k / 2048)));
  // Slaney filterbank, librosa.filters.mel(sr=32000, n_fft=2048, n_mels=128, fmin=0, fmax=16000,
  // htk=False, norm=None): float64 ramps stored as float32
  const int nm = kNMels;
  std::vector<double> edges(nm + 2);
  const double mlo = hz_to_mel(0.0), mhi = hz_to_mel(16000.0);
  for (int i = 0; i < nm + 2; ++i) edges[i] = mel_to_hz(mlo + (mhi - mlo) * i / (nm + 1));
  std::vector<float> w;
  std::vector<int> start(nm), len(nm), off(nm);
  for (int m = 0; m < nm; ++m) {
    int s = -1, e = -1;
    std::vector<float> row(kNBins);
    for (int k = 0; k < kNBins; ++k) {
      double f = (double)k * kSampleRate / kNFFT;
      double lower = (f - edges[m]) / (edges[m + 1] - edges[m]);
      double upper = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
      double v = fmax(0.0, fmin(lower, upper));
      row[k] = (float)v;
      if (row[k] != 0.f) {
        if (s < 0) s = k;
        e = k;
      }
    }
    if (s < 0) {
      s = 0;
      e = -1;
    }
    start[m] = s;
    len[m] = e - s + 1;
    off[m] = (int)w.size();
    for (int k = s; k <= e; ++k) w.push_back(row[k]);
  }
  h->mel_nnz = (int)w.size();
  BSED_REQUIRE(h->mel_nnz <= 2048, "mel filterbank has %d weights (> 2048)", h->mel_nnz);
  if (w.empty()) w.push_back(0.f);
  // The same filterbank by INTERVALS between consecutive band edges: a bin in [e_j, e_j+1) lies on the rising slope of
  // band j and on the falling slope of band j - 1 and on no other band, so one magnitude load serves both:
  //   U_j = sum_k up[k] |X[k]|,  D_j = sum_k down[k] |X[k]|  over the bins of interval j;  mel[m] = U_m + D_(m+1).
  // iv_w[k] = (W[j][k], W[j-1][k]) for the bins in interval order (= bin order), iv_start / iv_len per interval j = 0..128.
  {
    std::vector<float2> ivw;
    std::vector<int> ivs(nm + 1), ivl(nm + 1);
    int nnz = 0;
    int k = 0;
    auto weight = [&](int m, int kk) -> float {
      if (m < 0 || m >= nm) return 0.f;
      double f = (double)kk * kSampleRate / kNFFT;
      double lower = (f - edges[m]) / (edges[m + 1] - edges[m]);
      double upper = (edges[m + 2] - f) / (edges[m + 2] - edges[m + 1]);
      return (float)fmax(0.0, fmin(lower, upper));
    };
    for (int j = 0; j <= nm; ++j) {
      while (k < kNBins && (double)k * kSampleRate / kNFFT < edges[j]) ++k;
      ivs[j] = (int)ivw.size();
      int kk = k;
      while (kk < kNBins && (double)kk * kSampleRate / kNFFT < edges[j + 1]) {
        float2 w2 = make_float2(weight(j, kk), weight(j - 1, kk));
        nnz += (w2.x != 0.f) + (w2.y != 0.f);
        ivw.push_back(w2);
        ++kk;
      }
      ivl[j] = kk - k;
      // bins of interval j are k .. kk-1; their position in iv_w equals (bin - first bin of interval 0)
      k = kk;
    }
    BSED_REQUIRE(nnz == h->mel_nnz, "mel interval tables hold %d non-zero weights, the filterbank %d", nnz, h->mel_nnz);
    BSED_REQUIRE(ivl[nm] <= 32 && (int)ivw.size() <= 1028, "mel interval tables: top interval %d bins, %zu entries", ivl[nm], ivw.size());
    // first bin of interval 0 (bins below e_0 = 0 Hz: none) -- iv_w index i corresponds to bin iv_bin0 + i
    int bin0 = 0;
    while (bin0 < kNBins && (double)bin0 * kSampleRate / kNFFT < edges[0]) ++bin0;
    h->mel_iv_bin0 = bin0;
    h->mel_iv_n = (int)ivw.size();
    if (ivw.empty()) ivw.push_back(make_float2(0.f, 0.f));
    BSED_TRY(upload(&h->mel_iv_w, ivw));
    BSED_TRY(upload(&h->mel_iv_start, ivs));
    BSED_TRY(upload(&h->mel_iv_len, ivl));
  }
  BSED_TRY(upload(&h->window, win));
  BSED_TRY(upload(&h->tw1024, t1));
  BSED_TRY(upload(&h->tw2048, t2));
  BSED_TRY(upload(&h->mel_w, w));
  BSED_TRY(upload(&h->mel_start, start));
  BSED_TRY(upload(&h->mel_len, len));
  BSED_TRY(upload(&h->mel_off, off));
  return BSED_OK;
}

}  // namespace

extern "C" int bsed_create(int device, bsed_handle* out) {
  BSED_REQUIRE(out, "bsed_create: null out");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    bsed_set_error("bsed_create: no usable CUDA device (%s); this library has no CPU fallback",
                   e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return BSED_E_CUDA;
  }
  BSED_REQUIRE(device >= 0 && device < count, "bsed_create: device %d out of range [0,%d)", device, count);
  BSED_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BSED_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    bsed_set_error("bsed_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                   prop.minor);
    return BSED_E_CUDA;
  }
  bsed_context* h = new bsed_context();
  memset(h, 0, sizeof(*h));
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->disc_precision = BSED_PRECISION_FP32;
  int r = build_tables(h);
  if (r != BSED_OK) {
    delete h;
    return r;
  }
  *out = h;
  return BSED_OK;
}

extern "C" int bsed_destroy(bsed_handle h) {
  if (!h) return BSED_OK;
  cudaFree(h->window);
  cudaFree(h->tw1024);
  cudaFree(h->tw2048);
  cudaFree(h->mel_w);
  cudaFree(h->mel_start);
  cudaFree(h->mel_iv_w);
  cudaFree(h->mel_iv_start);
  cudaFree(h->mel_iv_len);
  cudaFree(h->mel_len);
  cudaFree(h->mel_off);
  delete h;
  return BSED_OK;
}

// ---------------------------------------------------------------------------------------------
// frontend / post-processing
// ---------------------------------------------------------------------------------------------
extern "C" int bsed_frontend_n_frames(int n_samples) { return n_samples < 0 ? -1 : 1 + n_samples / kHop; }

extern "C" int bsed_melspec(bsed_handle h, const float* audio, int B, int n_samples, float* mel, void* stream) {
  BSED_REQUIRE(h && audio && mel, "bsed_melspec: null argument");
  return melspec(h, audio, B, n_samples, mel, as_stream(stream));
}

extern "C" size_t bsed_amp_to_db_workspace_bytes(int B) { return B > 0 ? sizeof(double) * (size_t)B * (kNMels + 1) : 0; }

extern "C" int bsed_logmel(bsed_handle h, const float* audio, int B, int n_samples, int frames, const float* scaler_mean,
                           const float* scaler_std, float* mel, float* out, void* workspace, size_t workspace_bytes,
                           void* stream) {
  BSED_REQUIRE(h && audio && mel && out && workspace, "bsed_logmel: null argument");
  BSED_REQUIRE(B > 0 && n_samples >= kNFFT / 2 + 1, "bsed_logmel: B=%d n_samples=%d", B, n_samples);
  return logmel(h, audio, B, n_samples, frames, scaler_mean, scaler_std, mel, out, workspace, workspace_bytes, as_stream(stream));
}

extern "C" int bsed_amp_to_db(bsed_handle h, const float* mel, const float* unit_noise, float snr_db, int B, int t_in,
                              int frames, const float* scaler_mean, const float* scaler_std, float* out,
                              void* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(h && mel && out && workspace, "bsed_amp_to_db: null argument");
  return amp_to_db(mel, unit_noise, snr_db, B, t_in, frames, scaler_mean, scaler_std, out, workspace, workspace_bytes,
                   as_stream(stream));
}

extern "C" int bsed_median_decode(bsed_handle h, const float* strong, int B, int T, int C, float threshold, int win,
                                  int32_t* events, int max_events, int32_t* n_events, void* stream) {
  BSED_REQUIRE(h && strong && events && n_events, "bsed_median_decode: null argument");
  return median_decode(strong, B, T, C, threshold, win, events, max_events, n_events, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------
// losses / optimiser
// ---------------------------------------------------------------------------------------------
extern "C" int bsed_mt_loss(bsed_handle h, const float* strong, const float* weak, int B, int T, int C, int syn_first,
                            int syn_n, const float* syn_target, int real_first, int real_n, const float* strong_ema,
                            const float* weak_ema, float cons_w, float* losses, float* d_strong, float* d_weak,
                            void* stream) {
  BSED_REQUIRE(h && strong && weak && losses && d_strong && d_weak, "bsed_mt_loss: null argument");
  BSED_REQUIRE(syn_n == 0 || syn_target, "bsed_mt_loss: syn_target missing");
  BSED_REQUIRE(real_n == 0 || (strong_ema && weak_ema), "bsed_mt_loss: teacher outputs missing");
  BSED_REQUIRE(syn_first >= 0 && syn_first + syn_n <= B && real_first >= 0 && real_first + real_n <= B,
               "bsed_mt_loss: clip ranges outside [0,B)");
  return mt_loss(strong, weak, B, T, C, syn_first, syn_n, syn_target, real_first, real_n, strong_ema, weak_ema, cons_w,
                 h->step_state, losses, d_strong, d_weak, as_stream(stream));
}

extern "C" int bsed_loss_terms(bsed_handle h, const float* strong, const float* weak, int B, int T, int C,
                               const bsed_loss_term* terms, int n_terms, float* losses, int n_slots, float* d_strong,
                               float* d_weak, void* stream) {
  BSED_REQUIRE(h && strong && weak && terms && losses && d_strong && d_weak, "bsed_loss_terms: null argument");
  BSED_REQUIRE(B >= 1 && T >= 1 && n_terms >= 1 && n_slots >= 1, "bsed_loss_terms: B=%d T=%d terms=%d slots=%d", B, T, n_terms,
               n_slots);
  return loss_terms(strong, weak, B, T, C, terms, n_terms, losses, n_slots, d_strong, d_weak, as_stream(stream));
}

extern "C" int bsed_roll_clips(bsed_handle h, const float* x, const int32_t* shift_t, const int32_t* shift_f, float* out,
                               int B, int T, int F, void* stream) {
  BSED_REQUIRE(h && x && out && x != out, "bsed_roll_clips: null or aliased argument");
  BSED_REQUIRE(B >= 1 && T >= 1 && F >= 1, "bsed_roll_clips: B=%d T=%d F=%d", B, T, F);
  return roll_clips(x, shift_t, shift_f, out, B, T, F, as_stream(stream));
}

extern "C" int bsed_step_state_advance(bsed_handle h, bsed_step_state* state, const bsed_step_cfg* cfg, void* stream) {
  BSED_REQUIRE(h && state && cfg, "bsed_step_state_advance: null argument");
  return step_state_advance(state, cfg, as_stream(stream));
}

extern "C" int bsed_set_step_state(bsed_handle h, const bsed_step_state* state) {
  BSED_REQUIRE(h, "bsed_set_step_state: null handle");
  h->step_state = state;
  return BSED_OK;
}

extern "C" int bsed_scale_f32(bsed_handle h, float* dst, const float* src, int64_t n, float alpha, void* stream) {
  BSED_REQUIRE(h && dst && src, "bsed_scale_f32: null argument");
  return scale_f32(dst, src, n, alpha, as_stream(stream));
}

extern "C" int bsed_opt_ema_step(bsed_handle h, float* params, const float* grads, float* m, float* v, float* ema,
                                 int64_t n, const bsed_opt_cfg* cfg, void* stream) {
  BSED_REQUIRE(h && params && grads && m && cfg, "bsed_opt_ema_step: null argument");
  BSED_REQUIRE(cfg->kind != 0 || v, "bsed_opt_ema_step: Adam needs v");
  return opt_ema_step(params, grads, m, v, ema, n, cfg, h->step_state, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------
// peer memory (CUDA IPC) for the fused data-parallel step
// ---------------------------------------------------------------------------------------------
extern "C" int bsed_ipc_export(bsed_handle h, const void* dev_ptr, unsigned char* handle, uint64_t* offset) {
  BSED_REQUIRE(h && dev_ptr && handle && offset, "bsed_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == BSED_IPC_HANDLE_BYTES, "IPC handle size");
  typedef int (*GetRange)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  BSED_CHECK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
  BSED_REQUIRE(fn && q == cudaDriverEntryPointSuccess, "bsed_ipc_export: cuMemGetAddressRange not available");
  unsigned long long base = 0;
  size_t size = 0;
  int r = reinterpret_cast<GetRange>(fn)(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
  BSED_REQUIRE(r == 0, "bsed_ipc_export: cuMemGetAddressRange failed (%d)", r);
  cudaIpcMemHandle_t hd;
  BSED_CHECK_CUDA(cudaIpcGetMemHandle(&hd, reinterpret_cast<void*>((uintptr_t)base)));
  memcpy(handle, &hd, sizeof(hd));
  *offset = (uint64_t)((uintptr_t)dev_ptr - (uintptr_t)base);
  return BSED_OK;
}

extern "C" int bsed_ipc_open(bsed_handle h, const unsigned char* handle, uint64_t offset, void** mapped) {
  BSED_REQUIRE(h && handle && mapped, "bsed_ipc_open: null argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle, sizeof(hd));
  void* base = nullptr;
  BSED_CHECK_CUDA(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
  *mapped = reinterpret_cast<unsigned char*>(base) + offset;
  return BSED_OK;
}

extern "C" int bsed_ipc_close(bsed_handle h, void* mapped, uint64_t offset) {
  BSED_REQUIRE(h && mapped, "bsed_ipc_close: null argument");
  BSED_CHECK_CUDA(cudaIpcCloseMemHandle(reinterpret_cast<unsigned char*>(mapped) - offset));
  return BSED_OK;
}

extern "C" int bsed_dp_opt_ema_step(bsed_handle h, int rank, int world, const float* const* peer_grads,
                                    float* const* peer_params, float* const* peer_ema, int32_t* const* peer_flags,
                                    int64_t epoch, float* m, float* v, int64_t n, const bsed_opt_cfg* cfg, void* stream) {
  BSED_REQUIRE(h && peer_grads && peer_params && peer_flags && m && cfg, "bsed_dp_opt_ema_step: null argument");
  BSED_REQUIRE(cfg->kind != 0 || v, "bsed_dp_opt_ema_step: Adam needs v");
  return dp_opt_ema_step(rank, world, peer_grads, peer_params, peer_ema, reinterpret_cast<int* const*>(peer_flags), epoch, m, v,
                         n, cfg, h->step_state, h->num_sms, as_stream(stream));
}

extern "C" int bsed_ema_buffers(bsed_handle h, const float* bn_buffers, float* ema_bn_buffers, int64_t n,
                                const int64_t* nbt, int64_t* ema_nbt, int n_nbt, float ema_alpha, int64_t ema_step,
                                void* stream) {
  BSED_REQUIRE(h && bn_buffers && ema_bn_buffers, "bsed_ema_buffers: null argument");
  return ema_buffers(bn_buffers, ema_bn_buffers, n, nbt, ema_nbt, n_nbt, ema_alpha, ema_step, h->step_state, as_stream(stream));
}

// ---------------------------------------------------------------------------------------------
// generic kernels (unit tests)
// ---------------------------------------------------------------------------------------------
extern "C" int bsed_gemm_nn(bsed_handle h, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M,
                            int N, int K, const float* bias, int accumulate, void* stream) {
  BSED_REQUIRE(h && A && Bm && C, "bsed_gemm_nn: null argument");
  return gemm_nn(A, lda, Bm, ldb, C, ldc, M, N, K, bias, accumulate, as_stream(stream));
}

extern "C" int bsed_gemm_tn(bsed_handle h, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M,
                            int N, int K, void* stream) {
  BSED_REQUIRE(h && A && Bm && C, "bsed_gemm_tn: null argument");
  return gemm_tn(A, lda, Bm, ldb, C, ldc, 1, M, N, K, h->num_sms * 4, as_stream(stream));
}

extern "C" int bsed_conv3x3(bsed_handle h, const float* x, const float* weight, const float* bias, float* y, int B,
                            int T, int F, int Cin, int Cout, float* wpack, void* stream) {
  BSED_REQUIRE(h && x && weight && y && wpack, "bsed_conv3x3: null argument");
  PrepTable tb;
  memset(&tb, 0, sizeof(tb));
  tb.n = 1;
  tb.ops[0].type = PREP_CONV_PACK;
  tb.ops[0].src = weight;
  tb.ops[0].dst = wpack;
  tb.ops[0].d0 = Cout;
  tb.ops[0].d1 = Cin;
  BSED_TRY(run_prep(tb, as_stream(stream)));
  return conv3x3_nn(x, wpack, y, B, T, F, Cin, Cout, bias, 0, as_stream(stream));
}

// tensor-core variants of the unit-test entry points (kind::tf32, so ~1e-3 relative to fp32)
extern "C" int bsed_conv3x3_tc(bsed_handle h, const float* x, const float* weight, const float* bias, float* y, int B,
                               int T, int F, int Cin, int Cout, float* wpack, void* stream) {
  BSED_REQUIRE(h && x && weight && y && wpack, "bsed_conv3x3_tc: null argument");
  PrepTable tb;
  memset(&tb, 0, sizeof(tb));
  tb.n = 1;
  tb.ops[0].type = PREP_CONV_KMAJOR;
  tb.ops[0].src = weight;
  tb.ops[0].dst = wpack;
  tb.ops[0].d0 = Cout;
  tb.ops[0].d1 = Cin;
  BSED_TRY(run_prep(tb, as_stream(stream)));
  return tc_conv3x3(x, wpack, nullptr, y, B, T, F, Cin, Cout, bias, 0, h->num_sms, as_stream(stream));
}

extern "C" int bsed_conv3x3_tc3(bsed_handle h, const float* x, const float* weight, const float* bias, float* y, int B,
                                int T, int F, int Cin, int Cout, float* wpack, void* stream) {
  BSED_REQUIRE(h && x && weight && y && wpack, "bsed_conv3x3_tc3: null argument");
  PrepTable tb;
  memset(&tb, 0, sizeof(tb));
  tb.n = 1;
  tb.ops[0].type = PREP_CONV_KMAJOR;
  tb.ops[0].src = weight;
  tb.ops[0].dst = wpack;
  tb.ops[0].d0 = Cout;
  tb.ops[0].d1 = Cin;
  BSED_TRY(run_prep(tb, as_stream(stream)));
  const long long n = 9LL * Cin * Cout;
  BSED_TRY(split_hi_lo(wpack, wpack + n, n, as_stream(stream)));
  return tc_conv3x3(x, wpack, wpack + n, y, B, T, F, Cin, Cout, bias, 0, h->num_sms, as_stream(stream));
}

extern "C" int bsed_gemm_nt_tc3(bsed_handle h, const float* A, int lda, const float* Bk, int ldb, float* C, int ldc,
                                int M, int N, int K, const float* bias, int accumulate, float* split_ws, void* stream) {
  BSED_REQUIRE(h && A && Bk && C && split_ws, "bsed_gemm_nt_tc3: null argument");
  const long long n = (long long)N * ldb;
  BSED_CHECK_CUDA(cudaMemcpyAsync(split_ws, Bk, sizeof(float) * n, cudaMemcpyDeviceToDevice, as_stream(stream)));
  BSED_TRY(split_hi_lo(split_ws, split_ws + n, n, as_stream(stream)));
  return tc_gemm_nt(A, lda, split_ws, split_ws + n, ldb, C, ldc, M, N, K, bias, accumulate, h->num_sms, as_stream(stream));
}

// C[M][N] += sum_k A[k][M]^T Bm[k][N] on tcgen05 (kind::tf32): the weight-gradient kernel with one tap and rows = k
extern "C" int bsed_gemm_tn_tc(bsed_handle h, const float* A, int lda, const float* Bm, int ldb, float* C, int ldc, int M, int N,
                               int64_t K, float* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(h && A && Bm && C && workspace, "bsed_gemm_tn_tc: null argument");
  BSED_REQUIRE(M % 32 == 0 && M >= 32 && N % 32 == 0 && N >= 32 && (N < 128 || N % 128 == 0), "bsed_gemm_tn_tc: M=%d N=%d", M, N);
  BSED_REQUIRE(K >= 1 && K < (1ll << 31), "bsed_gemm_tn_tc: K=%lld", (long long)K);
  TcOperand a{A, lda, 0, M}, b{Bm, ldb, 0, N};
  return tc_wgrad_ex(a, b, 1, (int)K, 1, 1, 0, C, ldc, 1, 0, workspace, workspace_bytes, h->num_sms, as_stream(stream));
}

extern "C" int bsed_gemm_nt_tc(bsed_handle h, const float* A, int lda, const float* Bk, int ldb, float* C, int ldc,
                               int M, int N, int K, const float* bias, int accumulate, void* stream) {
  BSED_REQUIRE(h && A && Bk && C, "bsed_gemm_nt_tc: null argument");
  return tc_gemm_nt(A, lda, Bk, nullptr, ldb, C, ldc, M, N, K, bias, accumulate, h->num_sms, as_stream(stream));
}

// conv weight gradient: dw (Cout,Cin,3,3) += ...; unit-test entry (tensor_cores = 0 -> fp32 SIMT split-K)
extern "C" int bsed_conv3x3_wgrad(bsed_handle h, const float* x, const float* dy, float* dw, int B, int T, int F, int Cin,
                                  int Cout, int tensor_cores, float* workspace, size_t workspace_bytes, void* stream) {
  BSED_REQUIRE(h && x && dy && dw, "bsed_conv3x3_wgrad: null argument");
  if (!tensor_cores) return conv3x3_wgrad(x, dy, dw, B, T, F, Cin, Cout, h->num_sms * 4, as_stream(stream));
  BSED_REQUIRE(workspace, "bsed_conv3x3_wgrad: workspace required for the tensor-core path");
  return tc_wgrad(x, dy, dw, (long long)Cin * 9, 9, 1, B, T, F, Cin, Cout, 9, workspace, workspace_bytes, h->num_sms,
                  as_stream(stream));
}
extern "C" size_t bsed_conv3x3_wgrad_workspace_bytes(bsed_handle h) { return h ? tc_wgrad_workspace_bytes(h->num_sms) : 0; }
