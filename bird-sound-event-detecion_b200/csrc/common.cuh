// common.cuh -- shared device/host helpers for libbsed (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bsed.h"

// ----------------------------------------------------------------------------------------------
// error plumbing (thread-local message, no exceptions across the ABI)
// ----------------------------------------------------------------------------------------------
void bsed_set_error(const char* fmt, ...);
void bsed_count_launch();  // every kernel launch of this library passes through BSED_CHECK_LAUNCH

// optional per-class device timing (CUDA events on the launching stream), see api.cu
enum { PROF_NONE = 0, PROF_CONV = 1, PROF_WGRAD = 2, PROF_GEMM = 3, PROF_GEMM_TN = 4, PROF_MELSPEC = 5, PROF_GRU = 6,
       PROF_ELEMENTWISE = 7 };
void bsed_prof_begin(int cls, double flops, double bytes, cudaStream_t st);
void bsed_prof_end(int cls, cudaStream_t st);
struct ProfScope {
  int cls;
  cudaStream_t st;
  ProfScope(int c, double flops, double bytes, cudaStream_t s) : cls(c), st(s) { bsed_prof_begin(c, flops, bytes, s); }
  ~ProfScope() { bsed_prof_end(cls, st); }
};

#define BSED_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      bsed_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return BSED_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define BSED_CHECK_LAUNCH()                                                                    \
  do {                                                                                         \
    bsed_count_launch();                                                                       \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      bsed_set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return BSED_E_CUDA;                                                                      \
    }                                                                                          \
  } while (0)

#define BSED_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      bsed_set_error(__VA_ARGS__);   \
      return BSED_E_INVALID;         \
    }                                \
  } while (0)

#define BSED_TRY(expr)          \
  do {                          \
    int _r = (expr);            \
    if (_r != BSED_OK) return _r; \
  } while (0)

// ----------------------------------------------------------------------------------------------
// context
// ----------------------------------------------------------------------------------------------
constexpr int kNFFT = 2048;
constexpr int kHop = 255;
constexpr int kNBins = 1025;
constexpr int kNMels = 128;
constexpr int kSampleRate = 32000;

struct bsed_context {
  int device;
  int num_sms;
  // frontend tables (device)
  float* window;        // [2048] symmetric Hamming
  float2* tw1024;       // [32][32] e^{-2 pi i lane c / 1024} at [c][lane]
  float2* tw2048;       // [513]  e^{-2 pi i k / 2048}
  float* mel_w;         // packed non-zero filterbank weights
  int* mel_start;       // [128] first bin of band m
  int* mel_len;         // [128] number of bins of band m
  int* mel_off;         // [128] offset of band m inside mel_w
  int mel_nnz;
  float2* mel_iv_w;     // interval form of the filterbank: (rising weight of band j, falling weight of band j-1) per bin
  int* mel_iv_start;    // [129] first entry of interval j in mel_iv_w
  int* mel_iv_len;      // [129] bins in interval j
  int mel_iv_bin0;      // bin of entry 0
  int mel_iv_n;         // entries
  int disc_precision;   // BSED_PRECISION_* of the Clip_Discriminator GEMMs (default FP32, see bsed_disc_set_precision)
  const bsed_step_state* step_state;   // device-resident per-iteration scalars (bsed_set_step_state) or nullptr
};

// A dropout key by value (host-computed: bsed_mix_key) or by reference into the device-resident step state
struct DropKey {
  uint32_t key;
  const uint32_t* dev;
#ifdef __CUDACC__
  __device__ __forceinline__ uint32_t get() const { return dev ? *dev : key; }
#endif
};

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// stateless dropout rule; restated in numpy in oracle/crnn.py:keep_mask
__device__ __forceinline__ bool bsed_keep(uint32_t idx, uint32_t key, uint32_t thresh) {
  uint32_t h = (idx * 0x9E3779B1u) ^ key;
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h >= thresh;
}

// ex2.approx + rcp.approx: ~3e-7 relative, a third of the instructions of the IEEE division (the gate kernels issue one
// per element and are as much instruction- as HBM-bound)
__device__ __forceinline__ float sigmoidf_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// Two independent fp32 FMAs in one instruction (sm_100: FFMA2): d = a * b + d per component, each with the single
// rounding of fmaf -- bit-identical to two fmaf calls, at half the issue slots and register-file reads.  The dot products
// of the GRU recurrence are bound by exactly those.
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

#endif  // __CUDACC__

// host-side key mixing for the dropout hash (same as oracle/crnn.py:mix_key)
static inline uint32_t bsed_mix_key(uint64_t seed, uint64_t step, uint64_t stream) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + step * 0xD1B54A32D192ED03ull +
               stream * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
static inline uint32_t bsed_drop_thresh(float p) {
  double t = (double)p * 4294967296.0;
  if (t >= 4294967295.0) return 0xFFFFFFFFu;
  if (t <= 0.0) return 0u;
  return (uint32_t)t;
}
