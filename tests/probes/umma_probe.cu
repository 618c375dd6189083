// umma_probe.cu -- measurement probe (not part of the library): cycles per tcgen05.mma kind::tf32 (M = 128, K = 8) issued
// back to back by one elected thread, both operands in shared memory, on every SM at once.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I bird-sound-event-detecion_b200/csrc -I include \
//        -o tests/probes/umma_probe tests/probes/umma_probe.cu -lcudart
// MODE 0: one A tile, one B tile, one accumulator          (a GEMM k-loop)
// MODE 1: the 3xTF32 triple  a*w_hi, a*w_lo, a_lo*w_hi      (two A tiles, two B tiles, one accumulator)
// MODE 2: MODE 0 with two accumulators used alternately
// MODE 3: MODE 0 while the four epilogue warps stream shared memory (ld.shared.v4 + st.shared.v4 over a 32 KB buffer)
// MODE 4: MODE 0 with the A tile starting 1 / 2 rows into the 8-row swizzle atom (the halo trick of tc_conv.cu)
// MODE 5: MODE 0 while warp 0 streams global memory into shared memory with 16 KB bulk copies (what the TMA producer does)
// MODE 6: MODE 4 + MODE 5
// MODE 10: the issue pattern of tc_conv_col_kernel<.., X3>: per "tap" elect.sync, 12 MMAs (two accumulators, 3xTF32 triple,
//          descriptors = base + constant), one tcgen05.commit
// MODE 11: MODE 10 + what precedes each tap in the kernel: mbarrier wait (already complete), tcgen05.fence, __syncwarp
// MODE 12: MODE 11 without the per-tap commit
// MODE 13: MODE 11 issued by TWO warps, one accumulator each (6 MMAs per tap and warp)
#include <cstdio>

#include "tc_common.cuh"

using namespace bsed::tc;

template <int N, int ROWB, int MODE>
__global__ void __launch_bounds__(192, 1) probe_kernel(unsigned long long* out, int iters, const unsigned char* src) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = 128 * ROWB, B_BYTES = N * ROWB;
  unsigned char* sA = smem;                         // two A tiles
  unsigned char* sB = smem + 2 * A_BYTES;           // two B tiles
  unsigned char* extra = sB + 2 * B_BYTES;          // 32 KB for MODE 3
  uint64_t* bar = reinterpret_cast<uint64_t*>(extra + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 6);
  __shared__ volatile int stop;
  const int warp = threadIdx.x / 32;
  for (int i = threadIdx.x; i < (2 * A_BYTES + 2 * B_BYTES + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    mbar_init(bar + 2, MODE == 13 ? 2 : 1);
    mbar_init(bar + 3, 1);
    mbar_init(bar + 4, 1);
    mbar_init(bar + 5, 1);
    fence_barrier_init();
    stop = 0;
  }
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if ((warp == 1 || warp == 2) && MODE == 13) {
    constexpr uint32_t idesc = idesc_tf32(N, 0, 0);
    constexpr uint32_t DHI = kmajor_desc_hi<ROWB>();
    const int g = warp - 1;
    uint64_t* tapbar = bar + 2;      // per-tap commits of both warps land here (count 2, nobody waits)
    uint64_t* ready = bar + 3;
    uint64_t* fin = bar + 4 + g;
    const uint32_t a_w = kmajor_desc_lo(smem_u32(sA)), l_w = kmajor_desc_lo(smem_u32(sA) + 128 * ROWB);
    long long t0 = 0;
    if (threadIdx.x % 32 == 0) t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t b_w = kmajor_desc_lo(smem_u32(sB) + (it & 1) * 2 * N * ROWB / 2);
      mbar_wait(ready, 1);
      tc_fence_after();
      __syncwarp();
      if (elect_one()) {
        const uint32_t aoff = (uint32_t)((it % 3) * ROWB) >> 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          umma_tf32_w(tmem + g * N, a_w + aoff + k * 2, b_w + k * 2, DHI, idesc, 1u);
          umma_tf32_w(tmem + g * N, a_w + aoff + k * 2, b_w + 4 + k * 2, DHI, idesc, 1u);
          umma_tf32_w(tmem + g * N, l_w + aoff + k * 2, b_w + k * 2, DHI, idesc, 1u);
        }
        umma_commit(tapbar);
      }
      __syncwarp();
    }
    if (threadIdx.x % 32 == 0) {
      umma_commit(fin);
      mbar_wait(fin, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[4 + g] = (unsigned long long)(t1 - t0);
    }
  } else if (warp == 1 && MODE >= 10) {
    constexpr uint32_t idesc = idesc_tf32(N, 0, 0);
    constexpr uint32_t DHI = kmajor_desc_hi<ROWB>();
    uint64_t* tapbar = bar + 2;      // per-tap commits land here (nobody waits)
    uint64_t* ready = bar + 3;       // never armed: parity 1 is "complete"
    const uint32_t a_w = kmajor_desc_lo(smem_u32(sA)), l_w = kmajor_desc_lo(smem_u32(sA) + 128 * ROWB);
    long long t0 = 0;
    if (threadIdx.x == 32) t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint32_t b_w = kmajor_desc_lo(smem_u32(sB) + (it & 1) * 2 * N * ROWB / 2);
      if (MODE >= 11) {
        mbar_wait(ready, 1);
        tc_fence_after();
        __syncwarp();
      }
      if (elect_one()) {
        const uint32_t aoff = (uint32_t)((it % 3) * ROWB) >> 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int g = 0; g < 2; ++g) umma_tf32_w(tmem + g * N, a_w + aoff + k * 2, b_w + k * 2, DHI, idesc, 1u);
#pragma unroll
          for (int g = 0; g < 2; ++g) umma_tf32_w(tmem + g * N, a_w + aoff + k * 2, b_w + 4 + k * 2, DHI, idesc, 1u);
#pragma unroll
          for (int g = 0; g < 2; ++g) umma_tf32_w(tmem + g * N, l_w + aoff + k * 2, b_w + k * 2, DHI, idesc, 1u);
        }
        if (MODE != 12) umma_commit(tapbar);
      }
      __syncwarp();
    }
    if (threadIdx.x == 32) {
      umma_commit(bar);
      mbar_wait(bar, 0);
      const long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_tf32(N, 0, 0);
    if (elect_one()) {
      const uint32_t a0 = smem_u32(sA), a1 = a0 + A_BYTES, b0 = smem_u32(sB), b1 = b0 + B_BYTES;
      const long long t0 = clock64();
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ROWB / 32; ++k) {
          const uint64_t da = kmajor_desc<ROWB>(a0 + k * 32), db = kmajor_desc<ROWB>(b0 + k * 32);
          if (MODE == 1) {
            umma_tf32(tmem, da, db, idesc, 1u);
            umma_tf32(tmem, da, kmajor_desc<ROWB>(b1 + k * 32), idesc, 1u);
            umma_tf32(tmem, kmajor_desc<ROWB>(a1 + k * 32), db, idesc, 1u);
          } else if (MODE == 2) {
            umma_tf32(tmem + (k & 1) * 256, da, db, idesc, 1u);
          } else if (MODE == 4 || MODE == 6) {
            umma_tf32(tmem, kmajor_desc<ROWB>(a0 + (1 + (it & 1)) * ROWB + k * 32), db, idesc, 1u);
          } else {
            umma_tf32(tmem, da, db, idesc, 1u);
          }
        }
      }
      umma_commit(bar);
      mbar_wait(bar, 0);
      const long long t1 = clock64();
      stop = 1;
      if (blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
    }
  } else if ((MODE == 5 || MODE == 6) && warp == 0) {
    if (elect_one()) {
      uint64_t* cbar = bar + 1;
      uint32_t ph = 0;
      unsigned long long n = 0;
      while (!stop) {
        mbar_expect_tx(cbar, 32768);
        for (int j = 0; j < 2; ++j)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(extra) + j * 16384),
                       "l"(src + ((n++ * 16384 + (size_t)blockIdx.x * 262144) & ((64u << 20) - 1))), "r"(16384), "r"(smem_u32(cbar))
                       : "memory");
        mbar_wait(cbar, ph);
        ph ^= 1;
      }
      if (blockIdx.x == 0) out[2] = n * 16384;
    }
  } else if (MODE == 3 && warp >= 2) {
    const uint32_t base = smem_u32(extra);
    const int tid = threadIdx.x - 64;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    while (!stop) {
#pragma unroll 4
      for (int i = tid; i < 32768 / 16; i += 128) {
        const float4 v = lds128(base + i * 16);
        acc.x += v.x;
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + i * 16), "f"(v.y), "f"(v.x), "f"(v.w), "f"(v.z) : "memory");
      }
    }
    if (acc.x == 123.f) out[1] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N, int ROWB, int MODE>
static void run(unsigned long long* d_out, int sms, const unsigned char* src = nullptr) {
  static unsigned char* g_src = nullptr;
  if (!g_src) {
    cudaMalloc(&g_src, 64u << 20);
    cudaMemset(g_src, 0, 64u << 20);
  }
  src = g_src;
  const int iters = 512, per_it = MODE >= 10 ? 12 : (ROWB / 32) * (MODE == 1 ? 3 : 1);
  const size_t smem = 2 * 128 * ROWB + 2 * N * ROWB + 32768 + 64 + 1024;
  auto kern = probe_kernel<N, ROWB, MODE>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  unsigned long long h = 0;
  for (int rep = 0; rep < 2; ++rep) {
    kern<<<sms, 192, smem>>>(d_out, iters, src);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("N=%d ROWB=%d MODE=%d: %s\n", N, ROWB, MODE, cudaGetErrorString(e));
      return;
    }
  }
  unsigned long long hh[6] = {0, 0, 0, 0, 0, 0};
  cudaMemcpy(hh, d_out, sizeof(hh), cudaMemcpyDeviceToHost);
  h = MODE == 13 ? (hh[4] > hh[5] ? hh[4] : hh[5]) : hh[0];
  const double cyc = (double)h / (iters * per_it);
  if (MODE >= 5) printf("   (bulk copies into shared memory meanwhile: %.1f B/clk per SM)\n", (double)hh[2] / (double)h);
  printf("N=%3d rows of %3d B, mode %d: %7.1f clk per MMA (128 x %d x 8; %d MMAs), operand bytes per MMA %d -> %.1f B/clk\n", N, ROWB, MODE,
         cyc, N, iters * per_it, 128 * 32 + N * 32, (128 * 32 + N * 32) / cyc);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* d_out;
  cudaMalloc(&d_out, 64);
  cudaMemset(d_out, 0, 64);
  printf("umma probe on %d SMs\n", sms);
  run<64, 128, 0>(d_out, sms);
  run<128, 128, 0>(d_out, sms);
  run<256, 128, 0>(d_out, sms);
  run<64, 64, 0>(d_out, sms);
  run<128, 64, 0>(d_out, sms);
  run<256, 64, 0>(d_out, sms);
  run<64, 64, 1>(d_out, sms);
  run<128, 64, 1>(d_out, sms);
  run<128, 128, 1>(d_out, sms);
  run<128, 128, 2>(d_out, sms);
  run<128, 64, 2>(d_out, sms);
  run<128, 128, 3>(d_out, sms);
  run<128, 64, 3>(d_out, sms);
  run<256, 128, 3>(d_out, sms);
  run<128, 128, 4>(d_out, sms);
  run<128, 64, 4>(d_out, sms);
  run<64, 64, 4>(d_out, sms);
  run<128, 128, 5>(d_out, sms);
  run<128, 64, 5>(d_out, sms);
  run<128, 64, 6>(d_out, sms);
  run<64, 64, 6>(d_out, sms);
  run<128, 64, 10>(d_out, sms);
  run<128, 64, 11>(d_out, sms);
  run<128, 64, 12>(d_out, sms);
  run<128, 64, 13>(d_out, sms);
  run<64, 64, 13>(d_out, sms);
  run<64, 64, 10>(d_out, sms);
  run<64, 64, 11>(d_out, sms);
  run<16, 128, 0>(d_out, sms);
  run<32, 128, 0>(d_out, sms);
  return 0;
}
