// tc_conv.cu -- column-tiled implicit-GEMM 3x3 convolution (forward / data gradient) on tcgen05 tf32 with the nine
// taps served from ONE shared-memory copy of the input halo.
//
// The row-tiled kernel of tc_gemm.cu fetches every activation tile from L2 once per tap (9x), and with it the layer's
// weights once per 128 pixels: it is bound by the L2 -> SM TMA bandwidth (~8 TB/s chip-wide, profiles/r01_*), not by
// the tensor pipe.  Here a tile is a COLUMN segment: 128 consecutive time steps t at one frequency bin f, G = 2
// neighbouring bins per CTA step.  Per 32-channel chunk the producer loads the G + 2 columns f0-1 .. f0+G as boxes of
// 130 rows (t0-1 .. t0+128; out-of-range rows / columns are the TMA's zero fill = the convolution's padding).  A tap
// (dt, df) of bin f0+g is then just a descriptor into column copy g+df+1, starting (dt+1) rows in: 4 boxes of 130 rows
// replace 2 x 9 boxes of 128 rows (0.23x the activation traffic), and each weight tile feeds both bins (0.5x).
// Shifting a K-major SWIZZLE_128B operand by whole 128-byte rows needs nothing but the new start address: the tensor
// core applies the swizzle XOR to absolute shared-memory address bits (measured on B200: the descriptor's base-offset
// field must stay 0; setting it to (address >> 7) & 7 gives wrong products).
//
// Warp roles, pipelines and the epilogue (TMEM -> swizzled staging -> TMA store, batch statistics from the staged
// tile) follow tc_kmajor_kernel.
#include "tc_common.cuh"

namespace bsed {
namespace tc {

constexpr int kG = 2;               // frequency bins per CTA step (template G: 1 where 2 would leave half the SMs idle)
constexpr int kHaloRows = 130;      // t0-1 .. t0+128
constexpr int kHaloPad = 136;       // rows of one column copy, padded to the 8-row swizzle repeat
constexpr int kAStages = 2;
// Warp roles: 0 TMA producer, 1 and 6 MMA issuers (one frequency bin = one accumulator each), 2-5 epilogue, 7-8 operand
// splitters (3xTF32).  Two issuers because the tensor core's MMA queue is shallow: one thread that also waits on the weight
// ring, fences and commits between its MMAs leaves the pipe idle for all of that time (measured with tests/probes/
// umma_probe.cu: 74.6 clk per 128x128x8 MMA with one issuer, the 64.1 clk floor with two; in this kernel 104 clk).
constexpr int kColThreads = 7 * 32;
constexpr int kColThreadsX3 = 9 * 32;
// 3xTF32 (X3, see tc_gemm.cu): the input is walked in 16-channel chunks (64-byte rows, SWIZZLE_64B), so that the raw
// column copies, their low parts (one set per stage, written by two splitter warps) and a ring of (hi, lo) weight tiles
// fit the 227 KB of shared memory; every (tap, k-step) issues a*w_hi + a*w_lo + a_lo*w_hi.

struct CArgs {
  int n_tiles, fgroups, tblocks;   // tile = (clip, f group, t block)
  int T, F;
  int cpt;                         // 32-channel chunks of the input
  int rb;                          // weights resident in shared memory (all 9 * cpt tiles)
  int bstages;                     // weight ring depth (rb == 0)
  int debug;
  double* stats;
  int stats_groups, stats_c;       // statistics fold output column c onto channel c % stats_c
  int gfirst[kMaxGroups];
};

// K-major swizzled descriptor (rows of ROWB = 128 / 64 bytes) whose start may sit on any row of the swizzle repeat
template <int ROWB>
__device__ __forceinline__ uint64_t kmajor_desc_rows(uint32_t smem_addr) {
  constexpr uint64_t layout = ROWB == 128 ? 2 : 4;
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((8 * ROWB) >> 4) << 32) | (1ull << 46) |
         (layout << 61);
}

template <int N, bool X3, int G = kG>
struct CSmem {
  static constexpr int KCH = X3 ? 16 : 32;
  static constexpr int ROWB = KCH * 4;
  static constexpr int A_COPY = kHaloPad * ROWB;                      // one column copy (17 KB / 8.5 KB)
  static constexpr int A_STAGE = (G + 2) * A_COPY;
  static constexpr int LO_BYTES = X3 ? kAStages * A_STAGE : 0;        // low parts, one set per stage
  static constexpr int B_TILE = N * KCH * 4;                          // one tap, one chunk
  static constexpr int B_STRIDE = (B_TILE + 1023) / 1024 * 1024;
  static constexpr int NB = X3 ? 2 : 1;                               // weight tiles per (tap, chunk): hi [, lo]
  static constexpr int B_SLOT = NB * B_STRIDE;
  // columns staged per pass; 3xTF32 with 128 output channels stages 32 (16 KB) so that a fourth (hi, lo) weight slot fits
  static constexpr int HW = N < 64 ? N : (X3 && N == 128) ? 32 : 64;
  static constexpr int STG_BYTES = kBM * HW * 4;
  static constexpr int BAR_BYTES = 512;
};

template <int N, bool X3, int G>
__global__ void __launch_bounds__(X3 ? kColThreadsX3 : kColThreads, 1)
tc_conv_col_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   const __grid_constant__ CUtensorMap mapBlo, const __grid_constant__ CUtensorMap mapC,
                   const float* __restrict__ bias, CArgs a) {
  using S = CSmem<N, X3, G>;
  constexpr int KCH = S::KCH;
  constexpr int ROWB = S::ROWB;
  constexpr uint32_t TMEM_COLS = (2 * G * N <= 64) ? 64 : (2 * G * N <= 128) ? 128 : (2 * G * N <= 256) ? 256 : 512;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nk = 9 * a.cpt;
  const int nb = a.rb ? nk : a.bstages;                               // weight tiles held in shared memory
  unsigned char* alo = smem + kAStages * S::A_STAGE;                  // low parts of the column copies (X3)
  unsigned char* sB = alo + S::LO_BYTES;
  unsigned char* stg = sB + nb * S::B_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + S::STG_BYTES);
  uint64_t* afull = bars;
  uint64_t* aempty = bars + kAStages;
  uint64_t* bfull = bars + 2 * kAStages;          // up to 16 weight stages
  uint64_t* bempty = bfull + 16;
  uint64_t* tfull = bempty + 16;
  uint64_t* tempty = tfull + 2;
  uint64_t* rbfull = tempty + 2;
  uint64_t* lofull = rbfull + 1;                  // [kAStages] low parts written (X3)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(lofull + kAStages);
  float* sbias = reinterpret_cast<float*>(stg + S::STG_BYTES + S::BAR_BYTES);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    prefetch_tmap(&mapC);
    for (int s = 0; s < kAStages; ++s) {
      mbar_init(&afull[s], 1);
      mbar_init(&aempty[s], G);         // one tcgen05.commit per issuer warp
      mbar_init(&lofull[s], 1);
    }
    for (int s = 0; s < 16; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], G);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], G);
      mbar_init(&tempty[i], 4);
    }
    mbar_init(rbfull, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < N) sbias[threadIdx.x] = bias ? bias[threadIdx.x] : 0.f;
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int per_clip = a.fgroups * a.tblocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int sb = 0;
      uint32_t phb = 0;
      if (a.rb) {
        mbar_expect_tx(rbfull, S::NB * nk * S::B_TILE);
        for (int i = 0; i < nk; ++i) {
          tma_load_2d(&mapB, sB + i * S::B_SLOT, rbfull, i * KCH, 0);
          if (X3) tma_load_2d(&mapBlo, sB + i * S::B_SLOT + S::B_STRIDE, rbfull, i * KCH, 0);
        }
      }
      // Two independent request streams served by this one thread, polled without blocking: the activation columns of a
      // step are requested as soon as their stage is free (up to kAStages steps ahead of the MMAs), the weight tiles as
      // soon as their ring slot is.  (Blocking on the activation stage between two weight tiles drained the weight ring at
      // every step boundary: the MMAs of the next step then started behind a full L2 round trip.)  The column copies go
      // out one per loop turn, interleaved with the weight tiles, so that no 35-70 KB burst sits ahead of a weight tile
      // in the TMA queue.
      const int my_tiles = blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
      const int steps = my_tiles * a.cpt;
      const int total_b = a.rb ? 0 : steps * 9;
      int a_step = 0, a_copy = 0, b_idx = 0, b_tap = 0, b_ch = 0;
      uint32_t idle = 0;
      while (a_step < steps || b_idx < total_b) {
        bool progressed = false;
        if (a_step < steps) {
          const int stage = a_step % kAStages;
          if (a_copy > 0 || mbar_test(&aempty[stage], ((a_step / kAStages) & 1) ^ 1)) {
            const int tile = blockIdx.x + (a_step / a.cpt) * gridDim.x, ch = a_step % a.cpt;
            const int b = tile / per_clip, r = tile - b * per_clip;
            const int f0 = (r / a.tblocks) * G, t0 = (r % a.tblocks) * kBM;
            if (a_copy == 0) mbar_expect_tx(&afull[stage], (G + 2) * kHaloRows * KCH * 4);
            tma_load_4d(&mapA, smem + stage * S::A_STAGE + a_copy * S::A_COPY, &afull[stage], ch * KCH, f0 - 1 + a_copy, t0 - 1, b);
            if (++a_copy == G + 2) {
              a_copy = 0;
              ++a_step;
            }
            progressed = true;
          }
        }
        if (b_idx < total_b && mbar_test(&bempty[sb], phb ^ 1)) {
          if (a.debug == 2 && b_idx >= a.bstages) {   // experiment: the ring is never refilled
            mbar_arrive(&bfull[sb]);
          } else {
            mbar_expect_tx(&bfull[sb], S::NB * S::B_TILE);
            tma_load_2d(&mapB, sB + sb * S::B_SLOT, &bfull[sb], (b_tap * a.cpt + b_ch) * KCH, 0);
            if (X3) tma_load_2d(&mapBlo, sB + sb * S::B_SLOT + S::B_STRIDE, &bfull[sb], (b_tap * a.cpt + b_ch) * KCH, 0);
          }
          if (++sb == a.bstages) {
            sb = 0;
            phb ^= 1;
          }
          ++b_idx;
          if (++b_tap == 9) {
            b_tap = 0;
            if (++b_ch == a.cpt) b_ch = 0;
          }
          progressed = true;
        }
        if (progressed) idle = 0;
        else if (++idle > kSpinLimit) __trap();
      }
    }
  } else if (warp == 1 || (warp == 6 && G == 2)) {
    // ===================== MMA issuers: bin g = 0 (warp 1), g = 1 (warp 6) =====================
    const int g = warp == 1 ? 0 : 1;
    constexpr uint32_t idesc = idesc_tf32(N, 0, 0);
    int sa = 0, sb = 0;
    uint32_t pha = 0, phb = 0;
    int it = 0;
    if (a.rb) mbar_wait(rbfull, 0);
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t ab_ph = (it >> 1) & 1;
      mbar_wait(&tempty[ab], ab_ph ^ 1);
      tc_fence_after();
      for (int ch = 0; ch < a.cpt; ++ch) {
        mbar_wait(&afull[sa], pha);
        if (X3 && a.debug != 3) mbar_wait(&lofull[sa], pha);      // the splitters have written the low parts of this stage
        tc_fence_after();
        // descriptor low words ((address >> 4) | LBO): every tile of the step is these plus a compile-time constant
        const uint32_t a_w = kmajor_desc_lo(smem_u32(smem + sa * S::A_STAGE + g * S::A_COPY));
        const uint32_t l_w = kmajor_desc_lo(smem_u32(alo + sa * S::A_STAGE + g * S::A_COPY));
        const uint32_t d_tmem = tmem_base + (ab * G + g) * N;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dt = tap / 3 - 1, df = tap % 3 - 1;
          uint32_t b_w;
          if (a.rb) {
            b_w = kmajor_desc_lo(smem_u32(sB + (tap * a.cpt + ch) * S::B_SLOT));
          } else {
            mbar_wait(&bfull[sb], phb);
            tc_fence_after();
            b_w = kmajor_desc_lo(smem_u32(sB + sb * S::B_SLOT));
          }
          __syncwarp();
          if (elect_one()) {
            constexpr uint32_t DHI = kmajor_desc_hi<ROWB>();
            const uint32_t aoff = (uint32_t)((df + 1) * S::A_COPY + (dt + 1) * ROWB) >> 4;
#pragma unroll
            for (int k = 0; k < KCH / 8; ++k) {
              const uint32_t acc = (tap | k) != 0 ? 1u : (ch != 0 ? 1u : 0u);
              umma_tf32_w(d_tmem, a_w + aoff + ((k * 32) >> 4), b_w + ((k * 32) >> 4), DHI, idesc, acc);
              if (X3) {
                umma_tf32_w(d_tmem, a_w + aoff + ((k * 32) >> 4), b_w + ((S::B_STRIDE + k * 32) >> 4), DHI, idesc, 1u);
                umma_tf32_w(d_tmem, l_w + aoff + ((k * 32) >> 4), b_w + ((k * 32) >> 4), DHI, idesc, 1u);
              }
            }
            if (!a.rb) umma_commit(&bempty[sb]);
            if (tap == 8) {
              umma_commit(&aempty[sa]);
              if (ch == a.cpt - 1) umma_commit(&tfull[ab]);
            }
          }
          __syncwarp();
          if (!a.rb && ++sb == a.bstages) {
            sb = 0;
            phb ^= 1;
          }
        }
        if (++sa == kAStages) {
          sa = 0;
          pha ^= 1;
        }
      }
    }
  } else if (warp == 6) {
    // G == 1: no second accumulator, no second issuer
  } else if (X3 && warp >= 7) {
    // ===================== operand splitters (warps 7..8, 3xTF32) =====================
    // lo[sa] is free whenever A[sa] is: a stage is refilled only after the MMAs that read both have completed
    const int tid = threadIdx.x - 7 * 32;
    const int my_tiles = blockIdx.x < a.n_tiles ? (a.n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int steps = my_tiles * a.cpt;
    for (int step = 0; step < steps; ++step) {
      const int sa = step % kAStages;
      mbar_wait(&afull[sa], (step / kAStages) & 1);
      const uint32_t src = smem_u32(smem + sa * S::A_STAGE), dst = smem_u32(alo + sa * S::A_STAGE);
      split_lo_range<8>(src, dst, S::A_STAGE / 16, tid);
      fence_proxy_async();
      split_barrier();
      if (tid == 0) mbar_arrive(&lofull[sa]);
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    constexpr int HW = S::HW;
    constexpr int CW = N >= 32 ? 32 : 16;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const bool leader = threadIdx.x == 64;
    const uint32_t sbias_addr = smem_u32(sbias);
    const uint32_t stg_addr = smem_u32(stg);
    auto chunk_addr = [&](int r, int c4) -> uint32_t {
      if (CW == 32) return (uint32_t)(r * 128 + ((c4 ^ (r & 7)) << 4));
      return (uint32_t)(r * 64 + ((c4 ^ ((r >> 1) & 3)) << 4));
    };
    // thread `row` < HW owns output channels h * HW + row of every pass h for the statistics
    constexpr int NPASS = N / HW;
    double st_sum[NPASS], st_sq[NPASS];
#pragma unroll
    for (int h = 0; h < NPASS; ++h) st_sum[h] = st_sq[h] = 0.0;
    int st_grp = -1;
    auto flush_stats = [&]() {
      if (st_grp >= 0) {
#pragma unroll
        for (int h = 0; h < N / HW; ++h)
          if (row < HW) {
            const int c = (h * HW + row) % a.stats_c;
            atomicAdd(a.stats + ((size_t)st_grp * a.stats_c + c) * 2 + 0, st_sum[h]);
            atomicAdd(a.stats + ((size_t)st_grp * a.stats_c + c) * 2 + 1, st_sq[h]);
          }
      }
#pragma unroll
      for (int h = 0; h < NPASS; ++h) st_sum[h] = st_sq[h] = 0.0;
    };
    int it = 0, pass_no = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int ab = it & 1;
      const uint32_t ab_ph = (it >> 1) & 1;
      const int b = tile / per_clip, r = tile - b * per_clip;
      const int f0 = (r / a.tblocks) * G, t0 = (r % a.tblocks) * kBM;
      const int valid_rows = a.T - t0 < kBM ? a.T - t0 : kBM;
      if (a.stats) {
        int gi = 0;
#pragma unroll
        for (int k = 1; k < kMaxGroups; ++k)
          if (k < a.stats_groups && b >= a.gfirst[k]) gi = k;
        if (gi != st_grp) {
          flush_stats();
          st_grp = gi;
        }
      }
      mbar_wait(&tfull[ab], ab_ph);
      tc_fence_after();
      if (a.debug == 5) {   // experiment: accumulators dropped (main-loop time alone)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        continue;
      }
      for (int g = 0; g < G; ++g) {
        const uint32_t taddr = tmem_base + (ab * G + g) * N + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int h = 0; h < N / HW; ++h, ++pass_no) {
          if (pass_no > 0) {      // staging free: previous TMA stores have read it, statistics pass finished
            if (leader) bulk_wait_read0();
            epi_barrier();
          }
#pragma unroll 1
          for (int c0 = 0; c0 < HW; c0 += CW) {
            float v[CW];
            if constexpr (CW == 32) tmem_ld32(taddr + h * HW + c0, v);
            else tmem_ld16(taddr + h * HW + c0, v);
            const uint32_t sub = stg_addr + (c0 / CW) * (kBM * CW * 4);
            float4 bv[CW / 4];   // bias reads as one batch ahead of the stores (volatile asm keeps program order)
#pragma unroll
            for (int u = 0; u < CW / 4; ++u) bv[u] = lds128(sbias_addr + (h * HW + c0 + 4 * u) * 4);
#pragma unroll
            for (int u = 0; u < CW / 4; ++u)
              sts128(sub + chunk_addr(row, u), make_float4(v[4 * u] + bv[u].x, v[4 * u + 1] + bv[u].y, v[4 * u + 2] + bv[u].z, v[4 * u + 3] + bv[u].w));
          }
          if (g == G - 1 && h == N / HW - 1) {   // accumulators of this step fully read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[ab]);
          }
          fence_proxy_async();
          epi_barrier();
          if (leader && a.debug != 1) {
#pragma unroll 1
            for (int c0 = 0; c0 < HW; c0 += CW)
              tma_store_4d(&mapC, stg + (c0 / CW) * (kBM * CW * 4), h * HW + c0, f0 + g, t0, b, false);
            bulk_commit();
          }
          if (a.stats && row < HW) {
            const int c = row;
            const uint32_t cbase = stg_addr + (c / CW) * (kBM * CW * 4) + (c & 3) * 4;
            const int c4 = (c % CW) >> 2;
            float s1 = 0.f, s2 = 0.f;
            for (int rr = 0; rr < valid_rows; ++rr) {
              const float x = lds32(cbase + chunk_addr(rr, c4));
              s1 += x;
              s2 = fmaf(x, x, s2);
            }
            st_sum[h] += (double)s1;
            st_sq[h] += (double)s2;
          }
        }
      }
    }
    if (a.stats) flush_stats();
    if (leader) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int N, bool X3, int G>
static int launch_col(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mBlo, const CUtensorMap& mC,
                      const float* bias, CArgs& a, int sms, cudaStream_t st) {
  using S = CSmem<N, X3, G>;
  const int nk = 9 * a.cpt;
  const long long fixed = (long long)kAStages * S::A_STAGE + S::LO_BYTES + S::STG_BYTES + S::BAR_BYTES + 512 + 1024;
  const long long budget = 227 * 1024 - fixed;
  if ((long long)nk * S::B_SLOT <= budget && (long long)nk * S::B_SLOT <= 24 * 1024) {
    a.rb = 1;
    a.bstages = 0;
  } else {
    a.rb = 0;
    long long bs = budget / S::B_SLOT;
    a.bstages = (int)(bs > 9 ? 9 : bs);
    if (const char* e = getenv("BSED_COL_BSTAGES")) {      // measurement experiments only
      const int v = atoi(e);
      if (v >= 2 && v < a.bstages) a.bstages = v;
    }
    if (a.bstages < 2) {
      bsed_set_error("tc_conv_col: no room for the weight ring (N=%d)", N);
      return BSED_E_INVALID;
    }
  }
  const size_t smem_bytes = (size_t)(fixed + (long long)(a.rb ? nk : a.bstages) * S::B_SLOT);
  auto kern = tc_conv_col_kernel<N, X3, G>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    BSED_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  int grid = a.n_tiles < sms ? a.n_tiles : sms;
  kern<<<grid, X3 ? kColThreadsX3 : kColThreads, smem_bytes, st>>>(mA, mB, mBlo, mC, bias, a);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace tc

// The column-tiled kernel wins from F = 2 up in both precisions (128 -> 128 channels, 24 clips: F = 4 49 us against the
// row-tiled kernel's 110 us in 3xTF32, 27 against 35 single-pass; F = 2 with one bin per step 37 against 59 and 21 against 23).
static int col_min_f() {   // BSED_COL_MIN_F: measurement experiments
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BSED_COL_MIN_F");
    v = e ? atoi(e) : 2;
  }
  return v;
}

bool tc_conv_col_supported(int F, int Cin, int Cout) {
  return F >= col_min_f() && (F % tc::kG == 0 || Cout == 128) && Cin % 32 == 0 && (Cout == 16 || Cout == 32 || Cout == 64 || Cout == 128) &&
         !getenv("BSED_CONV_ROW_TILES");
}

// Y[B][T][F][Cout] = conv3x3(X[B][T][F][Cin], Wk) + bias ; Wk = K-major packed weights [Cout][9*Cin] (k = tap*Cin + ci)
// Wk_lo != nullptr selects 3xTF32 (Wk = tf32-rounded weights, Wk_lo = their fp32 remainders, same layout)
int tc_conv3x3_col(const float* X, const float* Wk, const float* Wk_lo, float* Y, int B, int T, int F, int Cin, int Cout,
                   const float* bias, double* stats, int stats_groups, const int* gfirst, int sms, cudaStream_t st, int stats_c) {
  BSED_REQUIRE(tc_conv_col_supported(F, Cin, Cout), "tc_conv3x3_col: F=%d Cin=%d Cout=%d", F, Cin, Cout);
  const int CW = Cout >= 32 ? 32 : 16;
  const bool x3 = Wk_lo != nullptr;
  const int KCH = x3 ? 16 : 32;
  CUtensorMap mA, mB, mBlo, mC;
  cuuint64_t dA[4] = {(cuuint64_t)Cin, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t sA[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)F * Cin * 4, (cuuint64_t)T * F * Cin * 4};
  cuuint32_t bA[4] = {(cuuint32_t)KCH, 1, (cuuint32_t)tc::kHaloRows, 1};
  BSED_TRY(tc::make_map(&mA, X, 4, dA, sA, bA, KCH * 4, x3));
  cuuint64_t dB[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
  cuuint64_t sB[1] = {(cuuint64_t)9 * Cin * 4};
  cuuint32_t bB[2] = {(cuuint32_t)KCH, (cuuint32_t)Cout};
  BSED_TRY(tc::make_map(&mB, Wk, 2, dB, sB, bB, KCH * 4, x3));
  BSED_TRY(tc::make_map(&mBlo, x3 ? Wk_lo : Wk, 2, dB, sB, bB, KCH * 4, x3));
  cuuint64_t dC[4] = {(cuuint64_t)Cout, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t sC[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)F * Cout * 4, (cuuint64_t)T * F * Cout * 4};
  cuuint32_t bC[4] = {(cuuint32_t)CW, 1, 128, 1};
  BSED_TRY(tc::make_map(&mC, Y, 4, dC, sC, bC, CW * 4, true));
  tc::CArgs a;
  // one bin per CTA step (G = 1, 128 output channels only) where two would leave half of the SMs without a tile
  a.tblocks = (T + 127) / 128;
  const bool g1 = Cout == 128 && (F % tc::kG != 0 || 2 * B * (F / tc::kG) * a.tblocks <= sms) && !getenv("BSED_COL_G2");
  a.fgroups = g1 ? F : F / tc::kG;
  a.n_tiles = B * a.fgroups * a.tblocks;
  a.T = T;
  a.F = F;
  a.cpt = Cin / KCH;
  a.debug = tc_debug();
  a.stats = stats;
  a.stats_groups = stats ? stats_groups : 0;
  a.stats_c = stats_c > 0 ? stats_c : Cout;
  for (int k = 0; k < kMaxGroups; ++k) a.gfirst[k] = stats && k < stats_groups ? gfirst[k] : 0;
  // algorithmic flops: in the pixel-pair view (stats_c = true channel count < Cout) half of the packed weight matrix
  // is structural zeros -- the convolution it stands for has Cin / 2 input channels
  const double algo = stats_c > 0 && stats_c < Cout ? 0.5 : 1.0;
  ProfScope prof(PROF_CONV, algo * 2.0 * B * T * F * Cout * 9.0 * Cin,
                 4.0 * ((double)B * T * F * Cin + (double)B * T * F * Cout + algo * 9.0 * Cin * Cout), st);
  if (x3) {
    switch (Cout) {
      case 16: return tc::launch_col<16, true, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
      case 32: return tc::launch_col<32, true, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
      case 64: return tc::launch_col<64, true, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
      default:
        return g1 ? tc::launch_col<128, true, 1>(mA, mB, mBlo, mC, bias, a, sms, st)
                  : tc::launch_col<128, true, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
    }
  }
  switch (Cout) {
    case 16: return tc::launch_col<16, false, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
    case 32: return tc::launch_col<32, false, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
    case 64: return tc::launch_col<64, false, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
    default:
      return g1 ? tc::launch_col<128, false, 1>(mA, mB, mBlo, mC, bias, a, sms, st)
                : tc::launch_col<128, false, 2>(mA, mB, mBlo, mC, bias, a, sms, st);
  }
}

}  // namespace bsed
