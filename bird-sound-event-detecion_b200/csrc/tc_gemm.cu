// tc_gemm.cu -- tcgen05 (5th-gen tensor core) GEMM family for sm_100a, operands fed by TMA.
//
//   D[128 x N] (fp32, TMEM) += A[128 x 8] * B[8 x N]      tcgen05.mma.cta_group::1.kind::tf32
//
// K-major kernel (tc_kmajor_kernel): implicit-GEMM 3x3 convolution forward / data gradient over
// channels-last activations, and plain row-major GEMMs (GLU linears).  One M-tile is 128 output
// pixels forming a th x tw rectangle of one clip; the A operand of tap (dt, df), channels
// [c, c+KCH) is ONE TMA box load of the 4-D activation tensor (C, F, T, B) at coordinates
// (c, df, t0 + dt, b): the box lands in shared memory as [128 pixels][KCH floats], i.e. exactly the
// K-major SWIZZLE_128B (KCH = 32) / SWIZZLE_64B (KCH = 16) operand tile the MMA wants, and the
// zero padding of the convolution is the TMA's out-of-bounds fill.  B is the [N][K] K-major packed
// weight matrix (2-D box).  Persistent CTAs, warp-specialised:
//   warp 0    TMA producer (one elected lane), STAGES-deep mbarrier ring
//   warp 1    TMEM allocation + MMA issue (one elected lane), double-buffered accumulators
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per warp) -> + bias -> global fp32
#include "tc_common.cuh"

namespace bsed {
namespace tc {

struct KArgs {
  int plain;           // 0: conv (4-D activation map), 1: plain GEMM (2-D row map)
  int n_tiles, tiles_per_clip, th;
  int T, F;            // conv geometry (tw == F)
  long long rows;      // plain: number of rows
  int ntaps, cpt;      // taps (9 or 1), k-chunks per tap
  int ldc;
  int accumulate;
  int debug;           // BSED_TC_DEBUG (measurement experiments only): 1 = skip the epilogue's global stores
  int n_chunks;        // plain mode: the output is n_chunks column blocks of N; tile = (row tile, column block)
  int stages;          // depth of the TMA ring (host: as many as fit the shared memory)
  int rb_bytes;        // bytes reserved for the resident weights (RB)
  // epi == 1: BatchNorm-backward epilogue (plain mode): C holds the direct gate path dxd on entry and
  //   dY = k * (dxd + acc - m1 - xhat * m2) on exit; (k, m1, m2) per group and column from `tab` [groups][3][N]
  int epi;
  // op_ring (epi == 1, N == 64 or 128): the two per-element operands (dxd = C on entry, xhat) reach the epilogue through a ring of
  // kOpSlots shared-memory slots filled by TMA one to two 32-column chunks ahead, and the result is written in place over
  // the dxd half of the slot and stored from there.  (Read by each thread straight from global memory they arrive one
  // chunk ahead at best: the epilogue then waits a DRAM round trip per chunk and the kernel runs at half the HBM rate.)
  int op_ring, ring_extra;   // ring_extra: bytes of the ring beyond the staging tile(s) it overlays
  const void* mX_host;       // host only: tensor map of xhat (same geometry as C's)
  const float* xh;           // xhat, same shape / leading dimension as C
  const float* tab;
  int tab_groups;
  long long rows_per_clip;   // rows of one clip (row -> clip -> group)
  int gfirst[kMaxGroups];    // first clip of each group, relative to row 0 (epi == 1) / to clip 0 (stats)
  // per-channel batch statistics of the output (conv mode): stats[(group * N + c) * 2 + {0,1}] += sum, sum of squares
  double* stats;
  int stats_groups;
  // epi == 2: GLU forward with the gate, dropout and average pool fused (4-D tiles, N = pack * C <= 64):
  //   lin = acc + bias (stored unless store_lin == 0);  o = lin * sigmoid(gamma * xhat + beta) * keep / (1 - p);
  //   pooled[to][fo][c] = mean of o over the pt x pf window.  tab = [gamma | beta], each replicated pack times.
  float* pooled;
  int pt, pf, pack, gC, To, Fo, store_lin;
  uint32_t drop_key, drop_thresh, drop_base;   // drop_base: element index of row 0 (the mask is keyed on absolute indices)
  float inv_keep;
};

// RB ("resident B"): the whole [N][K] weight matrix (<= kRbBytes) is loaded once per CTA and stays in shared
// memory; the ring then carries A tiles only.  Without it every 128-pixel tile re-fetches the weights from L2, which
// for the GLU linears and blocks 1-2 is 20-50 % of the L2 -> SM traffic these kernels are bound by.
constexpr int kRbBytes = 72 * 1024;

constexpr int kMaxStages = 24;
constexpr int kLoBufs = 3;
constexpr int kOpSlots = 3;                    // operand ring of the BatchNorm-backward epilogue (KArgs::op_ring)
constexpr int kOpSlotBytes = 2 * kBM * 32 * 4;   // one 128 x 32 chunk of dxd + one of xhat   // 3xTF32: low-part tiles in flight between the splitters and the MMA issuer

// X3 ("3xTF32"): error-compensated fp32-grade products on the tf32 tensor cores.  The activation tile arrives raw; two
// extra warps write its low part a - tf32(a) into a second tile; the weights come pre-split (hi, lo) from the prep pass;
// every k-step issues three MMAs: a*b_hi + a*b_lo + a_lo*b_hi.
template <int N, int KCH, bool RB, bool X3>
struct KSmem {
  static constexpr int A_BYTES = kBM * KCH * 4;
  static constexpr int B_BYTES = N * KCH * 4;
  static constexpr int B_STRIDE = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int NB = X3 ? 2 : 1;                                  // weight tiles per k-chunk (hi [, lo])
  static constexpr int STAGE = A_BYTES + (RB ? 0 : NB * B_STRIDE);
  static constexpr int LO_BYTES = X3 ? kLoBufs * A_BYTES : 0;
  static constexpr int STG_BYTES = kBM * N * 4;   // output tile staged for the TMA store
  static constexpr int STG2_BYTES = (!X3 && N <= 64) ? kBM * N * 4 : 0;   // gated tile of the fused GLU epilogue (epi == 2)
  static constexpr int BAR_BYTES = 512;   // 2 * stages + 5 + 2 * kLoBufs mbarriers + the TMEM slot
  static constexpr int TAB_BYTES = kMaxGroups * 3 * 128 * 4;   // BatchNorm-backward table
  static constexpr int BIAS_BYTES = 4096;   // up to 1024 bias values (n_chunks * N)
  static constexpr int FIXED = LO_BYTES + STG_BYTES + STG2_BYTES + 1024 /*align slack*/ + BAR_BYTES + BIAS_BYTES + TAB_BYTES;
  static_assert((2 * kMaxStages + 6 + 2 * kLoBufs + kOpSlots) * 8 + 8 <= BAR_BYTES, "barrier region too small");
};

template <int N, int KCH, bool RB, bool X3>
__global__ void __launch_bounds__(X3 ? kThreadsX3 : kThreads, 1)
tc_kmajor_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 const __grid_constant__ CUtensorMap mapBlo, const __grid_constant__ CUtensorMap mapC,
                 const __grid_constant__ CUtensorMap mapX, float* __restrict__ Y, const float* __restrict__ bias, KArgs a) {
  using S = KSmem<N, KCH, RB, X3>;
  constexpr int ROWB = KCH * 4;
  constexpr int NT = X3 ? kThreadsX3 : kThreads;
  constexpr uint32_t TMEM_COLS = (2 * N <= 32) ? 32 : (2 * N <= 64) ? 64 : (2 * N <= 128) ? 128 : (2 * N <= 256) ? 256 : 512;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int STAGES = a.stages;
  unsigned char* rb = smem + STAGES * S::STAGE;   // resident weights (RB): hi tiles, then (X3) lo tiles
  unsigned char* lo = rb + a.rb_bytes;            // low-part activation tiles (X3)
  unsigned char* stg = lo + S::LO_BYTES;          // output staging (1024-byte aligned: every region above is)
  unsigned char* stg2 = stg + S::STG_BYTES;
  unsigned char* after_stg = stg2 + S::STG2_BYTES + a.ring_extra;   // the operand ring overlays stg, stg2 and ring_extra
  uint64_t* bars = reinterpret_cast<uint64_t*>(after_stg);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint64_t* rbfull = bars + 2 * STAGES + 4;
  uint64_t* lofull = bars + 2 * STAGES + 5;
  uint64_t* loempty = lofull + kLoBufs;
  uint64_t* opfull = loempty + kLoBufs;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(opfull + kOpSlots);
  float* sbias = reinterpret_cast<float*>(after_stg + S::BAR_BYTES);   // bias staged once per CTA

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    prefetch_tmap(&mapC);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    for (int i = 0; i < kLoBufs; ++i) {
      mbar_init(&lofull[i], 1);
      mbar_init(&loempty[i], 1);
    }
    for (int i = 0; i < kOpSlots; ++i) mbar_init(&opfull[i], 1);
    mbar_init(rbfull, 1);
    fence_barrier_init();
    if (a.op_ring) prefetch_tmap(&mapX);
  }
  for (int i = threadIdx.x; i < N * a.n_chunks; i += NT) sbias[i] = bias ? bias[i] : 0.f;
  float* stab = sbias + S::BIAS_BYTES / 4;
  if (a.epi == 1)
    for (int i = threadIdx.x; i < a.tab_groups * 3 * N; i += NT) stab[i] = a.tab[i];
  if (a.epi == 2)
    for (int i = threadIdx.x; i < 2 * N; i += NT) stab[i] = a.tab[i];
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nk = a.ntaps * a.cpt;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      if (RB) {
        mbar_expect_tx(rbfull, S::NB * nk * S::B_BYTES);
        for (int i = 0; i < nk; ++i) tma_load_2d(&mapB, rb + i * S::B_STRIDE, rbfull, i * KCH, 0);
        if (X3)
          for (int i = 0; i < nk; ++i) tma_load_2d(&mapBlo, rb + (nk + i) * S::B_STRIDE, rbfull, i * KCH, 0);
      }
      for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
        int b = 0, t0 = 0;
        if (!a.plain) {
          b = tile / a.tiles_per_clip;
          t0 = (tile - b * a.tiles_per_clip) * a.th;
        }
        for (int tap = 0; tap < a.ntaps; ++tap) {
          const int dt = a.ntaps == 9 ? tap / 3 - 1 : 0, df = a.ntaps == 9 ? tap % 3 - 1 : 0;
          for (int ch = 0; ch < a.cpt; ++ch) {
            mbar_wait(&empty[s], ph ^ 1);
            unsigned char* sa = smem + s * S::STAGE;
            unsigned char* sb = sa + S::A_BYTES;
            mbar_expect_tx(&full[s], S::A_BYTES + (RB ? 0 : S::NB * S::B_BYTES));
            if (a.plain) tma_load_2d(&mapA, sa, &full[s], ch * KCH, (tile / a.n_chunks) * kBM);
            else tma_load_4d(&mapA, sa, &full[s], ch * KCH, df, t0 + dt, b);
            const int brow = a.plain ? (tile % a.n_chunks) * N : 0;     // column block of the output = row block of B
            if (!RB) tma_load_2d(&mapB, sb, &full[s], (tap * a.cpt + ch) * KCH, brow);
            if (!RB && X3) tma_load_2d(&mapBlo, sb + S::B_STRIDE, &full[s], (tap * a.cpt + ch) * KCH, brow);
            if (++s == STAGES) {
              s = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = idesc_tf32(N, 0, 0);
    int s = 0, lj = 0;
    uint32_t ph = 0, lph = 0;
    int it = 0;
    if (RB) mbar_wait(rbfull, 0);
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * N;
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&full[s], ph);
        if (X3) mbar_wait(&lofull[lj], lph);   // the splitters have written the low part of this stage
        tc_fence_after();
        __syncwarp();
        if (elect_one()) {   // one fixed lane issues the MMAs and their commits (commit tracks the issuing thread)
          const uint32_t sa = smem_u32(smem + s * S::STAGE);
          const uint32_t sb = RB ? smem_u32(rb + kc * S::B_STRIDE) : sa + S::A_BYTES;
          const uint32_t sblo = RB ? smem_u32(rb + (nk + kc) * S::B_STRIDE) : sb + S::B_STRIDE;
          const uint32_t salo = smem_u32(lo + lj * S::A_BYTES);
          // BSED_TC_DEBUG=4 (measurement experiment, wrong results): every other MMA accumulates into the OTHER buffer
          const uint32_t d_alt = a.debug == 4 ? tmem_base + (acc ^ 1) * N : d_tmem;
#pragma unroll
          for (int k = 0; k < KCH / 8; ++k) {
            uint64_t da = kmajor_desc<ROWB>(sa + k * 32);
            uint64_t db = kmajor_desc<ROWB>(sb + k * 32);
            umma_tf32((k & 1) ? d_alt : d_tmem, da, db, idesc, (kc | k) != 0 ? 1u : 0u);
            if (X3) {
              umma_tf32((k & 1) ? d_tmem : d_alt, da, kmajor_desc<ROWB>(sblo + k * 32), idesc, 1u);
              umma_tf32((k & 1) ? d_alt : d_tmem, kmajor_desc<ROWB>(salo + k * 32), db, idesc, 1u);
            }
          }
          umma_commit(&empty[s]);                      // frees the stage once these MMAs have read it
          if (X3) umma_commit(&loempty[lj]);
          if (kc == nk - 1) umma_commit(&tfull[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
        if (X3 && ++lj == kLoBufs) {
          lj = 0;
          lph ^= 1;
        }
      }
    }
  } else if (X3 && warp >= 6) {
    // ===================== operand splitters (warps 6..7, 3xTF32) =====================
    const int tid = threadIdx.x - 6 * 32;
    int s = 0, lj = 0;
    uint32_t ph = 0, lph = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < nk; ++kc) {
        mbar_wait(&full[s], ph);
        mbar_wait(&loempty[lj], lph ^ 1);
        const uint32_t src = smem_u32(smem + s * S::STAGE), dst = smem_u32(lo + lj * S::A_BYTES);
        split_lo_range<8>(src, dst, S::A_BYTES / 16, tid);
        fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
        split_barrier();
        if (tid == 0) mbar_arrive(&lofull[lj]);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
        if (++lj == kLoBufs) {
          lj = 0;
          lph ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    // TMEM -> registers -> (+ bias | BatchNorm backward) -> swizzled staging tile in shared memory -> TMA store
    // (TMA reduce-add in accumulate mode: C is never read by the SM).  Column statistics of the tile (conv forward in
    // train mode) are taken from the staged tile: thread c owns column c.
    constexpr int CW = N >= 32 ? 32 : 16;       // columns per TMEM load / per TMA store box
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;              // row of the 128-row tile
    const bool leader = threadIdx.x == 64;
    const uint32_t sbias_addr = smem_u32(sbias);
    const uint32_t stg_addr = smem_u32(stg);
    // swizzled position of 16-byte chunk c4 of this thread's row inside one [128][CW] staging sub-tile
    auto chunk_addr = [&](int r, int c4) -> uint32_t {
      if (CW == 32) return (uint32_t)(r * 128 + ((c4 ^ (r & 7)) << 4));
      return (uint32_t)(r * 64 + ((c4 ^ ((r >> 1) & 3)) << 4));
    };
    double st_sum = 0.0, st_sq = 0.0;           // running column statistics (thread = column `row`)
    int st_grp = -1;
    auto flush_stats = [&]() {
      if (st_grp >= 0 && row < N) {
        atomicAdd(a.stats + ((size_t)st_grp * N + row) * 2 + 0, st_sum);
        atomicAdd(a.stats + ((size_t)st_grp * N + row) * 2 + 1, st_sq);
      }
      st_sum = st_sq = 0.0;
    };
    // ---- operand ring of the BatchNorm-backward epilogue (a.op_ring): chunk n = 32-column block (n % NCH) of this CTA's
    // tile number n / NCH lives in slot n % kOpSlots; the leader requests chunk n + kOpSlots - 1 right after it has
    // handed chunk n's result to the TMA store (the slot it refills is the one whose store was committed one chunk earlier)
    constexpr int NCH = N / CW;
    const bool ring = a.op_ring != 0;
    const uint32_t ring_addr = stg_addr;
    int opn = 0;
    auto request_ops = [&](int n) {
      const int tl = blockIdx.x + (n / NCH) * gridDim.x;
      if (tl >= a.n_tiles) return;
      const int slot = n % kOpSlots;
      unsigned char* dst = stg + slot * kOpSlotBytes;
      mbar_expect_tx(&opfull[slot], kOpSlotBytes);
      tma_load_2d(&mapC, dst, &opfull[slot], (n % NCH) * CW, tl * kBM);
      tma_load_2d(&mapX, dst + kOpSlotBytes / 2, &opfull[slot], (n % NCH) * CW, tl * kBM);
    };
    if (ring && leader)
      for (int n = 0; n < kOpSlots - 1; ++n) request_ops(n);
    int it = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      long long grow;                       // global output row
      bool valid;
      int b = 0, t0 = 0, valid_rows = kBM;
      const int mt = a.plain ? tile / a.n_chunks : tile;            // row tile
      const int ncol0 = a.plain ? (tile % a.n_chunks) * N : 0;      // first output column of the tile
      if (a.plain) {
        grow = (long long)mt * kBM + row;
        valid = grow < a.rows;
        if (a.rows - (long long)mt * kBM < kBM) valid_rows = (int)(a.rows - (long long)mt * kBM);
      } else {
        b = tile / a.tiles_per_clip;
        t0 = (tile - b * a.tiles_per_clip) * a.th;
        long long in_clip = (long long)t0 * a.F + row;
        valid = in_clip < (long long)a.T * a.F;
        grow = (long long)b * a.T * a.F + in_clip;
        if ((long long)a.T * a.F - (long long)t0 * a.F < kBM) valid_rows = (int)((long long)a.T * a.F - (long long)t0 * a.F);
      }
      // operands of the BatchNorm-backward epilogue: requested before the accumulator is awaited, next chunk's while
      // the current one is combined
      const bool bn_rd = (a.epi == 1 || a.epi == 2) && valid && !ring;   // operands read per element: (dxd, xhat) / xhat
      const float* yrow = Y + grow * a.ldc;
      const float* xrow = a.xh + grow * a.ldc;
      uint32_t tab_addr = 0;
      if (a.epi == 1) {
        const int clip = (int)(grow / a.rows_per_clip);
        int gi = 0;
#pragma unroll
        for (int k = 1; k < kMaxGroups; ++k)
          if (k < a.tab_groups && clip >= a.gfirst[k]) gi = k;
        tab_addr = smem_u32(stab) + gi * 3 * N * 4;
      }
      float4 cpre[CW / 4], xpre[CW / 4];
      if (bn_rd) {
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          if (a.epi == 1) cpre[j] = *reinterpret_cast<const float4*>(yrow + 4 * j);
          xpre[j] = *reinterpret_cast<const float4*>(xrow + 4 * j);
        }
      }
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      if (a.debug == 5) {   // experiment: accumulators dropped (main-loop time alone)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        continue;
      }
      // the staging tile is free once the previous tile's TMA stores have read it (and every thread is done with its
      // statistics pass)
      if (it > 0 && !ring) {
        if (leader) bulk_wait_read0();
        epi_barrier();
      }
      const uint32_t taddr = tmem_base + acc * N + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += CW) {
        float v[CW];
        if (a.debug == 6) {   // experiment: no TMEM read
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] = 0.f;
        } else {
          if constexpr (CW == 32) tmem_ld32(taddr + c0, v);
          else tmem_ld16(taddr + c0, v);
        }
        if (a.debug == 7) {   // experiment: TMEM read only
          if (v[0] == 123.456f) sts128(stg_addr, make_float4(v[1], v[2], v[3], v[4]));
          continue;
        }
        if (ring) {
          if constexpr (CW == 32) {
            const int slot = opn % kOpSlots;
            mbar_wait(&opfull[slot], (opn / kOpSlots) & 1);
            const uint32_t dxd_addr = ring_addr + slot * kOpSlotBytes, xh_addr = dxd_addr + kOpSlotBytes / 2;
#pragma unroll
            for (int j0 = 0; j0 < CW; j0 += 8) {   // two float4 per batch: reads first, then the in-place stores
              float4 kk[2], m1[2], m2[2], cv[2], xv[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int j = j0 + 4 * u;
                kk[u] = lds128(tab_addr + (c0 + j) * 4);
                m1[u] = lds128(tab_addr + (N + c0 + j) * 4);
                m2[u] = lds128(tab_addr + (2 * N + c0 + j) * 4);
                cv[u] = lds128(dxd_addr + chunk_addr(row, j / 4));
                xv[u] = lds128(xh_addr + chunk_addr(row, j / 4));
              }
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int j = j0 + 4 * u;
                sts128(dxd_addr + chunk_addr(row, j / 4),
                       make_float4(kk[u].x * (cv[u].x + v[j] - m1[u].x - xv[u].x * m2[u].x),
                                   kk[u].y * (cv[u].y + v[j + 1] - m1[u].y - xv[u].y * m2[u].y),
                                   kk[u].z * (cv[u].z + v[j + 2] - m1[u].z - xv[u].z * m2[u].z),
                                   kk[u].w * (cv[u].w + v[j + 3] - m1[u].w - xv[u].w * m2[u].w)));
              }
            }
            if (c0 + CW >= N) {   // accumulator fully read
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[acc]);
            }
            fence_proxy_async();
            epi_barrier();
            if (leader) {
              if (a.debug != 1) tma_store_2d(&mapC, stg + slot * kOpSlotBytes, c0, mt * kBM, false);
              bulk_commit();
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // chunk opn - 1's store has left its slot
              request_ops(opn + kOpSlots - 1);
            }
            ++opn;
          }
          continue;
        }
        float4 ccur[CW / 4], xcur[CW / 4];
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
          ccur[j] = bn_rd ? cpre[j] : make_float4(0.f, 0.f, 0.f, 0.f);
          xcur[j] = bn_rd ? xpre[j] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (bn_rd && c0 + CW < N) {
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            if (a.epi == 1) cpre[j] = *reinterpret_cast<const float4*>(yrow + c0 + CW + 4 * j);
            xpre[j] = *reinterpret_cast<const float4*>(xrow + c0 + CW + 4 * j);
          }
        }
        const uint32_t sub = stg_addr + (c0 / CW) * (kBM * CW * 4);
        // The shared-memory wrappers are volatile asm (program order): table / bias reads are issued as a batch ahead of
        // the stores, not interleaved with them -- interleaved, every float4 paid a shared-memory round trip and this
        // loop, not HBM, set the pace of the streaming GEMMs (measured: 117 us with it, 60 us without, M = 963840).
        if (a.epi == 1) {   // paced by its two global operands per element: the per-float4 order is the faster one here
#pragma unroll
          for (int j = 0; j < CW; j += 4) {
            const float4 kk = lds128(tab_addr + (c0 + j) * 4);
            const float4 m1 = lds128(tab_addr + (N + c0 + j) * 4);
            const float4 m2 = lds128(tab_addr + (2 * N + c0 + j) * 4);
            const float4 cv = ccur[j / 4], xv = xcur[j / 4];
            sts128(sub + chunk_addr(row, j / 4),
                   make_float4(kk.x * (cv.x + v[j] - m1.x - xv.x * m2.x), kk.y * (cv.y + v[j + 1] - m1.y - xv.y * m2.y),
                               kk.z * (cv.z + v[j + 2] - m1.z - xv.z * m2.z), kk.w * (cv.w + v[j + 3] - m1.w - xv.w * m2.w)));
          }
        } else if (S::STG2_BYTES > 0 && a.epi == 2) {
#pragma unroll
          for (int j = 0; j < CW; j += 4) {
            const float4 bv = lds128(sbias_addr + (ncol0 + c0 + j) * 4);
            const float4 o = make_float4(v[j] + bv.x, v[j + 1] + bv.y, v[j + 2] + bv.z, v[j + 3] + bv.w);
            sts128(sub + chunk_addr(row, j / 4), o);
            // gated value into the second staging tile
            const float4 ga = lds128(smem_u32(stab) + (c0 + j) * 4);
            const float4 be = lds128(smem_u32(stab) + (N + c0 + j) * 4);
            const float4 xv = xcur[j / 4];
            float gv[4] = {o.x * sigmoidf_(fmaf(ga.x, xv.x, be.x)), o.y * sigmoidf_(fmaf(ga.y, xv.y, be.y)),
                           o.z * sigmoidf_(fmaf(ga.z, xv.z, be.z)), o.w * sigmoidf_(fmaf(ga.w, xv.w, be.w))};
            if (a.drop_thresh) {
              const uint32_t e0 = a.drop_base + (uint32_t)(grow * N + c0 + j);
#pragma unroll
              for (int u = 0; u < 4; ++u) gv[u] = bsed_keep(e0 + u, a.drop_key, a.drop_thresh) ? gv[u] * a.inv_keep : 0.f;
            }
            if (!valid) gv[0] = gv[1] = gv[2] = gv[3] = 0.f;
            sts128(smem_u32(stg2) + (c0 / CW) * (kBM * CW * 4) + chunk_addr(row, j / 4), make_float4(gv[0], gv[1], gv[2], gv[3]));
          }
        } else {
          float4 bv[CW / 4];
#pragma unroll
          for (int u = 0; u < CW / 4; ++u) bv[u] = lds128(sbias_addr + (ncol0 + c0 + 4 * u) * 4);
#pragma unroll
          for (int u = 0; u < CW / 4; ++u)
            sts128(sub + chunk_addr(row, u), make_float4(v[4 * u] + bv[u].x, v[4 * u + 1] + bv[u].y, v[4 * u + 2] + bv[u].z, v[4 * u + 3] + bv[u].w));
        }
      }
      if (ring) continue;   // stored chunk by chunk above
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);   // accumulator drained: the next MMAs may overwrite it
      fence_proxy_async();                         // staged tile visible to the TMA (async proxy)
      epi_barrier();
      if (leader && a.debug != 1 && a.debug != 6 && a.debug != 7 && !(a.epi == 2 && !a.store_lin)) {
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += CW) {
          const void* src = stg + (c0 / CW) * (kBM * CW * 4);
          if (a.plain) tma_store_2d(&mapC, src, ncol0 + c0, mt * kBM, a.accumulate != 0);
          else tma_store_4d(&mapC, src, c0, 0, t0, b, a.accumulate != 0);
        }
        bulk_commit();
      }
      if (S::STG2_BYTES > 0 && a.epi == 2) {
        // average pool from the gated tile: pooled float4 i = (tol, fo, c4); sources are pt x pf packed pixels
        const int c4n = a.gC / 4;
        const int n4 = (a.th / a.pt) * a.Fo * c4n;
        const float inv = 1.0f / (float)(a.pt * a.pf);
        const uint32_t s2 = smem_u32(stg2);
        for (int i = row; i < n4; i += kBM) {
          const int c4 = i % c4n, fo = (i / c4n) % a.Fo, tol = i / (c4n * a.Fo);
          const int to = t0 / a.pt + tol;
          if (to >= a.To) continue;
          float4 acc4 = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int dt = 0; dt < a.pt; ++dt)
            for (int df = 0; df < a.pf; ++df) {
              const int f = fo * a.pf + df;
              const int r = (tol * a.pt + dt) * a.F + f / a.pack;
              const int col = (f % a.pack) * a.gC + c4 * 4;
              const float4 x4 = lds128(s2 + (col / CW) * (kBM * CW * 4) + chunk_addr(r, (col % CW) >> 2));
              acc4.x += x4.x; acc4.y += x4.y; acc4.z += x4.z; acc4.w += x4.w;
            }
          acc4.x *= inv; acc4.y *= inv; acc4.z *= inv; acc4.w *= inv;
          *reinterpret_cast<float4*>(a.pooled + (((size_t)b * a.To + to) * a.Fo + fo) * a.gC + c4 * 4) = acc4;
        }
      }
      if (a.stats) {
        int gi = 0;
#pragma unroll
        for (int k = 1; k < kMaxGroups; ++k)
          if (k < a.stats_groups && b >= a.gfirst[k]) gi = k;
        if (gi != st_grp) {
          flush_stats();
          st_grp = gi;
        }
        if (row < N) {
          const int c = row;
          const uint32_t cbase = stg_addr + (c / CW) * (kBM * CW * 4) + (c & 3) * 4;
          const int c4 = (c % CW) >> 2;
          float s1 = 0.f, s2 = 0.f;
          for (int r = 0; r < valid_rows; ++r) {
            const float x = lds32(cbase + chunk_addr(r, c4));
            s1 += x;
            s2 = fmaf(x, x, s2);
          }
          st_sum += (double)s1;
          st_sq += (double)s2;
        }
      }
    }
    if (a.stats) flush_stats();
    if (leader) bulk_wait0();                      // shared memory must outlive the TMA reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------
constexpr int kSmemMax = 227 * 1024;

template <int N, int KCH, bool RB, bool X3>
static int launch_k2(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mBlo, const CUtensorMap& mC, float* Y,
                     const float* bias, KArgs& a, int sms, cudaStream_t st) {
  // as many stages as fit the shared memory (max 24): the small-tile GEMMs (GLU, block 1) are streaming kernels that
  // need tens of KB of loads in flight per SM to cover the HBM latency
  using S = KSmem<N, KCH, RB, X3>;
  a.rb_bytes = RB ? S::NB * a.ntaps * a.cpt * S::B_STRIDE : 0;
  a.ring_extra = 0;
  if (a.op_ring) {   // the operand ring overlays the staging tile(s); it needs a 4-stage A ring beside it
    const int extra = kOpSlots * kOpSlotBytes - S::STG_BYTES - S::STG2_BYTES;
    a.ring_extra = extra > 0 ? extra : 0;
    const int min_stages = N == 64 ? 4 : 3;
    if ((N != 64 && N != 128) || (kSmemMax - S::FIXED - a.rb_bytes - a.ring_extra) / S::STAGE < min_stages)
      a.op_ring = a.ring_extra = 0;
  }
  int stages = (kSmemMax - S::FIXED - a.rb_bytes - a.ring_extra) / S::STAGE;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) {
    bsed_set_error("tc gemm: no room for a 2-stage ring (N=%d KCH=%d rb=%d x3=%d)", N, KCH, a.rb_bytes, (int)X3);
    return BSED_E_INVALID;
  }
  a.stages = stages;
  const int total = stages * S::STAGE + a.rb_bytes + S::FIXED + a.ring_extra;
  auto kern = tc_kmajor_kernel<N, KCH, RB, X3>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    BSED_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
  int grid = a.n_tiles < sms ? a.n_tiles : sms;
  const CUtensorMap& mX = a.op_ring ? *static_cast<const CUtensorMap*>(a.mX_host) : mC;
  kern<<<grid, X3 ? kThreadsX3 : kThreads, total, st>>>(mA, mB, mBlo, mC, mX, Y, bias, a);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

template <int N, int KCH, bool X3>
static int launch_k(const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mBlo, const CUtensorMap& mC, float* Y,
                    const float* bias, KArgs& a, int sms, cudaStream_t st) {
  const bool rb = a.n_chunks == 1 && (long long)(X3 ? 2 : 1) * a.ntaps * a.cpt * KSmem<N, KCH, true, X3>::B_STRIDE <= kRbBytes;
  if (rb) return launch_k2<N, KCH, true, X3>(mA, mB, mBlo, mC, Y, bias, a, sms, st);
  return launch_k2<N, KCH, false, X3>(mA, mB, mBlo, mC, Y, bias, a, sms, st);
}

template <int KCH, bool X3>
static int dispatch_n(int N, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mBlo, const CUtensorMap& mC,
                      float* Y, const float* bias, KArgs& a, int sms, cudaStream_t st) {
  switch (N) {
    case 16: return launch_k<16, KCH, X3>(mA, mB, mBlo, mC, Y, bias, a, sms, st);
    case 32: return launch_k<32, KCH, X3>(mA, mB, mBlo, mC, Y, bias, a, sms, st);
    case 64: return launch_k<64, KCH, X3>(mA, mB, mBlo, mC, Y, bias, a, sms, st);
    case 128: return launch_k<128, KCH, X3>(mA, mB, mBlo, mC, Y, bias, a, sms, st);
  }
  bsed_set_error("tc gemm: N=%d unsupported (16/32/64/128)", N);
  return BSED_E_INVALID;
}

static int dispatch_k(int KCH, bool x3, int N, const CUtensorMap& mA, const CUtensorMap& mB, const CUtensorMap& mBlo,
                      const CUtensorMap& mC, float* Y, const float* bias, KArgs& a, int sms, cudaStream_t st) {
  if (x3) {
    if (KCH == 32) return dispatch_n<32, true>(N, mA, mB, mBlo, mC, Y, bias, a, sms, st);
    return dispatch_n<16, true>(N, mA, mB, mBlo, mC, Y, bias, a, sms, st);
  }
  if (KCH == 32) return dispatch_n<32, false>(N, mA, mB, mBlo, mC, Y, bias, a, sms, st);
  return dispatch_n<16, false>(N, mA, mB, mBlo, mC, Y, bias, a, sms, st);
}

// k-chunk width.  3xTF32 doubles the weight tiles and adds the low-part activation tiles: where the weights cannot
// stay resident, 16-wide chunks (64-byte rows) keep a >= 4-stage ring inside the 227 KB of shared memory.
static inline int pick_kch(int K, int N, int ntaps, bool x3) {
  if (K % 32 != 0) return 16;
  if (!x3) return 32;
  const long long rb32 = 2LL * ntaps * (K / 32) * ((N * 32 * 4 + 1023) / 1024 * 1024);
  return (rb32 <= kRbBytes || N <= 32) ? 32 : 16;
}

// ---------------------------------------------------------------------------------------------
// MN-major kernel: reductions over rows ("weight gradients").
//     D[m][n] (per tap) = sum_rows A[row][m] * Bm[row + shift(tap)][n]
// Both operands are "[rows][channels]" tensors, i.e. MN-major for this product (the reduction runs over
// the rows): A = dY tile (M = channels of A), B = shifted X tile (N = channels of B).  Same 4-D TMA boxes as
// the forward kernel -- (channel, f, t, clip) with zero fill outside the clip = the convolution's padding
// -- only the descriptors differ.  One CTA = (row range, tap, M/N tile); the accumulator stays in TMEM for
// the whole range; partial results go to a workspace [split][tap][tile][128][N] reduced afterwards.
// Modes:
//   W_CONV9   9 taps (dt, df) in {-1,0,1}^2                      conv weight gradient, Cin % 32 == 0
//   W_SINGLE  one tap (dt0, 0)                                   GLU / GRU reductions (dt0 = -1/+1: W_hh)
//   W_PAIR    operands viewed two pixels per row (rows of 2*C floats) for C = 16 channel tensors, whose
//             64-byte rows cannot be MN-major tf32 operands: 12 virtual taps (dt, j) with
//             j = 0: A even pixels, B pair g      -> cols 0..15 = tap df 0,  cols 16..31 = tap df +1
//             j = 1: A even pixels, B pair g - 1  ->                         cols 16..31 = tap df -1
//             j = 2: A odd pixels,  B pair g      -> cols 0..15 = tap df -1, cols 16..31 = tap df 0
//             j = 3: A odd pixels,  B pair g + 1  -> cols 0..15 = tap df +1
// ---------------------------------------------------------------------------------------------
constexpr int kWP = 64;   // rows per stage
enum { W_CONV9 = 0, W_SINGLE = 1, W_PAIR = 2 };

struct WArgs {
  int n_tiles, tiles_per_clip, th;   // 64-row tiles
  int T, F;
  int mode, dt0;
  int a_c0, b_c0;                    // first channel of A / B
  int m_total;                       // channels of A (multiple of 32); M tiles of 128
  int n_tiles_n;                     // N tiles (each N channels of B)
  int pair_stride;                   // W_PAIR: channel offset of the odd pixel inside an A row (= Cout)
  int tiles_per_split;
};

// MN-major descriptor for 32-bit operands.  tcgen05 accepts exactly one shared-memory layout for MN-major
// tf32: 128-byte rows (32 floats along M/N) whose 32-byte chunks are XOR-swizzled with the row index mod 4
// (layout type 1, SWIZZLE_128B_BASE32B <-> CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).  One swizzle atom is
// 4 k-rows x 128 B; a K = 8 instruction spans two atoms (SBO = 512 B apart); the next 32 floats along M/N
// live LBO bytes further (one TMA box each).
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (1ull << 61);
}

template <int N, int STAGES>
struct WSmem {
  static constexpr int NCH = 32;                           // floats per B row (128 B)
  static constexpr int A_CHUNK = kWP * 128;                // [64 rows][32 channels] fp32
  static constexpr int A_BYTES = 4 * A_CHUNK;              // room for M = 128 (4 chunks)
  static constexpr int B_CHUNK = kWP * NCH * 4;
  static constexpr int B_BYTES = (N / NCH) * B_CHUNK;
  static constexpr int STAGE = A_BYTES + (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int TOTAL = STAGES * STAGE + 1024 + 256;
};

template <int N, int STAGES>
__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                float* __restrict__ part, WArgs a) {
  using S = WSmem<N, STAGES>;
  constexpr int NCH = S::NCH;
  constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : N <= 64 ? 64 : 128;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int split = blockIdx.x, tap = blockIdx.y;
  const int mt = blockIdx.z / a.n_tiles_n, nt = blockIdx.z % a.n_tiles_n;
  const int tile_beg = split * a.tiles_per_split;
  int tile_end = tile_beg + a.tiles_per_split;
  if (tile_end > a.n_tiles) tile_end = a.n_tiles;
  const int my_tiles = tile_end > tile_beg ? tile_end - tile_beg : 0;
  int a_chunks = (a.m_total - mt * kBM) / 32;
  if (a_chunks > 4) a_chunks = 4;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  int dt, df, a_ch = a.a_c0 + mt * kBM;
  const int b_ch = a.b_c0 + nt * N;
  if (a.mode == W_CONV9) {
    dt = tap / 3 - 1;
    df = tap % 3 - 1;
  } else if (a.mode == W_SINGLE) {
    dt = a.dt0;
    df = 0;
  } else {
    const int j = tap & 3;
    dt = (tap >> 2) - 1;
    df = j == 1 ? -1 : j == 3 ? 1 : 0;
    if (j >= 2) a_ch += a.pair_stride;
  }

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = a_chunks * S::A_CHUNK + S::B_BYTES;
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        int b = tile / a.tiles_per_clip;
        int t0 = (tile - b * a.tiles_per_clip) * a.th;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* sa = smem + s * S::STAGE;
        unsigned char* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full[s], tx);
        for (int c = 0; c < a_chunks; ++c) tma_load_4d(&mapA, sa + c * S::A_CHUNK, &full[s], a_ch + c * 32, 0, t0, b);
#pragma unroll
        for (int c = 0; c < N / NCH; ++c) tma_load_4d(&mapB, sb + c * S::B_CHUNK, &full[s], b_ch + c * NCH, df, t0 + dt, b);
        if (++s == STAGES) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_tf32(N, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      __syncwarp();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * S::STAGE);
        const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
        for (int k = 0; k < kWP / 8; ++k) {
          uint64_t da = mnmajor_desc(sa + k * 1024, S::A_CHUNK);
          uint64_t db = mnmajor_desc(sb + k * 1024, S::B_CHUNK);
          umma_tf32(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty[s]);
        if (i == my_tiles - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      if (++s == STAGES) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;   // channel of A inside this M tile
    float* out = part + ((((size_t)split * gridDim.y + tap) * gridDim.z + blockIdx.z) * kBM + row) * N;
    if (my_tiles > 0) {
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + c0, v);
        if (row < a_chunks * 32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
    } else if (row < a_chunks * 32) {
      for (int j = 0; j < N; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------
// All-taps variant for the 3x3 weight gradient (W_CONV9 / W_PAIR).  The per-tap kernel above re-reads dY and X
// from L2 once per tap (9-12x), which is what bounds it; here ONE CTA owns a row range and a 32-channel slice of X
// for ALL taps: per 32-row stage it loads the dY tile once (both pixel parities in pair mode) plus the 9 shifted
// X tiles, and keeps 9 (12) accumulators of 32 columns in TMEM (288 / 384 of the 512 columns).
//   grid = (splits, 1, N tiles of 32 channels);   partial layout identical to tc_wgrad_kernel's with N = 32.
// ---------------------------------------------------------------------------------------------
constexpr int kW9Rows = 32;    // rows per stage
constexpr int kW9Stages = 3;
struct W9Smem {
  static constexpr int CHUNK = kW9Rows * 128;          // [32 rows][32 channels] fp32 = 4 KB
  static constexpr int A_BYTES = 8 * CHUNK;            // up to 4 chunks (M = 128) x 2 pixel parities
  static constexpr int B_BYTES = 9 * CHUNK;            // 9 shifted X tiles
  static constexpr int STAGE = (A_BYTES + B_BYTES + 1023) / 1024 * 1024;
  static constexpr int TOTAL = kW9Stages * STAGE + 1024 + 256;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_wgrad9_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                 float* __restrict__ part, WArgs a) {
  using S = W9Smem;
  constexpr int N = 32;
  constexpr uint32_t TMEM_COLS = 512;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kW9Stages * S::STAGE);
  uint64_t* full = bars;
  uint64_t* empty = bars + kW9Stages;
  uint64_t* tfull = bars + 2 * kW9Stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kW9Stages + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int split = blockIdx.x, nt = blockIdx.z;
  const bool pair = a.mode == W_PAIR;
  const int ntaps = pair ? 12 : 9;
  const int tile_beg = split * a.tiles_per_split;
  int tile_end = tile_beg + a.tiles_per_split;
  if (tile_end > a.n_tiles) tile_end = a.n_tiles;
  const int my_tiles = tile_end > tile_beg ? tile_end - tile_beg : 0;
  int a_chunks = a.m_total / 32;
  if (a_chunks > 4) a_chunks = 4;
  const int a_sets = pair ? 2 : 1;                      // pixel parities

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&mapA);
    prefetch_tmap(&mapB);
    for (int s = 0; s < kW9Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tfull[0], 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int b_ch = a.b_c0 + nt * N;

  if (warp == 0) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t tx = (a_sets * a_chunks + 9) * S::CHUNK;
      for (int tile = tile_beg; tile < tile_end; ++tile) {
        int b = tile / a.tiles_per_clip;
        int t0 = (tile - b * a.tiles_per_clip) * a.th;
        mbar_wait(&empty[s], ph ^ 1);
        unsigned char* sa = smem + s * S::STAGE;
        unsigned char* sb = sa + S::A_BYTES;
        mbar_expect_tx(&full[s], tx);
        for (int ps = 0; ps < a_sets; ++ps)
          for (int c = 0; c < a_chunks; ++c)
            tma_load_4d(&mapA, sa + (ps * 4 + c) * S::CHUNK, &full[s], a.a_c0 + ps * a.pair_stride + c * 32, 0, t0, b);
#pragma unroll
        for (int i = 0; i < 9; ++i)   // shift (dt, d) = (i / 3 - 1, i % 3 - 1) along (t, f) [pair mode: (t, pixel pair)]
          tma_load_4d(&mapB, sb + i * S::CHUNK, &full[s], b_ch, i % 3 - 1, t0 + i / 3 - 1, b);
        if (++s == kW9Stages) {
          s = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = idesc_tf32(N, 1, 1);
    int s = 0;
    uint32_t ph = 0;
    for (int i = 0; i < my_tiles; ++i) {
      mbar_wait(&full[s], ph);
      tc_fence_after();
      __syncwarp();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + s * S::STAGE);
        const uint32_t sb = sa + S::A_BYTES;
        for (int tap = 0; tap < ntaps; ++tap) {
          int aset = 0, bidx = tap;
          if (pair) {   // virtual tap (dt, j): see the table above
            const int j = tap & 3, dt = tap >> 2;
            aset = j >> 1;
            bidx = dt * 3 + (j == 1 ? 0 : j == 3 ? 2 : 1);
          }
          const uint32_t abase = sa + aset * 4 * S::CHUNK, bbase = sb + bidx * S::CHUNK;
#pragma unroll
          for (int k = 0; k < kW9Rows / 8; ++k) {
            uint64_t da = mnmajor_desc(abase + k * 1024, S::CHUNK);
            uint64_t db = mnmajor_desc(bbase + k * 1024, S::CHUNK);
            umma_tf32(tmem_base + tap * N, da, db, idesc, (i | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&empty[s]);
        if (i == my_tiles - 1) umma_commit(&tfull[0]);
      }
      __syncwarp();
      if (++s == kW9Stages) {
        s = 0;
        ph ^= 1;
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;   // channel of A
    if (my_tiles > 0) {
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
    }
    for (int tap = 0; tap < ntaps; ++tap) {
      float* out = part + ((((size_t)split * ntaps + tap) * gridDim.z + blockIdx.z) * kBM + row) * N;
      float v[32];
      if (my_tiles > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + tap * N, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      if (row < a_chunks * 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

struct RArgs {
  int splits, ntaps, ztiles, n_tiles_n, N;
  int m_total, n_total;
  int mode;
  long long rs, cs, ts;
};

// part [split][tap][z][128][N] -> dW.  W_CONV9 / W_SINGLE: dW[m*rs + n*cs + tap*ts] += sum_split.
// W_PAIR (C = 16 input channels, see the table above): dW[m*rs + ci*cs + (3*(dt+1) + df+1)*ts].
// mode 3 (GLU of a 16-channel block, both operands in pair view): dW[m*rs + n*cs] += P[m][n] + P[16+m][16+n].
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ part, RArgs r, float* dW) {
  // 32 output elements per CTA; the splits are spread over 8 thread groups (lane = element, warp = split group)
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sg = threadIdx.x >> 5;
  auto P = [&](int tap, int m, int n) {
    const int z = (m / kBM) * r.n_tiles_n + n / r.N;
    const float* p0 = part + (((size_t)tap * r.ztiles + z) * kBM + m % kBM) * r.N + n % r.N;
    const size_t stride = (size_t)r.ntaps * r.ztiles * kBM * r.N;
    float a[4] = {0.f, 0.f, 0.f, 0.f};   // independent loads in flight
    int s = sg;
    for (; s + 24 < r.splits; s += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) a[u] += p0[(size_t)(s + 8 * u) * stride];
    }
    for (; s < r.splits; s += 8) a[0] += p0[(size_t)s * stride];
    return (a[0] + a[1]) + (a[2] + a[3]);
  };
  float acc = 0.f;
  size_t dst = 0;
  bool live = false;
  if (r.mode == W_PAIR) {
    if (i < 9 * r.m_total * 16) {
      const int ci = i % 16, m = (i / 16) % r.m_total, tap = i / (16 * r.m_total);
      const int dt = tap / 3, df = tap % 3 - 1;
      if (df == 0) acc = P(dt * 4 + 0, m, ci) + P(dt * 4 + 2, m, 16 + ci);
      else if (df == 1) acc = P(dt * 4 + 0, m, 16 + ci) + P(dt * 4 + 3, m, ci);
      else acc = P(dt * 4 + 1, m, 16 + ci) + P(dt * 4 + 2, m, ci);
      dst = m * r.rs + ci * r.cs + tap * r.ts;
      live = true;
    }
  } else if (r.mode == 3) {
    if (i < 16 * 16) {
      const int n = i % 16, m = i / 16;
      acc = P(0, m, n) + P(0, 16 + m, 16 + n);
      dst = m * r.rs + n * r.cs;
      live = true;
    }
  } else if (i < r.ntaps * r.m_total * r.n_total) {
    const int n = i % r.n_total, m = (i / r.n_total) % r.m_total, tap = i / (r.n_total * r.m_total);
    acc = P(tap, m, n);
    dst = m * r.rs + n * r.cs + tap * r.ts;
    live = true;
  }
  __shared__ float red[8][32];
  red[sg][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sg == 0 && live) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    dW[dst] += t;
  }
}

// Vectorised split-K reduction of the plain modes (W_CONV9 / W_SINGLE): a thread owns four consecutive n of one
// (tap, m) row, so every partial is read as one float4; the splits are spread over 8 thread groups as above.
// 4x fewer threads and loads than the scalar kernel (which stays for the pair / 16-channel views).
__global__ void __launch_bounds__(256) wgrad_reduce4_kernel(const float* __restrict__ part, RArgs r, float* dW) {
  const int i4 = blockIdx.x * 32 + (threadIdx.x & 31);
  const int sg = threadIdx.x >> 5;
  const int nq = r.n_total / 4;
  const bool live = i4 < r.ntaps * r.m_total * nq;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int m = 0, n = 0, tap = 0;
  if (live) {
    n = (i4 % nq) * 4;
    m = (i4 / nq) % r.m_total;
    tap = i4 / (nq * r.m_total);
    const int z = (m / kBM) * r.n_tiles_n + n / r.N;
    const float* p0 = part + (((size_t)tap * r.ztiles + z) * kBM + m % kBM) * r.N + n % r.N;
    const size_t stride = (size_t)r.ntaps * r.ztiles * kBM * r.N;
    float4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = sg;
    for (; s + 24 < r.splits; s += 32) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 v = *reinterpret_cast<const float4*>(p0 + (size_t)(s + 8 * u) * stride);
        a[u].x += v.x;
        a[u].y += v.y;
        a[u].z += v.z;
        a[u].w += v.w;
      }
    }
    for (; s < r.splits; s += 8) {
      const float4 v = *reinterpret_cast<const float4*>(p0 + (size_t)s * stride);
      a[0].x += v.x;
      a[0].y += v.y;
      a[0].z += v.z;
      a[0].w += v.w;
    }
    acc = make_float4((a[0].x + a[1].x) + (a[2].x + a[3].x), (a[0].y + a[1].y) + (a[2].y + a[3].y),
                      (a[0].z + a[1].z) + (a[2].z + a[3].z), (a[0].w + a[1].w) + (a[2].w + a[3].w));
  }
  __shared__ float4 red[8][32];
  red[sg][threadIdx.x & 31] = acc;
  __syncthreads();
  if (sg == 0 && live) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 v = red[k][threadIdx.x];
      t.x += v.x;
      t.y += v.y;
      t.z += v.z;
      t.w += v.w;
    }
    float* d = dW + (size_t)m * r.rs + (size_t)n * r.cs + (size_t)tap * r.ts;
    d[0] += t.x;
    d[r.cs] += t.y;
    d[2 * r.cs] += t.z;
    d[3 * r.cs] += t.w;
  }
}

template <int N>
static int launch_w(const CUtensorMap& mA, const CUtensorMap& mB, float* part, const WArgs& a, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = N >= 128 ? 3 : 4;
  using S = WSmem<N, STAGES>;
  static_assert(S::TOTAL <= 227 * 1024, "wgrad stage ring exceeds shared memory");
  auto kern = tc_wgrad_kernel<N, STAGES>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    BSED_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
  kern<<<grid, kThreads, S::TOTAL, st>>>(mA, mB, part, a);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

}  // namespace tc

// Y[B][T][F][Cout] (+)= conv3x3(X[B][T][F][Cin], Wk) + bias ; Wk = K-major packed weights [Cout][9*Cin]
// (k = tap*Cin + ci).  Requires F in {2..128} dividing 128.
// Wk_lo != nullptr selects 3xTF32: Wk holds the tf32-rounded weights, Wk_lo their fp32 remainders (same layout).
int tc_conv3x3_stats(const float* X, const float* Wk, const float* Wk_lo, float* Y, int B, int T, int F, int Cin, int Cout,
                     const float* bias, int accumulate, double* stats, int stats_groups, const int* gfirst, int sms,
                     cudaStream_t st) {
  if (!accumulate && tc_conv_col_supported(F, Cin, Cout))
    return tc_conv3x3_col(X, Wk, Wk_lo, Y, B, T, F, Cin, Cout, bias, stats, stats_groups, gfirst, sms, st);
  BSED_REQUIRE(Cin % 16 == 0 && Cout % 16 == 0 && Cout <= 128, "tc_conv3x3: Cin=%d Cout=%d", Cin, Cout);
  BSED_REQUIRE(F >= 1 && F <= 128 && 128 % F == 0, "tc_conv3x3: F=%d must divide 128", F);
  const bool x3 = Wk_lo != nullptr;
  const int KCH = tc::pick_kch(Cin, Cout, 9, x3);
  const int th = 128 / F;
  const int CW = Cout >= 32 ? 32 : 16;
  CUtensorMap mA, mB, mBlo, mC;
  cuuint64_t dA[4] = {(cuuint64_t)Cin, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t sA[3] = {(cuuint64_t)Cin * 4, (cuuint64_t)F * Cin * 4, (cuuint64_t)T * F * Cin * 4};
  cuuint32_t bA[4] = {(cuuint32_t)KCH, (cuuint32_t)F, (cuuint32_t)th, 1};
  BSED_TRY(tc::make_map(&mA, X, 4, dA, sA, bA, KCH * 4, x3));
  cuuint64_t dB[2] = {(cuuint64_t)9 * Cin, (cuuint64_t)Cout};
  cuuint64_t sB[1] = {(cuuint64_t)9 * Cin * 4};
  cuuint32_t bB[2] = {(cuuint32_t)KCH, (cuuint32_t)Cout};
  BSED_TRY(tc::make_map(&mB, Wk, 2, dB, sB, bB, KCH * 4, x3));
  BSED_TRY(tc::make_map(&mBlo, x3 ? Wk_lo : Wk, 2, dB, sB, bB, KCH * 4, x3));
  cuuint64_t dC[4] = {(cuuint64_t)Cout, (cuuint64_t)F, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t sC[3] = {(cuuint64_t)Cout * 4, (cuuint64_t)F * Cout * 4, (cuuint64_t)T * F * Cout * 4};
  cuuint32_t bC[4] = {(cuuint32_t)CW, (cuuint32_t)F, (cuuint32_t)th, 1};
  BSED_TRY(tc::make_map(&mC, Y, 4, dC, sC, bC, CW * 4, true));
  tc::KArgs a;
  a.op_ring = a.ring_extra = 0;
  a.mX_host = nullptr;
  a.plain = 0;
  a.tiles_per_clip = (T + th - 1) / th;
  a.n_tiles = a.tiles_per_clip * B;
  a.th = th;
  a.T = T;
  a.F = F;
  a.rows = 0;
  a.ntaps = 9;
  a.n_chunks = 1;
  a.cpt = Cin / KCH;
  a.ldc = Cout;
  a.accumulate = accumulate;
  a.debug = tc_debug();
  a.epi = 0;
  a.xh = nullptr;
  a.tab = nullptr;
  a.tab_groups = 0;
  a.rows_per_clip = 1;
  for (int k = 0; k < kMaxGroups; ++k) a.gfirst[k] = stats && k < stats_groups ? gfirst[k] : 0;
  a.stats = stats;
  a.stats_groups = stats ? stats_groups : 0;
  a.pooled = nullptr;
  a.pt = a.pf = a.pack = 1;
  a.gC = a.To = a.Fo = 0;
  a.store_lin = 1;
  a.drop_key = a.drop_thresh = a.drop_base = 0;
  a.inv_keep = 1.f;
  ProfScope prof(PROF_CONV, 2.0 * B * T * F * Cout * 9.0 * Cin,
                 4.0 * ((double)B * T * F * Cin + (double)B * T * F * Cout + 9.0 * Cin * Cout), st);
  return tc::dispatch_k(KCH, x3, Cout, mA, mB, mBlo, mC, Y, bias, a, sms, st);
}

int tc_conv3x3(const float* X, const float* Wk, const float* Wk_lo, float* Y, int B, int T, int F, int Cin, int Cout,
               const float* bias, int accumulate, int sms, cudaStream_t st) {
  return tc_conv3x3_stats(X, Wk, Wk_lo, Y, B, T, F, Cin, Cout, bias, accumulate, nullptr, 0, nullptr, sms, st);
}

// C[M][N] (+)= A[M][K] * Bk^T + bias ; Bk = [N][K] K-major.  bnb != nullptr selects the BatchNorm-backward epilogue.
struct BnBwdEpi {
  const float* xhat;
  const float* tab;
  int groups;
  long long rows_per_clip;
  int gfirst[kMaxGroups];
};
// N > 128: the output is walked in column blocks of 128 inside ONE launch (tile = (row tile, column block), the row
// tile's A chunks re-read from L2), instead of one launch per block
static int tc_gemm_nt_impl(const float* A, int lda, const float* Bk, const float* Bk_lo, int ldb, float* C, int ldc, long long M,
                           int N_total, int K, const float* bias, int accumulate, const BnBwdEpi* bnb, int sms, cudaStream_t st) {
  const int n_chunks = N_total > 128 ? N_total / 128 : 1;
  const int N = N_total > 128 ? 128 : N_total;
  BSED_REQUIRE(K % 16 == 0 && N % 16 == 0 && N * n_chunks == N_total && N_total <= 1024 && lda % 4 == 0 && ldb % 4 == 0 &&
                   ldc % 4 == 0 && (n_chunks == 1 || !bnb),
               "tc_gemm_nt: M=%lld N=%d K=%d", M, N_total, K);
  const bool x3 = Bk_lo != nullptr;
  // BatchNorm-backward epilogue on 64 / 128 columns: operand ring (KArgs::op_ring); in 3xTF32 it fits beside 16-channel
  // chunks only
  const bool op_ring = bnb && (N_total == 64 || N_total == 128) && !accumulate && !getenv("BSED_NO_OP_RING");
  const int KCH = (op_ring && x3) ? 16 : tc::pick_kch(K, N, 1, x3);
  CUtensorMap mA, mB, mBlo;
  cuuint64_t dA[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t sA[1] = {(cuuint64_t)lda * 4};
  cuuint32_t bA[2] = {(cuuint32_t)KCH, 128};
  BSED_TRY(tc::make_map(&mA, A, 2, dA, sA, bA, KCH * 4, x3));
  cuuint64_t dB[2] = {(cuuint64_t)K, (cuuint64_t)N_total};
  cuuint64_t sB[1] = {(cuuint64_t)ldb * 4};
  cuuint32_t bB[2] = {(cuuint32_t)KCH, (cuuint32_t)N};
  BSED_TRY(tc::make_map(&mB, Bk, 2, dB, sB, bB, KCH * 4, x3));
  BSED_TRY(tc::make_map(&mBlo, x3 ? Bk_lo : Bk, 2, dB, sB, bB, KCH * 4, x3));
  const int CW = N >= 32 ? 32 : 16;
  CUtensorMap mC;
  cuuint64_t dC[2] = {(cuuint64_t)N_total, (cuuint64_t)M};
  cuuint64_t sC[1] = {(cuuint64_t)ldc * 4};
  cuuint32_t bC[2] = {(cuuint32_t)CW, 128};
  BSED_TRY(tc::make_map(&mC, C, 2, dC, sC, bC, CW * 4, true));
  CUtensorMap mX = mC;
  if (op_ring) BSED_TRY(tc::make_map(&mX, bnb->xhat, 2, dC, sC, bC, CW * 4, true));
  tc::KArgs a;
  a.op_ring = op_ring ? 1 : 0;
  a.mX_host = &mX;
  a.plain = 1;
  a.n_chunks = n_chunks;
  a.n_tiles = (int)((M + 127) / 128) * n_chunks;
  a.tiles_per_clip = 1;
  a.th = 1;
  a.T = 1;
  a.F = 128;
  a.rows = M;
  a.ntaps = 1;
  a.cpt = K / KCH;
  a.ldc = ldc;
  a.accumulate = accumulate;
  a.debug = tc_debug();
  a.epi = bnb ? 1 : 0;
  a.xh = bnb ? bnb->xhat : nullptr;
  a.tab = bnb ? bnb->tab : nullptr;
  a.tab_groups = bnb ? bnb->groups : 0;
  a.rows_per_clip = bnb ? bnb->rows_per_clip : 1;
  for (int k = 0; k < kMaxGroups; ++k) a.gfirst[k] = bnb ? bnb->gfirst[k] : 0;
  a.stats = nullptr;
  a.stats_groups = 0;
  a.pooled = nullptr;
  a.pt = a.pf = a.pack = 1;
  a.gC = a.To = a.Fo = 0;
  a.store_lin = 1;
  a.drop_key = a.drop_thresh = a.drop_base = 0;
  a.inv_keep = 1.f;
  ProfScope prof(PROF_GEMM, 2.0 * M * N_total * K, 4.0 * ((double)M * K + (double)K * N_total + (double)M * N_total), st);
  return tc::dispatch_k(KCH, x3, N, mA, mB, mBlo, mC, C, bias, a, sms, st);
}

int tc_gemm_nt(const float* A, int lda, const float* Bk, const float* Bk_lo, int ldb, float* C, int ldc, long long M, int N,
               int K, const float* bias, int accumulate, int sms, cudaStream_t st) {
  return tc_gemm_nt_impl(A, lda, Bk, Bk_lo, ldb, C, ldc, M, N, K, bias, accumulate, nullptr, sms, st);
}

// GLU forward of one block with gate, dropout and average pool fused (KArgs::epi == 2).  xhat / lin are the
// [B][T][F][C] tensors viewed as rows of CP = pack * C floats ([B][T][F/pack][CP]); Wk = block-diagonal folded weights
// [CP][CP] (K-major), bias [CP], tab = [gamma x pack | beta x pack]; pooled [B][T/pt][F/pf][C].
int tc_glu_gate_fwd(const float* xhat, const float* Wk, const float* bias, const float* tab, float* lin, float* pooled,
                    int B, int T, int F, int C, int pack, int pt, int pf, uint32_t key, uint32_t thresh, float inv_keep,
                    uint32_t drop_base, int store_lin, int sms, cudaStream_t st) {
  const int CP = C * pack, Fp = F / pack;
  BSED_REQUIRE(CP == 64 && F % pack == 0 && Fp <= 128 && 128 % Fp == 0, "tc_glu_gate_fwd: C=%d pack=%d F=%d", C, pack, F);
  const int th = 128 / Fp;
  BSED_REQUIRE(th % pt == 0 && F % pf == 0 && pf % 1 == 0, "tc_glu_gate_fwd: pooling %dx%d does not tile (th=%d)", pt, pf, th);
  CUtensorMap mA, mB, mC;
  cuuint64_t dA[4] = {(cuuint64_t)CP, (cuuint64_t)Fp, (cuuint64_t)T, (cuuint64_t)B};
  cuuint64_t sA[3] = {(cuuint64_t)CP * 4, (cuuint64_t)Fp * CP * 4, (cuuint64_t)T * Fp * CP * 4};
  cuuint32_t bA[4] = {32, (cuuint32_t)Fp, (cuuint32_t)th, 1};
  BSED_TRY(tc::make_map(&mA, xhat, 4, dA, sA, bA, 128));
  cuuint64_t dB[2] = {(cuuint64_t)CP, (cuuint64_t)CP};
  cuuint64_t sB[1] = {(cuuint64_t)CP * 4};
  cuuint32_t bB[2] = {32, (cuuint32_t)CP};
  BSED_TRY(tc::make_map(&mB, Wk, 2, dB, sB, bB, 128));
  BSED_TRY(tc::make_map(&mC, lin, 4, dA, sA, bA, 128, true));
  tc::KArgs a;
  a.op_ring = a.ring_extra = 0;
  a.mX_host = nullptr;
  a.plain = 0;
  a.tiles_per_clip = (T + th - 1) / th;
  a.n_tiles = a.tiles_per_clip * B;
  a.th = th;
  a.T = T;
  a.F = Fp;
  a.rows = 0;
  a.ntaps = 1;
  a.n_chunks = 1;
  a.cpt = CP / 32;
  a.ldc = CP;
  a.accumulate = 0;
  a.debug = tc_debug();
  a.epi = 2;
  a.xh = xhat;
  a.tab = tab;
  a.tab_groups = 0;
  a.rows_per_clip = 1;
  for (int k = 0; k < kMaxGroups; ++k) a.gfirst[k] = 0;
  a.stats = nullptr;
  a.stats_groups = 0;
  a.pooled = pooled;
  a.pt = pt;
  a.pf = pf;
  a.pack = pack;
  a.gC = C;
  a.To = T / pt;
  a.Fo = F / pf;
  a.store_lin = store_lin;
  a.drop_key = key;
  a.drop_thresh = thresh;
  a.drop_base = drop_base;
  a.inv_keep = inv_keep;
  const double M = (double)B * T * Fp;
  ProfScope prof(PROF_GEMM, 2.0 * M * CP * CP, 4.0 * (2.0 * M * CP + (double)CP * CP), st);
  return tc::dispatch_k(32, false, CP, mA, mB, mB, mC, lin, bias, a, sms, st);
}

// dY = k * (dxd + A * Bk^T - m1 - xhat * m2), in place on C (= dxd on entry); see KArgs::epi
int tc_gemm_nt_bnbwd(const float* A, const float* Bk, const float* Bk_lo, float* C, const float* xhat, long long M, int N,
                     int K, const float* tab, int groups, long long rows_per_clip, const int* gfirst, int sms, cudaStream_t st) {
  BnBwdEpi e;
  e.xhat = xhat;
  e.tab = tab;
  e.groups = groups;
  e.rows_per_clip = rows_per_clip;
  for (int k = 0; k < kMaxGroups; ++k) e.gfirst[k] = k < groups ? gfirst[k] : 0;
  return tc_gemm_nt_impl(A, K, Bk, Bk_lo, K, C, N, M, N, K, nullptr, 0, &e, sms, st);
}

}  // namespace bsed

namespace bsed {
// split-K partials: per-tap kernel <= (sms + 16) tiles of 128 x 128; all-taps kernel <= (sms + 16) x 12 taps x 128 x 32
size_t tc_wgrad_workspace_bytes(int sms) { return (size_t)(sms + 16) * 12 * 128 * 32 * sizeof(float); }

static int tc_wgrad9(TcOperand A, TcOperand Bm, int Bn, int T, int Fv, int kmode, float* dW, long long rs, long long cs,
                     long long ts, float* part, size_t part_bytes, int sms, cudaStream_t st) {
  const int ntaps = kmode == tc::W_PAIR ? 12 : 9;
  const int th = tc::kW9Rows / Fv;
  const int z = Bm.C / 32;
  CUtensorMap mA, mB;
  const int a_width = kmode == tc::W_PAIR ? 2 * A.C : A.c0 + A.C;
  cuuint64_t dA[4] = {(cuuint64_t)a_width, (cuuint64_t)Fv, (cuuint64_t)T, (cuuint64_t)Bn};
  cuuint64_t sA[3] = {(cuuint64_t)A.ld * 4, (cuuint64_t)Fv * A.ld * 4, (cuuint64_t)T * Fv * A.ld * 4};
  cuuint32_t bA[4] = {32, (cuuint32_t)Fv, (cuuint32_t)th, 1};
  BSED_TRY(tc::make_map(&mA, A.p, 4, dA, sA, bA, 132));
  cuuint64_t dB[4] = {(cuuint64_t)(Bm.c0 + Bm.C), (cuuint64_t)Fv, (cuuint64_t)T, (cuuint64_t)Bn};
  cuuint64_t sB[3] = {(cuuint64_t)Bm.ld * 4, (cuuint64_t)Fv * Bm.ld * 4, (cuuint64_t)T * Fv * Bm.ld * 4};
  cuuint32_t bB[4] = {32, (cuuint32_t)Fv, (cuuint32_t)th, 1};
  BSED_TRY(tc::make_map(&mB, Bm.p, 4, dB, sB, bB, 132));
  tc::WArgs a;
  a.tiles_per_clip = (T + th - 1) / th;
  a.n_tiles = a.tiles_per_clip * Bn;
  a.th = th;
  a.T = T;
  a.F = Fv;
  a.mode = kmode;
  a.dt0 = 0;
  a.a_c0 = A.c0;
  a.b_c0 = Bm.c0;
  a.m_total = A.C;
  a.n_tiles_n = z;
  a.pair_stride = A.C;
  int splits = sms / z;
  if (splits < 1) splits = 1;
  if (splits > a.n_tiles) splits = a.n_tiles;
  a.tiles_per_split = (a.n_tiles + splits - 1) / splits;
  splits = (a.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  size_t need = (size_t)splits * ntaps * z * 128 * 32 * sizeof(float);
  if (part_bytes < need) {
    bsed_set_error("tc_wgrad9: workspace %zu < %zu", part_bytes, need);
    return BSED_E_WORKSPACE;
  }
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured))
    BSED_CHECK_CUDA(cudaFuncSetAttribute(tc::tc_wgrad9_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::W9Smem::TOTAL));
  const double rows = (double)Bn * T * Fv;
  const double taps_real = 9.0;
  const double cin = kmode == tc::W_PAIR ? 16.0 : Bm.C;
  ProfScope prof(PROF_WGRAD, 2.0 * rows * (kmode == tc::W_PAIR ? 2.0 : 1.0) * A.C * taps_real * cin,
                 4.0 * (rows * A.C + rows * Bm.C + 9.0 * A.C * cin), st);
  dim3 grid(splits, 1, z);
  tc::tc_wgrad9_kernel<<<grid, tc::kThreads, tc::W9Smem::TOTAL, st>>>(mA, mB, part, a);
  BSED_CHECK_LAUNCH();
  tc::RArgs ra;
  ra.splits = splits;
  ra.ntaps = ntaps;
  ra.ztiles = z;
  ra.n_tiles_n = z;
  ra.N = 32;
  ra.m_total = A.C;
  ra.n_total = Bm.C;
  ra.mode = kmode;
  ra.rs = rs;
  ra.cs = cs;
  ra.ts = ts;
  const int n = kmode == tc::W_PAIR ? 9 * A.C * 16 : 9 * A.C * Bm.C;
  if (kmode == tc::W_CONV9)
    tc::wgrad_reduce4_kernel<<<ceil_div(n / 4, 32), 256, 0, st>>>(part, ra, dW);
  else
    tc::wgrad_reduce_kernel<<<ceil_div(n, 32), 256, 0, st>>>(part, ra, dW);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// D[m][n][tap] (+)= sum_rows A[row][a.c0 + m] * Bm[row + shift(tap)][b.c0 + n]; element (m, n, tap) lands at
// dW[m*rs + n*cs + tap*ts].  Rows are the pixels of a (Bn, T, Fv) grid, row stride ld floats.
// mode: 0 = 9 conv taps, 1 = one tap (dt0), 2 = pixel-pair view of a 16-channel B operand (A.C = Cout,
// operands passed in their pair views: A.ld = 2*Cout, Bm.ld = 32, Fv = F/2), 3 = 16-channel GLU reduction
// (both operands in pair view, A.C = Bm.C = 32; dW is the 16 x 16 matrix).
int tc_wgrad_ex(TcOperand A, TcOperand Bm, int Bn, int T, int Fv, int mode, int dt0, float* dW, long long rs,
                long long cs, long long ts, float* part, size_t part_bytes, int sms, cudaStream_t st) {
  BSED_REQUIRE(A.C % 32 == 0 && A.C >= 32 && Bm.C % 32 == 0 && Bm.C >= 32, "tc_wgrad: channels A=%d B=%d", A.C, Bm.C);
  BSED_REQUIRE(Fv >= 1 && Fv <= 64 && 64 % Fv == 0, "tc_wgrad: F=%d must divide 64", Fv);
  BSED_REQUIRE(A.ld % 4 == 0 && Bm.ld % 4 == 0, "tc_wgrad: leading dimensions must be multiples of 4");
  const int kmode = mode == 3 ? tc::W_SINGLE : mode;
  // all-taps kernel only where X is narrow (one 32-channel tile): every tcgen05.mma re-reads its 128 x 8 A tile from
  // shared memory whatever N is, so splitting wide X into 32-column MMAs costs more than the L2 traffic it saves
  if ((kmode == tc::W_CONV9 || kmode == tc::W_PAIR) && A.C <= 128 && Bm.C == 32 && 32 % Fv == 0 &&
      !getenv("BSED_WGRAD_PER_TAP"))
    return tc_wgrad9(A, Bm, Bn, T, Fv, kmode, dW, rs, cs, ts, part, part_bytes, sms, st);
  const int N = Bm.C >= 128 ? 128 : Bm.C;
  BSED_REQUIRE(Bm.C % N == 0 && (N == 32 || N == 64 || N == 128), "tc_wgrad: B channels %d", Bm.C);
  const int n_tiles_n = Bm.C / N, m_tiles = (A.C + 127) / 128;
  const int ntaps = kmode == tc::W_CONV9 ? 9 : kmode == tc::W_PAIR ? 12 : 1;
  const int th = tc::kWP / Fv;
  CUtensorMap mA, mB;
  const int a_width = mode == 2 ? 2 * A.C : A.c0 + A.C;   // channels the A map must cover
  cuuint64_t dA[4] = {(cuuint64_t)a_width, (cuuint64_t)Fv, (cuuint64_t)T, (cuuint64_t)Bn};
  cuuint64_t sA[3] = {(cuuint64_t)A.ld * 4, (cuuint64_t)Fv * A.ld * 4, (cuuint64_t)T * Fv * A.ld * 4};
  cuuint32_t bA[4] = {32, (cuuint32_t)Fv, (cuuint32_t)th, 1};
  BSED_TRY(tc::make_map(&mA, A.p, 4, dA, sA, bA, 132));
  cuuint64_t dB[4] = {(cuuint64_t)(Bm.c0 + Bm.C), (cuuint64_t)Fv, (cuuint64_t)T, (cuuint64_t)Bn};
  cuuint64_t sB[3] = {(cuuint64_t)Bm.ld * 4, (cuuint64_t)Fv * Bm.ld * 4, (cuuint64_t)T * Fv * Bm.ld * 4};
  cuuint32_t bB[4] = {32, (cuuint32_t)Fv, (cuuint32_t)th, 1};
  BSED_TRY(tc::make_map(&mB, Bm.p, 4, dB, sB, bB, 132));
  tc::WArgs a;
  a.tiles_per_clip = (T + th - 1) / th;
  a.n_tiles = a.tiles_per_clip * Bn;
  a.th = th;
  a.T = T;
  a.F = Fv;
  a.mode = kmode;
  a.dt0 = dt0;
  a.a_c0 = A.c0;
  a.b_c0 = Bm.c0;
  a.m_total = A.C;
  a.n_tiles_n = n_tiles_n;
  a.pair_stride = A.C;
  const int z = m_tiles * n_tiles_n;
  int splits = sms / (ntaps * z);
  if (splits < 1) splits = 1;
  if (splits > a.n_tiles) splits = a.n_tiles;
  a.tiles_per_split = (a.n_tiles + splits - 1) / splits;
  splits = (a.n_tiles + a.tiles_per_split - 1) / a.tiles_per_split;
  size_t need = (size_t)splits * ntaps * z * 128 * N * sizeof(float);
  if (part_bytes < need) {
    bsed_set_error("tc_wgrad: workspace %zu < %zu", part_bytes, need);
    return BSED_E_WORKSPACE;
  }
  const double rows = (double)Bn * T * Fv;
  ProfScope prof(PROF_WGRAD, 2.0 * rows * A.C * (double)ntaps * Bm.C,
                 4.0 * (rows * A.C + rows * Bm.C + (double)ntaps * A.C * Bm.C), st);
  dim3 grid(splits, ntaps, z);
  int r;
  switch (N) {
    case 32: r = tc::launch_w<32>(mA, mB, part, a, grid, st); break;
    case 64: r = tc::launch_w<64>(mA, mB, part, a, grid, st); break;
    default: r = tc::launch_w<128>(mA, mB, part, a, grid, st); break;
  }
  BSED_TRY(r);
  tc::RArgs ra;
  ra.splits = splits;
  ra.ntaps = ntaps;
  ra.ztiles = z;
  ra.n_tiles_n = n_tiles_n;
  ra.N = N;
  ra.m_total = A.C;
  ra.n_total = Bm.C;
  ra.mode = mode;
  ra.rs = rs;
  ra.cs = cs;
  ra.ts = ts;
  const int n = mode == 2 ? 9 * A.C * 16 : mode == 3 ? 256 : ntaps * A.C * Bm.C;
  if (mode == 0 || mode == 1)
    tc::wgrad_reduce4_kernel<<<ceil_div(n / 4, 32), 256, 0, st>>>(part, ra, dW);
  else
    tc::wgrad_reduce_kernel<<<ceil_div(n, 32), 256, 0, st>>>(part, ra, dW);
  BSED_CHECK_LAUNCH();
  return BSED_OK;
}

// conv weight gradient: dW (+)= sum_p dY[p][co] * X[p + tap][ci] for the 9 taps (ntaps = 9) or the plain
// product (ntaps = 1).  X [B][T][F][Cin], dY [B][T][F][Cout].  Cin = 16 goes through the pixel-pair view.
int tc_wgrad(const float* X, const float* dY, float* dW, long long rs, long long cs, long long ts, int B, int T, int F,
             int Cin, int Cout, int ntaps, float* part, size_t part_bytes, int sms, cudaStream_t st) {
  BSED_REQUIRE(ntaps == 9 || ntaps == 1, "tc_wgrad: ntaps=%d", ntaps);
  BSED_REQUIRE(Cout % 32 == 0 && Cout <= 128 && Cin <= 128, "tc_wgrad: Cin=%d Cout=%d", Cin, Cout);
  if (Cin == 16 && ntaps == 9) {
    BSED_REQUIRE(F % 2 == 0, "tc_wgrad: pair view needs an even F (got %d)", F);
    TcOperand A{dY, 2 * Cout, 0, Cout}, Bm{X, 32, 0, 32};
    return tc_wgrad_ex(A, Bm, B, T, F / 2, 2, 0, dW, rs, cs, ts, part, part_bytes, sms, st);
  }
  BSED_REQUIRE(Cin % 32 == 0, "tc_wgrad: Cin=%d", Cin);
  TcOperand A{dY, Cout, 0, Cout}, Bm{X, Cin, 0, Cin};
  return tc_wgrad_ex(A, Bm, B, T, F, ntaps == 9 ? 0 : 1, 0, dW, rs, cs, ts, part, part_bytes, sms, st);
}
}  // namespace bsed
