// tc_common.cuh -- PTX wrappers (mbarrier, TMA, tcgen05 / TMEM), UMMA descriptors and tensor-map construction shared
// by the tensor-core kernels of libbsed (tc_gemm.cu, tc_conv.cu).  sm_100a only.
#pragma once
#include <cuda.h>
#include <cstdlib>

#include "launch.h"

namespace bsed {
namespace tc {

constexpr int kBM = 128;
constexpr int kThreads = 192;     // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kThreadsX3 = 256;   // + warps 6-7: operand splitters of the 3xTF32 mode (8 warps keep the 255-register
                                  // budget of the epilogue warps; with 10 warps ptxas caps every thread at 168)
constexpr int kSplitThreads = kThreadsX3 - kThreads;
constexpr uint32_t kSpinLimit = 1u << 26;   // bounded mbarrier waits: trap instead of hanging the GPU

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > kSpinLimit) __trap();
  }
}
// non-blocking probe of a phase (a producer that serves two rings polls both instead of blocking on one)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {   // explicit shared-space load (a generic LD costs a long scoreboard)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, void* dst, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, void* dst, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, bool add) {
  if (add)
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// 3xTF32 operand split.  kind::tf32 reads the upper 19 bits of a 32-bit container (measured on B200: rewriting the tile
// as its truncated high part changes nothing), so a raw fp32 tile acts as its own high part hi = a & 0xffffe000; the low
// part is the exact fp32 difference a - hi (13 significant bits) rounded to NEAREST tf32, so that the tensor core's
// truncation of it is exact and the remaining error (<= 2^-22 |a|) is unbiased:
//   a * b ~= hi * b_hi + hi * b_lo + lo * b_hi.
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float4 tf32_lo4(float4 v) {
  return make_float4(tf32_rna(v.x - tf32_trunc(v.x)), tf32_rna(v.y - tf32_trunc(v.y)), tf32_rna(v.z - tf32_trunc(v.z)),
                     tf32_rna(v.w - tf32_trunc(v.w)));
}
__device__ __forceinline__ void split_barrier() { asm volatile("bar.sync 2, 64;" ::: "memory"); }
static_assert(kThreadsX3 - kThreads == 64, "split_barrier counts the splitter threads");
// low parts of n16 float4 starting at shared address src -> dst, one of kSplitThreads threads (tid).  The loads of a batch are
// issued together: the explicit ld.shared / st.shared wrappers are volatile asm and stay in program order, so a plain
// "load, convert, store" loop pays the whole shared-memory latency per float4 (measured: ~150 clk per float4 and thread,
// which made the splitters the bottleneck of every streaming 3xTF32 GEMM).
template <int BATCH>
__device__ __forceinline__ void split_lo_range(uint32_t src, uint32_t dst, int n16, int tid) {
  int i = tid;
  for (; i + (BATCH - 1) * kSplitThreads < n16; i += BATCH * kSplitThreads) {
    float4 v[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) v[b] = lds128(src + (i + b * kSplitThreads) * 16);
#pragma unroll
    for (int b = 0; b < BATCH; ++b) sts128(dst + (i + b * kSplitThreads) * 16, tf32_lo4(v[b]));
  }
  for (; i < n16; i += kSplitThreads) sts128(dst + i * 16, tf32_lo4(lds128(src + i * 16)));
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// One lane of a converged warp (elect.sync): unlike `lane == 0`, the compiler keeps the code under this predicate on the
// uniform datapath -- descriptors stay in uniform registers and every tcgen05.mma is issued without a per-instruction
// vector-to-uniform "waterfall" loop (ELECT / R2UR.BROADCAST / BRA.U.ANY, ~15 extra instructions per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same MMA with the descriptors given as (low word, shared high word): low word = (address >> 4) | LBO, so a tile
// `off` bytes further on is `lo + (off >> 4)` -- one uniform add per MMA in an issue loop instead of shift / mask / or.
__device__ __forceinline__ void umma_tf32_w(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 p, %5, 0;\n"
      "mov.b64 da, {%1, %3};\n"
      "mov.b64 db, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major, swizzled rows of ROWB bytes (128 or 64):
//   start address >> 4 | LBO (unused for swizzled K-major, 1) | SBO = 8 rows | version 1 | layout type
template <int ROWB>
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t smem_addr) {
  constexpr uint64_t sbo = (8 * ROWB) >> 4;
  constexpr uint64_t layout = ROWB == 128 ? 2 : 4;   // SWIZZLE_128B : SWIZZLE_64B
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// the two words of that descriptor: the low one carries the start address (shared memory is < 256 KB: no carry out of the
// 14-bit field when a byte offset >> 4 is added), the high one is the same for every tile of the layout
__device__ __forceinline__ uint32_t kmajor_desc_lo(uint32_t smem_addr) { return ((smem_addr >> 4) & 0x3FFF) | (1u << 16); }
template <int ROWB>
__host__ __device__ constexpr uint32_t kmajor_desc_hi() {
  return (uint32_t)((8 * ROWB) >> 4) | (1u << 14) | ((ROWB == 128 ? 2u : 4u) << 29);
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// rank-R fp32 map; dims/strides innermost first (strides in bytes for dims 1..R-1)
// `store` (and the operand maps of the 3xTF32 mode) use the plain FLOAT32 element type: the bytes land in shared memory
// untouched; the TFLOAT32 type of the 1xTF32 operand maps lets the copy engine prepare tf32 values.
static inline int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box, int row_bytes, bool store = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    bsed_set_error("cuTensorMapEncodeTiled entry point not available");
    return BSED_E_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = row_bytes == 128   ? CU_TENSOR_MAP_SWIZZLE_128B
                          : row_bytes == 132 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B   // 128 B rows, 32 B swizzle atoms
                                             : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r = fn(m, store ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    bsed_set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return BSED_E_CUDA;
  }
  return BSED_OK;
}

}  // namespace tc

static inline int tc_debug() {   // BSED_TC_DEBUG: measurement experiments only
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BSED_TC_DEBUG");
    v = e ? atoi(e) : 0;
  }
  return v;
}
}  // namespace bsed
