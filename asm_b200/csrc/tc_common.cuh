// tc_common.cuh -- what the tensor-core distance kernels (rf_i8.cu: int8, rf_f4.cu: block-scaled fp4) share: tile
// constants, mbarrier / TMA / tcgen05 wrappers, the cluster helpers, the in-kernel work queue of the 2-CTA kernels,
// and the host helpers for the tensor map and the Morton walk.
#pragma once
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "asm_internal.h"

namespace {


constexpr int kStages = 4;
constexpr int kBlockK = 128;                       // bytes (= int8 elements) per K block, one swizzle atom wide
constexpr int kTileN = 256;                        // columns per tile
constexpr int kABytes = kTileRows * kBlockK;       // 16 KB
constexpr int kBBytes = kTileN * kBlockK;          // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;     // 48 KB
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kAccCols = 256;                      // TMEM columns per accumulator
constexpr int kTmemCols = 512;
constexpr size_t kAuxBytes = 128 /*barriers + tmem ptr*/ + 2 * kTileN * 4 /*nb+1, double buffered*/ + 4 * kTileN * 4 /*col partials*/;
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + kAuxBytes;

// instruction descriptor for kind::i8 (cute::UMMA::InstrDescriptor): c_format S32 = 2 @ [4,6),
// a/b_format S8 = 1 @ [7,10)/[10,13), K-major A and B (bits 15,16 = 0), N>>3 @ [17,23), M>>4 @ [24,29)
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileRows >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// shared-memory matrix descriptor, K-major, 128B swizzle: start>>4 @ [0,14), LBO>>4 @ [16,30) (unused
// for swizzled K-major), SBO>>4 = 1024>>4 @ [32,46) (8 rows x 128 B per swizzle atom), version 1 @ [46,48),
// layout SWIZZLE_128B = 2 @ [61,64)  (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }


constexpr int kStages2 = 6;
constexpr int kStageBytes2 = 2 * kABytes;  // A half + B half: 32 KB
constexpr uint32_t kIdesc2 = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in CTA 0 of the pair

__device__ __forceinline__ bool elect_one() {  // one lane of the (converged) warp
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* leader_bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(leader_bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_i8_2sm(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {  // arrives on `bar` in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
struct PairInfo {
  int ib, jb;
  bool do_row, do_col;  // for THIS CTA's 128 rows
  int chain_rows, chain_cols;
};

// Work flags of block pair (ib <= jb) for the CTA of rank `rank`; the return value (any work at all) is the
// same in both CTAs of the pair.
__device__ __forceinline__ bool pair_flags(int ib, int jb, int rank, const int2* __restrict__ meta, PairInfo& pi) {
  pi.ib = ib;
  pi.jb = jb;
  const int2 mi0 = meta[2 * ib], mi1 = meta[2 * ib + 1], mj0 = meta[2 * jb], mj1 = meta[2 * jb + 1];
  const bool diag = ib == jb;
  const bool col_j = (mj0.y & kFlagCol) != 0, col_i = (mi0.y & kFlagCol) != 0;
  const bool rows_j = ((mj0.y | mj1.y) & kFlagRows) != 0;
  const int my = rank ? mi1.y : mi0.y;
  pi.do_row = (my & kFlagRows) && col_j;
  pi.do_col = !diag && col_i && rows_j;
  pi.chain_rows = mi0.x;
  pi.chain_cols = mj0.x;
  return (((mi0.y | mi1.y) & kFlagRows) && col_j) || pi.do_col;
}

// even bits of v, compacted (Morton decode)
__device__ __forceinline__ uint32_t compact_bits(unsigned long long v) {
  v &= 0x5555555555555555ULL;
  v = (v | (v >> 1)) & 0x3333333333333333ULL;
  v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
  v = (v | (v >> 4)) & 0x00FF00FF00FF00FFULL;
  v = (v | (v >> 8)) & 0x0000FFFF0000FFFFULL;
  v = (v | (v >> 16)) & 0x00000000FFFFFFFFULL;
  return (uint32_t)v;
}

__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_rank(uint64_t* bar, uint32_t rank) {  // arrive on CTA `rank`'s copy of `bar`
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank)
      : "memory");
}
// Arrivals that only hand a buffer back (the data was consumed into registers, `dep` is one of those registers so
// that the arrive cannot issue before the load has returned): no release fence.  A releasing arrive is a MEMBAR
// that waits for every outstanding memory operation of the thread, including the fire-and-forget atomics of the
// previous tile (ncu: a fifth of all stall samples of the fp4 kernel).
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar, int dep) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)), "r"(dep) : "memory");
}
__device__ __forceinline__ void mbar_arrive_rank_relaxed(uint64_t* bar, uint32_t rank, int dep) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
      "r"(rank), "r"(dep)
      : "memory");
}
__device__ __forceinline__ void st_int2_rank(int2* p, uint32_t rank, int a, int b) {  // store into CTA `rank`'s copy of *p
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.v2.s32 [ra], {%2, %3};\n\t}" ::"r"(smem_u32(p)),
      "r"(rank), "r"(a), "r"(b)
      : "memory");
}

constexpr int kQ = 4;  // depth of the in-kernel work queue
constexpr int kSchedConsumers = 2 + kEpiWarps + 1 + kEpiWarps;  // leader: TMA, MMA, 8 epilogue warps; peer: TMA, 8 epilogue warps
constexpr size_t kAuxBytes2 = 256 /*barriers, tmem ptr, work queue*/ + 2 * kTileN * 4 + 4 * kTileN * 4;
constexpr size_t kSmemBytes2x = (size_t)kStages2 * kStageBytes2 + kAuxBytes2;

// A role's view of the work queue: wait for entry q, read it, hand the entry back to the scheduler.
struct WorkQueue {
  uint64_t* sfull;
  uint64_t* sempty;
  const int2* slots;
  int q = 0;
  uint32_t ph = 0;
  int rank;
  // whole-warp call (every lane returns the entry; lane 0 releases it) or single-thread call (`solo`)
  __device__ __forceinline__ int2 next(bool solo) {
    mbar_wait_cluster(&sfull[q], ph);
    const int2 e = slots[q];
    if (!solo) __syncwarp();
    if (solo || (threadIdx.x & 31) == 0) {
      if (rank == 0)
        mbar_arrive_relaxed(&sempty[q], e.x);
      else
        mbar_arrive_rank_relaxed(&sempty[q], 0, e.x);
    }
    if (++q == kQ) {
      q = 0;
      ph ^= 1;
    }
    return e;
  }
};

// ---- host side ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// [rows][bytes_per_row] uint8, boxes of 128 rows x 128 bytes, 128B swizzle
int32_t make_tmap(asm_ctx* ctx, CUtensorMap* m, const void* base, uint64_t rows, uint64_t bytes_per_row) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return asm_fail(ctx, ASM_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  cuuint64_t dims[2] = {bytes_per_row, rows};
  cuuint64_t strides[1] = {bytes_per_row};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kTileRows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return asm_fail(ctx, ASM_E_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return ASM_OK;
}

}  // namespace

// The walk: 8x8 tiles of block pairs in Morton order, only the tiles that touch the triangle ib <= jb < nb
// (built on the host when nb changes; 4 bytes per 64 pairs).  Inside a tile the 64 pairs are again in Morton
// order, so any ~74 consecutive slots form a compact 2-D patch and almost every slot is a real pair.
static int32_t ensure_walk(asm_ctx* ctx, long long nb) {
  if (ctx->walk_nb == nb) return ASM_OK;
  const long long nt = (nb + 7) / 8;
  long long Pt = 1;
  while (Pt < nt) Pt <<= 1;
  auto compact = [](unsigned long long v) {
    v &= 0x5555555555555555ULL;
    v = (v | (v >> 1)) & 0x3333333333333333ULL;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFULL;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFULL;
    v = (v | (v >> 16)) & 0x00000000FFFFFFFFULL;
    return (uint32_t)v;
  };
  std::vector<uint32_t>& tiles = ctx->walk_host;
  tiles.clear();
  for (unsigned long long m = 0; m < (unsigned long long)(Pt * Pt); m++) {
    const uint32_t ti = compact(m), tj = compact(m >> 1);
    if (ti <= tj && (long long)tj < nt) tiles.push_back(ti | (tj << 16));
  }
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->walk_dev, sizeof(uint32_t) * tiles.size())) != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemcpyAsync(ctx->walk_dev.p, tiles.data(), sizeof(uint32_t) * tiles.size(), cudaMemcpyHostToDevice,
                                ctx->stream));
  ctx->walk_nb = nb;
  return ASM_OK;
}
