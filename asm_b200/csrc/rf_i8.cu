// rf_i8.cu -- K2: Robinson-Foulds distances as a dense int8 contraction on the 5th-generation tensor
// cores.  On 0/1 clade indicator rows RF(a,b) = |a| + |b| - 2 a.b, so the all-pairs matrix is
// A * A^T with K = D (the clade dictionary width), followed by an integer epilogue that turns the
// dot products into (d+1)^2 and folds them into the within-/between-chain sums of calcPSRF
// (TreePSRF.java:274-286).
//
// One persistent CTA per SM, 384 threads, warp-specialised:
//   warp 0      TMA producer: cp.async.bulk.tensor (UTMALDG) of a 128x128-byte A box and two
//               128x128-byte B boxes per K block into a 4-stage, 128B-swizzled shared-memory ring
//   warp 1      MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::i8 (UTCIMMA), M=128,
//               N=256, K=32 per instruction, accumulating s32 in TMEM; tcgen05.commit releases the
//               smem stage / publishes the accumulator
//   warp 2      TMEM allocator (512 columns = two 128x256 s32 accumulators, double-buffered)
//   warps 4-11  epilogue: tcgen05.ld (LDTM) 32 columns at a time, d = |a|+|b|-2 dot, v = (d+1)^2,
//               row sums stay in the thread (one thread = one row), column sums are warp-reduced with
//               REDUX, partials are added to S with 64-bit integer atomics
// The tile-pair matrix is walked over its upper triangle in 256-row blocks (two 128x256 tiles per
// block pair); an off-diagonal tile contributes row sums and column sums (symmetry), a diagonal block
// is evaluated in full and contributes row sums only.  Tiles are dealt round-robin to CTAs and,
// across GPUs, to shards.
//
// Bound: tensor pipe.  Algorithmic work D*T*(T-1) int8 flop (unordered pairs); see DESIGN.md.
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


// Linear work index -> block pair (ib <= jb) of the upper triangle, walked in G x G supertiles so that the
// 2G operand blocks a wave of CTAs touches stay resident in L2 (a plain column-major walk streams one
// operand from HBM with no reuse: 55 TB at cfg5).  Slots that fall outside the triangle are invalid.
__device__ __forceinline__ bool pair_from_slot(long long slot, int nb, int G, int& ib, int& jb) {
  if (G == 0) {  // plain column-major walk of the triangle
    long long j = (long long)((sqrt(8.0 * (double)slot + 1.0) - 1.0) * 0.5);
    while (j * (j + 1) / 2 > slot) j--;
    while ((j + 1) * (j + 2) / 2 <= slot) j++;
    jb = (int)j;
    ib = (int)(slot - j * (j + 1) / 2);
    return jb < nb;
  }
  const long long g2 = (long long)G * G;
  const long long st = slot / g2;
  const int local = (int)(slot - st * g2);
  long long sj = (long long)((sqrt(8.0 * (double)st + 1.0) - 1.0) * 0.5);
  while (sj * (sj + 1) / 2 > st) sj--;
  while ((sj + 1) * (sj + 2) / 2 <= st) sj++;
  const long long si = st - sj * (sj + 1) / 2;
  ib = (int)(si * G) + local / G;
  jb = (int)(sj * G) + local % G;
  return ib <= jb && jb < nb;
}

struct TileInfo {
  int row_tile;   // 128-row tile index of the A operand
  int col_block;  // 256-row block index of the B operand
  bool do_row, do_col;
  int chain_rows, chain_cols;  // chain of the A rows / of the B rows
};

// PSRF mode: tile t of the upper triangle over 256-row blocks; dump mode: rectangular enumeration.
template <bool kDump>
__device__ __forceinline__ bool decode_tile(long long t, int n_row_tiles, int nb, int G, const int2* __restrict__ meta, TileInfo& ti) {
  if (kDump) {
    ti.row_tile = (int)(t % n_row_tiles);
    ti.col_block = (int)(t / n_row_tiles);
    ti.do_row = ti.do_col = false;
    ti.chain_rows = ti.chain_cols = 0;
    return true;
  }
  const int half = (int)(t & 1);
  int ib, jb;
  if (!pair_from_slot(t >> 1, nb, G, ib, jb)) return false;
  ti.row_tile = 2 * ib + half;
  ti.col_block = jb;
  const int2 mi = meta[ti.row_tile], mj0 = meta[2 * jb], mj1 = meta[2 * jb + 1];
  const bool diag = ib == jb;
  ti.do_row = (mi.y & kFlagRows) && (mj0.y & kFlagCol);
  ti.do_col = !diag && (mi.y & kFlagCol) && ((mj0.y | mj1.y) & kFlagRows);
  ti.chain_rows = mi.x;
  ti.chain_cols = mj0.x;
  return ti.do_row || ti.do_col;
}

template <bool kDump>
__global__ void __launch_bounds__(kThreads, 1)
rf_i8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int nkb,
             const int* __restrict__ nbitsA, const int* __restrict__ nbitsB, const int2* __restrict__ tile_meta,
             long long n_tiles, int n_row_tiles, int nb, int G, int shard_index, int shard_count,
             unsigned long long* __restrict__ S, int C, uint16_t* __restrict__ dump, long long dump_ld) {
  extern __shared__ __align__(1024) unsigned char smem[];  // 128B-swizzle atoms need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  unsigned char* aux = smem + (size_t)kStages * kStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);      // [kStages]
  uint64_t* empty = full + kStages;                       // [kStages]
  uint64_t* tfull = empty + kStages;                      // [2]
  uint64_t* tempty = tfull + 2;                           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + 2);
  int* nb_s = reinterpret_cast<int*>(aux + 128);          // [2][256]  |b| + 1 of the tile's columns
  uint32_t* colpart = reinterpret_cast<uint32_t*>(aux + 128 + 2 * kTileN * 4);  // [4][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const long long first = blockIdx.x, stride = gridDim.x;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long k = first;; k += stride) {
        const long long t = k * shard_count + shard_index;
        if (t >= n_tiles) break;
        TileInfo ti;
        if (!decode_tile<kDump>(t, n_row_tiles, nb, G, tile_meta, ti)) continue;
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + (size_t)stage * kStageBytes;
          mbar_expect_tx(&full[stage], kStageBytes);
          tma_load_2d(st, &tmA, kb * kBlockK, ti.row_tile * kTileRows, &full[stage]);
          tma_load_2d(st + kABytes, &tmB, kb * kBlockK, ti.col_block * kTileN, &full[stage]);
          tma_load_2d(st + kABytes + kABytes, &tmB, kb * kBlockK, ti.col_block * kTileN + kTileRows, &full[stage]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long k = first;; k += stride) {
        const long long t = k * shard_count + shard_index;
        if (t >= n_tiles) break;
        TileInfo ti;
        if (!decode_tile<kDump>(t, n_row_tiles, nb, G, tile_meta, ti)) continue;
        mbar_wait(&tempty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * kAccCols);
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * kStageBytes);
          const uint64_t adesc = make_desc(a_addr), bdesc = make_desc(a_addr + kABytes);
#pragma unroll
          for (int kk = 0; kk < kBlockK / 32; kk++)  // UMMA_K = 32 bytes; advance the start address inside the atom
            umma_i8(tmem_c, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), kIdesc, (uint32_t)((kb | kk) != 0));
          umma_commit(&empty[stage]);  // smem stage is free once these MMAs have read it
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue =====
    const int ew = warp - 4;        // 0..7
    const int quad = warp & 3;      // TMEM lane quadrant this warp may touch
    const int half = ew >> 2;       // which 128 of the 256 columns
    const int et = threadIdx.x - 128;
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    for (long long k = first;; k += stride) {
      const long long t = k * shard_count + shard_index;
      if (t >= n_tiles) break;
      TileInfo ti;
      if (!decode_tile<kDump>(t, n_row_tiles, nb, G, tile_meta, ti)) continue;
      const long long col0 = (long long)ti.col_block * kTileN;
      const long long row = (long long)ti.row_tile * kTileRows + quad * 32 + lane;
      nb_s[buf * kTileN + et] = nbitsB[col0 + et] + 1;
      const int na = nbitsA[row];
      epi_barrier();
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccCols + half * 128);
      unsigned long long row_sum = 0;
#pragma unroll 1
      for (int cc = 0; cc < 4; cc++) {
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        const int* nb = nb_s + buf * kTileN + half * 128 + cc * 32;
        if (kDump) {
          uint16_t* o = dump + row * dump_ld + col0 + half * 128 + cc * 32;
#pragma unroll
          for (int j = 0; j < 32; j++) o[j] = (uint16_t)(na + nb[j] - 1 - 2 * (int)r[j]);
        } else {
          uint32_t chunk = 0, mycol = 0;
          if (ti.do_col) {
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const uint32_t e = (uint32_t)(na + nb[j] - 2 * (int)r[j]);  // d + 1
              const uint32_t v = e * e;
              chunk += v;
              const uint32_t s = __reduce_add_sync(0xffffffffu, v);  // over the 32 rows of this warp
              if (lane == j) mycol = s;
            }
            colpart[quad * kTileN + half * 128 + cc * 32 + lane] = mycol;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const uint32_t e = (uint32_t)(na + nb[j] - 2 * (int)r[j]);
              chunk += e * e;
            }
          }
          row_sum += chunk;
        }
      }
      // all TMEM reads of this accumulator are complete: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (!kDump) {
        if (ti.do_row) atomicAdd(&S[(size_t)row * C + ti.chain_cols], row_sum);
        if (ti.do_col) {
          epi_barrier();
          const uint32_t v = colpart[et] + colpart[kTileN + et] + colpart[2 * kTileN + et] + colpart[3 * kTileN + et];
          atomicAdd(&S[(size_t)(col0 + et) * C + ti.chain_rows], (unsigned long long)v);
        }
      }
      buf ^= 1;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}


// =====================================================================================================
// 2-CTA variant (cta_group::2): a cluster of two CTAs on one TPC computes a 256x256 block pair.  Each
// CTA stages its own 128 A rows and its own half (128 rows) of B, so a CTA moves 32 KB per K block
// instead of 48 KB for the same 128x256 outputs -- L2 -> SM traffic per MAC drops by a third (ncu on
// the 1-CTA kernel: lts__throughput 76 %, the limiter).  The leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256); both CTAs' TMA loads signal the leader's `full` barrier; the
// leader's commits are multicast to both CTAs' `empty` / `tmem_full` barriers; the peer's epilogue
// warps arrive remotely on the leader's `tmem_empty` barrier.
// Work distribution is dynamic: warp 3 of the leader draws the next block pair from a global counter and
// publishes it through a 4-deep shared-memory queue to every role of both CTAs (remote store + remote
// mbarrier arrive for the peer).  The pairs in flight across the GPU are therefore always ~74 consecutive
// points of a Morton (Z-order) walk over the triangle, i.e. a compact 2-D patch: operand blocks are shared
// through L2 by CTAs that use them at the same time (static striding let the clusters drift apart: ncu on
// a cfg5-shaped run showed 49 % L2 hit rate and 5.9 TB/s of DRAM reads).
// =====================================================================================================
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
rf_i8x2_kernel(const __grid_constant__ CUtensorMap tm, int nkb, const int* __restrict__ nbits,
               const int2* __restrict__ tile_meta, long long n_slots, const uint32_t* __restrict__ walk_tiles, int nb,
               int shard_index, int shard_count, unsigned long long* __restrict__ slot_counter,
               unsigned long long* __restrict__ S, int C) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  unsigned char* aux = smem + (size_t)kStages2 * kStageBytes2;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);  // [kStages2]   (used in CTA 0)
  uint64_t* empty = full + kStages2;                  // [kStages2]   (each CTA)
  uint64_t* tfull = empty + kStages2;                 // [2]          (each CTA)
  uint64_t* tempty = tfull + 2;                       // [2]          (used in CTA 0)
  uint64_t* sfull = tempty + 2;                       // [kQ]         (each CTA)
  uint64_t* sempty = sfull + kQ;                      // [kQ]         (used in CTA 0)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sempty + kQ);
  int2* slots = reinterpret_cast<int2*>(aux + 208);   // [kQ] work queue entries {ib, jb}, {-1,-1} = stop
  int* nb_s = reinterpret_cast<int*>(aux + 256);
  uint32_t* colpart = reinterpret_cast<uint32_t*>(aux + 256 + 2 * kTileN * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 2 * kEpiWarps);
    }
    for (int q = 0; q < kQ; q++) {
      mbar_init(&sfull[q], 1);
      mbar_init(&sempty[q], kSchedConsumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();  // the peer's barriers exist before anything can arrive on them
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  WorkQueue wq;
  wq.sfull = sfull;
  wq.sempty = sempty;
  wq.slots = slots;
  wq.rank = rank;

  if (warp == 3) {
    // ===== scheduler (leader CTA): draws the next block pair of the Morton walk for the whole cluster =====
    if (rank == 0 && lane == 0) {
      int q = 0;
      uint32_t ph = 0;
      for (;;) {
        mbar_wait_cluster(&sempty[q], ph ^ 1);
        int ib = -1, jb = -1;
        for (;;) {
          const long long t = (long long)atomicAdd(slot_counter, 1ULL) * shard_count + shard_index;
          if (t >= n_slots) {
            ib = jb = -1;
            break;
          }
          // slot -> (8x8 Morton tile of the precomputed walk, Morton position inside the tile)
          const uint32_t tile = walk_tiles[t >> 6];
          const unsigned long long local = (unsigned long long)(t & 63);
          ib = (int)(tile & 0xFFFFu) * 8 + (int)compact_bits(local);
          jb = (int)(tile >> 16) * 8 + (int)compact_bits(local >> 1);
          PairInfo tmp;
          if (ib <= jb && jb < nb && pair_flags(ib, jb, 0, tile_meta, tmp)) break;
        }
        slots[q] = make_int2(ib, jb);
        st_int2_rank(&slots[q], 1, ib, jb);
        mbar_arrive_rank(&sfull[q], 0);
        mbar_arrive_rank(&sfull[q], 1);
        if (ib < 0) break;
        if (++q == kQ) {
          q = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 0) {
    // ===== TMA producer (both CTAs; completion bytes go to CTA 0's full barrier) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (;;) {
        const int2 e = wq.next(true);
        if (e.x < 0) break;
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + (size_t)stage * kStageBytes2;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * kStageBytes2);  // bytes of both CTAs
          tma_load_2d_2sm(st, &tm, kb * kBlockK, e.x * kBlockRows + rank * kTileRows, &full[stage]);
          tma_load_2d_2sm(st + kABytes, &tm, kb * kBlockK, e.y * kBlockRows + rank * kTileRows, &full[stage]);
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    // The whole warp walks the loop (stage, phase and the descriptors stay warp-uniform: no R2UR per K block, no
    // elect loop around every tcgen05 instruction) and one elected lane issues.
    if (rank == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t smem_base = smem_u32(smem);
      for (;;) {
        const int2 e = wq.next(false);
        if (e.x < 0) break;
        mbar_wait_cluster(&tempty[acc], acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * kAccCols);
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + (uint32_t)stage * kStageBytes2;
            const uint64_t adesc = make_desc(a_addr), bdesc = make_desc(a_addr + kABytes);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 32; kk++)
              umma_i8_2sm(tmem_c, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), kIdesc2, (uint32_t)((kb | kk) != 0));
            umma_commit_2sm(&empty[stage]);
          }
          __syncwarp();
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_2sm(&tfull[acc]);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (each CTA: its 128 rows x 256 columns) =====
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int half = ew >> 2;
    const int et = threadIdx.x - 128;
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    for (;;) {
      const int2 e = wq.next(false);
      if (e.x < 0) break;
      PairInfo pi;
      pair_flags(e.x, e.y, rank, tile_meta, pi);
      const long long col0 = (long long)pi.jb * kTileN;
      const long long row = (long long)pi.ib * kBlockRows + rank * kTileRows + quad * 32 + lane;
      nb_s[buf * kTileN + et] = nbits[col0 + et] + 1;
      const int na = nbits[row];
      epi_barrier();
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kAccCols + half * 128);
      unsigned long long row_sum = 0;
#pragma unroll 1
      for (int cc = 0; cc < 4; cc++) {
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        const int* nb = nb_s + buf * kTileN + half * 128 + cc * 32;
        uint32_t chunk = 0, mycol = 0;
        if (pi.do_col) {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const uint32_t e2 = (uint32_t)(na + nb[j] - 2 * (int)r[j]);
            const uint32_t v = e2 * e2;
            chunk += v;
            const uint32_t sred = __reduce_add_sync(0xffffffffu, v);
            if (lane == j) mycol = sred;
          }
          colpart[quad * kTileN + half * 128 + cc * 32 + lane] = mycol;
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) {
            const uint32_t e2 = (uint32_t)(na + nb[j] - 2 * (int)r[j]);
            chunk += e2 * e2;
          }
        }
        row_sum += chunk;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0)
          mbar_arrive(&tempty[acc]);
        else
          mbar_arrive_rank(&tempty[acc], 0);
      }
      if (pi.do_row) atomicAdd(&S[(size_t)row * C + pi.chain_cols], row_sum);
      if (pi.do_col) {
        epi_barrier();
        const uint32_t v = colpart[et] + colpart[kTileN + et] + colpart[2 * kTileN + et] + colpart[3 * kTileN + et];
        atomicAdd(&S[(size_t)(col0 + et) * C + pi.chain_rows], (unsigned long long)v);
      }
      buf ^= 1;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}


// =====================================================================================================
// K2-fp4: the same contraction on the block-scaled FP4 path of the tensor cores, which issues twice the
// K per instruction.  A clade indicator is an e2m1 nibble (1.0 = 0x2), every scale factor is 2^0 (UE8M0
// 0x7F in every byte of the scale-factor columns of TMEM, so their layout does not matter), accumulators are
// f32: a dot product of two indicator rows is an integer below 2^24 and therefore exact (tools/fp4_probe.cu).
// Differences from rf_i8x2_kernel:
//   * TMEM holds two 240-column accumulators + 2 x 16 scale-factor columns (512 in all), so a tile is 256 rows x
//     240 columns; the operand layout has 240 trees per 256-row block (PsrfLayout::block_real), the 16 padding
//     rows of the row operand are masked out of the column sums;
//   * cta_group::2 splits B in halves of N/2 = 120 rows: CTA 1 loads the rows from 120 on;
//   * one K block is 128 bytes = 256 clades, 4 MMAs of K = 64.
// =====================================================================================================
constexpr int kTileNF = 240;
constexpr int kSfaCol = 480, kSfbCol = 496;
// The MMA needs half the time per tile, so the epilogue gets more warps: 12, three per TMEM lane quadrant,
// each with five of the 15 chunks of 16 columns.
constexpr int kEpiWarpsF = 12;
constexpr int kThreadsF = 128 + kEpiWarpsF * 32;
constexpr int kSchedConsumersF = 2 + kEpiWarpsF + 1 + kEpiWarpsF;
constexpr uint32_t kF32Magic = 0x4AC00000u;  // bits of 1.5 * 2^22 (ulp 0.5): see the epilogue
constexpr uint32_t kIdescF4 = (1u << 7) | (1u << 10) | ((uint32_t)(kTileNF >> 3) << 17) | (1u << 23) | ((uint32_t)(256 >> 4) << 24);

__device__ __forceinline__ void umma_f4_2sm(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate,
                                            uint32_t sfa, uint32_t sfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_c),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(sfa), "r"(sfb)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsF, 1)
rf_f4x2_kernel(const __grid_constant__ CUtensorMap tm, int nkb, const int* __restrict__ nbits,
               const int2* __restrict__ tile_meta, long long n_slots, const uint32_t* __restrict__ walk_tiles, int nb,
               int shard_index, int shard_count, unsigned long long* __restrict__ slot_counter,
               unsigned long long* __restrict__ S, int C, int dbg) {
  extern __shared__ __align__(1024) unsigned char smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  unsigned char* aux = smem + (size_t)kStages2 * kStageBytes2;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);  // [kStages2]   (used in CTA 0)
  uint64_t* empty = full + kStages2;                  // [kStages2]   (each CTA)
  uint64_t* tfull = empty + kStages2;                 // [2]          (each CTA)
  uint64_t* tempty = tfull + 2;                       // [2]          (used in CTA 0)
  uint64_t* sfull = tempty + 2;                       // [kQ]         (each CTA)
  uint64_t* sempty = sfull + kQ;                      // [kQ]         (used in CTA 0)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(sempty + kQ);
  int2* slots = reinterpret_cast<int2*>(aux + 208);   // [kQ] work queue entries {ib, jb}, {-1,-1} = stop
  int* nb_s = reinterpret_cast<int*>(aux + 256);
  uint32_t* colpart = reinterpret_cast<uint32_t*>(aux + 256 + 2 * kTileN * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 2 * kEpiWarpsF);
    }
    for (int q = 0; q < kQ; q++) {
      mbar_init(&sfull[q], 1);
      mbar_init(&sempty[q], kSchedConsumersF);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync_all();  // the peer's barriers exist before anything can arrive on them
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp >= 4 && warp < 8) {  // unit scale factors in both CTAs: 32 columns x 128 lanes of 0x7F bytes
    const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)kSfaCol;
    const uint32_t v = 0x7F7F7F7Fu;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(t),
        "r"(v)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's scale factors are in place before the leader issues for both CTAs
  tc_fence_after();

  WorkQueue wq;
  wq.sfull = sfull;
  wq.sempty = sempty;
  wq.slots = slots;
  wq.rank = rank;

  if (warp == 3) {
    // ===== scheduler (leader CTA): draws the next block pair of the Morton walk for the whole cluster =====
    if (rank == 0 && lane == 0) {
      int q = 0;
      uint32_t ph = 0;
      for (;;) {
        mbar_wait_cluster(&sempty[q], ph ^ 1);
        int ib = -1, jb = -1;
        for (;;) {
          const long long t = (long long)atomicAdd(slot_counter, 1ULL) * shard_count + shard_index;
          if (t >= n_slots) {
            ib = jb = -1;
            break;
          }
          const uint32_t tile = walk_tiles[t >> 6];
          const unsigned long long local = (unsigned long long)(t & 63);
          ib = (int)(tile & 0xFFFFu) * 8 + (int)compact_bits(local);
          jb = (int)(tile >> 16) * 8 + (int)compact_bits(local >> 1);
          PairInfo tmp;
          if (ib <= jb && jb < nb && pair_flags(ib, jb, 0, tile_meta, tmp)) break;
        }
        slots[q] = make_int2(ib, jb);
        st_int2_rank(&slots[q], 1, ib, jb);
        mbar_arrive_rank(&sfull[q], 0);
        mbar_arrive_rank(&sfull[q], 1);
        if (ib < 0) break;
        if (++q == kQ) {
          q = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 0) {
    // ===== TMA producer (both CTAs; completion bytes go to CTA 0's full barrier) =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (;;) {
        const int2 e = wq.next(true);
        if (e.x < 0) break;
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + (size_t)stage * kStageBytes2;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * kStageBytes2);  // bytes of both CTAs
          tma_load_2d_2sm(st, &tm, kb * kBlockK, e.x * kBlockRows + rank * kTileRows, &full[stage]);
          // B half of this CTA: columns [rank*120, rank*120 + 120) of the tile (the box brings 8 rows more)
          tma_load_2d_2sm(st + kABytes, &tm, kb * kBlockK, e.y * kBlockRows + rank * (kTileNF / 2), &full[stage]);
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    // The whole warp walks the loop (so that stage, phase and the descriptors are warp-uniform and live in the
    // uniform datapath: no R2UR per K block) and one elected lane issues.
    if (rank == 0) {
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      const uint32_t sfa = tmem_base + (uint32_t)kSfaCol, sfb = tmem_base + (uint32_t)kSfbCol;
      const uint32_t smem_base = smem_u32(smem);
      for (;;) {
        const int2 e = wq.next(false);
        if (e.x < 0) break;
        mbar_wait_cluster(&tempty[acc], acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * kTileNF);
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + (uint32_t)stage * kStageBytes2;
            const uint64_t adesc = make_desc(a_addr), bdesc = make_desc(a_addr + kABytes);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 32; kk++)
              umma_f4_2sm(tmem_c, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), kIdescF4, (uint32_t)((kb | kk) != 0),
                          sfa, sfb);
            umma_commit_2sm(&empty[stage]);
          }
          __syncwarp();
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit_2sm(&tfull[acc]);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (each CTA: its 128 rows x 240 columns) =====
    // Integer arithmetic on the f32 accumulators in three instructions per element and without F2I (quarter rate):
    // with c = 1.5*2^22 + |a|/2 (ulp 0.5, exact), bits(c - dot) = bits(1.5*2^22) + |a| - 2 dot, and nb_s holds
    // |b| + 1 - bits(1.5*2^22), so d + 1 = bits(c - dot) + nb_s (mod 2^32).
    // The accumulator is read into registers in one go and handed back to the MMA warp before the arithmetic
    // starts, so the tensor core never waits for the epilogue's math, only for its TMEM reads.
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int slice = ew >> 2;
    const int et = threadIdx.x - 128;
    constexpr int kSlices = kEpiWarpsF / 4, kChunks = kTileNF / 16, kCPW = kChunks / kSlices;
    static_assert(kCPW * kSlices == kChunks, "chunks must split evenly over the warps of a quadrant");
    const int c_begin = slice * kCPW;
    // rows 240..255 of a block are padding: they must not count as columns of the statistic in the column sums
    const bool pad_warp = rank == 1 && quad == 3;
    const uint32_t col_mask = (rank * kTileRows + quad * 32 + lane) < kTileNF ? 0xFFFFFFFFu : 0u;
    uint32_t* nbu = reinterpret_cast<uint32_t*>(nb_s);
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    for (;;) {
      const int2 e = wq.next(false);
      if (e.x < 0) break;
      PairInfo pi;
      pair_flags(e.x, e.y, rank, tile_meta, pi);
      const long long col0 = (long long)pi.jb * kBlockRows;
      const long long row = (long long)pi.ib * kBlockRows + rank * kTileRows + quad * 32 + lane;
      if (et < kTileNF) nbu[buf * kTileN + et] = (uint32_t)(nbits[col0 + et] + 1) - kF32Magic;
      const float cfa = __fadd_rn(__uint_as_float(kF32Magic), __fmul_rn(0.5f, (float)nbits[row]));
      asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarpsF * 32) : "memory");
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileNF + c_begin * 16);
      uint32_t r[kCPW][16];
#pragma unroll
      for (int cc = 0; cc < kCPW; cc++) tmem_ld16_nowait(taddr + cc * 16, r[cc]);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int cc = 0; cc < kCPW; cc++)  // the registers are defined only now: nothing that reads them may move above the wait
        asm volatile("" : "+r"(r[cc][0]), "+r"(r[cc][1]), "+r"(r[cc][2]), "+r"(r[cc][3]), "+r"(r[cc][4]), "+r"(r[cc][5]),
                          "+r"(r[cc][6]), "+r"(r[cc][7]), "+r"(r[cc][8]), "+r"(r[cc][9]), "+r"(r[cc][10]), "+r"(r[cc][11]),
                          "+r"(r[cc][12]), "+r"(r[cc][13]), "+r"(r[cc][14]), "+r"(r[cc][15]));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // the TMEM reads are complete and fenced: the accumulator goes back before the math
        if (rank == 0)
          mbar_arrive_relaxed(&tempty[acc], (int)r[0][0]);
        else
          mbar_arrive_rank_relaxed(&tempty[acc], 0, (int)r[0][0]);
      }
      unsigned long long row_sum = 0;
      if (!(dbg & 1)) {
#pragma unroll
        for (int cc = 0; cc < kCPW; cc++) {
          uint32_t nbv[16];
          {
            const uint4* nb4 = reinterpret_cast<const uint4*>(nbu + buf * kTileN + (c_begin + cc) * 16);
#pragma unroll
            for (int q = 0; q < 4; q++) {
              const uint4 t = nb4[q];
              nbv[4 * q] = t.x, nbv[4 * q + 1] = t.y, nbv[4 * q + 2] = t.z, nbv[4 * q + 3] = t.w;
            }
          }
          uint32_t chunk = 0;
          if (pi.do_col) {
            uint32_t cs[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const uint32_t e2 = __float_as_uint(__fsub_rn(cfa, __uint_as_float(r[cc][j]))) + nbv[j];
              const uint32_t v = e2 * e2;
              chunk += v;
              cs[j] = __reduce_add_sync(0xffffffffu, pad_warp ? (v & col_mask) : v);
            }
            if (lane == 0) {
              uint4* dst = reinterpret_cast<uint4*>(colpart + quad * kTileN + (c_begin + cc) * 16);
#pragma unroll
              for (int q = 0; q < 4; q++) dst[q] = make_uint4(cs[4 * q], cs[4 * q + 1], cs[4 * q + 2], cs[4 * q + 3]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const uint32_t e2 = __float_as_uint(__fsub_rn(cfa, __uint_as_float(r[cc][j]))) + nbv[j];
              chunk += e2 * e2;
            }
          }
          row_sum += chunk;
        }
      }
      if (pi.do_row && !(dbg & 2)) atomicAdd(&S[(size_t)row * C + pi.chain_cols], row_sum);
      if (pi.do_col && !(dbg & 2)) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarpsF * 32) : "memory");
        if (et < kTileNF) {
          const uint32_t v = colpart[et] + colpart[kTileN + et] + colpart[2 * kTileN + et] + colpart[3 * kTileN + et];
          atomicAdd(&S[(size_t)(col0 + et) * C + pi.chain_rows], (unsigned long long)v);
        }
      }
      buf ^= 1;
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other may still signal it
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

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

template <bool kDump>
int32_t set_smem_attr(asm_ctx* ctx) {
  const uint32_t bit = kDump ? kAttrI8Dump : kAttrI8;
  if (!(ctx->func_attr_done & bit)) {
    ASM_CUDA(ctx, cudaFuncSetAttribute(rf_i8_kernel<kDump>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    ctx->func_attr_done |= bit;
  }
  return ASM_OK;
}

// rows of one chain -> zero-padded 0/1 byte matrix + popcounts (operands of the dump-mode probe)
__global__ void __launch_bounds__(256) expand_rows_kernel(const uint64_t* __restrict__ src, int words, long long n_live,
                                                           long long n_pad, int Wp, uint4* __restrict__ out,
                                                           int* __restrict__ nbits) {
  const int lane = threadIdx.x & 31;
  const long long wpb = blockDim.x >> 5;
  for (long long p = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); p < n_pad; p += (long long)gridDim.x * wpb) {
    int cnt = 0;
    for (int w = lane; w < Wp; w += 32) {
      const uint64_t v = (p < n_live && w < words) ? src[(size_t)p * words + w] : 0ULL;
      cnt += __popcll(v);
      uint4 q[4];
      uint32_t* o = reinterpret_cast<uint32_t*>(q);
#pragma unroll
      for (int k = 0; k < 16; k++) o[k] = (((uint32_t)(v >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;
#pragma unroll
      for (int k = 0; k < 4; k++) out[((size_t)p * Wp + w) * 4 + k] = q[k];
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) nbits[p] = cnt;
  }
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

static bool use_two_cta() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ASM_I8_1CTA");
    v = (e && e[0] && e[0] != '0') ? 0 : 1;
  }
  return v == 1;
}

int32_t rf_i8_sums(asm_ctx* ctx, asm_shard shard) {
  PsrfLayout& L = ctx->layout;
  const long long nb = L.Tp / kBlockRows;
  const uint64_t Dp = (uint64_t)L.pitch8;
  L.bits_used = (int64_t)Dp;
  // The epilogue adds 32 values of (d+1)^2 in 32 bits (REDUX / per-chunk row sums): d + 1 <= 2*maxbits + 1 must stay
  // below sqrt(2^32 / 32) = 11585, i.e. rows with at most 5,791 clades.  Rows cannot have more bits than the
  // operand is wide, so narrow dictionaries need no check; wide ones are checked against the row popcounts.
  if (Dp >= 5792) {
    int32_t rcg = psrf_check_max_bits(ctx, 5791);
    if (rcg != ASM_OK) return rcg;
  }
  CUtensorMap tm;
  int32_t rc = make_tmap(ctx, &tm, ctx->opnd_i8.p, (uint64_t)L.Tp, Dp);
  if (rc != ASM_OK) return rc;
  // Supertile edge.  Measured (tools/psrf_probe.py, profiles/): G = 2..8 is best at every size (T = 40k: 1.47 ms vs
  // 1.46 column-major; T = 400k, D = 5116: 228 ms vs 285 ms); larger G loses more to the idle slots of the
  // diagonal supertiles than it gains in L2 reuse.
  int G = 4;
  if (const char* e = getenv("ASM_I8_SUPERTILE")) G = atoi(e);  // 0 = plain column-major walk (tuning knob)
  const long long n_pairs = nb * (nb + 1) / 2;
  const long long nst = G > 0 ? (nb + G - 1) / G : 0;
  const long long n_slots = G > 0 ? nst * (nst + 1) / 2 * G * G : n_pairs;  // incl. invalid slots outside the triangle
  if (use_two_cta()) {
    L.tiles_total = n_pairs;  // 256x256 block pairs of the upper triangle
    L.tiles_done = (n_pairs - shard.shard_index + shard.shard_count - 1) / shard.shard_count;
    if (!(ctx->func_attr_done & kAttrI8x2)) {
      ASM_CUDA(ctx, cudaFuncSetAttribute(rf_i8x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes2x));
      ctx->func_attr_done |= kAttrI8x2;
    }
    if ((rc = ensure_walk(ctx, nb)) != ASM_OK) return rc;
    const long long n_morton = (long long)ctx->walk_host.size() * 64;
    if ((rc = asm_reserve(ctx, ctx->tile_counter, sizeof(unsigned long long))) != ASM_OK) return rc;
    ASM_CUDA(ctx, cudaMemsetAsync(ctx->tile_counter.p, 0, sizeof(unsigned long long), ctx->stream));
    const int clusters = (int)std::min<long long>(L.tiles_done, ctx->sm_count / 2);
    if (clusters <= 0) return ASM_OK;
    {
      LaunchScope ls(ctx, ASM_K_RF_I8);
      rf_i8x2_kernel<<<2 * clusters, kThreads, kSmemBytes2x, ctx->stream>>>(
          tm, (int)(Dp / kBlockK), (const int*)ctx->nbits.p, (const int2*)ctx->blk_meta.p, n_morton,
          (const uint32_t*)ctx->walk_dev.p, (int)nb, shard.shard_index, shard.shard_count,
          (unsigned long long*)ctx->tile_counter.p,
          (unsigned long long*)ctx->S_pad.p, L.C);
    }
    ASM_CUDA(ctx, cudaGetLastError());
    return ASM_OK;
  }
  L.tiles_total = 2 * n_pairs;  // two 128x256 tiles per block pair of the upper triangle
  L.tiles_done = (2 * n_pairs - shard.shard_index + shard.shard_count - 1) / shard.shard_count;
  if ((rc = set_smem_attr<false>(ctx)) != ASM_OK) return rc;
  const int grid = (int)std::min<long long>(L.tiles_done, ctx->sm_count);
  if (grid <= 0) return ASM_OK;
  {
    LaunchScope ls(ctx, ASM_K_RF_I8);
    rf_i8_kernel<false><<<grid, kThreads, kSmemBytes, ctx->stream>>>(
        tm, tm, (int)(Dp / kBlockK), (const int*)ctx->nbits.p, (const int*)ctx->nbits.p, (const int2*)ctx->blk_meta.p,
        2 * n_slots, 0, (int)nb, G, shard.shard_index, shard.shard_count, (unsigned long long*)ctx->S_pad.p, L.C, nullptr,
        0);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int32_t rf_f4_sums(asm_ctx* ctx, asm_shard shard) {
  PsrfLayout& L = ctx->layout;
  if (L.block_real != kTileNF) return asm_fail(ctx, ASM_E_INVALID, "rf_f4: layout built for %d live rows per block", L.block_real);
  const long long nb = L.Tp / kBlockRows;
  const uint64_t pitch = (uint64_t)L.pitch4;
  L.bits_used = (int64_t)pitch * 2;
  if (pitch * 2 >= 5792) {  // same 32-bit epilogue sums as the int8 kernel
    int32_t rcg = psrf_check_max_bits(ctx, 5791);
    if (rcg != ASM_OK) return rcg;
  }
  CUtensorMap tm;
  int32_t rc = make_tmap(ctx, &tm, ctx->opnd_f4.p, (uint64_t)L.Tp, pitch);
  if (rc != ASM_OK) return rc;
  const long long n_pairs = nb * (nb + 1) / 2;
  L.tiles_total = n_pairs;
  L.tiles_done = (n_pairs - shard.shard_index + shard.shard_count - 1) / shard.shard_count;
  if (!(ctx->func_attr_done & kAttrF4x2)) {
    ASM_CUDA(ctx, cudaFuncSetAttribute(rf_f4x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes2x));
    ctx->func_attr_done |= kAttrF4x2;
  }
  if ((rc = ensure_walk(ctx, nb)) != ASM_OK) return rc;
  const long long n_morton = (long long)ctx->walk_host.size() * 64;
  if ((rc = asm_reserve(ctx, ctx->tile_counter, sizeof(unsigned long long))) != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemsetAsync(ctx->tile_counter.p, 0, sizeof(unsigned long long), ctx->stream));
  const int clusters = (int)std::min<long long>(L.tiles_done, ctx->sm_count / 2);
  if (clusters <= 0) return ASM_OK;
  {
    LaunchScope ls(ctx, ASM_K_RF_I8);
    rf_f4x2_kernel<<<2 * clusters, kThreadsF, kSmemBytes2x, ctx->stream>>>(
        tm, (int)(pitch / kBlockK), (const int*)ctx->nbits.p, (const int2*)ctx->blk_meta.p, n_morton,
        (const uint32_t*)ctx->walk_dev.p, (int)nb, shard.shard_index, shard.shard_count,
        (unsigned long long*)ctx->tile_counter.p, (unsigned long long*)ctx->S_pad.p, L.C,
        getenv("ASM_F4_DEBUG") ? atoi(getenv("ASM_F4_DEBUG")) : 0);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int32_t rf_i8_block(asm_ctx* ctx, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0, int64_t c1,
                    uint16_t* out_host) {
  const long long nr = r1 - r0, nc = c1 - c0;
  const long long nr_pad = (nr + kTileRows - 1) / kTileRows * kTileRows;
  const long long nc_pad = (nc + kTileN - 1) / kTileN * kTileN;
  const int Wp = (ctx->words + 15) / 16 * 16;
  const uint64_t Dp = (uint64_t)Wp * 64;
  int32_t rc;
  // scratch: A8 | B8 in opnd_i8, popcounts in nbits, distances in rf_out
  if ((rc = asm_reserve(ctx, ctx->opnd_i8, (size_t)(nr_pad + nc_pad) * Dp)) != ASM_OK) return rc;
  if ((rc = asm_reserve(ctx, ctx->nbits, sizeof(int) * (size_t)(nr_pad + nc_pad))) != ASM_OK) return rc;
  if ((rc = asm_reserve(ctx, ctx->rf_out, sizeof(uint16_t) * (size_t)(nr_pad * nc_pad))) != ASM_OK) return rc;
  unsigned char* A8 = (unsigned char*)ctx->opnd_i8.p;
  unsigned char* B8 = A8 + (size_t)nr_pad * Dp;
  int* nbA = (int*)ctx->nbits.p;
  int* nbB = nbA + nr_pad;
  {
    LaunchScope ls(ctx, ASM_K_GATHER);
    expand_rows_kernel<<<(unsigned)std::min<long long>((nr_pad + 7) / 8, ctx->sm_count * 8LL), 256, 0, ctx->stream>>>(
        ctx->trees[chainA].d_bits + (size_t)r0 * ctx->words, ctx->words, nr, nr_pad, Wp, (uint4*)A8, nbA);
  }
  {
    LaunchScope ls(ctx, ASM_K_GATHER);
    expand_rows_kernel<<<(unsigned)std::min<long long>((nc_pad + 7) / 8, ctx->sm_count * 8LL), 256, 0, ctx->stream>>>(
        ctx->trees[chainB].d_bits + (size_t)c0 * ctx->words, ctx->words, nc, nc_pad, Wp, (uint4*)B8, nbB);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap(ctx, &tmA, A8, (uint64_t)nr_pad, Dp)) != ASM_OK) return rc;
  if ((rc = make_tmap(ctx, &tmB, B8, (uint64_t)nc_pad, Dp)) != ASM_OK) return rc;
  if ((rc = set_smem_attr<true>(ctx)) != ASM_OK) return rc;
  const int n_row_tiles = (int)(nr_pad / kTileRows);
  const long long n_tiles = (long long)n_row_tiles * (nc_pad / kTileN);
  {
    LaunchScope ls(ctx, ASM_K_RF_I8);
    rf_i8_kernel<true><<<(int)std::min<long long>(n_tiles, ctx->sm_count), kThreads, kSmemBytes, ctx->stream>>>(
        tmA, tmB, (int)(Dp / kBlockK), nbA, nbB, nullptr, n_tiles, n_row_tiles, 0, 1, 0, 1, nullptr, 0, (uint16_t*)ctx->rf_out.p,
        nc_pad);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  ASM_CUDA(ctx, cudaMemcpy2DAsync(out_host, sizeof(uint16_t) * (size_t)nc, ctx->rf_out.p, sizeof(uint16_t) * (size_t)nc_pad,
                                  sizeof(uint16_t) * (size_t)nc, (size_t)nr, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}
