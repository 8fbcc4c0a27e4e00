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
#include "tc_common.cuh"

namespace {

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
          // each partial is a sum of 32 squares (< 2^32 by the clade-count guard); their sum of 128 may not be
          const unsigned long long v = (unsigned long long)colpart[et] + colpart[kTileN + et] + colpart[2 * kTileN + et] +
                                       colpart[3 * kTileN + et];
          atomicAdd(&S[(size_t)(col0 + et) * C + ti.chain_rows], v);
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
        // each partial is a sum of 32 squares (< 2^32 by the clade-count guard); their sum of 128 may not be
          const unsigned long long v = (unsigned long long)colpart[et] + colpart[kTileN + et] + colpart[2 * kTileN + et] +
                                       colpart[3 * kTileN + et];
        atomicAdd(&S[(size_t)(col0 + et) * C + pi.chain_rows], v);
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
