// rf_f4.cu -- K2-fp4: the all-pairs Robinson-Foulds contraction of rf_i8.cu on the block-scaled FP4 path of the tensor
// cores (exact on 0/1 clade indicators), fused with the PSRF sums.  Shared machinery: tc_common.cuh; design: DESIGN.md.
#include "tc_common.cuh"

namespace {

// =====================================================================================================
// K2-fp4: the same contraction on the block-scaled FP4 path of the tensor cores, which issues twice the
// K per instruction.  A clade indicator is an e2m1 nibble (1.0 = 0x2), every scale factor is 2^0 (UE8M0
// 0x7F in every byte of the scale-factor columns of TMEM, so their layout does not matter), accumulators are
// f32: a dot product of two indicator rows is an integer below 2^24 and therefore exact (tools/fp4_probe.cu).
// Differences from rf_i8x2_kernel:
//   * TMEM holds two 240-column accumulators + 2 x 16 scale-factor columns (512 in all), so a tile is 256 rows x
//     240 columns.  Rows and columns are blocked INDEPENDENTLY over the same dense operand rows (round 2): row blocks
//     of 256 trees, column blocks of 240 trees (a TMA box may start at any row), so neither operand carries padding
//     inside a segment (round 1 kept 240 trees per 256-row block: 1.14x the MMAs).  The pair space is then a
//     rectangle of (row block, column block) tiles instead of a triangle of equal blocks: every unordered pair of
//     padded rows {p < q} is owned by the tile that has p among its rows and q among its columns.  Tiles entirely
//     above the diagonal are evaluated in full (row sums + the symmetric column sums), tiles entirely below it are
//     skipped, the few tiles that straddle it mask the pairs they do not own (q < p) in the epilogue;
//   * cta_group::2 splits B in halves of N/2 = 120 rows: CTA 1 loads the rows from 120 on;
//   * one K block is 128 bytes = 256 clades, 4 MMAs of K = 64; the last K block issues only as many as the widest
//     row needs (the gather records the highest non-zero word).
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
// shared memory behind the TMA ring: barriers + tmem pointer (256 B), work queue (kQ x 16 B), nb+1 (double buffered),
// column partials
constexpr int kAuxQueueOff = 256, kAuxNbOff = 320;
constexpr size_t kAuxBytesF = kAuxNbOff + 2 * kTileN * 4 + 4 * kTileN * 4;
constexpr size_t kSmemBytesF = (size_t)kStages2 * kStageBytes2 + kAuxBytesF;

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
__device__ __forceinline__ void st_int4_rank(int4* p, uint32_t rank, int4 v) {  // store into CTA `rank`'s copy of *p
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.v4.s32 [ra], {%2, %3, %4, %5};\n\t}" ::"r"(smem_u32(p)),
      "r"(rank), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
      : "memory");
}

// Work item of the queue: {first row of the row block, first row of the column block, meta of the row block, meta of
// the column block}; meta = chain | flags << 8 (psrf_build_f4_blocks).  x < 0 = stop.
struct TileF4 {
  int row_base, col_base;
  int chain_rows, chain_cols;
  bool do_row, do_col;  // for THIS CTA's 128 rows
  bool diag;            // the tile straddles the diagonal: pairs with column < row belong to the transposed tile
};
__device__ __forceinline__ TileF4 decode_tile(int4 e, int rank) {
  TileF4 t;
  t.row_base = e.x;
  t.col_base = e.y;
  const uint32_t fr = (uint32_t)e.z >> 8, fc = (uint32_t)e.w >> 8;
  t.chain_rows = e.z & 0xFF;
  t.chain_cols = e.w & 0xFF;
  t.do_row = ((fr >> (1 + rank)) & 1u) && (fc & kFlagCol);
  t.do_col = (fr & kFlagCol) && (fc & 2u);
  t.diag = !(e.x + (kBlockRows - 1) < e.y);
  return t;
}
// Any work in the pair (row block r, column block c)?  The same answer in both CTAs of the pair.
__device__ __forceinline__ bool tile_has_work(int2 r, int2 c) {
  if (c.x + (kTileNF - 1) < r.x) return false;  // every column precedes every row: owned by the transposed tile
  const uint32_t fr = (uint32_t)r.y >> 8, fc = (uint32_t)c.y >> 8;
  return ((fr & 6u) && (fc & kFlagCol)) || ((fr & kFlagCol) && (fc & 2u));
}

struct WorkQueueF4 {
  uint64_t* sfull;
  uint64_t* sempty;
  const int4* slots;
  int q = 0;
  uint32_t ph = 0;
  int rank;
  __device__ __forceinline__ int4 next(bool solo) {
    mbar_wait_cluster(&sfull[q], ph);
    const int4 e = slots[q];
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

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsF, 1)
rf_f4x2_kernel(const __grid_constant__ CUtensorMap tm, const int* __restrict__ top_word, const int* __restrict__ nbits,
               const int2* __restrict__ blk_rows, const int2* __restrict__ blk_cols, int nR, int nC, long long n_slots,
               const uint32_t* __restrict__ walk_tiles, int shard_index, int shard_count,
               unsigned long long* __restrict__ slot_counter, unsigned long long* __restrict__ S, int C, int dbg) {
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
  int4* slots = reinterpret_cast<int4*>(aux + kAuxQueueOff);  // [kQ] work queue entries
  int* nb_s = reinterpret_cast<int*>(aux + kAuxNbOff);
  uint32_t* colpart = reinterpret_cast<uint32_t*>(aux + kAuxNbOff + 2 * kTileN * 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  // K extent: nkb whole 128-byte K blocks up to the highest non-zero word of any gathered row (found by the gather
  // kernel that ran before this one, so the host never waits for it); the last block issues `last_kk` of its four
  // K = 64 instructions
  const int used_bits = (__ldg(top_word) + 1) * 64;
  const int nkb = (used_bits + 2 * kBlockK - 1) / (2 * kBlockK);
  const int last_kk = (used_bits - (nkb - 1) * 2 * kBlockK + 63) / 64;

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

  WorkQueueF4 wq;
  wq.sfull = sfull;
  wq.sempty = sempty;
  wq.slots = slots;
  wq.rank = rank;

  if (warp == 3) {
    // ===== scheduler (leader CTA): draws the next tile of the Morton walk for the whole cluster =====
    if (rank == 0 && lane == 0) {
      int q = 0;
      uint32_t ph = 0;
      for (;;) {
        mbar_wait_cluster(&sempty[q], ph ^ 1);
        int4 e = make_int4(-1, -1, 0, 0);
        for (;;) {
          const long long t = (long long)atomicAdd(slot_counter, 1ULL) * shard_count + shard_index;
          if (t >= n_slots) break;
          const uint32_t tile = walk_tiles[t >> 6];
          const unsigned long long local = (unsigned long long)(t & 63);
          const int I = (int)(tile & 0xFFFFu) * 8 + (int)compact_bits(local);
          const int J = (int)(tile >> 16) * 8 + (int)compact_bits(local >> 1);
          if (I >= nR || J >= nC) continue;
          const int2 r = blk_rows[I], c = blk_cols[J];
          if (tile_has_work(r, c)) {
            e = make_int4(r.x, c.x, r.y, c.y);
            break;
          }
        }
        slots[q] = e;
        st_int4_rank(&slots[q], 1, e);
        mbar_arrive_rank(&sfull[q], 0);
        mbar_arrive_rank(&sfull[q], 1);
        if (e.x < 0) break;
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
        const int4 e = wq.next(true);
        if (e.x < 0) break;
        for (int kb = 0; kb < nkb; kb++) {
          mbar_wait(&empty[stage], phase ^ 1);
          unsigned char* st = smem + (size_t)stage * kStageBytes2;
          if (rank == 0) mbar_expect_tx(&full[stage], 2 * kStageBytes2);  // bytes of both CTAs
          tma_load_2d_2sm(st, &tm, kb * kBlockK, e.x + rank * kTileRows, &full[stage]);
          // B half of this CTA: columns [rank*120, rank*120 + 120) of the tile (the box brings 8 rows more)
          tma_load_2d_2sm(st + kABytes, &tm, kb * kBlockK, e.y + rank * (kTileNF / 2), &full[stage]);
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
        const int4 e = wq.next(false);
        if (e.x < 0) break;
        mbar_wait_cluster(&tempty[acc], acc_phase ^ 1);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_c = tmem_base + (uint32_t)(acc * kTileNF);
        for (int kb = 0; kb < nkb - 1; kb++) {
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
        {  // last K block: only the instructions that cover a non-zero word of some row
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_base + (uint32_t)stage * kStageBytes2;
            const uint64_t adesc = make_desc(a_addr), bdesc = make_desc(a_addr + kABytes);
#pragma unroll
            for (int kk = 0; kk < kBlockK / 32; kk++)
              if (kk < last_kk)
                umma_f4_2sm(tmem_c, adesc + (uint64_t)(kk * 2), bdesc + (uint64_t)(kk * 2), kIdescF4,
                            (uint32_t)((nkb > 1) | (kk != 0)), sfa, sfb);
            umma_commit_2sm(&empty[stage]);
            umma_commit_2sm(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == kStages2) {
            stage = 0;
            phase ^= 1;
          }
        }
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
    uint32_t* nbu = reinterpret_cast<uint32_t*>(nb_s);
    int acc = 0, buf = 0;
    uint32_t acc_phase = 0;
    for (;;) {
      const int4 e = wq.next(false);
      if (e.x < 0) break;
      const TileF4 pi = decode_tile(e, rank);
      const int col0 = pi.col_base;
      const int row = pi.row_base + rank * kTileRows + quad * 32 + lane;
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
        if (!pi.diag) {
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
                cs[j] = __reduce_add_sync(0xffffffffu, v);
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
        } else {
          // The tile straddles the diagonal.  Pair (row p, column q) is owned here iff p <= q (padded row indices):
          // q > p feeds the row sum of p and the column sum of q, q == p (the tree against itself, d + 1 = 1) the
          // row sum only, q < p belongs to the tile that has q among its rows.
#pragma unroll
          for (int cc = 0; cc < kCPW; cc++) {
            const uint32_t* nbp = nbu + buf * kTileN + (c_begin + cc) * 16;
            const int dq0 = col0 + (c_begin + cc) * 16 - row;  // column index minus row index of element j = 0
            uint32_t chunk = 0;
            uint32_t cs[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const uint32_t e2 = __float_as_uint(__fsub_rn(cfa, __uint_as_float(r[cc][j]))) + nbp[j];
              const uint32_t v = e2 * e2;
              chunk += (dq0 + j >= 0) ? v : 0u;
              cs[j] = __reduce_add_sync(0xffffffffu, (dq0 + j > 0) ? v : 0u);
            }
            if (lane == 0) {
              uint4* dst = reinterpret_cast<uint4*>(colpart + quad * kTileN + (c_begin + cc) * 16);
#pragma unroll
              for (int q = 0; q < 4; q++) dst[q] = make_uint4(cs[4 * q], cs[4 * q + 1], cs[4 * q + 2], cs[4 * q + 3]);
            }
            row_sum += chunk;
          }
        }
      }
      if (pi.do_row && !(dbg & 2)) atomicAdd(&S[(size_t)row * C + pi.chain_cols], row_sum);
      if (pi.do_col && !(dbg & 2)) {
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarpsF * 32) : "memory");
        if (et < kTileNF) {
          // each partial is a sum of 32 squares (< 2^32 by the clade-count guard); their sum of 128 may not be
          const unsigned long long v = (unsigned long long)colpart[et] + colpart[kTileN + et] + colpart[2 * kTileN + et] +
                                       colpart[3 * kTileN + et];
          atomicAdd(&S[(size_t)(col0 + et) * C + pi.chain_rows], v);
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

}  // namespace

// The walk: 8x8 tiles of (row block, column block) pairs in Morton order, only the tiles that reach the diagonal or lie
// above it (built on the host when the block tables change; 4 bytes per 64 pairs).  Inside a tile the 64 pairs are
// again in Morton order, so any ~74 consecutive slots form a compact 2-D patch and the clusters in flight share operand
// blocks in L2.
static int32_t ensure_walk_f4(asm_ctx* ctx) {
  const int nR = ctx->f4_nR, nC = ctx->f4_nC;
  const std::vector<int2>& blk = ctx->f4_blk_host;
  std::vector<long long> key;
  key.reserve(blk.size() + 2);
  key.push_back(nR);
  key.push_back(nC);
  for (const int2& b : blk) key.push_back(b.x);
  if (key == ctx->f4_walk_key) return ASM_OK;
  const long long ntR = (nR + 7) / 8, ntC = (nC + 7) / 8;
  long long P = 1;
  while (P < std::max(ntR, ntC)) P <<= 1;
  auto compact = [](unsigned long long v) {
    v &= 0x5555555555555555ULL;
    v = (v | (v >> 1)) & 0x3333333333333333ULL;
    v = (v | (v >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    v = (v | (v >> 4)) & 0x00FF00FF00FF00FFULL;
    v = (v | (v >> 8)) & 0x0000FFFF0000FFFFULL;
    v = (v | (v >> 16)) & 0x00000000FFFFFFFFULL;
    return (uint32_t)v;
  };
  std::vector<uint32_t>& tiles = ctx->f4_walk_host;
  tiles.clear();
  for (unsigned long long m = 0; m < (unsigned long long)(P * P); m++) {
    const uint32_t ti = compact(m), tj = compact(m >> 1);
    if ((long long)ti >= ntR || (long long)tj >= ntC) continue;
    // block bases ascend: the tile has work unless its last column block ends before its first row block starts
    const int first_row = blk[(size_t)ti * 8].x;
    const int last_col = blk[(size_t)nR + (size_t)std::min<long long>((long long)tj * 8 + 7, nC - 1)].x + (kTileNF - 1);
    if (last_col >= first_row) tiles.push_back(ti | (tj << 16));
  }
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->f4_walk_dev, sizeof(uint32_t) * std::max<size_t>(tiles.size(), 1))) != ASM_OK) return rc;
  if (!tiles.empty())
    ASM_CUDA(ctx, cudaMemcpyAsync(ctx->f4_walk_dev.p, tiles.data(), sizeof(uint32_t) * tiles.size(), cudaMemcpyHostToDevice,
                                  ctx->stream));
  ctx->f4_walk_key.swap(key);
  return ASM_OK;
}

int32_t rf_f4_sums(asm_ctx* ctx, asm_shard shard) {
  PsrfLayout& L = ctx->layout;
  if (L.col_block != kTileNF) return asm_fail(ctx, ASM_E_INVALID, "rf_f4: layout built for column blocks of %d trees", L.col_block);
  if (ctx->f4_nR >= 65536 * 8 || ctx->f4_nC >= 65536 * 8)
    return asm_fail(ctx, ASM_E_UNSUPPORTED, "rf_f4: %d row blocks exceed the 16-bit tile index of the walk", ctx->f4_nR);
  const uint64_t pitch = (uint64_t)L.pitch4;
  L.bits_used = (int64_t)pitch * 2;
  if (pitch * 2 >= 5792) {  // same 32-bit epilogue sums as the int8 kernel
    int32_t rcg = psrf_check_max_bits(ctx, 5791);
    if (rcg != ASM_OK) return rcg;
  }
  CUtensorMap tm;
  int32_t rc = make_tmap(ctx, &tm, ctx->opnd_f4.p, (uint64_t)L.Tp, pitch);
  if (rc != ASM_OK) return rc;
  if (!(ctx->func_attr_done & kAttrF4x2)) {
    ASM_CUDA(ctx, cudaFuncSetAttribute(rf_f4x2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesF));
    ctx->func_attr_done |= kAttrF4x2;
  }
  if ((rc = ensure_walk_f4(ctx)) != ASM_OK) return rc;
  const long long n_slots = (long long)ctx->f4_walk_host.size() * 64;
  L.tiles_total = n_slots;
  L.tiles_done = (n_slots - shard.shard_index + shard.shard_count - 1) / shard.shard_count;
  // the slot counter sits beside the gather's top word and was zeroed with it (psrf_gather_f4)
  unsigned long long* slot_counter = (unsigned long long*)((char*)ctx->top_word_dev.p + kF4CounterOff);
  // a slot is one of the 64 pairs of a walk tile; roughly half of them carry work
  const int clusters = (int)std::min<long long>(std::max<long long>(L.tiles_done / 2, 1), ctx->sm_count / 2);
  if (clusters <= 0 || n_slots == 0) return ASM_OK;
  const int2* blk = (const int2*)ctx->f4_blk_dev.p;
  L.bits_used = -1;  // known on the device only (gather_f4_kernel's top word): asm_psrf_last_layout reads it back
  {
    LaunchScope ls(ctx, ASM_K_RF_I8);
    rf_f4x2_kernel<<<2 * clusters, kThreadsF, kSmemBytesF, ctx->stream>>>(
        tm, (const int*)ctx->top_word_dev.p, (const int*)ctx->nbits.p, blk, blk + ctx->f4_nR, ctx->f4_nR, ctx->f4_nC, n_slots,
        (const uint32_t*)ctx->f4_walk_dev.p, shard.shard_index, shard.shard_count, slot_counter,
        (unsigned long long*)ctx->S_pad.p, L.C, getenv("ASM_F4_DEBUG") ? atoi(getenv("ASM_F4_DEBUG")) : 0);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}
