// psrf_layout.cu -- operand layout, gather / expand, S compaction and the PSRF finish kernels.
//
// Data layout in HBM for one PSRF evaluation:
//   operand  bits [Tp][Wp] uint64   thinned trees of all chains, chain-major; per chain an optional
//                                   "A" segment (burn-in trees that are rows but not columns) and a
//                                   "B" segment (column trees), each zero-padded to 256 rows, so that
//                                   every 128-row tile belongs to one chain and one class
//   operand  i8   [Tp][pitch8] int8 the same matrix as 0/1 bytes (tcgen05 path; written instead of `bits`)
//   nbits    [Tp] int32             popcount of every row (= number of clades of the tree)
//   S_pad    [Tp][C] int64          sum over the column trees i of chain c of (RF(p,i)+1)^2
//   S        [C][n_rows][C] int64   the rows the caller asked for, zero-padding removed in closed form
#include <algorithm>
#include <cmath>
#include <cstring>

#include "asm_internal.h"
#include "seq_sum.cuh"

namespace {

struct SegTableDev {
  int32_t n_seg;
  int32_t delta;
  int32_t words;   // stored words per source row
  int32_t Wp;      // padded words per operand row
  int32_t block_real;  // live rows per 256-row block (256, or 240 for the fp4 operand)
  int64_t Tp;
  const uint64_t* chain_bits[ASM_MAX_CHAINS];
  Segment seg[kMaxSegments];
};


// padded row p -> (segment, live?, index of the tree inside the segment).  A segment is a run of 256-row blocks
// of which the first `block_real` rows are trees and the rest zero padding.
struct RowRef {
  const Segment* seg;
  bool live;
  int64_t k;
};
__device__ __forceinline__ RowRef locate_row(const SegTableDev* tab, int64_t p) {
  int lo = 0, hi = tab->n_seg - 1;
  while (lo < hi) {  // last segment with dst_off <= p
    int mid = (lo + hi + 1) >> 1;
    if (tab->seg[mid].dst_off <= p) lo = mid; else hi = mid - 1;
  }
  const Segment& s = tab->seg[lo];
  const int64_t j = p - s.dst_off;
  const int64_t o = j & (kBlockRows - 1);
  const int64_t k = (j >> 8) * tab->block_real + o;
  RowRef r;
  r.seg = &s;
  r.k = k;
  r.live = o < tab->block_real && k < s.n;
  return r;
}

// One warp per padded row: copy (zero-extended) or zero-fill, and count the set bits.
__global__ void __launch_bounds__(256) gather_bits_kernel(const __grid_constant__ SegTableDev tab_,
                                                           uint64_t* __restrict__ out, int32_t* __restrict__ nbits) {
  const SegTableDev* tab = &tab_;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t Tp = tab->Tp;
  const int Wp = tab->Wp, words = tab->words, delta = tab->delta;
  for (int64_t p = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < Tp;
       p += (int64_t)gridDim.x * warps_per_block) {
    const RowRef rr = locate_row(tab, p);
    const Segment& s = *rr.seg;
    const bool live = rr.live;
    const uint64_t* src = live ? tab->chain_bits[s.chain] + (size_t)(s.src_first + rr.k * delta) * words : nullptr;
    int cnt = 0;
    for (int w = lane; w < Wp; w += 32) {
      uint64_t v = (live && w < words) ? __ldg(src + w) : 0ULL;
      out[(size_t)p * Wp + w] = v;
      cnt += __popcll(v);
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) nbits[p] = cnt;
  }
}


// int8 path: the same gather, written straight as 0/1 bytes (one 64-bit word -> 64 bytes, four 16-byte
// stores per lane, a warp writes 2 KB contiguous); the bit matrix itself is never materialised.
__global__ void __launch_bounds__(256) gather_i8_kernel(const __grid_constant__ SegTableDev tab_, uint4* __restrict__ out,
                                                         int W8, int32_t* __restrict__ nbits) {
  const SegTableDev* tab = &tab_;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t Tp = tab->Tp;
  const int words = tab->words, delta = tab->delta;
  for (int64_t p = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < Tp;
       p += (int64_t)gridDim.x * warps_per_block) {
    const RowRef rr = locate_row(tab, p);
    const Segment& s = *rr.seg;
    const bool live = rr.live;
    const uint64_t* src = live ? tab->chain_bits[s.chain] + (size_t)(s.src_first + rr.k * delta) * words : nullptr;
    int cnt = 0;
    for (int w = lane; w < W8; w += 32) {
      const uint64_t v = (live && w < words) ? __ldg(src + w) : 0ULL;
      cnt += __popcll(v);
      uint4 q[4];
      uint32_t* o = reinterpret_cast<uint32_t*>(q);
#pragma unroll
      for (int k = 0; k < 16; k++) o[k] = (((uint32_t)(v >> (4 * k)) & 0xFu) * 0x00204081u) & 0x01010101u;
#pragma unroll
      for (int k = 0; k < 4; k++) out[((size_t)p * W8 + w) * 4 + k] = q[k];
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) nbits[p] = cnt;
  }
}

// fp4 path: the same gather, written as e2m1 nibbles (clade k -> nibble k of the row, 1.0 = 0x2, 0.0 = 0x0; the
// even clade in the low nibble).  One source byte becomes one 32-bit word by spreading its bits to the nibbles with
// three shift-or-mask steps (a 256-entry table in shared memory was bound by its bank conflicts: ncu showed the kernel
// at 10 % of DRAM and 39 % of L2 throughput); one 64-bit word -> 32 bytes, two 16-byte stores per lane.
__device__ __forceinline__ uint32_t byte_to_e2m1(uint32_t b) {
  uint32_t x = b & 0xFFu;
  x = (x | (x << 12)) & 0x000F000Fu;
  x = (x | (x << 6)) & 0x03030303u;
  x = (x | (x << 3)) & 0x11111111u;
  return x << 1;  // 1.0 in e2m1 is 0x2
}
__global__ void __launch_bounds__(256) gather_f4_kernel(const __grid_constant__ SegTableDev tab_, uint4* __restrict__ out,
                                                         int W4, int32_t* __restrict__ nbits, int* __restrict__ top_word) {
  __shared__ int top_s;
  if (threadIdx.x == 0) top_s = 0;
  __syncthreads();
  const SegTableDev* tab = &tab_;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t Tp = tab->Tp;
  const int words = tab->words, delta = tab->delta;
  int top = 0;  // highest non-zero word this lane has seen: the K extent of the contraction (rf_f4.cu)
  for (int64_t p = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < Tp;
       p += (int64_t)gridDim.x * warps_per_block) {
    const RowRef rr = locate_row(tab, p);
    const Segment& s = *rr.seg;
    const bool live = rr.live;
    const uint64_t* src = live ? tab->chain_bits[s.chain] + (size_t)(s.src_first + rr.k * delta) * words : nullptr;
    // a lane takes 32 clades at a time: a coalesced 4-byte read, one 16-byte store, consecutive lanes write
    // consecutive 16 bytes (two 16-byte stores per lane, 32 bytes apart from lane to lane, left every 32-byte sector
    // half written by each store instruction)
    const uint32_t* src32 = reinterpret_cast<const uint32_t*>(src);
    int cnt = 0;
    for (int h = lane; h < 2 * W4; h += 32) {
      const uint32_t v = (live && h < 2 * words) ? __ldg(src32 + h) : 0u;
      cnt += __popc(v);
      if (v) top = max(top, h >> 1);
      out[(size_t)p * W4 * 2 + h] = make_uint4(byte_to_e2m1(v), byte_to_e2m1(v >> 8), byte_to_e2m1(v >> 16), byte_to_e2m1(v >> 24));
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) nbits[p] = cnt;
  }
  top = __reduce_max_sync(0xffffffffu, top);
  if (lane == 0 && top > 0) atomicMax(&top_s, top);
  __syncthreads();
  if (threadIdx.x == 0 && top_s > 0) atomicMax(top_word, top_s);  // one global atomic per block
}

__global__ void __launch_bounds__(256) max_bits_kernel(const int32_t* __restrict__ nbits, long long n, int* __restrict__ out) {
  int m = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    m = max(m, nbits[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

struct CompactParams {
  int32_t C, n_rows, subtract_pad;
  int64_t segA_off[ASM_MAX_CHAINS], segB_off[ASM_MAX_CHAINS], rowA0[ASM_MAX_CHAINS], rowB0[ASM_MAX_CHAINS];
  int64_t padC[ASM_MAX_CHAINS], padR[ASM_MAX_CHAINS];
  int32_t rowsA[ASM_MAX_CHAINS];
};

__global__ void __launch_bounds__(256) compact_kernel(const __grid_constant__ CompactParams prm_,
                                                       const long long* __restrict__ S_pad,
                                                       const int32_t* __restrict__ nbits, long long* __restrict__ out) {
  const CompactParams* prm = &prm_;
  const int C = prm->C, n_rows = prm->n_rows;
  const int64_t total = (int64_t)C * n_rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t ar = i / C;
    int r = (int)(ar % n_rows);
    int a = (int)(ar / n_rows);
    const bool inA = r < prm->rowsA[a];
    const int64_t k = inA ? prm->rowA0[a] + r : prm->rowB0[a] + (r - prm->rowsA[a]);
    const int64_t p = (inA ? prm->segA_off[a] : prm->segB_off[a]) + k;
    long long v = S_pad[p * C + c];
    if (prm->subtract_pad) {
      // Every pair of (padded) rows is evaluated once, as (row = the earlier one, column = the later one).  The zero
      // rows behind chain c's column trees come after this tree when a <= c (they were columns of its tiles: padC of
      // them) and before it when a > c (they were rows, and reached it through the column sums: padR of them).
      long long e = (long long)nbits[p] + 1;  // RF(tree, empty row) + 1
      v -= (a <= c ? prm->padC[c] : prm->padR[c]) * e * e;
    }
    out[i] = v;
  }
}

// Row values (TreePSRF.java:287-292) of every ordered chain pair, one thread per row (a, r) of the table:
// vals[a*C+b][r] for the sums below and, when the caller asked for the rows, psrf_rows[a][r][b].
__global__ void __launch_bounds__(256) psrf_rows_kernel(const long long* __restrict__ S, int C, int n_rows,
                                                         double* __restrict__ vals, double* __restrict__ psrf_rows) {
  const long long total = (long long)C * n_rows;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int a = (int)(i / n_rows), r = (int)(i % n_rows);
    const long long* row = S + (size_t)i * C;
    const double varIn = (double)row[a];
#pragma unroll 4
    for (int b = 0; b < C; b++) {
      const double varBetween = (double)row[b];
      double psrf;
      if (varIn != 0) {
        psrf = sqrt(varBetween / varIn);
      } else {
        psrf = sqrt(varBetween);
      }
      vals[((size_t)a * C + b) * n_rows + r] = psrf;
      if (psrf_rows) psrf_rows[(size_t)i * C + b] = psrf;
    }
  }
}

// One block per ordered chain pair (a, b).  The mean is the sum of the row values in ascending row order, one
// rounding per add (TreePSRF.mean, TreePSRF.java:264-271), divided by the row count.  The order is kept and the
// chain of dependent adds is not: seq_sum.cuh evaluates the same roundings as an integer scan over the block (plain
// DADDs only where the running sum changes binade).
__global__ void __launch_bounds__(seqsum::kThreads) psrf_mean_kernel(const double* __restrict__ vals, int n_rows,
                                                                      double* __restrict__ psrf_mean, int planned) {
  __shared__ seqsum::Shared sh;
  const double* v = vals + (size_t)blockIdx.x * n_rows;
  const double sum = seqsum::block_sum(sh, n_rows, [&](int k) { return v[k]; }, planned != 0);
  if (threadIdx.x == 0) psrf_mean[blockIdx.x] = sum / (double)n_rows;  // 0/0 = NaN for no rows, as in Java
}

// The same walk over doubles in memory (asm_seq_sum).
__global__ void __launch_bounds__(seqsum::kThreads) seq_sum_kernel(const double* __restrict__ v, int n, double* __restrict__ out,
                                                                    int planned) {
  __shared__ seqsum::Shared sh;
  const double sum = seqsum::block_sum(sh, n, [&](int k) { return v[k]; }, planned != 0);
  if (threadIdx.x == 0) *out = sum;
}

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// ASM_SEQSUM_PLANNED=0: the binade-by-binade walk alone (tuning / test knob; same bits)
inline int seq_sum_planned() {
  const char* e = getenv("ASM_SEQSUM_PLANNED");
  return !(e && e[0] == '0');
}

}  // namespace

int32_t psrf_build_layout(asm_ctx* ctx, const int32_t* burnin, int32_t row_start, int32_t end, int32_t delta,
                          int32_t col_block) {
  PsrfLayout& L = ctx->layout;
  const int C = ctx->n_chains;
  if (delta < 1) return asm_fail(ctx, ASM_E_INVALID, "psrf: delta must be >= 1, not %d", delta);
  if (row_start < 0 || row_start % delta != 0)
    return asm_fail(ctx, ASM_E_INVALID, "psrf: row_start %d must be a non-negative multiple of delta %d", row_start, delta);
  if (end < 0) return asm_fail(ctx, ASM_E_INVALID, "psrf: negative end");
  for (int c = 0; c < C; c++) {
    if (burnin[c] < 0) return asm_fail(ctx, ASM_E_RANGE, "psrf: negative burn-in for chain %d", c);
    if (end > ctx->trees[c].n)
      return asm_fail(ctx, ASM_E_RANGE, "psrf: end %d past chain %d's %lld trees", end, c, (long long)ctx->trees[c].n);
  }
  L = PsrfLayout();
  L.C = C;
  L.delta = delta;
  L.row_start = row_start;
  L.end = end;
  L.n_rows = end > row_start ? (end - row_start + delta - 1) / delta : 0;
  L.Wp = (int32_t)round_up(ctx->words > 0 ? ctx->words : 1, kBitsPad / 64);
  L.pitch8 = (int32_t)round_up((int64_t)(ctx->words > 0 ? ctx->words : 1) * 64, 128);  // int8 operand row pitch: whole 128-byte K blocks
  L.pitch4 = (int32_t)round_up((int64_t)(ctx->words > 0 ? ctx->words : 1) * 32, 128);  // fp4: two clades per byte
  L.block_real = kBlockRows;
  L.col_block = col_block;
  // rows a segment occupies: whole 256-row blocks, and (fp4) room for the last 240-tree column block, whose second
  // 128-row TMA box starts at row 120 of the block
  auto padded = [&](int64_t n) {
    int64_t ext = round_up(n, kBlockRows);
    if (col_block != kBlockRows) ext = std::max(ext, round_up(round_up(n, col_block) + (kTileRows - col_block / 2), kBlockRows));
    return ext;
  };
  if (L.n_rows == 0) return ASM_OK;
  int64_t off = 0;
  for (int c = 0; c < C; c++) {
    int64_t q = round_up(burnin[c], delta);         // TreePSRF.java:296-301
    int64_t u = std::min<int64_t>(row_start, q);
    int64_t m = (end - u + delta - 1) / delta;       // thinned trees u, u+delta, ... < end
    int64_t kq = std::min<int64_t>((q - u) / delta, m);
    int64_t kr = (row_start - u) / delta;
    int64_t nA = kq, nB = m - kq;
    int64_t offA = off;
    if (nA > 0) {
      L.seg[L.n_seg++] = Segment{off, nA, u, c, 0};
      off += padded(nA);
    }
    int64_t offB = off;
    if (nB > 0) {
      L.seg[L.n_seg++] = Segment{off, nB, u + kq * delta, c, 1};
      off += padded(nB);
      L.padC[c] = round_up(nB, col_block) - nB;   // zero rows among the rows the kernels use as columns ...
      L.padR[c] = round_up(nB, kBlockRows) - nB;  // ... and as rows
    } else {
      L.padC[c] = L.padR[c] = 0;
    }
    L.rowsA[c] = (int32_t)std::max<int64_t>(0, kq - kr);
    L.segA_off[c] = offA;
    L.segB_off[c] = offB;
    L.rowA0[c] = kr;
    L.rowB0[c] = std::max<int64_t>(kr - kq, 0);
  }
  L.Tp = off;
  return ASM_OK;
}

// Seg table, per-tile metadata and the zeroed S table: common to both distance kernels.
// (`tile_meta`: the int8 / popcount kernels' per-tile table; the fp4 kernel has block tables of its own.)
static int32_t psrf_prepare(asm_ctx* ctx, SegTableDev& tab, bool tile_meta = true) {
  PsrfLayout& L = ctx->layout;
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->nbits, sizeof(int32_t) * (size_t)L.Tp)) != ASM_OK) return rc;
  const int64_t n_tiles = L.Tp / kTileRows;
  tab.n_seg = L.n_seg;
  tab.delta = L.delta;
  tab.words = ctx->words;
  tab.Wp = L.Wp;
  tab.block_real = L.block_real;
  tab.Tp = L.Tp;
  for (int c = 0; c < ctx->n_chains; c++) tab.chain_bits[c] = ctx->trees[c].d_bits;
  for (int s = 0; s < L.n_seg; s++) tab.seg[s] = L.seg[s];
  if ((rc = asm_reserve(ctx, ctx->S_pad, sizeof(int64_t) * (size_t)L.Tp * L.C)) != ASM_OK) return rc;
  // S_pad starts at zero; the distance kernels accumulate into it with integer atomics
  ASM_CUDA(ctx, cudaMemsetAsync(ctx->S_pad.p, 0, sizeof(int64_t) * (size_t)L.Tp * L.C, ctx->stream));
  if (!tile_meta) return ASM_OK;

  // per-tile metadata {chain, flags}
  if ((rc = asm_reserve(ctx, ctx->blk_meta, sizeof(int2) * (size_t)n_tiles)) != ASM_OK) return rc;
  std::vector<int2>& meta = ctx->meta_host;
  meta.assign((size_t)n_tiles, make_int2(0, 0));
  for (int s = 0; s < L.n_seg; s++) {
    const Segment& g = L.seg[s];
    const int64_t n_tiles_seg = (g.n + L.block_real - 1) / L.block_real * (kBlockRows / kTileRows);
    // first tree (index within the segment) that is an output row
    int64_t first_row;
    if (!g.is_col) {
      first_row = (L.rowsA[g.chain] > 0) ? L.rowA0[g.chain] : g.n;
    } else {
      first_row = L.rowB0[g.chain];
      if (L.n_rows - L.rowsA[g.chain] <= 0) first_row = g.n;
    }
    for (int64_t t = 0; t < n_tiles_seg; t++) {
      // trees of this 128-row tile: block t/2, rows [0,128) or [128, block_real) of it
      const int64_t base = (t / 2) * L.block_real;
      const int64_t lo = base + (t & 1) * kTileRows;
      const int64_t hi = std::min<int64_t>(base + std::min<int64_t>(L.block_real, (t & 1) * kTileRows + kTileRows), g.n);
      uint32_t f = g.is_col ? kFlagCol : 0;
      if (hi > lo && hi > first_row) f |= kFlagRows;
      meta[(size_t)(g.dst_off / kTileRows + t)] = make_int2(g.chain, (int)f);
    }
  }
  // pageable source: the runtime stages the bytes before returning, so no host sync is needed here
  ASM_CUDA(ctx, cudaMemcpyAsync(ctx->blk_meta.p, meta.data(), sizeof(int2) * (size_t)n_tiles, cudaMemcpyHostToDevice,
                                ctx->stream));
  return ASM_OK;
}

int32_t psrf_gather_bits(asm_ctx* ctx) {
  PsrfLayout& L = ctx->layout;
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->opnd_bits, sizeof(uint64_t) * (size_t)L.Tp * L.Wp)) != ASM_OK) return rc;
  SegTableDev tab;
  if ((rc = psrf_prepare(ctx, tab)) != ASM_OK) return rc;
  {
    LaunchScope ls(ctx, ASM_K_GATHER);
    int blocks = (int)std::min<int64_t>((L.Tp + 7) / 8, (int64_t)ctx->sm_count * 16);
    gather_bits_kernel<<<blocks, 256, 0, ctx->stream>>>(tab, (uint64_t*)ctx->opnd_bits.p, (int32_t*)ctx->nbits.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int32_t psrf_gather_i8(asm_ctx* ctx) {
  PsrfLayout& L = ctx->layout;
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->opnd_i8, (size_t)L.Tp * L.pitch8)) != ASM_OK) return rc;
  SegTableDev tab;
  if ((rc = psrf_prepare(ctx, tab)) != ASM_OK) return rc;
  {
    LaunchScope ls(ctx, ASM_K_GATHER);
    int blocks = (int)std::min<int64_t>((L.Tp + 7) / 8, (int64_t)ctx->sm_count * 16);
    gather_i8_kernel<<<blocks, 256, 0, ctx->stream>>>(tab, (uint4*)ctx->opnd_i8.p, L.pitch8 / 64, (int32_t*)ctx->nbits.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

// Block tables of the fp4 kernel: the dense rows of every segment cut into row blocks of 256 trees (operand A of a
// tile, one half per CTA of the pair) and, independently, into column blocks of col_block = 240 trees (operand B:
// TMEM has room for two 240-column accumulators beside the scale factors).  Entry = {first padded row,
// chain | flags << 8}; flags: kFlagCol (the segment's trees are columns of the statistic), bit 1 / bit 2 = the first /
// second 128 rows of a row block hold output rows, bit 1 = a column block holds output rows.
static int32_t psrf_build_f4_blocks(asm_ctx* ctx) {
  PsrfLayout& L = ctx->layout;
  std::vector<int2>& blk = ctx->f4_blk_host;
  blk.clear();
  std::vector<int2> cols;
  for (int s = 0; s < L.n_seg; s++) {
    const Segment& g = L.seg[s];
    int64_t first_row;  // first tree (index within the segment) that is an output row
    if (!g.is_col) {
      first_row = (L.rowsA[g.chain] > 0) ? L.rowA0[g.chain] : g.n;
    } else {
      first_row = L.rowB0[g.chain];
      if (L.n_rows - L.rowsA[g.chain] <= 0) first_row = g.n;
    }
    auto has_rows = [&](int64_t lo, int64_t hi) { return std::min<int64_t>(hi, g.n) > std::max<int64_t>(lo, first_row); };
    const uint32_t base_flags = g.is_col ? kFlagCol : 0;
    for (int64_t k = 0; k < g.n; k += kBlockRows) {
      uint32_t f = base_flags;
      if (has_rows(k, k + kTileRows)) f |= 2;
      if (has_rows(k + kTileRows, k + kBlockRows)) f |= 4;
      blk.push_back(make_int2((int)(g.dst_off + k), (int)((uint32_t)g.chain | (f << 8))));
    }
    for (int64_t k = 0; k < g.n; k += L.col_block) {
      uint32_t f = base_flags;
      if (has_rows(k, k + L.col_block)) f |= 2;
      cols.push_back(make_int2((int)(g.dst_off + k), (int)((uint32_t)g.chain | (f << 8))));
    }
  }
  ctx->f4_nR = (int32_t)blk.size();
  ctx->f4_nC = (int32_t)cols.size();
  blk.insert(blk.end(), cols.begin(), cols.end());
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->f4_blk_dev, sizeof(int2) * std::max<size_t>(blk.size(), 1))) != ASM_OK) return rc;
  if (!blk.empty()) {
    // through the context's page-locked staging: a pageable source is copied by the driver before the call returns
    // (microseconds of host time at the head of every step); the event keeps a later call from overwriting the
    // staging while this copy is still queued (the *_async entry points do not synchronise)
    const size_t bytes = sizeof(int2) * blk.size();
    if (ctx->h_stage_bytes < bytes) {
      if (ctx->h_stage) {
        ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ASM_CUDA(ctx, cudaFreeHost(ctx->h_stage));
        ctx->h_stage = nullptr;
        ctx->h_stage_bytes = 0;
      }
      const size_t want = std::max<size_t>(bytes * 2, 1 << 16);
      ASM_CUDA(ctx, cudaMallocHost(&ctx->h_stage, want));
      ctx->h_stage_bytes = want;
    } else {
      ASM_CUDA(ctx, cudaEventSynchronize(ctx->ev_stage));
    }
    memcpy(ctx->h_stage, blk.data(), bytes);
    ASM_CUDA(ctx, cudaMemcpyAsync(ctx->f4_blk_dev.p, ctx->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ASM_CUDA(ctx, cudaEventRecord(ctx->ev_stage, ctx->stream));
  }
  return ASM_OK;
}

int32_t psrf_gather_f4(asm_ctx* ctx) {
  PsrfLayout& L = ctx->layout;
  int32_t rc;
  if (L.Tp >= (int64_t)1 << 31) return asm_fail(ctx, ASM_E_UNSUPPORTED, "psrf: %lld padded rows exceed the fp4 kernel's 32-bit row index", (long long)L.Tp);
  if ((rc = asm_reserve(ctx, ctx->opnd_f4, (size_t)L.Tp * L.pitch4)) != ASM_OK) return rc;
  SegTableDev tab;
  if ((rc = psrf_prepare(ctx, tab, /*tile_meta=*/false)) != ASM_OK) return rc;
  if ((rc = psrf_build_f4_blocks(ctx)) != ASM_OK) return rc;
  // the two scalars of the fp4 path share 16 bytes and one memset: [0] the gather's top word, [8] the contraction's
  // work-slot counter (kF4CounterOff)
  if ((rc = asm_reserve(ctx, ctx->top_word_dev, 16)) != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemsetAsync(ctx->top_word_dev.p, 0, 16, ctx->stream));
  {
    LaunchScope ls(ctx, ASM_K_GATHER);
    int blocks = (int)std::min<int64_t>((L.Tp + 7) / 8, (int64_t)ctx->sm_count * 16);
    gather_f4_kernel<<<blocks, 256, 0, ctx->stream>>>(tab, (uint4*)ctx->opnd_f4.p, L.pitch4 / 32, (int32_t*)ctx->nbits.p,
                                                     (int*)ctx->top_word_dev.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int32_t psrf_check_max_bits(asm_ctx* ctx, int32_t limit) {
  PsrfLayout& L = ctx->layout;
  int32_t rc = asm_reserve(ctx, ctx->misc, sizeof(int));
  if (rc != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemsetAsync(ctx->misc.p, 0, sizeof(int), ctx->stream));
  {
    LaunchScope ls(ctx, ASM_K_MISC);
    max_bits_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>((const int32_t*)ctx->nbits.p, L.Tp, (int*)ctx->misc.p);
  }
  int m = 0;
  ASM_CUDA(ctx, cudaMemcpyAsync(&m, ctx->misc.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (m > limit)
    return asm_fail(ctx, ASM_E_UNSUPPORTED, "ASM_RF_I8 supports trees with at most %d clades (found %d); use ASM_RF_POPC", limit, m);
  return ASM_OK;
}

int32_t psrf_compact(asm_ctx* ctx, asm_shard shard, int64_t* d_S_out) {
  PsrfLayout& L = ctx->layout;
  CompactParams prm;
  prm.C = L.C;
  prm.n_rows = L.n_rows;
  prm.subtract_pad = shard.shard_index == 0 ? 1 : 0;  // the closed-form pad term is removed exactly once
  for (int c = 0; c < L.C; c++) {
    prm.segA_off[c] = L.segA_off[c];
    prm.segB_off[c] = L.segB_off[c];
    prm.rowA0[c] = L.rowA0[c];
    prm.rowB0[c] = L.rowB0[c];
    prm.padC[c] = L.padC[c];
    prm.padR[c] = L.padR[c];
    prm.rowsA[c] = L.rowsA[c];
  }
  {
    LaunchScope ls(ctx, ASM_K_PSRF_FINISH);
    int64_t total = (int64_t)L.C * L.n_rows * L.C;
    int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16);
    compact_kernel<<<blocks, 256, 0, ctx->stream>>>(prm, (const long long*)ctx->S_pad.p,
                                                    (const int32_t*)ctx->nbits.p, (long long*)d_S_out);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int32_t psrf_finish(asm_ctx* ctx, const int64_t* d_S, int32_t n_rows, double* psrf_rows_host, double* psrf_mean_host) {
  const int C = ctx->n_chains;
  int32_t rc;
  if (!psrf_mean_host) return asm_fail(ctx, ASM_E_INVALID, "psrf_finish: null output");
  const size_t rows_bytes = sizeof(double) * (size_t)C * (size_t)n_rows * (size_t)C;
  if ((rc = asm_reserve(ctx, ctx->psrf_mean, sizeof(double) * (size_t)C * C)) != ASM_OK) return rc;
  double* d_rows = nullptr;
  if (psrf_rows_host && rows_bytes) {
    if ((rc = asm_reserve(ctx, ctx->psrf_rows, rows_bytes)) != ASM_OK) return rc;
    d_rows = (double*)ctx->psrf_rows.p;
  }
  const size_t vals_bytes = sizeof(double) * (size_t)C * (size_t)C * (size_t)n_rows;
  if ((rc = asm_reserve(ctx, ctx->psrf_vals, vals_bytes ? vals_bytes : 8)) != ASM_OK) return rc;
  if (n_rows > 0) {
    LaunchScope ls(ctx, ASM_K_PSRF_FINISH);
    const int64_t total = (int64_t)C * n_rows;
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 8);
    psrf_rows_kernel<<<blocks, 256, 0, ctx->stream>>>((const long long*)d_S, C, n_rows, (double*)ctx->psrf_vals.p, d_rows);
  }
  {
    LaunchScope ls(ctx, ASM_K_PSRF_FINISH);
    psrf_mean_kernel<<<C * C, seqsum::kThreads, 0, ctx->stream>>>((const double*)ctx->psrf_vals.p, n_rows, (double*)ctx->psrf_mean.p,
                                                                seq_sum_planned());
  }
  ASM_CUDA(ctx, cudaGetLastError());
  ASM_CUDA(ctx, cudaMemcpyAsync(psrf_mean_host, ctx->psrf_mean.p, sizeof(double) * (size_t)C * C, cudaMemcpyDeviceToHost,
                                ctx->stream));
  if (d_rows) ASM_CUDA(ctx, cudaMemcpyAsync(psrf_rows_host, d_rows, rows_bytes, cudaMemcpyDeviceToHost, ctx->stream));
  // an ids upload that named a clade outside its dictionary fails the evaluation: its flag comes back with the results
  if (ctx->ids_check_pending) return trees_ids_verify(ctx);  // synchronises
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t seq_sum(asm_ctx* ctx, const double* values_host, int64_t n, double* sum_host) {
  if (n < 0 || n > INT32_MAX || (n > 0 && !values_host) || !sum_host) return asm_fail(ctx, ASM_E_INVALID, "seq_sum: bad argument");
  int32_t rc = asm_reserve(ctx, ctx->misc, sizeof(double) * ((size_t)n + 1));
  if (rc != ASM_OK) return rc;
  double* d = (double*)ctx->misc.p;
  if (n) ASM_CUDA(ctx, cudaMemcpyAsync(d + 1, values_host, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
  {
    LaunchScope ls(ctx, ASM_K_MISC);
    seq_sum_kernel<<<1, seqsum::kThreads, 0, ctx->stream>>>(d + 1, (int)n, d, seq_sum_planned());
  }
  ASM_CUDA(ctx, cudaGetLastError());
  ASM_CUDA(ctx, cudaMemcpyAsync(sum_host, d, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}
