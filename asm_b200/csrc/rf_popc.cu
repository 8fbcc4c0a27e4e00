// rf_popc.cu -- K1: Robinson-Foulds distances as XOR + POPC over clade bit rows, with the
// (d+1)^2 within-/between-chain reductions of calcPSRF (TreePSRF.java:274-286) fused in.
//
// Work decomposition: the padded operand matrix bits[Tp][Wp] is cut into 128-row tiles; the kernel
// walks the upper triangle of the tile-pair matrix (i <= j).  An off-diagonal pair contributes twice,
// as row sums (rows of i over the columns of j) and as column sums (rows of j over the columns of
// i), which is the symmetry RF(a,b) = RF(b,a); a diagonal pair is evaluated in full and contributes
// row sums only.  Persistent CTAs pull pair indices from a global counter.
//
// One CTA = 256 threads = a 16x16 thread grid, each thread an 8x8 micro-tile of distances held in
// registers.  K (the clade dimension) is streamed in chunks of 16 words (128 bytes per row) through a
// 3-stage shared-memory ring filled with cp.async.bulk (the TMA engine's 1-D bulk copy; UBLKCP in
// SASS) and tracked with mbarrier transaction counts.  Rows are padded to 144 bytes in shared memory
// so that the 16-byte row reads of a quarter warp fall into 8 different bank groups.
//
// Bound: integer pipes (POPC issue first, LOP3 second), not HBM: a tile pair reads 2*128*Wp*8 bytes and does
// 128*128*Wp 64-bit XOR-popcounts.
#include <algorithm>
#include <cmath>

#include "asm_internal.h"

namespace {

constexpr int kThreads = 256;
constexpr int kKC = 16;                 // words per K chunk
constexpr int kRowStride = kKC + 2;     // padded smem row, in words (144 bytes)
constexpr int kStages = 3;
constexpr int kStageWords = 2 * kTileRows * kRowStride;  // A tile + B tile
constexpr size_t kSmemBytes = sizeof(uint64_t) * kStages * kStageWords + 64 /*barriers*/ + sizeof(unsigned long long) * 8 * kTileRows;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// tile-pair index t -> (i, j), i <= j, column-major over the upper triangle: t = j(j+1)/2 + i
__device__ __forceinline__ void pair_from_index(long long t, int& i, int& j) {
  long long jj = (long long)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while (jj * (jj + 1) / 2 > t) jj--;
  while ((jj + 1) * (jj + 2) / 2 <= t) jj++;
  j = (int)jj;
  i = (int)(t - jj * (jj + 1) / 2);
}

__global__ void __launch_bounds__(kThreads, 1)
rf_popc_psrf_kernel(const uint64_t* __restrict__ bits, int Wp, const int2* __restrict__ tile_meta, long long n_pairs,
                    int shard_index, int shard_count, unsigned long long* __restrict__ pair_counter,
                    unsigned long long* __restrict__ S, int C) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* ring = reinterpret_cast<uint64_t*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + sizeof(uint64_t) * kStages * kStageWords);
  unsigned long long* col_part = reinterpret_cast<unsigned long long*>(smem_raw + sizeof(uint64_t) * kStages * kStageWords + 64);
  __shared__ long long s_pair;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int warp = tid >> 5;
  const int nk = Wp / kKC;

  if (tid == 0) {
    for (int s = 0; s < kStages; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  uint32_t phase_bits = 0;  // one parity bit per stage
  int stage_head = 0;       // stage of the next chunk to consume (global over all pairs)

  for (;;) {
    if (tid == 0) s_pair = (long long)atomicAdd(pair_counter, 1ULL);
    __syncthreads();
    const long long local = s_pair;
    const long long t = local * shard_count + shard_index;
    if (t >= n_pairs) break;
    int ti, tj;
    pair_from_index(t, ti, tj);
    const int2 mi = tile_meta[ti], mj = tile_meta[tj];
    const bool diag = ti == tj;
    const bool do_row = (mi.y & kFlagRows) && (mj.y & kFlagCol);
    const bool do_col = !diag && (mj.y & kFlagRows) && (mi.y & kFlagCol);
    if (!do_row && !do_col) {
      __syncthreads();
      continue;
    }
    const uint64_t* Ag = bits + (size_t)ti * kTileRows * Wp;
    const uint64_t* Bg = bits + (size_t)tj * kTileRows * Wp;

    // every thread copies one 128-byte row segment per chunk: threads 0..127 the A rows, 128..255 the B rows
    const uint64_t* my_src = (tid < kTileRows ? Ag + (size_t)tid * Wp : Bg + (size_t)(tid - kTileRows) * Wp);
    const int my_dst = tid * kRowStride;
    auto issue = [&](int kc, int stage) {
      if (tid == 0) mbar_expect_tx(&full[stage], 2 * kTileRows * kKC * 8);
      __syncthreads();  // expect_tx is posted before any complete_tx can arrive
      bulk_g2s(ring + (size_t)stage * kStageWords + my_dst, my_src + (size_t)kc * kKC, kKC * 8, &full[stage]);
    };

    const int pre = min(kStages, nk);
    for (int c = 0; c < pre; c++) issue(c, (stage_head + c) % kStages);

    int acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
      for (int b = 0; b < 8; b++) acc[a][b] = 0;

    for (int kc = 0; kc < nk; kc++) {
      const int stage = (stage_head + kc) % kStages;
      mbar_wait(&full[stage], (phase_bits >> stage) & 1u);
      phase_bits ^= 1u << stage;
      const uint64_t* As = ring + (size_t)stage * kStageWords;
      const uint64_t* Bs = As + kTileRows * kRowStride;
#pragma unroll 1
      for (int kk = 0; kk < kKC; kk += 2) {
        ulonglong2 a[8], b[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
          a[r] = *reinterpret_cast<const ulonglong2*>(As + (ty + 16 * r) * kRowStride + kk);
          b[r] = *reinterpret_cast<const ulonglong2*>(Bs + (tx + 16 * r) * kRowStride + kk);
        }
        // 3:2 carry-save step: the four 32-bit XOR words w0..w3 of a 128-bit slice need three POPCs instead of four
        // (popc(w0)+popc(w1)+popc(w2) = popc(w0^w1^w2) + 2*popc(maj(w0,w1,w2))); POPC is the binding pipe (16/clk/SM),
        // the extra LOP3s go to the ALU pipe (64/clk/SM).
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int c = 0; c < 8; c++) {
            const unsigned long long x = a[r].x ^ b[c].x, y = a[r].y ^ b[c].y;
            const uint32_t w0 = (uint32_t)x, w1 = (uint32_t)(x >> 32), w2 = (uint32_t)y, w3 = (uint32_t)(y >> 32);
            const uint32_t sum = w0 ^ w1 ^ w2;
            const uint32_t carry = (w0 & w1) | (w2 & (w0 ^ w1));
            acc[r][c] += __popc(sum) + __popc(w3) + 2 * __popc(carry);
          }
      }
      __syncthreads();  // all reads of this stage are done before it is refilled
      if (kc + kStages < nk) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        issue(kc + kStages, stage);
      }
    }
    stage_head = (stage_head + nk) % kStages;

    // ---- fused epilogue: (d+1)^2, row sums over tx, column sums over ty ----------------------------
    unsigned long long rs[8], cs[8];
#pragma unroll
    for (int r = 0; r < 8; r++) rs[r] = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) cs[c] = 0;
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
      for (int c = 0; c < 8; c++) {
        unsigned long long e = (unsigned long long)(acc[r][c] + 1);
        unsigned long long v = e * e;
        rs[r] += v;
        cs[c] += v;
      }
    if (do_row) {
#pragma unroll
      for (int r = 0; r < 8; r++) {
        unsigned long long v = rs[r];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);  // over the 16 tx lanes
        if (tx == 0) atomicAdd(&S[((size_t)ti * kTileRows + ty + 16 * r) * C + mj.x], v);
      }
    }
    if (do_col) {
#pragma unroll
      for (int c = 0; c < 8; c++) {
        unsigned long long v = cs[c];
        v += __shfl_xor_sync(0xffffffffu, v, 16);  // the two ty values of this warp
        if ((tid & 16) == 0) col_part[warp * kTileRows + tx + 16 * c] = v;
      }
      __syncthreads();
      if (tid < kTileRows) {
        unsigned long long v = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) v += col_part[w * kTileRows + tid];
        atomicAdd(&S[((size_t)tj * kTileRows + tid) * C + mi.x], v);
      }
    }
    __syncthreads();
  }
}

// Parity probe: one thread per (row, column) pair straight from the chain buffers.
__global__ void __launch_bounds__(256) rf_block_kernel(const uint64_t* __restrict__ A, const uint64_t* __restrict__ B,
                                                        int words, long long nr, long long nc,
                                                        uint16_t* __restrict__ out) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; idx < nr * nc; idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx / nc, c = idx - r * nc;
    const uint64_t* a = A + (size_t)r * words;
    const uint64_t* b = B + (size_t)c * words;
    int d = 0;
    for (int w = 0; w < words; w++) d += __popcll(a[w] ^ b[w]);
    out[idx] = (uint16_t)d;
  }
}

}  // namespace

int32_t rf_popc_sums(asm_ctx* ctx, asm_shard shard) {
  PsrfLayout& L = ctx->layout;
  const long long n_tiles = L.Tp / kTileRows;
  const long long n_pairs = n_tiles * (n_tiles + 1) / 2;
  L.tiles_total = n_pairs;
  L.bits_used = (int64_t)L.Wp * 64;
  L.tiles_done = (n_pairs - shard.shard_index + shard.shard_count - 1) / shard.shard_count;
  int32_t rc = asm_reserve(ctx, ctx->tile_counter, sizeof(unsigned long long));
  if (rc != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemsetAsync(ctx->tile_counter.p, 0, sizeof(unsigned long long), ctx->stream));
  if (!(ctx->func_attr_done & kAttrPopc)) {
    ASM_CUDA(ctx, cudaFuncSetAttribute(rf_popc_psrf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    ctx->func_attr_done |= kAttrPopc;
  }
  const int grid = (int)std::min<long long>(L.tiles_done, ctx->sm_count);
  if (grid <= 0) return ASM_OK;
  {
    LaunchScope ls(ctx, ASM_K_RF_POPC);
    rf_popc_psrf_kernel<<<grid, kThreads, kSmemBytes, ctx->stream>>>(
        (const uint64_t*)ctx->opnd_bits.p, L.Wp, (const int2*)ctx->blk_meta.p, n_pairs, shard.shard_index,
        shard.shard_count, (unsigned long long*)ctx->tile_counter.p, (unsigned long long*)ctx->S_pad.p, L.C);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

int32_t rf_popc_block(asm_ctx* ctx, int32_t chainA, int64_t r0, int64_t r1, int32_t chainB, int64_t c0, int64_t c1,
                      uint16_t* out_host) {
  const long long nr = r1 - r0, nc = c1 - c0;
  int32_t rc = asm_reserve(ctx, ctx->rf_out, sizeof(uint16_t) * (size_t)(nr * nc));
  if (rc != ASM_OK) return rc;
  {
    LaunchScope ls(ctx, ASM_K_RF_POPC);
    int blocks = (int)std::min<long long>((nr * nc + 255) / 256, (long long)ctx->sm_count * 16);
    rf_block_kernel<<<blocks, 256, 0, ctx->stream>>>(ctx->trees[chainA].d_bits + (size_t)r0 * ctx->words,
                                                     ctx->trees[chainB].d_bits + (size_t)c0 * ctx->words, ctx->words, nr,
                                                     nc, (uint16_t*)ctx->rf_out.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  ASM_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->rf_out.p, sizeof(uint16_t) * (size_t)(nr * nc), cudaMemcpyDeviceToHost,
                                ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}
