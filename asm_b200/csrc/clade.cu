// clade.cu -- K4: per-chain clade counts = column sums of the bit matrix
// (CladeDifference.process, CladeDifference.java:147-159: how many trees of a chain contain each clade).
//
// HBM-bound: every word of rows [from,to) is read once (coalesced: consecutive lanes read consecutive
// words of one row); each thread keeps the 64 counters of its word column in registers over a slab of
// trees and flushes them with integer atomics.
#include <algorithm>

#include "asm_internal.h"

namespace {

constexpr int kSlab = 1024;  // trees per block in y

__global__ void __launch_bounds__(128) clade_counts_kernel(const uint64_t* __restrict__ bits, int words, long long from,
                                                            long long to, int* __restrict__ out) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words) return;
  const long long t0 = from + (long long)blockIdx.y * kSlab;
  const long long t1 = min(t0 + (long long)kSlab, to);
  int cnt[64];
#pragma unroll
  for (int b = 0; b < 64; b++) cnt[b] = 0;
  for (long long t = t0; t < t1; t++) {
    const uint64_t v = bits[(size_t)t * words + w];
    if (v == 0) continue;  // rows are sparse: n-1 bits out of D
#pragma unroll
    for (int b = 0; b < 64; b++) cnt[b] += (int)((v >> b) & 1ULL);
  }
#pragma unroll
  for (int b = 0; b < 64; b++)
    if (cnt[b]) atomicAdd(&out[w * 64 + b], cnt[b]);
}

}  // namespace

int32_t clade_counts(asm_ctx* ctx, int32_t chain, int64_t from, int64_t to, int32_t* out_host) {
  const int words = ctx->words;
  const size_t bytes = sizeof(int32_t) * (size_t)words * 64;
  if (words == 0) return ASM_OK;
  int32_t rc = asm_reserve(ctx, ctx->misc, bytes);
  if (rc != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemsetAsync(ctx->misc.p, 0, bytes, ctx->stream));
  // one rank of N counts its contiguous share of the trees (every rank holds every tree); the int32 all-reduce
  // below adds the shares
  const int64_t per = (to - from + ctx->world - 1) / ctx->world;
  const int64_t lo = std::min<int64_t>(from + (int64_t)ctx->rank * per, to), hi = std::min<int64_t>(lo + per, to);
  if (hi > lo) {
    LaunchScope ls(ctx, ASM_K_CLADE_COUNTS);
    dim3 grid((unsigned)((words + 127) / 128), (unsigned)((hi - lo + kSlab - 1) / kSlab));
    clade_counts_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->trees[chain].d_bits, words, lo, hi, (int*)ctx->misc.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  if ((rc = comm_all_reduce_i32(ctx, (int32_t*)ctx->misc.p, (size_t)words * 64)) != ASM_OK) return rc;
  if (out_host) ASM_CUDA(ctx, cudaMemcpyAsync(out_host, ctx->misc.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}
