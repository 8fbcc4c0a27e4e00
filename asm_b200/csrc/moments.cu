// moments.cu -- per-(chain, column) sequential sum and sum of squares of the trace logs:
// the two accumulation loops of GelmanRubin.calcGRStat (GelmanRubin.java:61-73),
//     mean += d; sumsq += d * d;        for i in [0, sampleCount)
// kept in their Java order (one dependent DADD per sample per chain of adds, DMUL separate from DADD) so that
// the host can finish the statistic bit-identically.  One warp per (chain, column): the lanes stage 256 samples
// in shared memory, lane 0 runs both chains.  Latency-bound and tiny next to the ESS kernels.
#include <algorithm>

#include "asm_internal.h"

namespace {

constexpr int kWarps = 4;

struct MomentParams {
  int32_t n_chains, n_cols, end;
  const double* vals[ASM_MAX_CHAINS];  // row-major [n_samples][n_cols]
};

__global__ void __launch_bounds__(32 * kWarps) trace_moments_kernel(const __grid_constant__ MomentParams prm_,
                                                                    double* __restrict__ sums, double* __restrict__ sumsq) {
  const MomentParams* prm = &prm_;
  __shared__ double bufs[kWarps][2][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int item = blockIdx.x * kWarps + wid;
  if (item >= prm->n_chains * prm->n_cols) return;
  const int chain = item / prm->n_cols, col = item % prm->n_cols;
  const double* __restrict__ x = prm->vals[chain] + col;
  const long long stride = prm->n_cols, n = prm->end;
  double(*buf)[256] = bufs[wid];
  double s = 0.0, q = 0.0;
  int cur = 0;
  for (int k = lane; k < 256 && k < n; k += 32) buf[0][k] = x[k * stride];
  __syncwarp();
  for (long long base = 0; base < n; base += 256) {
    const int m = (int)min(256LL, n - base);
    const long long nb = base + 256;
    if (lane != 0) {
      for (long long k = nb + lane - 1; k < nb + 256 && k < n; k += 31) buf[cur ^ 1][k - nb] = x[k * stride];
    } else {
#pragma unroll 4
      for (int k = 0; k < m; k++) {
        const double d = buf[cur][k];
        s = __dadd_rn(s, d);                 // mean1 += d      (:64)
        q = __dadd_rn(q, __dmul_rn(d, d));   // sumsq1 += d * d (:65)
      }
    }
    __syncwarp();
    cur ^= 1;
  }
  if (lane == 0) {
    sums[item] = s;
    sumsq[item] = q;
  }
}

}  // namespace

int32_t trace_moments(asm_ctx* ctx, int32_t end, double* sums_host, double* sumsq_host) {
  const int C = ctx->n_chains;
  const int n_cols = ctx->traces[0].n_cols;
  MomentParams prm;
  prm.n_chains = C;
  prm.n_cols = n_cols;
  prm.end = end;
  for (int c = 0; c < C; c++) prm.vals[c] = ctx->traces[c].d_vals;
  const int items = C * n_cols;
  int32_t rc = asm_reserve(ctx, ctx->misc, sizeof(double) * 2 * (size_t)items);
  if (rc != ASM_OK) return rc;
  double* d_s = (double*)ctx->misc.p;
  double* d_q = d_s + items;
  {
    LaunchScope ls(ctx, ASM_K_MISC);
    trace_moments_kernel<<<(items + kWarps - 1) / kWarps, 32 * kWarps, 0, ctx->stream>>>(prm, d_s, d_q);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  ASM_CUDA(ctx, cudaMemcpyAsync(sums_host, d_s, sizeof(double) * (size_t)items, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaMemcpyAsync(sumsq_host, d_q, sizeof(double) * (size_t)items, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}
