// ess.cu -- K5/K6: batched autocorrelation time / effective sample size, bit-exact with
// TraceESS.ACT (TraceESS.java:130-179).
//
// Only the last outer iteration of the Java loop reaches the result (autoCorrelation[] is
// overwritten on every iteration), and everything that iteration reads is a
// sequential floating-point chain:
//     squareLaggedSums[lag] = sum_{i=lag}^{n-1} trace[i-lag]*trace[i]      (:149)  i ascending
//     sum                   = sum_{i<n} trace[i]                            (:137)  i ascending
//     sum1, sum2            = sum - trace[n-1] - ... , sum - trace[0] - ... (:156-157)
// The (trace, lag) chains are independent of each other, so the kernel keeps every chain in its
// Java order -- one accumulator per (trace, lag), samples visited in ascending i, a separate
// multiply and add per step (DMUL + DADD, never DFMA) -- and spreads the chains over threads.
//
// K5 ess_lag_kernel: one CTA = 128 threads x 8 lags = 1024 consecutive lags of one trace.  The
// trace streams through a 4096-sample shared-memory ring (cp.async, one chunk of 512 samples in
// flight while the previous one is consumed).  Per group of 8 samples a thread reads 8 new window
// values and the 8 current samples (LDS.128, XOR-swizzled 64-byte blocks so that the 8 lanes of a
// quarter warp hit 8 different bank groups) and issues 64 DMUL + 64 DADD on a 15-value register
// window.  Bound: FP64 pipe; arithmetic intensity ~500 flop/byte of HBM.
// ess_sum_kernel: the sequential `sum` chain, on a side stream.  K6 ess_finish_kernel: one thread per trace:
// the autocorrelation normalisation (:154-155), the pair-sum truncated integral (:161-172) and ACT/ESS (:178,:126-128).
#include <algorithm>
#include <cmath>

#include "asm_internal.h"

namespace {

constexpr int kNT = 128;             // threads per CTA
constexpr int kR = 8;                // lags per thread
constexpr int kLagsPerCta = kNT * kR;  // 1024
constexpr int kChunk = 512;          // samples per cp.async group
constexpr int kRing = 4096;          // ring size in samples (>= max lag + 8 + 2*kChunk)

// logical ring index -> physical: 64-byte blocks, the four 16-byte sub-blocks XOR-swizzled
__device__ __forceinline__ int ring_phys(int e) {
  int g = e >> 3;
  int sub = (e >> 1) & 3;
  return (g << 3) | (((sub ^ ((g >> 1) & 3)) << 1) | (e & 1));
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// read the aligned 8-sample block that starts at logical index e0 (multiple of 8)
__device__ __forceinline__ void load_block8(const double* ring, int e0, double (&v)[8]) {
  const int g = (e0 & (kRing - 1)) >> 3;
  const int sw = (g >> 1) & 3;
  const double2* base = reinterpret_cast<const double2*>(ring + (g << 3));
#pragma unroll
  for (int s = 0; s < 4; s++) {
    double2 d = base[s ^ sw];
    v[2 * s] = d.x;
    v[2 * s + 1] = d.y;
  }
}

__global__ void __launch_bounds__(kNT) ess_lag_kernel(const double* __restrict__ traces, long long stride, long long n,
                                                      int lag_blocks, double* __restrict__ sls) {
  __shared__ __align__(16) double ring[kRing];
  const int tid = threadIdx.x;
  const long long trace = blockIdx.x / lag_blocks;
  const int lb = blockIdx.x % lag_blocks;
  const double* __restrict__ x = traces + trace * stride;
  const int l0 = lb * kLagsPerCta + tid * kR;

  for (int e = tid; e < kRing; e += kNT) ring[e] = 0.0;  // x[i] = 0 for i < 0
  __syncthreads();

  const long long n8 = (n + 7) & ~7LL;
  const long long n_chunks = (n8 + kChunk - 1) / kChunk;

  auto load_chunk = [&](long long c) {
    const long long base = c * kChunk;
#pragma unroll
    for (int e = tid; e < kChunk; e += kNT) {
      const long long i = base + e;
      double* dst = ring + ring_phys((int)(i & (kRing - 1)));
      if (i < n)
        cp_async8(dst, x + i);
      else
        *dst = 0.0;  // tail padding: adds +-0.0, which leaves every accumulator unchanged
    }
    cp_async_commit();
  };

  double acc[kR];
  double w[2 * kR - 1];
#pragma unroll
  for (int r = 0; r < kR; r++) acc[r] = 0.0;
#pragma unroll
  for (int k = 0; k < 2 * kR - 1; k++) w[k] = 0.0;

  if (n_chunks > 0) load_chunk(0);
  for (long long c = 0; c < n_chunks; c++) {
    if (c + 1 < n_chunks) {
      load_chunk(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const long long i_begin = c * kChunk;
    const long long i_end = min(i_begin + (long long)kChunk, n8);
    for (long long i = i_begin; i < i_end; i += 8) {
      double xi[8], wn[8];
      load_block8(ring, (int)(i & (kRing - 1)), xi);
      load_block8(ring, (int)((i - l0) & (kRing - 1)), wn);
#pragma unroll
      for (int k = 0; k < 8; k++) w[kR - 1 + k] = wn[k];
      // w[k] = x[i - l0 - 7 + k]; lag l0+r pairs x[i+di] with x[i+di-l0-r] = w[7 + di - r]
#pragma unroll
      for (int di = 0; di < 8; di++)
#pragma unroll
        for (int r = 0; r < kR; r++) acc[r] = __dadd_rn(acc[r], __dmul_rn(w[kR - 1 + di - r], xi[di]));
#pragma unroll
      for (int k = 0; k < kR - 1; k++) w[k] = w[k + 8];
    }
  }
  double* out = sls + (trace * lag_blocks + lb) * (long long)kLagsPerCta + tid * kR;
#pragma unroll
  for (int r = 0; r < kR; r++) out[r] = acc[r];
}

// The sequential `sum` chain (TraceESS.java:137): one warp per trace, 8 traces per block (few, fat blocks,
// so that the lag kernel keeps its CTA slots).  The lanes stage 256 samples in shared memory while lane 0
// adds the previous 256 in order.  Latency-bound (one dependent DADD per sample): it runs on a side stream
// underneath the lag kernel.
constexpr int kSumWarps = 8;
__global__ void __launch_bounds__(32 * kSumWarps) ess_sum_kernel(const double* __restrict__ traces, long long stride,
                                                                 long long n, int n_traces, double* __restrict__ sums) {
  __shared__ double bufs[kSumWarps][2][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long trace = (long long)blockIdx.x * kSumWarps + wid;
  if (trace >= n_traces) return;
  double(*buf)[256] = bufs[wid];
  const double* __restrict__ x = traces + trace * stride;
  double sum = 0.0;
  int cur = 0;
  for (int k = lane; k < 256 && k < n; k += 32) buf[0][k] = x[k];
  __syncwarp();
  for (long long base = 0; base < n; base += 256) {
    const int m = (int)min(256LL, n - base);
    const long long nb = base + 256;
    if (lane != 0) {  // lanes 1..31 prefetch the next block while lane 0 adds
      for (long long k = nb + lane - 1; k < nb + 256 && k < n; k += 31) buf[cur ^ 1][k - nb] = x[k];
    } else {
#pragma unroll 8
      for (int k = 0; k < m; k++) sum = __dadd_rn(sum, buf[cur][k]);
    }
    __syncwarp();
    cur ^= 1;
  }
  if (lane == 0) sums[trace] = sum;
}

// One thread per trace: normalisation (:154-155), pair-sum truncated integral (:161-172), ACT and ESS.
__global__ void __launch_bounds__(128) ess_finish_kernel(const double* __restrict__ traces, long long stride, long long n,
                                                         int n_traces, int max_lag, int lags_stride,
                                                         const double* __restrict__ sls, const double* __restrict__ sums,
                                                         double* __restrict__ act_out, double* __restrict__ ess_out,
                                                         int* __restrict__ lags_used) {
  const long long trace = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (trace >= n_traces) return;
  const double* __restrict__ x = traces + trace * stride;
  const double sum = sums[trace];
  const long long L = n < (long long)max_lag ? n : (long long)max_lag;  // :160
  const double* s = sls + trace * (long long)lags_stride;
  double act;
  int used = 0;
  if (L == 0) {
    act = __longlong_as_double(0x7ff8000000000000LL);  // 0.0 / 0.0
  } else {
    const double mean = __ddiv_rn(sum, (double)n);  // :139 at i = n-1
    const double mm = __dmul_rn(mean, mean);
    double sum1 = sum, sum2 = sum;
    double integral = 0.0, ac0 = 0.0, prev = 0.0;
    for (long long lag = 0; lag < L; lag++) {
      // :154-155  ((SLS - (sum1+sum2)*mean) + (mean*mean)*(n-lag)) / (n+start-lag), start = 0
      double a = __dsub_rn(s[lag], __dmul_rn(__dadd_rn(sum1, sum2), mean));
      a = __dadd_rn(a, __dmul_rn(mm, (double)(n - lag)));
      a = __ddiv_rn(a, (double)(n - lag));
      if (lag == 0) {  // :163-164
        integral = a;
        ac0 = a;
        used = 1;
      } else if ((lag & 1) == 0) {  // :165-171
        used = (int)lag + 1;
        const double pair = __dadd_rn(prev, a);
        if (pair > 0)
          integral = __dadd_rn(integral, __dmul_rn(2.0, pair));
        else
          break;
      }
      prev = a;
      sum1 = __dsub_rn(sum1, x[n - 1 - lag]);  // :156
      sum2 = __dsub_rn(sum2, x[lag]);          // :157
    }
    act = __ddiv_rn(integral, ac0);  // :178
  }
  if (act_out) act_out[trace] = act;
  if (ess_out) ess_out[trace] = __ddiv_rn((double)n, act);  // :126-128
  if (lags_used) lags_used[trace] = used;
}

struct ConcatParams {
  int32_t n_chains, n_cols, end;
  int64_t total, stride;
  int32_t burnin[ASM_MAX_CHAINS];
  int64_t offset[ASM_MAX_CHAINS];  // start of chain i inside the concatenated trace
  const double* vals[ASM_MAX_CHAINS];
  int32_t src_cols[ASM_MAX_CHAINS];
  int32_t cols[64];
};

// trace[j][offset[i] + (x - burnin[i])] = logLines[i][cols[j]].get(x)   (TraceESS.java:55-60)
__global__ void __launch_bounds__(256) concat_kernel(const ConcatParams* __restrict__ prm, double* __restrict__ out) {
  const long long total = prm->total;
  const long long all = total * prm->n_cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < all;
       idx += (long long)gridDim.x * blockDim.x) {
    int j = (int)(idx / total);
    long long k = idx - (long long)j * total;
    int i = 0;
    while (i + 1 < prm->n_chains && k >= prm->offset[i + 1]) i++;
    long long xs = prm->burnin[i] + (k - prm->offset[i]);
    out[(long long)j * prm->stride + k] = prm->vals[i][xs * prm->src_cols[i] + prm->cols[j]];
  }
}

}  // namespace

int32_t ess_batch_device(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* d_traces,
                         int32_t max_lag, double* act_out_host, double* ess_out_host) {
  const long long L = std::min<long long>(n_samples, max_lag);
  const int lag_blocks = (int)std::max<long long>(1, (L + kLagsPerCta - 1) / kLagsPerCta);
  const int lags_stride = lag_blocks * kLagsPerCta;
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->ess_sls, sizeof(double) * (size_t)n_traces * lags_stride)) != ASM_OK) return rc;
  if ((rc = asm_reserve(ctx, ctx->ess_out, sizeof(double) * 3 * (size_t)n_traces + sizeof(int) * ((size_t)n_traces + 2))) != ASM_OK)
    return rc;
  double* d_act = (double*)ctx->ess_out.p;
  double* d_ess = d_act + n_traces;
  int* d_used = (int*)(d_ess + n_traces);
  double* d_sums = (double*)(d_used + n_traces + (n_traces & 1));
  // sum chain on the side stream, underneath the lag kernel
  ASM_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  ASM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream2, ctx->ev_fork, 0));
  ctx->launches++;
  ess_sum_kernel<<<(n_traces + kSumWarps - 1) / kSumWarps, 32 * kSumWarps, 0, ctx->stream2>>>(d_traces, stride, n_samples,
                                                                                              n_traces, d_sums);
  ASM_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->stream2));
  if (n_samples > 0) {
    LaunchScope ls(ctx, ASM_K_ESS_LAG);
    ess_lag_kernel<<<n_traces * lag_blocks, kNT, 0, ctx->stream>>>(d_traces, stride, n_samples, lag_blocks,
                                                                   (double*)ctx->ess_sls.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  ASM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
  {
    LaunchScope ls(ctx, ASM_K_ESS_FINISH);
    ess_finish_kernel<<<(n_traces + 127) / 128, 128, 0, ctx->stream>>>(d_traces, stride, n_samples, n_traces, max_lag,
                                                                       lags_stride, (const double*)ctx->ess_sls.p, d_sums,
                                                                       d_act, d_ess, d_used);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  if (act_out_host)
    ASM_CUDA(ctx, cudaMemcpyAsync(act_out_host, d_act, sizeof(double) * (size_t)n_traces, cudaMemcpyDeviceToHost, ctx->stream));
  if (ess_out_host)
    ASM_CUDA(ctx, cudaMemcpyAsync(ess_out_host, d_ess, sizeof(double) * (size_t)n_traces, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

int32_t ess_eval_stream(asm_ctx* ctx, const int32_t* burnin, int32_t end, const int32_t* cols, int32_t n_cols,
                        double* ess_out_host) {
  if (n_cols > 64) return asm_fail(ctx, ASM_E_UNSUPPORTED, "ess_eval: at most 64 columns per call, not %d", n_cols);
  ConcatParams prm;
  prm.n_chains = ctx->n_chains;
  prm.n_cols = n_cols;
  prm.end = end;
  int64_t total = 0;
  for (int i = 0; i < ctx->n_chains; i++) {  // TraceESS.java:40-43
    prm.burnin[i] = burnin[i];
    prm.offset[i] = total;
    prm.vals[i] = ctx->traces[i].d_vals;
    prm.src_cols[i] = ctx->traces[i].n_cols;
    total += std::max(0, end - burnin[i]);
  }
  for (int j = 0; j < n_cols; j++) prm.cols[j] = cols[j];
  prm.total = total;
  prm.stride = std::max<int64_t>(2, (total + 1) & ~int64_t(1));
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->ess_tr, sizeof(double) * (size_t)n_cols * (size_t)prm.stride)) != ASM_OK) return rc;
  if ((rc = asm_reserve(ctx, ctx->misc, sizeof prm)) != ASM_OK) return rc;
  ASM_CUDA(ctx, cudaMemcpyAsync(ctx->misc.p, &prm, sizeof prm, cudaMemcpyHostToDevice, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (total > 0) {
    LaunchScope ls(ctx, ASM_K_MISC);
    int blocks = (int)std::min<int64_t>((total * n_cols + 255) / 256, (int64_t)ctx->sm_count * 8);
    concat_kernel<<<blocks, 256, 0, ctx->stream>>>((const ConcatParams*)ctx->misc.p, (double*)ctx->ess_tr.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ess_batch_device(ctx, n_cols, total, prm.stride, (const double*)ctx->ess_tr.p, ASM_MAX_LAG, nullptr, ess_out_host);
}
