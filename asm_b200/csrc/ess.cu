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
// K5 ess_lag_kernel<R>: one CTA = blockDim.x threads x R consecutive lags of one trace.  The trace
// streams through a 4096-sample shared-memory ring (cp.async, one chunk of 512 samples in flight while
// the previous one is consumed).  Per group of R samples a thread reads R new window values and the R
// current samples (LDS.128, XOR-swizzled 64-byte blocks so that the 8 lanes of a quarter warp hit 8
// different bank groups) and issues R*R DMUL + R*R DADD on a (2R-1)-value register window.
// Bound: FP64 pipe; arithmetic intensity ~500 flop/byte of HBM.
//
// Staged evaluation: the integral of TraceESS.java:161-172 stops at the first even lag whose pair sum
// autoCorrelation[lag-1] + autoCorrelation[lag] is not positive, so lags beyond that point cannot change
// the result.  Lags are therefore evaluated in stages [0,128), [128,256), [256,768), [768,1792), [1792,2048); after each stage
// K6 ess_scan_kernel advances every still-open trace's integral and retires the traces whose loop has
// stopped (or has reached min(n, 2000) lags).  The output is bit-identical to evaluating all lags
// (asm_ess_set_adaptive(ctx, 0) / ASM_ESS_ADAPTIVE=0 does exactly that, in one stage).
// The sequential `sum` chain rides along in the first warp of each trace's first lag block (first stage).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <thread>
#include <type_traits>

#include "asm_internal.h"

namespace {

constexpr int kChunkMax = 512;  // samples per cp.async group: 512, or 256 where that lets the ring be half as large
// Ring size in samples >= largest lag of the stage + 8 + 3*chunk: a thread that has passed the barrier of chunk c
// issues the copy of chunk c+1 while a slower thread of its group may still be reading the window of chunk c-1.

// Ring layout.  A thread reads its R new window values as R/2 16-byte loads; the 8 lanes of a quarter warp sit
// 8*R bytes apart and must hit 8 different 16-byte bank groups.  Instead of an XOR swizzle (one LOP3 per load) the
// ring is padded: 16 bytes after every 8 samples (R = 8: blocks 80 bytes apart, 5b mod 8 is a permutation) or after
// every 4 samples (R = 4: half blocks 48 bytes apart, 3h mod 8 is a permutation); with 2 or 1 lags per thread a
// quarter warp reads contiguous bytes and needs no padding.  An aligned block therefore starts at e * off(R) / R
// bytes and its 16-byte pieces follow at immediates.  The first `chunk` samples are mirrored behind the ring's end
// so that the consecutive window blocks of one chunk never wrap: the inner loop only adds a constant to
// its two addresses.
template <int R>
struct Ring {
  static constexpr int kPadShift = R == 8 ? 3 : (R == 4 ? 2 : 0);
  __host__ __device__ static constexpr uint32_t off(uint32_t e) {
    return e * 8u + (kPadShift ? (e >> kPadShift) * 16u : 0u);
  }
  // ring of `ring_size` samples + mirror of the first chunk + one block of read-ahead
  __host__ __device__ static constexpr uint32_t bytes(uint32_t ring_size, uint32_t chunk) {
    return (off(ring_size + chunk + 16) + 127u) & ~127u;
  }
};

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// read N consecutive samples (no padding in between) at byte offset `o` of the ring
template <int N, bool kAligned16>
__device__ __forceinline__ void load_block(const char* ring, uint32_t o, double (&v)[N]) {
  if constexpr (!kAligned16) {
#pragma unroll
    for (int s = 0; s < N; s++) v[s] = *reinterpret_cast<const double*>(ring + o + 8 * s);
  } else {
#pragma unroll
    for (int s = 0; s < N / 2; s++) {
      double2 d = *reinterpret_cast<const double2*>(ring + o + 16 * s);
      v[2 * s] = d.x;
      v[2 * s + 1] = d.y;
    }
  }
}

// A CTA is always 4 (or 8) warps so that all four SM sub-partitions issue FP64 work; it is cut into groups of
// `gthreads` threads, one group = gthreads*R consecutive lags of one trace with its own ring in shared memory.
// grid: x = ceil(n_open / groups per CTA), y = lag block of this stage.
template <int R>
__global__ void __launch_bounds__(256) ess_lag_kernel(const double* __restrict__ traces, long long stride, long long n,
                                                      int lag_begin, const int* __restrict__ trace_ids, int n_open,
                                                      int lags_stride, double* __restrict__ sls, int gthreads,
                                                      int ring_size, int chunk, double* __restrict__ sums) {
  extern __shared__ __align__(128) char ring_all[];
  const int group = threadIdx.x / gthreads, tid = threadIdx.x - group * gthreads, NT = gthreads;
  const int groups = blockDim.x / gthreads;
  const int idx = blockIdx.x * groups + group;
  if (idx >= n_open) return;  // whole group idle; groups synchronise among themselves only
  const uint32_t ring_bytes = Ring<R>::bytes((uint32_t)ring_size, (uint32_t)chunk);
  char* ring = ring_all + (size_t)group * ring_bytes;
  const int mask = ring_size - 1;
  const uint32_t mirror_off = Ring<R>::off((uint32_t)ring_size);
  const long long trace = trace_ids ? trace_ids[idx] : idx;
  const double* __restrict__ x = traces + trace * stride;
  const int l0 = lag_begin + (int)blockIdx.y * NT * R + tid * R;
  auto group_sync = [&]() {
    if (NT == 32)
      __syncwarp();
    else
      asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(NT) : "memory");
  };

  for (uint32_t o = tid * 16; o < ring_bytes; o += NT * 16)  // x[i] = 0 for i < 0
    *reinterpret_cast<double2*>(ring + o) = make_double2(0.0, 0.0);
  group_sync();

  constexpr int kRound = R >= 4 ? R : 8;  // samples per inner-loop step
  const long long nR = (n + kRound - 1) / kRound * kRound;
  const long long n_chunks = (nR + chunk - 1) / chunk;

  auto load_chunk = [&](long long c) {
    const long long base = c * chunk;
    const uint32_t e0 = (uint32_t)(base & mask);  // the chunk occupies [e0, e0 + chunk), no wrap
    const bool mirror = e0 == 0;
    for (int e = tid; e < chunk; e += NT) {
      const long long i = base + e;
      char* dst = ring + Ring<R>::off(e0 + (uint32_t)e);
      if (i < n) {
        cp_async8(dst, x + i);
        if (mirror) cp_async8(dst + mirror_off, x + i);
      } else {  // tail padding: adds +-0.0, which leaves every accumulator unchanged
        *reinterpret_cast<double*>(dst) = 0.0;
        if (mirror) *reinterpret_cast<double*>(dst + mirror_off) = 0.0;
      }
    }
    cp_async_commit();
  };

  // SB samples per inner-loop step: R for the wide shapes; 8 for 1 or 2 lags per thread, whose steps would
  // otherwise be shorter than the shared-memory latency that the one-step read-ahead has to cover.
  constexpr int SB = R >= 4 ? R : 8;
  constexpr bool kWinAligned = R >= 2;  // window blocks start at multiples of R samples
  double acc[R];
  double w[R - 1 + SB];
#pragma unroll
  for (int r = 0; r < R; r++) acc[r] = 0.0;
#pragma unroll
  for (int k = 0; k < R - 1 + SB; k++) w[k] = 0.0;

  // The sequential `sum` chain (TraceESS.java:137) rides along in the first warp of the trace's first lag block
  // of the first stage: one more dependent DADD per sample in that warp, instead of a latency-bound kernel of
  // its own whose warps would take a quarter of the FP64 issue slots of every sub-partition they sit on.
  const bool do_sum = sums != nullptr && blockIdx.y == 0 && tid < 32;
  double sum = 0.0;
  constexpr uint32_t kStep = Ring<R>::off(SB);  // bytes from one block of SB samples to the next
  auto compute_chunk = [&](long long c, auto with_sum) {
    const long long i_begin = c * chunk;
    const int n_blocks = (int)((min(i_begin + (long long)chunk, nR) - i_begin) / SB);
    // x[i..i+SB) sits at xo, the window block x[i-l0..i-l0+SB) at wo; both advance by kStep per block.
    uint32_t xo = Ring<R>::off((uint32_t)(i_begin & mask));
    uint32_t wo = Ring<R>::off((uint32_t)((i_begin - l0) & mask));
    // software pipeline: the loads of block k+1 are issued before the arithmetic of block k, so that a warp
    // that has an SM sub-partition almost to itself does not stall on shared-memory latency (the read-ahead of
    // the last block lands in the next chunk / the mirror and is discarded)
    double xi[SB], wn[SB], xi_n[SB], wn_n[SB];
    load_block<SB, true>(ring, xo, xi);
    load_block<SB, kWinAligned>(ring, wo, wn);
#pragma unroll(R == 1 ? 1 : 2)
    for (int k = 0; k < n_blocks; k++) {
      xo += kStep;
      wo += kStep;
      load_block<SB, true>(ring, xo, xi_n);
      load_block<SB, kWinAligned>(ring, wo, wn_n);
#pragma unroll
      for (int k2 = 0; k2 < SB; k2++) w[R - 1 + k2] = wn[k2];
      // w[k] = x[i - l0 - (R-1) + k]; lag l0+r pairs x[i+di] with x[i+di-l0-r] = w[R-1 + di - r]
#pragma unroll
      for (int di = 0; di < SB; di++)
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = __dadd_rn(acc[r], __dmul_rn(w[R - 1 + di - r], xi[di]));
      if constexpr (decltype(with_sum)::value) {  // the zero padding behind sample n-1 adds +0.0 to a sum that is never -0.0
#pragma unroll
        for (int di = 0; di < SB; di++) sum = __dadd_rn(sum, xi[di]);
      }
#pragma unroll
      for (int k2 = 0; k2 < R - 1; k2++) w[k2] = w[k2 + SB];
#pragma unroll
      for (int k2 = 0; k2 < SB; k2++) {
        xi[k2] = xi_n[k2];
        wn[k2] = wn_n[k2];
      }
    }
  };
  if (n_chunks > 0) load_chunk(0);
  for (long long c = 0; c < n_chunks; c++) {
    if (c + 1 < n_chunks) {
      load_chunk(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    group_sync();
    // Two copies of the chunk loop, with and without the sum chain: a predicated-off DADD still takes its FP64
    // issue slot, so the warps that do not carry the sum must not see those instructions at all.
    if (do_sum)
      compute_chunk(c, std::true_type{});
    else
      compute_chunk(c, std::false_type{});
  }
  if (do_sum && tid == 0) sums[trace] = sum;
  if (l0 + R <= lags_stride) {
    double* out = sls + trace * (long long)lags_stride + l0;
#pragma unroll
    for (int r = 0; r < R; r++) out[r] = acc[r];
  }
}


// ---------------------------------------------------------------------------------------------------------------
// K5t ess_ticket_kernel<R>: the same chains, cut into tickets.
//
// ess_lag_kernel gives a group of threads one (trace, lag block) for the whole trace: a launch with 1.7 warps per SM
// sub-partition takes as long as one with 2, and nothing else can use the slots a short group leaves behind.  Here a
// JOB is one lag block (gthreads*R lags) of one trace and a TICKET is one segment of `seg_len` samples of a job.
// Workers (groups of `gthreads` threads of a persistent grid, each with its own shared-memory ring) draw tickets from
// one atomic counter, segment-major: every segment 0 first, then every segment 1, ...  A chain's additions stay in
// their Java order because segment k of a job starts from the accumulators segment k-1 parked in global memory (an
// acquire/release flag per job orders the two; a ticket only ever waits for a lower-numbered one, which a running
// worker has already drawn, so the wait is short and cannot deadlock whatever else occupies the GPU).  The ring is
// re-primed per ticket with the `lag_hi + R` samples before the segment (zeros before sample 0).  The jobs of two
// trace lists with different lag ranges (the "short" and "long" halves of the pilot split) share one launch, so the
// early stages keep every sub-partition busy and finish together.  Results are bit-identical to ess_lag_kernel's.
// ---------------------------------------------------------------------------------------------------------------
struct EssJobSet {
  const int* ids;  // trace ids (nullptr: 0..n-1)
  int n;           // traces
  int lag_begin;
  int n_blocks;    // lag blocks of gthreads*R lags per trace
  int first_job;   // jobs [first_job, first_job + n*n_blocks): job j -> trace ids[(j-first_job) % n], lag block (j-first_job) / n
  int with_sum;    // the job of lag block 0 carries the trace's sequential sum (first stage of its group)
  int R;           // lags per thread of this set's jobs (the sets of a launch share `gthreads`, not R)
  size_t park_off; // this set's slice of `park`, in doubles: [n*n_blocks][gthreads*R + 1]
};
struct EssTicketParams {
  const double* traces;
  long long stride, n;
  EssJobSet set[2];  // the trace lists of up to two groups (the "short" and "long" halves of the pilot split) share a launch
  int n_sets, n_jobs;
  int seg_len, n_seg;  // samples per ticket (a multiple of `chunk`), tickets per job
  int lags_stride;
  double* sls;
  double* sums;
  double* park;        // accumulators (+ the running sum) between the segments of a job (EssJobSet::park_off)
  size_t ring_stride;  // bytes between the rings of the workers of a CTA
  int* progress;       // [n_jobs]: segments completed
  unsigned long long* ticket_counter;
  int gthreads, ring_size, chunk;
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// One ticket: segment k of job j of set `js`, R lags per thread.
template <int R>
__device__ __forceinline__ void ess_ticket_body(const EssTicketParams& prm, const EssJobSet& js, int k, int j, char* ring, int NT,
                                                int group, int tid) {
  const int ring_size = prm.ring_size, chunk = prm.chunk;
  const long long mask = ring_size - 1;
  const uint32_t mirror_off = Ring<R>::off((uint32_t)ring_size);
  const long long n = prm.n;
  constexpr int kRound = R >= 4 ? R : 8;  // samples per inner-loop step
  constexpr int SB = kRound;
  constexpr bool kWinAligned = R >= 2;
  constexpr uint32_t kStep = Ring<R>::off(SB);
  const long long nR = (n + 7) / 8 * 8;
  auto group_sync = [&]() {
    if (NT == 32)
      __syncwarp();
    else
      asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(NT) : "memory");
  };
  const int jj = j - js.first_job;
  const int blk = jj / js.n;
  const int ti = jj - blk * js.n;
  const long long trace = js.ids ? js.ids[ti] : ti;
  const double* __restrict__ x = prm.traces + trace * prm.stride;
  const int l0 = js.lag_begin + blk * NT * R + tid * R;
  const int lag_hi = js.lag_begin + (blk + 1) * NT * R;  // exclusive
  const bool do_sum = js.with_sum && blk == 0 && tid < 32;
  const long long i0 = (long long)k * prm.seg_len;
  const long long i1 = min(nR, i0 + (long long)prm.seg_len);

  auto load_chunk = [&](long long c) {
    const long long base = c * chunk;
    const uint32_t e0 = (uint32_t)(base & mask);  // the chunk occupies [e0, e0 + chunk), no wrap
    const bool mirror = e0 == 0;
    for (int e = tid; e < chunk; e += NT) {
      const long long i = base + e;
      char* dst = ring + Ring<R>::off(e0 + (uint32_t)e);
      if (i >= 0 && i < n) {
        cp_async8(dst, x + i);
        if (mirror) cp_async8(dst + mirror_off, x + i);
      } else {  // x[i] = 0 before the trace; behind it the padding adds +-0.0, which leaves every accumulator unchanged
        *reinterpret_cast<double*>(dst) = 0.0;
        if (mirror) *reinterpret_cast<double*>(dst + mirror_off) = 0.0;
      }
    }
    cp_async_commit();
  };

  double acc[R];
  double w[R - 1 + SB];
  double sum = 0.0;
  double* park = prm.park + js.park_off + (size_t)jj * ((size_t)NT * R + 1);
  if (k == 0) {
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = 0.0;
  } else {
    if (tid == 0) {
      uint32_t spins = 0;
      while (ld_acquire(prm.progress + j) < k) {
        if (++spins > (1u << 30)) __trap();  // a protocol error fails the launch instead of hanging the GPU
      }
    }
    group_sync();
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = __ldcg(park + tid * R + r);
    if (do_sum) sum = __ldcg(park + NT * R);
  }

  auto compute_chunk = [&](long long c, auto with_sum) {
    const long long i_begin = c * chunk;
    const int n_blocks = (int)((min(i_begin + (long long)chunk, i1) - i_begin) / SB);
    uint32_t xo = Ring<R>::off((uint32_t)(i_begin & mask));
    uint32_t wo = Ring<R>::off((uint32_t)((i_begin - l0) & mask));
    double xi[SB], wn[SB], xi_n[SB], wn_n[SB];
    load_block<SB, true>(ring, xo, xi);
    load_block<SB, kWinAligned>(ring, wo, wn);
#pragma unroll(R == 1 ? 1 : 2)
    for (int kb = 0; kb < n_blocks; kb++) {
      xo += kStep;
      wo += kStep;
      load_block<SB, true>(ring, xo, xi_n);
      load_block<SB, kWinAligned>(ring, wo, wn_n);
#pragma unroll
      for (int k2 = 0; k2 < SB; k2++) w[R - 1 + k2] = wn[k2];
#pragma unroll
      for (int di = 0; di < SB; di++)
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = __dadd_rn(acc[r], __dmul_rn(w[R - 1 + di - r], xi[di]));
      if constexpr (decltype(with_sum)::value) {
#pragma unroll
        for (int di = 0; di < SB; di++) sum = __dadd_rn(sum, xi[di]);
      }
#pragma unroll
      for (int k2 = 0; k2 < R - 1; k2++) w[k2] = w[k2 + SB];
#pragma unroll
      for (int k2 = 0; k2 < SB; k2++) {
        xi[k2] = xi_n[k2];
        wn[k2] = wn_n[k2];
      }
    }
  };

  // chunks [cb, c0) prime the ring with the window behind the segment, chunks [c0, c1) are computed
  const long long c0 = i0 / chunk, c1 = (i1 + chunk - 1) / chunk;
  const long long cb = c0 - (lag_hi + R + 8 + chunk - 1) / chunk;
  load_chunk(cb);
  for (long long c = cb; c < c1; c++) {
    if (c + 1 < c1) {
      load_chunk(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    group_sync();
    if (c < c0) continue;
    if (c == c0) {  // w[k] = x[i0 - l0 - (R-1) + k]: the R-1 window values a running walk would carry into this block
#pragma unroll
      for (int k2 = 0; k2 < R - 1; k2++)
        w[k2] = *reinterpret_cast<const double*>(ring + Ring<R>::off((uint32_t)((i0 - l0 - (R - 1) + k2) & mask)));
    }
    if (do_sum)
      compute_chunk(c, std::true_type{});
    else
      compute_chunk(c, std::false_type{});
  }

  if (k + 1 == prm.n_seg) {
    if (do_sum && tid == 0) prm.sums[trace] = sum;
    if (l0 + R <= prm.lags_stride) {
      double* out = prm.sls + trace * (long long)prm.lags_stride + l0;
#pragma unroll
      for (int r = 0; r < R; r++) out[r] = acc[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; r++) __stcg(park + tid * R + r, acc[r]);
    if (do_sum && tid == 0) __stcg(park + NT * R, sum);
    __threadfence();
    group_sync();
    if (tid == 0) st_release(prm.progress + j, k + 1);
  }
}

__global__ void __launch_bounds__(128) ess_ticket_kernel(const __grid_constant__ EssTicketParams prm) {
  extern __shared__ __align__(128) char ring_all[];
  __shared__ long long ticket_s[4];
  const int NT = prm.gthreads;
  const int group = threadIdx.x / NT, tid = threadIdx.x - group * NT;
  char* ring = ring_all + (size_t)group * prm.ring_stride;
  const long long total = (long long)prm.n_jobs * prm.n_seg;
  for (;;) {
    // every thread is done with the previous ticket's ring and has read ticket_s
    if (NT == 32)
      __syncwarp();
    else
      asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(NT) : "memory");
    if (tid == 0) ticket_s[group] = (long long)atomicAdd(prm.ticket_counter, 1ULL);
    if (NT == 32)
      __syncwarp();
    else
      asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(NT) : "memory");
    const long long t = ticket_s[group];
    if (t >= total) break;
    const int k = (int)(t / prm.n_jobs), j = (int)(t - (long long)k * prm.n_jobs);
    const EssJobSet& js = prm.set[(prm.n_sets > 1 && j >= prm.set[1].first_job) ? 1 : 0];
    switch (js.R) {  // the same for every thread of the worker
      case 8: ess_ticket_body<8>(prm, js, k, j, ring, NT, group, tid); break;
      case 4: ess_ticket_body<4>(prm, js, k, j, ring, NT, group, tid); break;
      case 2: ess_ticket_body<2>(prm, js, k, j, ring, NT, group, tid); break;
      default: ess_ticket_body<1>(prm, js, k, j, ring, NT, group, tid); break;
    }
  }
}

// per-trace state of the integral loop carried from stage to stage
struct EssState {
  double integral, prev, sum1, sum2, ac0;
  int next_lag;
};

// K6: one thread per open trace: advance the loop of TraceESS.java:161-172 over the lags [next_lag, stage_end)
// whose sums are now available; retire the trace (write ACT, ESS) when the loop stops or all L lags are used,
// otherwise append it to the next stage's work list.
__global__ void __launch_bounds__(128) ess_scan_kernel(const double* __restrict__ traces, long long stride, long long n,
                                                       int n_open, const int* __restrict__ ids_in, int max_lag, int stage_end,
                                                       int lags_stride, const double* __restrict__ sls,
                                                       const double* __restrict__ sums, EssState* __restrict__ state,
                                                       double* __restrict__ act_out, double* __restrict__ ess_out,
                                                       int* __restrict__ lags_used, int* __restrict__ ids_out,
                                                       int* __restrict__ n_out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_open) return;
  const long long trace = ids_in ? ids_in[k] : k;
  const double* __restrict__ x = traces + trace * stride;
  const double sum = sums[trace];
  const long long L = n < (long long)max_lag ? n : (long long)max_lag;  // :160
  const double* s = sls + trace * (long long)lags_stride;
  EssState st = state[trace];
  bool finished = false;
  if (L == 0) {
    st.integral = 0.0;
    st.ac0 = 0.0;  // 0.0 / 0.0 = NaN below
    finished = true;
  } else {
    const double mean = __ddiv_rn(sum, (double)n);  // :139 at i = n-1
    const double mm = __dmul_rn(mean, mean);
    if (st.next_lag == 0) {
      st.sum1 = sum;
      st.sum2 = sum;
      st.integral = 0.0;
      st.prev = 0.0;
      st.ac0 = 0.0;
    }
    const long long end = L < (long long)stage_end ? L : (long long)stage_end;
    long long lag = st.next_lag;
    for (; lag < end; lag++) {
      // :154-155  ((SLS - (sum1+sum2)*mean) + (mean*mean)*(n-lag)) / (n+start-lag), start = 0
      double a = __dsub_rn(s[lag], __dmul_rn(__dadd_rn(st.sum1, st.sum2), mean));
      a = __dadd_rn(a, __dmul_rn(mm, (double)(n - lag)));
      a = __ddiv_rn(a, (double)(n - lag));
      if (lag == 0) {  // :163-164
        st.integral = a;
        st.ac0 = a;
      } else if ((lag & 1) == 0) {  // :165-171
        const double pair = __dadd_rn(st.prev, a);
        if (pair > 0) {
          st.integral = __dadd_rn(st.integral, __dmul_rn(2.0, pair));
        } else {
          finished = true;  // break
          lag++;
          break;
        }
      }
      st.prev = a;
      st.sum1 = __dsub_rn(st.sum1, x[n - 1 - lag]);  // :156
      st.sum2 = __dsub_rn(st.sum2, x[lag]);          // :157
    }
    st.next_lag = (int)lag;
    if (lag >= L) finished = true;
  }
  if (finished) {
    const double act = __ddiv_rn(st.integral, st.ac0);  // :178
    if (act_out) act_out[trace] = act;
    if (ess_out) ess_out[trace] = __ddiv_rn((double)n, act);  // :126-128
    if (lags_used) lags_used[trace] = st.next_lag;
  } else {
    state[trace] = st;
    ids_out[atomicAdd(n_out, 1)] = (int)trace;
  }
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
__global__ void __launch_bounds__(256) concat_kernel(const __grid_constant__ ConcatParams prm_, double* __restrict__ out) {
  const ConcatParams* prm = &prm_;
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

// Pilot of the staged evaluation: one warp per trace estimates the autocorrelation at lag `lag` from the first P
// samples and files the trace under "long" (slowly decaying: its integral loop will run far past the first stages)
// or "short".  Only the schedule depends on it -- which lags are evaluated in one launch -- never a result.
__global__ void __launch_bounds__(256) ess_pilot_kernel(const double* __restrict__ traces, long long stride, int n_traces,
                                                         int P, int lag, double threshold, int* __restrict__ ids_short,
                                                         int* __restrict__ ids_long, int* __restrict__ counts) {
  __shared__ double red[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x;
  const double* __restrict__ x = traces + (long long)t * stride;
  auto block_sum = [&](double v, int slot) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[slot][warp] = v;
    __syncthreads();
    double r = 0.0;
    for (int k = 0; k < 8; k++) r += red[slot][k];
    __syncthreads();
    return r;
  };
  double m = 0.0;
  for (int i = threadIdx.x; i < P; i += 256) m += x[i];
  m = block_sum(m, 0) / (double)P;
  double v = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < P; i += 256) {
    const double d = x[i] - m;
    v += d * d;
    if (i + lag < P) c += d * (x[i + lag] - m);
  }
  v = block_sum(v, 1);
  c = block_sum(c, 2);
  if (threadIdx.x == 0) {
    const bool is_long = v > 0.0 && c > threshold * v;
    if (is_long)
      ids_long[atomicAdd(&counts[1], 1)] = t;
    else
      ids_short[atomicAdd(&counts[0], 1)] = t;
  }
}

}  // namespace

namespace {

// One group = a contiguous range of traces processed on its own stream.  Groups of one batch run concurrently:
// while one group is in a late, thinly populated stage another one is still in its wide first stage (or still
// being uploaded), which keeps the FP64 pipes busy.
struct EssGroup {
  int t0 = 0, nt = 0;
  cudaStream_t st = nullptr, st2 = nullptr;
  cudaEvent_t ready = nullptr;  // traces of this group are on the device (may be null)
  cudaEvent_t fork = nullptr, join = nullptr;
  int stage = 0, n_open = 0;
  const int* ids_in = nullptr;
  bool done = false;
  // A group may be a SUBSET of its trace range (the "short" / "long" halves the pilot makes of one range): ids0 lists
  // its n0 traces (relative to t0); `slot` selects its private work-list buffers and counter; the state of the range is
  // cleared and the outputs are copied by the caller, once, for both halves.
  const int* ids0 = nullptr;
  int n0 = 0;
  int slot = -1;
  bool subset = false;
  int peer_ctas = 0;  // CTAs of the other groups' first stages launched beside this one (for the CTA spreading)
  int bounds[12] = {0, 2048};  // stage s evaluates the lags [bounds[s], bounds[s+1]); a stage may be narrowed
  int n_stage = 1;
  void plan(std::initializer_list<int> b) {
    n_stage = (int)b.size() - 1;
    std::copy(b.begin(), b.end(), bounds);
  }
};

struct EssScratch {
  double *act, *ess, *sums, *sls;
  EssState* state;
  int *used, *ids_base, *counter;
  size_t nd;
  int* ids(int slot, int k) const { return ids_base + ((size_t)slot * 2 + (size_t)k) * nd; }
  int* h_counter;  // pinned
};

// Thread shape of one stage: R lags per thread, groups of `gthreads` threads per (trace, lag block), CTAs of 128
// threads (one warp per SM sub-partition).  8 lags per thread halves the shared-memory traffic per flop and runs
// at the FP64 issue rate; fewer lags per thread give more, lighter warps for the thinly populated late stages.
struct EssShape {
  int R, gthreads, blocks_y, groups, ring, chunk;
};

// groups, grid rows and ring of a stage of `lags` lags ending at `lag_end`, R lags per thread
EssShape ess_shape_for(int R, int lags, int lag_end) {
  EssShape sh;
  sh.R = R;
  sh.gthreads = 128;  // the widest group whose lag blocks tile the stage
  while (sh.gthreads > 32 && (sh.gthreads * R > lags || lags % (sh.gthreads * R) != 0)) sh.gthreads /= 2;
  sh.blocks_y = lags / (sh.gthreads * R);
  sh.groups = std::max(1, 128 / sh.gthreads);
  auto ring_for = [&](int chunk) {
    int ring = 1024;
    while (ring < lag_end + 8 + 3 * chunk) ring *= 2;
    return ring;
  };
  sh.chunk = kChunkMax;
  sh.ring = ring_for(kChunkMax);
  if (ring_for(kChunkMax / 2) < sh.ring) sh.chunk = kChunkMax / 2, sh.ring = ring_for(kChunkMax / 2);
  return sh;
}

// Picks the shape of the stage that starts at lag `lag_begin` and is planned to end at `*lag_end`.  May move
// `*lag_end` down (see below); the caller evaluates [lag_begin, *lag_end) and leaves the rest to the next stage.
EssShape ess_pick_shape(const asm_ctx* ctx, int n_open, int lag_begin, int* lag_end, int stage, bool can_narrow, bool last) {
  const int lags = *lag_end - lag_begin;
  int R = 0, width = lags;
  if (const char* env = getenv("ASM_ESS_R")) {  // tuning aid: "R0,R1,R2,R3" per stage, 0 = automatic
    const char* p = env;
    for (int k = 0; k < stage && p; k++) {
      p = strchr(p, ',');
      if (p) p++;
    }
    if (p) R = atoi(p);
    if ((R != 8 && R != 4 && R != 2 && R != 1) || lags % (32 * R) != 0) R = 0;
  }
  if (R == 0) {
    // Cost model, milliseconds per million samples on the busiest SM sub-partition.  A warp with R lags per
    // thread issues 2R FP64 instructions per sample at one per 2.07 clocks (profiles/r01_microbench_pipes.json);
    // `eff` is the share of the FP64 issue rate it reaches (ncu, profiles/r01_ncu_ess_lag_padded_ring_summary.md).
    // A sub-partition runs ceil(warps / slots) warps.  With 1 or 2 lags per thread a warp is bound by the latency
    // of its dependent adds, not by issue: measured 8.0-8.4 ms per million samples however few warps run (one floor
    // for both, so that the 2-lag shape with half the warps and loads wins the tie).
    // Stages are short-lived, so the least modelled time wins; ties go to the larger R.
    //
    // Wave cliff: a warp stays on its sub-partition for the whole walk, so a stage with a few warps more than
    // slots takes twice as long as one with a few less.  When an R needs two or more waves for the planned
    // width but at least three quarters of that width fit in one, the narrower stage is a candidate too (at a small
    // handicap); the lags it drops are picked up by the next stage, which has fewer traces.
    static const int kR[4] = {8, 4, 2, 1};
    static const double kEff[4] = {0.85, 0.76, 0.70, 0.65};
    static const double kFloor[4] = {0.0, 0.0, 8.4, 8.4};
    const double slots = 4.0 * ctx->sm_count;
    double best = 0;
    for (int k = 0; k < 4; k++) {
      const int r = kR[k];
      if (lags % (32 * r) != 0) continue;
      const double warps = (double)n_open * lags / (32.0 * r);
      const double waves = std::ceil(warps / slots);
      const double t = std::max(kFloor[k], waves * r * 2.11 / kEff[k]);
      if (R == 0 || t < best - 1e-9) best = t, R = r, width = lags;
      if (can_narrow && waves >= 2 && n_open > 0) {
        const int one_wave = (int)(slots / n_open) * 32 * r;  // widest stage of this R that is a single wave
        if (one_wave >= 32 * r && 4 * one_wave >= 3 * lags) {
          // handicap: the dropped lags cost the next stage little, unless a stage has to be added for them
          const double t1 = std::max(kFloor[k], r * 2.11 / kEff[k]) + (last ? 9.0 : 1.0);
          if (t1 < best - 1e-9) best = t1, R = r, width = one_wave;
        }
      }
    }
  }
  *lag_end = lag_begin + width;
  return ess_shape_for(R, width, *lag_end);
}

int32_t ess_launch_stage(asm_ctx* ctx, EssGroup& g, const EssScratch& sc, const double* d_traces, int64_t stride,
                         int64_t n_samples, int32_t max_lag, int lags_stride) {
  const long long L = std::min<long long>(n_samples, max_lag);
  const int s = g.stage;
  const int b = g.bounds[s], e = g.bounds[s + 1];
  const double* tr = d_traces + (size_t)g.t0 * stride;
  double* sls = sc.sls + (size_t)g.t0 * lags_stride;
  if (s == 0) {
    if (g.ready) ASM_CUDA(ctx, cudaStreamWaitEvent(g.st, g.ready, 0));
    if (!g.subset) ASM_CUDA(ctx, cudaMemsetAsync(sc.state + g.t0, 0, sizeof(EssState) * (size_t)g.nt, g.st));
  }
  (void)L;
  if (n_samples > 0) {
    int e_fit = e;
    const bool last = s + 1 == g.n_stage;
    const bool room = g.n_stage + 2 <= (int)(sizeof(g.bounds) / sizeof(g.bounds[0]));
    const EssShape sh = ess_pick_shape(ctx, g.n_open, b, &e_fit, s, ctx->ess_adaptive && (!last || room), last);
    if (e_fit < e) {  // narrowed to a single wave: the next stage starts where this one stops
      if (last) {     // ... and if there was none, one is added for the rest
        g.bounds[g.n_stage + 1] = g.bounds[g.n_stage];
        g.n_stage++;
      }
      g.bounds[s + 1] = e_fit;
    }
    const int lags = g.bounds[s + 1] - b;
    const int R = sh.R, gthreads = sh.gthreads;
    dim3 grid((unsigned)((g.n_open + sh.groups - 1) / sh.groups), (unsigned)sh.blocks_y);
    // A CTA stays where it is placed for the whole walk over the samples, and the block scheduler is free to
    // stack the CTAs of a small grid on few SMs when other groups' kernels are resident.  Asking for just too
    // much shared memory for one more CTA than an even spread needs keeps them apart.
    const int per_sm = ((int)(grid.x * grid.y) + (s == 0 ? g.peer_ctas : 0) + ctx->sm_count - 1) / ctx->sm_count;
    const size_t smem_spread = std::min<size_t>((size_t)ctx->smem_optin, (size_t)ctx->smem_optin / (per_sm + 1) + 1024);
    LaunchScope ls(ctx, ASM_K_ESS_LAG, g.st);
#define ASM_ESS_LAUNCH(RR)                                                                                     \
  ess_lag_kernel<RR><<<grid, sh.groups * gthreads,                                                             \
                       std::max(smem_spread, (size_t)sh.groups * Ring<RR>::bytes(sh.ring, sh.chunk)), g.st>>>(  \
      tr, stride, n_samples, b, g.ids_in, g.n_open, lags_stride, sls, gthreads, sh.ring, sh.chunk,              \
      s == 0 ? sc.sums + g.t0 : nullptr)
    switch (R) {
      case 8: ASM_ESS_LAUNCH(8); break;
      case 4: ASM_ESS_LAUNCH(4); break;
      case 2: ASM_ESS_LAUNCH(2); break;
      default: ASM_ESS_LAUNCH(1); break;
    }
#undef ASM_ESS_LAUNCH
    ctx->ess_lag_slots += (long long)g.n_open * lags;
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ASM_OK;
}

}  // namespace

// Scratch of one batch: [act n][ess n][sums n][state n][used n (int)][work lists: 2 slots x 2 buffers x n (int)]
// [pilot lists: 2 x n (int)][counters 32 (int)]
static int32_t ess_scratch(asm_ctx* ctx, int32_t n_traces, EssScratch& sc) {
  const size_t nd = (size_t)n_traces;
  const size_t bytes = sizeof(double) * 3 * nd + sizeof(EssState) * nd + sizeof(int) * (7 * nd + 32);
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->ess_out, bytes)) != ASM_OK) return rc;
  sc.nd = nd;
  sc.act = (double*)ctx->ess_out.p;
  sc.ess = sc.act + nd;
  sc.sums = sc.ess + nd;
  sc.state = (EssState*)(sc.sums + nd);
  sc.used = (int*)(sc.state + nd);
  sc.ids_base = sc.used + nd;
  sc.counter = sc.used + 7 * nd;
  sc.sls = (double*)ctx->ess_sls.p;
  if (!ctx->h_pinned) ASM_CUDA(ctx, cudaMallocHost(&ctx->h_pinned, 256));
  sc.h_counter = (int*)ctx->h_pinned;
  return ASM_OK;
}

// Runs `n_groups` groups (each a trace range with its own stream) through the staged evaluation.
static int32_t ess_run_groups(asm_ctx* ctx, EssGroup* groups, int n_groups, int32_t n_traces, int64_t n_samples,
                              int64_t stride, const double* d_traces, int32_t max_lag, double* act_out_host,
                              double* ess_out_host) {
  const long long L = std::min<long long>(n_samples, max_lag);
  const int lags_stride = 2048;  // >= ASM_MAX_LAG rounded up to the widest stage
  int32_t rc;
  if ((rc = asm_reserve(ctx, ctx->ess_sls, sizeof(double) * (size_t)n_traces * lags_stride)) != ASM_OK) return rc;
  EssScratch sc;
  if ((rc = ess_scratch(ctx, n_traces, sc)) != ASM_OK) return rc;

  if (!(ctx->func_attr_done & kAttrEss)) {
    const int optin = ctx->smem_optin;
    ASM_CUDA(ctx, cudaFuncSetAttribute(ess_lag_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    ASM_CUDA(ctx, cudaFuncSetAttribute(ess_lag_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    ASM_CUDA(ctx, cudaFuncSetAttribute(ess_lag_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    ASM_CUDA(ctx, cudaFuncSetAttribute(ess_lag_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
    ctx->func_attr_done |= kAttrEss;
  }
  ctx->ess_lag_slots = 0;

  auto launch = [&](int gi) -> int32_t {
    EssGroup& g = groups[gi];
    int32_t r = ess_launch_stage(ctx, g, sc, d_traces, stride, n_samples, max_lag, lags_stride);
    if (r != ASM_OK) return r;
    const int e = g.bounds[g.stage + 1];
    ASM_CUDA(ctx, cudaMemsetAsync(sc.counter + gi, 0, sizeof(int), g.st));
    {
      LaunchScope ls(ctx, ASM_K_ESS_FINISH, g.st);
      ess_scan_kernel<<<(g.n_open + 127) / 128, 128, 0, g.st>>>(
          d_traces + (size_t)g.t0 * stride, stride, n_samples, g.n_open, g.ids_in, max_lag, e, lags_stride,
          sc.sls + (size_t)g.t0 * lags_stride, sc.sums + g.t0, sc.state + g.t0, sc.act + g.t0, sc.ess + g.t0,
          sc.used + g.t0, sc.ids(g.slot, g.stage & 1) + g.t0, sc.counter + gi);
    }
    ASM_CUDA(ctx, cudaGetLastError());
    ASM_CUDA(ctx, cudaMemcpyAsync(sc.h_counter + gi, sc.counter + gi, sizeof(int), cudaMemcpyDeviceToHost, g.st));
    return ASM_OK;
  };

  int remaining = 0;
  for (int gi = 0; gi < n_groups; gi++) {
    groups[gi].stage = 0;
    if (groups[gi].slot < 0) groups[gi].slot = 0;
    groups[gi].n_open = groups[gi].subset ? groups[gi].n0 : groups[gi].nt;
    groups[gi].ids_in = groups[gi].subset ? groups[gi].ids0 : nullptr;
    groups[gi].done = groups[gi].n_open == 0;
    if (!groups[gi].done) {
      if ((rc = launch(gi)) != ASM_OK) return rc;
      remaining++;
    }
  }
  // Whichever group finishes its stage first is served first: with several groups in flight the host polls their
  // streams instead of blocking on them in turn (a blocked host would hold back the short late stages of one
  // group behind the long early stage of another).
  const bool trace_on = getenv("ASM_ESS_TRACE") != nullptr;  // tuning aid: host-side timeline on stderr
  const auto t_begin = std::chrono::steady_clock::now();
  auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  bool seen_ready[2 * asm_ctx::kMaxGroups] = {};
  while (remaining > 0) {
    bool progressed = false;
    for (int gi = 0; gi < n_groups; gi++) {
      EssGroup& g = groups[gi];
      if (trace_on && g.ready && !seen_ready[gi] && cudaEventQuery(g.ready) == cudaSuccess) {
        seen_ready[gi] = true;
        fprintf(stderr, "[ess %8.2f ms] group %d (%d traces) uploaded\n", now_ms(), gi, g.nt);
      }
      if (g.done) continue;
      if (remaining == 1) {
        ASM_CUDA(ctx, cudaStreamSynchronize(g.st));  // nothing else to serve: block
      } else {
        const cudaError_t q = cudaStreamQuery(g.st);
        if (q == cudaErrorNotReady) continue;
        ASM_CUDA(ctx, q);
      }
      progressed = true;  // stage finished: the open count is on the host
      const int cnt = sc.h_counter[gi];
      if (trace_on)
        fprintf(stderr, "[ess %8.2f ms] group %d stage %d [%d,%d) done, %d of %d traces still open\n", now_ms(), gi, g.stage,
                g.bounds[g.stage], g.bounds[g.stage + 1], cnt, g.n_open);
      const int next = g.stage + 1;
      if (next < g.n_stage && cnt > 0 && g.bounds[next] < L) {
        g.ids_in = sc.ids(g.slot, g.stage & 1) + g.t0;
        g.n_open = cnt;
        g.stage = next;
        if ((rc = launch(gi)) != ASM_OK) return rc;
      } else {
        g.done = true;
        remaining--;
        if (g.subset) continue;  // the two halves of a range: the caller copies the range out once both are done
        if (act_out_host)
          ASM_CUDA(ctx, cudaMemcpyAsync(act_out_host + g.t0, sc.act + g.t0, sizeof(double) * (size_t)g.nt,
                                        cudaMemcpyDeviceToHost, g.st));
        if (ess_out_host)
          ASM_CUDA(ctx, cudaMemcpyAsync(ess_out_host + g.t0, sc.ess + g.t0, sizeof(double) * (size_t)g.nt,
                                        cudaMemcpyDeviceToHost, g.st));
      }
    }
    if (!progressed) std::this_thread::sleep_for(std::chrono::microseconds(20));
  }
  for (int gi = 0; gi < n_groups; gi++) ASM_CUDA(ctx, cudaStreamSynchronize(groups[gi].st));
  return ASM_OK;
}

static void ess_env_bounds(const char* name, EssGroup& g) {  // tuning aid: "0,128,256,...,2048"
  const char* env = getenv(name);
  if (!env) return;
  int k = 0;
  for (const char* q = env; q && *q && k < 10; k++) {
    g.bounds[k] = atoi(q);
    q = strchr(q, ',');
    if (q) q++;
  }
  if (k >= 2 && g.bounds[0] == 0 && g.bounds[k - 1] == 2048) g.n_stage = k - 1;
}

// ---- ticket rounds (resident path) ------------------------------------------------------------------------------
// One ROUND = one ess_ticket_kernel launch that covers the current stage of every group still open (at most two: the
// "short" and "long" halves of the pilot split, or one group), then one ess_scan_kernel per group, then one host
// synchronisation for the open counts.

// dynamic shared memory a ticket CTA may ask for: the opt-in maximum less the kernel's static ticket slots
static int ess_ticket_smem(const asm_ctx* ctx) { return ctx->smem_optin - 1024; }

struct EssTicketShape {
  int R[2] = {0, 0}, gthreads = 0, chunk = 0, ring = 0, n_ctas = 0;
  size_t ring_bytes = 0;
  double model_ms = 0;
};

static size_t ess_ring_bytes(int R, int max_lag_end, int* ring_out, int* chunk_out) {
  int chunk = kChunkMax;
  auto ring_for = [&](int ck) {
    int ring = 1024;
    while (ring < max_lag_end + 2 * R + 16 + 3 * ck) ring *= 2;
    return ring;
  };
  int ring = ring_for(kChunkMax);
  if (ring_for(kChunkMax / 2) < ring) chunk = kChunkMax / 2, ring = ring_for(kChunkMax / 2);
  if (const char* e = getenv("ASM_ESS_CHUNK")) {  // tuning knob: force the cp.async chunk (256 or 512 samples)
    const int c = atoi(e);
    if (c == kChunkMax || c == kChunkMax / 2) chunk = c, ring = ring_for(c);
  }
  *ring_out = ring;
  *chunk_out = chunk;
  switch (R) {
    case 8: return Ring<8>::bytes((uint32_t)ring, (uint32_t)chunk);
    case 4: return Ring<4>::bytes((uint32_t)ring, (uint32_t)chunk);
    case 2: return Ring<2>::bytes((uint32_t)ring, (uint32_t)chunk);
    default: return Ring<1>::bytes((uint32_t)ring, (uint32_t)chunk);
  }
}

// Shape of one round: workers of `gthreads` threads, R[a] lags per thread for the jobs of set a, `n_ctas` CTAs of four
// warps (one per sub-partition).  A chain is sequential, so at most one worker per job is busy at any time: the grid is
// never larger than the number of jobs, and it is a whole number of CTAs per SM so that every sub-partition carries the
// same number of busy warps.  Model (milliseconds per million samples): the FP64 issue time of the round over all
// sub-partitions at the efficiency each R reaches (ncu, profiles/; less with a single warp per sub-partition, which
// cannot hide the latency of its own dependent adds), but never less than one job's own walk (2R FP64 instructions per
// sample at one per 2.07 clocks, or the 8.4-clock latency of its adds).
static EssTicketShape ess_ticket_shape(const asm_ctx* ctx, const int* n_open, const int* width, int n_sets, int max_lag_end) {
  static const int kR[4] = {8, 4, 2, 1};
  static const double kEff[4] = {0.90, 0.76, 0.68, 0.62};
  EssTicketShape best;
  int forced_R = 0, forced_g = 0;
  if (const char* env = getenv("ASM_ESS_TICKET_R")) forced_R = atoi(env);
  if (const char* env = getenv("ASM_ESS_TICKET_G")) forced_g = atoi(env);
  const double clk_ghz = 1.9;
  const bool mixed = getenv("ASM_ESS_TICKET_MIXED") != nullptr;
  for (int g = 128; g >= 32; g /= 2) {
    if (forced_g && g != forced_g) continue;
    for (int k0 = 0; k0 < 4; k0++)
      for (int k1 = 0; k1 < (n_sets > 1 ? 4 : 1); k1++) {
        const int ks[2] = {k0, k1};
        // Measured (cfg3, profiles/r02_ess_ticket_rounds.md): with one ring per warp a round that mixes 8 and 4 lags per
        // thread is slower than the same round at 4 throughout, so both sets take the same R unless asked otherwise.
        if (n_sets > 1 && k1 != k0 && !mixed) continue;
        bool ok = true;
        long long jobs = 0;
        double issue = 0, chain = 0;
        size_t ring_bytes = 0;
        EssTicketShape sh;
        sh.gthreads = g;
        for (int a = 0; a < n_sets && ok; a++) {
          const int R = kR[ks[a]];
          if (forced_R && R != forced_R) ok = false;
          if (width[a] % (g * R) != 0) ok = false;
          if (!ok) break;
          sh.R[a] = R;
          jobs += (long long)n_open[a] * (width[a] / (g * R));
          issue += (double)n_open[a] * width[a] * 2.0 / 32.0 * 2.07 / kEff[ks[a]];
          chain = std::max(chain, std::max(2.0 * R * 2.07, 8.4));
          int ring, chunk;
          ring_bytes = std::max(ring_bytes, ess_ring_bytes(R, max_lag_end, &ring, &chunk));
          sh.ring = ring;  // the same for every R: it depends on max_lag_end (+ 2R)
          sh.chunk = chunk;
        }
        if (!ok || jobs == 0) continue;
        sh.ring_bytes = ring_bytes;
        const int groups = 128 / g;
        const int fit = (int)std::min<size_t>(8, (size_t)ess_ticket_smem(ctx) / (groups * ring_bytes + 1024));
        if (fit < 1) continue;
        long long ctas = std::min<long long>((long long)ctx->sm_count * fit, (jobs + groups - 1) / groups);
        if (ctas > ctx->sm_count) ctas -= ctas % ctx->sm_count;
        sh.n_ctas = (int)ctas;
        const double wps = (double)ctas / ctx->sm_count;  // busy warps per sub-partition: every worker always has a job
        const double hide = wps >= 2 ? 1.0 : (wps >= 1 ? 0.85 : 0.85 * wps);
        sh.model_ms = std::max(issue / (4.0 * ctx->sm_count) / hide, chain) / clk_ghz;
        if (best.gthreads == 0 || sh.model_ms < best.model_ms - 1e-9) best = sh;
      }
  }
  return best;
}

static void ess_ticket_launch(asm_ctx* ctx, const EssTicketParams& prm, const EssTicketShape& sh, cudaStream_t st) {
  const int groups = 128 / sh.gthreads;
  // spread a grid of fewer CTAs than fit over the SMs instead of stacking it: ask for just too much shared memory for
  // one more CTA per SM than an even spread needs
  const int per_sm = (sh.n_ctas + ctx->sm_count - 1) / ctx->sm_count;
  const size_t smem = std::max<size_t>(groups * sh.ring_bytes,
                                       std::min<size_t>((size_t)ess_ticket_smem(ctx), (size_t)ess_ticket_smem(ctx) / (per_sm + 1) + 1024));
  LaunchScope ls(ctx, ASM_K_ESS_LAG, st);
  ess_ticket_kernel<<<(unsigned)sh.n_ctas, 128, smem, st>>>(prm);
}

static int32_t ess_run_rounds(asm_ctx* ctx, EssGroup* groups, int n_groups, int32_t n_traces, int64_t n_samples,
                              int64_t stride, const double* d_traces, int32_t max_lag, const EssScratch& sc) {
  const long long L = std::min<long long>(n_samples, max_lag);
  const int lags_stride = 2048;
  cudaStream_t st = ctx->stream;
  int32_t rc;
  if (!(ctx->func_attr_done & kAttrEssTicket)) {
    ASM_CUDA(ctx, cudaFuncSetAttribute(ess_ticket_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ess_ticket_smem(ctx)));
    ctx->func_attr_done |= kAttrEssTicket;
  }
  ctx->ess_lag_slots = 0;
  for (int gi = 0; gi < n_groups; gi++) {
    EssGroup& g = groups[gi];
    g.stage = 0;
    if (g.slot < 0) g.slot = gi;
    g.n_open = g.subset ? g.n0 : g.nt;
    g.ids_in = g.subset ? g.ids0 : nullptr;
    g.done = g.n_open == 0;
  }
  const bool trace_on = getenv("ASM_ESS_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto now_ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count(); };
  const int seg_len = 65536;
  const long long nR = (n_samples + 7) / 8 * 8;
  const int n_seg = (int)std::max<long long>(1, (nR + seg_len - 1) / seg_len);
  for (int round = 0;; round++) {
    int act[2], n_act = 0, n_open[2], width[2], max_lag_end = 0;
    for (int gi = 0; gi < n_groups; gi++)
      if (!groups[gi].done) {
        const EssGroup& g = groups[gi];
        n_open[n_act] = g.n_open;
        width[n_act] = g.bounds[g.stage + 1] - g.bounds[g.stage];
        max_lag_end = std::max(max_lag_end, g.bounds[g.stage + 1]);
        act[n_act++] = gi;
      }
    if (n_act == 0) break;
    const EssTicketShape sh = ess_ticket_shape(ctx, n_open, width, n_act, max_lag_end);
    if (sh.gthreads == 0) return asm_fail(ctx, ASM_E_UNSUPPORTED, "ess: no thread shape tiles the stages of this round");
    EssTicketParams prm{};
    prm.traces = d_traces;
    prm.stride = stride;
    prm.n = n_samples;
    prm.seg_len = seg_len;
    prm.n_seg = n_seg;
    prm.lags_stride = lags_stride;
    prm.sls = sc.sls;
    prm.sums = sc.sums;
    prm.gthreads = sh.gthreads;
    prm.ring_size = sh.ring;
    prm.chunk = sh.chunk;
    prm.ring_stride = sh.ring_bytes;
    int first = 0;
    size_t park_doubles = 0;
    for (int a = 0; a < n_act; a++) {
      const EssGroup& g = groups[act[a]];
      EssJobSet& js = prm.set[a];
      js.ids = g.ids_in;
      js.n = g.n_open;
      js.lag_begin = g.bounds[g.stage];
      js.R = sh.R[a];
      js.n_blocks = width[a] / (sh.gthreads * js.R);
      js.first_job = first;
      js.with_sum = g.stage == 0 ? 1 : 0;
      js.park_off = park_doubles;
      park_doubles += (size_t)js.n * js.n_blocks * ((size_t)sh.gthreads * js.R + 1);
      first += js.n * js.n_blocks;
      ctx->ess_lag_slots += (long long)g.n_open * width[a];
    }
    prm.n_sets = n_act;
    prm.n_jobs = first;
    if (n_samples > 0 && prm.n_jobs > 0) {
      if ((rc = asm_reserve(ctx, ctx->ess_park, sizeof(double) * park_doubles + sizeof(int) * (size_t)prm.n_jobs + 64)) != ASM_OK) return rc;
      prm.park = (double*)ctx->ess_park.p;
      prm.progress = (int*)(prm.park + park_doubles);
      prm.ticket_counter = (unsigned long long*)(((uintptr_t)(prm.progress + prm.n_jobs) + 15) & ~(uintptr_t)15);
      ASM_CUDA(ctx, cudaMemsetAsync(prm.progress, 0, sizeof(int) * (size_t)prm.n_jobs + 48, st));
      ess_ticket_launch(ctx, prm, sh, st);
      ASM_CUDA(ctx, cudaGetLastError());
    }
    for (int a = 0; a < n_act; a++) {
      const int gi = act[a];
      EssGroup& g = groups[gi];
      ASM_CUDA(ctx, cudaMemsetAsync(sc.counter + gi, 0, sizeof(int), st));
      {
        LaunchScope ls(ctx, ASM_K_ESS_FINISH, st);
        ess_scan_kernel<<<(g.n_open + 127) / 128, 128, 0, st>>>(d_traces, stride, n_samples, g.n_open, g.ids_in, max_lag,
                                                               g.bounds[g.stage + 1], lags_stride, sc.sls, sc.sums, sc.state,
                                                               sc.act, sc.ess, sc.used, sc.ids(g.slot, g.stage & 1),
                                                               sc.counter + gi);
      }
      ASM_CUDA(ctx, cudaGetLastError());
      ASM_CUDA(ctx, cudaMemcpyAsync(sc.h_counter + gi, sc.counter + gi, sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    ASM_CUDA(ctx, cudaStreamSynchronize(st));
    for (int a = 0; a < n_act; a++) {
      const int gi = act[a];
      EssGroup& g = groups[gi];
      const int cnt = sc.h_counter[gi];
      if (trace_on)
        fprintf(stderr, "[ess %8.2f ms] round %d (R = %d x %d threads, %d CTAs, model %.1f): group %d stage %d [%d,%d) done, %d of %d traces still open\n",
                now_ms(), round, sh.R[a], sh.gthreads, sh.n_ctas, sh.model_ms, gi, g.stage, g.bounds[g.stage], g.bounds[g.stage + 1], cnt,
                g.n_open);
      const int next = g.stage + 1;
      if (next < g.n_stage && cnt > 0 && g.bounds[next] < L) {
        g.ids_in = sc.ids(g.slot, g.stage & 1);
        g.n_open = cnt;
        g.stage = next;
      } else {
        g.done = true;
      }
    }
  }
  return ASM_OK;
}

int32_t ess_batch_device(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* d_traces,
                         int32_t max_lag, double* act_out_host, double* ess_out_host) {
  EssGroup g[2];
  g[0].t0 = 0;
  g[0].nt = n_traces;
  g[0].st = ctx->stream;
  g[0].st2 = ctx->stream2;
  g[0].fork = ctx->ev_fork;
  g[0].join = ctx->ev_join;
  if (ctx->ess_adaptive) g[0].plan({0, 128, 256, 768, 1792, 2048});
  ess_env_bounds("ASM_ESS_BOUNDS", g[0]);
  // Resident path: pilot split + ticket rounds (ess_run_rounds).  ASM_ESS_TICKETS=0 / ASM_ESS_PILOT=0 fall back to one
  // whole-trace warp group per lag block (ess_lag_kernel, the kernels the host-pointer entry overlaps with its upload).
  // The pilot split is what the rounds are for; without it (staging switched off, short traces, few traces, tuning
  // knobs) the whole-trace kernels run as before.
  const char* te = getenv("ASM_ESS_TICKETS");
  const char* pe = getenv("ASM_ESS_PILOT");
  const bool pilot = ctx->ess_adaptive && !(te && te[0] == '0') && !(pe && pe[0] == '0') && n_traces >= 8 && n_samples >= 65536 &&
                     std::min<int64_t>(n_samples, max_lag) >= 1024 && !getenv("ASM_ESS_BOUNDS");
  if (!pilot) return ess_run_groups(ctx, g, 1, n_traces, n_samples, stride, d_traces, max_lag, act_out_host, ess_out_host);
  int32_t rc;
  const int lags_stride = 2048;
  if ((rc = asm_reserve(ctx, ctx->ess_sls, sizeof(double) * (size_t)n_traces * lags_stride)) != ASM_OK) return rc;
  EssScratch sc;
  if ((rc = ess_scratch(ctx, n_traces, sc)) != ASM_OK) return rc;
  int n_groups = 1;
  // Pilot split.  With the staged plan a slowly decaying trace walks through every stage, and each thin stage costs at
  // least one latency-bound walk of a chain over all samples while most of the GPU idles.  A cheap estimate of the
  // lag-128 autocorrelation (first 128k samples) picks those traces out beforehand; they form a second group whose
  // first stage covers [0, 768) at once, in the SAME launch as the first stage of the others.  A wrong guess only costs
  // time: a "short" trace that is still open after a stage simply goes on to the next one, a "long" one that stops
  // early has computed lags it did not need.  Results do not depend on the split.
  {
    int* ids_short = sc.ids_base + 4 * sc.nd;
    int* ids_long = sc.ids_base + 5 * sc.nd;
    int* counts = sc.counter + 16;
    ASM_CUDA(ctx, cudaMemsetAsync(counts, 0, 2 * sizeof(int), ctx->stream));
    {
      LaunchScope ls(ctx, ASM_K_ESS_FINISH);
      const int P = (int)std::min<int64_t>(n_samples, 131072);
      const double thr = getenv("ASM_ESS_PILOT_THRESHOLD") ? atof(getenv("ASM_ESS_PILOT_THRESHOLD")) : 0.15;
      ess_pilot_kernel<<<n_traces, 256, 0, ctx->stream>>>(d_traces, stride, n_traces, P, 128, thr, ids_short, ids_long, counts);
    }
    ASM_CUDA(ctx, cudaGetLastError());
    ASM_CUDA(ctx, cudaMemcpyAsync(sc.h_counter + 16, counts, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int n_short = sc.h_counter[16], n_long = sc.h_counter[17];
    if (getenv("ASM_ESS_TRACE")) fprintf(stderr, "[ess] pilot: %d short, %d long traces\n", n_short, n_long);
    g[0].subset = true;
    g[0].ids0 = ids_short;
    g[0].n0 = n_short;
    g[0].slot = 0;
    g[1] = g[0];
    g[1].ids0 = ids_long;
    g[1].n0 = n_long;
    g[1].slot = 1;
    if (n_long == 0) {  // nothing to split off: the staged plan with the whole-trace kernels, over the whole range
      g[0].subset = false;
      g[0].ids0 = nullptr;
      return ess_run_groups(ctx, g, 1, n_traces, n_samples, stride, d_traces, max_lag, act_out_host, ess_out_host);
    }
    g[1].plan({0, 768, 2048});
    ess_env_bounds("ASM_ESS_LONG_BOUNDS", g[1]);
    n_groups = 2;
  }
  ASM_CUDA(ctx, cudaMemsetAsync(sc.state, 0, sizeof(EssState) * (size_t)n_traces, ctx->stream));
  if ((rc = ess_run_rounds(ctx, g, n_groups, n_traces, n_samples, stride, d_traces, max_lag, sc)) != ASM_OK) return rc;
  if (act_out_host)
    ASM_CUDA(ctx, cudaMemcpyAsync(act_out_host, sc.act, sizeof(double) * (size_t)n_traces, cudaMemcpyDeviceToHost, ctx->stream));
  if (ess_out_host)
    ASM_CUDA(ctx, cudaMemcpyAsync(ess_out_host, sc.ess, sizeof(double) * (size_t)n_traces, cudaMemcpyDeviceToHost, ctx->stream));
  ASM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return ASM_OK;
}

// Host-pointer entry: the traces are uploaded in up to four chunks on the copy stream, each chunk is a group on
// its own stream that starts as soon as its upload has landed.
int32_t ess_batch_host(asm_ctx* ctx, int32_t n_traces, int64_t n_samples, int64_t stride, const double* traces,
                       int64_t n_pad, double* d_tr, int32_t max_lag, double* act_out_host, double* ess_out_host) {
  int n_groups = 1;
  if (n_traces >= 512 && (int64_t)n_traces * n_pad * 8 >= (1LL << 30)) n_groups = 4;
  if (const char* env = getenv("ASM_ESS_H2D_GROUPS")) {  // tuning aid
    const int v = atoi(env);
    if (v >= 1 && v <= asm_ctx::kMaxGroups) n_groups = v;
  }
  EssGroup groups[asm_ctx::kMaxGroups];
  // Chunk plan.  The upload (PCIe) is the long pole, so the chunks that land last are on the critical path.  Every
  // stage costs at least one walk of its slowest warp over all samples, however few traces it has, so the last
  // two chunks are small and use two wide stages (twice the lag sums of the five-stage plan, but for few traces);
  // the early chunks are large and get the full five-stage plan, whose latency hides under the rest of the upload.
  int sizes[asm_ctx::kMaxGroups];
  const bool tapered = n_groups == 4 && !getenv("ASM_ESS_H2D_GROUPS") && ctx->ess_adaptive;
  if (tapered) {
    const int one_wave = std::max(8, (4 * ctx->sm_count / 8) & ~7);
    sizes[3] = std::min(one_wave, n_traces / 8);
    sizes[2] = std::min(2 * one_wave, n_traces / 4);
    sizes[1] = n_traces / 4;
    sizes[0] = n_traces - sizes[1] - sizes[2] - sizes[3];
  } else {
    const int chunk = (n_traces + n_groups - 1) / n_groups;
    for (int k = 0; k < n_groups; k++) sizes[k] = std::max(0, std::min(chunk, n_traces - k * chunk));
  }
  int codes[asm_ctx::kMaxGroups];
  for (int k = 0; k < n_groups; k++) codes[k] = (tapered && k >= 2) ? 2 : 5;
  if (const char* env = getenv("ASM_ESS_H2D_PLAN")) {  // tuning aid: "traces:stages,traces:stages,..." (last takes the rest)
    int k = 0, used = 0;
    for (const char* q = env; q && *q && k < asm_ctx::kMaxGroups; k++) {
      int m = 0, c = 5;
      if (sscanf(q, "%d:%d", &m, &c) < 1) break;
      sizes[k] = std::max(0, std::min(m, n_traces - used));
      codes[k] = c;
      used += sizes[k];
      q = strchr(q, ',');
      if (q) q++;
    }
    if (k > 0) {
      sizes[k - 1] += n_traces - used;
      n_groups = k;
    }
  }
  int t0 = 0;
  for (int k = 0; k < n_groups; k++) {
    EssGroup& g = groups[k];
    g.t0 = t0;
    g.nt = sizes[k];
    t0 += sizes[k];
    if (ctx->ess_adaptive) {
      switch (codes[k]) {
        case 1: g.plan({0, 2048}); break;
        case 2: g.plan({0, 256, 2048}); break;
        case 3: g.plan({0, 128, 512, 2048}); break;
        default: g.plan({0, 128, 256, 768, 1792, 2048}); break;
      }
    }
    g.st = ctx->group_stream[k];
    g.st2 = ctx->group_stream2[k];
    g.ready = ctx->group_event[k];
    g.fork = ctx->group_fork[k];
    g.join = ctx->group_join[k];
    if (g.nt > 0 && n_samples > 0)
      ASM_CUDA(ctx, cudaMemcpy2DAsync(d_tr + (size_t)g.t0 * n_pad, sizeof(double) * (size_t)n_pad, traces + (size_t)g.t0 * stride,
                                      sizeof(double) * (size_t)stride, sizeof(double) * (size_t)n_samples, (size_t)g.nt,
                                      cudaMemcpyHostToDevice, ctx->stream_copy));
    ASM_CUDA(ctx, cudaEventRecord(g.ready, ctx->stream_copy));
  }
  int32_t rc = ess_run_groups(ctx, groups, n_groups, n_traces, n_samples, n_pad, d_tr, max_lag, act_out_host, ess_out_host);
  cudaStreamSynchronize(ctx->stream_copy);
  return rc;
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
  if (total > 0) {
    LaunchScope ls(ctx, ASM_K_MISC);
    int blocks = (int)std::min<int64_t>((total * n_cols + 255) / 256, (int64_t)ctx->sm_count * 8);
    concat_kernel<<<blocks, 256, 0, ctx->stream>>>(prm, (double*)ctx->ess_tr.p);
  }
  ASM_CUDA(ctx, cudaGetLastError());
  return ess_batch_device(ctx, n_cols, total, prm.stride, (const double*)ctx->ess_tr.p, ASM_MAX_LAG, nullptr, ess_out_host);
}
