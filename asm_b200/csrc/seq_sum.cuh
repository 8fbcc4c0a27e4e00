// seq_sum.cuh -- the left-to-right double sum  s = (((0 + v0) + v1) + ...)  evaluated in parallel, bit for bit.
//
// TreePSRF.mean (TreePSRF.java:264-271) adds the n row values in order, one rounding per add, and the result must
// be the Java's to the last bit, so the order cannot change.  A dependent DADD costs 8.3 clocks on B200
// (profiles/r01_microbench_pipes.json): 10,000 rows are 42 us of one idle thread, and the chain grows with the
// number of trees while everything around it scales.
//
// What makes the chain splittable: while the running sum s stays inside one binade [2^e, 2^(e+1)) it is an integer
// multiple M of u = 2^(e-52), and adding a value 0 <= v < 2^(e+1) is integer arithmetic,
//     v = q*u + r (0 <= r < u)   ->   M' = M + q + (r > u/2  or  (r == u/2 and M+q odd)),      round-to-nearest-even
// i.e. a step depends on the running sum only through ITS PARITY, and only for exact ties.  A run of elements is
// therefore a pair (d0, d1) = ulps added when the run starts on an even / odd M; two runs compose as
//     (a then b).d_p = a.d_p + b.d_{p xor (a.d_p & 1)}
// which is associative: threads fold their own elements, an ordered reduction combines the threads.  A run ends
// where a conservative bound M + sum(q_k + 1) could reach 2^53 (the sum would leave the binade), or at an element
// that is negative, non-finite or >= 2^(e+1); that element is added with a plain DADD and the walk restarts in the
// new binade.  Binade changes are rare (log2 of the sum), so nearly all adds take the parallel path.
//
// Two layers on the device (block_sum): that walk, one block-wide pass per binade, correct on any input; and in front
// of it plan_tile, which predicts the binade before every element from a parallel prefix sum, folds a whole tile in one
// pass and lets thread 0 stitch the pieces with the exact running sum -- every piece is checked against the exact sum
// before it is applied, and whatever fails the check is left to the walk.  tests/c/seq_sum_check.cpp walks both layers
// on the host with this header's arithmetic.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define SEQ_HD __host__ __device__ __forceinline__
#else
#define SEQ_HD inline
#endif

namespace seqsum {

struct Seg {  // ulps a run of elements adds to a running sum that starts even (d0) / odd (d1)
  unsigned long long d0, d1;
};

SEQ_HD Seg then(Seg a, Seg b) {
  Seg r;
  r.d0 = a.d0 + ((a.d0 & 1ull) ? b.d1 : b.d0);
  r.d1 = a.d1 + ((a.d1 & 1ull) ? b.d0 : b.d1);  // started odd: still odd after `a` iff a.d1 is even
  return r;
}

SEQ_HD unsigned long long bits_of(double v) {
#if defined(__CUDA_ARCH__)
  return (unsigned long long)__double_as_longlong(v);
#else
  unsigned long long b;
  std::memcpy(&b, &v, sizeof b);
  return b;
#endif
}
SEQ_HD double double_of(unsigned long long b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)b);
#else
  double v;
  std::memcpy(&v, &b, sizeof v);
  return v;
#endif
}

constexpr unsigned long long kTwo52 = 1ull << 52, kTwo53 = 1ull << 53, kBoundCap = 1ull << 54;
constexpr int kMinExp = -900;  // binades below stay on the plain path (u = 2^(e-52) must be far from subnormal)

// Can the running sum s take the integer path?  (positive, normal, not tiny.)  On success e = its exponent and
// M = s / 2^(e-52) in [2^52, 2^53).
SEQ_HD bool binade_of(double s, int* e, unsigned long long* M) {
  const unsigned long long b = bits_of(s);
  const int ef = (int)((b >> 52) & 0x7ff);
  if ((b >> 63) || ef == 0 || ef == 0x7ff) return false;
  *e = ef - 1023;
  if (*e < kMinExp) return false;
  *M = (b & (kTwo52 - 1)) | kTwo52;
  return true;
}
SEQ_HD double from_binade(int e, unsigned long long M) {  // M in [2^52, 2^53)
  return double_of(((unsigned long long)(e + 1023) << 52) | (M & (kTwo52 - 1)));
}

// One element against a running sum in binade e.  false: it needs the plain add.  Otherwise seg = what it does to
// M, bound = an upper bound of that (q + 1, or 0 for an element that cannot change M).  Branch-free: eight of these
// per thread run interleaved.
SEQ_HD bool element(double v, int e, Seg* seg, unsigned long long* bound) {
  const unsigned long long b = bits_of(v);
  const int ef = (int)((b >> 52) & 0x7ff);
  const unsigned long long frac = b & (kTwo52 - 1);
  const bool zero = (b << 1) == 0;                       // +0.0 and -0.0 leave a positive sum as it is
  const bool bad = (b >> 63) != 0 || ef == 0x7ff;        // negative, infinite, NaN
  const unsigned long long m = ef ? (frac | kTwo52) : frac;
  const int ek = ef ? ef - 1023 : -1022;
  const int d = e - ek;                                  // v = m * 2^-d ulps of the running sum
  const bool small = zero || d >= 64;                    // far below half an ulp (m < 2^53)
  const int dc = d < 0 ? 0 : (d > 63 ? 63 : d);
  const unsigned long long q = m >> dc;
  const unsigned long long low = dc ? (m << (64 - dc)) : 0ull;  // the bits shifted out, left-aligned
  const unsigned long long half = 1ull << 63;
  const unsigned long long up = low > half ? 1ull : 0ull, tie = low == half ? 1ull : 0ull;
  seg->d0 = small ? 0ull : q + (up | (tie & q & 1ull));          // exact tie: to the even neighbour
  seg->d1 = small ? 0ull : q + (up | (tie & ~q & 1ull));
  *bound = small ? 0ull : q + 1;
  return zero || (!bad && d >= 0);                       // d < 0: the element is above the binade
}

SEQ_HD unsigned long long sat_add(unsigned long long a, unsigned long long b) {
  const unsigned long long s = a + b;  // both <= kBoundCap: no wrap
  return s > kBoundCap ? kBoundCap : s;
}

// ---- shape of the block-wide walk (shared with the host check, tests/c/seq_sum_check.cpp)
constexpr int kThreads = 512, kPerThread = 4, kTile = kThreads * kPerThread, kWarps = kThreads / 32;
constexpr int kPrelude = 64;   // the first adds cross a binade almost every time: plain
constexpr int kPlainRun = 32;  // after a run that made little progress, this many plain adds before trying again
constexpr int kMinWindow = 256;
// thread t's 4 values sit 5 doubles apart in shared memory (odd stride: no bank conflicts wherever the run starts)
SEQ_HD int pad(int k) { return k + (k >> 2); }
// A run is cut where the sum may leave its binade, and with values of similar size that happens when the number of
// elements added so far has doubled: looking further ahead than that decomposes elements against the wrong binade.
SEQ_HD int window_of(int consumed) {
  return consumed < kMinWindow ? kMinWindow : (consumed > kTile ? kTile : consumed);
}
// ---- the planned pass (plan_tile): one thread's 4 consecutive elements form a group; a group is either a RUN group
// (its elements are decomposed against the binade the running sum is predicted to be in) or a PLAIN group (added with
// DADDs).  Consecutive run groups of one binade form an event, every plain group is an event of its own.
constexpr int kMaxEvents = 16;       // more events than this in a tile: the tile is left to the walk
constexpr int kPlainCls = 1 << 20;   // class of a plain group (a run group's class is its binade exponent)
constexpr int kParts = kMaxEvents + kWarps;  // (event, warp) pieces of the run events: one per lane of the warp that adds them up
static_assert(kParts <= 32, "one piece per lane");

}  // namespace seqsum

#if defined(__CUDACC__)
namespace seqsum {

struct Shared {
  double vals[kTile + kTile / 4 + 4];
  unsigned long long warp_u64[kWarps];
  Seg warp_seg[kWarps];
  int warp_cut[kWarps];
  double s;
  int lo;
  // plan_tile
  double warp_sum[kWarps];
  int warp_cls[kWarps], warp_heads[kWarps];
  int ev_first[kMaxEvents], ev_cls[kMaxEvents];
  Seg part_seg[kParts];
  int part_ev[kParts];
  Seg ev_tot[kMaxEvents];
};

// The planned pass over the elements [lo, tn) of the tile in shared memory, all threads.
//
// The walk below (block_sum's loop) learns where the running sum changes binade by running into it, one block-wide
// pass per binade.  But a sum of doubles in ANY order is within ~n ulps of the left-to-right one, so a parallel prefix
// sum predicts the binade of the running sum before every element, wrong only when the sum passes within ~1e-12 of a
// power of two.  So: (1) approximate prefix sums B (before a thread's 4 elements) and A (after them); (2) a group
// whose B and A share a binade e is a run group and is decomposed against e, any other group is plain; (3) run
// groups are folded per event with ONE segmented warp scan, the pieces (event, warp) land in shared memory in
// order; (4) thread 0 stitches the events in order with the EXACT running sum: a run event is applied only if the
// exact sum is in the predicted binade and stays in it (M + d < 2^53; the partial sums are monotone, so then every
// element was decomposed against the right binade), a plain group is 4 DADDs.  The prediction is never trusted:
// the first event that fails the check ends the pass, sh.lo points at it, and the walk takes over from there.
// Tiles with negative or non-finite values, or with more than kMaxEvents events, are left to the walk entirely.
__device__ __forceinline__ void plan_tile(Shared& sh, int tn) {
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lo = sh.lo;
  const double s = sh.s;
  if (lo >= tn) return;  // uniform
  const int first = lo + tid * kPerThread;
  const bool active = first < tn;
  double v[kPerThread];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < kPerThread; i++) {
    const int k = first + i;
    v[i] = k < tn ? sh.vals[pad(k)] : 0.0;
    const unsigned long long b = bits_of(v[i]);
    bad = bad || (((b >> 63) != 0 && (b << 1) != 0) || ((b >> 52) & 0x7ff) == 0x7ff);  // negative (not -0.0), infinite, NaN
  }
  if (tid < kParts) sh.part_ev[tid] = -1;
  if (__syncthreads_or(bad)) return;
  // (1) approximate running sum before / after this thread's elements
  double a = v[0];
#pragma unroll
  for (int i = 1; i < kPerThread; i++) a += v[i];
  double incl = a;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) sh.warp_sum[warp] = incl;
  __syncthreads();
  double B = s;
  for (int w = 0; w < warp; w++) B += sh.warp_sum[w];
  B += incl - a;
  const double A = B + a;
  // (2) class of the group and, for a run group, the fold of its elements
  int cls = kPlainCls;
  Seg mine{0, 0};
  {
    int e = 0, eA = 0;
    unsigned long long M = 0, MA = 0;
    if (binade_of(B, &e, &M) && binade_of(A, &eA, &MA) && e == eA) {
      bool ok = true;
#pragma unroll
      for (int i = 0; i < kPerThread; i++) {
        Seg sg;
        unsigned long long bnd;
        ok = element(v[i], e, &sg, &bnd) && ok;
        mine = then(mine, sg);
      }
      if (ok) cls = e;
    }
  }
  // (3) events: a group starts one if it is plain or its class differs from the group before it
  if (lane == 31) sh.warp_cls[warp] = cls;
  __syncthreads();
  int prev_cls = __shfl_up_sync(0xffffffffu, cls, 1);
  if (lane == 0) prev_cls = warp ? sh.warp_cls[warp - 1] : kPlainCls;
  const bool head = active && (cls == kPlainCls || cls != prev_cls);
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  if (lane == 0) sh.warp_heads[warp] = __popc(heads);
  __syncthreads();
  int ev = -1, n_events = 0;
  for (int w = 0; w < kWarps; w++) {
    if (w == warp) ev = n_events;
    n_events += sh.warp_heads[w];
  }
  if (n_events > kMaxEvents) return;  // uniform
  ev += __popc(heads & (0xffffffffu >> (31 - lane))) - 1;  // index of the event this group belongs to
  if (head) {
    sh.ev_first[ev] = tid;
    sh.ev_cls[ev] = cls;
  }
  // segmented fold within the warp: event indices ascend with the lane, so "same event" = no head in between
  const int key = active ? ev : -2 - lane;  // inactive groups: keys of their own, never written
  if (cls == kPlainCls) mine = Seg{0, 0};
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    Seg other;
    other.d0 = __shfl_up_sync(0xffffffffu, mine.d0, o);
    other.d1 = __shfl_up_sync(0xffffffffu, mine.d1, o);
    const int okey = __shfl_up_sync(0xffffffffu, key, o);
    if (lane >= o && okey == key) mine = then(other, mine);
  }
  const int next_key = __shfl_down_sync(0xffffffffu, key, 1);
  if (active && cls != kPlainCls && (lane == 31 || next_key != key)) {  // the last group of this event in this warp
    sh.part_seg[ev + warp] = mine;  // (event, warp) pieces in order: ev + warp ascends strictly from piece to piece
    sh.part_ev[ev + warp] = ev;
  }
  __syncthreads();
  // the pieces of one event sit in consecutive slots: warp 0 adds them up in order, one slot per lane, with the same
  // segmented scan, so that the stitch reads one total per event
  if (warp == 0) {
    const int pkey = lane < kParts ? sh.part_ev[lane] : -1;
    Seg tot = pkey >= 0 ? sh.part_seg[lane] : Seg{0, 0};
    const int skey = pkey >= 0 ? pkey : -2 - lane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      Seg other;
      other.d0 = __shfl_up_sync(0xffffffffu, tot.d0, o);
      other.d1 = __shfl_up_sync(0xffffffffu, tot.d1, o);
      const int okey = __shfl_up_sync(0xffffffffu, skey, o);
      if (lane >= o && okey == skey) tot = then(other, tot);
    }
    const int nkey = __shfl_down_sync(0xffffffffu, skey, 1);
    if (pkey >= 0 && (lane == 31 || nkey != skey)) sh.ev_tot[pkey] = tot;
    __syncwarp();
  }
  // (4) stitch with the exact running sum
  if (tid == 0) {
    double s1 = s;
    int done = tn;
    for (int j = 0; j < n_events; j++) {
      const int c = sh.ev_cls[j], g = sh.ev_first[j];
      if (c == kPlainCls) {
#pragma unroll
        for (int i = 0; i < kPerThread; i++) {
          const int k = lo + g * kPerThread + i;
          if (k < tn) s1 = __dadd_rn(s1, sh.vals[pad(k)]);
        }
        continue;
      }
      const Seg tot = sh.ev_tot[j];
      int e = 0;
      unsigned long long M = 0;
      const bool in_binade = binade_of(s1, &e, &M) && e == c;
      const unsigned long long d = (M & 1ull) ? tot.d1 : tot.d0;
      if (!in_binade || d >= kTwo53 || M + d >= kTwo53) {  // the prediction was off: the walk takes over here
        done = lo + g * kPerThread;
        break;
      }
      s1 = from_binade(e, M + d);
    }
    sh.s = s1;
    sh.lo = done;
  }
  __syncthreads();
}

// Sum of value(0) .. value(n-1) in that order, every add rounded as a scalar loop would; all kThreads threads of
// the block call it, the result is returned to every thread.  value(k) is evaluated once per element (by some
// thread, a tile ahead of its use) and may have side effects for that k.
template <class Value>
__device__ double block_sum(Shared& sh, int n, Value&& value, bool planned = true) {
  const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    sh.s = 0.0;
    sh.lo = 0;
  }
  double ahead[kPerThread];  // this thread's share of the next tile, in flight while the current one is summed
#pragma unroll
  for (int j = 0; j < kPerThread; j++) {
    const int k = j * kThreads + tid;
    ahead[j] = k < n ? value(k) : 0.0;
  }
  for (int t0 = 0; t0 < n; t0 += kTile) {
    const int tn = min(kTile, n - t0);
    __syncthreads();  // the previous tile's values are consumed
#pragma unroll
    for (int j = 0; j < kPerThread; j++) {
      const int k = j * kThreads + tid;
      if (k < tn) sh.vals[pad(k)] = ahead[j];
    }
#pragma unroll
    for (int j = 0; j < kPerThread; j++) {
      const int k = t0 + kTile + j * kThreads + tid;
      ahead[j] = k < n ? value(k) : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
      double s = sh.s;
      int lo = 0;
      if (t0 == 0)
        for (; lo < min(kPrelude, tn); lo++) s = __dadd_rn(s, sh.vals[pad(lo)]);
      sh.s = s;
      sh.lo = lo;
    }
    __syncthreads();
    if (planned) plan_tile(sh, tn);  // normally consumes the whole tile; the walk below takes what it leaves
    for (;;) {
      const int lo = sh.lo;  // uniform
      if (lo >= tn) break;
      const double s = sh.s;
      int e = 0;
      unsigned long long M = 0;
      const bool integer_path = binade_of(s, &e, &M);  // uniform
      int cut = lo;  // first element of [lo, wn) that does not go the integer way
      Seg mine{0, 0};
      if (integer_path) {
        const int wn = min(tn, lo + window_of(t0 + lo));
        // this thread's elements: lo + tid*4 .. +3, i.e. the run starts at `lo`, not at a multiple of 4
        const int first = lo + tid * kPerThread;
        Seg segs[kPerThread];
        unsigned long long bnd[kPerThread];
        bool ok[kPerThread];
        unsigned long long local = 0;
#pragma unroll
        for (int i = 0; i < kPerThread; i++) {
          const int k = first + i;
          const double v = k < wn ? sh.vals[pad(k)] : 0.0;  // +0.0: no change, no bound, never a cut
          ok[i] = element(v, e, &segs[i], &bnd[i]);
          local = sat_add(local, bnd[i]);  // bnd <= 2^53
        }
        // exclusive prefix of the bounds over the threads; saturating adds of non-negative terms = min(sum, cap)
        unsigned long long incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl = sat_add(incl, up);
        }
        const unsigned long long prev = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 31) sh.warp_u64[warp] = incl;
        __syncthreads();
        unsigned long long run = lane ? prev : 0ull;
        for (int w = 0; w < warp; w++) run = sat_add(run, sh.warp_u64[w]);
        int my_cut = wn;
#pragma unroll
        for (int i = 0; i < kPerThread; i++) {
          const int k = first + i;
          run = sat_add(run, bnd[i]);
          if (k < my_cut && (!ok[i] || M + run >= kTwo53)) my_cut = k;  // k ascends: the first hit stays
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_cut = min(my_cut, __shfl_xor_sync(0xffffffffu, my_cut, o));
        if (lane == 0) sh.warp_cut[warp] = my_cut;
        __syncthreads();
        cut = wn;
#pragma unroll
        for (int w = 0; w < kWarps; w++) cut = min(cut, sh.warp_cut[w]);
        // fold this thread's elements below the cut, then the threads in order
#pragma unroll
        for (int i = 0; i < kPerThread; i++)
          if (first + i < cut) mine = then(mine, segs[i]);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          Seg other;
          other.d0 = __shfl_down_sync(0xffffffffu, mine.d0, o);
          other.d1 = __shfl_down_sync(0xffffffffu, mine.d1, o);
          if (lane + o < 32) mine = then(mine, other);  // lanes that are not a multiple of 2*o go unused from here on
        }
        if (lane == 0) sh.warp_seg[warp] = mine;
      }
      __syncthreads();
      if (tid == 0) {
        double s1 = s;
        int next = cut;
        if (integer_path && cut > lo) {
          Seg all = sh.warp_seg[0];
          const int used = (cut - lo + 32 * kPerThread - 1) / (32 * kPerThread);  // warps with elements below the cut
          for (int w = 1; w < used; w++) all = then(all, sh.warp_seg[w]);
          s1 = from_binade(e, M + ((M & 1ull) ? all.d1 : all.d0));
        }
        // the element that ended the run, and more of them if the run was short (negative values, a sum that is
        // still tiny): plain adds
        const int plain = (cut - lo < 16) ? kPlainRun : 1;
        for (int i = 0; i < plain && next < tn; i++, next++) s1 = __dadd_rn(s1, sh.vals[pad(next)]);
        sh.s = s1;
        sh.lo = next;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  return sh.s;
}

}  // namespace seqsum
#endif
