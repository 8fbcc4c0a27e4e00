// seq_sum_check.cpp -- host check of asm_b200/csrc/seq_sum.cuh: the binade / parity arithmetic that lets the GPU
// evaluate TreePSRF.mean's left-to-right sum in parallel, against the plain loop, on inputs chosen to hit every
// branch (exact ties in both parities, zeros, subnormals, binade crossings on every add, huge and tiny values,
// negative / non-finite values that must fall back to the plain add).  Test infrastructure; no GPU.
//
// The walk below is the one the kernel does (prelude, tiles, look-ahead window, conservative cut, ordered fold, plain adds
// after a short run) with its threads folded one after the other; the arithmetic is the header's own.
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../asm_b200/csrc/seq_sum.cuh"

using namespace seqsum;

static double plain_sum(const std::vector<double>& v) {
  volatile double s = 0.0;  // one rounding per add, no reassociation
  for (double x : v) s = s + x;
  return s;
}

static long g_integer_adds = 0, g_plain_adds = 0;

static double walked_sum(const std::vector<double>& v, int threads, int per_thread, int prelude, int plain_run, bool windowed,
                         int planned = 0, std::mt19937_64* jitter = nullptr);
static int planned_tile(const double* vals, int lo, int tn, double& s_io, std::mt19937_64* jitter);

static double walked_sum(const std::vector<double>& v, int threads, int per_thread, int prelude, int plain_run, bool windowed,
                         int planned, std::mt19937_64* jitter) {
  const int n = (int)v.size(), tile = threads * per_thread;
  double s = 0.0;
  for (int t0 = 0; t0 < n; t0 += tile) {
    const int tn = std::min(tile, n - t0);
    const double* vals = v.data() + t0;
    int lo = 0;
    if (t0 == 0)
      for (; lo < std::min(prelude, tn); lo++) {
        volatile double t = s + vals[lo];
        s = t;
        g_plain_adds++;
      }
    if (planned) lo = planned_tile(vals, lo, tn, s, jitter);
    while (lo < tn) {
      int e = 0;
      unsigned long long M = 0;
      const bool integer_path = binade_of(s, &e, &M);
      int cut = lo;
      if (integer_path) {
        // per thread: bounds and cut, with the exclusive prefix of the earlier threads
        unsigned long long run = 0;
        const int wn = windowed ? std::min(tn, lo + window_of(t0 + lo)) : tn;  // the kernel looks ahead this far
        cut = wn;
        for (int k = lo; k < wn; k++) {
          Seg sg;
          unsigned long long b;
          const bool ok = element(vals[k], e, &sg, &b);
          run = sat_add(run, b);
          if (!ok || M + run >= kTwo53) {
            cut = k;
            break;
          }
        }
        // fold per thread, then the threads in order (a tree would give the same: `then` is associative; checked below)
        Seg all{0, 0};
        for (int t = 0; t < threads; t++) {
          Seg mine{0, 0};
          for (int i = 0; i < per_thread; i++) {
            const int k = lo + t * per_thread + i;
            if (k < cut) {
              Seg sg;
              unsigned long long b;
              element(vals[k], e, &sg, &b);
              mine = then(mine, sg);
            }
          }
          all = then(all, mine);
        }
        if (cut > lo) {
          s = from_binade(e, M + ((M & 1ull) ? all.d1 : all.d0));
          g_integer_adds += cut - lo;
        }
      }
      int next = cut;
      const int plain = (cut - lo < 16) ? plain_run : 1;
      for (int i = 0; i < plain && next < tn; i++, next++) {
        volatile double t = s + vals[next];
        s = t;
        g_plain_adds++;
      }
      lo = next;
    }
  }
  return s;
}

// The planned pass of the kernel (plan_tile) over [lo, tn) of one tile, its 512 threads and 16 warps walked one after
// the other with the kernel's own order of operations.  `jitter` perturbs the approximate prefix sums (0: none), to
// drive the prediction wrong on purpose: the result must not depend on it, only the share of work left to the walk.
// Returns the new lo; s is updated.
static long g_planned_adds = 0, g_plan_fallbacks = 0;
static int planned_tile(const double* vals, int lo, int tn, double& s_io, std::mt19937_64* jitter) {
  const int T = kThreads, W = kWarps;
  const double s = s_io;
  if (lo >= tn) return lo;
  std::vector<double> a(T), incl(T), B(T), A(T);
  std::vector<int> cls(T), ev(T, -1);
  std::vector<Seg> mine(T, Seg{0, 0});
  std::vector<char> active(T), head(T);
  bool bad = false;
  for (int t = 0; t < T; t++) {
    const int first = lo + t * kPerThread;
    active[t] = first < tn;
    double acc = 0;
    for (int i = 0; i < kPerThread; i++) {
      const int k = first + i;
      const double v = k < tn ? vals[k] : 0.0;
      const unsigned long long b = bits_of(v);
      bad = bad || (((b >> 63) != 0 && (b << 1) != 0) || ((b >> 52) & 0x7ff) == 0x7ff);
      acc = i ? acc + v : v;
    }
    a[t] = acc;
  }
  if (bad) return lo;
  std::vector<double> warp_sum(W);
  for (int w = 0; w < W; w++) {  // Hillis-Steele inclusive scan of a warp's 32 lanes
    double x[32];
    for (int l = 0; l < 32; l++) x[l] = a[w * 32 + l];
    for (int o = 1; o < 32; o <<= 1) {
      double y[32];
      for (int l = 0; l < 32; l++) y[l] = l >= o ? x[l] + x[l - o] : x[l];
      for (int l = 0; l < 32; l++) x[l] = y[l];
    }
    for (int l = 0; l < 32; l++) incl[w * 32 + l] = x[l];
    warp_sum[w] = x[31];
  }
  for (int t = 0; t < T; t++) {
    double b = s;
    for (int w = 0; w < t / 32; w++) b += warp_sum[w];
    b += incl[t] - a[t];
    B[t] = b;
    A[t] = b + a[t];
    if (jitter) {  // a wrong prediction now and then
      const unsigned r = (unsigned)((*jitter)() & 1023);
      if (r < 8) B[t] *= (r & 1) ? 0.49 : 2.1;
      else if (r < 16) A[t] *= (r & 1) ? 0.49 : 2.1;
      else if (r < 20) B[t] *= 2.0, A[t] *= 2.0;   // the whole group predicted one binade up / down: the stitch must notice
      else if (r < 24) B[t] *= 0.5, A[t] *= 0.5;
      else if (r < 48) B[t] *= 1.0 + ((r & 1) ? 1e-9 : -1e-9), A[t] = B[t] + a[t];
    }
  }
  for (int t = 0; t < T; t++) {
    cls[t] = kPlainCls;
    int e = 0, eA = 0;
    unsigned long long M = 0, MA = 0;
    if (binade_of(B[t], &e, &M) && binade_of(A[t], &eA, &MA) && e == eA) {
      bool ok = true;
      Seg m{0, 0};
      for (int i = 0; i < kPerThread; i++) {
        const int k = lo + t * kPerThread + i;
        Seg sg;
        unsigned long long bnd;
        ok = element(k < tn ? vals[k] : 0.0, e, &sg, &bnd) && ok;
        m = then(m, sg);
      }
      if (ok) cls[t] = e, mine[t] = m;
    }
  }
  int n_events = 0;
  for (int t = 0; t < T; t++) {
    const int prev = t ? cls[t - 1] : kPlainCls;
    head[t] = active[t] && (cls[t] == kPlainCls || cls[t] != prev);
    if (head[t]) n_events++;
    ev[t] = n_events - 1;
  }
  if (n_events > kMaxEvents) return lo;
  std::vector<int> ev_first(n_events), ev_cls(n_events);
  for (int t = 0; t < T; t++)
    if (head[t]) ev_first[ev[t]] = t, ev_cls[ev[t]] = cls[t];
  // pieces (event, warp): segmented Hillis-Steele scan per warp, the last lane of an event in a warp writes slot ev + warp
  std::vector<Seg> part_seg(kParts, Seg{0, 0});
  std::vector<int> part_ev(kParts, -1);
  for (int w = 0; w < W; w++) {
    Seg x[32];
    int key[32];
    for (int l = 0; l < 32; l++) {
      const int t = w * 32 + l;
      key[l] = active[t] ? ev[t] : -2 - l;
      x[l] = cls[t] == kPlainCls ? Seg{0, 0} : mine[t];
    }
    for (int o = 1; o < 32; o <<= 1) {
      Seg y[32];
      for (int l = 0; l < 32; l++) y[l] = (l >= o && key[l - o] == key[l]) ? then(x[l - o], x[l]) : x[l];
      for (int l = 0; l < 32; l++) x[l] = y[l];
    }
    for (int l = 0; l < 32; l++) {
      const int t = w * 32 + l;
      if (active[t] && cls[t] != kPlainCls && (l == 31 || key[l + 1] != key[l])) {
        if (part_ev[ev[t] + w] != -1) std::printf("FAIL piece slot reused\n"), std::exit(1);
        part_seg[ev[t] + w] = x[l];
        part_ev[ev[t] + w] = ev[t];
      }
    }
  }
  // stitch
  double s1 = s;
  int slot = 0, done = tn;
  for (int j = 0; j < n_events; j++) {
    const int c = ev_cls[j], g = ev_first[j];
    if (c == kPlainCls) {
      for (int i = 0; i < kPerThread; i++) {
        const int k = lo + g * kPerThread + i;
        if (k < tn) {
          volatile double t = s1 + vals[k];
          s1 = t;
          g_plain_adds++;
        }
      }
      continue;
    }
    Seg tot{0, 0};
    while (slot < kParts && part_ev[slot] < j) slot++;
    while (slot < kParts && part_ev[slot] == j) tot = then(tot, part_seg[slot++]);
    int e = 0;
    unsigned long long M = 0;
    const bool in_binade = binade_of(s1, &e, &M) && e == c;
    const unsigned long long d = (M & 1ull) ? tot.d1 : tot.d0;
    if (!in_binade || d >= kTwo53 || M + d >= kTwo53) {
      done = lo + g * kPerThread;
      g_plan_fallbacks++;
      break;
    }
    s1 = from_binade(e, M + d);
    const int g_end = j + 1 < n_events ? ev_first[j + 1] : T;
    g_planned_adds += std::min(tn, lo + g_end * kPerThread) - (lo + g * kPerThread);
  }
  s_io = s1;
  return done;
}

static int g_fail = 0, g_cases = 0;
static void check(const char* name, const std::vector<double>& v) {
  const double want = plain_sum(v);
  // the kernel's own shape first (with its look-ahead window), then shapes that stress the fold and the tiling
  const int shapes[][5] = {{kThreads, kPerThread, kPrelude, kPlainRun, 1}, {256, 8, 64, 32, 0}, {4, 2, 0, 1, 0}, {32, 8, 3, 5, 0},
                           {1, 1, 0, 1, 0}};
  // the kernel as shipped (planned pass, then the walk for what it leaves), and the same with the prediction jittered
  static std::mt19937_64 jit(777);
  for (int mode = 0; mode < 3; mode++) {
    const double got = walked_sum(v, kThreads, kPerThread, kPrelude, kPlainRun, true, 1, mode ? &jit : nullptr);
    g_cases++;
    if (bits_of(got) != bits_of(want) && !(got != got && want != want)) {
      g_fail++;
      std::printf("FAIL %s (planned, jitter %d): got %.17g want %.17g\n", name, mode, got, want);
    }
  }
  for (const auto& sh : shapes) {
    const double got = walked_sum(v, sh[0], sh[1], sh[2], sh[3], sh[4] != 0);
    g_cases++;
    if (bits_of(got) != bits_of(want) && !(got != got && want != want)) {
      g_fail++;
      std::printf("FAIL %s (threads %d x %d): got %.17g (%016" PRIx64 ") want %.17g (%016" PRIx64 ")\n", name, sh[0], sh[1], got,
                  (uint64_t)bits_of(got), want, (uint64_t)bits_of(want));
    }
  }
}

int main() {
  std::mt19937_64 rng(20240611);
  auto uni = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
  // `then` is associative
  for (int it = 0; it < 200000; it++) {
    auto seg = [&]() {
      Seg s;
      s.d0 = rng() >> 12;
      s.d1 = (rng() & 3) ? s.d0 : s.d0 + (rng() & 1 ? 1 : -(long long)(s.d0 > 0));
      return s;
    };
    const Seg a = seg(), b = seg(), c = seg();
    const Seg l = then(then(a, b), c), r = then(a, then(b, c));
    if (l.d0 != r.d0 || l.d1 != r.d1) {
      g_fail++;
      std::printf("FAIL associativity\n");
      break;
    }
  }
  // PSRF-like rows: sqrt of ratios of integers near 1
  for (int n : {1, 2, 63, 64, 65, 100, 2047, 2048, 2049, 10000, 28285, 50000}) {
    std::vector<double> v((size_t)n);
    for (auto& x : v) x = std::sqrt((double)(1000000 + rng() % 400000) / (double)(1000000 + rng() % 400000));
    check("psrf-like", v);
    for (auto& x : v) x = std::sqrt((double)(rng() % 5000000000ull));  // the varIn == 0 branch: sqrt of an integer
    check("sqrt-int", v);
  }
  // exact ties: values with few mantissa bits against sums of every parity
  for (int rep = 0; rep < 200; rep++) {
    const int n = 200 + (int)(rng() % 6000);
    std::vector<double> v((size_t)n);
    const int e0 = (int)(rng() % 40) - 20;
    for (auto& x : v) {
      const int kind = (int)(rng() % 6);
      const double small = std::ldexp(1.0, e0 - 40 - (int)(rng() % 16));   // around half an ulp of the running sum
      if (kind == 0) x = std::ldexp((double)(1 + rng() % 7), e0);
      else if (kind == 1) x = small;
      else if (kind == 2) x = 3 * small;
      else if (kind == 3) x = std::ldexp(1.0, e0) + small;
      else if (kind == 4) x = 0.0;
      else x = std::ldexp((double)(rng() % 1024), e0 - 10);
    }
    check("ties", v);
  }
  // the tie positions exactly: s = 2^52 * k region, adding 0.5, 1.5, 2.5 ulps
  for (int rep = 0; rep < 100; rep++) {
    std::vector<double> v;
    for (int i = 0; i < 70; i++) v.push_back(std::ldexp(1.0, 46));  // prelude: reach 2^52 + ...
    for (int i = 0; i < 5000; i++) {
      const int r = (int)(rng() % 8);
      v.push_back(r == 0 ? 0.5 : r == 1 ? 1.5 : r == 2 ? 2.5 : r == 3 ? 1.0 : r == 4 ? 0.25 : r == 5 ? 0.75 : r == 6 ? 3.0 : 0.5000000000000001);
    }
    check("half-ulp", v);
  }
  // wide dynamic range, zeros, subnormals, tiny sums
  for (int rep = 0; rep < 100; rep++) {
    const int n = 1 + (int)(rng() % 9000);
    std::vector<double> v((size_t)n);
    for (auto& x : v) {
      const int kind = (int)(rng() % 10);
      if (kind == 0) x = 0.0;
      else if (kind == 1) x = -0.0;
      else if (kind == 2) x = std::ldexp(uni(1, 2), -1060 - (int)(rng() % 14));  // subnormal
      else x = std::ldexp(uni(1, 2), (int)(rng() % 120) - 60);
    }
    check("wide", v);
    for (auto& x : v) x = std::ldexp(x, -1000);  // everything tiny: mostly below the integer path's floor
    check("tiny", v);
    for (auto& x : v) x = std::ldexp(x, 1000 + 960);  // towards overflow
    check("huge", v);
  }
  // values that must take the plain add: negative, NaN, infinite
  for (int rep = 0; rep < 100; rep++) {
    const int n = 1 + (int)(rng() % 5000);
    std::vector<double> v((size_t)n);
    for (auto& x : v) {
      x = uni(0, 3);
      const int kind = (int)(rng() % 50);
      if (kind == 0) x = -x;
      if (kind == 1) x = -1e3 * x;
    }
    check("signed", v);
    if (rep % 10 == 0) {
      v[(size_t)(rng() % n)] = INFINITY;
      check("inf", v);
      v[(size_t)(rng() % n)] = NAN;
      check("nan", v);
    }
  }
  // constant rows (every add the same value: binade crossings at exact powers of two)
  for (double c : {1.0, 0.5, 3.0, 1.0 / 3, 1e-3, 1024.0}) {
    std::vector<double> v(40000, c);
    check("constant", v);
  }
  std::printf("%s: %d cases, %d failures; last run took %ld integer-path adds, %ld plain adds; planned pass: %ld adds, %ld hand-overs to the walk\n",
              g_fail ? "FAILED" : "ok", g_cases, g_fail, g_integer_adds, g_plain_adds, g_planned_adds, g_plan_fallbacks);
  // share of plain adds on a PSRF-like row of 28,285 values
  {
    std::vector<double> v(28285);
    for (auto& x : v) x = std::sqrt((double)(1000000 + rng() % 400000) / (double)(1000000 + rng() % 400000));
    g_integer_adds = g_plain_adds = g_planned_adds = g_plan_fallbacks = 0;
    walked_sum(v, kThreads, kPerThread, kPrelude, kPlainRun, true, 1);
    std::printf("psrf-like n=28285: %ld integer-path adds, %ld plain adds (planned pass: %ld adds, %ld fallbacks to the walk)\n",
                g_integer_adds + g_planned_adds, g_plain_adds, g_planned_adds, g_plan_fallbacks);
  }
  return g_fail ? 1 : 0;
}
