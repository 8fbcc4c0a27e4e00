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

static double walked_sum(const std::vector<double>& v, int threads, int per_thread, int prelude, int plain_run, bool windowed) {
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

static int g_fail = 0, g_cases = 0;
static void check(const char* name, const std::vector<double>& v) {
  const double want = plain_sum(v);
  // the kernel's own shape first (with its look-ahead window), then shapes that stress the fold and the tiling
  const int shapes[][5] = {{kThreads, kPerThread, kPrelude, kPlainRun, 1}, {256, 8, 64, 32, 0}, {4, 2, 0, 1, 0}, {32, 8, 3, 5, 0},
                           {1, 1, 0, 1, 0}};
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
  std::printf("%s: %d cases, %d failures; last run took %ld integer-path adds, %ld plain adds\n", g_fail ? "FAILED" : "ok", g_cases,
              g_fail, g_integer_adds, g_plain_adds);
  // share of plain adds on a PSRF-like row of 28,285 values
  {
    std::vector<double> v(28285);
    for (auto& x : v) x = std::sqrt((double)(1000000 + rng() % 400000) / (double)(1000000 + rng() % 400000));
    g_integer_adds = g_plain_adds = 0;
    walked_sum(v, kThreads, kPerThread, kPrelude, kPlainRun, true);
    std::printf("psrf-like n=28285: %ld integer-path adds, %ld plain adds\n", g_integer_adds, g_plain_adds);
  }
  return g_fail ? 1 : 0;
}
