// host_encoder.cpp -- trees -> clades -> dictionary -> bit rows.  CPU only; part of libasmgpu.so but
// needs no GPU.
//
// Clade identity follows the reference: the set of leaf numbers under an internal node, root
// included (CladeDifference.java:107-141).  Instead of the reference's "3,7,12," key strings a
// clade is keyed by a 128-bit Zobrist hash (XOR of per-leaf random words) and every dictionary
// hit is confirmed against the stored leaf list, so the mapping is exact, not probabilistic.
#include <algorithm>
#include <atomic>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "host_encoder.h"

// Post-order walk (left subtree, right subtree, node) without recursion limits: emits one clade id
// per internal node.  Returns false if the arrays do not describe a binary tree on n_taxa leaves.
// `resolve(key, leaf set, position in the post-order) -> id` names a clade: the serial path inserts into the dictionary,
// the parallel path looks it up read-only and records a miss.
template <class Resolve>
static bool encode_tree_with(const asm_encoder* enc, const int32_t* left, const int32_t* right, int32_t root, int32_t* ids_out,
                             Resolve&& resolve) {
  const int n = enc->n_taxa, nn = 2 * n - 1, W = enc->set_words;
  if (root < 0 || root >= nn) return false;
  struct Frame {
    int32_t node, stage;
  };
  // scratch kept across calls (one encoder = one caller thread)
  static thread_local std::vector<Frame> st;
  static thread_local std::vector<uint64_t> sets;  // leaf sets of completed subtrees, stacked (W words each)
  static thread_local std::vector<int32_t> sizes;
  static thread_local std::vector<Key128> keys;
  st.clear();
  sets.clear();
  sizes.clear();
  keys.clear();
  int n_out = 0, visited = 0;
  st.push_back({root, 0});
  while (!st.empty()) {
    Frame& f = st.back();
    const int32_t v = f.node;
    if (v < 0 || v >= nn || ++visited > 4 * nn) return false;
    if (left[v] < 0) {  // leaf: getNr() == node index, must be < n_taxa
      if (v >= n) return false;
      sets.resize(sets.size() + (size_t)W, 0);
      sets[sets.size() - (size_t)W + (size_t)(v >> 6)] = 1ULL << (v & 63);
      sizes.push_back(1);
      keys.push_back(enc->zobrist[(size_t)v]);
      st.pop_back();
      continue;
    }
    if (f.stage == 0) {
      f.stage = 1;
      const int32_t c = left[v];
      st.push_back({c, 0});
      continue;
    }
    if (f.stage == 1) {
      f.stage = 2;
      const int32_t c = right[v];
      st.push_back({c, 0});
      continue;
    }
    // both children done: the node's leaf set is the union of the two sets on top of the value stack
    if (sizes.size() < 2) return false;
    const int32_t nr = sizes.back();
    sizes.pop_back();
    const int32_t nl = sizes.back();
    sizes.pop_back();
    const Key128 kr = keys.back();
    keys.pop_back();
    const Key128 kl = keys.back();
    keys.pop_back();
    uint64_t* l = sets.data() + (sets.size() - 2 * (size_t)W);
    const uint64_t* r = l + W;
    uint64_t overlap = 0;
    for (int w = 0; w < W; w++) {
      overlap |= l[w] & r[w];
      l[w] |= r[w];
    }
    if (overlap) return false;  // a leaf reached twice: not a tree
    sets.resize(sets.size() - (size_t)W);
    const Key128 k{kl.a ^ kr.a, kl.b ^ kr.b};
    if (n_out >= n - 1) return false;
    ids_out[n_out] = resolve(k, sets.data() + (sets.size() - (size_t)W), n_out);
    n_out++;
    sizes.push_back(nl + nr);
    keys.push_back(k);
    st.pop_back();
  }
  return n_out == n - 1 && sizes.size() == 1 && sizes[0] == n;
}

bool asm_encode_tree(asm_encoder* enc, const int32_t* left, const int32_t* right, int32_t root, int32_t* ids_out) {
  return encode_tree_with(enc, left, right, root, ids_out,
                          [enc](const Key128& k, const uint64_t* set, int) { return enc->get_or_add_set(k, set); });
}

namespace {

int encoder_threads(int64_t n_trees) {
  int t = (int)std::thread::hardware_concurrency();
  if (const char* e = getenv("ASM_ENCODER_THREADS")) t = atoi(e);
  if (t < 1) t = 1;
  if (t > 64) t = 64;
  if (n_trees < 512) t = 1;  // the live runner adds one tree per chain per check
  return t;
}

// Two-phase batch encoder.  Phase 1 (parallel over the trees of a batch, dictionary frozen): walk every tree, look each
// clade up read-only -- the exact look-up of the serial path, key + stored leaf set -- and record the clades that are
// not in the dictionary yet.  Phase 2 (serial, tree order, post-order inside a tree): insert the recorded clades.  A
// clade gets the id the serial encoder would have given it: ids of known clades do not change, new ones are numbered
// in the same first-seen order.  `walk(t, ids_out, resolve)` encodes tree t of the batch.
template <class Walk>
int32_t encode_batch(asm_encoder* enc, int64_t n_trees, int32_t* ids_out, int threads, Walk&& walk) {
  const int64_t k = enc->n_taxa - 1;
  const int W = enc->set_words;
  if (threads <= 1) {
    for (int64_t t = 0; t < n_trees; t++)
      if (!walk(t, ids_out + t * k, [enc](const Key128& key, const uint64_t* set, int) { return enc->get_or_add_set(key, set); }))
        return ASM_E_INVALID;
    return ASM_OK;
  }
  constexpr int64_t kBatch = 8192;  // largest batch; the first ones are small (see below)
  struct PerThread {
    std::vector<asm_encoder_miss> misses;
    std::vector<uint64_t> pool;
  };
  std::vector<PerThread> per((size_t)threads);
  std::vector<int32_t> owner((size_t)kBatch), first((size_t)kBatch), count((size_t)kBatch);
  // Batches grow geometrically: the first trees of a run meet an empty dictionary, so nearly all of their clades are
  // misses and go through the serial phase; a short first batch warms the dictionary up for the next, longer one.
  int64_t batch = 8 * threads;
  for (int64_t b0 = 0, nb = 0; b0 < n_trees; b0 += nb, batch = std::min(kBatch, batch * 2)) {
    nb = std::min(batch, n_trees - b0);
    std::atomic<int64_t> next{0};
    std::atomic<int> bad{0};
    for (auto& p : per) {
      p.misses.clear();
      p.pool.clear();
    }
    auto work = [&](int tid) {
      PerThread& me = per[(size_t)tid];
      for (;;) {
        const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
        if (i >= nb || bad.load(std::memory_order_relaxed)) break;
        owner[(size_t)i] = tid;
        first[(size_t)i] = (int32_t)me.misses.size();
        const bool ok = walk(b0 + i, ids_out + (b0 + i) * k, [&](const Key128& key, const uint64_t* set, int pos) {
          const int32_t id = enc->find_set(key, set);
          if (id >= 0) return id;
          me.misses.push_back(asm_encoder_miss{pos, key, me.pool.size()});
          me.pool.insert(me.pool.end(), set, set + W);
          return (int32_t)-1;
        });
        count[(size_t)i] = (int32_t)me.misses.size() - first[(size_t)i];
        if (!ok) bad.store(1, std::memory_order_relaxed);
      }
    };
    std::vector<std::thread> pool;
    for (int tid = 1; tid < threads; tid++) pool.emplace_back(work, tid);
    work(0);
    for (auto& th : pool) th.join();
    if (bad.load()) return ASM_E_INVALID;
    for (int64_t i = 0; i < nb; i++) {  // phase 2: first-seen order
      const PerThread& o = per[(size_t)owner[(size_t)i]];
      for (int32_t m = 0; m < count[(size_t)i]; m++) {
        const asm_encoder_miss& ms = o.misses[(size_t)(first[(size_t)i] + m)];
        ids_out[(b0 + i) * k + ms.pos] = enc->get_or_add_set(ms.key, o.pool.data() + ms.set_off);
      }
    }
  }
  return ASM_OK;
}

}  // namespace

extern "C" {

int32_t asm_encoder_create(asm_encoder** out, int32_t n_taxa) {
  if (!out || n_taxa < 2) return ASM_E_INVALID;
  auto* e = new (std::nothrow) asm_encoder();
  if (!e) return ASM_E_NOMEM;
  e->n_taxa = n_taxa;
  e->set_words = (n_taxa + 63) / 64;
  e->zobrist.resize((size_t)n_taxa);
  uint64_t s = 0x5EEDC1ADE5ULL;
  for (auto& z : e->zobrist) {
    z.a = splitmix64(s);
    z.b = splitmix64(s);
  }
  *out = e;
  return ASM_OK;
}

int32_t asm_encoder_destroy(asm_encoder* enc) {
  delete enc;
  return ASM_OK;
}

int32_t asm_encoder_add_topologies(asm_encoder* enc, int64_t n_trees, const int32_t* left, const int32_t* right,
                                   const int32_t* root, int32_t* clade_ids_out) {
  if (!enc || n_trees < 0 || !left || !right || !root || !clade_ids_out) return ASM_E_INVALID;
  const int64_t nn = 2 * (int64_t)enc->n_taxa - 1;
  return encode_batch(enc, n_trees, clade_ids_out, encoder_threads(n_trees), [&](int64_t t, int32_t* out, auto&& resolve) {
    return encode_tree_with(enc, left + t * nn, right + t * nn, root[t], out, resolve);
  });
}

}  // extern "C"

// Minimal Newick reader for BEAST tree logs: "(...)label:length[&meta]...;" with integer leaf labels.  Fills the child
// arrays (left/right: 2n-1 entries, -1 for leaves) and the root; ASM_OK or the error the caller reports.
static int32_t parse_newick(int n, const char* s, int32_t label_offset, std::vector<int32_t>& left, std::vector<int32_t>& right,
                            int32_t* root_out) {
  const int nn = 2 * n - 1;
  left.assign((size_t)nn, -1);
  right.assign((size_t)nn, -1);
  static thread_local std::vector<int32_t> stack;          // completed subtree roots
  static thread_local std::vector<int32_t> open_marks;     // stack depth at each '('
  stack.clear();
  open_marks.clear();
  int next_internal = n;
  const char* p = s;
  auto skip_meta = [&]() {
    for (;;) {
      while (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r') p++;
      if (*p == '[') {
        while (*p && *p != ']') p++;
        if (*p) p++;
      } else if (*p == ':') {
        p++;
        while (*p == '[') {  // ":[&rate=..]0.1"
          while (*p && *p != ']') p++;
          if (*p) p++;
        }
        while (*p && (std::isdigit((unsigned char)*p) || *p == '.' || *p == 'e' || *p == 'E' || *p == '-' || *p == '+'))
          p++;
      } else {
        break;
      }
    }
  };
  static thread_local std::vector<uint8_t> seen;
  seen.assign((size_t)n, 0);
  while (*p && *p != ';') {
    skip_meta();
    if (*p == '(') {
      open_marks.push_back((int32_t)stack.size());
      p++;
    } else if (*p == ',') {
      p++;
    } else if (*p == ')') {
      p++;
      if (open_marks.empty()) return ASM_E_INVALID;
      int32_t mark = open_marks.back();
      open_marks.pop_back();
      if ((int32_t)stack.size() - mark != 2) return ASM_E_UNSUPPORTED;  // binary trees only (Node.getLeft/getRight)
      if (next_internal >= nn) return ASM_E_INVALID;
      int32_t r = stack.back();
      stack.pop_back();
      int32_t l = stack.back();
      stack.pop_back();
      left[(size_t)next_internal] = l;
      right[(size_t)next_internal] = r;
      stack.push_back(next_internal++);
      // an internal node may carry a label; skip it
      while (*p && *p != ':' && *p != ',' && *p != ')' && *p != ';' && *p != '[') p++;
    } else if (*p == ';' || *p == 0) {
      break;
    } else {
      char* endp = nullptr;
      long v = std::strtol(p, &endp, 10);
      if (endp == p) return ASM_E_INVALID;  // non-numeric label: needs the translate map, caller's job
      p = endp;
      v -= label_offset;
      if (v < 0 || v >= n || seen[(size_t)v]) return ASM_E_INVALID;
      seen[(size_t)v] = 1;
      stack.push_back((int32_t)v);
    }
  }
  if (stack.size() != 1 || next_internal != nn) return ASM_E_INVALID;
  *root_out = stack[0];
  return ASM_OK;
}

extern "C" {

int32_t asm_encoder_add_newick(asm_encoder* enc, const char* s, int32_t label_offset, int32_t* clade_ids_out) {
  if (!enc || !s || !clade_ids_out) return ASM_E_INVALID;
  std::vector<int32_t> left, right;
  int32_t root = -1;
  const int32_t rc = parse_newick(enc->n_taxa, s, label_offset, left, right, &root);
  if (rc != ASM_OK) return rc;
  return asm_encode_tree(enc, left.data(), right.data(), root, clade_ids_out) ? ASM_OK : ASM_E_INVALID;
}

// n_trees Newick strings at once (an offline caller that has read whole .trees files, TreeConvergenceCheck.java:94-112):
// parsing and the clade walks run on every host core, the dictionary ids are those the one-by-one calls would give.
int32_t asm_encoder_add_newick_batch(asm_encoder* enc, int64_t n_trees, const char* const* newicks, int32_t label_offset,
                                     int32_t* clade_ids_out) {
  if (!enc || n_trees < 0 || (n_trees > 0 && (!newicks || !clade_ids_out))) return ASM_E_INVALID;
  std::atomic<int32_t> first_err{ASM_OK};
  const int32_t rc = encode_batch(enc, n_trees, clade_ids_out, encoder_threads(n_trees), [&](int64_t t, int32_t* out, auto&& resolve) {
    static thread_local std::vector<int32_t> left, right;
    int32_t root = -1;
    const int32_t prc = newicks[t] ? parse_newick(enc->n_taxa, newicks[t], label_offset, left, right, &root) : ASM_E_INVALID;
    if (prc != ASM_OK) {
      int32_t expect = ASM_OK;
      first_err.compare_exchange_strong(expect, prc);
      return false;
    }
    return encode_tree_with(enc, left.data(), right.data(), root, out, resolve);
  });
  return rc == ASM_OK ? ASM_OK : (first_err.load() != ASM_OK ? first_err.load() : rc);
}

int64_t asm_encoder_num_clades(const asm_encoder* enc) { return enc ? enc->num_clades() : 0; }

int32_t asm_encoder_clade_leaves(const asm_encoder* enc, int64_t id, int32_t* leaves_out, int32_t cap) {
  if (!enc || id < 0 || id >= enc->num_clades()) return ASM_E_INVALID;
  const int32_t n = enc->clade_size(id);
  if (n > cap || !leaves_out) return ASM_E_RANGE;
  const uint64_t* p = enc->clade_set(id);
  int32_t k = 0;
  for (int w = 0; w < enc->set_words; w++)  // ascending leaf numbers
    for (uint64_t m = p[w]; m; m &= m - 1) leaves_out[k++] = w * 64 + __builtin_ctzll(m);
  return n;
}

int32_t asm_encoder_words(const asm_encoder* enc) {
  if (!enc) return 0;
  int64_t w = (enc->num_clades() + 63) / 64;
  w = (w + 3) / 4 * 4;
  return (int32_t)(w < 4 ? 4 : w);
}

int32_t asm_encode_bits(const int32_t* clade_ids, int64_t n_trees, int32_t ids_per_tree, int32_t words_per_tree,
                        uint64_t* bits_out) {
  if (!clade_ids || !bits_out || n_trees < 0 || ids_per_tree < 0 || words_per_tree <= 0) return ASM_E_INVALID;
  std::atomic<int> bad{0};
  auto rows = [&](int64_t t0, int64_t t1) {
    for (int64_t t = t0; t < t1; t++) {
      uint64_t* row = bits_out + t * words_per_tree;
      std::memset(row, 0, sizeof(uint64_t) * (size_t)words_per_tree);
      for (int32_t j = 0; j < ids_per_tree; j++) {
        int32_t id = clade_ids[t * ids_per_tree + j];
        if (id < 0 || (id >> 6) >= words_per_tree) {
          bad.store(1);
          return;
        }
        row[id >> 6] |= 1ULL << (id & 63);
      }
    }
  };
  const int threads = encoder_threads(n_trees);
  if (threads <= 1) {
    rows(0, n_trees);
  } else {
    std::vector<std::thread> pool;
    const int64_t per = (n_trees + threads - 1) / threads;
    for (int k = 0; k < threads; k++) pool.emplace_back(rows, std::min<int64_t>(k * per, n_trees), std::min<int64_t>((k + 1) * per, n_trees));
    for (auto& th : pool) th.join();
  }
  return bad.load() ? ASM_E_RANGE : ASM_OK;
}

}  // extern "C"
