// host_encoder.cpp -- trees -> clades -> dictionary -> bit rows, and the synthetic inputs of the
// named benchmark shapes.  CPU only; part of libasmgpu.so but needs no GPU.
//
// Clade identity follows the reference: the set of leaf numbers under an internal node, root
// included (CladeDifference.java:107-141).  Instead of the reference's "3,7,12," key strings a
// clade is keyed by a 128-bit Zobrist hash (XOR of per-leaf random words) and every dictionary
// hit is confirmed against the stored leaf list, so the mapping is exact, not probabilistic.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/asmgpu.h"

namespace {

inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// counter-based: value i of stream `key`
inline uint64_t mix_at(uint64_t key, uint64_t i) {
  uint64_t s = key + i * 0xD1342543DE82EF95ULL;
  return splitmix64(s);
}

struct Key128 {
  uint64_t a, b;
  bool operator==(const Key128& o) const { return a == o.a && b == o.b; }
};
struct KeyHash {
  size_t operator()(const Key128& k) const { return (size_t)(k.a ^ (k.b * 0x9E3779B97F4A7C15ULL)); }
};

}  // namespace

struct asm_encoder {
  int32_t n_taxa = 0;
  int32_t set_words = 0;        // words of a leaf-set bitset: ceil(n_taxa / 64)
  std::vector<Key128> zobrist;  // per leaf
  // open-addressing table, linear probing: (key, id) with id < 0 = empty.  A node-based std::unordered_map costs
  // a cache miss or two per lookup, and there is one lookup per internal node of every tree.
  struct Slot {
    Key128 key;
    int32_t id;
  };
  std::vector<Slot> slots;
  size_t mask = 0;
  // id -> leaf set as a bitset over the taxa (bit v = leaf number v).  Bitsets make a node's set the OR of its
  // children's (n/64 words) where sorted leaf lists cost a merge of the whole subtree at every node.
  std::vector<uint64_t> set_pool;
  std::vector<uint64_t> scratch;
  std::string err;

  int64_t num_clades() const { return set_words ? (int64_t)(set_pool.size() / (size_t)set_words) : 0; }
  const uint64_t* clade_set(int64_t id) const { return set_pool.data() + (size_t)id * (size_t)set_words; }
  int32_t clade_size(int64_t id) const {
    int32_t c = 0;
    const uint64_t* p = clade_set(id);
    for (int w = 0; w < set_words; w++) c += __builtin_popcountll(p[w]);
    return c;
  }
  void grow() {
    const size_t cap = slots.empty() ? 4096 : slots.size() * 2;
    std::vector<Slot> old;
    old.swap(slots);
    slots.assign(cap, Slot{{0, 0}, -1});
    mask = cap - 1;
    for (const Slot& s : old) {
      if (s.id < 0) continue;
      size_t h = KeyHash()(s.key) & mask;
      while (slots[h].id >= 0) h = (h + 1) & mask;
      slots[h] = s;
    }
  }
  int32_t get_or_add_set(const Key128& k, const uint64_t* bits) {
    if ((size_t)(num_clades() + 1) * 2 > slots.size()) grow();  // load factor <= 1/2
    size_t h = KeyHash()(k) & mask;
    for (;; h = (h + 1) & mask) {
      const Slot& s = slots[h];
      if (s.id < 0) break;
      // equal keys are confirmed against the stored leaf set: the mapping is exact, not probabilistic
      if (s.key == k && std::memcmp(clade_set(s.id), bits, sizeof(uint64_t) * (size_t)set_words) == 0) return s.id;
    }
    const int32_t id = (int32_t)num_clades();
    set_pool.insert(set_pool.end(), bits, bits + set_words);
    slots[h] = Slot{k, id};
    return id;
  }
  // the same from a list of leaf numbers (any order)
  int32_t get_or_add(const Key128& k, const int32_t* leaves, int32_t n) {
    scratch.assign((size_t)set_words, 0);
    for (int32_t i = 0; i < n; i++) scratch[(size_t)(leaves[i] >> 6)] |= 1ULL << (leaves[i] & 63);
    return get_or_add_set(k, scratch.data());
  }
};

namespace {

// Post-order walk (left subtree, right subtree, node) without recursion limits: emits one clade id
// per internal node.  Returns false if the arrays do not describe a binary tree on n_taxa leaves.
bool encode_tree(asm_encoder* enc, const int32_t* left, const int32_t* right, int32_t root, int32_t* ids_out) {
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
    ids_out[n_out++] = enc->get_or_add_set(k, sets.data() + (sets.size() - (size_t)W));
    sizes.push_back(nl + nr);
    keys.push_back(k);
    st.pop_back();
  }
  return n_out == n - 1 && sizes.size() == 1 && sizes[0] == n;
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
  const int64_t nn = 2 * (int64_t)enc->n_taxa - 1, k = enc->n_taxa - 1;
  for (int64_t t = 0; t < n_trees; t++)
    if (!encode_tree(enc, left + t * nn, right + t * nn, root[t], clade_ids_out + t * k)) return ASM_E_INVALID;
  return ASM_OK;
}

// Minimal Newick reader for BEAST tree logs: "(...)label:length[&meta]...;" with integer leaf labels.
int32_t asm_encoder_add_newick(asm_encoder* enc, const char* s, int32_t label_offset, int32_t* clade_ids_out) {
  if (!enc || !s || !clade_ids_out) return ASM_E_INVALID;
  const int n = enc->n_taxa, nn = 2 * n - 1;
  std::vector<int32_t> left((size_t)nn, -1), right((size_t)nn, -1);
  std::vector<int32_t> stack;          // completed subtree roots
  std::vector<int32_t> open_marks;     // stack depth at each '('
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
  std::vector<uint8_t> seen((size_t)n, 0);
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
  int32_t root = stack[0];
  return encode_tree(enc, left.data(), right.data(), root, clade_ids_out) ? ASM_OK : ASM_E_INVALID;
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
  std::memset(bits_out, 0, sizeof(uint64_t) * (size_t)n_trees * (size_t)words_per_tree);
  for (int64_t t = 0; t < n_trees; t++) {
    uint64_t* row = bits_out + t * words_per_tree;
    for (int32_t j = 0; j < ids_per_tree; j++) {
      int32_t id = clade_ids[t * ids_per_tree + j];
      if (id < 0 || (id >> 6) >= words_per_tree) return ASM_E_RANGE;
      row[id >> 6] |= 1ULL << (id & 63);
    }
  }
  return ASM_OK;
}

// ------------------------------------------------------------------------------------------------
// Synthetic chains
// ------------------------------------------------------------------------------------------------

int32_t asm_synth_tree_chain(int32_t n_taxa, int64_t n_trees, uint64_t start_seed, uint64_t chain_seed,
                             int32_t moves_per_sample, double beta, int32_t* left_out, int32_t* right_out,
                             int32_t* root_out, asm_encoder* enc, int32_t* clade_ids_out) {
  if (n_taxa < 3 || n_trees < 0 || moves_per_sample < 0) return ASM_E_INVALID;
  if (enc && (enc->n_taxa != n_taxa || !clade_ids_out)) return ASM_E_INVALID;
  const int n = n_taxa, nn = 2 * n - 1;
  std::vector<int32_t> L((size_t)nn, -1), R((size_t)nn, -1), P((size_t)nn, -1);
  // shared start tree: join two random lineages until one is left
  {
    uint64_t s = start_seed * 0x9E3779B97F4A7C15ULL + 12345;
    std::vector<int32_t> live((size_t)n);
    for (int i = 0; i < n; i++) live[(size_t)i] = i;
    int next = n;
    while (live.size() > 1) {
      size_t i = (size_t)(splitmix64(s) % live.size());
      int32_t a = live[i];
      live[i] = live.back();
      live.pop_back();
      size_t j = (size_t)(splitmix64(s) % live.size());
      int32_t b = live[j];
      L[(size_t)next] = a;
      R[(size_t)next] = b;
      P[(size_t)a] = next;
      P[(size_t)b] = next;
      live[j] = next++;
    }
  }
  int32_t root = nn - 1;
  // per-node Zobrist hash of the leaf set (own table: the walk must not depend on an encoder)
  std::vector<Key128> Z((size_t)n), H((size_t)nn);
  {
    uint64_t s = 0xC0FFEE1234ULL;
    for (auto& z : Z) {
      z.a = splitmix64(s);
      z.b = splitmix64(s);
    }
  }
  for (int v = 0; v < n; v++) H[(size_t)v] = Z[(size_t)v];
  for (int v = n; v < nn; v++) {  // children are always created before parents in the join order
    H[(size_t)v].a = H[(size_t)L[(size_t)v]].a ^ H[(size_t)R[(size_t)v]].a;
    H[(size_t)v].b = H[(size_t)L[(size_t)v]].b ^ H[(size_t)R[(size_t)v]].b;
  }
  std::unordered_map<Key128, char, KeyHash> home;  // clades of the start tree
  for (int v = n; v < nn; v++) home[H[(size_t)v]] = 1;
  std::vector<char> in_home((size_t)nn, 1);

  // encoder ids per internal node (fast path: only the node an NNI changes is re-looked-up)
  std::vector<int32_t> node_id((size_t)nn, -1);
  std::vector<int32_t> tmp_leaves;
  auto collect_leaves = [&](int32_t v) {
    tmp_leaves.clear();
    std::vector<int32_t> stk{v};
    while (!stk.empty()) {
      int32_t u = stk.back();
      stk.pop_back();
      if (L[(size_t)u] < 0)
        tmp_leaves.push_back(u);
      else {
        stk.push_back(L[(size_t)u]);
        stk.push_back(R[(size_t)u]);
      }
    }
    std::sort(tmp_leaves.begin(), tmp_leaves.end());
  };
  auto enc_key = [&](int32_t v) {  // hash under the ENCODER's Zobrist table
    Key128 k{0, 0};
    for (int32_t leaf : tmp_leaves) {
      k.a ^= enc->zobrist[(size_t)leaf].a;
      k.b ^= enc->zobrist[(size_t)leaf].b;
    }
    (void)v;
    return k;
  };
  if (enc) {
    // assign ids of the start tree in post-order so that they agree with asm_encoder_add_topologies
    std::vector<int32_t> ids((size_t)(n - 1));
    if (!encode_tree(enc, L.data(), R.data(), root, ids.data())) return ASM_E_INVALID;
    // map back: redo the post-order to know which node got which id
    std::vector<std::pair<int32_t, int>> stk{{root, 0}};
    int k = 0;
    while (!stk.empty()) {
      auto& f = stk.back();
      int32_t v = f.first;
      if (L[(size_t)v] < 0) {
        stk.pop_back();
        continue;
      }
      if (f.second == 0) {
        f.second = 1;
        stk.push_back({L[(size_t)v], 0});
      } else if (f.second == 1) {
        f.second = 2;
        stk.push_back({R[(size_t)v], 0});
      } else {
        node_id[(size_t)v] = ids[(size_t)k++];
        stk.pop_back();
      }
    }
  }

  uint64_t rs = chain_seed * 0xD6E8FEB86659FD93ULL + 777;
  auto uniform01 = [&]() { return (double)(splitmix64(rs) >> 11) * (1.0 / 9007199254740992.0); };
  const double p_out = std::exp(-beta);  // accept prob. for a move that leaves the home clade set

  for (int64_t t = 0; t < n_trees; t++) {
    for (int m = 0; m < moves_per_sample; m++) {
      // random internal non-root node v with parent u and sibling s; swap s with a random child c of v
      int32_t v = n + (int32_t)(splitmix64(rs) % (uint64_t)(n - 1));
      if (v == root) continue;
      int32_t u = P[(size_t)v];
      int32_t s = (L[(size_t)u] == v) ? R[(size_t)u] : L[(size_t)u];
      bool pick_left = (splitmix64(rs) & 1) != 0;
      int32_t c = pick_left ? L[(size_t)v] : R[(size_t)v];
      int32_t o = pick_left ? R[(size_t)v] : L[(size_t)v];  // child that stays
      Key128 nk{H[(size_t)o].a ^ H[(size_t)s].a, H[(size_t)o].b ^ H[(size_t)s].b};
      bool new_home = home.find(nk) != home.end();
      int dE = (new_home ? 0 : 1) - (in_home[(size_t)v] ? 0 : 1);
      if (dE > 0 && uniform01() >= p_out) continue;  // Metropolis reject
      // apply: v keeps o, gains s; u keeps v, gains c
      if (pick_left) L[(size_t)v] = s; else R[(size_t)v] = s;
      if (L[(size_t)u] == s) L[(size_t)u] = c; else R[(size_t)u] = c;
      P[(size_t)s] = v;
      P[(size_t)c] = u;
      H[(size_t)v] = nk;
      in_home[(size_t)v] = new_home ? 1 : 0;
      if (enc) {
        collect_leaves(v);
        node_id[(size_t)v] = enc->get_or_add(enc_key(v), tmp_leaves.data(), (int32_t)tmp_leaves.size());
      }
    }
    if (left_out) std::memcpy(left_out + t * nn, L.data(), sizeof(int32_t) * (size_t)nn);
    if (right_out) std::memcpy(right_out + t * nn, R.data(), sizeof(int32_t) * (size_t)nn);
    if (root_out) root_out[t] = root;
    if (enc) {
      int32_t* dst = clade_ids_out + t * (n - 1);
      for (int v = n; v < nn; v++) dst[v - n] = node_id[(size_t)v];
    }
  }
  return ASM_OK;
}

int32_t asm_synth_ar1(double* out, int64_t n, double mu, double phi, uint64_t seed) {
  if (!out || n < 0 || !(std::fabs(phi) < 1.0)) return ASM_E_INVALID;
  const uint64_t key = seed * 0xA0761D6478BD642FULL + 0xE7037ED1A0B428DBULL;
  auto normal_at = [&](int64_t i) {
    // Box-Muller on two counter-based uniforms
    double u1 = ((double)(mix_at(key, 2 * (uint64_t)i) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    double u2 = ((double)(mix_at(key, 2 * (uint64_t)i + 1) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
  };
  double x = 0.0;
  for (int64_t i = 0; i < n; i++) {
    double e = normal_at(i);
    x = (i == 0) ? e / std::sqrt(1.0 - phi * phi) : phi * x + e;
    out[i] = mu + x;
  }
  return ASM_OK;
}

}  // extern "C"
