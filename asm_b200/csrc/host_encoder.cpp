// host_encoder.cpp -- trees -> clades -> dictionary -> bit rows.  CPU only; part of libasmgpu.so but
// needs no GPU.
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

#include "host_encoder.h"

// Post-order walk (left subtree, right subtree, node) without recursion limits: emits one clade id
// per internal node.  Returns false if the arrays do not describe a binary tree on n_taxa leaves.
bool asm_encode_tree(asm_encoder* enc, const int32_t* left, const int32_t* right, int32_t root, int32_t* ids_out) {
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
    if (!asm_encode_tree(enc, left + t * nn, right + t * nn, root[t], clade_ids_out + t * k)) return ASM_E_INVALID;
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
  return asm_encode_tree(enc, left.data(), right.data(), root, clade_ids_out) ? ASM_OK : ASM_E_INVALID;
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

}  // extern "C"
