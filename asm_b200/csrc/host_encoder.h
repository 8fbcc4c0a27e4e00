// host_encoder.h -- internals of the host encoder shared by host_encoder.cpp (libasmgpu.so) and the synthetic
// input generators (synthgen/synth.cpp, which builds its own copy into libasmsynth.so).  Not part of the ABI.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/asmgpu.h"

namespace asm_enc_detail {

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

}  // namespace asm_enc_detail
using asm_enc_detail::Key128;
using asm_enc_detail::KeyHash;
using asm_enc_detail::splitmix64;
using asm_enc_detail::mix_at;

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
  // read-only lookup (no insertion, safe from several threads while nobody inserts): id or -1
  int32_t find_set(const Key128& k, const uint64_t* bits) const {
    if (slots.empty()) return -1;
    size_t h = KeyHash()(k) & mask;
    for (;; h = (h + 1) & mask) {
      const Slot& s = slots[h];
      if (s.id < 0) return -1;
      if (s.key == k && std::memcmp(clade_set(s.id), bits, sizeof(uint64_t) * (size_t)set_words) == 0) return s.id;
    }
  }
  // the same from a list of leaf numbers (any order)
  int32_t get_or_add(const Key128& k, const int32_t* leaves, int32_t n) {
    scratch.assign((size_t)set_words, 0);
    for (int32_t i = 0; i < n; i++) scratch[(size_t)(leaves[i] >> 6)] |= 1ULL << (leaves[i] & 63);
    return get_or_add_set(k, scratch.data());
  }
};


// Post-order walk of one tree (host_encoder.cpp): one clade id per internal node; false if the arrays do not
// describe a binary tree on n_taxa leaves.
bool asm_encode_tree(asm_encoder* enc, const int32_t* left, const int32_t* right, int32_t root, int32_t* ids_out);

// A clade met during the parallel phase that the dictionary (frozen for that phase) does not hold yet.
struct asm_encoder_miss {
  int32_t pos;      // position of the clade in the tree's post-order
  Key128 key;
  size_t set_off;   // leaf set: set_words words at this offset of the recording thread's pool
};
