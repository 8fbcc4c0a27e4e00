// synth.cpp -- synthetic inputs of the named benchmark shapes (BASELINE.json configs).
// Test and benchmark infrastructure, NOT part of the product: built into synthgen/libasmsynth.so together with its
// own copy of the host encoder (asm_b200/csrc/host_encoder.cpp), so that the reference arm of bench.py maps nothing
// of libasmgpu.so.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_map>
#include <vector>

#include "../asm_b200/csrc/host_encoder.h"
#include "asmsynth.h"

extern "C" {

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
    if (!asm_encode_tree(enc, L.data(), R.data(), root, ids.data())) return ASM_E_INVALID;
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
