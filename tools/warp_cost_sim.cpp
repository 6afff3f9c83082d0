// Planning tool (CPU, not a product path): how much of K1's lane under-utilisation could an ordering of the
// frontier remove?  Runs a sequential BFS over a lowered model, and for every level computes, for several
// orderings of that level's states, the number of (warp, item) pairs in which at least one of the warp's 32
// states enables the item (item = one top-level guarded block of the lowered Next; a warp executes a body as
// soon as one lane enables it).  ideal = enabled (state, item) pairs / 32.
//   g++ -O2 -std=c++17 -DKMC_MODEL_HEADER='"build/models/<m>/model.h"' tools/warp_cost_sim.cpp -o /tmp/wcs
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <numeric>
#include <unordered_set>
#include <vector>
#include KMC_MODEL_HEADER

namespace M = kmc_model;
using M::State;
struct H {
  size_t operator()(const State& s) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < M::W; ++i) { h ^= s.w[i] + (h << 6) + (h >> 2); h *= 0xff51afd7ed558ccdull; h ^= h >> 33; }
    return (size_t)h;
  }
};
struct Eq { bool operator()(const State& a, const State& b) const { return !memcmp(a.w, b.w, sizeof a.w); } };
struct Sink {
  std::vector<State>* out; std::vector<uint8_t>* act; int failed = 0;
  void emit(const State& n, int a) { out->push_back(n); act->push_back((uint8_t)a); }
  void fail(int c) { failed = c; }
};
template <int G> struct MaskLoop {
  static uint64_t run(const State& s) {
    return (M::group_guard_mask(M::GroupTag<G>{}, s) << M::GROUP_ITEM_BEGIN[G]) | MaskLoop<G + 1>::run(s);
  }
};
template <> struct MaskLoop<M::NUM_GROUPS> { static uint64_t run(const State&) { return 0; } };

static uint64_t warp_pairs(const std::vector<uint64_t>& mask, const std::vector<uint32_t>& order) {
  uint64_t pairs = 0;
  for (size_t i = 0; i < order.size(); i += 32) {
    uint64_t any = 0;
    for (size_t j = i; j < std::min(order.size(), i + 32); ++j) any |= mask[order[j]];
    pairs += __builtin_popcountll(any);
  }
  return pairs;
}

int main(int argc, char** argv) {
  uint64_t max_states = argc > 1 ? strtoull(argv[1], 0, 10) : 40000000ull;
  std::unordered_set<State, H, Eq> seen;
  std::vector<State> frontier, next, succ;
  std::vector<uint8_t> fact, nact, sact;
  for (int i = 0; i < M::NUM_INIT; ++i) {
    State s; memcpy(s.w, M::INIT_STATES[i], sizeof s.w);
    if (seen.insert(s).second) { frontier.push_back(s); fact.push_back(0); }
  }
  uint64_t tot_en = 0, tot_gen = 0, tot_mask = 0, tot_act = 0, tot_states = 0;
  int level = 1;
  while (!frontier.empty() && seen.size() < max_states) {
    size_t n = frontier.size();
    std::vector<uint64_t> mask(n);
    uint64_t enabled = 0;
    for (size_t i = 0; i < n; ++i) { mask[i] = MaskLoop<0>::run(frontier[i]); enabled += __builtin_popcountll(mask[i]); }
    std::vector<uint32_t> ord(n);
    std::iota(ord.begin(), ord.end(), 0u);
    uint64_t p_gen = warp_pairs(mask, ord);
    std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return fact[a] < fact[b]; });
    uint64_t p_act = warp_pairs(mask, ord);
    std::iota(ord.begin(), ord.end(), 0u);
    std::sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return mask[a] < mask[b]; });
    uint64_t p_mask = warp_pairs(mask, ord);
    if (n >= 100000)
      printf("level %2d  states %9zu  enabled/state %.2f  lanes per executed body: generation order %.1f, by producing action %.1f, "
             "by guard mask %.1f\n", level, n, (double)enabled / n, (double)enabled / p_gen, (double)enabled / p_act,
             (double)enabled / p_mask);
    tot_en += enabled; tot_gen += p_gen; tot_act += p_act; tot_mask += p_mask; tot_states += n;
    next.clear(); nact.clear();
    for (size_t i = 0; i < n; ++i) {
      succ.clear(); sact.clear();
      Sink sink{&succ, &sact};
      M::expand(frontier[i], sink);
      for (size_t k = 0; k < succ.size(); ++k)
        if (M::in_model(succ[k]) && seen.insert(succ[k]).second) { next.push_back(succ[k]); nact.push_back(sact[k]); }
    }
    frontier.swap(next); fact.swap(nact);
    ++level;
  }
  printf("TOTAL states %llu  items %d  enabled pairs %llu  warp-level body executions: generation order %llu (%.1f lanes), "
         "by producing action %llu (%.1f lanes), by guard mask %llu (%.1f lanes)\n",
         (unsigned long long)tot_states, M::NUM_ITEMS, (unsigned long long)tot_en, (unsigned long long)tot_gen,
         (double)tot_en / tot_gen, (unsigned long long)tot_act, (double)tot_en / tot_act, (unsigned long long)tot_mask,
         (double)tot_en / tot_mask);
  return 0;
}
