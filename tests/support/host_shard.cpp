// Test stand-in for ONE RANK of the sharded engine, on the host (NOT a product path): it gives the
// multi-rank driver (kafka_specification_b200/sharded.py) something to drive in the CPU-only gloo
// tests.  Same row format and owner function as kmc_engine.cu, std::unordered_set instead of the
// HBM hash set.
#include <stdint.h>
#include <string.h>
#include <unordered_set>
#include <vector>
#include KMC_MODEL_HEADER

using kmc_model::State;
static constexpr int W = kmc_model::W;
static constexpr int ROW = W + 1;
static constexpr uint64_t NO_PARENT = 0x0000FFFFFFFFFFFFull;

static inline uint64_t fmix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
static inline uint64_t fingerprint(const State& s) {
  if (kmc_model::STATE_BITS <= 63) return fmix64(s.w[0] + 1);
  uint64_t h = fmix64(s.w[0] + 0x9E3779B97F4A7C15ull);
  for (int i = 1; i < W; ++i) h = fmix64(h ^ (s.w[i] + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1)));
  return h ? h : 1;
}
static inline uint32_t owner_of(uint64_t fp, uint32_t world) { return (uint32_t)(((fp >> 32) * (uint64_t)world) >> 32); }

namespace {
struct StateHash { size_t operator()(const State& s) const { return (size_t)fingerprint(s); } };
struct StateEq { bool operator()(const State& a, const State& b) const { return memcmp(a.w, b.w, sizeof(a.w)) == 0; } };
struct Rank {
  uint32_t rank, world;
  std::unordered_set<State, StateHash, StateEq> seen;
  std::vector<State> store;
  std::vector<uint64_t> parent;
  State viol_state;
  uint64_t viol_meta = 0, viol_fp = ~0ull, viol_level = 0, level = 0;
  std::vector<std::vector<uint64_t>> cand;
  std::vector<uint64_t> recv;
  uint64_t generated = 0, deadlocks = 0, level_first = 0, level_count = 0;
  int64_t viol_inv = -1;
  int fail = 0;
};
struct Sink {
  Rank* r; uint64_t parent; int n = 0;
  void emit(const State& s, int action) {
    ++n;
    uint32_t d = r->world > 1 ? owner_of(fingerprint(s), r->world) : 0;
    for (int i = 0; i < W; ++i) r->cand[d].push_back(s.w[i]);
    r->cand[d].push_back(parent | ((uint64_t)action << 56));
  }
  void fail(int code) { r->fail = code; }
};
}  // namespace

extern "C" {
void* hs_create(uint32_t rank, uint32_t world) {
  Rank* r = new Rank();
  r->rank = rank; r->world = world; r->cand.resize(world);
  return r;
}
void hs_destroy(void* h) { delete (Rank*)h; }
void hs_begin(void* h) {
  Rank* r = (Rank*)h;
  r->seen.clear(); r->store.clear(); r->parent.clear();
  r->viol_fp = ~0ull; r->viol_level = 0; r->level = 0;
  for (auto& c : r->cand) c.clear();
  r->generated = r->deadlocks = r->level_first = r->level_count = 0;
  r->viol_inv = -1; r->fail = 0;
}
void hs_seed_init(void* h) {
  Rank* r = (Rank*)h;
  if (r->rank != 0) return;
  for (int i = 0; i < kmc_model::NUM_INIT; ++i) {
    State s; memcpy(s.w, kmc_model::INIT_STATES[i], sizeof(s.w));
    uint32_t d = r->world > 1 ? owner_of(fingerprint(s), r->world) : 0;
    for (int k = 0; k < W; ++k) r->cand[d].push_back(s.w[k]);
    r->cand[d].push_back(NO_PARENT);
    r->generated++;
  }
}
void hs_reset_cand(void* h) { for (auto& c : ((Rank*)h)->cand) c.clear(); }
void hs_expand(void* h, uint64_t first, uint64_t count) {
  Rank* r = (Rank*)h;
  for (uint64_t i = first; i < first + count; ++i) {
    Sink sink{r, i | ((uint64_t)r->rank << 40)};
    kmc_model::expand(r->store[i], sink);
    r->generated += sink.n;
    if (sink.n == 0) r->deadlocks++;
  }
}
void hs_counts(void* h, uint64_t* out) {
  Rank* r = (Rank*)h;
  for (uint32_t d = 0; d < r->world; ++d) out[d] = r->cand[d].size() / ROW;
}
uint64_t* hs_send_ptr(void* h, uint32_t dest) { return ((Rank*)h)->cand[dest].data(); }
uint64_t* hs_recv_ptr(void* h, uint64_t rows) {
  Rank* r = (Rank*)h;
  if (r->recv.size() < rows * ROW) r->recv.resize(rows * ROW);
  return r->recv.data();
}
void hs_insert(void* h, const uint64_t* rows, uint64_t n) {
  Rank* r = (Rank*)h;
  for (uint64_t i = 0; i < n; ++i) {
    State s; memcpy(s.w, rows + i * ROW, sizeof(s.w));
    bool inmodel = kmc_model::in_model(s);
    bool is_new = inmodel && r->seen.insert(s).second;
    uint64_t meta = rows[i * ROW + W];
    if (is_new) { r->store.push_back(s); r->parent.push_back(meta); }
    if (is_new || !inmodel) {
      int inv = kmc_model::first_violated_invariant(s);
      uint64_t fp = fingerprint(s);
      if (inv >= 0 && (r->viol_inv < 0 || r->viol_level == r->level + 1) && fp < r->viol_fp) {
        if (r->viol_inv < 0) r->viol_level = r->level + 1;
        r->viol_inv = inv; r->viol_fp = fp; r->viol_state = s; r->viol_meta = meta;
      }
    }
  }
}
void hs_level_done(void* h, uint64_t* first, uint64_t* count) {
  Rank* r = (Rank*)h;
  r->level_first += r->level_count;
  r->level_count = r->store.size() - r->level_first;
  r->level++;
  *first = r->level_first; *count = r->level_count;
}
void hs_stats(void* h, uint64_t* out) {
  Rank* r = (Rank*)h;
  out[0] = r->store.size(); out[1] = r->generated; out[2] = r->deadlocks; out[3] = (uint64_t)r->fail;
  out[4] = (uint64_t)r->viol_inv;
}
int hs_row_words() { return ROW; }
// violation record: out[0..W) words, out[W] parent word, out[W+1] fingerprint, out[W+2] level; returns invariant or -1
int hs_violation_record(void* h, uint64_t* out) {
  Rank* r = (Rank*)h;
  if (r->viol_inv < 0) return -1;
  memcpy(out, r->viol_state.w, sizeof(r->viol_state.w));
  out[W] = r->viol_meta; out[W + 1] = r->viol_fp; out[W + 2] = r->viol_level;
  return (int)r->viol_inv;
}
void hs_state_and_parent(void* h, uint64_t idx, uint64_t* out) {
  Rank* r = (Rank*)h;
  memcpy(out, r->store[idx].w, sizeof(uint64_t) * W);
  out[W] = r->parent[idx];
}
int hs_is_successor(const uint64_t* a, const uint64_t* b) {
  struct Find { const State* t; int found = -1; void emit(const State& n, int act) { if (found < 0 && memcmp(n.w, t->w, sizeof(n.w)) == 0) found = act; } void fail(int) {} };
  State s, t; memcpy(s.w, a, sizeof(s.w)); memcpy(t.w, b, sizeof(t.w));
  Find f; f.t = &t;
  kmc_model::expand(s, f);
  return f.found;
}
int hs_is_init(const uint64_t* a) {
  for (int i = 0; i < kmc_model::NUM_INIT; ++i) if (memcmp(kmc_model::INIT_STATES[i], a, sizeof(uint64_t) * W) == 0) return 1;
  return 0;
}
}
