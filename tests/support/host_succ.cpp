// Test harness: "is t a successor of s under the lowered Next?" with the host build of a model header.
#include <stdint.h>
#include <string.h>
#include KMC_MODEL_HEADER
using kmc_model::State;
namespace {
struct Sink {
  const State* target;
  int found = -1;
  void emit(const State& n, int a) {
    if (memcmp(n.w, target->w, sizeof(n.w)) == 0 && found < 0) found = a;
  }
  void fail(int) {}
};
}  // namespace
extern "C" int kmc_host_is_successor(const uint64_t* s, const uint64_t* t) {
  State a, b;
  memcpy(a.w, s, sizeof(a.w));
  memcpy(b.w, t, sizeof(b.w));
  Sink sink;
  sink.target = &b;
  kmc_model::expand(a, sink);
  return sink.found;
}
extern "C" int kmc_host_is_init(const uint64_t* s) {
  for (int i = 0; i < kmc_model::NUM_INIT; ++i)
    if (memcmp(kmc_model::INIT_STATES[i], s, sizeof(uint64_t) * kmc_model::W) == 0) return 1;
  return 0;
}
extern "C" int kmc_host_first_violated(const uint64_t* s) {
  State a;
  memcpy(a.w, s, sizeof(a.w));
  return kmc_model::first_violated_invariant(a);
}
