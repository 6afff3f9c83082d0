// kmc_engine.cu -- the BFS frontier-expansion engine, compiled once per lowered model:
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared \
//        -Xcompiler -fPIC -include <model>.h kmc_engine.cu -o libkmc_<model>.so
//
// It replaces TLC's Worker next-state loop, FPSet and StateQueue:
//
//   k_expand  (K1)  lowered Next over the frontier.  The unrolled Next is hundreds of KB of SASS, so it
//                   is cut into NUM_GROUPS groups; one 1024-thread CTA per SM sweeps one group at a
//                   time over a tile of states (all warps in the same group => the group's code stays in
//                   the instruction cache).  Successor rows (state words + parent/action word) are
//                   staged per warp in shared memory and flushed in bulk: into the local candidate
//                   buffer (one rank), into per-owner regions (NCCL exchange), or -- fused exchange --
//                   straight into the owner rank's inbox through a CUDA-IPC peer mapping (NVLink stores).
//   k_insert  (K2)  one thread per candidate: 64-bit fingerprint (of the orbit representative under
//                   SYMMETRY), open-addressing hash set in HBM with 32-byte buckets (4 fingerprints =
//                   one DRAM sector, two 128-bit loads), CAS insertion, warp ballot/popc compaction of
//                   the winners into the state store (= next frontier), parent link.  k_insert_inbox is
//                   the same over the regions the peers filled.
//   k_invariants (K3) the cfg's INVARIANTs on the new states of a level (compacted => every lane busy).
//                   Violating states (rare, terminal) go to a small ring; the host reports the one with
//                   the smallest fingerprint, so the counterexample is deterministic.
//   k_expand2       opt-in two-phase form of K1 (guard masks, CTA-wide compaction, bodies); measured
//                   slower, see DESIGN.md section 4.
//
// HBM layout (per rank):   table  u64[2^table_log2]            fingerprints, 0 = empty
//                          store  u64[max_states][W]           all distinct states, BFS order;
//                                                              level k is a contiguous range
//                          parent u64[max_states]              parent ref | action << 56
//                          cand   u64[world][region_rows][W+1] successors of one frontier chunk
//
// There is no CPU fallback anywhere in this file: without a CUDA device kmc_create fails with
// KMC_E_NO_GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "kspecmc.h"

namespace M = kmc_model;
using M::State;

static constexpr int W = M::W;
static constexpr int ROW = W + 1;
static constexpr int MAX_WORLD = 8;
static constexpr uint64_t NO_PARENT = 0x0000FFFFFFFFFFFFull;
static constexpr uint64_t IDX_MASK = 0x000000FFFFFFFFFFull;
static constexpr bool EXACT64 = (M::STATE_BITS <= 63);
static_assert(M::NUM_ACTIONS <= 255, "the parent word holds the action id in 8 bits");

#define KMC_FAIL_TABLE_FULL 2
#define KMC_FAIL_STORE_FULL 3
#define KMC_FAIL_CAND_FULL 4
#define KMC_FAIL_PEER_TIMEOUT 5

// ----------------------------------------------------------------------------------------
// fingerprints
// ----------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}

// Identity of a state in the set.
//   one-word models (<= 63 bits): the fingerprint is a bijection of the state, stored in 8-byte slots   -> exact
//   two-word models             : the key IS the packed state (16-byte slots, 128-bit CAS)              -> exact
//                                 (all-ones marks an empty slot; the lowering proves no valid state packs to it)
//   wider models                : a 128-bit fingerprint (two independent 64-bit chains) in 16-byte slots;
//                                 collision probability ~ n^2 / 2^129 (TLC's FP64 contract squared)
// The 64-bit fingerprint also picks the bucket and the owner rank and orders counterexamples.
static constexpr bool KEY128 = (W >= 2);
static constexpr bool EXACT_SET = EXACT64 || (W == 2 && !M::ALL_ONES_POSSIBLE);
#ifndef KMC_BUCKET_SLOTS
#define KMC_BUCKET_SLOTS 2            // 16-byte slots per bucket: 2 = one 32 B sector, 4 = one 64 B DRAM burst
#endif
static constexpr int BUCKET_SLOTS = KEY128 ? KMC_BUCKET_SLOTS : 4;
static constexpr int SLOT_BYTES = KEY128 ? 16 : 8;

struct alignas(16) Key128 {
  unsigned long long lo, hi;
};
__host__ __device__ __forceinline__ bool key_eq(const Key128& a, const Key128& b) { return a.lo == b.lo && a.hi == b.hi; }
__host__ __device__ __forceinline__ bool key_empty(const Key128& a) { return (a.lo & a.hi) == ~0ull; }

__host__ __device__ __forceinline__ uint64_t fingerprint(const State& s) {
  if (EXACT64) return fmix64(s.w[0] + 1);
  uint64_t h = fmix64(s.w[0] + 0x9E3779B97F4A7C15ull);
#pragma unroll
  for (int i = 1; i < W; ++i) h = fmix64(h ^ (s.w[i] + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1)));
  return h ? h : 1;
}
__host__ __device__ __forceinline__ uint64_t fmix64b(uint64_t x) {     // a second, unrelated finaliser (splitmix64's)
  x ^= x >> 30;
  x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27;
  x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ Key128 key_of(const State& s, uint64_t fp) {
  Key128 k;
  if (W == 2 && !M::ALL_ONES_POSSIBLE) {
    k.lo = s.w[0];
    k.hi = s.w[W - 1];
  } else {
    uint64_t h = fmix64b(s.w[0] ^ 0xD6E8FEB86659FD93ull);
#pragma unroll
    for (int i = 1; i < W; ++i) h = fmix64b((h << 7 | h >> 57) ^ s.w[i]);
    k.lo = fp;
    k.hi = h;
    if (key_empty(k)) k.hi ^= 1;
  }
  return k;
}

// With SYMMETRY the identity is that of the orbit representative (smallest packed image under the symmetry
// group, TLC's symmetry reduction); the state that is stored, expanded and shown in traces stays the one that
// was actually reached.  canonicalize() is a few thousand instructions (n!-1 permuted images); kept out of
// line so that it exists once per kernel instead of once per call site (nvcc time of a symmetric model: 16 min -> 2 min).
struct Ident {
  uint64_t fp;
  Key128 key;
};
__host__ __device__ __noinline__ void canonical_ident(const State& s, Ident& id) {
  State c;
  M::canonicalize(s, c);
  id.fp = fingerprint(c);
  id.key = key_of(c, id.fp);
}
__host__ __device__ __forceinline__ Ident state_ident(const State& s) {
  Ident id;
  if (M::HAS_SYMMETRY) {
    canonical_ident(s, id);
  } else {
    id.fp = fingerprint(s);
    id.key = key_of(s, id.fp);
  }
  return id;
}
__host__ __device__ __forceinline__ uint64_t state_fp(const State& s) { return state_ident(s).fp; }

__host__ __device__ __forceinline__ uint32_t owner_of(uint64_t fp, uint32_t world) {
  return (uint32_t)(((fp >> 32) * (uint64_t)world) >> 32);
}

// ----------------------------------------------------------------------------------------
// device-side counters
// ----------------------------------------------------------------------------------------
struct DevCounters {
  unsigned long long cand_count[MAX_WORLD];
  unsigned long long store_tail;
  unsigned long long generated;
  unsigned long long deadlocks;
  unsigned long long out_of_model;
  unsigned long long probes;
  unsigned long long fail;
  unsigned long long max_fanout_seen;
  unsigned long long viol_count;         // rows claimed in the violator ring
  unsigned long long action_counts[64];
};

// Violating states are rare and terminal, so they go to a small ring: W state words, the
// parent/action word, the fingerprint and the invariant index (~0 = deadlock).  The host picks
// the entry with the smallest fingerprint -> the reported counterexample is deterministic
// (as long as the first violating level has <= VIOL_RING violators).
static constexpr int VIOL_RING = 1 << 16;     // rows; (W + 3) * 8 bytes each: 2.6 MB for a two-word model
static constexpr int VIOL_ROW = W + 3;

struct Params {
  void* table;              // 8-byte slots (one-word models) or 16-byte slots
  uint64_t bucket_mask;     // #buckets - 1
  uint64_t* store;
  uint64_t* parent;
  uint64_t max_states;      // capacity of the device-resident store (a ring when spilling)
  uint64_t store_mask;      // spill: max_states - 1 (power of two), device slot = global index & mask; else ~0
  uint64_t store_base;      // spill: global index of the oldest state still on the device (older ones are on the host)
  uint64_t* cand;
  uint64_t region_rows;
  DevCounters* ctr;
  uint64_t* viol_ring;
  uint32_t rank, world;
  uint32_t check_deadlock;
  uint32_t count_actions;
  uint32_t prefetch;        // K1 issues an L2 prefetch of every candidate's first bucket (see flush_stage)
  uint32_t cand_slot;       // single rank: which half of the candidate buffer (and which counter) this launch uses
  // fused exchange (world > 1, after kmc_shard_open_peers): every rank's inbox, mapped into this
  // process through CUDA IPC.  An inbox is two buffers (double buffering); a buffer is an 8-word
  // header (rows sent by each source rank) followed by world regions of region_rows rows.
  uint64_t* peer_inbox[MAX_WORLD];
  uint64_t inbox_stride;    // words per inbox buffer
  uint32_t p2p;             // 1: expand stores rows straight into the owners' inboxes
  uint32_t inbox_buf;       // which of the two buffers this round uses
};

static constexpr int INBOX_HEADER = 8;
// Sync page in front of every rank's inbox (same allocation, so peers map it with the inbox).  Peers PUSH into it
// (posted NVLink stores) and its owner polls it locally:
//   ready[src]   round number src has finished storing rows (and their counts) for, into this rank's inbox
//   done[dst]    round number dst has finished inserting from ITS inbox (so its buffer of that round may be reused)
//   board[r][8]  rank r's level summary {level id, new states, violations, store tail, generated, fail, deadlocks, -}
// All counters are monotonic over the life of the context (never reset), so no reset can race with a peer.
static constexpr int SYNC_WORDS = 256;
static constexpr int SYNC_READY = 0, SYNC_DONE = 8, SYNC_BOARD = 16, BOARD_WORDS = 8;

__device__ __forceinline__ unsigned lane_id() {
  unsigned r;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(r));
  return r;
}

__device__ __noinline__ void record_violation(const Params& p, const State& s, uint64_t meta, uint64_t fp, uint64_t inv) {
  unsigned long long slot = atomicAdd(&p.ctr->viol_count, 1ull);
  if (slot >= (unsigned long long)VIOL_RING) return;
  uint64_t* row = p.viol_ring + slot * VIOL_ROW;
#pragma unroll
  for (int k = 0; k < W; ++k) row[k] = s.w[k];
  row[W] = meta;
  row[W + 1] = fp;
  row[W + 2] = inv;
}

// ----------------------------------------------------------------------------------------
// K1: expand
// ----------------------------------------------------------------------------------------
// ----------------------------------------------------------------------------------------
// K2 primitives: the fingerprint set
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ ulonglong2 ld_cg128(const void* p) {
  // L2-coherent 128-bit load, no L1 allocation: buckets are touched once per probe
  ulonglong2 v;
  asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ uint64_t bucket_of(uint64_t fp, uint64_t bucket_mask) { return (fp ^ (fp >> 31)) & bucket_mask; }

// One bucket = BUCKET_SLOTS slots, loaded with 128-bit loads: 8-byte slots -> 2 loads of 2 slots each (32 B);
// 16-byte slots -> one load per slot.
static constexpr int BUCKET_LOADS = KEY128 ? BUCKET_SLOTS : 2;
struct Bucket {
  ulonglong2 v[BUCKET_LOADS];
};
__device__ __forceinline__ const char* bucket_addr(const void* table, uint64_t b) {
  return static_cast<const char*>(table) + b * (uint64_t)(BUCKET_SLOTS * SLOT_BYTES);
}
__device__ __forceinline__ Bucket ld_bucket(const void* table, uint64_t b) {
  Bucket k;
  const char* base = bucket_addr(table, b);
#pragma unroll
  for (int i = 0; i < BUCKET_LOADS; ++i) k.v[i] = ld_cg128(base + 16 * i);
  return k;
}

// returns 1 = inserted (new), 0 = already present, -1 = table full.  `bk` = the first bucket, loaded by the
// caller ahead of time.  Correctness of the lock-free insert: slots never return to empty and every inserter
// of a key scans the same probe sequence without skipping an unverified slot, so a key occupies at most one
// slot; a stale read is harmless (the CAS decides).
__device__ __forceinline__ int set_insert_pre(void* table, uint64_t bucket_mask, const Ident& id, Bucket bk, unsigned& probes) {
  uint64_t b = bucket_of(id.fp, bucket_mask);
  for (int attempt = 0; attempt < 512; ++attempt) {
    if (attempt) bk = ld_bucket(table, b);
    ++probes;
    if constexpr (KEY128) {
      Key128* base = reinterpret_cast<Key128*>(const_cast<char*>(bucket_addr(table, b)));
#pragma unroll
      for (int k = 0; k < BUCKET_SLOTS; ++k)
        if (bk.v[k].x == id.key.lo && bk.v[k].y == id.key.hi) return 0;
#pragma unroll
      for (int k = 0; k < BUCKET_SLOTS; ++k) {
        if ((bk.v[k].x & bk.v[k].y) == ~0ull) {
          const Key128 empty{~0ull, ~0ull};
          Key128 old = atomicCAS(base + k, empty, id.key);          // ATOMG.E.CAS.128
          if (key_empty(old)) return 1;
          if (key_eq(old, id.key)) return 0;
          // another key took the slot: keep scanning (slots never empty again)
        }
      }
    } else {
      uint64_t* base = reinterpret_cast<uint64_t*>(const_cast<char*>(bucket_addr(table, b)));
      const uint64_t fp = id.fp;
      uint64_t v[4] = {bk.v[0].x, bk.v[0].y, bk.v[1].x, bk.v[1].y};
      if (v[0] == fp || v[1] == fp || v[2] == fp || v[3] == fp) return 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (v[k] == 0) {
          unsigned long long old = atomicCAS((unsigned long long*)(base + k), 0ull, (unsigned long long)fp);
          if (old == 0) return 1;
          if (old == fp) return 0;
        }
      }
    }
    b = (b + 1) & bucket_mask;
  }
  return -1;
}

__device__ __forceinline__ int set_insert(void* table, uint64_t bucket_mask, const Ident& id, unsigned& probes) {
  return set_insert_pre(table, bucket_mask, id, ld_bucket(table, bucket_of(id.fp, bucket_mask)), probes);
}

__device__ __forceinline__ int set_contains(const void* table, uint64_t bucket_mask, const Ident& id) {
  uint64_t b = bucket_of(id.fp, bucket_mask);
  for (int attempt = 0; attempt < 512; ++attempt) {
    Bucket bk = ld_bucket(table, b);
    bool any_empty = false;
    if constexpr (KEY128) {
#pragma unroll
      for (int k = 0; k < BUCKET_SLOTS; ++k) {
        if (bk.v[k].x == id.key.lo && bk.v[k].y == id.key.hi) return 1;
        any_empty |= (bk.v[k].x & bk.v[k].y) == ~0ull;
      }
    } else {
      const uint64_t fp = id.fp;
      if (bk.v[0].x == fp || bk.v[0].y == fp || bk.v[1].x == fp || bk.v[1].y == fp) return 1;
      any_empty = bk.v[0].x == 0 || bk.v[0].y == 0 || bk.v[1].x == 0 || bk.v[1].y == 0;
    }
    if (any_empty) return 0;
    b = (b + 1) & bucket_mask;
  }
  return 0;
}

// Warp-collective insert of one candidate row per lane (invalid lanes pass valid = false):
// constraint check, identity, bucket probe + CAS, ballot/popc compaction of the winners into the
// state store, parent link.
struct Prefetched {
  Ident id;
  Bucket bk;
  bool inmodel;
};

// first half of an insert: identity + issue the bucket loads (no dependent use yet)
__device__ __forceinline__ Prefetched prefetch_row(const Params& p, const State& s, bool valid) {
  Prefetched f;
  f.id.fp = 0;
  f.id.key = Key128{0, 0};
#pragma unroll
  for (int i = 0; i < BUCKET_LOADS; ++i) f.bk.v[i] = make_ulonglong2(0, 0);
  f.inmodel = false;
  if (valid) {
    f.inmodel = (M::NUM_CONSTRAINTS == 0) || M::in_model(s);
    f.id = state_ident(s);
    if (f.inmodel) f.bk = ld_bucket(p.table, bucket_of(f.id.fp, p.bucket_mask));
  }
  return f;
}

__device__ __forceinline__ void insert_row(const Params& p, const State& s, uint64_t meta, bool valid, const Prefetched& f,
                                            unsigned& probes, unsigned& oom, int& failed) {
  bool is_new = false;
  if (valid) {
    if (f.inmodel) {
      int r = set_insert_pre(p.table, p.bucket_mask, f.id, f.bk, probes);
      if (r < 0) failed = KMC_FAIL_TABLE_FULL;
      is_new = r > 0;
    } else {
      ++oom;
      // TLC also checks invariants on successors discarded by a CONSTRAINT; they are not stored,
      // so that (rare) case is handled here.  New in-model states are checked by k_invariants (K3).
      if (M::NUM_INVARIANTS > 0) {
        int inv = M::first_violated_invariant(s);
        if (inv >= 0) record_violation(p, s, meta, f.id.fp, (uint64_t)inv);
      }
    }
  }
  unsigned mask = __ballot_sync(0xffffffffu, is_new);
  if (mask) {
    unsigned lane = lane_id();
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if ((int)lane == leader) base = atomicAdd(&p.ctr->store_tail, (unsigned long long)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (is_new) {
      uint64_t idx = base + __popc(mask & ((1u << lane) - 1));
      if (idx - p.store_base < p.max_states) {
        uint64_t* dst = p.store + (idx & p.store_mask) * W;
#pragma unroll
        for (int k = 0; k < W; ++k) dst[k] = s.w[k];
        p.parent[idx & p.store_mask] = meta;
      } else {
        failed = KMC_FAIL_STORE_FULL;
      }
    }
  }
}

// ----------------------------------------------------------------------------------------
// K1: expand
// ----------------------------------------------------------------------------------------
// Successor rows are staged per warp in shared memory and flushed in bulk: one global slot claim
// per flush (instead of one ~600-cycle atomic round trip per emit site), coalesced row stores,
// and -- multi-rank -- the owner computation (a fingerprint) done with all 32 lanes busy instead
// of inside the emit site.  The stage is addressed through 32-bit shared-window addresses and
// st.shared / atom.shared so that no generic-address store (ST + QSPC) is ever generated.
#ifndef EXPAND_BLOCK_THREADS
#define EXPAND_BLOCK_THREADS 1024
#endif
#ifndef EXPAND_CTAS_PER_SM
#define EXPAND_CTAS_PER_SM 1          // 2 (with 512 threads): one CTA's barrier waits are covered by the other CTA
#endif
static constexpr int EXPAND_BLOCK = EXPAND_BLOCK_THREADS;
static constexpr int NWARPS = EXPAND_BLOCK / 32;
static constexpr int STAGE_ROWS = 64;            // rows per warp; a body pass adds <= 32, a flush empties it
static constexpr int STAGE_FLUSH = 32;           // flush once at least this many rows are staged
static constexpr int LIST_CAP = 12288;           // (state, site) pairs of one scatter round, 16-bit tile slots
static constexpr int MAX_GROUP_SITES = 64;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts64(uint32_t a, uint64_t v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t lds64(uint32_t a) {
  uint64_t v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ unsigned atoms_add(uint32_t a, unsigned v) {
  unsigned old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(a), "r"(v) : "memory");
  return old;
}

// one row per lane (valid lanes only) into the owner's region, slot claims aggregated per owner
__device__ __forceinline__ void claim_and_store(const Params& p, const State& s, uint64_t meta, bool valid, int& failed) {
  uint32_t dest = 0xFFu;
  if (valid) dest = owner_of(state_fp(s), p.world);
  unsigned peers = __match_any_sync(0xffffffffu, dest);
  unsigned lane = lane_id();
  int leader = __ffs(peers) - 1;
  unsigned long long base = 0;
  if (valid && (int)lane == leader) base = atomicAdd(&p.ctr->cand_count[dest], (unsigned long long)__popc(peers));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (!valid) return;
  unsigned long long pos = base + __popc(peers & ((1u << lane) - 1));
  if (pos >= p.region_rows) {
    failed = KMC_FAIL_CAND_FULL;
    return;
  }
  // p2p: the row goes straight into region `rank` of the owner's inbox (a peer store over NVLink
  // when dest != rank); otherwise into the local per-owner candidate region for a later exchange
  uint64_t* row = p.p2p ? p.peer_inbox[dest] + (uint64_t)p.inbox_buf * p.inbox_stride + INBOX_HEADER +
                              ((uint64_t)p.rank * p.region_rows + pos) * ROW
                        : p.cand + ((uint64_t)dest * p.region_rows + pos) * ROW;
#pragma unroll
  for (int i = 0; i < W; ++i) row[i] = s.w[i];
  row[W] = meta;
}

// called by all 32 lanes of a warp at a converged point
__device__ __forceinline__ void flush_stage(const Params& p, uint32_t wbuf, uint32_t wcnt, bool force, int& failed) {
  __syncwarp();
  unsigned n;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(n) : "r"(wcnt));
  if (n > (unsigned)STAGE_ROWS) n = STAGE_ROWS;
  if (n == 0 || (!force && n < (unsigned)STAGE_FLUSH)) return;
  unsigned lane = lane_id();
  if (p.world == 1) {
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&p.ctr->cand_count[p.cand_slot], (unsigned long long)n);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base + n > p.region_rows) {
      failed = KMC_FAIL_CAND_FULL;
    } else {
      uint64_t* dst = p.cand + base * ROW;
      for (unsigned k = lane; k < n * ROW; k += 32) dst[k] = lds64(wbuf + k * 8);      // coalesced
      if (p.prefetch && !M::HAS_SYMMETRY) {
        // Software pipelining across kernels through the L2: the bucket this candidate will probe in K2 is
        // requested now, while K1 still has integer work to hide the DRAM latency behind.  With frontier chunks
        // sized so that a chunk's buckets fit the 126 MB L2, K2 then probes L2-resident sectors.
        for (unsigned r = lane; r < n; r += 32) {
          State t;
#pragma unroll
          for (int q = 0; q < W; ++q) t.w[q] = lds64(wbuf + (r * ROW + q) * 8);
          const char* a = bucket_addr(p.table, bucket_of(fingerprint(t), p.bucket_mask));
          asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a));
        }
      }
    }
  } else {
    for (unsigned r0 = 0; r0 < n; r0 += 32) {
      unsigned r = r0 + lane;
      bool valid = r < n;
      State s;
      const uint32_t row = wbuf + (valid ? r : 0) * (ROW * 8);
#pragma unroll
      for (int k = 0; k < W; ++k) s.w[k] = lds64(row + k * 8);
      claim_and_store(p, s, lds64(row + W * 8), valid, failed);
    }
  }
  __syncwarp();
  if (lane == 0) asm volatile("st.shared.u32 [%0], %1;" ::"r"(wcnt), "r"(0u) : "memory");
  __syncwarp();
}

// The sink handed to the lowered bodies.  It holds values only (no reference to the kernel parameters and no
// out-of-line member): anything that makes the object addressable sends it -- and every pointer in it --
// through local memory and turns the stage stores / counter atomics into generic-address operations.
struct CandSink {
  uint64_t parent_ref;
  uint32_t wbuf;       // this warp's staging rows (shared-window address)
  uint32_t wcnt;       // rows staged by this warp (shared-window address)
  unsigned long long* action_counts;   // per-action counters (global memory), or nullptr
  int n;
  int failed;
#ifdef KMC_ONE_PHASE
  uint64_t* cand;                      // one-phase comparison kernel only (single rank): overflow rows go straight
  unsigned long long* cand_count;      // to the candidate buffer
  uint64_t region_rows;
#endif

  __device__ __forceinline__ void emit(const State& s, int action) {
    ++n;
    const uint64_t meta = parent_ref | ((uint64_t)action << 56);
    unsigned active = __activemask();
    unsigned lane = lane_id();
    int leader = __ffs(active) - 1;
    unsigned base = 0;
    if ((int)lane == leader) base = atoms_add(wcnt, (unsigned)__popc(active));
    base = __shfl_sync(active, base, leader);
    unsigned pos = base + __popc(active & ((1u << lane) - 1));
    if (pos < (unsigned)STAGE_ROWS) {
      const uint32_t row = wbuf + pos * (ROW * 8);
#pragma unroll
      for (int i = 0; i < W; ++i) sts64(row + i * 8, s.w[i]);
      sts64(row + W * 8, meta);
    } else {
#ifdef KMC_ONE_PHASE
      if (cand != nullptr) {           // one-phase kernel: a burst of emits between two flushes
        unsigned long long gpos = atomicAdd(cand_count, 1ull);
        if (gpos >= region_rows) {
          failed = KMC_FAIL_CAND_FULL;
        } else {
          uint64_t* grow = cand + gpos * ROW;
#pragma unroll
          for (int i = 0; i < W; ++i) grow[i] = s.w[i];
          grow[W] = meta;
        }
      } else
#endif
      // two-phase kernel: cannot happen (a body pass adds <= 32 rows to a stage that is flushed at >= STAGE_FLUSH)
      failed = KMC_FAIL_CAND_FULL;
    }
    if (action_counts != nullptr && action < 64) atomicAdd(action_counts + action, 1ull);
  }
  __device__ __forceinline__ void fail(int code) { failed = code; }
};

__device__ __forceinline__ void load_state(State& s, const uint64_t* src) {
#pragma unroll
  for (int k = 0; k < W; ++k) s.w[k] = __ldg(src + k);
}

// ----------------------------------------------------------------------------------------
// K1, two-phase.  The round-1 kernel ran every thread through the whole lowered Next of its own
// states: ncu showed 12 of 32 lanes per issued instruction, because an action body is enabled for
// a few per cent of the states of a warp.  Here the lowering provides, per site group (<= 64 emit
// sites), site_mask(s) = the COMPLETE path condition of every site (cheap compares, evaluated by
// all lanes on their own states) and site_body<i>(s) = the straight-line successor construction.
// One 1024-thread CTA per SM works on a tile of EXPAND_BLOCK x SPT states held in shared memory:
//   A1  every thread evaluates the group's masks for its SPT states; per-site totals via ballot/popc
//       and one shared-memory atomic per warp and site                                  -- barrier --
//   A2  every warp derives the same segment offsets from the totals (segments padded to 32) and
//       scatters its enabled (site, tile slot) pairs into the sorted list                -- barrier --
//   B   the list is consumed 32 entries at a time; a chunk belongs to exactly one site, so the body
//       dispatch is warp-uniform and the body runs with (almost) all lanes on the same code.
//       B of group g overlaps A1 of group g+1 (double-buffered totals): two barriers per group.
// Successor counts per state (deadlock detection, "states generated") are popc(mask).
// ----------------------------------------------------------------------------------------
static constexpr int STAGE_BYTES = NWARPS * STAGE_ROWS * ROW * 8;
static constexpr int FIXED_SMEM_BYTES = STAGE_BYTES + LIST_CAP * 2 + (4 * MAX_GROUP_SITES + NWARPS + 8) * 4;
static constexpr int SPT_FIT = (227 * 1024 / EXPAND_CTAS_PER_SM - 1024 - FIXED_SMEM_BYTES) / (EXPAND_BLOCK * W * 8);
static constexpr int SPT = SPT_FIT > 4 ? 4 : SPT_FIT;
static_assert(SPT >= 1, "state too wide for the expand kernel's shared-memory tile");
static constexpr int TILE = EXPAND_BLOCK * SPT;
static_assert(TILE <= LIST_CAP, "a site's segment (<= TILE pairs) must fit one scatter round");
static constexpr size_t EXPAND_SMEM_BYTES = (size_t)TILE * W * 8 + FIXED_SMEM_BYTES;

struct TileCtx {
  uint32_t tile;        // [TILE][W] states (shared-window addresses throughout)
  uint32_t list;        // [LIST_CAP] u16 tile slots, per-site segments
  uint32_t cnt;         // [2][MAX_GROUP_SITES] enabled pairs per site (double-buffered across groups)
  uint32_t cur;         // [MAX_GROUP_SITES] scatter cursors
  uint32_t seg;         // [MAX_GROUP_SITES] first chunk of each site's segment
  uint32_t wbuf, wcnt;
  uint64_t first, tile_base;
  unsigned nvalid;      // states in this tile
};

template <int LO, int HI>
struct SiteDispatch {
  static __device__ __forceinline__ void run(int i, const State& s, CandSink& sink) {
    if constexpr (HI - LO == 1) {
      // the opaque copy keeps the compiler from hoisting every body's unpacking above the dispatch
      State t = s;
#pragma unroll
      for (int k = 0; k < W; ++k) asm volatile("" : "+l"(t.w[k]));
      M::site_body(M::SiteTag<LO>{}, t, sink);
    } else {
      constexpr int MID = (LO + HI) / 2;
      if (i < MID) SiteDispatch<LO, MID>::run(i, s, sink);
      else SiteDispatch<MID, HI>::run(i, s, sink);
    }
  }
};

__device__ __forceinline__ unsigned lds32(uint32_t a) {
  unsigned v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

template <int G>
struct SiteGroupRunner {
  static __device__ __forceinline__ void run(const Params& p, const TileCtx& c, unsigned (&nsucc)[SPT], int& failed) {
    constexpr int BEGIN = M::SITE_GROUP_BEGIN[G];
    constexpr int END = M::SITE_GROUP_BEGIN[G + 1];
    constexpr int NS = END - BEGIN;
    static_assert(NS >= 1 && NS <= MAX_GROUP_SITES, "site group size");
    const unsigned lane = lane_id();
    const unsigned warp = threadIdx.x >> 5;
    const uint32_t cnt = c.cnt + (G & 1) * (MAX_GROUP_SITES * 4);
    // ---- A1: masks of this thread's states; every enabled (state, site) pair bumps the site's total.
    // (One shared-memory atomic per pair: ~3 per state.  A ballot/popc census over all sites and states cost
    // 60 % of the kernel's instructions in the first version of this kernel, profiles/README.md r2a.)
    uint64_t masks[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const unsigned slot = (unsigned)j * EXPAND_BLOCK + threadIdx.x;
      masks[j] = 0;
      if (slot < c.nvalid) {
        State s;
#pragma unroll
        for (int k = 0; k < W; ++k) s.w[k] = lds64(c.tile + (slot * W + k) * 8);
        masks[j] = M::site_mask(M::SiteGroupTag<G>{}, s);
        nsucc[j] += (unsigned)__popcll(masks[j]);
      }
    }
    // a state enables ~0.6 sites of a group on average: three predicated steps (no loop control, no
    // reconvergence stack) cover nearly every mask; the loop behind them is the rare tail
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      uint64_t m = masks[j];
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        if (m) {
          atoms_add(cnt + (__ffsll((long long)m) - 1) * 4, 1u);
          m &= m - 1;
        }
      }
      while (m) {
        atoms_add(cnt + (__ffsll((long long)m) - 1) * 4, 1u);
        m &= m - 1;
      }
    }
    __syncthreads();                                   // totals complete; B of the previous group finished
    // the other totals buffer (read last by B of group G-1) is cleared for A1 of group G+1
    if (threadIdx.x < MAX_GROUP_SITES) sts32(c.cnt + ((G + 1) & 1) * (MAX_GROUP_SITES * 4) + threadIdx.x * 4, 0u);
    // ---- every warp: the same padded segment layout, sites 2*lane and 2*lane+1 per lane
    const unsigned c0 = (2 * lane < (unsigned)NS) ? lds32(cnt + (2 * lane) * 4) : 0u;
    const unsigned c1 = (2 * lane + 1 < (unsigned)NS) ? lds32(cnt + (2 * lane + 1) * 4) : 0u;
    const unsigned ch0 = (c0 + 31) >> 5, ch1 = (c1 + 31) >> 5;       // chunks of 32 pairs
    unsigned incl = ch0 + ch1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)lane >= o) incl += t;
    }
    const unsigned start0 = incl - ch0 - ch1, start1 = start0 + ch0;  // first chunk of each site
    const unsigned total_chunks = __shfl_sync(0xffffffffu, incl, 31);
    // segment starts where every thread of the CTA can look them up by site (all warps store the same values)
    sts32(c.seg + (2 * lane) * 4, start0);
    sts32(c.seg + (2 * lane + 1) * 4, start1);
    __syncwarp();
    // Scatter rounds.  The list holds LIST_CAP pairs; a round covers the chunks [r0, r_end) = the sites [k_lo, k_hi),
    // and a segment is never split across rounds (a segment has <= TILE/32 chunks, so every round makes
    // progress).  One round is the rule; more are needed only when a tile enables more than LIST_CAP pairs here.
    constexpr unsigned ROUND_CHUNKS = LIST_CAP / 32;
    unsigned r0 = 0;
#pragma unroll 1
    do {
      // end of this round: the start of the first segment that does not fit any more
      unsigned r_end = total_chunks;
      if (total_chunks > r0 + ROUND_CHUNKS) {
        unsigned cand0 = (ch0 && start0 >= r0 && start0 + ch0 > r0 + ROUND_CHUNKS) ? start0 : 0xFFFFFFFFu;
        unsigned cand1 = (ch1 && start1 >= r0 && start1 + ch1 > r0 + ROUND_CHUNKS) ? start1 : 0xFFFFFFFFu;
        unsigned nxt = min(cand0, cand1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nxt = min(nxt, __shfl_xor_sync(0xffffffffu, nxt, o));
        r_end = min(r_end, nxt);
      }
      // sites of this round = those whose (non-empty) segment starts in [r0, r_end): a contiguous index range
      const unsigned in0 = __ballot_sync(0xffffffffu, ch0 && start0 >= r0 && start0 < r_end);
      const unsigned in1 = __ballot_sync(0xffffffffu, ch1 && start1 >= r0 && start1 < r_end);
      int k_lo = 64, k_hi = 0;
      if (in0) { k_lo = min(k_lo, 2 * (__ffs(in0) - 1)); k_hi = max(k_hi, 2 * (31 - __clz(in0)) + 1); }
      if (in1) { k_lo = min(k_lo, 2 * (__ffs(in1) - 1) + 1); k_hi = max(k_hi, 2 * (31 - __clz(in1)) + 2); }
      if (r0) __syncthreads();                         // later rounds: the previous round's list is consumed
      // ---- A2: every thread scatters its own pairs: slot = cursor[site]++ inside the site's segment
      auto scatter_one = [&](uint64_t& m, int j) {
        const int k = __ffsll((long long)m) - 1;
        m &= m - 1;
        if (k >= k_lo && k < k_hi) {                   // (a later round takes the sites outside)
          const unsigned st = lds32(c.seg + k * 4);
          const unsigned pos = (st - r0) * 32 + atoms_add(c.cur + k * 4, 1u);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(c.list + pos * 2), "h"((unsigned short)((unsigned)j * EXPAND_BLOCK + threadIdx.x)) : "memory");
        }
      };
#pragma unroll
      for (int j = 0; j < SPT; ++j) {
        uint64_t m = masks[j];
#pragma unroll
        for (int it = 0; it < 3; ++it)
          if (m) scatter_one(m, j);
        while (m) scatter_one(m, j);
      }
      __syncthreads();                                 // list complete (also orders the clearing of the other totals buffer)
      if (threadIdx.x < MAX_GROUP_SITES) sts32(c.cur + threadIdx.x * 4, 0u);   // cursors ready for the next scatter
      // ---- B: bodies, one chunk (= 32 pairs of one site) per warp and step
#pragma unroll 1
      for (unsigned ch = r0 + warp; ch < r_end; ch += NWARPS) {
        // site of this chunk: the last site whose first chunk is <= ch (an empty site shares its successor's start)
        const int below = __popc(__ballot_sync(0xffffffffu, start0 <= ch && 2 * lane < (unsigned)NS)) +
                          __popc(__ballot_sync(0xffffffffu, start1 <= ch && 2 * lane + 1 < (unsigned)NS));
        const int k = below - 1;
        const unsigned st = __shfl_sync(0xffffffffu, (k & 1) ? start1 : start0, k >> 1);
        const unsigned ck = __shfl_sync(0xffffffffu, (k & 1) ? c1 : c0, k >> 1);
        const unsigned e = (ch - st) * 32 + lane;      // index inside the site's segment
        if (e < ck) {
          unsigned short slot16;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(slot16) : "r"(c.list + ((st - r0) * 32 + e) * 2));
          const unsigned slot = slot16;
          State s;
#pragma unroll
          for (int q = 0; q < W; ++q) s.w[q] = lds64(c.tile + (slot * W + q) * 8);
          CandSink sink{(c.first + c.tile_base + slot) | ((uint64_t)p.rank << 40), c.wbuf, c.wcnt,
                        p.count_actions ? p.ctr->action_counts : nullptr, 0, 0
#ifdef KMC_ONE_PHASE
                        , nullptr, nullptr, 0
#endif
          };
          SiteDispatch<BEGIN, END>::run(k + BEGIN, s, sink);
          failed |= sink.failed;
        }
        flush_stage(p, c.wbuf, c.wcnt, false, failed);
      }
      r0 = r_end;
    } while (r0 < total_chunks);
    SiteGroupRunner<G + 1>::run(p, c, nsucc, failed);
  }
};
template <>
struct SiteGroupRunner<M::NUM_SITE_GROUPS> {
  static __device__ __forceinline__ void run(const Params&, const TileCtx&, unsigned (&)[SPT], int&) {}
};

__global__ void __launch_bounds__(EXPAND_BLOCK, EXPAND_CTAS_PER_SM) k_expand(Params p, uint64_t first, uint64_t count, unsigned tile_states) {
  extern __shared__ __align__(16) uint64_t smem[];   // tile | stage | list | cnt[2][64] | cur[64] | seg[64] | wcnt[NWARPS]
  const int warp = threadIdx.x >> 5;
  TileCtx c;
  c.tile = smem_addr(smem);
  const uint32_t stage = c.tile + TILE * W * 8;
  c.wbuf = stage + warp * (STAGE_ROWS * ROW * 8);
  c.list = stage + STAGE_BYTES;
  c.cnt = c.list + LIST_CAP * 2;
  c.cur = c.cnt + 2 * MAX_GROUP_SITES * 4;
  c.seg = c.cur + MAX_GROUP_SITES * 4;
  c.wcnt = c.seg + MAX_GROUP_SITES * 4 + warp * 4;
  c.first = first;
  if (lane_id() == 0) sts32(c.wcnt, 0u);
  unsigned long long gen = 0, dead = 0;
  unsigned maxfan = 0;
  int failed = 0;
  // tile_states <= TILE: small levels use smaller tiles so that every SM still gets one
  for (uint64_t tile_base = (uint64_t)blockIdx.x * tile_states; tile_base < count; tile_base += (uint64_t)gridDim.x * tile_states) {
    c.tile_base = tile_base;
    c.nvalid = (unsigned)min((uint64_t)tile_states, count - tile_base);
    __syncthreads();                                   // every body of the previous tile has read its state
    if (threadIdx.x < 3 * MAX_GROUP_SITES) sts32(c.cnt + threadIdx.x * 4, 0u);      // cnt[2][64] and cur[64]
    {
      // frontier tile -> shared memory, coalesced (128-bit loads when the rows are 16-byte aligned)
      const uint64_t* src = p.store + ((first + tile_base) & p.store_mask) * W;      // (a chunk never crosses the ring's wrap)
      const unsigned nwords = c.nvalid * W;
      if (((W & 1) == 0)) {
        const ulonglong2* src2 = reinterpret_cast<const ulonglong2*>(src);
        for (unsigned i = threadIdx.x; i < nwords / 2; i += EXPAND_BLOCK) {
          ulonglong2 v = __ldg(src2 + i);
          asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(c.tile + i * 16), "l"(v.x), "l"(v.y) : "memory");
        }
      } else {
        for (unsigned i = threadIdx.x; i < nwords; i += EXPAND_BLOCK) sts64(c.tile + i * 8, __ldg(src + i));
      }
    }
    __syncthreads();
    unsigned nsucc[SPT];
#pragma unroll
    for (int j = 0; j < SPT; ++j) nsucc[j] = 0;
    SiteGroupRunner<0>::run(p, c, nsucc, failed);
    flush_stage(p, c.wbuf, c.wcnt, true, failed);
#pragma unroll
    for (int j = 0; j < SPT; ++j) {
      const unsigned slot = (unsigned)j * EXPAND_BLOCK + threadIdx.x;
      if (slot < c.nvalid) {
        gen += nsucc[j];
        maxfan = max(maxfan, nsucc[j]);
        if (nsucc[j] == 0) {
          ++dead;
          if (p.check_deadlock) {
            State s;
            const uint64_t gi = (first + tile_base + slot) & p.store_mask;
            load_state(s, p.store + gi * W);
            record_violation(p, s, p.parent[gi], fingerprint(s), ~0ull);
          }
        }
      }
    }
  }
  // warp reduce the statistics, one atomic per warp
  for (int o = 16; o > 0; o >>= 1) {
    gen += __shfl_xor_sync(0xffffffffu, gen, o);
    dead += __shfl_xor_sync(0xffffffffu, dead, o);
    maxfan = max(maxfan, __shfl_xor_sync(0xffffffffu, maxfan, o));
    failed = max(failed, __shfl_xor_sync(0xffffffffu, failed, o));
  }
  if (lane_id() == 0) {
    if (gen) atomicAdd(&p.ctr->generated, gen);
    if (dead) atomicAdd(&p.ctr->deadlocks, dead);
    if (maxfan) atomicMax(&p.ctr->max_fanout_seen, (unsigned long long)maxfan);
    if (failed) atomicCAS(&p.ctr->fail, 0ull, (unsigned long long)failed);
  }
}

#ifdef KMC_ONE_PHASE
// ----------------------------------------------------------------------------------------
// K1, one-phase (round 1): comparison build only (-DKMC_ONE_PHASE, option "one_phase").  The
// lowered Next is cut into NUM_GROUPS groups of ~1k instructions and the CTA sweeps ONE group at
// a time over a tile of EXPAND_BLOCK x spt states (__syncthreads between groups keeps all warps
// of the SM in the same group, so a fetched instruction line serves every warp).
// ----------------------------------------------------------------------------------------
static constexpr int EXPAND_SPT1 = 4;
template <int G>
struct GroupRunner {
  static __device__ __forceinline__ void run(const Params& p, uint64_t first, uint64_t tile_base, uint64_t count,
                                              int spt, unsigned (&nsucc)[EXPAND_SPT1], int& failed, uint32_t wbuf, uint32_t wcnt) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < EXPAND_SPT1; ++j) {
      if (j < spt) {
        uint64_t i = tile_base + (uint64_t)j * EXPAND_BLOCK + threadIdx.x;
        if (i < count) {
          State s;
          load_state(s, p.store + ((first + i) & p.store_mask) * W);
          CandSink sink{(first + i) | ((uint64_t)p.rank << 40), wbuf, wcnt, p.count_actions ? p.ctr->action_counts : nullptr, 0, 0,
                        p.world == 1 ? p.cand : nullptr, &p.ctr->cand_count[p.cand_slot], p.region_rows};
          M::expand_group(M::GroupTag<G>{}, s, sink);
          nsucc[j] += (unsigned)sink.n;
          failed |= sink.failed;
        }
        flush_stage(p, wbuf, wcnt, false, failed);
      }
    }
    GroupRunner<G + 1>::run(p, first, tile_base, count, spt, nsucc, failed, wbuf, wcnt);
  }
};
template <>
struct GroupRunner<M::NUM_GROUPS> {
  static __device__ __forceinline__ void run(const Params&, uint64_t, uint64_t, uint64_t, int, unsigned (&)[EXPAND_SPT1], int&,
                                              uint32_t, uint32_t) {}
};

__global__ void __launch_bounds__(EXPAND_BLOCK, 1) k_expand1(Params p, uint64_t first, uint64_t count, int spt) {
  extern __shared__ __align__(16) uint64_t smem[];          // [warps][STAGE_ROWS][ROW] then [warps] counters
  const int warp = threadIdx.x >> 5;
  const uint32_t wbuf = smem_addr(smem) + warp * (STAGE_ROWS * ROW * 8);
  const uint32_t wcnt = smem_addr(smem) + STAGE_BYTES + warp * 4;
  if (lane_id() == 0) sts32(wcnt, 0u);
  __syncwarp();
  unsigned long long gen = 0, dead = 0;
  unsigned maxfan = 0;
  int failed = 0;
  const uint64_t tile = (uint64_t)EXPAND_BLOCK * spt;
  for (uint64_t tile_base = (uint64_t)blockIdx.x * tile; tile_base < count; tile_base += (uint64_t)gridDim.x * tile) {
    unsigned nsucc[EXPAND_SPT1];
#pragma unroll
    for (int j = 0; j < EXPAND_SPT1; ++j) nsucc[j] = 0;
    GroupRunner<0>::run(p, first, tile_base, count, spt, nsucc, failed, wbuf, wcnt);
    flush_stage(p, wbuf, wcnt, true, failed);
#pragma unroll
    for (int j = 0; j < EXPAND_SPT1; ++j) {
      uint64_t i = tile_base + (uint64_t)j * EXPAND_BLOCK + threadIdx.x;
      if (j >= spt || i >= count) continue;
      gen += nsucc[j];
      if (nsucc[j] > maxfan) maxfan = nsucc[j];
      if (nsucc[j] == 0) {
        ++dead;
        if (p.check_deadlock) {
          State s;
          load_state(s, p.store + ((first + i) & p.store_mask) * W);
          record_violation(p, s, p.parent[(first + i) & p.store_mask], fingerprint(s), ~0ull);
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    gen += __shfl_xor_sync(0xffffffffu, gen, o);
    dead += __shfl_xor_sync(0xffffffffu, dead, o);
    maxfan = max(maxfan, __shfl_xor_sync(0xffffffffu, maxfan, o));
    failed = max(failed, __shfl_xor_sync(0xffffffffu, failed, o));
  }
  if (lane_id() == 0) {
    if (gen) atomicAdd(&p.ctr->generated, gen);
    if (dead) atomicAdd(&p.ctr->deadlocks, dead);
    if (maxfan) atomicMax(&p.ctr->max_fanout_seen, (unsigned long long)maxfan);
    if (failed) atomicCAS(&p.ctr->fail, 0ull, (unsigned long long)failed);
  }
}
#endif  // KMC_ONE_PHASE

__device__ __forceinline__ void load_row(State& s, uint64_t& meta, const uint64_t* rows, uint64_t i, bool valid) {
  meta = 0;
  if (valid) {
    const uint64_t* row = rows + i * ROW;
#pragma unroll
    for (int k = 0; k < W; ++k) s.w[k] = __ldcs(row + k);
    meta = __ldcs(row + W);
  }
}

// One candidate row per thread per iteration.  (Two rows per thread -- two sectors in flight per
// lane -- was measured slower: 72 vs 63 ms on the 340 M-state model; the extra registers cost more
// occupancy than the added memory-level parallelism gains.)
__global__ void __launch_bounds__(256) k_insert(Params p, const uint64_t* rows, const unsigned long long* n_ptr,
                                                 uint64_t n_fixed) {
  uint64_t n = n_ptr ? (uint64_t)*n_ptr : n_fixed;
  unsigned probes = 0, oom = 0;
  int failed = 0;
  if (n_ptr && n > p.region_rows) {
    // the expand kernel's slot claims ran past the region (it reports KMC_FAIL_CAND_FULL itself; the
    // counter keeps counting): never read beyond the rows that were actually written
    n = p.region_rows;
    failed = KMC_FAIL_CAND_FULL;
  }
  const uint64_t n_round = (n + 31) & ~31ull;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool v0 = i < n;
    State s0;
    uint64_t m0;
    load_row(s0, m0, rows, i, v0);
    Prefetched f0 = prefetch_row(p, s0, v0);
    insert_row(p, s0, m0, v0, f0, probes, oom, failed);
  }
  for (int o = 16; o > 0; o >>= 1) {
    probes += __shfl_xor_sync(0xffffffffu, probes, o);
    oom += __shfl_xor_sync(0xffffffffu, oom, o);
    failed = max(failed, __shfl_xor_sync(0xffffffffu, failed, o));
  }
  if (lane_id() == 0) {
    if (probes) atomicAdd(&p.ctr->probes, (unsigned long long)probes);
    if (oom) atomicAdd(&p.ctr->out_of_model, (unsigned long long)oom);
    if (failed) atomicCAS(&p.ctr->fail, 0ull, (unsigned long long)failed);
  }
}

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t* sync_page(const Params& p, uint32_t r) { return p.peer_inbox[r] - SYNC_WORDS; }

// Fused exchange, step 2: tell every owner how many rows this rank stored in its inbox region and (round > 0)
// raise this rank's ready flag there.  The expand kernel that stored the rows is complete (stream order); the
// system-scope fence + release store order its peer writes before the flag.
__global__ void k_publish_counts(Params p, uint64_t round) {
  unsigned d = threadIdx.x;
  if (d < p.world) {
    p.peer_inbox[d][(uint64_t)p.inbox_buf * p.inbox_stride + p.rank] = p.ctr->cand_count[d];
    if (round) {
      __threadfence_system();
      st_release_sys(sync_page(p, d) + SYNC_READY + p.rank, round);
    }
  }
}

// A device-side wait is bounded: a peer that never arrives (its process died, its context failed) turns into
// KMC_E_PEER_TIMEOUT on this rank after PEER_TIMEOUT_NS instead of a kernel that spins for ever and takes the
// GPU with it.  Once the flag is up, later waits of the run return at once (the run is lost anyway).
static constexpr unsigned long long PEER_TIMEOUT_NS = 30ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_flag(const uint64_t* flag, uint64_t value, DevCounters* ctr) {
  if (ld_acquire_sys(flag) >= value) return;
  if (*(volatile unsigned long long*)&ctr->fail == KMC_FAIL_PEER_TIMEOUT) return;
  const unsigned long long t0 = globaltimer_ns();
  unsigned ns = 32;
  while (ld_acquire_sys(flag) < value) {
    __nanosleep(ns);
    if (ns < 1024) ns <<= 1;
    if (globaltimer_ns() - t0 > PEER_TIMEOUT_NS) {
      atomicCAS(&ctr->fail, 0ull, (unsigned long long)KMC_FAIL_PEER_TIMEOUT);
      return;
    }
  }
}

// Device-side wait (one warp): lane i waits until flags[i] >= value.  The flags live in this rank's own memory
// (peers push), so the polling never crosses NVLink.
__global__ void k_wait_flags(const uint64_t* flags, unsigned n, uint64_t value, DevCounters* ctr) {
  unsigned i = threadIdx.x;
  if (i < n) wait_flag(flags + i, value, ctr);
}

// after the insert of a round: every source may now reuse this rank's inbox buffer of that round
__global__ void k_publish_done(Params p, uint64_t round) {
  unsigned s = threadIdx.x;
  if (s < p.world) {
    __threadfence_system();
    st_release_sys(sync_page(p, s) + SYNC_DONE + p.rank, round);
  }
}

// Level end: this rank's summary goes to every rank's board (peer stores), ...
__global__ void k_publish_level(Params p, uint64_t level_id, uint64_t prev_tail) {
  unsigned d = threadIdx.x;
  if (d < p.world) {
    uint64_t* e = sync_page(p, d) + SYNC_BOARD + p.rank * BOARD_WORDS;
    const unsigned long long tail = p.ctr->store_tail;
    e[1] = tail - prev_tail;
    e[2] = p.ctr->viol_count;
    e[3] = tail;
    e[4] = p.ctr->generated;
    e[5] = p.ctr->fail;
    e[6] = p.ctr->deadlocks;
    __threadfence_system();
    st_release_sys(e, level_id);
  }
}
// ... and once every rank's entry of this level has arrived the whole board is copied to pinned host memory:
// the host's single synchronisation per level is the stream sync after this kernel.
__global__ void k_gather_level(Params p, uint64_t level_id, uint64_t* host_out) {
  unsigned r = threadIdx.x;
  const uint64_t* board = sync_page(p, p.rank) + SYNC_BOARD;
  if (r < p.world) {
    wait_flag(board + r * BOARD_WORDS, level_id, p.ctr);
    for (int k = 0; k < BOARD_WORDS; ++k) host_out[r * BOARD_WORDS + k] = board[r * BOARD_WORDS + k];
    // a timed-out wait of this level (or of one of its rounds) reaches the host through this rank's own entry
    if (r == p.rank) {
      const unsigned long long f = *(volatile unsigned long long*)&p.ctr->fail;
      if (f == KMC_FAIL_PEER_TIMEOUT) host_out[r * BOARD_WORDS + 5] = f;
    }
  }
}

// Fused exchange, step 3 (after a cross-rank barrier): insert the rows of all source regions of
// this rank's inbox buffer.  Row counts come from the header the sources wrote -- the host never
// sees them.
__global__ void __launch_bounds__(256) k_insert_inbox(Params p) {
  const uint64_t* inbox = p.peer_inbox[p.rank] + (uint64_t)p.inbox_buf * p.inbox_stride;
  uint64_t starts[MAX_WORLD + 1];
  starts[0] = 0;
#pragma unroll
  for (int r = 0; r < MAX_WORLD; ++r) {
    uint64_t n = (r < (int)p.world) ? inbox[r] : 0;
    if (n > p.region_rows) n = p.region_rows;
    starts[r + 1] = starts[r] + n;
  }
  const uint64_t n = starts[MAX_WORLD];
  const uint64_t n_round = (n + 31) & ~31ull;
  unsigned probes = 0, oom = 0;
  int failed = 0;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
    const bool v0 = i < n;
    State s0;
    uint64_t m0 = 0;
    if (v0) {
      int src = 0;
#pragma unroll
      for (int r = 1; r < MAX_WORLD; ++r) src += (i >= starts[r]) ? 1 : 0;
      const uint64_t* row = inbox + INBOX_HEADER + ((uint64_t)src * p.region_rows + (i - starts[src])) * ROW;
#pragma unroll
      for (int k = 0; k < W; ++k) s0.w[k] = __ldcs(row + k);
      m0 = __ldcs(row + W);
    }
    Prefetched f0 = prefetch_row(p, s0, v0);
    insert_row(p, s0, m0, v0, f0, probes, oom, failed);
  }
  for (int o = 16; o > 0; o >>= 1) {
    probes += __shfl_xor_sync(0xffffffffu, probes, o);
    oom += __shfl_xor_sync(0xffffffffu, oom, o);
    failed = max(failed, __shfl_xor_sync(0xffffffffu, failed, o));
  }
  if (lane_id() == 0) {
    if (probes) atomicAdd(&p.ctr->probes, (unsigned long long)probes);
    if (oom) atomicAdd(&p.ctr->out_of_model, (unsigned long long)oom);
    if (failed) atomicCAS(&p.ctr->fail, 0ull, (unsigned long long)failed);
  }
}

// K3: invariants on the new states of a level.  They sit compacted in the store, so every lane
// has work (inside k_insert only the ~1/3 of lanes holding a new state would be active).
__global__ void __launch_bounds__(256) k_invariants(Params p, uint64_t first, const unsigned long long* end_ptr) {
  uint64_t end = (uint64_t)*end_ptr;
  if (end - p.store_base > p.max_states) end = p.store_base + p.max_states;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
    State s;
    const uint64_t* src = p.store + (i & p.store_mask) * W;
#pragma unroll
    for (int k = 0; k < W; ++k) s.w[k] = __ldcs(src + k);
    int inv = M::first_violated_invariant(s);
    if (inv >= 0) record_violation(p, s, p.parent[i & p.store_mask], fingerprint(s), (uint64_t)inv);
  }
}

// -recover: the set is not part of a checkpoint; it is rebuilt from the stored states (one insert each)
__global__ void __launch_bounds__(256) k_rebuild(Params p, const uint64_t* states, uint64_t n) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  int failed = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    State s;
#pragma unroll
    for (int k = 0; k < W; ++k) s.w[k] = states[i * W + k];
    unsigned probes = 0;
    if (set_insert(p.table, p.bucket_mask, state_ident(s), probes) < 0) failed = KMC_FAIL_TABLE_FULL;
  }
  if (failed) atomicCAS(&p.ctr->fail, 0ull, (unsigned long long)failed);
}

// The set alone (FPSet.put / contains): the caller's 64-bit fingerprints are the identities; with 16-byte
// slots the key is the fingerprint and its second mix.
__global__ void k_fpset_put(void* table, uint64_t bucket_mask, const uint64_t* fps, uint64_t n, uint8_t* seen,
                            DevCounters* ctr, int insert) {
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    Ident id;
    id.fp = fps[i] ? fps[i] : 1;
    id.key = Key128{id.fp, fmix64b(id.fp)};
    if (insert) {
      unsigned probes = 0;
      int r = set_insert(table, bucket_mask, id, probes);
      if (r < 0) atomicCAS(&ctr->fail, 0ull, (unsigned long long)KMC_FAIL_TABLE_FULL);
      if (r > 0) atomicAdd(&ctr->store_tail, 1ull);
      seen[i] = r == 0;
    } else {
      seen[i] = (uint8_t)set_contains(table, bucket_mask, id);
    }
  }
}

// ----------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------
struct LaunchRec {
  int kind;  // 0 expand, 1 insert, 2 other
  cudaEvent_t a, b;
};

struct Engine {
  int device = 0;
  int sms = 148;
  uint32_t rank = 0, world = 1;
  int table_log2 = 0;
  uint64_t max_states = 0;
  uint64_t cand_bytes = 0;
  bool cont = false;
  bool check_deadlock = M::CHECK_DEADLOCK;
  bool timing = true;
  bool count_actions = false;
  uint64_t stop_after_states = 0;   // bounded run: stop at the first level end with >= this many states
  int l2_fetch = 0;                 // cudaLimitMaxL2FetchGranularity hint (32/64/128), 0 = leave the default
  uint32_t fanout_bound = 0;        // successors per state assumed when sizing a frontier chunk (0: min(MAX_FANOUT, 32))
  // spill / checkpoint (single rank): the device store is a ring over the live window [store_base, tail); the
  // levels below the one being expanded move to host memory (TLC's DiskStateQueue / trace file on disk)
  bool spill = false;
  uint64_t store_base = 0;
  std::vector<uint64_t> host_store, host_parent;
  std::string checkpoint_dir, recover_dir;
  double checkpoint_minutes = 0;    // 0: a checkpoint after every level (when checkpoint_dir is set)
  std::chrono::steady_clock::time_point last_checkpoint;
  bool overlap = false;             // single rank: K2 of chunk i runs on a second stream while K1 expands chunk i+1
                                    // (K1 is issue-bound, K2 waits on random DRAM sectors: they use different units)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_exp[2] = {nullptr, nullptr}, ev_ins[2] = {nullptr, nullptr};
  bool prefetch = false;            // K1 prefetches candidate buckets into L2 (pair with a small chunk_states)
  uint64_t chunk_states_opt = 0;    // frontier states per K1/K2 launch pair (0: as many as the candidate buffer allows)
  bool one_phase = false;           // comparison only: the round-1 one-phase K1 (needs a -DKMC_ONE_PHASE build)

  void* table = nullptr;
  uint64_t table_slots = 0;             // slots of SLOT_BYTES each
  uint64_t* store = nullptr;
  uint64_t* parent = nullptr;
  uint64_t* cand = nullptr;
  uint64_t region_rows = 0;
  uint64_t* recv = nullptr;
  uint64_t recv_rows = 0;
  uint64_t* inbox_alloc = nullptr;    // sync page + 2 inbox buffers, shared through CUDA IPC
  uint64_t* board_host = nullptr;     // pinned: the level board as gathered by k_gather_level
  uint64_t round = 0, level_id = 0;   // monotonic over the life of the context (see SYNC_WORDS)
  uint64_t prev_tail = 0;
  uint64_t* inbox = nullptr;          // fused exchange: 2 x (header + world regions)
  uint64_t inbox_stride = 0;
  uint64_t* peer_inbox[MAX_WORLD] = {};
  bool peers_open = false;
  bool peers_direct = false;          // peer pointers given directly (same process, cudaDeviceEnablePeerAccess)
  uint32_t inbox_buf = 0;
  uint64_t exchanged_rows = 0;
  DevCounters* ctr = nullptr;
  uint64_t* viol_ring = nullptr;
  cudaStream_t stream = nullptr;
  bool own_stream = true;       // false: the caller's stream (option "stream"), e.g. torch's current stream
  uint64_t chunk_states = 0;

  std::vector<cudaEvent_t> event_pool;
  size_t events_used = 0;
  std::vector<LaunchRec> launches;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;

  // results
  mutable std::mutex mu;
  kmc_stats_t stats{};
  std::vector<uint64_t> widths;
  std::vector<uint64_t> action_counts;
  kmc_violation_t viol{};
  std::vector<std::vector<uint64_t>> trace;
  std::vector<uint32_t> trace_actions;
  std::vector<uint64_t> viol_words;     // the offending state itself and its parent/action word
  uint64_t viol_meta = 0;
  bool ran = false;
  uint64_t level_first = 0, level_count = 0;   // shard API
  uint64_t shard_levels = 0;
  std::string last_error;

  Params params() const {
    Params p;
    p.table = table;
    p.bucket_mask = table_slots / BUCKET_SLOTS - 1;
    p.store = store;
    p.parent = parent;
    p.max_states = max_states;
    p.store_mask = spill ? max_states - 1 : ~0ull;
    p.store_base = store_base;
    p.cand = cand;
    p.region_rows = region_rows;
    p.ctr = ctr;
    p.viol_ring = viol_ring;
    p.rank = rank;
    p.world = world;
    p.check_deadlock = check_deadlock ? 1 : 0;
    p.count_actions = count_actions ? 1 : 0;
    p.prefetch = prefetch ? 1 : 0;
    p.cand_slot = 0;
    for (int r = 0; r < MAX_WORLD; ++r) p.peer_inbox[r] = peer_inbox[r];
    p.inbox_stride = inbox_stride;
    p.p2p = 0;
    p.inbox_buf = inbox_buf;
    return p;
  }
};

struct kmcm_ctx {
  Engine e;
  // option "gpus": N > 1 -- this context drives N GPUs of the process: one sub-context (rank) per device, peers
  // mapped directly (cudaDeviceEnablePeerAccess), one host thread per rank inside kmcm_run.  `e` then only
  // holds the aggregated results.
  std::vector<kmcm_ctx*> ranks;
};

#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess) {                                                                             \
      E.last_error = std::string(#call) + ": " + cudaGetErrorString(_e);                                 \
      return (_e == cudaErrorMemoryAllocation) ? KMC_E_OOM : KMC_E_CUDA;                                 \
    }                                                                                                    \
  } while (0)

static bool json_find(const char* js, const char* key, const char** val) {
  if (!js) return false;
  std::string pat = std::string("\"") + key + "\"";
  const char* p = strstr(js, pat.c_str());
  if (!p) return false;
  p += pat.size();
  while (*p == ' ' || *p == '\t' || *p == '\n') ++p;
  if (*p != ':') return false;
  ++p;
  while (*p == ' ' || *p == '\t' || *p == '\n') ++p;
  *val = p;
  return true;
}
static bool json_num(const char* js, const char* key, double* out) {
  const char* v;
  if (!json_find(js, key, &v)) return false;
  char* end;
  double d = strtod(v, &end);
  if (end == v) return false;
  *out = d;
  return true;
}
static bool json_str(const char* js, const char* key, std::string* out) {
  const char* v;
  if (!json_find(js, key, &v) || *v != '"') return false;
  const char* e = strchr(v + 1, '"');
  if (!e) return false;
  out->assign(v + 1, e);
  return true;
}
static bool json_bool(const char* js, const char* key, bool* out) {
  const char* v;
  if (!json_find(js, key, &v)) return false;
  if (!strncmp(v, "true", 4)) { *out = true; return true; }
  if (!strncmp(v, "false", 5)) { *out = false; return true; }
  double d;
  if (json_num(js, key, &d)) { *out = d != 0; return true; }
  return false;
}

static cudaEvent_t get_event(Engine& E) {
  if (E.events_used == E.event_pool.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    E.event_pool.push_back(ev);
  }
  return E.event_pool[E.events_used++];
}

struct TimedLaunch {
  Engine& E;
  int kind;
  cudaStream_t on;
  cudaEvent_t a = nullptr, b = nullptr;
  TimedLaunch(Engine& e, int k, cudaStream_t s = nullptr) : E(e), kind(k), on(s ? s : e.stream) {
    if (E.timing) {
      a = get_event(E);
      b = get_event(E);
      cudaEventRecord(a, on);
    }
  }
  ~TimedLaunch() {
    if (E.timing) {
      cudaEventRecord(b, on);
      E.launches.push_back({kind, a, b});
    }
  }
};

static int grid_for(const Engine& E, uint64_t n, int block, int per_sm) {
  uint64_t g = (n + block - 1) / block;
  uint64_t cap = (uint64_t)E.sms * per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

static int engine_alloc(Engine& E) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    E.last_error = "no CUDA device visible; this library has no CPU fallback";
    return KMC_E_NO_GPU;
  }
  if (E.device >= ndev) {
    E.last_error = "device index out of range";
    return KMC_E_BADARG;
  }
  CK(cudaSetDevice(E.device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, E.device));
  E.sms = prop.multiProcessorCount;
  if (E.l2_fetch) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)E.l2_fetch));
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  // Default sizing from the memory that is actually free: candidate buffers first, then set + store + parent
  // links share the rest (2.5 slots per state: load <= 0.4, rounded to a power of two).
  if (E.cand_bytes == 0) E.cand_bytes = std::min<uint64_t>(free_b / 8, (E.world > 1 ? 24ull : 12ull) << 30);
  const uint64_t cand_total = E.cand_bytes * (E.world > 1 ? 4 : 1);          // + recv + two inbox buffers
  const uint64_t budget = free_b > cand_total + (512ull << 20) ? (uint64_t)((free_b - cand_total) * 0.94) : 0;
  const uint64_t per_state = (uint64_t)W * 8 + 8;
  if (E.table_log2 == 0 && E.max_states == 0) {
    int lg = 34;
    while (lg > 16 && ((uint64_t)SLOT_BYTES << lg) + (uint64_t)((1ull << lg) / 2.5) * per_state > budget) --lg;
    E.table_log2 = lg;
    E.max_states = (uint64_t)((1ull << lg) / 2.5);
  } else if (E.table_log2 == 0) {
    int lg = 16;
    while (lg < 34 && (1ull << lg) < (uint64_t)(2.5 * (double)E.max_states)) ++lg;
    E.table_log2 = lg;
  }
  if (E.spill && E.max_states) {
    uint64_t p2 = 1;
    while (p2 * 2 <= E.max_states) p2 *= 2;
    E.max_states = p2;                                  // the spilling store is a power-of-two ring
  }
  E.table_slots = 1ull << E.table_log2;
  if (E.max_states == 0) {
    const uint64_t table_bytes = E.table_slots * SLOT_BYTES;
    const uint64_t room = budget > table_bytes ? (budget - table_bytes) / per_state : 0;
    E.max_states = std::max<uint64_t>(1024, std::min<uint64_t>(E.table_slots / 2, room));
    if (E.spill) {
      uint64_t p2 = 1;
      while (p2 * 2 <= E.max_states) p2 *= 2;
      E.max_states = p2;
    }
  }
  uint64_t rows_total = E.cand_bytes / (ROW * 8);
  if (E.world > 1) E.overlap = false;                    // (the fused exchange double-buffers on its own)
  E.region_rows = rows_total / (E.overlap ? 2 : E.world);
  if (E.region_rows < (uint64_t)M::MAX_FANOUT) E.region_rows = M::MAX_FANOUT;
  // A chunk of frontier states is sized for `fanout_bound` successors per state on average *per owner region*.
  // MAX_FANOUT (emit sites in expand) is a safe but very loose bound -- reachable states enable a small
  // fraction of the sites (max 15 successors seen on the Kafka models, MAX_FANOUT ~100); the default
  // assumes <= 32 and relies on the kernel's overflow check (KMC_E_CAND_FULL, nothing is lost silently).
  if (E.fanout_bound == 0) E.fanout_bound = std::min<uint32_t>((uint32_t)M::MAX_FANOUT, 32u);
  E.chunk_states = std::max<uint64_t>(1, E.region_rows / E.fanout_bound);
  if (E.chunk_states_opt) E.chunk_states = std::min<uint64_t>(E.chunk_states, E.chunk_states_opt);
  if (E.own_stream) CK(cudaStreamCreateWithFlags(&E.stream, cudaStreamNonBlocking));
  if (E.overlap) {
    CK(cudaStreamCreateWithFlags(&E.stream2, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CK(cudaEventCreateWithFlags(&E.ev_exp[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&E.ev_ins[i], cudaEventDisableTiming));
    }
  }
  CK(cudaMalloc(&E.table, E.table_slots * SLOT_BYTES));
  CK(cudaMalloc(&E.store, E.max_states * W * 8));
  CK(cudaMalloc(&E.parent, E.max_states * 8));
  CK(cudaMalloc(&E.cand, E.region_rows * (E.overlap ? 2 : E.world) * ROW * 8));
  if (E.world > 1) {
    E.recv_rows = E.region_rows * E.world;
    CK(cudaMalloc(&E.recv, E.recv_rows * ROW * 8));
    E.inbox_stride = INBOX_HEADER + E.region_rows * E.world * ROW;
    CK(cudaMalloc(&E.inbox_alloc, (SYNC_WORDS + 2 * E.inbox_stride) * 8));
    CK(cudaMemset(E.inbox_alloc, 0, (SYNC_WORDS + 2 * E.inbox_stride) * 8));
    E.inbox = E.inbox_alloc + SYNC_WORDS;
    CK(cudaHostAlloc(&E.board_host, MAX_WORLD * BOARD_WORDS * 8, cudaHostAllocMapped));
    memset(E.board_host, 0, MAX_WORLD * BOARD_WORDS * 8);
  }
  // the expand kernel keeps its state tile, the successor stage and the pair list in > 48 KB of dynamic shared memory
  CK(cudaFuncSetAttribute(k_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXPAND_SMEM_BYTES));
#ifdef KMC_ONE_PHASE
  CK(cudaFuncSetAttribute(k_expand1, cudaFuncAttributeMaxDynamicSharedMemorySize, STAGE_BYTES + NWARPS * 4));
#else
  if (E.one_phase) {
    E.last_error = "option one_phase needs a library built with -DKMC_ONE_PHASE";
    return KMC_E_BADARG;
  }
#endif
  CK(cudaMalloc(&E.ctr, sizeof(DevCounters)));
  CK(cudaMalloc(&E.viol_ring, (size_t)VIOL_RING * VIOL_ROW * 8));
  CK(cudaEventCreate(&E.ev_begin));
  CK(cudaEventCreate(&E.ev_end));
  return KMC_OK;
}

static int engine_reset(Engine& E) {
  CK(cudaSetDevice(E.device));
  CK(cudaMemsetAsync(E.table, KEY128 ? 0xFF : 0, E.table_slots * SLOT_BYTES, E.stream));      // empty marker: all-ones keys / zero fingerprints
  DevCounters h;
  memset(&h, 0, sizeof(h));
  CK(cudaMemcpyAsync(E.ctr, &h, sizeof(h), cudaMemcpyHostToDevice, E.stream));
  CK(cudaStreamSynchronize(E.stream));
  E.events_used = 0;
  E.launches.clear();
  E.widths.clear();
  E.trace.clear();
  E.trace_actions.clear();
  memset(&E.viol, 0, sizeof(E.viol));
  E.viol.invariant = -1;
  E.level_first = E.level_count = 0;
  E.shard_levels = 0;
  E.store_base = 0;
  E.host_store.clear();
  E.host_parent.clear();
  return KMC_OK;
}

static int read_counters(Engine& E, DevCounters* h) {
  CK(cudaMemcpyAsync(h, E.ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, E.stream));
  CK(cudaStreamSynchronize(E.stream));
  return KMC_OK;
}

static int fail_to_error(unsigned long long f) {
  switch (f) {
    case 0: return KMC_OK;
    case KMC_FAIL_LAYOUT: return KMC_E_LAYOUT_OVERFLOW;
    case KMC_FAIL_TABLE_FULL: return KMC_E_TABLE_FULL;
    case KMC_FAIL_STORE_FULL: return KMC_E_STORE_FULL;
    case KMC_FAIL_CAND_FULL: return KMC_E_CAND_FULL;
    case KMC_FAIL_PEER_TIMEOUT: return KMC_E_PEER_TIMEOUT;
    default: return KMC_E_CUDA;
  }
}

// writes the init states as candidate rows (one region per owner) -- host side, tiny
static int seed_init(Engine& E) {
  std::vector<uint64_t> rows[MAX_WORLD];
  unsigned long long counts[MAX_WORLD] = {0};
  for (int i = 0; i < M::NUM_INIT; ++i) {
    State s;
    memcpy(s.w, M::INIT_STATES[i], sizeof(s.w));
    uint32_t d = E.world > 1 ? owner_of(state_fp(s), E.world) : 0;
    // every rank seeds the same init states but only rank 0 contributes them, so that the
    // generated count and the exchange see each init state exactly once
    if (E.rank != 0) continue;
    for (int k = 0; k < W; ++k) rows[d].push_back(s.w[k]);
    rows[d].push_back(NO_PARENT);
    counts[d]++;
  }
  for (uint32_t d = 0; d < E.world; ++d) {
    if (counts[d] == 0) continue;
    if (counts[d] > E.region_rows) return KMC_E_OOM;
    CK(cudaMemcpyAsync(E.cand + (uint64_t)d * E.region_rows * ROW, rows[d].data(), rows[d].size() * 8,
                       cudaMemcpyHostToDevice, E.stream));
  }
  CK(cudaMemcpyAsync(E.ctr, counts, sizeof(counts), cudaMemcpyHostToDevice, E.stream));
  unsigned long long gen = E.rank == 0 ? M::NUM_INIT : 0;
  CK(cudaMemcpyAsync(&E.ctr->generated, &gen, sizeof(gen), cudaMemcpyHostToDevice, E.stream));
  CK(cudaStreamSynchronize(E.stream));
  return KMC_OK;
}

static int launch_insert(Engine& E, const uint64_t* rows, const unsigned long long* n_dev, uint64_t n_fixed,
                         uint64_t n_bound, cudaStream_t on = nullptr) {
  Params p = E.params();
  {
    TimedLaunch t(E, 1, on);
    k_insert<<<grid_for(E, n_bound, 256, 8), 256, 0, on ? on : E.stream>>>(p, rows, n_dev, n_fixed);
  }
  CK(cudaGetLastError());
  return KMC_OK;
}

static int launch_invariants(Engine& E, uint64_t first, uint64_t count_bound) {
  if (M::NUM_INVARIANTS == 0) return KMC_OK;
  Params p = E.params();
  TimedLaunch t(E, 2);
  k_invariants<<<grid_for(E, std::max<uint64_t>(count_bound, 1), 256, 8), 256, 0, E.stream>>>(p, first, &E.ctr->store_tail);
  CK(cudaGetLastError());
  return KMC_OK;
}

static int launch_expand(Engine& E, uint64_t first, uint64_t count, bool p2p = false, uint32_t slot = 0) {
  Params p = E.params();
  p.p2p = p2p ? 1 : 0;
  p.cand_slot = slot;
  p.cand = E.cand + (uint64_t)slot * E.region_rows * ROW;
  if (count == 0) return KMC_OK;
  TimedLaunch t(E, 0);
#ifdef KMC_ONE_PHASE
  if (E.one_phase) {
    int spt = EXPAND_SPT1;
    while (spt > 1 && count < (uint64_t)E.sms * EXPAND_BLOCK * spt) spt >>= 1;
    uint64_t tiles = (count + (uint64_t)EXPAND_BLOCK * spt - 1) / ((uint64_t)EXPAND_BLOCK * spt);
    int grid = (int)std::min<uint64_t>(std::max<uint64_t>(tiles, 1), (uint64_t)E.sms);
    k_expand1<<<grid, EXPAND_BLOCK, STAGE_BYTES + NWARPS * 4, E.stream>>>(p, first, count, spt);
    CK(cudaGetLastError());
    return KMC_OK;
  }
#endif
  // small levels: smaller tiles so that every SM still gets one (a tile is a multiple of 32 states)
  const uint64_t ctas = (uint64_t)E.sms * EXPAND_CTAS_PER_SM;
  uint64_t per_cta = (count + ctas - 1) / ctas;
  unsigned tile_states = (unsigned)std::min<uint64_t>((uint64_t)TILE, std::max<uint64_t>(32, (per_cta + 31) & ~31ull));
  uint64_t tiles = (count + tile_states - 1) / tile_states;
  int grid = (int)std::min<uint64_t>(tiles, ctas);
  k_expand<<<grid, EXPAND_BLOCK, EXPAND_SMEM_BYTES, E.stream>>>(p, first, count, tile_states);
  CK(cudaGetLastError());
  return KMC_OK;
}

static int reset_cand(Engine& E) {
  CK(cudaMemsetAsync(E.ctr->cand_count, 0, sizeof(unsigned long long) * MAX_WORLD, E.stream));
  return KMC_OK;
}

// State / parent word by GLOBAL index: from the host spill below store_base, else from the device ring.
static int fetch_state(Engine& E, uint64_t idx, uint64_t* words, uint64_t* meta) {
  if (idx < E.store_base) {
    memcpy(words, E.host_store.data() + idx * W, W * 8);
    *meta = E.host_parent[idx];
    return KMC_OK;
  }
  const uint64_t slot = E.spill ? (idx & (E.max_states - 1)) : idx;
  CK(cudaMemcpy(words, E.store + slot * W, W * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(meta, E.parent + slot, 8, cudaMemcpyDeviceToHost));
  return KMC_OK;
}

// device ring range [first, first + count) (global indices, all on the device) -> host buffers
static int copy_ring_range(Engine& E, uint64_t first, uint64_t count, uint64_t* states, uint64_t* parents) {
  uint64_t done = 0;
  while (done < count) {
    const uint64_t g = first + done;
    const uint64_t slot = E.spill ? (g & (E.max_states - 1)) : g;
    const uint64_t n = E.spill ? std::min<uint64_t>(count - done, E.max_states - slot) : count - done;
    if (states) CK(cudaMemcpy(states + done * W, E.store + slot * W, n * W * 8, cudaMemcpyDeviceToHost));
    if (parents) CK(cudaMemcpy(parents + done, E.parent + slot, n * 8, cudaMemcpyDeviceToHost));
    done += n;
  }
  return KMC_OK;
}

// Spill: everything below the level that is expanded next moves to host memory and its ring slots become free.
static int spill_below(Engine& E, uint64_t level_first) {
  if (!E.spill || level_first <= E.store_base) return KMC_OK;
  const uint64_t n = level_first - E.store_base;
  E.host_store.resize(level_first * W);
  E.host_parent.resize(level_first);
  int rc = copy_ring_range(E, E.store_base, n, E.host_store.data() + E.store_base * W, E.host_parent.data() + E.store_base);
  if (rc) return rc;
  E.store_base = level_first;
  return KMC_OK;
}

// ---- checkpoint / recover (TLC -checkpoint / -recover): written at a level boundary -------------------------------
// <dir>/checkpoint.meta  text: key value per line;  <dir>/checkpoint.bin  states [0, tail) then parent words [0, tail).
// The fingerprint set is not stored: it is rebuilt from the states on recover (and may then have another size).
struct LevelCursor {
  uint64_t level_first, level_end, level;
};
static int write_checkpoint(Engine& E, const DevCounters& h, const LevelCursor& lc) {
  const std::string meta = E.checkpoint_dir + "/checkpoint.meta", bin = E.checkpoint_dir + "/checkpoint.bin";
  const std::string tmp = bin + ".tmp";
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) {
    E.last_error = "cannot write " + tmp;
    return KMC_E_BADARG;
  }
  const uint64_t tail = h.store_tail;
  std::vector<uint64_t> st((size_t)(tail - E.store_base) * W), pa((size_t)(tail - E.store_base));
  int rc = copy_ring_range(E, E.store_base, tail - E.store_base, st.data(), pa.data());
  if (rc) { fclose(f); return rc; }
  bool ok = fwrite(E.host_store.data(), 8, (size_t)E.store_base * W, f) == (size_t)E.store_base * W &&
            fwrite(st.data(), 8, st.size(), f) == st.size() &&
            fwrite(E.host_parent.data(), 8, (size_t)E.store_base, f) == (size_t)E.store_base &&
            fwrite(pa.data(), 8, pa.size(), f) == pa.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok || rename(tmp.c_str(), bin.c_str()) != 0) {
    E.last_error = "short write on " + tmp;
    return KMC_E_BADARG;
  }
  f = fopen((meta + ".tmp").c_str(), "w");
  if (!f) return KMC_E_BADARG;
  fprintf(f, "model %s\ndigest %s\nwords %d\ntail %llu\nlevel_first %llu\nlevel_end %llu\nlevel %llu\ngenerated %llu\n"
             "deadlocks %llu\nout_of_model %llu\nprobes %llu\nwidths",
          KMC_MODEL_NAME, KMC_MODEL_DIGEST, W, (unsigned long long)tail, (unsigned long long)lc.level_first,
          (unsigned long long)lc.level_end, (unsigned long long)lc.level, (unsigned long long)h.generated,
          (unsigned long long)h.deadlocks, (unsigned long long)h.out_of_model, (unsigned long long)h.probes);
  for (uint64_t w : E.widths) fprintf(f, " %llu", (unsigned long long)w);
  fprintf(f, "\n");
  fclose(f);
  if (rename((meta + ".tmp").c_str(), meta.c_str()) != 0) return KMC_E_BADARG;
  E.last_checkpoint = std::chrono::steady_clock::now();
  return KMC_OK;
}

static int read_checkpoint(Engine& E, DevCounters* h, LevelCursor* lc) {
  const std::string meta = E.recover_dir + "/checkpoint.meta", bin = E.recover_dir + "/checkpoint.bin";
  FILE* f = fopen(meta.c_str(), "r");
  if (!f) {
    E.last_error = "cannot read " + meta;
    return KMC_E_BADARG;
  }
  char key[64], val[256];
  unsigned long long tail = 0, words = 0;
  std::string digest;
  memset(h, 0, sizeof(*h));
  E.widths.clear();
  while (fscanf(f, "%63s", key) == 1) {
    if (!strcmp(key, "widths")) {
      unsigned long long w;
      while (fscanf(f, "%llu", &w) == 1) E.widths.push_back(w);
      break;
    }
    if (fscanf(f, "%255s", val) != 1) break;
    const unsigned long long v = strtoull(val, nullptr, 10);
    if (!strcmp(key, "digest")) digest = val;
    else if (!strcmp(key, "words")) words = v;
    else if (!strcmp(key, "tail")) tail = v;
    else if (!strcmp(key, "level_first")) lc->level_first = v;
    else if (!strcmp(key, "level_end")) lc->level_end = v;
    else if (!strcmp(key, "level")) lc->level = v;
    else if (!strcmp(key, "generated")) h->generated = v;
    else if (!strcmp(key, "deadlocks")) h->deadlocks = v;
    else if (!strcmp(key, "out_of_model")) h->out_of_model = v;
    else if (!strcmp(key, "probes")) h->probes = v;
  }
  fclose(f);
  if (digest != KMC_MODEL_DIGEST || words != (unsigned long long)W) {
    E.last_error = "checkpoint belongs to another model (digest mismatch)";
    return KMC_E_MODEL;
  }
  h->store_tail = tail;
  // states below the level to expand go to the host (spill) or, without spill, everything to the device
  const uint64_t keep_from = E.spill ? lc->level_first : 0;
  if (tail - keep_from > E.max_states) {
    E.last_error = "checkpoint does not fit the state store (raise max_states or use spill)";
    return KMC_E_STORE_FULL;
  }
  f = fopen(bin.c_str(), "rb");
  if (!f) {
    E.last_error = "cannot read " + bin;
    return KMC_E_BADARG;
  }
  std::vector<uint64_t> st((size_t)tail * W), pa((size_t)tail);
  bool ok = fread(st.data(), 8, st.size(), f) == st.size() && fread(pa.data(), 8, pa.size(), f) == pa.size();
  fclose(f);
  if (!ok) {
    E.last_error = "short read on " + bin;
    return KMC_E_BADARG;
  }
  E.store_base = keep_from;
  E.host_store.assign(st.begin(), st.begin() + (size_t)keep_from * W);
  E.host_parent.assign(pa.begin(), pa.begin() + (size_t)keep_from);
  // device window
  for (uint64_t g = keep_from; g < tail;) {
    const uint64_t slot = E.spill ? (g & (E.max_states - 1)) : g;
    const uint64_t n = E.spill ? std::min<uint64_t>(tail - g, E.max_states - slot) : tail - g;
    CK(cudaMemcpyAsync(E.store + slot * W, st.data() + g * W, n * W * 8, cudaMemcpyHostToDevice, E.stream));
    CK(cudaMemcpyAsync(E.parent + slot, pa.data() + g, n * 8, cudaMemcpyHostToDevice, E.stream));
    g += n;
  }
  // rebuild the set: every stored state is inserted once, in batches through the candidate buffer
  Params p = E.params();
  const uint64_t batch = E.region_rows * ROW / W;
  for (uint64_t g = 0; g < tail; g += batch) {
    const uint64_t n = std::min<uint64_t>(batch, tail - g);
    CK(cudaMemcpyAsync(E.cand, st.data() + g * W, n * W * 8, cudaMemcpyHostToDevice, E.stream));
    k_rebuild<<<grid_for(E, n, 256, 8), 256, 0, E.stream>>>(p, E.cand, n);
    CK(cudaStreamSynchronize(E.stream));
  }
  CK(cudaMemcpyAsync(&E.ctr->store_tail, &h->store_tail, 8, cudaMemcpyHostToDevice, E.stream));
  CK(cudaMemcpyAsync(&E.ctr->generated, &h->generated, 8, cudaMemcpyHostToDevice, E.stream));
  CK(cudaMemcpyAsync(&E.ctr->deadlocks, &h->deadlocks, 8, cudaMemcpyHostToDevice, E.stream));
  CK(cudaMemcpyAsync(&E.ctr->out_of_model, &h->out_of_model, 8, cudaMemcpyHostToDevice, E.stream));
  CK(cudaMemcpyAsync(&E.ctr->probes, &h->probes, 8, cudaMemcpyHostToDevice, E.stream));
  CK(cudaStreamSynchronize(E.stream));
  DevCounters now;
  int rc = read_counters(E, &now);
  if (rc) return rc;
  return fail_to_error(now.fail);
}

// Picks the violator with the smallest fingerprint (deadlocks, which belong to the level being
// expanded, before invariant violations of the next level) and walks its parent links back to an
// initial state.  `level` is the level being expanded (0 for the init insert).
static int build_trace(Engine& E, const DevCounters& h, uint64_t level) {
  uint64_t n = std::min<uint64_t>(h.viol_count, VIOL_RING);
  if (n == 0) return KMC_OK;
  std::vector<uint64_t> ring(n * VIOL_ROW);
  CK(cudaMemcpy(ring.data(), E.viol_ring, ring.size() * 8, cudaMemcpyDeviceToHost));
  const uint64_t* best = nullptr;
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t* r = ring.data() + i * VIOL_ROW;
    if (!best) { best = r; continue; }
    bool r_dead = r[W + 2] == ~0ull, b_dead = best[W + 2] == ~0ull;
    if (r_dead != b_dead) { if (r_dead) best = r; continue; }
    if (r[W + 1] < best[W + 1] || (r[W + 1] == best[W + 1] && r[W] < best[W])) best = r;
  }
  std::vector<std::vector<uint64_t>> rev;
  std::vector<uint32_t> rev_act;
  uint64_t meta = best[W];
  E.viol_words.assign(best, best + W);
  E.viol_meta = meta;
  rev.push_back(std::vector<uint64_t>(best, best + W));
  rev_act.push_back((uint32_t)(meta >> 56));
  uint64_t guard = 0;
  while ((meta & 0x0000FFFFFFFFFFFFull) != NO_PARENT && guard++ < 100000) {
    uint64_t idx = meta & IDX_MASK;
    uint32_t prank = (uint32_t)((meta >> 40) & 0xFF);
    if (prank != E.rank || idx - E.store_base >= E.max_states && idx >= E.store_base) break;  // parent lives on another rank
    std::vector<uint64_t> st(W);
    {
      int frc = fetch_state(E, idx, st.data(), &meta);
      if (frc) return frc;
    }
    rev.push_back(st);
    rev_act.push_back((uint32_t)(meta >> 56));
  }
  E.trace.assign(rev.rbegin(), rev.rend());
  E.trace_actions.assign(rev_act.rbegin(), rev_act.rend());
  bool dead = best[W + 2] == ~0ull;
  E.viol.kind = dead ? KMC_RESULT_DEADLOCK : KMC_RESULT_INVARIANT;
  E.viol.invariant = dead ? -1 : (int32_t)best[W + 2];
  E.viol.level = dead ? level : level + 1;
  E.viol.trace_len = E.trace.size();
  E.viol.fingerprint = best[W + 1];
  return KMC_OK;
}

static void accumulate_timing(Engine& E, kmc_stats_t& st) {
  st.gpu_ms_expand = st.gpu_ms_insert = st.gpu_ms_invariant = 0;
  st.launches_expand = st.launches_insert = st.launches_other = 0;
  for (const LaunchRec& r : E.launches) {
    float ms = 0;
    cudaEventElapsedTime(&ms, r.a, r.b);
    if (r.kind == 0) { st.gpu_ms_expand += ms; st.launches_expand++; }
    else if (r.kind == 1) { st.gpu_ms_insert += ms; st.launches_insert++; }
    else { st.gpu_ms_invariant += ms; st.launches_other++; }
  }
}

static int engine_run(Engine& E) {
  auto t0 = std::chrono::steady_clock::now();
  int rc = engine_reset(E);
  if (rc) return rc;
  if (E.world != 1) {
    E.last_error = "kmc_run drives one rank; use the kmc_shard_* calls for world > 1";
    return KMC_E_BADARG;
  }
  CK(cudaEventRecord(E.ev_begin, E.stream));
  DevCounters h;
  uint64_t level_first = 0, level_end = 0, level = 1;
  bool stopped = false;
  int err = 0;
  E.last_checkpoint = std::chrono::steady_clock::now();
  if (!E.recover_dir.empty()) {
    // -recover: continue from the level boundary a checkpoint was written at
    LevelCursor lc{0, 0, 1};
    if ((rc = read_checkpoint(E, &h, &lc))) return rc;
    level_first = lc.level_first;
    level_end = lc.level_end;
    level = lc.level;
  } else {
    rc = seed_init(E);
    if (rc) return rc;
    rc = launch_insert(E, E.cand, &E.ctr->cand_count[0], 0, M::NUM_INIT);
    if (rc) return rc;
    if ((rc = launch_invariants(E, 0, M::NUM_INIT))) return rc;
    rc = read_counters(E, &h);
    if (rc) return rc;
    level_end = h.store_tail;
    err = fail_to_error(h.fail);
    if (!err && h.viol_count) {
      build_trace(E, h, 0);
      if (!E.cont) stopped = true;
    }
  }
  while (!err && !stopped && level_end > level_first) {
    E.widths.push_back(level_end - level_first);
    if ((rc = spill_below(E, level_first))) return rc;         // (no-op unless spilling)
    // a chunk never crosses the wrap of the ring store
    auto chunk_len = [&](uint64_t off) {
      uint64_t cnt = std::min<uint64_t>(E.chunk_states, level_end - off);
      if (E.spill) cnt = std::min<uint64_t>(cnt, E.max_states - (off & (E.max_states - 1)));
      return cnt;
    };
    if (E.overlap) {
      // K1 of chunk i+1 (stream) overlaps K2 of chunk i (stream2); the two halves of the candidate buffer alternate
      uint32_t slot = 0;
      for (uint64_t off = level_first, cnt; off < level_end; off += cnt, slot ^= 1) {
        cnt = chunk_len(off);
        CK(cudaStreamWaitEvent(E.stream, E.ev_ins[slot], 0));            // the K2 that read this half has finished
        CK(cudaMemsetAsync(&E.ctr->cand_count[slot], 0, sizeof(unsigned long long), E.stream));
        if ((rc = launch_expand(E, off, cnt, false, slot))) return rc;
        CK(cudaEventRecord(E.ev_exp[slot], E.stream));
        CK(cudaStreamWaitEvent(E.stream2, E.ev_exp[slot], 0));
        if ((rc = launch_insert(E, E.cand + (uint64_t)slot * E.region_rows * ROW, &E.ctr->cand_count[slot], 0,
                                cnt * (uint64_t)E.fanout_bound, E.stream2))) return rc;
        CK(cudaEventRecord(E.ev_ins[slot], E.stream2));
      }
      CK(cudaStreamWaitEvent(E.stream, E.ev_ins[0], 0));
      CK(cudaStreamWaitEvent(E.stream, E.ev_ins[1], 0));
    } else {
      for (uint64_t off = level_first, cnt; off < level_end; off += cnt) {
        cnt = chunk_len(off);
        if ((rc = reset_cand(E))) return rc;
        if ((rc = launch_expand(E, off, cnt))) return rc;
        if ((rc = launch_insert(E, E.cand, &E.ctr->cand_count[0], 0, cnt * (uint64_t)E.fanout_bound))) return rc;
      }
    }
    if ((rc = launch_invariants(E, level_end, (level_end - level_first) * 2))) return rc;
    if ((rc = read_counters(E, &h))) return rc;
    err = fail_to_error(h.fail);
    {
      std::lock_guard<std::mutex> g(E.mu);
      E.stats.distinct = E.spill ? h.store_tail : std::min<uint64_t>(h.store_tail, E.max_states);
      E.stats.generated = h.generated;
      E.stats.depth = level;
      E.stats.queue = h.store_tail - level_end;
    }
    if (!err && h.viol_count && E.viol.kind == KMC_RESULT_OK) {
      build_trace(E, h, level);
      if (!E.cont) {
        stopped = true;
        level_first = level_end;
        level_end = h.store_tail;
        break;
      }
    }
    level_first = level_end;
    level_end = h.store_tail;
    ++level;
    if (!err && !E.checkpoint_dir.empty() && level_end > level_first) {
      const double mins = std::chrono::duration<double>(std::chrono::steady_clock::now() - E.last_checkpoint).count() / 60.0;
      if (mins >= E.checkpoint_minutes) {
        if ((rc = spill_below(E, level_first))) return rc;
        if ((rc = write_checkpoint(E, h, LevelCursor{level_first, level_end, level}))) return rc;
      }
    }
    if (E.stop_after_states && h.store_tail >= E.stop_after_states && level_end > level_first) {
      stopped = true;          // bounded throughput run: the queue is reported, no error
      break;
    }
  }
  CK(cudaEventRecord(E.ev_end, E.stream));
  CK(cudaStreamSynchronize(E.stream));
  float total_ms = 0;
  cudaEventElapsedTime(&total_ms, E.ev_begin, E.ev_end);
  auto t1 = std::chrono::steady_clock::now();
  {
    std::lock_guard<std::mutex> g(E.mu);
    kmc_stats_t& st = E.stats;
    st.distinct = E.spill ? h.store_tail : std::min<uint64_t>(h.store_tail, E.max_states);
    st.generated = h.generated;
    st.queue = stopped ? (level_end - level_first) : 0;
    st.depth = E.widths.size();
    st.deadlocks = h.deadlocks;
    st.out_of_model = h.out_of_model;
    st.probes = h.probes;
    st.levels = E.widths.size();
    st.gpu_ms_total = total_ms;
    st.wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    st.table_slots = E.table_slots;
    st.slot_bytes = SLOT_BYTES;
    st.max_states = E.max_states;
    st.complete = (!err && !stopped) ? 1 : 0;
    if (E.timing) accumulate_timing(E, st);
    E.action_counts.assign(h.action_counts, h.action_counts + 64);
    E.ran = true;
  }
  return err;
}

// ----------------------------------------------------------------------------------------
// exported per-model ABI (the dispatcher libkspecmc.so forwards kmc_* to these)
// ----------------------------------------------------------------------------------------
static int multi_create(kmcm_ctx* c, const char* options_json, int gpus);
static int multi_run(kmcm_ctx* c);

#define E (c->e)
extern "C" {

int kmcm_create(const char* options_json, kmcm_ctx** out) {
  if (!out) return KMC_E_BADARG;
  kmcm_ctx* c = new kmcm_ctx();
  double d;
  bool b;
  if (json_num(options_json, "device", &d)) E.device = (int)d;
  if (json_num(options_json, "table_log2", &d)) E.table_log2 = (int)d;
  if (json_num(options_json, "max_states", &d)) E.max_states = (uint64_t)d;
  if (json_num(options_json, "cand_bytes", &d)) E.cand_bytes = (uint64_t)d;
  if (json_num(options_json, "rank", &d)) E.rank = (uint32_t)d;
  if (json_num(options_json, "world", &d)) E.world = (uint32_t)d;
  if (json_bool(options_json, "continue", &b)) E.cont = b;
  if (json_bool(options_json, "check_deadlock", &b)) E.check_deadlock = b;
  if (json_bool(options_json, "timing", &b)) E.timing = b;
  if (json_bool(options_json, "count_actions", &b)) E.count_actions = b;
  if (json_num(options_json, "stop_after_states", &d)) E.stop_after_states = (uint64_t)d;
  if (json_num(options_json, "l2_fetch", &d)) E.l2_fetch = (int)d;
  if (json_bool(options_json, "one_phase", &b)) E.one_phase = b;
  if (json_bool(options_json, "prefetch", &b)) E.prefetch = b;
  if (json_bool(options_json, "overlap", &b)) E.overlap = b;
  if (json_bool(options_json, "spill", &b)) E.spill = b;
  json_str(options_json, "checkpoint_dir", &E.checkpoint_dir);
  json_str(options_json, "recover", &E.recover_dir);
  if (json_num(options_json, "checkpoint_minutes", &d)) E.checkpoint_minutes = d;
  if (json_num(options_json, "chunk_states", &d)) E.chunk_states_opt = (uint64_t)d;
  if (json_num(options_json, "fanout_bound", &d)) E.fanout_bound = (uint32_t)d;
  if (json_num(options_json, "stream", &d) && d != 0) {
    // a cudaStream_t handle of the calling process (e.g. torch.cuda.current_stream().cuda_stream): engine
    // kernels are then ordered with the caller's own work (NCCL exchange) without host synchronisation
    E.stream = reinterpret_cast<cudaStream_t>((uintptr_t)d);
    E.own_stream = false;
  }
  if (E.world < 1 || E.world > MAX_WORLD || E.rank >= E.world || (E.table_log2 && (E.table_log2 < 4 || E.table_log2 > 34))) {
    delete c;
    return KMC_E_BADARG;
  }
  if (json_num(options_json, "gpus", &d) && d > 1) {
    *out = c;
    return multi_create(c, options_json, (int)d);
  }
  int rc = engine_alloc(E);
  *out = c;   // returned even on failure so that the caller can read the error text
  if (rc == KMC_OK) rc = engine_reset(E);
  return rc;
}

void kmcm_destroy(kmcm_ctx* c) {
  if (!c) return;
  if (!c->ranks.empty()) {
    for (kmcm_ctx* r : c->ranks) kmcm_destroy(r);
    delete c;
    return;
  }
  cudaSetDevice(E.device);
  cudaFree(E.table);
  cudaFree(E.store);
  cudaFree(E.parent);
  cudaFree(E.cand);
  cudaFree(E.recv);
  if (E.peers_open)
    for (uint32_t r = 0; r < E.world; ++r)
      if (r != E.rank && E.peer_inbox[r] && !E.peers_direct) cudaIpcCloseMemHandle(E.peer_inbox[r] - SYNC_WORDS);
  cudaFree(E.inbox_alloc);
  if (E.board_host) cudaFreeHost(E.board_host);
  cudaFree(E.ctr);
  cudaFree(E.viol_ring);
  for (cudaEvent_t ev : E.event_pool) cudaEventDestroy(ev);
  if (E.ev_begin) cudaEventDestroy(E.ev_begin);
  if (E.ev_end) cudaEventDestroy(E.ev_end);
  if (E.stream && E.own_stream) cudaStreamDestroy(E.stream);
  if (E.stream2) cudaStreamDestroy(E.stream2);
  for (int i = 0; i < 2; ++i) {
    if (E.ev_exp[i]) cudaEventDestroy(E.ev_exp[i]);
    if (E.ev_ins[i]) cudaEventDestroy(E.ev_ins[i]);
  }
  delete c;
}

int kmcm_model_info(const kmcm_ctx*, kmc_model_info_t* out) {
  if (!out) return KMC_E_BADARG;
  memset(out, 0, sizeof(*out));
  out->words = W;
  out->state_bits = M::STATE_BITS;
  out->num_actions = M::NUM_ACTIONS;
  out->num_invariants = M::NUM_INVARIANTS;
  out->num_init = M::NUM_INIT;
  out->max_fanout = M::MAX_FANOUT;
  out->check_deadlock = M::CHECK_DEADLOCK;
  out->exact = EXACT_SET ? 1 : 0;
  strncpy(out->name, KMC_MODEL_NAME, sizeof(out->name) - 1);
  strncpy(out->digest, KMC_MODEL_DIGEST, sizeof(out->digest) - 1);
  return KMC_OK;
}

int kmcm_run(kmcm_ctx* c) {
  if (!c) return KMC_E_BADARG;
  if (!c->ranks.empty()) return multi_run(c);
  return engine_run(E);
}

int kmcm_stats(const kmcm_ctx* c, kmc_stats_t* out) {
  if (!c || !out) return KMC_E_BADARG;
  std::lock_guard<std::mutex> g(E.mu);
  *out = E.stats;
  return KMC_OK;
}

int kmcm_level_widths(const kmcm_ctx* c, uint64_t* out, size_t cap, size_t* n) {
  if (!c || !n) return KMC_E_BADARG;
  std::lock_guard<std::mutex> g(E.mu);
  *n = E.widths.size();
  for (size_t i = 0; i < E.widths.size() && i < cap; ++i) out[i] = E.widths[i];
  return KMC_OK;
}

int kmcm_action_counts(const kmcm_ctx* c, uint64_t* out, size_t cap, size_t* n) {
  if (!c || !n) return KMC_E_BADARG;
  std::lock_guard<std::mutex> g(E.mu);
  *n = std::min<size_t>(M::NUM_ACTIONS, E.action_counts.size());
  for (size_t i = 0; i < *n && i < cap; ++i) out[i] = E.action_counts[i];
  return KMC_OK;
}

int kmcm_violation(const kmcm_ctx* c, kmc_violation_t* out) {
  if (!c || !out) return KMC_E_BADARG;
  if (!E.ran && E.shard_levels == 0) return KMC_E_STATE;
  *out = E.viol;
  return KMC_OK;
}

int kmcm_trace_state(const kmcm_ctx* c, uint32_t i, uint64_t* buf, size_t cap_words, uint32_t* action_id) {
  if (!c || !buf) return KMC_E_BADARG;
  if (i >= E.trace.size() || cap_words < (size_t)W) return KMC_E_BADARG;
  memcpy(buf, E.trace[i].data(), W * 8);
  if (action_id) *action_id = E.trace_actions[i];
  return KMC_OK;
}

// the offending state of this rank (before any cross-rank trace walk): packed words + parent/action word
int kmcm_violation_record(const kmcm_ctx* c, uint64_t* words, size_t cap_words, uint64_t* parent_meta) {
  if (!c || !words || !parent_meta || cap_words < (size_t)W) return KMC_E_BADARG;
  if (E.viol.kind == KMC_RESULT_OK || E.viol_words.size() != (size_t)W) return KMC_E_STATE;
  memcpy(words, E.viol_words.data(), W * 8);
  *parent_meta = E.viol_meta;
  return KMC_OK;
}

int kmcm_copy_parents(const kmcm_ctx* c_, uint64_t first, uint64_t count, uint64_t* buf) {
  kmcm_ctx* c = const_cast<kmcm_ctx*>(c_);
  if (!c || !buf) return KMC_E_BADARG;
  if (!c->ranks.empty()) return KMC_E_STATE;      // per-rank stores: address a rank's own context
  CK(cudaSetDevice(E.device));
  if (E.spill) {
    std::vector<uint64_t> tmp(W);
    for (uint64_t i = 0; i < count; ++i) {
      int rc = fetch_state(E, first + i, tmp.data(), buf + i);
      if (rc) return rc;
    }
    return KMC_OK;
  }
  if (first + count > E.max_states) return KMC_E_BADARG;
  CK(cudaMemcpy(buf, E.parent + first, count * 8, cudaMemcpyDeviceToHost));
  return KMC_OK;
}

int kmcm_copy_states(const kmcm_ctx* c_, uint64_t first, uint64_t count, uint64_t* buf) {
  kmcm_ctx* c = const_cast<kmcm_ctx*>(c_);
  if (!c || !buf) return KMC_E_BADARG;
  if (!c->ranks.empty()) return KMC_E_STATE;
  CK(cudaSetDevice(E.device));
  if (E.spill) {
    uint64_t meta;
    for (uint64_t i = 0; i < count; ++i) {
      int rc = fetch_state(E, first + i, buf + i * W, &meta);
      if (rc) return rc;
    }
    return KMC_OK;
  }
  if (first + count > E.max_states) return KMC_E_BADARG;
  CK(cudaMemcpy(buf, E.store + first * W, count * W * 8, cudaMemcpyDeviceToHost));
  return KMC_OK;
}

const char* kmcm_strerror(const kmcm_ctx* c, int code) {
  switch (code) {
    case KMC_OK: return "ok";
    case KMC_E_BADARG: return (c && !E.last_error.empty()) ? E.last_error.c_str() : "bad argument";
    case KMC_E_CUDA: return (c && !E.last_error.empty()) ? E.last_error.c_str() : "CUDA error";
    case KMC_E_OOM: return "out of device memory";
    case KMC_E_TABLE_FULL: return "fingerprint set is full (raise table_log2)";
    case KMC_E_STORE_FULL: return "state store is full (raise max_states)";
    case KMC_E_LAYOUT_OVERFLOW: return "a successor value does not fit the packed state layout";
    case KMC_E_MODEL: return "cannot load the lowered model library";
    case KMC_E_STATE: return "call sequence error";
    case KMC_E_NO_GPU: return "no CUDA device visible; this library has no CPU fallback";
    case KMC_E_CAND_FULL: return "candidate buffer overflow (raise cand_bytes or fanout_bound)";
    case KMC_E_PEER_TIMEOUT: return "a peer rank did not arrive at a device-side synchronisation point within 30 s";
    default: return "unknown error";
  }
}

// ---- fingerprint set alone ---------------------------------------------------------------
static int fpset_call(kmcm_ctx* c, const uint64_t* fps, size_t n, uint8_t* out, int insert) {
  if (!c || (!fps && n) || (!out && n)) return KMC_E_BADARG;
  if (n == 0) return KMC_OK;
  CK(cudaSetDevice(E.device));
  uint64_t* d_fps = nullptr;
  uint8_t* d_out = nullptr;
  CK(cudaMalloc(&d_fps, n * 8));
  CK(cudaMalloc(&d_out, n));
  CK(cudaMemcpyAsync(d_fps, fps, n * 8, cudaMemcpyHostToDevice, E.stream));
  k_fpset_put<<<grid_for(E, n, 256, 8), 256, 0, E.stream>>>(E.table, E.table_slots / BUCKET_SLOTS - 1, d_fps, n, d_out, E.ctr, insert);
  CK(cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, E.stream));
  CK(cudaStreamSynchronize(E.stream));
  cudaFree(d_fps);
  cudaFree(d_out);
  unsigned long long f = 0;
  CK(cudaMemcpy(&f, &E.ctr->fail, 8, cudaMemcpyDeviceToHost));
  return fail_to_error(f);
}
int kmcm_fpset_put(kmcm_ctx* c, const uint64_t* fps, size_t n, uint8_t* out_seen) { return fpset_call(c, fps, n, out_seen, 1); }
int kmcm_fpset_contains(kmcm_ctx* c, const uint64_t* fps, size_t n, uint8_t* out) { return fpset_call(c, fps, n, out, 0); }
int kmcm_fpset_size(const kmcm_ctx* c_, uint64_t* out) {
  kmcm_ctx* c = const_cast<kmcm_ctx*>(c_);
  if (!c || !out) return KMC_E_BADARG;
  unsigned long long t = 0;
  CK(cudaSetDevice(E.device));
  CK(cudaMemcpy(&t, &E.ctr->store_tail, 8, cudaMemcpyDeviceToHost));
  *out = t;
  return KMC_OK;
}

// ---- sharded (multi-rank) building blocks -------------------------------------------------
int kmcm_shard_begin(kmcm_ctx* c) {
  if (!c) return KMC_E_BADARG;
  int rc = engine_reset(E);
  if (rc) return rc;
  E.inbox_buf = 0;
  E.shard_levels = 0;
  E.ran = false;
  CK(cudaEventRecord(E.ev_begin, E.stream));
  return KMC_OK;
}

int kmcm_shard_buffers(kmcm_ctx* c, kmc_shard_buffers_t* out) {
  if (!c || !out) return KMC_E_BADARG;
  out->cand = E.cand;
  out->region_rows = E.region_rows;
  out->cand_counts = (uint64_t*)E.ctr->cand_count;
  out->recv = E.world > 1 ? E.recv : E.cand;
  out->recv_rows_cap = E.world > 1 ? E.recv_rows : E.region_rows;
  out->row_words = ROW;
  return KMC_OK;
}

int kmcm_shard_seed_init(kmcm_ctx* c) {
  if (!c) return KMC_E_BADARG;
  return seed_init(E);
}

int kmcm_shard_expand(kmcm_ctx* c, uint64_t first, uint64_t count) {
  if (!c) return KMC_E_BADARG;
  if (count > E.chunk_states) {
    E.last_error = "expand chunk larger than chunk_states";
    return KMC_E_BADARG;
  }
  if (count == 0) return KMC_OK;
  return launch_expand(E, first, count);
}

int kmcm_shard_counts(kmcm_ctx* c, uint64_t* host_counts) {
  if (!c || !host_counts) return KMC_E_BADARG;
  unsigned long long tmp[MAX_WORLD];
  CK(cudaMemcpyAsync(tmp, E.ctr->cand_count, sizeof(tmp), cudaMemcpyDeviceToHost, E.stream));
  CK(cudaStreamSynchronize(E.stream));
  for (uint32_t d = 0; d < E.world; ++d) host_counts[d] = tmp[d];
  return KMC_OK;
}

int kmcm_shard_reset_cand(kmcm_ctx* c) {
  if (!c) return KMC_E_BADARG;
  return reset_cand(E);
}

int kmcm_shard_insert(kmcm_ctx* c, const uint64_t* rows_dev, uint64_t rows, uint64_t* new_tail) {
  if (!c) return KMC_E_BADARG;
  if (rows) {
    int rc = launch_insert(E, rows_dev, nullptr, rows, rows);
    if (rc) return rc;
  }
  if (new_tail) {
    DevCounters h;
    int rc = read_counters(E, &h);
    if (rc) return rc;
    *new_tail = h.store_tail;
    return fail_to_error(h.fail);
  }
  return KMC_OK;
}

int kmcm_shard_level_done(kmcm_ctx* c, uint64_t* level_first, uint64_t* level_count) {
  if (!c) return KMC_E_BADARG;
  DevCounters h;
  int rc = launch_invariants(E, E.level_first + E.level_count, std::max<uint64_t>(E.level_count * 2, 1024));
  if (rc) return rc;
  rc = read_counters(E, &h);
  if (rc) return rc;
  uint64_t prev_end = E.level_first + E.level_count;
  E.level_first = prev_end;
  E.level_count = h.store_tail - prev_end;
  E.shard_levels++;
  if (level_first) *level_first = E.level_first;
  if (level_count) *level_count = E.level_count;
  {
    std::lock_guard<std::mutex> g(E.mu);
    E.stats.distinct = h.store_tail;
    E.stats.generated = h.generated;
    E.stats.deadlocks = h.deadlocks;
    E.stats.out_of_model = h.out_of_model;
    E.stats.probes = h.probes;
    E.stats.table_slots = E.table_slots;
    E.stats.slot_bytes = SLOT_BYTES;
    E.stats.max_states = E.max_states;
    if (E.level_count) E.widths.push_back(E.level_count);
    E.stats.levels = E.stats.depth = E.widths.size();
  }
  if (h.viol_count && E.viol.kind == KMC_RESULT_OK) build_trace(E, h, E.shard_levels - 1);
  return fail_to_error(h.fail);
}

// ---- fused exchange over peer memory ---------------------------------------------------------
int kmcm_shard_ipc_handle(kmcm_ctx* c, void* out64) {
  if (!c || !out64 || !E.inbox) return KMC_E_BADARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  CK(cudaSetDevice(E.device));
  CK(cudaIpcGetMemHandle(&h, E.inbox_alloc));
  memcpy(out64, &h, 64);
  return KMC_OK;
}

int kmcm_shard_open_peers(kmcm_ctx* c, const void* handles, uint32_t world) {
  if (!c || !handles || world != E.world || !E.inbox) return KMC_E_BADARG;
  CK(cudaSetDevice(E.device));
  for (uint32_t r = 0; r < world; ++r) {
    if (r == E.rank) {
      E.peer_inbox[r] = E.inbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, 64);
    void* ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    E.peer_inbox[r] = (uint64_t*)ptr + SYNC_WORDS;
  }
  E.peers_open = true;
  return KMC_OK;
}

// expand a frontier chunk, storing every successor row directly into its owner's inbox, then
// publish the per-owner row counts into the owners' inbox headers (both on the engine stream)
int kmcm_shard_expand_p2p(kmcm_ctx* c, uint64_t first, uint64_t count) {
  if (!c || !E.peers_open) return KMC_E_STATE;
  if (count > E.chunk_states) return KMC_E_BADARG;
  int rc = reset_cand(E);
  if (rc) return rc;
  if ((rc = launch_expand(E, first, count, true))) return rc;
  Params p = E.params();
  p.p2p = 1;
  {
    TimedLaunch t(E, 2);
    k_publish_counts<<<1, 32, 0, E.stream>>>(p, 0);
  }
  CK(cudaGetLastError());
  return KMC_OK;
}

// seed: the initial states go through the same inbox path (rank 0 contributes them)
static int seed_p2p(kmcm_ctx* c, bool publish) {
  if (!c || !E.peers_open) return KMC_E_STATE;
  CK(cudaSetDevice(E.device));
  unsigned long long counts[MAX_WORLD] = {0};
  if (E.rank == 0) {
    for (int i = 0; i < M::NUM_INIT; ++i) {
      State s;
      memcpy(s.w, M::INIT_STATES[i], sizeof(s.w));
      uint32_t d = owner_of(state_fp(s), E.world);
      uint64_t row[ROW];
      for (int k = 0; k < W; ++k) row[k] = s.w[k];
      row[W] = NO_PARENT;
      uint64_t* dst = E.peer_inbox[d] + (uint64_t)E.inbox_buf * E.inbox_stride + INBOX_HEADER +
                      ((uint64_t)E.rank * E.region_rows + counts[d]) * ROW;
      CK(cudaMemcpyAsync(dst, row, sizeof(row), cudaMemcpyHostToDevice, E.stream));
      CK(cudaStreamSynchronize(E.stream));
      counts[d]++;
    }
    unsigned long long gen = M::NUM_INIT;
    CK(cudaMemcpyAsync(&E.ctr->generated, &gen, sizeof(gen), cudaMemcpyHostToDevice, E.stream));
  }
  CK(cudaMemcpyAsync(E.ctr, counts, sizeof(counts), cudaMemcpyHostToDevice, E.stream));
  if (publish) {
    Params p = E.params();
    p.p2p = 1;
    k_publish_counts<<<1, 32, 0, E.stream>>>(p, 0);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(E.stream));
  return KMC_OK;
}
int kmcm_shard_seed_p2p(kmcm_ctx* c) { return seed_p2p(c, true); }

// insert everything the peers stored into the current inbox buffer, then switch buffers.
// The caller must have put a cross-rank barrier on the stream between expand_p2p and this call.
int kmcm_shard_insert_p2p(kmcm_ctx* c) {
  if (!c || !E.peers_open) return KMC_E_STATE;
  Params p = E.params();
  {
    TimedLaunch t(E, 1);
    k_insert_inbox<<<E.sms * 8, 256, 0, E.stream>>>(p);
  }
  CK(cudaGetLastError());
  E.inbox_buf ^= 1;
  return KMC_OK;
}

// One expand -> exchange -> insert round with device-side cross-rank synchronisation (no NCCL, no host wait):
//   wait until every destination has consumed the buffer this round reuses (done >= round - 2)
//   expand (or, seed != 0, store the initial states) straight into the owners' inboxes; publish counts + ready
//   wait until every source is ready for this round; insert from the own inbox; publish done
// Every rank must call it the same number of times (count = 0 on ranks without work).
int kmcm_shard_round_p2p(kmcm_ctx* c, uint64_t first, uint64_t count, int seed) {
  if (!c || !E.peers_open) return KMC_E_STATE;
  if (count > E.chunk_states) return KMC_E_BADARG;
  CK(cudaSetDevice(E.device));
  const uint64_t round = ++E.round;
  E.inbox_buf = (uint32_t)(round & 1);
  int rc;
  if (round > 2) k_wait_flags<<<1, 32, 0, E.stream>>>(E.inbox_alloc + SYNC_DONE, E.world, round - 2, E.ctr);
  if (seed) {
    if ((rc = seed_p2p(c, false))) return rc;
  } else {
    if ((rc = reset_cand(E))) return rc;
    if ((rc = launch_expand(E, first, count, true))) return rc;
  }
  Params p = E.params();
  p.p2p = 1;
  k_publish_counts<<<1, 32, 0, E.stream>>>(p, round);
  k_wait_flags<<<1, 32, 0, E.stream>>>(E.inbox_alloc + SYNC_READY, E.world, round, E.ctr);
  {
    TimedLaunch t(E, 1);
    k_insert_inbox<<<E.sms * 8, 256, 0, E.stream>>>(p);
  }
  k_publish_done<<<1, 32, 0, E.stream>>>(p, round);
  CK(cudaGetLastError());
  return KMC_OK;
}

// Level end on all ranks at once: invariants on this rank's new states, publish the summary to every board, wait
// for all summaries, copy the board to pinned host memory, ONE stream synchronisation.  board_out receives
// world x 8 words: {level id, new states, violations, store tail, generated, fail, deadlocks, -} per rank.
int kmcm_shard_level_sync(kmcm_ctx* c, uint64_t* board_out) {
  if (!c || !E.peers_open || !board_out) return KMC_E_STATE;
  CK(cudaSetDevice(E.device));
  int rc = launch_invariants(E, E.level_first + E.level_count, std::max<uint64_t>(E.level_count * 2, 1024));
  if (rc) return rc;
  const uint64_t level_id = ++E.level_id;
  Params p = E.params();
  k_publish_level<<<1, 32, 0, E.stream>>>(p, level_id, E.level_first + E.level_count);
  uint64_t* dev_view = nullptr;
  CK(cudaHostGetDevicePointer((void**)&dev_view, E.board_host, 0));
  k_gather_level<<<1, 32, 0, E.stream>>>(p, level_id, dev_view);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(E.stream));
  memcpy(board_out, E.board_host, (size_t)E.world * BOARD_WORDS * 8);
  const uint64_t* mine = E.board_host + (size_t)E.rank * BOARD_WORDS;
  const uint64_t prev_end = E.level_first + E.level_count;
  E.level_first = prev_end;
  E.level_count = mine[1];
  E.shard_levels++;
  {
    std::lock_guard<std::mutex> g(E.mu);
    E.stats.distinct = mine[3];
    E.stats.generated = mine[4];
    E.stats.deadlocks = mine[6];
    E.stats.table_slots = E.table_slots;
    E.stats.slot_bytes = SLOT_BYTES;
    E.stats.max_states = E.max_states;
    if (E.level_count) E.widths.push_back(E.level_count);
    E.stats.levels = E.stats.depth = E.widths.size();
  }
  if (mine[2] && E.viol.kind == KMC_RESULT_OK) {
    DevCounters h;
    if ((rc = read_counters(E, &h))) return rc;
    build_trace(E, h, E.shard_levels - 1);
  }
  return fail_to_error(mine[5]);
}

// same-process peers (one context per GPU in one process): direct pointers instead of CUDA IPC handles.
// inboxes[r] = the value kmcm_shard_inbox_ptr returned for rank r's context.
int kmcm_shard_inbox_ptr(kmcm_ctx* c, void** out) {
  if (!c || !out || !E.inbox_alloc) return KMC_E_BADARG;
  *out = E.inbox_alloc;
  return KMC_OK;
}
int kmcm_shard_open_peers_direct(kmcm_ctx* c, void* const* inboxes, const int* devices, uint32_t world) {
  if (!c || !inboxes || !devices || world != E.world || !E.inbox_alloc) return KMC_E_BADARG;
  CK(cudaSetDevice(E.device));
  for (uint32_t r = 0; r < world; ++r) {
    if (r != E.rank) {
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        E.last_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
        return KMC_E_CUDA;
      }
      cudaGetLastError();
    }
    E.peer_inbox[r] = (uint64_t*)inboxes[r] + SYNC_WORDS;
  }
  E.peers_open = true;
  E.peers_direct = true;
  return KMC_OK;
}

int kmcm_shard_sync(kmcm_ctx* c) {
  if (!c) return KMC_E_BADARG;
  CK(cudaEventRecord(E.ev_end, E.stream));
  DevCounters hc;
  {
    int rc = read_counters(E, &hc);              // synchronises the stream
    if (rc) return rc;
    std::lock_guard<std::mutex> g(E.mu);
    E.stats.probes = hc.probes;
    E.stats.out_of_model = hc.out_of_model;
    E.stats.generated = hc.generated;
    E.stats.deadlocks = hc.deadlocks;
  }
  float total_ms = 0;
  cudaEventElapsedTime(&total_ms, E.ev_begin, E.ev_end);
  std::lock_guard<std::mutex> g(E.mu);
  E.stats.gpu_ms_total = total_ms;
  if (E.timing) accumulate_timing(E, E.stats);
  return KMC_OK;
}

}  // extern "C"
#undef E

// ----------------------------------------------------------------------------------------
// N GPUs behind one context (kmc_create option "gpus": N): tlc2 -workers N, or any C / JNI caller.
// The ranks are ordinary sub-contexts driven through the same kmcm_shard_* entry points the multi-process
// driver uses; they see each other's inboxes through direct peer pointers.  Every rank thread reads the same
// level board, so all of them take the same decisions without any host-side barrier.
// ----------------------------------------------------------------------------------------
static int multi_create(kmcm_ctx* c, const char* options_json, int gpus) {
  Engine& A = c->e;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    A.last_error = "no CUDA device visible; this library has no CPU fallback";
    return KMC_E_NO_GPU;
  }
  if (gpus > ndev || gpus > MAX_WORLD) {
    A.last_error = "option gpus exceeds the visible devices";
    return KMC_E_BADARG;
  }
  std::string base = options_json ? options_json : "{}";
  size_t close = base.rfind('}');
  if (close == std::string::npos) return KMC_E_BADARG;
  const int dev0 = A.device;
  for (int r = 0; r < gpus; ++r) {
    // the sub-context's own keys go first: json_find takes the first occurrence of a key
    std::string opt = "{\"device\": " + std::to_string(dev0 + r) + ", \"rank\": " + std::to_string(r) + ", \"world\": " +
                      std::to_string(gpus) + ", \"gpus\": 0, " + base.substr(base.find('{') + 1);
    kmcm_ctx* sub = nullptr;
    int rc = kmcm_create(opt.c_str(), &sub);
    if (sub) c->ranks.push_back(sub);
    if (rc) {
      A.last_error = sub ? sub->e.last_error : "cannot create a rank context";
      return rc;
    }
  }
  void* inboxes[MAX_WORLD] = {};
  int devices[MAX_WORLD] = {};
  for (int r = 0; r < gpus; ++r) {
    kmcm_shard_inbox_ptr(c->ranks[r], &inboxes[r]);
    devices[r] = dev0 + r;
  }
  for (int r = 0; r < gpus; ++r) {
    int rc = kmcm_shard_open_peers_direct(c->ranks[r], inboxes, devices, (uint32_t)gpus);
    if (rc) {
      A.last_error = c->ranks[r]->e.last_error;
      return rc;
    }
  }
  A.world = (uint32_t)gpus;
  return KMC_OK;
}

struct RankOutcome {
  int rc = KMC_OK;
  std::vector<uint64_t> levels;
  bool stopped = false;
  uint64_t board[MAX_WORLD * BOARD_WORDS] = {};
};

static void rank_loop(kmcm_ctx* sub, bool cont, uint64_t stop_after, RankOutcome* out) {
  Engine& R = sub->e;
  const uint32_t world = R.world, rank = R.rank;
  uint64_t* board = out->board;
  auto fail = [&](int rc) { out->rc = rc; };
  int rc;
  if ((rc = kmcm_shard_begin(sub))) return fail(rc);
  if ((rc = kmcm_shard_round_p2p(sub, 0, 0, 1))) return fail(rc);
  rc = kmcm_shard_level_sync(sub, board);
  uint64_t first = 0;
  for (;;) {
    // a failure flag of ANY rank ends the run on every rank (they all read the same board)
    int err = rc;
    uint64_t total = 0, viol = 0, max_new = 0, distinct = 0;
    for (uint32_t r = 0; r < world; ++r) {
      const uint64_t* b = board + r * BOARD_WORDS;
      total += b[1];
      viol += b[2];
      distinct += b[3];
      max_new = std::max(max_new, b[1]);
      if (!err && b[5]) err = fail_to_error(b[5]);
    }
    if (err) return fail(err);
    if (viol && !cont) { out->stopped = true; break; }
    if (total == 0) break;
    out->levels.push_back(total);
    if (stop_after && distinct >= stop_after) { out->stopped = true; break; }
    const uint64_t count = board[rank * BOARD_WORDS + 1];
    const uint64_t n_chunks = (max_new + R.chunk_states - 1) / R.chunk_states;
    for (uint64_t ci = 0; ci < n_chunks; ++ci) {
      const uint64_t off = ci * R.chunk_states;
      const uint64_t n = off < count ? std::min<uint64_t>(R.chunk_states, count - off) : 0;
      if ((rc = kmcm_shard_round_p2p(sub, first + off, n, 0))) return fail(rc);
    }
    first += count;
    rc = kmcm_shard_level_sync(sub, board);
  }
  if ((rc = kmcm_shard_sync(sub))) return fail(rc);
}

static int multi_run(kmcm_ctx* c) {
  Engine& A = c->e;
  const size_t n = c->ranks.size();
  auto t0 = std::chrono::steady_clock::now();
  std::vector<RankOutcome> out(n);
  std::vector<std::thread> th;
  for (size_t r = 0; r < n; ++r) th.emplace_back(rank_loop, c->ranks[r], A.cont, A.stop_after_states, &out[r]);
  for (auto& t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  int rc = KMC_OK;
  for (size_t r = 0; r < n; ++r)
    if (out[r].rc && !rc) {
      rc = out[r].rc;
      A.last_error = c->ranks[r]->e.last_error;
    }
  std::lock_guard<std::mutex> g(A.mu);
  kmc_stats_t& st = A.stats;
  memset(&st, 0, sizeof(st));
  A.widths = out[0].levels;
  for (size_t r = 0; r < n; ++r) {
    const kmc_stats_t& s = c->ranks[r]->e.stats;
    st.distinct += s.distinct;
    st.generated += s.generated;
    st.deadlocks += s.deadlocks;
    st.out_of_model += s.out_of_model;
    st.probes += s.probes;
    st.table_slots += s.table_slots;
    st.max_states += s.max_states;
    st.launches_expand += s.launches_expand;
    st.launches_insert += s.launches_insert;
    st.launches_other += s.launches_other;
    st.gpu_ms_total = std::max(st.gpu_ms_total, s.gpu_ms_total);
    st.gpu_ms_expand = std::max(st.gpu_ms_expand, s.gpu_ms_expand);
    st.gpu_ms_insert = std::max(st.gpu_ms_insert, s.gpu_ms_insert);
    st.gpu_ms_invariant = std::max(st.gpu_ms_invariant, s.gpu_ms_invariant);
    st.slot_bytes = s.slot_bytes;
  }
  st.depth = st.levels = A.widths.size();
  st.complete = (!rc && !out[0].stopped) ? 1 : 0;
  if (out[0].stopped && !rc) {
    // the states of the last level were inserted but not expanded: they are the queue
    for (size_t r = 0; r < n; ++r) st.queue += out[0].board[r * BOARD_WORDS + 1];
  }
  st.wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
  // violation: deadlocks (they belong to the level being expanded) before invariant violations, then the smallest
  // fingerprint, then the rank -- the single-GPU rule -- and the parent links are followed from store to store
  memset(&A.viol, 0, sizeof(A.viol));
  A.viol.invariant = -1;
  A.trace.clear();
  A.trace_actions.clear();
  const Engine* best = nullptr;
  for (size_t r = 0; r < n; ++r) {
    const Engine& R = c->ranks[r]->e;
    if (R.viol.kind == KMC_RESULT_OK) continue;
    if (!best) { best = &R; continue; }
    const bool rd = R.viol.kind == KMC_RESULT_DEADLOCK, bd = best->viol.kind == KMC_RESULT_DEADLOCK;
    if (rd != bd) { if (rd) best = &R; continue; }
    if (R.viol.fingerprint < best->viol.fingerprint) best = &R;
  }
  if (best && !rc) {
    std::vector<std::vector<uint64_t>> rev;
    std::vector<uint32_t> rev_act;
    uint64_t meta = best->viol_meta;
    rev.push_back(best->viol_words);
    rev_act.push_back((uint32_t)(meta >> 56));
    uint64_t guard = 0;
    while ((meta & 0x0000FFFFFFFFFFFFull) != NO_PARENT && guard++ < 100000) {
      const uint64_t idx = meta & IDX_MASK;
      const uint32_t prank = (uint32_t)((meta >> 40) & 0xFF);
      if (prank >= n) break;
      const Engine& P = c->ranks[prank]->e;
      if (idx >= P.max_states) break;
      std::vector<uint64_t> sw(W);
      if (cudaSetDevice(P.device) != cudaSuccess ||
          cudaMemcpy(sw.data(), P.store + idx * W, W * 8, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(&meta, P.parent + idx, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
        break;
      rev.push_back(sw);
      rev_act.push_back((uint32_t)(meta >> 56));
    }
    A.trace.assign(rev.rbegin(), rev.rend());
    A.trace_actions.assign(rev_act.rbegin(), rev_act.rend());
    A.viol = best->viol;
    A.viol.trace_len = A.trace.size();
    A.viol_words = best->viol_words;
    A.viol_meta = best->viol_meta;
  }
  A.ran = true;
  return rc;
}

