/* kspec_oracle.c -- Oracle B: hand-written plain-C restatement of the reference's TLA+ specs
 * plus an exact (full-state identity), multi-threaded, level-synchronous BFS.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs may build, load or run this file; nothing under kafka_specification_b200/ does.
 *
 * PARITY UNPINNED: the reference (hachikuji/kafka-specification @ d68782b) ships no TLC output,
 * no .cfg and no golden counts, and TLC (third-party tla2tools.jar, not vendored, version not
 * pinned) cannot run here (no JVM).  This restatement is pinned against: analytic closed forms
 * (IdSequence: MaxId+2 states; FiniteReplicatedLog: (sum_{e<=L} R^e)^n states, depth n*L+1), the
 * reference's own qualitative claims (Kip320.tla:168-171 hold; the TruncateToHW / Kip101 /
 * Kip279 / Kip320FirstTry header comments say StrongIsr fails), and bit-exact agreement with
 * Oracle A (oracle/tla_interp.py), which evaluates the unchanged .tla text.
 *
 * It is written by hand from the spec text, independently of the TLA+ front-end and of the
 * lowering, so that a parser or lowering bug cannot cancel out.  Every action cites the
 * reference lines it restates.  Successor multiplicities follow TLC's rule (every
 * positive-position \/ and bounded \E branches) so that "states generated"
 * is comparable, e.g. the double emission in Kip279.tla:47-51 and Kip320.tla:82-83.
 *
 * Build:  gcc -O3 -march=native -std=c11 -shared -fPIC -pthread kspec_oracle.c -o libkspec_oracle.so
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NMAX 5
#define LMAX 6
#define EMAX 7
#define NONE 255
#define MAX_SUCC 256

enum { M_IDSEQ = 0, M_FRL = 1, M_TRUNCHW = 2, M_KIP101 = 3, M_KIP279 = 4, M_KIP320 = 5, M_FIRSTTRY = 6, M_ASYNCISR = 7,
       M_KIP320_279 = 8 /* Kip320's Next with BecomeFollowerTruncateKip279 swapped in, the experiment Kip320.tla:126-133 suggests */ };
#define IS_KAFKA(m) (((m) >= M_TRUNCHW && (m) <= M_FIRSTTRY) || (m) == M_KIP320_279)
enum { INV_WEAKISR = 0, INV_STRONGISR = 1, INV_LEADERINISR = 2, INV_VALIDHW = 3, INV_COUNT = 4 };

/* ------------------------------------------------------------------------------------------
 * state records (canonical: all padding zero, unwritten log slots zero, request set sorted)
 * ---------------------------------------------------------------------------------------- */
typedef struct { int8_t epoch; uint8_t leader; uint8_t isr; } Req;   /* QuorumState record, KafkaReplication.tla:87-89 */

typedef struct {
  uint8_t end[NMAX];              /* replicaLog[r].endOffset          FiniteReplicatedLog.tla:41 */
  uint8_t rec_id[NMAX][LMAX];     /* replicaLog[r].records[o].id      KafkaReplication.tla:82    */
  uint8_t rec_ep[NMAX][LMAX];     /* replicaLog[r].records[o].epoch; slots >= end are Nil (0,0)  */
  uint8_t hw[NMAX];               /* replicaState[r].hw               KafkaReplication.tla:96-99 */
  int8_t rs_epoch[NMAX];          /* replicaState[r].leaderEpoch, Nil = -1                       */
  uint8_t rs_leader[NMAX];        /* replicaState[r].leader, None = NONE                         */
  uint8_t rs_isr[NMAX];           /* replicaState[r].isr as a bit set                            */
  uint8_t next_record_id;         /* KafkaReplication.tla:55                                     */
  uint8_t next_leader_epoch;      /* KafkaReplication.tla:59                                     */
  int8_t q_epoch;                 /* quorumState, KafkaReplication.tla:73                        */
  uint8_t q_leader;
  uint8_t q_isr;
  uint8_t nreq;                   /* leaderAndIsrRequests, KafkaReplication.tla:66               */
  Req req[EMAX + 1];
} KState;

typedef struct {
  uint8_t end[NMAX];
  uint8_t rec[NMAX][LMAX];        /* record index 0..R-1, slots >= end are Nil (0) */
} FState;

typedef struct { uint8_t next_id; } IState;

typedef struct {
  uint8_t c_isr, c_ver;           /* controllerState  AsyncIsr.tla:48-51 */
  uint8_t l_isr, l_ver, l_pisr;   /* leaderState      AsyncIsr.tla:40-46 */
  int8_t l_pver;                  /* pendingVersion, Nil = -1            */
  uint8_t off[NMAX];
  uint64_t requests[4];           /* bit (version * 2^n + isr)           */
  uint64_t updates[4];
} AState;

typedef struct {
  int model;
  int n, L, R, E;                 /* Kafka family / FRL (R = |LogRecords|) ; IdSeq: E = MaxId      */
  int M, V;                       /* AsyncIsr: MaxOffset, MaxVersion                                */
  size_t ssize;
} Cfg;

typedef struct {
  uint8_t* buf;                   /* MAX_SUCC successor records */
  int n;
  const Cfg* cfg;
} Out;

static inline void* out_slot(Out* o) {
  if (o->n >= MAX_SUCC) { fprintf(stderr, "oracle: MAX_SUCC exceeded\n"); abort(); }
  return o->buf + (size_t)o->n * o->cfg->ssize;
}

/* ------------------------------------------------------------------------------------------
 * IdSequence.tla
 * ---------------------------------------------------------------------------------------- */
/* Next == \E id \in IdSet : NextId(id)                       IdSequence.tla:39
 * NextId(id) == id <= MaxId /\ id = nextId /\ nextId' = nextId + 1      :30-33 */
static void idseq_expand(const Cfg* c, const IState* s, Out* o) {
  if (s->next_id <= c->E) {
    IState* t = out_slot(o);
    memset(t, 0, sizeof(*t));
    t->next_id = s->next_id + 1;
    o->n++;
  }
}

/* ------------------------------------------------------------------------------------------
 * FiniteReplicatedLog.tla (stand-alone root, config #2)
 * ---------------------------------------------------------------------------------------- */
static void frl_expand(const Cfg* c, const FState* s, Out* o) {
  /* Next == \E replica \in Replicas : ...                    FiniteReplicatedLog.tla:115-118 */
  for (int r = 0; r < c->n; ++r) {
    /* \/ \E record \in LogRecords, offset \in Offsets : Append(replica, record, offset)   :116
     * Append: ~IsFull /\ offset = endOffset /\ records[offset] := record, endOffset + 1   :99-103 */
    if (s->end[r] < c->L) {
      for (int rec = 0; rec < c->R; ++rec) {
        FState* t = out_slot(o);
        *t = *s;
        t->rec[r][s->end[r]] = (uint8_t)rec;
        t->end[r] = s->end[r] + 1;
        o->n++;
      }
    }
    /* \/ \E offset \in Offsets : TruncateTo(replica, offset)                                :117
     * TruncateTo: newEnd <= endOffset; slots >= newEnd := Nil; endOffset := newEnd          :105-109
     * (offset = endOffset is a generated self-loop; Offsets = 0..LogSize-1)                        */
    for (int off = 0; off < c->L; ++off) {
      if (off <= s->end[r]) {
        FState* t = out_slot(o);
        *t = *s;
        for (int k = off; k < LMAX; ++k) t->rec[r][k] = 0;
        t->end[r] = (uint8_t)off;
        o->n++;
      }
    }
    /* \/ \E otherReplica \in Replicas \ {replica} : ReplicateTo(replica, otherReplica)      :118
     * ReplicateTo(from, to) == \E offset, record : HasEntry(from, record, offset) /\ Append(to, record, offset)  :111-113 */
    for (int q = 0; q < c->n; ++q) {
      if (q == r) continue;
      int off = s->end[q];
      if (off < c->L && off < s->end[r]) {
        FState* t = out_slot(o);
        *t = *s;
        t->rec[q][off] = s->rec[r][off];
        t->end[q] = (uint8_t)(off + 1);
        o->n++;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * KafkaReplication.tla -- shared helpers
 * ---------------------------------------------------------------------------------------- */
static inline int presumes(const KState* s, int r) { return s->rs_leader[r] == r; }                 /* :126 */
static inline int following(const KState* s, int f, int l) { return s->rs_leader[f] == l; }          /* :127 */
static inline int is_true_leader(const KState* s, int l) {                                           /* :128-131 */
  return s->q_leader == l && presumes(s, l) && s->rs_epoch[l] == s->q_epoch;
}
static inline int same_record(const KState* s, int r1, int r2, int off) {
  return s->rec_id[r1][off] == s->rec_id[r2][off] && s->rec_ep[r1][off] == s->rec_ep[r2][off];
}

static int req_cmp(const Req* a, const Req* b) {
  if (a->epoch != b->epoch) return a->epoch < b->epoch ? -1 : 1;
  if (a->leader != b->leader) return a->leader < b->leader ? -1 : 1;
  if (a->isr != b->isr) return a->isr < b->isr ? -1 : 1;
  return 0;
}
static void req_insert(KState* t, Req q) {      /* set union with one element, kept sorted */
  int i = 0;
  while (i < t->nreq && req_cmp(&t->req[i], &q) < 0) ++i;
  if (i < t->nreq && req_cmp(&t->req[i], &q) == 0) return;
  if (t->nreq >= EMAX + 1) { fprintf(stderr, "oracle: request set overflow\n"); abort(); }
  for (int k = t->nreq; k > i; --k) t->req[k] = t->req[k - 1];
  t->req[i] = q;
  t->nreq++;
}

/* ControllerUpdateIsr(newLeader, newIsr)                       KafkaReplication.tla:138-145
 * \E newLeaderEpoch \in IdSet : NextId(newLeaderEpoch) [<= MaxLeaderEpoch, = nextLeaderEpoch, +1]
 * quorumState' = [leader, leaderEpoch, isr]; requests' = requests \union {that record}        */
static void controller_update_isr(const Cfg* c, const KState* s, int new_leader, int new_isr, Out* o) {
  if (s->next_leader_epoch > c->E) return;
  KState* t = out_slot(o);
  *t = *s;
  Req q;
  memset(&q, 0, sizeof(q));
  q.epoch = (int8_t)s->next_leader_epoch;
  q.leader = (uint8_t)new_leader;
  q.isr = (uint8_t)new_isr;
  t->next_leader_epoch = s->next_leader_epoch + 1;
  t->q_epoch = q.epoch;
  t->q_leader = q.leader;
  t->q_isr = q.isr;
  req_insert(t, q);
  o->n++;
}

/* ControllerShrinkIsr                                          KafkaReplication.tla:158-168 */
static void controller_shrink_isr(const Cfg* c, const KState* s, Out* o) {
  for (int r = 0; r < c->n; ++r) {
    int bit = 1 << r;
    if (s->q_leader == r && s->q_isr == bit) controller_update_isr(c, s, NONE, s->q_isr, o);          /* :159-161 */
    if (s->q_leader == r && s->q_isr != bit) controller_update_isr(c, s, NONE, s->q_isr & ~bit, o);   /* :162-164 */
    if (s->q_leader != r && (s->q_isr & bit)) controller_update_isr(c, s, s->q_leader, s->q_isr & ~bit, o); /* :165-167 */
  }
}

/* ControllerElectLeader == \E newLeader \in quorumState.isr : leader # newLeader /\ ...  :176-179 */
static void controller_elect_leader(const Cfg* c, const KState* s, Out* o) {
  for (int r = 0; r < c->n; ++r)
    if ((s->q_isr >> r) & 1)
      if (s->q_leader != r) controller_update_isr(c, s, r, s->q_isr, o);
}

/* BecomeLeader == \E request \in leaderAndIsrRequests : ...    KafkaReplication.tla:186-195 */
static void become_leader(const Cfg* c, const KState* s, Out* o) {
  (void)c;
  for (int i = 0; i < s->nreq; ++i) {
    const Req* q = &s->req[i];
    if (q->leader == NONE) continue;                                   /* leader # None            */
    int l = q->leader;
    if (!(q->epoch > s->rs_epoch[l])) continue;                        /* request epoch > local    */
    KState* t = out_slot(o);
    *t = *s;
    t->rs_epoch[l] = q->epoch;                                         /* hw kept (:191)           */
    t->rs_leader[l] = (uint8_t)l;
    t->rs_isr[l] = q->isr;
    o->n++;
  }
}

/* LeaderWrite == \E replica, id, offset : presumes /\ NextId(id) /\ Append(replica, [id, epoch], offset)  :202-207 */
static void leader_write(const Cfg* c, const KState* s, Out* o) {
  for (int r = 0; r < c->n; ++r) {
    if (!presumes(s, r)) continue;
    if (s->next_record_id > c->R - 1) continue;                        /* RecordSeq!NextId: id <= MaxRecords-1 (:78) */
    if (s->end[r] >= c->L) continue;                                   /* ~IsFull                  */
    KState* t = out_slot(o);
    *t = *s;
    t->rec_id[r][s->end[r]] = s->next_record_id;
    t->rec_ep[r][s->end[r]] = (uint8_t)s->rs_epoch[r];                 /* presumes => epoch >= 0   */
    t->end[r] = s->end[r] + 1;
    t->next_record_id = s->next_record_id + 1;
    o->n++;
  }
}

/* QuorumUpdateLeaderAndIsr(leader, newIsr)                     KafkaReplication.tla:213-217 */
static void quorum_update(const KState* s, int l, int new_isr, Out* o) {
  if (!is_true_leader(s, l)) return;
  KState* t = out_slot(o);
  *t = *s;
  t->q_isr = (uint8_t)new_isr;
  t->rs_isr[l] = (uint8_t)new_isr;
  o->n++;
}

/* IsFollowerCaughtUp(leader, follower, endOffset)              KafkaReplication.tla:219-225 */
static int is_follower_caught_up(const KState* s, int l, int f, int end_offset) {
  if (!following(s, f, l)) return 0;
  if (end_offset == 0) return 1;
  int off = end_offset - 1;
  return off < s->end[l] && off < s->end[f];     /* leader has an entry at off, follower has the offset */
}

/* LeaderShrinkIsr                                              KafkaReplication.tla:233-239 */
static void leader_shrink_isr(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    int isr = s->rs_isr[l];
    for (int r = 0; r < c->n; ++r) {
      if (r == l || !((isr >> r) & 1)) continue;
      if (is_follower_caught_up(s, l, r, s->end[l])) continue;
      quorum_update(s, l, isr & ~(1 << r), o);
    }
  }
}

/* LeaderExpandIsr                                              KafkaReplication.tla:248-254 */
static void leader_expand_isr(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    int isr = s->rs_isr[l];
    for (int r = 0; r < c->n; ++r) {
      if ((isr >> r) & 1) continue;
      if (!is_follower_caught_up(s, l, r, s->hw[l])) continue;
      quorum_update(s, l, isr | (1 << r), o);
    }
  }
}

/* LeaderIncHighWatermark == \E offset \in Offsets, leader : ...  KafkaReplication.tla:264-271 */
static void leader_inc_hw(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    if (!presumes(s, l)) continue;
    int off = s->hw[l];
    if (off > c->L - 1) continue;                                      /* offset \in Offsets       */
    int ok = 1;
    for (int f = 0; f < c->n && ok; ++f)
      if ((s->rs_isr[l] >> f) & 1) ok = following(s, f, l) && off < s->end[f];
    if (!ok) continue;
    KState* t = out_slot(o);
    *t = *s;
    t->hw[l] = s->hw[l] + 1;
    o->n++;
  }
}

static void truncate_log(KState* t, int r, int new_end) {   /* FiniteReplicatedLog.tla:105-109 */
  for (int k = new_end; k < LMAX; ++k) { t->rec_id[r][k] = 0; t->rec_ep[r][k] = 0; }
  t->end[r] = (uint8_t)new_end;
}

/* BecomeFollowerAndTruncateTo(leader, replica, truncationOffset)  KafkaReplication.tla:281-294
 * (callers bind leader \in Replicas, so the `leader = None` branch :285-286 never fires)          */
static void become_follower_and_truncate_to(const KState* s, int l, int r, int trunc, Out* o) {
  if (l == r) return;
  for (int i = 0; i < s->nreq; ++i) {
    const Req* q = &s->req[i];
    if (q->leader != l) continue;
    if (!(q->epoch > s->rs_epoch[r])) continue;
    if (!(trunc <= s->end[r])) continue;                               /* TruncateTo guard :106    */
    KState* t = out_slot(o);
    *t = *s;
    truncate_log(t, r, trunc);
    t->rs_epoch[r] = q->epoch;
    t->rs_leader[r] = (uint8_t)l;
    t->rs_isr[r] = q->isr;
    t->hw[r] = (uint8_t)(trunc < s->hw[r] ? trunc : s->hw[r]);         /* Min({trunc, @.hw}) :293  */
    o->n++;
  }
}

/* ReplicateTo(leader, follower) on replicaLog                   FiniteReplicatedLog.tla:111-113 */
static int can_replicate(const Cfg* c, const KState* s, int l, int f) {
  return s->end[f] < c->L && s->end[f] < s->end[l];
}
static void do_replicate(KState* t, const KState* s, int l, int f) {
  int off = s->end[f];
  t->rec_id[f][off] = s->rec_id[l][off];
  t->rec_ep[f][off] = s->rec_ep[l][off];
  t->end[f] = (uint8_t)(off + 1);
  int new_end = off + 1;                                               /* KafkaReplication.tla:306-309 */
  t->hw[f] = (uint8_t)(s->hw[l] < new_end ? s->hw[l] : new_end);
}

/* FollowerReplicate == \E follower, leader : ...                KafkaReplication.tla:302-310 */
static void follower_replicate(const Cfg* c, const KState* s, Out* o) {
  for (int f = 0; f < c->n; ++f)
    for (int l = 0; l < c->n; ++l) {
      if (!presumes(s, l) || !following(s, f, l)) continue;
      if (!can_replicate(c, s, l, f)) continue;
      KState* t = out_slot(o);
      *t = *s;
      do_replicate(t, s, l, f);
      o->n++;
    }
}

/* ------------------------------------------------------------------------------------------
 * variants
 * ---------------------------------------------------------------------------------------- */
/* KafkaTruncateToHighWatermark.tla:29-31 */
static void become_follower_truncate_hw(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l)
    for (int r = 0; r < c->n; ++r) become_follower_and_truncate_to(s, l, r, s->hw[r], o);
}

/* Kip101.tla:31-39  LookupOffsetForEpoch(leader, follower, epoch) */
static int lookup_offset_for_epoch(const KState* s, int l, int f, int epoch) {
  if (s->end[l] == 0) return s->hw[f];
  if (s->rec_ep[l][s->end[l] - 1] == epoch) return s->end[l];
  for (int off = 0; off < s->end[l]; ++off)                    /* Min(OffsetsWithLargerEpochs) :27-29 */
    if (s->rec_ep[l][off] > epoch) return off;
  return s->hw[f];
}
/* Kip101.tla:41-47 */
static void become_follower_truncate_kip101(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l)
    for (int r = 0; r < c->n; ++r) {
      if (s->end[r] == 0) {
        become_follower_and_truncate_to(s, l, r, 0, o);                       /* :42-43 */
      } else {
        int ep = s->rec_ep[r][s->end[r] - 1];                                  /* IsLatestRecord :45 */
        become_follower_and_truncate_to(s, l, r, lookup_offset_for_epoch(s, l, r, ep), o);
      }
    }
}

/* Kip279.tla:39-45  FirstNonMatchingOffsetFromTail(leader, follower) */
static int first_non_matching_from_tail(const KState* s, int l, int f) {
  if (s->end[l] == 0) return 0;
  int best = -1;                                    /* Max(MatchingOffsets(follower, leader)) :27-30 */
  for (int off = 0; off < s->end[f]; ++off)
    if (off < s->end[l] && same_record(s, f, l, off)) best = off;
  return best + 1;
}
/* Kip279.tla:47-51 -- the two disjuncts are NOT exclusive: an empty replica fires both */
static void become_follower_truncate_kip279(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l)
    for (int r = 0; r < c->n; ++r) {
      if (s->end[r] == 0) become_follower_and_truncate_to(s, l, r, 0, o);
      become_follower_and_truncate_to(s, l, r, first_non_matching_from_tail(s, l, r), o);
    }
}

/* Kip320.tla:39-42 */
static int following_leader_epoch(const KState* s, int l, int f) {
  return presumes(s, l) && s->rs_leader[f] == l && s->rs_epoch[f] == s->rs_epoch[l];
}
/* Kip320.tla:49-56 FencedFollowerFetch */
static void fenced_follower_fetch(const Cfg* c, const KState* s, Out* o) {
  for (int f = 0; f < c->n; ++f)
    for (int l = 0; l < c->n; ++l) {
      if (!following_leader_epoch(s, l, f)) continue;
      if (!can_replicate(c, s, l, f)) continue;
      KState* t = out_slot(o);
      *t = *s;
      do_replicate(t, s, l, f);
      o->n++;
    }
}
/* Kip320.tla:63-70 FencedLeaderIncHighWatermark */
static void fenced_leader_inc_hw(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    int hw = s->hw[l];
    if (!(hw < s->end[l])) continue;                                   /* HasOffset(leader, hw)    */
    int ok = 1;
    for (int f = 0; f < c->n && ok; ++f)
      if ((s->rs_isr[l] >> f) & 1) ok = following_leader_epoch(s, l, f) && hw < s->end[f];
    if (!ok) continue;
    KState* t = out_slot(o);
    *t = *s;
    t->hw[l] = (uint8_t)(hw + 1);
    o->n++;
  }
}
/* Kip320.tla:78-85 FencedLeaderShrinkIsr -- inner \/ (:82-83) is not exclusive: both -> 2 successors */
static void fenced_leader_shrink_isr(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    int isr = s->rs_isr[l];
    for (int f = 0; f < c->n; ++f) {
      if (f == l || !((isr >> f) & 1)) continue;
      if (!following_leader_epoch(s, l, f)) quorum_update(s, l, isr & ~(1 << f), o);
      if (s->end[f] < s->end[l]) quorum_update(s, l, isr & ~(1 << f), o);
    }
  }
}
/* Kip320.tla:87-92 / Kip320FirstTry.tla:122-127 */
static int hw_reached_current_epoch(const KState* s, int l) {
  int hw = s->hw[l];
  if (hw == s->end[l]) return 1;
  return hw < s->end[l] && s->rec_ep[l][hw] == s->rs_epoch[l];
}
/* Kip320.tla:94-98 */
static int follower_reached_hw(const KState* s, int l, int f) {
  int hw = s->hw[l];
  return hw == 0 || (hw - 1 < s->end[f]);
}
/* Kip320.tla:110-117 FencedLeaderExpandIsr */
static void fenced_leader_expand_isr(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    int isr = s->rs_isr[l];
    for (int f = 0; f < c->n; ++f) {
      if ((isr >> f) & 1) continue;
      if (!following_leader_epoch(s, l, f)) continue;
      if (!follower_reached_hw(s, l, f)) continue;
      if (!hw_reached_current_epoch(s, l)) continue;
      quorum_update(s, l, isr | (1 << f), o);
    }
  }
}
/* Kip320.tla:134-148 FencedBecomeFollowerAndTruncate (leader \in Replicas: the None branch is dead) */
static void fenced_become_follower_and_truncate(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l)
    for (int r = 0; r < c->n; ++r) {
      if (l == r) continue;
      for (int i = 0; i < s->nreq; ++i) {
        const Req* q = &s->req[i];
        if (q->leader != l) continue;
        if (!(q->epoch > s->rs_epoch[r])) continue;
        if (!presumes(s, l)) continue;                                 /* :142 */
        if (s->rs_epoch[l] != q->epoch) continue;                      /* :143 */
        int trunc = first_non_matching_from_tail(s, l, r);             /* :144 */
        if (!(trunc <= s->end[r])) continue;
        KState* t = out_slot(o);
        *t = *s;
        truncate_log(t, r, trunc);
        t->rs_epoch[r] = q->epoch;                                     /* BecomeFollower :119-124 */
        t->rs_leader[r] = q->leader;
        t->rs_isr[r] = q->isr;
        t->hw[r] = (uint8_t)(trunc < s->hw[r] ? trunc : s->hw[r]);
        o->n++;
      }
    }
}

/* Kip320FirstTry.tla:49-57 */
static int caught_up_to_leader_epoch(const KState* s, int l, int f, int end_offset) {
  if (!presumes(s, l) || !following(s, f, l)) return 0;
  if (end_offset == 0) return 1;
  int off = end_offset - 1;
  return off < s->end[l] && off < s->end[f] && s->rec_ep[f][off] == s->rec_ep[l][off];
}
/* Kip320FirstTry.tla:64-69 */
static int follower_needs_truncation(const KState* s, int f, int l) {
  if (s->end[f] > s->end[l]) return 1;
  if (s->end[f] == 0) return 0;
  int off = s->end[f] - 1;
  return off < s->end[l] && s->rec_ep[l][off] != s->rec_ep[f][off];
}
/* Kip320FirstTry.tla:75-82 FollowerTruncate */
static void follower_truncate(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l)
    for (int f = 0; f < c->n; ++f) {
      if (!presumes(s, l) || !following(s, f, l)) continue;
      if (!follower_needs_truncation(s, f, l)) continue;
      int trunc = first_non_matching_from_tail(s, l, f);
      if (!(trunc <= s->end[f])) continue;
      KState* t = out_slot(o);
      *t = *s;
      truncate_log(t, f, trunc);
      t->hw[f] = (uint8_t)(trunc < s->hw[f] ? trunc : s->hw[f]);
      o->n++;
    }
}
/* Kip320FirstTry.tla:90-97 ImprovedLeaderIncHighWatermark */
static void improved_leader_inc_hw(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    if (!presumes(s, l)) continue;
    int hw = s->hw[l];
    if (!(hw < s->end[l])) continue;                                   /* \E record: HasEntry(leader, record, hw) */
    int ok = 1;
    for (int f = 0; f < c->n && ok; ++f)
      if ((s->rs_isr[l] >> f) & 1) ok = caught_up_to_leader_epoch(s, l, f, hw + 1);
    if (!ok) continue;
    KState* t = out_slot(o);
    *t = *s;
    t->hw[l] = (uint8_t)(hw + 1);
    o->n++;
  }
}
/* Kip320FirstTry.tla:103-111 FollowerFetch */
static void follower_fetch(const Cfg* c, const KState* s, Out* o) {
  for (int f = 0; f < c->n; ++f)
    for (int l = 0; l < c->n; ++l) {
      if (!caught_up_to_leader_epoch(s, l, f, s->end[f])) continue;
      if (!can_replicate(c, s, l, f)) continue;
      KState* t = out_slot(o);
      *t = *s;
      do_replicate(t, s, l, f);
      o->n++;
    }
}
/* Kip320FirstTry.tla:114-120 */
static void leader_shrink_isr_better(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    int isr = s->rs_isr[l];
    for (int r = 0; r < c->n; ++r) {
      if (r == l || !((isr >> r) & 1)) continue;
      if (caught_up_to_leader_epoch(s, l, r, s->end[l])) continue;
      quorum_update(s, l, isr & ~(1 << r), o);
    }
  }
}
/* Kip320FirstTry.tla:134-141 */
static void leader_expand_isr_better(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l) {
    int isr = s->rs_isr[l];
    for (int r = 0; r < c->n; ++r) {
      if ((isr >> r) & 1) continue;
      if (!caught_up_to_leader_epoch(s, l, r, s->hw[l])) continue;
      if (!hw_reached_current_epoch(s, l)) continue;
      quorum_update(s, l, isr | (1 << r), o);
    }
  }
}
/* Kip320FirstTry.tla:148-157 BecomeFollower (no truncation, hw kept) */
static void first_try_become_follower(const Cfg* c, const KState* s, Out* o) {
  for (int l = 0; l < c->n; ++l)
    for (int r = 0; r < c->n; ++r) {
      if (l == r) continue;
      for (int i = 0; i < s->nreq; ++i) {
        const Req* q = &s->req[i];
        if (q->leader != l) continue;
        if (!(q->epoch > s->rs_epoch[r])) continue;
        KState* t = out_slot(o);
        *t = *s;
        t->rs_epoch[r] = q->epoch;
        t->rs_leader[r] = (uint8_t)l;
        t->rs_isr[r] = q->isr;
        o->n++;
      }
    }
}

static void kafka_expand(const Cfg* c, const KState* s, Out* o) {
  switch (c->model) {
    case M_TRUNCHW:   /* KafkaTruncateToHighWatermark.tla:33-42 */
    case M_KIP101:    /* Kip101.tla:49-58 */
    case M_KIP279:    /* Kip279.tla:53-62 */
      controller_elect_leader(c, s, o);
      controller_shrink_isr(c, s, o);
      become_leader(c, s, o);
      leader_expand_isr(c, s, o);
      leader_shrink_isr(c, s, o);
      leader_write(c, s, o);
      leader_inc_hw(c, s, o);
      if (c->model == M_TRUNCHW) become_follower_truncate_hw(c, s, o);
      else if (c->model == M_KIP101) become_follower_truncate_kip101(c, s, o);
      else become_follower_truncate_kip279(c, s, o);
      follower_replicate(c, s, o);
      break;
    case M_KIP320:    /* Kip320.tla:150-159 */
      controller_elect_leader(c, s, o);
      controller_shrink_isr(c, s, o);
      become_leader(c, s, o);
      fenced_leader_expand_isr(c, s, o);
      fenced_leader_shrink_isr(c, s, o);
      leader_write(c, s, o);
      fenced_leader_inc_hw(c, s, o);
      fenced_become_follower_and_truncate(c, s, o);
      fenced_follower_fetch(c, s, o);
      break;
    case M_KIP320_279:   /* Kip320.tla:150-159 with :157 replaced by Kip279.tla:47-51 (models/MCKip320With279.tla) */
      controller_elect_leader(c, s, o);
      controller_shrink_isr(c, s, o);
      become_leader(c, s, o);
      fenced_leader_expand_isr(c, s, o);
      fenced_leader_shrink_isr(c, s, o);
      leader_write(c, s, o);
      fenced_leader_inc_hw(c, s, o);
      become_follower_truncate_kip279(c, s, o);
      fenced_follower_fetch(c, s, o);
      break;
    case M_FIRSTTRY:  /* Kip320FirstTry.tla:159-169 */
      controller_elect_leader(c, s, o);
      controller_shrink_isr(c, s, o);
      become_leader(c, s, o);
      leader_expand_isr_better(c, s, o);
      leader_shrink_isr_better(c, s, o);
      leader_write(c, s, o);
      improved_leader_inc_hw(c, s, o);
      first_try_become_follower(c, s, o);
      follower_fetch(c, s, o);
      follower_truncate(c, s, o);
      break;
  }
}

/* WeakIsr / StrongIsr                                         KafkaReplication.tla:320-326, 334-340 */
static int isr_property(const Cfg* c, const KState* s, int strong) {
  for (int r1 = 0; r1 < c->n; ++r1) {
    if (!presumes(s, r1)) continue;
    int hw = s->hw[r1];
    if (hw == 0) continue;
    int isr = strong ? s->q_isr : s->rs_isr[r1];
    for (int r2 = 0; r2 < c->n; ++r2) {
      if (!((isr >> r2) & 1)) continue;
      for (int off = 0; off < hw; ++off) {
        if (!(off < s->end[r1] && off < s->end[r2] && same_record(s, r1, r2, off))) return 0;
      }
    }
  }
  return 1;
}
/* LeaderInIsr == quorumState.leader \in quorumState.isr       KafkaReplication.tla:345 */
static int leader_in_isr(const KState* s) {
  return s->q_leader != NONE && ((s->q_isr >> s->q_leader) & 1);
}

/* ------------------------------------------------------------------------------------------
 * SYMMETRY Permutations(Replicas) (TLC's symmetry reduction; models/MCKip320Sym.tla).  Replicas are
 * used only through equality and membership (KafkaReplication.tla:126-131,158-179), so permuting them
 * maps reachable states to reachable states.  The orbit representative is the permuted image with
 * the smallest byte string.
 * ---------------------------------------------------------------------------------------- */
static inline uint8_t perm_bits(uint8_t m, const int* pi, int n) {
  uint8_t o = 0;
  for (int r = 0; r < n; ++r) if ((m >> r) & 1) o |= (uint8_t)(1u << pi[r]);
  return o;
}
static void kafka_permute(const Cfg* c, const KState* s, const int* pi, KState* t) {
  memset(t, 0, sizeof(*t));
  for (int r = 0; r < c->n; ++r) {
    int q = pi[r];
    t->end[q] = s->end[r];
    memcpy(t->rec_id[q], s->rec_id[r], LMAX);
    memcpy(t->rec_ep[q], s->rec_ep[r], LMAX);
    t->hw[q] = s->hw[r];
    t->rs_epoch[q] = s->rs_epoch[r];
    t->rs_leader[q] = s->rs_leader[r] == NONE ? NONE : (uint8_t)pi[s->rs_leader[r]];
    t->rs_isr[q] = perm_bits(s->rs_isr[r], pi, c->n);
  }
  t->next_record_id = s->next_record_id;
  t->next_leader_epoch = s->next_leader_epoch;
  t->q_epoch = s->q_epoch;
  t->q_leader = s->q_leader == NONE ? NONE : (uint8_t)pi[s->q_leader];
  t->q_isr = perm_bits(s->q_isr, pi, c->n);
  for (int i = 0; i < s->nreq; ++i) {
    Req q;
    memset(&q, 0, sizeof(q));
    q.epoch = s->req[i].epoch;
    q.leader = s->req[i].leader == NONE ? NONE : (uint8_t)pi[s->req[i].leader];
    q.isr = perm_bits(s->req[i].isr, pi, c->n);
    req_insert(t, q);
  }
}
static int next_perm(int* a, int n) {          /* lexicographic next permutation */
  int i = n - 2;
  while (i >= 0 && a[i] > a[i + 1]) --i;
  if (i < 0) return 0;
  int j = n - 1;
  while (a[j] < a[i]) --j;
  int x = a[i]; a[i] = a[j]; a[j] = x;
  for (int l = i + 1, r = n - 1; l < r; ++l, --r) { x = a[l]; a[l] = a[r]; a[r] = x; }
  return 1;
}
static void kafka_canonicalize(const Cfg* c, KState* s) {
  int pi[NMAX];
  for (int r = 0; r < c->n; ++r) pi[r] = r;
  KState best = *s, t;
  while (next_perm(pi, c->n)) {
    kafka_permute(c, s, pi, &t);
    if (memcmp(&t, &best, sizeof(KState)) < 0) best = t;
  }
  *s = best;
}

/* ------------------------------------------------------------------------------------------
 * AsyncIsr.tla (+ the Bound constraint of models/MCAsyncIsr.tla); Leader = replica 0
 * ---------------------------------------------------------------------------------------- */
static inline int msg_bit(const Cfg* c, int isr, int ver) { return ver * (1 << c->n) + isr; }
static inline void bit_set(uint64_t* m, int b) { m[b >> 6] |= 1ull << (b & 63); }
static inline int bit_get(const uint64_t* m, int b) { return (int)((m[b >> 6] >> (b & 63)) & 1); }

/* HighWatermark == Min({offsets[r] : r \in isr \union pendingIsr})   AsyncIsr.tla:58-60 */
static int async_hw(const Cfg* c, const AState* s) {
  int set = s->l_isr | s->l_pisr, best = 1 << 30;
  for (int r = 0; r < c->n; ++r)
    if ((set >> r) & 1) if (s->off[r] < best) best = s->off[r];
  return best;
}
static void async_expand(const Cfg* c, const AState* s, Out* o) {
  int nver = c->V + 2;   /* versions representable: 0..V+1 (one past the constraint) */
  /* ControllerShrinkIsr   AsyncIsr.tla:72-79 */
  for (int r = 1; r < c->n; ++r) {
    if (!((s->c_isr >> r) & 1)) continue;
    AState* t = out_slot(o);
    *t = *s;
    t->c_ver = s->c_ver + 1;
    t->c_isr = s->c_isr & ~(1 << r);
    bit_set(t->updates, msg_bit(c, t->c_isr, t->c_ver));
    o->n++;
  }
  /* ControllerHandleRequest   AsyncIsr.tla:81-86 */
  for (int ver = 0; ver < nver; ++ver)
    for (int isr = 0; isr < (1 << c->n); ++isr) {
      if (!bit_get(s->requests, msg_bit(c, isr, ver))) continue;
      if (ver != s->c_ver) continue;
      AState* t = out_slot(o);
      *t = *s;
      t->c_ver = s->c_ver + 1;
      t->c_isr = (uint8_t)isr;
      bit_set(t->updates, msg_bit(c, isr, t->c_ver));
      o->n++;
    }
  /* LeaderRequestShrinkIsr   AsyncIsr.tla:88-100 */
  for (int r = 1; r < c->n; ++r) {
    if (!((s->l_isr >> r) & 1)) continue;
    int isr = s->l_isr & ~(1 << r);
    AState* t = out_slot(o);
    *t = *s;
    bit_set(t->requests, msg_bit(c, isr, s->l_ver));
    t->l_pisr = s->l_pisr | isr;
    t->l_pver = (int8_t)s->l_ver;
    o->n++;
  }
  /* LeaderRequestExpandIsr   AsyncIsr.tla:102-115 */
  for (int r = 0; r < c->n; ++r) {
    if ((s->l_isr >> r) & 1) continue;
    if (!(s->off[r] >= async_hw(c, s))) continue;
    int isr = s->l_isr | (1 << r);
    AState* t = out_slot(o);
    *t = *s;
    bit_set(t->requests, msg_bit(c, isr, s->l_ver));
    t->l_pisr = s->l_pisr | isr;
    t->l_pver = (int8_t)s->l_ver;
    o->n++;
  }
  /* LeaderWrite   AsyncIsr.tla:117-119 (unguarded) */
  {
    AState* t = out_slot(o);
    *t = *s;
    t->off[0] = s->off[0] + 1;
    o->n++;
  }
  /* LeaderHandleUpdate   AsyncIsr.tla:121-129 */
  for (int ver = 0; ver < nver; ++ver)
    for (int isr = 0; isr < (1 << c->n); ++isr) {
      if (!bit_get(s->updates, msg_bit(c, isr, ver))) continue;
      if (!(ver > s->l_ver)) continue;
      AState* t = out_slot(o);
      *t = *s;
      t->l_isr = (uint8_t)isr;
      t->l_ver = (uint8_t)ver;
      t->l_pisr = 0;
      t->l_pver = -1;
      o->n++;
    }
  /* FollowerReplicate   AsyncIsr.tla:131-135 */
  for (int r = 1; r < c->n; ++r) {
    if (!(s->off[r] < s->off[0])) continue;
    AState* t = out_slot(o);
    *t = *s;
    t->off[r] = s->off[r] + 1;
    o->n++;
  }
}
/* ValidHighWatermark   AsyncIsr.tla:161-162 */
static int async_valid_hw(const Cfg* c, const AState* s) {
  int hw = async_hw(c, s);
  for (int r = 0; r < c->n; ++r)
    if ((s->c_isr >> r) & 1) if (!(s->off[r] >= hw)) return 0;
  return 1;
}
/* Bound (models/MCAsyncIsr.tla) */
static int async_in_model(const Cfg* c, const AState* s) { return s->off[0] <= c->M && s->c_ver <= c->V; }

/* ------------------------------------------------------------------------------------------
 * dispatch
 * ---------------------------------------------------------------------------------------- */
static void expand(const Cfg* c, const void* s, Out* o) {
  o->n = 0;
  switch (c->model) {
    case M_IDSEQ: idseq_expand(c, s, o); break;
    case M_FRL: frl_expand(c, s, o); break;
    case M_ASYNCISR: async_expand(c, s, o); break;
    default: kafka_expand(c, s, o); break;
  }
}
static int in_model(const Cfg* c, const void* s) {
  return c->model == M_ASYNCISR ? async_in_model(c, s) : 1;
}
/* bit mask of violated invariants among those requested */
static unsigned violated(const Cfg* c, const void* s, unsigned want) {
  unsigned v = 0;
  if (IS_KAFKA(c->model)) {
    if ((want >> INV_WEAKISR) & 1) if (!isr_property(c, s, 0)) v |= 1u << INV_WEAKISR;
    if ((want >> INV_STRONGISR) & 1) if (!isr_property(c, s, 1)) v |= 1u << INV_STRONGISR;
    if ((want >> INV_LEADERINISR) & 1) if (!leader_in_isr(s)) v |= 1u << INV_LEADERINISR;
  } else if (c->model == M_ASYNCISR) {
    if ((want >> INV_VALIDHW) & 1) if (!async_valid_hw(c, s)) v |= 1u << INV_VALIDHW;
  }
  return v;
}
static void init_state(const Cfg* c, void* s) {
  memset(s, 0, c->ssize);
  if (IS_KAFKA(c->model)) {      /* KafkaReplication.tla:109-120 */
    KState* k = s;
    for (int r = 0; r < NMAX; ++r) { k->rs_epoch[r] = 0; k->rs_leader[r] = 0; }
    for (int r = 0; r < c->n; ++r) { k->rs_epoch[r] = -1; k->rs_leader[r] = NONE; }
    k->q_epoch = -1;
    k->q_leader = NONE;
    k->q_isr = (uint8_t)((1 << c->n) - 1);
  } else if (c->model == M_ASYNCISR) {                        /* AsyncIsr.tla:137-150 */
    AState* a = s;
    a->c_isr = a->l_isr = (uint8_t)((1 << c->n) - 1);
    a->l_pver = -1;
  }
}

/* ------------------------------------------------------------------------------------------
 * exact parallel BFS
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  uint64_t distinct, generated, depth, deadlocks, out_of_model, complete;
  uint64_t levels[256];
  uint64_t first_violation_level[INV_COUNT];
  uint64_t violating_states[INV_COUNT];
  double seconds;
  uint64_t state_size;
} kso_result;

typedef struct {
  Cfg cfg;
  uint8_t* store;            /* all states, fixed-size records */
  uint64_t store_cap;
  _Atomic uint64_t tail;
  _Atomic uint32_t* table;   /* 0 empty, else state index + 1 (u32: up to 4e9 states) */
  uint64_t table_mask;
  uint8_t* dead;             /* 1: slot allocated by a loser of an insertion race (hole) */
  unsigned want_inv;
  int symmetry;
  uint64_t stop_at;
  /* per level */
  uint64_t lvl_first, lvl_end, level;
  _Atomic uint64_t cursor;
  _Atomic uint64_t generated, deadlocks, oom, holes;
  _Atomic uint64_t viol_count[INV_COUNT];
  _Atomic uint64_t viol_level[INV_COUNT];
  _Atomic int overflow;
} Bfs;

static inline uint64_t hash_bytes(const uint8_t* p, size_t n) {
  uint64_t h = 0xcbf29ce484222325ull ^ n;
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t w;
    memcpy(&w, p + i, 8);
    h = (h ^ w) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
  }
  for (; i < n; ++i) { h ^= p[i]; h *= 0x100000001b3ull; }
  h ^= h >> 32; h *= 0xd6e8feb86659fd93ull; h ^= h >> 32;
  return h;
}

static void note_violation(Bfs* b, unsigned v, uint64_t level) {
  for (int i = 0; i < INV_COUNT; ++i)
    if ((v >> i) & 1) {
      atomic_fetch_add(&b->viol_count[i], 1);
      uint64_t cur = atomic_load(&b->viol_level[i]);
      while ((cur == 0 || level < cur) && !atomic_compare_exchange_weak(&b->viol_level[i], &cur, level)) {}
    }
}

/* returns 1 if the state was new.  A store slot is claimed only when an empty table slot is
 * reached; if the CAS then loses to a thread inserting the very same state, the claimed slot
 * becomes a hole (skipped by later levels). */
static int insert_state(Bfs* b, const uint8_t* st) {
  const size_t sz = b->cfg.ssize;
  uint64_t idx = UINT64_MAX;
  uint64_t h = hash_bytes(st, sz) & b->table_mask;
  for (;;) {
    uint32_t cur = atomic_load_explicit(&b->table[h], memory_order_acquire);
    if (cur == 0) {
      if (idx == UINT64_MAX) {
        idx = atomic_fetch_add(&b->tail, 1);
        if (idx >= b->store_cap) { atomic_store(&b->overflow, 1); return 0; }
        memcpy(b->store + idx * sz, st, sz);
      }
      uint32_t expect = 0;
      if (atomic_compare_exchange_strong_explicit(&b->table[h], &expect, (uint32_t)(idx + 1), memory_order_release,
                                                  memory_order_acquire))
        return 1;
      cur = expect;
    }
    if (memcmp(b->store + (uint64_t)(cur - 1) * sz, st, sz) == 0) {
      if (idx != UINT64_MAX) {
        b->dead[idx] = 1;
        atomic_fetch_add(&b->holes, 1);
      }
      return 0;
    }
    h = (h + 1) & b->table_mask;
  }
}

static void* worker(void* arg) {
  Bfs* b = arg;
  const Cfg* c = &b->cfg;
  const size_t sz = c->ssize;
  Out out;
  out.buf = malloc((size_t)MAX_SUCC * sz);
  out.cfg = c;
  out.n = 0;
  uint64_t gen = 0, dead = 0, oom = 0;
  for (;;) {
    if (atomic_load_explicit(&b->tail, memory_order_relaxed) > b->stop_at) break;   /* truncated run */
    uint64_t start = atomic_fetch_add(&b->cursor, 256);
    if (start >= b->lvl_end) break;
    uint64_t stop = start + 256 < b->lvl_end ? start + 256 : b->lvl_end;
    for (uint64_t i = start; i < stop; ++i) {
      if (b->dead[i]) continue;
      expand(c, b->store + i * sz, &out);
      gen += out.n;
      if (out.n == 0) ++dead;
      for (int k = 0; k < out.n; ++k) {
        uint8_t* t = out.buf + (size_t)k * sz;
        if (b->symmetry) kafka_canonicalize(c, (KState*)t);
        if (!in_model(c, t)) {
          ++oom;
          unsigned v = b->want_inv ? violated(c, t, b->want_inv) : 0;
          if (v) note_violation(b, v, b->level + 1);
          continue;
        }
        if (insert_state(b, t)) {
          unsigned v = b->want_inv ? violated(c, t, b->want_inv) : 0;
          if (v) note_violation(b, v, b->level + 1);
        }
      }
    }
  }
  atomic_fetch_add(&b->generated, gen);
  atomic_fetch_add(&b->deadlocks, dead);
  atomic_fetch_add(&b->oom, oom);
  free(out.buf);
  return NULL;
}

static size_t state_size(int model) {
  switch (model) {
    case M_IDSEQ: return sizeof(IState);
    case M_FRL: return sizeof(FState);
    case M_ASYNCISR: return sizeof(AState);
    default: return sizeof(KState);
  }
}

/* params: Kafka family {n, L, R, E}; FRL {n, L, R}; IdSequence {MaxId}; AsyncIsr {n, MaxOffset, MaxVersion}.
 * inv_mask: bit i = evaluate invariant i on every new state (statistics only; the search never stops).
 * dump: optional buffer receiving all distinct state records (dump_cap records).                    */
int kso_run_sym(int model, const int* params, int threads, uint64_t max_states, unsigned inv_mask, kso_result* res,
                uint8_t* dump, uint64_t dump_cap, int symmetry);

int kso_run(int model, const int* params, int threads, uint64_t max_states, unsigned inv_mask, kso_result* res,
            uint8_t* dump, uint64_t dump_cap) {
  return kso_run_sym(model, params, threads, max_states, inv_mask, res, dump, dump_cap, 0);
}

/* symmetry != 0: SYMMETRY Permutations(Replicas) (Kafka family only) */
int kso_run_sym(int model, const int* params, int threads, uint64_t max_states, unsigned inv_mask, kso_result* res,
                uint8_t* dump, uint64_t dump_cap, int symmetry) {
  if (symmetry && !IS_KAFKA(model)) return -3;
  Bfs* b = calloc(1, sizeof(Bfs));
  b->symmetry = symmetry;
  Cfg* c = &b->cfg;
  c->model = model;
  c->ssize = state_size(model);
  switch (model) {
    case M_IDSEQ: c->E = params[0]; break;
    case M_FRL: c->n = params[0]; c->L = params[1]; c->R = params[2]; break;
    case M_ASYNCISR: c->n = params[0]; c->M = params[1]; c->V = params[2]; break;
    default: c->n = params[0]; c->L = params[1]; c->R = params[2]; c->E = params[3]; break;
  }
  if (c->n > NMAX || c->L > LMAX || (model != M_IDSEQ && c->E > EMAX) || (model == M_ASYNCISR && (c->n > 4 || c->V > 13)))
    return -1;
  if (threads < 1) threads = 1;
  if (max_states == 0) max_states = 1ull << 26;
  b->store_cap = max_states + max_states / 8 + 1024;
  b->store = malloc(b->store_cap * c->ssize);
  b->dead = malloc(b->store_cap);
  uint64_t tcap = 1;
  while (tcap < b->store_cap * 2) tcap <<= 1;
  b->table = malloc(tcap * sizeof(uint32_t));
  b->table_mask = tcap - 1;
  b->want_inv = inv_mask;
  b->stop_at = max_states;
  if (!b->store || !b->dead || !b->table) return -2;
  /* allocation and first-touch page faults are outside the timed region (as on the GPU side,
   * where the set and the store are allocated before the clock starts) */
  memset((void*)b->table, 0, tcap * sizeof(uint32_t));
  memset(b->dead, 0, b->store_cap);
  for (uint64_t off = 0; off < b->store_cap * c->ssize; off += 4096) b->store[off] = 0;
  memset(res, 0, sizeof(*res));
  res->state_size = c->ssize;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);

  uint8_t* init = calloc(1, c->ssize);
  init_state(c, init);
  if (symmetry) kafka_canonicalize(c, (KState*)init);
  atomic_store(&b->generated, 1);
  if (in_model(c, init)) insert_state(b, init);
  unsigned v0 = inv_mask ? violated(c, init, inv_mask) : 0;
  if (v0) note_violation(b, v0, 1);
  free(init);

  b->lvl_first = 0;
  b->lvl_end = atomic_load(&b->tail);
  b->level = 1;
  pthread_t* th = malloc(sizeof(pthread_t) * threads);
  int complete = 1;
  while (b->lvl_end > b->lvl_first) {
    uint64_t holes_before = atomic_load(&b->holes);
    (void)holes_before;
    /* width of this level = live states in [lvl_first, lvl_end) */
    uint64_t width = 0;
    for (uint64_t i = b->lvl_first; i < b->lvl_end; ++i) width += !b->dead[i];
    if (width == 0) break;               /* only holes left: the previous level was the last one */
    if (b->level <= 256) res->levels[b->level - 1] = width;
    atomic_store(&b->cursor, b->lvl_first);
    int use = threads;                    /* small levels do not pay for idle threads */
    if (width < 64ull * (uint64_t)use) use = (int)(width / 64) + 1;
    if (use > threads) use = threads;
    if (use == 1) {
      worker(b);
    } else {
      for (int t = 0; t < use; ++t) pthread_create(&th[t], NULL, worker, b);
      for (int t = 0; t < use; ++t) pthread_join(th[t], NULL);
    }
    if (atomic_load(&b->overflow)) { complete = 0; break; }
    b->lvl_first = b->lvl_end;
    b->lvl_end = atomic_load(&b->tail);
    b->level++;
    if (atomic_load(&b->tail) - atomic_load(&b->holes) > max_states) { complete = 0; break; }
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  {
    uint64_t tail = atomic_load(&b->tail);
    if (tail > b->store_cap) tail = b->store_cap;
    res->distinct = tail - atomic_load(&b->holes);
  }
  res->generated = atomic_load(&b->generated);
  res->depth = b->level - 1 + (complete ? 0 : 1);
  res->deadlocks = atomic_load(&b->deadlocks);
  res->out_of_model = atomic_load(&b->oom);
  res->complete = (uint64_t)complete;
  for (int i = 0; i < INV_COUNT; ++i) {
    res->first_violation_level[i] = atomic_load(&b->viol_level[i]);
    res->violating_states[i] = atomic_load(&b->viol_count[i]);
  }
  res->seconds = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
  if (dump) {
    uint64_t k = 0, tail = atomic_load(&b->tail);
    if (tail > b->store_cap) tail = b->store_cap;
    for (uint64_t i = 0; i < tail && k < dump_cap; ++i)
      if (!b->dead[i]) { memcpy(dump + k * c->ssize, b->store + i * c->ssize, c->ssize); k++; }
  }
  free(th);
  free(b->store);
  free(b->dead);
  free((void*)b->table);
  free(b);
  return 0;
}

size_t kso_state_size(int model) { return state_size(model); }

/* ------------------------------------------------------------------------------------------
 * single-state entry points (tests/: every step of an error trace printed by the product is
 * re-checked here as a successor under THIS restatement of Next, not under the lowered code)
 * ---------------------------------------------------------------------------------------- */
static int fill_cfg(Cfg* c, int model, const int* params) {
  memset(c, 0, sizeof(*c));
  c->model = model;
  c->ssize = state_size(model);
  switch (model) {
    case M_IDSEQ: c->E = params[0]; break;
    case M_FRL: c->n = params[0]; c->L = params[1]; c->R = params[2]; break;
    case M_ASYNCISR: c->n = params[0]; c->M = params[1]; c->V = params[2]; break;
    default: c->n = params[0]; c->L = params[1]; c->R = params[2]; c->E = params[3]; break;
  }
  if (c->n > NMAX || c->L > LMAX || (model != M_IDSEQ && c->E > EMAX)) return -1;
  return 0;
}
/* successors of *state (TLC multiplicity: duplicates kept) into out_buf; returns their number, < 0 on error */
int kso_successors(int model, const int* params, const void* state, uint8_t* out_buf, int cap) {
  Cfg c;
  if (fill_cfg(&c, model, params)) return -1;
  uint8_t* tmp = malloc((size_t)MAX_SUCC * c.ssize);
  Out o = {tmp, 0, &c};
  expand(&c, state, &o);
  int n = o.n;
  if (n > cap) { free(tmp); return -2; }
  memcpy(out_buf, tmp, (size_t)n * c.ssize);
  free(tmp);
  return n;
}
int kso_init_state(int model, const int* params, void* out) {
  Cfg c;
  if (fill_cfg(&c, model, params)) return -1;
  init_state(&c, out);
  return 0;
}
unsigned kso_violated(int model, const int* params, const void* state, unsigned want) {
  Cfg c;
  if (fill_cfg(&c, model, params)) return 0;
  return violated(&c, state, want);
}
