"""Fingerprint-sharded BFS across ranks: one process per GPU, ``torch.distributed`` for the plumbing.

This is the multi-GPU form of the hot path (TLC's analogue is the
fingerprint-partitioned FPSet of ``tlc2.tool.distributed``).  Every rank owns the slice of the
fingerprint space ``owner = (fp >> 32) * world >> 32`` -- its own hash set, state store and
frontier.  Per BFS level and per frontier chunk:

  1. K1 expands the local frontier chunk and buckets each successor row by owner rank
     (``kmc_shard_expand``; the bucketing happens inside the expand kernel);
  2. ranks exchange per-destination row counts (all_gather) and then the rows themselves
     (all-to-all-v as one batch of isend/irecv -- NCCL over NVLink on GPUs, gloo in the CPU tests);
  3. K2 inserts what a rank received into its own set and appends new states to its own store
     (``kmc_shard_insert``).

A level ends with an all-reduce of (new states, violation flag); the search ends when no rank
found a new state.  Distinct = sum of per-rank set sizes, depth = number of levels: both are
independent of the partition, hence bit-exact across 1/2/4/8 ranks.

The class is written against a small engine interface (``ShardEngine``) so that the same driver
code runs over the CUDA engine (``CudaShardEngine``, via the C ABI) and, in the CPU-only tests,
over a host stand-in built from the test harness.
"""
from __future__ import annotations

import ctypes
import time
from dataclasses import dataclass, field

import torch
import torch.distributed as dist

from .runtime import Checker, KmcError, ShardBuffers


class ShardEngine:
    """What the driver needs from one rank's engine."""
    world: int
    rank: int
    row_words: int
    chunk_states: int
    device: torch.device

    def begin(self): ...
    def seed_init(self): ...
    def expand(self, first: int, count: int): ...
    def counts(self) -> list[int]: ...                     # rows produced for each owner (syncs)
    def send_view(self, dest: int, rows: int) -> torch.Tensor: ...   # int64 [rows * row_words]
    def reserve_recv(self, rows: int): ...                 # called once per round before recv_view
    def recv_view(self, offset_rows: int, rows: int) -> torch.Tensor: ...
    def insert_received(self, rows: int): ...              # rows contiguous at the start of recv
    def insert_local(self, rows: int): ...                 # world == 1 fast path: rows in region 0
    def reset_cand(self): ...
    def level_done(self) -> tuple[int, int]: ...           # (first, count) of the new level
    def finish(self): ...
    def stats(self) -> dict: ...
    def violation(self): ...


class _DevArray:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 3}


class CudaShardEngine(ShardEngine):
    """One rank of the CUDA engine, driven through the kmc_shard_* entry points."""

    def __init__(self, model: str, rank: int, world: int, device: int, **options):
        self.device = torch.device("cuda", device)
        # run the engine on a torch stream that is also the current stream of every collective the
        # driver issues: kernels and the NCCL exchange then order themselves on the device, the host
        # only synchronises where it needs a value.  (The legacy default stream has handle 0 and
        # cannot be passed, hence a dedicated stream.)
        self.stream = torch.cuda.Stream(self.device)
        options.setdefault("stream", self.stream.cuda_stream)
        self.ck = Checker(model, device=device, rank=rank, world=world, **options)
        self.rank, self.world = rank, world
        self.lib = self.ck.lib
        b = ShardBuffers()
        self.ck._check(self.lib.kmc_shard_buffers(self.ck.ctx, ctypes.byref(b)))
        self.row_words = b.row_words
        self.region_rows = b.region_rows
        self.recv_rows_cap = b.recv_rows_cap
        self.chunk_states = max(1, self.region_rows // min(self.ck.info.max_fanout, int(options.get("fanout_bound", 32))))
        n_cand = self.region_rows * world * self.row_words
        self.cand = torch.as_tensor(_DevArray(b.cand, n_cand), device=self.device)
        self.recv = (torch.as_tensor(_DevArray(b.recv, self.recv_rows_cap * self.row_words), device=self.device)
                     if world > 1 else self.cand)
        self._recv_ptr = b.recv if world > 1 else b.cand
        self._cand_ptr = b.cand
        self.counts_dev = torch.as_tensor(_DevArray(b.cand_counts, 8), device=self.device)[:world]
        self._matrix = torch.empty(world * world, dtype=torch.int64, device=self.device)
        self.p2p = False
        self.device_sync = bool(options.get("device_sync", True))
        if world > 1 and options.get("p2p", True) and dist.is_initialized():
            self._open_peers()

    def _open_peers(self):
        """Map every rank's inbox into this process (CUDA IPC) so that the expand kernel can store
        successor rows straight into their owner's memory over NVLink."""
        h = (ctypes.c_ubyte * 64)()
        self.ck._check(self.lib.kmc_shard_ipc_handle(self.ck.ctx, h))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(h))
        blob = b"".join(handles)
        rc = self.lib.kmc_shard_open_peers(self.ck.ctx, blob, self.world)
        ok = torch.tensor([1 if rc == 0 else 0], dtype=torch.int64, device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        self.p2p = bool(ok.item())
        if not self.p2p and self.rank == 0:
            print(f"[kspec-mc] CUDA IPC peer mapping unavailable (rc={rc}); using the NCCL all-to-all exchange")
        self._flag = torch.zeros(1, dtype=torch.int64, device=self.device)

    def seed_p2p(self):
        self.ck._check(self.lib.kmc_shard_seed_p2p(self.ck.ctx))

    def expand_p2p(self, first, count):
        self.ck._check(self.lib.kmc_shard_expand_p2p(self.ck.ctx, first, count))

    def insert_p2p(self):
        self.ck._check(self.lib.kmc_shard_insert_p2p(self.ck.ctx))

    def round_p2p(self, first, count, seed=False):
        self.ck._check(self.lib.kmc_shard_round_p2p(self.ck.ctx, first, count, 1 if seed else 0))

    def level_sync(self):
        """Level end with device-side synchronisation: returns the board, one row per rank:
        [level id, new states, violations, store tail, generated, fail, deadlocks, -]."""
        buf = (ctypes.c_uint64 * (8 * self.world))()
        self.ck._check(self.lib.kmc_shard_level_sync(self.ck.ctx, buf))
        return [[int(buf[r * 8 + k]) for k in range(8)] for r in range(self.world)]

    def barrier_on_stream(self, group=None):
        """Cross-rank barrier ordered on the engine's stream; the host does not wait."""
        dist.all_reduce(self._flag, group=group)

    def begin(self):
        self.ck._check(self.lib.kmc_shard_begin(self.ck.ctx))

    def seed_init(self):
        self.ck._check(self.lib.kmc_shard_seed_init(self.ck.ctx))

    def expand(self, first, count):
        self.ck._check(self.lib.kmc_shard_expand(self.ck.ctx, first, count))

    def counts(self):
        buf = (ctypes.c_uint64 * 8)()
        self.ck._check(self.lib.kmc_shard_counts(self.ck.ctx, buf))
        return [int(buf[d]) for d in range(self.world)]

    def count_matrix(self, group=None):
        """world x world matrix of rows produced per (source, owner): one collective on the device
        counters, one device->host copy (the only host synchronisation of a round)."""
        dist.all_gather_into_tensor(self._matrix, self.counts_dev, group=group)
        m = self._matrix.cpu().tolist()
        return [m[r * self.world:(r + 1) * self.world] for r in range(self.world)]

    def send_view(self, dest, rows):
        off = dest * self.region_rows * self.row_words
        return self.cand[off: off + rows * self.row_words]

    def reserve_recv(self, rows):
        if rows > self.recv_rows_cap:
            raise RuntimeError(f"rank {self.rank}: {rows} incoming rows exceed the receive buffer "
                               f"({self.recv_rows_cap} rows); raise cand_bytes")

    def recv_view(self, offset_rows, rows):
        return self.recv[offset_rows * self.row_words: (offset_rows + rows) * self.row_words]

    def insert_received(self, rows):
        self.ck._check(self.lib.kmc_shard_insert(self.ck.ctx, self._recv_ptr, rows, None))

    def insert_local(self, rows):
        self.ck._check(self.lib.kmc_shard_insert(self.ck.ctx, self._cand_ptr, rows, None))

    def reset_cand(self):
        self.ck._check(self.lib.kmc_shard_reset_cand(self.ck.ctx))

    def level_done(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self.ck._check(self.lib.kmc_shard_level_done(self.ck.ctx, ctypes.byref(a), ctypes.byref(b)))
        return int(a.value), int(b.value)

    def finish(self):
        self.ck._check(self.lib.kmc_shard_sync(self.ck.ctx))

    def stats(self):
        return self.ck.stats()

    def violation(self):
        return self.ck.violation()

    def violation_record(self):
        return self.ck.violation_record()

    def state_and_parent(self, idx):
        return self.ck.state_and_parent(idx)

    def describe_state(self, words):
        return self.ck.decoder.text(words)

    def action_name(self, aid):
        acts = self.ck.meta["actions"]
        return acts[aid]["name"] if aid < len(acts) else None

    def trace(self):
        return self.ck.trace()

    def close(self):
        self.ck.close()


@dataclass
class ShardedResult:
    distinct: int
    generated: int
    depth: int
    deadlocks: int
    levels: list[int]
    complete: bool
    violation: dict | None
    per_rank_distinct: list[int]
    seconds: float
    exchanged_rows: int
    stats: dict = field(default_factory=dict)
    trace: list = field(default_factory=list)      # error trace across ranks: [{"words", "action", "rank", "text"}]


class ShardedChecker:
    """Collective driver: every rank constructs one and calls ``run()`` together."""

    def __init__(self, engine: ShardEngine, group=None, cont: bool = False):
        self.e = engine
        self.group = group
        self.cont = cont
        self.world, self.rank = engine.world, engine.rank

    # -- collectives -----------------------------------------------------------
    def _all_gather_counts(self, counts: list[int]) -> list[list[int]]:
        if self.world == 1:
            return [counts]
        t = torch.tensor(counts, dtype=torch.int64, device=self.e.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return [o.tolist() for o in out]

    def _all_reduce(self, vals: list[int], op=dist.ReduceOp.SUM) -> list[int]:
        if self.world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.int64, device=self.e.device)
        dist.all_reduce(t, op=op, group=self.group)
        return t.tolist()

    def _exchange(self, counts, matrix=None) -> int:
        """all-to-all-v of candidate rows; returns the number of rows now in the recv buffer."""
        if matrix is None:
            matrix = self._all_gather_counts(counts)       # matrix[src][dst]
        incoming = [matrix[src][self.rank] for src in range(self.world)]
        self.e.reserve_recv(sum(incoming))
        ops, off = [], 0
        for src in range(self.world):
            n = incoming[src]
            if n:
                view = self.e.recv_view(off, n)
                if src == self.rank:
                    view.copy_(self.e.send_view(self.rank, n))
                else:
                    ops.append(dist.P2POp(dist.irecv, view, src, group=self.group))
            off += n
        for dst in range(self.world):
            n = counts[dst]
            if n and dst != self.rank:
                ops.append(dist.P2POp(dist.isend, self.e.send_view(dst, n), dst, group=self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.exchanged += sum(counts) - counts[self.rank]
        return off

    def _round(self, first: int, count: int, init: bool = False):
        """One expand -> exchange -> insert round on a frontier chunk (count may be 0 on idle ranks)."""
        e = self.e
        if getattr(e, "p2p", False):
            # fused path: rows were stored into the owners' inboxes by the expand kernel itself
            if init:
                e.seed_p2p()
            else:
                e.expand_p2p(first, count)
            e.barrier_on_stream(self.group)
            e.insert_p2p()
            return
        if not init:
            e.reset_cand()
            if count:
                e.expand(first, count)
        if self.world == 1:
            counts = e.counts()
            if counts[0]:
                e.insert_local(counts[0])
            return
        if hasattr(e, "count_matrix"):
            matrix = e.count_matrix(self.group)
            counts = matrix[self.rank]
            # Every rank holds the same matrix, so every rank takes the same decision here: an overflowed owner
            # region (the expand kernel's counters keep counting past region_rows) or an inbox that cannot take
            # what is coming ends the run on ALL ranks before any row is exchanged -- never on one rank only,
            # with the others blocked inside the all-to-all.
            cap, rcap = getattr(e, "region_rows", None), getattr(e, "recv_rows_cap", None)
            if cap is not None and any(x > cap for row in matrix for x in row):
                raise KmcError(-10, "candidate buffer overflow (raise cand_bytes or fanout_bound): an owner region of "
                                    f"{cap} rows was asked to hold {max(x for row in matrix for x in row)}")
            if rcap is not None:
                worst = max(sum(matrix[src][dst] for src in range(self.world)) for dst in range(self.world))
                if worst > rcap:
                    raise KmcError(-10, f"{worst} incoming rows exceed a rank's receive buffer ({rcap} rows); raise cand_bytes")
        else:
            matrix, counts = None, e.counts()
        rows = self._exchange(counts, matrix)
        if rows:
            e.insert_received(rows)

    # -- the search ------------------------------------------------------------
    def _level_summary(self, count: int) -> tuple[int, int, int]:
        """One collective per level: (total new states, any violation, max #chunks over ranks)."""
        e = self.e
        my_chunks = (count + e.chunk_states - 1) // e.chunk_states
        if self.world == 1:
            return count, 1 if e.violation() else 0, my_chunks
        rows = self._all_gather_counts([count, 1 if e.violation() else 0, my_chunks])
        return sum(r[0] for r in rows), max(r[1] for r in rows), max(r[2] for r in rows)

    def run(self) -> ShardedResult:
        stream = getattr(self.e, "stream", None)
        if stream is not None:
            with torch.cuda.stream(stream):
                return self._run()
        return self._run()

    def _run_device_sync(self) -> ShardedResult:
        """The fused path with device-side synchronisation: per level one host synchronisation (the level board)
        and no collective; per round no host wait at all (kmc_shard_round_p2p)."""
        e = self.e
        self.exchanged = 0
        dist.barrier(group=self.group)
        t0 = time.perf_counter()
        e.begin()
        e.round_p2p(0, 0, seed=True)
        board = e.level_sync()
        first, levels, stopped = 0, [], False
        while True:
            count = board[self.rank][1]
            total = sum(b[1] for b in board)
            if any(b[2] for b in board) and not self.cont:
                stopped = True
                break
            if total == 0:
                break
            levels.append(total)
            n_chunks = max((b[1] + e.chunk_states - 1) // e.chunk_states for b in board)
            for c in range(n_chunks):
                off = c * e.chunk_states
                e.round_p2p(first + off, max(0, min(e.chunk_states, count - off)))
            first += count
            board = e.level_sync()
        e.finish()
        st = e.stats()
        rows = [[b[3], b[4], b[6]] for b in board]
        seconds = self._all_reduce_max_float(time.perf_counter() - t0)
        any_viol = any(b[2] for b in board)
        viol, trace = (self._global_violation() if any_viol else (None, []))
        return ShardedResult(distinct=sum(r[0] for r in rows), generated=sum(r[1] for r in rows), depth=len(levels),
                             deadlocks=sum(r[2] for r in rows), levels=levels, complete=not stopped,
                             violation=viol, per_rank_distinct=[r[0] for r in rows], seconds=seconds,
                             # rows that crossed NVLink: no counter on the fused path (the rows leave inside K1's
                             # flush); with the uniform owner hash it is generated * (world - 1) / world
                             exchanged_rows=sum(r[1] for r in rows) * (self.world - 1) // self.world, stats=st, trace=trace)

    def _run(self) -> ShardedResult:
        e = self.e
        if getattr(e, "p2p", False) and getattr(e, "device_sync", True) and hasattr(e, "round_p2p"):
            return self._run_device_sync()
        self.exchanged = 0
        if self.world > 1:
            dist.barrier(group=self.group)
        t0 = time.perf_counter()
        e.begin()
        if not getattr(e, "p2p", False):
            e.seed_init()
        self._round(0, 0, init=True)
        first, count = e.level_done()
        levels: list[int] = []
        stopped = False
        while True:
            total, viol, n_chunks = self._level_summary(count)
            if viol and not self.cont:
                stopped = True
                break
            if total == 0:
                break
            levels.append(total)
            for c in range(n_chunks):
                off = c * e.chunk_states
                n = max(0, min(e.chunk_states, count - off))
                self._round(first + off, n)
            first, count = e.level_done()
        e.finish()
        st = e.stats()
        if self.world > 1:
            rows = self._all_gather_counts([st["distinct"], st["generated"], st["deadlocks"]])
        else:
            rows = [[st["distinct"], st["generated"], st["deadlocks"]]]
        seconds = time.perf_counter() - t0
        if self.world > 1:
            seconds = self._all_reduce_max_float(seconds)
        any_viol = max(self._all_reduce([1 if e.violation() else 0], op=dist.ReduceOp.MAX)) if self.world > 1 \
            else (1 if e.violation() else 0)
        viol, trace = (self._global_violation() if (any_viol and hasattr(e, "violation_record")) else (e.violation(), []))
        return ShardedResult(distinct=sum(r[0] for r in rows), generated=sum(r[1] for r in rows), depth=len(levels),
                             deadlocks=sum(r[2] for r in rows), levels=levels, complete=not stopped,
                             violation=viol, per_rank_distinct=[r[0] for r in rows], seconds=seconds,
                             exchanged_rows=self.exchanged, stats=st, trace=trace)

    # -- error trace across ranks ----------------------------------------------
    NO_PARENT = 0x0000FFFFFFFFFFFF

    def _gather_objects(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def _global_violation(self):
        """All ranks agree on one offending state: deadlocks (they belong to the level being expanded)
        before invariant violations, then the smallest fingerprint -- the same rule as on one GPU --
        and walk its parent links back to an initial state, hopping ranks as the links do."""
        e = self.e
        v = e.violation()
        rec = e.violation_record() if v else None
        mine = None
        if v and rec:
            mine = {"rank": self.rank, "kind": v["kind"], "invariant": v.get("invariant"), "level": v["level"],
                    "fingerprint": v["fingerprint"], "words": rec[0], "meta": rec[1]}
        cands = [c for c in self._gather_objects(mine) if c]
        if not cands:
            return None, []
        best = min(cands, key=lambda c: (0 if c["kind"] == "deadlock" else 1, c["fingerprint"], c["rank"]))
        chain = [(best["words"], best["meta"], best["rank"])]
        meta = best["meta"]
        for _ in range(100000):
            if (meta & self.NO_PARENT) == self.NO_PARENT:
                break
            prank, idx = (meta >> 40) & 0xFF, meta & 0xFFFFFFFFFF
            fetched = e.state_and_parent(idx) if prank == self.rank else None
            got = [g for g in self._gather_objects(fetched) if g]
            words, pmeta = got[0]
            chain.append((words, pmeta, prank))
            meta = pmeta
        chain.reverse()
        trace = []
        for i, (words, m, r) in enumerate(chain):
            aid = (m >> 56) & 0xFF
            entry = {"words": words, "rank": r, "action": None if i == 0 else aid}
            if hasattr(e, "describe_state"):
                entry["text"] = e.describe_state(words)
                entry["action_name"] = None if i == 0 else e.action_name(aid)
            trace.append(entry)
        viol = {k: best[k] for k in ("kind", "invariant", "level", "fingerprint", "rank")}
        viol["trace_len"] = len(trace)
        return viol, trace

    def _all_reduce_max_float(self, x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=self.e.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return float(t.item())
