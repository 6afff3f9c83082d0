"""ctypes wrapper of Oracle B (oracle/kspec_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = {"idsequence": 0, "frl": 1, "trunchw": 2, "kip101": 3, "kip279": 4, "kip320": 5, "firsttry": 6, "asyncisr": 7, "kip320_279": 8}
MODULE_TO_MODEL = {"IdSequence": "idsequence", "FiniteReplicatedLog": "frl", "KafkaTruncateToHighWatermark": "trunchw",
                   "Kip101": "kip101", "Kip279": "kip279", "Kip320": "kip320", "Kip320FirstTry": "firsttry",
                   "AsyncIsr": "asyncisr", "MCAsyncIsr": "asyncisr", "MCKip320With279": "kip320_279"}
INVARIANTS = ["WeakIsr", "StrongIsr", "LeaderInIsr", "ValidHighWatermark"]


class Result(ctypes.Structure):
    _fields_ = [("distinct", ctypes.c_uint64), ("generated", ctypes.c_uint64), ("depth", ctypes.c_uint64),
                ("deadlocks", ctypes.c_uint64), ("out_of_model", ctypes.c_uint64), ("complete", ctypes.c_uint64),
                ("levels", ctypes.c_uint64 * 256), ("first_violation_level", ctypes.c_uint64 * 4),
                ("violating_states", ctypes.c_uint64 * 4), ("seconds", ctypes.c_double),
                ("state_size", ctypes.c_uint64)]


def build() -> str:
    so = os.path.join(HERE, "_build", "libkspec_oracle.so")
    src = os.path.join(HERE, "kspec_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", HERE])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.kso_run.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_uint64,
                                 ctypes.c_uint, ctypes.POINTER(Result), ctypes.c_void_p, ctypes.c_uint64]
        _lib.kso_run.restype = ctypes.c_int
        _lib.kso_run_sym.argtypes = _lib.kso_run.argtypes + [ctypes.c_int]
        _lib.kso_run_sym.restype = ctypes.c_int
        _lib.kso_state_size.argtypes = [ctypes.c_int]
        _lib.kso_state_size.restype = ctypes.c_size_t
        _lib.kso_successors.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib.kso_init_state.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]
        _lib.kso_violated.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_void_p, ctypes.c_uint]
        _lib.kso_violated.restype = ctypes.c_uint
    return _lib


def run(model: str, params: list[int], threads: int = 0, max_states: int = 0, invariants: list[str] | None = None,
        dump: bool = False, symmetry: bool = False) -> dict:
    """params: Kafka family [n, L, R, E]; frl [n, L, R]; idsequence [MaxId]; asyncisr [n, MaxOffset, MaxVersion]."""
    threads = threads or os.cpu_count() or 1
    mask = 0
    for inv in invariants or []:
        if inv in INVARIANTS:
            mask |= 1 << INVARIANTS.index(inv)
    arr = (ctypes.c_int * 4)(*(list(params) + [0] * (4 - len(params))))
    res = Result()
    buf = None
    cap = 0
    if dump:
        cap = max_states or (1 << 22)
        buf = np.zeros(cap * lib().kso_state_size(MODELS[model]), dtype=np.uint8)
    rc = lib().kso_run_sym(MODELS[model], arr, threads, max_states, mask, ctypes.byref(res),
                           buf.ctypes.data if dump else None, cap, 1 if symmetry else 0)
    if rc != 0:
        raise RuntimeError(f"kso_run failed: {rc}")
    depth = int(res.depth)
    out = {
        "distinct": int(res.distinct), "generated": int(res.generated), "depth": depth,
        "deadlocks": int(res.deadlocks), "out_of_model": int(res.out_of_model), "complete": bool(res.complete),
        "levels": [int(res.levels[i]) for i in range(min(depth, 256))],
        "first_violation_level": {INVARIANTS[i]: (int(res.first_violation_level[i]) or None) for i in range(4)
                                  if (mask >> i) & 1},
        "violating_states": {INVARIANTS[i]: int(res.violating_states[i]) for i in range(4) if (mask >> i) & 1},
        "seconds": float(res.seconds), "threads": threads,
    }
    if dump:
        out["records"] = buf.reshape(cap, -1)[: out["distinct"]]
    return out


# ---- single states (error-trace validation) ---------------------------------------------------------------
NMAX, LMAX, EMAX, NONE = 5, 6, 7, 255


class Req(ctypes.Structure):
    _fields_ = [("epoch", ctypes.c_int8), ("leader", ctypes.c_uint8), ("isr", ctypes.c_uint8)]


class KState(ctypes.Structure):
    """Mirror of kspec_oracle.c's KState (Kafka family)."""
    _fields_ = [("end", ctypes.c_uint8 * NMAX), ("rec_id", (ctypes.c_uint8 * LMAX) * NMAX),
                ("rec_ep", (ctypes.c_uint8 * LMAX) * NMAX), ("hw", ctypes.c_uint8 * NMAX),
                ("rs_epoch", ctypes.c_int8 * NMAX), ("rs_leader", ctypes.c_uint8 * NMAX),
                ("rs_isr", ctypes.c_uint8 * NMAX), ("next_record_id", ctypes.c_uint8),
                ("next_leader_epoch", ctypes.c_uint8), ("q_epoch", ctypes.c_int8), ("q_leader", ctypes.c_uint8),
                ("q_isr", ctypes.c_uint8), ("nreq", ctypes.c_uint8), ("req", Req * (EMAX + 1))]


def kstate_from_tla(state: dict, replicas: list) -> bytes:
    """Oracle B's canonical record of a decoded TLA+ state of the Kafka family (values as produced by
    kafka_specification_b200.runtime.StateDecoder: FnVal / frozenset / int / model values); `replicas` fixes the
    replica numbering (sorted model-value names, the order Oracle B uses)."""
    assert ctypes.sizeof(KState) == lib().kso_state_size(MODELS["kip320"])
    idx = {r: i for i, r in enumerate(replicas)}

    def rid(x):
        return NONE if x == "NONE" else idx[x]

    def bits(s):
        return sum(1 << idx[r] for r in s)

    k = KState()
    for r, i in idx.items():
        log = state["replicaLog"].apply(r)
        end = log.apply("endOffset")
        k.end[i] = end
        recs = log.apply("records")
        for o in range(end):
            rec = recs.apply(o)
            k.rec_id[i][o] = rec.apply("id")
            k.rec_ep[i][o] = rec.apply("epoch")
        rs = state["replicaState"].apply(r)
        k.hw[i] = rs.apply("hw")
        k.rs_epoch[i] = rs.apply("leaderEpoch")
        k.rs_leader[i] = rid(rs.apply("leader"))
        k.rs_isr[i] = bits(rs.apply("isr"))
    k.next_record_id = state["nextRecordId"]
    k.next_leader_epoch = state["nextLeaderEpoch"]
    q = state["quorumState"]
    k.q_epoch, k.q_leader, k.q_isr = q.apply("leaderEpoch"), rid(q.apply("leader")), bits(q.apply("isr"))
    reqs = sorted(((x.apply("leaderEpoch"), rid(x.apply("leader")), bits(x.apply("isr"))) for x in state["leaderAndIsrRequests"]))
    k.nreq = len(reqs)
    for j, (e, l, s) in enumerate(reqs):
        k.req[j].epoch, k.req[j].leader, k.req[j].isr = e, l, s
    return bytes(k)


def _params(params):
    return (ctypes.c_int * 4)(*(list(params) + [0] * (4 - len(params))))


def successors(model: str, params: list[int], state: bytes) -> list[bytes]:
    size = lib().kso_state_size(MODELS[model])
    buf = ctypes.create_string_buffer(256 * size)
    src = ctypes.create_string_buffer(state, size)
    n = lib().kso_successors(MODELS[model], _params(params), src, buf, 256)
    if n < 0:
        raise RuntimeError(f"kso_successors failed: {n}")
    return [buf.raw[i * size:(i + 1) * size] for i in range(n)]


def init_state(model: str, params: list[int]) -> bytes:
    size = lib().kso_state_size(MODELS[model])
    buf = ctypes.create_string_buffer(size)
    if lib().kso_init_state(MODELS[model], _params(params), buf) != 0:
        raise RuntimeError("kso_init_state failed")
    return buf.raw


def violated(model: str, params: list[int], state: bytes, invariants: list[str]) -> list[str]:
    mask = sum(1 << INVARIANTS.index(i) for i in invariants if i in INVARIANTS)
    size = lib().kso_state_size(MODELS[model])
    v = lib().kso_violated(MODELS[model], _params(params), ctypes.create_string_buffer(state, size), mask)
    return [INVARIANTS[i] for i in range(4) if (v >> i) & 1]


if __name__ == "__main__":
    import json
    import sys
    m = sys.argv[1]
    p = [int(x) for x in sys.argv[2].split(",")]
    invs = sys.argv[3].split(",") if len(sys.argv) > 3 else []
    r = run(m, p, max_states=int(sys.argv[4]) if len(sys.argv) > 4 else 0, invariants=invs)
    print(json.dumps(r))
