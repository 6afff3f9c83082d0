"""ctypes wrapper of Oracle B (oracle/kspec_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = {"idsequence": 0, "frl": 1, "trunchw": 2, "kip101": 3, "kip279": 4, "kip320": 5, "firsttry": 6, "asyncisr": 7}
MODULE_TO_MODEL = {"IdSequence": "idsequence", "FiniteReplicatedLog": "frl", "KafkaTruncateToHighWatermark": "trunchw",
                   "Kip101": "kip101", "Kip279": "kip279", "Kip320": "kip320", "Kip320FirstTry": "firsttry",
                   "AsyncIsr": "asyncisr", "MCAsyncIsr": "asyncisr"}
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


if __name__ == "__main__":
    import json
    import sys
    m = sys.argv[1]
    p = [int(x) for x in sys.argv[2].split(",")]
    invs = sys.argv[3].split(",") if len(sys.argv) > 3 else []
    r = run(m, p, max_states=int(sys.argv[4]) if len(sys.argv) > 4 else 0, invariants=invs)
    print(json.dumps(r))
