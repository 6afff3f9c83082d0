"""Test support: build a lowered model for the host and run the sequential BFS harness."""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from kafka_specification_b200.lower.model import LoweredModel, lower_model  # noqa: E402

BUILD = os.path.join(ROOT, "build", "hosttest")


def build_host(model: LoweredModel) -> ctypes.CDLL:
    os.makedirs(BUILD, exist_ok=True)
    tag = hashlib.sha256((model.header + open(os.path.join(HERE, "host_bfs.cpp")).read()).encode()).hexdigest()[:16]
    hdr = os.path.join(BUILD, f"{model.name}_{tag}.h")
    so = os.path.join(BUILD, f"{model.name}_{tag}.so")
    if not os.path.exists(so):
        with open(hdr, "w") as f:
            f.write(model.header)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f'-DKMC_MODEL_HEADER="{hdr}"',
                               os.path.join(HERE, "host_bfs.cpp"), "-o", so])
    lib = ctypes.CDLL(so)
    lib.kmc_host_bfs.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_uint64]
    return lib


def run_host(model: LoweredModel, max_states: int = 0, dump: bool = False, items: bool = False) -> dict:
    lib = build_host(model)
    lib.kmc_host_use_items(1 if items else 0)
    st = np.zeros(320, dtype=np.uint64)
    buf = None
    cap = 0
    if dump:
        cap = max_states or 2_000_000
        buf = np.zeros((cap, model.words), dtype=np.uint64)
    lib.kmc_host_bfs(st.ctypes.data, max_states, buf.ctypes.data if dump else None, cap)
    ninv = len(model.invariants)
    depth = int(st[2])
    res = {
        "distinct": int(st[0]), "generated": int(st[1]), "depth": depth, "deadlocks": int(st[3]),
        "fail": int(st[4]), "complete": bool(st[5]),
        "first_violated": None if st[6] == np.uint64(2**64 - 1) else model.invariants[int(st[6])],
        "first_violated_level": int(st[7]), "max_fanout_seen": int(st[8]),
        "first_violation_level": {model.invariants[i]: (int(st[16 + i]) or None) for i in range(ninv)},
        "levels": [int(x) for x in st[64:64 + min(depth, 192)]],
        "per_action": {a["name"]: int(st[256 + i]) for i, a in enumerate(model.actions)},
    }
    if dump:
        res["states"] = buf[: res["distinct"]]
    return res
