"""CPU stand-in for one rank of the sharded engine (tests only), see host_shard.cpp."""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build(name: str) -> ctypes.CDLL:
    hdr = os.path.join(ROOT, "build", "models", name, "model.h")
    tag = hashlib.sha256(open(hdr, "rb").read() + open(os.path.join(HERE, "host_shard.cpp"), "rb").read()).hexdigest()[:12]
    out = os.path.join(ROOT, "build", "hosttest")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, f"shard_{name}_{tag}.so")
    if not os.path.exists(so):
        tmp = so + f".{os.getpid()}.tmp"
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f'-DKMC_MODEL_HEADER="{hdr}"',
                               os.path.join(HERE, "host_shard.cpp"), "-o", tmp])
        os.replace(tmp, so)
    lib = ctypes.CDLL(so)
    lib.hs_create.restype = ctypes.c_void_p
    lib.hs_create.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
    for fn in ("hs_destroy", "hs_begin", "hs_seed_init", "hs_reset_cand"):
        getattr(lib, fn).argtypes = [ctypes.c_void_p]
        getattr(lib, fn).restype = None
    lib.hs_expand.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64]
    lib.hs_expand.restype = None
    lib.hs_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.hs_counts.restype = None
    lib.hs_send_ptr.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    lib.hs_send_ptr.restype = ctypes.POINTER(ctypes.c_int64)
    lib.hs_recv_ptr.argtypes = [ctypes.c_void_p, ctypes.c_uint64]
    lib.hs_recv_ptr.restype = ctypes.POINTER(ctypes.c_int64)
    lib.hs_insert.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    lib.hs_insert.restype = None
    lib.hs_level_done.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.hs_level_done.restype = None
    lib.hs_stats.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.hs_stats.restype = None
    return lib


class HostShardEngine:
    def __init__(self, name: str, rank: int, world: int, chunk_states: int = 1000):
        self.lib = build(name)
        self.rank, self.world = rank, world
        self.h = self.lib.hs_create(rank, world)
        self.row_words = self.lib.hs_row_words()
        self.chunk_states = chunk_states
        self.device = torch.device("cpu")
        self._recv_rows = 0
        self._recv = None
        self._incoming_total = 0

    def begin(self):
        self.lib.hs_begin(self.h)

    def seed_init(self):
        self.lib.hs_seed_init(self.h)

    def expand(self, first, count):
        self.lib.hs_expand(self.h, first, count)

    def counts(self):
        buf = np.zeros(8, dtype=np.uint64)
        self.lib.hs_counts(self.h, buf.ctypes.data)
        self._recv = None
        return [int(buf[d]) for d in range(self.world)]

    def _view(self, ptr, nwords):
        if nwords == 0:
            return torch.empty(0, dtype=torch.int64)
        return torch.from_numpy(np.ctypeslib.as_array(ptr, shape=(nwords,)))

    def send_view(self, dest, rows):
        return self._view(self.lib.hs_send_ptr(self.h, dest), rows * self.row_words)

    def reserve_recv(self, rows):
        if self._recv is None or self._recv.numel() < rows * self.row_words:
            self._recv = torch.zeros(max(rows, 1024) * self.row_words, dtype=torch.int64)

    def recv_view(self, offset_rows, rows):
        return self._recv[offset_rows * self.row_words: (offset_rows + rows) * self.row_words]

    def insert_received(self, rows):
        buf = self._recv[: rows * self.row_words].contiguous().numpy()
        self.lib.hs_insert(self.h, buf.ctypes.data, rows)

    def insert_local(self, rows):
        ptr = self.lib.hs_send_ptr(self.h, 0)
        self.lib.hs_insert(self.h, ctypes.cast(ptr, ctypes.c_void_p), rows)

    def reset_cand(self):
        self.lib.hs_reset_cand(self.h)

    def level_done(self):
        a, b = ctypes.c_uint64(), ctypes.c_uint64()
        self.lib.hs_level_done(self.h, ctypes.byref(a), ctypes.byref(b))
        return int(a.value), int(b.value)

    def finish(self):
        pass

    def _raw(self):
        st = np.zeros(8, dtype=np.uint64)
        self.lib.hs_stats(self.h, st.ctypes.data)
        return st

    def stats(self):
        st = self._raw()
        return {"distinct": int(st[0]), "generated": int(st[1]), "deadlocks": int(st[2]), "fail": int(st[3])}

    def _record(self):
        out = np.zeros(self.row_words + 4, dtype=np.uint64)
        self.lib.hs_violation_record.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        inv = self.lib.hs_violation_record(self.h, out.ctypes.data)
        return inv, out

    def violation(self):
        inv, out = self._record()
        if inv < 0:
            return None
        W = self.row_words - 1
        return {"kind": "invariant", "invariant": inv, "level": int(out[W + 2]), "fingerprint": int(out[W + 1])}

    def violation_record(self):
        inv, out = self._record()
        if inv < 0:
            return None
        W = self.row_words - 1
        return [int(x) for x in out[:W]], int(out[W])

    def state_and_parent(self, idx):
        out = np.zeros(self.row_words, dtype=np.uint64)
        self.lib.hs_state_and_parent.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        self.lib.hs_state_and_parent(self.h, idx, out.ctypes.data)
        return [int(x) for x in out[:-1]], int(out[-1])

    def is_successor(self, a, b) -> int:
        self.lib.hs_is_successor.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        x, y = np.array(a, dtype=np.uint64), np.array(b, dtype=np.uint64)
        return self.lib.hs_is_successor(x.ctypes.data, y.ctypes.data)

    def is_init(self, a) -> int:
        self.lib.hs_is_init.argtypes = [ctypes.c_void_p]
        x = np.array(a, dtype=np.uint64)
        return self.lib.hs_is_init(x.ctypes.data)
