"""ctypes binding of ``libkspecmc.so`` (include/kspecmc.h) -- the host side above the C ABI.

Mirrors the reference-facing interface of the replaced path (TLC's ``ModelChecker``: run a
model, read "states generated / distinct states / depth", fetch the error trace).  There is no
CPU fallback: if the library, the lowered model or the GPU is missing, construction raises.
"""
from __future__ import annotations

import ctypes
import json
import os
from dataclasses import dataclass, field

import numpy as np

from .frontend.cfg import ModelValue
from .frontend.values import FnVal, fmt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")

KMC_ERRORS = {
    0: "KMC_OK", -1: "KMC_E_BADARG", -2: "KMC_E_CUDA", -3: "KMC_E_OOM", -4: "KMC_E_TABLE_FULL",
    -5: "KMC_E_STORE_FULL", -6: "KMC_E_LAYOUT_OVERFLOW", -7: "KMC_E_MODEL", -8: "KMC_E_STATE", -9: "KMC_E_NO_GPU",
    -10: "KMC_E_CAND_FULL", -11: "KMC_E_PEER_TIMEOUT",
}


class KmcError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"{KMC_ERRORS.get(code, code)}: {text}")
        self.code = code


class Stats(ctypes.Structure):
    _fields_ = [
        ("distinct", ctypes.c_uint64), ("generated", ctypes.c_uint64), ("queue", ctypes.c_uint64),
        ("depth", ctypes.c_uint64), ("deadlocks", ctypes.c_uint64), ("out_of_model", ctypes.c_uint64),
        ("probes", ctypes.c_uint64), ("levels", ctypes.c_uint64),
        ("gpu_ms_total", ctypes.c_double), ("gpu_ms_expand", ctypes.c_double), ("gpu_ms_insert", ctypes.c_double),
        ("launches_expand", ctypes.c_uint64), ("launches_insert", ctypes.c_uint64), ("launches_other", ctypes.c_uint64),
        ("wall_ms", ctypes.c_double), ("table_slots", ctypes.c_uint64), ("max_states", ctypes.c_uint64),
        ("complete", ctypes.c_uint64), ("gpu_ms_invariant", ctypes.c_double), ("slot_bytes", ctypes.c_uint64),
    ]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


class Violation(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("invariant", ctypes.c_int32), ("level", ctypes.c_uint64),
                ("trace_len", ctypes.c_uint64), ("fingerprint", ctypes.c_uint64)]


class ModelInfo(ctypes.Structure):
    _fields_ = [("words", ctypes.c_int32), ("state_bits", ctypes.c_int32), ("num_actions", ctypes.c_int32),
                ("num_invariants", ctypes.c_int32), ("num_init", ctypes.c_int32), ("max_fanout", ctypes.c_int32),
                ("check_deadlock", ctypes.c_int32), ("exact", ctypes.c_int32),
                ("name", ctypes.c_char * 128), ("digest", ctypes.c_char * 32)]


class ShardBuffers(ctypes.Structure):
    _fields_ = [("cand", ctypes.c_void_p), ("region_rows", ctypes.c_uint64), ("cand_counts", ctypes.c_void_p),
                ("recv", ctypes.c_void_p), ("recv_rows_cap", ctypes.c_uint64), ("row_words", ctypes.c_int32)]


_LIB = None


def load_library(path: str | None = None) -> ctypes.CDLL:
    """Loads libkspecmc.so; fails loudly when it has not been built."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    path = path or os.path.join(BUILD, "libkspecmc.so")
    if not os.path.exists(path):
        raise KmcError(-7, f"{path} is missing: run `python -m kafka_specification_b200.build --all` "
                           f"(there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    vp, u64p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)
    lib.kmc_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.kmc_destroy.argtypes = [vp]
    lib.kmc_destroy.restype = None
    lib.kmc_model_info.argtypes = [vp, ctypes.POINTER(ModelInfo)]
    lib.kmc_run.argtypes = [vp]
    lib.kmc_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    lib.kmc_level_widths.argtypes = [vp, u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    lib.kmc_action_counts.argtypes = [vp, u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
    lib.kmc_violation.argtypes = [vp, ctypes.POINTER(Violation)]
    lib.kmc_trace_state.argtypes = [vp, ctypes.c_uint32, u64p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_uint32)]
    lib.kmc_copy_states.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, vp]
    lib.kmc_copy_parents.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, vp]
    lib.kmc_violation_record.argtypes = [vp, u64p, ctypes.c_size_t, u64p]
    lib.kmc_strerror.argtypes = [vp, ctypes.c_int]
    lib.kmc_strerror.restype = ctypes.c_char_p
    lib.kmc_fpset_put.argtypes = [vp, vp, ctypes.c_size_t, vp]
    lib.kmc_fpset_contains.argtypes = [vp, vp, ctypes.c_size_t, vp]
    lib.kmc_fpset_size.argtypes = [vp, u64p]
    lib.kmc_shard_begin.argtypes = [vp]
    lib.kmc_shard_buffers.argtypes = [vp, ctypes.POINTER(ShardBuffers)]
    lib.kmc_shard_seed_init.argtypes = [vp]
    lib.kmc_shard_expand.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64]
    lib.kmc_shard_counts.argtypes = [vp, u64p]
    lib.kmc_shard_reset_cand.argtypes = [vp]
    lib.kmc_shard_insert.argtypes = [vp, vp, ctypes.c_uint64, u64p]
    lib.kmc_shard_level_done.argtypes = [vp, u64p, u64p]
    lib.kmc_shard_sync.argtypes = [vp]
    lib.kmc_shard_ipc_handle.argtypes = [vp, vp]
    lib.kmc_shard_open_peers.argtypes = [vp, vp, ctypes.c_uint32]
    lib.kmc_shard_seed_p2p.argtypes = [vp]
    lib.kmc_shard_expand_p2p.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64]
    lib.kmc_shard_insert_p2p.argtypes = [vp]
    lib.kmc_shard_round_p2p.argtypes = [vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int]
    lib.kmc_shard_level_sync.argtypes = [vp, u64p]
    lib.kmc_shard_inbox_ptr.argtypes = [vp, ctypes.POINTER(vp)]
    lib.kmc_shard_open_peers_direct.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(ctypes.c_int), ctypes.c_uint32]
    for fn in ("kmc_create", "kmc_model_info", "kmc_run", "kmc_stats", "kmc_level_widths", "kmc_action_counts",
               "kmc_violation", "kmc_trace_state", "kmc_copy_states", "kmc_copy_parents", "kmc_violation_record", "kmc_fpset_put", "kmc_fpset_contains",
               "kmc_fpset_size", "kmc_shard_begin", "kmc_shard_buffers", "kmc_shard_seed_init", "kmc_shard_expand",
               "kmc_shard_counts", "kmc_shard_reset_cand", "kmc_shard_insert", "kmc_shard_level_done", "kmc_shard_sync",
               "kmc_shard_ipc_handle", "kmc_shard_open_peers", "kmc_shard_seed_p2p", "kmc_shard_expand_p2p",
               "kmc_shard_insert_p2p", "kmc_shard_round_p2p", "kmc_shard_level_sync", "kmc_shard_inbox_ptr",
               "kmc_shard_open_peers_direct"):
        getattr(lib, fn).restype = ctypes.c_int
    _LIB = lib
    return lib


def model_paths(name: str) -> tuple[str, str]:
    d = os.path.join(BUILD, "models", name)
    return os.path.join(d, f"libkmc_{name}.so"), os.path.join(d, "model.json")


# ---------------------------------------------------------------------------
# decoding packed states from model.json (no lowering needed at run time)
# ---------------------------------------------------------------------------
def _parse_atom(text: str):
    if text.startswith('"'):
        return text[1:-1]
    try:
        return int(text)
    except ValueError:
        return ModelValue(text)


def _bits_for(card: int) -> int:
    return max(0, (card - 1).bit_length())


class StateDecoder:
    """Rebuilds TLA+ values from packed words using the layout description in model.json."""

    def __init__(self, meta: dict):
        self.meta = meta
        self.lay = meta["layout"]
        self.atoms = self.lay["atoms"]
        self.variables = self.lay["variables"]

    def _card(self, t) -> int:
        k = t["t"]
        if k == "int":
            return t["hi"] - t["lo"] + 1
        if k == "bool":
            return 2
        if k == "enum":
            return len(t["values"])
        if k == "rec":
            c = 1
            for f in t["fields"].values():
                c *= self._card(f)
            return c
        if k == "fn":
            return self._card(t["elem"]) ** len(t["keys"])
        if k == "tuple":
            c = 1
            for e in t["elems"]:
                c *= self._card(e)
            return c
        if k == "union":
            return sum(self._card(a) for a in t["alts"])
        if k == "set":
            return (1 << self._card(t["elem"])) - (1 if t.get("nonempty") else 0)
        raise ValueError(k)

    def _dec(self, t, code: int):
        k = t["t"]
        if k == "int":
            return code + t["lo"]
        if k == "bool":
            return bool(code)
        if k == "enum":
            return _parse_atom(t["values"][code])
        if k == "rec":
            d = {}
            for f, ft in t["fields"].items():
                c = self._card(ft)
                d[f] = self._dec(ft, code % c)
                code //= c
            return FnVal(d)
        if k == "fn":
            d = {}
            c = self._card(t["elem"])
            for key in t["keys"]:
                d[_parse_atom(key)] = self._dec(t["elem"], code % c)
                code //= c
            return FnVal(d)
        if k == "tuple":
            out = []
            for e in t["elems"]:
                c = self._card(e)
                out.append(self._dec(e, code % c))
                code //= c
            return tuple(out)
        if k == "union":
            off = 0
            for a in t["alts"]:
                c = self._card(a)
                if code < off + c:
                    return self._dec(a, code - off)
                off += c
            raise ValueError("bad union code")
        if k == "set":
            c = self._card(t["elem"])
            if t.get("nonempty"):
                code += 1
            return frozenset(self._dec(t["elem"], j) for j in range(c) if (code >> j) & 1)
        raise ValueError(k)

    def _read(self, t, codes: list[int], pos: list[int]):
        k = t["t"]
        if k == "rec":
            d = {}
            for f, ft in t["fields"].items():
                if ft["t"] == "prefixfn":
                    # entries at index >= the record's length field are the nil value; the others hold the inner code
                    n, vals = d[ft["len"]], {}
                    for key in ft["keys"]:
                        kk = _parse_atom(key)
                        if _bits_for(self._card(ft["inner"])) == 0:
                            x = self._dec(ft["inner"], 0)
                        else:
                            x = self._dec(ft["inner"], codes[pos[0]])
                            pos[0] += 1
                        vals[kk] = x if kk < n else _parse_atom(ft["nil"])
                    d[f] = FnVal(vals)
                else:
                    d[f] = self._read(ft, codes, pos)
            return FnVal(d)
        if k == "fn":
            return FnVal({_parse_atom(key): self._read(t["elem"], codes, pos) for key in t["keys"]})
        if k == "tuple":
            return tuple(self._read(e, codes, pos) for e in t["elems"])
        if k == "seq":
            n = 0
            if t["cap"] > 0:
                n = codes[pos[0]]
                pos[0] += 1
            items = []
            for _ in range(t["cap"]):              # one scalar code per slot (mixed radix for record elements)
                if _bits_for(self._card(t["elem"])) == 0:
                    items.append(self._dec(t["elem"], 0))
                else:
                    items.append(self._dec(t["elem"], codes[pos[0]]))
                    pos[0] += 1
            return tuple(items[:n])
        if k == "set":
            if t["repr"] == "keyed":
                key, fields = t["key"], t["elem"]["fields"]
                rest = {"t": "rec", "fields": {f: ft for f, ft in fields.items() if f != key}}
                out = []
                for j in range(self._card(fields[key])):
                    c = codes[pos[0]]
                    pos[0] += 1
                    if c:
                        d = dict(self._dec(rest, c - 1).items)
                        d[key] = self._dec(fields[key], j)
                        out.append(FnVal({f: d[f] for f in fields}))
                return frozenset(out)
            ecard = self._card(t["elem"])
            if t["repr"] == "bitmap":
                out, base, n = [], 0, ecard
                while n > 0:
                    b = min(32, n)
                    m = codes[pos[0]]
                    pos[0] += 1
                    out += [self._dec(t["elem"], base + j) for j in range(b) if (m >> j) & 1]
                    base += b
                    n -= b
                return frozenset(out)
            cnt = codes[pos[0]]
            pos[0] += 1
            slots = codes[pos[0]: pos[0] + t["cap"]]
            pos[0] += t["cap"]
            return frozenset(self._dec(t["elem"], c) for c in slots[:cnt])
        card = self._card(t)
        if _bits_for(card) == 0:
            return self._dec(t, 0)
        c = codes[pos[0]]
        pos[0] += 1
        return self._dec(t, c)

    def decode(self, words) -> dict:
        codes = [(int(words[a["word"]]) >> a["shift"]) & ((1 << a["bits"]) - 1) for a in self.atoms]
        pos = [0]
        return {v: self._read(self.lay["types"][v], codes, pos) for v in self.variables}

    def text(self, words) -> str:
        st = self.decode(words)
        return "\n".join(f"/\\ {v} = {fmt(st[v])}" for v in self.variables)

    def _spans(self) -> list[tuple[int, int]]:
        """Atom index range [begin, end) each variable reads (static: no type consumes a value-dependent number of
        codes), found by decoding the all-zero code vector once."""
        if getattr(self, "_span_cache", None) is None:
            zeros, pos, spans = [0] * len(self.atoms), [0], []
            for v in self.variables:
                b = pos[0]
                self._read(self.lay["types"][v], zeros, pos)
                spans.append((b, pos[0]))
            self._span_cache = spans
        return self._span_cache

    def texts(self, rows) -> list[str]:
        """TLC-style text of MANY packed states (rows: [n, W] uint64): the atom codes are extracted with numpy, and
        every variable is decoded once per DISTINCT value it takes in the batch (a few thousand for millions of
        states), so that state-set digests of 10^6..10^7 states take seconds instead of minutes."""
        rows = np.ascontiguousarray(rows, dtype=np.uint64).reshape(-1, self.lay["words"])
        n = rows.shape[0]
        if n == 0:
            return []
        codes = np.empty((n, len(self.atoms)), dtype=np.int64)
        for i, a in enumerate(self.atoms):
            codes[:, i] = ((rows[:, a["word"]] >> np.uint64(a["shift"])) & np.uint64((1 << a["bits"]) - 1)).astype(np.int64)
        parts = []
        for v, (b, e) in zip(self.variables, self._spans()):
            ty = self.lay["types"][v]
            if e == b:
                txt = f"/\\ {v} = {fmt(self._read(ty, [], [0]))}"
                parts.append([txt] * n)
                continue
            uniq, inv = np.unique(codes[:, b:e], axis=0, return_inverse=True)
            table = [f"/\\ {v} = {fmt(self._read(ty, [int(x) for x in u], [0]))}" for u in uniq]
            inv = np.asarray(inv).reshape(-1)
            parts.append([table[j] for j in inv])
        return ["\n".join(p) for p in zip(*parts)]


# ---------------------------------------------------------------------------
@dataclass
class RunResult:
    distinct: int
    generated: int
    depth: int
    queue: int
    deadlocks: int
    complete: bool
    levels: list[int]
    stats: dict
    violation: dict | None = None
    trace: list[dict] = field(default_factory=list)


class Checker:
    """One GPU-resident model checker instance (one ``kmc_ctx``)."""

    def __init__(self, model: str, model_lib: str | None = None, model_json: str | None = None, **options):
        self.lib = load_library()
        lib_path, json_path = model_paths(model)
        self.model_lib = model_lib or lib_path
        json_path = model_json or json_path
        if not os.path.exists(self.model_lib):
            raise KmcError(-7, f"lowered model library {self.model_lib} is missing (build it first; no CPU fallback)")
        with open(json_path) as f:
            self.meta = json.load(f)
        self.decoder = StateDecoder(self.meta)
        self.ctx = ctypes.c_void_p()
        opts = {("continue" if k == "cont" else k): v for k, v in options.items() if k not in ("p2p", "device_sync")}
        rc = self.lib.kmc_create(self.model_lib.encode(), json.dumps(opts).encode(), ctypes.byref(self.ctx))
        if rc != 0:
            msg = self.lib.kmc_strerror(self.ctx, rc).decode() if self.ctx else "kmc_create failed"
            if self.ctx:
                self.lib.kmc_destroy(self.ctx)
                self.ctx = None
            raise KmcError(rc, msg)
        self.info = ModelInfo()
        self._check(self.lib.kmc_model_info(self.ctx, ctypes.byref(self.info)))
        if self.info.digest.decode() != self.meta["digest"]:
            raise KmcError(-7, "model.json does not belong to the loaded model library (digest mismatch)")
        self.words = self.info.words

    def _check(self, rc: int):
        if rc != 0:
            raise KmcError(rc, self.lib.kmc_strerror(self.ctx, rc).decode())

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.kmc_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- full BFS ------------------------------------------------------------
    def stats(self) -> dict:
        st = Stats()
        self._check(self.lib.kmc_stats(self.ctx, ctypes.byref(st)))
        return st.as_dict()

    def level_widths(self) -> list[int]:
        buf = (ctypes.c_uint64 * 4096)()
        n = ctypes.c_size_t()
        self._check(self.lib.kmc_level_widths(self.ctx, buf, 4096, ctypes.byref(n)))
        return [int(buf[i]) for i in range(min(n.value, 4096))]

    def action_counts(self) -> dict:
        buf = (ctypes.c_uint64 * 64)()
        n = ctypes.c_size_t()
        self._check(self.lib.kmc_action_counts(self.ctx, buf, 64, ctypes.byref(n)))
        return {a["name"]: int(buf[i]) for i, a in enumerate(self.meta["actions"]) if i < n.value}

    def violation(self) -> dict | None:
        v = Violation()
        self._check(self.lib.kmc_violation(self.ctx, ctypes.byref(v)))
        if v.kind == 0:
            return None
        return {"kind": "invariant" if v.kind == 1 else "deadlock",
                "invariant": self.meta["invariants"][v.invariant] if v.kind == 1 else None,
                "level": int(v.level), "trace_len": int(v.trace_len), "fingerprint": int(v.fingerprint)}

    def trace(self) -> list[dict]:
        v = self.violation()
        if v is None:
            return []
        out = []
        buf = (ctypes.c_uint64 * self.words)()
        act = ctypes.c_uint32()
        for i in range(v["trace_len"]):
            self._check(self.lib.kmc_trace_state(self.ctx, i, buf, self.words, ctypes.byref(act)))
            words = [int(buf[k]) for k in range(self.words)]
            a = self.meta["actions"][act.value] if (i > 0 and act.value < len(self.meta["actions"])) else None
            out.append({"words": words, "action": a, "state": self.decoder.decode(words),
                        "text": self.decoder.text(words)})
        return out

    def run(self, raise_on_error: bool = True) -> RunResult:
        self.last_rc = self.lib.kmc_run(self.ctx)
        if self.last_rc != 0 and raise_on_error:
            self._check(self.last_rc)
        return self.result()

    def error_text(self, rc: int) -> str:
        return self.lib.kmc_strerror(self.ctx, rc).decode()

    def result(self) -> RunResult:
        st = self.stats()
        viol = self.violation()
        return RunResult(distinct=st["distinct"], generated=st["generated"], depth=st["depth"], queue=st["queue"],
                         deadlocks=st["deadlocks"], complete=bool(st["complete"]), levels=self.level_widths(),
                         stats=st, violation=viol, trace=self.trace() if viol else [])

    def violation_record(self):
        """(packed words, parent word) of this rank's offending state, or None."""
        buf = (ctypes.c_uint64 * self.words)()
        meta = ctypes.c_uint64()
        rc = self.lib.kmc_violation_record(self.ctx, buf, self.words, ctypes.byref(meta))
        if rc != 0:
            return None
        return [int(buf[k]) for k in range(self.words)], int(meta.value)

    def state_and_parent(self, idx: int):
        st = self.copy_states(idx, 1)[0]
        par = np.empty(1, dtype=np.uint64)
        self._check(self.lib.kmc_copy_parents(self.ctx, idx, 1, par.ctypes.data))
        return [int(x) for x in st], int(par[0])

    def copy_states(self, first: int, count: int) -> np.ndarray:
        buf = np.empty((count, self.words), dtype=np.uint64)
        if count:
            self._check(self.lib.kmc_copy_states(self.ctx, first, count, buf.ctypes.data))
        return buf

    # -- fingerprint set alone (FPSet.put / contains / size) -----------------
    def fpset_put(self, fps: np.ndarray) -> np.ndarray:
        fps = np.ascontiguousarray(fps, dtype=np.uint64)
        seen = np.zeros(len(fps), dtype=np.uint8)
        self._check(self.lib.kmc_fpset_put(self.ctx, fps.ctypes.data, len(fps), seen.ctypes.data))
        return seen.astype(bool)

    def fpset_contains(self, fps: np.ndarray) -> np.ndarray:
        fps = np.ascontiguousarray(fps, dtype=np.uint64)
        out = np.zeros(len(fps), dtype=np.uint8)
        self._check(self.lib.kmc_fpset_contains(self.ctx, fps.ctypes.data, len(fps), out.ctypes.data))
        return out.astype(bool)

    def fpset_size(self) -> int:
        n = ctypes.c_uint64()
        self._check(self.lib.kmc_fpset_size(self.ctx, ctypes.byref(n)))
        return int(n.value)
