"""TLA+ value representation shared by the front-end consumers.

Sets are ``frozenset``; functions and records are :class:`FnVal` (a record is a function whose
domain is a set of strings); model values are :class:`ModelValue`; tuples are Python tuples.
``fmt`` prints values the way TLC does and doubles as the canonical serialisation used to
compare state sets between independent implementations.
"""
from __future__ import annotations

from .cfg import ModelValue


class EvalError(Exception):
    pass


def sort_key(v):
    if isinstance(v, bool):
        return (0, int(v))
    if isinstance(v, int):
        return (1, v)
    if isinstance(v, str):
        return (2, v)
    if isinstance(v, ModelValue):
        return (3, v.name)
    if isinstance(v, FnVal):
        return (4, tuple((sort_key(k), sort_key(x)) for k, x in v.items))
    if isinstance(v, frozenset):
        return (5, tuple(sorted(sort_key(x) for x in v)))
    if isinstance(v, tuple):
        return (6, tuple(sort_key(x) for x in v))
    raise EvalError(f"unsortable value {v!r}")


class FnVal:
    """Immutable TLA+ function (records are functions with string domain)."""
    __slots__ = ("items", "_d", "_h")

    def __init__(self, mapping: dict):
        self._d = dict(mapping)
        self.items = tuple(sorted(self._d.items(), key=lambda kv: sort_key(kv[0])))
        self._h = hash(self.items)

    def __hash__(self):
        return self._h

    def __eq__(self, other):
        return isinstance(other, FnVal) and self._h == other._h and self.items == other.items

    def __ne__(self, other):
        return not self.__eq__(other)

    def apply(self, k):
        try:
            return self._d[k]
        except KeyError:
            raise EvalError(f"function applied outside its domain: {k!r} not in {list(self._d)}")

    def domain(self):
        return frozenset(self._d)

    def updated(self, k, v):
        if k not in self._d:
            raise EvalError(f"EXCEPT on key {k!r} outside domain")
        d = dict(self._d)
        d[k] = v
        return FnVal(d)

    def __repr__(self):
        if self._d and all(isinstance(k, str) for k in self._d):
            return "[" + ", ".join(f"{k} |-> {fmt(v)}" for k, v in self.items) + "]"
        return "(" + " @@ ".join(f"{fmt(k)} :> {fmt(v)}" for k, v in self.items) + ")"


def fmt(v) -> str:
    """TLC-style value printing (also the canonical serialisation used for state digests)."""
    if isinstance(v, bool):
        return "TRUE" if v else "FALSE"
    if isinstance(v, int):
        return str(v)
    if isinstance(v, str):
        return '"' + v + '"'
    if isinstance(v, ModelValue):
        return v.name
    if isinstance(v, frozenset):
        return "{" + ", ".join(fmt(x) for x in sorted(v, key=sort_key)) + "}"
    if isinstance(v, tuple):
        return "<<" + ", ".join(fmt(x) for x in v) + ">>"
    return repr(v)


