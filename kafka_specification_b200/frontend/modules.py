"""Module loading, EXTENDS flattening, INSTANCE substitution and LOCAL scoping.

Name resolution rules implemented here (TLA+ semantics; the reference exercises all of them):

* ``EXTENDS`` is textual inclusion: Kip320 -> Kip279 -> KafkaReplication -> {Integers, Util}
  (Kip320.tla:37, Kip279.tla:25, KafkaReplication.tla:30).  Declarations and non-LOCAL
  definitions of the extended modules become visible; a ``LOCAL`` definition is visible only
  to definitions of the module that declares it (``LOCAL Next`` in Kip279.tla:53 does not
  clash with ``Next`` in Kip320.tla:150).
* ``I == INSTANCE M WITH a <- e`` creates a fresh namespace for M in which constant/variable
  ``a`` means ``e`` evaluated in the instantiating context, and every constant/variable of M
  not mentioned in WITH is substituted by the same-named symbol of the instantiating context
  (KafkaReplication.tla:84 instantiates FiniteReplicatedLog with implicit
  ``Replicas/LogRecords/Nil/LogSize`` where ``LogRecords`` and ``Nil`` are *definitions*).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

from .tla_parser import Def, Instance, Module, parse_module_text

STANDARD_MODULES = {"Integers", "Naturals", "FiniteSets", "Sequences", "TLC", "Reals"}


class ModuleError(Exception):
    pass


class Loader:
    """Finds ``Name.tla`` in a list of directories and caches parsed modules."""

    def __init__(self, search_dirs: list[str]):
        self.search_dirs = list(search_dirs)
        self.cache: dict[str, Module] = {}

    def load(self, name: str) -> Module:
        if name in self.cache:
            return self.cache[name]
        for d in self.search_dirs:
            p = os.path.join(d, name + ".tla")
            if os.path.isfile(p):
                with open(p, "r", encoding="utf-8") as f:
                    mod = parse_module_text(f.read(), name)
                if mod.name != name:
                    raise ModuleError(f"{p}: module is named {mod.name}")
                self.cache[name] = mod
                return mod
        raise ModuleError(f"module {name}.tla not found in {self.search_dirs}")

    def extends_closure(self, name: str) -> list[Module]:
        """Modules reachable through EXTENDS, extended-first, root last, no duplicates."""
        out: list[Module] = []
        seen: set[str] = set()

        def visit(n: str):
            if n in seen or n in STANDARD_MODULES:
                return
            seen.add(n)
            m = self.load(n)
            for e in m.extends:
                visit(e)
            out.append(m)

        visit(name)
        return out


@dataclass
class Resolved:
    kind: str          # 'def' | 'inst' | 'const' | 'var' | 'subst'
    name: str
    ctx: "ModuleContext"
    defn: Def | None = None
    expr: tuple | None = None      # for 'subst': expression, evaluated in ctx
    inst: "ModuleContext | None" = None
    from_module: str | None = None  # module whose LOCAL scope applies when evaluating expr


class ModuleContext:
    """One instantiation of a module (the root, or a named INSTANCE of it)."""

    def __init__(self, loader: Loader, module_name: str,
                 substitutions: dict[str, tuple[tuple, "ModuleContext"]] | None = None,
                 parent: "ModuleContext | None" = None, path: str = ""):
        self.loader = loader
        self.module_name = module_name
        self.parent = parent
        self.path = path or module_name
        self.modules = loader.extends_closure(module_name)
        self.module_names = [m.name for m in self.modules]
        self.constants: list[str] = []
        self.variables: list[str] = []
        self.defs: dict[str, list[Def]] = {}
        self.instance_decls: dict[str, tuple[Instance, str]] = {}
        self.assumes: list[tuple[tuple, str]] = []
        for m in self.modules:
            for c in m.constants:
                if c not in self.constants:
                    self.constants.append(c)
            for v in m.variables:
                if v not in self.variables:
                    self.variables.append(v)
            for d in m.defs:
                self.defs.setdefault(d.name, []).append(d)
            for inst in m.instances:
                if inst.name is None:
                    raise ModuleError(f"{m.name}: unnamed INSTANCE is not supported")
                self.instance_decls[inst.name] = (inst, m.name)
            for a in m.assumes:
                self.assumes.append((a, m.name))
        self.substitutions = substitutions  # None for the root
        self._inst_cache: dict[str, ModuleContext] = {}
        self.const_overrides: dict[str, str] = {}   # cfg "C <- Op" (root only)
        self.home_module: str | None = None         # module containing the INSTANCE statement

    # ------------------------------------------------------------------
    def is_root(self) -> bool:
        return self.parent is None

    def find_def(self, name: str, from_module: str | None) -> Def | None:
        cands = self.defs.get(name)
        if not cands:
            return None
        # LOCAL definitions of the asking module win, then non-LOCAL, latest module last.
        if from_module is not None:
            for d in reversed(cands):
                if d.local and d.module == from_module:
                    return d
        pub = [d for d in cands if not d.local]
        if pub:
            return pub[-1]
        # cfg-level lookups (from_module=None) may name a LOCAL definition of the root
        # module itself (``LOCAL Next`` in Kip101.tla:49 / Kip279.tla:53).
        if from_module is None:
            for d in reversed(cands):
                if d.module == self.module_name:
                    return d
        return None

    def resolve(self, name: str, from_module: str | None) -> Resolved | None:
        d = self.find_def(name, from_module)
        if d is not None:
            return Resolved("def", name, self, defn=d)
        if name in self.instance_decls:
            return Resolved("inst", name, self, inst=self.instance(name))
        if name in self.constants or name in self.variables:
            if self.substitutions is None:
                if name in self.const_overrides:
                    return Resolved("subst", name, self, expr=("id", self.const_overrides[name]))
                return Resolved("const" if name in self.constants else "var", name, self)
            if name in self.substitutions:
                expr, ctx = self.substitutions[name]
                return Resolved("subst", name, ctx, expr=expr, from_module=self.home_module)
            # implicit same-name substitution from the instantiating context
            return Resolved("subst", name, self.parent, expr=("id", name), from_module=self.home_module)
        return None

    def instance(self, name: str) -> "ModuleContext":
        if name not in self._inst_cache:
            inst, home = self.instance_decls[name]
            subs = {lhs: (expr, self) for lhs, expr in inst.substitutions}
            ctx = ModuleContext(self.loader, inst.module, subs, self, f"{self.path}!{name}")
            ctx.home_module = home
            self._inst_cache[name] = ctx
        return self._inst_cache[name]

    def root(self) -> "ModuleContext":
        c = self
        while c.parent is not None:
            c = c.parent
        return c


def load_root(module_name: str, search_dirs: list[str]) -> ModuleContext:
    return ModuleContext(Loader(search_dirs), module_name)
