"""Recursive-descent parser for the TLA+ subset of SURVEY.md Appendix A.

Input surface: the reference's ``.tla`` files, unchanged (IdSequence.tla, Util.tla,
FiniteReplicatedLog.tla, KafkaReplication.tla and its five variants, AsyncIsr.tla)
plus the small ``MC*.tla`` wrapper modules under ``models/``.

AST = plain tuples, first element the node kind:

  ('num', n) ('str', s) ('bool', b) ('id', name) ('at',)
  ('app', name, [args])                 user/builtin operator application  Op(a, b)
  ('inst', inst_name, op_name, [args])  I!Op(args)   (args may be empty)
  ('binop', op, a, b)   op in = # < > <= >= + - * \\div \\in \\notin \\subseteq \\union \\intersect \\ .. => <=>
  ('neg', a) ('not', a)
  ('and', [items]) ('or', [items])      junction lists and infix chains, flattened
  ('quant', 'E'|'A', [([names], set)], body)
  ('choose', name, set, body)
  ('if', c, t, e) ('let', [Def], body) ('case', [(guard, e)], other|None)
  ('setenum', [items]) ('setmap', expr, [([names], set)]) ('setfilter', name, set, pred)
  ('subset', e) ('union_all', e) ('domain', e)
  ('fnlit', [([names], set)], body) ('fnapp', f, [args]) ('fnset', S, T)
  ('rec', [(field, e)]) ('recset', [(field, S)]) ('dot', e, field)
  ('except', f, [(path, e)])            path = [('idx', e) | ('fld', name)]
  ('tuple', [items]) ('prime', e) ('unchanged', e) ('enabled', e)
  temporal (parsed, never evaluated): ('box', e) ('diamond', e) ('actionbox', a, v) ('fair', 'SF'|'WF', v, a)
"""
from __future__ import annotations

from dataclasses import dataclass, field

from .tla_lexer import Token, TlaSyntaxError, tokenize


@dataclass
class Def:
    name: str
    params: list[str]
    body: tuple
    local: bool = False
    module: str = ""
    line: int = 0
    col: int = 0
    end_line: int = 0
    end_col: int = 0


@dataclass
class Instance:
    name: str | None            # None for an unnamed INSTANCE (not used by the reference)
    module: str
    substitutions: list[tuple[str, tuple]]
    local: bool = False


@dataclass
class Module:
    name: str
    extends: list[str] = field(default_factory=list)
    constants: list[str] = field(default_factory=list)
    variables: list[str] = field(default_factory=list)
    defs: list[Def] = field(default_factory=list)
    instances: list[Instance] = field(default_factory=list)
    assumes: list[tuple] = field(default_factory=list)
    theorems: list[tuple] = field(default_factory=list)


# binary operator precedence (low, high) ranges follow the TLA+ book table; we only
# need a consistent total order for the operators that occur.
_BINOPS = {
    "=>": (1, "R"), "<=>": (2, "N"),
    "\\/": (3, "L"), "/\\": (3, "L"),
    "=": (5, "N"), "#": (5, "N"), "/=": (5, "N"), "<": (5, "N"), ">": (5, "N"),
    "<=": (5, "N"), "=<": (5, "N"), ">=": (5, "N"),
    "\\in": (5, "N"), "\\notin": (5, "N"), "\\subseteq": (5, "N"),
    "\\union": (8, "L"), "\\intersect": (8, "L"), "\\": (8, "L"),
    "..": (9, "N"),
    "+": (10, "L"), "-": (11, "L"), "%": (11, "L"),
    "\\X": (12, "L"),                      # n-ary: A \X B \X C is the set of triples
    "*": (13, "L"), "\\div": (13, "L"), "\\o": (13, "L"),
}

_UNIT_STARTERS = {"CONSTANT", "CONSTANTS", "VARIABLE", "VARIABLES", "ASSUME", "ASSUMPTION",
                  "AXIOM", "THEOREM", "LEMMA", "INSTANCE", "LOCAL", "EXTENDS", "RECURSIVE"}


class Parser:
    def __init__(self, toks: list[Token], module_hint: str = "?"):
        self.toks = toks
        self.i = 0
        self.barriers: list[int] = []   # columns of enclosing junction lists
        self.module_hint = module_hint

    # -- token helpers ---------------------------------------------------
    @property
    def tok(self) -> Token:
        return self.toks[self.i]

    def peek(self, k: int = 1) -> Token:
        j = min(self.i + k, len(self.toks) - 1)
        return self.toks[j]

    def advance(self) -> Token:
        t = self.toks[self.i]
        self.i += 1
        return t

    def error(self, msg: str):
        raise TlaSyntaxError(f"{self.module_hint}: {msg} at {self.tok!r}")

    def is_op(self, text: str) -> bool:
        return self.tok.kind == "op" and self.tok.text == text

    def is_kw(self, text: str) -> bool:
        return self.tok.kind == "kw" and self.tok.text == text

    def expect_op(self, text: str) -> Token:
        if not self.is_op(text):
            self.error(f"expected {text!r}")
        return self.advance()

    def expect_kw(self, text: str) -> Token:
        if not self.is_kw(text):
            self.error(f"expected {text}")
        return self.advance()

    def expect_id(self) -> str:
        if self.tok.kind != "id":
            self.error("expected identifier")
        return self.advance().text

    def blocked(self) -> bool:
        """True if the current token cannot continue the expression being parsed:
        it sits at or left of the innermost junction-list column."""
        t = self.tok
        if t.kind in ("eof", "end", "sep"):
            return True
        return bool(self.barriers) and t.col <= self.barriers[-1]

    # -- module level ----------------------------------------------------
    def parse_module(self) -> Module:
        if self.tok.kind != "sep":
            self.error("expected ---- MODULE header")
        self.advance()
        self.expect_kw("MODULE")
        mod = Module(self.expect_id())
        self.module_hint = mod.name
        if self.tok.kind != "sep":
            self.error("expected ---- after module name")
        self.advance()
        while True:
            t = self.tok
            if t.kind == "end" or t.kind == "eof":
                break
            if t.kind == "sep":
                self.advance()
                continue
            self.parse_unit(mod)
        return mod

    def parse_unit(self, mod: Module):
        local = False
        if self.is_kw("LOCAL"):
            self.advance()
            local = True
        t = self.tok
        if t.kind == "kw" and t.text == "EXTENDS":
            self.advance()
            mod.extends.append(self.expect_id())
            while self.is_op(","):
                self.advance()
                mod.extends.append(self.expect_id())
        elif t.kind == "kw" and t.text in ("CONSTANT", "CONSTANTS"):
            self.advance()
            mod.constants.extend(self.parse_decl_names())
        elif t.kind == "kw" and t.text in ("VARIABLE", "VARIABLES"):
            self.advance()
            mod.variables.extend(self.parse_decl_names())
        elif t.kind == "kw" and t.text in ("ASSUME", "ASSUMPTION", "AXIOM"):
            self.advance()
            mod.assumes.append(self.parse_expr())
        elif t.kind == "kw" and t.text in ("THEOREM", "LEMMA"):
            self.advance()
            mod.theorems.append(self.parse_expr())
        elif t.kind == "kw" and t.text == "INSTANCE":
            mod.instances.append(self.parse_instance(None, local))
        elif t.kind == "kw" and t.text == "RECURSIVE":
            # RECURSIVE Op(_, _), Op2(_): forward declarations; definitions are looked up by name when called,
            # so nothing needs recording
            self.advance()
            while True:
                self.expect_id()
                if self.is_op("("):
                    self.advance()
                    while not self.is_op(")"):
                        self.advance()
                    self.expect_op(")")
                if self.is_op(","):
                    self.advance()
                    continue
                break
        elif t.kind == "id":
            self.parse_definition(mod, local)
        else:
            self.error("unexpected token at module level")

    def parse_decl_names(self) -> list[str]:
        names = [self.expect_id()]
        while self.is_op(","):
            self.advance()
            names.append(self.expect_id())
        return names

    def parse_instance(self, name: str | None, local: bool) -> Instance:
        self.expect_kw("INSTANCE")
        module = self.expect_id()
        subs: list[tuple[str, tuple]] = []
        if self.is_kw("WITH"):
            self.advance()
            while True:
                lhs = self.expect_id()
                self.expect_op("<-")
                subs.append((lhs, self.parse_expr()))
                if self.is_op(","):
                    self.advance()
                    continue
                break
        return Instance(name, module, subs, local)

    def parse_definition(self, mod: Module, local: bool):
        start = self.tok
        name = self.expect_id()
        params: list[str] = []
        if self.is_op("("):
            self.advance()
            params.append(self.expect_id())
            while self.is_op(","):
                self.advance()
                params.append(self.expect_id())
            self.expect_op(")")
        self.expect_op("==")
        if self.is_kw("INSTANCE"):
            if params:
                self.error("parametrised INSTANCE not supported")
            mod.instances.append(self.parse_instance(name, local))
            return
        body = self.parse_expr()
        last = self.toks[self.i - 1]
        mod.defs.append(Def(name, params, body, local, mod.name, start.line, start.col,
                            last.line, last.col + len(last.text) - 1))

    # -- expressions -----------------------------------------------------
    def parse_expr(self, min_prec: int = 0) -> tuple:
        lhs = self.parse_prefix()
        while True:
            if self.blocked():
                return lhs
            t = self.tok
            if t.kind != "op" or t.text not in _BINOPS:
                return lhs
            prec, assoc = _BINOPS[t.text]
            if prec < min_prec:
                return lhs
            op = self.advance().text
            next_min = prec + 1 if assoc in ("L", "N") else prec
            # '-' binds tighter on the right than '+', both left-assoc: prec+1 is right for both
            rhs = self.parse_expr(next_min)
            if op == "/\\":
                lhs = ("and", _flat("and", lhs) + _flat("and", rhs))
            elif op == "\\/":
                lhs = ("or", _flat("or", lhs) + _flat("or", rhs))
            else:
                if op == "=<":
                    op = "<="
                if op == "/=":
                    op = "#"
                if op == "\\X":
                    lhs = ("cross", (lhs[1] if lhs[0] == "cross" else [lhs]) + [rhs])
                else:
                    lhs = ("binop", op, lhs, rhs)

    def parse_junction_list(self) -> tuple:
        bullet = self.tok
        kind = "and" if bullet.text == "/\\" else "or"
        col = bullet.col
        items = []
        while self.tok.kind == "op" and self.tok.text == bullet.text and self.tok.col == col:
            # a bullet at this column is only ours if no *inner* barrier forbids it
            if self.barriers and col <= self.barriers[-1]:
                break
            self.advance()
            self.barriers.append(col)
            try:
                items.append(self.parse_expr())
            finally:
                self.barriers.pop()
        # NB: items are NOT flattened into each other: a nested list of the other
        # kind is a single item.  Same-kind infix chains inside an item stay nested too;
        # that is semantically identical.
        return (kind, items) if len(items) > 1 else items[0]

    def parse_prefix(self) -> tuple:
        if self.blocked():
            self.error("expression expected (junction-list indentation?)")
        t = self.tok
        if t.kind == "op":
            if t.text in ("/\\", "\\/"):
                return self.parse_junction_list()
            if t.text == "~":
                self.advance()
                return ("not", self.parse_expr(4))
            if t.text == "-":
                self.advance()
                e = self.parse_expr(12)
                return ("num", -e[1]) if e[0] == "num" else ("neg", e)
            if t.text in ("\\E", "\\A"):
                return self.parse_quant()
            if t.text == "[]":
                self.advance()
                return ("box", self.parse_expr(4))
            if t.text == "<>":
                self.advance()
                return ("diamond", self.parse_expr(4))
        if t.kind == "kw":
            if t.text == "IF":
                self.advance()
                c = self.parse_expr()
                self.expect_kw("THEN")
                a = self.parse_expr()
                self.expect_kw("ELSE")
                b = self.parse_expr()
                return ("if", c, a, b)
            if t.text == "LET":
                return self.parse_let()
            if t.text == "CASE":
                # CASE p1 -> e1 [] p2 -> e2 [] OTHER -> e   ==> ('case', [(p, e)...], other | None)
                self.advance()
                arms, other = [], None
                while True:
                    if self.is_kw("OTHER"):
                        self.advance()
                        self.expect_op("->")
                        other = self.parse_expr()
                    else:
                        g = self.parse_expr()
                        self.expect_op("->")
                        arms.append((g, self.parse_expr()))
                    if self.is_op("[]") and not self.blocked():
                        self.advance()
                        continue
                    break
                return ("case", arms, other)
            if t.text == "CHOOSE":
                self.advance()
                name = self.expect_id()
                self.expect_op("\\in")
                s = self.parse_expr(6)
                self.expect_op(":")
                return ("choose", name, s, self.parse_expr())
            if t.text == "SUBSET":
                self.advance()
                return ("subset", self.parse_expr(9))
            if t.text == "UNION":
                self.advance()
                return ("union_all", self.parse_expr(9))
            if t.text == "DOMAIN":
                self.advance()
                return ("domain", self.parse_expr(9))
            if t.text == "UNCHANGED":
                self.advance()
                return ("unchanged", self.parse_expr(5))
            if t.text == "ENABLED":
                self.advance()
                return ("enabled", self.parse_expr(5))
        return self.parse_postfix(self.parse_atom())

    def parse_quant(self) -> tuple:
        q = "E" if self.advance().text == "\\E" else "A"
        bounds = self.parse_bounds()
        self.expect_op(":")
        return ("quant", q, bounds, self.parse_expr())

    def parse_bounds(self) -> list[tuple[list[str], tuple]]:
        """x \\in S | x, y \\in S | x \\in S, y \\in T"""
        bounds = []
        while True:
            names = [self.expect_id()]
            while self.is_op(","):
                self.advance()
                names.append(self.expect_id())
            self.expect_op("\\in")
            s = self.parse_expr(6)
            bounds.append((names, s))
            if self.is_op(","):
                self.advance()
                continue
            return bounds

    def parse_let(self) -> tuple:
        self.expect_kw("LET")
        defs: list[Def] = []
        # LET definitions are not subject to the enclosing junction barrier in
        # practice (they are always indented further); parse until IN.
        while not self.is_kw("IN"):
            start = self.tok
            name = self.expect_id()
            params: list[str] = []
            if self.is_op("("):
                self.advance()
                params.append(self.expect_id())
                while self.is_op(","):
                    self.advance()
                    params.append(self.expect_id())
                self.expect_op(")")
            self.expect_op("==")
            body = self.parse_expr()
            defs.append(Def(name, params, body, True, self.module_hint, start.line, start.col))
        self.expect_kw("IN")
        return ("let", defs, self.parse_expr())

    def parse_atom(self) -> tuple:
        t = self.tok
        if t.kind == "num":
            self.advance()
            return ("num", int(t.text))
        if t.kind == "str":
            self.advance()
            return ("str", t.text)
        if t.kind == "kw" and t.text in ("TRUE", "FALSE"):
            self.advance()
            return ("bool", t.text == "TRUE")
        if t.kind == "kw" and t.text in ("BOOLEAN", "STRING"):
            self.advance()
            return ("id", t.text)
        if t.kind == "id":
            self.advance()
            name = t.text
            if name.startswith(("SF_", "WF_")) and self.is_op("("):
                self.advance()
                a = self.parse_expr()
                self.expect_op(")")
                return ("fair", name[:2], ("id", name[3:]), a)
            if self.is_op("!") and not self.blocked():
                # I!Op or I!Op(args)
                self.advance()
                op = self.expect_id()
                args = self.parse_call_args()
                return ("inst", name, op, args)
            if self.is_op("(") and not self.blocked():
                return ("app", name, self.parse_call_args())
            return ("id", name)
        if t.kind == "op":
            if t.text == "@":
                self.advance()
                return ("at",)
            if t.text == "(":
                self.advance()
                saved, self.barriers = self.barriers, []
                try:
                    e = self.parse_expr()
                finally:
                    self.barriers = saved
                self.expect_op(")")
                return e
            if t.text == "{":
                return self.parse_brace()
            if t.text == "[":
                return self.parse_bracket()
            if t.text == "<<":
                self.advance()
                saved, self.barriers = self.barriers, []
                try:
                    items = []
                    if not self.is_op(">>"):
                        items.append(self.parse_expr())
                        while self.is_op(","):
                            self.advance()
                            items.append(self.parse_expr())
                finally:
                    self.barriers = saved
                self.expect_op(">>")
                return ("tuple", items)
        self.error("unexpected token in expression")

    def parse_call_args(self) -> list[tuple]:
        if not self.is_op("(") or self.blocked():
            return []
        self.advance()
        saved, self.barriers = self.barriers, []
        try:
            args = [self.parse_expr()]
            while self.is_op(","):
                self.advance()
                args.append(self.parse_expr())
        finally:
            self.barriers = saved
        self.expect_op(")")
        return args

    def parse_postfix(self, e: tuple) -> tuple:
        while not self.blocked():
            t = self.tok
            if t.kind == "op" and t.text == "[":
                self.advance()
                saved, self.barriers = self.barriers, []
                try:
                    args = [self.parse_expr()]
                    while self.is_op(","):
                        self.advance()
                        args.append(self.parse_expr())
                finally:
                    self.barriers = saved
                self.expect_op("]")
                e = ("fnapp", e, args)
            elif t.kind == "op" and t.text == "." and self.peek().kind == "id":
                self.advance()
                e = ("dot", e, self.expect_id())
            elif t.kind == "op" and t.text == "'":
                self.advance()
                e = ("prime", e)
            else:
                break
        return e

    def parse_brace(self) -> tuple:
        self.expect_op("{")
        saved, self.barriers = self.barriers, []
        try:
            if self.is_op("}"):
                self.advance()
                return ("setenum", [])
            # {x \in S : p}  vs  {e : x \in S}  vs  {a, b}
            if self.tok.kind == "id" and self.peek().kind == "op" and self.peek().text == "\\in":
                save_i = self.i
                name = self.expect_id()
                self.expect_op("\\in")
                s = self.parse_expr(6)
                if self.is_op(":"):
                    self.advance()
                    p = self.parse_expr()
                    self.expect_op("}")
                    return ("setfilter", name, s, p)
                self.i = save_i
            first = self.parse_expr()
            if self.is_op(":"):
                self.advance()
                bounds = self.parse_bounds()
                self.expect_op("}")
                return ("setmap", first, bounds)
            items = [first]
            while self.is_op(","):
                self.advance()
                items.append(self.parse_expr())
            self.expect_op("}")
            return ("setenum", items)
        finally:
            self.barriers = saved

    def parse_bracket(self) -> tuple:
        self.expect_op("[")
        saved, self.barriers = self.barriers, []
        try:
            # record constructor / record set: [f |-> e, ...]  /  [f : S, ...]
            if self.tok.kind == "id" and self.peek().kind == "op" and self.peek().text in ("|->", ":"):
                is_set = self.peek().text == ":"
                fields = []
                while True:
                    f = self.expect_id()
                    self.expect_op(":" if is_set else "|->")
                    fields.append((f, self.parse_expr()))
                    if self.is_op(","):
                        self.advance()
                        continue
                    break
                self.expect_op("]")
                return ("recset" if is_set else "rec", fields)
            # function constructor [x \in S |-> e]
            if self.tok.kind == "id" and self.peek().kind == "op" and self.peek().text in ("\\in", ","):
                save_i = self.i
                try:
                    bounds = self.parse_bounds()
                    if self.is_op("|->"):
                        self.advance()
                        body = self.parse_expr()
                        self.expect_op("]")
                        return ("fnlit", bounds, body)
                except TlaSyntaxError:
                    pass
                self.i = save_i
            e = self.parse_expr()
            if self.is_op("->"):
                self.advance()
                t = self.parse_expr()
                self.expect_op("]")
                return ("fnset", e, t)
            if self.is_kw("EXCEPT"):
                self.advance()
                updates = []
                while True:
                    self.expect_op("!")
                    path = []
                    while True:
                        if self.is_op("["):
                            self.advance()
                            path.append(("idx", self.parse_expr()))
                            self.expect_op("]")
                        elif self.is_op("."):
                            self.advance()
                            path.append(("fld", self.expect_id()))
                        else:
                            break
                    self.expect_op("=")
                    updates.append((path, self.parse_expr()))
                    if self.is_op(","):
                        self.advance()
                        continue
                    break
                self.expect_op("]")
                return ("except", e, updates)
            self.expect_op("]")
            # [A]_v  (action box; temporal formulas only)
            if self.tok.kind == "id" and self.tok.text.startswith("_"):
                v = self.advance().text[1:]
                return ("actionbox", e, ("id", v))
            self.error("unsupported bracket expression")
        finally:
            self.barriers = saved


def _flat(kind: str, e: tuple) -> list[tuple]:
    # keep bullet-list structure intact: only infix chains are merged by the caller
    return [e]


def parse_module_text(text: str, hint: str = "?") -> Module:
    return Parser(tokenize(text), hint).parse_module()


def parse_expression_text(text: str) -> tuple:
    """Parse a stand-alone expression (used for cfg pragma values)."""
    toks = tokenize("---- MODULE _X ----\n" + text + "\n====")
    p = Parser(toks, "<expr>")
    p.i = 4  # skip: sep MODULE _X sep
    e = p.parse_expr()
    if p.tok.kind not in ("end", "eof"):
        p.error("trailing tokens in expression")
    return e
