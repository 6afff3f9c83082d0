"""Column-aware tokenizer for the TLA+ subset used by the Kafka replication specs.

The parser needs every token's (line, col) because TLA+ conjunction/disjunction
bullet lists are delimited by indentation (see e.g. the three-deep list at
KafkaReplication.tla:158-168 in the reference).

Only the text between the ``---- MODULE M ----`` header and the closing ``====``
is tokenised; the Toolbox "Modification History" footers after ``====`` are
ignored (Util.tla:26-29).
"""
from __future__ import annotations

import re
from dataclasses import dataclass


class TlaSyntaxError(Exception):
    pass


@dataclass(frozen=True)
class Token:
    kind: str   # 'id', 'num', 'str', 'op', 'kw', 'sep', 'end', 'eof'
    text: str
    line: int   # 1-based
    col: int    # 1-based

    def __repr__(self) -> str:  # compact, for error messages
        return f"{self.kind}:{self.text!r}@{self.line}:{self.col}"


KEYWORDS = {
    "MODULE", "EXTENDS", "CONSTANT", "CONSTANTS", "VARIABLE", "VARIABLES",
    "ASSUME", "ASSUMPTION", "AXIOM", "THEOREM", "LEMMA", "INSTANCE", "WITH", "LOCAL",
    "LET", "IN", "IF", "THEN", "ELSE", "CHOOSE", "EXCEPT", "SUBSET", "UNION",
    "DOMAIN", "UNCHANGED", "ENABLED", "CASE", "OTHER", "RECURSIVE", "LAMBDA",
    "TRUE", "FALSE", "BOOLEAN", "STRING",
}

# backslash-words are normalised to these operator spellings
BACKSLASH_WORDS = {
    "\\E": "\\E", "\\A": "\\A", "\\in": "\\in", "\\notin": "\\notin",
    "\\leq": "<=", "\\geq": ">=", "\\union": "\\union", "\\cup": "\\union",
    "\\intersect": "\\intersect", "\\cap": "\\intersect",
    "\\subseteq": "\\subseteq", "\\lnot": "~", "\\neg": "~",
    "\\land": "/\\", "\\lor": "\\/", "\\div": "\\div", "\\X": "\\X", "\\times": "\\X",
    "\\o": "\\o", "\\circ": "\\o",
}

_TOKEN_RE = re.compile(
    r"""
    (?P<ws>[ \t\r]+)
  | (?P<nl>\n)
  | (?P<sep>-{4,})
  | (?P<end>={4,})
  | (?P<num>\d+)
  | (?P<str>"(?:[^"\\]|\\.)*")
  | (?P<id>[A-Za-z_][A-Za-z0-9_]*)
  | (?P<bsword>\\[A-Za-z]+)
  | (?P<op>/\\|\\/|\|->|<=>|->|<-|==|=>|=<|<=|>=|/=|\.\.|<<|>>|\[\]|<>|::|[=<>+\-*/()\[\]{},:.!@'~\#\\^|&%])
    """,
    re.VERBOSE,
)


def strip_comments(text: str) -> str:
    """Replace comments by spaces, preserving line/column positions.

    ``(* ... *)`` comments nest; ``\\*`` comments run to end of line.  String
    literals are respected (the reference has only ``"NONE"``).
    """
    out = list(text)
    i, n, depth = 0, len(text), 0
    while i < n:
        c = text[i]
        if depth == 0 and c == '"':
            j = i + 1
            while j < n and text[j] != '"' and text[j] != "\n":
                j += 2 if text[j] == "\\" else 1
            i = j + 1
            continue
        if text.startswith("(*", i):
            depth += 1
            out[i] = out[i + 1] = " "
            i += 2
            continue
        if depth > 0 and text.startswith("*)", i):
            depth -= 1
            out[i] = out[i + 1] = " "
            i += 2
            continue
        if depth == 0 and text.startswith("\\*", i):
            while i < n and text[i] != "\n":
                out[i] = " "
                i += 1
            continue
        if depth > 0 and c != "\n":
            out[i] = " "
        i += 1
    if depth != 0:
        raise TlaSyntaxError("unterminated (* comment")
    return "".join(out)


def tokenize(text: str) -> list[Token]:
    """Tokenise one module's source text (header through ``====``)."""
    text = strip_comments(text)
    toks: list[Token] = []
    line, line_start, pos, n = 1, 0, 0, len(text)
    in_module = False
    while pos < n:
        m = _TOKEN_RE.match(text, pos)
        if m is None:
            raise TlaSyntaxError(f"unexpected character {text[pos]!r} at {line}:{pos - line_start + 1}")
        kind = m.lastgroup
        s = m.group()
        col = pos - line_start + 1
        pos = m.end()
        if kind == "ws":
            continue
        if kind == "nl":
            line += 1
            line_start = pos
            continue
        if not in_module:
            # prose before the module header is legal TLA+; skip to "---- MODULE"
            if kind == "sep":
                in_module = True
                toks.append(Token("sep", s, line, col))
            continue
        if kind == "end":
            toks.append(Token("end", s, line, col))
            break
        if kind == "id":
            toks.append(Token("kw" if s in KEYWORDS else "id", s, line, col))
        elif kind == "bsword":
            if s not in BACKSLASH_WORDS:
                raise TlaSyntaxError(f"unsupported operator {s} at {line}:{col}")
            toks.append(Token("op", BACKSLASH_WORDS[s], line, col))
        elif kind == "str":
            toks.append(Token("str", s[1:-1], line, col))
        else:
            toks.append(Token(kind, s, line, col))
    toks.append(Token("eof", "", line + 1, 0))
    return toks
