"""Front-end / lowering coverage beyond the ten reference files.

tests/specs/MiniLock.tla is a synthetic spec using CASE/OTHER, Cardinality, `x' \\in S`, a primed
variable read on a right-hand side and BOOLEAN variables.  Oracle A (direct interpreter) and the
lowered model (host build of the generated header) must agree state for state."""
import os

from conftest import ROOT
from golden.make_golden import state_digest
from hostmodel import lower_model, run_host
from kafka_specification_b200.frontend.tla_parser import parse_expression_text

SPECS = os.path.join(ROOT, "tests", "specs")


def test_case_expression_parses():
    e = parse_expression_text("CASE x < 1 -> 0 [] x = 1 -> 5 [] OTHER -> x")
    assert e[0] == "case" and len(e[1]) == 2 and e[2] == ("id", "x")
    e = parse_expression_text("CASE a -> 1 [] b -> 2")
    assert e[0] == "case" and len(e[1]) == 2 and e[2] is None


def test_minilock_oracle_vs_lowering():
    import tla_interp
    cfg = open(os.path.join(SPECS, "MiniLock.cfg")).read()
    a = tla_interp.run_bfs("MiniLock", [SPECS], cfg, collect_states=True, stop_on_violation=False)
    assert (a["distinct"], a["generated"], a["depth"]) == (76, 169, 14)
    m = lower_model("MiniLock", [SPECS], cfg)
    assert not m.warnings and m.words == 1
    for items in (False, True):
        r = run_host(m, dump=True, max_states=1000, items=items)
        assert (r["distinct"], r["generated"], r["depth"], r["levels"], r["deadlocks"]) == (
            a["distinct"], a["generated"], a["depth"], a["levels"], a["deadlocks"])
        assert r["first_violated"] is None
        assert state_digest([m.state_text(row) for row in r["states"]]) == state_digest(a["states"])


def test_minilock_violation_levels_agree():
    """A deliberately false invariant: both sides must find it at the same BFS level."""
    import tla_interp
    cfg = open(os.path.join(SPECS, "MiniLock.cfg")).read().replace("INVARIANTS TypeOk Bounded HolderNotWaiting",
                                                                    "INVARIANTS TypeOk NeverTwo")
    src = open(os.path.join(SPECS, "MiniLock.tla")).read()
    tmp = os.path.join(ROOT, "build", "hosttest", "specs")
    os.makedirs(tmp, exist_ok=True)
    with open(os.path.join(tmp, "MiniLock.tla"), "w") as f:
        f.write(src.replace("HolderNotWaiting ==", "NeverTwo == Cardinality(waiting) < 2\nHolderNotWaiting =="))
    a = tla_interp.run_bfs("MiniLock", [tmp], cfg, stop_on_violation=False)
    lvl = a["first_violation_level"]["NeverTwo"]
    assert lvl is not None
    m = lower_model("MiniLock", [tmp], cfg)
    r = run_host(m)
    assert r["first_violated"] == "NeverTwo" and r["first_violated_level"] == lvl


# ---------------------------------------------------------------------------------------------------------------
# module Sequences, tuples as values, \X, RECURSIVE (tests/specs/MiniQueue.tla)
# ---------------------------------------------------------------------------------------------------------------
def test_sequence_syntax_parses():
    e = parse_expression_text("A \\X B \\X C")
    assert e[0] == "cross" and len(e[1]) == 3
    e = parse_expression_text("Append(s, 1) \\o <<2, 3>>")
    assert e == ("binop", "\\o", ("app", "Append", [("id", "s"), ("num", 1)]), ("tuple", [("num", 2), ("num", 3)]))
    from kafka_specification_b200.frontend.tla_parser import parse_module_text
    m = parse_module_text("---- MODULE R ----\nRECURSIVE F(_), G(_, _)\nF(n) == IF n = 0 THEN 0 ELSE n + F(n - 1)\n"
                          "G(a, b) == a\n====\n")
    assert [d.name for d in m.defs] == ["F", "G"]


def test_miniqueue_oracle_vs_lowering():
    """Sequences with a run-time length (Append / Tail / SubSeq / \\o / EXCEPT / DOMAIN), a tuple-valued variable typed
    by a Cartesian product and a RECURSIVE operator unfolded over the bounded sequence: Oracle A and the lowered
    model agree state for state, in both forms of the lowered Next."""
    import tla_interp
    cfg = open(os.path.join(SPECS, "MiniQueue.cfg")).read()
    a = tla_interp.run_bfs("MiniQueue", [SPECS], cfg, collect_states=True, stop_on_violation=False)
    assert (a["distinct"], a["generated"], a["depth"], a["levels"]) == (160, 729, 6, [1, 3, 12, 36, 27, 81])
    assert all(v is None for v in a["first_violation_level"].values())
    m = lower_model("MiniQueue", [SPECS], cfg)
    assert not m.warnings and m.words == 1 and m.state_bits == 14
    assert m.layout["types"]["queue"] == {"t": "seq", "cap": 3, "elem": {"t": "int", "lo": 0, "hi": 2}}
    assert m.layout["types"]["last"]["t"] == "tuple"
    for items in (False, True):
        r = run_host(m, dump=True, max_states=10000, items=items)
        assert (r["distinct"], r["generated"], r["depth"], r["levels"], r["deadlocks"]) == (
            a["distinct"], a["generated"], a["depth"], a["levels"], a["deadlocks"])
        assert r["first_violated"] is None
        assert state_digest([m.state_text(row) for row in r["states"]]) == state_digest(a["states"])
    # the run-time decoder (model.json only, what traces on the GPU box use) prints the same text
    from kafka_specification_b200.runtime import StateDecoder
    dec = StateDecoder(m.meta())
    assert [dec.text(row) for row in r["states"]] == [m.state_text(row) for row in r["states"]]
    assert any("queue = <<2, 1, 0>>" in dec.text(row) for row in r["states"])


def test_miniqueue_violation_and_capacity_trap():
    import tla_interp
    from kafka_specification_b200.lower.svals import LowerError
    src = open(os.path.join(SPECS, "MiniQueue.tla")).read()
    tmp = os.path.join(ROOT, "build", "hosttest", "specs_q")
    os.makedirs(tmp, exist_ok=True)
    with open(os.path.join(tmp, "MiniQueue.tla"), "w") as f:
        f.write(src.replace("SumOk ==", "NeverFull == Len(queue) < Cap \\/ Head(queue) # MaxVal\nSumOk =="))
    cfg = open(os.path.join(SPECS, "MiniQueue.cfg")).read().replace("INVARIANTS TypeOk SumOk LastOk Bounded",
                                                                     "INVARIANTS TypeOk NeverFull")
    a = tla_interp.run_bfs("MiniQueue", [tmp], cfg, stop_on_violation=False)
    lvl = a["first_violation_level"]["NeverFull"]
    assert lvl == 4
    m = lower_model("MiniQueue", [tmp], cfg)
    r = run_host(m)
    assert r["first_violated"] == "NeverFull" and r["first_violated_level"] == lvl
    # the bound of a sequence comes from `\* kspec: CAPACITY`, else from a conjunct `Len(v) <= N` of the type invariant
    # (MiniQueue's TypeOk has one) ...
    plain = open(os.path.join(SPECS, "MiniQueue.cfg")).read()
    m1 = lower_model("MiniQueue", [SPECS], plain.replace("\\* kspec: CAPACITY queue = Cap", ""))
    assert m1.layout["types"]["queue"]["cap"] == 3 and m1.layout == lower_model("MiniQueue", [SPECS], plain).layout
    # ... and a sequence variable with neither is rejected with the hint to give one (MiniMsgs' TypeOk has no Len bound)
    try:
        lower_model("MiniMsgs", [SPECS], open(os.path.join(SPECS, "MiniMsgs.cfg")).read().replace("\\* kspec: CAPACITY inbox = MaxLen", ""))
        assert False, "expected a LowerError"
    except LowerError as e:
        assert "CAPACITY" in str(e)
    # ... and a bound that is too small is a checked hint: the run traps instead of wrapping
    m2 = lower_model("MiniQueue", [SPECS], open(os.path.join(SPECS, "MiniQueue.cfg")).read().replace("CAPACITY queue = Cap", "CAPACITY queue = 2"))
    r2 = run_host(m2)
    assert r2["fail"] == 1                        # KMC_FAIL_LAYOUT


def test_minimsgs_sequence_of_records():
    """tests/specs/MiniMsgs.tla: a sequence of records (one mixed-radix code per slot), indexing with a run-time
    index, a two-parameter RECURSIVE operator, \\o with a run-time-length operand."""
    import tla_interp
    cfg = open(os.path.join(SPECS, "MiniMsgs.cfg")).read()
    a = tla_interp.run_bfs("MiniMsgs", [SPECS], cfg, collect_states=True, stop_on_violation=False)
    assert (a["distinct"], a["generated"], a["depth"]) == (548, 1409, 10)
    m = lower_model("MiniMsgs", [SPECS], cfg)
    assert not m.warnings and m.words == 1 and m.state_bits == 12
    for items in (False, True):
        r = run_host(m, dump=True, max_states=100000, items=items)
        assert (r["distinct"], r["generated"], r["depth"], r["levels"], r["deadlocks"]) == (
            a["distinct"], a["generated"], a["depth"], a["levels"], a["deadlocks"])
        assert r["first_violated"] is None
        assert state_digest([m.state_text(row) for row in r["states"]]) == state_digest(a["states"])
    from kafka_specification_b200.runtime import StateDecoder
    dec = StateDecoder(m.meta())
    assert [dec.text(row) for row in r["states"]] == [m.state_text(row) for row in r["states"]]


def test_miniwindow_runtime_bounds_and_products():
    """tests/specs/MiniWindow.tla: SubSeq with a run-time lower bound, Len of a concatenation, \\E over 1 .. Len(s) with
    a run-time index, a record with a tuple-valued field, membership in Nat \\X Nat."""
    import numpy as np
    import tla_interp
    from kafka_specification_b200.runtime import StateDecoder
    cfg = open(os.path.join(SPECS, "MiniWindow.cfg")).read()
    a = tla_interp.run_bfs("MiniWindow", [SPECS], cfg, collect_states=True, stop_on_violation=False)
    assert (a["distinct"], a["generated"], a["depth"]) == (189, 371, 7)
    m = lower_model("MiniWindow", [SPECS], cfg)
    assert not m.warnings and m.state_bits == 11
    for items in (False, True):
        r = run_host(m, dump=True, max_states=100000, items=items)
        assert (r["distinct"], r["generated"], r["depth"], r["levels"], r["fail"]) == (
            a["distinct"], a["generated"], a["depth"], a["levels"], 0)
        assert r["first_violated"] is None
        assert state_digest([m.state_text(row) for row in r["states"]]) == state_digest(a["states"])
    assert StateDecoder(m.meta()).texts(np.array(r["states"], dtype=np.uint64)) == [m.state_text(row) for row in r["states"]]


def test_integer_division_and_modulus(tmp_path):
    """Integers: \\div (floor) and % with constant and run-time operands, negative dividends included."""
    import tla_interp
    (tmp_path / "DivMod.tla").write_text("""---- MODULE DivMod ----
EXTENDS Integers
VARIABLES x, y, d
TypeOk == x \\in -7 .. 7 /\\ y \\in 0 .. 6 /\\ d \\in 1 .. 3
Init == x = -7 /\\ y = 0 /\\ d = 1
Step == /\\ x < 7
        /\\ x' = x + 1
        /\\ y' = (x + 7) % 5 + (x \\div 4 + 2) % 2
        /\\ d' = (d % 3) + 1
Wrap == /\\ x = 7
        /\\ x' = (x \\div d) - 7
        /\\ y' = x % d
        /\\ UNCHANGED d
Next == Step \\/ Wrap
Sane == y = y % 7 /\\ (x \\div d) * d + x % d = x
====
""")
    cfg = "INIT Init\nNEXT Next\nINVARIANTS TypeOk Sane\nCHECK_DEADLOCK FALSE\n"
    a = tla_interp.run_bfs("DivMod", [str(tmp_path)], cfg, collect_states=True, stop_on_violation=False)
    assert all(v is None for v in a["first_violation_level"].values()) and a["distinct"] >= 15
    m = lower_model("DivMod", [str(tmp_path)], cfg)
    r = run_host(m, dump=True, max_states=10000)
    assert (r["distinct"], r["generated"], r["depth"], r["levels"], r["fail"]) == (
        a["distinct"], a["generated"], a["depth"], a["levels"], 0)
    assert r["first_violated"] is None
    assert state_digest([m.state_text(row) for row in r["states"]]) == state_digest(a["states"])
