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
