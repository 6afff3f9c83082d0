"""Front-end: the reference's .tla files parse unchanged; cfg grammar; scoping rules."""
import os

import pytest

from conftest import REFERENCE, ROOT, needs_reference
from kafka_specification_b200.frontend.cfg import ModelValue, parse_cfg
from kafka_specification_b200.frontend.modules import Loader, load_root
from kafka_specification_b200.frontend.tla_lexer import strip_comments, tokenize
from kafka_specification_b200.frontend.tla_parser import parse_expression_text, parse_module_text

ALL = ["IdSequence", "Util", "FiniteReplicatedLog", "KafkaReplication", "KafkaTruncateToHighWatermark",
       "Kip101", "Kip279", "Kip320", "Kip320FirstTry", "AsyncIsr"]


@needs_reference
@pytest.mark.parametrize("name", ALL)
def test_reference_modules_parse(name):
    m = Loader([REFERENCE]).load(name)
    assert m.name == name
    assert m.defs or m.instances


@needs_reference
def test_definition_inventory():
    L = Loader([REFERENCE])
    k = L.load("KafkaReplication")
    names = [d.name for d in k.defs]
    for want in ["TypeOk", "Init", "ControllerShrinkIsr", "ControllerElectLeader", "BecomeLeader", "LeaderWrite",
                 "LeaderShrinkIsr", "LeaderExpandIsr", "LeaderIncHighWatermark", "BecomeFollowerAndTruncateTo",
                 "FollowerReplicate", "WeakIsr", "StrongIsr", "LeaderInIsr"]:
        assert want in names
    assert "Next" not in names                                  # KafkaReplication.tla has no Next
    assert [(i.name, i.module) for i in k.instances] == [
        ("LeaderEpochSeq", "IdSequence"), ("RecordSeq", "IdSequence"), ("ReplicaLog", "FiniteReplicatedLog")]
    assert k.constants == ["Replicas", "LogSize", "MaxRecords", "MaxLeaderEpoch"]
    assert len(k.variables) == 6


@needs_reference
def test_local_next_scoping():
    # Kip320 EXTENDS Kip279, whose Next is LOCAL (Kip279.tla:53); the cfg-level Next must be Kip320's.
    root = load_root("Kip320", [REFERENCE])
    d = root.find_def("Next", None)
    assert d.module == "Kip320" and not d.local
    # a root module whose own Next is LOCAL is still addressable from the cfg
    root = load_root("Kip279", [REFERENCE])
    d = root.find_def("Next", None)
    assert d.module == "Kip279" and d.local
    # LOCAL helper of Kip320 invisible from Kip279's definitions
    assert load_root("Kip320", [REFERENCE]).find_def("IsFollowingLeaderEpoch", "Kip279") is None
    assert load_root("Kip320", [REFERENCE]).find_def("IsFollowingLeaderEpoch", "Kip320") is not None


@needs_reference
def test_instance_substitution():
    root = load_root("Kip320", [REFERENCE])
    inst = root.instance("ReplicaLog")
    r = inst.resolve("logs", None)
    assert r.kind == "subst" and r.expr == ("id", "replicaLog")
    r = inst.resolve("LogRecords", None)          # implicit same-name substitution by a *definition*
    assert r.kind == "subst" and r.expr == ("id", "LogRecords") and r.ctx is root
    seq = root.instance("RecordSeq")
    assert seq.resolve("MaxId", None).expr == ("binop", "-", ("id", "MaxRecords"), ("num", 1))


def test_junction_lists_by_column():
    src = """---- MODULE T ----
A == /\\ x = 1
     /\\ \\/ y = 2
        \\/ /\\ y = 3
           /\\ z = 4
     /\\ w = 5
B == a /\\ b \\/ c
====
"""
    m = parse_module_text(src)
    a = m.defs[0].body
    assert a[0] == "and" and len(a[1]) == 3
    assert a[1][1][0] == "or" and len(a[1][1][1]) == 2
    assert a[1][1][1][1][0] == "and" and len(a[1][1][1][1][1]) == 2
    assert m.defs[1].name == "B"


def test_nested_comments_and_footer():
    src = "(* a (* nested *) b *) ---- MODULE T ----\nX == 1 \\* tail\n====\nModification History garbage == (("
    m = parse_module_text(src)
    assert [d.name for d in m.defs] == ["X"]
    assert "nested" not in strip_comments("(* a (* nested *) b *) x")


def test_expression_forms():
    e = parse_expression_text("[f EXCEPT ![a].b[c] = @ + 1, !.d = {}]")
    assert e[0] == "except" and len(e[2]) == 2 and e[2][0][0] == [("idx", ("id", "a")), ("fld", "b"), ("idx", ("id", "c"))]
    assert parse_expression_text("[x \\in S |-> x]")[0] == "fnlit"
    assert parse_expression_text("[S -> T]")[0] == "fnset"
    assert parse_expression_text("[a : S, b : T]")[0] == "recset"
    assert parse_expression_text("{x \\in S : x > 1}")[0] == "setfilter"
    assert parse_expression_text("{x + 1 : x \\in S}")[0] == "setmap"
    assert parse_expression_text("\\E a, b \\in S, c \\in T : a = c")[2] == [(["a", "b"], ("id", "S")), (["c"], ("id", "T"))]
    assert parse_expression_text("I!Op(1, 2)") == ("inst", "I", "Op", [("num", 1), ("num", 2)])
    assert parse_expression_text("-1") == ("num", -1)
    assert parse_expression_text("Init /\\ [][Next]_vars /\\ WF_vars(A)")[0] == "and"


def test_cfg_parser():
    cfg = parse_cfg("""
\\* comment
CONSTANTS
    Replicas = {r1, r2, r3}   LogSize = 4
    Nil = nil   Name = "x"  Flag <- Other
INIT Init NEXT Next
INVARIANTS TypeOk WeakIsr
CONSTRAINT Bound
CHECK_DEADLOCK FALSE
\\* kspec: LAYOUT Layout
\\* kspec: CAPACITY reqs = MaxLeaderEpoch + 1
""")
    assert cfg.constants["Replicas"] == frozenset({ModelValue("r1"), ModelValue("r2"), ModelValue("r3")})
    assert cfg.constants["LogSize"] == 4 and cfg.constants["Nil"] == ModelValue("nil") and cfg.constants["Name"] == "x"
    assert cfg.overrides == {"Flag": "Other"}
    assert (cfg.init, cfg.next) == ("Init", "Next")
    assert cfg.invariants == ["TypeOk", "WeakIsr"] and cfg.constraints == ["Bound"]
    assert cfg.check_deadlock is False and cfg.layout == "Layout"
    assert cfg.capacities == {"reqs": "MaxLeaderEpoch + 1"}
    assert parse_cfg("SPECIFICATION Spec").check_deadlock is True


def test_models_directory_cfgs_parse():
    d = os.path.join(ROOT, "models")
    n = 0
    for f in os.listdir(d):
        if f.endswith(".cfg"):
            cfg = parse_cfg(open(os.path.join(d, f)).read())
            assert cfg.init == "Init" and cfg.next in ("Next", "NextWith279")     # (MCKip320With279: the reference's proposed experiment)
            n += 1
    assert n >= 10
