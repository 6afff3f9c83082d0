"""The two oracles against each other, the closed forms, the goldens and the reference's own claims."""
import os

import pytest

import kso
from conftest import REFERENCE, ROOT, needs_reference


def test_closed_forms_oracle_b():
    # IdSequence.tla:30-39: MaxId + 2 states, one deadlock state
    for n in (0, 1, 4, 9):
        r = kso.run("idsequence", [n], max_states=1000)
        assert (r["distinct"], r["generated"], r["depth"], r["deadlocks"]) == (n + 2, n + 2, n + 2, 1)
    # FiniteReplicatedLog.tla:97-118: (sum_{e<=L} R^e)^n states, depth n*L+1, no deadlock
    for (n, L, R) in [(1, 1, 1), (2, 2, 2), (3, 2, 2), (2, 3, 3), (3, 3, 2), (3, 4, 2)]:
        r = kso.run("frl", [n, L, R], max_states=200000)
        assert r["distinct"] == sum(R ** e for e in range(L + 1)) ** n
        assert r["depth"] == n * L + 1 and r["deadlocks"] == 0


def test_oracle_b_matches_goldens(goldens):
    for name, g in goldens.items():
        if g["distinct"] > 3_000_000 or not g.get("kso"):
            continue                      # headline configs: minutes and tens of GB; synthetic specs: no C restatement
        model, params = g["kso"]
        invs = [i for i in g["first_violation_level"] if i != "TypeOk"]
        r = kso.run(model, params, max_states=4_000_000, invariants=invs, symmetry=bool(g.get("symmetry")))
        for k in ("distinct", "generated", "depth", "levels", "deadlocks"):
            assert r[k] == g[k], (name, k)
        for inv in invs:
            assert r["first_violation_level"][inv] == g["first_violation_level"][inv], (name, inv)


def test_oracle_b_thread_count_independence():
    a = kso.run("kip320", [3, 2, 2, 2], threads=1, max_states=1_000_000)
    b = kso.run("kip320", [3, 2, 2, 2], threads=8, max_states=1_000_000)
    for k in ("distinct", "generated", "depth", "levels", "deadlocks"):
        assert a[k] == b[k]


def test_reference_qualitative_claims(goldens):
    # Kip320.tla:168-171: TypeOk, WeakIsr, StrongIsr hold
    # (with Oracle A merged the entry also carries its TypeOk verdict; both oracles are among the sources)
    assert goldens["kip320_small"]["first_violation_level"] == {"WeakIsr": None, "StrongIsr": None, "TypeOk": None}
    assert set(goldens["kip320_small"]["sources"]) >= {"oracle_a", "oracle_b"}
    # KafkaTruncateToHighWatermark.tla:23-27, Kip279.tla:20-23 (about Kip101), Kip320.tla:126-133 (Kip279
    # truncation without fencing), Kip320FirstTry.tla:27-33: StrongIsr is violated
    for m in ("trunchw_small", "kip101_small", "kip279_small", "firsttry_small"):
        assert goldens[m]["first_violation_level"]["StrongIsr"] is not None, m
    # KafkaReplication.tla:117-119,345: LeaderInIsr is false in the initial state
    assert goldens["leaderinisr_init"]["first_violation_level"]["LeaderInIsr"] == 1


@needs_reference
@pytest.mark.parametrize("name", ["idsequence", "frl_tiny", "asyncisr_v2", "kip320_n2"])
def test_oracle_a_reproduces_goldens(name, goldens, registry):
    """Re-run the direct interpreter on the small cases (the larger oracle_a goldens take minutes)."""
    import tla_interp
    from golden.make_golden import state_digest
    g, spec = goldens[name], registry[name]
    cfg_text = open(os.path.join(ROOT, spec["cfg"])).read() + "\nCHECK_DEADLOCK FALSE\n"
    a = tla_interp.run_bfs(spec["module"], [REFERENCE, os.path.join(ROOT, "models")], cfg_text,
                           collect_states=True, stop_on_violation=False)
    for k in ("distinct", "generated", "depth", "levels", "deadlocks"):
        assert a[k] == g[k], k
    assert a["first_violation_level"] == g["first_violation_level"]
    assert state_digest(a["states"]) == g["state_digest"]


@needs_reference
def test_oracle_a_error_trace_and_init_violation():
    import tla_interp
    cfg = open(os.path.join(ROOT, "models", "LeaderInIsr_init.cfg")).read()
    r = tla_interp.run_bfs("Kip320", [REFERENCE], cfg)
    assert r["violation"]["invariant"] == "LeaderInIsr" and r["violation"]["level"] == 1
    assert len(r["violation"]["trace"]) == 1
    # AsyncIsr's own TypeOk is false initially: pendingVersion = Nil = -1 is not in Nat (AsyncIsr.tla:44,146)
    cfg = """CONSTANTS Replicas = {r1, r2} Leader = r1 MaxOffset = 1
INIT Init NEXT Next INVARIANT TypeOk CHECK_DEADLOCK FALSE"""
    r = tla_interp.run_bfs("AsyncIsr", [REFERENCE], cfg)
    assert r["violation"]["invariant"] == "TypeOk" and r["violation"]["level"] == 1
    # deadlock checking on: IdSequence stops at nextId = MaxId + 1 with a 6-state trace
    r = tla_interp.run_bfs("IdSequence", [REFERENCE], "CONSTANT MaxId = 4 INIT Init NEXT Next")
    assert r["violation"]["invariant"] == "<deadlock>" and len(r["violation"]["trace"]) == 6
