"""Folds tests/golden/oracle_a/<name>.json (written by run_oracle_a.py: hours of Python per 3-replica model) into
goldens.json.  Every number Oracle A reports is first CHECKED against the entry that is already there (Oracle B's):
distinct, generated, depth, per-level widths, deadlocks and the first-violation level of every invariant Oracle B
knows.  Only then the entry gains "oracle_a" among its sources, the interpreter's TypeOk verdict and the
order-independent digest of the reachable state set (canonical TLC text), which the lowered model and the GPU
must reproduce state for state.

    python tests/golden/merge_oracle_a.py            # all files present
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    gp = os.path.join(HERE, "goldens.json")
    goldens = json.load(open(gp))
    d = os.path.join(HERE, "oracle_a")
    merged = []
    for fn in sorted(os.listdir(d)):
        name = fn[:-5]
        if not fn.endswith(".json") or name not in goldens:
            continue
        a = json.load(open(os.path.join(d, fn)))
        g = goldens[name]
        for k in ("distinct", "generated", "depth", "levels", "deadlocks"):
            if a[k] != g[k]:
                sys.exit(f"{name}: Oracle A disagrees with the golden on {k}: {a[k]} vs {g[k]}")
        for inv, lvl in g["first_violation_level"].items():
            if inv in a["first_violation_level"] and a["first_violation_level"][inv] != lvl:
                sys.exit(f"{name}: first violation of {inv}: Oracle A level {a['first_violation_level'][inv]}, golden {lvl}")
        g["first_violation_level"] = {**g["first_violation_level"], **a["first_violation_level"]}
        if "state_digest" in a:
            if g.get("state_digest", a["state_digest"]) != a["state_digest"]:
                sys.exit(f"{name}: state digest differs from the one recorded")
            g["state_digest"] = a["state_digest"]
        if "oracle_a" not in g["sources"]:
            g["sources"].append("oracle_a")
        g["oracle_a_seconds"] = a.get("seconds")
        merged.append(name)
    with open(gp, "w") as f:
        json.dump(goldens, f, indent=1, sort_keys=True)
    print("merged:", merged)


if __name__ == "__main__":
    main()
