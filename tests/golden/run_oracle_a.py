"""Runs Oracle A (oracle/tla_interp.py, direct interpreter of the unchanged .tla text) on ONE model of
models/MODELS.json and writes tests/golden/oracle_a/<name>.json: counts, per-level widths, first-violation
levels and the order-independent digest of the reachable state set (canonical TLC-style text).

These runs take from minutes to hours of Python (about 400 states/s), so they are kept out of
make_golden.py; tests/golden/merge_oracle_a.py folds the files into goldens.json after checking every
number against the Oracle-B entry that is already there.  Run HERE (needs /root/reference):

    nice -n 19 python tests/golden/run_oracle_a.py kip320_small
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)

import tla_interp  # noqa: E402
from make_golden import state_digest  # noqa: E402


def main():
    name = sys.argv[1]
    reg = json.load(open(os.path.join(ROOT, "models", "MODELS.json")))
    spec = reg[name]
    cfg_text = open(os.path.join(ROOT, spec["cfg"])).read() + "\nCHECK_DEADLOCK FALSE\n"
    dirs = ["/root/reference", os.path.join(ROOT, "models"), os.path.join(ROOT, "tests", "specs")]
    t0 = time.time()
    a = tla_interp.run_bfs(spec["module"], dirs, cfg_text, collect_states=True, stop_on_violation=False)
    out = {k: a[k] for k in ("distinct", "generated", "depth", "levels", "deadlocks", "first_violation_level")}
    out["violating_states"] = a.get("violating_states")
    out["seconds"] = round(time.time() - t0, 1)
    if not spec.get("symmetry"):
        out["state_digest"] = state_digest(a["states"])
    os.makedirs(os.path.join(HERE, "oracle_a"), exist_ok=True)
    with open(os.path.join(HERE, "oracle_a", name + ".json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(name, {k: v for k, v in out.items() if k != "levels"}, flush=True)


if __name__ == "__main__":
    main()
