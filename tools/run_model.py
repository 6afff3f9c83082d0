"""Run prebuilt lowered models on the GPU through the C ABI and print one JSON line each."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kafka_specification_b200.runtime import Checker  # noqa: E402


def main():
    args = sys.argv[1:]
    opts = {}
    names = []
    for a in args:
        if "=" in a:
            k, v = a.split("=", 1)
            opts[k] = json.loads(v)
        else:
            names.append(a)
    for name in names:
        t0 = time.time()
        with Checker(name, **opts) as ck:
            r = ck.run(raise_on_error=False)
            out = {"model": name, "distinct": r.distinct, "generated": r.generated, "depth": r.depth,
                   "deadlocks": r.deadlocks, "complete": r.complete, "violation": r.violation,
                   "gpu_ms": round(r.stats["gpu_ms_total"], 3), "expand_ms": round(r.stats["gpu_ms_expand"], 3),
                   "insert_ms": round(r.stats["gpu_ms_insert"], 3), "invariant_ms": round(r.stats["gpu_ms_invariant"], 3), "wall_ms": round(r.stats["wall_ms"], 1),
                   "probes": r.stats["probes"], "levels": r.levels, "total_s": round(time.time() - t0, 2)}
            print(json.dumps(out), flush=True)
            if r.violation:
                for i, t in enumerate(r.trace):
                    print(f"State {i + 1}: <{t['action']['name'] if t['action'] else 'Initial predicate'}>")
                    print(t["text"])


if __name__ == "__main__":
    main()
