"""A/B runs of one prebuilt model on the GPU through the C ABI: each variant = options (+ optional library
variant suffix, e.g. "1p").  Every run is checked against the golden counts; prints one JSON line per variant
(best of `reps` by device time).  Usage:

    python tools/bench_variants.py kip320_3x4_r4e3 3 '{"tag":"default"}' '{"tag":"1p","lib":"1p","one_phase":true}'
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kafka_specification_b200.runtime import Checker, model_paths  # noqa: E402


def main():
    name, reps = sys.argv[1], int(sys.argv[2])
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "goldens.json"))).get(name)
    for spec in sys.argv[3:]:
        v = json.loads(spec)
        tag = v.pop("tag", "run")
        lib = v.pop("lib", "")
        mdir = v.pop("model", name)                  # lowering variant: its own model directory (<model>@<tag>)
        so, js = model_paths(mdir)
        if lib:
            so = so[:-3] + "." + lib + ".so"
        opts = dict(v)
        if gold and "table_log2" not in opts:
            need = max(gold["distinct"] * 3, 1 << 20)
            opts["table_log2"] = max(20, (need - 1).bit_length())
            opts.setdefault("max_states", int(gold["distinct"] * 1.02) + 4096)
        best = None
        try:
            with Checker(mdir, model_lib=so, model_json=js, cont=True, **opts) as ck:
                for _ in range(reps):
                    r = ck.run(raise_on_error=False)
                    st = r.stats
                    ok = (not gold) or (r.distinct, r.generated, r.depth, r.levels) == (
                        gold["distinct"], gold["generated"], gold["depth"], gold["levels"])
                    row = {"model": name, "tag": tag, "ok": bool(ok and r.complete), "rc": ck.last_rc,
                           "distinct": r.distinct, "generated": r.generated, "depth": r.depth,
                           "gpu_ms": round(st["gpu_ms_total"], 3), "expand_ms": round(st["gpu_ms_expand"], 3),
                           "insert_ms": round(st["gpu_ms_insert"], 3), "invariant_ms": round(st["gpu_ms_invariant"], 3),
                           "probes": st["probes"], "opts": opts}
                    if best is None or row["gpu_ms"] < best["gpu_ms"] or not row["ok"]:
                        best = row
                    if not row["ok"]:
                        break
        except Exception as e:  # noqa: BLE001
            best = {"model": name, "tag": tag, "ok": False, "error": str(e)[:300]}
        print(json.dumps(best), flush=True)


if __name__ == "__main__":
    main()
