"""Builds the A/B variants of one model used by the round-2 GPU experiments (run HERE, needs /root/reference):
engine variants (build.VARIANTS: library suffix) and lowering variants (own model directory <model>@<tag>)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kafka_specification_b200 import build as B  # noqa: E402

LOWER = {"g200": {"guard_lines": 200}, "g3000": {"guard_lines": 3000}, "g800": {"guard_lines": 800}, "g1600s32": {"guard_lines": 1600, "max_group_sites": 32}}


def main():
    model = sys.argv[1]
    reg = B.registry()[model]
    B.build_dispatcher()
    for v in sys.argv[2:]:
        if v in B.VARIANTS:
            print(B.compile_model(model, variant=v), flush=True)
        elif v in LOWER:
            name = f"{model}@{v}"
            B.lower_to_dir(reg["module"], os.path.join(ROOT, reg["cfg"]), name, **LOWER[v])
            print(B.compile_model(name), flush=True)
        else:
            raise SystemExit(f"unknown variant {v}")


if __name__ == "__main__":
    main()
