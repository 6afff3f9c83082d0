"""Static SASS census of a model library's kernels (no GPU needed): instruction count and the mnemonics that matter
for the design claims -- generic stores / QSPC / local memory in k_expand (must be absent), 128-bit loads, the 128-bit
CAS of the exact set, shared-memory atomics of the pair bookkeeping, peer-synchronisation primitives.

    python tools/sass_summary.py kip320_3x4_r4e3 > profiles/r2_sass_kip320_3x4_r4e3.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["QSPC", "LDL", "STL", "ST.E", "STG.E", "STS", "LDS", "LDG.E.128", "LDG.E.64", "LD.E", "ATOMS", "ATOMG.E.CAS.128",
         "ATOMG.E.CAS.64", "ATOMG", "RED", "BAR.SYNC", "VOTE", "MATCH", "SHFL", "POPC", "NANOSLEEP", "CS2R", "CCTL", "MEMBAR",
         "IMAD", "LOP3", "ISETP", "BRA", "BSSY"]


def main():
    model = sys.argv[1]
    so = os.path.join(ROOT, "build", "models", model, f"libkmc_{model}.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    fn, per = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if fn and m:
            op = m.group(1)
            per[fn]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w.count(".") and op.startswith(w)):
                    per[fn][w] += 1
    regs = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
    print(f"# {model}: sm_100a SASS census (cuobjdump -sass), {os.path.basename(so)}")
    for fn, c in per.items():
        print(f"\n{fn}: {c['_total']} instructions")
        print("   " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
    print("\n# resource usage (cuobjdump -res-usage)")
    for line in regs.splitlines():
        if "Function" in line or "REG" in line:
            print("  " + line.strip()[:200])


if __name__ == "__main__":
    main()
