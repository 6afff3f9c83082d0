#!/usr/bin/env python
"""Where does K1 (k_expand) spend its time?  Per-group SM cycles of the lowered Next.

    python tools/group_clocks.py --build kip320_3x4_r4e3      # here (CPU): nvcc -DKMC_GROUP_CLOCKS -> *_gclk.so
    python tools/group_clocks.py kip320_3x4_r4e3              # on a B200: run it, print the table (+ JSON)

The diagnostic library is a separate file next to the production one (libkmc_<model>_gclk.so); the
production kernels carry no clock reads.  A group's cycles are those of the slowest warp of each CTA
(the CTA moves through the groups in lock step), summed over CTAs and tiles.
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kafka_specification_b200 import build  # noqa: E402


def gclk_lib(name):
    return os.path.join(build.model_dir(name), f"libkmc_{name}_gclk.so")


def group_table(name):
    """group -> (lines, emit sites, action ids) read off the lowered header."""
    hdr = open(os.path.join(build.model_dir(name), "model.h")).read()
    starts = [(m.start(), int(m.group(1))) for m in re.finditer(r"void expand_group\(GroupTag<(\d+)>", hdr)]
    end = hdr.index("KMC_HD void expand(")
    out = {}
    for k, (pos, g) in enumerate(starts):
        body = hdr[pos:(starts[k + 1][0] if k + 1 < len(starts) else end)]
        acts = [int(a) for a in re.findall(r"sink\.emit\(n, (\d+)\)", body)]
        out[g] = (body.count("\n"), len(acts), sorted(set(acts)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--build", action="store_true", help="compile the diagnostic library (no GPU needed)")
    ap.add_argument("--table-log2", type=int, default=0)
    ap.add_argument("--max-states", type=int, default=0)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    lib = gclk_lib(a.model)
    if a.build:
        cmd = [build.nvcc_path(), *build.NVCC_ARCH, "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
               "-diag-suppress", "177", "-DKMC_GROUP_CLOCKS", f"-I{build.INCLUDE}", "-include",
               os.path.join(build.model_dir(a.model), "model.h"), os.path.join(build.CSRC, "kmc_engine.cu"), "-o", lib]
        subprocess.run(cmd, check=True)
        print(lib)
        return
    from kafka_specification_b200.runtime import Checker
    reg = json.load(open(os.path.join(ROOT, "models", "MODELS.json")))[a.model]
    opts = {"timing": True}
    if a.table_log2:
        opts["table_log2"] = a.table_log2
    ms = a.max_states or reg.get("max_states", 0)
    if ms:
        opts["max_states"] = ms
    ck = Checker(a.model, model_lib=lib, **opts)
    res = ck.run()
    raw = ctypes.CDLL(lib)
    n = ctypes.c_size_t()
    buf = (ctypes.c_uint64 * 256)()
    raw.kmcm_group_clocks.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_size_t,
                                      ctypes.POINTER(ctypes.c_size_t)]
    # the dispatcher's ctx wraps the model's ctx; the clocks live in a __device__ symbol of the model
    # library, so any non-null ctx pointer is accepted
    rc = raw.kmcm_group_clocks(ck.ctx, buf, 256, ctypes.byref(n))
    if rc != 0:
        sys.exit(f"kmcm_group_clocks: rc={rc} (library not built with -DKMC_GROUP_CLOCKS?)")
    clocks = [int(buf[i]) for i in range(n.value)]
    total = sum(clocks) or 1
    tab = group_table(a.model)
    names = ck.meta.get("actions", [])
    print(f"{a.model}: distinct={res.distinct} generated={res.generated} depth={res.depth} "
          f"expand={res.stats.get('gpu_ms_expand', 0.0):.1f} ms (diagnostic build)")
    print(f"{'group':>5} {'share':>7} {'lines':>6} {'emits':>6}  actions")
    rows = []
    for g, c in enumerate(clocks):
        lines, emits, acts = tab.get(g, (0, 0, []))
        an = [names[i] if i < len(names) else str(i) for i in acts]
        an = [x["name"] if isinstance(x, dict) else x for x in an]
        rows.append({"group": g, "share": c / total, "cycles": c, "lines": lines, "emit_sites": emits, "actions": an})
        print(f"{g:>5} {100.0 * c / total:6.1f}% {lines:>6} {emits:>6}  {', '.join(an)}")
    if a.json:
        with open(a.json, "w") as f:
            json.dump({"model": a.model, "distinct": res.distinct, "groups": rows}, f, indent=1)


if __name__ == "__main__":
    main()
