"""The C-ABI library: loads, exports every symbol include/kspecmc.h declares, fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "kspecmc.h")
LIB = os.path.join(ROOT, "build", "libkspecmc.so")


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(kmc_[a-z_0-9]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from kafka_specification_b200 import build as B
    B.build_dispatcher()
    return ctypes.CDLL(LIB)


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for s in ("kmc_create", "kmc_run", "kmc_stats", "kmc_violation", "kmc_trace_state", "kmc_strerror", "kmc_destroy",
              "kmc_fpset_put", "kmc_fpset_contains", "kmc_fpset_size", "kmc_shard_expand", "kmc_shard_insert"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_model_libraries_export_the_per_model_abi():
    models = os.path.join(ROOT, "build", "models")
    if not os.path.isdir(models):
        pytest.skip("models not built")
    want = {s.replace("kmc_", "kmcm_", 1) for s in declared_symbols()}
    n = 0
    for name in os.listdir(models):
        so = os.path.join(models, name, f"libkmc_{name}.so")
        if not os.path.exists(so):
            continue
        out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
        have = set(re.findall(r"\bT (kmcm_[a-z_0-9]+)", out))
        assert want <= have, (name, sorted(want - have))
        n += 1
    assert n > 0


def test_model_library_is_sm100a_device_code():
    so = os.path.join(ROOT, "build", "models", "kip320_small", "libkmc_kip320_small.so")
    if not os.path.exists(so):
        pytest.skip("model not built")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_bad_model_path_fails_loudly(lib):
    lib.kmc_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.kmc_strerror.restype = ctypes.c_char_p
    lib.kmc_strerror.argtypes = [ctypes.c_void_p, ctypes.c_int]
    ctx = ctypes.c_void_p()
    rc = lib.kmc_create(b"/nonexistent/libkmc_x.so", b"{}", ctypes.byref(ctx))
    assert rc == -7                                     # KMC_E_MODEL
    assert b"cannot load" in lib.kmc_strerror(ctx, rc)
    lib.kmc_destroy.argtypes = [ctypes.c_void_p]
    lib.kmc_destroy(ctx)


def test_no_cpu_fallback_without_gpu():
    """On a machine without a CUDA device the product path must refuse to run (KMC_E_NO_GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    so = os.path.join(ROOT, "build", "models", "idsequence", "libkmc_idsequence.so")
    if not os.path.exists(so):
        pytest.skip("model not built")
    from kafka_specification_b200.runtime import Checker, KmcError
    with pytest.raises(KmcError) as e:
        Checker("idsequence")
    assert e.value.code == -9 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """Nothing under the package may import, load or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "kafka_specification_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cpp", ".h")):
                continue
            for line in open(os.path.join(dp, f)):
                low = line.lower()
                if "oracle" not in low:
                    continue
                assert not re.search(r"\b(import|from|include|cdll|dlopen|subprocess|open)\b", low), (f, line.strip())


@pytest.mark.gpu
def test_two_gpus_behind_kmc_create_and_kmc_run_only(goldens):
    """One kmc_ctx drives several GPUs (option "gpus"): no torch, no NCCL, no second process -- host threads inside
    the library, peers mapped with cudaDeviceEnablePeerAccess, device-side round/level synchronisation."""
    import ctypes
    import json
    lib = ctypes.CDLL(os.path.join(ROOT, "build", "libkspecmc.so"))
    ndev = ctypes.c_int(0)
    try:
        cudart = ctypes.CDLL("libcudart.so")
        cudart.cudaGetDeviceCount(ctypes.byref(ndev))
    except OSError:                              # (no unversioned libcudart on the loader path: ask torch instead)
        import torch
        ndev.value = torch.cuda.device_count()
    if ndev.value < 2:
        pytest.skip("needs 2 GPUs")
    from kafka_specification_b200.runtime import Checker
    g = goldens["kip320_small"]
    with Checker("kip320_small", gpus=2, table_log2=22, cont=True) as ck:
        r = ck.run()
    assert r.complete and r.violation is None
    assert (r.distinct, r.generated, r.depth, r.deadlocks, r.levels) == (
        g["distinct"], g["generated"], g["depth"], g["deadlocks"], g["levels"])
    # a violating model: same verdict, level and a complete trace across the two stores
    g = goldens["trunchw_small"]
    first = min(l for l in g["first_violation_level"].values() if l)
    with Checker("trunchw_small", gpus=2, table_log2=22) as ck:
        r = ck.run()
    assert not r.complete and r.violation["kind"] == "invariant" and r.violation["level"] == first
    assert len(r.trace) == first and r.trace[0]["action"] is None
