"""The tlc2.TLC-compatible command line (kafka_specification_b200/tlc2.py)."""
import os
import subprocess
import sys

import pytest

from conftest import REFERENCE, ROOT, needs_reference

SPECS = os.path.join(ROOT, "tests", "specs")


def run_cli(*args, timeout=600):
    p = subprocess.run([sys.executable, "-m", "kafka_specification_b200.tlc2", *args], cwd=ROOT,
                       capture_output=True, text=True, timeout=timeout)
    return p.returncode, p.stdout + p.stderr


def test_cli_spec_errors_have_tlc_exit_codes(tmp_path):
    rc, out = run_cli("-config", str(tmp_path / "missing.cfg"), os.path.join(SPECS, "MiniLock"))
    assert rc == 150 and "cannot read the configuration file" in out
    bad = tmp_path / "Bad.tla"
    bad.write_text("---- MODULE Bad ----\nVARIABLE x\nInit == x = \nNext == x' = x\n====\n")
    (tmp_path / "Bad.cfg").write_text("INIT Init\nNEXT Next\n")
    rc, out = run_cli(str(bad))
    assert rc == 150 and "Error:" in out


@needs_reference
def test_cli_false_assume_is_exit_10(tmp_path):
    cfg = tmp_path / "A.cfg"
    cfg.write_text(open(os.path.join(ROOT, "models", "MCAsyncIsr.cfg")).read().replace("MaxOffset = 2", "MaxOffset = 0"))
    rc, out = run_cli("-config", str(cfg), "-I", REFERENCE, os.path.join(ROOT, "models", "MCAsyncIsr"))
    assert rc == 10 and "ASSUME" in out


def test_cli_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    rc, out = run_cli("-config", os.path.join(SPECS, "MiniLock.cfg"), os.path.join(SPECS, "MiniLock"))
    assert rc == 1 and "KMC_E_NO_GPU" in out and "no CPU fallback" in out


@pytest.mark.gpu
def test_cli_end_to_end_on_gpu(tmp_path):
    """.tla + .cfg in, TLC's summary lines and exit status out (lowering + nvcc + GPU run inside)."""
    rc, out = run_cli("-config", os.path.join(SPECS, "MiniLock.cfg"), "-deadlock", os.path.join(SPECS, "MiniLock"))
    assert rc == 0, out
    assert "Model checking completed. No error has been found." in out
    assert "169 states generated, 76 distinct states found, 0 states left on queue." in out
    assert "The depth of the complete state graph search is 14." in out
    # a false invariant: exit 12, TLC-style trace with action headers
    spec = tmp_path / "MiniLock.tla"
    spec.write_text(open(os.path.join(SPECS, "MiniLock.tla")).read().replace(
        "HolderNotWaiting ==", "NeverTwo == Cardinality(waiting) < 2\nHolderNotWaiting =="))
    cfg = tmp_path / "MiniLock.cfg"
    cfg.write_text(open(os.path.join(SPECS, "MiniLock.cfg")).read().replace(
        "INVARIANTS TypeOk Bounded HolderNotWaiting", "INVARIANTS TypeOk NeverTwo"))
    rc, out = run_cli("-config", str(cfg), str(spec))
    assert rc == 12, out
    assert "Error: Invariant NeverTwo is violated." in out
    assert "State 1: <Initial predicate>" in out and "State 3: <Request line" in out
    assert "/\\ waiting = {" in out
    # deadlock checking is on by default (TLC): MiniLock never deadlocks, IdSequence-style specs do
    dl = tmp_path / "Stop.tla"
    dl.write_text("---- MODULE Stop ----\nEXTENDS Integers\nVARIABLE x\nInit == x = 0\nNext == x < 3 /\\ x' = x + 1\n"
                  "TypeOk == x \\in 0 .. 3\n====\n")
    (tmp_path / "Stop.cfg").write_text("INIT Init\nNEXT Next\nINVARIANT TypeOk\n")
    rc, out = run_cli(str(dl))
    assert rc == 11 and "Error: Deadlock reached." in out and "State 4: <Next" in out
    rc, out = run_cli("-deadlock", str(dl))
    assert rc == 0 and "4 distinct states found" in out


@pytest.mark.gpu
def test_cli_tool_mode_wraps_messages_in_tlc_markers(tmp_path):
    """-tool: every message carries TLC's @!@!@STARTMSG code:class / @!@!@ENDMSG code markers; trace states are class 4."""
    dl = tmp_path / "Stop.tla"
    dl.write_text("---- MODULE Stop ----\nEXTENDS Integers\nVARIABLE x\nInit == x = 0\nNext == x < 3 /\\ x' = x + 1\n"
                  "TypeOk == x \\in 0 .. 3\n====\n")
    (tmp_path / "Stop.cfg").write_text("INIT Init\nNEXT Next\nINVARIANT TypeOk\n")
    rc, out = run_cli("-tool", str(dl))
    assert rc == 11
    assert "@!@!@STARTMSG 2114:1 @!@!@\nError: Deadlock reached.\n@!@!@ENDMSG 2114 @!@!@" in out
    assert out.count("@!@!@STARTMSG 2217:4 @!@!@") == 4 and "@!@!@STARTMSG 2199:0 @!@!@" in out
    assert out.count("@!@!@STARTMSG") == out.count("@!@!@ENDMSG")


@pytest.mark.gpu
def test_cli_workers_n_runs_on_n_gpus(tmp_path):
    """-workers 2: the library shards the search over two GPUs of this process (kmc_create option gpus)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rc, out = run_cli("-workers", "2", "-config", os.path.join(SPECS, "MiniLock.cfg"), "-deadlock", os.path.join(SPECS, "MiniLock"))
    assert rc == 0, out
    assert "on 2 GPUs" in out
    assert "169 states generated, 76 distinct states found, 0 states left on queue." in out
    assert "The depth of the complete state graph search is 14." in out
