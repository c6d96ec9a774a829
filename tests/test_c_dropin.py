"""The single-call drop-in claim of INTEGRATION.md section 1, exercised from C: tests/c_dropin/dropin_test.c includes ONLY the reference's
public headers and is linked against libsecp256k1_b200.so.  CPU tier: it compiles and links (needs /root/reference/include, skipped
otherwise); GPU tier: the binary built here (it travels with the snapshot) runs and every check passes."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "c_dropin", "dropin_test.c")
BIN = os.path.join(HERE, "c_dropin", "dropin_test")
LIBDIR = os.path.join(ROOT, "secp256k1-zkp_b200")
REF_INC = os.environ.get("REF", "/root/reference") + "/include"


def test_dropin_compiles_against_reference_headers():
    if not os.path.isdir(REF_INC):
        pytest.skip("reference headers not present (GPU box)")
    assert os.path.exists(os.path.join(LIBDIR, "libsecp256k1_b200.so")), "run `python __graft_entry__.py build` first"
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-O1", "-I", REF_INC, "-o", BIN, SRC, "-L", LIBDIR, "-lsecp256k1_b200", "-Wl,-rpath,$ORIGIN/../../secp256k1-zkp_b200"]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    # every secp256k1_* symbol the program needs is satisfied by our library alone
    und = subprocess.run(["nm", "-u", BIN], capture_output=True, text=True).stdout
    needed = sorted(l.split()[-1] for l in und.splitlines() if "secp256k1_" in l)
    exported = subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIBDIR, "libsecp256k1_b200.so")], capture_output=True, text=True).stdout
    for sym in needed:
        assert sym.split("@")[0] in exported, sym
    assert len(needed) >= 12


@pytest.mark.gpu
def test_dropin_runs_on_gpu():
    if not os.path.exists(BIN):
        pytest.skip("tests/c_dropin/dropin_test was not built (run the CPU tier where the reference headers exist)")
    p = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "dropin ok" in p.stdout, (p.returncode, p.stdout, p.stderr)
