"""CPU tier: the C-ABI library loads and exports every symbol include/secp256k1_b200.h declares;
without a GPU the product refuses to work (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    inc = os.path.join(ROOT, "include")
    for f in sorted(os.listdir(inc)):
        if f.endswith(".h"):
            text = open(os.path.join(inc, f)).read()
            names += re.findall(r"SECP256K1_B200_API\s+[\w\s\*]+?\b(secp256k1_\w+)\s*\(", text)
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    import secp256k1_zkp_b200 as b200
    path = b200.library_path()
    assert os.path.exists(path), "run `python __graft_entry__.py build` first"
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "missing export: %s" % n


def test_no_cpu_fallback():
    import torch
    import secp256k1_zkp_b200 as b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b200.B200Error):
        b200.Engine()
