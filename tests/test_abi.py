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


def test_halfagg_randomizer_chain_host_paths():
    """the host walker of the half-aggregation randomizer chain (SHA-extension and portable paths, no GPU needed) against the
    device-header implementation compiled for the host, and z_1 against hashlib"""
    import hashlib
    import numpy as np
    from tests import hostemu_lib
    import secp256k1_zkp_b200 as b200
    lib = ctypes.CDLL(b200.library_path())
    emu = hostemu_lib.load()
    rng = np.random.default_rng(11)
    for n in (0, 1, 2, 3, 64, 1001):
        pk, msg, agg = rng.integers(0, 256, 32 * n + 1, dtype=np.uint8).tobytes(), rng.integers(0, 256, 32 * n + 1, dtype=np.uint8).tobytes(), rng.integers(0, 256, 32 * (n + 1), dtype=np.uint8).tobytes()
        want = ctypes.create_string_buffer(32 * n + 32)
        emu.emu_halfagg_z_chain(want, pk, msg, ctypes.c_size_t(n), agg)
        for force in (0, 1):
            got = ctypes.create_string_buffer(32 * n + 32)
            used = lib.secp256k1_b200_halfagg_randomizers(got, pk, msg, ctypes.c_size_t(n), agg, force)
            assert got.raw[:32 * n] == want.raw[:32 * n], (n, force)
            assert not (force and used)
        if n >= 2:
            th = hashlib.sha256(b"HalfAgg/randomizer").digest()
            h = hashlib.sha256(th + th + b"".join(agg[32 * i:32 * i + 32] + pk[32 * i:32 * i + 32] + msg[32 * i:32 * i + 32] for i in range(2)))
            assert want.raw[32:64] == h.digest() and want.raw[:32] == (1).to_bytes(32, "big")


def test_selftest_host_part_runs_without_a_device():
    """secp256k1_selftest (src/selftest.h:12-34): the SHA-256 known-answer test runs on the host whether or not a GPU is present; a wrong
    digest would abort the process through the default error callback.  (With a B200 the same call also runs device known answers.)"""
    import secp256k1_zkp_b200 as b200
    lib = ctypes.CDLL(b200.library_path())
    lib.secp256k1_selftest.restype = None
    lib.secp256k1_selftest()
