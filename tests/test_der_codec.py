"""CPU tier: the host-only parts of the reference-named API (no engine, no GPU): strict DER signature parsing / serialisation,
compact serialisation and low-s normalisation, against the 463 Wycheproof vectors (190 of them invalid encodings) and against the
unmodified reference on mutated inputs."""
import ctypes
import json
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


@pytest.fixture(scope="module")
def libs():
    import secp256k1_zkp_b200 as b200
    from tests.refshim import RefLib
    mine = ctypes.CDLL(b200.library_path())
    ref = RefLib()
    for lib in (mine, ref.lib):
        lib.secp256k1_context_create.restype = ctypes.c_void_p
    mctx = ctypes.c_void_p.in_dll(mine, "secp256k1_context_static")
    rctx = ctypes.c_void_p.in_dll(ref.lib, "secp256k1_context_static")
    return (mine, mctx), (ref.lib, rctx)


def _roundtrip(lib, ctx, der):
    sig = ctypes.create_string_buffer(64)
    ok = lib.secp256k1_ecdsa_signature_parse_der(ctx, sig, der, ctypes.c_size_t(len(der)))
    comp = ctypes.create_string_buffer(64)
    lib.secp256k1_ecdsa_signature_serialize_compact(ctx, comp, sig)
    out = ctypes.create_string_buffer(80); ln = ctypes.c_size_t(80)
    s = lib.secp256k1_ecdsa_signature_serialize_der(ctx, out, ctypes.byref(ln), sig)
    small = ctypes.create_string_buffer(8); ln2 = ctypes.c_size_t(8)
    s2 = lib.secp256k1_ecdsa_signature_serialize_der(ctx, small, ctypes.byref(ln2), sig)
    return ok, sig.raw, comp.raw, s, out.raw[:ln.value], s2, ln2.value


def test_der_wycheproof_and_mutations(libs):
    (mine, mctx), (theirs, rctx) = libs
    vecs = json.load(open(os.path.join(GOLD, "wycheproof_ecdsa.json")))["vectors"]
    assert sum(1 for v in vecs if not v["der_ok"]) == 190
    rnd = random.Random(3)
    n_bad = 0
    for v in vecs:
        der = bytes.fromhex(v["der"])
        a = _roundtrip(mine, mctx, der)
        assert a[0] == v["der_ok"], v["tcId"]
        if v["der_ok"]:
            assert a[2].hex() == v["compact"], v["tcId"]
        assert a == _roundtrip(theirs, rctx, der), v["tcId"]
        n_bad += not a[0]
        for _ in range(3):                                       # byte mutations, truncations, extensions
            m = bytearray(der)
            k = rnd.randrange(4)
            if k == 0 and m: m[rnd.randrange(len(m))] ^= 1 << rnd.randrange(8)
            elif k == 1 and m: del m[rnd.randrange(len(m))]
            elif k == 2: m.insert(rnd.randrange(len(m) + 1), rnd.randrange(256))
            else: m = m[:rnd.randrange(len(m) + 1)]
            assert _roundtrip(mine, mctx, bytes(m)) == _roundtrip(theirs, rctx, bytes(m)), (v["tcId"], bytes(m).hex())
    assert n_bad == 190


def test_der_of_boundary_scalars(libs):
    (mine, mctx), (theirs, rctx) = libs
    vals = [0, 1, 0x7F, 0x80, 0xFF, 0x100, 2**255 - 1, 2**255, N - 1, N, N + 1, 2**256 - 1]
    for r in vals:
        for s in vals:
            for lib, ctx in ((mine, mctx), (theirs, rctx)):
                pass
            def enc(x):
                b = x.to_bytes((x.bit_length() + 7) // 8 or 1, "big")
                if b[0] & 0x80: b = b"\x00" + b
                return b"\x02" + bytes([len(b)]) + b
            body = enc(r) + enc(s)
            der = b"\x30" + bytes([len(body)]) + body
            assert _roundtrip(mine, mctx, der) == _roundtrip(theirs, rctx, der), (hex(r), hex(s))
