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


def test_host_only_helpers(libs):
    """secp256k1_ec_pubkey_cmp / xonly_pubkey_cmp (on opaque objects parsed by the reference: the two libraries share the object layout),
    secp256k1_ec_seckey_verify, secp256k1_tagged_sha256"""
    import hashlib
    (mine, mctx), (theirs, rctx) = libs
    vecs = json.load(open(os.path.join(GOLD, "wycheproof_ecdsa.json")))["vectors"]
    keys = sorted({v["pk65"] for v in vecs})[:12]
    objs, xobjs = [], []
    for k in keys:
        o = ctypes.create_string_buffer(64)
        assert theirs.secp256k1_ec_pubkey_parse(rctx, o, bytes.fromhex(k), ctypes.c_size_t(65)) == 1
        objs.append(o)
        x = ctypes.create_string_buffer(64)
        if theirs.secp256k1_xonly_pubkey_parse(rctx, x, bytes.fromhex(k)[1:33]) == 1:
            xobjs.append(x)
    sign = lambda v: (v > 0) - (v < 0)
    for a in objs:
        for b in objs:
            assert sign(mine.secp256k1_ec_pubkey_cmp(mctx, a, b)) == sign(theirs.secp256k1_ec_pubkey_cmp(rctx, a, b))
    for a in xobjs:
        for b in xobjs:
            assert sign(mine.secp256k1_xonly_pubkey_cmp(mctx, a, b)) == sign(theirs.secp256k1_xonly_pubkey_cmp(rctx, a, b))
    for k in (0, 1, N - 1, N, N + 1, 2**256 - 1, 12345):
        kb = k.to_bytes(32, "big")
        assert mine.secp256k1_ec_seckey_verify(mctx, kb) == theirs.secp256k1_ec_seckey_verify(rctx, kb) == (1 if 0 < k < N else 0)
    rnd = random.Random(5)
    for taglen, msglen in ((0, 0), (1, 1), (17, 32), (64, 55), (65, 56), (100, 200)):
        tag = bytes(rnd.randrange(256) for _ in range(taglen)) or b""
        msg = bytes(rnd.randrange(256) for _ in range(msglen)) or b""
        outs = []
        for lib, ctx in ((mine, mctx), (theirs, rctx)):
            o = ctypes.create_string_buffer(32)
            assert lib.secp256k1_tagged_sha256(ctx, o, tag if tag else b"\x00", ctypes.c_size_t(taglen), msg if msg else b"\x00", ctypes.c_size_t(msglen)) == 1
            outs.append(o.raw)
        th = hashlib.sha256(tag).digest()
        assert outs[0] == outs[1] == hashlib.sha256(th + th + msg).digest()
