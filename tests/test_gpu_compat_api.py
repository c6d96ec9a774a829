"""GPU tier: the reference-NAMED single-call API exported by libsecp256k1_b200.so
(include/secp256k1_b200_compat.h) against the same calls on the unmodified reference (oracle/_ref):
same arguments, same return values, same output bytes, same illegal-callback behaviour."""
import ctypes
import hashlib
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_void_p)


class Api:
    """the subset of the reference C API both libraries export, driven through ctypes"""

    def __init__(self, lib):
        self.lib = lib
        lib.secp256k1_context_create.restype = ctypes.c_void_p
        lib.secp256k1_context_create.argtypes = [ctypes.c_uint]
        self.ctx = ctypes.c_void_p(lib.secp256k1_context_create(1))
        self.static = ctypes.c_void_p.in_dll(lib, "secp256k1_context_static")
        self.illegal = 0
        self._cb = CB(self._on_illegal)
        lib.secp256k1_context_set_illegal_callback.argtypes = [ctypes.c_void_p, CB, ctypes.c_void_p]
        lib.secp256k1_context_set_illegal_callback(self.ctx, self._cb, None)
        self.gen_h = ctypes.string_at(ctypes.c_void_p.in_dll(lib, "secp256k1_generator_h").value, 64)

    def _on_illegal(self, msg, data):
        self.illegal += 1

    def xonly_parse(self, b, ctx=None):
        out = ctypes.create_string_buffer(64)
        return self.lib.secp256k1_xonly_pubkey_parse(ctx or self.ctx, out, b), out.raw

    def ec_parse(self, b):
        out = ctypes.create_string_buffer(64)
        return self.lib.secp256k1_ec_pubkey_parse(self.ctx, out, b, ctypes.c_size_t(len(b))), out.raw

    def ec_serialize(self, pk64, compressed):
        out = ctypes.create_string_buffer(65)
        ln = ctypes.c_size_t(65)
        r = self.lib.secp256k1_ec_pubkey_serialize(self.ctx, out, ctypes.byref(ln), pk64, 258 if compressed else 2)
        return r, out.raw[:ln.value]

    def schnorr_verify(self, sig, msg, pk64, ctx=None):
        return self.lib.secp256k1_schnorrsig_verify(ctx or self.ctx, sig, msg if msg else None, ctypes.c_size_t(len(msg)), pk64)

    def ecdsa_parse_compact(self, c):
        out = ctypes.create_string_buffer(64)
        return self.lib.secp256k1_ecdsa_signature_parse_compact(self.ctx, out, c), out.raw

    def ecdsa_verify(self, sig64, msg32, pk64):
        return self.lib.secp256k1_ecdsa_verify(self.ctx, sig64, msg32, pk64)

    def commitment_parse(self, c33):
        out = ctypes.create_string_buffer(64)
        return self.lib.secp256k1_pedersen_commitment_parse(self.ctx, out, c33), out.raw

    def pedersen_commit(self, blind, value, gen, ctx=None):
        out = ctypes.create_string_buffer(64)
        r = self.lib.secp256k1_pedersen_commit(ctx or self.ctx, out, blind, ctypes.c_uint64(value), gen)
        ser = ctypes.create_string_buffer(33)
        self.lib.secp256k1_pedersen_commitment_serialize(self.ctx, ser, out)
        return r, ser.raw

    def generator_roundtrip(self, g33):
        out = ctypes.create_string_buffer(64)
        r = self.lib.secp256k1_generator_parse(self.ctx, out, g33)
        ser = ctypes.create_string_buffer(33)
        if r:
            self.lib.secp256k1_generator_serialize(self.ctx, ser, out)
        return r, out.raw, ser.raw

    def rangeproof_verify(self, commit64, proof, gen, extra=b""):
        mn, mx = ctypes.c_uint64(0), ctypes.c_uint64(0)
        r = self.lib.secp256k1_rangeproof_verify(self.ctx, ctypes.byref(mn), ctypes.byref(mx), commit64, proof, ctypes.c_size_t(len(proof)),
                                                 extra if extra else None, ctypes.c_size_t(len(extra)), gen)
        return r, mn.value, mx.value

    def rangeproof_info(self, proof):
        e, m = ctypes.c_int(0), ctypes.c_int(0)
        mn, mx = ctypes.c_uint64(0), ctypes.c_uint64(0)
        r = self.lib.secp256k1_rangeproof_info(self.ctx, ctypes.byref(e), ctypes.byref(m), ctypes.byref(mn), ctypes.byref(mx), proof, ctypes.c_size_t(len(proof)))
        return r, e.value, m.value, mn.value, mx.value


@pytest.fixture(scope="module")
def apis(ref):
    import secp256k1_zkp_b200 as b200
    mine = Api(ctypes.CDLL(b200.library_path()))
    theirs = Api(ref.lib)
    yield mine, theirs
    mine.lib.secp256k1_b200_global_engine_shutdown()


def test_generator_h_constant(apis):
    mine, theirs = apis
    assert mine.gen_h == theirs.gen_h


def test_pubkey_parse_serialize_and_bip340(apis):
    mine, theirs = apis
    rnd = random.Random(1)
    for v in json.load(open(os.path.join(GOLD, "bip340.json")))["vectors"]:
        pk = bytes.fromhex(v["pk"])
        a, b = mine.xonly_parse(pk), theirs.xonly_parse(pk)
        assert a == b
        if a[0]:
            sig, msg = bytes.fromhex(v["sig"]), bytes.fromhex(v["msg"])
            assert mine.schnorr_verify(sig, msg, a[1]) == theirs.schnorr_verify(sig, msg, b[1]) == v["expected"]
            # verification works with the static context too (SURVEY 8b "context kinds")
            assert mine.schnorr_verify(sig, msg, a[1], ctx=mine.static) == v["expected"]
    for i in range(40):
        x = hashlib.sha256(b"cx%d" % i).digest()
        for enc in (bytes([2]) + x, bytes([3]) + x, bytes([4]) + x + x, bytes([5]) + x, bytes([2]) + b"\xff" * 32):
            a, b = mine.ec_parse(enc), theirs.ec_parse(enc)
            assert a == b
            if a[0]:
                for comp in (True, False):
                    assert mine.ec_serialize(a[1], comp) == theirs.ec_serialize(b[1], comp)
                # uncompressed + hybrid encodings of the same key
                r, unc = theirs.ec_serialize(b[1], False)
                hyb = bytes([6 + (unc[64] & 1)]) + unc[1:]
                bad = bytes([7 - (unc[64] & 1)]) + unc[1:]
                for e2 in (unc, hyb, bad, unc[:64] + bytes([unc[64] ^ 1])):
                    assert mine.ec_parse(e2) == theirs.ec_parse(e2)


def test_ecdsa_wycheproof_through_compat_api(apis):
    mine, theirs = apis
    vecs = [v for v in json.load(open(os.path.join(GOLD, "wycheproof_ecdsa.json")))["vectors"] if v["der_ok"]][::3]
    for v in vecs:
        pk = bytes.fromhex(v["pk65"])
        (ra, pa), (rb, pb) = mine.ec_parse(pk), theirs.ec_parse(pk)
        assert (ra, pa) == (rb, pb) and ra == 1
        c = bytes.fromhex(v["compact"])
        sa, sb = mine.ecdsa_parse_compact(c), theirs.ecdsa_parse_compact(c)
        assert sa == sb
        m = bytes.fromhex(v["msghash"])
        assert mine.ecdsa_verify(sa[1], m, pa) == theirs.ecdsa_verify(sb[1], m, pb) == v["expected"]
    over = (N).to_bytes(32, "big") + (1).to_bytes(32, "big")
    assert mine.ecdsa_parse_compact(over) == theirs.ecdsa_parse_compact(over)


def test_pedersen_generator_rangeproof_compat(apis, ref):
    mine, theirs = apis
    rnd = random.Random(2)
    gen = theirs.gen_h
    for i in range(12):
        blind = (rnd.getrandbits(256) % N).to_bytes(32, "big")
        val = rnd.getrandbits(64)
        a, b = mine.pedersen_commit(blind, val, gen), theirs.pedersen_commit(blind, val, gen)
        assert a == b and a[0] == 1
        assert mine.commitment_parse(a[1]) == theirs.commitment_parse(b[1])
        badc = bytes([a[1][0] ^ 4]) + a[1][1:]
        assert mine.commitment_parse(badc) == theirs.commitment_parse(badc)
        g33 = bytes([10 + (i & 1)]) + hashlib.sha256(b"gen%d" % i).digest()
        assert mine.generator_roundtrip(g33) == theirs.generator_roundtrip(g33)
    over = (N).to_bytes(32, "big")
    assert mine.pedersen_commit(over, 5, gen)[0] == theirs.pedersen_commit(over, 5, gen)[0] == 0
    for v in json.load(open(os.path.join(GOLD, "rangeproof.json")))["vectors"]:
        proof, c33 = bytes.fromhex(v["proof"]), bytes.fromhex(v["commit"])
        (ra, ca), (rb, cb) = mine.commitment_parse(c33), theirs.commitment_parse(c33)
        assert ra == rb == 1
        assert mine.rangeproof_verify(ca, proof, gen) == theirs.rangeproof_verify(cb, proof, gen) == (1, v["min_value"], v["max_value"])
        assert mine.rangeproof_info(proof) == theirs.rangeproof_info(proof)
        bad = bytearray(proof); bad[len(bad) // 2] ^= 1
        assert mine.rangeproof_verify(ca, bytes(bad), gen) == theirs.rangeproof_verify(cb, bytes(bad), gen)
        assert mine.rangeproof_verify(ca, proof, gen, b"xx") == theirs.rangeproof_verify(cb, proof, gen, b"xx")


def test_illegal_callback_semantics(apis):
    mine, theirs = apis
    for api in (mine, theirs):
        before = api.illegal
        out = ctypes.create_string_buffer(64)
        assert api.lib.secp256k1_xonly_pubkey_parse(api.ctx, out, None) == 0          # NULL input -> illegal callback, 0
        assert api.lib.secp256k1_schnorrsig_verify(api.ctx, None, b"", ctypes.c_size_t(0), out) == 0
        assert api.illegal == before + 2
        # an all-zero (never parsed) pubkey object is API misuse: callback + 0 (src/secp256k1.c:282)
        assert api.lib.secp256k1_schnorrsig_verify(api.ctx, bytes(64), b"", ctypes.c_size_t(0), bytes(64)) == 0
        assert api.illegal == before + 3


def test_surjection_and_tally_compat(apis, ref):
    """secp256k1_surjectionproof_parse/serialize/verify and secp256k1_pedersen_verify_tally through both libraries"""
    from tests.refshim import surjection_make
    mine, theirs = apis
    rnd = random.Random(5)

    class SJ(ctypes.Structure):
        _fields_ = [("n_inputs", ctypes.c_size_t), ("used_inputs", ctypes.c_ubyte * 32), ("data", ctypes.c_ubyte * (32 * 257))]
    for (n_in, n_use) in [(1, 1), (3, 2), (6, 3)]:
        pr, tin, tout = surjection_make(ref, n_in, n_use, rnd.randrange(n_in), hashlib.sha256(b"csj%d" % n_in).digest())
        for mu in (None, len(pr) - 3, 1):
            p2 = bytearray(pr)
            if mu is not None:
                p2[mu] ^= 2
            res = []
            for api in (mine, theirs):
                sj = SJ()
                ok = api.lib.secp256k1_surjectionproof_parse(api.ctx, ctypes.byref(sj), bytes(p2), ctypes.c_size_t(len(p2)))
                v = api.lib.secp256k1_surjectionproof_verify(api.ctx, ctypes.byref(sj), tin, ctypes.c_size_t(n_in), tout) if ok else -1
                out = ctypes.create_string_buffer(len(p2) + 8); ln = ctypes.c_size_t(len(p2) + 8)
                s = api.lib.secp256k1_surjectionproof_serialize(api.ctx, out, ctypes.byref(ln), ctypes.byref(sj)) if ok else 0
                res.append((ok, v, s, out.raw[:ln.value] if s else b""))
            assert res[0] == res[1], (n_in, mu)
            if mu is None:
                assert res[0][1] == 1 and res[0][3] == bytes(pr)
    gen = theirs.gen_h
    bl = [rnd.getrandbits(256) % N for _ in range(3)]
    cs = []
    for b, v in ((bl[0], 70), (bl[1], 30), (bl[2], 55), ((bl[0] + bl[1] - bl[2]) % N, 45)):
        r, c33 = theirs.pedersen_commit(b.to_bytes(32, "big"), v, gen)
        assert r == 1
        cs.append(c33)
    for api in (mine, theirs):
        objs = []
        for c33 in cs:
            r, o = api.commitment_parse(c33)
            assert r == 1
            objs.append(ctypes.create_string_buffer(o, 64))
        arr_t = ctypes.c_void_p * 2
        neg = arr_t(ctypes.addressof(objs[0]), ctypes.addressof(objs[1]))
        pos = arr_t(ctypes.addressof(objs[2]), ctypes.addressof(objs[3]))
        assert api.lib.secp256k1_pedersen_verify_tally(api.ctx, pos, ctypes.c_size_t(2), neg, ctypes.c_size_t(2)) == 1
        assert api.lib.secp256k1_pedersen_verify_tally(api.ctx, pos, ctypes.c_size_t(1), neg, ctypes.c_size_t(2)) == 0


def test_ecdsa_recover_compat(apis, ref):
    """secp256k1_ecdsa_recoverable_signature_{parse,serialize}_compact / _convert / secp256k1_ecdsa_recover through both libraries"""
    from tests.test_hostemu_logic import _recover_cases
    mine, theirs = apis
    ill0 = (mine.illegal, theirs.illegal)
    gold = json.load(open(os.path.join(GOLD, "recovery.json")))["vectors"]
    cases = [(bytes.fromhex(v["sig64"]), rid, bytes.fromhex(v["msg32"])) for v in gold for rid in range(4)]
    sig, rid, msg = _recover_cases(ref, 60, b"compat")
    cases += [(sig[64 * i:64 * i + 64], rid[i], msg[32 * i:32 * i + 32]) for i in range(60) if rid[i] <= 3]
    n_ok = 0
    for s64, r, m32 in cases:
        res = []
        for api in (mine, theirs):
            obj = ctypes.create_string_buffer(65)
            p = api.lib.secp256k1_ecdsa_recoverable_signature_parse_compact(api.ctx, obj, s64, r)
            pk = ctypes.create_string_buffer(64)
            rec = api.lib.secp256k1_ecdsa_recover(api.ctx, pk, obj, m32) if p else -1
            out = ctypes.create_string_buffer(64); rr = ctypes.c_int(-1)
            api.lib.secp256k1_ecdsa_recoverable_signature_serialize_compact(api.ctx, out, ctypes.byref(rr), obj)
            plain = ctypes.create_string_buffer(64)
            api.lib.secp256k1_ecdsa_recoverable_signature_convert(api.ctx, plain, obj)
            ser = api.ec_serialize(pk.raw, False)[1] if rec == 1 else b""
            res.append((p, rec, obj.raw, pk.raw, out.raw, rr.value, plain.raw, ser))
        assert res[0] == res[1], (s64.hex(), r)
        n_ok += res[0][1] == 1
    assert n_ok > 20
    assert (mine.illegal, theirs.illegal) == ill0
    obj = ctypes.create_string_buffer(65)
    for api in (mine, theirs):
        assert api.lib.secp256k1_ecdsa_recoverable_signature_parse_compact(api.ctx, obj, cases[0][0], 4) == 0
    assert (mine.illegal, theirs.illegal) == (ill0[0] + 1, ill0[1] + 1)


def test_halfagg_compat(apis, ref):
    """secp256k1_schnorrsig_aggverify with parsed secp256k1_xonly_pubkey objects through both libraries"""
    from tests.test_hostemu_logic import _halfagg_cases
    mine, theirs = apis
    ill0 = (mine.illegal, theirs.illegal)
    for n in (0, 1, 4):
        for pk, msg, agg, label in _halfagg_cases(ref, n, b"cagg%d" % n):
            res = []
            for api in (mine, theirs):
                objs = ctypes.create_string_buffer(64 * max(n, 1))
                parsed = all(api.lib.secp256k1_xonly_pubkey_parse(api.ctx, ctypes.byref(objs, 64 * i), pk[32 * i:32 * i + 32]) for i in range(n))
                r = api.lib.secp256k1_schnorrsig_aggverify(api.ctx, objs if n else None, msg if n else None, ctypes.c_size_t(n), agg, ctypes.c_size_t(len(agg))) if parsed else -1
                res.append((parsed, r))
            assert res[0] == res[1], (n, label)
            assert (label == "valid") == (res[0][1] == 1) or label.startswith("key tampered")
    assert (mine.illegal, theirs.illegal) == ill0


def test_rangeproof_rewind_compat(apis, ref):
    """secp256k1_rangeproof_rewind through both libraries: same verdict, blind, value, message, min/max"""
    from tests.test_hostemu_logic import _rewind_cases
    mine, theirs = apis
    ill0 = (mine.illegal, theirs.illegal)
    c = _rewind_cases(ref, 4, 10)
    for i in range(10):
        for variant in ("ok", "wrong nonce", "no message buffer"):
            nonce = bytearray(c["nonce"][32 * i:32 * i + 32])
            if variant == "wrong nonce":
                nonce[3] ^= 1
            res = []
            for api in (mine, theirs):
                r, cobj = api.commitment_parse(c["commit"][i].tobytes())
                assert r == 1
                blind = ctypes.create_string_buffer(32); val = ctypes.c_uint64(0); mn = ctypes.c_uint64(0); mx = ctypes.c_uint64(0)
                msg = ctypes.create_string_buffer(4096); ml = ctypes.c_size_t(4096)
                nomsg = variant == "no message buffer"
                proof = c["proofs"][i, :int(c["plen"][i])].tobytes()
                ret = api.lib.secp256k1_rangeproof_rewind(api.ctx, blind, ctypes.byref(val), None if nomsg else msg, None if nomsg else ctypes.byref(ml), bytes(nonce),
                                                          ctypes.byref(mn), ctypes.byref(mx), cobj, proof, ctypes.c_size_t(len(proof)), None, ctypes.c_size_t(0), api.gen_h)
                res.append((ret, blind.raw, val.value, msg.raw[:ml.value] if (ret and not nomsg) else b"", mn.value, mx.value) if ret else (ret, mn.value, mx.value))
            assert res[0] == res[1], (i, variant)
            assert (res[0][0] == 1) == (variant != "wrong nonce")
    assert (mine.illegal, theirs.illegal) == ill0


def test_ecdsa_signature_normalize_and_serialize_compat(apis):
    """secp256k1_ecdsa_signature_normalize / _serialize_compact (host-only byte arithmetic) through both libraries"""
    mine, theirs = apis
    rnd = random.Random(9)
    half = N // 2
    svals = [1, half, half + 1, N - 1, 2, N - 2] + [rnd.randrange(1, N) for _ in range(40)]
    for sv in svals:
        c = rnd.randrange(1, N).to_bytes(32, "big") + sv.to_bytes(32, "big")
        res = []
        for api in (mine, theirs):
            ok, obj = api.ecdsa_parse_compact(c)
            out = ctypes.create_string_buffer(64)
            r = api.lib.secp256k1_ecdsa_signature_normalize(api.ctx, out, obj)
            r2 = api.lib.secp256k1_ecdsa_signature_normalize(api.ctx, None, obj)
            ser = ctypes.create_string_buffer(64)
            api.lib.secp256k1_ecdsa_signature_serialize_compact(api.ctx, ser, out)
            res.append((ok, r, r2, out.raw, ser.raw))
        assert res[0] == res[1], hex(sv)
        assert res[0][1] == (1 if sv > half else 0)
        assert int.from_bytes(res[0][4][32:], "big") == (N - sv if sv > half else sv)
