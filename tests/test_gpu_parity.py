"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI vs the reference
compiled into oracle/_ref, bit-exact on serialized outputs."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
G64 = bytes.fromhex("79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798"
                    "483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8")


def b32(x):
    return int(x).to_bytes(32, "big")


def ints(arr):
    return [int.from_bytes(bytes(r), "big") for r in arr]


FE_EDGE = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2, 2**32 + 977, 2**32 + 976, 2**255, P - 2, 2**128, 2**224 - 1,
           (1 << 256) - (1 << 32), 0xFFFFFFFF, 0xFFFFFFFF00000000]


def test_fe_ops_vs_bigint_and_portable(engine):
    import secp256k1_zkp_b200 as b200
    rnd = random.Random(7)
    vals = FE_EDGE + [rnd.getrandbits(256) for _ in range(4000)] + [rnd.getrandbits(256) | (2**256 - 2**200) for _ in range(200)]
    a = [rnd.choice(vals) for _ in range(20000)] + [x for x in FE_EDGE for _ in FE_EDGE]
    b = [rnd.choice(vals) for _ in range(20000)] + [y for _ in FE_EDGE for y in FE_EDGE]
    A = b"".join(b32(x) for x in a)
    B = b"".join(b32(x) for x in b)
    inv2 = pow(2, -1, P)
    checks = [
        (b200.FE_MUL, lambda x, y: x * y % P), (b200.FE_SQR, lambda x, y: x * x % P),
        (b200.FE_ADD, lambda x, y: (x + y) % P), (b200.FE_SUB, lambda x, y: (x - y) % P),
        (b200.FE_HALF, lambda x, y: x * inv2 % P), (b200.FE_NEG, lambda x, y: (-x) % P),
        (b200.FE_MUL_PORTABLE, lambda x, y: x * y % P), (b200.FE_SQR_PORTABLE, lambda x, y: x * x % P),
    ]
    for op, f in checks:
        got = ints(engine.fe_op(op, A, B))
        exp = [f(x, y) for x, y in zip(a, b)]
        assert got == exp, "fe op %d mismatch" % op
    # iterated (dependent chain) multiply: r = a * b^k
    got = ints(engine.fe_op(b200.FE_MUL, A[:32 * 512], B[:32 * 512], iters=9))
    assert got == [x * pow(y, 9, P) % P for x, y in zip(a[:512], b[:512])]
    sub = 2048
    got = ints(engine.fe_op(b200.FE_INV, A[:32 * sub]))
    assert got == [pow(x, P - 2, P) for x in a[:sub]]
    got = ints(engine.fe_op(b200.FE_SQRT, A[:32 * sub]))
    assert got == [pow(x, (P + 1) // 4, P) for x in a[:sub]]


def _points(ref, ks):
    n = len(ks)
    return ref.ecmult_batch(G64 * n, b"".join(b32(0) for _ in ks), b"".join(b32(k) for k in ks))


def test_ecmult_batch_vs_reference(engine, ref):
    rnd = random.Random(11)
    lam = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
    edge = [0, 1, 2, 3, N - 1, N - 2, 2**128, 2**128 - 1, 2**127, N // 2, N // 2 + 1, lam, N - lam, lam + 1, 2**255, 15, 16, 17]
    n = 6000
    na = [rnd.choice(edge) if i % 5 == 0 else rnd.getrandbits(256) % N for i in range(n)]
    ng = [rnd.choice(edge) if i % 7 == 0 else rnd.getrandbits(256) % N for i in range(n)]
    pk = [rnd.choice([1, 2, 3, 5, N - 1, N - 2]) if i % 11 == 0 else rnd.getrandbits(256) % N for i in range(n)]
    pts = _points(ref, pk).copy()
    for i in range(0, n, 97):
        pts[i] = 0  # point at infinity
    # structured cases: na*A cancels ng*G  (A = k*G, na = 1, ng = N-k) -> infinity
    for i in range(1, n, 101):
        na[i] = 1
        ng[i] = (N - pk[i]) % N
    NA = b"".join(b32(x) for x in na)
    NG = b"".join(b32(x) for x in ng)
    exp = ref.ecmult_batch(pts, NA, NG)
    got = engine.ecmult(pts, NA, NG)
    assert np.array_equal(got, exp)
    assert not exp[1].any()          # the cancelling case really is infinity
    # ng = NULL path
    exp = ref.ecmult_batch(pts[:512], NA[:32 * 512], None)
    got = engine.ecmult(pts[:512], NA[:32 * 512], None)
    assert np.array_equal(got, exp)


def _schnorr_batch(ref, n, seed):
    sk = b"".join(hashlib.sha256(b"sk%d-%d" % (seed, i)).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"msg%d-%d" % (seed, i)).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    msg = np.frombuffer(msg, dtype=np.uint8).copy()
    for i in range(n):
        if i % 3 == 1:
            sig[64 * i + (i * 7) % 64] ^= 1 << (i % 8)
        if i % 7 == 2:
            msg[32 * i + i % 32] ^= 0x10
        if i % 11 == 3:
            pk[32 * i + i % 32] ^= 0x01
        if i % 53 == 5:
            sig[64 * i:64 * i + 32] = 0xFF      # r >= p
        if i % 59 == 6:
            sig[64 * i + 32:64 * i + 64] = 0xFF  # s >= n
        if i % 61 == 7:
            pk[32 * i:32 * i + 32] = 0xFF        # x >= p
    return sig, msg, pk


def test_schnorrsig_verify_batch_vs_reference(engine, ref):
    import secp256k1_zkp_b200 as b200
    n = 20000
    sig, msg, pk = _schnorr_batch(ref, n, 1)
    exp = ref.schnorr_verify(sig, msg, pk)
    got = engine.schnorrsig_verify(sig, msg, pk, b200.PK_XONLY32)
    assert np.array_equal(got, exp)
    assert 0.3 < exp.mean() < 0.7


def test_ecdsa_verify_batch_vs_reference(engine, ref):
    import secp256k1_zkp_b200 as b200
    n = 8000
    sk = b"".join(hashlib.sha256(b"esk%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"emsg%d" % i).digest() for i in range(n))
    sig, pk = ref.ecdsa_make(sk, msg)
    msg = np.frombuffer(msg, dtype=np.uint8).copy()
    for i in range(n):
        if i % 3 == 1:
            sig[64 * i + (i * 7) % 64] ^= 1 << (i % 8)
        if i % 7 == 2:
            msg[32 * i + i % 32] ^= 0x10
        if i % 11 == 3:
            pk[33 * i + 1 + i % 32] ^= 0x01
        if i % 13 == 4:
            s = int.from_bytes(bytes(sig[64 * i + 32:64 * i + 64]), "big")
            sig[64 * i + 32:64 * i + 64] = np.frombuffer(b32((N - s) % N), dtype=np.uint8)
        if i % 17 == 5:
            pk[33 * i] ^= 0x01   # other parity
        if i % 97 == 6:
            sig[64 * i:64 * i + 32] = 0
    exp = ref.ecdsa_verify(sig, msg, pk)
    got = engine.ecdsa_verify(sig, msg, pk, b200.SIG_COMPACT64, b200.PK_COMPRESSED33)
    assert np.array_equal(got, exp)
    assert 0.2 < exp.mean() < 0.7


def test_ecdsa_recover_batch_vs_reference(engine, ref):
    """secp256k1_b200_ecdsa_recover_batch: bit-exact recovered keys and accept set vs the reference, incl. the reference's edge-case vectors"""
    from tests.test_hostemu_logic import _recover_cases
    gold = json.load(open(os.path.join(GOLD, "recovery.json")))["vectors"]
    sig = b"".join(bytes.fromhex(v["sig64"]) * 4 for v in gold)
    msg = b"".join(bytes.fromhex(v["msg32"]) * 4 for v in gold)
    rid = bytes([0, 1, 2, 3] * len(gold))
    ok, _ = engine.ecdsa_recover(sig, rid, msg)
    assert ok.tolist() == [x for v in gold for x in v["recover_ok"]]
    n = 6000
    sig, rid, msg = _recover_cases(ref, n, b"gpu")
    eok, epk = ref.ecdsa_recover(sig, rid, msg)
    ok, pk = engine.ecdsa_recover(sig, rid, msg)
    assert np.array_equal(ok, eok)
    assert np.array_equal(pk, epk)
    assert 0.2 < eok.mean() < 0.9
    # round trip at scale: every recovered key verifies its signature (normalised to low-s as secp256k1_ecdsa_verify demands)
    import secp256k1_zkp_b200 as b200
    good = np.nonzero(ok)[0]
    s_arr = np.frombuffer(sig, dtype=np.uint8).reshape(n, 64)[good].copy()
    for j in range(len(good)):
        sv = int.from_bytes(bytes(s_arr[j, 32:]), "big")
        if sv > N // 2:
            s_arr[j, 32:] = np.frombuffer(b32(N - sv), dtype=np.uint8)
    comp = np.concatenate([(2 + (pk[good, 63] & 1))[:, None].astype(np.uint8), pk[good, :32]], axis=1)
    m_arr = np.frombuffer(msg, dtype=np.uint8).reshape(n, 32)[good]
    assert engine.ecdsa_verify(s_arr, m_arr, comp, b200.SIG_COMPACT64, b200.PK_COMPRESSED33).all()


def test_halfagg_verify_vs_reference(engine, ref):
    """secp256k1_b200_schnorrsig_aggverify (host chain + MSM) and the _dev entry point (device chain) vs the reference: spec vectors,
    valid aggregates of several sizes, every tampering of _halfagg_cases"""
    import torch
    from tests.test_hostemu_logic import _halfagg_cases

    def dev_verify(pk, msg, agg):
        n = len(pk) // 32
        if len(agg) != 32 * (n + 1):
            return None                      # the _dev call has no length argument: the caller guarantees 32*(n+1) bytes
        t = lambda b: torch.frombuffer(bytearray(b if len(b) else b"\0"), dtype=torch.uint8).cuda()
        d_pk, d_msg, d_agg = t(pk), t(msg), t(agg)
        wsb = engine.schnorrsig_aggverify_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        ok = torch.zeros(1, dtype=torch.uint8, device="cuda")
        engine.schnorrsig_aggverify_dev(ok.data_ptr(), d_pk.data_ptr(), d_msg.data_ptr(), n, d_agg.data_ptr(), ws.data_ptr(), wsb, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return int(ok[0])
    for v in json.load(open(os.path.join(GOLD, "halfagg.json")))["vectors"]:
        pk, msg, agg = bytes.fromhex(v["pubkeys"]), bytes.fromhex(v["msgs32"]), bytes.fromhex(v["aggsig"])
        assert engine.schnorrsig_aggverify(pk, msg, agg) == v["expected"]
        assert dev_verify(pk, msg, agg) == v["expected"]
    seen = set()
    for n in (0, 1, 2, 33, 300):
        for pk, msg, agg, label in _halfagg_cases(ref, n, b"gpu%d" % n):
            exp = ref.aggverify(pk, msg, agg)
            want = 1 if exp == 1 else 0
            assert engine.schnorrsig_aggverify(pk, msg, agg) == want, (n, label)
            d = dev_verify(pk, msg, agg)
            assert d is None or d == want, (n, label)
            seen.add(exp)
    assert seen >= {0, 1}


def test_halfagg_verify_large(engine, ref):
    """one aggregate over 20000 signatures (reference-signed, device-side randomizer chain skipped: host chain + 40001-term MSM)"""
    n = 20000
    sk = b"".join(hashlib.sha256(b"bigagg%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"bigaggm%d" % i).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    agg = ref.aggregate(pk.tobytes(), msg, sig.tobytes())
    assert engine.schnorrsig_aggverify(pk.tobytes(), msg, agg) == 1
    bad = bytearray(agg); bad[32 * 12345 + 7] ^= 0x20
    assert engine.schnorrsig_aggverify(pk.tobytes(), msg, bytes(bad)) == 0
    bad = bytearray(msg); bad[32 * 19999] ^= 1
    assert engine.schnorrsig_aggverify(pk.tobytes(), bytes(bad), agg) == 0


def test_rangeproof_rewind_batch(engine, ref):
    """secp256k1_b200_rangeproof_rewind_batch vs the reference: reproducible vectors (nonce -> blind, value, 0xFF message), assorted shapes with
    messages, wrong nonce, tampered proof, short / absent message buffer; then 4096 64-bit proofs of the fixture (known value and blind)"""
    from tests.test_hostemu_logic import _rewind_cases
    gen = ref.generator_h()
    gold = json.load(open(os.path.join(GOLD, "rangeproof_rewind.json")))["vectors"]
    stride = 5134
    pr = np.zeros((len(gold), stride), dtype=np.uint8)
    for i, v in enumerate(gold):
        b = bytes.fromhex(v["proof"]); pr[i, :len(b)] = np.frombuffer(b, dtype=np.uint8)
    ok, bl, val, msg, ml, mn, mx = engine.rangeproof_rewind(b"".join(bytes.fromhex(v["nonce"]) for v in gold), b"".join(bytes.fromhex(v["commit"]) for v in gold),
                                                            pr, [len(v["proof"]) // 2 for v in gold], gen)
    for i, v in enumerate(gold):
        assert ok[i] == 1 and bl[i].tobytes().hex() == v["blind"] and int(val[i]) == int(v["value"]) and int(ml[i]) >= v["message_len"], v["name"]
        assert msg[i, :v["message_len"]].tobytes() == bytes([255]) * v["message_len"] and int(mn[i]) == int(v["min_value"]) and int(mx[i]) == int(v["max_value"])
    n = 48
    c = _rewind_cases(ref, 3, n)
    nonce = bytearray(c["nonce"]); nonce[32 * 7 + 1] ^= 0x80
    proofs = c["proofs"].copy(); proofs[11, 40] ^= 2; proofs[12, int(c["plen"][12]) - 33] ^= 1
    for cap in (4096, 33, 0):
        eok, ebl, eval_, emsg, emn, emx = ref.rangeproof_rewind(bytes(nonce), c["commit"], proofs, c["plen"], c["gen"], msg_cap=cap)
        ok, bl, val, msg, ml, mn, mx = engine.rangeproof_rewind(bytes(nonce), c["commit"].tobytes(), proofs, c["plen"], c["gen"], msg_cap=cap)
        assert np.array_equal(ok, eok) and eok[7] == 0 and eok[11] == 0 and eok[12] == 0 and eok.sum() == n - 3
        for i in range(n):
            if ok[i]:
                assert bl[i].tobytes() == ebl[i].tobytes() and val[i] == eval_[i] and mn[i] == emn[i] and mx[i] == emx[i], (cap, i)
                assert (msg[i, :ml[i]].tobytes() if cap else b"") == emsg[i], (cap, i)
                if cap == 4096:
                    assert emsg[i][:len(c["msgs"][i])] == c["msgs"][i]
            else:
                assert not bl[i].any() and val[i] == 0 and ml[i] == 0
    fx = np.load(os.path.join(GOLD, "rangeproof64_fixture.npz"))
    reps = 64
    proofs = np.tile(fx["proofs"], (reps, 1)); commits = np.tile(fx["commits"], (reps, 1))
    nonces = commits[:, :32].copy()                      # the fixture was signed with nonce = commitment bytes (bench_rangeproof.c:32-34)
    bad = np.arange(5, 64 * reps, 97)
    nonces[bad, 9] ^= 1
    ok, bl, val, msg, ml, mn, mx = engine.rangeproof_rewind(nonces.tobytes(), commits.tobytes(), proofs, np.full(64 * reps, 5126), fx["gen64"], msg_cap=64)
    exp_ok = np.ones(64 * reps, dtype=np.uint8); exp_ok[bad] = 0
    assert np.array_equal(ok, exp_ok)
    good = ok.astype(bool)
    assert np.array_equal(val[good], np.tile(fx["values"], reps)[good]) and np.array_equal(bl[good], np.tile(fx["blinds"], (reps, 1))[good])


# ---- the reference's own golden vectors through the C ABI ------------------------------------------
import ctypes
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_bip340_vectors_gpu(engine):
    import secp256k1_zkp_b200 as b200
    vecs = json.load(open(os.path.join(GOLD, "bip340.json")))["vectors"]
    for v in vecs:
        pk = bytes.fromhex(v["pk"])
        if v["pk_parse_fails"]:
            assert int(engine.schnorrsig_verify(bytes(64), bytes(32), pk, b200.PK_XONLY32)[0]) == 0
            continue
        msg = bytes.fromhex(v["msg"])
        got = int(engine.schnorrsig_verify(bytes.fromhex(v["sig"]), msg, pk, b200.PK_XONLY32, msglen=len(msg))[0])
        assert got == v["expected"], v["index"]


def test_wycheproof_ecdsa_vectors_gpu(engine):
    import secp256k1_zkp_b200 as b200
    vecs = [v for v in json.load(open(os.path.join(GOLD, "wycheproof_ecdsa.json")))["vectors"] if v["der_ok"]]
    sig = b"".join(bytes.fromhex(v["compact"]) for v in vecs)
    msg = b"".join(bytes.fromhex(v["msghash"]) for v in vecs)
    pk65 = [bytes.fromhex(v["pk65"]) for v in vecs]
    opaque = b"".join(p[1:33][::-1] + p[33:65][::-1] for p in pk65)
    comp = b"".join(bytes([2 + (p[64] & 1)]) + p[1:33] for p in pk65)
    exp = np.array([v["expected"] for v in vecs], dtype=np.uint8)
    assert np.array_equal(engine.ecdsa_verify(sig, msg, opaque, b200.SIG_COMPACT64, b200.PK_OPAQUE64), exp)
    assert np.array_equal(engine.ecdsa_verify(sig, msg, comp, b200.SIG_COMPACT64, b200.PK_COMPRESSED33), exp)


def _digest(points):
    acc = hashlib.sha256()
    for xy in points:
        xy = bytes(xy)
        acc.update(b"\0" if xy == bytes(64) else b"\x04" + xy)
    return acc.hexdigest()


def test_ecmult_constants_digests_gpu(engine):
    """run_ecmult_constants (src/tests.c:5800-5960), all three scalar families, each computed two ways
    (comb table for G; GLV/fixed-window path with A = G)."""
    g = json.load(open(os.path.join(GOLD, "ecmult.json")))["constants"]
    fam = {}
    for key in ("sha_6bit20", "sha_8bit8"):
        c = g[key]
        fam[key] = ([0, 1, N - 1] + [int.from_bytes(hashlib.sha256(c["prefix"].to_bytes(4, "little") + i.to_bytes(2, "little")).digest(), "big") % N
                                      for i in range(c["iter"])], c["digest"])
    two = []
    for i in range(37):
        two += [i, (N - i) % N]
    for i in range(256):
        for j in range(1, 256, 2):
            two.append((j << i) % N)
    fam["2bit"] = (two, g["2bit"])
    for key, (scalars, digest) in fam.items():
        S = b"".join(b32(x) for x in scalars)
        n = len(scalars)
        via_g = engine.ecmult(bytes(64 * n), bytes(32 * n), S)
        via_a = engine.ecmult(G64 * n, S, bytes(32 * n))
        assert _digest(via_g) == digest, key
        assert _digest(via_a) == digest, key


def test_ecmult_chain_endpoint_gpu(engine):
    c = json.load(open(os.path.join(GOLD, "ecmult.json")))["chain"]
    xn, gn, xf, gf = int(c["xn"], 16), int(c["gn"], 16), int(c["xf"], 16), int(c["gf"], 16)
    ae, ge = 1, 0
    for _ in range(c["iterations"]):
        ae = ae * xn % N
        ge = (ge * xn + gn) % N
        xn = xn * xf % N
        gn = gn * gf % N
    out = engine.ecmult(bytes.fromhex(c["a_xy"]), b32(ae), b32(ge))
    assert bytes(out[0]).hex() == c["expected_xy"].lower()
    # and 64 real chain steps, one dependent launch each, against the reference
    x = bytes.fromhex(c["a_xy"])
    xn, gn = int(c["xn"], 16), int(c["gn"], 16)
    for _ in range(64):
        x = bytes(engine.ecmult(x, b32(xn), b32(gn))[0])
        xn = xn * xf % N
        gn = gn * gf % N
    ae, ge = 1, 0
    xn, gn = int(c["xn"], 16), int(c["gn"], 16)
    for _ in range(64):
        ae = ae * xn % N
        ge = (ge * xn + gn) % N
        xn = xn * xf % N
        gn = gn * gf % N
    assert x == bytes(engine.ecmult(bytes.fromhex(c["a_xy"]), b32(ae), b32(ge))[0])


def test_testdata_signer_matches_reference(engine, ref):
    n = 3000
    sk = b"".join(hashlib.sha256(b"tsk%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"tmsg%d" % i).digest() for i in range(n))
    sig, pk = engine.testdata_schnorrsig_sign(sk, msg)
    rsig, rpk = ref.schnorr_make(sk, msg)
    assert np.array_equal(sig, rsig) and np.array_equal(pk, rpk)


def test_full_size_batch_properties(engine):
    """2^20-item batch (BASELINE size): signer -> verify round trip, then every 16th item corrupted
    must be exactly the rejected set (size-independent property, no oracle needed)."""
    import secp256k1_zkp_b200 as b200
    n = 1 << 20
    rnd = np.random.default_rng(5)
    sk = rnd.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sk[:, 0] &= 0x7F
    sk[:, 31] |= 1
    msg = rnd.integers(0, 256, size=(n, 32), dtype=np.uint8)
    sig, pk = engine.testdata_schnorrsig_sign(sk, msg)
    ok = engine.schnorrsig_verify(sig, msg, pk, b200.PK_XONLY32)
    assert ok.all()
    sig = sig.reshape(n, 64).copy()
    bad = np.arange(7, n, 16)
    bit = bad % 512
    sig[bad, bit // 8] ^= (1 << (bit % 8)).astype(np.uint8)
    ok = engine.schnorrsig_verify(sig, msg, pk, b200.PK_XONLY32)
    exp = np.ones(n, dtype=np.uint8)
    exp[bad] = 0
    assert np.array_equal(ok, exp)


# ---- Pedersen commitments and range proofs ----------------------------------------------------------
def test_pedersen_commit_batch_vs_reference(engine, ref):
    rnd = random.Random(21)
    gen = ref.generator_h()
    n = 3000
    vals = [0, 1, 2**64 - 1, 2**63] + [rnd.getrandbits(64) for _ in range(n - 4)]
    blinds = [0, 1, N - 1, N, 2**256 - 1] + [rnd.getrandbits(256) % N for _ in range(n - 5)]   # incl. blind = 0 (allowed) and >= n (rejected)
    B = b"".join(b32(x) for x in blinds)
    okr, cr = ref.pedersen_commit(B, vals, gen)
    okg, cg = engine.pedersen_commit(B, vals, gen)
    assert np.array_equal(okg, okr) and np.array_equal(cg, cr)
    assert okr[3] == 0 and okr[4] == 0 and okr[0] == 0   # 0*G + 0*H = infinity is rejected too
    # per-item generators: a few other valid generators = serialized points k*G as 64-byte x||y
    gens = ref.ecmult_batch(G64 * 8, bytes(32 * 8), b"".join(b32(rnd.getrandbits(200) + 2) for _ in range(8))).tobytes()
    for k in range(8):
        g = gens[64 * k:64 * k + 64]
        okr, cr = ref.pedersen_commit(B[:32 * 64], vals[:64], g)
        okg, cg = engine.pedersen_commit(B[:32 * 64], vals[:64], g)
        assert np.array_equal(okg, okr) and np.array_equal(cg, cr)


def _rp_check(engine, ref, proofs, plen, commits, gen, extra=b""):
    er = ref.rangeproof_verify(commits.tobytes(), proofs, plen, gen, extra)
    if extra:
        n = proofs.shape[0]
        ex = np.frombuffer(extra * n, dtype=np.uint8).reshape(n, len(extra))
        eg = engine.rangeproof_verify(commits, proofs, plen, gen, ex, [len(extra)] * n)
    else:
        eg = engine.rangeproof_verify(commits, proofs, plen, gen)
    for a, b_ in zip(er, eg):
        assert np.array_equal(a, b_), (a[:8], b_[:8])
    return er[0]


def test_rangeproof_golden_vectors_gpu(engine, ref):
    gen = ref.generator_h()
    vecs = json.load(open(os.path.join(GOLD, "rangeproof.json")))["vectors"]
    stride = max(len(v["proof"]) // 2 for v in vecs) + 1
    for v in vecs:
        proof = np.frombuffer(bytes.fromhex(v["proof"]), dtype=np.uint8)
        commit = np.frombuffer(bytes.fromhex(v["commit"]), dtype=np.uint8)
        buf = np.zeros((1, stride), dtype=np.uint8)
        buf[0, :proof.size] = proof
        ok, mn, mx = engine.rangeproof_verify(commit, buf, [proof.size], gen)
        assert (int(ok[0]), int(mn[0]), int(mx[0])) == (1, v["min_value"], v["max_value"]), v["name"]
        ok, _, _ = engine.rangeproof_verify(commit, buf, [proof.size + 1], gen)      # trailing byte
        assert int(ok[0]) == 0


def test_rangeproof_every_bit_flip_of_one_proof(engine, ref):
    """rangeproof/tests_impl.h:306-310: every single bit of a proof flipped must fail (here: all bits of a
    3-bit proof and of the golden 64-bit proof's first 128 and last 64 bytes), min/max outputs compared too."""
    gen = ref.generator_h()
    blind = b32(0x1234567890ABCDEF1234567890ABCDEF)
    proofs, plen, commits = ref.rangeproof_make(blind, [5], gen, min_value=0, exp=0, min_bits=3)
    L = int(plen[0])
    n = 8 * L
    batch = np.repeat(proofs[:1, :L], n, axis=0).copy()
    idx = np.arange(n)
    batch[idx, idx // 8] ^= (1 << (idx % 8)).astype(np.uint8)
    ok = _rp_check(engine, ref, batch, [L] * n, np.repeat(commits[:1], n, axis=0), gen)
    assert not ok.any()
    v = [x for x in json.load(open(os.path.join(GOLD, "rangeproof.json")))["vectors"] if x["name"] == "reproducible_0"][0]
    proof = np.frombuffer(bytes.fromhex(v["proof"]), dtype=np.uint8)
    commit = np.frombuffer(bytes.fromhex(v["commit"]), dtype=np.uint8).reshape(1, 33)
    pos = np.concatenate([np.arange(0, 128 * 8), np.arange((proof.size - 64) * 8, proof.size * 8)])
    batch = np.repeat(proof.reshape(1, -1), pos.size, axis=0).copy()
    k = np.arange(pos.size)
    batch[k, pos // 8] ^= (1 << (pos % 8)).astype(np.uint8)
    ok = _rp_check(engine, ref, batch, [proof.size] * pos.size, np.repeat(commit, pos.size, axis=0), gen)
    assert not ok.any()


def test_rangeproof_random_parameters_vs_reference(engine, ref):
    rnd = random.Random(33)
    gen = ref.generator_h()
    other_gen = ref.ecmult_batch(G64, bytes(32), b32(0xC0FFEE)).tobytes()
    cases = [(0, 0, 64), (13, 1, 8), (100, 2, 13), (0, 0, 1), (0, 3, 0), (7, 0, 17), (0, 18, 3), (1, 0, 63)]
    for (mv, ex, mb) in cases:
        for g in (gen, other_gen):
            m = 12
            span = 1 if mb == 0 else min(2**mb, 2**62)
            vals = [mv + (rnd.randrange(span) * (10**ex)) % (2**63) for _ in range(m)]
            blinds = b"".join(b32(rnd.getrandbits(256) % N or 1) for _ in range(m))
            try:
                proofs, plen, commits = ref.rangeproof_make(blinds, vals, g, min_value=mv, exp=ex, min_bits=mb)
            except AssertionError:
                continue
            proofs = proofs.copy()
            for i in range(m):
                if i % 3 == 1:
                    p = rnd.randrange(int(plen[i]))
                    proofs[i, p] ^= 1 << rnd.randrange(8)
                if i % 5 == 2:
                    commits[i, 1 + rnd.randrange(32)] ^= 1
            ok = _rp_check(engine, ref, proofs, plen, commits, g)
            ok2 = _rp_check(engine, ref, proofs, plen, commits, g, extra=b"extra-commit-data")
            assert not ok2.any()            # proofs were made without extra_commit
            assert ok.sum() >= m // 3


def test_rangeproof_64bit_batch_vs_reference(engine, ref):
    """SURVEY 8d config 4 shape (64-bit proofs, 5126 bytes), small batch with 1/8 corrupted."""
    rnd = random.Random(44)
    gen = ref.generator_h()
    n = 256
    vals = [rnd.getrandbits(64) for _ in range(n)]
    blinds = b"".join(b32(rnd.getrandbits(256) % N or 1) for _ in range(n))
    proofs, plen, commits = ref.rangeproof_make(blinds, vals, gen, min_value=0, exp=0, min_bits=64)
    assert int(plen[0]) == 5126
    for i in range(3, n, 8):
        bit = int.from_bytes(hashlib.sha256(b"%d" % i).digest()[:4], "big") % (8 * 5126)
        proofs[i, bit // 8] ^= 1 << (bit % 8)
    ok = _rp_check(engine, ref, proofs, plen, commits, gen)
    assert int(ok.sum()) == n - len(range(3, n, 8))
    okc, cg = engine.pedersen_commit(blinds, vals, gen)
    assert okc.all() and np.array_equal(cg, commits)


# ---- multi-scalar multiplication ---------------------------------------------------------------------
def test_ecmult_multi_vs_reference(engine, ref):
    rnd = random.Random(55)
    for m in (0, 1, 2, 87, 88, 300, 5000):
        ks = [rnd.getrandbits(256) % N for _ in range(m)]
        pts = bytearray(ref.ecmult_batch(G64 * m, bytes(32 * m), b"".join(b32(k) for k in ks)).tobytes()) if m else bytearray()
        scs = [rnd.choice([0, 1, N - 1, 2**255, 32768, 32769, 65535 << 16, N - 32768]) if i % 4 == 0 else rnd.getrandbits(256) % N for i in range(m)]
        if m >= 87:
            pts[64 * 3:64 * 4] = bytes(64)                       # point at infinity
            pts[64 * 5:64 * 6] = pts[64 * 4:64 * 5]              # cancelling pair
            scs[5] = (N - scs[4]) % N
            for i in range(10, 40):                              # many points in ONE bucket (same scalar)
                scs[i] = scs[10]
        sc = b"".join(b32(x) for x in scs)
        for g in (None, b32(rnd.getrandbits(256) % N), b32(0)):
            ok, exp = ref.ecmult_multi(sc, bytes(pts), g, scratch_bytes=1 << 24)
            assert ok == 1
            assert engine.ecmult_multi(sc, bytes(pts), g) == exp, (m, g is None)


def test_ecmult_multi_2pow20_bench_identity(engine, ref):
    """bench_ecmult's teardown identity at the BASELINE size (src/bench_ecmult.c:73-93, :250-260): with P_i = sk_i*G the
    MSM must equal (sum s_i*sk_i mod n)*G -- checked with scalars only, no oracle MSM needed -- plus a 2^15-term prefix
    against the reference's own Pippenger."""
    n = 1 << 20
    rng = np.random.default_rng(77)
    sk = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); sk[:, 0] &= 0x7F
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); s[:, 0] &= 0x7F
    pts = engine.ecmult(np.zeros((n, 64), dtype=np.uint8), np.zeros((n, 32), dtype=np.uint8), sk)       # sk_i * G
    got = engine.ecmult_multi(s, pts)
    ski = [int.from_bytes(bytes(r), "big") for r in sk]
    si = [int.from_bytes(bytes(r), "big") for r in s]
    tot = sum(a * b for a, b in zip(ski, si)) % N
    exp = bytes(engine.ecmult(bytes(64), bytes(32), b32(tot))[0])
    assert got == exp
    m = 1 << 15
    ok, exp = ref.ecmult_multi(s[:m].tobytes(), pts[:m].tobytes(), None, scratch_bytes=64 << 20)
    assert ok == 1 and engine.ecmult_multi(s[:m], pts[:m]) == exp


# ---- BP++ norm-product argument ---------------------------------------------------------------------------
def test_bppp_golden_vectors_gpu(engine):
    d = json.load(open(os.path.join(GOLD, "bppp_verify.json")))
    gens = bytes.fromhex(d["gens_xy"])
    for v in d["vectors"]:
        g, h = v["g_len"], v["h_len"]
        pr = np.frombuffer(bytes.fromhex(v["proof"]), dtype=np.uint8).reshape(1, -1)
        cm, rho, cv = bytes.fromhex(v["commit_xy"]), bytes.fromhex(v["r32"]), bytes.fromhex(v["c_vec32"])
        if g == 0:
            tr = engine.bppp_transcripts(cm, rho, b"", 0, cv, h, mode="plain")
            ok = engine._lib.secp256k1_b200_bppp_norm_verify_batch  # shape guard path: g_len = 0 -> rejected
            out = np.ones(1, dtype=np.uint8)
            assert ok(engine._h, out.ctypes.data, pr.ctypes.data, pr.shape[1], np.array([pr.shape[1]], dtype=np.uint32).ctypes.data, tr.ctypes.data,
                      np.frombuffer(rho, dtype=np.uint8).ctypes.data, None, 0, np.frombuffer(cv, dtype=np.uint8).ctypes.data, h,
                      np.frombuffer(cm, dtype=np.uint8).ctypes.data, 1) == 1
            assert int(out[0]) == 0
            continue
        tr = engine.bppp_transcripts(cm, rho, gens[:64 * (g + h)], g, cv, h, mode="plain")
        got = engine.bppp_norm_verify(pr, [pr.shape[1]], tr, rho, gens[:64 * (g + h)], g, cv, h, cm)
        assert int(got[0]) == v["expected"], v["index"]


def test_bppp_norm_verify_batch_vs_reference(engine, ref):
    """SURVEY 8d config 5 geometry (16 G + 8 H generators), proofs made by the reference prover, 1/4 tampered;
    plus other geometries."""
    L = ref.lib
    L.secp256k1_bppp_generators_create.restype = ctypes.c_void_p
    rnd = random.Random(66)
    for (g, h, n) in [(16, 8, 96), (1, 1, 8), (64, 1, 4), (1, 64, 4), (8, 16, 8), (32, 32, 4)]:
        gs = ctypes.c_void_p(L.secp256k1_bppp_generators_create(ref.ctx, ctypes.c_size_t(g + h)))
        gxy = ctypes.create_string_buffer(64 * (g + h)); L.ref_bppp_generators_xy(gxy, gs)
        proofs, rhos, cvs, cms, exp = [], [], [], [], []
        for i in range(n):
            nv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(g)); lv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(h))
            cv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(h)); rho = b32(rnd.getrandbits(256) % N or 1)
            proof = ctypes.create_string_buffer(1000); plen = ctypes.c_size_t(1000); cm = ctypes.create_string_buffer(64)
            assert L.ref_bppp_norm_prove(ref.ctx, proof, ctypes.byref(plen), cm, rho, gs, nv, ctypes.c_size_t(g), lv, cv, ctypes.c_size_t(h)) == 1
            pr = bytearray(proof.raw[:plen.value])
            if i % 4 == 1:
                pr[rnd.randrange(len(pr))] ^= 1 << rnd.randrange(8)
            if i % 16 == 5:
                pr[-1] ^= 1          # the SURVEY 8d pattern: flip the last byte
            e = L.ref_bppp_norm_verify(ref.ctx, bytes(pr), ctypes.c_size_t(len(pr)), cm.raw, rho, gs, ctypes.c_size_t(g), cv, ctypes.c_size_t(h))
            proofs.append(bytes(pr)); rhos.append(rho); cvs.append(cv); cms.append(cm.raw); exp.append(e)
        P = np.frombuffer(b"".join(proofs), dtype=np.uint8).reshape(n, -1)
        tr = engine.bppp_transcripts(b"".join(cms), b"".join(rhos), gxy.raw, g, b"".join(cvs), h, mode="tests")
        got = engine.bppp_norm_verify(P, [P.shape[1]] * n, tr, b"".join(rhos), gxy.raw, g, b"".join(cvs), h, b"".join(cms))
        assert list(got) == exp, (g, h)
        assert sum(exp) >= n // 2


# ---- BASELINE.json full-size batches through size-independent properties -----------------------------------
def test_rangeproof_65536_batch_properties(engine):
    """config 4 size: 65536 64-bit proofs (reference-signed fixture tiled).  One flipped bit in every 32nd proof, at a
    position derived from the index: exactly those must be rejected, min/max must be the header values everywhere."""
    fx = np.load(os.path.join(GOLD, "rangeproof64_fixture.npz"))
    n = 65536
    proofs = np.tile(fx["proofs"], (n // 64, 1))
    commits = np.tile(fx["commits"], (n // 64, 1))
    bad = np.arange(3, n, 32)
    pos = (bad * 7919 + 11) % (5126 * 8)
    pos = np.where(pos < 16, pos + 64, pos)            # keep the header bytes intact so min/max stay comparable
    proofs[bad, pos // 8] ^= (1 << (pos % 8)).astype(np.uint8)
    ok, mn, mx = engine.rangeproof_verify(commits, proofs, np.full(n, 5126, dtype=np.uint32), fx["gen64"].tobytes())
    exp = np.ones(n, dtype=np.uint8); exp[bad] = 0
    assert np.array_equal(ok, exp)
    assert (mn == 0).all() and (mx == np.uint64(2**64 - 1)).all()
    okc, cg = engine.pedersen_commit(np.tile(fx["blinds"], (n // 64, 1)), np.tile(fx["values"], n // 64), fx["gen64"].tobytes())
    assert okc.all() and np.array_equal(cg, commits)


def test_bppp_4096_batch_properties(engine):
    """config 5 size: 4096 norm-product arguments at 16 G + 8 H generators (reference-made fixture tiled), last byte
    flipped for i = 5 (mod 16)."""
    fb = np.load(os.path.join(GOLD, "bppp_fixture.npz"))
    n = 4096
    r = n // 64
    bp, rho, cv, cm = (np.tile(fb[k], (r, 1)) for k in ("proofs", "rho", "c_vec", "commits"))
    bad = np.arange(5, n, 16)
    bp[bad, 323] ^= 1
    tr = engine.bppp_transcripts(cm, rho, fb["gens"], 16, cv, 8, mode="tests")
    ok = engine.bppp_norm_verify(bp, np.full(n, 324, dtype=np.uint32), tr, rho, fb["gens"], 16, cv, 8, cm)
    exp = np.ones(n, dtype=np.uint8); exp[bad] = 0
    assert np.array_equal(ok, exp)


def test_bppp_packed_large_batches_vs_reference(engine, ref):
    """From 4096 proofs on, k_bppp_norm_verify_packed runs (4 proofs per warp, comb tables for the batch's generators built on the stream):
    128 distinct reference-made arguments per geometry, 1/4 tampered at random positions (incl. malformed point encodings), tiled to 4224
    items so that the four proofs of a warp differ and the last warp is ragged; accept bits against the reference's."""
    L = ref.lib
    L.secp256k1_bppp_generators_create.restype = ctypes.c_void_p
    rnd = random.Random(67)
    for (g, h) in [(16, 8), (8, 16), (32, 32), (1, 1), (64, 2), (2, 64)]:
        n0 = 128 if g + h <= 32 else 32
        gs = ctypes.c_void_p(L.secp256k1_bppp_generators_create(ref.ctx, ctypes.c_size_t(g + h)))
        gxy = ctypes.create_string_buffer(64 * (g + h)); L.ref_bppp_generators_xy(gxy, gs)
        proofs, rhos, cvs, cms, exp = [], [], [], [], []
        for i in range(n0):
            nv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(g)); lv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(h))
            cv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(h)); rho = b32(rnd.getrandbits(256) % N or 1)
            proof = ctypes.create_string_buffer(1000); plen = ctypes.c_size_t(1000); cm = ctypes.create_string_buffer(64)
            assert L.ref_bppp_norm_prove(ref.ctx, proof, ctypes.byref(plen), cm, rho, gs, nv, ctypes.c_size_t(g), lv, cv, ctypes.c_size_t(h)) == 1
            pr = bytearray(proof.raw[:plen.value])
            if i % 4 == 1 and len(pr):
                pr[rnd.randrange(len(pr))] ^= 1 << rnd.randrange(8)
            if i % 16 == 6 and len(pr) >= 65:
                pr[0] = 4                                   # malformed sign byte of the first point pair
            e = L.ref_bppp_norm_verify(ref.ctx, bytes(pr), ctypes.c_size_t(len(pr)), cm.raw, rho, gs, ctypes.c_size_t(g), cv, ctypes.c_size_t(h))
            proofs.append(bytes(pr)); rhos.append(rho); cvs.append(cv); cms.append(cm.raw); exp.append(e)
        n = 4224 + 3                                        # >= 4096, not a multiple of 4
        idx = [i % n0 for i in range(n)]
        P = np.frombuffer(b"".join(proofs[i] for i in idx), dtype=np.uint8).reshape(n, -1)
        rh = b"".join(rhos[i] for i in idx); cvv = b"".join(cvs[i] for i in idx); cmm = b"".join(cms[i] for i in idx)
        tr = engine.bppp_transcripts(cmm, rh, gxy.raw, g, cvv, h, mode="tests")
        got = engine.bppp_norm_verify(P, [P.shape[1]] * n, tr, rh, gxy.raw, g, cvv, h, cmm)
        assert list(got) == [exp[i] for i in idx], (g, h)
        small = engine.bppp_norm_verify(P[:n0], [P.shape[1]] * n0, tr[:n0], rh[:32 * n0], gxy.raw, g, cvv[:32 * h * n0], h, cmm[:64 * n0])
        assert list(small) == exp, (g, h, "warp-per-proof kernel")
        assert 0 < sum(exp) < n0


def test_rangeproof_header_fuzz_vs_reference(engine, ref):
    """Every value of each of the first 11 proof bytes (header: flags/exp, mantissa, min_value, first sign byte) and
    truncated / extended lengths: accept bit AND min/max outputs must match the reference (SURVEY 8b edge semantics:
    exp > 18, mantissa > 64, max overflow, sign-bit padding, plen < 65, outputs written before the proof fails)."""
    gen = ref.generator_h()
    rnd = random.Random(99)
    for (mv, ex, mb, val) in [(0, 0, 64, 12345678901234567), (13, 1, 8, 13 + 70), (5, 18, 1, 5), (0, 0, 3, 5)]:
        blind = b32(rnd.getrandbits(256) % N or 1)
        proofs, plen, commits = ref.rangeproof_make(blind, [val], gen, min_value=mv, exp=ex, min_bits=mb)
        L = int(plen[0])
        base = proofs[0, :L].copy()
        variants, lens = [], []
        for pos in range(11):
            for v in range(256):
                p = base.copy(); p[pos] = v
                variants.append(p); lens.append(L)
        for cut in (0, 1, 32, 64, 65, 66, L - 33, L - 32, L - 1):
            if 0 <= cut <= L:
                variants.append(base.copy()); lens.append(cut)
        stride = L + 40
        batch = np.zeros((len(variants), stride), dtype=np.uint8)
        for i, p in enumerate(variants):
            batch[i, :L] = p
        lens += [L + 1, L + 32]
        batch = np.concatenate([batch, np.zeros((2, stride), dtype=np.uint8)]); batch[-2, :L] = base; batch[-1, :L] = base
        cm = np.repeat(commits[:1], batch.shape[0], axis=0)
        er = ref.rangeproof_verify(cm.tobytes(), batch, lens, gen)
        eg = engine.rangeproof_verify(cm, batch, lens, gen)
        for a, b_ in zip(er, eg):
            assert np.array_equal(a, b_), (mv, ex, mb, np.nonzero(a != b_)[0][:5])
        assert er[0].sum() >= 1


# ---- surjection proofs / tallies ----------------------------------------------------------
def test_surjectionproof_verify_batch_vs_reference(engine, ref):
    from tests.refshim import surjection_make, surjection_verify
    rnd = random.Random(77)
    proofs, tags, offs, outs, exp = [], [], [0], [], []
    for k in range(160):
        n_in = rnd.choice([1, 2, 3, 3, 5, 8, 13])
        n_use = rnd.randint(1, min(3, n_in))
        made = surjection_make(ref, n_in, n_use, rnd.randrange(n_in), hashlib.sha256(b"gsj%d" % k).digest())
        assert made is not None
        pr, tin, tout = made
        pr = bytearray(pr)
        if k % 4 == 1:
            pr[rnd.randrange(len(pr))] ^= 1 << rnd.randrange(8)
        if k % 9 == 2:
            tin = tin[64:] + tin[:64]                  # rotate the input tags
        if k % 11 == 3:
            tout = tin[:64]
        n_tags = n_in
        if k % 13 == 4 and n_in > 1:
            tin = tin[:-64]; n_tags = n_in - 1         # wrong number of tags
        exp.append(surjection_verify(ref, bytes(pr), tin, n_tags, tout))
        proofs.append(bytes(pr)); tags.append(tin); offs.append(offs[-1] + n_tags); outs.append(tout)
    stride = max(len(p) for p in proofs) + 7
    P = np.zeros((len(proofs), stride), dtype=np.uint8)
    for i, p in enumerate(proofs):
        P[i, :len(p)] = np.frombuffer(p, dtype=np.uint8)
    got = engine.surjectionproof_verify(P, [len(p) for p in proofs], b"".join(tags), offs, b"".join(outs))
    assert list(got) == exp
    assert 60 < sum(exp) < 150


def test_pedersen_verify_tally_batch_vs_reference(engine, ref):
    rnd = random.Random(88)
    gen = ref.generator_h()
    pos, neg, po, no, exp = [], [], [0], [0], []
    for k in range(200):
        nn, np_ = rnd.randint(1, 4), rnd.randint(1, 4)
        vals_n = [rnd.getrandbits(40) for _ in range(nn)]
        bl_n = [rnd.getrandbits(256) % N for _ in range(nn)]
        vals_p = [rnd.getrandbits(30) for _ in range(np_ - 1)]
        bl_p = [rnd.getrandbits(256) % N for _ in range(np_ - 1)]
        last = sum(vals_n) - sum(vals_p)
        if last < 0:
            vals_p = [0] * (np_ - 1); last = sum(vals_n)
        vals_p.append(last); bl_p.append((sum(bl_n) - sum(bl_p)) % N)
        cp = bytearray(ref.pedersen_commit(b"".join(b32(x) for x in bl_p), vals_p, gen)[1].tobytes())
        cn = bytearray(ref.pedersen_commit(b"".join(b32(x) for x in bl_n), vals_n, gen)[1].tobytes())
        if k % 3 == 1:
            cp[1 + rnd.randrange(32)] ^= 1
        if k % 7 == 2:
            cn[0] ^= 1                                   # flip the sign of one commitment
        exp.append(ref.lib.ref_pedersen_verify_tally(ref.ctx, bytes(cp), ctypes.c_size_t(np_), bytes(cn), ctypes.c_size_t(nn)))
        pos.append(bytes(cp)); neg.append(bytes(cn)); po.append(po[-1] + np_); no.append(no[-1] + nn)
    got = engine.pedersen_verify_tally(b"".join(pos), po, b"".join(neg), no)
    assert list(got) == exp and 80 < sum(exp) < 160


def test_empty_ragged_and_unaligned_batches(engine, ref):
    """edge shapes of the batch ABI: n = 0 on every entry point, n = 1 and n = 129 (not a multiple of any block size), and device
    buffers that start at odd addresses (the kernels fall back from 16-byte to byte loads) -- same verdicts as the aligned run"""
    import torch
    import secp256k1_zkp_b200 as b200
    e = engine
    z = np.zeros(0, dtype=np.uint8)
    assert e.schnorrsig_verify(z, z, z, b200.PK_XONLY32).size == 0
    assert e.ecdsa_verify(z, z, z, b200.SIG_COMPACT64, b200.PK_COMPRESSED33).size == 0
    ok, pk = e.ecdsa_recover(z, z, z); assert ok.size == 0 and pk.shape == (0, 64)
    assert e.ecmult(z, z, z).shape == (0, 64)
    assert e.schnorrsig_aggverify(b"", b"", bytes(32)) == 1          # the empty aggregate with s = 0 verifies (spec vector 0)
    assert e.schnorrsig_aggverify(b"", b"", bytes(31) + b"\x01") == 0
    out = e.ecmult_multi(np.zeros((0, 32), dtype=np.uint8), np.zeros((0, 64), dtype=np.uint8))
    assert out == bytes(64)                                            # empty sum = infinity = all-zero encoding
    for n in (1, 129):
        sk = b"".join(hashlib.sha256(b"rag%d-%d" % (n, i)).digest() for i in range(n))
        msg = b"".join(hashlib.sha256(b"ragm%d-%d" % (n, i)).digest() for i in range(n))
        sig, pk = ref.schnorr_make(sk, msg)
        sig = sig.copy(); sig[64 * (n // 2) + 5] ^= 1
        exp = ref.schnorr_verify(sig, msg, pk)
        assert np.array_equal(e.schnorrsig_verify(sig, msg, pk, b200.PK_XONLY32), exp) and exp.sum() == n - 1
        # the same batch from device buffers at odd byte offsets
        st = torch.cuda.current_stream().cuda_stream

        def odd(a, off):
            t = torch.zeros(len(a) + 16, dtype=torch.uint8, device="cuda")
            t[off:off + len(a)] = torch.frombuffer(bytearray(bytes(a)), dtype=torch.uint8).cuda()
            return t, t.data_ptr() + off
        ts, ps = odd(sig.tobytes(), 1); tm, pm = odd(msg, 3); tp, pp = odd(pk.tobytes(), 7)
        okd = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
        e.schnorrsig_verify_dev(okd.data_ptr() + 5, ps, pm, 32, pp, b200.PK_XONLY32, n, st)
        torch.cuda.synchronize()
        assert np.array_equal(okd[5:5 + n].cpu().numpy(), exp) and not okd[:5].any() and not okd[5 + n:].any()
        rsig, rid = ref.ecdsa_sign_recoverable(sk, msg)
        eok, epk = ref.ecdsa_recover(rsig, rid, msg)
        tr, pr = odd(rsig.tobytes(), 9); ti, pi = odd(rid.tobytes(), 1)
        outk = torch.zeros(64 * n + 16, dtype=torch.uint8, device="cuda")
        e.ecdsa_recover_dev(okd.data_ptr() + 5, outk.data_ptr() + 3, pr, pi, pm, n, st)
        torch.cuda.synchronize()
        assert np.array_equal(okd[5:5 + n].cpu().numpy(), eok) and np.array_equal(outk[3:3 + 64 * n].cpu().numpy().reshape(n, 64), epk)
