"""CPU tier: the device headers compiled for the host (portable limb code) against the reference in
oracle/_ref and against the reference's own golden vectors.  This checks the LOGIC the kernels run
(field folding, scalar reduction, GLV split, signed-window recoding, group law corner cases, verify
epilogues); the PTX limb code itself is checked on the GPU tier (test_gpu_parity.py)."""
import ctypes
import hashlib
import json
import os
import random

import numpy as np
import pytest

from tests.hostemu_lib import load

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
G64 = bytes.fromhex("79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798"
                    "483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def b32(x):
    return int(x).to_bytes(32, "big")


@pytest.fixture(scope="module")
def emu():
    return load()


def test_field_ops(emu):
    rnd = random.Random(1)
    edge = [0, 1, 2, P - 1, P, P + 1, 2**256 - 1, 2**256 - 2, 2**32 + 977, 2**32 + 976, 2**255, P - 2, 2**128, 2**224 - 1]
    vals = edge + [rnd.getrandbits(256) for _ in range(200)] + [rnd.getrandbits(256) | (2**256 - 2**200) for _ in range(40)]
    r = ctypes.create_string_buffer(32)
    inv2 = pow(2, -1, P)
    for x in vals:
        for y in rnd.sample(vals, 12) + edge:
            emu.emu_fe_mul(r, b32(x), b32(y)); assert int.from_bytes(r.raw, "big") == x * y % P
            emu.emu_fe_add(r, b32(x), b32(y)); assert int.from_bytes(r.raw, "big") == (x + y) % P
            emu.emu_fe_sub(r, b32(x), b32(y)); assert int.from_bytes(r.raw, "big") == (x - y) % P
        emu.emu_fe_sqr(r, b32(x)); assert int.from_bytes(r.raw, "big") == x * x % P
        emu.emu_fe_neg(r, b32(x)); assert int.from_bytes(r.raw, "big") == (-x) % P
        emu.emu_fe_half(r, b32(x)); assert int.from_bytes(r.raw, "big") == x * inv2 % P
        for k in (2, 3, 4, 8, 21, 977, 2**20 - 1):
            emu.emu_fe_mul_small(r, b32(x), k); assert int.from_bytes(r.raw, "big") == x * k % P
        assert emu.emu_fe_is_zero(b32(x)) == (x % P == 0)
    for x in vals[:60]:
        emu.emu_fe_inv(r, b32(x)); assert int.from_bytes(r.raw, "big") == pow(x, P - 2, P)
        ok = emu.emu_fe_sqrt(r, b32(x))
        assert ok == (pow(x % P, (P - 1) // 2, P) in (0, 1))
        assert int.from_bytes(r.raw, "big") == pow(x, (P + 1) // 4, P)


def test_safegcd_inverse_and_jacobi(emu):
    """modinv.cuh (division-step inverse mod p and mod n, Jacobi symbol) against Python's pow on edge values, every bit length and
    12 000 random inputs -- the values secp256k1_fe_inv_var / secp256k1_scalar_inverse_var / secp256k1_fe_is_square_var return"""
    rnd = random.Random(4)
    edge = [0, 1, 2, 3, P - 1, P - 2, P, P + 1, 2**256 - 1, 2**255, 2**128, 2**30, 2**30 - 1, 2**60, 2**240, (P + 1) // 2, (P - 1) // 2, 977,
            2**32 + 977, N, N - 1, N + 1, (N - 1) // 2]
    vals = edge + [rnd.getrandbits(256) for _ in range(12000)] + [rnd.getrandbits(k) for k in range(1, 256) for _ in range(8)]
    n = len(vals)
    A = b"".join(b32(v) for v in vals)
    out = ctypes.create_string_buffer(32 * n)
    emu.emu_fe_inv_safegcd(out, A, ctypes.c_size_t(n))
    assert all(int.from_bytes(out.raw[32 * i:32 * i + 32], "big") == pow(v % P, P - 2, P) for i, v in enumerate(vals))
    emu.emu_sc_inverse_safegcd(out, A, ctypes.c_size_t(n))
    assert all(int.from_bytes(out.raw[32 * i:32 * i + 32], "big") == pow(v % N, N - 2, N) for i, v in enumerate(vals))
    sq = ctypes.create_string_buffer(n)
    emu.emu_fe_is_square_jacobi(sq, A, ctypes.c_size_t(n))
    assert all(sq.raw[i] == (1 if (v % P == 0 or pow(v % P, (P - 1) // 2, P) == 1) else 0) for i, v in enumerate(vals))


def test_scalar_ops_and_sha(emu, ref):
    rnd = random.Random(2)
    edge = [0, 1, 2, N - 1, N, N + 1, 2**256 - 1, N // 2, N // 2 + 1, N // 2 + 2, 2**128, 2**128 - 1]
    vals = edge + [rnd.getrandbits(256) for _ in range(150)]
    r, r2, q, q2 = (ctypes.create_string_buffer(32) for _ in range(4))
    for x in vals:
        for y in rnd.sample(vals, 8) + edge:
            emu.emu_sc_mul(r, b32(x), b32(y)); assert int.from_bytes(r.raw, "big") == (x % N) * (y % N) % N
            emu.emu_sc_add(r, b32(x), b32(y)); assert int.from_bytes(r.raw, "big") == ((x % N) + (y % N)) % N
        emu.emu_sc_negate(r, b32(x)); assert int.from_bytes(r.raw, "big") == (-x) % N
        ov = emu.emu_sc_set_b32(r, b32(x)); assert ov == (x >= N) and int.from_bytes(r.raw, "big") == x % N
        assert emu.emu_sc_is_high(b32(x)) == ((x % N) > N // 2)
        emu.emu_sc_split_lambda(r, r2, b32(x)); ref.lib.ref_scalar_split_lambda(q, q2, b32(x % N))
        assert r.raw == q.raw and r2.raw == q2.raw
    for x in vals[:30]:
        emu.emu_sc_inverse(r, b32(x)); assert int.from_bytes(r.raw, "big") == pow(x % N, N - 2, N)
    for ln in [0, 1, 55, 56, 63, 64, 65, 119, 120, 128, 1000]:
        d = bytes(rnd.getrandbits(8) for _ in range(ln))
        emu.emu_sha256(r, d, ctypes.c_size_t(ln)); assert r.raw == hashlib.sha256(d).digest()


def _refmul(ref, k):
    o = ctypes.create_string_buffer(64)
    ref.lib.ref_ecmult(o, G64, b32(0), b32(k))
    return o.raw


def test_group_law_and_ecmult(emu, ref):
    rnd = random.Random(3)
    L = ref.lib
    o, o2 = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
    pts = [_refmul(ref, k) for k in [1, 2, 3, 5, N - 1, N - 2, rnd.getrandbits(200)]] + [bytes(64)]
    for a in pts:
        emu.emu_ge_double(o, a); L.ref_ge_double(o2, a); assert o.raw == o2.raw
        for c in pts:
            emu.emu_ge_add(o, a, c); L.ref_ge_add(o2, a, c); assert o.raw == o2.raw
            if a != bytes(64):
                emu.emu_ge_add_mixed(o, a, c); assert o.raw == o2.raw
    lam = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
    scal = [0, 1, 2, 3, N - 1, N - 2, 2**128, 2**128 - 1, 2**127, N // 2, N // 2 + 1, lam, N - lam] + \
           [rnd.getrandbits(256) % N for _ in range(60)] + [rnd.getrandbits(k) for k in (8, 64, 120, 129, 130)]
    for i, na in enumerate(scal):
        for ng in rnd.sample(scal, 4) + [0, 1]:
            a = pts[i % len(pts)]
            e1 = emu.emu_ecmult(o, a, b32(na), b32(ng)); e2 = L.ref_ecmult(o2, a, b32(na), b32(ng))
            assert o.raw == o2.raw and e1 == e2


def test_bip340_vectors(emu):
    vecs = json.load(open(os.path.join(GOLD, "bip340.json")))["vectors"]
    for v in vecs:
        pk = bytes.fromhex(v["pk"])
        if v["pk_parse_fails"]:
            # parsing is fused into the verify: any signature must be rejected
            assert emu.emu_schnorr_verify(bytes(64), bytes(32), 32, pk, 1) == 0
            continue
        msg = bytes.fromhex(v["msg"])
        got = emu.emu_schnorr_verify(bytes.fromhex(v["sig"]), msg if msg else b"\0", len(msg), pk, 1)
        assert got == v["expected"], v["index"]


def test_wycheproof_ecdsa_vectors(emu):
    vecs = json.load(open(os.path.join(GOLD, "wycheproof_ecdsa.json")))["vectors"]
    assert len(vecs) == 463
    for v in vecs:
        if not v["der_ok"]:
            assert v["expected"] == 0
            continue
        pk65 = bytes.fromhex(v["pk65"])
        opaque = pk65[1:33][::-1] + pk65[33:65][::-1]     # 64-byte opaque pubkey = x,y little-endian
        got = emu.emu_ecdsa_verify(bytes.fromhex(v["compact"]), 1, bytes.fromhex(v["msghash"]), opaque, 0)
        assert got == v["expected"], v["tcId"]
        comp = bytes([2 + (pk65[64] & 1)]) + pk65[1:33]
        assert emu.emu_ecdsa_verify(bytes.fromhex(v["compact"]), 1, bytes.fromhex(v["msghash"]), comp, 2) == v["expected"]


def _constants_sha_digest(mul, prefix, iters):
    acc = hashlib.sha256()
    scalars = [0, 1, N - 1] + [int.from_bytes(hashlib.sha256(prefix.to_bytes(4, "little") + i.to_bytes(2, "little")).digest(), "big") % N
                               for i in range(iters)]
    for x in scalars:
        xy = mul(x)
        acc.update(b"\0" if xy == bytes(64) else b"\x04" + xy)
    return acc.hexdigest()


def test_ecmult_constants_digests(emu):
    """run_ecmult_constants (src/tests.c:5800-5960): digests computed by an independent implementation."""
    g = json.load(open(os.path.join(GOLD, "ecmult.json")))["constants"]
    o = ctypes.create_string_buffer(64)

    def mul_g(x):       # ng*G through the comb table
        emu.emu_ecmult(o, bytes(64), b32(0), b32(x)); return o.raw

    def mul_a(x):       # na*A with A = G through the GLV / fixed-window path
        emu.emu_ecmult(o, G64, b32(x), b32(0)); return o.raw
    c = g["sha_6bit20"]
    assert _constants_sha_digest(mul_g, c["prefix"], c["iter"]) == c["digest"]
    assert _constants_sha_digest(mul_a, c["prefix"], c["iter"]) == c["digest"]
    c = g["sha_8bit8"]
    assert _constants_sha_digest(mul_a, c["prefix"], c["iter"]) == c["digest"]


def test_ecmult_chain_endpoint_closed_form(emu):
    """run_ecmult_chain (src/tests.c:4617-4674): X_{i+1} = xn_i*X_i + gn_i*G; after 20000 steps the
    reference hard-codes the endpoint.  ae/ge are accumulated with Python integers, then ONE ecmult
    ae*A + ge*G must hit the expected point."""
    c = json.load(open(os.path.join(GOLD, "ecmult.json")))["chain"]
    xn, gn = int(c["xn"], 16), int(c["gn"], 16)
    xf, gf = int(c["xf"], 16), int(c["gf"], 16)
    ae, ge = 1, 0
    for _ in range(c["iterations"]):
        ae = ae * xn % N
        ge = (ge * xn + gn) % N
        xn = xn * xf % N
        gn = gn * gf % N
    o = ctypes.create_string_buffer(64)
    emu.emu_ecmult(o, bytes.fromhex(c["a_xy"]), b32(ae), b32(ge))
    assert o.raw.hex() == c["expected_xy"].lower()


def test_schnorr_and_ecdsa_random_vs_reference(emu, ref):
    n = 200
    sk = b"".join(hashlib.sha256(b"sk%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"m%d" % i).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    sig, pk, m = bytearray(sig.tobytes()), bytearray(pk.tobytes()), bytearray(msg)
    for i in range(n):
        if i % 3 == 1: sig[64 * i + (i * 7) % 64] ^= 1 << (i % 8)
        if i % 7 == 2: m[32 * i + i % 32] ^= 0x10
        if i % 11 == 3: pk[32 * i + i % 32] ^= 0x01
    exp = ref.schnorr_verify(bytes(sig), bytes(m), bytes(pk))
    for i in range(n):
        assert emu.emu_schnorr_verify(bytes(sig[64 * i:64 * i + 64]), bytes(m[32 * i:32 * i + 32]), 32, bytes(pk[32 * i:32 * i + 32]), 1) == exp[i]
    sig, pk = ref.ecdsa_make(sk, msg)
    sig, pk, m = bytearray(sig.tobytes()), bytearray(pk.tobytes()), bytearray(msg)
    for i in range(n):
        if i % 3 == 1: sig[64 * i + (i * 7) % 64] ^= 1 << (i % 8)
        if i % 7 == 2: m[32 * i + i % 32] ^= 0x10
        if i % 11 == 3: pk[33 * i + 1 + i % 32] ^= 0x01
        if i % 13 == 4:
            s = int.from_bytes(sig[64 * i + 32:64 * i + 64], "big"); sig[64 * i + 32:64 * i + 64] = b32((N - s) % N)
    exp = ref.ecdsa_verify(bytes(sig), bytes(m), bytes(pk))
    for i in range(n):
        assert emu.emu_ecdsa_verify(bytes(sig[64 * i:64 * i + 64]), 1, bytes(m[32 * i:32 * i + 32]), bytes(pk[33 * i:33 * i + 33]), 2) == exp[i]


# ---- rangeproof / pedersen / MSM / BP++ logic (the warp-cooperative kernels, lanes run sequentially) ---------------
def _emu_rp(emu, commit, proof, gen, extra=b""):
    mn, mx = ctypes.c_uint64(0), ctypes.c_uint64(0)
    ok = emu.emu_rangeproof_verify(ctypes.byref(mn), ctypes.byref(mx), bytes(commit), bytes(proof), ctypes.c_size_t(len(proof)), bytes(gen),
                                   extra if extra else None, ctypes.c_size_t(len(extra)))
    return ok, mn.value, mx.value


def test_rangeproof_and_pedersen_logic(emu, ref):
    gen = ref.generator_h()
    for v in json.load(open(os.path.join(GOLD, "rangeproof.json")))["vectors"]:
        pr, c = bytes.fromhex(v["proof"]), bytes.fromhex(v["commit"])
        assert _emu_rp(emu, c, pr, gen) == (1, v["min_value"], v["max_value"]), v["name"]
        bad = bytearray(pr); bad[len(bad) // 3] ^= 8
        assert _emu_rp(emu, c, bytes(bad), gen)[0] == 0
        assert _emu_rp(emu, c, pr + b"\0", gen)[0] == 0
    rnd = random.Random(3)
    vals = [0, 1, 2**64 - 1] + [rnd.getrandbits(64) for _ in range(4)]
    blinds = b"".join(b32(rnd.getrandbits(256) % N) for _ in vals)
    okr, cr = ref.pedersen_commit(blinds, vals, gen)
    for i, v in enumerate(vals):
        out = ctypes.create_string_buffer(33)
        assert emu.emu_pedersen_commit(out, blinds[32 * i:32 * i + 32], ctypes.c_uint64(v), gen) == 1 and out.raw == bytes(cr[i])
    for (mv, ex, mb, vs) in [(13, 1, 8, [77, 200]), (100, 2, 13, [100 + 4500, 100 + 123400]), (0, 0, 1, [0, 1]), (0, 3, 0, [7000, 9000])]:
        proofs, plen, commits = ref.rangeproof_make(blinds[:64], vs, gen, min_value=mv, exp=ex, min_bits=mb)
        for i in range(2):
            pr = bytes(proofs[i][:int(plen[i])])
            e = ref.rangeproof_verify_one(bytes(commits[i]), pr, gen)
            assert e == _emu_rp(emu, bytes(commits[i]), pr, gen) and e[0] == 1
            for pos in random.Random(i).sample(range(len(pr)), 8):
                bad = bytearray(pr); bad[pos] ^= 1 << (pos % 8)
                assert ref.rangeproof_verify_one(bytes(commits[i]), bytes(bad), gen) == _emu_rp(emu, bytes(commits[i]), bytes(bad), gen)


def _msm_case(rnd, ref, m, same_scalar_from=None):
    ks = [rnd.getrandbits(256) % N for _ in range(m)]
    pts = bytearray(ref.ecmult_batch(G64 * m, bytes(32 * m), b"".join(b32(k) for k in ks)).tobytes()) if m else bytearray()
    edge = [0, 1, N - 1, 2**255, 32768, 32769, 65535 << 16, N - 32768, 2**128 - 1, 2**128, 2**127, (N - 1) // 2, (N + 1) // 2, LAMBDA, N - LAMBDA]
    scs = [rnd.choice(edge) if i % 4 == 0 else rnd.getrandbits(256) % N for i in range(m)]
    if m >= 7:
        pts[64 * 3:64 * 4] = bytes(64)                                       # point at infinity
        pts[64 * 5:64 * 6] = pts[64 * 4:64 * 5]; scs[5] = (N - scs[4]) % N   # cancelling pair
        pts[64 * 6:64 * 7] = pts[64 * 4:64 * 5]; scs[6] = scs[4]             # equal points in one bucket (doubling branch)
    if same_scalar_from is not None:
        for i in range(same_scalar_from, m):
            scs[i] = scs[same_scalar_from]                                   # hot buckets: many terms share every digit
    return b"".join(b32(x) for x in scs), bytes(pts)


LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72


def test_borromean_challenge_hash_layouts(emu):
    """borromean_hash (src/modules/rangeproof/borromean_impl.h:23-37) = SHA256(e || m || be32(ring) || be32(idx)) reduced mod n: the kernels
    assemble the two blocks as words for e of 32 bytes (e0) and 33 bytes (a serialized point); any other length takes the byte path"""
    rnd = random.Random(31)
    emu.emu_borromean_challenge.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32]
    for _ in range(300):
        for elen in (32, 33, 20, 64):
            e = bytes(rnd.getrandbits(8) for _ in range(elen))
            m = bytes(rnd.getrandbits(8) for _ in range(32))
            ring, idx = rnd.choice([0, 1, 31, 0x01020304, 0xFFFFFFFF]), rnd.choice([0, 3, 0x80000000, 0xA1B2C3D4])
            out = ctypes.create_string_buffer(32)
            ov = emu.emu_borromean_challenge(out, e, elen, m, ring, idx)
            d = int.from_bytes(hashlib.sha256(e + m + ring.to_bytes(4, "big") + idx.to_bytes(4, "big")).digest(), "big")
            assert int.from_bytes(out.raw, "big") == d % N and ov == int(d >= N), (elen, ring, idx)


def test_msm_pipeline_logic(emu, ref):
    """(also covers MSM_R2_PARTS > 1: c = 16 has 8192 chunk sums in the top window)
    the GPU MSM pipeline (k_msm.cu) executed sequentially on the host: every window width, forced small pieces (multi-piece buckets +
    combine), hot buckets, edge scalars incl. the GLV split bounds -- against the reference's ecmult_multi_var"""
    rnd = random.Random(5)
    emu.emu_msm2.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32]
    cases = [(0, 0, 0), (1, 0, 0), (7, 0, 0), (100, 0, 0), (300, 0, 0)]
    cases += [(23, c, 0) for c in range(4, 17)]                              # every window width (top-window geometry differs for each)
    cases += [(150, 5, 2), (150, 8, 3), (90, 13, 1)]                         # tiny pieces: multi-piece buckets, parts, combine
    for (m, c, tp) in cases:
        sc, pts = _msm_case(rnd, ref, m, same_scalar_from=(40 if tp else None))
        for g in (None, b32(rnd.getrandbits(256) % N)):
            out = ctypes.create_string_buffer(64)
            emu.emu_msm2(out, g, sc if m else None, pts if m else None, m, c, tp)
            ok, exp = ref.ecmult_multi(sc, pts, g, scratch_bytes=1 << 22)
            assert ok == 1 and out.raw == exp, (m, c, tp, g is None)
    # window sharding: the partial results of the shards add up to the MSM, whatever the number of shards (more shards than windows included)
    emu.emu_msm_window_shards.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_uint32]
    for (m, c) in ((60, 0), (40, 4), (33, 16), (21, 11)):
        sc, pts = _msm_case(rnd, ref, m)
        g = b32(rnd.getrandbits(256) % N)
        ok, exp = ref.ecmult_multi(sc, pts, g, scratch_bytes=1 << 22)
        for shards in (1, 2, 3, 8, 70):
            out = ctypes.create_string_buffer(64)
            emu.emu_msm_window_shards(out, g, sc, pts, m, c, shards)
            assert ok == 1 and out.raw == exp, (m, c, shards)
    # the dispatcher: window width grows with n, c = 8 below 2^11 terms, 16 at 2^20
    cs = [emu.emu_msm_choose_c(ctypes.c_size_t(n)) for n in (1, 88, 2048, 1 << 13, 1 << 15, 1 << 17, 1 << 20, 1 << 24)]
    assert cs == [8, 8, 8, 10, 11, 12, 16, 16], cs


def test_bppp_norm_verify_logic(emu, ref):
    d = json.load(open(os.path.join(GOLD, "bppp_verify.json")))
    gens = bytes.fromhex(d["gens_xy"])
    for v in d["vectors"]:
        pr = bytes.fromhex(v["proof"])
        got = emu.emu_bppp_norm_verify(pr, ctypes.c_size_t(len(pr)), bytes.fromhex(v["commit_xy"]), bytes.fromhex(v["r32"]), gens, v["g_len"],
                                       bytes.fromhex(v["c_vec32"]), v["h_len"], 0)
        assert got == v["expected"], v["index"]
    L = ref.lib
    L.secp256k1_bppp_generators_create.restype = ctypes.c_void_p
    rnd = random.Random(8)
    for (g, h) in [(16, 8), (1, 1), (8, 16)]:
        gs = ctypes.c_void_p(L.secp256k1_bppp_generators_create(ref.ctx, ctypes.c_size_t(g + h)))
        gxy = ctypes.create_string_buffer(64 * (g + h)); L.ref_bppp_generators_xy(gxy, gs)
        nv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(g)); lv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(h))
        cv = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(h)); rho = b32(rnd.getrandbits(256) % N)
        proof = ctypes.create_string_buffer(1000); plen = ctypes.c_size_t(1000); cm = ctypes.create_string_buffer(64)
        assert L.ref_bppp_norm_prove(ref.ctx, proof, ctypes.byref(plen), cm, rho, gs, nv, ctypes.c_size_t(g), lv, cv, ctypes.c_size_t(h)) == 1
        pr = proof.raw[:plen.value]
        for mut in (None, plen.value - 1, plen.value - 33, 0, 7):
            p2 = bytearray(pr)
            if mut is not None: p2[mut] ^= 1
            p2 = bytes(p2)
            e = L.ref_bppp_norm_verify(ref.ctx, p2, ctypes.c_size_t(len(p2)), cm.raw, rho, gs, ctypes.c_size_t(g), cv, ctypes.c_size_t(h))
            assert e == emu.emu_bppp_norm_verify(p2, ctypes.c_size_t(len(p2)), cm.raw, rho, gxy.raw, g, cv, h, 1), (g, h, mut)


def test_rangeproof_header_fuzz_logic(emu, ref):
    gen = ref.generator_h()
    blind = b32(0x1234567)
    proofs, plen, commits = ref.rangeproof_make(blind, [13 + 70], gen, min_value=13, exp=1, min_bits=8)
    L = int(plen[0]); base = bytes(proofs[0][:L]); c = bytes(commits[0])
    cases = [(bytes(base[:pos]) + bytes([v]) + bytes(base[pos + 1:]), L) for pos in range(11) for v in range(0, 256, 5)]
    cases += [(base, cut) for cut in (0, 1, 64, 65, 66, L - 32, L - 1)] + [(base + bytes(32), L + 1), (base + bytes(32), L + 32)]
    for p, ln in cases:
        e = ref.rangeproof_verify_one(c, p[:ln], gen)
        assert e == _emu_rp(emu, c, p[:ln], gen), (ln, p[:12].hex())


def test_surjection_and_tally_logic(emu, ref):
    from tests.refshim import surjection_make, surjection_verify
    rnd = random.Random(12)
    for (n_in, n_use) in [(1, 1), (3, 1), (3, 3), (5, 2), (8, 3), (17, 3)]:
        made = surjection_make(ref, n_in, n_use, rnd.randrange(n_in), hashlib.sha256(b"sj%d-%d" % (n_in, n_use)).digest())
        assert made is not None
        pr, tin, tout = made
        for mu in [None] + [rnd.randrange(len(pr)) for _ in range(6)] + [0, 1, 2, len(pr) - 1]:
            p2 = bytearray(pr)
            if mu is not None: p2[mu] ^= 1 << rnd.randrange(8)
            p2 = bytes(p2)
            e = surjection_verify(ref, p2, tin, n_in, tout)
            assert e == emu.emu_surjection_verify(p2, ctypes.c_size_t(len(p2)), tin, n_in, tout), (n_in, n_use, mu)
            if mu is None: assert e == 1
        for (pp, tt, nn, oo) in [(pr, tin[::-1], n_in, tout), (pr[:-1], tin, n_in, tout), (pr, tin, n_in, tin[:64])]:
            assert surjection_verify(ref, pp, tt, nn, oo) == emu.emu_surjection_verify(pp, ctypes.c_size_t(len(pp)), tt, nn, oo)
    gen = ref.generator_h()
    vals_n, bl_n = [1000, 234], [rnd.getrandbits(256) % N for _ in range(2)]
    vals_p, bl_p = [1200, 34], [rnd.getrandbits(256) % N]
    bl_p.append((sum(bl_n) - sum(bl_p)) % N)
    cp = ref.pedersen_commit(b"".join(b32(x) for x in bl_p), vals_p, gen)[1].tobytes()
    cn = ref.pedersen_commit(b"".join(b32(x) for x in bl_n), vals_n, gen)[1].tobytes()
    assert ref.lib.ref_pedersen_verify_tally(ref.ctx, cp, ctypes.c_size_t(2), cn, ctypes.c_size_t(2)) == emu.emu_pedersen_tally(cp, 2, cn, 2) == 1
    bad = bytearray(cp); bad[40] ^= 1
    assert ref.lib.ref_pedersen_verify_tally(ref.ctx, bytes(bad), ctypes.c_size_t(2), cn, ctypes.c_size_t(2)) == emu.emu_pedersen_tally(bytes(bad), 2, cn, 2) == 0
    assert emu.emu_pedersen_tally(cp, 1, cn, 2) == ref.lib.ref_pedersen_verify_tally(ref.ctx, cp, ctypes.c_size_t(1), cn, ctypes.c_size_t(2)) == 0


# ---- ECDSA public-key recovery --------------------------------------------------------------------------
def _recover_cases(ref, n, tag):
    """compact signatures + recids from the reference signer, then tampered: wrong recid, flipped bits, r/s = 0, r + n variants"""
    sk = b"".join(hashlib.sha256(b"%ssk%d" % (tag, i)).digest() for i in range(n))
    msg = bytearray(b"".join(hashlib.sha256(b"%sm%d" % (tag, i)).digest() for i in range(n)))
    sig, rid = ref.ecdsa_sign_recoverable(sk, bytes(msg))
    sig, rid = bytearray(sig.tobytes()), bytearray(rid.tobytes())
    for i in range(n):
        if i % 3 == 1: rid[i] = (rid[i] + 1 + i % 3) & 3
        if i % 5 == 2: sig[64 * i + (i * 7) % 64] ^= 1 << (i % 8)
        if i % 7 == 3: msg[32 * i + i % 32] ^= 0x10
        if i % 23 == 4: sig[64 * i:64 * i + 32] = bytes(32)
        if i % 29 == 5: sig[64 * i + 32:64 * i + 64] = bytes(32)
        if i % 31 == 6: sig[64 * i:64 * i + 32] = b32(N + 5 - (i % 7))                 # r overflows: parse failure
        if i % 37 == 7: sig[64 * i:64 * i + 32] = b32(1 + (i % 5)); rid[i] = 2 + (i & 1)  # tiny r with the "r + n" bit
        if i % 41 == 8: rid[i] = 4 + (i % 200)                                         # out-of-range recid
    return bytes(sig), bytes(rid), bytes(msg)


def test_ecdsa_recover_vs_reference_and_golden(emu, ref):
    out = ctypes.create_string_buffer(64)
    for v in json.load(open(os.path.join(GOLD, "recovery.json")))["vectors"]:
        for rid in range(4):
            assert emu.emu_ecdsa_recover(out, bytes.fromhex(v["sig64"]), rid, bytes.fromhex(v["msg32"])) == v["recover_ok"][rid]
    n = 150
    sig, rid, msg = _recover_cases(ref, n, b"emu")
    eok, epk = ref.ecdsa_recover(sig, rid, msg)
    assert 0.2 < eok.mean() < 0.9
    for i in range(n):
        ok = emu.emu_ecdsa_recover(out, sig[64 * i:64 * i + 64], rid[i], msg[32 * i:32 * i + 32])
        assert ok == eok[i], i
        assert out.raw == epk[i].tobytes(), i


# ---- BIP-340 half-aggregation -----------------------------------------------------------------------------
def _halfagg_cases(ref, n, tag):
    """(pk32, msgs32, aggsig, label) list: a valid aggregate of n reference signatures plus tampered variants"""
    sk = b"".join(hashlib.sha256(b"%shsk%d" % (tag, i)).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"%shm%d" % (tag, i)).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    sig, pk = sig.tobytes(), pk.tobytes()
    agg = ref.aggregate(pk, msg, sig)
    assert len(agg) == 32 * (n + 1)
    out = [(pk, msg, agg, "valid")]

    def flip(b, pos, bit=1):
        b = bytearray(b); b[pos] ^= bit; return bytes(b)
    if n:
        out.append((pk, msg, flip(agg, 32 * n + 31), "s+-1"))
        out.append((pk, msg, flip(agg, 5, 4), "r_0 tampered"))
        out.append((pk, flip(msg, 32 * (n - 1) + 3), agg, "last msg tampered"))
        out.append((flip(pk, 32 * (n // 2) + 31), msg, agg, "key tampered (may stop parsing)"))
        out.append((pk, msg, agg[:32 * (n // 2)] + b"\xff" * 32 + agg[32 * (n // 2) + 32:], "r >= p"))
        out.append((pk, msg, agg[:32 * n] + b32(N), "s = n (overflow)"))
        out.append((pk, msg, agg[:-32], "aggsig one block short"))
        out.append((pk, msg, agg + bytes(32), "aggsig one block long"))
        out.append((pk, msg, agg + b"\x00", "aggsig_len not a multiple of 32"))
        if n > 1:   # swap two (key, msg, r) triples: the randomizers depend on the order
            i, j = 0, n - 1
            def sw(b, w):
                b = bytearray(b); b[w * i:w * i + w], b[w * j:w * j + w] = b[w * j:w * j + w], b[w * i:w * i + w]; return bytes(b)
            out.append((sw(pk, 32), sw(msg, 32), sw(agg[:32 * n], 32) + agg[32 * n:], "order swapped"))
    return out


def test_halfagg_vs_reference_and_golden(emu, ref):
    for v in json.load(open(os.path.join(GOLD, "halfagg.json")))["vectors"]:
        pk, msg, agg = bytes.fromhex(v["pubkeys"]), bytes.fromhex(v["msgs32"]), bytes.fromhex(v["aggsig"])
        assert emu.emu_halfagg_verify(pk, msg, ctypes.c_size_t(len(pk) // 32), agg, ctypes.c_size_t(len(agg))) == v["expected"]
    seen = set()
    for n in (0, 1, 2, 5, 17):
        for pk, msg, agg, label in _halfagg_cases(ref, n, b"emu%d" % n):
            exp = ref.aggverify(pk, msg, agg)
            got = emu.emu_halfagg_verify(pk, msg, ctypes.c_size_t(n), agg, ctypes.c_size_t(len(agg)))
            assert got == (1 if exp == 1 else 0), (n, label)     # exp == -1: a key the reference cannot even parse
            assert (label == "valid") == (exp == 1) or label.startswith("key tampered"), (n, label, exp)
            seen.add(exp)
    assert seen >= {0, 1}


# ---- range-proof rewinding ----------------------------------------------------------------------------------
def _rewind_cases(ref, tag, count):
    """reference-signed proofs of assorted shapes with embedded messages; returns dict of arrays + the inputs that went in"""
    rnd = random.Random(tag)
    gen = ref.generator_h()
    shapes = [(0, 64, 0), (0, 8, 0), (1, 3, 1), (0, 1, 5), (2, 16, 0), (0, 0, 0), (-1, 0, 7), (18, 64, 0)]   # (exp, min_bits, min_value)
    values, mins, exps, bits, msgs = [], [], [], [], []
    for i in range(count):
        exp, mb, mv = shapes[i % len(shapes)]
        if exp == -1:
            v = mv + rnd.randrange(1000); msg = b""            # exact-value proof: single ring, no room for a message
        elif exp == 18:
            v = 2**64 - 1 - rnd.randrange(10); msg = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 31, 100, 3968])))
        else:
            v = mv + rnd.randrange(1 << max(mb - 1, 1)) * (10 ** exp)
            msg = bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 1, 32, 33, 64])))
        values.append(v); mins.append(mv); exps.append(exp); bits.append(mb); msgs.append(msg)
    blind = b"".join(hashlib.sha256(b"rwb%d-%d" % (tag, i)).digest() for i in range(count))
    nonce = b"".join(hashlib.sha256(b"rwn%d-%d" % (tag, i)).digest() for i in range(count))
    rows = []
    for i in range(count):      # sign one by one: a message that does not fit the proof's shape makes the reference signer refuse -> retry without
        args = (blind[32 * i:32 * i + 32], [values[i]], [mins[i]], [exps[i]], [bits[i]], nonce[32 * i:32 * i + 32])
        try:
            rows.append(ref.rangeproof_make_msg(*args, [msgs[i]], gen))
        except AssertionError:
            msgs[i] = b""
            rows.append(ref.rangeproof_make_msg(*args, [b""], gen))
    proofs = np.concatenate([r[0] for r in rows]); plen = np.concatenate([r[1] for r in rows]); commit = np.concatenate([r[2] for r in rows])
    return {"proofs": proofs, "plen": plen, "commit": commit, "nonce": nonce, "blind": blind, "values": values, "msgs": msgs, "gen": gen}


def _emu_rewind(emu, nonce, commit, proof, gen, cap=4096, extra=b""):
    bl = ctypes.create_string_buffer(32); val = ctypes.c_uint64(0); msg = ctypes.create_string_buffer(max(cap, 1)); ml = ctypes.c_size_t(0)
    mn, mx = ctypes.c_uint64(0), ctypes.c_uint64(0)
    ok = emu.emu_rangeproof_rewind(bl, ctypes.byref(val), msg, ctypes.c_size_t(cap), ctypes.byref(ml), ctypes.byref(mn), ctypes.byref(mx), bytes(nonce),
                                   bytes(commit), bytes(proof), ctypes.c_size_t(len(proof)), bytes(gen), extra if extra else None, ctypes.c_size_t(len(extra)))
    return ok, bl.raw, val.value, msg.raw[:ml.value], mn.value, mx.value


def test_rangeproof_rewind_golden_and_reference(emu, ref):
    gen = ref.generator_h()
    for v in json.load(open(os.path.join(GOLD, "rangeproof_rewind.json")))["vectors"]:
        ok, bl, val, msg, mn, mx = _emu_rewind(emu, bytes.fromhex(v["nonce"]), bytes.fromhex(v["commit"]), bytes.fromhex(v["proof"]), gen)
        assert ok == 1 and bl.hex() == v["blind"] and val == int(v["value"]), v["name"]
        assert msg[:v["message_len"]] == bytes([255]) * v["message_len"] and mn == int(v["min_value"]) and mx == int(v["max_value"]), v["name"]
    c = _rewind_cases(ref, 1, 16)
    n = 16
    # wrong nonce / tampered proof / small message buffer variants ride along
    nonce = bytearray(c["nonce"]); nonce[32 * 3] ^= 1
    proofs = c["proofs"].copy(); proofs[5, int(c["plen"][5]) - 1] ^= 1
    for cap in (4096, 40, 0):
        eok, ebl, eval_, emsg, emn, emx = ref.rangeproof_rewind(bytes(nonce), c["commit"], proofs, c["plen"], c["gen"], msg_cap=cap)
        assert eok[3] == 0 and eok[5] == 0 and eok.sum() == n - 2
        for i in range(n):
            ok, bl, val, msg, mn, mx = _emu_rewind(emu, nonce[32 * i:32 * i + 32], c["commit"][i], proofs[i, :int(c["plen"][i])], c["gen"], cap=cap)
            assert ok == eok[i], (cap, i)
            if ok:
                assert bl == ebl[i].tobytes() == c["blind"][32 * i:32 * i + 32] and val == eval_[i] == c["values"][i], (cap, i)
                assert msg == emsg[i] and mn == emn[i] and mx == emx[i], (cap, i)
                if cap == 4096:
                    assert msg[:len(c["msgs"][i])] == c["msgs"][i]
