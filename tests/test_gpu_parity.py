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
