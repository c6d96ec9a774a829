"""C-level multi-device engine (secp256k1_b200_multi_*, csrc/multi.cu): sharded host-buffer calls must return exactly what the reference
returns item by item, whatever the number of devices.  With one visible GPU the W = 1 path runs (no NCCL); on a box with >= 2 GPUs
(`gpurun --gpus 2`) the same tests run with every device: ncclCommInitAll, the table broadcast, the all-gather of MSM partial sums."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
G64 = bytes.fromhex("79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798"
                    "483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8")


def b32(x):
    return int(x).to_bytes(32, "big")


@pytest.fixture(scope="module")
def multi():
    import torch
    import secp256k1_zkp_b200 as b200
    m = b200.MultiEngine(n_devices=torch.cuda.device_count())
    yield m
    m.close()


def test_multi_reports_devices_and_nccl(multi):
    import torch
    assert multi.device_count == torch.cuda.device_count()
    if multi.device_count > 1:
        assert multi.nccl_version >= 21000          # NCCL was loaded and initialised (ncclCommInitAll + table broadcast succeeded)


def test_multi_schnorr_ecdsa_ecmult_vs_reference(multi, ref):
    import secp256k1_zkp_b200 as b200
    n = 3001                                        # not a multiple of the device count: ragged shards
    sk = b"".join(hashlib.sha256(b"multi-sk%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"multi-msg%d" % i).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    sig = sig.copy()
    for i in range(0, n, 7):
        sig[64 * i + (i % 64)] ^= 1
    assert np.array_equal(multi.schnorrsig_verify(sig, msg, pk, b200.PK_XONLY32), ref.schnorr_verify(sig, msg, pk))
    esig, epk = ref.ecdsa_make(sk, msg)
    esig = np.array(esig, dtype=np.uint8).copy().reshape(-1)
    for i in range(3, n, 11):
        esig[64 * i + 40] ^= 4
    assert np.array_equal(multi.ecdsa_verify(esig, msg, epk), ref.ecdsa_verify(esig, msg, epk))
    pts = ref.ecmult_batch(G64 * n, bytes(32 * n), sk)
    assert np.array_equal(multi.ecmult(pts, msg, sk), ref.ecmult_batch(pts, msg, sk))


def test_multi_pedersen_and_rangeproof_vs_reference(multi, ref):
    n = 37
    gen = ref.generator_h()
    blinds = b"".join(hashlib.sha256(b"multi-b%d" % i).digest() for i in range(n))
    vals = [int.from_bytes(hashlib.sha256(b"multi-v%d" % i).digest()[:8], "little") for i in range(n)]
    okc, commits = multi.pedersen_commit(blinds, vals, gen)
    assert okc.all() and np.array_equal(commits, ref.pedersen_commit(blinds, vals, gen)[1])
    proofs, plen, rc = ref.rangeproof_make(blinds, vals, gen)
    assert np.array_equal(np.asarray(rc).reshape(n, 33), commits)
    proofs = np.array(proofs, dtype=np.uint8).copy()
    for i in range(2, n, 5):
        proofs[i, 100 + i] ^= 1
    ok, mn, mx = multi.rangeproof_verify(commits, proofs, plen, gen)
    eok, emn, emx = ref.rangeproof_verify(commits, proofs, plen, gen)
    assert np.array_equal(ok, eok) and np.array_equal(mn, emn) and np.array_equal(mx, emx)
    assert 0 < int(ok.sum()) < n


def test_multi_sharded_msm_vs_reference(multi, ref):
    rnd = random.Random(77)
    for m in (0, 1, 5, 300, 5000):
        ks = [rnd.getrandbits(256) % N for _ in range(m)]
        pts = ref.ecmult_batch(G64 * m, bytes(32 * m), b"".join(b32(k) for k in ks)).tobytes() if m else b""
        scs = [rnd.choice([0, 1, N - 1, 2**128]) if i % 9 == 0 else rnd.getrandbits(256) % N for i in range(m)]
        sc = b"".join(b32(x) for x in scs)
        for g in (None, b32(rnd.getrandbits(256) % N)):
            ok, exp = ref.ecmult_multi(sc, pts, g, scratch_bytes=64 << 20)
            assert ok == 1 and multi.ecmult_multi(sc, pts, g) == exp, (m, g is None)
