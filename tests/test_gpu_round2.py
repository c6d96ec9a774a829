"""GPU parity tests added in round 2 (run with -m gpu on a B200), all through the C ABI and bit-exact against the reference compiled into
oracle/_ref:
  * device unit kernels of the scalar / group primitives on the PTX limb path (and the 4-lane forms of the MSM tail) on >= 10^5 inputs with
    deterministic degenerate cases (a = b, a = -b, infinity on either side, z != 1), the reference's scalar edge vectors
    (src/tests.c:2456 run_scalar_tests) and the GLV near-split-bound scalars (src/tests.c:4779 run_ecmult_near_split_bound);
  * the size-adaptive MSM (window from n, every forced window), hot buckets, the sharded partial/combine path, bench_ecmult's point set."""
import hashlib
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
LAMBDA = 0x5363AD4CC05C30E0A5261C028812645A122E22EA20816678DF02967C1B23BD72
G64 = bytes.fromhex("79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798"
                    "483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8")
U_SC_MUL, U_SC_ADD, U_SC_NEG, U_SC_INV, U_SC_SPLIT, U_SC_RED = 0, 1, 2, 3, 4, 5
U_DBL, U_ADD, U_ADD_GE, U_DBL_G4, U_ADD_G4, U_ADD_GE_G4, U_DBL_W10, U_DBL16_W10, U_ADD_W10 = 8, 9, 10, 24, 25, 26, 27, 28, 29


def b32(x):
    return int(x).to_bytes(32, "big")


# scalars on which the reference's own tests concentrate: 0, 1, n-1, the halves of n, lambda, the GLV split bounds
# (src/tests.c:4779-4800 run_ecmult_near_split_bound: multiples of the lattice vectors around which split_lambda changes branch)
def _edge_scalars():
    a1 = 0x3086D221A7D46BCDE86C90E49284EB15
    b1 = 0xE4437ED6010E88286F547FA90ABFE4C3          # -b1 of the lattice basis
    a2 = 0x114CA50F7A8E2F3F657C1108D9D44CFD8
    edge = [0, 1, 2, N - 1, N - 2, (N - 1) // 2, (N + 1) // 2, LAMBDA, N - LAMBDA, (LAMBDA * LAMBDA) % N, 2**128, 2**128 - 1, 2**127, 2**255, 2**256 - 1,
            N, N + 1, 2**256 - 2**128, 0xFFFFFFFF, 2**64, 2**192 - 1]
    for v in (a1, b1, a2):
        for m in (1, 2, 3, 2**64, 2**100, 2**127):
            for d in (-2, -1, 0, 1, 2):
                edge.append((v * m + d) % 2**256)
                edge.append((N - v * m + d) % 2**256)
    # run_scalar_tests (src/tests.c:2456) style patterns: runs of 0x00 / 0xFF bytes
    rnd = random.Random(99)
    for _ in range(200):
        bs = bytearray(32)
        i = 0
        while i < 32:
            run = rnd.randint(1, 8)
            fill = rnd.choice([0x00, 0xFF, rnd.randint(0, 255)])
            for j in range(i, min(32, i + run)):
                bs[j] = fill
            i += run
        edge.append(int.from_bytes(bs, "big"))
    return edge


def test_unit_scalar_ops_vs_reference(engine, ref):
    rnd = random.Random(21)
    edge = _edge_scalars()
    n = 100000
    a = [rnd.choice(edge) if i % 3 == 0 else rnd.getrandbits(256) for i in range(n)] + [x for x in edge for _ in range(4)]
    b = [rnd.choice(edge) if i % 5 == 0 else rnd.getrandbits(256) for i in range(n)] + [rnd.choice(edge) for x in edge for _ in range(4)]
    A = b"".join(b32(x) for x in a)
    B = b"".join(b32(x) for x in b)
    for op in (U_SC_MUL, U_SC_ADD):
        assert np.array_equal(engine.unit_op(op, A, B), ref.unit_op(op, A, B)), op
    for op in (U_SC_NEG, U_SC_SPLIT, U_SC_RED):
        assert np.array_equal(engine.unit_op(op, A), ref.unit_op(op, A)), op
    m = 20000 + len(edge) * 4
    assert np.array_equal(engine.unit_op(U_SC_INV, A[-32 * m:]), ref.unit_op(U_SC_INV, A[-32 * m:]))
    # split_lambda invariants on the GPU output itself: k = r1 + lambda*r2 (mod n), both halves < 2^128 in absolute value
    out = engine.unit_op(U_SC_SPLIT, A).reshape(-1, 64)
    for i in range(0, len(a), 997):
        r1 = int.from_bytes(bytes(out[i, :32]), "big"); r2 = int.from_bytes(bytes(out[i, 32:]), "big")
        assert (r1 + LAMBDA * r2 - a[i]) % N == 0
        assert min(r1, N - r1) < 2**128 and min(r2, N - r2) < 2**128


def _points(ref, ks):
    n = len(ks)
    return np.frombuffer(ref.ecmult_batch(G64 * n, bytes(32 * n), b"".join(b32(k) for k in ks)).tobytes(), dtype=np.uint8).reshape(n, 64)


def test_unit_group_ops_vs_reference(engine, ref):
    """gej_double / gej_add / gej_add_ge and their 4-lane forms: 10^5 operand pairs over a pool of 2048 points, with every degenerate
    relation placed deterministically: equal points with different z, opposite points, infinity left / right / both, z = 1 and z = p - 1"""
    rnd = random.Random(22)
    pool_k = [rnd.getrandbits(256) % N for _ in range(2048)]
    pool = _points(ref, pool_k)
    n = 100000
    ia = np.array([rnd.randrange(2048) for _ in range(n)])
    ib = np.array([rnd.randrange(2048) for _ in range(n)])
    A = np.zeros((n, 96), dtype=np.uint8); B = np.zeros((n, 96), dtype=np.uint8)
    A[:, :64] = pool[ia]; B[:, :64] = pool[ib]
    zs = np.frombuffer(b"".join(b32(rnd.choice([1, 2, P - 1, P - 2, 3]) if i % 7 == 0 else rnd.getrandbits(256)) for i in range(2 * n)), dtype=np.uint8).reshape(2 * n, 32)
    A[:, 64:] = zs[:n]; B[:, 64:] = zs[n:]
    kinds = {0: 0, 1: 0, 2: 0, 3: 0, 4: 0, 5: 0}
    for i in range(0, n, 4):
        kind = (i // 4) % 8
        if kind == 0:                                   # a == b (doubling branch), different z
            B[i, :64] = A[i, :64]
        elif kind == 1:                                 # a == -b (result infinity)
            B[i, :32] = A[i, :32]
            y = int.from_bytes(bytes(A[i, 32:64]), "big")
            B[i, 32:64] = np.frombuffer(b32(P - y), dtype=np.uint8)
        elif kind == 2:
            A[i, :64] = 0                               # infinity + b
        elif kind == 3:
            B[i, :64] = 0                               # a + infinity
        elif kind == 4:
            A[i, :64] = 0; B[i, :64] = 0
        elif kind == 5:                                 # a == b and equal z too
            B[i] = A[i]
        kinds[kind if kind < 6 else 0] += 1
    assert all(v > 1000 for v in kinds.values())
    Ab, Bb = A.tobytes(), B.tobytes()
    exp_dbl, exp_add, exp_mix = ref.unit_op(U_DBL, Ab), ref.unit_op(U_ADD, Ab, Bb), ref.unit_op(U_ADD_GE, Ab, Bb)
    assert np.array_equal(engine.unit_op(U_DBL, Ab), exp_dbl)
    assert np.array_equal(engine.unit_op(U_ADD, Ab, Bb), exp_add)
    assert np.array_equal(engine.unit_op(U_ADD_GE, Ab, Bb), exp_mix)
    assert np.array_equal(engine.unit_op(U_DBL_G4, Ab), exp_dbl)
    assert np.array_equal(engine.unit_op(U_ADD_G4, Ab, Bb), exp_add)
    assert np.array_equal(engine.unit_op(U_ADD_GE_G4, Ab, Bb), exp_mix)
    # sanity of the expectation itself: the a == -b items really are infinity, the a == b items are the doubling
    e = exp_add.reshape(n, 64)
    assert not e[4].any() and np.array_equal(e[0], exp_dbl.reshape(n, 64)[0])
    # the warp-wide doubling chain (ge_w10.cuh: one 26-bit limb per lane): 2 a, and 2^16 a against 16 reference doublings
    m = 20000
    Aw = A[:m].copy()
    Aw[0, 64:] = np.frombuffer(b32(P - 1), dtype=np.uint8); Aw[1, 64:] = np.frombuffer(b32(1), dtype=np.uint8)
    assert np.array_equal(engine.unit_op(U_DBL_W10, Aw.tobytes()), ref.unit_op(U_DBL, Aw.tobytes()))
    cur = Aw.copy()
    for _ in range(16):
        d = ref.unit_op(U_DBL, cur.tobytes()).reshape(m, 64)
        cur[:, :64] = d
        cur[:, 64:] = 0                                   # s = 0 is taken as 1
    assert np.array_equal(engine.unit_op(U_DBL16_W10, Aw.tobytes()).reshape(m, 64), cur[:, :64])
    assert np.array_equal(engine.unit_op(U_ADD_W10, A[:m].tobytes(), B[:m].tobytes()), exp_add[:64 * m])


# ---- MSM ------------------------------------------------------------------------------------------------------------------------
def _msm_inputs(ref, rnd, m, hot_from=None):
    ks = [rnd.getrandbits(256) % N for _ in range(m)]
    pts = bytearray(_points(ref, ks).tobytes()) if m else bytearray()
    edge = [0, 1, N - 1, 2**255, 32768, 32769, 65535 << 16, N - 32768, 2**128 - 1, 2**128, (N - 1) // 2, (N + 1) // 2, LAMBDA, N - LAMBDA]
    scs = [rnd.choice(edge) if i % 4 == 0 else rnd.getrandbits(256) % N for i in range(m)]
    if m >= 8:
        pts[64 * 3:64 * 4] = bytes(64)
        pts[64 * 5:64 * 6] = pts[64 * 4:64 * 5]; scs[5] = (N - scs[4]) % N
        pts[64 * 6:64 * 7] = pts[64 * 4:64 * 5]; scs[6] = scs[4]
    if hot_from is not None:
        for i in range(hot_from, m):
            scs[i] = scs[hot_from]
    return b"".join(b32(x) for x in scs), bytes(pts)


def test_msm_dispatch_sizes_vs_reference(engine, ref):
    """every size class of the dispatcher (secp256k1_ecmult_multi_var picks Strauss below 88 points and a bucket window from n,
    src/ecmult_impl.h:597-621,:823-867): window chosen from n, with and without the generator term"""
    rnd = random.Random(56)
    engine.ecmult_multi_set_window(0)
    for m in (0, 1, 2, 8, 87, 88, 300, 2047, 4096, 9000, 40000):
        sc, pts = _msm_inputs(ref, rnd, m, hot_from=(m // 2 if m in (300, 9000) else None))
        for g in (None, b32(rnd.getrandbits(256) % N), b32(0)):
            ok, exp = ref.ecmult_multi(sc, pts, g, scratch_bytes=64 << 20)
            assert ok == 1
            assert engine.ecmult_multi(sc, pts, g) == exp, (m, g is None)


def test_msm_every_window_width(engine, ref):
    rnd = random.Random(57)
    try:
        for m, cs in ((150, range(4, 17)), (6000, range(8, 17))):
            sc, pts = _msm_inputs(ref, rnd, m, hot_from=m - 70)
            g = b32(rnd.getrandbits(256) % N)
            ok, exp = ref.ecmult_multi(sc, pts, g, scratch_bytes=64 << 20)
            assert ok == 1
            for c in cs:
                engine.ecmult_multi_set_window(c)
                assert engine.ecmult_multi(sc, pts, g) == exp, (m, c)
    finally:
        engine.ecmult_multi_set_window(0)


def test_msm_bench_ecmult_point_set(engine, ref):
    """config 3 of BASELINE.json as src/bench_ecmult.c builds it (:262-276, :362-371): s_i = SHA256("ecmult" || LE32(i)) mod n,
    P_i = 2^i * G; expected = (sum s_i 2^i) * G with scalars only (:286-298), and the reference's own ecmult_multi_var"""
    m = 32768
    scs = [int.from_bytes(hashlib.sha256(b"ecmult" + i.to_bytes(4, "little")).digest(), "big") % N for i in range(m)]
    ks = [pow(2, i, N) for i in range(m)]
    pts = _points(ref, ks).tobytes()
    sc = b"".join(b32(x) for x in scs)
    tot = sum(s * k for s, k in zip(scs, ks)) % N
    exp = bytes(engine.ecmult(bytes(64), bytes(32), b32(tot))[0])
    assert engine.ecmult_multi(sc, pts) == exp
    ok, exp2 = ref.ecmult_multi(sc, pts, None, scratch_bytes=64 << 20)
    assert ok == 1 and exp2 == exp


def _dev_msm(engine, torch, sc_t, pt_t, m, partial=False):
    wsb = engine.ecmult_multi_workspace_bytes(m)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    out = torch.zeros(96 if partial else 64, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    fn = engine.ecmult_multi_partial_dev if partial else engine.ecmult_multi_dev
    fn(out.data_ptr(), None, sc_t.data_ptr(), pt_t.data_ptr(), m, ws.data_ptr(), wsb, st)
    return out, ws, wsb, st


def test_msm_sharded_partial_and_combine(engine, ref):
    """an MSM shards by terms; W partial Jacobian sums are gathered and added once.  Shard sizes are ragged on purpose."""
    import torch
    rnd = random.Random(58)
    m = 3000
    sc, pts = _msm_inputs(ref, rnd, m)
    ok, exp = ref.ecmult_multi(sc, pts, None, scratch_bytes=64 << 20)
    sc_t = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda().reshape(m, 32)
    pt_t = torch.frombuffer(bytearray(pts), dtype=torch.uint8).cuda().reshape(m, 64)
    for world in (1, 2, 3, 8):
        cuts = [m * r // world for r in range(world + 1)]
        cuts[1] = min(cuts[1], 1) if world == 3 else cuts[1]             # a one-term shard
        parts = torch.zeros((world, 96), dtype=torch.uint8, device="cuda")
        keep = []
        for r in range(world):
            lo, hi = cuts[r], cuts[r + 1]
            o, ws, wsb, st = _dev_msm(engine, torch, sc_t[lo:hi].contiguous(), pt_t[lo:hi].contiguous(), hi - lo, partial=True)
            torch.cuda.synchronize()
            parts[r] = o
            keep.append(ws)
        out = torch.zeros(64, dtype=torch.uint8, device="cuda")
        engine.ecmult_multi_combine_dev(out.data_ptr(), parts.data_ptr(), world, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert out.cpu().numpy().tobytes() == exp, world
    # all partials infinity / zero partials
    z = torch.zeros((4, 96), dtype=torch.uint8, device="cuda")
    out = torch.ones(64, dtype=torch.uint8, device="cuda")
    engine.ecmult_multi_combine_dev(out.data_ptr(), z.data_ptr(), 4, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert not out.cpu().numpy().any()


def test_msm_window_shards_add_up(engine, ref):
    """the other sharding (terms resident everywhere, bucket windows spread over the shards): for every shard count the partial results add up
    to the reference's MSM -- with the generator term, with every dispatcher window width, also when there are more shards than windows"""
    import torch
    rnd = random.Random(59)
    st = torch.cuda.current_stream().cuda_stream
    try:
        for m, c in ((3000, 0), (3000, 4), (5000, 13), (40000, 16), (9, 8), (1 << 18, 0)):
            if m <= 40000:
                sc, pts = _msm_inputs(ref, rnd, m)
            else:                                       # big enough that the piece size (and with it the workspace layout) depends on the shard count
                _, p4k = _msm_inputs(ref, rnd, 4096)
                pts = p4k * (m // 4096)
                sc = b"".join(b32(rnd.getrandbits(256) % N) for _ in range(m))
            g = b32(rnd.getrandbits(256) % N)
            ok, exp = ref.ecmult_multi(sc, pts, g, scratch_bytes=64 << 20)
            assert ok == 1
            sc_t = torch.frombuffer(bytearray(sc), dtype=torch.uint8).cuda()
            pt_t = torch.frombuffer(bytearray(pts), dtype=torch.uint8).cuda()
            g_t = torch.frombuffer(bytearray(g), dtype=torch.uint8).cuda()
            wsb = engine.ecmult_multi_workspace_bytes(m)
            ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
            engine.ecmult_multi_set_window(c)
            for shards in (1, 2, 8, 5, 70):
                parts = torch.zeros((shards, 96), dtype=torch.uint8, device="cuda")
                for s in range(shards):
                    engine.ecmult_multi_window_shard_dev(parts[s].data_ptr(), g_t.data_ptr(), sc_t.data_ptr(), pt_t.data_ptr(), m, s, shards, ws.data_ptr(), wsb, st)
                out = torch.zeros(64, dtype=torch.uint8, device="cuda")
                engine.ecmult_multi_combine_dev(out.data_ptr(), parts.data_ptr(), shards, st)
                torch.cuda.synchronize()
                assert out.cpu().numpy().tobytes() == exp, (m, c, shards)
    finally:
        engine.ecmult_multi_set_window(0)


def test_msm_adversarial_skew_time_and_value(engine, ref):
    """2^20 IDENTICAL scalars: every window has ONE bucket holding every term.  Value: k * sum(P_i).  Time: at most 2x the random-scalar
    MSM of the same size (pieces spread a hot bucket over the whole GPU; round 1 put it on one thread)."""
    import torch
    n = 1 << 20
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    sk = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen); sk[:, 0] &= 0x7F
    z64, z32 = torch.zeros((n, 64), dtype=torch.uint8, device=dev), torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    pts = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    engine.ecmult_dev(pts.data_ptr(), z64.data_ptr(), z32.data_ptr(), sk.data_ptr(), n, st)
    rnd_sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen)
    k = random.Random(3).getrandbits(256) % N
    same = torch.frombuffer(bytearray(b32(k)), dtype=torch.uint8).to(dev).repeat(n, 1).contiguous()
    wsb = engine.ecmult_multi_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.zeros(64, dtype=torch.uint8, device=dev)

    def run(sc):
        for _ in range(2):
            engine.ecmult_multi_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), n, ws.data_ptr(), wsb, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            engine.ecmult_multi_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), n, ws.data_ptr(), wsb, st)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 3, out.cpu().numpy().tobytes()
    t_rand, _ = run(rnd_sc)
    t_same, got = run(same)
    tot = sum(int.from_bytes(bytes(r), "big") for r in sk.cpu().numpy()) % N
    exp = bytes(engine.ecmult(bytes(64), bytes(32), b32(tot * k % N))[0])
    assert got == exp
    print("MSM 2^20: random %.3f ms, identical scalars %.3f ms" % (t_rand, t_same))
    assert t_same <= 2.0 * t_rand, (t_rand, t_same)


# ---- cross-item isolation of the block-shared inversions (a crafted off-curve key must not touch its neighbours) -----------
def test_offcurve_key_does_not_poison_its_block(engine, ref):
    import secp256k1_zkp_b200 as b200
    n = 512
    sk = b"".join(hashlib.sha256(b"iso-sk%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"iso-msg%d" % i).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    okc, opaque = engine.point_codec(pk, 32, 1)                      # x-only -> 64-byte opaque secp256k1_xonly_pubkey (x, y little-endian)
    assert okc.all()
    opaque = opaque.copy(); sig = sig.copy()
    for j in (5, 130, 131, 300):                                     # (x, 0): not on the curve; doubling it gives Z = 0
        opaque[j, 32:] = 0
        sig[64 * j:64 * j + 32] = 0                                  # r = 0: what a zeroed shared inverse would have made "valid"
    got = engine.schnorrsig_verify(sig, msg, opaque, b200.PK_OPAQUE64)
    exp = np.ones(n, dtype=np.uint8); exp[[5, 130, 131, 300]] = 0
    assert np.array_equal(got, exp)
    # raw ecmult (points taken as given): the off-curve items report infinity, every other item is untouched
    ks = [int.from_bytes(hashlib.sha256(b"iso-k%d" % i).digest(), "big") % N for i in range(n)]
    pts = _points(ref, ks).copy()
    na = b"".join(b32(k) for k in ks)
    clean = engine.ecmult(pts, na, na)
    assert np.array_equal(clean, np.frombuffer(ref.ecmult_batch(pts.tobytes(), na, na).tobytes(), dtype=np.uint8).reshape(n, 64))
    bad = pts.copy()
    for j in (7, 129, 400):
        bad[j, 32:] = 0
    out = engine.ecmult(bad, na, na)
    mask = np.ones(n, dtype=bool); mask[[7, 129, 400]] = False
    assert np.array_equal(out[mask], clean[mask])


# ---- range proofs: every mantissa, and a large batch of DISTINCT 64-bit proofs, item by item against the reference ---------------------
def _threaded(fn, n, threads=16):
    import threading
    per = (n + threads - 1) // threads
    ts = [threading.Thread(target=fn, args=(t * per, min(n, (t + 1) * per))) for t in range(threads) if t * per < n]
    for t in ts:
        t.start()
    for t in ts:
        t.join()


def test_rangeproof_every_mantissa_and_4096_distinct(engine, ref):
    gen = ref.generator_h()
    # (a) mantissa 1..64 (ring counts 1..32, odd mantissas end in a 2-member ring), 8 proofs each, half of them damaged
    proofs, plens, commits = [], [], []
    for m in range(1, 65):
        blinds = b"".join(hashlib.sha256(b"mant-b%d-%d" % (m, i)).digest() for i in range(8))
        vals = [int.from_bytes(hashlib.sha256(b"mant-v%d-%d" % (m, i)).digest()[:8], "little") % (1 << m) for i in range(8)]
        p, pl, c = ref.rangeproof_make(blinds, vals, gen, min_bits=m)
        proofs.append(p); plens.append(pl); commits.append(c)
    proofs = np.concatenate(proofs); plens = np.concatenate(plens); commits = np.concatenate(commits)
    n = proofs.shape[0]
    for i in range(1, n, 2):
        off = int.from_bytes(hashlib.sha256(b"mant-flip%d" % i).digest()[:4], "big") % (8 * int(plens[i]))
        proofs[i, off // 8] ^= 1 << (off % 8)
    ok, mn, mx = engine.rangeproof_verify(commits, proofs, plens, gen)
    eok, emn, emx = ref.rangeproof_verify(commits, proofs, plens, gen)
    hdr_ok = np.array([(int(p[0]) & 128) == 0 for p in proofs])           # a rejected header leaves min / max unspecified in the reference
    assert np.array_equal(ok, eok)
    assert np.array_equal(mn[hdr_ok & (eok == 1)], emn[hdr_ok & (eok == 1)]) and np.array_equal(mx[hdr_ok & (eok == 1)], emx[hdr_ok & (eok == 1)])
    assert int(eok[0::2].sum()) == n // 2 and len(set(int(x) for x in plens)) >= 32
    # (b) 4096 distinct 64-bit proofs (SURVEY 8d config 4 at one sixteenth), 1/8 with one flipped bit
    n = 4096
    blinds = [hashlib.sha256(b"rp-b" + i.to_bytes(4, "little")).digest() for i in range(n)]
    vals = [int.from_bytes(hashlib.sha256(b"rp-v" + i.to_bytes(4, "little")).digest()[:8], "little") for i in range(n)]
    P_ = np.zeros((n, 5126), dtype=np.uint8); C_ = np.zeros((n, 33), dtype=np.uint8)

    def sign(lo, hi):
        p, pl, c = ref.rangeproof_make(b"".join(blinds[lo:hi]), vals[lo:hi], gen, stride=5126)
        P_[lo:hi] = p; C_[lo:hi] = c
    _threaded(sign, n)
    for i in range(3, n, 8):
        off = int.from_bytes(hashlib.sha256(b"%d" % i).digest()[:4], "big") % (8 * 5126)
        P_[i, off // 8] ^= 1 << (off % 8)
    pl = np.full(n, 5126, dtype=np.uint64)
    ok, mn, mx = engine.rangeproof_verify(C_, P_, pl, gen)
    eok = np.zeros(n, dtype=np.uint8); emx = np.zeros(n, dtype=np.uint64)

    def check(lo, hi):
        a, _, c = ref.rangeproof_verify(C_[lo:hi], P_[lo:hi], pl[lo:hi], gen)
        eok[lo:hi] = a; emx[lo:hi] = c
    _threaded(check, n)
    assert np.array_equal(ok, eok) and int(eok.sum()) >= n - len(range(3, n, 8)) - 2 and int(eok.sum()) < n
    assert np.array_equal(mx[eok == 1], emx[eok == 1])
    okc, cg = engine.pedersen_commit(b"".join(blinds), vals, gen)          # the same (blind, value) pairs through the batched commit
    assert okc.all() and np.array_equal(cg, C_)


# ---- batched parsing / serialization: every format of secp256k1_b200_point_codec_batch at scale -----------------------
def test_point_codec_batch_all_formats_vs_reference_parsers(engine, ref):
    import ctypes
    L = ref.lib
    n = 100000
    rnd = random.Random(41)
    ks = [rnd.getrandbits(256) % N or 1 for _ in range(4096)]
    pool = _points(ref, ks)                                            # valid affine points
    idx = np.array([rnd.randrange(4096) for _ in range(n)])
    xs = pool[idx, :32].copy(); ys = pool[idx, 32:].copy()
    parity = (ys[:, 31] & 1).astype(np.uint8)
    bad_kind = np.array([i % 10 for i in range(n)])                    # 0: x not on the curve, 1: x >= p, 2: bad tag, rest valid
    for i in range(n):
        if bad_kind[i] == 0:
            xs[i] = np.frombuffer(hashlib.sha256(b"codec%d" % i).digest(), dtype=np.uint8)      # ~half of these have no y
        elif bad_kind[i] == 1:
            xs[i] = np.frombuffer((P + (i % 977)).to_bytes(32, "big"), dtype=np.uint8)

    def ref_loop(fn, recs, inlen, out_len=64):
        ok = np.zeros(n, dtype=np.uint8); out = np.zeros((n, 64), dtype=np.uint8)
        buf = ctypes.create_string_buffer(64)
        for i in range(n):
            ok[i] = fn(ref.ctx, buf, recs[i].tobytes()[:inlen], *([ctypes.c_size_t(inlen)] if fn is L.secp256k1_ec_pubkey_parse else []))
            if ok[i]:
                out[i, :out_len] = np.frombuffer(buf.raw[:out_len], dtype=np.uint8)
        return ok, out
    # format 1: 32-byte x-only keys -> opaque secp256k1_xonly_pubkey
    ok, out = engine.point_codec(xs.tobytes(), 32, 1)
    rok, rout = ref_loop(L.secp256k1_xonly_pubkey_parse, xs, 32)
    assert np.array_equal(ok, rok) and np.array_equal(out[ok == 1], rout[rok == 1]) and 0 < int(ok.sum()) < n
    # format 2: SEC1 compressed (33) and uncompressed (65) keys -> opaque secp256k1_pubkey
    tag = (2 + parity).astype(np.uint8); tag[bad_kind == 2] = 5
    c33 = np.concatenate([tag[:, None], xs], axis=1)
    ok, out = engine.point_codec(c33.tobytes(), 33, 2)
    rok, rout = ref_loop(L.secp256k1_ec_pubkey_parse, c33, 33)
    assert np.array_equal(ok, rok) and np.array_equal(out[ok == 1], rout[rok == 1])
    t65 = np.full(n, 4, dtype=np.uint8); t65[bad_kind == 2] = 1; t65[bad_kind == 3] = 6 + parity[bad_kind == 3]     # hybrid encodings are legal
    u65 = np.concatenate([t65[:, None], xs, ys], axis=1)
    ok, out = engine.point_codec(u65.tobytes(), 65, 2)
    rok, rout = ref_loop(L.secp256k1_ec_pubkey_parse, u65, 65)
    assert np.array_equal(ok, rok) and np.array_equal(out[ok == 1], rout[rok == 1]) and 0 < int(ok.sum()) < n
    # format 4: 33-byte generators (tags 10 / 11 = y is a square / is not) -> 64-byte secp256k1_generator; format 5 = serialize back
    gtag = np.array([10 + (i & 1) for i in range(n)], dtype=np.uint8); gtag[bad_kind == 2] = 9
    g33 = np.concatenate([gtag[:, None], xs], axis=1)
    ok, out = engine.point_codec(g33.tobytes(), 33, 4)
    rok, rout = ref_loop(L.secp256k1_generator_parse, g33, 33)
    assert np.array_equal(ok, rok) and np.array_equal(out[ok == 1], rout[rok == 1]) and 0 < int(ok.sum()) < n
    good = out[ok == 1]
    ok5, ser = engine.point_codec(good.tobytes(), 64, 5)
    assert ok5.all() and np.array_equal(ser[:, :33], g33[ok == 1])
    # format 3: 33-byte commitments (tags 8 / 9) -> validity + affine point; the reference's parser accepts exactly the same encodings
    ctag = np.array([8 + (i & 1) for i in range(n)], dtype=np.uint8); ctag[bad_kind == 2] = 10
    c33 = np.concatenate([ctag[:, None], xs], axis=1)
    ok, out = engine.point_codec(c33.tobytes(), 33, 3)
    rok, _ = ref_loop(L.secp256k1_pedersen_commitment_parse, c33, 33, out_len=33)
    assert np.array_equal(ok, rok) and 0 < int(ok.sum()) < n
    for i in np.nonzero(ok)[0][::97]:
        x = int.from_bytes(bytes(out[i, :32]), "big"); y = int.from_bytes(bytes(out[i, 32:]), "big")
        assert x == int.from_bytes(bytes(xs[i]), "big") and (y * y - x * x * x - 7) % P == 0
        assert (pow(y, (P - 1) // 2, P) == 1) == (ctag[i] == 8)          # tag 8: y is the square root that is itself a square (:266-273)


# ---- division-step inverse / Jacobi symbol on the device (csrc/modinv.cuh) ------------------------------------------------------------
def test_safegcd_inverse_and_jacobi_gpu(engine, ref):
    import secp256k1_zkp_b200 as b200
    rnd = random.Random(31)
    edge = [0, 1, 2, 3, P - 1, P - 2, P, P + 1, 2**256 - 1, 2**255, 2**128, 2**30, 2**30 - 1, 2**60, 2**240, (P + 1) // 2, (P - 1) // 2, 977, 2**32 + 977]
    vals = edge + [rnd.getrandbits(256) for _ in range(30000)] + [rnd.getrandbits(k) for k in range(1, 256) for _ in range(8)]
    A = b"".join(b32(v) for v in vals)
    inv = [int.from_bytes(bytes(r), "big") for r in engine.fe_op(b200.FE_INV, A)]
    assert inv == [pow(v % P, P - 2, P) for v in vals]
    fer = engine.fe_op(b200.FE_INV_FERMAT, A)
    assert [int.from_bytes(bytes(r), "big") for r in fer] == inv
    sq = engine.fe_op(b200.FE_IS_SQUARE, A)[:, 31]
    exp = np.array([1 if (v % P == 0 or pow(v % P, (P - 1) // 2, P) == 1) else 0 for v in vals], dtype=np.uint8)
    assert np.array_equal(sq, exp)
    assert np.array_equal(engine.fe_op(b200.FE_IS_SQUARE_SQRT, A)[:, 31], exp)
    # the reference's own fe_inv_var / fe_is_square_var on a sample
    for v in vals[:200]:
        assert ref.fe1("ref_fe_inv", b32(v))[0] == b32(pow(v % P, P - 2, P))
        assert ref.lib.ref_fe_is_square(b32(v)) == int(exp[vals.index(v)])
