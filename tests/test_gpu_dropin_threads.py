"""GPU tier: the drop-in boundary under concurrency and on its error paths (SURVEY 8b "Threading", "Error conventions").
  * many host threads calling the reference-named functions at once -- directly (the engine serializes them) and through the opt-in
    coalescing front-end (concurrent calls share one batch launch) -- must get exactly the reference's results;
  * an engine failure reaches the context's error callback and the call returns 0;
  * static / cloned contexts; secp256k1_selftest as a device known-answer test; the reference's order of checks for degenerate inputs."""
import ctypes
import hashlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
CB = ctypes.CFUNCTYPE(None, ctypes.c_char_p, ctypes.c_void_p)
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141


@pytest.fixture(scope="module")
def lib():
    import secp256k1_zkp_b200 as b200
    L = ctypes.CDLL(b200.library_path())
    L.secp256k1_context_create.restype = ctypes.c_void_p
    L.secp256k1_context_create.argtypes = [ctypes.c_uint]
    L.secp256k1_context_clone.restype = ctypes.c_void_p
    L.secp256k1_context_clone.argtypes = [ctypes.c_void_p]
    L.secp256k1_context_destroy.argtypes = [ctypes.c_void_p]
    L.secp256k1_context_set_error_callback.argtypes = [ctypes.c_void_p, CB, ctypes.c_void_p]
    L.secp256k1_context_set_illegal_callback.argtypes = [ctypes.c_void_p, CB, ctypes.c_void_p]
    L.secp256k1_schnorrsig_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    L.secp256k1_ecdsa_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_xonly_pubkey_parse.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_ec_pubkey_parse.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    L.secp256k1_ecdsa_signature_parse_compact.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_rangeproof_verify.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_char_p,
                                              ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    L.secp256k1_pedersen_commitment_parse.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_b200_coalesce_configure.argtypes = [ctypes.c_uint, ctypes.c_size_t]
    L.secp256k1_b200_test_fail_next_gpu_calls.argtypes = [ctypes.c_int]
    return L


def _workload(ref, lib, ctx, n):
    """n Schnorr + n ECDSA items (1/5 corrupted) as the opaque objects the single calls take, with the reference's verdicts"""
    sk = b"".join(hashlib.sha256(b"thr-sk%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"thr-msg%d" % i).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    sig = sig.copy()
    for i in range(0, n, 5):
        sig[64 * i + (i % 64)] ^= 1
    s_exp = ref.schnorr_verify(sig, msg, pk)
    esig, epk = ref.ecdsa_make(sk, msg)
    esig = np.array(esig, dtype=np.uint8).copy()
    for i in range(2, n, 5):
        esig[64 * i + 40] ^= 2
    e_exp = ref.ecdsa_verify(esig, msg, epk)
    xo, ec, es = [], [], []
    for i in range(n):
        o = ctypes.create_string_buffer(64)
        assert lib.secp256k1_xonly_pubkey_parse(ctx, o, bytes(pk[32 * i:32 * i + 32])) == 1
        xo.append(o.raw)
        o = ctypes.create_string_buffer(64)
        assert lib.secp256k1_ec_pubkey_parse(ctx, o, bytes(epk[33 * i:33 * i + 33]), 33) == 1
        ec.append(o.raw)
        o = ctypes.create_string_buffer(64)
        lib.secp256k1_ecdsa_signature_parse_compact(ctx, o, bytes(esig[64 * i:64 * i + 64]))
        es.append(o.raw)
    sigs = [bytes(sig[64 * i:64 * i + 64]) for i in range(n)]
    msgs = [msg[32 * i:32 * i + 32] for i in range(n)]
    return sigs, msgs, xo, s_exp, es, ec, e_exp


def _hammer(lib, ctx, wl, threads, calls_per_thread):
    sigs, msgs, xo, s_exp, es, ec, e_exp = wl
    n = len(sigs)
    bad = []

    def body(t):
        for k in range(calls_per_thread):
            i = (t * 7919 + k * 31) % n
            if k & 1:
                r = lib.secp256k1_schnorrsig_verify(ctx, sigs[i], msgs[i], 32, xo[i])
                if r != int(s_exp[i]):
                    bad.append(("schnorr", i, r))
            else:
                r = lib.secp256k1_ecdsa_verify(ctx, es[i], msgs[i], ec[i])
                if r != int(e_exp[i]):
                    bad.append(("ecdsa", i, r))
    ts = [threading.Thread(target=body, args=(t,)) for t in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    return bad


def test_threads_direct_and_coalesced_match_reference(lib, ref):
    ctx = ctypes.c_void_p(lib.secp256k1_context_create(1))
    wl = _workload(ref, lib, ctx, 2000)
    assert 0 < int(wl[3].sum()) < 2000 and 0 < int(wl[6].sum()) < 2000
    lib.secp256k1_b200_coalesce_configure(0, 0)
    assert _hammer(lib, ctx, wl, threads=32, calls_per_thread=60) == []           # direct path: one launch per call, serialized by the engine
    lib.secp256k1_b200_coalesce_configure(50, 4096)
    try:
        assert _hammer(lib, ctx, wl, threads=32, calls_per_thread=10000) == []    # 320 000 mixed calls through shared batch launches
        assert _hammer(lib, ctx, wl, threads=1, calls_per_thread=50) == []        # a lone caller still gets served (after the window)
    finally:
        lib.secp256k1_b200_coalesce_configure(0, 0)
    lib.secp256k1_context_destroy(ctx)


def test_coalesced_rangeproof_verify(lib, ref):
    gen = ref.generator_h()
    n = 24
    blinds = b"".join(hashlib.sha256(b"thr-b%d" % i).digest() for i in range(n))
    vals = [1000 + 37 * i for i in range(n)]
    proofs, plen, commits = ref.rangeproof_make(blinds, vals, gen, min_bits=16)       # short proofs of different lengths are fine too
    proofs = np.array(proofs, dtype=np.uint8)
    for i in range(1, n, 4):
        proofs[i, 70] ^= 1
    eok, emn, emx = ref.rangeproof_verify(np.asarray(commits).reshape(n, 33), proofs, plen, gen)
    ctx = ctypes.c_void_p(lib.secp256k1_context_create(1))
    cm = []
    for i in range(n):
        o = ctypes.create_string_buffer(64)
        assert lib.secp256k1_pedersen_commitment_parse(ctx, o, bytes(np.asarray(commits).reshape(n, 33)[i])) == 1
        cm.append(o.raw)
    lib.secp256k1_b200_coalesce_configure(200, 64)
    res = [None] * n
    try:
        def body(i):
            mn, mx = ctypes.c_uint64(7), ctypes.c_uint64(7)
            r = lib.secp256k1_rangeproof_verify(ctx, ctypes.byref(mn), ctypes.byref(mx), cm[i], bytes(proofs[i][:int(plen[i])]), int(plen[i]), None, 0, gen)
            res[i] = (r, mn.value, mx.value)
        ts = [threading.Thread(target=body, args=(i,)) for i in range(n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    finally:
        lib.secp256k1_b200_coalesce_configure(0, 0)
    assert [r[0] for r in res] == [int(x) for x in eok]
    assert [r[1] for r in res] == [int(x) for x in emn] and [r[2] for r in res] == [int(x) for x in emx]
    lib.secp256k1_context_destroy(ctx)


def test_engine_failure_reaches_error_callback(lib, ref):
    ctx = ctypes.c_void_p(lib.secp256k1_context_create(1))
    hits = []
    cb = CB(lambda msg, data: hits.append(msg))
    lib.secp256k1_context_set_error_callback(ctx, cb, None)
    wl = _workload(ref, lib, ctx, 5)
    lib.secp256k1_b200_test_fail_next_gpu_calls(1)
    assert lib.secp256k1_schnorrsig_verify(ctx, wl[0][1], wl[1][1], 32, wl[2][1]) == 0 and len(hits) == 1
    assert lib.secp256k1_schnorrsig_verify(ctx, wl[0][1], wl[1][1], 32, wl[2][1]) == int(wl[3][1]) and len(hits) == 1    # next call is healthy again
    lib.secp256k1_context_destroy(ctx)


def test_static_and_cloned_contexts_and_selftest(lib, ref):
    lib.secp256k1_selftest()                                        # device known-answer test: aborts the process on failure
    static = ctypes.c_void_p.in_dll(lib, "secp256k1_context_static")
    ctx = ctypes.c_void_p(lib.secp256k1_context_create(1))
    wl = _workload(ref, lib, ctx, 6)
    clone = ctypes.c_void_p(lib.secp256k1_context_clone(ctx))
    for c in (static, clone):
        for i in range(6):
            assert lib.secp256k1_schnorrsig_verify(c, wl[0][i], wl[1][i], 32, wl[2][i]) == int(wl[3][i])
            assert lib.secp256k1_ecdsa_verify(c, wl[4][i], wl[1][i], wl[5][i]) == int(wl[6][i])
    lib.secp256k1_context_destroy(clone)
    # the reference's order of checks: a high s (ECDSA) or r >= p / s >= n (Schnorr) returns 0 BEFORE pubkey_load's ARG_CHECK on a zeroed key
    illegal = []
    cb = CB(lambda msg, data: illegal.append(msg))
    lib.secp256k1_context_set_illegal_callback(ctx, cb, None)
    zero_pk = bytes(64)
    high_s = bytes(32) + (N - 1).to_bytes(32, "little")              # opaque signature: r = 0 (LE), s = n - 1 (LE) > n/2
    assert lib.secp256k1_ecdsa_verify(ctx, high_s, wl[1][0], zero_pk) == 0 and illegal == []
    assert lib.secp256k1_schnorrsig_verify(ctx, b"\xff" * 64, wl[1][0], 32, zero_pk) == 0 and illegal == []
    assert lib.secp256k1_schnorrsig_verify(ctx, wl[0][1], wl[1][1], 32, zero_pk) == 0 and len(illegal) == 1    # in range: now the ARG_CHECK fires
    lib.secp256k1_context_destroy(ctx)
