"""Single-call latency of the reference-named functions: this library (a batch of one on the GPU) next to the reference on a host core.

    python tools/single_call_latency.py [--coalesce-threads T]     -> JSON lines (p50 / p99 / mean in microseconds)

With --coalesce-threads T it also measures aggregate throughput of T host threads calling secp256k1_schnorrsig_verify concurrently, with the
coalescing front-end off and on (secp256k1_b200_coalesce_configure).  Numbers go to profiles/r2_single_call_latency.md; INTEGRATION.md quotes them."""
import ctypes
import hashlib
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def stats(us):
    us = sorted(us)
    return {"calls": len(us), "p50_us": round(us[len(us) // 2], 1), "p99_us": round(us[int(len(us) * 0.99) - 1], 1), "mean_us": round(sum(us) / len(us), 1)}


def timed(fn, reps):
    out = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        out.append((time.perf_counter() - t0) * 1e6)
    return stats(out)


def api(L):
    L.secp256k1_context_create.restype = ctypes.c_void_p
    L.secp256k1_context_create.argtypes = [ctypes.c_uint]
    ctx = ctypes.c_void_p(L.secp256k1_context_create(1))
    L.secp256k1_schnorrsig_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    L.secp256k1_ecdsa_verify.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_xonly_pubkey_parse.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_ec_pubkey_parse.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
    L.secp256k1_ecdsa_signature_parse_compact.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_pedersen_commit.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
    L.secp256k1_pedersen_commitment_parse.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    L.secp256k1_rangeproof_verify.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_char_p,
                                              ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
    return ctx


def main():
    import numpy as np
    import secp256k1_zkp_b200 as b200
    from tests.refshim import RefLib
    ref = RefLib()
    ours = ctypes.CDLL(b200.library_path())
    libs = {"b200": (ours, api(ours)), "reference": (ref.lib, api(ref.lib))}
    n = 64
    sk = b"".join(hashlib.sha256(b"lat-sk%d" % i).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"lat-msg%d" % i).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    esig, epk = ref.ecdsa_make(sk, msg)
    gen = ref.generator_h()
    proofs, plen, commits = ref.rangeproof_make(sk[:32 * 8], [12345 + i for i in range(8)], gen)
    commits = np.asarray(commits).reshape(-1, 33)
    for name, (L, ctx) in libs.items():
        xo = ctypes.create_string_buffer(64); L.secp256k1_xonly_pubkey_parse(ctx, xo, bytes(pk[:32]))
        ec = ctypes.create_string_buffer(64); L.secp256k1_ec_pubkey_parse(ctx, ec, bytes(epk[:33]), 33)
        es = ctypes.create_string_buffer(64); L.secp256k1_ecdsa_signature_parse_compact(ctx, es, bytes(esig[:64]))
        cm = ctypes.create_string_buffer(64); L.secp256k1_pedersen_commitment_parse(ctx, cm, bytes(commits[0]))
        out = ctypes.create_string_buffer(64)
        mn, mx = ctypes.c_uint64(0), ctypes.c_uint64(0)
        pr = bytes(proofs[0][:int(plen[0])])
        calls = {
            "secp256k1_schnorrsig_verify": lambda: L.secp256k1_schnorrsig_verify(ctx, bytes(sig[:64]), msg[:32], 32, xo),
            "secp256k1_ecdsa_verify": lambda: L.secp256k1_ecdsa_verify(ctx, es, msg[:32], ec),
            "secp256k1_xonly_pubkey_parse": lambda: L.secp256k1_xonly_pubkey_parse(ctx, out, bytes(pk[:32])),
            "secp256k1_ec_pubkey_parse": lambda: L.secp256k1_ec_pubkey_parse(ctx, out, bytes(epk[:33]), 33),
            "secp256k1_pedersen_commit": lambda: L.secp256k1_pedersen_commit(ctx, out, sk[:32], 12345, gen),
            "secp256k1_rangeproof_verify (64-bit)": lambda: L.secp256k1_rangeproof_verify(ctx, ctypes.byref(mn), ctypes.byref(mx), cm, pr, len(pr), None, 0, gen),
        }
        for fname, fn in calls.items():
            assert fn() == 1, (name, fname)
            for _ in range(20):
                fn()
            print(json.dumps(dict({"lib": name, "function": fname}, **timed(fn, 300 if "rangeproof" not in fname else 60))), flush=True)
    if "--coalesce-threads" in sys.argv:
        T = int(sys.argv[sys.argv.index("--coalesce-threads") + 1])
        L, ctx = libs["b200"]
        L.secp256k1_b200_coalesce_configure.argtypes = [ctypes.c_uint, ctypes.c_size_t]
        xos = []
        for i in range(n):
            o = ctypes.create_string_buffer(64); L.secp256k1_xonly_pubkey_parse(ctx, o, bytes(pk[32 * i:32 * i + 32])); xos.append(o.raw)
        sigs = [bytes(sig[64 * i:64 * i + 64]) for i in range(n)]; msgs = [msg[32 * i:32 * i + 32] for i in range(n)]
        for window, per in ((0, 100), (50, 5000), (200, 5000)):
            L.secp256k1_b200_coalesce_configure(window, 4096)

            def body(t):
                for k in range(per):
                    i = (t + k) % n
                    assert L.secp256k1_schnorrsig_verify(ctx, sigs[i], msgs[i], 32, xos[i]) == 1
            ts = [threading.Thread(target=body, args=(t,)) for t in range(T)]
            t0 = time.perf_counter()
            for t in ts:
                t.start()
            for t in ts:
                t.join()
            dt = time.perf_counter() - t0
            print(json.dumps({"lib": "b200", "function": "secp256k1_schnorrsig_verify", "threads": T, "coalesce_window_us": window, "calls": T * per,
                              "calls_per_s": round(T * per / dt), "mean_latency_us": round(dt / per * 1e6, 1)}), flush=True)
        L.secp256k1_b200_coalesce_configure(0, 0)
        Lr, cr = libs["reference"]
        xor = ctypes.create_string_buffer(64); Lr.secp256k1_xonly_pubkey_parse(cr, xor, bytes(pk[:32]))

        def rbody(t):
            for k in range(2000):
                Lr.secp256k1_schnorrsig_verify(cr, sigs[0], msgs[0], 32, xor)
        ts = [threading.Thread(target=rbody, args=(t,)) for t in range(T)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        dt = time.perf_counter() - t0
        print(json.dumps({"lib": "reference", "function": "secp256k1_schnorrsig_verify", "threads": T, "calls": T * 2000, "calls_per_s": round(T * 2000 / dt)}), flush=True)


if __name__ == "__main__":
    main()
