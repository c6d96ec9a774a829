"""ctypes access to oracle/_ref/libsecp256k1_ref.so (the unmodified reference + oracle/ref_shim.c).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libsecp256k1_ref.so")

SECP256K1_CONTEXT_NONE = 1


def _buf(a):
    if a is None:
        return None
    if isinstance(a, (bytes, bytearray)):
        return bytes(a)
    return np.ascontiguousarray(a, dtype=np.uint8).tobytes()


class RefLib:
    def __init__(self, path=REF_SO):
        self.lib = ctypes.CDLL(path)
        L = self.lib
        L.secp256k1_context_create.restype = ctypes.c_void_p
        L.secp256k1_context_create.argtypes = [ctypes.c_uint]
        self.ctx = ctypes.c_void_p(L.secp256k1_context_create(SECP256K1_CONTEXT_NONE))
        for name in ("ref_ecmult_batch", "ref_schnorrsig_verify_batch", "ref_ecdsa_verify_batch"):
            getattr(L, name).restype = None

    # ---- field / scalar helpers (32-byte big-endian in/out) ----
    def fe2(self, name, a, b):
        r = ctypes.create_string_buffer(32)
        getattr(self.lib, name)(r, a, b)
        return r.raw

    def fe1(self, name, a):
        r = ctypes.create_string_buffer(32)
        ret = getattr(self.lib, name)(r, a)
        return r.raw, ret

    # ---- ecmult ----
    def ecmult_batch(self, a64, na32, ng32=None):
        na32 = _buf(na32)
        n = len(na32) // 32
        out = ctypes.create_string_buffer(64 * n)
        self.lib.ref_ecmult_batch(out, _buf(a64), na32, _buf(ng32), ctypes.c_size_t(n))
        return np.frombuffer(out.raw, dtype=np.uint8).reshape(n, 64)

    def unit_op(self, op, a, b=None):
        """reference side of secp256k1_b200_unit_op_batch (same op codes and operand layouts)"""
        scalar = op < 8
        a = np.ascontiguousarray(np.frombuffer(a, dtype=np.uint8) if isinstance(a, (bytes, bytearray)) else a, dtype=np.uint8).reshape(-1)
        n = a.size // (32 if scalar else 96)
        out = np.zeros(n * (32 if (scalar and op != 4) else 64), dtype=np.uint8)
        if b is not None:
            b = np.ascontiguousarray(np.frombuffer(b, dtype=np.uint8) if isinstance(b, (bytes, bytearray)) else b, dtype=np.uint8).reshape(-1)
        self.lib.ref_unit_op_batch(ctypes.c_int(op), out.ctypes.data_as(ctypes.c_void_p), a.ctypes.data_as(ctypes.c_void_p),
                                   b.ctypes.data_as(ctypes.c_void_p) if b is not None else None, ctypes.c_size_t(n))
        return out

    def ecmult_multi(self, sc32, pt64, g_sc32=None, scratch_bytes=0):
        sc32 = _buf(sc32)
        n = len(sc32) // 32
        out = ctypes.create_string_buffer(64)
        ok = self.lib.ref_ecmult_multi(out, _buf(g_sc32), sc32, _buf(pt64), ctypes.c_size_t(n), ctypes.c_size_t(scratch_bytes))
        return ok, out.raw

    # ---- signatures ----
    def schnorr_make(self, sk32, msg32):
        sk32 = _buf(sk32)
        n = len(sk32) // 32
        sig = ctypes.create_string_buffer(64 * n)
        pk = ctypes.create_string_buffer(32 * n)
        assert self.lib.ref_schnorrsig_make_batch(self.ctx, sig, pk, sk32, _buf(msg32), ctypes.c_size_t(n))
        return np.frombuffer(sig.raw, dtype=np.uint8).copy(), np.frombuffer(pk.raw, dtype=np.uint8).copy()

    def schnorr_verify(self, sig64, msg32, pk32):
        sig64 = _buf(sig64)
        n = len(sig64) // 64
        ok = ctypes.create_string_buffer(n)
        self.lib.ref_schnorrsig_verify_batch(self.ctx, ok, sig64, _buf(msg32), _buf(pk32), ctypes.c_size_t(n))
        return np.frombuffer(ok.raw, dtype=np.uint8).copy()

    def schnorr_verify_one(self, sig64, msg, pk32):
        """public API, arbitrary message length; returns (parse_ok, verify_ok)"""
        pk = ctypes.create_string_buffer(64)
        if not self.lib.secp256k1_xonly_pubkey_parse(self.ctx, pk, bytes(pk32)):
            return 0, 0
        m = bytes(msg)
        return 1, self.lib.secp256k1_schnorrsig_verify(self.ctx, bytes(sig64), m if len(m) else None, ctypes.c_size_t(len(m)), pk)

    # ---- pedersen / rangeproof (generator = 64-byte opaque secp256k1_generator = x||y big-endian) ----
    def generator_h(self):
        p = ctypes.c_void_p.in_dll(self.lib, "secp256k1_generator_h")
        return ctypes.string_at(p.value, 64)

    def pedersen_commit(self, blind32, values, gen64):
        blind32 = _buf(blind32)
        n = len(blind32) // 32
        vals = (ctypes.c_uint64 * n)(*[int(v) for v in values])
        ok = ctypes.create_string_buffer(n)
        out = ctypes.create_string_buffer(33 * n)
        self.lib.ref_pedersen_commit_batch.restype = None
        self.lib.ref_pedersen_commit_batch(self.ctx, ok, out, blind32, vals, bytes(gen64), ctypes.c_size_t(n))
        return np.frombuffer(ok.raw, dtype=np.uint8).copy(), np.frombuffer(out.raw, dtype=np.uint8).reshape(n, 33).copy()

    def rangeproof_make(self, blind32, values, gen64, min_value=0, exp=0, min_bits=64, stride=5134):
        blind32 = _buf(blind32)
        n = len(blind32) // 32
        vals = (ctypes.c_uint64 * n)(*[int(v) for v in values])
        proofs = ctypes.create_string_buffer(stride * n)
        plen = (ctypes.c_size_t * n)()
        commit = ctypes.create_string_buffer(33 * n)
        ok = self.lib.ref_rangeproof_make_batch(self.ctx, proofs, ctypes.c_size_t(stride), plen, commit, blind32, vals,
                                                ctypes.c_uint64(min_value), exp, min_bits, bytes(gen64), ctypes.c_size_t(n))
        assert ok == 1
        return (np.frombuffer(proofs.raw, dtype=np.uint8).reshape(n, stride).copy(), np.array(list(plen), dtype=np.uint64),
                np.frombuffer(commit.raw, dtype=np.uint8).reshape(n, 33).copy())

    def rangeproof_make_msg(self, blind32, values, min_values, exps, min_bits, nonce32, msgs, gen64, extra=b"", stride=5134):
        """reference-signed proofs with embedded messages: msgs = list of bytes.  -> (proofs[n,stride], plen[n], commit33[n,33])"""
        n = len(values)
        msg_stride = max([len(m) for m in msgs] + [1])
        mbuf = b"".join(bytes(m).ljust(msg_stride, b"\0") for m in msgs)
        proofs = ctypes.create_string_buffer(stride * n)
        plen = (ctypes.c_size_t * n)()
        commit = ctypes.create_string_buffer(33 * n)
        ok = self.lib.ref_rangeproof_make_msg_batch(
            self.ctx, proofs, ctypes.c_size_t(stride), plen, commit, _buf(blind32), (ctypes.c_uint64 * n)(*[int(v) for v in values]),
            (ctypes.c_uint64 * n)(*[int(v) for v in min_values]), (ctypes.c_int * n)(*exps), (ctypes.c_int * n)(*min_bits), _buf(nonce32), mbuf,
            ctypes.c_size_t(msg_stride), (ctypes.c_size_t * n)(*[len(m) for m in msgs]), bytes(extra) if extra else None, ctypes.c_size_t(len(extra)),
            bytes(gen64), ctypes.c_size_t(n))
        assert ok == 1
        return (np.frombuffer(proofs.raw, dtype=np.uint8).reshape(n, stride).copy(), np.array(list(plen), dtype=np.uint64),
                np.frombuffer(commit.raw, dtype=np.uint8).reshape(n, 33).copy())

    def rangeproof_rewind(self, nonce32, commit33, proofs, plen, gen64, msg_cap=4096, extra=b""):
        """-> ok[n], blind[n,32], value[n], msgs (list of bytes), min[n], max[n]"""
        proofs = np.ascontiguousarray(proofs, dtype=np.uint8)
        n, stride = proofs.shape
        ok = ctypes.create_string_buffer(n); blind = ctypes.create_string_buffer(32 * n)
        val = (ctypes.c_uint64 * n)(); mn = (ctypes.c_uint64 * n)(); mx = (ctypes.c_uint64 * n)()
        msg = ctypes.create_string_buffer(max(msg_cap, 1) * n); ml = (ctypes.c_size_t * n)()
        self.lib.ref_rangeproof_rewind_batch.restype = None
        self.lib.ref_rangeproof_rewind_batch(self.ctx, ok, blind, val, msg, ctypes.c_size_t(msg_cap), ml, mn, mx, _buf(nonce32), _buf(commit33),
                                             proofs.tobytes(), ctypes.c_size_t(stride), (ctypes.c_size_t * n)(*[int(x) for x in plen]),
                                             bytes(extra) if extra else None, ctypes.c_size_t(len(extra)), bytes(gen64), ctypes.c_size_t(n))
        msgs = [msg.raw[msg_cap * i:msg_cap * i + ml[i]] for i in range(n)]
        return (np.frombuffer(ok.raw, dtype=np.uint8).copy(), np.frombuffer(blind.raw, dtype=np.uint8).reshape(n, 32).copy(),
                np.array(list(val), dtype=np.uint64), msgs, np.array(list(mn), dtype=np.uint64), np.array(list(mx), dtype=np.uint64))

    def rangeproof_verify(self, commit33, proofs, plen, gen64, extra=b""):
        proofs = np.ascontiguousarray(proofs, dtype=np.uint8)
        n, stride = proofs.shape
        ok = ctypes.create_string_buffer(n)
        mn = (ctypes.c_uint64 * n)()
        mx = (ctypes.c_uint64 * n)()
        pl = (ctypes.c_size_t * n)(*[int(x) for x in plen])
        self.lib.ref_rangeproof_verify_batch.restype = None
        self.lib.ref_rangeproof_verify_batch(self.ctx, ok, mn, mx, _buf(commit33), proofs.tobytes(), ctypes.c_size_t(stride), pl,
                                             bytes(extra) if extra else None, ctypes.c_size_t(len(extra)), bytes(gen64), ctypes.c_size_t(n))
        return (np.frombuffer(ok.raw, dtype=np.uint8).copy(), np.array(list(mn), dtype=np.uint64), np.array(list(mx), dtype=np.uint64))

    def rangeproof_verify_one(self, commit33, proof, gen64=None, extra=b""):
        gen64 = gen64 or self.generator_h()
        proof = np.frombuffer(bytes(proof), dtype=np.uint8).reshape(1, -1)
        ok, mn, mx = self.rangeproof_verify(bytes(commit33), proof, [proof.shape[1]], gen64, extra)
        return int(ok[0]), int(mn[0]), int(mx[0])

    def ecdsa_make(self, sk32, msg32):
        sk32 = _buf(sk32)
        n = len(sk32) // 32
        sig = ctypes.create_string_buffer(64 * n)
        pk = ctypes.create_string_buffer(33 * n)
        assert self.lib.ref_ecdsa_make_batch(self.ctx, sig, pk, sk32, _buf(msg32), ctypes.c_size_t(n))
        return np.frombuffer(sig.raw, dtype=np.uint8).copy(), np.frombuffer(pk.raw, dtype=np.uint8).copy()

    def ecdsa_sign_recoverable(self, sk32, msg32):
        sk32 = _buf(sk32)
        n = len(sk32) // 32
        sig = ctypes.create_string_buffer(64 * n)
        rid = ctypes.create_string_buffer(n)
        assert self.lib.ref_ecdsa_sign_recoverable_batch(self.ctx, sig, rid, sk32, _buf(msg32), ctypes.c_size_t(n))
        return np.frombuffer(sig.raw, dtype=np.uint8).copy(), np.frombuffer(rid.raw, dtype=np.uint8).copy()

    def ecdsa_recover(self, sig64, recid, msg32):
        sig64 = _buf(sig64)
        n = len(sig64) // 64
        ok = ctypes.create_string_buffer(n)
        out = ctypes.create_string_buffer(64 * n)
        self.lib.ref_ecdsa_recover_batch(self.ctx, ok, out, sig64, _buf(recid), _buf(msg32), ctypes.c_size_t(n))
        return np.frombuffer(ok.raw, dtype=np.uint8).copy(), np.frombuffer(out.raw, dtype=np.uint8).reshape(n, 64).copy()

    def aggregate(self, pk32, msgs32, sigs64):
        n = len(_buf(pk32)) // 32
        agg = ctypes.create_string_buffer(32 * (n + 1))
        ln = ctypes.c_size_t(32 * (n + 1))
        assert self.lib.ref_schnorrsig_aggregate(self.ctx, agg, ctypes.byref(ln), _buf(pk32), _buf(msgs32), _buf(sigs64), ctypes.c_size_t(n))
        return agg.raw[:ln.value]

    def aggverify(self, pk32, msgs32, aggsig):
        n = len(_buf(pk32)) // 32
        return self.lib.ref_schnorrsig_aggverify(self.ctx, _buf(pk32), _buf(msgs32), ctypes.c_size_t(n), _buf(aggsig), ctypes.c_size_t(len(_buf(aggsig))))

    def ecdsa_verify(self, sig64, msg32, pk33):
        sig64 = _buf(sig64)
        n = len(sig64) // 64
        ok = ctypes.create_string_buffer(n)
        self.lib.ref_ecdsa_verify_batch(self.ctx, ok, sig64, _buf(msg32), _buf(pk33), ctypes.c_size_t(n))
        return np.frombuffer(ok.raw, dtype=np.uint8).copy()


PORT_SO = os.path.join(ROOT, "oracle", "libsecp_oracle.so")


class PortLib:
    """oracle/libsecp_oracle.so -- our plain-C restatement (oracle/secp_oracle.c)."""

    def __init__(self, path=PORT_SO):
        self.lib = ctypes.CDLL(path)
        self.lib.orc_ecmult_batch.restype = None
        self.lib.orc_schnorrsig_verify_batch.restype = None

    def ecmult_batch(self, a64, na32, ng32=None):
        na32 = _buf(na32)
        n = len(na32) // 32
        out = ctypes.create_string_buffer(64 * n)
        self.lib.orc_ecmult_batch(out, _buf(a64), na32, _buf(ng32), ctypes.c_size_t(n))
        return np.frombuffer(out.raw, dtype=np.uint8).reshape(n, 64)

    def schnorr_verify(self, sig64, msg32, pk32):
        sig64 = _buf(sig64)
        n = len(sig64) // 64
        ok = ctypes.create_string_buffer(n)
        self.lib.orc_schnorrsig_verify_batch(ok, sig64, _buf(msg32), _buf(pk32), ctypes.c_size_t(n))
        return np.frombuffer(ok.raw, dtype=np.uint8).copy()

    def rangeproof_verify(self, commit33, proof, gen64, extra=b""):
        mn, mx = ctypes.c_uint64(0), ctypes.c_uint64(0)
        proof = bytes(proof)
        ok = self.lib.orc_rangeproof_verify(ctypes.byref(mn), ctypes.byref(mx), bytes(commit33), proof, ctypes.c_size_t(len(proof)),
                                            bytes(extra) if extra else None, ctypes.c_size_t(len(extra)), bytes(gen64))
        return ok, mn.value, mx.value


def surjection_make(ref, n_in, n_use, real, seed32):
    """(serialized proof, ephemeral input tags n_in*64, ephemeral output tag 64) made by the reference, or None"""
    proof = ctypes.create_string_buffer(9000)
    plen = ctypes.c_size_t(9000)
    tin = ctypes.create_string_buffer(64 * n_in)
    tout = ctypes.create_string_buffer(64)
    ok = ref.lib.ref_surjection_make(ref.ctx, proof, ctypes.byref(plen), tin, tout, ctypes.c_size_t(n_in), ctypes.c_size_t(n_use), ctypes.c_size_t(real), bytes(seed32))
    return (proof.raw[:plen.value], tin.raw, tout.raw) if ok else None


def surjection_verify(ref, proof, tin, n_in, tout):
    return ref.lib.ref_surjection_verify(ref.ctx, bytes(proof), ctypes.c_size_t(len(proof)), bytes(tin), ctypes.c_size_t(n_in), bytes(tout))
