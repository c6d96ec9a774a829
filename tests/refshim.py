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

    def ecdsa_make(self, sk32, msg32):
        sk32 = _buf(sk32)
        n = len(sk32) // 32
        sig = ctypes.create_string_buffer(64 * n)
        pk = ctypes.create_string_buffer(33 * n)
        assert self.lib.ref_ecdsa_make_batch(self.ctx, sig, pk, sk32, _buf(msg32), ctypes.c_size_t(n))
        return np.frombuffer(sig.raw, dtype=np.uint8).copy(), np.frombuffer(pk.raw, dtype=np.uint8).copy()

    def ecdsa_verify(self, sig64, msg32, pk33):
        sig64 = _buf(sig64)
        n = len(sig64) // 64
        ok = ctypes.create_string_buffer(n)
        self.lib.ref_ecdsa_verify_batch(self.ctx, ok, sig64, _buf(msg32), _buf(pk33), ctypes.c_size_t(n))
        return np.frombuffer(ok.raw, dtype=np.uint8).copy()
