"""ctypes mirror of include/secp256k1_b200.h.

Argument meaning and error behaviour follow the reference's C API for the hot path
(include/secp256k1.h, include/secp256k1_schnorrsig.h, ...): every verify returns 1 = valid,
0 = invalid per item; API misuse raises.  Engine/CUDA failures raise B200Error carrying
secp256k1_b200_last_error().
"""
import ctypes
import os

import numpy as np

PK_OPAQUE64, PK_XONLY32, PK_COMPRESSED33 = 0, 1, 2
SIG_OPAQUE64, SIG_COMPACT64 = 0, 1
FE_MUL, FE_SQR, FE_ADD, FE_SUB, FE_HALF, FE_INV, FE_SQRT, FE_NEG = range(8)
FE_INV_FERMAT, FE_IS_SQUARE, FE_IS_SQUARE_SQRT = 8, 9, 10
FE_MUL_PORTABLE, FE_SQR_PORTABLE = 16, 17

_PKLEN = {PK_OPAQUE64: 64, PK_XONLY32: 32, PK_COMPRESSED33: 33}


class B200Error(RuntimeError):
    pass


def library_path():
    if os.environ.get("SECP256K1_B200_LIB"):       # development override (A/B of two builds)
        return os.environ["SECP256K1_B200_LIB"]
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "libsecp256k1_b200.so")


_LIB = None


def load_library():
    """Loads libsecp256k1_b200.so (built in-tree by build.py).  Fails loudly if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise B200Error("libsecp256k1_b200.so is not built (%s); run `python __graft_entry__.py build`. "
                        "There is no CPU fallback." % path)
    lib = ctypes.CDLL(path)
    vp, sz, ci, u8p = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p
    lib.secp256k1_b200_last_error.restype = ctypes.c_char_p
    lib.secp256k1_b200_engine_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci]
    lib.secp256k1_b200_engine_destroy.argtypes = [vp]
    lib.secp256k1_b200_engine_destroy.restype = None
    lib.secp256k1_b200_engine_table_ptr.argtypes = [vp]
    lib.secp256k1_b200_engine_table_ptr.restype = vp
    lib.secp256k1_b200_engine_table_bytes.argtypes = [vp]
    lib.secp256k1_b200_engine_table_bytes.restype = sz
    lib.secp256k1_b200_engine_table_build.argtypes = [vp]
    lib.secp256k1_b200_engine_launch_count.argtypes = [vp]
    lib.secp256k1_b200_engine_launch_count.restype = ctypes.c_uint64
    lib.secp256k1_b200_engine_device.argtypes = [vp]
    lib.secp256k1_b200_fe_op_batch.argtypes = [vp, ci, u8p, u8p, u8p, sz, ci]
    lib.secp256k1_b200_fe_op_batch_dev.argtypes = [vp, ci, u8p, u8p, u8p, sz, ci, vp]
    lib.secp256k1_b200_engine_table_export_dev.argtypes = [vp, vp, vp]
    lib.secp256k1_b200_engine_table_import_dev.argtypes = [vp, vp, vp]
    lib.secp256k1_b200_testdata_schnorrsig_sign_batch.argtypes = [vp, u8p, u8p, u8p, u8p, sz, sz]
    lib.secp256k1_b200_testdata_schnorrsig_sign_batch_dev.argtypes = [vp, u8p, u8p, u8p, u8p, sz, sz, vp]
    u64p = ctypes.c_void_p
    lib.secp256k1_b200_pedersen_commit_batch.argtypes = [vp, u8p, u8p, u8p, u64p, u8p, sz, sz]
    lib.secp256k1_b200_pedersen_commit_batch_dev.argtypes = [vp, u8p, u8p, u8p, u64p, u8p, sz, sz, vp]
    lib.secp256k1_b200_rangeproof_verify_batch.argtypes = [vp, u8p, u64p, u64p, u8p, u8p, sz, vp, u8p, sz, u8p, sz, vp, sz]
    lib.secp256k1_b200_rangeproof_verify_batch_dev.argtypes = [vp, u8p, u64p, u64p, u8p, u8p, sz, vp, u8p, sz, u8p, sz, vp, sz, vp]
    lib.secp256k1_b200_ecmult_multi_workspace_bytes.argtypes = [sz]
    lib.secp256k1_b200_ecmult_multi_workspace_bytes.restype = sz
    lib.secp256k1_b200_ecmult_multi.argtypes = [vp, u8p, u8p, u8p, u8p, sz]
    lib.secp256k1_b200_ecmult_multi_dev.argtypes = [vp, u8p, u8p, u8p, u8p, sz, vp, sz, vp]
    lib.secp256k1_b200_ecmult_multi_partial_dev.argtypes = [vp, u8p, u8p, u8p, u8p, sz, vp, sz, vp]
    lib.secp256k1_b200_ecmult_multi_window_shard_dev.argtypes = [vp, u8p, u8p, u8p, u8p, sz, ctypes.c_uint, ctypes.c_uint, vp, sz, vp]
    lib.secp256k1_b200_ecmult_multi_combine_dev.argtypes = [vp, u8p, u8p, sz, vp]
    lib.secp256k1_b200_ecmult_multi_set_window.argtypes = [vp, ci]
    lib.secp256k1_b200_unit_op_batch.argtypes = [vp, ci, u8p, u8p, u8p, sz]
    lib.secp256k1_b200_point_codec_batch.argtypes = [vp, u8p, u8p, u8p, sz, ctypes.c_uint32, ci, sz]
    lib.secp256k1_b200_sha256_state_init.argtypes = [u8p]
    lib.secp256k1_b200_sha256_state_init.restype = None
    lib.secp256k1_b200_bppp_transcript_init.argtypes = [u8p, u8p, u8p, u8p, sz, sz, u8p, sz]
    lib.secp256k1_b200_bppp_transcript_init.restype = None
    lib.secp256k1_b200_bppp_norm_verify_batch.argtypes = [vp, u8p, u8p, sz, vp, u8p, u8p, u8p, sz, u8p, sz, u8p, sz]
    lib.secp256k1_b200_bppp_norm_verify_batch_dev.argtypes = [vp, u8p, u8p, sz, vp, u8p, u8p, u8p, sz, u8p, sz, u8p, sz, vp]
    lib.secp256k1_b200_surjectionproof_verify_batch.argtypes = [vp, u8p, u8p, sz, vp, u8p, vp, u8p, sz]
    lib.secp256k1_b200_surjectionproof_verify_batch_dev.argtypes = [vp, u8p, u8p, sz, vp, u8p, vp, u8p, sz, vp]
    lib.secp256k1_b200_pedersen_verify_tally_batch.argtypes = [vp, u8p, u8p, vp, u8p, vp, sz]
    lib.secp256k1_b200_pedersen_verify_tally_batch_dev.argtypes = [vp, u8p, u8p, vp, u8p, vp, sz, vp]
    lib.secp256k1_b200_imad_probe_dev.argtypes = [vp, vp, ci, ci, ci, vp]
    lib.secp256k1_b200_ecmult_batch.argtypes = [vp, u8p, u8p, u8p, u8p, sz]
    lib.secp256k1_b200_ecmult_batch_dev.argtypes = [vp, u8p, u8p, u8p, u8p, sz, vp]
    lib.secp256k1_b200_schnorrsig_verify_batch.argtypes = [vp, u8p, u8p, u8p, sz, u8p, ci, sz]
    lib.secp256k1_b200_schnorrsig_verify_batch_dev.argtypes = [vp, u8p, u8p, u8p, sz, u8p, ci, sz, vp]
    lib.secp256k1_b200_ecdsa_verify_batch.argtypes = [vp, u8p, u8p, ci, u8p, u8p, ci, sz]
    lib.secp256k1_b200_ecdsa_verify_batch_dev.argtypes = [vp, u8p, u8p, ci, u8p, u8p, ci, sz, vp]
    lib.secp256k1_b200_rangeproof_rewind_workspace_bytes.argtypes = [sz]
    lib.secp256k1_b200_rangeproof_rewind_workspace_bytes.restype = sz
    lib.secp256k1_b200_rangeproof_rewind_batch.argtypes = [vp, u8p, u8p, vp, u8p, sz, vp, vp, vp, u8p, u8p, u8p, sz, vp, u8p, sz, u8p, sz, vp, sz]
    lib.secp256k1_b200_rangeproof_rewind_batch_dev.argtypes = [vp, u8p, u8p, vp, u8p, sz, vp, vp, vp, u8p, u8p, u8p, sz, vp, u8p, sz, u8p, sz, vp, sz, vp, sz, vp]
    lib.secp256k1_b200_schnorrsig_aggverify.argtypes = [vp, ctypes.POINTER(ci), u8p, u8p, sz, u8p, sz]
    lib.secp256k1_b200_schnorrsig_aggverify_workspace_bytes.argtypes = [sz]
    lib.secp256k1_b200_schnorrsig_aggverify_workspace_bytes.restype = sz
    lib.secp256k1_b200_schnorrsig_aggverify_dev.argtypes = [vp, u8p, u8p, u8p, sz, u8p, vp, sz, vp]
    lib.secp256k1_b200_ecdsa_recover_batch.argtypes = [vp, u8p, u8p, u8p, u8p, u8p, sz]
    lib.secp256k1_b200_ecdsa_recover_batch_dev.argtypes = [vp, u8p, u8p, u8p, u8p, u8p, sz, vp]
    _LIB = lib
    return lib


def _u8(buf, nbytes, name):
    """bytes / bytearray / uint8 ndarray -> contiguous uint8 ndarray of exactly nbytes."""
    a = np.frombuffer(buf, dtype=np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else np.ascontiguousarray(buf, dtype=np.uint8).reshape(-1)
    if a.size != nbytes:
        raise ValueError("%s: expected %d bytes, got %d" % (name, nbytes, a.size))
    return a


def _p(a):
    return a.ctypes.data if a is not None else None


class Engine:
    """One engine per process per GPU (the reference context is stateless for verification; device
    state -- the generator comb table, streams, scratch -- lives here, outside any secp256k1_context)."""

    def __init__(self, device=-1, gen_window_bits=0, build_tables=True):
        self._lib = load_library()
        h = ctypes.c_void_p()
        if not self._lib.secp256k1_b200_engine_create(ctypes.byref(h), device, gen_window_bits, 1 if build_tables else 0):
            raise B200Error(self._err())
        self._h = h

    def _err(self):
        return self._lib.secp256k1_b200_last_error().decode()

    def close(self):
        if getattr(self, "_h", None):
            self._lib.secp256k1_b200_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- table plumbing (NCCL broadcast at init is done by the caller on this buffer) --
    @property
    def table_ptr(self):
        return self._lib.secp256k1_b200_engine_table_ptr(self._h)

    @property
    def table_bytes(self):
        return self._lib.secp256k1_b200_engine_table_bytes(self._h)

    def build_tables(self):
        if not self._lib.secp256k1_b200_engine_table_build(self._h):
            raise B200Error(self._err())

    @property
    def launch_count(self):
        return int(self._lib.secp256k1_b200_engine_launch_count(self._h))

    @property
    def device(self):
        return self._lib.secp256k1_b200_engine_device(self._h)

    def _ck(self, ok):
        if not ok:
            raise B200Error(self._err())

    # -- host-buffer entry points ---------------------------------------------------------------
    def fe_op(self, op, a32, b32=None, iters=1):
        a = _u8(a32, len(a32) if not hasattr(a32, "size") else a32.size, "a32")
        n = a.size // 32
        b = _u8(b32, 32 * n, "b32") if b32 is not None else None
        out = np.empty(32 * n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_fe_op_batch(self._h, op, _p(out), _p(a), _p(b), n, iters))
        return out.reshape(n, 32)

    def ecmult(self, a64, na32, ng32=None, out=None):
        """out: optional preallocated uint8 array of 64*n bytes (e.g. a view of pinned memory: written by the D2H copies directly)"""
        na = _u8(na32, len(na32) if not hasattr(na32, "size") else na32.size, "na32")
        n = na.size // 32
        a = _u8(a64, 64 * n, "a64")
        ng = _u8(ng32, 32 * n, "ng32") if ng32 is not None else None
        out = np.empty(64 * n, dtype=np.uint8) if out is None else out.reshape(-1)
        self._ck(self._lib.secp256k1_b200_ecmult_batch(self._h, _p(out), _p(a), _p(na), _p(ng), n))
        return out.reshape(n, 64)

    def schnorrsig_verify(self, sig64, msg, pk, pkfmt=PK_XONLY32, msglen=32):
        sig = _u8(sig64, len(sig64) if not hasattr(sig64, "size") else sig64.size, "sig64")
        n = sig.size // 64
        m = _u8(msg, msglen * n, "msg") if msglen else np.zeros(1, dtype=np.uint8)
        p = _u8(pk, _PKLEN[pkfmt] * n, "pk")
        ok = np.empty(n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_schnorrsig_verify_batch(self._h, _p(ok), _p(sig), _p(m), msglen, _p(p), pkfmt, n))
        return ok

    def ecdsa_verify(self, sig64, msghash32, pk, sigfmt=SIG_COMPACT64, pkfmt=PK_COMPRESSED33):
        sig = _u8(sig64, len(sig64) if not hasattr(sig64, "size") else sig64.size, "sig64")
        n = sig.size // 64
        m = _u8(msghash32, 32 * n, "msghash32")
        p = _u8(pk, _PKLEN[pkfmt] * n, "pk")
        ok = np.empty(n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_ecdsa_verify_batch(self._h, _p(ok), _p(sig), sigfmt, _p(m), _p(p), pkfmt, n))
        return ok

    def rangeproof_rewind(self, nonce32, commit33, proofs, plen, gen64, msg_cap=4096, extra=None, extra_len=None):
        """secp256k1_rangeproof_rewind over a batch: proofs[n, stride], one 32-byte nonce per proof.
        -> ok[n], blind[n,32], value[n], message[n,msg_cap], message_len[n], min[n], max[n]"""
        proofs = np.ascontiguousarray(proofs, dtype=np.uint8)
        n, stride = proofs.shape
        c = _u8(commit33, 33 * n, "commit33"); nz = _u8(nonce32, 32 * n, "nonce32")
        pl = np.ascontiguousarray(plen, dtype=np.uint32)
        g = _u8(gen64, len(gen64) if not hasattr(gen64, "size") else gen64.size, "gen64")
        gen_stride = 0 if g.size == 64 else 64
        ok = np.zeros(n, dtype=np.uint8); blind = np.zeros((n, 32), dtype=np.uint8); val = np.zeros(n, dtype=np.uint64)
        mn = np.zeros(n, dtype=np.uint64); mx = np.zeros(n, dtype=np.uint64)
        msg = np.zeros((n, max(msg_cap, 1)), dtype=np.uint8) if msg_cap else None
        ml = np.zeros(n, dtype=np.uint32)
        ex = np.ascontiguousarray(extra, dtype=np.uint8) if extra is not None else None
        exl = np.ascontiguousarray(extra_len, dtype=np.uint32) if extra_len is not None else None
        self._ck(self._lib.secp256k1_b200_rangeproof_rewind_batch(
            self._h, _p(ok), _p(blind), _p(val), _p(msg), msg_cap, _p(ml), _p(mn), _p(mx), _p(nz), _p(c), _p(proofs), stride, _p(pl), _p(g), gen_stride,
            _p(ex), ex.shape[1] if ex is not None else 0, _p(exl), n))
        return ok, blind, val, msg, ml, mn, mx

    def schnorrsig_aggverify(self, pubkeys32, msgs32, aggsig):
        """secp256k1_schnorrsig_aggverify: one half-aggregate signature over n (key, message) pairs -> 0/1"""
        agg = _u8(aggsig, len(aggsig) if not hasattr(aggsig, "size") else aggsig.size, "aggsig")
        pk = np.ascontiguousarray(np.frombuffer(pubkeys32, dtype=np.uint8) if isinstance(pubkeys32, (bytes, bytearray)) else pubkeys32, dtype=np.uint8).reshape(-1)
        n = pk.size // 32
        m = _u8(msgs32, 32 * n, "msgs32") if n else np.zeros(1, dtype=np.uint8)
        if n == 0:
            pk = np.zeros(1, dtype=np.uint8)
        ok = ctypes.c_int(0)
        self._ck(self._lib.secp256k1_b200_schnorrsig_aggverify(self._h, ctypes.byref(ok), _p(pk), _p(m), n, _p(agg), agg.size))
        return ok.value

    def schnorrsig_aggverify_workspace_bytes(self, n):
        return int(self._lib.secp256k1_b200_schnorrsig_aggverify_workspace_bytes(n))

    def schnorrsig_aggverify_dev(self, d_ok, d_pk32, d_msgs32, n, d_aggsig, d_ws, ws_bytes, stream=0):
        self._ck(self._lib.secp256k1_b200_schnorrsig_aggverify_dev(self._h, d_ok, d_pk32, d_msgs32, n, d_aggsig, d_ws, ws_bytes, stream))

    def ecdsa_recover(self, sig64, recid, msghash32, out=None):
        """(ok[n], pubkey64[n,64]) -- secp256k1_ecdsa_recover on compact signatures + recovery ids"""
        sig = _u8(sig64, len(sig64) if not hasattr(sig64, "size") else sig64.size, "sig64")
        n = sig.size // 64
        rid = _u8(recid, n, "recid")
        m = _u8(msghash32, 32 * n, "msghash32")
        ok = np.empty(n, dtype=np.uint8)
        out = np.empty((n, 64), dtype=np.uint8) if out is None else out.reshape(n, 64)
        self._ck(self._lib.secp256k1_b200_ecdsa_recover_batch(self._h, _p(ok), _p(out), _p(sig), _p(rid), _p(m), n))
        return ok, out

    def ecmult_multi(self, scalars32, points64, g_scalar32=None):
        """sum_i scalars[i]*points[i] (+ g_scalar*G); returns 64 bytes x||y (zeros = infinity)"""
        sc = _u8(scalars32, len(scalars32) if not hasattr(scalars32, "size") else scalars32.size, "scalars32")
        n = sc.size // 32
        pt = _u8(points64, 64 * n, "points64")
        g = _u8(g_scalar32, 32, "g_scalar32") if g_scalar32 is not None else None
        out = np.zeros(64, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_ecmult_multi(self._h, _p(out), _p(g), _p(sc) if n else None, _p(pt) if n else None, n))
        return out.tobytes()

    def ecmult_multi_workspace_bytes(self, n):
        return int(self._lib.secp256k1_b200_ecmult_multi_workspace_bytes(n))

    def ecmult_multi_dev(self, d_out64, d_g_scalar32, d_scalars32, d_points64, n, d_workspace, workspace_bytes, stream=0):
        self._ck(self._lib.secp256k1_b200_ecmult_multi_dev(self._h, d_out64, d_g_scalar32, d_scalars32, d_points64, n, d_workspace, workspace_bytes, stream))

    def ecmult_multi_partial_dev(self, d_out96, d_g_scalar32, d_scalars32, d_points64, n, d_workspace, workspace_bytes, stream=0):
        """one shard of a sharded MSM: 96-byte Jacobian partial sum (no inversion); combine with ecmult_multi_combine_dev"""
        self._ck(self._lib.secp256k1_b200_ecmult_multi_partial_dev(self._h, d_out96, d_g_scalar32, d_scalars32, d_points64, n, d_workspace, workspace_bytes, stream))

    def ecmult_multi_window_shard_dev(self, d_out96, d_g_scalar32, d_scalars32, d_points64, n, shard, shards, d_workspace, workspace_bytes, stream=0):
        """one WINDOW shard of an MSM over all n terms (bucket windows wi with wi % shards == shard): 96-byte Jacobian partial result"""
        self._ck(self._lib.secp256k1_b200_ecmult_multi_window_shard_dev(self._h, d_out96, d_g_scalar32, d_scalars32, d_points64, n, shard, shards,
                                                                        d_workspace, workspace_bytes, stream))

    def ecmult_multi_combine_dev(self, d_out64, d_partials96, count, stream=0):
        self._ck(self._lib.secp256k1_b200_ecmult_multi_combine_dev(self._h, d_out64, d_partials96, count, stream))

    def ecmult_multi_set_window(self, window_bits):
        """0 = bucket window chosen from n (default); 4..16 forces it"""
        self._ck(self._lib.secp256k1_b200_ecmult_multi_set_window(self._h, window_bits))

    def unit_op(self, op, a, b=None):
        """unit kernels of the scalar / group primitives (secp256k1_b200_unit_op_batch)"""
        scalar = op < 8
        a = _u8(a, len(a) if not hasattr(a, "size") else a.size, "a")
        n = a.size // (32 if scalar else 96)
        bb = _u8(b, a.size, "b") if b is not None else None
        out = np.empty(n * (32 if (scalar and op != 4) else 64), dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_unit_op_batch(self._h, op, _p(out), _p(a), _p(bb), n))
        return out

    def point_codec(self, data, in_len, fmt):
        a = _u8(data, len(data) if not hasattr(data, "size") else data.size, "data")
        n = a.size // in_len
        ok = np.empty(n, dtype=np.uint8)
        out = np.empty(64 * n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_point_codec_batch(self._h, _p(ok), _p(out), _p(a), in_len, in_len, fmt, n))
        return ok, out.reshape(n, 64)

    def bppp_transcripts(self, commits64, rho32, gens64, g_len, c_vec32, h_len, mode="tests"):
        """104-byte SHA-256 transcript objects: mode 'plain' = fresh SHA-256 (the reference's verify vectors),
        'tests' = commit to the initial public data like modules/bppp/tests_impl.h:273-311 (host-side hashing)"""
        cm = _u8(commits64, len(commits64) if not hasattr(commits64, "size") else commits64.size, "commits64")
        n = cm.size // 64
        out = np.zeros((n, 104), dtype=np.uint8)
        if mode == "plain":
            self._lib.secp256k1_b200_sha256_state_init(_p(out))
            out[:] = out[0]
            return out
        rho = _u8(rho32, 32 * n, "rho32")
        g = _u8(gens64, 64 * (g_len + h_len), "gens64")
        cv = _u8(c_vec32, 32 * h_len * n, "c_vec32")
        for i in range(n):
            self._lib.secp256k1_b200_bppp_transcript_init(out[i].ctypes.data, cm[64 * i:].ctypes.data, rho[32 * i:].ctypes.data, _p(g), g_len,
                                                          g_len + h_len, cv[32 * h_len * i:].ctypes.data, h_len)
        return out

    def bppp_norm_verify(self, proofs, plen, transcripts104, rho32, gens64, g_len, c_vec32, h_len, commits64):
        pr = np.ascontiguousarray(proofs, dtype=np.uint8)
        n, stride = pr.shape
        pl = np.ascontiguousarray(plen, dtype=np.uint32).reshape(-1)
        tr = _u8(transcripts104, 104 * n, "transcripts104")
        rho = _u8(rho32, 32 * n, "rho32")
        g = _u8(gens64, 64 * (g_len + h_len), "gens64")
        cv = _u8(c_vec32, 32 * h_len * n, "c_vec32")
        cm = _u8(commits64, 64 * n, "commits64")
        ok = np.empty(n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_bppp_norm_verify_batch(self._h, _p(ok), _p(pr), stride, _p(pl), _p(tr), _p(rho), _p(g), g_len, _p(cv), h_len, _p(cm), n))
        return ok

    def surjectionproof_verify(self, proofs, plen, in_tags64, tag_offset, out_tags64):
        """proofs (n, stride) serialized; in_tags64 flat (ntags*64); tag_offset n+1 entries; out_tags64 (n*64)"""
        pr = np.ascontiguousarray(proofs, dtype=np.uint8)
        n, stride = pr.shape
        pl = np.ascontiguousarray(plen, dtype=np.uint32).reshape(-1)
        off = np.ascontiguousarray(tag_offset, dtype=np.uint32).reshape(-1)
        tags = _u8(in_tags64, 64 * int(off[-1]), "in_tags64") if int(off[-1]) else np.zeros(64, dtype=np.uint8)
        outs = _u8(out_tags64, 64 * n, "out_tags64")
        ok = np.empty(n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_surjectionproof_verify_batch(self._h, _p(ok), _p(pr), stride, _p(pl), _p(tags), _p(off), _p(outs), n))
        return ok

    def pedersen_verify_tally(self, pos33, pos_offset, neg33, neg_offset):
        po = np.ascontiguousarray(pos_offset, dtype=np.uint32).reshape(-1)
        no = np.ascontiguousarray(neg_offset, dtype=np.uint32).reshape(-1)
        n = po.size - 1
        p = _u8(pos33, 33 * int(po[-1]), "pos33") if int(po[-1]) else np.zeros(33, dtype=np.uint8)
        q = _u8(neg33, 33 * int(no[-1]), "neg33") if int(no[-1]) else np.zeros(33, dtype=np.uint8)
        ok = np.empty(n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_pedersen_verify_tally_batch(self._h, _p(ok), _p(p), _p(po), _p(q), _p(no), n))
        return ok

    def pedersen_commit(self, blind32, values, gen64, out=None):
        """gen64: 64 bytes (shared) or n*64 bytes (one generator per item); out: optional preallocated 33*n uint8 array"""
        b = _u8(blind32, len(blind32) if not hasattr(blind32, "size") else blind32.size, "blind32")
        n = b.size // 32
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1)
        if v.size != n:
            raise ValueError("values: expected %d entries" % n)
        g = _u8(gen64, len(gen64) if not hasattr(gen64, "size") else gen64.size, "gen64")
        stride = 0 if g.size == 64 else 64
        if stride and g.size != 64 * n:
            raise ValueError("gen64: expected 64 or %d bytes" % (64 * n))
        ok = np.empty(n, dtype=np.uint8)
        out = np.empty(33 * n, dtype=np.uint8) if out is None else out.reshape(-1)
        self._ck(self._lib.secp256k1_b200_pedersen_commit_batch(self._h, _p(ok), _p(out), _p(b), _p(v), _p(g), stride, n))
        return ok, out.reshape(n, 33)

    def rangeproof_verify(self, commit33, proofs, plen, gen64, extra=None, extra_len=None):
        """proofs: uint8 array (n, stride); plen: n lengths; gen64: 64 or n*64 bytes; extra: (n, estride) or None"""
        pr = np.ascontiguousarray(proofs, dtype=np.uint8)
        n, stride = pr.shape
        c = _u8(commit33, 33 * n, "commit33")
        pl = np.ascontiguousarray(plen, dtype=np.uint32).reshape(-1)
        g = _u8(gen64, len(gen64) if not hasattr(gen64, "size") else gen64.size, "gen64")
        gstride = 0 if g.size == 64 else 64
        ex = el = None
        estride = 0
        if extra is not None:
            ex = np.ascontiguousarray(extra, dtype=np.uint8)
            estride = ex.shape[1]
            el = np.ascontiguousarray(extra_len, dtype=np.uint32).reshape(-1)
        ok = np.empty(n, dtype=np.uint8)
        mn = np.zeros(n, dtype=np.uint64)
        mx = np.zeros(n, dtype=np.uint64)
        self._ck(self._lib.secp256k1_b200_rangeproof_verify_batch(self._h, _p(ok), _p(mn), _p(mx), _p(c), _p(pr), stride, _p(pl), _p(g), gstride,
                                                                  _p(ex), estride, _p(el), n))
        return ok, mn, mx

    def rangeproof_verify_dev(self, d_ok, d_min, d_max, d_commit33, d_proofs, stride, d_plen, d_gen64, gen_stride, n, stream=0,
                              d_extra=None, extra_stride=0, d_extra_len=None):
        self._ck(self._lib.secp256k1_b200_rangeproof_verify_batch_dev(self._h, d_ok, d_min, d_max, d_commit33, d_proofs, stride, d_plen, d_gen64,
                                                                      gen_stride, d_extra, extra_stride, d_extra_len, n, stream))

    def pedersen_commit_dev(self, d_ok, d_commit33, d_blind32, d_value, d_gen64, gen_stride, n, stream=0):
        self._ck(self._lib.secp256k1_b200_pedersen_commit_batch_dev(self._h, d_ok, d_commit33, d_blind32, d_value, d_gen64, gen_stride, n, stream))

    def testdata_schnorrsig_sign(self, sk32, msg, msglen=32):
        sk = _u8(sk32, len(sk32) if not hasattr(sk32, "size") else sk32.size, "sk32")
        n = sk.size // 32
        m = _u8(msg, msglen * n, "msg") if msglen else np.zeros(1, dtype=np.uint8)
        sig = np.empty(64 * n, dtype=np.uint8)
        pk = np.empty(32 * n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_testdata_schnorrsig_sign_batch(self._h, _p(sig), _p(pk), _p(sk), _p(m), msglen, n))
        return sig, pk

    def testdata_schnorrsig_sign_dev(self, d_sig64, d_pk32, d_sk32, d_msg, msglen, n, stream=0):
        self._ck(self._lib.secp256k1_b200_testdata_schnorrsig_sign_batch_dev(self._h, d_sig64, d_pk32, d_sk32, d_msg, msglen, n, stream))

    def table_export_dev(self, d_dst, stream=0):
        self._ck(self._lib.secp256k1_b200_engine_table_export_dev(self._h, d_dst, stream))

    def table_import_dev(self, d_src, stream=0):
        self._ck(self._lib.secp256k1_b200_engine_table_import_dev(self._h, d_src, stream))

    # -- device-pointer entry points (ints: tensor.data_ptr(); stream: torch stream .cuda_stream) ---
    def fe_op_dev(self, op, d_out, d_a, d_b, n, iters=1, stream=0):
        self._ck(self._lib.secp256k1_b200_fe_op_batch_dev(self._h, op, d_out, d_a, d_b, n, iters, stream))

    def imad_probe_dev(self, d_out, blocks, threads, iters, stream=0):
        self._ck(self._lib.secp256k1_b200_imad_probe_dev(self._h, d_out, blocks, threads, iters, stream))

    def ecmult_dev(self, d_out64, d_a64, d_na32, d_ng32, n, stream=0):
        self._ck(self._lib.secp256k1_b200_ecmult_batch_dev(self._h, d_out64, d_a64, d_na32, d_ng32, n, stream))

    def schnorrsig_verify_dev(self, d_ok, d_sig64, d_msg, msglen, d_pk, pkfmt, n, stream=0):
        self._ck(self._lib.secp256k1_b200_schnorrsig_verify_batch_dev(self._h, d_ok, d_sig64, d_msg, msglen, d_pk, pkfmt, n, stream))

    def ecdsa_recover_dev(self, d_ok, d_pubkey64, d_sig64, d_recid, d_msg32, n, stream=0):
        self._ck(self._lib.secp256k1_b200_ecdsa_recover_batch_dev(self._h, d_ok, d_pubkey64, d_sig64, d_recid, d_msg32, n, stream))

    def ecdsa_verify_dev(self, d_ok, d_sig64, sigfmt, d_msg32, d_pk, pkfmt, n, stream=0):
        self._ck(self._lib.secp256k1_b200_ecdsa_verify_batch_dev(self._h, d_ok, d_sig64, sigfmt, d_msg32, d_pk, pkfmt, n, stream))


class MultiEngine:
    """One process, several GPUs: ctypes mirror of the secp256k1_b200_multi_* calls (include/secp256k1_b200.h).  Host arrays are sharded by
    contiguous index ranges over the devices inside the C library; results have the semantics of the single-engine calls."""

    def __init__(self, devices=None, n_devices=0, gen_window_bits=0):
        self._lib = load_library()
        L, vp, sz, ci = self._lib, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.secp256k1_b200_multi_create.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(ci), ci, ci]
        L.secp256k1_b200_multi_destroy.argtypes = [vp]
        L.secp256k1_b200_multi_destroy.restype = None
        L.secp256k1_b200_multi_device_count.argtypes = [vp]
        L.secp256k1_b200_multi_nccl_version.argtypes = [vp]
        L.secp256k1_b200_multi_engine.argtypes = [vp, ci]
        L.secp256k1_b200_multi_engine.restype = vp
        L.secp256k1_b200_multi_launch_count.argtypes = [vp]
        L.secp256k1_b200_multi_launch_count.restype = ctypes.c_uint64
        L.secp256k1_b200_multi_schnorrsig_verify_batch.argtypes = [vp, vp, vp, vp, sz, vp, ci, sz]
        L.secp256k1_b200_multi_ecdsa_verify_batch.argtypes = [vp, vp, vp, ci, vp, vp, ci, sz]
        L.secp256k1_b200_multi_ecmult_batch.argtypes = [vp, vp, vp, vp, vp, sz]
        L.secp256k1_b200_multi_pedersen_commit_batch.argtypes = [vp, vp, vp, vp, vp, vp, sz, sz]
        L.secp256k1_b200_multi_rangeproof_verify_batch.argtypes = [vp, vp, vp, vp, vp, vp, sz, vp, vp, sz, vp, sz, vp, sz]
        L.secp256k1_b200_multi_bppp_norm_verify_batch.argtypes = [vp, vp, vp, sz, vp, vp, vp, vp, sz, vp, sz, vp, sz]
        L.secp256k1_b200_multi_ecmult_multi.argtypes = [vp, vp, vp, vp, vp, sz]
        h = vp()
        if devices is not None:
            arr = (ci * len(devices))(*devices)
            ok = L.secp256k1_b200_multi_create(ctypes.byref(h), arr, len(devices), gen_window_bits)
        else:
            ok = L.secp256k1_b200_multi_create(ctypes.byref(h), None, n_devices, gen_window_bits)
        if not ok:
            raise B200Error(L.secp256k1_b200_last_error().decode())
        self._h = h

    def _ck(self, ok):
        if not ok:
            raise B200Error(self._lib.secp256k1_b200_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._lib.secp256k1_b200_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_count(self):
        return self._lib.secp256k1_b200_multi_device_count(self._h)

    @property
    def nccl_version(self):
        return self._lib.secp256k1_b200_multi_nccl_version(self._h)

    @property
    def launch_count(self):
        return int(self._lib.secp256k1_b200_multi_launch_count(self._h))

    def schnorrsig_verify(self, sig64, msg, pk, pkfmt=PK_XONLY32, msglen=32, out=None):
        sig = _u8(sig64, len(sig64) if not hasattr(sig64, "size") else sig64.size, "sig64")
        n = sig.size // 64
        m = _u8(msg, msglen * n, "msg") if msglen else np.zeros(1, dtype=np.uint8)
        p = _u8(pk, _PKLEN[pkfmt] * n, "pk")
        ok = out if out is not None else np.empty(n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_multi_schnorrsig_verify_batch(self._h, _p(ok), _p(sig), _p(m), msglen, _p(p), pkfmt, n))
        return ok

    def ecdsa_verify(self, sig64, msghash32, pk, sigfmt=SIG_COMPACT64, pkfmt=PK_COMPRESSED33):
        sig = _u8(sig64, len(sig64) if not hasattr(sig64, "size") else sig64.size, "sig64")
        n = sig.size // 64
        m = _u8(msghash32, 32 * n, "msghash32")
        p = _u8(pk, _PKLEN[pkfmt] * n, "pk")
        ok = np.empty(n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_multi_ecdsa_verify_batch(self._h, _p(ok), _p(sig), sigfmt, _p(m), _p(p), pkfmt, n))
        return ok

    def ecmult(self, a64, na32, ng32=None):
        na = _u8(na32, len(na32) if not hasattr(na32, "size") else na32.size, "na32")
        n = na.size // 32
        a = _u8(a64, 64 * n, "a64")
        ng = _u8(ng32, 32 * n, "ng32") if ng32 is not None else None
        out = np.empty(64 * n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_multi_ecmult_batch(self._h, _p(out), _p(a), _p(na), _p(ng), n))
        return out.reshape(n, 64)

    def pedersen_commit(self, blind32, values, gen64):
        b = _u8(blind32, len(blind32) if not hasattr(blind32, "size") else blind32.size, "blind32")
        n = b.size // 32
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1)
        g = _u8(gen64, len(gen64) if not hasattr(gen64, "size") else gen64.size, "gen64")
        ok = np.empty(n, dtype=np.uint8)
        out = np.zeros(33 * n, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_multi_pedersen_commit_batch(self._h, _p(ok), _p(out), _p(b), _p(v), _p(g), 0 if g.size == 64 else 64, n))
        return ok, out.reshape(n, 33)

    def rangeproof_verify(self, commit33, proofs, plen, gen64):
        pr = np.ascontiguousarray(proofs, dtype=np.uint8)
        n, stride = pr.shape
        c = _u8(commit33, 33 * n, "commit33")
        pl = np.ascontiguousarray(plen, dtype=np.uint32).reshape(-1)
        g = _u8(gen64, len(gen64) if not hasattr(gen64, "size") else gen64.size, "gen64")
        ok = np.empty(n, dtype=np.uint8)
        mn = np.zeros(n, dtype=np.uint64)
        mx = np.zeros(n, dtype=np.uint64)
        self._ck(self._lib.secp256k1_b200_multi_rangeproof_verify_batch(self._h, _p(ok), _p(mn), _p(mx), _p(c), _p(pr), stride, _p(pl), _p(g),
                                                                        0 if g.size == 64 else 64, None, 0, None, n))
        return ok, mn, mx

    def ecmult_multi(self, scalars32, points64, g_scalar32=None):
        sc = _u8(scalars32, len(scalars32) if not hasattr(scalars32, "size") else scalars32.size, "scalars32")
        n = sc.size // 32
        pt = _u8(points64, 64 * n, "points64")
        g = _u8(g_scalar32, 32, "g_scalar32") if g_scalar32 is not None else None
        out = np.zeros(64, dtype=np.uint8)
        self._ck(self._lib.secp256k1_b200_multi_ecmult_multi(self._h, _p(out), _p(g), _p(sc) if n else None, _p(pt) if n else None, n))
        return out.tobytes()
