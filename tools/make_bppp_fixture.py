"""tests/golden/bppp_fixture.npz: 64 norm-product arguments at the 64-bit geometry (16 G + 8 H generators, 324-byte proofs)
made by the UNMODIFIED reference prover through oracle/_ref (transcript as modules/bppp/tests_impl.h:273-311).
bench.py tiles them to the batch size."""
import ctypes
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.refshim import RefLib  # noqa: E402

N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
ref = RefLib()
L = ref.lib
L.secp256k1_bppp_generators_create.restype = ctypes.c_void_p
g, h, count = 16, 8, 64
gs = ctypes.c_void_p(L.secp256k1_bppp_generators_create(ref.ctx, ctypes.c_size_t(g + h)))
gxy = ctypes.create_string_buffer(64 * (g + h))
L.ref_bppp_generators_xy(gxy, gs)


def sc(tag, i, j=0):
    return (int.from_bytes(hashlib.sha256(b"%s|%d|%d" % (tag, i, j)).digest(), "big") % N or 1).to_bytes(32, "big")


proofs, rhos, cvs, cms = [], [], [], []
for i in range(count):
    nv = b"".join(sc(b"bp-n", i, j) for j in range(g))
    lv = b"".join(sc(b"bp-l", i, j) for j in range(h))
    cv = b"".join(sc(b"bp-c", i, j) for j in range(h))
    rho = sc(b"bp-rho", i)
    proof = ctypes.create_string_buffer(1000)
    plen = ctypes.c_size_t(1000)
    cm = ctypes.create_string_buffer(64)
    assert L.ref_bppp_norm_prove(ref.ctx, proof, ctypes.byref(plen), cm, rho, gs, nv, ctypes.c_size_t(g), lv, cv, ctypes.c_size_t(h)) == 1
    assert plen.value == 324
    assert L.ref_bppp_norm_verify(ref.ctx, proof.raw[:324], ctypes.c_size_t(324), cm.raw, rho, gs, ctypes.c_size_t(g), cv, ctypes.c_size_t(h)) == 1
    proofs.append(proof.raw[:324]); rhos.append(rho); cvs.append(cv); cms.append(cm.raw)
out = os.path.join(ROOT, "tests", "golden", "bppp_fixture.npz")
u8 = lambda b, *s: np.frombuffer(b, dtype=np.uint8).reshape(*s)
np.savez_compressed(out, proofs=u8(b"".join(proofs), count, 324), rho=u8(b"".join(rhos), count, 32), c_vec=u8(b"".join(cvs), count, 32 * h),
                    commits=u8(b"".join(cms), count, 64), gens=u8(gxy.raw, g + h, 64), g_len=np.array(g), h_len=np.array(h))
print(out, os.path.getsize(out))
