"""Generates tests/golden/rangeproof64_fixture.npz: 64 distinct valid 64-bit Back-Maxwell range proofs
(5126 bytes each, min_value 0, exp 0, min_bits 64, generator secp256k1_generator_h) with their commitments,
made by the UNMODIFIED reference (secp256k1_rangeproof_sign through oracle/_ref), following SURVEY 8d config 4:
value_i = SHA256("rp-v"||LE32(i))[0:8], blind_i = SHA256("rp-b"||LE32(i)) mod n, nonce = commitment bytes
(src/bench_rangeproof.c:32-34).  bench.py tiles these to the batch size (signing 65536 proofs is ~5 CPU-minutes and the
GPU box must not depend on the oracle for its inputs)."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.refshim import RefLib  # noqa: E402

N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
ref = RefLib()
gen = ref.generator_h()
count = 64
vals = [int.from_bytes(hashlib.sha256(b"rp-v" + i.to_bytes(4, "little")).digest()[:8], "little") for i in range(count)]
blinds = b"".join((int.from_bytes(hashlib.sha256(b"rp-b" + i.to_bytes(4, "little")).digest(), "big") % N or 1).to_bytes(32, "big") for i in range(count))
proofs, plen, commits = ref.rangeproof_make(blinds, vals, gen, min_value=0, exp=0, min_bits=64, stride=5126)
assert (plen == 5126).all()
ok, mn, mx = ref.rangeproof_verify(commits.tobytes(), proofs, plen, gen)
assert ok.all() and (mn == 0).all() and (mx == 2**64 - 1).all()
out = os.path.join(ROOT, "tests", "golden", "rangeproof64_fixture.npz")
np.savez_compressed(out, proofs=proofs, commits=commits, values=np.array(vals, dtype=np.uint64), blinds=np.frombuffer(blinds, dtype=np.uint8).reshape(count, 32),
                    gen64=np.frombuffer(gen, dtype=np.uint8))
print(out, os.path.getsize(out))
