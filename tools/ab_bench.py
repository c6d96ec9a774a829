"""A/B two builds of the library on one GPU: correctness spot-check + kernel timings.  usage: ab_bench.py libA.so libB.so ..."""
import os, subprocess, sys
code = r'''
import os, sys, json, hashlib
import numpy as np, torch
sys.path.insert(0, os.getcwd())
import secp256k1_zkp_b200 as b200
from tests.refshim import RefLib
eng = b200.Engine(); ref = RefLib()
n = 4096
sk = b"".join(hashlib.sha256(b"ab%d" % i).digest() for i in range(n)); msg = b"".join(hashlib.sha256(b"abm%d" % i).digest() for i in range(n))
sig, pk = eng.testdata_schnorrsig_sign(sk, msg); rs, rp = ref.schnorr_make(sk, msg)
good = bool(np.array_equal(rs, sig)) and bool(np.array_equal(rp, pk))
sig = sig.copy(); sig[64 * 5 + 3] ^= 1
ok = eng.schnorrsig_verify(sig, msg, pk, b200.PK_XONLY32)
good = good and int(ok.sum()) == n - 1 and int(ok[5]) == 0
P = 2**256 - 2**32 - 977
a = np.random.default_rng(1).integers(0, 256, size=(2048, 32), dtype=np.uint8); b = np.random.default_rng(2).integers(0, 256, size=(2048, 32), dtype=np.uint8)
a[0] = 255; b[0] = 255; a[1] = 255
r = eng.fe_op(b200.FE_MUL, a, b)
good = good and all(int.from_bytes(bytes(x), "big") * int.from_bytes(bytes(y), "big") % P == int.from_bytes(bytes(z), "big") for x, y, z in zip(a, b, r))
st = torch.cuda.current_stream().cuda_stream
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); s.record(); fn(); e.record(); torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    return best
m = 148 * 2048
A = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device="cuda"); B = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device="cuda"); O = torch.empty_like(A)
res = {"ok": good}
for op, name in ((b200.FE_MUL, "fe_mul_G/s"), (b200.FE_SQR, "fe_sqr_G/s")):
    ms = timed(lambda: eng.fe_op_dev(op, O.data_ptr(), A.data_ptr(), B.data_ptr(), m, 2000, st)); res[name] = round(m * 2000 / ms / 1e6, 1)
N = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(3)
SK = torch.randint(0, 256, (N, 32), dtype=torch.uint8, device="cuda", generator=g); SK[:, 0] &= 0x7F; SK[:, 31] |= 1
M = torch.randint(0, 256, (N, 32), dtype=torch.uint8, device="cuda", generator=g)
S = torch.empty((N, 64), dtype=torch.uint8, device="cuda"); PK = torch.empty((N, 32), dtype=torch.uint8, device="cuda"); OK = torch.empty(N, dtype=torch.uint8, device="cuda")
eng.testdata_schnorrsig_sign_dev(S.data_ptr(), PK.data_ptr(), SK.data_ptr(), M.data_ptr(), 32, N, st)
ms = timed(lambda: eng.schnorrsig_verify_dev(OK.data_ptr(), S.data_ptr(), M.data_ptr(), 32, PK.data_ptr(), b200.PK_XONLY32, N, st)); res["schnorr_M/s"] = round(N / ms / 1e3, 2); res["all_valid"] = bool(OK.all())
print(json.dumps(res))
'''
for lib in sys.argv[1:]:
    env = dict(os.environ, SECP256K1_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(os.path.basename(lib), out.stdout.strip().split("\n")[-1] if out.stdout.strip() else out.stderr[-400:])
