"""One launch of each hot kernel on synthetic inputs, for `ncu --set full -k regex:...` captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import secp256k1_zkp_b200 as b200  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 18
n = 1 << lg
eng = b200.Engine()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda"); g.manual_seed(1)
sk = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g); sk[:, 0] &= 0x7F; sk[:, 31] |= 1
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
sig = torch.empty((n, 64), dtype=torch.uint8, device="cuda"); pk = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
eng.testdata_schnorrsig_sign_dev(sig.data_ptr(), pk.data_ptr(), sk.data_ptr(), msg.data_ptr(), 32, n, st)
ok = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.schnorrsig_verify_dev(ok.data_ptr(), sig.data_ptr(), msg.data_ptr(), 32, pk.data_ptr(), b200.PK_XONLY32, n, st)
a = torch.randint(0, 256, (148 * 2048, 32), dtype=torch.uint8, device="cuda", generator=g)
o = torch.empty_like(a)
for op in (b200.FE_MUL, b200.FE_SQR):
    for _ in range(2):
        eng.fe_op_dev(op, o.data_ptr(), a.data_ptr(), msg.data_ptr(), a.shape[0], 1000, st)
torch.cuda.synchronize()
print("accepted", int(ok.sum()), "of", n)

# ---- secondary kernels: range proof (1024 proofs from the fixture), MSM 2^18, BP++ 512
import numpy as np
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
fx = np.load(os.path.join(gold, "rangeproof64_fixture.npz"))
per = 1024
proofs = torch.from_numpy(np.tile(fx["proofs"], (per // 64, 1))).cuda()
commits = torch.from_numpy(np.tile(fx["commits"], (per // 64, 1))).cuda()
plen = torch.full((per,), 5126, dtype=torch.int32, device="cuda")
g64 = torch.from_numpy(fx["gen64"].copy()).cuda()
okr = torch.zeros(per, dtype=torch.uint8, device="cuda"); mn = torch.zeros(per, dtype=torch.int64, device="cuda"); mx = torch.zeros(per, dtype=torch.int64, device="cuda")
for _ in range(2):
    eng.rangeproof_verify_dev(okr.data_ptr(), mn.data_ptr(), mx.data_ptr(), commits.data_ptr(), proofs.data_ptr(), 5126, plen.data_ptr(), g64.data_ptr(), 0, per, st)
m = 1 << 18
sc = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device="cuda", generator=g)
pts = torch.empty((m, 64), dtype=torch.uint8, device="cuda")
eng.ecmult_dev(pts.data_ptr(), torch.zeros((m, 64), dtype=torch.uint8, device="cuda").data_ptr(), torch.zeros((m, 32), dtype=torch.uint8, device="cuda").data_ptr(), sc.data_ptr(), m, st)
wsb = eng.ecmult_multi_workspace_bytes(m); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); out = torch.empty(64, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.ecmult_multi_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), m, ws.data_ptr(), wsb, st)
torch.cuda.synchronize()
print("rangeproof accepted", int(okr.sum()), "of", per)
