"""One launch of each secondary kernel at the bench.py sizes, for ncu metric captures (profiles/r1g_*): 8192 64-bit range proofs,
2^20-term MSM (full-range scalars), 2^18 ECDSA recoveries, one 2^18-signature half-aggregate (device chain skipped: z uploaded)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import secp256k1_zkp_b200 as b200  # noqa: E402

eng = b200.Engine()
st = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device="cuda"); g.manual_seed(1)
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
fx = np.load(os.path.join(gold, "rangeproof64_fixture.npz"))
per = 8192
proofs = torch.from_numpy(np.tile(fx["proofs"], (per // 64, 1))).cuda()
commits = torch.from_numpy(np.tile(fx["commits"], (per // 64, 1))).cuda()
plen = torch.full((per,), 5126, dtype=torch.int32, device="cuda")
g64 = torch.from_numpy(fx["gen64"].copy()).cuda()
okr = torch.zeros(per, dtype=torch.uint8, device="cuda"); mn = torch.zeros(per, dtype=torch.int64, device="cuda"); mx = torch.zeros(per, dtype=torch.int64, device="cuda")
for _ in range(2):
    eng.rangeproof_verify_dev(okr.data_ptr(), mn.data_ptr(), mx.data_ptr(), commits.data_ptr(), proofs.data_ptr(), 5126, plen.data_ptr(), g64.data_ptr(), 0, per, st)
m = 1 << 20
sc = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device="cuda", generator=g)
sk = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device="cuda", generator=g); sk[:, 0] &= 0x7F
pts = torch.empty((m, 64), dtype=torch.uint8, device="cuda")
eng.ecmult_dev(pts.data_ptr(), torch.zeros((m, 64), dtype=torch.uint8, device="cuda").data_ptr(), torch.zeros((m, 32), dtype=torch.uint8, device="cuda").data_ptr(), sk.data_ptr(), m, st)
wsb = eng.ecmult_multi_workspace_bytes(m); ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); out = torch.empty(64, dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.ecmult_multi_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), m, ws.data_ptr(), wsb, st)
n = 1 << 18
msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
sig = torch.empty((n, 64), dtype=torch.uint8, device="cuda"); pk = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
sk2 = sk[:n].clone(); sk2[:, 31] |= 1
eng.testdata_schnorrsig_sign_dev(sig.data_ptr(), pk.data_ptr(), sk2.data_ptr(), msg.data_ptr(), 32, n, st)
rsig = sig.clone(); rsig[:, 32] &= 0x7F
rid = torch.randint(0, 2, (n,), dtype=torch.uint8, device="cuda", generator=g)
ok = torch.empty(n, dtype=torch.uint8, device="cuda"); rec = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
for _ in range(2):
    eng.ecdsa_recover_dev(ok.data_ptr(), rec.data_ptr(), rsig.data_ptr(), rid.data_ptr(), msg.data_ptr(), n, st)
agg = torch.cat([sig[:, :32].reshape(-1), torch.zeros(32, dtype=torch.uint8, device="cuda")]).contiguous()
for _ in range(2):
    eng.schnorrsig_aggverify(pk.cpu().numpy(), msg.cpu().numpy(), agg.cpu().numpy())
torch.cuda.synchronize()
print("rangeproof accepted", int(okr.sum()), "of", per, "recovered", int(ok.sum()), "of", n)
