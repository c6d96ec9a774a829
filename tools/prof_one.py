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
