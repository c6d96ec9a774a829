"""Developer benchmarks for the SURVEY 8(f) rank-3 additions on one B200 (run under gpurun): ECDSA public-key recovery and
BIP-340 half-aggregation verification.  Prints JSON lines."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import secp256k1_zkp_b200 as b200  # noqa: E402
from tools.devbench import timed  # noqa: E402


def main():
    eng = b200.Engine()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    for lg in (16, 20):
        n = 1 << lg
        sk = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g); sk[:, 0] &= 0x7F; sk[:, 31] |= 1
        msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda", generator=g)
        sig = torch.empty((n, 64), dtype=torch.uint8, device="cuda"); pk = torch.empty((n, 32), dtype=torch.uint8, device="cuda")
        eng.testdata_schnorrsig_sign_dev(sig.data_ptr(), pk.data_ptr(), sk.data_ptr(), msg.data_ptr(), 32, n, st)
        # recovery: r = a valid curve x (the nonce point of the BIP-340 signature), s random, recid in {0,1}: every item does the full work
        rsig = sig.clone(); rsig[:, 32] &= 0x7F
        rid = torch.randint(0, 2, (n,), dtype=torch.uint8, device="cuda", generator=g)
        ok = torch.empty(n, dtype=torch.uint8, device="cuda"); out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
        ms = timed(lambda: eng.ecdsa_recover_dev(ok.data_ptr(), out.data_ptr(), rsig.data_ptr(), rid.data_ptr(), msg.data_ptr(), n, st), warm=1, reps=3)
        print(json.dumps({"bench": "ecdsa_recover", "n": n, "ms": ms, "ops_s": n / ms * 1e3, "recovered": int(ok.sum())}))
        # half-aggregation: r_i of real signatures, s arbitrary (the verdict is "reject", the work is the same)
        agg = torch.cat([sig[:, :32].reshape(-1), torch.zeros(32, dtype=torch.uint8, device="cuda")]).contiguous()
        wsb = eng.schnorrsig_aggverify_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda"); okb = torch.zeros(1, dtype=torch.uint8, device="cuda")
        if lg <= 16:
            ms = timed(lambda: eng.schnorrsig_aggverify_dev(okb.data_ptr(), pk.data_ptr(), msg.data_ptr(), n, agg.data_ptr(), ws.data_ptr(), wsb, st), warm=1, reps=2)
            print(json.dumps({"bench": "halfagg_verify_dev (device randomizer chain, 1 thread)", "n": n, "ms": ms, "sigs_s": n / ms * 1e3}))
        h_pk, h_msg, h_agg = pk.cpu().numpy(), msg.cpu().numpy(), agg.cpu().numpy()
        eng.schnorrsig_aggverify(h_pk, h_msg, h_agg)
        t0 = time.perf_counter()
        for _ in range(3):
            eng.schnorrsig_aggverify(h_pk, h_msg, h_agg)
        ms = (time.perf_counter() - t0) / 3 * 1e3
        print(json.dumps({"bench": "halfagg_verify (host buffers: host randomizer chain + H2D + terms + MSM)", "n": n, "ms": ms, "sigs_s": n / ms * 1e3}))


def rewind_bench():
    """8192 64-bit proofs of the fixture (nonce = commitment bytes): verify + replay + value/blind recovery, host buffers"""
    eng = b200.Engine()
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    fx = np.load(os.path.join(gold, "rangeproof64_fixture.npz"))
    reps = 128
    proofs = np.tile(fx["proofs"], (reps, 1)); commits = np.tile(fx["commits"], (reps, 1)); n = proofs.shape[0]
    nonces = commits[:, :32].copy()
    plen = np.full(n, 5126)
    for cap in (0, 4096):
        eng.rangeproof_rewind(nonces.tobytes(), commits.tobytes(), proofs, plen, fx["gen64"], msg_cap=cap)
        t0 = time.perf_counter()
        ok, bl, val, msg, ml, mn, mx = eng.rangeproof_rewind(nonces.tobytes(), commits.tobytes(), proofs, plen, fx["gen64"], msg_cap=cap)
        ms = (time.perf_counter() - t0) * 1e3
        good = bool(ok.all()) and bool(np.array_equal(val, np.tile(fx["values"], reps)))
        print(json.dumps({"bench": "rangeproof_rewind 64-bit (host buffers, message capacity %d)" % cap, "n": n, "ms": ms, "proofs_s": n / ms * 1e3, "values_ok": good}))


if __name__ == "__main__":
    rewind_bench()
    main()
