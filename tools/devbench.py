"""Developer micro-benchmarks on one B200 (run under gpurun): IMAD peak probe, fe_mul probe,
ecmult / schnorr / ecdsa batch kernels with inputs resident in HBM.  Prints JSON lines."""
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import secp256k1_zkp_b200 as b200  # noqa: E402


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return best


def main():
    torch.cuda.init()
    t0 = time.time()
    eng = b200.Engine()
    print(json.dumps({"engine_create_s": round(time.time() - t0, 3), "table_mib": eng.table_bytes / 2**20}))
    stream = torch.cuda.current_stream().cuda_stream
    # IMAD probe
    blocks, threads, iters = 148 * 8, 256, 4096
    out = torch.empty(blocks * threads, dtype=torch.int64, device="cuda")
    ms = timed(lambda: eng.imad_probe_dev(out.data_ptr(), blocks, threads, iters, stream))
    macs = blocks * threads * iters * 16
    print(json.dumps({"probe": "imad_wide", "ms": ms, "tmac_s": macs / ms / 1e9}))
    # fe_mul probe
    n = 148 * 2048
    a = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
    b = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
    o = torch.empty_like(a)
    for op, name in ((b200.FE_MUL, "fe_mul"), (b200.FE_SQR, "fe_sqr"), (b200.FE_MUL_PORTABLE, "fe_mul_portable"), (b200.FE_ADD, "fe_add"), (b200.FE_SUB, "fe_sub")):
        k = 2000
        ms = timed(lambda: eng.fe_op_dev(op, o.data_ptr(), a.data_ptr(), b.data_ptr(), n, k, stream))
        print(json.dumps({"probe": name, "ms": ms, "gops_s": n * k / ms / 1e6, "tmac124_s": n * k * 124 / ms / 1e9}))
    # ecmult / schnorr (random inputs: invalid signatures do the same work up to the final compare, like bench_rangeproof)
    for lg in (16, 18, 20):
        n = 1 << lg
        sig = torch.randint(0, 256, (n, 64), dtype=torch.uint8, device="cuda")
        sig[:, 0] = 0x7F
        sig[:, 32] = 0x7F
        msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
        # valid x-only pubkeys: take x of k*G computed on the device by ecmult
        na = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
        ng = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
        ng[:, 0] = 0x7F
        a64 = torch.zeros((n, 64), dtype=torch.uint8, device="cuda")
        pts = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
        ms = timed(lambda: eng.ecmult_dev(pts.data_ptr(), a64.data_ptr(), na.data_ptr(), ng.data_ptr(), n, stream), warm=1, reps=3)
        print(json.dumps({"bench": "ecmult_gen_only", "n": n, "ms": ms, "ops_s": n / ms * 1e3}))
        na2 = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
        na2[:, 0] = 0x7F
        out = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
        ms = timed(lambda: eng.ecmult_dev(out.data_ptr(), pts.data_ptr(), na2.data_ptr(), ng.data_ptr(), n, stream), warm=1, reps=3)
        print(json.dumps({"bench": "ecmult", "n": n, "ms": ms, "ops_s": n / ms * 1e3, "tmac_s": n * 1.81e5 / ms / 1e9}))
        pk = pts[:, :32].contiguous()
        ok = torch.empty(n, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: eng.schnorrsig_verify_dev(ok.data_ptr(), sig.data_ptr(), msg.data_ptr(), 32, pk.data_ptr(), b200.PK_XONLY32, n, stream), warm=1, reps=3)
        print(json.dumps({"bench": "schnorr_verify_xonly32", "n": n, "ms": ms, "ops_s": n / ms * 1e3, "accepted": int(ok.sum())}))
        pk33 = torch.cat([torch.full((n, 1), 2, dtype=torch.uint8, device="cuda"), pk], dim=1).contiguous()
        ms = timed(lambda: eng.ecdsa_verify_dev(ok.data_ptr(), sig.data_ptr(), b200.SIG_COMPACT64, msg.data_ptr(), pk33.data_ptr(), b200.PK_COMPRESSED33, n, stream), warm=1, reps=3)
        print(json.dumps({"bench": "ecdsa_verify_comp33", "n": n, "ms": ms, "ops_s": n / ms * 1e3, "accepted": int(ok.sum())}))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def bench_msm_and_rangeproof():
    import hashlib
    torch.cuda.init()
    eng = b200.Engine()
    stream = torch.cuda.current_stream().cuda_stream
    for lg in (16, 20):
        n = 1 << lg
        ng = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda"); ng[:, 0] = 0x7F
        pts = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
        eng.ecmult_dev(pts.data_ptr(), torch.zeros((n, 64), dtype=torch.uint8, device="cuda").data_ptr(),
                       torch.zeros((n, 32), dtype=torch.uint8, device="cuda").data_ptr(), ng.data_ptr(), n, stream)
        sc = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device="cuda")
        wsb = eng.ecmult_multi_workspace_bytes(n)
        ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
        out = torch.empty(64, dtype=torch.uint8, device="cuda")
        ms = timed(lambda: eng.ecmult_multi_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), n, ws.data_ptr(), wsb, stream), warm=1, reps=3)
        print(json.dumps({"bench": "ecmult_multi_pippenger", "n": n, "ms": ms, "terms_s": n / ms * 1e3, "workspace_mib": wsb / 2**20}))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "msm":
    bench_msm_and_rangeproof()
