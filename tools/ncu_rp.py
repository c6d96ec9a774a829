"""Workload for an ncu capture of k_rangeproof_verify / k_ecdsa_verify / k_bppp_norm_verify: one device-resident launch of each at bench sizes.

    ncu --set full --import-source on --clock-control none -k regex:'k_rangeproof_verify|k_ecdsa_verify|k_bppp_norm_verify' \
        -o gpurun_out/r2m_kernels python tools/ncu_rp.py

Inputs are the ones bench.py uses (8192 distinct reference-signed 64-bit proofs, 2^18 signatures, 512 distinct norm arguments)."""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import bench
    import secp256k1_zkp_b200 as b200
    from tests.refshim import RefLib
    ref = RefLib()
    hi = bench.host_info()
    eng = b200.Engine(0)
    dev = torch.device("cuda:0")
    per = int(os.environ.get("RP_N", "8192"))
    proofs_h, commits_h, _, _, genh = bench.make_rangeproofs(ref, per, b"ncu", hi)
    proofs = torch.from_numpy(proofs_h).to(dev); commits = torch.from_numpy(commits_h).to(dev)
    plen = torch.full((per,), 5126, dtype=torch.int32, device=dev)
    g64 = torch.frombuffer(bytearray(genh), dtype=torch.uint8).to(dev)
    okr = torch.zeros(per, dtype=torch.uint8, device=dev)
    mn = torch.zeros(per, dtype=torch.int64, device=dev); mx = torch.zeros(per, dtype=torch.int64, device=dev)
    for _ in range(2):
        eng.rangeproof_verify_dev(okr.data_ptr(), mn.data_ptr(), mx.data_ptr(), commits.data_ptr(), proofs.data_ptr(), 5126, plen.data_ptr(), g64.data_ptr(), 0, per, 0)
    torch.cuda.synchronize()
    assert bool(okr.all()), "range proofs rejected"
    if os.environ.get("RP_TIME"):                            # A/B of two builds (SECP256K1_B200_LIB): CUDA-event time of the launch, nothing else
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ts = []
        for _ in range(5):
            ev[0].record()
            eng.rangeproof_verify_dev(okr.data_ptr(), mn.data_ptr(), mx.data_ptr(), commits.data_ptr(), proofs.data_ptr(), 5126, plen.data_ptr(), g64.data_ptr(), 0, per, 0)
            ev[1].record(); torch.cuda.synchronize()
            ts.append(ev[0].elapsed_time(ev[1]))
        ev[0].record()                                       # back to back, as bench.py times it (no synchronisation between the calls)
        for _ in range(3):
            eng.rangeproof_verify_dev(okr.data_ptr(), mn.data_ptr(), mx.data_ptr(), commits.data_ptr(), proofs.data_ptr(), 5126, plen.data_ptr(), g64.data_ptr(), 0, per, 0)
        ev[1].record(); torch.cuda.synchronize()
        print("rangeproof_verify %d proofs: %s ms -> %.1f k/s; 3 calls back to back %.2f ms each (%s)" % (
            per, ["%.2f" % t for t in ts], per / min(ts), ev[0].elapsed_time(ev[1]) / 3, os.environ.get("SECP256K1_B200_LIB", "default")))
        return

    if os.environ.get("RP_ONLY"):
        return
    n = 1 << 18
    sk = np.frombuffer(b"".join(hashlib.sha256(b"ncu-sk%d" % i).digest() for i in range(4096)), dtype=np.uint8)
    msg = np.frombuffer(b"".join(hashlib.sha256(b"ncu-m%d" % i).digest() for i in range(4096)), dtype=np.uint8)
    sig, pk = ref.ecdsa_make(sk.tobytes(), msg.tobytes())
    sig = np.tile(np.frombuffer(bytes(sig), dtype=np.uint8), n // 4096); pk = np.tile(np.frombuffer(bytes(pk), dtype=np.uint8), n // 4096)
    msgs = np.tile(msg, n // 4096)
    dsig = torch.from_numpy(sig.copy()).to(dev); dpk = torch.from_numpy(pk.copy()).to(dev); dmsg = torch.from_numpy(msgs.copy()).to(dev)
    oke = torch.zeros(n, dtype=torch.uint8, device=dev)
    for _ in range(2):
        eng.ecdsa_verify_dev(oke.data_ptr(), dsig.data_ptr(), b200.SIG_COMPACT64, dmsg.data_ptr(), dpk.data_ptr(), b200.PK_COMPRESSED33, n, 0)
    torch.cuda.synchronize()
    assert bool(oke.all()), "ECDSA signatures rejected"

    nb = 512
    pr, rhos, cvs, cms, gens, _ = bench.make_bppp(ref, nb, b"ncu", hi)
    tr = eng.bppp_transcripts(cms, rhos, gens, 16, cvs, 8, mode="tests")
    for _ in range(2):
        okb = eng.bppp_norm_verify(pr, np.full(nb, 324, dtype=np.uint32), tr, rhos, gens, 16, cvs, 8, cms)
    assert int(okb.sum()) == nb, "norm arguments rejected"
    print("done")


if __name__ == "__main__":
    main()
