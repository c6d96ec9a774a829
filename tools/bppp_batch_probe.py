"""BP++ norm verification: ms per call at several batch sizes (CUDA events, distinct reference-made arguments tiled beyond 8192).
SECP256K1_B200_BPPP_PACK_MIN=<n> moves the point where the packed kernel (4 proofs per warp, generator comb tables) takes over."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import numpy as np
    import torch
    import bench
    import secp256k1_zkp_b200 as b200
    from tests.refshim import RefLib
    ref = RefLib(); hi = bench.host_info(); eng = b200.Engine(0); dev = torch.device("cuda:0")
    sizes = [int(x) for x in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]
    base = bench.make_bppp(ref, min(max(sizes), 8192), b"probe", hi)
    for big in sizes:
        bp, rho, cv, cm = (np.tile(x, ((big + x.shape[0] - 1) // x.shape[0], 1))[:big] for x in base[:4])
        gens = base[4]
        tr = eng.bppp_transcripts(cm, rho, gens, 16, cv, 8, mode="tests")
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in (("p", bp), ("rho", rho), ("cv", cv), ("cm", cm), ("tr", tr), ("g", gens))}
        pl = torch.full((big,), 324, dtype=torch.int32, device=dev); ok = torch.zeros(big, dtype=torch.uint8, device=dev)

        def f():
            eng._ck(eng._lib.secp256k1_b200_bppp_norm_verify_batch_dev(eng._h, ok.data_ptr(), d["p"].data_ptr(), 324, pl.data_ptr(), d["tr"].data_ptr(),
                                                                       d["rho"].data_ptr(), d["g"].data_ptr(), 16, d["cv"].data_ptr(), 8, d["cm"].data_ptr(), big, 0))
        f(); f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("%6d proofs: %.3f ms -> %.0f k/s, all accepted %s (pack_min %s)" % (big, ms, big / ms, bool(ok.all()), os.environ.get("SECP256K1_B200_BPPP_PACK_MIN", "default")))


if __name__ == "__main__":
    main()
