"""MSM timing sweep on one GPU (CUDA events, inputs resident in HBM):  python tools/msm_bench.py [--sweep]
One JSON line per (n, window): ms per MSM, terms/s.  --sweep also forces every allowed window width (tuning of msm_choose_c)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import secp256k1_zkp_b200 as b200
    sweep = "--sweep" in sys.argv
    dev = torch.device("cuda", 0)
    eng = b200.Engine(device=0)
    st = torch.cuda.current_stream().cuda_stream
    nmax = 1 << 20
    gen = torch.Generator(device=dev); gen.manual_seed(1)
    sk = torch.randint(0, 256, (nmax, 32), dtype=torch.uint8, device=dev, generator=gen); sk[:, 0] &= 0x7F
    sc = torch.randint(0, 256, (nmax, 32), dtype=torch.uint8, device=dev, generator=gen)
    z64, z32 = torch.zeros((nmax, 64), dtype=torch.uint8, device=dev), torch.zeros((nmax, 32), dtype=torch.uint8, device=dev)
    pts = torch.empty((nmax, 64), dtype=torch.uint8, device=dev)
    eng.ecmult_dev(pts.data_ptr(), z64.data_ptr(), z32.data_ptr(), sk.data_ptr(), nmax, st)
    wsb = eng.ecmult_multi_workspace_bytes(nmax)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.zeros(96, dtype=torch.uint8, device=dev)
    sizes = (8, 88, 1024, 4096, 1 << 14, 1 << 15, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20)
    if "--n" in sys.argv:                    # one size, one launch sequence after a warm-up (for an ncu launch list)
        n = int(sys.argv[sys.argv.index("--n") + 1])
        for _ in range(2):
            eng.ecmult_multi_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), n, ws.data_ptr(), wsb, st)
        torch.cuda.synchronize()
        eng.close()
        return
    for n in sizes:
        cs = [0] + (list(range(8 if n > 4095 else 4, 17)) if sweep else [])
        for c in cs:
            eng.ecmult_multi_set_window(c)
            reps = 20 if n <= (1 << 17) else 5

            def run():
                eng.ecmult_multi_dev(out.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), n, ws.data_ptr(), wsb, st)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(json.dumps({"n": n, "window": c, "ms": round(ms, 4), "terms_per_s": n / ms * 1e3}), flush=True)
    eng.ecmult_multi_set_window(0)
    eng.close()


if __name__ == "__main__":
    main()
