"""bench.py -- headline benchmark: batch 2^20 independent BIP-340 schnorrsig_verify per GPU
(BASELINE.json configs[1]; one verify = one a*P + b*G ecmult + x-only pubkey parse, exactly what the
reference's `bench schnorrsig_verify` times, src/modules/schnorrsig/bench_impl.h:39-48).

  python bench.py --gpus N --steps K --warmup W            our arm (torchrun launches N ranks for N > 1)
  python bench.py --impl reference --gpus N ...            the reference's CPU implementation (oracle/_ref)

A step = one pass of the hot path over the rank's batch.  `value` = whole-job verifies/s with inputs
resident in HBM (CUDA events, max over ranks); `e2e` = the same through the host-buffer C-ABI call
(pinned host inputs, H2D and D2H inside the timed region).  Weak scaling: every rank verifies its own
2^20 batch; the only collectives are the init-time NCCL broadcast of the generator table and the
per-step all-gather of the packed accept bitmaps.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAC_PER_ECMULT = 1.81e5      # reference-algorithm 32x32->64 MACs per a*P + b*G
BYTES_PER_VERIFY = 64 + 32 + 32 + 1


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=20, help="verifies per GPU per step = 2^log2n")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (rangeproof / MSM / BP++)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.15)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) >= 9]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.rows[0][2]), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference (CPU) legs: oracle/_ref = the unmodified reference compiled from its own sources
def _make_cpu_workload(ref, n):
    import hashlib
    import numpy as np
    sk = b"".join(hashlib.sha256(b"bench-sk" + i.to_bytes(4, "little")).digest() for i in range(n))
    msg = b"".join(hashlib.sha256(b"bench-msg" + i.to_bytes(4, "little")).digest() for i in range(n))
    sig, pk = ref.schnorr_make(sk, msg)
    for i in range(7, n, 16):
        bit = i % 512
        sig[64 * i + bit // 8] ^= 1 << (bit % 8)
    return sig, np.frombuffer(msg, dtype=np.uint8).copy(), pk


def _cpu_verify_threads(ref, sig, msg, pk, n, threads):
    """verifies items [0, n) split over `threads` host threads (ctypes releases the GIL); returns seconds"""
    import numpy as np
    per = (n + threads - 1) // threads
    outs = [None] * threads

    def work(t):
        lo, hi = t * per, min(n, (t + 1) * per)
        if lo < hi:
            outs[t] = ref.schnorr_verify(sig[64 * lo:64 * hi], msg[32 * lo:32 * hi], pk[32 * lo:32 * hi])
    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    ok = np.concatenate([o for o in outs if o is not None])
    return dt, ok


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from tests.refshim import RefLib
    ref = RefLib()
    cores = cpu_cores()
    unique = 4096                                  # distinct signed items, tiled to the sample size
    sig, msg, pk = _make_cpu_workload(ref, unique)
    per_thread = 1024
    sample = per_thread * cores
    import numpy as np
    reps = (sample + unique - 1) // unique
    sigs, msgs, pks = np.tile(sig, reps)[:64 * sample], np.tile(msg, reps)[:32 * sample], np.tile(pk, reps)[:32 * sample]
    for _ in range(args.warmup):
        _cpu_verify_threads(ref, sigs, msgs, pks, sample, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, ok = _cpu_verify_threads(ref, sigs, msgs, pks, sample, cores)
        t += dt
    ms = t / args.steps * 1e3
    value = sample / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "schnorrsig_verify/s (= ecmult/s)", "value": value, "unit": "verify/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (5x52 limbs, x86-64 int128)", "data": "synthetic",
        "config": {"workload": "batch 2^%d independent BIP-340 schnorrsig_verify incl. xonly_pubkey_parse (random keys/msgs, 1/16 corrupted)" % args.log2n,
                   "step": "bounded sample: %d verifies over %d host threads" % (sample, cores)},
        "cpu_baseline": {"value": value, "unit": "verify/s", "cores": cores, "kind": "reference",
                         "sample": "%d verifies per step (%d per thread), reference secp256k1_schnorrsig_verify + xonly_pubkey_parse from oracle/_ref" % (sample, per_thread)},
        "e2e": {"value": value, "unit": "verify/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "accepted_fraction": float(ok.mean()),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# secondary workloads of BASELINE.json (configs 2..4): MSM, 64-bit range proofs, BP++ norm argument.  Inputs: GPU-generated
# points / committed fixtures made by the reference prover (tools/make_*_fixture.py), tiled to the batch size.
def _time_dev(torch, dist, world, fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def run_extra_workloads(eng, b200_mod, torch, dist, D, dev, stream, rank, world, imad_peak):
    import numpy as np
    out = {}
    gold = os.path.join(ROOT, "tests", "golden")
    # ---- config 2: Pippenger 2^20-term MSM, strong scaling: rank r owns terms [r*N/W, (r+1)*N/W), partial sums all-gathered
    n_total = 1 << 20
    lo, hi = D.shard_range(n_total, rank, world)
    m = hi - lo
    gen = torch.Generator(device=dev); gen.manual_seed(777 + rank)
    sk = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device=dev, generator=gen); sk[:, 0] &= 0x7F
    sc = torch.randint(0, 256, (m, 32), dtype=torch.uint8, device=dev, generator=gen)
    pts = torch.empty((m, 64), dtype=torch.uint8, device=dev)
    z64, z32 = torch.zeros((m, 64), dtype=torch.uint8, device=dev), torch.zeros((m, 32), dtype=torch.uint8, device=dev)
    eng.ecmult_dev(pts.data_ptr(), z64.data_ptr(), z32.data_ptr(), sk.data_ptr(), m, stream)     # P_i = sk_i * G
    wsb = eng.ecmult_multi_workspace_bytes(max(m, 64))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    part = torch.empty(96 if world > 1 else 64, dtype=torch.uint8, device=dev)
    final = torch.empty(64, dtype=torch.uint8, device=dev)
    allp = torch.empty((world, 96), dtype=torch.uint8, device=dev)

    def msm_step():
        if world == 1:
            eng.ecmult_multi_dev(part.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), m, ws.data_ptr(), wsb, stream)
        else:   # shard sum stays Jacobian (no inversion), ONE all-gather of 96 bytes per rank, W-1 additions + one inversion in one block
            eng.ecmult_multi_partial_dev(part.data_ptr(), None, sc.data_ptr(), pts.data_ptr(), m, ws.data_ptr(), wsb, stream)
            D.allgather_partials(part, allp)
            eng.ecmult_multi_combine_dev(final.data_ptr(), allp.data_ptr(), world, stream)
    ms = _time_dev(torch, dist, world, msm_step)
    # identity check of bench_ecmult's teardown (src/bench_ecmult.c:73-93): MSM == (sum s_i*sk_i) * G, scalars on the host
    res = (final if world > 1 else part).cpu().numpy().tobytes()
    N_ORD = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
    skh, sch = sk.cpu().numpy(), sc.cpu().numpy()
    tot = sum(int.from_bytes(a.tobytes(), "big") * int.from_bytes(b.tobytes(), "big") for a, b in zip(skh, sch)) % N_ORD
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, tot)
        tot = sum(parts) % N_ORD
    exp = eng.ecmult(bytes(64), bytes(32), tot.to_bytes(32, "big"))[0].tobytes()
    out["msm_pippenger"] = {"terms": n_total, "ms": ms, "terms_per_s": n_total / ms * 1e3, "scaling": "strong", "identity_check_ok": res == exp,
                            "roofline_frac_ref_mac": n_total / world * 2.5e4 / (ms / 1e3) / 1e12 / imad_peak if world == 1 else None}
    del ws, pts, sk, sc, z64, z32
    # ---- config 3: 64-bit range proofs, 8192 per GPU (65536 on 8 GPUs), reference-made fixture tiled, 1/32 corrupted
    fx = np.load(os.path.join(gold, "rangeproof64_fixture.npz"))
    per = 8192
    reps = per // fx["proofs"].shape[0]
    proofs = torch.from_numpy(np.tile(fx["proofs"], (reps, 1))).to(dev)
    commits = torch.from_numpy(np.tile(fx["commits"], (reps, 1))).to(dev)
    bad = torch.arange(3, per, 32, device=dev)
    proofs[bad, (bad * 7919) % 5126] ^= 1
    plen = torch.full((per,), 5126, dtype=torch.int32, device=dev)
    g64 = torch.from_numpy(fx["gen64"].copy()).to(dev)
    okr = torch.zeros(per, dtype=torch.uint8, device=dev)
    mn = torch.zeros(per, dtype=torch.int64, device=dev); mx = torch.zeros(per, dtype=torch.int64, device=dev)

    def rp_step():
        eng.rangeproof_verify_dev(okr.data_ptr(), mn.data_ptr(), mx.data_ptr(), commits.data_ptr(), proofs.data_ptr(), 5126, plen.data_ptr(),
                                  g64.data_ptr(), 0, per, stream)
        if world > 1:
            D.allgather_bitmaps(D.pack_bitmap(okr))
    ms = _time_dev(torch, dist, world, rp_step, warm=1, reps=3)
    expok = torch.ones(per, dtype=torch.uint8, device=dev); expok[bad] = 0
    out["rangeproof_verify_64bit"] = {"proofs": per * world, "per_gpu": per, "ms": ms, "verify_per_s": per * world / ms * 1e3, "scaling": "weak",
                                      "accept_pattern_ok": bool(torch.equal(okr, expok)), "max_value_ok": bool((mx == -1).all()),
                                      "roofline_frac_ref_mac": per * 2.7e7 / (ms / 1e3) / 1e12 / imad_peak,
                                      "data": "64 distinct reference-signed proofs (tests/golden/rangeproof64_fixture.npz) tiled"}
    del proofs, commits
    # ---- config 4: BP++ norm-product argument, 16 G + 8 H generators, 512 per GPU (4096 on 8 GPUs)
    fb = np.load(os.path.join(gold, "bppp_fixture.npz"))
    per = 512
    reps = per // fb["proofs"].shape[0]
    bp = np.tile(fb["proofs"], (reps, 1)); rho = np.tile(fb["rho"], (reps, 1)); cv = np.tile(fb["c_vec"], (reps, 1)); cm = np.tile(fb["commits"], (reps, 1))
    badi = np.arange(5, per, 16)
    bp[badi, 323] ^= 1
    tr = eng.bppp_transcripts(cm, rho, fb["gens"], 16, cv, 8, mode="tests")
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in (("p", bp), ("rho", rho), ("cv", cv), ("cm", cm), ("tr", tr), ("g", fb["gens"]))}
    pl = torch.full((per,), 324, dtype=torch.int32, device=dev)
    okb = torch.zeros(per, dtype=torch.uint8, device=dev)

    def bp_step():
        eng._ck(eng._lib.secp256k1_b200_bppp_norm_verify_batch_dev(eng._h, okb.data_ptr(), d["p"].data_ptr(), 324, pl.data_ptr(), d["tr"].data_ptr(),
                                                                   d["rho"].data_ptr(), d["g"].data_ptr(), 16, d["cv"].data_ptr(), 8, d["cm"].data_ptr(), per, stream))
        if world > 1:
            D.allgather_bitmaps(D.pack_bitmap(okb))
    ms = _time_dev(torch, dist, world, bp_step, warm=1, reps=3)
    expb = np.ones(per, dtype=np.uint8); expb[badi] = 0
    out["bppp_norm_verify_16g8h"] = {"proofs": per * world, "per_gpu": per, "ms": ms, "verify_per_s": per * world / ms * 1e3, "scaling": "weak",
                                     "accept_pattern_ok": bool(np.array_equal(okb.cpu().numpy(), expb)),
                                     "note": "norm-product argument only: this reference snapshot has no full BP++ range proof (SURVEY fact 3)"}
    del d
    # ---- ECDSA public-key recovery, 2^18 per GPU: r = x of a BIP-340 nonce point (always liftable), s random
    per = 1 << 18
    gen = torch.Generator(device=dev); gen.manual_seed(4242 + rank)
    sk = torch.randint(0, 256, (per, 32), dtype=torch.uint8, device=dev, generator=gen); sk[:, 0] &= 0x7F; sk[:, 31] |= 1
    msg = torch.randint(0, 256, (per, 32), dtype=torch.uint8, device=dev, generator=gen)
    sig = torch.empty((per, 64), dtype=torch.uint8, device=dev); pk = torch.empty((per, 32), dtype=torch.uint8, device=dev)
    eng.testdata_schnorrsig_sign_dev(sig.data_ptr(), pk.data_ptr(), sk.data_ptr(), msg.data_ptr(), 32, per, stream)
    rsig = sig.clone(); rsig[:, 32] &= 0x7F
    rid = torch.randint(0, 2, (per,), dtype=torch.uint8, device=dev, generator=gen)
    okc = torch.empty(per, dtype=torch.uint8, device=dev); rec = torch.empty((per, 64), dtype=torch.uint8, device=dev)

    def rec_step():
        eng.ecdsa_recover_dev(okc.data_ptr(), rec.data_ptr(), rsig.data_ptr(), rid.data_ptr(), msg.data_ptr(), per, stream)
    ms = _time_dev(torch, dist, world, rec_step, warm=1, reps=3)
    # property check: a recovered key verifies the signature it was recovered from (low-s form, as secp256k1_ecdsa_verify demands)
    chk = 4096
    s_host = rsig[:chk].cpu().numpy().copy()
    N_ORD = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
    for j in range(chk):
        sv = int.from_bytes(s_host[j, 32:].tobytes(), "big")
        if sv > N_ORD // 2:
            s_host[j, 32:] = np.frombuffer((N_ORD - sv).to_bytes(32, "big"), dtype=np.uint8)
    rk = rec[:chk].cpu().numpy()
    comp = np.concatenate([(2 + (rk[:, 63] & 1))[:, None].astype(np.uint8), rk[:, :32]], axis=1)
    vok = eng.ecdsa_verify(s_host, msg[:chk].cpu().numpy(), comp, b200_mod.SIG_COMPACT64, b200_mod.PK_COMPRESSED33)
    out["ecdsa_recover"] = {"signatures": per * world, "per_gpu": per, "ms": ms, "recover_per_s": per * world / ms * 1e3, "scaling": "weak",
                            "all_recovered": bool(okc.all()), "recovered_keys_verify": bool(vok.all())}
    # ---- one half-aggregate signature over the same 2^18 (key, message, r) triples per GPU, device part only
    # (terms + 2n+1-term MSM; the serial randomizer hash is host work and is reported by tools/devbench2.py)
    agg = torch.cat([sig[:, :32].reshape(-1), torch.zeros(32, dtype=torch.uint8, device=dev)]).contiguous()
    wsb = eng.schnorrsig_aggverify_workspace_bytes(per)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev); okb1 = torch.zeros(1, dtype=torch.uint8, device=dev)
    h_ok = eng.schnorrsig_aggverify(pk.cpu().numpy(), msg.cpu().numpy(), agg.cpu().numpy())       # s = 0: must be rejected
    t0 = time.perf_counter()
    eng.schnorrsig_aggverify(pk.cpu().numpy(), msg.cpu().numpy(), agg.cpu().numpy())
    out["schnorrsig_aggverify"] = {"signatures_per_aggregate": per, "host_call_ms": (time.perf_counter() - t0) * 1e3, "rejects_wrong_s": h_ok == 0,
                                   "note": "host-buffer call incl. the serial randomizer hash chain on the host; each rank verifies its own aggregate"}
    return out


# ------------------------------------------------------------------------------------------------
# measured with ncu on the 2^20 launch: DRAM bytes per launch (profiles/r1h_*, tables + spills in local memory), IMAD.WIDE.U32(.X) executed per verify (profiles/r1e_*)
NCU_DRAM_BYTES_2P20 = 1615576832 + 1516754176
IMAD_WIDE_EXECUTED_PER_VERIFY = 136155


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    import secp256k1_zkp_b200 as b200

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with %d ranks" % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- engine + generator table: rank 0 builds, ONE NCCL broadcast replicates it ---------------
    from secp256k1_zkp_b200.host import dist as D
    eng = b200.Engine(device=local_rank, build_tables=(rank == 0))
    D.replicate_engine_table(eng, rank, dev, stream)

    # ---- synthetic workload, generated on the GPU -----------------------------------------------------
    n = 1 << args.log2n
    gen = torch.Generator(device=dev)
    gen.manual_seed(20260923 + rank)
    sk = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen)
    sk[:, 0] &= 0x7F
    sk[:, 31] |= 1
    msg = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen)
    sig = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    pk = torch.empty((n, 32), dtype=torch.uint8, device=dev)
    eng.testdata_schnorrsig_sign_dev(sig.data_ptr(), pk.data_ptr(), sk.data_ptr(), msg.data_ptr(), 32, n, stream)
    bad = torch.arange(7, n, 16, device=dev)
    bit = bad % 512
    sig[bad, bit // 8] ^= (1 << (bit % 8)).to(torch.uint8)
    expected = torch.ones(n, dtype=torch.uint8, device=dev)
    expected[bad] = 0
    ok = torch.zeros(n, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * n // 8, dtype=torch.uint8, device=dev) if world > 1 else None
    torch.cuda.synchronize()

    kernel_ms = []

    def step(record=False):
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        eng.schnorrsig_verify_dev(ok.data_ptr(), sig.data_ptr(), msg.data_ptr(), 32, pk.data_ptr(), b200.PK_XONLY32, n, stream)
        if record:
            e1.record()
            kernel_ms.append((e0, e1))
        if world > 1:
            D.allgather_bitmaps(D.pack_bitmap(ok), gathered)

    # ---- integer-pipe peak, measured live (register-resident IMAD.WIDE.U32 saturation) ------------------
    pblocks, pthreads, piters = 148 * 8, 256, 4096
    pout = torch.empty(pblocks * pthreads, dtype=torch.int64, device=dev)
    for _ in range(2):
        eng.imad_probe_dev(pout.data_ptr(), pblocks, pthreads, piters, stream)
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    eng.imad_probe_dev(pout.data_ptr(), pblocks, pthreads, piters, stream)
    pe1.record()
    torch.cuda.synchronize()
    imad_peak_tmacs = pblocks * pthreads * piters * 16 / pe0.elapsed_time(pe1) / 1e9

    # ---- timed region: `value` (inputs resident in HBM) ---------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = eng.launch_count
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(record=True)
    t1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count - launches0
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1) / args.steps
    kms = sum(a.elapsed_time(b) for a, b in kernel_ms) / len(kernel_ms)
    tms = torch.tensor([ms, kms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms, kms = float(tms[0]), float(tms[1])
    full_ok = bool(torch.equal(ok, expected))

    # ---- e2e: the host-buffer C-ABI call, pinned host inputs, H2D + D2H inside the timed region -------------
    h_sig, h_msg, h_pk = sig.cpu().pin_memory(), msg.cpu().pin_memory(), pk.cpu().pin_memory()
    h_ok = None

    def e2e_step():
        return eng.schnorrsig_verify(h_sig.numpy(), h_msg.numpy(), h_pk.numpy(), b200.PK_XONLY32)
    for _ in range(2):
        h_ok = e2e_step()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 10))
    tw = time.perf_counter()
    for _ in range(e2e_steps):
        h_ok = e2e_step()
    e2e_ms = (time.perf_counter() - tw) / e2e_steps * 1e3
    te = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms = float(te[0])
    e2e_ok = bool(np.array_equal(h_ok, expected.cpu().numpy()))

    # ---- parity sample against the oracle + CPU baseline (rank 0) -------------------------------------------
    parity = {"full_batch_matches_construction": full_ok and e2e_ok}
    cpu_baseline = None
    if rank == 0:
        try:
            from tests.refshim import RefLib
            ref = RefLib()
            m = 4096
            idx = slice(0, m)
            exp = ref.schnorr_verify(h_sig.numpy()[idx].reshape(-1), h_msg.numpy()[idx].reshape(-1), h_pk.numpy()[idx].reshape(-1))
            parity.update({"oracle": "oracle/_ref (reference)", "oracle_checked": m, "oracle_mismatches": int((exp != h_ok[:m]).sum())})
            if world == 1 and not args.no_cpu_baseline:
                cores = cpu_cores()
                per_thread = 8192 if cores <= 16 else 4096
                sample = min(n, per_thread * cores)
                dt, okc = _cpu_verify_threads(ref, h_sig.numpy().reshape(-1), h_msg.numpy().reshape(-1), h_pk.numpy().reshape(-1), sample, cores)
                cpu_baseline = {"value": sample / dt, "unit": "verify/s", "cores": cores, "kind": "reference",
                                "sample": "first %d items of the same batch, %d host threads, reference xonly_pubkey_parse + schnorrsig_verify (oracle/_ref), %.1f s wall" % (sample, cores, dt),
                                "mismatches_vs_gpu": int((okc != h_ok[:sample]).sum())}
        except OSError as e:
            parity.update({"oracle": "unavailable: %s" % e})

    extra = None
    if not args.no_extra:
        try:
            extra = run_extra_workloads(eng, b200, torch, dist, D, dev, stream, rank, world, imad_peak_tmacs)
        except Exception as ex:  # the headline line must survive a failing secondary workload
            extra = {"error": repr(ex)}

    if rank == 0:
        total = world * n
        value = total / (ms / 1e3)
        achieved = n * MAC_PER_ECMULT / (kms / 1e3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "schnorrsig_verify/s (= ecmult/s)", "value": value, "unit": "verify/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8x32-bit limbs, IMAD.WIDE.U32)", "data": "synthetic",
            "config": {"workload": "batch 2^%d independent BIP-340 schnorrsig_verify per GPU incl. x-only pubkey parse (random keys/msgs, 1/16 corrupted), %d GPU(s)" % (args.log2n, world),
                       "per_gpu_batch": n, "global_batch": total, "l2": "inputs %.0f MiB per pass > 126 MB L2" % (n * 128 / 2**20),
                       "parallelism": "independent shards, table broadcast at init, bitmap all-gather per step"},
            "clocks": clocks,
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "verify/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 128 * n, "d2h_bytes_per_step": n,
                    "api": "secp256k1_b200_schnorrsig_verify_batch (host buffers, pinned)"},
            "gpu_launches": launches,
            "roofline": {"bound": "int32 multiply pipe (IMAD.WIDE.U32); not hbm, not tensor", "achieved": achieved, "peak": imad_peak_tmacs, "unit": "TMAC/s",
                         "frac": achieved / imad_peak_tmacs, "traffic": NCU_DRAM_BYTES_2P20 if args.log2n == 20 else None,
                         "how": "algorithmic 1.81e5 MAC per verify (SURVEY 8d, reference algorithm) x %d verifies / %.3f ms kernel (CUDA events); peak = live IMAD.WIDE.U32 probe "
                                "(16 independent accumulators, distinct multiplicands); traffic = dram read+write bytes of one 2^20 launch (ncu, profiles/r1h_ncu_schnorr_2p20_metrics.csv): "
                                "inputs 135 MB + generator-table misses + local-memory tables/spills written back, 129 GB/s = 2 %% of HBM, not a bound" % (n, kms),
                         "executed": {"imad_wide_per_verify": IMAD_WIDE_EXECUTED_PER_VERIFY, "tmac_s": n * IMAD_WIDE_EXECUTED_PER_VERIFY / (kms / 1e3) / 1e12,
                                      "frac_of_peak": n * IMAD_WIDE_EXECUTED_PER_VERIFY / (kms / 1e3) / 1e12 / imad_peak_tmacs,
                                      "source": "ncu SourceCounters, IMAD.WIDE.U32(.X) thread instructions per verify (profiles/r1e_ncu_summary.md)"},
                         "kernel": "k_schnorr_verify", "kernel_ms": kms,
                         "hbm": {"achieved_gbs": n * BYTES_PER_VERIFY / (kms / 1e3) / 1e9, "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback",
                                 "frac": n * BYTES_PER_VERIFY / (kms / 1e3) / 1e9 / hbm_peak}},
            "parity": parity,
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if extra is not None:
            line["extra"] = extra
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
