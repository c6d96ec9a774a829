"""bench.py -- headline benchmark: batch 2^20 independent BIP-340 schnorrsig_verify per GPU
(BASELINE.json configs[1]; one verify = one a*P + b*G ecmult + x-only pubkey parse, exactly what the
reference's `bench schnorrsig_verify` times, src/modules/schnorrsig/bench_impl.h:39-48).

  python bench.py --gpus N --steps K --warmup W            our arm (torchrun launches N ranks for N > 1)
  python bench.py --impl reference --gpus N ...            the reference's CPU implementation (oracle/_ref)

A step = one pass of the hot path over the rank's batch.  `value` = whole-job verifies/s with inputs resident in HBM (CUDA events, max
over ranks); `e2e` = the same through the reference-facing C ABI with HOST buffers (pinned inputs, H2D and D2H inside the timed region):
at N = 1 `secp256k1_b200_schnorrsig_verify_batch`, at N > 1 the single-process multi-device call
`secp256k1_b200_multi_schnorrsig_verify_batch` driven by rank 0 over all N GPUs (the per-rank single-device figure is kept beside it).
Weak scaling: every rank verifies its own 2^20 batch; the only collectives are the init-time NCCL broadcast of the generator table and the
per-step all-gather of the packed accept bitmaps.  `extra` holds the other BASELINE.json configs (ecmult, MSM, range proofs, BP++ norm
argument, Pedersen commitments, ECDSA, recovery, half-aggregation), each with its device-timed figure, an end-to-end figure through the
host-buffer call, its reference-MAC roofline fraction and -- at N = 1 -- the reference's CPU figure (1 pinned thread and one thread per
physical core) on a bounded sample of the same inputs.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# reference-algorithm 32x32->64 multiply-accumulates per unit (what `roofline.achieved` counts)
MAC_PER_ECMULT = 1.81e5
MAC_PER_MSM_TERM = 2.5e4
MAC_PER_RANGEPROOF64 = 2.7e7
MAC_PER_BPPP = 1.8e6
MAC_PER_COMMIT = 1.81e5          # blind*G + value*H: one ecmult-sized unit (the reference runs ecmult_gen + ecmult_const)
BYTES_PER_VERIFY = 64 + 32 + 32 + 1
N_ORD = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
# architectural integer-multiply peak of a B200: 148 SMs x 4 sub-partitions x 8 IMAD.WIDE lanes per clock x max SM clock
ARCH_SUBPARTITIONS = 148 * 4


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2n", type=int, default=20, help="verifies per GPU per step = 2^log2n")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads (ecmult / MSM / rangeproof / BP++ / ...)")
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
# host description and the CPU legs (oracle/_ref = the unmodified reference compiled from its own sources)
def host_info():
    """CPU model, logical CPUs this process may use, and one logical CPU per physical core among them"""
    model, phys = "unknown", {}
    try:
        cur = {}
        for line in open("/proc/cpuinfo"):
            if ":" not in line:
                if "processor" in cur:
                    phys.setdefault((cur.get("physical id", "0"), cur.get("core id", cur["processor"])), []).append(int(cur["processor"]))
                cur = {}
                continue
            k, v = [x.strip() for x in line.split(":", 1)]
            cur[k] = v
            if k == "model name":
                model = v
        if "processor" in cur:
            phys.setdefault((cur.get("physical id", "0"), cur.get("core id", cur["processor"])), []).append(int(cur["processor"]))
    except OSError:
        pass
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except Exception:
        allowed = list(range(os.cpu_count() or 1))
    aset = set(allowed)
    one_per_core = sorted(min(c for c in cpus if c in aset) for cpus in phys.values() if any(c in aset for c in cpus)) or allowed
    return {"cpu_model": model, "logical_cpus": len(allowed), "physical_cores": len(one_per_core), "allowed": allowed, "core_cpus": one_per_core}


def run_pinned(work, n_units, cpus):
    """work(lo, hi) over [0, n_units) split evenly over len(cpus) threads, thread t pinned to cpus[t] (ctypes calls release the GIL).
    Returns wall seconds."""
    T = len(cpus)
    per = (n_units + T - 1) // T
    errs = []

    def body(t):
        try:
            try:
                os.sched_setaffinity(0, {cpus[t]})          # pid 0 = the calling thread
            except Exception:
                pass
            lo, hi = t * per, min(n_units, (t + 1) * per)
            if lo < hi:
                work(lo, hi)
        except Exception as e:                               # pragma: no cover
            errs.append(repr(e))
    ts = [threading.Thread(target=body, args=(t,)) for t in range(T)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    dt = time.perf_counter() - t0
    if errs:
        raise RuntimeError(errs[0])
    return dt


def cpu_leg(hi_info, work, unit, n_one, n_per_thread, what, parallel=True):
    """reference CPU figure for one workload: 1 pinned thread over n_one units, then one thread per physical core over n_per_thread each"""
    cpus = hi_info["core_cpus"]
    work(0, min(n_one, 8))                                   # touch code and data once
    dt1 = run_pinned(work, n_one, cpus[:1])
    leg = {"unit": unit, "kind": "reference", "cpu_model": hi_info["cpu_model"],
           "one_thread": {"value": n_one / dt1, "units": n_one, "seconds": round(dt1, 3), "pinned_to_cpu": cpus[0]}}
    if parallel and len(cpus) > 1:
        n_mt = n_per_thread * len(cpus)
        dtm = run_pinned(work, n_mt, cpus)
        leg.update({"value": n_mt / dtm, "cores": len(cpus), "sample": "%s: %d units, one pinned thread per physical core (%d), %.2f s wall; 1 thread: %d units in %.2f s"
                    % (what, n_mt, len(cpus), dtm, n_one, dt1)})
    else:
        leg.update({"value": n_one / dt1, "cores": 1, "sample": "%s: %d units on 1 pinned thread, %.2f s" % (what, n_one, dt1)})
    return leg


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from tests.refshim import RefLib
    ref = RefLib()
    hi_info = host_info()
    cores = hi_info["logical_cpus"]
    unique = 4096                                            # distinct signed items, tiled to the sample size
    sk = b"".join(hashlib.sha256(b"bench-sk" + i.to_bytes(4, "little")).digest() for i in range(unique))
    msg = b"".join(hashlib.sha256(b"bench-msg" + i.to_bytes(4, "little")).digest() for i in range(unique))
    sig, pk = ref.schnorr_make(sk, msg)
    for i in range(7, unique, 16):
        bit = i % 512
        sig[64 * i + bit // 8] ^= 1 << (bit % 8)
    msg = np.frombuffer(msg, dtype=np.uint8).copy()
    per_thread = 1024
    sample = per_thread * cores
    reps = (sample + unique - 1) // unique
    sigs, msgs, pks = np.tile(sig, reps)[:64 * sample], np.tile(msg, reps)[:32 * sample], np.tile(pk, reps)[:32 * sample]
    acc = []

    def work(lo, hi):
        acc.append(int(ref.schnorr_verify(sigs[64 * lo:64 * hi], msgs[32 * lo:32 * hi], pks[32 * lo:32 * hi]).sum()))
    for _ in range(args.warmup):
        run_pinned(work, sample, hi_info["allowed"])
    t = 0.0
    acc.clear()
    for _ in range(args.steps):
        t += run_pinned(work, sample, hi_info["allowed"])     # every logical CPU this process may use
    ms = t / args.steps * 1e3
    value = sample / (ms / 1e3)
    one = cpu_leg(hi_info, work, "verify/s", 4096, 1024, "reference xonly_pubkey_parse + schnorrsig_verify")
    line = {
        "impl": "reference", "metric": "schnorrsig_verify/s (= ecmult/s)", "value": value, "unit": "verify/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64 (5x52 limbs, x86-64 int128)", "data": "synthetic",
        "config": {"workload": "batch 2^%d independent BIP-340 schnorrsig_verify incl. xonly_pubkey_parse (random keys/msgs, 1/16 corrupted)" % args.log2n,
                   "step": "bounded sample: %d verifies over %d host threads (every logical CPU available to the process)" % (sample, cores)},
        "cpu_baseline": {"value": value, "unit": "verify/s", "cores": cores, "kind": "reference", "cpu_model": hi_info["cpu_model"],
                         "physical_cores": hi_info["physical_cores"], "one_thread": one["one_thread"], "one_thread_per_physical_core": one.get("value"),
                         "sample": "%d verifies per step (%d per thread), reference secp256k1_schnorrsig_verify + xonly_pubkey_parse from oracle/_ref" % (sample, per_thread)},
        "e2e": {"value": value, "unit": "verify/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "accepted_fraction": sum(acc) / float(sample * args.steps),
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# device timing helpers
def time_dev(ctx, fn, warm=2, reps=5):
    """ms per call of fn (CUDA events on the launching stream, max over ranks)"""
    torch, dist = ctx["torch"], ctx["dist"]
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if ctx["world"] > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=ctx["dev"])
    if ctx["world"] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def time_host(ctx, fn, warm=1, reps=3):
    """ms per host-buffer call (wall clock around the blocking C call, max over ranks)"""
    torch, dist = ctx["torch"], ctx["dist"]
    r = None
    for _ in range(warm):
        r = fn()
    if ctx["world"] > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    ms = (time.perf_counter() - t0) / reps * 1e3
    t = torch.tensor([ms], dtype=torch.float64, device=ctx["dev"])
    if ctx["world"] > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), r


def pinned(torch, arr):
    """numpy array -> pinned host copy (numpy view of a pinned torch tensor; the tensor is kept alive by the caller)"""
    t = torch.from_numpy(arr.copy()).pin_memory()
    return t, t.numpy()


def roofline_entry(units_per_gpu, mac_per_unit, ms, peaks):
    ach = units_per_gpu * mac_per_unit / (ms / 1e3) / 1e12
    return {"bound": "int32 multiply pipe", "unit": "TMAC/s", "achieved_ref_mac": ach, "peak_probe": peaks["probe"], "peak_architectural": peaks["arch"],
            "frac": ach / peaks["probe"], "frac_of_architectural": ach / peaks["arch"], "mac_per_unit": mac_per_unit}


def load_ncu_constants(kernel):
    """measured-once constants of the dominant kernel (DRAM bytes of one 2^20 launch, executed IMAD.WIDE per verify), valid only while the
    kernel's sources are unchanged: profiles/ncu_constants.json stores the sha256 of the source files next to the numbers
    (tools/ncu_constants.py writes it from the committed ncu CSVs)."""
    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    try:
        rec = json.load(open(path))[kernel]
    except Exception as e:
        return None, "profiles/ncu_constants.json has no entry for %s (%s)" % (kernel, e)
    h = hashlib.sha256()
    for f in rec["source_files"]:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    if h.hexdigest() != rec["source_sha256"]:
        return None, "STALE: the sources of %s changed since %s was captured -- re-run the ncu pass and tools/ncu_constants.py" % (kernel, rec.get("from"))
    return rec, None


# ------------------------------------------------------------------------------------------------
# secondary workloads
def make_rangeproofs(ref, n, seed, hi_info):
    """n DISTINCT 64-bit range proofs signed by the reference prover (secp256k1_rangeproof_sign, exp 0, min_bits 64, min_value 0, nonce = commit
    -- src/bench_rangeproof.c:32-34 at 64 bits), in parallel over the host cores; value_i, blind_i seeded"""
    import numpy as np
    gen = ref.generator_h()
    blinds = [hashlib.sha256(b"rp-b" + seed + i.to_bytes(4, "little")).digest() for i in range(n)]
    vals = [int.from_bytes(hashlib.sha256(b"rp-v" + seed + i.to_bytes(4, "little")).digest()[:8], "little") for i in range(n)]
    proofs = np.zeros((n, 5126), dtype=np.uint8)
    commits = np.zeros((n, 33), dtype=np.uint8)

    def work(lo, hi):
        p, pl, c = ref.rangeproof_make(b"".join(blinds[lo:hi]), vals[lo:hi], gen, stride=5126)
        assert (np.asarray(pl) == 5126).all()
        proofs[lo:hi] = p
        commits[lo:hi] = np.asarray(c).reshape(-1, 33)
    run_pinned(work, n, hi_info["allowed"])
    return proofs, commits, np.frombuffer(b"".join(blinds), dtype=np.uint8).reshape(n, 32).copy(), np.array(vals, dtype=np.uint64), gen


def make_bppp(ref, n, seed, hi_info):
    """n DISTINCT norm-product arguments at the 64-bit geometry (16 G + 8 H generators, 324-byte proofs) from the reference prover
    (tools/make_bppp_fixture.py does the same for the committed 64-item fixture)"""
    import ctypes
    import numpy as np
    L = ref.lib
    L.secp256k1_bppp_generators_create.restype = ctypes.c_void_p
    g, h = 16, 8
    gs = ctypes.c_void_p(L.secp256k1_bppp_generators_create(ref.ctx, ctypes.c_size_t(g + h)))
    gxy = ctypes.create_string_buffer(64 * (g + h))
    L.ref_bppp_generators_xy(gxy, gs)

    def sc(tag, i, j=0):
        return (int.from_bytes(hashlib.sha256(b"%s|%s|%d|%d" % (tag, seed, i, j)).digest(), "big") % N_ORD or 1).to_bytes(32, "big")
    proofs = np.zeros((n, 324), dtype=np.uint8); rhos = np.zeros((n, 32), dtype=np.uint8)
    cvs = np.zeros((n, 32 * h), dtype=np.uint8); cms = np.zeros((n, 64), dtype=np.uint8)

    def work(lo, hi):
        for i in range(lo, hi):
            nv = b"".join(sc(b"bp-n", i, j) for j in range(g)); lv = b"".join(sc(b"bp-l", i, j) for j in range(h))
            cv = b"".join(sc(b"bp-c", i, j) for j in range(h)); rho = sc(b"bp-rho", i)
            proof = ctypes.create_string_buffer(1000); plen = ctypes.c_size_t(1000); cm = ctypes.create_string_buffer(64)
            assert L.ref_bppp_norm_prove(ref.ctx, proof, ctypes.byref(plen), cm, rho, gs, nv, ctypes.c_size_t(g), lv, cv, ctypes.c_size_t(h)) == 1 and plen.value == 324
            proofs[i] = np.frombuffer(proof.raw[:324], dtype=np.uint8); rhos[i] = np.frombuffer(rho, dtype=np.uint8)
            cvs[i] = np.frombuffer(cv, dtype=np.uint8); cms[i] = np.frombuffer(cm.raw, dtype=np.uint8)
    run_pinned(work, n, hi_info["allowed"])
    gens = np.frombuffer(gxy.raw, dtype=np.uint8).reshape(g + h, 64).copy()

    def cpu_verify(lo, hi):
        for j in range(lo, hi):
            i = j % n                                                # the CPU sample may be longer than the batch
            L.ref_bppp_norm_verify(ref.ctx, proofs[i].tobytes(), ctypes.c_size_t(324), cms[i].tobytes(), rhos[i].tobytes(), gs, ctypes.c_size_t(g), cvs[i].tobytes(), ctypes.c_size_t(h))
    return proofs, rhos, cvs, cms, gens, cpu_verify


def run_extra_workloads(ctx, peaks, ref, hi_info, want_cpu):
    import numpy as np
    torch, dist, D, eng, b200 = ctx["torch"], ctx["dist"], ctx["D"], ctx["eng"], ctx["b200"]
    dev, stream, rank, world = ctx["dev"], ctx["stream"], ctx["rank"], ctx["world"]
    out = {}
    seed = b"r%d" % rank

    def finish(name, entry, units_total, ms, unit_key, e2e_ms=None, h2d=0, d2h=0, mac=None, units_per_gpu=None, cpu=None, api=None):
        entry.update({"ms": ms, unit_key: units_total / ms * 1e3})
        if mac is not None:
            entry["roofline"] = roofline_entry(units_per_gpu if units_per_gpu is not None else units_total / world, mac, ms, peaks)
        if e2e_ms is not None:
            entry["e2e"] = {"value": units_total / e2e_ms * 1e3, "unit": unit_key.replace("_per_s", "/s"), "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                            "d2h_bytes_per_step": d2h, "api": api}
        if cpu is not None:
            entry["cpu_baseline"] = cpu
        out[name] = entry

    # ---- raw ecmult a*P + b*G (north_star target 1: secp256k1_ecmult/s), 2^20 per GPU ------------------------------------------------
    n = 1 << 20
    gen = torch.Generator(device=dev); gen.manual_seed(31337 + rank)
    sk = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen); sk[:, 0] &= 0x7F
    na = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen)
    ng = torch.randint(0, 256, (n, 32), dtype=torch.uint8, device=dev, generator=gen)
    z64, z32 = torch.zeros((n, 64), dtype=torch.uint8, device=dev), torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    pts = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    eng.ecmult_dev(pts.data_ptr(), z64.data_ptr(), z32.data_ptr(), sk.data_ptr(), n, stream)                 # P_i = sk_i * G
    res = torch.empty((n, 64), dtype=torch.uint8, device=dev)
    ms = time_dev(ctx, lambda: eng.ecmult_dev(res.data_ptr(), pts.data_ptr(), na.data_ptr(), ng.data_ptr(), n, stream))
    tp, hp = pinned(torch, pts.cpu().numpy()); tn, hn = pinned(torch, na.cpu().numpy()); tg, hg = pinned(torch, ng.cpu().numpy())
    tout = torch.empty((n, 64), dtype=torch.uint8).pin_memory()                       # pinned result buffer: D2H copies land in it directly
    e2e_ms, hres = time_host(ctx, lambda: eng.ecmult(hp, hn, hg, out=tout.numpy()))
    entry = {"items": n * world, "per_gpu": n, "scaling": "weak", "e2e_matches_device": bool(np.array_equal(hres, res.cpu().numpy()))}
    cpu = None
    if want_cpu:
        m = 4096
        exp = ref.ecmult_batch(hp[:m].tobytes(), hn[:m].tobytes(), hg[:m].tobytes())
        entry["oracle_mismatches"] = int((np.asarray(exp).reshape(m, 64) != hres[:m]).any(axis=1).sum())
        cpu = cpu_leg(hi_info, lambda lo, hi: ref.ecmult_batch(hp[lo:hi].tobytes(), hn[lo:hi].tobytes(), hg[lo:hi].tobytes()), "ecmult/s", 4096, 1024,
                      "reference secp256k1_ecmult a*P + b*G (what bench_ecmult's ecmult_1p_g row times twice per iteration, src/bench_ecmult.c:184-191)")
    finish("ecmult", entry, n * world, ms, "ecmult_per_s", e2e_ms, 128 * n, 64 * n, MAC_PER_ECMULT, n, cpu, "secp256k1_b200_ecmult_batch")
    del res, na, ng, z64, z32, tp, tn, tg, tout

    # ---- config 3: Pippenger 2^20-term MSM, strong scaling: rank r owns terms [r*N/W, (r+1)*N/W), partial sums all-gathered --------------
    n_total = 1 << 20
    lo, hi = D.shard_range(n_total, rank, world)
    m = hi - lo
    sc = torch.randint(0, 256, (n_total, 32), dtype=torch.uint8, device=dev, generator=gen)[lo:hi].contiguous()
    mpts = pts[lo:hi].contiguous(); msk = sk[lo:hi].contiguous()
    wsb = eng.ecmult_multi_workspace_bytes(max(m, 64))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    part = torch.empty(96 if world > 1 else 64, dtype=torch.uint8, device=dev)
    final = torch.empty(64, dtype=torch.uint8, device=dev)
    allp = torch.empty((world, 96), dtype=torch.uint8, device=dev)

    def msm_step():
        if world == 1:
            eng.ecmult_multi_dev(part.data_ptr(), None, sc.data_ptr(), mpts.data_ptr(), m, ws.data_ptr(), wsb, stream)
        else:   # shard sum stays Jacobian (no inversion), ONE all-gather of 96 bytes per rank, W-1 additions + one inversion in one block
            eng.ecmult_multi_partial_dev(part.data_ptr(), None, sc.data_ptr(), mpts.data_ptr(), m, ws.data_ptr(), wsb, stream)
            D.allgather_partials(part, allp)
            eng.ecmult_multi_combine_dev(final.data_ptr(), allp.data_ptr(), world, stream)
    ms = time_dev(ctx, msm_step)
    # identity check of bench_ecmult's teardown (src/bench_ecmult.c:73-93): MSM == (sum s_i*sk_i) * G, scalars on the host
    got = (final if world > 1 else part).cpu().numpy().tobytes()
    skh, sch = msk.cpu().numpy(), sc.cpu().numpy()
    tot = sum(int.from_bytes(a.tobytes(), "big") * int.from_bytes(b.tobytes(), "big") for a, b in zip(skh, sch)) % N_ORD
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, tot)
        tot = sum(parts) % N_ORD
    exp = eng.ecmult(bytes(64), bytes(32), tot.to_bytes(32, "big"))[0].tobytes()
    entry = {"terms": n_total, "scaling": "strong", "identity_check_ok": got == exp}
    if world > 1 and n_total % world == 0:
        # the other sharding: all 2^20 terms resident on EVERY GPU, rank r sums the bucket windows wi with wi % W == r (1/W of the bucket additions
        # AND of the bucket reduction, which sharding by terms repeats on every GPU); same all-gather of 96-byte partial results
        full_sc = torch.empty((n_total, 32), dtype=torch.uint8, device=dev); full_pts = torch.empty((n_total, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(full_sc, sc); dist.all_gather_into_tensor(full_pts, mpts)
        wsb_f = eng.ecmult_multi_workspace_bytes(n_total)
        ws_f = torch.empty(wsb_f, dtype=torch.uint8, device=dev)

        def msm_step_windows():
            eng.ecmult_multi_window_shard_dev(part.data_ptr(), None, full_sc.data_ptr(), full_pts.data_ptr(), n_total, rank, world, ws_f.data_ptr(), wsb_f, stream)
            D.allgather_partials(part, allp)
            eng.ecmult_multi_combine_dev(final.data_ptr(), allp.data_ptr(), world, stream)
        ms_w = time_dev(ctx, msm_step_windows)
        entry["sharding"] = {"by_terms_ms": ms, "by_windows_ms": ms_w, "by_windows_identity_check_ok": final.cpu().numpy().tobytes() == exp,
                             "reported": "by_windows" if ms_w < ms else "by_terms",
                             "note": "by_terms: rank r holds terms [rN/W, (r+1)N/W) only; by_windows: all terms resident on every GPU (replicated inputs), "
                                     "rank r sums the bucket windows wi = r mod W"}
        ms = min(ms, ms_w)
        del full_sc, full_pts, ws_f
    tps, hps = pinned(torch, mpts.cpu().numpy()); tss, hss = pinned(torch, sch)
    e2e_ms, hout = time_host(ctx, lambda: eng.ecmult_multi(hss, hps))
    if world == 1:
        entry["e2e_matches_device"] = hout == got
    cpu = None
    if want_cpu:
        def msm_cpu(k):
            t0 = time.perf_counter()
            ok, pt = ref.ecmult_multi(hss[:k].tobytes(), hps[:k].tobytes(), None, scratch_bytes=(420 << 20) if k > 100000 else (16 << 20))
            return time.perf_counter() - t0, ok, pt
        d1, ok1, _ = msm_cpu(32767)
        d2, ok2, pt2 = msm_cpu(n_total)
        entry["oracle_full_msm_matches"] = bool(ok2 == 1 and pt2 == got)
        cpu = {"value": n_total / d2, "unit": "terms/s", "cores": 1, "kind": "reference", "cpu_model": hi_info["cpu_model"],
               "sample": "reference secp256k1_ecmult_multi_var (Pippenger), single thread by construction: 2^20 terms with a 420 MB scratch in %.2f s; "
                         "32767 terms (bench_ecmult's largest row, src/bench_ecmult.c:387-391) in %.3f s = %.0f terms/s" % (d2, d1, 32767 / d1)}
    finish("msm_pippenger", entry, n_total, ms, "terms_per_s", e2e_ms if world == 1 else None, 96 * m, 64, MAC_PER_MSM_TERM, n_total / world, cpu,
           "secp256k1_b200_ecmult_multi")
    if world > 1:   # the per-rank host call above times a 1/W shard on one GPU; the sharded host-buffer call is the multi engine's (rank 0, below)
        out["msm_pippenger"]["e2e_note"] = "see e2e_multi_device.msm"
    del ws, sc, tps, tss

    # ---- config 4: 64-bit range proofs, 8192 DISTINCT reference-signed proofs per GPU (65536 on 8 GPUs), 1/32 corrupted ------------------
    per = 8192
    t0 = time.perf_counter()
    proofs_h, commits_h, blinds_h, vals_h, genh = make_rangeproofs(ref, per, seed, hi_info)
    gen_s = time.perf_counter() - t0
    bad = np.arange(3, per, 32)
    for i in bad:
        off = int.from_bytes(hashlib.sha256(b"%d" % i).digest()[:4], "big") % (8 * 5126)
        proofs_h[i, off // 8] ^= 1 << (off % 8)
    tpr, hpr = pinned(torch, proofs_h); tcm, hcm = pinned(torch, commits_h)
    proofs = tpr.to(dev); commits = tcm.to(dev)
    plen_h = np.full(per, 5126, dtype=np.uint32)
    plen = torch.from_numpy(plen_h.astype(np.int32)).to(dev)
    g64 = torch.frombuffer(bytearray(genh), dtype=torch.uint8).to(dev)
    okr = torch.zeros(per, dtype=torch.uint8, device=dev)
    mn = torch.zeros(per, dtype=torch.int64, device=dev); mx = torch.zeros(per, dtype=torch.int64, device=dev)

    def rp_step():
        eng.rangeproof_verify_dev(okr.data_ptr(), mn.data_ptr(), mx.data_ptr(), commits.data_ptr(), proofs.data_ptr(), 5126, plen.data_ptr(),
                                  g64.data_ptr(), 0, per, stream)
        if world > 1:
            D.allgather_bitmaps(D.pack_bitmap(okr))
    ms = time_dev(ctx, rp_step, warm=1, reps=3)
    expok = np.ones(per, dtype=np.uint8); expok[bad] = 0
    e2e_ms, hres = time_host(ctx, lambda: eng.rangeproof_verify(hcm, hpr, plen_h, genh), warm=1, reps=2)
    entry = {"proofs": per * world, "per_gpu": per, "scaling": "weak", "accept_pattern_ok": bool(np.array_equal(okr.cpu().numpy(), expok)),
             "max_value_ok": bool((mx == -1).all()), "e2e_matches_device": bool(np.array_equal(hres[0], okr.cpu().numpy())),
             "data": "%d distinct 64-bit proofs per GPU signed by the reference prover in %.1f s (seeded values / blinds, SURVEY 8d config 4), 1/32 with one flipped bit" % (per, gen_s)}
    cpu = None
    if want_cpu:
        k = 256
        eok = ref.rangeproof_verify(hcm[:k], hpr[:k], plen_h[:k], genh)
        entry["oracle_mismatches"] = int((eok[0] != hres[0][:k]).sum() + (eok[2] != hres[2][:k]).sum())
        cpu = cpu_leg(hi_info, lambda lo, hi: ref.rangeproof_verify(hcm[lo:hi], hpr[lo:hi], plen_h[lo:hi], genh), "verify/s", 128, 32,
                      "reference secp256k1_rangeproof_verify on 64-bit proofs (bench_rangeproof is hard-wired to 32 bits, src/bench_rangeproof.c:59)")
    finish("rangeproof_verify_64bit", entry, per * world, ms, "verify_per_s", e2e_ms, per * (5126 + 33 + 4) + 64, per * 17, MAC_PER_RANGEPROOF64, per, cpu,
           "secp256k1_b200_rangeproof_verify_batch")

    # ---- Pedersen commitments over the same (blind, value) pairs: must reproduce the prover's commitments ------------------------------
    reps_c = 32
    nb = per * reps_c                                                      # 2^18 commitments per GPU
    bl = torch.from_numpy(np.tile(blinds_h, (reps_c, 1))).to(dev)
    # distinct values per copy so that no two commitments coincide
    vv = (np.tile(vals_h, reps_c) + np.repeat(np.arange(reps_c, dtype=np.uint64), per))
    vt = torch.from_numpy(vv.astype(np.int64)).to(dev)
    okc = torch.zeros(nb, dtype=torch.uint8, device=dev); cm33 = torch.zeros((nb, 33), dtype=torch.uint8, device=dev)
    ms = time_dev(ctx, lambda: eng.pedersen_commit_dev(okc.data_ptr(), cm33.data_ptr(), bl.data_ptr(), vt.data_ptr(), g64.data_ptr(), 0, nb, stream), warm=1, reps=3)
    tb, hb = pinned(torch, np.tile(blinds_h, (reps_c, 1)))
    tco = torch.empty((nb, 33), dtype=torch.uint8).pin_memory()
    e2e_ms, hcres = time_host(ctx, lambda: eng.pedersen_commit(hb, vv, genh, out=tco.numpy()))
    entry = {"commitments": nb * world, "per_gpu": nb, "scaling": "weak", "all_ok": bool(okc.all()),
             "reproduces_prover_commitments": bool(np.array_equal(cm33[:per].cpu().numpy(), commits_h)),
             "e2e_matches_device": bool(np.array_equal(hcres[1], cm33.cpu().numpy()))}
    cpu = None
    if want_cpu:
        cpu = cpu_leg(hi_info, lambda lo, hi: ref.pedersen_commit(hb[lo:hi].tobytes(), vv[lo:hi], genh), "commit/s", 4096, 1024,
                      "reference secp256k1_pedersen_commit (ecmult_gen + ecmult_const, constant time)")
    finish("pedersen_commit", entry, nb * world, ms, "commit_per_s", e2e_ms, nb * 40 + 64, nb * 34, MAC_PER_COMMIT, nb, cpu, "secp256k1_b200_pedersen_commit_batch")
    del proofs, commits, tpr, tcm, bl, vt, cm33, tb

    # ---- config 5: BP++ norm-product argument, 16 G + 8 H generators, 512 DISTINCT reference-made arguments per GPU (4096 on 8 GPUs) ------
    per = 512
    bp, rho, cv, cm, gens, bp_cpu = make_bppp(ref, per, seed, hi_info)
    badi = np.arange(5, per, 16)
    bp[badi, 323] ^= 1
    tr = eng.bppp_transcripts(cm, rho, gens, 16, cv, 8, mode="tests")
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in (("p", bp), ("rho", rho), ("cv", cv), ("cm", cm), ("tr", tr), ("g", gens))}
    pl = torch.full((per,), 324, dtype=torch.int32, device=dev)
    okb = torch.zeros(per, dtype=torch.uint8, device=dev)

    def bp_step():
        eng._ck(eng._lib.secp256k1_b200_bppp_norm_verify_batch_dev(eng._h, okb.data_ptr(), d["p"].data_ptr(), 324, pl.data_ptr(), d["tr"].data_ptr(),
                                                                   d["rho"].data_ptr(), d["g"].data_ptr(), 16, d["cv"].data_ptr(), 8, d["cm"].data_ptr(), per, stream))
        if world > 1:
            D.allgather_bitmaps(D.pack_bitmap(okb))
    ms = time_dev(ctx, bp_step, warm=1, reps=3)
    expb = np.ones(per, dtype=np.uint8); expb[badi] = 0
    plh = np.full(per, 324, dtype=np.uint32)
    e2e_ms, hb_ok = time_host(ctx, lambda: eng.bppp_norm_verify(bp, plh, tr, rho, gens, 16, cv, 8, cm))
    entry = {"proofs": per * world, "per_gpu": per, "scaling": "weak", "accept_pattern_ok": bool(np.array_equal(okb.cpu().numpy(), expb)),
             "e2e_matches_device": bool(np.array_equal(hb_ok, okb.cpu().numpy())),
             "data": "%d distinct reference-made arguments per GPU, 1/16 corrupted" % per,
             "note": "norm-product argument only: this reference snapshot has no full BP++ range proof (SURVEY fact 3); 512 proofs = 0.3 waves of warps, "
                     "the kernel is latency bound at this batch size"}
    cpu = cpu_leg(hi_info, bp_cpu, "verify/s", 256, 16, "reference secp256k1_bppp_rangeproof_norm_product_verify (no reference bench exists)") if want_cpu else None
    # the same kernel with enough warps to fill the GPU: 8192 DISTINCT arguments per GPU (5.5 waves of warps) -- what the kernel sustains
    big = 8192
    bp2, rho2, cv2, cm2, gens2, _ = make_bppp(ref, big, seed + b"big", hi_info)
    tr2 = eng.bppp_transcripts(cm2, rho2, gens2, 16, cv2, 8, mode="tests")
    d2 = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in (("p", bp2), ("rho", rho2), ("cv", cv2), ("cm", cm2), ("tr", tr2), ("g", gens2))}
    pl2 = torch.full((big,), 324, dtype=torch.int32, device=dev)
    okb2 = torch.zeros(big, dtype=torch.uint8, device=dev)
    ms2 = time_dev(ctx, lambda: eng._ck(eng._lib.secp256k1_b200_bppp_norm_verify_batch_dev(
        eng._h, okb2.data_ptr(), d2["p"].data_ptr(), 324, pl2.data_ptr(), d2["tr"].data_ptr(), d2["rho"].data_ptr(), d2["g"].data_ptr(), 16,
        d2["cv"].data_ptr(), 8, d2["cm"].data_ptr(), big, stream)), warm=1, reps=3)
    entry["batch_8192"] = {"proofs_per_gpu": big, "all_accepted": bool(okb2.all()), "ms": ms2, "verify_per_s": big * world / ms2 * 1e3,
                           "roofline": roofline_entry(big, MAC_PER_BPPP, ms2, peaks)}
    finish("bppp_norm_verify_16g8h", entry, per * world, ms, "verify_per_s", e2e_ms, per * (324 + 104 + 32 + 256 + 64 + 4) + 24 * 64, per, MAC_PER_BPPP, per, cpu,
           "secp256k1_b200_bppp_norm_verify_batch")
    del d, d2

    # ---- ECDSA verify / recover, 2^18 per GPU: r = x of a BIP-340 nonce point (always liftable), s random ------------------------------
    per = 1 << 18
    gen2 = torch.Generator(device=dev); gen2.manual_seed(4242 + rank)
    sk2 = torch.randint(0, 256, (per, 32), dtype=torch.uint8, device=dev, generator=gen2); sk2[:, 0] &= 0x7F; sk2[:, 31] |= 1
    msg = torch.randint(0, 256, (per, 32), dtype=torch.uint8, device=dev, generator=gen2)
    sig = torch.empty((per, 64), dtype=torch.uint8, device=dev); pk = torch.empty((per, 32), dtype=torch.uint8, device=dev)
    eng.testdata_schnorrsig_sign_dev(sig.data_ptr(), pk.data_ptr(), sk2.data_ptr(), msg.data_ptr(), 32, per, stream)
    rsig = sig.clone(); rsig[:, 32] &= 0x3F                                 # low s
    rid = torch.randint(0, 2, (per,), dtype=torch.uint8, device=dev, generator=gen2)
    okc = torch.empty(per, dtype=torch.uint8, device=dev); rec = torch.empty((per, 64), dtype=torch.uint8, device=dev)
    ms = time_dev(ctx, lambda: eng.ecdsa_recover_dev(okc.data_ptr(), rec.data_ptr(), rsig.data_ptr(), rid.data_ptr(), msg.data_ptr(), per, stream), warm=1, reps=3)
    h_sig, h_msg, h_rid = rsig.cpu().numpy(), msg.cpu().numpy(), rid.cpu().numpy()
    trec = torch.empty((per, 64), dtype=torch.uint8).pin_memory()
    ths, thm = rsig.cpu().pin_memory(), msg.cpu().pin_memory()
    h_sig, h_msg = ths.numpy(), thm.numpy()
    e2e_ms, hrec = time_host(ctx, lambda: eng.ecdsa_recover(h_sig, h_rid, h_msg, out=trec.numpy()))
    finish("ecdsa_recover", {"signatures": per * world, "per_gpu": per, "scaling": "weak", "all_recovered": bool(okc.all()),
                             "e2e_matches_device": bool(np.array_equal(hrec[1], rec.cpu().numpy()))},
           per * world, ms, "recover_per_s", e2e_ms, per * 97, per * 65, MAC_PER_ECMULT, per, None, "secp256k1_b200_ecdsa_recover_batch")
    # the recovered keys verify the signatures they were recovered from: the ECDSA verify workload (compressed keys, compact signatures)
    rk = rec.cpu().numpy()
    comp = np.concatenate([(2 + (rk[:, 63] & 1))[:, None].astype(np.uint8), rk[:, :32]], axis=1)
    dcomp = torch.from_numpy(comp).to(dev); okv = torch.empty(per, dtype=torch.uint8, device=dev)
    ms = time_dev(ctx, lambda: eng.ecdsa_verify_dev(okv.data_ptr(), rsig.data_ptr(), b200.SIG_COMPACT64, msg.data_ptr(), dcomp.data_ptr(), b200.PK_COMPRESSED33, per, stream),
                  warm=1, reps=3)
    e2e_ms, hv = time_host(ctx, lambda: eng.ecdsa_verify(h_sig, h_msg, comp, b200.SIG_COMPACT64, b200.PK_COMPRESSED33))
    cpu = None
    if want_cpu:
        cpu = cpu_leg(hi_info, lambda lo, hi: ref.ecdsa_verify(h_sig[lo:hi].tobytes(), h_msg[lo:hi].tobytes(), comp[lo:hi].tobytes()), "verify/s", 4096, 1024,
                      "reference ec_pubkey_parse + ecdsa_signature_parse_compact + ecdsa_verify")
    finish("ecdsa_verify", {"signatures": per * world, "per_gpu": per, "scaling": "weak", "recovered_keys_verify": bool(okv.all()), "e2e_all_valid": bool(hv.all())},
           per * world, ms, "verify_per_s", e2e_ms, per * 129, per, MAC_PER_ECMULT, per, cpu, "secp256k1_b200_ecdsa_verify_batch")

    # ---- one half-aggregate signature over the same 2^18 (key, message, r) triples per GPU -----------------------------
    agg = torch.cat([sig[:, :32].reshape(-1), torch.zeros(32, dtype=torch.uint8, device=dev)]).contiguous()
    h_pk, h_agg = pk.cpu().numpy(), agg.cpu().numpy()
    h_ok = eng.schnorrsig_aggverify(h_pk, h_msg, h_agg)                     # s = 0: must be rejected
    t0 = time.perf_counter()
    eng.schnorrsig_aggverify(h_pk, h_msg, h_agg)
    out["schnorrsig_aggverify"] = {"signatures_per_aggregate": per, "host_call_ms": (time.perf_counter() - t0) * 1e3, "rejects_wrong_s": h_ok == 0,
                                   "note": "host-buffer call incl. the serial randomizer hash chain on the host; each rank verifies its own aggregate"}
    return out


# ------------------------------------------------------------------------------------------------
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
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")           # host-side rendezvous (a NCCL barrier would keep a kernel spinning on every waiting GPU)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- engine + generator table: rank 0 builds, ONE NCCL broadcast replicates it ---------------
    from secp256k1_zkp_b200.host import dist as D
    eng = b200.Engine(device=local_rank, build_tables=(rank == 0))
    D.replicate_engine_table(eng, rank, dev, stream)
    ctx = {"torch": torch, "dist": dist, "D": D, "eng": eng, "b200": b200, "dev": dev, "stream": stream, "rank": rank, "world": world}

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
    arch_peak = ARCH_SUBPARTITIONS * 8 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
    peaks = {"probe": imad_peak_tmacs, "arch": arch_peak}

    # ---- e2e, per rank: the host-buffer C-ABI call, pinned host inputs, H2D + D2H inside the timed region -------------
    h_sig, h_msg, h_pk = sig.cpu().pin_memory(), msg.cpu().pin_memory(), pk.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms, h_ok = time_host(ctx, lambda: eng.schnorrsig_verify(h_sig.numpy(), h_msg.numpy(), h_pk.numpy(), b200.PK_XONLY32), warm=2, reps=e2e_steps)
    e2e_ok = bool(np.array_equal(h_ok, expected.cpu().numpy()))
    e2e = {"value": world * n / (e2e_ms / 1e3), "unit": "verify/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": 128 * n * world, "d2h_bytes_per_step": n * world,
           "api": "secp256k1_b200_schnorrsig_verify_batch (host buffers, pinned), one process per GPU"}

    # ---- e2e at N > 1 through the C-level multi-device engine: ONE process (rank 0) drives all N GPUs, the other ranks wait on the host --------
    e2e_multi = None
    if world > 1:
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                me = b200.MultiEngine(devices=list(range(world)))
                big = [t.repeat(world, 1).pin_memory() for t in (h_sig, h_msg, h_pk)]        # N x 2^20 items in ONE set of host arrays
                outb = np.empty(world * n, dtype=np.uint8)
                l0 = me.launch_count
                for _ in range(2):
                    me.schnorrsig_verify(big[0].numpy(), big[1].numpy(), big[2].numpy(), b200.PK_XONLY32, out=outb)
                tw = time.perf_counter()
                for _ in range(e2e_steps):
                    me.schnorrsig_verify(big[0].numpy(), big[1].numpy(), big[2].numpy(), b200.PK_XONLY32, out=outb)
                mms = (time.perf_counter() - tw) / e2e_steps * 1e3
                e2e_multi = {"value": world * n / (mms / 1e3), "unit": "verify/s", "ms_per_step": mms, "h2d_bytes_per_step": 128 * n * world,
                             "d2h_bytes_per_step": n * world, "devices": me.device_count, "nccl_version": me.nccl_version,
                             "kernel_launches": (me.launch_count - l0) // (e2e_steps + 2),
                             "all_match": bool(np.array_equal(outb, np.tile(expected.cpu().numpy(), world))),
                             "api": "secp256k1_b200_multi_schnorrsig_verify_batch: one process, %d devices (ncclCommInitAll + one table broadcast at create), "
                                    "host arrays sharded by contiguous ranges" % world}
                # the sharded MSM through the same engine: 2^20 terms in host arrays, partial sums all-gathered with NCCL, combined on device 0
                mgen = np.random.default_rng(99)
                msc = mgen.integers(0, 256, size=(1 << 20, 32), dtype=np.uint8)
                mpt = eng.ecmult(np.zeros((1 << 20, 64), dtype=np.uint8), np.zeros((1 << 20, 32), dtype=np.uint8), mgen.integers(0, 128, size=(1 << 20, 32), dtype=np.uint8))
                tsc, tpt = torch.from_numpy(msc).pin_memory(), torch.from_numpy(mpt.copy()).pin_memory()
                r1 = me.ecmult_multi(tsc.numpy(), tpt.numpy())
                tw = time.perf_counter()
                for _ in range(3):
                    r1 = me.ecmult_multi(tsc.numpy(), tpt.numpy())
                e2e_multi["msm"] = {"terms": 1 << 20, "ms_per_call": (time.perf_counter() - tw) / 3 * 1e3, "matches_single_device": r1 == eng.ecmult_multi(msc, mpt),
                                    "api": "secp256k1_b200_multi_ecmult_multi (host buffers: 96 MiB H2D per call inside the timed region)"}
                me.close()
            except Exception as ex:
                e2e_multi = {"error": repr(ex)}
        dist.barrier(group=cpu_group)

    # ---- parity sample against the oracle + CPU baseline (rank 0) -------------------------------------------
    parity = {"full_batch_matches_construction": full_ok and e2e_ok}
    cpu_baseline = None
    ref = None
    hi_info = host_info()
    try:
        from tests.refshim import RefLib
        ref = RefLib()
    except OSError as e:
        parity.update({"oracle": "unavailable: %s" % e})
    want_cpu = (rank == 0 and world == 1 and not args.no_cpu_baseline and ref is not None)
    if rank == 0 and ref is not None:
        m = 4096
        hs, hm, hp = h_sig.numpy(), h_msg.numpy(), h_pk.numpy()
        exp = ref.schnorr_verify(hs[:m].reshape(-1), hm[:m].reshape(-1), hp[:m].reshape(-1))
        parity.update({"oracle": "oracle/_ref (reference)", "oracle_checked": m, "oracle_mismatches": int((exp != h_ok[:m]).sum())})
        if want_cpu:
            cpu_baseline = cpu_leg(hi_info, lambda lo, hi: ref.schnorr_verify(hs[lo:hi].reshape(-1), hm[lo:hi].reshape(-1), hp[lo:hi].reshape(-1)), "verify/s",
                                   8192, 2048, "first items of the same batch, reference xonly_pubkey_parse + schnorrsig_verify (oracle/_ref)")
            cpu_baseline["physical_cores"] = hi_info["physical_cores"]
            cpu_baseline["logical_cpus"] = hi_info["logical_cpus"]

    extra = None
    if not args.no_extra:
        try:
            if ref is None:
                raise RuntimeError("oracle/_ref is needed to make the distinct range-proof / BP++ inputs")
            extra = run_extra_workloads(ctx, peaks, ref, hi_info, want_cpu)
        except Exception as ex:  # the headline line must survive a failing secondary workload
            import traceback
            extra = {"error": repr(ex), "trace": traceback.format_exc()[-1500:]}

    if rank == 0:
        total = world * n
        value = total / (ms / 1e3)
        achieved = n * MAC_PER_ECMULT / (kms / 1e3) / 1e12
        mpeaks = {}
        try:
            mpeaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = mpeaks.get("hbm_gbs", 6650.0)
        ncu, ncu_note = load_ncu_constants("k_schnorr_verify")
        traffic = ncu["dram_bytes_2p20"] if (ncu and args.log2n == 20) else None
        executed = None
        if ncu:
            ex_rate = n * ncu["imad_wide_per_verify"] / (kms / 1e3) / 1e12
            executed = {"imad_wide_per_verify": ncu["imad_wide_per_verify"], "tmac_s": ex_rate, "frac_of_probe": ex_rate / imad_peak_tmacs,
                        "frac_of_architectural": ex_rate / arch_peak, "pipe_fmaheavy_pct": ncu.get("pipe_fmaheavy_pct"), "source": ncu.get("from")}
        line = {
            "metric": "schnorrsig_verify/s (= ecmult/s)", "value": value, "unit": "verify/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (8x32-bit limbs, IMAD.WIDE.U32)", "data": "synthetic",
            "config": {"workload": "batch 2^%d independent BIP-340 schnorrsig_verify per GPU incl. x-only pubkey parse (random keys/msgs, 1/16 corrupted), %d GPU(s)" % (args.log2n, world),
                       "per_gpu_batch": n, "global_batch": total, "l2": "inputs %.0f MiB per pass > 126 MB L2" % (n * 128 / 2**20),
                       "parallelism": "independent shards, table broadcast at init, bitmap all-gather per step"},
            "clocks": clocks,
            "e2e": e2e if (world == 1 or not e2e_multi or "error" in e2e_multi) else dict(e2e_multi, per_rank_processes=e2e),
            "gpu_launches": launches,
            "roofline": {"bound": "int32 multiply pipe (IMAD.WIDE.U32); not hbm, not tensor", "achieved": achieved, "peak": imad_peak_tmacs, "unit": "TMAC/s",
                         "frac": achieved / imad_peak_tmacs, "peak_architectural": arch_peak, "frac_of_architectural": achieved / arch_peak, "traffic": traffic,
                         "traffic_note": ncu_note,
                         "how": "algorithmic 1.81e5 MAC per verify (SURVEY 8d, reference algorithm) x %d verifies / %.3f ms kernel (CUDA events); peak = live IMAD.WIDE.U32 probe "
                                "(16 independent accumulators, distinct multiplicands); peak_architectural = 592 sub-partitions x 8 lanes x max SM clock; traffic = dram "
                                "read+write bytes of one 2^20 launch from the committed ncu capture (profiles/ncu_constants.json, checked against the kernel's source hash)" % (n, kms),
                         "executed": executed, "kernel": "k_schnorr_verify", "kernel_ms": kms,
                         "hbm": {"achieved_gbs": n * BYTES_PER_VERIFY / (kms / 1e3) / 1e9, "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json (measured)" if mpeaks else "fallback",
                                 "frac": n * BYTES_PER_VERIFY / (kms / 1e3) / 1e9 / hbm_peak}},
            "parity": parity,
            "host": {k: hi_info[k] for k in ("cpu_model", "logical_cpus", "physical_cores")},
        }
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        if extra is not None:
            line["extra"] = extra
        if e2e_multi is not None:
            line["e2e_multi_device"] = e2e_multi
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
