"""Writes profiles/ncu_constants.json from an `ncu --set full --import-source on` capture of ONE launch of a kernel:

    python tools/ncu_constants.py <report.ncu-rep> <kernel name> <units in that launch> <source files...> [--from profiles/<summary file>]

Stores, next to the numbers bench.py prints (DRAM bytes of the launch, executed IMAD.WIDE per unit, multiply-pipe utilisation), the sha256 of
the kernel's source files, so that bench.py can tell when the numbers no longer describe the code (it then reports traffic = null and says so)."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--from")]
    src_note = None
    if "--from" in sys.argv:
        src_note = sys.argv[sys.argv.index("--from") + 1]
        args = [a for a in args if a != src_note]
    rep, kernel, units = args[0], args[1], int(args[2])
    files = args[3:]
    raw = ncu_csv(rep, "raw")
    H = raw[0]
    row = [r for r in raw[2:] if any(kernel in c for c in r)][0]
    val = dict(zip(H, row))
    unit = dict(zip(H, raw[1]))

    def metric(name):
        v = float(val[name].replace(",", ""))
        u = unit.get(name, "")
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
        return v * scale
    src = ncu_csv(rep, "source")
    hdr = [i for i, r in enumerate(src) if r and r[0] == "Address"][0]
    SH = src[hdr]
    si, ti = SH.index("Source"), SH.index("Thread Instructions Executed")
    wide = other = 0
    for r in src[hdr + 1:]:
        if len(r) <= ti:
            continue
        toks = r[si].split()
        op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
        if op.startswith("IMAD.WIDE"):
            wide += int(r[ti])
        else:
            other += int(r[ti])
    h = hashlib.sha256()
    for f in files:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    rec = {"source_files": files, "source_sha256": h.hexdigest(), "units_in_launch": units,
           "dram_bytes_2p20": int(metric("dram__bytes_read.sum") + metric("dram__bytes_write.sum")),
           "imad_wide_per_verify": wide / units, "thread_instructions_per_unit": (wide + other) / units,
           "pipe_fmaheavy_pct": float(val["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]),
           "duration_ms": metric("gpu__time_duration.sum") if unit.get("gpu__time_duration.sum") == "ms" else float(val["gpu__time_duration.sum"]),
           "from": src_note or os.path.relpath(rep, ROOT)}
    path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    allc = json.load(open(path)) if os.path.exists(path) else {}
    allc[kernel] = rec
    json.dump(allc, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
