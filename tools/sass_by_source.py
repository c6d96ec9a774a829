"""Where a kernel's instructions go, by source location: joins the per-instruction counters of an ncu capture (--import-source on) with the
inline line information nvdisasm prints for the same cubin.

    python tools/sass_by_source.py <report.ncu-rep> <kernel regex> <cubin> [--by file|line|root] [--top N] [--only-root file:line]

Per group: static SASS instructions (code size), warp instructions executed, share of the kernel, stall samples and the share of
`stall_no_inst` samples (instruction-cache misses).  Groups: --by file = innermost source file; --by line = innermost file:line;
--by root = the line of the kernel's own file (or its first inlined-at ancestor) the instruction was inlined into."""
import collections
import csv
import io
import re
import subprocess
import sys


def disasm_map(cubin, kernel_re):
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", cubin], capture_output=True, text=True).stdout.splitlines()
    out = {}
    inside = False
    chain = []
    pending = []
    for ln in txt:
        if ln.startswith(".text."):
            inside = re.search(kernel_re, ln) is not None
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pending.append((m.group(1), int(m.group(2)), m.group(3), int(m.group(4)) if m.group(4) else None))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            if pending:
                chain = pending
                pending = []
            out[int(m.group(1), 16)] = chain
    return out


def main():
    rep, kre, cubin = sys.argv[1:4]
    by = sys.argv[sys.argv.index("--by") + 1] if "--by" in sys.argv else "file"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"],
                            capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(csvtxt)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    H = rows[h]
    ia, ie, isamp, inoi = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples"), H.index("stall_no_inst")
    base = None
    only = sys.argv[sys.argv.index("--only-root") + 1] if "--only-root" in sys.argv else None
    lines = disasm_map(cubin, kre)
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0, 0]
    for r in rows[h + 1:]:
        if r and r[0] == "Address":
            break                                            # a second launch of the same kernel follows: one is enough
        if len(r) <= inoi or not r[ia]:
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        chain = lines.get(addr - base, [])
        if only and chain:
            f, l = chain[-1][2] or chain[-1][0], chain[-1][3] or chain[-1][1]
            if "%s:%d" % (f.split("/")[-1], l) != only:
                continue
        if not chain:
            key = "?"
        elif by == "file":
            key = chain[0][0].split("/")[-1]
        elif by == "line":
            key = "%s:%d" % (chain[0][0].split("/")[-1], chain[0][1])
        else:
            f, l = chain[-1][2] or chain[-1][0], chain[-1][3] or chain[-1][1]
            key = "%s:%d" % (f.split("/")[-1], l)
        v = [1, int(r[ie] or 0), int(r[isamp] or 0), int(r[inoi] or 0)]
        for k in range(4):
            agg[key][k] += v[k]
            tot[k] += v[k]
    print("%-34s %9s %14s %7s %9s %9s" % ("group", "sass", "warp instrs", "share", "samples", "no_inst"))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%-34s %9d %14d %6.1f%% %9d %9d" % (key, v[0], v[1], 100.0 * v[1] / max(tot[1], 1), v[2], v[3]))
    print("%-34s %9d %14d %7s %9d %9d" % ("total", tot[0], tot[1], "", tot[2], tot[3]))


if __name__ == "__main__":
    main()
