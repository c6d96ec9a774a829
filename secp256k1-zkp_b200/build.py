"""Builds libsecp256k1_b200.so (sm_100a only) in-tree with nvcc; incremental and parallel.

Usage: python secp256k1-zkp_b200/build.py [--force]
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# development A/B builds: B200_EXTRA_NVCC_FLAGS="-DFOO" B200_BUILD_TAG=foo -> build_foo/ and libsecp256k1_b200_foo.so
_TAG = os.environ.get("B200_BUILD_TAG", "")
EXTRA = os.environ.get("B200_EXTRA_NVCC_FLAGS", "").split()
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libsecp256k1_b200" + ("_" + _TAG if _TAG else "") + ".so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-diag-suppress", "550", "-diag-suppress", "177"]

# translation unit -> extra flags
UNITS = {
    "engine.cu": [],
    "k_fe.cu": ["-DB200_INLINE_FE"],
    "k_fe_portable.cu": ["-DB200_PORTABLE_BN", "-DB200_NS=b200_portable", "-DB200_INLINE_FE"],
    "k_gtable.cu": [],
    "k_verify.cu": [],
}


def _newest_header():
    t = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cuh", ".h")):
                t = max(t, os.path.getmtime(os.path.join(root, f)))
    return t


CXX = os.environ.get("CXX", "g++")


def _compile(unit, flags, force, hdr_time):
    src = os.path.join(CSRC, unit)
    obj = os.path.join(OBJ, unit.replace(".cu", ".o").replace(".cpp", ".o"))
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
            and os.path.getmtime(obj) > hdr_time):
        return obj, ""
    if unit.endswith(".cpp"):      # plain host C++ (no CUDA front end: x86 intrinsics headers stay out of cudafe's way)
        cmd = [CXX, "-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden"] + flags + ["-c", "-o", obj, src]
    else:
        cmd = [NVCC] + ARCH + COMMON + EXTRA + flags + ["-c", "-o", obj, src]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (unit, p.stdout, p.stderr))
    return obj, p.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    units = dict(UNITS)
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cpp")) and f not in units:
            units[f] = []
    hdr_time = _newest_header()
    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(units))) as ex:
        futs = [ex.submit(_compile, u, fl, force, hdr_time) for u, fl in units.items()]
        for f in futs:
            obj, log = f.result()
            objs.append(obj)
            if verbose and log:
                sys.stderr.write(log)
    newest_obj = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest_obj:
        cmd = [NVCC] + ARCH + ["-shared", "-cudart", "static", "-Xlinker", "-Bsymbolic", "-o", LIB] + objs
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (p.stdout, p.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
