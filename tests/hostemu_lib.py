"""Builds (g++) and loads tests/hostemu/libhostemu.so -- the TEST-ONLY host build of the device
headers (portable limb path).  Lets the CPU tier exercise the exact group-law / ecmult / verify logic
that the kernels run.  Not part of the product; the product library never loads it."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "hostemu.cpp")
SO = os.path.join(HERE, "hostemu", "libhostemu.so")
CSRC = os.path.join(os.path.dirname(HERE), "secp256k1-zkp_b200", "csrc")


def _stale():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    srcs = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    return any(os.path.getmtime(s) > t for s in srcs)


_LIB = None


def load(gw=8):
    global _LIB
    if _LIB is None:
        if _stale():
            subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", SO, SRC])
        _LIB = ctypes.CDLL(SO)
        _LIB.emu_init(gw)
    return _LIB
