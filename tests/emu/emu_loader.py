"""TEST-ONLY: build/load the host-loop emulation of the kernel bodies (g++ -DVFS_EMU).

Lets the parity suite exercise the *logic* of every kernel functor against the oracle in a
container without a GPU.  It is not part of the product: nothing under vfs-wind_b200/ references
this file or libvfs_emu.so, and bench.py / smoke() never load it.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "libvfs_emu.so")
SRC = os.path.join(ROOT, "vfs-wind_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "include", "vfs_b200.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    subprocess.check_call(["g++", "-x", "c++", "-std=c++17", "-O2", "-ffp-contract=off", "-DVFS_EMU", "-fPIC", "-shared", "-o", SO,
                           os.path.join(SRC, "vfs_ctx.cu")])
    return SO


def load(capi):
    return capi._bind(C.CDLL(build()))
