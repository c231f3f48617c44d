"""In-tree build of liblqft.so (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "liblqft.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]
LINK = ["-lcudart_static", "-ldl", "-lpthread", "-lrt"]


def _sources():
    out = []
    for d, _, files in os.walk(CSRC):
        out += [os.path.join(d, f) for f in files if f.endswith((".cu", ".cuh", ".h", ".hpp", ".cpp"))]
    out.append(os.path.join(ROOT, "include", "lqft.h"))
    return out


def stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(os.path.join(PKG, "lqft_sim")):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile lattice_qft_b200/csrc/lqft.cu -> lattice_qft_b200/liblqft.so."""
    if not force and not stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
        "-I", os.path.join(ROOT, "include"), "-o", LIB + ".tmp", os.path.join(CSRC, "lqft.cu")] + LINK
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(LIB + ".tmp", LIB)
    if verbose:
        print(res.stderr)
    build_host_driver()
    return LIB


SIM = os.path.join(PKG, "lqft_sim")


def build_host_driver() -> str:
    """g++ -> lattice_qft_b200/lqft_sim: src/bin/sim.rs written against the C++ host mirror."""
    cmd = ["g++", "-O2", "-std=c++17", "-o", SIM + ".tmp", os.path.join(CSRC, "host", "sim.cpp"), "-L", PKG, "-llqft",
           "-Wl,-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    os.replace(SIM + ".tmp", SIM)
    return SIM


if __name__ == "__main__":
    print(build(force=True, verbose=True))
