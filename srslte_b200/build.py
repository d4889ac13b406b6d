"""Builds libsrslte_b200.so in-tree with nvcc for sm_100a.  No torch involved: the library is plain CUDA runtime + C ABI."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsrslte_b200.so")

SOURCES = [
    "b200_runtime.cu",
    "tdec_kernels.cu",
    "tdec_host.cu",
    "rm_kernels.cu",
    "sch_host.cu",
    "ofdm_kernels.cu",
    "ofdm_host.cu",
    "demod_kernels.cu",
    "pusch_kernels.cu",
    "enb_ul.cu",
    "srsran_compat.cu",
    "synth.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function,-Wno-unknown-pragmas",
    "-DSRSLTE_B200_BUILD",
    "-shared",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, "..", "include", "srslte_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB_PATH
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB_PATH, *sources(), "-lcudart"]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
