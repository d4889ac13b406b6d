"""Builds libsrslte_b200.so in-tree with nvcc for sm_100a.  No torch involved: the library is plain CUDA runtime + C ABI.

Every source is compiled to its own object (in parallel, only when it or a header changed) and the objects are linked
into one shared library; `tools/libsrslte_b200_synth.so` (synthetic test/bench input generator, NOT part of the product
library) is built beside it from synth.cu."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libsrslte_b200.so")
SYNTH_DIR = os.path.join(ROOT, "tools", "synth")
SYNTH_LIB_PATH = os.path.join(SYNTH_DIR, "libsrslte_b200_synth.so")

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
    "uci_host.cu",
    "enb_ul.cu",
    "srsran_compat.cu",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function,-Wno-unknown-pragmas",
    "-DSRSLTE_B200_BUILD",
]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources() -> list[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".inc"))]
    inc = os.path.join(ROOT, "include")
    hs += [os.path.join(inc, f) for f in os.listdir(inc)]
    return hs


def _newest(paths) -> float:
    return max((os.path.getmtime(p) for p in paths if os.path.exists(p)), default=0.0)


def stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return _newest(sources() + _headers()) > t


def _compile(src: str, obj: str, verbose: bool) -> None:
    cmd = [nvcc_path(), *NVCC_FLAGS, "-c", "-o", obj, src]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
        print(" ".join(cmd), file=sys.stderr)
    subprocess.check_call(cmd, cwd=CSRC)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale() and os.path.exists(SYNTH_LIB_PATH):
        return LIB_PATH
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_t = _newest(_headers())
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for f in [ex.submit(_compile, s, o, verbose) for s, o in jobs]:
            f.result()
    subprocess.check_call([nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_PATH, *objs, "-lcudart"], cwd=CSRC)
    build_synth(force or bool(jobs))
    return LIB_PATH


def build_synth(force: bool = False) -> str:
    """The synthetic-input generator (test vectors for bench/tests) lives in its own library, outside the product."""
    src = os.path.join(SYNTH_DIR, "synth.cu")
    if not os.path.exists(src):
        return ""
    deps = [src] + _headers()
    if force or not os.path.exists(SYNTH_LIB_PATH) or os.path.getmtime(SYNTH_LIB_PATH) < _newest(deps):
        subprocess.check_call([nvcc_path(), *NVCC_FLAGS, "-shared", "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
                               "-o", SYNTH_LIB_PATH, src, "-lcudart"], cwd=SYNTH_DIR)
    return SYNTH_LIB_PATH


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
