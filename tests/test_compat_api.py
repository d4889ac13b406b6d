"""The reference-named C API (include/srslte_b200_srsran_api.h) driven from a plain C program the way the reference's own
unit tests drive it; every result is checked against the oracle inside the program (tests/c/compat_test.c)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "compat_test.c")
EXE = os.path.join(ROOT, "tests", "c", "compat_test")


def build():
    from srslte_b200 import build as b

    b.build_library()
    from oracle import loader

    loader.port()
    subprocess.check_call(["gcc", "-O1", "-std=gnu99", "-Wall", "-I" + os.path.join(ROOT, "include"), SRC, "-o", EXE,
                           "-L" + os.path.join(ROOT, "srslte_b200"), "-lsrslte_b200", "-L" + os.path.join(ROOT, "oracle"),
                           "-loracle_port", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "srslte_b200"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle"),
                           "-Wl,-rpath-link,/usr/local/cuda/lib64"])


def test_compat_program_compiles_and_links():
    """Header is valid C, and every reference-named symbol it declares resolves against the library (no GPU needed)."""
    build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_compat_program_runs():
    build()
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
