"""CPU emulation of the CUDA decoder data path (tests/emu/tdec_emu.cpp runs the very per-lane code of
srslte_b200/csrc/tdec_core.h with host versions of the packed int16 ops) against the oracle.  Covers the tile layout,
checkpoint/recompute schedule, in-place extrinsic exchange, CRC syndrome, early stop and freezing of finished blocks
without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import coded_llrs, npass_of

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libtdec_emu.so")
    src = os.path.join(HERE, "emu", "tdec_emu.cpp")
    hdrs = [os.path.join(ROOT, "srslte_b200", "csrc", h) for h in ("tdec_core.h", "packed16.h", "lte_tables.h")]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in [src] + hdrs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I" + os.path.join(ROOT, "srslte_b200", "csrc"), "-o", so, src])
    return C.CDLL(so)


def run_emu(emu, llr, K, max_pass, crc_kind, early, force16=False, split=47):
    ncb = llr.shape[0]
    out = np.zeros((ncb, K // 8), np.uint8)
    ok = np.zeros(ncb, np.uint8)
    nc = np.zeros(ncb, np.uint8)
    nr = np.zeros(ncb, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu.emu_tdec_batch2(p(llr), ncb, K, max_pass, crc_kind, int(early), int(force16), split, p(out), p(ok), p(nc), p(nr)) == 0
    return out, ok, nc, nr


@pytest.mark.parametrize("K,ncb,sigma,scale,clip", [(40, 70, 0.8, 16, 31), (48, 3, 1.0, 16, 31), (504, 5, 0.9, 16, 31),
                                                    (1024, 66, 1.0, 16, 31), (6144, 3, 0.92, 16, 31),
                                                    (6144, 2, 0.8, 8000, 30000), (2048, 3, 1.2, 500, 2000)])
def test_emulated_kernels_match_oracle(emu, port, K, ncb, sigma, scale, clip):
    llr, _ = coded_llrs(port, K, ncb, sigma, scale, clip, seed=K + ncb)
    for early in (True, False):
        for mp in (8, 5, 1):
            o1, k1, n1, _ = port.decode_batch(llr, K, mp, "B", 0, early)
            for force16, split in ((False, 47), (True, 47), (False, 1), (True, 99)):
                o2, k2, nc, nr = run_emu(emu, llr, K, mp, 0, early, force16, split)
                assert (o1 == o2).all(), (early, mp, force16, split)
                assert (k1 == k2).all() and (n1 == npass_of(k2, nc, nr)).all(), (early, mp, force16, split)


def test_emulation_all_188_sizes(emu, port):
    """config 3 of BASELINE.json in miniature: every LTE QPP size, two blocks each, mixed outcomes."""
    for i, K in enumerate(port.cb_sizes()):
        K = int(K)
        llr, _ = coded_llrs(port, K, 2, 0.85 + 0.3 * (i % 3), 16, 31, seed=i)
        o1, k1, n1, _ = port.decode_batch(llr, K, 4, "B", 0, True)
        o2, k2, nc, nr = run_emu(emu, llr, K, 4, 0, True)
        assert (o1 == o2).all() and (k1 == k2).all() and (n1 == npass_of(k2, nc, nr)).all(), K


def test_emulation_crc24a_and_no_crc(emu, port):
    K = 1024
    llr, _ = coded_llrs(port, K, 4, 0.8, 16, 31, seed=3, crc="A")
    o1, k1, n1, _ = port.decode_batch(llr, K, 6, "A", K, True)
    o2, k2, nc, nr = run_emu(emu, llr, K, 6, 1, True)
    assert (o1 == o2).all() and (k1 == k2).all() and (n1 == npass_of(k2, nc, nr)).all()
    assert k1.all()
    o1, k1, n1, _ = port.decode_batch(llr, K, 3, None, 0, True)
    o2, k2, nc, nr = run_emu(emu, llr, K, 3, 2, True)
    assert (o1 == o2).all() and not k2.any() and (nr == 3).all()


def test_emulation_mixed_int8_and_int16_tiles(emu, port):
    """Per-tile input format (int8 when every channel LLR of the tile fits, else int16) inside one batch."""
    K = 512
    llr, _ = coded_llrs(port, K, 64 * 3 + 5, 0.9, 16, 31, seed=21)
    llr[:64] = np.clip(llr[:64].astype(np.int32) * 4, -128, 127).astype(np.int16)
    llr[64:128] = (llr[64:128].astype(np.int32) * 5).astype(np.int16)
    llr[130, 3 * K + 1] = 128
    llr[190, 3 * K + 6] = 3000  # encoder 2's systematic tail is kept in int16 regardless
    o1, k1, n1, _ = port.decode_batch(llr, K, 5, "B", 0, True)
    o2, k2, nc, nr = run_emu(emu, llr, K, 5, 0, True)
    assert (o1 == o2).all() and (k1 == k2).all() and (n1 == npass_of(k2, nc, nr)).all()
