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


def run_emu_mixed(emu, llrs, Ks, max_pass, crc_kind, early, compact=0, force16=False, split=51):
    """llrs: list of (ncb_g, 3K_g+12) arrays.  Returns per-group (out, ok, npass) plus the number of lanes moved."""
    ncbs = [a.shape[0] for a in llrs]
    flat = np.ascontiguousarray(np.concatenate([a.ravel() for a in llrs]).astype(np.int16))
    tot = sum(ncbs)
    out = np.zeros(sum(n * K // 8 for n, K in zip(ncbs, Ks)), np.uint8)
    ok = np.zeros(tot, np.uint8)
    nc = np.zeros(tot, np.uint8)
    nr = np.zeros(tot, np.uint8)
    moved = C.c_uint32(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    Ka = np.array(Ks, np.uint32)
    na = np.array(ncbs, np.uint32)
    assert emu.emu_tdec_mixed(p(flat), len(Ks), p(Ka), p(na), max_pass, crc_kind, int(early), int(force16), split, compact,
                              p(out), p(ok), p(nc), p(nr), C.byref(moved)) == 0
    res, o0, c0 = [], 0, 0
    for n, K in zip(ncbs, Ks):
        res.append((out[o0:o0 + n * K // 8].reshape(n, K // 8), ok[c0:c0 + n], npass_of(ok[c0:c0 + n], nc[c0:c0 + n], nr[c0:c0 + n])))
        o0 += n * K // 8
        c0 += n
    return res, moved.value


def test_emulation_mixed_sizes_in_one_batch(emu, port):
    """BASELINE config 3: several code block lengths in ONE batch (tiles carry their own K, ordered by length)."""
    Ks = [512, 40, 6144, 1024, 48, 2048]
    ncbs = [70, 3, 2, 65, 130, 1]
    llrs = [coded_llrs(port, K, n, 0.9, 16, 31, seed=K)[0] for K, n in zip(Ks, ncbs)]
    for early in (True, False):
        res, _ = run_emu_mixed(emu, llrs, Ks, 5, 0, early)
        for (o2, k2, n2), llr, K in zip(res, llrs, Ks):
            o1, k1, n1, _ = port.decode_batch(llr, K, 5, "B", 0, early)
            assert (o1 == o2).all() and (k1 == k2).all() and (n1 == n2).all(), (K, early)


@pytest.mark.parametrize("K,ntile,compact", [(40, 9, 1), (104, 6, 1), (40, 12, 3)])
def test_emulation_lane_repacking_between_passes(emu, port, K, ntile, compact):
    """Block-granular early stop: the lanes that still run are re-packed into fewer tiles after every pass (compact_* of
    tdec_core.h).  Blocks with very different convergence share tiles, so lanes really move; the result must be what
    the oracle (and the unpacked run) gives for every block."""
    ncb = 64 * ntile - 5
    rng = np.random.default_rng(K)
    parts = []
    for i in range(ncb):  # one block in ~six is hard (stays for many passes), the rest converge quickly
        sigma = 1.25 if rng.random() < 0.17 else 0.55
        parts.append(coded_llrs(port, K, 1, sigma, 16, 31, seed=1000 * K + i)[0])
    llr = np.concatenate(parts)
    # a second group of another length in the same batch: lanes never cross groups
    llr2, _ = coded_llrs(port, 48, 150, 0.9, 16, 31, seed=5)
    o1, k1, n1, _ = port.decode_batch(llr, K, 8, "B", 0, True)
    p1, q1, r1, _ = port.decode_batch(llr2, 48, 8, "B", 0, True)
    (a, b), moved = run_emu_mixed(emu, [llr, llr2], [K, 48], 8, 0, True, compact=compact)
    assert moved > 0
    assert (a[0] == o1).all() and (a[1] == k1).all() and (a[2] == n1).all()
    assert (b[0] == p1).all() and (b[1] == q1).all() and (b[2] == r1).all()
    assert 1 < n1.min() + 0 < n1.max() or n1.max() > n1.min()  # mixed convergence, otherwise nothing was exercised
