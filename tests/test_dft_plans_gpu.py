"""srsran_dft_plan_* / srsran_dft_run_* of this library (the reference-named DFT plan API, lib/src/phy/dft/dft_fftw.c) against the
REFERENCE's own dft_fftw.c (oracle/_ref/libsrsref.so, compiled over the float64 FFTW shim): every knob the plan carries --
direction, mirror (both directions), dc, norm, dB, re-planning to a smaller size, and guru plans with strides between
transforms (the form the reference's OFDM transmitter uses).  Both libraries are loaded side by side with ctypes; the same
plan structure (dft.h:54-68) is handed to each.  -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Plan(C.Structure):
    _fields_ = [("init_size", C.c_int), ("size", C.c_int), ("in_", C.c_void_p), ("out", C.c_void_p), ("p", C.c_void_p),
                ("is_guru", C.c_bool), ("forward", C.c_bool), ("mirror", C.c_bool), ("db", C.c_bool), ("norm", C.c_bool), ("dc", C.c_bool),
                ("dir", C.c_int), ("mode", C.c_int)]


FWD, BWD = 0, 1


@pytest.fixture(scope="module")
def libs(ref):
    from oracle import loader
    from srslte_b200.build import LIB_PATH

    ours = C.CDLL(LIB_PATH)
    theirs = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsrsref.so"))
    for L in (ours, theirs):
        L.srsran_dft_plan_c.argtypes = [C.POINTER(Plan), C.c_int, C.c_int]
        L.srsran_dft_plan_guru_c.argtypes = [C.POINTER(Plan), C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.srsran_dft_replan.argtypes = [C.POINTER(Plan), C.c_int]
        L.srsran_dft_plan_free.argtypes = [C.POINTER(Plan)]
        L.srsran_dft_plan_free.restype = None
        for f in ("mirror", "db", "norm", "dc"):
            fn = getattr(L, "srsran_dft_plan_set_" + f)
            fn.argtypes = [C.POINTER(Plan), C.c_bool]
            fn.restype = None
        L.srsran_dft_run_c.argtypes = [C.POINTER(Plan), C.c_void_p, C.c_void_p]
        L.srsran_dft_run_c.restype = None
        L.srsran_dft_run_c_zerocopy.argtypes = [C.POINTER(Plan), C.c_void_p, C.c_void_p]
        L.srsran_dft_run_c_zerocopy.restype = None
        L.srsran_dft_run_guru_c.argtypes = [C.POINTER(Plan)]
        L.srsran_dft_run_guru_c.restype = None
    return ours, theirs


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def run_c(L, N, direction, x, mirror=False, dc=False, norm=False, db=False, replan_to=None):
    p = Plan()
    assert L.srsran_dft_plan_c(C.byref(p), N, direction) == 0
    L.srsran_dft_plan_set_mirror(C.byref(p), mirror)
    L.srsran_dft_plan_set_dc(C.byref(p), dc)
    L.srsran_dft_plan_set_norm(C.byref(p), norm)
    L.srsran_dft_plan_set_db(C.byref(p), db)
    n = N
    if replan_to is not None:
        assert L.srsran_dft_replan(C.byref(p), replan_to) == 0
        n = replan_to
    xin = np.ascontiguousarray(x[:n], np.complex64)
    out = np.zeros(n, np.complex64)
    L.srsran_dft_run_c(C.byref(p), xin.ctypes.data, out.ctypes.data)
    L.srsran_dft_plan_free(C.byref(p))
    return out


@pytest.mark.parametrize("N", [128, 1536, 2048, 1200, 72])
def test_plan_knobs_match_the_reference(libs, N):
    ours, theirs = libs
    rng = np.random.default_rng(N)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    for direction in (FWD, BWD):
        for mirror in (False, True):
            for dc in (False, True):
                for norm in (False, True):
                    a = run_c(ours, N, direction, x, mirror, dc, norm)
                    b = run_c(theirs, N, direction, x, mirror, dc, norm)
                    assert rel(a, b) < 1e-4, (N, direction, mirror, dc, norm, rel(a, b))
    # dB output: 10 log10 of the REAL part (the reference hands a complex value to a float function, dft_fftw.c:346-349):
    # compared where the real part is comfortably positive, NaN where it is negative in both
    a = run_c(ours, N, FWD, x, db=True, norm=True)
    b = run_c(theirs, N, FWD, x, db=True, norm=True)
    lin = run_c(theirs, N, FWD, x, norm=True).real
    good = lin > 0.05
    assert good.sum() > N // 8 and np.allclose(a.real[good], b.real[good], atol=2e-3)
    assert (np.isnan(a.real) == np.isnan(b.real)).all() or (np.isnan(a.real[lin < -0.05])).all()
    assert (a.imag[good] == 0).all()


def test_replan_to_a_smaller_size(libs):
    ours, theirs = libs
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(2048) + 1j * rng.standard_normal(2048)).astype(np.complex64)
    for new in (1536, 1024, 600, 128):
        a = run_c(ours, 2048, FWD, x, replan_to=new)
        b = run_c(theirs, 2048, FWD, x, replan_to=new)
        assert rel(a, b) < 1e-4, new
    p = Plan()
    assert ours.srsran_dft_plan_c(C.byref(p), 512, FWD) == 0
    assert ours.srsran_dft_replan(C.byref(p), 1024) == -1   # larger than the size it was created with (dft_fftw.c:92-104)
    ours.srsran_dft_plan_free(C.byref(p))


@pytest.mark.parametrize("N,how_many,idist,odist,direction", [(2048, 7, 2048 + 144, 2048, FWD), (2048, 7, 2048, 2048 + 144, BWD),
                                                              (128, 6, 128 + 32, 128, FWD), (1536, 3, 1536, 1536, BWD)])
def test_guru_plans_with_gaps_between_transforms(libs, N, how_many, idist, odist, direction):
    """srsran_dft_plan_guru_c / srsran_dft_run_guru_c: how_many transforms idist / odist apart on caller-owned buffers -- the
    OFDM receiver's input layout (symbols separated by cyclic prefixes) and the transmitter's output layout (room left for
    them).  What lies between the transforms in the output buffer must be left alone."""
    ours, theirs = libs
    rng = np.random.default_rng(N + how_many)
    nin, nout = (how_many - 1) * idist + N, (how_many - 1) * odist + N
    x = (rng.standard_normal(nin) + 1j * rng.standard_normal(nin)).astype(np.complex64)
    outs = []
    for L in (ours, theirs):
        xin = x.copy()
        out = np.full(nout, 7 - 3j, np.complex64)
        p = Plan()
        assert L.srsran_dft_plan_guru_c(C.byref(p), N, direction, xin.ctypes.data, out.ctypes.data, 1, 1, how_many, idist, odist) == 0
        L.srsran_dft_run_guru_c(C.byref(p))
        L.srsran_dft_plan_free(C.byref(p))
        outs.append(out)
    assert rel(outs[0], outs[1]) < 1e-4
    if odist > N:
        gaps = np.ones(nout, bool)
        for i in range(how_many):
            gaps[i * odist:i * odist + N] = False
        assert (outs[0][gaps] == 7 - 3j).all()


def test_zerocopy_run(libs):
    ours, theirs = libs
    rng = np.random.default_rng(9)
    x = (rng.standard_normal(600) + 1j * rng.standard_normal(600)).astype(np.complex64)
    outs = []
    for L in (ours, theirs):
        p = Plan()
        assert L.srsran_dft_plan_c(C.byref(p), 600, BWD) == 0
        out = np.zeros(600, np.complex64)
        L.srsran_dft_run_c_zerocopy(C.byref(p), x.ctypes.data, out.ctypes.data)
        L.srsran_dft_plan_free(C.byref(p))
        outs.append(out)
    assert rel(outs[0], outs[1]) < 1e-4
