"""OFDM receive and soft demapper on the GPU against the oracle / golden vectors, through the C ABI.  -m gpu."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-4  # BASELINE.json north_star: OFDM outputs within 1e-4 relative L2 of the FFTW-backed reference


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_golden_vectors_from_reference_ofdm_c():
    """Outputs recorded from the reference's ofdm.c + dft_fftw.c (over the float64 DFT shim), tools/gen_golden.py."""
    from srslte_b200.ofdm import OfdmRx

    g = np.load(os.path.join(GOLD, "ofdm_demod.npz"))
    i = 0
    while f"cfg{i}" in g:
        prb, N, cp, fs, wo, nm, kd = g[f"cfg{i}"]
        q = OfdmRx(int(prb), bool(cp), int(N), float(fs), float(wo), bool(nm), bool(kd))
        y = q.rx_sf(g[f"in{i}"])
        q.close()
        assert rel(y, g[f"out{i}"]) < TOL, (i, rel(y, g[f"out{i}"]))
        i += 1
    assert i >= 6


@pytest.mark.parametrize("prb,N", [(6, 0), (15, 0), (25, 0), (50, 0), (75, 0), (100, 0), (6, 128), (15, 256), (25, 512), (50, 1024),
                                   (75, 1536), (100, 2048), (110, 2048), (100, 4096)])
@pytest.mark.parametrize("variant", ["plain", "ul", "extcp_norm", "shift_only_keepdc"])
def test_every_symbol_size_vs_oracle(port, prb, N, variant):
    """All sizes the reference can request (phy_common.c:342-385: 128..2048 incl. 384/768/1536) plus the forced 4096 of
    ofdm_test -N 4096, in the variants dft/test/CMakeLists.txt:28-33 registers: normal/extended CP, -s 0.5 half-subcarrier
    shift, -o 0.5 window offset; 3 subframes each."""
    from srslte_b200.ofdm import OfdmRx

    cp, fs, wo, nm, kd = {"plain": (0, 0.0, 0.0, 0, 0), "ul": (0, -0.5, 0.5, 0, 0), "extcp_norm": (1, 0.0, 0.25, 1, 0),
                          "shift_only_keepdc": (0, 0.5, 0.0, 0, 1)}[variant]
    n = N or port.symbol_sz(prb)
    rng = np.random.default_rng(prb * 7 + n)
    x = (rng.normal(size=3 * 15 * n) + 1j * rng.normal(size=3 * 15 * n)).astype(np.complex64)
    want, _ = port.ofdm_rx(x, prb, bool(cp), N, fs, wo, bool(nm), bool(kd))
    q = OfdmRx(prb, bool(cp), N, fs, wo, bool(nm), bool(kd))
    assert q.symbol_sz == n
    got = q.rx_sf(x)
    q.close()
    assert got.shape == want.shape
    assert rel(got, want) < TOL, rel(got, want)


def test_loopback_like_ofdm_test(port):
    """ofdm_test.c:120-179: Tx -> Rx loop-back must return the transmitted grid (RMS error < 1e-4).  The time-domain signal
    is synthesised here as the exact inverse of the receive definition (normalised, no shift)."""
    from srslte_b200.ofdm import OfdmRx

    prb, N = 25, 512
    R = 12 * prb
    rng = np.random.default_rng(3)
    grid = ((rng.integers(0, 2, (14, R)) * 2 - 1) + 1j * (rng.integers(0, 2, (14, R)) * 2 - 1)).astype(np.complex64) / np.sqrt(2)
    cp1, cp2 = int(np.ceil(160 * N / 2048)), int(np.ceil(144 * N / 2048))
    sf = []
    for l in range(14):
        X = np.zeros(N, np.complex128)
        X[N - R // 2:] = grid[l, :R // 2]
        X[1:1 + R // 2] = grid[l, R // 2:]
        t = np.fft.ifft(X) * np.sqrt(N)
        cp = cp1 if l % 7 == 0 else cp2
        sf.append(np.concatenate([t[-cp:], t]))
    x = np.concatenate(sf).astype(np.complex64)
    assert x.size == 15 * N
    q = OfdmRx(prb, False, N, 0.0, 0.0, True, False)
    y = q.rx_sf(x)[0]
    q.close()
    assert np.sqrt(np.mean(np.abs(y - grid) ** 2)) < 1e-4


def test_invalid_configurations():
    from srslte_b200.ofdm import OfdmRx

    with pytest.raises(RuntimeError):
        OfdmRx(0)  # ofdm.c:41-45 "Invalid number of PRB"
    with pytest.raises(RuntimeError):
        OfdmRx(111)
    with pytest.raises(RuntimeError):
        OfdmRx(6, symbol_sz=1001)  # 7 x 11 x 13: not 2^a 3^b 5^c


def test_demod_bit_exact(port):
    from srslte_b200.ofdm import demod_soft_s

    g = np.load(os.path.join(GOLD, "ofdm_demod.npz"))
    assert (demod_soft_s(1, g["qpsk_sym"]) == g["qpsk_llr"]).all()
    assert (demod_soft_s(2, g["qam16_sym"]) == g["qam16_llr"]).all()
    assert (demod_soft_s(3, g["qam64_sym"]) == g["qam64_llr"]).all()
    rng = np.random.default_rng(11)
    for mod in (1, 2, 3):
        for n in (1, 4, 7, 9, 16, 1000, 14401):
            s = ((rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.8).astype(np.complex64)
            s[:1] = 47.0 - 46.9j  # saturates the int16 conversion
            assert (demod_soft_s(mod, s) == port.demod_s(mod, s)).all(), (mod, n)
    # batch standing for several reference calls: the vector-body / scalar-tail split is per call
    s = ((rng.normal(size=3 * 1001) + 1j * rng.normal(size=3 * 1001)) * 0.8).astype(np.complex64)
    for mod in (1, 2, 3):
        want = np.concatenate([port.demod_s(mod, s[i * 1001:(i + 1) * 1001]) for i in range(3)])
        assert (demod_soft_s(mod, s, symbols_per_call=1001) == want).all()


@pytest.mark.parametrize("prb,N", [(100, 2048), (100, 0), (25, 512), (6, 128), (75, 0)])
def test_int16_iq_input_equals_float_input(prb, N):
    """SRSRAN_B200_FLAG_IQ_INT16: the radio's int16 I/Q pairs give bit for bit what the float entry gives on x / 32768
    (the conversion is exact), on the specialised and the generic kernel, host and device pointers."""
    import ctypes as C

    import torch

    from srslte_b200 import _lib
    from srslte_b200.ofdm import OfdmRx

    rx = OfdmRx(prb, False, N, -0.5, 0.5, False, False)
    rng = np.random.default_rng(prb + N)
    nsf = 3
    iq = rng.integers(-20000, 20000, (nsf, rx.sf_sz, 2)).astype(np.int16)
    iq[0, :4] = [[-32768, 32767], [0, -1], [1, 0], [32767, -32768]]
    xf = (iq[..., 0].astype(np.float32) / np.float32(32768.0) + 1j * (iq[..., 1].astype(np.float32) / np.float32(32768.0))).astype(np.complex64)
    want = rx.rx_sf(xf.reshape(-1))
    out = np.zeros_like(want)
    rc = rx._lib.srsran_b200_ofdm_rx_sf_batch(rx._h, iq.ctypes.data, out.ctypes.data, nsf, _lib.FLAG_IQ_INT16, None)
    assert rc == 0
    assert (out.view(np.uint32) == want.view(np.uint32)).all()
    d_out = torch.empty((nsf, rx.nof_symbols, rx.nof_re), dtype=torch.complex64, device="cuda")
    rx.rx_sf_device(torch.from_numpy(iq).cuda(), d_out, nsf)
    torch.cuda.synchronize()
    assert (d_out.cpu().numpy().view(np.uint32) == want.view(np.uint32)).all()
    rx.close()
