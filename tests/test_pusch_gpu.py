"""Identity-channel PUSCH receive pipeline (BASELINE config 4): OFDM rx -> soft demap -> de-match -> turbo decode on the
GPU, against the transmitted payload and, stage by stage, against the oracle.  -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
SB = 18600


@pytest.fixture(scope="module")
def subframes():
    from srslte_b200 import synth_pusch as sp

    iq, payload, G = sp.make_subframes(100, 2048, 75376, 6, 0, sp.qpp_interleaver(5824), 3, snr_db=23.0, seed=11)
    return iq, payload, G


def test_config4_pipeline_decodes_and_matches_oracle_stage_by_stage(port, subframes):
    import torch

    from srslte_b200.pusch import DATA_SYMBOL_MASK, PuschRx

    iq, payload, G = subframes
    nsf, tbs = iq.shape[0], 75376
    rx = PuschRx(100, tbs, 3, llr_shift=4, max_noi=8, symbol_sz=2048)
    x = torch.from_numpy(iq).cuda()
    ok, its = rx.run(x, nsf)
    torch.cuda.synchronize()
    data = rx.data[:nsf, :tbs // 8 + 3].cpu().numpy()
    assert ok.all()
    assert (data == payload).all()

    # stage 1: OFDM grid vs the oracle's receiver (float: 1e-4 relative L2, BASELINE north_star)
    grid = rx.grid[:nsf].cpu().numpy()
    want, _ = port.ofdm_rx(iq.reshape(-1), 100, False, 2048, -0.5, 0.5, False, False)
    assert np.linalg.norm(grid - want) / np.linalg.norm(want) < 1e-4
    # stage 2: soft bits from the GPU's own grid are bit-exact with the reference demapper (one call per subframe, pusch.c:449)
    llr = rx.llr[:nsf].cpu().numpy()
    data_syms = [l for l in range(14) if (DATA_SYMBOL_MASK >> l) & 1]
    for s in range(nsf):
        q = port.demod_s(3, grid[s, data_syms, :].reshape(-1))
        assert ((q >> 4) == llr[s]).all()
    # stage 3: decode_tb of the oracle on the same soft bits: same bytes, verdict and pass count
    for s in range(nsf):
        soft = np.zeros(13 * SB, np.int16)
        cbcrc = np.zeros(13, np.uint8)
        d = np.zeros(tbs // 8 + 3 + 768, np.uint8)
        ret, iters = port.decode_tb(llr[s], tbs, 6, 0, 8, soft, cbcrc, d)
        assert ret == 0 and (d[:tbs // 8 + 3] == data[s]).all()
        assert abs(iters / 13 - its[s]) < 1e-6
    rx.close()


def test_config4_pipeline_fails_cleanly_when_too_noisy(subframes):
    import torch

    from srslte_b200.pusch import PuschRx

    iq, _, _ = subframes
    rng = np.random.default_rng(0)
    noisy = iq + ((rng.normal(size=iq.shape) + 1j * rng.normal(size=iq.shape)) * 0.01).astype(np.complex64)
    rx = PuschRx(100, 75376, 3, llr_shift=4, max_noi=4, symbol_sz=2048)
    ok, its = rx.run(torch.from_numpy(noisy).cuda(), iq.shape[0])
    assert not ok.any() and (its == 4).all()
    rx.close()


@pytest.mark.parametrize("mod,shift", [(1, 0), (2, 3), (3, 0)])
def test_demap_batch_other_modulations(port, mod, shift):
    import ctypes as C

    import torch

    from srslte_b200 import _lib
    from srslte_b200.pusch import DATA_SYMBOL_MASK

    L = _lib.lib()
    L.srsran_b200_pusch_demap_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    rng = np.random.default_rng(mod)
    nsf, nre = 3, 73  # odd sizes: the reference's SIMD body / scalar tail split must follow the per-subframe call length
    grid = (rng.normal(size=(nsf, 14, nre)) + 1j * rng.normal(size=(nsf, 14, nre))).astype(np.complex64) * 0.8
    g = torch.from_numpy(grid).cuda()
    out = torch.zeros((nsf, 12 * nre * 2 * mod), dtype=torch.int16, device="cuda")
    assert L.srsran_b200_pusch_demap_batch(0, mod, g.data_ptr(), out.data_ptr(), nsf, 14, nre, DATA_SYMBOL_MASK, shift,
                                           _lib.FLAG_DEVICE_PTRS, None) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    data_syms = [l for l in range(14) if (DATA_SYMBOL_MASK >> l) & 1]
    for s in range(nsf):
        assert ((port.demod_s(mod, grid[s, data_syms, :].reshape(-1)) >> shift) == got[s]).all()
    assert L.srsran_b200_pusch_demap_batch(0, mod, g.data_ptr(), out.data_ptr(), nsf, 14, nre, DATA_SYMBOL_MASK, shift, 0, None) == -2
