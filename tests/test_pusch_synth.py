"""The numpy transmit chain used to synthesise PUSCH subframes (srslte_b200/synth_pusch.py, bench/tests only) against the
oracle's encoder, rate matcher and CRC, and against the oracle's OFDM receiver (loop-back)."""
import numpy as np

from srslte_b200 import synth_pusch as sp


def test_crc_and_qpp(port):
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, (3, 6120)).astype(np.uint8)
    for poly, kind in ((sp.CRC24A, "A"), (sp.CRC24B, "B")):
        par = sp.crc24(bits, poly)
        for r in range(3):
            want = port.crc24(kind, np.packbits(bits[r]), 6120)
            assert sum(int(b) << (23 - i) for i, b in enumerate(par[r])) == want
    for K in (40, 1024, 5824, 6144):
        fwd, _ = port.interleaver(K)
        assert (sp.qpp_interleaver(K) == fwd).all()


def test_turbo_encoder_and_rate_matching(port):
    rng = np.random.default_rng(2)
    for K in (40, 504, 5824):
        bits = rng.integers(0, 2, (2, K)).astype(np.uint8)
        d = sp.turbo_encode(bits, sp.qpp_interleaver(K))
        for r in range(2):
            cw = port.tcod_encode(bits[r])  # natural order [3k+j], k = 0..K+3
            assert (d[r].T.reshape(-1) == cw).all(), K
            for rv in range(4):
                for E in (int(0.4 * 3 * K), 3 * K + 12, int(1.7 * 3 * K)):
                    assert (sp.rate_match(d[r:r + 1], E, rv)[0] == port.rm_tx(cw, K, E, rv)).all(), (K, rv, E)


def test_transport_block_matches_helper_chain(port):
    """Same segmentation / E split as the oracle-built chain of tests/helpers.make_tb (sch.c:240-350 restated)."""
    tbs, Qm, G = 75376, 6, 86400
    assert sp.segment(tbs) == (13, 5824)
    f, payload = sp.make_transport_blocks(tbs, Qm, G, 0, sp.qpp_interleaver(5824), 1, seed=3)
    s = port.cbsegm(tbs)
    tb = np.unpackbits(payload[0])[:tbs + 24]
    C, K = s["C"], s["K1"]
    Gp = G // Qm
    gamma = Gp % C
    tx = []
    for c in range(C):
        bits = tb[c * (K - 24):(c + 1) * (K - 24)]
        r = port.crc24("B", np.packbits(bits), K - 24)
        bits = np.concatenate([bits, np.array([(r >> (23 - i)) & 1 for i in range(24)], np.uint8)])
        n_e = Qm * (Gp // C) if c <= C - gamma - 1 else Qm * -(-Gp // C)
        tx.append(port.rm_tx(port.tcod_encode(bits), K, n_e, 0))
    assert (np.concatenate(tx) == f[0]).all()
    assert port.crc24("A", payload[0], tbs + 24) == 0


def test_ofdm_loopback_through_oracle_receiver(port):
    """ofdm_modulate is the inverse of the eNB uplink receive configuration (shift -0.5, window offset 0.5, no normalisation)."""
    rng = np.random.default_rng(4)
    for prb, N in ((6, 128), (100, 2048)):
        R = 12 * prb
        grid = (rng.normal(size=(2, 14, R)) + 1j * rng.normal(size=(2, 14, R))).astype(np.complex64)
        x = sp.ofdm_modulate(grid, N)
        got, _ = port.ofdm_rx(x.reshape(-1), prb, False, N, -0.5, 0.5, False, False)
        assert np.linalg.norm(got - grid) / np.linalg.norm(grid) < 1e-4


def test_soft_bits_have_the_right_sign(port):
    """64QAM/16QAM/QPSK mapper vs the reference's soft demapper convention (LLR > 0 <=> bit 1)."""
    rng = np.random.default_rng(5)
    for mod, Qm in ((1, 2), (2, 4), (3, 6)):
        bits = rng.integers(0, 2, 240 * Qm).astype(np.uint8)
        sym = sp._qam(bits, Qm).astype(np.complex64)
        llr = port.demod_s(mod, sym)
        assert ((llr > 0).astype(np.uint8) == bits).all(), mod


def test_full_pusch_transmitter_matches_the_reference(ref):
    """The numpy transmitter used by bench.py's full-chain leg produces the resource grid of srsran_pusch_encode + DMRS."""
    from oracle import loader
    from srslte_b200 import synth_pusch as sp

    for kw, tbs, K in ((dict(cell_id=1, nof_prb=100, L_prb=100, n_prb=0, mod=3), 75376, 5824),
                       (dict(cell_id=150, nof_prb=50, L_prb=24, n_prb=13, mod=2, cyclic_shift=3), 9912, 4992)):
        rnti = np.array([62, 40000], np.uint32)
        tti = np.array([3, 18], np.uint32)
        links = [loader.pusch_link(rnti=int(r), tti=int(t), tbs=tbs, **kw) for r, t in zip(rnti, tti)]
        dm = {int(t % 10): ref.dmrs_pusch_gen(lk).reshape(2, -1) for t, lk in zip(tti, links)}
        grid, payload = sp.make_pusch_grids(kw["cell_id"], kw["nof_prb"], kw["L_prb"], kw["n_prb"], tbs, 2 * kw["mod"], 0,
                                            sp.qpp_interleaver(K), 2, rnti, tti, lambda sf: dm[sf], seed=5)
        for s in range(2):
            want = ref.pusch_encode(links[s], payload[s, :tbs // 8])
            assert np.linalg.norm(grid[s] - want) / np.linalg.norm(want) < 1e-5


def test_bench_ofdm_cpu_substitute_demodulates_what_the_synthesiser_sends():
    """bench.py's labelled stand-in for the CPU speed of srsran_ofdm_rx_sf (FFTW is not available here) must at least
    be the same arithmetic: eNB uplink settings, N = 2048, 100 PRB."""
    import os
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from srslte_b200 import synth_pusch as sp

    rng = np.random.default_rng(0)
    g = (rng.standard_normal((3, 14, 1200)) + 1j * rng.standard_normal((3, 14, 1200))).astype(np.complex64)
    out = bench.ofdm_cpu_substitute(sp.ofdm_modulate(g, 2048), 2, True)
    assert np.abs(out - g).max() < 1e-5 * np.abs(g).max()
