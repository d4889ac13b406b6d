"""Pins the oracle port to the reference itself: every function of oracle/oracle_port.c against the same call into
oracle/_ref/libsrsref.so (the reference's unmodified sources, compiled in place by oracle/Makefile)."""
import numpy as np
import pytest

from helpers import coded_llrs


def test_tables(port, ref):
    assert (port.cb_sizes() == ref.cb_sizes()).all()
    for K in (40, 48, 104, 512, 528, 1024, 2048, 5824, 6144):
        a, b = port.interleaver(K)
        c, d = ref.interleaver(K)
        assert (a == c).all() and (b == d).all()
    for K in (1, 39, 40, 41, 513, 6144, 6145):
        assert port.cbindex(K) == ref.cbindex(K)
    for tbs in list(range(16, 6300, 97)) + [6120, 6121, 6144, 12960, 36696, 75376, 97896]:
        assert port.cbsegm(tbs) == ref.cbsegm(tbs), tbs


def test_rm_tables_all(port, ref):
    for i in range(188):
        for rv in range(4):
            assert (port.rm_table(i, rv) == ref.rm_table(i, rv)).all(), (i, rv)


def test_rm_matches_float_spec_implementation(port, ref):
    """rm_turbo_test.c:172-188: the LUT path must equal the float spec implementation srsran_rm_turbo_rx exactly."""
    import ctypes as C

    rng = np.random.default_rng(5)
    for cb_idx in (0, 7, 59, 100, 187):
        K = int(port.cb_sizes()[cb_idx])
        n = 3 * K + 12
        for rv in range(4):
            for E in (n // 2, n, n + 500):
                e = rng.integers(-50, 50, E).astype(np.int16)
                soft = np.zeros(n + 64, np.int16)
                port.rm_rx(e, soft, cb_idx, rv)
                ef = e.astype(np.float32)
                of = np.zeros(n, np.float32)
                assert ref.lib.ref_rm_rx_float(ef.ctypes.data_as(C.c_void_p), C.c_uint32(E), of.ctypes.data_as(C.c_void_p),
                                               C.c_uint32(n), C.c_uint32(rv)) == 0
                assert (soft[:n] == of.astype(np.int16)).all(), (cb_idx, rv, E)


@pytest.mark.parametrize("K", [40, 504, 1024, 6144])
def test_encode_rm_decode(port, ref, K):
    rng = np.random.default_rng(K)
    bits = rng.integers(0, 2, K).astype(np.uint8)
    cw = port.tcod_encode(bits)
    assert (cw == ref.tcod_encode(bits)).all()
    n = 3 * K + 12
    for rv in range(4):
        for E in (int(0.4 * n), n, int(1.7 * n)):
            assert (port.rm_tx(cw, K, E, rv) == ref.rm_tx(cw, K, E, rv)).all()
            e = rng.integers(-40, 40, E).astype(np.int16)
            s1 = rng.integers(-100, 100, n + 64).astype(np.int16)
            s2 = s1.copy()
            port.rm_rx(e, s1, port.cbindex(K), rv)
            ref.rm_rx(e, s2, port.cbindex(K), rv)
            assert (s1 == s2).all()


@pytest.mark.parametrize("K,sigma,scale,clip", [(40, 0.8, 16, 31), (504, 0.95, 16, 31), (1024, 1.0, 16, 31), (6144, 0.93, 16, 31),
                                                (6144, 1.3, 32, 63), (2048, 0.8, 500, 2000), (1024, 3.0, 8000, 30000)])
def test_generic_decoder_bit_exact_including_wraparound(port, ref, K, sigma, scale, clip):
    """Every pass's decision equals the reference's generic int16 decoder, also where its arithmetic wraps (|LLR| >= 100,
    SURVEY.md section 0.2)."""
    llr, _ = coded_llrs(port, K, 3, sigma, scale, clip, seed=K + clip)
    for c in range(3):
        assert (port.tdec_passes(llr[c], K, 8) == ref.tdec_passes(llr[c], K, 8)).all()
    for es in (True, False):
        a = port.decode_batch(llr, K, 8, "B", 0, es)
        b = ref.decode_batch(llr, K, 8, "B", 0, es)
        for x, y in zip(a[:3], b[:3]):
            assert (x == y).all()


def test_crc(port, ref):
    rng = np.random.default_rng(9)
    d = rng.integers(0, 256, 4096).astype(np.uint8)
    for n in (8, 16, 24, 32, 6144, 32768):
        for k in "AB":
            assert port.crc24(k, d, n) == ref.crc24(k, d, n)


@pytest.mark.parametrize("prb,N,cp,fs,wo,nm,kd", [(6, 0, 0, 0.0, 0.0, 0, 0), (25, 0, 0, -0.5, 0.5, 0, 0), (100, 2048, 0, -0.5, 0.5, 0, 0),
                                                  (50, 0, 1, 0.0, 0.0, 1, 0), (100, 0, 0, 0.5, 0.0, 0, 1), (15, 0, 0, 0.0, 0.3, 0, 0),
                                                  (75, 1536, 1, -0.5, 0.5, 1, 0), (6, 4096, 0, 0.0, 0.0, 0, 0)])
def test_ofdm_rx(port, ref, prb, N, cp, fs, wo, nm, kd):
    """ofdm_test.c:170-179 accepts < 1e-4; same bound between the float64 restatement and the reference's ofdm.c."""
    n = N or port.symbol_sz(prb)
    rng = np.random.default_rng(prb + n)
    x = (rng.normal(size=2 * 15 * n) + 1j * rng.normal(size=2 * 15 * n)).astype(np.complex64)
    a, _ = port.ofdm_rx(x, prb, bool(cp), N, fs, wo, bool(nm), bool(kd))
    b, _ = ref.ofdm_rx(x, prb, bool(cp), N, fs, wo, bool(nm), bool(kd))
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-4


def test_demod(port, ref):
    rng = np.random.default_rng(11)
    for mod in (1, 2, 3):
        for n in (1, 4, 7, 9, 16, 1000, 14401):
            s = ((rng.normal(size=n) + 1j * rng.normal(size=n)) * 0.8).astype(np.complex64)
            s[:1] = 47.0 - 46.9j
            assert (port.demod_s(mod, s) == ref.demod_s(mod, s)).all()


# ---- PUSCH chain between OFDM and de-matching ------------------------------------------------
def test_pusch_descrambling_and_deinterleave(port, ref):
    rng = np.random.default_rng(31)
    x = rng.integers(-32768, 32768, 86400).astype(np.int16)
    x[:5] = -32768
    for rnti, ns, cid in ((62, 6, 1), (0xFFFF, 18, 503), (1, 0, 0), (4660, 9, 301)):
        for n in (86400, 23, 24, 1, 4801):
            assert (ref.pusch_seq_apply_s(x[:n], rnti, ns, cid) == port.pusch_seq_apply_s(x[:n], rnti, ns, cid)).all()
    for Qm, nre, nsym in ((6, 14400, 12), (2, 12 * 36, 12), (4, 10 * 12 * 5, 10), (6, 12 * 1200, 12)):
        q = rng.integers(-3000, 3000, nre * Qm).astype(np.int16)
        assert (ref.ulsch_deinterleave(q, Qm, nsym) == port.ulsch_deinterleave(q, Qm, nsym)).all()


@pytest.mark.parametrize("L", [1, 2, 3, 5, 6, 25, 27, 60, 81, 100])
def test_dft_precoding(port, ref, L):
    rng = np.random.default_rng(L)
    z = (rng.standard_normal(12 * 12 * L) + 1j * rng.standard_normal(12 * 12 * L)).astype(np.complex64)
    for tx in (False, True):
        a, b = ref.dft_precoding(z, L, tx), port.dft_precoding(z, L, tx)
        assert np.linalg.norm(a - b) / np.linalg.norm(a) < 1e-6
    # the receiver's transform undoes the transmitter's
    assert np.linalg.norm(ref.dft_precoding(ref.dft_precoding(z, L, True), L, False) - z) / np.linalg.norm(z) < 1e-5


def test_predecoding_single(port, ref):
    rng = np.random.default_rng(32)
    for n in (14400, 40, 33, 8):  # AVX body (> 32 symbols) and the generic loop (precoding.c:372-377)
        y = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        h = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        for n0 in (0.0, 0.01, 0.5):
            a, b = ref.predecoding_single(y, h, n0), port.predecoding_single(y, h, n0)
            assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()


PUSCH_LINKS = [dict(), dict(cell_id=301, tti=7, n_dmrs=3, cyclic_shift=5, delta_ss=11), dict(L_prb=6, group_hopping=1, tti=4, tbs=2216, mod=2),
               dict(L_prb=50, n_prb=20, sequence_hopping=1, tti=9, cell_id=77, tbs=14112, mod=2), dict(L_prb=3, nof_prb=6, cell_id=9, tbs=392, mod=1),
               dict(L_prb=5, nof_prb=15, cell_id=42, cp_ext=1, tti=12, tbs=1000, mod=2), dict(L_prb=81, n_prb=3, cell_id=503, delta_ss=29, tti=5, tbs=51024),
               # 1 and 2 PRB: the base sequences are the phi(n) tables of TS 36.211 5.5.1.2 instead of Zadoff-Chu
               dict(L_prb=1, nof_prb=6, n_prb=4, cell_id=12, tti=3, tbs=136, mod=1), dict(L_prb=2, nof_prb=25, n_prb=11, cell_id=250, tti=8, tbs=328, mod=2, n_dmrs=5),
               dict(L_prb=1, nof_prb=15, n_prb=0, cell_id=499, group_hopping=1, delta_ss=17, tti=6, tbs=56, mod=1), dict(L_prb=2, nof_prb=6, n_prb=2, cell_id=88, cp_ext=1, cyclic_shift=2, tbs=256, mod=1)]


@pytest.mark.parametrize("kw", PUSCH_LINKS)
def test_dmrs_and_chest(port, ref, kw):
    from oracle import loader

    lk = loader.pusch_link(**kw)
    a, b = ref.dmrs_pusch_gen(lk), port.dmrs_pusch_gen(lk)
    assert np.abs(a - b).max() < 1e-6
    assert np.abs(np.abs(a) - 1).max() < 1e-5
    rng = np.random.default_rng(int(lk[0]))
    data = rng.integers(0, 256, int(lk[12]) // 8, dtype=np.uint8)
    grid = ref.pusch_encode(lk, data)
    noise = (rng.standard_normal(grid.shape) + 1j * rng.standard_normal(grid.shape)).astype(np.complex64) * np.float32(0.03)
    rx = (grid * np.complex64(0.8 * np.exp(0.7j)) + noise).astype(np.complex64)
    ce_r, m_r = ref.chest_ul_pusch(lk, rx)
    ce_p, m_p = port.chest_ul_pusch(lk, rx, b)
    assert np.linalg.norm(ce_r - ce_p) / np.linalg.norm(ce_r) < 1e-6
    assert abs(m_r[0] - m_p[0]) <= 1e-4 * m_r[0] and abs(m_r[1] - m_p[1]) <= 2e-4 * m_r[1]
    assert abs(m_r[2] - m_p[2]) <= 1e-3 * max(1.0, abs(m_r[2]))


def test_reference_pusch_link_closes(ref):
    """The reference's own transmitter and receiver (the end-to-end oracle of tests/test_pusch_chain_gpu.py) agree with each
    other, and its receiver's buffers relate the way the port's stage functions say they do."""
    from oracle import loader

    port = loader.api("port")
    lk = loader.pusch_link(tti=3, rnti=1234)
    rng = np.random.default_rng(33)
    data = rng.integers(0, 256, 75376 // 8, dtype=np.uint8)
    grid = ref.pusch_encode(lk, data)
    noise = (rng.standard_normal(grid.shape) + 1j * rng.standard_normal(grid.shape)).astype(np.complex64) * np.float32(0.02)
    rx = (grid * np.complex64(0.9 * np.exp(-0.4j)) + noise).astype(np.complex64)
    r = ref.pusch_decode(lk, rx)
    assert r["ret"] == 0 and r["crc"] and (r["data"] == data).all()
    data_syms = [l for l in range(14) if l not in (3, 10)]
    y = np.concatenate([rx[l] for l in data_syms])
    h = np.concatenate([r["ce"][l] for l in data_syms])
    d = port.dft_precoding(port.predecoding_single(y, h, r["noise"]), 100, False)
    assert np.linalg.norm(d - r["d"]) / np.linalg.norm(r["d"]) < 1e-6
    q = port.pusch_seq_apply_s(port.demod_s(3, r["d"]), 1234, 6, 1)
    assert (q == r["q"]).all()
    assert (port.ulsch_deinterleave(q, 6, 12) == r["g"]).all()
