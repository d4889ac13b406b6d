"""Synthetic PUSCH subframes for benchmarks and tests (numpy only; never on the receive path).

A plain restatement of the TRANSMIT side from the 3GPP text, independent of the receive kernels it feeds:
  36.212 5.1.1 CRC24A/B, 5.1.2 code block segmentation (standard transport block sizes: no filler bits, one block size),
  5.1.3.2 turbo encoder (g0 = 1+D^2+D^3, g1 = 1+D+D^3, trellis termination, QPP interleaver), 5.1.4.1 rate matching
  (32-column sub-block interleaver, circular buffer, k0(rv)), 5.1.5 concatenation; 36.211 7.1 modulation mapper and 5.6
  SC-FDMA baseband signal generation (half-subcarrier shift, normal cyclic prefix).
What BASELINE.json config 4 leaves out is left out here as well: no scrambling, no channel interleaver, no transform
precoding, identity channel; the DMRS symbols (3 and 10) carry unit-modulus filler.  The receive chain of the reference
that undoes this is enb_ul.c:58,153 (srsran_ofdm_rx_sf), pusch.c:449 (soft demapper), sch.c:507-572 (decode_tb).
"""
from __future__ import annotations

import numpy as np

CRC24A = 0x1864CFB
CRC24B = 0x1800063
# 36.212 5.1.4.1.1 inter-column permutation pattern of the sub-block interleaver
_P32 = np.array([0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30, 1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31])


def qpp_interleaver(K: int) -> np.ndarray:
    """PI(i) = (f1 i + f2 i^2) mod K with (f1, f2) of 36.212 table 5.1.3-3 (the table file the library is built from)."""
    import os
    import re

    txt = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "qpp_table.inc")).read()
    rows = {int(a): (int(b), int(c)) for a, b, c in re.findall(r"\{(\d+),(\d+),(\d+)\}", txt)}
    f1, f2 = rows[K]
    i = np.arange(K, dtype=np.int64)
    return ((f1 * i + f2 * i * i) % K).astype(np.int64)


def _crc_table(poly: int) -> np.ndarray:
    t = np.zeros(256, np.uint32)
    for b in range(256):
        r = b << 16
        for _ in range(8):
            r = ((r << 1) ^ (poly if r & 0x800000 else 0)) & 0xFFFFFF
        t[b] = r
    return t


_T = {CRC24A: _crc_table(CRC24A), CRC24B: _crc_table(CRC24B)}


def crc24(bits: np.ndarray, poly: int) -> np.ndarray:
    """bits: (n, nbits) uint8, nbits multiple of 8.  Returns the 24 parity bits (n, 24), MSB first."""
    by = np.packbits(bits, axis=1).astype(np.uint32)
    r = np.zeros(bits.shape[0], np.uint32)
    t = _T[poly]
    for j in range(by.shape[1]):
        r = ((r << 8) & 0xFFFFFF) ^ t[((r >> 16) ^ by[:, j]) & 0xFF]
    return ((r[:, None] >> (23 - np.arange(24))) & 1).astype(np.uint8)


def segment(tbs: int):
    """36.212 5.1.2 for transport block sizes without filler bits.  Returns (C, K)."""
    B = tbs + 24
    if B <= 6144:
        return 1, B
    C = -(-B // (6144 - 24))
    Bp = B + 24 * C
    assert Bp % C == 0, "not a standard transport block size (filler bits needed)"
    return C, Bp // C


def _rsc(c: np.ndarray):
    """Constituent encoder over (n, K) bits: parity (n, K), and the 3 termination (x, z) pairs (n, 3) each."""
    n, K = c.shape
    s1 = np.zeros(n, np.uint8)
    s2 = np.zeros(n, np.uint8)
    s3 = np.zeros(n, np.uint8)
    z = np.zeros((n, K), np.uint8)
    for k in range(K):
        a = c[:, k] ^ s2 ^ s3          # feedback 1 + D^2 + D^3
        z[:, k] = a ^ s1 ^ s3          # feed-forward 1 + D + D^3
        s1, s2, s3 = a, s1, s2
    xt = np.zeros((n, 3), np.uint8)
    zt = np.zeros((n, 3), np.uint8)
    for t in range(3):                 # switch in the lower position: the input is the feedback itself
        x = s2 ^ s3
        a = np.zeros(n, np.uint8)
        xt[:, t] = x
        zt[:, t] = a ^ s1 ^ s3
        s1, s2, s3 = a, s1, s2
    return z, xt, zt


def turbo_encode(c: np.ndarray, qpp: np.ndarray):
    """c: (n, K) bits; qpp: PI(i).  Returns d (n, 3, K+4): the three output streams incl. the multiplexed tail."""
    n, K = c.shape
    z, xt, zt = _rsc(c)
    zp, xpt, zpt = _rsc(c[:, qpp])
    d = np.zeros((n, 3, K + 4), np.uint8)
    d[:, 0, :K], d[:, 1, :K], d[:, 2, :K] = c, z, zp
    d[:, 0, K], d[:, 1, K], d[:, 2, K] = xt[:, 0], zt[:, 0], xt[:, 1]
    d[:, 0, K + 1], d[:, 1, K + 1], d[:, 2, K + 1] = zt[:, 1], xt[:, 2], zt[:, 2]
    d[:, 0, K + 2], d[:, 1, K + 2], d[:, 2, K + 2] = xpt[:, 0], zpt[:, 0], xpt[:, 1]
    d[:, 0, K + 3], d[:, 1, K + 3], d[:, 2, K + 3] = zpt[:, 1], xpt[:, 2], zpt[:, 2]
    return d


def rate_match(d: np.ndarray, E: int, rv: int) -> np.ndarray:
    """36.212 5.1.4.1.  d: (n, 3, D) -> (n, E) bits."""
    n, _, D = d.shape
    R = -(-D // 32)
    Kp = 32 * R
    ND = Kp - D
    NULL = 2
    y = np.full((n, 3, Kp), NULL, np.uint8)
    y[:, :, ND:] = d
    k = np.arange(Kp)
    # v0, v1: written row by row into R x 32, columns permuted, read column by column
    src01 = (k % R) * 32 + _P32[k // R]
    src2 = (_P32[k // R] + 32 * (k % R) + 1) % Kp
    v0, v1, v2 = y[:, 0, src01], y[:, 1, src01], y[:, 2, src2]
    w = np.empty((n, 3 * Kp), np.uint8)
    w[:, :Kp] = v0
    w[:, Kp::2] = v1
    w[:, Kp + 1::2] = v2
    Ncb = 3 * Kp
    k0 = R * (2 * (-(-Ncb // (8 * R))) * rv + 2)
    idx = (k0 + np.arange(2 * Ncb + E)) % Ncb       # enough positions to skip every dummy
    keep = idx[w[0, idx] != NULL][:E]               # the dummy pattern is the same for every block
    assert keep.size == E
    return w[:, keep]


def _qam(bits: np.ndarray, Qm: int) -> np.ndarray:
    """36.211 7.1.2-7.1.4.  bits (..., Qm*nsym) -> complex symbols of unit average power."""
    b = bits.reshape(bits.shape[:-1] + (-1, Qm)).astype(np.float64)
    s = 1 - 2 * b
    if Qm == 2:
        i, q, nrm = s[..., 0], s[..., 1], np.sqrt(2.0)
    elif Qm == 4:
        i, q, nrm = s[..., 0] * (2 - s[..., 2]), s[..., 1] * (2 - s[..., 3]), np.sqrt(10.0)
    else:
        i = s[..., 0] * (4 - s[..., 2] * (2 - s[..., 4]))
        q = s[..., 1] * (4 - s[..., 3] * (2 - s[..., 5]))
        nrm = np.sqrt(42.0)
    return (i + 1j * q) / nrm


_TB_CACHE: dict = {}


def make_transport_blocks(tbs: int, Qm: int, G: int, rv: int, qpp: np.ndarray, n: int, seed: int):
    """n random transport blocks -> (coded bits (n, G) uint8, payload bytes incl. the TB CRC (n, tbs/8+3)).
    The result for the last few argument sets is kept: the cells of the multi-cell benchmark transmit the same payloads and
    differ in scrambling, reference signals and noise, so the (slow, pure numpy) turbo encoding runs once."""
    key = (tbs, Qm, G, rv, n, seed, int(qpp.size))
    if key in _TB_CACHE:
        return _TB_CACHE[key]
    if len(_TB_CACHE) > 4:
        _TB_CACHE.clear()
    _TB_CACHE[key] = _make_transport_blocks(tbs, Qm, G, rv, qpp, n, seed)
    return _TB_CACHE[key]


def _make_transport_blocks(tbs: int, Qm: int, G: int, rv: int, qpp: np.ndarray, n: int, seed: int):
    rng = np.random.default_rng(seed)
    C, K = segment(tbs)
    assert qpp.size == K
    payload = rng.integers(0, 2, (n, tbs)).astype(np.uint8)
    tb = np.concatenate([payload, crc24(payload, CRC24A)], axis=1)
    rlen = K if C == 1 else K - 24
    blocks = tb.reshape(n * C, rlen)
    if C > 1:
        blocks = np.concatenate([blocks, crc24(blocks, CRC24B)], axis=1)
    d = turbo_encode(blocks, qpp).reshape(n, C, 3, K + 4)
    Gp = G // Qm
    gamma = Gp % C
    out = []
    for r in range(C):
        E = Qm * (Gp // C) if r <= C - gamma - 1 else Qm * (-(-Gp // C))
        out.append(rate_match(d[:, r], E, rv))
    f = np.concatenate(out, axis=1)
    assert f.shape[1] == G
    return f, np.packbits(tb, axis=1)


def ofdm_modulate(grid: np.ndarray, N: int, half_shift: bool = True) -> np.ndarray:
    """grid (nsf, 14, R) -> time samples (nsf, 15 N) complex64, normal CP, such that an unnormalised receive DFT returns
    the grid.  36.211 5.6: s_l(t) = sum_k a_k exp(j 2 pi (k + 1/2) df (t - Ncp Ts)), k = -R/2 .. R/2-1."""
    nsf, nsym, R = grid.shape
    assert nsym == 14
    cp = [-(-160 * N // 2048)] + [-(-144 * N // 2048)] * 6
    k = np.arange(R) - R // 2
    out = np.zeros((nsf, 15 * N), np.complex64)
    pos = 0
    for l in range(14):
        c = cp[l % 7]
        X = np.zeros((nsf, N), np.complex128)
        X[:, k % N] = grid[:, l, :]
        x = np.fft.ifft(X, axis=1)                      # (1/N) sum X[k] e^{+j2pi kn/N}
        n = np.arange(-c, N)
        s = x[:, n % N]
        if half_shift:
            s = s * np.exp(1j * np.pi * n / N)[None, :]  # (k + 1/2): not N-periodic, so applied over CP + body
        out[:, pos:pos + c + N] = s
        pos += c + N
    assert pos == 15 * N
    return out


def make_subframes(nof_prb: int, N: int, tbs: int, Qm: int, rv: int, qpp: np.ndarray, n: int, snr_db: float, seed: int):
    """n distinct PUSCH subframes of BASELINE config 4.  Returns (iq (n, 15N) complex64, payload bytes (n, tbs/8+3), G)."""
    R = 12 * nof_prb
    G = 12 * R * Qm
    f, payload = make_transport_blocks(tbs, Qm, G, rv, qpp, n, seed)
    data = _qam(f, Qm).reshape(n, 12, R)
    rng = np.random.default_rng(seed + 1)
    grid = np.zeros((n, 14, R), np.complex128)
    data_syms = [l for l in range(14) if l not in (3, 10)]
    grid[:, data_syms, :] = data
    grid[:, [3, 10], :] = np.exp(2j * np.pi * rng.random((n, 2, R)))
    iq = ofdm_modulate(grid, N).astype(np.complex128)
    sigma_f = 10 ** (-snr_db / 20.0)                     # per-RE noise std (complex), unit signal power
    sigma_t = sigma_f / np.sqrt(N)
    iq += (rng.normal(size=iq.shape) + 1j * rng.normal(size=iq.shape)) * (sigma_t / np.sqrt(2.0))
    return iq.astype(np.complex64), payload, G


# ---- full PUSCH transmitter (36.212 5.2.2.7-8, 36.211 5.3.1-5.3.4, 5.5.2.1) -------------------------------------------
def gold_bits(c_init: np.ndarray, nbits: int) -> np.ndarray:
    """36.211 7.2 pseudo-random sequence c(0..nbits-1) for every seed in c_init -> (n, nbits) uint8."""
    c_init = np.atleast_1d(np.asarray(c_init, np.uint32))
    n = c_init.size
    x1 = np.zeros((n, 1600 + nbits + 31), np.uint8)
    x2 = np.zeros_like(x1)
    x1[:, 0] = 1
    x2[:, :31] = (c_init[:, None] >> np.arange(31)[None, :]) & 1
    # the recurrences reach back 28..31 samples, so 28 new samples can be produced per vector step
    step = 28
    for p in range(31, x1.shape[1], step):
        e = min(p + step, x1.shape[1])
        w = e - p
        x1[:, p:e] = x1[:, p - 28:p - 28 + w] ^ x1[:, p - 31:p - 31 + w]
        x2[:, p:e] = x2[:, p - 28:p - 28 + w] ^ x2[:, p - 29:p - 29 + w] ^ x2[:, p - 30:p - 30 + w] ^ x2[:, p - 31:p - 31 + w]
    return x1[:, 1600:1600 + nbits] ^ x2[:, 1600:1600 + nbits]


def make_pusch_grids(cell_id: int, cell_nof_prb: int, L_prb: int, n_prb: int, tbs: int, Qm: int, rv: int, qpp: np.ndarray, n: int,
                     rnti: np.ndarray, tti: np.ndarray, dmrs, seed: int):
    """n PUSCH subframes as resource grids (n, 14, 12 cell_nof_prb) complex128 before the channel: transport block coding,
    channel interleaver (no control information), scrambling, modulation, transform precoding, mapping, DMRS.
    dmrs(sf_idx) -> (2, 12 L_prb) known reference symbols.  Returns (grid, payload bytes incl. TB CRC)."""
    M, R = 12 * L_prb, 12 * cell_nof_prb
    G = 12 * M * Qm
    f, payload = make_transport_blocks(tbs, Qm, G, rv, qpp, n, seed)
    # 36.212 5.2.2.8: the matrix is written row by row (rows = M, columns = 12 symbols, Qm bits per entry), read column by column
    q = f.reshape(n, M, 12, Qm).transpose(0, 2, 1, 3).reshape(n, G)
    c = gold_bits((np.asarray(rnti, np.uint32) << 14) + ((np.asarray(tti, np.uint32) % 10) << 9) + np.uint32(cell_id), G)
    d = _qam(q ^ c, Qm).reshape(n, 12, M)
    z = np.fft.fft(d, axis=2) / np.sqrt(M)               # 36.211 5.3.3 transform precoding
    grid = np.zeros((n, 14, R), np.complex128)
    data_syms = [l for l in range(14) if l not in (3, 10)]
    grid[:, data_syms, 12 * n_prb:12 * n_prb + M] = z
    for s in range(n):
        r = dmrs(int(tti[s] % 10))
        grid[s, 3, 12 * n_prb:12 * n_prb + M] = r[0]
        grid[s, 10, 12 * n_prb:12 * n_prb + M] = r[1]
    return grid, payload


def make_subframes_full(cell_id: int, nof_prb: int, N: int, tbs: int, Qm: int, rv: int, qpp: np.ndarray, n: int, rnti, tti, dmrs,
                        snr_db: float, seed: int, fading: bool = True, noise: bool = True, return_gain: bool = False):
    """Time-domain PUSCH subframes through a per-subframe flat complex gain with a small timing offset (linear phase over the
    subcarriers) plus AWGN whose level follows the gain, so that every subframe is received at snr_db.
    Returns (iq (n, 15N) complex64, payload bytes (n, tbs/8+3), G)."""
    grid, payload = make_pusch_grids(cell_id, nof_prb, nof_prb, 0, tbs, Qm, rv, qpp, n, rnti, tti, dmrs, seed)
    rng = np.random.default_rng(seed + 1)
    R = 12 * nof_prb
    gain = np.ones(n, np.complex128)
    if fading:
        gain = (0.7 + 0.6 * rng.random(n)) * np.exp(2j * np.pi * rng.random(n))
        slope = 2 * np.pi * (rng.random(n) - 0.5) * 2e-3   # radians per subcarrier: a timing offset of up to +-2 samples at N=2048
        h = gain[:, None] * np.exp(1j * slope[:, None] * (np.arange(R)[None, :] - R / 2))
        grid = grid * h[:, None, :]
    iq = ofdm_modulate(grid, N).astype(np.complex128)
    sigma_t = 10 ** (-snr_db / 20.0) / np.sqrt(N)
    if noise:
        iq += (rng.normal(size=iq.shape) + 1j * rng.normal(size=iq.shape)) * (sigma_t / np.sqrt(2.0)) * np.abs(gain)[:, None]
    if return_gain:  # (noiseless subframes + per-subframe |gain| and the time-domain noise sigma: the caller adds its own noise)
        return iq.astype(np.complex64), payload, 12 * R * Qm, np.abs(gain).astype(np.float32), float(sigma_t)
    return iq.astype(np.complex64), payload, 12 * R * Qm
