"""Shared test-vector helpers (numpy + the oracle's encoder)."""
from __future__ import annotations

import numpy as np


def coded_llrs(api, K: int, ncb: int, sigma: float, scale: float = 16.0, clip: int = 31, seed: int = 0,
               crc: str | None = "B", zero_frac: float = 0.0):
    """ncb random code blocks: payload (+CRC), turbo-encoded by `api`, BPSK + AWGN, quantised to int16.

    Returns (llr (ncb, 3K+12) int16, bits (ncb, K) uint8)."""
    rng = np.random.default_rng(seed)
    llr = np.zeros((ncb, 3 * K + 12), np.int16)
    allbits = np.zeros((ncb, K), np.uint8)
    for c in range(ncb):
        bits = rng.integers(0, 2, K).astype(np.uint8)
        if crc is not None:
            r = api.crc24(crc, np.packbits(bits[:K - 24]), K - 24)
            bits[K - 24:] = [(r >> (23 - i)) & 1 for i in range(24)]
        cw = api.tcod_encode(bits)
        y = (2.0 * cw - 1.0) + rng.normal(size=cw.size) * sigma
        q = np.clip(np.rint(scale * y), -clip, clip)
        if zero_frac > 0:  # punctured positions carry no information
            q[rng.random(q.size) < zero_frac] = 0
        llr[c] = q.astype(np.int16)
        allbits[c] = bits
    return llr, allbits


def npass_of(ok, npass_crc, npass_run):
    return np.where(ok == 1, npass_crc, npass_run)


def make_tb(api, tbs: int, Qm: int, G: int, rv: int, sigma: float, scale: float = 16.0, clip: int = 31, seed: int = 0):
    """One transport block through the transmit chain of sch.c:240-350 restated with the oracle's pieces: TB CRC24A,
    segmentation (standard TBS only: F == 0, C2 == 0), CB CRC24B, turbo encoding, rate matching with the encoder-side
    E split, BPSK-like +-1 mapping + AWGN, int16 quantisation.

    Returns (e_bits int16 (G,), expected data bytes (tbs/8+3,), segm dict)."""
    rng = np.random.default_rng(seed)
    s = api.cbsegm(tbs)
    assert s["F"] == 0 and s["C2"] == 0, "use a standard TBS"
    C, K = s["C"], s["K1"]
    payload = rng.integers(0, 2, tbs).astype(np.uint8)
    par = api.crc24("A", np.packbits(payload), tbs)
    tb = np.concatenate([payload, np.array([(par >> (23 - i)) & 1 for i in range(24)], np.uint8)])
    Gp = G // Qm
    gamma = Gp % C
    rlen = K if C == 1 else K - 24
    tx = []
    for c in range(C):
        bits = tb[c * rlen:(c + 1) * rlen]
        if C > 1:
            r = api.crc24("B", np.packbits(bits), rlen)
            bits = np.concatenate([bits, np.array([(r >> (23 - i)) & 1 for i in range(24)], np.uint8)])
        cw = api.tcod_encode(bits)
        n_e = Qm * (Gp // C) if c <= C - gamma - 1 else Qm * -(-Gp // C)
        tx.append(api.rm_tx(cw, K, n_e, rv))
    tx = np.concatenate(tx)
    tx = np.concatenate([tx, np.zeros(G - tx.size, np.uint8)])[:G]
    y = (2.0 * tx - 1.0) + rng.normal(size=G) * sigma
    e = np.clip(np.rint(scale * y), -clip, clip).astype(np.int16)
    return e, np.packbits(tb), s
