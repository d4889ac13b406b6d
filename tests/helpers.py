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
