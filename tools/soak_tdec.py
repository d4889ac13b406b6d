#!/usr/bin/env python3
"""Randomised differential soak of the batched decoder against the oracle port (developer tool; the fixed cases live in tests/).
usage: python tools/soak_tdec.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import coded_llrs  # noqa: E402
from oracle import loader  # noqa: E402
from srslte_b200 import TurboDecoderBatch  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
port = loader.api("port")
Ks = port.cb_sizes()
dec = TurboDecoderBatch(device=0)
t0, n, blocks = time.time(), 0, 0
while time.time() - t0 < budget:
    K = int(rng.choice(Ks))
    ncb = int(rng.choice([1, 2, 63, 64, 65, 127, 128, 129, 200]))
    if K > 3000:
        ncb = min(ncb, 65)
    sigma = float(rng.choice([0.5, 0.8, 0.95, 1.1, 1.5]))
    scale, clip = [(16.0, 31), (8.0, 15), (32.0, 127), (40.0, 200), (500.0, 2000), (8000.0, 30000)][int(rng.integers(0, 6))]
    mp = int(rng.choice([1, 2, 3, 5, 8, 10]))
    early = bool(rng.integers(0, 2))
    crc = ["B", "A", None][int(rng.integers(0, 3))]
    llr, _ = coded_llrs(port, K, ncb, sigma, scale, clip, seed=int(rng.integers(0, 1 << 30)))
    if rng.integers(0, 4) == 0:  # full-range garbage in a few blocks
        k = int(rng.integers(0, ncb))
        llr[k] = rng.integers(-32768, 32768, llr.shape[1]).astype(np.int16)
    out, ok, npass = dec.decode(llr, K, mp, crc, early)
    o1, k1, n1, _ = port.decode_batch(llr, K, mp, crc, K if crc == "A" else 0, early, nthreads=8)  # CRC24A runs over the K = tbs+24 bits
    if not ((out == o1).all() and (ok == k1).all() and (npass == n1).all()):
        print("MISMATCH", dict(K=K, ncb=ncb, sigma=sigma, scale=scale, clip=clip, mp=mp, early=early, crc=crc), flush=True)
        sys.exit(1)
    n += 1
    blocks += ncb
print(f"soak ok: {n} random batches, {blocks} code blocks, {time.time()-t0:.0f} s")
