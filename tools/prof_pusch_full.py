#!/usr/bin/env python3
"""Tiny driver for profiler captures of the complete PUSCH receive chain: NSF subframes (100 PRB, N=2048) through OFDM rx,
channel estimation, equaliser + transform de-precoding, demap + descrambling + de-interleave, de-matching and decoding, twice."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200 import synth_pusch as sp  # noqa: E402
from srslte_b200.pusch import PuschRxFull  # noqa: E402

nsf = int(os.environ.get("NSF", 2048))
rx = PuschRxFull(1, 100, 75376, 3, llr_shift=4, max_noi=8, symbol_sz=2048)
rnti8 = np.arange(8, dtype=np.uint32) * 97 + 62
tti8 = np.arange(8, dtype=np.uint32) * 3
iq8, payload8, G = sp.make_subframes_full(1, 100, 2048, 75376, 6, 0, sp.qpp_interleaver(5824), 8, rnti8, tti8, lambda sf: rx.chain.dmrs(sf, 0),
                                          23.0, seed=1)
x = torch.from_numpy(np.ascontiguousarray(np.tile(iq8, (nsf // 8, 1)))).cuda()
rnti, tti = np.tile(rnti8, nsf // 8), np.tile(tti8, nsf // 8)
for _ in range(2):
    ok, its = rx.run(x, nsf, rnti, tti)
torch.cuda.synchronize()
print("ok", bool(ok.all()), float(its.mean()))
