#!/usr/bin/env python3
"""Tiny driver for profiler captures of the PUSCH front end: NSF subframes (100 PRB, N=2048) through OFDM rx + demap +
decode, twice."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200 import synth_pusch as sp  # noqa: E402
from srslte_b200.pusch import PuschRx  # noqa: E402

nsf = int(os.environ.get("NSF", 2048))
iq8, payload8, G = sp.make_subframes(100, 2048, 75376, 6, 0, sp.qpp_interleaver(5824), 8, 23.0, seed=1)
x = torch.from_numpy(np.ascontiguousarray(np.tile(iq8, (nsf // 8, 1)))).cuda()
rx = PuschRx(100, 75376, 3, llr_shift=4, max_noi=8, symbol_sz=2048)
for _ in range(2):
    ok, its = rx.run(x, nsf)
torch.cuda.synchronize()
print("ok", bool(ok.all()), float(its.mean()))
