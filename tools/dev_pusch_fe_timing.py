#!/usr/bin/env python3
"""Developer timing of the PUSCH chain kernels on random device data (not the contract bench)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200.pusch import PuschChain  # noqa: E402

nsf = int(os.environ.get("NSF", 4096))
L = int(os.environ.get("L_PRB", 100))
ch = PuschChain(1, 100, False, L, 0, 3, 4)
g = torch.randn((nsf, 14, 1200, 2), device="cuda")
grid = torch.view_as_complex(g)
tti = np.arange(nsf, dtype=np.uint32)
rnti = np.arange(nsf, dtype=np.uint32) + 61


def timed(f, n=10):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


ce, meas = ch.chest(grid, tti)
d = ch.equalize_deprecode(grid, ce, meas)
out = torch.empty((nsf, ch.nof_bits), dtype=torch.int16, device="cuda")
M = 12 * L
print(f"nsf={nsf} L_prb={L}")
t = timed(lambda: ch.chest(grid, tti))
print(f"chest              {t:.3f} ms  {nsf*(2*M*8*2+2*M*8)/t/1e6:.0f} GB/s")
t = timed(lambda: ch.equalize_deprecode(grid, ce, meas))
print(f"equalize+deprecode {t:.3f} ms  {nsf*(12*M*8*2+2*M*8)/t/1e6:.0f} GB/s")
t = timed(lambda: ch.demod_descramble(d, rnti, tti, out=out))
print(f"demod+descr+deint  {t:.3f} ms  {nsf*12*M*(8+12+1.5)/t/1e6:.0f} GB/s")
