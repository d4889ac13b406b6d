#!/usr/bin/env python3
"""Times the PUSCH front-end stages one by one (CUDA events, random input: the kernels' run time does not depend on the values):
OFDM receive, channel estimation, equaliser + transform de-precoding, demap + descramble + de-interleave, for nsf subframes of
100 PRB / 64QAM.  Development tool for the front-end kernels; the judged numbers are bench.py's `pusch_full.front_end`."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from srslte_b200.ofdm import OfdmRx
from srslte_b200.pusch import PuschChain

ap = argparse.ArgumentParser()
ap.add_argument("--nsf", type=int, default=4096)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--prb", type=int, default=100)
ap.add_argument("--mod", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
nsf, M = a.nsf, 12 * a.prb
ofdm = OfdmRx(a.prb, symbol_sz=2048 if a.prb == 100 else 0)
ch = PuschChain(cell_id=1, cell_nof_prb=a.prb, L_prb=a.prb, n_prb=0, mod=a.mod, llr_shift=4)
g = torch.Generator(device=dev).manual_seed(1)
iq = torch.view_as_complex(torch.randn((nsf, ofdm.sf_sz, 2), device=dev, generator=g) * 0.1)
grid = torch.empty((nsf, 14, M), dtype=torch.complex64, device=dev)
rnti = np.arange(nsf, dtype=np.uint32) % 60000 + 1
tti = np.arange(nsf, dtype=np.uint32) % 10240
st = torch.cuda.current_stream(dev).cuda_stream
ofdm.rx_sf_device(iq, grid, nsf, st)
ce, meas = ch.chest(grid, tti)
d = ch.equalize_deprecode(grid, ce, meas)
llr = ch.demod_descramble(d, rnti, tti)
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    ts = []
    for _ in range(a.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


Qm = 2 * a.mod
res = {
    "ofdm_ms": timeit(lambda: ofdm.rx_sf_device(iq, grid, nsf, st)),
    "chest_ms": timeit(lambda: ch.chest(grid, tti, out=(ce, meas))),
    "equalize_deprecode_ms": timeit(lambda: ch.equalize_deprecode(grid, ce, meas, out=d)),
    "demod_ms": timeit(lambda: ch.demod_descramble(d, rnti, tti, out=llr)),
}
byt = {"ofdm_ms": nsf * (ofdm.sf_sz * 8 + 14 * M * 8), "chest_ms": nsf * 2 * M * 8 * 3, "equalize_deprecode_ms": nsf * (12 * M * 8 * 2 + 2 * M * 8),
       "demod_ms": nsf * 12 * M * (8 + 2 * Qm + Qm / 8.0 * 2)}
for k in list(res):
    res[k.replace("_ms", "_gbs")] = round(byt[k] / (res[k] * 1e-3) / 1e9, 1)
print(json.dumps(res))
