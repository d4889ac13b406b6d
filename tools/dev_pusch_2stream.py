#!/usr/bin/env python3
"""Developer experiment: T host threads, each with its own complete PUSCH pipeline and CUDA stream, NSF/T subframes per step each."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200 import synth_pusch as sp  # noqa: E402
from srslte_b200.pusch import PuschRxFull  # noqa: E402

NSF = int(os.environ.get("NSF", 4096))
T = int(os.environ.get("T", 2))
STEPS = int(os.environ.get("STEPS", 10))
nsf = NSF // T
rnti8 = np.arange(8, dtype=np.uint32) * 97 + 62
tti8 = np.arange(8, dtype=np.uint32) * 3
pipes = []
for t in range(T):
    rx = PuschRxFull(1, 100, 75376, 3, llr_shift=4, max_noi=8, symbol_sz=2048)
    if t == 0:
        iq8, payload8, G = sp.make_subframes_full(1, 100, 2048, 75376, 6, 0, sp.qpp_interleaver(5824), 8, rnti8, tti8,
                                                  lambda sf: rx.chain.dmrs(sf, 0), 23.0, seed=1)
    x = torch.from_numpy(np.ascontiguousarray(np.tile(iq8, (nsf // 8, 1)))).cuda()
    pipes.append((rx, x, torch.cuda.Stream()))
rnti, tti = np.tile(rnti8, nsf // 8), np.tile(tti8, nsf // 8)
oks = [None] * T


def work(t, steps):
    rx, x, st = pipes[t]
    with torch.cuda.stream(st):
        for _ in range(steps):
            ok, _ = rx.run(x, nsf, rnti, tti)
        oks[t] = bool(ok.all())


def run(steps):
    th = [threading.Thread(target=work, args=(t, steps)) for t in range(T)]
    t0 = time.perf_counter()
    for h in th:
        h.start()
    for h in th:
        h.join()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


run(2)
dt = run(STEPS)
print(f"T={T} nsf/thread={nsf}: {dt/STEPS*1e3:.3f} ms per {NSF} subframes -> {NSF*STEPS/dt:.0f} subframes/s, all ok {all(oks)}")
