#!/usr/bin/env python3
"""Developer measurement: wall time of one srsran_b200_enb_ul_pusch_batch call for small batches (latency, not throughput)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200 import synth_pusch as sp  # noqa: E402
from srslte_b200.pusch import PUSCH_RES_DTYPE, EnbUl, PuschChain  # noqa: E402

tbs = 75376
ch = PuschChain(1, 100, False, 100, 0, 3, 4)
dm = {sf: ch.dmrs(sf, 0) for sf in range(10)}
ch.close()
rnti8 = np.arange(8, dtype=np.uint32) * 97 + 62
tti8 = np.arange(8, dtype=np.uint32) * 3
iq8, payload8, _ = sp.make_subframes_full(1, 100, 2048, tbs, 6, 0, sp.qpp_interleaver(5824), 8, rnti8, tti8, lambda sf: dm[sf], 23.0, seed=1)
enb = EnbUl(1, 100, tbs, 3, llr_shift=4, max_noi=8, symbol_sz=2048)
for nsf in (1, 2, 8, 32, 128, 512):
    reps = -(-nsf // 8)
    h_iq = torch.from_numpy(np.ascontiguousarray(np.tile(iq8, (reps, 1))[:nsf])).pin_memory()
    rnti, tti = np.tile(rnti8, reps)[:nsf], np.tile(tti8, reps)[:nsf]
    out = torch.empty((nsf, enb.tb_bytes), dtype=torch.uint8).pin_memory()
    res = np.zeros(nsf, PUSCH_RES_DTYPE)
    for _ in range(3):
        enb.run_ptr(h_iq.data_ptr(), nsf, rnti, tti, out.data_ptr(), res)
    t0 = time.perf_counter()
    n = 20
    for _ in range(n):
        enb.run_ptr(h_iq.data_ptr(), nsf, rnti, tti, out.data_ptr(), res)
    ms = (time.perf_counter() - t0) / n * 1e3
    print(f"nsf={nsf:4d}: {ms:7.3f} ms per call, {nsf/ms*1e3:9.0f} subframes/s, all ok {bool(res['crc_ok'].all())}, mean passes {res['avg_iterations'].mean():.2f}")
