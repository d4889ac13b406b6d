#!/usr/bin/env python3
"""Consistency soak of the one-call receiver's host logic: a random sequence of calls (batch sizes around the chunk and group
boundaries, first transmissions and retransmissions mixed per subframe, host float / int16 samples, the one-shot and the
begin / finish form) on one object; prints a digest of every byte and result it returned.  Run it twice, once with
SRSLTE_B200_ENB_UL_NO_GROUPS=1 (one decode over the whole batch): the digests must be equal.
usage: soak_enb_ul.py [calls] [seed]"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from srslte_b200 import synth_pusch as sp  # noqa: E402
from srslte_b200.pusch import EnbUl, PUSCH_RES_DTYPE, PuschChain  # noqa: E402

ncalls = int(sys.argv[1]) if len(sys.argv) > 1 else 16
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
tbs, nd = 1544, 32
ch = PuschChain(33, 15, False, 15, 0, 1, 1)
dm = {sf: ch.dmrs(sf, 0) for sf in range(10)}
ch.close()
rnti_d = rng.integers(1, 65000, nd).astype(np.uint32)
tti_d = rng.integers(0, 10240, nd).astype(np.uint32)
qpp = sp.qpp_interleaver(1568)
# the same transport blocks as rv 0 and rv 2 at an SNR where a first transmission often fails and the combination mostly works
pool = {rv: sp.make_subframes_full(33, 15, 256, tbs, 2, rv, qpp, nd, rnti_d, tti_d, lambda sf: dm[sf], 0.5, seed=9, fading=False)[0] for rv in (0, 2)}
peak = max(np.abs(p.view(np.float32)).max() for p in pool.values())
enb = EnbUl(33, 15, tbs, 1, llr_shift=1, max_noi=8, symbol_sz=256)
h = hashlib.sha256()
cap = 3200
which = np.zeros(cap, np.int64)            # which pool entry lives in each HARQ slot
failed = np.zeros(cap, bool)               # slots whose last transmission failed (candidates for a retransmission)
ok_total = sf_total = 0
for call in range(ncalls):
    nsf = int(rng.choice([1, 37, 511, 600, 1024, 1100, 2047, 2048, 2597, 3100]))
    retx = failed[:nsf] & (rng.random(nsf) < 0.7) & (call > 0)
    which[:nsf] = np.where(retx, which[:nsf], rng.integers(0, nd, nsf))
    rv = np.where(retx, 2, 0).astype(np.uint32)
    new_data = (~retx).astype(np.uint32)
    iq = np.where(retx[:, None], pool[2][which[:nsf]], pool[0][which[:nsf]])
    rnti, tti = rnti_d[which[:nsf]], tti_d[which[:nsf]]
    use16 = bool(rng.integers(0, 2))
    if use16:
        samples = np.round(iq.view(np.float32).reshape(nsf, -1, 2) * (16384.0 / peak)).astype(np.int16)
    else:
        samples = np.ascontiguousarray(iq)
    hs = torch.from_numpy(samples).pin_memory()
    data = torch.zeros((nsf, enb.tb_bytes), dtype=torch.uint8).pin_memory()
    res = np.zeros(nsf, PUSCH_RES_DTYPE)
    fl = 8 if use16 else 0
    if rng.integers(0, 2):
        enb.run_ptr(hs.data_ptr(), nsf, rnti, tti, data.data_ptr(), res, rv=rv, new_data=new_data, flags=fl)
    else:
        enb.begin_ptr(hs.data_ptr(), nsf, rnti, tti, data.data_ptr(), res, rv=rv, new_data=new_data, flags=fl)
        enb.finish()
    failed[:nsf] = res["crc_ok"] == 0
    h.update(data.numpy().tobytes())
    h.update(res["crc_ok"].tobytes())
    h.update(res["avg_iterations"].tobytes())
    ok_total += int((res["crc_ok"] != 0).sum())
    sf_total += nsf
enb.close()
print(f"soak_enb_ul: {ncalls} calls, {sf_total} subframes, {ok_total} decoded, digest {h.hexdigest()[:24]}")
