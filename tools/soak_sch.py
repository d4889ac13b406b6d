#!/usr/bin/env python3
"""Randomised differential soak of the transport-block decode loop (srsran_b200_sch_decode_batch and its begin / finish halves,
host buffers) against the oracle port's decode_tb: random standard transport block sizes, modulations, G, noise levels, first
transmissions followed by an rv-2 retransmission onto the kept soft buffers and CRC masks.
usage: soak_sch.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import make_tb  # noqa: E402
from oracle import loader  # noqa: E402
from srslte_b200 import SchDecoder  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
port = loader.api("port")
SB = 18600
Ks = port.cb_sizes()
# standard sizes without filler bits: one code block (K - 24) or C equal blocks (C (K - 24) - 24)
sizes = [k - 24 for k in Ks if k >= 64] + [c * (k - 24) - 24 for c in (2, 3, 5) for k in Ks if k >= 3136 and k % 64 == 0]
dec = SchDecoder(device=0, max_noi=8)
t0, ncalls, ntb = time.time(), 0, 0
while time.time() - t0 < budget:
    n = int(rng.integers(1, 7))
    tbs_l, Qm_l, G_l, e_l, sig = [], [], [], [], []
    for t in range(n):
        tbs = int(rng.choice(sizes))
        s = port.cbsegm(tbs)
        if s["F"] or s["C2"]:
            continue
        Qm = int(rng.choice([2, 4, 6]))
        rate = float(rng.choice([0.35, 0.5, 0.75, 1.2]))          # received bits per coded bit: puncturing ... repetition
        G = max(Qm * s["C"], int((3 * (tbs + 24 * s["C"] + 24) * rate) // (Qm * s["C"])) * Qm * s["C"] + Qm * int(rng.integers(0, s["C"])))
        tbs_l.append(tbs); Qm_l.append(Qm); G_l.append(G); sig.append(float(rng.choice([0.6, 0.85, 1.0, 1.3])))
    if not tbs_l:
        continue
    n = len(tbs_l)
    Cs = [port.cbsegm(t)["C"] for t in tbs_l]
    soft_off = np.concatenate([[0], np.cumsum([c * SB for c in Cs])]).astype(np.int64)
    strides = [(t // 8 + 3 + 768 + 15) // 16 * 16 for t in tbs_l]
    data_off = np.concatenate([[0], np.cumsum(strides)]).astype(np.int64)
    soft = np.zeros(int(soft_off[-1]), np.int16)
    soft_o = soft.copy()
    masks = [0] * n
    cbcrc = [np.zeros(c, np.uint8) for c in Cs]
    seed0 = int(rng.integers(0, 1 << 30))
    data = np.zeros(int(data_off[-1]) + 1024, np.uint8)   # kept across the two calls: decoded blocks keep their bytes (sch.c:466-471)
    live = list(range(n))
    for rv, new_data in ((0, 1), (2, 0)):
        if not live:
            break
        es = {t: make_tb(port, tbs_l[t], Qm_l[t], G_l[t], rv, sig[t], seed=seed0 + t)[0] for t in live}
        e_off, pos = {}, 0
        for t in live:
            e_off[t] = pos
            pos += G_l[t]
        e_all = np.concatenate([es[t] for t in live])
        rc, res = dec.decode(e_all, soft, data, [dict(tbs=tbs_l[t], Qm=Qm_l[t], rv=rv, nof_e_bits=G_l[t], e_offset=int(e_off[t]),
                                                      soft_offset=int(soft_off[t]), data_offset=int(data_off[t]), new_data=new_data,
                                                      cb_crc_mask=masks[t]) for t in live])
        assert rc == 0
        failed = []
        for i, t in enumerate(live):
            C = Cs[t]
            so = soft_o[soft_off[t]:soft_off[t + 1]]
            want = np.zeros(strides[t] + 1024, np.uint8)
            if not new_data:  # blocks decoded in the first transmission keep their bytes (sch.c:466-471)
                want[:strides[t]] = prev_data[t]
            ret, iters = port.decode_tb(es[t], tbs_l[t], Qm_l[t], rv, 8, so, cbcrc[t], want)
            got = data[data_off[t]:data_off[t] + tbs_l[t] // 8 + 3]
            assert res[i]["result"] == ret, ("result", tbs_l[t], Qm_l[t], G_l[t], rv, res[i], ret)
            assert abs(res[i]["avg_iterations"] - iters / C) < 1e-6, ("iterations", tbs_l[t], rv)
            assert (soft[soft_off[t]:soft_off[t + 1]] == so).all(), ("soft", tbs_l[t], rv)
            masks[t] = sum(int(b) << c for c, b in enumerate(cbcrc[t]))
            assert res[i]["cb_crc_mask"] == masks[t], ("mask", tbs_l[t], rv)
            if ret == 0:
                assert (got == want[:tbs_l[t] // 8 + 3]).all(), ("bytes", tbs_l[t], rv)
            elif masks[t] != (1 << C) - 1:
                failed.append(t)  # HARQ: only a block that is still missing something is sent again
            ntb += 1
        prev_data = {t: data[data_off[t]:data_off[t] + strides[t]].copy() for t in range(n)}
        live = failed
        ncalls += 1
dec.close()
print(f"soak_sch ok: {ncalls} calls, {ntb} transport blocks, {time.time() - t0:.0f} s")
