#!/usr/bin/env python3
"""Randomised soak of the control-information path: random allocations (1..100 PRB, both prefixes, SRS subframes, all
modulations), random field combinations and offsets, reference transmitter -> fading + noise -> GPU chain, compared with the
reference receiver: Q', de-interleaved stream (exact against the reference's rules applied to the plain stream, last-bit
tolerance against q->g), decided values.  Usage: soak_uci.py [seconds]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import loader  # noqa: E402
from srslte_b200.pusch import PuschChain, uci_cfg  # noqa: E402
import test_pusch_uci_gpu as T  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
ref = loader.api("ref")
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))


def five_smooth(n: int) -> bool:  # srsran_dft_precoding_valid_prb: 2^a 3^b 5^c
    for p in (2, 3, 5):
        while n % p == 0:
            n //= p
    return n == 1


valid_L = [L for L in range(1, 101) if five_smooth(L)]
t_end, it, nsub, soft_mism = time.time() + budget, 0, 0, 0
while time.time() < t_end:
    L = int(rng.choice(valid_L))
    nprb = int(rng.choice([p for p in (6, 15, 25, 50, 75, 100) if p >= L]))
    mod = int(rng.integers(1, 4))
    kw = dict(cell_id=int(rng.integers(0, 504)), cell_nof_prb=nprb, L_prb=L, n_prb=int(rng.integers(0, nprb - L + 1)), mod=mod,
              cp_ext=bool(rng.integers(0, 2)), shortened=bool(rng.integers(0, 2)), cyclic_shift=int(rng.integers(0, 8)), delta_ss=int(rng.integers(0, 30)))
    ch = PuschChain(llr_shift=0, **kw)
    Qm, M, nd = 2 * mod, ch.M, ch.nd
    # a transport block that fits comfortably: a standard size (no filler bits) at a code rate around 0.4
    nbits = M * nd * Qm
    cands = [k - 24 for k in loader.api("port").cb_sizes() if k - 24 <= 0.45 * nbits and k - 24 >= 16]
    if not cands:
        ch.close()
        continue
    tbs = int(cands[-1]) if nbits * 0.45 < 6120 else int(rng.choice([6712, 9144, 12216, 15264, 18336, 21384, 24496, 30576]))
    if tbs > 0.6 * nbits:
        ch.close()
        continue
    nsf = 6
    cases = []
    for s in range(nsf):
        c = {}
        if rng.random() < 0.7:
            c["nof_ack"] = int(rng.choice([1, 2, 3, 5, 10]))
            c["ack_bits"] = int(rng.integers(0, 1 << c["nof_ack"]))
            c["I_offset_ack"] = int(rng.integers(0, 15))
        if rng.random() < 0.5:
            c["ri_len"], c["ri"], c["I_offset_ri"] = 1, int(rng.integers(0, 2)), int(rng.integers(0, 13))
        if rng.random() < 0.6:
            kind = int(rng.integers(1, 4))
            c.update(cqi_kind=kind, cqi_wb=int(rng.integers(0, 16)), I_offset_cqi=int(rng.integers(2, 16)))
            if kind == 3:
                c.update(cqi_N=int(rng.integers(4, 21)), cqi_sb=int(rng.integers(0, 1 << 30)))
            elif kind == 2:
                c["cqi_sb"] = int(rng.integers(0, 4))
        cases.append(c)
    rnti = rng.integers(1, 65000, nsf).astype(np.uint32)
    tti = rng.integers(0, 10240, nsf).astype(np.uint32)
    n_dmrs = rng.integers(0, 8, nsf).astype(np.uint32)
    ucfg = uci_cfg(nsf)
    grids, refs = [], []
    for s, c in enumerate(cases):
        u = loader.pusch_uci(**c)
        ucfg[s] = (c.get("nof_ack", 0), c.get("ri_len", 0), T.cqi_len_of(c), u[8], u[9], u[10])
        lk = T.link_of(loader, kw, int(rnti[s]), int(tti[s]), int(n_dmrs[s]), tbs)
        tx = ref.pusch_encode_uci(lk, u, rng.integers(0, 256, tbs // 8, dtype=np.uint8))
        sigma = np.float32(rng.choice([0.01, 0.05, 0.2, 0.5]))
        h = np.complex64((0.5 + rng.random()) * np.exp(2j * np.pi * rng.random()))
        rxg = (tx * h + (rng.standard_normal(tx.shape) + 1j * rng.standard_normal(tx.shape)).astype(np.complex64) * sigma).astype(np.complex64)
        grids.append(rxg)
        refs.append(ref.pusch_decode_uci(lk, u, rxg))
    grid = torch.from_numpy(np.stack(grids)).cuda()
    g_plain = ch.rx(grid, rnti, tti, n_dmrs).cpu().numpy()
    g_uci = ch.rx_uci(grid, rnti, tti, n_dmrs, tbs, ucfg)
    val = ch.uci_collect(nsf)
    g_uci = g_uci.cpu().numpy()
    for s, c in enumerate(cases):
        r, v = refs[s], val[s]
        n_valid = (M * nd - int(v["Q_prime_ri"])) * Qm
        want, _, _ = T.reference_g_from_plain(g_plain[s], M, nd, Qm, int(v["Q_prime_ack"]), int(v["Q_prime_ri"]))
        assert (g_uci[s, :n_valid] == want).all(), (kw, tbs, c, "stream differs from the rules applied to the plain stream")
        diff = np.abs(g_uci[s, :n_valid].astype(np.int32) - r["g"][:n_valid].astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-2, (kw, tbs, c, diff.max(), (diff != 0).mean())
        na, Lc = c.get("nof_ack", 0), T.cqi_len_of(c)
        same = (v["ack_value"][:na] == r["ack"][:na]).all() and (not na or bool(v["ack_valid"]) == r["ack_valid"])
        same = same and (not c.get("ri_len") or v["ri"] == r["ri"])
        same = same and (not Lc or (bool(v["cqi_crc"]) == r["cqi_crc"] and (not r["cqi_crc"] or (v["cqi_bits"][:Lc] == r["cqi_bits"][:Lc]).all())))
        if not same:
            # only acceptable where the two float paths' soft bits differ in the last bit somewhere in the subframe
            assert (diff != 0).any(), (kw, tbs, c, v, r["ack"], r["ack_valid"], r["ri"], r["cqi_crc"], r["cqi_bits"])
            soft_mism += 1
        nsub += 1
    ch.close()
    it += 1
print(f"soak_uci: {it} random configurations, {nsub} subframes, {soft_mism} decisions differing next to a last-bit difference of the soft bits, none otherwise")
