#!/usr/bin/env python3
"""Randomised differential soak of the PUSCH chain stages against the oracle port (developer tool).
usage: python tools/soak_pusch.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import loader  # noqa: E402
from srslte_b200.pusch import PuschChain  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
port = loader.api("port")
valid = []
for _L in range(1, 101):  # srsran_dft_precoding_valid_prb: 12 L = 2^a 3^b 5^c
    _n = _L
    for _f in (2, 3, 5):
        while _n % _f == 0:
            _n //= _f
    if _n == 1:
        valid.append(_L)

def rel(a, b):
    return float(np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)) / max(np.linalg.norm(b.astype(np.complex128)), 1e-30))


t0, n = time.time(), 0
while time.time() - t0 < budget:
    L = int(rng.choice(valid))
    cell_prb = int(rng.integers(L, 111)) if L < 100 else int(rng.integers(100, 111))
    n_prb = int(rng.integers(0, cell_prb - L + 1))
    mod = int(rng.integers(1, 4))
    shift = int(rng.choice([0, 2, 4]))
    cp_ext = bool(rng.integers(0, 4) == 0)
    kw = dict(cell_id=int(rng.integers(0, 504)), cell_nof_prb=cell_prb, cp_ext=cp_ext, L_prb=L, n_prb=n_prb, mod=mod, llr_shift=shift,
              cyclic_shift=int(rng.integers(0, 8)), delta_ss=int(rng.integers(0, 30)), group_hopping=bool(rng.integers(0, 2)),
              sequence_hopping=bool(rng.integers(0, 2)), shortened=bool(rng.integers(0, 3) == 0))
    ch = PuschChain(**kw)
    nsf = int(rng.integers(1, 6))
    rnti = rng.integers(0, 65536, nsf).astype(np.uint32)
    tti = rng.integers(0, 10240, nsf).astype(np.uint32)
    n_dmrs = rng.integers(0, 8, nsf).astype(np.uint32)
    grid = (rng.standard_normal((nsf, ch.nsym, ch.R)) + 1j * rng.standard_normal((nsf, ch.nsym, ch.R))).astype(np.complex64)
    g_t = torch.from_numpy(grid).cuda()
    ce, meas = ch.chest(g_t, tti, n_dmrs)
    d = ch.equalize_deprecode(g_t, ce, meas)
    g = ch.demod_descramble(d, rnti, tti)
    torch.cuda.synchronize()
    ce, meas, d, g = ce.cpu().numpy(), meas.cpu().numpy(), d.cpu().numpy(), g.cpu().numpy()
    off, half = 12 * n_prb, ch.nsym // 2
    data_syms = [l for l in range(ch.nsym) if l not in (half - 4, ch.nsym - 4) and not (kw["shortened"] and l == ch.nsym - 1)]
    for s in range(nsf):
        lk = loader.pusch_link(kw["cell_id"], cell_prb, int(cp_ext), kw["cyclic_shift"], kw["delta_ss"], int(kw["group_hopping"]),
                               int(kw["sequence_hopping"]), int(rnti[s]), int(tti[s]), L, n_prb, mod, 0, 0, int(n_dmrs[s]), 8)
        dm = port.dmrs_pusch_gen(lk)
        assert np.abs(ch.dmrs(int(tti[s] % 10), int(n_dmrs[s])).reshape(-1) - dm).max() < 1e-6, ("dmrs", kw)
        want_ce, want_meas = port.chest_ul_pusch(lk, grid[s], dm)
        want_ce = want_ce.reshape(ch.nsym, ch.R)
        for slot in range(2):
            assert rel(ce[s, slot], want_ce[(slot + 1) * half - 4, off:off + ch.M]) < 1e-5, ("ce", kw)
        assert abs(meas[s, 0] - want_meas[0]) <= 3e-4 * abs(want_meas[0]), ("noise", kw)
        y = np.concatenate([grid[s, l, off:off + ch.M] for l in data_syms])
        h = np.concatenate([ce[s, l // half] for l in data_syms])
        want_d = port.dft_precoding(port.predecoding_single(y, h, float(meas[s, 0])), L, False)
        assert rel(d[s], want_d) < 1e-4, ("d", kw, rel(d[s], want_d))
        q = port.pusch_seq_apply_s(port.demod_s(mod, d[s]) >> shift, int(rnti[s]), 2 * int(tti[s] % 10), kw["cell_id"])
        assert (g[s] == port.ulsch_deinterleave(q, 2 * mod, ch.nd)).all(), ("g", kw)
    ch.close()
    n += 1
print(f"soak ok: {n} random PUSCH configurations, {time.time()-t0:.0f} s")
