"""Transport-block level oracle (orc_decode_tb, oracle_port.c) against a composition of reference-pinned pieces, following
sch.c:370-572 literally in Python: per code block dematch with the REFERENCE's srsran_rm_turbo_rx_lut_, the REFERENCE's
generic decoder pass by pass, CRC checks, payload assembly.  No GPU."""
import numpy as np
import pytest

from helpers import make_tb

SB = 18600


def compose_with_reference(ref, e, tbs, Qm, rv, max_iter):
    s = ref.cbsegm(tbs)
    C, K, Ki = s["C"], s["K1"], s["K1_idx"]
    Gp = e.size // Qm
    gamma = Gp % C
    n_e = Qm * (Gp // C)
    data = np.zeros(tbs // 8 + 3 + 768, np.uint8)
    iters = 0
    all_ok = True
    for cb in range(C):
        rlen = K if C == 1 else K - 24
        rp, n_e2 = cb * n_e, n_e
        if cb > C - gamma:
            n_e2 = n_e + Qm
            rp = (C - gamma) * n_e + (cb - (C - gamma)) * n_e2
        soft = np.zeros(SB, np.int16)
        ref.rm_rx(e[rp:rp + n_e2], soft, Ki, rv)
        per_pass = ref.tdec_passes(soft[:3 * K + 12], K, max_iter)
        ok = False
        for p in range(max_iter):
            iters += 1
            out = per_pass[p]
            data[cb * rlen // 8: cb * rlen // 8 + K // 8] = out
            crc = ref.crc24("B", out, K) if C > 1 else ref.crc24("A", out, tbs + 24)
            if crc == 0:
                ok = True
                break
        all_ok &= ok
    par_rx = ref.crc24("A", data, tbs)
    par_tx = (int(data[tbs // 8]) << 16) | (int(data[tbs // 8 + 1]) << 8) | int(data[tbs // 8 + 2])
    ret = 0 if (all_ok and par_rx == par_tx and par_rx != 0) else -1
    return ret, iters, data[:tbs // 8 + 3]


@pytest.mark.parametrize("tbs,Qm,G,rv,sigma", [(6120, 2, 14400, 0, 0.6), (2216, 2, 3000, 0, 0.5), (12960, 4, 28800, 0, 0.75),
                                               (36696, 6, 57600, 0, 0.55), (75376, 6, 86400, 0, 0.35), (12960, 2, 14406, 2, 0.4)])
def test_decode_tb_port_equals_reference_composition(port, ref, tbs, Qm, G, rv, sigma):
    e, expect, s = make_tb(port, tbs, Qm, G, rv, sigma, seed=tbs)
    C = s["C"]
    soft = np.zeros(C * SB, np.int16)
    cbcrc = np.zeros(C, np.uint8)
    data = np.zeros(tbs // 8 + 3 + 768, np.uint8)
    ret, iters = port.decode_tb(e, tbs, Qm, rv, 8, soft, cbcrc, data)
    r2, it2, d2 = compose_with_reference(ref, e, tbs, Qm, rv, 8)
    assert ret == r2 and iters == it2
    assert (data[:tbs // 8 + 3] == d2).all()
    if ret == 0:
        assert (data[:tbs // 8 + 3] == expect).all()
