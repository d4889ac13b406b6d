"""Host half of the control information on the PUSCH (srslte_b200/csrc/uci_host.cu through srsran_b200_uci_decide, no GPU
involved) against the reference's own field decoders (srsran_uci_decode_ack_ri, srsran_uci_decode_cqi_pusch: uci.c) from
oracle/_ref, on random and on coded-plus-noise soft bits."""
import ctypes as C

import numpy as np
import pytest

from oracle import loader

pytestmark = pytest.mark.skipif(not loader.have_ref(), reason="oracle/_ref not built")

BETA_ACK = [2.0, 2.5, 3.125, 4.0, 5.0, 6.25, 8.0, 10.0, 12.625, 15.875, 20.0, 31.0, 50.0, 80.0, 126.0]
BETA_CQI = [None, None, 1.125, 1.25, 1.375, 1.625, 1.75, 2.0, 2.25, 2.5, 2.875, 3.125, 3.5, 4.0, 5.0, 6.25]
BASIS = [0x403, 0x607, 0x749, 0x50d, 0x48f, 0x5d3, 0x755, 0x599, 0x69b, 0x65d, 0x6e5, 0x567, 0x7a9, 0x6ab, 0x4b1, 0x6f3,
         0x277, 0x139, 0x0fb, 0x061, 0x445, 0x60b, 0x591, 0x717, 0x3df, 0x4e3, 0x32d, 0x3af, 0x175, 0x1fd, 0x7ff, 0x001]


@pytest.fixture(scope="module")
def lib():
    from srslte_b200 import _lib

    return _lib.lib()


def decide(lib, Qm, nof_ack=0, ri_len=0, cqi_len=0, Qa=0, Qr=0, Qc=0, ack=None, ri=None, cqi=None):
    from srslte_b200.pusch import UCI_CFG_DTYPE, UCI_VALUE_DTYPE

    cfg = np.zeros(1, UCI_CFG_DTYPE)
    cfg["nof_ack"], cfg["ri_len"], cfg["cqi_len"] = nof_ack, ri_len, cqi_len
    out = np.zeros(1, UCI_VALUE_DTYPE)
    ptr = lambda a: None if a is None else np.ascontiguousarray(a, np.int16).ctypes.data
    keep = [np.ascontiguousarray(a, np.int16) if a is not None else None for a in (ack, ri, cqi)]
    rc = lib.srsran_b200_uci_decide(cfg.ctypes.data, Qm, Qa, Qr, Qc, *(None if a is None else a.ctypes.data for a in keep), out.ctypes.data)
    assert rc == 0
    return out[0]


def field_positions(M, nd, Qm, Qp, is_ri):
    """uci.c:346-393: positions of the field's soft bits in the column-major q order"""
    cols = ([1, 4, 7, 10] if is_ri else [2, 3, 8, 9]) if nd > 10 else ([0, 3, 5, 8] if is_ri else [1, 2, 6, 7])
    pos = []
    for a in range(Qp):
        row, col = M - 1 - a // 4, cols[(3 * a) % 4]
        pos += [row * Qm + M * col * Qm + k for k in range(Qm)]
    return np.array(pos, np.int64)


@pytest.mark.parametrize("mod,L_prb,tbs,cp_ext", [(1, 6, 600, 0), (2, 25, 6200, 0), (3, 100, 75376, 0), (2, 15, 4584, 1)])
def test_ack_and_ri_decisions_follow_the_reference(lib, mod, L_prb, tbs, cp_ext):
    ref = loader.api("ref")
    Qm, M, nd = 2 * mod, 12 * L_prb, 10 if cp_ext else 12
    lk = loader.pusch_link(nof_prb=100, L_prb=L_prb, mod=mod, tbs=tbs, cp_ext=cp_ext)
    rng = np.random.default_rng(1000 * mod + L_prb)
    mism = 0
    for trial in range(60):
        nbits = int(rng.choice([1, 2, 3, 4, 7, 10]))
        is_ri = bool(trial % 3 == 0) and nbits == 1
        beta = BETA_ACK[int(rng.integers(0, 13))]
        scale = float(rng.choice([30.0, 300.0, 3000.0, 20000.0]))  # up to where the int16 accumulators clip and wrap
        q = np.clip(np.round(rng.standard_normal(M * nd * Qm) * scale), -32768, 32767).astype(np.int16)
        # bias the field towards a code word so that `valid` comes out both ways
        c = np.zeros(M * nd * Qm, np.uint8)
        Qp, want_bits, want_valid = ref.uci_decode_ack_ri(lk, q, c, beta, nbits, is_ri)
        pos = field_positions(M, nd, Qm, Qp, is_ri)
        if trial % 2:
            word = int(rng.integers(0, 1 << nbits))
            if nbits > 2:
                code = np.array([bin(word & BASIS[i % 32]).count("1") & 1 for i in range(pos.size)])
                q[pos] = np.clip(q[pos].astype(np.int32) + (2 * code - 1) * int(scale), -32768, 32767).astype(np.int16)
            Qp, want_bits, want_valid = ref.uci_decode_ack_ri(lk, q, c, beta, nbits, is_ri)
        llr = q[pos]
        v = decide(lib, Qm, nof_ack=0 if is_ri else nbits, ri_len=1 if is_ri else 0, Qa=0 if is_ri else Qp, Qr=Qp if is_ri else 0,
                   ack=None if is_ri else llr, ri=llr if is_ri else None)
        if is_ri:
            mism += int(v["ri"] != want_bits[0])
        else:
            mism += int((v["ack_value"][:nbits] != want_bits).any() or bool(v["ack_valid"]) != want_valid)
            assert (v["ack_value"][nbits:] == 2).all()
    assert mism == 0


@pytest.mark.parametrize("mod,L_prb,tbs", [(1, 6, 600), (2, 25, 6200), (3, 100, 75376)])
def test_cqi_decisions_follow_the_reference(lib, mod, L_prb, tbs):
    """Block-coded reports (<= 11 bits) and the CRC-8 + tail-biting convolutional code above 11 bits (the library restates the
    decoder the reference's x86 build selects, viterbi37_avx2_16bit.c, with its quantisation, wrap-around metrics and
    traceback): verdict and bits must agree on every input, also where the noise makes the CRC fail or pass on wrong bits."""
    ref = loader.api("ref")
    Qm, M, nd = 2 * mod, 12 * L_prb, 12
    lk = loader.pusch_link(nof_prb=100, L_prb=L_prb, mod=mod, tbs=tbs)
    rng = np.random.default_rng(77 + mod)
    short_mism = long_mism = long_ok = long_total = 0
    for trial in range(80):
        cqi_len = int(rng.choice([1, 4, 6, 10, 11, 12, 18, 22, 30, 44]))
        beta = BETA_CQI[int(rng.integers(2, 16))]
        sigma = float(rng.choice([0.3, 0.8, 1.3, 2.0]))
        # Q' from the reference itself (it only depends on the sizes); then code + noise of that length
        Qp, _, _ = ref.uci_decode_cqi(lk, np.zeros(M * nd * Qm, np.int16), beta, 0, cqi_len)
        n = Qp * Qm
        bits = rng.integers(0, 2, cqi_len).astype(np.uint8)
        if cqi_len <= 11:
            word = sum(int(b) << i for i, b in enumerate(bits))
            code = np.array([bin(word & BASIS[i % 32]).count("1") & 1 for i in range(n)])
        else:
            code = conv_code(bits, n)
        amp = 400.0
        q = np.clip(np.round((2.0 * code - 1.0) * amp + rng.standard_normal(n) * amp * sigma), -32768, 32767).astype(np.int16)
        _, want_bits, want_crc = ref.uci_decode_cqi(lk, q, beta, 0, cqi_len)
        v = decide(lib, Qm, cqi_len=cqi_len, Qc=Qp, cqi=q)
        if cqi_len <= 11:
            short_mism += int((v["cqi_bits"][:cqi_len] != want_bits).any() or not v["cqi_crc"])
        else:
            long_total += 1
            same = bool(v["cqi_crc"]) == want_crc and (not want_crc or (v["cqi_bits"][:cqi_len] == want_bits).all())
            long_mism += int(not same)
            if want_crc:
                long_ok += 1
                assert (want_bits == bits).all() or sigma >= 1.3  # (a passing CRC-8 on wrong bits needs heavy noise)
    assert short_mism == 0
    assert long_ok >= long_total // 3 and long_total - long_ok >= 3, (long_ok, long_total)  # both verdicts are exercised
    assert long_mism == 0, (long_mism, long_total)


def conv_code(bits, n_out):
    """TS 36.212 5.2.2.6.4 for more than 11 bits: CRC-8, rate 1/3 tail-biting convolutional code (5.1.3.1), rate matching (5.1.4.2)"""
    r = 0
    for b in bits:
        r = ((r << 1) | int(b))
        if r & 0x100:
            r ^= 0x19B
    for _ in range(8):
        r <<= 1
        if r & 0x100:
            r ^= 0x19B
    c = np.concatenate([bits, [(r >> (7 - i)) & 1 for i in range(8)]]).astype(np.int64)
    F = c.size
    polys = [0o133, 0o171, 0o165]
    d = np.zeros((3, F), np.int64)
    state = [int(c[F - 1 - i]) for i in range(6)]  # tail biting: the register starts with the last six bits
    for t in range(F):
        reg = [int(c[t])] + state  # reg[i] = input i steps ago
        for s, g in enumerate(polys):
            taps = [(g >> (6 - i)) & 1 for i in range(7)]
            d[s, t] = sum(reg[i] & taps[i] for i in range(7)) & 1
        state = reg[:6]
    perm = [1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31, 0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30]
    nrows = (F - 1) // 32 + 1
    ndummy = nrows * 32 - F
    w = []
    for s in range(3):
        y = np.concatenate([np.full(ndummy, -1), d[s]]).reshape(nrows, 32)
        w.append(y[:, perm].T.reshape(-1))
    w = np.concatenate(w)
    out, j = [], 0
    while len(out) < n_out:
        if w[j] >= 0:
            out.append(int(w[j]))
        j = (j + 1) % w.size
    return np.array(out)
