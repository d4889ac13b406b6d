"""Control information multiplexed into the PUSCH (HARQ-ACK, RI, CQI: TS 36.212 5.2.2.6-5.2.2.8) through the batched chain,
against the reference's own transmitter and receiver (srsran_pusch_encode / srsran_pusch_decode with cfg->uci_cfg set:
sch.c:1022-1195, uci.c) from oracle/_ref."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def link_of(loader, kw, rnti, tti, n_dmrs, tbs):
    return loader.pusch_link(kw["cell_id"], kw["cell_nof_prb"], int(kw.get("cp_ext", False)), kw.get("cyclic_shift", 0), kw.get("delta_ss", 0),
                             0, 0, rnti, tti, kw["L_prb"], kw["n_prb"], kw["mod"], tbs, 0, n_dmrs, 8, int(kw.get("shortened", False)))


# harness parameter block (oracle/loader.py: pusch_uci) and the cqi payload length it stands for
UCI_CASES = [
    dict(),
    dict(nof_ack=1, ack_bits=1),
    dict(nof_ack=1, ack_bits=0, I_offset_ack=12),
    dict(nof_ack=2, ack_bits=0b10),
    dict(nof_ack=2, ack_bits=0b01, ri_len=1, ri=1),
    dict(ri_len=1, ri=0, I_offset_ri=9),
    dict(nof_ack=4, ack_bits=0b1011, cqi_kind=1, cqi_wb=11),
    dict(nof_ack=7, ack_bits=0b1010011, I_offset_ack=11),
    dict(cqi_kind=2, cqi_wb=6, cqi_sb=2, I_offset_cqi=10),
    dict(nof_ack=2, ack_bits=0b11, ri_len=1, ri=1, cqi_kind=3, cqi_N=7, cqi_wb=9, cqi_sb=0x2D5A, I_offset_cqi=12),
    dict(ri_len=1, ri=1, cqi_kind=3, cqi_N=13, cqi_wb=3, cqi_sb=0x2A5F0C3, I_offset_ri=12, I_offset_cqi=15),
    dict(nof_ack=10, ack_bits=0b1100101101, ri_len=1, ri=1, cqi_kind=1, cqi_wb=5, I_offset_ack=14, I_offset_ri=12),
]


def cqi_len_of(c):
    k = c.get("cqi_kind", 0)
    return {0: 0, 1: 4, 2: 6, 3: 4 + 2 * c.get("cqi_N", 0)}[k]


def reference_g_from_plain(g_plain, M, nd, Qm, Q_ack, Q_ri):
    """sch.c:993-1119 restated on the de-interleaved stream of a subframe WITHOUT control information: back to the interleaver
    matrix, zero the HARQ-ACK positions, leave the RI positions out, and the scatter quirk at element 0."""
    norm = nd > 10
    ri_cols = [1, 4, 7, 10] if norm else [0, 3, 5, 8]
    ack_cols = [2, 3, 8, 9] if norm else [1, 2, 6, 7]
    mat = g_plain.reshape(M, nd, Qm).copy()  # [row j][col i][k]: g index (j*nd + i)*Qm + k
    ack = np.zeros((M, nd), bool)
    ri = np.zeros((M, nd), bool)
    for a in range(Q_ack):
        ack[M - 1 - a // 4, ack_cols[(3 * a) % 4]] = True
    for r in range(Q_ri):
        ri[M - 1 - r // 4, ri_cols[(3 * r) % 4]] = True
    ack_llr = np.array([mat[M - 1 - a // 4, ack_cols[(3 * a) % 4]] for a in range(Q_ack)], np.int16).reshape(-1)
    ri_llr = np.array([mat[M - 1 - r // 4, ri_cols[(3 * r) % 4]] for r in range(Q_ri)], np.int16).reshape(-1)
    mat[ack] = 0
    out = mat[~ri].reshape(-1).copy()
    if Q_ri:
        # the largest RI position of the column-major q order: last RI column, bottom row
        cols = sorted({ri_cols[(3 * r) % 4] for r in range(Q_ri)})
        c = cols[-1]
        rows = [M - 1 - r // 4 for r in range(Q_ri) if ri_cols[(3 * r) % 4] == c]
        out[0] = mat[max(rows), c, Qm - 1]
    return out, ack_llr, ri_llr


CONFIGS = [
    (dict(cell_id=3, cell_nof_prb=25, L_prb=10, n_prb=5, mod=2), 2536),
    (dict(cell_id=1, cell_nof_prb=100, L_prb=100, n_prb=0, mod=3), 75376),
    (dict(cell_id=42, cell_nof_prb=15, L_prb=15, n_prb=0, mod=1, cp_ext=True, delta_ss=7), 1544),
    (dict(cell_id=200, cell_nof_prb=50, L_prb=4, n_prb=30, mod=2, cyclic_shift=4), 1000),
    (dict(cell_id=5, cell_nof_prb=6, L_prb=2, n_prb=1, mod=1), 208),   # 24 subcarriers: the fields fill whole columns of the matrix
    (dict(cell_id=61, cell_nof_prb=50, L_prb=20, n_prb=8, mod=2, shortened=True), 5160),               # SRS subframe: 11 columns
    (dict(cell_id=42, cell_nof_prb=15, L_prb=12, n_prb=1, mod=1, cp_ext=True, shortened=True), 1000),  # ... and 9
]


@pytest.mark.parametrize("kw,tbs", CONFIGS)
def test_uci_chain_against_the_reference_link(ref, port, kw, tbs):
    import torch
    from oracle import loader
    from srslte_b200.pusch import PuschChain, uci_cfg
    from srslte_b200.sch import SOFTBUFFER_SIZE, SchDecoder

    ch0 = PuschChain(llr_shift=0, **kw)
    Qm = 2 * kw["mod"]
    nsf = len(UCI_CASES)
    rng = np.random.default_rng(tbs)
    rnti = rng.integers(1, 65000, nsf).astype(np.uint32)
    tti = rng.integers(0, 10240, nsf).astype(np.uint32)
    n_dmrs = rng.integers(0, 8, nsf).astype(np.uint32)
    ucfg = uci_cfg(nsf)
    grids, datas, refs = [], [], []
    for s, c in enumerate(UCI_CASES):
        u = loader.pusch_uci(**c)
        ucfg[s] = (c.get("nof_ack", 0), c.get("ri_len", 0), cqi_len_of(c), u[8], u[9], u[10])
        lk = link_of(loader, kw, int(rnti[s]), int(tti[s]), int(n_dmrs[s]), tbs)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        tx = ref.pusch_encode_uci(lk, u, data)
        h = np.complex64((0.7 + 0.05 * s) * np.exp(1j * (0.3 + s)))
        noise = (rng.standard_normal(tx.shape) + 1j * rng.standard_normal(tx.shape)).astype(np.complex64) * np.float32(0.012)
        rxg = (tx * h + noise).astype(np.complex64)
        grids.append(rxg)
        datas.append(data)
        refs.append(ref.pusch_decode_uci(lk, u, rxg))
        assert refs[-1]["ret"] == 0, s
        assert not refs[-1]["crc"] or (refs[-1]["data"] == data).all(), s
    # (64QAM at rate 0.87 and the 2-PRB allocation do not survive the heaviest puncturing of the list: there the verdicts must agree)
    assert sum(r["crc"] for r in refs) >= nsf - 3
    grid = torch.from_numpy(np.stack(grids)).cuda()
    g_plain = ch0.rx(grid, rnti, tti, n_dmrs).cpu().numpy()
    g_uci = ch0.rx_uci(grid, rnti, tti, n_dmrs, tbs, ucfg)
    val = ch0.uci_collect(nsf)
    g_uci = g_uci.cpu().numpy()
    M, nd = ch0.M, ch0.nd
    for s, c in enumerate(UCI_CASES):
        r, v = refs[s], val[s]
        geo = ch0.uci_geometry(tbs, ucfg[s:s + 1])
        assert (geo["Q_prime_ack"], geo["Q_prime_ri"], geo["Q_prime_cqi"]) == (v["Q_prime_ack"], v["Q_prime_ri"], v["Q_prime_cqi"])
        if c.get("nof_ack"):
            seg = port.cbsegm(tbs)
            K = seg["C1"] * seg["K1"] + seg["C2"] * seg["K2"]
            beta = [2.0, 2.5, 3.125, 4.0, 5.0, 6.25, 8.0, 10.0, 12.625, 15.875, 20.0, 31.0, 50.0, 80.0, 126.0][int(ucfg[s]["I_offset_ack"])]
            assert v["Q_prime_ack"] == ref.qprime_ack(kw["L_prb"], nd, K, c["nof_ack"], beta)
        n_valid = (M * nd - int(v["Q_prime_ri"])) * Qm
        assert v["nof_e_bits"] == n_valid - int(v["Q_prime_cqi"]) * Qm and v["e_offset"] == int(v["Q_prime_cqi"]) * Qm
        # exact: the kernel's stream equals the reference's de-interleaver rules applied to the plain stream of the same symbols
        want, _, _ = reference_g_from_plain(g_plain[s], M, nd, Qm, int(v["Q_prime_ack"]), int(v["Q_prime_ri"]))
        assert (g_uci[s, :n_valid] == want).all(), (s, int((g_uci[s, :n_valid] != want).sum()))
        # and it is the reference receiver's q->g up to the float path's last-bit differences
        diff = np.abs(g_uci[s, :n_valid].astype(np.int32) - r["g"][:n_valid].astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 3e-3, (s, diff.max(), (diff != 0).mean())
        # decided values
        na = c.get("nof_ack", 0)
        assert (v["ack_value"][:na] == r["ack"][:na]).all() and (v["ack_value"][na:] == 2).all(), (s, v["ack_value"], r["ack"])
        if na:
            assert bool(v["ack_valid"]) == r["ack_valid"] and r["ack_valid"]
            assert (v["ack_value"][:na] == [(c["ack_bits"] >> a) & 1 for a in range(na)]).all()
        if c.get("ri_len"):
            assert v["ri"] == r["ri"] == c["ri"]
        L = cqi_len_of(c)
        if L:
            assert bool(v["cqi_crc"]) == r["cqi_crc"] and r["cqi_crc"]
            assert (v["cqi_bits"][:L] == r["cqi_bits"][:L]).all(), (s, v["cqi_bits"][:L], r["cqi_bits"][:L])
    ch0.close()
    # the transport blocks decode from what the control information leaves (soft bits scaled into the decoder's envelope);
    # the decided control values do not depend on the shift
    ch4 = PuschChain(llr_shift=4 if kw["mod"] == 3 else 3 if kw["mod"] == 2 else 1, **kw)
    g = ch4.rx_uci(grid, rnti, tti, n_dmrs, tbs, ucfg)
    val4 = ch4.uci_collect(nsf)
    assert val4.tobytes() == val.tobytes()
    sch = SchDecoder(0, 8)
    seg = port.cbsegm(tbs)
    stride = (tbs // 8 + 3 + 768 + 15) // 16 * 16
    soft = np.zeros((nsf, seg["C"] * SOFTBUFFER_SIZE), np.int16)
    out = np.zeros((nsf, stride), np.uint8)
    rc, res = sch.decode(g.cpu().numpy().reshape(-1), soft.reshape(-1), out.reshape(-1),
                         [dict(tbs=tbs, Qm=Qm, rv=0, nof_e_bits=int(val[s]["nof_e_bits"]), e_offset=s * ch4.nof_bits + int(val[s]["e_offset"]),
                               soft_offset=s * soft.shape[1], data_offset=s * stride) for s in range(nsf)])
    assert rc == 0
    for s in range(nsf):
        assert (res[s]["result"] == 0) == refs[s]["crc"], s
        if refs[s]["crc"]:
            assert (out[s, :tbs // 8] == datas[s]).all()
    sch.close()
    ch4.close()


def test_uci_in_noise_matches_the_reference_decisions(ref):
    """Low SNR: the decisions (including `valid` going false and the CRC-8 of the long CQI failing) still follow the reference,
    because they are taken on the same soft bits.  The soft bits of the two float paths differ in the last bit at a few
    positions, so a decision a hair from its threshold may differ: allow a small number of mismatches."""
    import torch
    from oracle import loader
    from srslte_b200.pusch import PuschChain, uci_cfg

    kw, tbs = dict(cell_id=11, cell_nof_prb=25, L_prb=8, n_prb=2, mod=1), 1096
    ch = PuschChain(llr_shift=0, **kw)
    cases = [dict(nof_ack=1, ack_bits=1, I_offset_ack=2), dict(nof_ack=2, ack_bits=2, I_offset_ack=0), dict(nof_ack=5, ack_bits=0b10110, I_offset_ack=3),
             dict(ri_len=1, ri=1, I_offset_ri=0, cqi_kind=1, cqi_wb=9, I_offset_cqi=2), dict(cqi_kind=3, cqi_N=9, cqi_wb=7, cqi_sb=0x155AA, I_offset_cqi=2)]
    nrep = 24
    nsf = len(cases) * nrep
    rng = np.random.default_rng(5)
    ucfg = uci_cfg(nsf)
    rnti = rng.integers(1, 65000, nsf).astype(np.uint32)
    tti = rng.integers(0, 10240, nsf).astype(np.uint32)
    grids, refs, cs = [], [], []
    for s in range(nsf):
        c = cases[s % len(cases)]
        u = loader.pusch_uci(**c)
        ucfg[s] = (c.get("nof_ack", 0), c.get("ri_len", 0), cqi_len_of(c), u[8], u[9], u[10])
        lk = link_of(loader, kw, int(rnti[s]), int(tti[s]), 0, tbs)
        tx = ref.pusch_encode_uci(lk, u, rng.integers(0, 256, tbs // 8, dtype=np.uint8))
        sigma = np.float32(0.40 + 0.30 * rng.random())
        rxg = (tx + (rng.standard_normal(tx.shape) + 1j * rng.standard_normal(tx.shape)).astype(np.complex64) * sigma).astype(np.complex64)
        grids.append(rxg)
        refs.append(ref.pusch_decode_uci(lk, u, rxg))
        cs.append(c)
    ch.rx_uci(torch.from_numpy(np.stack(grids)).cuda(), rnti, tti, None, tbs, ucfg)
    val = ch.uci_collect(nsf)
    mism, invalid, crc_fail = 0, 0, 0
    for s in range(nsf):
        c, r, v = cs[s], refs[s], val[s]
        na, L = c.get("nof_ack", 0), cqi_len_of(c)
        same = (v["ack_value"][:na] == r["ack"][:na]).all() and (not na or bool(v["ack_valid"]) == r["ack_valid"])
        same = same and (not c.get("ri_len") or v["ri"] == r["ri"])
        same = same and (not L or (bool(v["cqi_crc"]) == r["cqi_crc"] and (not r["cqi_crc"] or (v["cqi_bits"][:L] == r["cqi_bits"][:L]).all())))
        mism += 0 if same else 1
        invalid += 1 if (na and not r["ack_valid"]) else 0
        crc_fail += 1 if (L > 11 and not r["cqi_crc"]) else 0
    assert invalid > 0 and crc_fail > 0, (invalid, crc_fail)  # the noise level does exercise the negative verdicts
    assert mism <= 2, mism
    ch.close()


def test_uci_rejects_what_the_reference_does_not_carry():
    from srslte_b200.pusch import PuschChain, uci_cfg

    ch = PuschChain(cell_id=1, cell_nof_prb=25, L_prb=4, n_prb=0, mod=1)
    for bad in (uci_cfg(1, nof_ack=11), uci_cfg(1, ri_len=2), uci_cfg(1, cqi_len=60)):
        with pytest.raises(RuntimeError):
            ch.uci_geometry(1000, bad)
    # so much control information that no symbol is left for the transport block: accepted like the reference does (the block
    # then fails its CRC), with an empty UL-SCH span
    geo = ch.uci_geometry(40, uci_cfg(1, cqi_len=40, I_offset_cqi=15))
    assert geo["nof_e_bits"] == 0 and geo["Q_prime_cqi"] == 12 * 48
    with pytest.raises(RuntimeError):  # nothing pending
        ch.uci_collect(3)
    ch.close()


def test_enb_ul_one_call_with_uci(ref):
    """srsran_b200_enb_ul_pusch_uci_batch: time samples in, transport blocks and control information out."""
    from oracle import loader
    from srslte_b200.pusch import EnbUl, uci_cfg
    from srslte_b200 import synth_pusch as sp

    kw, tbs = dict(cell_id=21, cell_nof_prb=25, L_prb=25, n_prb=0, mod=2), 6200
    cases = [UCI_CASES[i] for i in (0, 1, 4, 6, 9, 11)]
    nsf = 2 * len(cases)
    rng = np.random.default_rng(17)
    rnti = rng.integers(1, 65000, nsf).astype(np.uint32)
    tti = rng.integers(0, 10240, nsf).astype(np.uint32)
    ucfg = uci_cfg(nsf)
    grids, datas, cs = [], [], []
    for s in range(nsf):
        c = cases[s % len(cases)]
        u = loader.pusch_uci(**c)
        ucfg[s] = (c.get("nof_ack", 0), c.get("ri_len", 0), cqi_len_of(c), u[8], u[9], u[10])
        lk = link_of(loader, kw, int(rnti[s]), int(tti[s]), 0, tbs)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        grids.append(ref.pusch_encode_uci(lk, u, data))
        datas.append(data)
        cs.append(c)
    enb = EnbUl(cell_id=21, nof_prb=25, tbs=tbs, mod=2, llr_shift=3)
    iq = sp.ofdm_modulate(np.stack(grids), enb.sf_sz // 15)
    iq = (iq + (rng.standard_normal(iq.shape) + 1j * rng.standard_normal(iq.shape)).astype(np.complex64) * np.float32(0.002)).astype(np.complex64)
    data, res, val = enb.run(iq, rnti, tti, uci=ucfg)
    for s in range(nsf):
        c = cs[s]
        assert res[s]["crc_ok"] == 1, s
        assert (data[s, :tbs // 8] == datas[s]).all(), s
        na = c.get("nof_ack", 0)
        assert (val[s]["ack_value"][:na] == [(c["ack_bits"] >> a) & 1 for a in range(na)]).all() and (not na or val[s]["ack_valid"])
        if c.get("ri_len"):
            assert val[s]["ri"] == c["ri"]
        if cqi_len_of(c):
            assert val[s]["cqi_crc"] == 1
            assert (val[s]["cqi_bits"][:4] == [(c["cqi_wb"] >> (3 - b)) & 1 for b in range(4)]).all()
    # the plain entry still works on the same object afterwards (the per-subframe spans are reset)
    data2, res2 = enb.run(iq[:1], rnti[:1], tti[:1])
    assert res2[0]["crc_ok"] == 1 and (data2[0, :tbs // 8] == datas[0]).all()
    enb.close()
