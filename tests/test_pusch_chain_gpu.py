"""PUSCH receive chain between OFDM and de-matching on the GPU: DMRS, channel estimation, equaliser +
transform de-precoding, soft demapping + descrambling + UL-SCH de-interleaving, against the oracle stage by stage and, where the
reference build is available, end to end against srsran_pusch_encode / srsran_chest_ul_estimate_pusch / srsran_pusch_decode.
-m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# float stages: relative L2 error bound (the north star's 1e-4 for FFT-type outputs)
TOL = 1e-4


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.complex128) - b.astype(np.complex128)) / max(np.linalg.norm(b.astype(np.complex128)), 1e-30))


def link_of(loader, ch, rnti=0, tti=0, n_dmrs=0, tbs=0):
    c = ch.cfg
    return loader.pusch_link(c.cell_id, c.cell_nof_prb, c.cp_ext, c.dmrs_cyclic_shift, c.dmrs_delta_ss, c.group_hopping_en,
                             c.sequence_hopping_en, rnti, tti, c.L_prb, c.n_prb, c.modulation, tbs, 0, n_dmrs, 8, c.shortened)


CONFIGS = [
    dict(cell_id=1, cell_nof_prb=100, L_prb=100, n_prb=0, mod=3),
    dict(cell_id=301, cell_nof_prb=50, L_prb=25, n_prb=10, mod=2, cyclic_shift=5, delta_ss=11),
    dict(cell_id=77, cell_nof_prb=25, L_prb=6, n_prb=19, mod=1, group_hopping=True),
    dict(cell_id=503, cell_nof_prb=100, L_prb=81, n_prb=3, mod=3, sequence_hopping=True, delta_ss=29),
    dict(cell_id=9, cell_nof_prb=6, L_prb=3, n_prb=2, mod=2),
    dict(cell_id=42, cell_nof_prb=15, L_prb=5, n_prb=0, mod=3, cp_ext=True),
    # 1 and 2 PRB: table base sequences (TS 36.211 5.5.1.2), 12- and 24-point transforms
    dict(cell_id=12, cell_nof_prb=6, L_prb=1, n_prb=4, mod=1),
    dict(cell_id=250, cell_nof_prb=25, L_prb=2, n_prb=11, mod=2, group_hopping=True, delta_ss=17),
    dict(cell_id=88, cell_nof_prb=6, L_prb=2, n_prb=2, mod=1, cp_ext=True, cyclic_shift=2),
]


@pytest.mark.parametrize("kw", CONFIGS)
def test_dmrs_table_matches_oracle(port, kw):
    from oracle import loader
    from srslte_b200.pusch import PuschChain

    ch = PuschChain(**kw)
    for sf_idx, n_dmrs in ((0, 0), (3, 1), (7, 5), (9, 7)):
        want = port.dmrs_pusch_gen(link_of(loader, ch, tti=sf_idx, n_dmrs=n_dmrs))
        got = ch.dmrs(sf_idx, n_dmrs).reshape(-1)
        assert np.abs(got - want).max() < 1e-6, (kw, sf_idx, n_dmrs)
    ch.close()


@pytest.mark.parametrize("kw", CONFIGS)
def test_chest_matches_oracle(port, kw):
    import torch
    from oracle import loader
    from srslte_b200.pusch import PuschChain

    ch = PuschChain(**kw)
    rng = np.random.default_rng(5)
    nsf = 5
    tti = np.array([0, 3, 14, 7, 29], np.uint32)
    n_dmrs = np.array([0, 2, 7, 1, 4], np.uint32)
    grid = (rng.standard_normal((nsf, ch.nsym, ch.R)) + 1j * rng.standard_normal((nsf, ch.nsym, ch.R))).astype(np.complex64) * 0.05
    off = 12 * kw["n_prb"]
    # a smooth channel over the allocation plus noise on the two DMRS symbols
    for s in range(nsf):
        h = (0.7 + 0.2 * s) * np.exp(1j * (0.3 * s + 2 * np.pi * 0.0007 * np.arange(ch.M)))
        r = ch.dmrs(int(tti[s] % 10), int(n_dmrs[s]))
        for slot in range(2):
            l = (slot + 1) * (ch.nsym // 2) - 4
            grid[s, l, off:off + ch.M] += (r[slot] * h * np.exp(1j * 0.05 * slot)).astype(np.complex64)
    ce, meas = ch.chest(torch.from_numpy(grid).cuda(), tti, n_dmrs)
    torch.cuda.synchronize()
    ce, meas = ce.cpu().numpy(), meas.cpu().numpy()
    for s in range(nsf):
        lk = link_of(loader, ch, tti=int(tti[s]), n_dmrs=int(n_dmrs[s]))
        want_ce, want_meas = port.chest_ul_pusch(lk, grid[s], port.dmrs_pusch_gen(lk))
        want_ce = want_ce.reshape(ch.nsym, ch.R)
        for slot in range(2):
            l = (slot + 1) * (ch.nsym // 2) - 4
            assert rel(ce[s, slot], want_ce[l, off:off + ch.M]) < 1e-5
            # the reference copies the slot's estimate to every symbol of the slot (chest_ul.c:246-259)
            assert (want_ce[slot * (ch.nsym // 2), off:off + ch.M] == want_ce[l, off:off + ch.M]).all()
        assert abs(meas[s, 0] - want_meas[0]) <= 2e-4 * abs(want_meas[0])
        assert abs(meas[s, 1] - want_meas[1]) <= 4e-4 * abs(want_meas[1])
        assert abs(meas[s, 2] - want_meas[2]) <= 1e-3 * max(abs(want_meas[2]), 1.0)
    ch.close()


@pytest.mark.parametrize("kw", CONFIGS)
def test_equalize_deprecode_matches_oracle(port, kw):
    import torch
    from srslte_b200.pusch import PuschChain

    ch = PuschChain(**kw)
    rng = np.random.default_rng(6)
    nsf = 3
    grid = (rng.standard_normal((nsf, ch.nsym, ch.R)) + 1j * rng.standard_normal((nsf, ch.nsym, ch.R))).astype(np.complex64)
    ce = (rng.standard_normal((nsf, 2, ch.M)) + 1j * rng.standard_normal((nsf, 2, ch.M))).astype(np.complex64)
    meas = np.zeros((nsf, 4), np.float32)
    meas[:, 0] = [0.0, 0.01, 0.3]
    d = ch.equalize_deprecode(torch.from_numpy(grid).cuda(), torch.from_numpy(ce).cuda(), torch.from_numpy(meas).cuda())
    torch.cuda.synchronize()
    d = d.cpu().numpy()
    off = 12 * kw["n_prb"]
    half = ch.nsym // 2
    data_syms = [l for l in range(ch.nsym) if l not in (half - 4, ch.nsym - 4)]
    for s in range(nsf):
        y = np.concatenate([grid[s, l, off:off + ch.M] for l in data_syms])
        h = np.concatenate([ce[s, l // half] for l in data_syms])
        z = port.predecoding_single(y, h, float(meas[s, 0]))
        want = port.dft_precoding(z, kw["L_prb"], False)
        assert rel(d[s], want) < TOL, (kw, s, rel(d[s], want))
    ch.close()


@pytest.mark.parametrize("kw", CONFIGS)
@pytest.mark.parametrize("shift", [0, 4])
def test_demod_descramble_deinterleave_is_bit_exact(port, kw, shift):
    import torch
    from srslte_b200.pusch import PuschChain

    ch = PuschChain(llr_shift=shift, **kw)
    rng = np.random.default_rng(7)
    nsf = 4
    rnti = np.array([62, 0xFFFF, 1, 0x1234], np.uint32)
    tti = np.array([0, 9, 13, 5], np.uint32)
    d = (rng.standard_normal((nsf, ch.nof_re)) + 1j * rng.standard_normal((nsf, ch.nof_re))).astype(np.complex64)
    d[0, :7] *= 100.0  # saturating values
    g = ch.demod_descramble(torch.from_numpy(d).cuda(), rnti, tti)
    torch.cuda.synchronize()
    g = g.cpu().numpy()
    Qm = 2 * kw["mod"]
    for s in range(nsf):
        q = port.demod_s(kw["mod"], d[s]) >> shift
        q = port.pusch_seq_apply_s(q, int(rnti[s]), 2 * int(tti[s] % 10), kw["cell_id"])
        want = port.ulsch_deinterleave(q, Qm, ch.nd)
        assert (g[s] == want).all(), (kw, shift, s, int((g[s] != want).sum()))
    ch.close()


def test_unsupported_configurations_fail_cleanly():
    from srslte_b200.pusch import PuschChain

    for kw in (dict(L_prb=7, cell_nof_prb=25), dict(L_prb=0, cell_nof_prb=6), dict(L_prb=50, n_prb=60), dict(mod=4)):
        with pytest.raises(RuntimeError):
            PuschChain(**kw)


@pytest.mark.parametrize("kw,tbs", [(dict(cell_id=1, cell_nof_prb=100, L_prb=100, n_prb=0, mod=3), 75376),
                                    (dict(cell_id=150, cell_nof_prb=50, L_prb=24, n_prb=13, mod=2, cyclic_shift=3), 9912),
                                    (dict(cell_id=7, cell_nof_prb=25, L_prb=10, n_prb=5, mod=1), 1544),
                                    (dict(cell_id=42, cell_nof_prb=15, L_prb=15, n_prb=0, mod=2, cp_ext=True, delta_ss=7), 4584),
                                    (dict(cell_id=333, cell_nof_prb=75, L_prb=72, n_prb=2, mod=3, group_hopping=True), 46888),
                                    (dict(cell_id=12, cell_nof_prb=6, L_prb=1, n_prb=4, mod=1), 136),
                                    (dict(cell_id=250, cell_nof_prb=25, L_prb=2, n_prb=11, mod=2, cyclic_shift=6), 328),
                                    # SRS subframes: the last symbol is not PUSCH (11 / 9 data symbols)
                                    (dict(cell_id=61, cell_nof_prb=50, L_prb=20, n_prb=8, mod=3, shortened=True), 11448),
                                    (dict(cell_id=42, cell_nof_prb=15, L_prb=12, n_prb=1, mod=1, cp_ext=True, shortened=True), 1000)])
def test_chain_against_the_reference_link(ref, port, kw, tbs):
    """Reference transmitter (srsran_pusch_encode + DMRS) -> flat fading + AWGN -> GPU chain, compared with the buffers of the
    reference receiver (chest + srsran_pusch_decode) and decoded down to the transport block."""
    import torch
    from oracle import loader
    from srslte_b200.pusch import PuschChain
    from srslte_b200.sch import SOFTBUFFER_SIZE, SchDecoder

    ch0 = PuschChain(llr_shift=0, **kw)
    nsf = 3
    rng = np.random.default_rng(8)
    rnti = np.array([62, 4097, 65000], np.uint32)
    tti = np.array([3, 18, 4], np.uint32)
    n_dmrs = np.array([0, 6, 3], np.uint32)
    grids, datas, refs = [], [], []
    for s in range(nsf):
        lk = link_of(loader, ch0, rnti=int(rnti[s]), tti=int(tti[s]), n_dmrs=int(n_dmrs[s]), tbs=tbs)
        data = rng.integers(0, 256, tbs // 8, dtype=np.uint8)
        tx = ref.pusch_encode(lk, data)
        h = np.complex64((0.6 + 0.3 * s) * np.exp(1j * (0.4 + s)))
        noise = (rng.standard_normal(tx.shape) + 1j * rng.standard_normal(tx.shape)).astype(np.complex64) * np.float32(0.012)
        rxg = (tx * h + noise).astype(np.complex64)
        grids.append(rxg)
        datas.append(data)
        refs.append(ref.pusch_decode(lk, rxg))
        assert refs[-1]["crc"] and (refs[-1]["data"] == data).all()
    grid = torch.from_numpy(np.stack(grids)).cuda()
    # stage by stage against the reference receiver's own buffers
    ce, meas = ch0.chest(grid, tti, n_dmrs)
    d = ch0.equalize_deprecode(grid, ce, meas)
    g0 = ch0.demod_descramble(d, rnti, tti)
    torch.cuda.synchronize()
    off, half = 12 * kw["n_prb"], ch0.nsym // 2
    for s in range(nsf):
        r = refs[s]
        for slot in range(2):
            assert rel(ce[s, slot].cpu().numpy(), r["ce"][(slot + 1) * half - 4, off:off + ch0.M]) < 1e-5
        assert abs(float(meas[s, 0]) - r["noise"]) <= 2e-4 * r["noise"]
        assert rel(d[s].cpu().numpy(), r["d"]) < TOL
        # soft bits: the float path differs in the last bits, so a value next to a rounding boundary may move by one step
        diff = np.abs(g0[s].cpu().numpy().astype(np.int32) - r["g"].astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 2e-3, (diff.max(), (diff != 0).mean())
    # the one-call entry gives the same soft bits as the three stages
    g1 = ch0.rx(grid, rnti, tti, n_dmrs)
    torch.cuda.synchronize()
    assert torch.equal(g0, g1)
    ch0.close()
    # ... and the transport blocks decode (soft bits scaled into the generic decoder's envelope)
    ch4 = PuschChain(llr_shift=4 if kw["mod"] == 3 else 3 if kw["mod"] == 2 else 1, **kw)
    g = ch4.rx(grid, rnti, tti, n_dmrs)
    torch.cuda.synchronize()
    sch = SchDecoder(0, 8)
    seg = port.cbsegm(tbs)
    stride = (tbs // 8 + 3 + 768 + 15) // 16 * 16
    soft = np.zeros((nsf, seg["C"] * SOFTBUFFER_SIZE), np.int16)
    out = np.zeros((nsf, stride), np.uint8)
    rc, res = sch.decode(g.cpu().numpy().reshape(-1), soft.reshape(-1), out.reshape(-1),
                         [dict(tbs=tbs, Qm=2 * kw["mod"], rv=0, nof_e_bits=ch4.nof_bits, e_offset=s * ch4.nof_bits,
                               soft_offset=s * soft.shape[1], data_offset=s * stride) for s in range(nsf)])
    assert rc == 0
    for s in range(nsf):
        assert res[s]["result"] == 0, s
        assert (out[s, :tbs // 8] == datas[s]).all()
    sch.close()
    ch4.close()


def test_full_pipeline_from_time_samples(port):
    """bench.py's full-chain leg in small: numpy transmitter (checked against srsran_pusch_encode in tests/test_pusch_synth.py) ->
    fading + timing offset + AWGN -> OFDM rx -> chain -> de-matching -> turbo decoding; payload bytes and the SNR estimate."""
    import torch
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import PuschRxFull

    tbs, nsf = 75376, 4
    rx = PuschRxFull(17, 100, tbs, 3, llr_shift=4, max_noi=8, symbol_sz=2048)
    rnti = np.array([62, 159, 4000, 65535], np.uint32)
    tti = np.array([0, 13, 26, 9], np.uint32)
    iq, payload, G = sp.make_subframes_full(17, 100, 2048, tbs, 6, 0, sp.qpp_interleaver(5824), nsf, rnti, tti,
                                            lambda sf: rx.chain.dmrs(sf, 0), 23.0, seed=3)
    ok, its = rx.run(torch.from_numpy(iq).cuda(), nsf, rnti, tti)
    torch.cuda.synchronize()
    assert ok.all()
    assert (rx.data[:nsf, :tbs // 8 + 3].cpu().numpy() == payload).all()
    meas = rx.meas[:nsf].cpu().numpy()
    snr_db = 10 * np.log10(meas[:, 1])
    assert (np.abs(snr_db - 23.0) < 1.5).all(), snr_db
    # wrong RNTI: the scrambling sequence differs and nothing decodes
    ok2, _ = rx.run(torch.from_numpy(iq).cuda(), nsf, rnti + 1, tti)
    assert not ok2.any()
    rx.close()


def test_native_enb_ul_pipeline_with_harq(port):
    """srsran_b200_enb_ul_pusch_batch: host samples in (float and int16), transport-block bytes out, results, and a HARQ
    retransmission into the same slots (first transmission too noisy to decode, second one combines)."""
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import EnbUl, PuschChain

    tbs, nsf = 75376, 5
    ch = PuschChain(17, 100, False, 100, 0, 3, 4)
    dm = {sf: ch.dmrs(sf, 0) for sf in range(10)}
    ch.close()
    rnti = np.array([62, 159, 4000, 65535, 7], np.uint32)
    tti = np.array([0, 13, 26, 9, 1234], np.uint32)
    qpp = sp.qpp_interleaver(5824)
    iq, payload, _ = sp.make_subframes_full(17, 100, 2048, tbs, 6, 0, qpp, nsf, rnti, tti, lambda sf: dm[sf], 23.0, seed=3)
    enb = EnbUl(17, 100, tbs, 3, llr_shift=4, max_noi=8, symbol_sz=2048)
    data, res = enb.run(iq, rnti, tti)
    assert res["crc_ok"].all() and (data == payload).all()
    assert (np.abs(10 * np.log10(res["snr"]) - 23.0) < 1.5).all()
    # int16 samples (scaled to about half of full scale)
    peak = np.abs(iq.view(np.float32)).max()
    iq16 = np.round(iq.view(np.float32).reshape(nsf, -1, 2) * (16384.0 / peak)).astype(np.int16)
    data16, res16 = enb.run(iq16, rnti, tti)
    assert res16["crc_ok"].all() and (data16 == payload).all()
    # HARQ: rv 0 at 14 dB fails, rv 2 of the same transport blocks into the same slots succeeds
    iq_a, payload2, _ = sp.make_subframes_full(17, 100, 2048, tbs, 6, 0, qpp, nsf, rnti, tti, lambda sf: dm[sf], 14.0, seed=9, fading=False)
    iq_b, payload2b, _ = sp.make_subframes_full(17, 100, 2048, tbs, 6, 2, qpp, nsf, rnti, tti, lambda sf: dm[sf], 14.0, seed=9, fading=False)
    assert (payload2 == payload2b).all()
    _, r1 = enb.run(iq_a, rnti, tti, rv=np.zeros(nsf, np.uint32), new_data=np.ones(nsf, np.uint32))
    assert not r1["crc_ok"].any()
    d2, r2 = enb.run(iq_b, rnti, tti, rv=np.full(nsf, 2, np.uint32), new_data=np.zeros(nsf, np.uint32))
    assert r2["crc_ok"].all() and (d2 == payload2).all()
    # ... while the retransmission alone does not decode
    _, r3 = enb.run(iq_b, rnti, tti, rv=np.full(nsf, 2, np.uint32), new_data=np.ones(nsf, np.uint32))
    assert not r3["crc_ok"].any()
    enb.close()


def test_native_enb_ul_large_batch_is_decoded_in_groups(port):
    """From 2048 subframes of host samples on, the one-call receiver decodes in groups while later samples are still being
    copied (enb_ul.cu): same bytes and results as the same subframes in small batches, also on a second call (cached plans) and
    for a batch small enough to run as one group."""
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import EnbUl, PuschChain

    tbs, nd, nsf = 1544, 24, 2560 + 37  # 15 PRB QPSK, one code block; 2597 subframes = 6 chunks in 2 groups, the last one ragged
    ch = PuschChain(33, 15, False, 15, 0, 1, 1)
    dm = {sf: ch.dmrs(sf, 0) for sf in range(10)}
    ch.close()
    rng = np.random.default_rng(4)
    rnti_d = rng.integers(1, 65000, nd).astype(np.uint32)
    tti_d = rng.integers(0, 10240, nd).astype(np.uint32)
    qpp = sp.qpp_interleaver(1568)
    iq_d, payload_d, _ = sp.make_subframes_full(33, 15, 256, tbs, 2, 0, qpp, nd, rnti_d, tti_d, lambda sf: dm[sf], 9.0, seed=5)
    pick = rng.integers(0, nd, nsf)
    iq, rnti, tti, want = np.ascontiguousarray(iq_d[pick]), rnti_d[pick], tti_d[pick], payload_d[pick]
    enb = EnbUl(33, 15, tbs, 1, llr_shift=1, max_noi=8, symbol_sz=256)
    for _ in range(2):
        data, res = enb.run(iq, rnti, tti)
        assert res["crc_ok"].all() and (data == want).all()
    small, res_s = enb.run(iq[:nd], rnti[:nd], tti[:nd])
    assert (small == want[:nd]).all()
    assert (res["avg_iterations"][:nd] == res_s["avg_iterations"]).all() and np.allclose(res["snr"][:nd], res_s["snr"])
    # HARQ across the groups: rv 0 too noisy to decode, rv 2 of the same blocks into the same slots combines; per subframe the large
    # (grouped) batch must give exactly what the same subframe gives in a small single-group batch on a fresh object
    iq_a, pay_a, _ = sp.make_subframes_full(33, 15, 256, tbs, 2, 0, qpp, nd, rnti_d, tti_d, lambda sf: dm[sf], 0.0, seed=9, fading=False)
    iq_b, pay_b, _ = sp.make_subframes_full(33, 15, 256, tbs, 2, 2, qpp, nd, rnti_d, tti_d, lambda sf: dm[sf], 0.0, seed=9, fading=False)
    assert (pay_a == pay_b).all()
    ref_enb = EnbUl(33, 15, tbs, 1, llr_shift=1, max_noi=8, symbol_sz=256)
    o, z, r = np.ones(nd, np.uint32), np.zeros(nd, np.uint32), np.full(nd, 2, np.uint32)
    _, s1 = ref_enb.run(iq_a, rnti_d, tti_d, rv=z, new_data=o)
    sd2, s2 = ref_enb.run(iq_b, rnti_d, tti_d, rv=r, new_data=z)
    ref_enb.close()
    ones, zeros, rv2 = np.ones(nsf, np.uint32), np.zeros(nsf, np.uint32), np.full(nsf, 2, np.uint32)
    _, r1 = enb.run(np.ascontiguousarray(iq_a[pick]), rnti, tti, rv=zeros, new_data=ones)
    d2, r2 = enb.run(np.ascontiguousarray(iq_b[pick]), rnti, tti, rv=rv2, new_data=zeros)
    assert (r1["crc_ok"] == s1["crc_ok"][pick]).all() and (r1["avg_iterations"] == s1["avg_iterations"][pick]).all()
    assert (r2["crc_ok"] == s2["crc_ok"][pick]).all() and (r2["avg_iterations"] == s2["avg_iterations"][pick]).all()
    okm = r2["crc_ok"] != 0
    assert (d2[okm] == pay_a[pick][okm]).all() and (d2 == sd2[pick]).all()
    assert s1["crc_ok"].mean() < 0.5 and s2["crc_ok"].mean() > s1["crc_ok"].mean() + 0.3, (s1["crc_ok"].mean(), s2["crc_ok"].mean())
    enb.close()


def test_native_enb_ul_two_cells_from_one_thread(port):
    """srsran_b200_enb_ul_pusch_batch_begin / _finish: one thread queues the batches of two cells (two objects) before it waits for
    either; same bytes and results as the one-shot call; a second begin on a busy object is refused; finish without begin is a no-op."""
    import torch
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import EnbUl, PUSCH_RES_DTYPE, PuschChain

    tbs, nsf = 4584, 40
    cells = []
    for cell_id in (42, 43):
        ch = PuschChain(cell_id, 25, False, 25, 0, 2, 3)
        dm = {sf: ch.dmrs(sf, 0) for sf in range(10)}
        ch.close()
        rnti = (np.arange(nsf, dtype=np.uint32) * 31 + cell_id) % 65000 + 1
        tti = (np.arange(nsf, dtype=np.uint32) * 7 + cell_id) % 10240
        enb = EnbUl(cell_id, 25, tbs, 2, llr_shift=3, max_noi=8)
        iq, payload, _ = sp.make_subframes_full(cell_id, 25, enb.sf_sz // 15, tbs, 4, 0, sp.qpp_interleaver(port.cbsegm(tbs)["K1"]), nsf, rnti, tti,
                                                lambda sf: dm[sf], 16.0, seed=cell_id)
        want_data, want_res = enb.run(iq, rnti, tti)
        assert want_res["crc_ok"].all() and (want_data == payload).all()
        h_iq = torch.from_numpy(iq).pin_memory()
        cells.append(dict(enb=enb, iq=h_iq, rnti=rnti, tti=tti, want=want_data, want_res=want_res,
                          data=torch.zeros((nsf, enb.tb_bytes), dtype=torch.uint8).pin_memory(), res=np.zeros(nsf, PUSCH_RES_DTYPE)))
    assert cells[0]["enb"].finish() is None  # nothing begun
    for c in cells:
        c["enb"].begin_ptr(c["iq"].data_ptr(), nsf, c["rnti"], c["tti"], c["data"].data_ptr(), c["res"])
    with pytest.raises(RuntimeError):
        c = cells[0]
        c["enb"].begin_ptr(c["iq"].data_ptr(), nsf, c["rnti"], c["tti"], c["data"].data_ptr(), c["res"])
    for c in cells:
        c["enb"].finish()
        assert (c["data"].numpy() == c["want"]).all()
        assert (c["res"]["crc_ok"] == 1).all() and (c["res"]["avg_iterations"] == c["want_res"]["avg_iterations"]).all()
        c["enb"].close()


def test_chain_against_the_committed_golden_fixtures():
    """tests/golden/pusch_chain.npz (the reference's receiver buffers, tools/gen_golden.py): needs neither oracle/_ref nor the
    port at run time.  Five links, among them the 1- and 2-PRB allocations."""
    import os
    import torch
    from srslte_b200.pusch import PuschChain

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pusch_chain.npz"))
    i = 0
    while f"link{i}" in g:
        lk = [int(v) for v in g[f"link{i}"]]
        ch = PuschChain(cell_id=lk[0], cell_nof_prb=lk[1], cp_ext=bool(lk[2]), L_prb=lk[9], n_prb=lk[10], mod=lk[11], llr_shift=0,
                        cyclic_shift=lk[3], delta_ss=lk[4], group_hopping=bool(lk[5]), sequence_hopping=bool(lk[6]))
        rnti, tti, n_dmrs = np.array([lk[7]], np.uint32), np.array([lk[8]], np.uint32), np.array([lk[14]], np.uint32)
        assert np.abs(ch.dmrs(lk[8] % 10, lk[14]).reshape(-1) - g[f"dmrs{i}"]).max() < 1e-6
        grid = torch.from_numpy(g[f"rx{i}"][None]).cuda()
        ce, meas = ch.chest(grid, tti, n_dmrs)
        d = ch.equalize_deprecode(grid, ce, meas)
        soft = ch.demod_descramble(d, rnti, tti)
        torch.cuda.synchronize()
        off, half = 12 * lk[10], ch.nsym // 2
        for slot in range(2):
            assert rel(ce[0, slot].cpu().numpy(), g[f"ce{i}"][(slot + 1) * half - 4, off:off + ch.M]) < 1e-5, (i, slot)
        assert abs(float(meas[0, 0]) - g[f"meas{i}"][0]) <= 2e-4 * g[f"meas{i}"][0]
        assert rel(d[0].cpu().numpy(), g[f"d{i}"]) < TOL, i
        diff = np.abs(soft[0].cpu().numpy().astype(np.int32) - g[f"g{i}"].astype(np.int32))
        assert diff.max() <= 1 and (diff != 0).mean() < 3e-3, (i, diff.max(), (diff != 0).mean())
        ch.close()
        i += 1
    assert i >= 5
