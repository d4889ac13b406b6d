"""Rate de-matching and the transport-block decode loop on the GPU against the oracle, through the C ABI.  -m gpu."""
import numpy as np
import pytest

from helpers import make_tb

pytestmark = pytest.mark.gpu
SB = 18600


@pytest.fixture(autouse=True, params=["fused", "unfused"])
def dematching_kernel(request, monkeypatch):
    """Every test of this file runs with the separate de-matching + load kernels (the default) and with the de-matching kernel that
    writes the decoder's tiles itself (SRSLTE_B200_RM_FUSED=1)."""
    if request.param == "fused":
        monkeypatch.setenv("SRSLTE_B200_RM_FUSED", "1")
    else:
        monkeypatch.delenv("SRSLTE_B200_RM_FUSED", raising=False)
    return request.param


@pytest.fixture(scope="module")
def sch():
    from srslte_b200 import SchDecoder

    d = SchDecoder(device=0, max_noi=8)
    yield d
    d.close()


def test_rm_rx_all_sizes_and_rvs(sch, port):
    """rm_turbo_test's sweep (-c k -i rv over 188 x 4) with puncturing (E<N), exact fit and repetition (E>N), on top of a
    non-zero soft buffer (HARQ combining) and from a fresh one."""
    rng = np.random.default_rng(0)
    Ks = port.cb_sizes()
    jobs, e_all, expect = [], [], []
    soft_pool = np.zeros(188 * 4 * 3 * SB, np.int16)
    e_off = 0
    slot = 0
    for i, K in enumerate(Ks):
        n = 3 * int(K) + 12
        for rv in range(4):
            for frac in (0.4, 1.0, 2.3):
                E = max(1, int(frac * n))
                e = rng.integers(-3000, 3000, E).astype(np.int16)
                fresh = (slot % 2) == 0
                init = rng.integers(-20000, 20000, n).astype(np.int16)
                soft_pool[slot * SB: slot * SB + n] = init
                want = np.zeros(n + 64, np.int16) if fresh else np.concatenate([init, np.zeros(64, np.int16)])
                port.rm_rx(e, want, i, rv)
                expect.append((slot, n, want[:n].copy()))
                jobs.append(dict(cb_idx=i, rv=rv, E=E, new_data=int(fresh), in_offset=e_off, soft_offset=slot * SB))
                e_all.append(e)
                e_off += E
                slot += 1
    e_all = np.concatenate(e_all)
    assert sch.rm_rx(e_all, soft_pool, jobs) == 0
    for slot, n, want in expect:
        assert (soft_pool[slot * SB: slot * SB + n] == want).all(), slot


def test_rm_rx_invalid_inputs(sch):
    soft = np.zeros(SB, np.int16)
    e = np.zeros(100, np.int16)
    assert sch.rm_rx(e, soft, [dict(cb_idx=0, rv=4, E=100, in_offset=0, soft_offset=0)]) == -2
    assert sch.rm_rx(e, soft, [dict(cb_idx=188, rv=0, E=100, in_offset=0, soft_offset=0)]) == -2


@pytest.mark.parametrize("cases", [[(6120, 2, 14400, 0, 0.6), (2216, 2, 3000, 0, 0.5), (12960, 4, 28800, 0, 0.75)],
                                   [(75376, 6, 86400, 0, 0.35), (36696, 6, 57600, 0, 0.55), (75376, 6, 86400, 0, 0.52),
                                    (12960, 2, 14406, 2, 0.4), (75376, 4, 57600, 0, 0.3)]])
def test_decode_tb_batch_matches_oracle(sch, port, cases):
    """Mixed batch of transport blocks (incl. the 100-PRB 64QAM case of BASELINE config 4: TBS 75376 -> 13 x K=5824 with the
    reference's E split off-by-one) -- bytes, TB verdict, per-block CRC mask and average iterations must equal the oracle."""
    e_all, tbs_desc, want = [], [], []
    e_off = soft_off = data_off = 0
    for i, (tbs, Qm, G, rv, sigma) in enumerate(cases):
        e, _, s = make_tb(port, tbs, Qm, G, rv, sigma, seed=1000 + i)
        C = s["C"]
        soft = np.zeros(C * SB, np.int16)
        cbcrc = np.zeros(C, np.uint8)
        data = np.zeros(tbs // 8 + 3 + 768, np.uint8)
        ret, iters = port.decode_tb(e, tbs, Qm, rv, 8, soft, cbcrc, data)
        want.append((ret, iters / C, cbcrc.copy(), data[:tbs // 8 + 3].copy(), soft.copy(), data_off, soft_off, C, tbs))
        tbs_desc.append(dict(tbs=tbs, Qm=Qm, rv=rv, nof_e_bits=G, e_offset=e_off, soft_offset=soft_off, data_offset=data_off, new_data=1))
        e_all.append(e)
        e_off += G
        soft_off += C * SB
        data_off += tbs // 8 + 3 + 768 + 13
    e_all = np.concatenate(e_all)
    soft_pool = np.full(soft_off, 77, np.int16)  # stale contents: new_data must ignore them
    data = np.zeros(data_off + 1024, np.uint8)
    rc, res = sch.decode(e_all, soft_pool, data, tbs_desc)
    assert rc == 0
    for r, (ret, avg, cbcrc, d, soft, doff, soff, C, tbs) in zip(res, want):
        assert r["result"] == ret
        assert r["nof_cb"] == C
        assert abs(r["avg_iterations"] - avg) < 1e-6
        assert r["cb_crc_mask"] == sum(int(b) << c for c, b in enumerate(cbcrc))
        assert (data[doff: doff + tbs // 8 + 3] == d).all()
        K = port.cbsegm(tbs)["K1"]
        for c in range(C):  # combined soft buffers are the HARQ state: must match too
            assert (soft_pool[soff + c * SB: soff + c * SB + 3 * K + 12] == soft[c * SB: c * SB + 3 * K + 12]).all()


def test_harq_retransmission_combining(sch, port):
    """First transmission too noisy, second (rv 2) combined on the kept soft buffers; already-decoded blocks are skipped
    through cb_crc_mask (sch.c:390,466-471)."""
    tbs, Qm, G = 36696, 6, 45000
    s = port.cbsegm(tbs)
    C = s["C"]
    e0, expect, _ = make_tb(port, tbs, Qm, G, 0, 0.95, seed=5)
    e2, expect2, _ = make_tb(port, tbs, Qm, G, 2, 0.95, seed=5)  # same payload (same seed), other redundancy version
    assert (expect == expect2).all()
    soft_o = np.zeros(C * SB, np.int16)
    cb_o = np.zeros(C, np.uint8)
    data_o = np.zeros(tbs // 8 + 3 + 768, np.uint8)
    r0, _ = port.decode_tb(e0, tbs, Qm, 0, 8, soft_o, cb_o, data_o)
    mask0 = sum(int(b) << c for c, b in enumerate(cb_o))
    r1, it1 = port.decode_tb(e2, tbs, Qm, 2, 8, soft_o, cb_o, data_o)

    soft = np.zeros(C * SB, np.int16)
    data = np.zeros(tbs // 8 + 3 + 768, np.uint8)
    rc, res = sch.decode(e0, soft, data, [dict(tbs=tbs, Qm=Qm, rv=0, nof_e_bits=G, e_offset=0, soft_offset=0, data_offset=0, new_data=1)])
    assert rc == 0 and res[0]["result"] == r0 and res[0]["cb_crc_mask"] == mask0
    rc, res = sch.decode(e2, soft, data, [dict(tbs=tbs, Qm=Qm, rv=2, nof_e_bits=G, e_offset=0, soft_offset=0, data_offset=0, new_data=0,
                                               cb_crc_mask=mask0)])
    assert rc == 0 and res[0]["result"] == r1
    assert (data[:tbs // 8 + 3] == data_o[:tbs // 8 + 3]).all()
    assert (soft == soft_o).all()
    if r1 == 0:
        assert (data[:tbs // 8 + 3] == expect).all()


def test_filler_bits_rejected(sch):
    """Non-standard TBS needing filler bits: SRSRAN_ERROR_INVALID_INPUTS like sch.c:521-524."""
    tbs = 6200  # B = 6224 > 6144 -> C = 2, B' = 6272, K+ = 3136 -> F = 0 ; pick one with F != 0 instead
    from oracle import loader

    p = loader.api("port")
    while p.cbsegm(tbs)["F"] == 0:
        tbs += 8
    soft = np.zeros(4 * SB, np.int16)
    data = np.zeros(4096, np.uint8)
    rc, res = sch.decode(np.zeros(20000, np.int16), soft, data, [dict(tbs=tbs, Qm=2, rv=0, nof_e_bits=20000, e_offset=0, soft_offset=0,
                                                                     data_offset=0)])
    assert rc == 0 and res[0]["result"] == -2


def test_config2_all_188_sizes_in_one_dematch_and_decode_batch(sch, port):
    """BASELINE.json configs[2]: every LTE QPP size K = 40..6144 in ONE batch through rate de-matching + turbo decoding, with
    punctured (E < 3K+12), exactly fitting and repeated (E > 3K+12) transmissions and all four redundancy versions.
    One single-code-block transport block per size (TBS = K - 24, CRC24A): bytes, verdicts, pass counts and the combined soft
    buffers must equal the oracle's decode_tb."""
    Ks = port.cb_sizes()
    assert len(Ks) == 188
    e_all, tbs_desc, want = [], [], []
    e_off = soft_off = data_off = 0
    for i, K in enumerate(Ks):
        K = int(K)
        tbs, Qm = K - 24, 2
        ratio = (0.45, 1.0, 1.7)[i % 3]
        G = max(2 * Qm, int(ratio * (3 * K + 12)) // Qm * Qm)
        if ratio == 1.0:
            G = 3 * K + 12
        rv = i % 4 if ratio > 0.9 else 0          # heavily punctured blocks only decode from rv 0
        sigma = 0.55 if ratio < 0.9 else 0.85
        e, _, s = make_tb(port, tbs, Qm, G, rv, sigma, seed=3000 + i)
        assert s["C"] == 1 and s["K1"] == K and s["F"] == 0
        soft = np.zeros(SB, np.int16)
        cbcrc = np.zeros(1, np.uint8)
        data = np.zeros(tbs // 8 + 3 + 768, np.uint8)
        ret, iters = port.decode_tb(e, tbs, Qm, rv, 8, soft, cbcrc, data)
        want.append((ret, iters, int(cbcrc[0]), data[:tbs // 8 + 3].copy(), soft[:3 * K + 12].copy(), data_off, soft_off, tbs, K))
        tbs_desc.append(dict(tbs=tbs, Qm=Qm, rv=rv, nof_e_bits=G, e_offset=e_off, soft_offset=soft_off, data_offset=data_off, new_data=1))
        e_all.append(e)
        e_off += G
        soft_off += SB
        data_off += (tbs // 8 + 3 + 768 + 15) // 16 * 16
    e_all = np.concatenate(e_all)
    soft_pool = np.zeros(soft_off, np.int16)
    data = np.zeros(data_off + 1024, np.uint8)
    rc, res = sch.decode(e_all, soft_pool, data, tbs_desc)
    assert rc == 0
    n_ok = 0
    for r, (ret, iters, cbok, d, soft, doff, soff, tbs, K) in zip(res, want):
        assert r["result"] == ret and r["nof_cb"] == 1, K
        assert abs(r["avg_iterations"] - iters) < 1e-6, K
        assert r["cb_crc_mask"] == cbok, K
        assert (data[doff: doff + tbs // 8 + 3] == d).all(), K
        assert (soft_pool[soff: soff + 3 * K + 12] == soft).all(), K
        n_ok += ret == 0
    assert n_ok > 150  # the point is parity, but most of the batch should decode


def test_soft_values_beyond_int8_are_decoded_with_int16_tiles_on_demand(port):
    """The decode loop's decoder workspace is carved without the int16 copies of the channel LLRs until a batch needs them
    (int16 on demand): a FRESH object first sees a batch that fits int8, then one whose soft values reach +-400 (the batch is
    flagged on the device and decoded again with int16 tiles), then an int8 batch again.  Every result must equal the
    oracle's decode_tb."""
    from srslte_b200 import SchDecoder

    d = SchDecoder(device=0, max_noi=8)
    try:
        for scale, clip, seed in ((16.0, 31, 1), (150.0, 400, 2), (16.0, 31, 3), (150.0, 400, 4)):
            tbs, Qm, G = 12960, 4, 28800
            e, _, s = make_tb(port, tbs, Qm, G, 0, 0.7, scale=scale, clip=clip, seed=seed)
            assert np.abs(e).max() == clip
            soft = np.zeros(s["C"] * SB, np.int16)
            cbcrc = np.zeros(s["C"], np.uint8)
            want = np.zeros(tbs // 8 + 3 + 768, np.uint8)
            ret, iters = port.decode_tb(e, tbs, Qm, 0, 8, soft, cbcrc, want)
            soft_pool = np.zeros(s["C"] * SB, np.int16)
            data = np.zeros(tbs // 8 + 3 + 1024, np.uint8)
            rc, res = d.decode(e, soft_pool, data, [dict(tbs=tbs, Qm=Qm, rv=0, nof_e_bits=G, e_offset=0, soft_offset=0, data_offset=0)])
            assert rc == 0 and res[0]["result"] == ret and abs(res[0]["avg_iterations"] - iters / s["C"]) < 1e-6, (scale, res, ret, iters)
            assert (data[:tbs // 8 + 3] == want[:tbs // 8 + 3]).all(), scale
            assert (soft_pool == soft).all()
    finally:
        d.close()


def test_repeated_lists_reuse_the_plan_but_follow_the_data(port):
    """decode_batch keeps what it derives from the transport block list (segmentation, descriptors, decoder groups) and reuses
    it when the next list repeats the same inputs.  The kept plan must only ever stand for the descriptors: new soft bits under
    the same list, a changed redundancy version, a de-matching call in between and a HARQ retransmission must all give the
    oracle's results."""
    from srslte_b200 import SchDecoder

    d = SchDecoder(device=0, max_noi=8)
    try:
        tbs, Qm, G = 12960, 4, 28800
        C = port.cbsegm(tbs)["C"]

        def oracle(e, rv, soft, cbcrc):
            data = np.zeros(tbs // 8 + 3 + 768, np.uint8)
            ret, iters = port.decode_tb(e, tbs, Qm, rv, 8, soft, cbcrc, data)
            return ret, iters / C, data[:tbs // 8 + 3].copy()

        def ours(e, rv, soft_pool, new_data=1, mask=0):
            data = np.zeros(tbs // 8 + 3 + 1024, np.uint8)
            rc, res = d.decode(e, soft_pool, data, [dict(tbs=tbs, Qm=Qm, rv=rv, nof_e_bits=G, e_offset=0, soft_offset=0, data_offset=0,
                                                         new_data=new_data, cb_crc_mask=mask)])
            assert rc == 0
            return res[0], data[:tbs // 8 + 3].copy()

        seq = [(0, 0.7, 11), (0, 0.7, 11), (0, 0.75, 12), (2, 0.7, 13), (0, 0.7, 11)]   # (rv, sigma, seed): same list, new data, new rv, back
        for n, (rv, sigma, seed) in enumerate(seq):
            e, _, _ = make_tb(port, tbs, Qm, G, rv, sigma, seed=seed)
            want = oracle(e, rv, np.zeros(C * SB, np.int16), np.zeros(C, np.uint8))
            if n == 3:  # a de-matching call through the same object in between (it shares the descriptor arena)
                assert d.rm_rx(np.zeros(100, np.int16), np.zeros(SB, np.int16), [dict(cb_idx=0, rv=0, E=100, new_data=1, in_offset=0, soft_offset=0)]) == 0
            got, data = ours(e, rv, np.zeros(C * SB, np.int16))
            assert got["result"] == want[0] and abs(got["avg_iterations"] - want[1]) < 1e-6 and (data == want[2]).all(), n
        # HARQ: a noisy first transmission, then rv 2 combined on the kept buffers with the updated mask (a different list)
        e0, _, _ = make_tb(port, tbs, Qm, G, 0, 1.05, seed=21)
        e2, _, _ = make_tb(port, tbs, Qm, G, 2, 1.05, seed=21)
        soft_o, cb_o = np.zeros(C * SB, np.int16), np.zeros(C, np.uint8)
        w0 = oracle(e0, 0, soft_o, cb_o)
        mask0 = sum(int(b) << c for c, b in enumerate(cb_o))
        w1 = oracle(e2, 2, soft_o, cb_o)
        soft = np.zeros(C * SB, np.int16)
        g0, _ = ours(e0, 0, soft)
        assert g0["result"] == w0[0] and g0["cb_crc_mask"] == mask0
        g1, _ = ours(e2, 2, soft, new_data=0, mask=mask0)
        assert g1["result"] == w1[0] and (soft == soft_o).all()
    finally:
        d.close()


def test_decode_in_two_halves(port):
    """srsran_b200_sch_decode_begin / _finish: two objects with a batch in flight each give what the one-shot call gives; a second
    begin on a busy object, host pointers and a finish without a begin behave as documented."""
    import torch
    from srslte_b200 import SchDecoder, _lib
    from srslte_b200.pusch import TB_DTYPE

    L = _lib.lib()
    tbs, Qm, G, ntb = 12960, 4, 28800, 6
    C = port.cbsegm(tbs)["C"]
    es, want = [], []
    for t in range(ntb):
        e, _, _ = make_tb(port, tbs, Qm, G, 0, 0.7 + 0.08 * t, seed=40 + t)
        es.append(e)
        data = np.zeros(tbs // 8 + 3 + 768, np.uint8)
        ret, iters = port.decode_tb(e, tbs, Qm, 0, 8, np.zeros(C * SB, np.int16), np.zeros(C, np.uint8), data)
        want.append((ret, iters / C, data[:tbs // 8 + 3].copy()))
    stride = (tbs // 8 + 3 + 768 + 15) // 16 * 16
    e_dev = torch.from_numpy(np.concatenate(es)).cuda()

    def new_set():
        tb = np.zeros(ntb, TB_DTYPE)
        i = np.arange(ntb, dtype=np.uint64)
        tb["tbs"], tb["Qm"], tb["nof_e_bits"], tb["new_data"] = tbs, Qm, G, 1
        tb["e_offset"], tb["soft_offset"], tb["data_offset"] = i * np.uint64(G), i * np.uint64(C * SB), i * np.uint64(stride)
        return dict(q=SchDecoder(device=0, max_noi=8), tb=tb, soft=torch.zeros(ntb * C * SB, dtype=torch.int16, device="cuda"),
                    data=torch.zeros(ntb * stride, dtype=torch.uint8, device="cuda"))

    def begin(s, flags=_lib.FLAG_DEVICE_PTRS):
        return L.srsran_b200_sch_decode_begin(s["q"]._h, e_dev.data_ptr(), e_dev.numel(), s["soft"].data_ptr(), s["soft"].numel(),
                                              s["data"].data_ptr(), s["data"].numel(), s["tb"].ctypes.data, ntb, flags)

    a, b = new_set(), new_set()
    try:
        assert L.srsran_b200_sch_decode_finish(a["q"]._h) == 0          # nothing begun: nothing to wait for
        assert begin(a, 0) == -2                                        # host pointers are refused
        assert begin(a) == 0 and begin(b) == 0                          # two batches in flight
        assert begin(a) == -2                                           # one batch per object at a time
        assert L.srsran_b200_sch_decode_finish(a["q"]._h) == 0 and L.srsran_b200_sch_decode_finish(b["q"]._h) == 0
        for s in (a, b):
            d = s["data"].cpu().numpy().reshape(ntb, stride)
            for t in range(ntb):
                assert s["tb"]["result"][t] == want[t][0] and abs(s["tb"]["avg_iterations"][t] - want[t][1]) < 1e-6, t
                assert (d[t, :tbs // 8 + 3] == want[t][2]).all(), t
        assert begin(a) == 0 and L.srsran_b200_sch_decode_finish(a["q"]._h) == 0   # and again on the same object (cached plan)
        assert (a["tb"]["result"] == [w[0] for w in want]).all()
    finally:
        a["q"].close()
        b["q"].close()
