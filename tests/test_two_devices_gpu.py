"""Two GPUs driven from ONE process: every engine takes an explicit device, and the kernels' function attributes (the shared
memory opt-ins above 48 KB, the carve-out preferences) are per device -- an object created on the second GPU after the first
one has been used must work exactly like the first.  Needs a box with two GPUs (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

from helpers import coded_llrs

pytestmark = pytest.mark.gpu


def _two_gpus():
    import torch

    return torch.cuda.device_count() >= 2


def test_decoder_on_the_second_device_matches_the_first(port):
    if not _two_gpus():
        pytest.skip("one GPU")
    from srslte_b200 import TurboDecoderBatch

    K, ncb = 6144, 70
    llr, _ = coded_llrs(port, K, ncb, 0.93, 16, 31, seed=21)
    outs = []
    for dev in (0, 1, 0):
        d = TurboDecoderBatch(device=dev)
        outs.append(d.decode(llr, K, 8, "B", True))
        # mixed lengths go through the load / decide kernels with the largest shared-memory opt-ins
        Ks = [6144, 40, 1024, 5824]
        ls = [coded_llrs(port, k, 10, 0.9, 16, 31, seed=k)[0] for k in Ks]
        outs.append(d.decode_mixed(ls, Ks, 8, "B", True))
        d.close()
    for out, ok, npass in (outs[0], outs[2], outs[4]):
        assert (out == outs[0][0]).all() and (ok == outs[0][1]).all() and (npass == outs[0][2]).all()
    for a, b in ((outs[1], outs[3]), (outs[1], outs[5])):
        for x, y in zip(a, b):
            for u, v in zip(x, y):
                assert (np.asarray(u) == np.asarray(v)).all()
    assert outs[0][1].mean() > 0.5


def test_pusch_receiver_on_the_second_device(port):
    if not _two_gpus():
        pytest.skip("one GPU")
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import EnbUl, PuschChain

    tbs, nsf = 4584, 6
    res = []
    for dev in (0, 1):
        ch = PuschChain(42, 25, False, 25, 0, 2, 3, device=dev)
        dm = {sf: ch.dmrs(sf, 0) for sf in range(10)}
        ch.close()
        rnti = np.arange(1, nsf + 1, dtype=np.uint32) * 77
        tti = np.arange(nsf, dtype=np.uint32) * 3
        seg = port.cbsegm(tbs)
        qpp = sp.qpp_interleaver(seg["K1"])
        enb = EnbUl(42, 25, tbs, 2, llr_shift=3, max_noi=8, device=dev)
        iq, payload, _ = sp.make_subframes_full(42, 25, enb.sf_sz // 15, tbs, 4, 0, qpp, nsf, rnti, tti, lambda sf: dm[sf], 18.0, seed=2)
        data, r = enb.run(iq, rnti, tti)
        assert r["crc_ok"].all() and (data == payload).all(), dev
        res.append((data, r))
        enb.close()
    assert (res[0][0] == res[1][0]).all() and (res[0][1]["avg_iterations"] == res[1][1]["avg_iterations"]).all()
