"""Parity of the CUDA turbo decoder with the oracle, through the C ABI (srsran_b200_tdec_run).  Run with -m gpu."""
import os

import numpy as np
import pytest

from helpers import coded_llrs

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True, params=["throughput", "low_latency", "auto"])
def siso_kernel(request, monkeypatch):
    """Every test of this file runs three times: forced onto the throughput SISO kernel (two warps per tile), forced onto
    the low-latency kernel (16 warps per tile, all windows in parallel between two recursion warps), and with the engine's
    own choice (small batches: low latency; large early-stop batches: throughput first, low latency for the tail passes)."""
    if request.param == "throughput":
        monkeypatch.setenv("SRSLTE_B200_TDEC_LL", "0")
    elif request.param == "low_latency":
        monkeypatch.setenv("SRSLTE_B200_TDEC_LL", "1")
    else:
        monkeypatch.delenv("SRSLTE_B200_TDEC_LL", raising=False)
    return request.param


@pytest.fixture(scope="module")
def dec():
    from srslte_b200 import TurboDecoderBatch

    d = TurboDecoderBatch(device=0)
    yield d
    d.close()


def test_golden_vectors_from_reference(dec):
    """Decoded bytes / CRC flags / pass counts recorded from the reference's generic int16 decoder (tools/gen_golden.py)."""
    g = np.load(os.path.join(GOLD, "tdec.npz"))
    ci = 0
    while f"c{ci}_meta" in g:
        K, ncb, _ = (int(v) for v in g[f"c{ci}_meta"])
        llr = g[f"c{ci}_llr"]
        for es in (0, 1):
            out, ok, npass = dec.decode(llr, K, 8, "B", bool(es))
            assert (out == g[f"c{ci}_loop{es}_out"]).all(), (ci, es)
            assert (ok == g[f"c{ci}_loop{es}_ok"]).all(), (ci, es)
            assert (npass == g[f"c{ci}_loop{es}_npass"]).all(), (ci, es)
        for p in range(1, 9):  # decision after exactly p passes
            out, _, _ = dec.decode(llr, K, p, None, False)
            assert (out == g[f"c{ci}_per_pass"][:, p - 1, :]).all(), (ci, p)
        ci += 1
    assert ci >= 6


@pytest.mark.parametrize("K,ncb,sigma,scale,clip", [(40, 200, 0.8, 16, 31), (48, 65, 1.0, 16, 31), (504, 130, 0.9, 16, 31),
                                                    (1024, 100, 1.0, 16, 31), (6144, 70, 0.93, 16, 31), (6144, 9, 1.3, 32, 63),
                                                    (6144, 5, 0.8, 8000, 30000), (2048, 40, 1.2, 500, 2000)])
def test_bit_exact_vs_oracle(dec, port, K, ncb, sigma, scale, clip):
    """Same seeded quantised LLRs into both; bytes, CRC outcome and pass count must be identical, also where the
    reference's int16 arithmetic wraps (clip >= 100)."""
    llr, _ = coded_llrs(port, K, ncb, sigma, scale, clip, seed=K * 7 + ncb)
    for early in (True, False):
        for mp in (8, 3):
            o1, k1, n1, _ = port.decode_batch(llr, K, mp, "B", 0, early, nthreads=8)
            o2, k2, n2 = dec.decode(llr, K, mp, "B", early)
            assert (o1 == o2).all(), (early, mp, np.argwhere((o1 != o2).any(axis=1)).ravel()[:8])
            assert (k1 == k2).all() and (n1 == n2).all(), (early, mp)


def test_all_188_sizes(dec, port):
    """BASELINE.json config 3 (every LTE QPP size in one run), bit-exact."""
    for i, K in enumerate(port.cb_sizes()):
        K = int(K)
        llr, _ = coded_llrs(port, K, 3, 0.85 + 0.3 * (i % 3), 16, 31, seed=i)
        o1, k1, n1, _ = port.decode_batch(llr, K, 4, "B", 0, True)
        o2, k2, n2 = dec.decode(llr, K, 4, "B", True)
        assert (o1 == o2).all() and (k1 == k2).all() and (n1 == n2).all(), K


def test_edge_cases(dec, port):
    K = 512
    # empty batch
    out, ok, npass = dec.decode(np.zeros((0, 3 * K + 12), np.int16), K)
    assert out.shape == (0, K // 8)
    # all-zero LLRs: decodes to all-zero bits, which passes the CB CRC
    z = np.zeros((3, 3 * K + 12), np.int16)
    o1, k1, n1, _ = port.decode_batch(z, K, 8, "B", 0, True)
    o2, k2, n2 = dec.decode(z, K, 8, "B", True)
    assert (o1 == o2).all() and (k1 == k2).all() and (n1 == n2).all()
    # extreme values
    rng = np.random.default_rng(0)
    x = rng.integers(-32768, 32767, (5, 3 * K + 12)).astype(np.int16)
    o1, k1, n1, _ = port.decode_batch(x, K, 6, "B", 0, False)
    o2, k2, n2 = dec.decode(x, K, 6, "B", False)
    assert (o1 == o2).all() and (k1 == k2).all() and (n1 == n2).all()
    # invalid K is rejected like srsran_tdec_new_cb (turbodecoder.c:517-521)
    with pytest.raises(RuntimeError):
        dec.decode(np.zeros((1, 3 * 100 + 12), np.int16), 100)
    # CRC24A variant (single-code-block transport block)
    llr, _ = coded_llrs(port, 1024, 10, 0.8, 16, 31, seed=3, crc="A")
    o1, k1, n1, _ = port.decode_batch(llr, 1024, 6, "A", 1024, True)
    o2, k2, n2 = dec.decode(llr, 1024, 6, "A", True)
    assert (o1 == o2).all() and (k1 == k2).all() and (n1 == n2).all() and k2.all()


def test_device_pointer_path_and_full_size_properties(dec, port):
    """Device-resident path at a size the oracle cannot replay in full: 16,384 blocks of K=6144 generated on the GPU.
    Properties: every block whose CRC matched equals the transmitted bits; early stop and fixed-pass runs agree on every
    block that converged; a sampled subset is bit-exact against the oracle."""
    import torch

    from srslte_b200.tdec import synth_llr

    K, ncb = 6144, 16384
    llr, truth = synth_llr(0, ncb, K, sigma=0.79, scale=16.0, clip=31, seed=77)
    out = torch.empty((ncb, K // 8), dtype=torch.uint8, device="cuda")
    ok = torch.empty(ncb, dtype=torch.uint8, device="cuda")
    npass = torch.empty(ncb, dtype=torch.uint8, device="cuda")
    dec.decode_device(llr, K, out, ok, npass, 8, "B", True)
    torch.cuda.synchronize()
    okb = ok.bool()
    assert okb.float().mean().item() > 0.9
    assert (out[okb] == truth[okb]).all()
    out8 = torch.empty_like(out)
    ok8 = torch.empty_like(ok)
    np8 = torch.empty_like(npass)
    dec.decode_device(llr, K, out8, ok8, np8, 8, "B", False)
    torch.cuda.synchronize()
    assert (ok8 == ok).all() and (np8 == npass).all()
    idx = torch.arange(0, ncb, 683)
    sub = llr[idx].cpu().numpy()
    o1, k1, n1, _ = port.decode_batch(sub, K, 8, "B", 0, True, nthreads=8)
    assert (o1 == out[idx].cpu().numpy()).all()
    assert (k1 == ok[idx].cpu().numpy()).all() and (n1 == npass[idx].cpu().numpy()).all()
    o1, _, _, _ = port.decode_batch(sub, K, 8, "B", 0, False, nthreads=8)
    assert (o1 == out8[idx].cpu().numpy()).all()


def test_mixed_int8_and_int16_tiles_in_one_batch(dec, port):
    """The decoder keeps a tile's channel LLRs as int8 when all 64 blocks fit and as int16 otherwise (fmt per tile): one
    batch with both kinds of tiles, values sitting exactly on the int8 boundary, and a tile that leaves int8 only through
    a tail value must decode bit-exactly like the oracle either way."""
    K = 1024
    llr, _ = coded_llrs(port, K, 64 * 4 + 7, 0.9, 16, 31, seed=21)
    llr[:64] = np.clip(llr[:64].astype(np.int32) * 4, -128, 127).astype(np.int16)       # tile 0: int8, touching both ends
    llr[64:128] = (llr[64:128].astype(np.int32) * 5).astype(np.int16)                   # tile 1: up to +-155 -> int16
    llr[130, 3 * K + 1] = 128                                                            # tile 2: one parity tail value
    llr[200, 3 * K + 6] = 3000                                                           # tile 3: encoder-2 systematic tail only
    for early in (True, False):
        o1, k1, n1, _ = port.decode_batch(llr, K, 6, "B", 0, early, nthreads=8)
        o2, k2, n2 = dec.decode(llr, K, 6, "B", early)
        assert (o1 == o2).all() and (k1 == k2).all() and (n1 == n2).all(), early


def test_int8_llr_container_decodes_like_int16(dec, port):
    """SRSRAN_B200_FLAG_LLR_INT8: the same values in an int8 buffer (host and device path) give the int16 entry's results."""
    import ctypes as C

    import torch

    from srslte_b200 import _lib
    from helpers import coded_llrs

    for K, ncb in ((6144, 70), (40, 129), (1008, 65)):
        llr, _ = coded_llrs(port, K, ncb, 0.9, 16.0, 31, seed=K)
        assert np.abs(llr).max() <= 127
        want = dec.decode(llr, K, 8, "B", True)
        l8 = np.ascontiguousarray(llr.astype(np.int8))
        out = np.zeros((ncb, K // 8), np.uint8)
        ok = np.zeros(ncb, np.uint8)
        npass = np.zeros(ncb, np.uint8)
        rc = dec._lib.srsran_b200_tdec_run(dec._h, l8.ctypes.data, ncb, K, 8, _lib.CRC24B, 1, out.ctypes.data, ok.ctypes.data,
                                           npass.ctypes.data, _lib.FLAG_LLR_INT8, None)
        assert rc == 0
        assert (out == want[0]).all() and (ok == want[1]).all() and (npass == want[2]).all()
        d8 = torch.from_numpy(l8).cuda()
        o = torch.empty((ncb, K // 8), dtype=torch.uint8, device="cuda")
        k = torch.empty(ncb, dtype=torch.uint8, device="cuda")
        n = torch.empty(ncb, dtype=torch.uint8, device="cuda")
        dec.decode_device(d8, K, o, k, n, 8, "B", True)
        torch.cuda.synchronize()
        assert (o.cpu().numpy() == want[0]).all() and (k.cpu().numpy() == want[1]).all() and (n.cpu().numpy() == want[2]).all()


def test_mixed_sizes_in_one_batch(dec, port):
    """BASELINE config 3 through srsran_b200_tdec_run_mixed: all 188 lengths in ONE batch (one launch per pass over every
    tile, ordered by length), 64+ blocks for a spread of sizes and a few blocks for every other one, bit-exact per block."""
    sizes = [int(k) for k in port.cb_sizes()]
    big = set(sizes[::12] + [40, 6144])
    Ks, llrs = [], []
    for i, K in enumerate(sizes):
        n = (66 if K <= 1024 else 64) if K in big else 2
        if K > 3000 and K in big:
            n = 64
        Ks.append(K)
        llrs.append(coded_llrs(port, K, n, 0.85 + 0.3 * (i % 3), 16, 31, seed=100 + i)[0])
    for early in (True, False):
        res = dec.decode_mixed(llrs, Ks, 6, "B", early)
        for (o2, k2, n2), llr, K in zip(res, llrs, Ks):
            o1, k1, n1, _ = port.decode_batch(llr, K, 6, "B", 0, early, nthreads=8)
            assert (o1 == o2).all() and (k1 == k2).all() and (n1 == n2).all(), (K, early)
    # an invalid length anywhere in the list is rejected like srsran_tdec_new_cb
    with pytest.raises(RuntimeError):
        dec.decode_mixed([np.zeros((1, 3 * 100 + 12), np.int16)], [100])


def test_lane_repacking_between_passes(dec, port, monkeypatch):
    """Block-granular early stop on the device: running lanes are re-packed into fewer tiles after every pass.  Blocks with
    very different convergence share tiles so lanes really move; results must equal the oracle's per block, and equal the
    run with re-packing switched off."""
    K, ncb = 104, 64 * 40 - 7
    rng = np.random.default_rng(9)
    parts = []
    for i in range(ncb):
        sigma = 1.3 if rng.random() < 0.15 else 0.5
        parts.append(coded_llrs(port, K, 1, sigma, 16, 31, seed=7000 + i)[0])
    llr = np.concatenate(parts)
    llr2, _ = coded_llrs(port, 40, 700, 1.0, 16, 31, seed=11)
    monkeypatch.setenv("SRSLTE_B200_TDEC_COMPACT_MIN_TILES", "1")
    dec.profile_reset(True)
    (a, b) = dec.decode_mixed([llr, llr2], [K, 40], 8, "B", True)
    prof = dec.profile_get()
    dec.profile_reset(False)
    assert prof["repack_launches"] == 7
    o1, k1, n1, _ = port.decode_batch(llr, K, 8, "B", 0, True, nthreads=8)
    p1, q1, r1, _ = port.decode_batch(llr2, 40, 8, "B", 0, True, nthreads=8)
    assert n1.max() > n1.min() + 2
    assert (a[0] == o1).all() and (a[1] == k1).all() and (a[2] == n1).all()
    assert (b[0] == p1).all() and (b[1] == q1).all() and (b[2] == r1).all()
    monkeypatch.setenv("SRSLTE_B200_TDEC_NO_COMPACT", "1")
    (c, d) = dec.decode_mixed([llr, llr2], [K, 40], 8, "B", True)
    assert (c[0] == a[0]).all() and (c[2] == a[2]).all() and (d[0] == b[0]).all()


def test_lane_repacking_full_size_properties(dec, port):
    """65,536 blocks of K=6144 near the waterfall (blocks finish after very different numbers of passes): with and without
    re-packing the decoder must return the same bytes, flags and pass counts; a sample is checked against the oracle."""
    import torch

    from srslte_b200.tdec import synth_llr

    K, ncb = 6144, 65536
    llr, truth = synth_llr(0, ncb, K, sigma=0.90, scale=16.0, clip=31, seed=123)
    outs = []
    for no_compact in (False, True):
        if no_compact:
            os.environ["SRSLTE_B200_TDEC_NO_COMPACT"] = "1"
        try:
            out = torch.empty((ncb, K // 8), dtype=torch.uint8, device="cuda")
            ok = torch.empty(ncb, dtype=torch.uint8, device="cuda")
            npass = torch.empty(ncb, dtype=torch.uint8, device="cuda")
            dec.decode_device(llr, K, out, ok, npass, 8, "B", True)
            torch.cuda.synchronize()
            outs.append((out, ok, npass))
        finally:
            os.environ.pop("SRSLTE_B200_TDEC_NO_COMPACT", None)
    (o_a, k_a, n_a), (o_b, k_b, n_b) = outs
    assert (k_a == k_b).all() and (n_a == n_b).all() and (o_a == o_b).all()
    okb = k_a.bool()
    assert 0.5 < okb.float().mean().item() and n_a.max().item() >= n_a.min().item() + 3
    assert (o_a[okb] == truth[okb]).all()
    idx = torch.arange(0, ncb, 2731)
    sub = llr[idx].cpu().numpy()
    o1, k1, n1, _ = port.decode_batch(sub, K, 8, "B", 0, True, nthreads=8)
    assert (o1 == o_a[idx].cpu().numpy()).all() and (k1 == k_a[idx].cpu().numpy()).all() and (n1 == n_a[idx].cpu().numpy()).all()
