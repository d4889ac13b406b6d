"""The 8-bit LLR entries of the reference-named API: srsran_tdec_run_all_8bit / srsran_tdec_iteration_8bit /
srsran_rm_turbo_rx_lut_8bit.  This library widens the int8 values and decodes with the generic int16 arithmetic -- the route the
reference itself takes when its 8-bit window decoders cannot handle a length (convert_8_to_16, turbodecoder.c:441-470).  So:
  * parity: bit-exact with the 16-bit entry on the widened values (and therefore with the reference's generic decoder);
  * BER-equivalence criterion against the reference's own 8-bit path (its saturating 8-bit window decoders where they apply,
    libsrsref.so loaded side by side): on the same quantised inputs the block error rate must not be worse, beyond the
    statistical slack of the sample;
  * rate de-matching: the int8 accumulate wraps modulo 256 like the reference's scalar form.
-m gpu."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import coded_llrs

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class TdecHandle:
    """srsran_tdec_t as an opaque, generously sized buffer (turbodecoder.h:63-95 is ~18 KB: it embeds 4 x 188 interleaver tables)."""

    def __init__(self, L, max_k=6144):
        self.L = L
        self.buf = C.create_string_buffer(1 << 16)
        L.srsran_tdec_init.argtypes = [C.c_void_p, C.c_uint32]
        L.srsran_tdec_free.argtypes = [C.c_void_p]
        L.srsran_tdec_free.restype = None
        L.srsran_tdec_force_not_sb.argtypes = [C.c_void_p]
        L.srsran_tdec_force_not_sb.restype = None
        L.srsran_tdec_run_all.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.srsran_tdec_run_all_8bit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.srsran_tdec_new_cb.argtypes = [C.c_void_p, C.c_uint32]
        L.srsran_tdec_iteration_8bit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.srsran_tdec_iteration_8bit.restype = None
        assert L.srsran_tdec_init(self.buf, max_k) == 0
        L.srsran_tdec_force_not_sb(self.buf)  # natural input layout (what the rate de-matcher of this library produces)

    def run8(self, llr8, K, iters):
        out = np.zeros(K // 8, np.uint8)
        x = np.ascontiguousarray(llr8, np.int8)
        assert self.L.srsran_tdec_run_all_8bit(self.buf, x.ctypes.data, out.ctypes.data, iters, K) == 0
        return out

    def run16(self, llr16, K, iters):
        out = np.zeros(K // 8, np.uint8)
        x = np.ascontiguousarray(llr16, np.int16)
        assert self.L.srsran_tdec_run_all(self.buf, x.ctypes.data, out.ctypes.data, iters, K) == 0
        return out

    def close(self):
        self.L.srsran_tdec_free(self.buf)


@pytest.fixture(scope="module")
def libs(ref):
    from srslte_b200.build import LIB_PATH

    return C.CDLL(LIB_PATH), C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libsrsref.so"))


@pytest.mark.parametrize("K,sigma,nblk", [(504, 0.80, 120), (2048, 0.84, 80), (6144, 0.86, 60)])
def test_8bit_decoder_entries(libs, port, K, sigma, nblk):
    ours, theirs = libs
    llr, bits = coded_llrs(port, K, nblk, sigma, 16.0, 31, seed=K)   # |LLR| <= 31: fits the 8-bit container
    a, b = TdecHandle(ours), TdecHandle(theirs)
    try:
        want = np.packbits(bits, axis=1)
        err_ours = err_ref = 0
        for i in range(nblk):
            o8 = a.run8(llr[i].astype(np.int8), K, 8)
            o16 = a.run16(llr[i], K, 8)
            assert (o8 == o16).all(), i                                # parity: the 16-bit entry on the widened values
            if i < 8:                                                  # ... which is the oracle's generic decoder
                o1, _, _, _ = port.decode_batch(llr[i:i + 1], K, 8, None, 0, False)
                assert (o1[0] == o8).all()
            r8 = b.run8(llr[i].astype(np.int8), K, 8)                  # the reference's own 8-bit path on the same values
            err_ours += int((o8 != want[i]).any())
            err_ref += int((r8 != want[i]).any())
        # BER-equivalence criterion: not worse than the reference's 8-bit decoders, beyond two standard deviations of the sample
        slack = 2.0 * np.sqrt(max(err_ref, 1))
        assert err_ours <= err_ref + slack, (err_ours, err_ref)
        # iteration by iteration (sch.c:426: srsran_tdec_iteration_8bit), decisions after every pass
        x8 = np.ascontiguousarray(llr[0].astype(np.int8))
        assert ours.srsran_tdec_new_cb(a.buf, K) == 0
        out = np.zeros(K // 8, np.uint8)
        for p in range(1, 5):
            ours.srsran_tdec_iteration_8bit(a.buf, x8.ctypes.data, out.ctypes.data)
            o1, _, _, _ = port.decode_batch(llr[0:1], K, p, None, 0, False)
            assert (o1[0] == out).all(), p
    finally:
        a.close()
        b.close()


def test_8bit_rate_dematching_wraps_like_the_reference(libs, port):
    ours, theirs = libs
    for L in (ours, theirs):
        L.srsran_rm_turbo_rx_lut_8bit.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.srsran_rm_turbo_gentables.restype = None
        L.srsran_rm_turbo_gentables()
    rng = np.random.default_rng(3)
    sizes = [int(k) for k in port.cb_sizes()]
    for cb_idx in (0, 5, 20, 44):        # K <= 400: the reference's 8-bit entry uses the natural-order table there too
        K = sizes[cb_idx]
        n = 3 * K + 12
        for rv in range(4):
            for E in (n // 2, n, 3 * n + 17):
                e = rng.integers(-128, 128, E).astype(np.int8)
                init = rng.integers(-128, 128, n).astype(np.int8)
                a, b = init.copy(), init.copy()
                assert ours.srsran_rm_turbo_rx_lut_8bit(e.ctypes.data, a.ctypes.data, E, cb_idx, rv) == 0
                assert theirs.srsran_rm_turbo_rx_lut_8bit(e.ctypes.data, b.ctypes.data, E, cb_idx, rv) == 0
                assert (a == b).all(), (K, rv, E)
    # larger blocks: natural layout (this library never uses the sub-block layouts), checked against the 16-bit oracle modulo 256
    for cb_idx in (100, 187):
        K = sizes[cb_idx]
        n = 3 * K + 12
        e = rng.integers(-128, 128, 2 * n + 5).astype(np.int8)
        init = rng.integers(-128, 128, n).astype(np.int8)
        a = init.copy()
        assert ours.srsran_rm_turbo_rx_lut_8bit(e.ctypes.data, a.ctypes.data, e.size, cb_idx, 1) == 0
        want = np.concatenate([init.astype(np.int16), np.zeros(64, np.int16)])
        port.rm_rx(e.astype(np.int16), want, cb_idx, 1)
        assert (a == want[:n].astype(np.int8)).all()
    assert ours.srsran_rm_turbo_rx_lut_8bit(e.ctypes.data, a.ctypes.data, 10, 0, 4) == -2
