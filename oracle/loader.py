"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the two CPU checkers.

* ``port()``  -> oracle/liboracle_port.so : this repo's plain-C restatement (oracle_port.c)
* ``ref()``   -> oracle/_ref/libsrsref.so : the reference's own sources compiled in place (Makefile target ``ref``)

Only tests/, ``__graft_entry__.smoke()`` and bench.py's ``cpu_baseline`` / ``--impl reference`` legs may import
this module.  The product package ``srslte_b200`` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "libsrsref.so")
REFERENCE_ROOT = os.environ.get("SRSLTE_REFERENCE_ROOT", "/root/reference")

c_vp = C.c_void_p


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_vp)


def build(verbose: bool = False) -> None:
    """Compile the port always, and the reference library when the reference tree is present."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "port"], stdout=out)
    if os.path.exists(os.path.join(REFERENCE_ROOT, "lib", "include", "srsran", "config.h")):
        subprocess.check_call(["make", "-C", HERE, "ref", f"REF={REFERENCE_ROOT}", "-j8"], stdout=out)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


@lru_cache(maxsize=None)
def port() -> C.CDLL:
    if not os.path.exists(PORT_SO):
        build()
    lib = C.CDLL(PORT_SO)
    lib.orc_crc24.restype = C.c_uint32
    return lib


@lru_cache(maxsize=None)
def ref() -> C.CDLL:
    if not os.path.exists(REF_SO):
        build()
    lib = C.CDLL(REF_SO)
    lib.ref_crc_byte.restype = C.c_uint32
    lib.ref_crc_bits.restype = C.c_uint32
    return lib


CRC24A = 0x1864CFB
CRC24B = 0x1800063

# srsran_tdec_impl_type_t (turbodecoder_impl.h:27-37)
TDEC_AUTO, TDEC_GENERIC, TDEC_SSE, TDEC_SSE_WINDOW, TDEC_NEON_WINDOW, TDEC_AVX_WINDOW = 0, 1, 2, 3, 4, 5


class _Api:
    """Same Python surface over either library, so tests can diff them call by call."""

    def __init__(self, which: str):
        self.which = which
        self.lib = port() if which == "port" else ref()
        self.pfx = "orc_" if which == "port" else "ref_"

    def f(self, name):
        return getattr(self.lib, self.pfx + name)

    # -- tables ---------------------------------------------------------------------------------------
    def cb_sizes(self) -> np.ndarray:
        return np.array([self.f("cbsize")(C.c_uint32(i)) for i in range(self.f("nof_cb_sizes")())], dtype=np.int64)

    def cbindex(self, K: int) -> int:
        return int(self.f("cbindex")(C.c_uint32(K)))

    def cbsegm(self, tbs: int) -> dict:
        out = np.zeros(9, np.uint32)
        r = self.f("cbsegm")(C.c_uint32(tbs), _p(out))
        keys = ["F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2", "tbs"]
        d = {k: int(v) for k, v in zip(keys, out)}
        d["ret"] = int(r)
        return d

    def interleaver(self, K: int):
        fwd = np.zeros(K, np.uint16)
        rev = np.zeros(K, np.uint16)
        r = self.f("interleaver")(C.c_uint32(K), _p(fwd), _p(rev))
        assert r == 0
        return fwd, rev

    def crc24(self, kind: str, data: np.ndarray, nbits: int) -> int:
        data = np.ascontiguousarray(data, np.uint8)
        if self.which == "port":
            return int(self.lib.orc_crc24(C.c_int(0 if kind == "A" else 1), _p(data), C.c_int(nbits)))
        return int(self.lib.ref_crc_byte(C.c_uint32(CRC24A if kind == "A" else CRC24B), C.c_int(24), _p(data), C.c_int(nbits)))

    # -- coding ---------------------------------------------------------------------------------------
    def tcod_encode(self, bits: np.ndarray) -> np.ndarray:
        bits = np.ascontiguousarray(bits, np.uint8)
        K = bits.size
        out = np.zeros(3 * K + 12, np.uint8)
        r = self.f("tcod_encode")(_p(bits), _p(out), C.c_uint32(K))
        assert r == 0
        return out

    def rm_table(self, cb_idx: int, rv: int) -> np.ndarray:
        K = int(self.f("cbsize")(C.c_uint32(cb_idx)))
        t = np.zeros(3 * K + 12, np.uint16)
        r = self.f("rm_table")(C.c_uint32(cb_idx), C.c_uint32(rv), _p(t))
        assert r == 0
        return t

    def rm_tx(self, coded: np.ndarray, K: int, E: int, rv: int) -> np.ndarray:
        coded = np.ascontiguousarray(coded, np.uint8)
        out = np.zeros(E, np.uint8)
        r = self.f("rm_tx")(_p(coded), C.c_uint32(K), _p(out), C.c_uint32(E), C.c_uint32(rv))
        assert r == 0
        return out

    def rm_rx(self, e: np.ndarray, soft: np.ndarray, cb_idx: int, rv: int) -> int:
        """soft (int16, >= 3K+12, modified in place) += dematch(e); natural layout."""
        e = np.ascontiguousarray(e, np.int16)
        assert soft.dtype == np.int16
        if self.which == "port":
            return int(self.lib.orc_rm_rx(_p(e), _p(soft), C.c_uint32(e.size), C.c_uint32(cb_idx), C.c_uint32(rv)))
        return int(self.lib.ref_rm_rx(_p(e), _p(soft), C.c_uint32(e.size), C.c_uint32(cb_idx), C.c_uint32(rv), C.c_int(1)))

    # -- decoder --------------------------------------------------------------------------------------
    def tdec_passes(self, llr: np.ndarray, K: int, npass: int, impl: int = TDEC_GENERIC) -> np.ndarray:
        """Decided bytes after each of npass passes: (npass, K/8) uint8."""
        llr = np.ascontiguousarray(llr, np.int16)
        assert llr.size == 3 * K + 12
        out = np.zeros((npass, K // 8), np.uint8)
        if self.which == "port":
            r = self.lib.orc_tdec_passes(_p(llr), C.c_uint32(K), C.c_uint32(npass), _p(out))
        else:
            r = self.lib.ref_tdec_passes(C.c_int(impl), _p(llr), C.c_uint32(K), C.c_uint32(npass), _p(out))
        assert r == 0
        return out

    def decode_batch(self, llr: np.ndarray, K: int, max_pass: int = 8, crc: str = "B", crc_len: int = 0,
                     early_stop: bool = True, nthreads: int = 1, impl: int = TDEC_GENERIC):
        """decode_tb_cb-style loop over (ncb, 3K+12) int16.  Returns bytes (ncb,K/8), crc_ok, npass, seconds."""
        llr = np.ascontiguousarray(llr, np.int16).reshape(-1, 3 * K + 12)
        ncb = llr.shape[0]
        out = np.zeros((ncb, K // 8), np.uint8)
        ok = np.zeros(ncb, np.uint8)
        npass = np.zeros(ncb, np.uint8)
        sec = C.c_double(0)
        kind = {"B": 0, "A": 1, None: 2, "none": 2}[crc]
        args = [_p(llr), C.c_uint32(ncb), C.c_uint32(K), C.c_uint32(max_pass), C.c_int(kind), C.c_uint32(crc_len),
                C.c_int(1 if early_stop else 0), _p(out), _p(ok), _p(npass), C.c_int(nthreads), C.byref(sec)]
        if self.which == "port":
            r = self.lib.orc_decode_batch(*args)
        else:
            r = self.lib.ref_decode_batch(C.c_int(impl), *args)
        assert r == 0
        return out, ok, npass, sec.value

    def decode_tb(self, e_bits: np.ndarray, tbs: int, Qm: int, rv: int, max_iter: int, soft: np.ndarray, cb_crc: np.ndarray,
                  data: np.ndarray):
        """Port only: decode_tb/decode_tb_cb (sch.c:370-572).  soft (C*18600 int16), cb_crc (C uint8), data (bytes) are in/out.
        Returns (ret, iter_sum)."""
        assert self.which == "port"
        e_bits = np.ascontiguousarray(e_bits, np.int16)
        it = C.c_uint32(0)
        r = self.lib.orc_decode_tb(_p(e_bits), C.c_uint32(e_bits.size), C.c_uint32(tbs), C.c_uint32(Qm), C.c_uint32(rv),
                                   C.c_uint32(max_iter), _p(soft), _p(cb_crc), _p(data), C.byref(it))
        return int(r), int(it.value)

    # -- OFDM / demap ---------------------------------------------------------------------------------
    def ofdm_rx(self, x: np.ndarray, nof_prb: int, cp_ext: bool = False, symbol_sz: int = 0, freq_shift: float = 0.0,
                rx_window_offset: float = 0.0, normalize: bool = False, keep_dc: bool = False):
        x = np.ascontiguousarray(x, np.complex64)
        N = symbol_sz or self.symbol_sz(nof_prb)
        nsf = x.size // (15 * N)
        nsym = 12 if cp_ext else 14
        out = np.zeros((nsf, nsym, 12 * nof_prb), np.complex64)
        args = [C.c_uint32(nof_prb), C.c_int(int(cp_ext)), C.c_uint32(symbol_sz), C.c_float(freq_shift),
                C.c_float(rx_window_offset), C.c_int(int(normalize)), C.c_int(int(keep_dc)), _p(x), _p(out), C.c_uint32(nsf)]
        if self.which == "port":
            r = self.lib.orc_ofdm_rx(*args)
            sec = 0.0
        else:
            s = C.c_double(0)
            r = self.lib.ref_ofdm_rx(*args, C.byref(s))
            sec = s.value
        assert r == 0
        return out, sec

    def symbol_sz(self, nof_prb: int) -> int:
        if self.which == "port":
            return int(self.lib.orc_symbol_sz(C.c_uint32(nof_prb), C.c_int(0)))
        return int(self.lib.ref_symbol_sz(C.c_uint32(nof_prb)))

    def demod_s(self, mod: int, sym: np.ndarray) -> np.ndarray:
        sym = np.ascontiguousarray(sym, np.complex64)
        bps = {1: 2, 2: 4, 3: 6, 4: 8}[mod]
        # the reference's SSE path uses aligned loads/stores: hand it 16-byte aligned buffers
        raw_in = np.zeros(sym.size * 2 + 8, np.float32)
        off = (-raw_in.ctypes.data // 4) % 4
        a = raw_in[off:off + 2 * sym.size]
        a[:] = sym.view(np.float32)
        raw_out = np.zeros(sym.size * bps + 16, np.int16)
        off2 = (-raw_out.ctypes.data // 2) % 8
        o = raw_out[off2:off2 + sym.size * bps]
        r = self.f("demod_s")(C.c_int(mod), a.ctypes.data_as(c_vp), o.ctypes.data_as(c_vp), C.c_int(sym.size))
        assert r == 0
        return o.copy()

    # -- PUSCH chain between OFDM and de-matching ---------------------------------------------
    def pusch_seq_apply_s(self, x: np.ndarray, rnti: int, nslot: int, cell_id: int) -> np.ndarray:
        x = _aligned_copy(x, np.int16)
        out = _aligned(x.size, np.int16)
        self.f("pusch_seq_apply_s")(_p(x), _p(out), C.c_uint32(rnti), C.c_uint32(nslot), C.c_uint32(cell_id), C.c_uint32(x.size))
        return out.copy()

    def ulsch_deinterleave(self, q: np.ndarray, Qm: int, nof_symb: int) -> np.ndarray:
        q = _aligned_copy(q, np.int16)
        g = _aligned(q.size, np.int16)
        r = self.f("ulsch_deinterleave")(_p(q), _p(g), C.c_uint32(Qm), C.c_uint32(q.size // Qm), C.c_uint32(nof_symb))
        assert r == 0
        return g.copy()

    def dft_precoding(self, x: np.ndarray, nof_prb: int, is_tx: bool) -> np.ndarray:
        x = _aligned_copy(x, np.complex64)
        nsym = x.size // (12 * nof_prb)
        out = _aligned(x.size, np.complex64)
        r = self.f("dft_precoding")(_p(x), _p(out), C.c_uint32(nof_prb), C.c_uint32(nsym), C.c_int(int(is_tx)))
        assert r == 0
        return out.copy()

    def predecoding_single(self, y: np.ndarray, h: np.ndarray, noise: float, scaling: float = 1.0) -> np.ndarray:
        y = _aligned_copy(y, np.complex64)
        h = _aligned_copy(h, np.complex64)
        x = _aligned(y.size, np.complex64)
        r = self.f("predecoding_single")(_p(y), _p(h), _p(x), C.c_int(y.size), C.c_float(scaling), C.c_float(noise))
        assert r == y.size
        return x.copy()

    def dmrs_pusch_gen(self, link: np.ndarray) -> np.ndarray:
        link = np.ascontiguousarray(link, np.uint32)
        r = _aligned(2 * 12 * int(link[9]), np.complex64)
        ret = self.f("dmrs_pusch_gen")(_p(link), _p(r))
        assert ret == 0, ret
        return r.copy()

    def chest_ul_pusch(self, link: np.ndarray, grid: np.ndarray, dmrs: np.ndarray | None = None):
        """Returns (ce grid, [noise_estimate, snr, cfo_hz, ta_us]).  The port takes the known DMRS as an argument."""
        link = np.ascontiguousarray(link, np.uint32)
        grid = _aligned_copy(grid, np.complex64)
        ce = _aligned(grid.size, np.complex64)
        meas = np.zeros(8, np.float32)
        if self.which == "port":
            dmrs = _aligned_copy(dmrs, np.complex64)
            r = self.lib.orc_chest_ul_pusch(_p(link), _p(grid), _p(dmrs), _p(ce), _p(meas))
        else:
            r = self.lib.ref_chest_ul_pusch(_p(link), _p(grid), _p(ce), _p(meas))
        assert r == 0, r
        return ce.copy(), meas[:4].copy()

    def pusch_encode(self, link: np.ndarray, data: np.ndarray) -> np.ndarray:
        """Reference only: srsran_pusch_encode + DMRS into one subframe's resource grid (nsymb*2, 12*nof_prb)."""
        assert self.which == "ref"
        link = np.ascontiguousarray(link, np.uint32)
        nsym = 12 if link[2] else 14
        data = np.concatenate([np.ascontiguousarray(data, np.uint8), np.zeros(8, np.uint8)])
        grid = _aligned(nsym * 12 * int(link[1]), np.complex64)
        r = self.lib.ref_pusch_encode(_p(link), _p(data), _p(grid))
        assert r == 0, r
        return grid.reshape(nsym, -1).copy()

    def pusch_encode_uci(self, link: np.ndarray, uci: np.ndarray, data: np.ndarray) -> np.ndarray:
        """Reference only: srsran_pusch_encode with control information (pusch_uci) multiplexed in, + DMRS."""
        assert self.which == "ref"
        link = np.ascontiguousarray(link, np.uint32)
        uci = np.ascontiguousarray(uci, np.uint32)
        nsym = 12 if link[2] else 14
        data = np.concatenate([np.ascontiguousarray(data, np.uint8), np.zeros(8, np.uint8)])
        grid = _aligned(nsym * 12 * int(link[1]), np.complex64)
        r = self.lib.ref_pusch_encode_uci(_p(link), _p(uci), _p(data), _p(grid))
        assert r == 0, r
        return grid.reshape(nsym, -1).copy()

    def pusch_decode_uci(self, link: np.ndarray, uci: np.ndarray, grid: np.ndarray, identity_ce: bool = False) -> dict:
        """Reference only: (chest +) srsran_pusch_decode of one subframe that carries HARQ-ACK / RI / CQI."""
        assert self.which == "ref"
        link = np.ascontiguousarray(link, np.uint32)
        uci = np.ascontiguousarray(uci, np.uint32)
        Qm = {1: 2, 2: 4, 3: 6}[int(link[11])]
        nsym = 12 if link[2] else 14
        nre = (nsym - 2 - (int(link[16]) if link.size > 16 else 0)) * 12 * int(link[9])
        grid = _aligned_copy(grid, np.complex64)
        data = np.zeros(int(link[12]) // 8 + 16, np.uint8)
        crc = C.c_int(0)
        meas = np.zeros(8, np.float32)
        q = _aligned(nre * Qm, np.int16)
        g = _aligned(nre * Qm, np.int16)
        out = np.zeros(14 + 256, np.int32)
        r = self.lib.ref_pusch_decode_uci(_p(link), _p(uci), _p(grid), C.c_int(int(identity_ce)), _p(data), C.byref(crc), _p(meas),
                                          _p(q), _p(g), _p(out))
        return dict(ret=r, crc=bool(crc.value), data=data[:int(link[12]) // 8].copy(), noise=float(meas[0]), q=q.copy(), g=g.copy(),
                    ack=out[:10].copy(), ack_valid=bool(out[10]), ri=int(out[11]), cqi_crc=bool(out[12]),
                    cqi_bits=out[14:14 + int(out[13])].astype(np.uint8))

    def uci_decode_ack_ri(self, link: np.ndarray, q_bits: np.ndarray, c_seq: np.ndarray, beta: float, nof_bits: int, is_ri: bool):
        """Reference only: srsran_uci_decode_ack_ri on a whole descrambled subframe.  Returns (Q', data bits, valid)."""
        assert self.which == "ref"
        link = np.ascontiguousarray(link, np.uint32)
        q = _aligned_copy(q_bits, np.int16)
        c = _aligned_copy(c_seq, np.uint8)
        data = np.full(16, 2, np.uint8)
        valid = C.c_int(0)
        r = self.lib.ref_uci_decode_ack_ri(_p(link), _p(q), _p(c), C.c_float(beta), C.c_uint32(nof_bits), C.c_int(int(is_ri)), _p(data),
                                           C.byref(valid))
        return r, data[:nof_bits].copy(), bool(valid.value)

    def uci_decode_cqi(self, link: np.ndarray, q_bits: np.ndarray, beta: float, Q_prime_ri: int, cqi_len: int):
        """Reference only: srsran_uci_decode_cqi_pusch on the front of the de-interleaved stream.  Returns (Q', bits, crc)."""
        assert self.which == "ref"
        link = np.ascontiguousarray(link, np.uint32)
        q = _aligned_copy(q_bits, np.int16)
        data = np.zeros(64, np.uint8)
        crc = C.c_int(0)
        r = self.lib.ref_uci_decode_cqi(_p(link), _p(q), C.c_float(beta), C.c_uint32(Q_prime_ri), C.c_uint32(cqi_len), _p(data), C.byref(crc))
        return r, data[:cqi_len].copy(), bool(crc.value)

    def qprime_ack(self, L_prb: int, nof_symbols: int, K_segm: int, nof_ack: int, beta: float) -> int:
        assert self.which == "ref"
        f = self.lib.ref_qprime_ack
        f.restype = C.c_uint32
        return int(f(C.c_uint32(L_prb), C.c_uint32(nof_symbols), C.c_uint32(K_segm), C.c_uint32(nof_ack), C.c_float(beta)))

    def pusch_decode(self, link: np.ndarray, grid: np.ndarray, identity_ce: bool = False) -> dict:
        """Reference only: (chest +) srsran_pusch_decode of one subframe, with the object's intermediate buffers."""
        assert self.which == "ref"
        link = np.ascontiguousarray(link, np.uint32)
        Qm = {1: 2, 2: 4, 3: 6}[int(link[11])]
        nsym = 12 if link[2] else 14
        nre = (nsym - 2 - (int(link[16]) if link.size > 16 else 0)) * 12 * int(link[9])
        grid = _aligned_copy(grid, np.complex64)
        data = np.zeros(int(link[12]) // 8 + 16, np.uint8)
        crc = C.c_int(0)
        meas = np.zeros(8, np.float32)
        d = _aligned(nre, np.complex64)
        q = _aligned(nre * Qm, np.int16)
        g = _aligned(nre * Qm, np.int16)
        ce = _aligned(grid.size, np.complex64)
        r = self.lib.ref_pusch_decode(_p(link), _p(grid), C.c_int(int(identity_ce)), _p(data), C.byref(crc), _p(meas), _p(d), _p(q),
                                      _p(g), _p(ce))
        return dict(ret=int(r), crc=bool(crc.value), data=data[: int(link[12]) // 8].copy(), noise=float(meas[0]), snr=float(meas[1]),
                    cfo_hz=float(meas[2]), avg_iter=float(meas[4]), d=d.copy(), q=q.copy(), g=g.copy(), ce=ce.reshape(nsym, -1).copy())

    def pusch_rx_bench(self, link: np.ndarray, grids: np.ndarray, nthreads: int):
        """Reference only: chest + srsran_pusch_decode over grids (nsf, nsym, 12 nof_prb) on nthreads workers.
        Returns (crc verdicts (nsf,), seconds of the decode loops)."""
        assert self.which == "ref"
        link = np.ascontiguousarray(link, np.uint32)
        nsf = grids.shape[0]
        g = _aligned_copy(grids, np.complex64)
        ok = np.zeros(nsf, np.uint8)
        sec = C.c_double(0)
        r = self.lib.ref_pusch_rx_bench(_p(link), _p(g), C.c_uint32(nsf), C.c_int(nthreads), _p(ok), C.byref(sec))
        assert r == 0, r
        return ok, sec.value


def pusch_link(cell_id=1, nof_prb=100, cp_ext=0, cyclic_shift=0, delta_ss=0, group_hopping=0, sequence_hopping=0, rnti=62, tti=0,
               L_prb=100, n_prb=0, mod=3, tbs=75376, rv=0, n_dmrs=0, max_iter=8, shortened=0) -> np.ndarray:
    """The uint32 parameter block shared by ref_harness.c and oracle_port.c (mod: 1 QPSK, 2 16QAM, 3 64QAM)."""
    return np.array([cell_id, nof_prb, cp_ext, cyclic_shift, delta_ss, group_hopping, sequence_hopping, rnti, tti, L_prb, n_prb, mod,
                     tbs, rv, n_dmrs, max_iter, shortened], np.uint32)


def pusch_uci(nof_ack=0, ack_bits=0, ri_len=0, ri=0, cqi_kind=0, cqi_N=0, cqi_wb=0, cqi_sb=0, I_offset_ack=9, I_offset_ri=5,
              I_offset_cqi=6) -> np.ndarray:
    """The uint32 UCI parameter block of ref_harness.c (cqi_kind: 0 none, 1 wideband, 2 wideband + PMI, 3 higher-layer subband)."""
    return np.array([nof_ack, ack_bits, ri_len, ri, cqi_kind, cqi_N, cqi_wb, cqi_sb, I_offset_ack, I_offset_ri, I_offset_cqi], np.uint32)


def _aligned(n: int, dtype, align: int = 64) -> np.ndarray:
    """Zeroed 1-D array whose data pointer is `align`-byte aligned (the reference's AVX kernels use aligned loads)."""
    item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * item].view(dtype)


def _aligned_copy(a: np.ndarray, dtype) -> np.ndarray:
    a = np.asarray(a, dtype).reshape(-1)
    out = _aligned(a.size, dtype)
    out[:] = a
    return out


def api(which: str = "port") -> _Api:
    return _Api(which)
