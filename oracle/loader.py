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


def api(which: str = "port") -> _Api:
    return _Api(which)
