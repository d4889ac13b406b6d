"""Python mirror of the OFDM receive and soft-demapper entries of include/srslte_b200.h."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class OfdmCfg(C.Structure):
    _fields_ = [("nof_prb", C.c_uint32), ("cp_ext", C.c_int), ("symbol_sz", C.c_uint32), ("freq_shift_f", C.c_float),
                ("rx_window_offset", C.c_float), ("normalize", C.c_int), ("keep_dc", C.c_int)]


class OfdmRx:
    """srsran_ofdm_t on the receive side (ofdm.h:69-86): srsran_ofdm_rx_init_cfg + srsran_ofdm_rx_sf for batches of subframes."""

    def __init__(self, nof_prb: int, cp_ext: bool = False, symbol_sz: int = 0, freq_shift_f: float = 0.0,
                 rx_window_offset: float = 0.0, normalize: bool = False, keep_dc: bool = False, device: int = 0):
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        self.device = device
        cfg = OfdmCfg(nof_prb, int(cp_ext), symbol_sz, freq_shift_f, rx_window_offset, int(normalize), int(keep_dc))
        rc = self._lib.srsran_b200_ofdm_rx_init(C.byref(self._h), device, C.byref(cfg))
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_ofdm_rx_init failed ({rc})")
        n, sf, ns, nre = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._lib.srsran_b200_ofdm_rx_geometry(self._h, C.byref(n), C.byref(sf), C.byref(ns), C.byref(nre))
        self.symbol_sz, self.sf_sz, self.nof_symbols, self.nof_re = n.value, sf.value, ns.value, nre.value

    def close(self):
        if self._h:
            self._lib.srsran_b200_ofdm_rx_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rx_sf(self, x: np.ndarray) -> np.ndarray:
        """x: (nsf*sf_sz,) complex64 host samples -> (nsf, nof_symbols, nof_re) complex64."""
        x = np.ascontiguousarray(x, np.complex64)
        nsf = x.size // self.sf_sz
        out = np.zeros((nsf, self.nof_symbols, self.nof_re), np.complex64)
        rc = self._lib.srsran_b200_ofdm_rx_sf_batch(self._h, x.ctypes.data, out.ctypes.data, nsf, 0, None)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_ofdm_rx_sf_batch failed ({rc})")
        return out

    def rx_sf_device(self, x, out, nsf: int, stream_ptr=None):
        """torch CUDA complex64 tensors; asynchronous on torch's current stream."""
        import torch

        if stream_ptr is None:
            stream_ptr = torch.cuda.current_stream(x.device).cuda_stream
        # int16 tensors (nsf, sf_sz, 2) are I/Q pairs in the radio's wire format (SRSRAN_B200_FLAG_IQ_INT16)
        fl = _lib.FLAG_DEVICE_PTRS | (_lib.FLAG_IQ_INT16 if x.dtype == torch.int16 else 0)
        rc = self._lib.srsran_b200_ofdm_rx_sf_batch(self._h, x.data_ptr(), out.data_ptr(), nsf, fl, stream_ptr)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_ofdm_rx_sf_batch failed ({rc})")


def demod_soft_s(mod: int, symbols: np.ndarray, symbols_per_call: int = 0, device: int = 0) -> np.ndarray:
    """srsran_demod_soft_demodulate_s for mod 1 (QPSK), 2 (16QAM), 3 (64QAM); host arrays."""
    symbols = np.ascontiguousarray(symbols, np.complex64)
    bps = {1: 2, 2: 4, 3: 6}[mod]
    out = np.zeros(symbols.size * bps, np.int16)
    rc = _lib.lib().srsran_b200_demod_soft_demodulate_s(device, mod, symbols.ctypes.data, out.ctypes.data, symbols.size,
                                                        symbols_per_call, 0, None)
    if rc != _lib.SUCCESS:
        raise RuntimeError(f"srsran_b200_demod_soft_demodulate_s failed ({rc})")
    return out
