"""Identity-channel PUSCH receive pipeline of BASELINE.json config 4 / 5 on device buffers, built from the C ABI entries:

    srsran_b200_ofdm_rx_sf_batch  ->  srsran_b200_pusch_demap_batch  ->  srsran_b200_sch_decode_batch

i.e. what enb_ul.c:153 (srsran_ofdm_rx_sf), pusch.c:449 (srsran_demod_soft_demodulate_s) and sch.c:507-572 (decode_tb) do
for one subframe, for a batch of (cell, subframe) pairs.  Channel estimation, equaliser, transform de-precoding,
descrambling and UL-SCH de-interleaving are not part of it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .ofdm import OfdmRx
from .sch import SOFTBUFFER_SIZE, SchDecoder, Tb

# numpy image of srsran_b200_tb_t (include/srslte_b200.h)
TB_DTYPE = np.dtype([("tbs", "<u4"), ("Qm", "<u4"), ("rv", "<u4"), ("nof_e_bits", "<u4"), ("e_offset", "<u8"), ("soft_offset", "<u8"),
                     ("data_offset", "<u8"), ("new_data", "<u4"), ("cb_crc_mask", "<u4"), ("result", "<i4"), ("nof_cb", "<u4"),
                     ("avg_iterations", "<f4")], align=True)
assert TB_DTYPE.itemsize == C.sizeof(Tb)

DATA_SYMBOL_MASK = 0x3FFF & ~((1 << 3) | (1 << 10))  # normal CP: symbols 3 and 10 carry the DMRS


class PuschRx:
    def __init__(self, nof_prb: int = 100, tbs: int = 75376, mod: int = 3, llr_shift: int = 4, max_noi: int = 8, device: int = 0,
                 symbol_sz: int = 0):
        import torch

        self.torch = torch
        self.device = device
        self.dev = torch.device("cuda", device)
        self._lib = _lib.lib()
        self._lib.srsran_b200_pusch_demap_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                                            C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        # eNB uplink configuration (enb_ul.c:50-58): half-subcarrier shift -0.5, window advanced by half a CP, no normalisation
        self.ofdm = OfdmRx(nof_prb, False, symbol_sz, -0.5, 0.5, False, False, device)
        self.sch = SchDecoder(device, max_noi)
        self.mod, self.Qm, self.tbs, self.llr_shift = mod, 2 * mod, tbs, llr_shift
        self.nof_re = self.ofdm.nof_re
        self.G = 12 * self.nof_re * self.Qm
        self.data_stride = (tbs // 8 + 3 + 768 + 15) // 16 * 16
        self.C = None
        self._cap = 0

    def close(self):
        self.ofdm.close()
        self.sch.close()

    def _reserve(self, nsf: int):
        if nsf <= self._cap:
            return
        t = self.torch
        self.grid = t.empty((nsf, 14, self.nof_re), dtype=t.complex64, device=self.dev)
        self.llr = t.empty((nsf, self.G), dtype=t.int16, device=self.dev)
        self.soft = t.zeros((nsf, 13 * SOFTBUFFER_SIZE), dtype=t.int16, device=self.dev)  # up to 13 code blocks per TB here
        self.data = t.zeros((nsf, self.data_stride), dtype=t.uint8, device=self.dev)
        # transport block descriptors (srsran_b200_tb_t), filled once; only the in/out fields are reset per call
        self.tb_np = np.zeros(nsf, TB_DTYPE)
        i = np.arange(nsf, dtype=np.uint64)
        self.tb_np["tbs"], self.tb_np["Qm"], self.tb_np["nof_e_bits"] = self.tbs, self.Qm, self.G
        self.tb_np["e_offset"] = i * np.uint64(self.G)
        self.tb_np["soft_offset"] = i * np.uint64(self.soft.shape[1])
        self.tb_np["data_offset"] = i * np.uint64(self.data_stride)
        self._cap = nsf

    def front_end(self, iq, nsf: int):
        """OFDM demodulation + soft demapping of nsf subframes (asynchronous, torch's current stream)."""
        t = self.torch
        self._reserve(nsf)
        st = t.cuda.current_stream(self.dev).cuda_stream
        self.ofdm.rx_sf_device(iq, self.grid, nsf, st)
        rc = self._lib.srsran_b200_pusch_demap_batch(self.device, self.mod, self.grid.data_ptr(), self.llr.data_ptr(), nsf, 14,
                                                     self.nof_re, DATA_SYMBOL_MASK, self.llr_shift, _lib.FLAG_DEVICE_PTRS, st)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_demap_batch failed ({rc})")

    def decode(self, nsf: int, rv: int = 0):
        """Rate de-matching + transport block decode loop of the nsf subframes demapped last (synchronous)."""
        t = self.torch
        d = self.tb_np
        soft_stride = self.soft.shape[1]
        d["rv"][:nsf], d["new_data"][:nsf], d["cb_crc_mask"][:nsf] = rv, 1, 0
        t.cuda.current_stream(self.dev).synchronize()  # the decode loop runs on the library's own stream
        rc = self._lib.srsran_b200_sch_decode_batch(self.sch._h, self.llr.data_ptr(), nsf * self.G, self.soft.data_ptr(),
                                                    nsf * soft_stride, self.data.data_ptr(), nsf * self.data_stride,
                                                    d.ctypes.data, nsf, _lib.FLAG_DEVICE_PTRS)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_sch_decode_batch failed ({rc})")
        return d["result"][:nsf] == 0, d["avg_iterations"][:nsf].copy()

    def run(self, iq, nsf: int, rv: int = 0):
        """iq: torch CUDA complex64 (nsf, sf_sz).  Returns (tb_ok (nsf,), avg passes (nsf,)); bytes are in self.data[:, :tbs/8+3]."""
        self.front_end(iq, nsf)
        return self.decode(nsf, rv)
