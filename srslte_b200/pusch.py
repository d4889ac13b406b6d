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
        # the decode loop runs on the library's own stream: order it after the front end instead of synchronising, so that
        # the call's host-side bookkeeping (segmentation, de-matching descriptors) overlaps the front-end kernels
        self._lib.srsran_b200_sch_decode_after(self.sch._h, t.cuda.current_stream(self.dev).cuda_stream)
        rc = self._lib.srsran_b200_sch_decode_batch(self.sch._h, self.llr.data_ptr(), nsf * self.G, self.soft.data_ptr(),
                                                    nsf * soft_stride, self.data.data_ptr(), nsf * self.data_stride,
                                                    d.ctypes.data, nsf, _lib.FLAG_DEVICE_PTRS)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_sch_decode_batch failed ({rc})")
        return d["result"][:nsf] == 0, d["avg_iterations"][:nsf].copy()

    def decode_begin(self, nsf: int, rv: int = 0):
        """decode() without the wait (srsran_b200_sch_decode_begin): returns once the batch is queued behind the front end."""
        t = self.torch
        d = self.tb_np
        d["rv"][:nsf], d["new_data"][:nsf], d["cb_crc_mask"][:nsf] = rv, 1, 0
        self._lib.srsran_b200_sch_decode_after(self.sch._h, t.cuda.current_stream(self.dev).cuda_stream)
        rc = self._lib.srsran_b200_sch_decode_begin(self.sch._h, self.llr.data_ptr(), nsf * self.G, self.soft.data_ptr(),
                                                    nsf * self.soft.shape[1], self.data.data_ptr(), nsf * self.data_stride,
                                                    d.ctypes.data, nsf, _lib.FLAG_DEVICE_PTRS)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_sch_decode_begin failed ({rc})")
        self._begun = nsf

    def decode_finish(self):
        """Waits for the batch of decode_begin; returns what decode() returns."""
        rc = self._lib.srsran_b200_sch_decode_finish(self.sch._h)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_sch_decode_finish failed ({rc})")
        nsf, d = self._begun, self.tb_np
        return d["result"][:nsf] == 0, d["avg_iterations"][:nsf].copy()

    def run(self, iq, nsf: int, rv: int = 0):
        """iq: torch CUDA complex64 (nsf, sf_sz).  Returns (tb_ok (nsf,), avg passes (nsf,)); bytes are in self.data[:, :tbs/8+3]."""
        self.front_end(iq, nsf)
        return self.decode(nsf, rv)


class PuschCfg(C.Structure):
    """srsran_b200_pusch_cfg_t (include/srslte_b200.h)"""
    _fields_ = [("cell_id", C.c_uint32), ("cell_nof_prb", C.c_uint32), ("cp_ext", C.c_int), ("L_prb", C.c_uint32), ("n_prb", C.c_uint32),
                ("modulation", C.c_int), ("llr_shift", C.c_uint32), ("dmrs_cyclic_shift", C.c_uint32), ("dmrs_delta_ss", C.c_uint32),
                ("group_hopping_en", C.c_int), ("sequence_hopping_en", C.c_int), ("shortened", C.c_int)]


# srsran_b200_uci_cfg_t / srsran_b200_uci_value_t (include/srslte_b200.h)
UCI_CFG_DTYPE = np.dtype([("nof_ack", "<u4"), ("ri_len", "<u4"), ("cqi_len", "<u4"), ("I_offset_ack", "<u4"), ("I_offset_ri", "<u4"),
                          ("I_offset_cqi", "<u4")])
UCI_VALUE_DTYPE = np.dtype([("ack_value", "u1", (10,)), ("ack_valid", "u1"), ("ri", "u1"), ("cqi_crc", "u1"), ("reserved", "u1"),
                            ("cqi_bits", "u1", (64,)), ("Q_prime_ack", "<u4"), ("Q_prime_ri", "<u4"), ("Q_prime_cqi", "<u4"),
                            ("e_offset", "<u4"), ("nof_e_bits", "<u4")], align=True)
assert UCI_VALUE_DTYPE.itemsize == 100


def uci_cfg(nsf: int = 1, nof_ack=0, ri_len=0, cqi_len=0, I_offset_ack=9, I_offset_ri=5, I_offset_cqi=6) -> np.ndarray:
    """nsf equal srsran_b200_uci_cfg_t entries (edit single rows afterwards for a mixed batch)."""
    a = np.zeros(nsf, UCI_CFG_DTYPE)
    a["nof_ack"], a["ri_len"], a["cqi_len"] = nof_ack, ri_len, cqi_len
    a["I_offset_ack"], a["I_offset_ri"], a["I_offset_cqi"] = I_offset_ack, I_offset_ri, I_offset_cqi
    return a


class PuschChain:
    """Mirror of the srsran_b200_pusch_* entries: channel estimation -> equaliser + transform de-precoding -> soft demapping +
    descrambling + UL-SCH de-interleaving for a batch of subframes sharing one allocation (chest_ul.c:370, pusch.c:392-443,
    sch.c:993).  All tensors are torch CUDA tensors; rnti / tti / n_dmrs are numpy uint32 arrays (host)."""

    def __init__(self, cell_id=1, cell_nof_prb=100, cp_ext=False, L_prb=100, n_prb=0, mod=3, llr_shift=0, cyclic_shift=0, delta_ss=0,
                 group_hopping=False, sequence_hopping=False, device=0, shortened=False):
        import torch

        self.torch = torch
        self.device = device
        self.dev = torch.device("cuda", device)
        self._lib = _lib.lib()
        self.cfg = PuschCfg(cell_id, cell_nof_prb, int(cp_ext), L_prb, n_prb, mod, llr_shift, cyclic_shift, delta_ss, int(group_hopping),
                            int(sequence_hopping), int(shortened))
        self._h = C.c_void_p()
        rc = self._lib.srsran_b200_pusch_init(C.byref(self._h), device, C.byref(self.cfg))
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_init failed ({rc})")
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._lib.srsran_b200_pusch_geometry(self._h, C.byref(a), C.byref(b), C.byref(c))
        self.nof_re, self.nof_bits, self.nd = a.value, b.value, c.value
        self.M = 12 * L_prb
        self.nsym = 12 if cp_ext else 14
        self.R = 12 * cell_nof_prb

    def close(self):
        if self._h:
            self._lib.srsran_b200_pusch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _u32(a, n):
        if a is None:
            return None, None
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, np.uint32), (n,)))
        return a, a.ctypes.data

    def _st(self):
        return self.torch.cuda.current_stream(self.dev).cuda_stream

    def dmrs(self, sf_idx: int, n_dmrs: int = 0) -> np.ndarray:
        r = np.zeros(2 * self.M, np.complex64)
        rc = self._lib.srsran_b200_refsignal_dmrs_pusch_gen(self._h, sf_idx, n_dmrs, r.ctypes.data)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_refsignal_dmrs_pusch_gen failed ({rc})")
        return r.reshape(2, self.M)

    def chest(self, grid, tti, n_dmrs=None, out=None):
        t = self.torch
        nsf = grid.shape[0]
        ce, meas = out if out is not None else (t.empty((nsf, 2, self.M), dtype=t.complex64, device=self.dev),
                                                t.empty((nsf, 4), dtype=t.float32, device=self.dev))
        k1, p1 = self._u32(tti, nsf)
        k2, p2 = self._u32(n_dmrs, nsf)
        rc = self._lib.srsran_b200_chest_ul_pusch_batch(self._h, grid.data_ptr(), nsf, p1, p2, ce.data_ptr(), meas.data_ptr(),
                                                        _lib.FLAG_DEVICE_PTRS, self._st())
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_chest_ul_pusch_batch failed ({rc})")
        return ce, meas

    def equalize_deprecode(self, grid, ce, meas, out=None):
        t = self.torch
        nsf = grid.shape[0]
        d = out if out is not None else t.empty((nsf, self.nof_re), dtype=t.complex64, device=self.dev)
        rc = self._lib.srsran_b200_pusch_equalize_deprecode_batch(self._h, grid.data_ptr(), ce.data_ptr(),
                                                                  meas.data_ptr() if meas is not None else None, d.data_ptr(), nsf,
                                                                  _lib.FLAG_DEVICE_PTRS, self._st())
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_equalize_deprecode_batch failed ({rc})")
        return d

    def demod_descramble(self, d, rnti, tti, out=None):
        t = self.torch
        nsf = d.shape[0]
        g = out if out is not None else t.empty((nsf, self.nof_bits), dtype=t.int16, device=self.dev)
        k1, p1 = self._u32(rnti, nsf)
        k2, p2 = self._u32(tti, nsf)
        rc = self._lib.srsran_b200_pusch_demod_descramble_batch(self._h, d.data_ptr(), g.data_ptr(), nsf, p1, p2, _lib.FLAG_DEVICE_PTRS,
                                                                self._st())
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_demod_descramble_batch failed ({rc})")
        return g

    def rx(self, grid, rnti, tti, n_dmrs=None, out=None, meas=None):
        """grid (nsf, nsym, 12*cell_nof_prb) complex64 -> g (nsf, nof_bits) int16 (+ meas (nsf, 4) if a tensor is passed)."""
        t = self.torch
        nsf = grid.shape[0]
        g = out if out is not None else t.empty((nsf, self.nof_bits), dtype=t.int16, device=self.dev)
        k1, p1 = self._u32(rnti, nsf)
        k2, p2 = self._u32(tti, nsf)
        k3, p3 = self._u32(n_dmrs, nsf)
        rc = self._lib.srsran_b200_pusch_rx_batch(self._h, grid.data_ptr(), g.data_ptr(), meas.data_ptr() if meas is not None else None,
                                                  nsf, p1, p2, p3, _lib.FLAG_DEVICE_PTRS, self._st())
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_rx_batch failed ({rc})")
        return g


    def uci_geometry(self, tbs: int, uci: np.ndarray) -> np.ndarray:
        """Q' of the three fields and the UL-SCH span (e_offset, nof_e_bits) of one grant; host only."""
        uci = np.ascontiguousarray(uci, UCI_CFG_DTYPE).reshape(-1)
        out = np.zeros(1, UCI_VALUE_DTYPE)
        rc = self._lib.srsran_b200_pusch_uci_geometry(self._h, tbs, uci.ctypes.data, out.ctypes.data)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_uci_geometry failed ({rc})")
        return out[0]

    def rx_uci(self, grid, rnti, tti, n_dmrs, tbs, uci: np.ndarray, out=None, meas=None):
        """rx() for subframes with HARQ-ACK / RI / CQI multiplexed in: uci = UCI_CFG_DTYPE array (nsf,), tbs scalar or (nsf,).
        Returns g; the decided values come from uci_collect()."""
        t = self.torch
        nsf = grid.shape[0]
        g = out if out is not None else t.zeros((nsf, self.nof_bits), dtype=t.int16, device=self.dev)
        k1, p1 = self._u32(rnti, nsf)
        k2, p2 = self._u32(tti, nsf)
        k3, p3 = self._u32(n_dmrs, nsf)
        k4, p4 = self._u32(tbs, nsf)
        uci = np.ascontiguousarray(uci, UCI_CFG_DTYPE).reshape(nsf)
        rc = self._lib.srsran_b200_pusch_rx_uci_batch(self._h, grid.data_ptr(), g.data_ptr(), meas.data_ptr() if meas is not None else None,
                                                      nsf, p1, p2, p3, p4, uci.ctypes.data, _lib.FLAG_DEVICE_PTRS, self._st())
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_rx_uci_batch failed ({rc})")
        return g

    def uci_collect(self, nsf: int) -> np.ndarray:
        out = np.zeros(nsf, UCI_VALUE_DTYPE)
        rc = self._lib.srsran_b200_pusch_uci_collect(self._h, out.ctypes.data, nsf)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_pusch_uci_collect failed ({rc})")
        return out


class PuschRxFull(PuschRx):
    """Complete PUSCH receive pipeline on device buffers, one full-band allocation per (cell, subframe):

        srsran_b200_ofdm_rx_sf_batch -> srsran_b200_pusch_rx_batch (channel estimation, equaliser, transform de-precoding,
        soft demapping, descrambling, UL-SCH de-interleaving) -> srsran_b200_sch_decode_batch

    i.e. enb_ul.c:151-154 (srsran_enb_ul_fft) + enb_ul.c:262-290 (get_pusch: srsran_chest_ul_estimate_pusch, srsran_pusch_decode)
    for a batch of subframes of one cell."""

    def __init__(self, cell_id: int = 1, nof_prb: int = 100, tbs: int = 75376, mod: int = 3, llr_shift: int = 4, max_noi: int = 8,
                 device: int = 0, symbol_sz: int = 0, cyclic_shift: int = 0, delta_ss: int = 0):
        super().__init__(nof_prb, tbs, mod, llr_shift, max_noi, device, symbol_sz)
        self.chain = PuschChain(cell_id, nof_prb, False, nof_prb, 0, mod, llr_shift, cyclic_shift, delta_ss, False, False, device)
        assert self.chain.nof_bits == self.G
        self.meas = None

    def close(self):
        self.chain.close()
        super().close()

    def front_end(self, iq, nsf: int, rnti=None, tti=None, n_dmrs=None):
        t = self.torch
        self._reserve(nsf)
        if self.meas is None or self.meas.shape[0] < nsf:
            self.meas = t.empty((nsf, 4), dtype=t.float32, device=self.dev)
        st = t.cuda.current_stream(self.dev).cuda_stream
        self.ofdm.rx_sf_device(iq, self.grid, nsf, st)
        self.chain.rx(self.grid[:nsf], rnti if rnti is not None else 0, tti if tti is not None else 0, n_dmrs, out=self.llr,
                      meas=self.meas)

    def run(self, iq, nsf: int, rnti=None, tti=None, n_dmrs=None, rv: int = 0):
        self.front_end(iq, nsf, rnti, tti, n_dmrs)
        return self.decode(nsf, rv)

    def run_begin(self, iq, nsf: int, rnti=None, tti=None, n_dmrs=None, rv: int = 0):
        """run() in two halves: a caller thread that alternates between two objects (each on its own torch stream) keeps two
        batches in flight, so one batch's last passes -- a handful of blocks that fail their CRC and run all the passes -- overlap
        the bulk of the other."""
        self.front_end(iq, nsf, rnti, tti, n_dmrs)
        self.decode_begin(nsf, rv)

    def run_finish(self):
        return self.decode_finish()


class EnbUlCfg(C.Structure):
    """srsran_b200_enb_ul_cfg_t"""
    _fields_ = [("cell_id", C.c_uint32), ("cell_nof_prb", C.c_uint32), ("cp_ext", C.c_int), ("symbol_sz", C.c_uint32),
                ("dmrs_cyclic_shift", C.c_uint32), ("dmrs_delta_ss", C.c_uint32), ("group_hopping_en", C.c_int),
                ("sequence_hopping_en", C.c_int), ("L_prb", C.c_uint32), ("n_prb", C.c_uint32), ("modulation", C.c_int), ("tbs", C.c_uint32),
                ("llr_shift", C.c_uint32), ("max_iterations", C.c_uint32), ("shortened", C.c_int)]


PUSCH_RES_DTYPE = np.dtype([("crc_ok", "<i4"), ("avg_iterations", "<f4"), ("noise_estimate", "<f4"), ("snr", "<f4"), ("cfo_hz", "<f4")])


class EnbUl:
    """srsran_b200_enb_ul_*: the native one-call PUSCH receiver (time samples in, transport-block bytes out)."""

    def __init__(self, cell_id=1, nof_prb=100, tbs=75376, mod=3, llr_shift=4, max_noi=8, device=0, symbol_sz=0, L_prb=None, n_prb=0,
                 cyclic_shift=0, delta_ss=0, cp_ext=False, shortened=False):
        self._lib = _lib.lib()
        self.cfg = EnbUlCfg(cell_id, nof_prb, int(cp_ext), symbol_sz, cyclic_shift, delta_ss, 0, 0, L_prb if L_prb is not None else nof_prb,
                            n_prb, mod, tbs, llr_shift, max_noi, int(shortened))
        self._h = C.c_void_p()
        rc = self._lib.srsran_b200_enb_ul_init(C.byref(self._h), device, C.byref(self.cfg))
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_enb_ul_init failed ({rc})")
        a, b = C.c_uint32(), C.c_uint32()
        self._lib.srsran_b200_enb_ul_geometry(self._h, C.byref(a), C.byref(b))
        self.sf_sz, self.tb_bytes = a.value, b.value

    def close(self):
        if self._h:
            self._lib.srsran_b200_enb_ul_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_ptr(self, samples_ptr: int, nsf: int, rnti, tti, data_ptr: int, res: np.ndarray, n_dmrs=None, rv=None, new_data=None,
                flags: int = 0):
        """Raw pointers (host unless FLAG_DEVICE_PTRS); res: numpy array of PUSCH_RES_DTYPE with nsf entries."""
        k = [PuschChain._u32(a, nsf) for a in (rnti, tti, n_dmrs, rv, new_data)]
        rc = self._lib.srsran_b200_enb_ul_pusch_batch(self._h, samples_ptr, nsf, k[0][1], k[1][1], k[2][1], k[3][1], k[4][1], data_ptr,
                                                      res.ctypes.data, flags)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_enb_ul_pusch_batch failed ({rc})")

    def begin_ptr(self, samples_ptr: int, nsf: int, rnti, tti, data_ptr: int, res: np.ndarray, n_dmrs=None, rv=None, new_data=None,
                  flags: int = 0):
        """run_ptr without the wait (srsran_b200_enb_ul_pusch_batch_begin); finish() completes it.  The buffers behind data_ptr and
        res must stay alive until then."""
        k = [PuschChain._u32(a, nsf) for a in (rnti, tti, n_dmrs, rv, new_data)]
        rc = self._lib.srsran_b200_enb_ul_pusch_batch_begin(self._h, samples_ptr, nsf, k[0][1], k[1][1], k[2][1], k[3][1], k[4][1], None,
                                                            data_ptr, res.ctypes.data, None, flags)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_enb_ul_pusch_batch_begin failed ({rc})")

    def finish(self):
        rc = self._lib.srsran_b200_enb_ul_pusch_batch_finish(self._h)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_enb_ul_pusch_batch_finish failed ({rc})")

    def run(self, samples: np.ndarray, rnti, tti, n_dmrs=None, rv=None, new_data=None, uci: np.ndarray | None = None):
        """samples: (nsf, sf_sz) complex64 or (nsf, sf_sz, 2) int16 host array.  Returns (bytes (nsf, tb_bytes) uint8, results)
        and, with uci (UCI_CFG_DTYPE array (nsf,)), the decided control information (UCI_VALUE_DTYPE array) as a third item."""
        samples = np.ascontiguousarray(samples)
        nsf = samples.shape[0]
        fl = _lib.FLAG_IQ_INT16 if samples.dtype == np.int16 else 0
        data = np.zeros((nsf, self.tb_bytes), np.uint8)
        res = np.zeros(nsf, PUSCH_RES_DTYPE)
        if uci is None:
            self.run_ptr(samples.ctypes.data, nsf, rnti, tti, data.ctypes.data, res, n_dmrs, rv, new_data, fl)
            return data, res
        uci = np.ascontiguousarray(uci, UCI_CFG_DTYPE).reshape(nsf)
        val = np.zeros(nsf, UCI_VALUE_DTYPE)
        k = [PuschChain._u32(a, nsf) for a in (rnti, tti, n_dmrs, rv, new_data)]
        rc = self._lib.srsran_b200_enb_ul_pusch_uci_batch(self._h, samples.ctypes.data, nsf, k[0][1], k[1][1], k[2][1], k[3][1], k[4][1],
                                                          uci.ctypes.data, data.ctypes.data, res.ctypes.data, val.ctypes.data, fl)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_enb_ul_pusch_uci_batch failed ({rc})")
        return data, res, val
