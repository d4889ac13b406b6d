"""Python mirror of the shared-channel entries of include/srslte_b200.h (srsran_b200_sch_*, srsran_b200_rm_turbo_rx_batch)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

SOFTBUFFER_SIZE = 18600


class RmCb(C.Structure):
    _fields_ = [("cb_idx", C.c_uint32), ("rv", C.c_uint32), ("E", C.c_uint32), ("new_data", C.c_uint32),
                ("in_offset", C.c_uint64), ("soft_offset", C.c_uint64)]


class Tb(C.Structure):
    _fields_ = [("tbs", C.c_uint32), ("Qm", C.c_uint32), ("rv", C.c_uint32), ("nof_e_bits", C.c_uint32),
                ("e_offset", C.c_uint64), ("soft_offset", C.c_uint64), ("data_offset", C.c_uint64),
                ("new_data", C.c_uint32), ("cb_crc_mask", C.c_uint32), ("result", C.c_int32), ("nof_cb", C.c_uint32),
                ("avg_iterations", C.c_float)]


class SchDecoder:
    """Receive-side counterpart of srsran_sch_t: batched de-matching + transport block decode loop."""

    def __init__(self, device: int = 0, max_noi: int = 0):
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        rc = self._lib.srsran_b200_sch_init(C.byref(self._h), device)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_sch_init failed ({rc}): no usable CUDA device {device}?")
        if max_noi:
            self.set_max_noi(max_noi)

    def set_max_noi(self, n: int):
        self._lib.srsran_b200_sch_set_max_noi(self._h, n)

    def close(self):
        if self._h:
            self._lib.srsran_b200_sch_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def rm_rx(self, e_bits: np.ndarray, soft_pool: np.ndarray, jobs: list[dict]) -> int:
        """Host path.  jobs: dicts with cb_idx, rv, E, new_data, in_offset, soft_offset.  soft_pool is updated in place."""
        e_bits = np.ascontiguousarray(e_bits, np.int16)
        assert soft_pool.dtype == np.int16 and soft_pool.flags["C_CONTIGUOUS"]
        arr = (RmCb * len(jobs))(*[RmCb(j["cb_idx"], j["rv"], j["E"], int(j.get("new_data", 0)), j["in_offset"], j["soft_offset"])
                                  for j in jobs])
        return int(self._lib.srsran_b200_rm_turbo_rx_batch(self._h, e_bits.ctypes.data, e_bits.size, soft_pool.ctypes.data,
                                                           soft_pool.size, arr, len(jobs), 0, None))

    def decode(self, e_bits: np.ndarray, soft_pool: np.ndarray, data: np.ndarray, tbs: list[dict]):
        """Host path.  tbs: dicts with tbs, Qm, rv, nof_e_bits, e_offset, soft_offset, data_offset, new_data, cb_crc_mask.
        Returns (rc, list of result dicts); soft_pool and data are updated in place."""
        e_bits = np.ascontiguousarray(e_bits, np.int16)
        assert soft_pool.dtype == np.int16 and data.dtype == np.uint8
        arr = (Tb * len(tbs))(*[Tb(t["tbs"], t["Qm"], t["rv"], t["nof_e_bits"], t["e_offset"], t["soft_offset"], t["data_offset"],
                                   int(t.get("new_data", 1)), int(t.get("cb_crc_mask", 0)), 0, 0, 0.0) for t in tbs])
        rc = int(self._lib.srsran_b200_sch_decode_batch(self._h, e_bits.ctypes.data, e_bits.size, soft_pool.ctypes.data, soft_pool.size,
                                                        data.ctypes.data, data.size, arr, len(tbs), 0))
        res = [{"result": a.result, "nof_cb": a.nof_cb, "avg_iterations": a.avg_iterations, "cb_crc_mask": a.cb_crc_mask} for a in arr]
        return rc, res
