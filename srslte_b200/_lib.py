"""ctypes binding of libsrslte_b200.so (the C ABI of include/srslte_b200.h).

There is no Python or CPU fallback: if the shared library is missing or cannot be loaded this module raises, so a
GPU box can never silently run something else.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

from .build import LIB_PATH

SUCCESS = 0
ERROR = -1
ERROR_INVALID_INPUTS = -2

FLAG_DEVICE_PTRS = 0x1
FLAG_LLR_INT8 = 0x4
FLAG_IQ_INT16 = 0x8
CRC_NONE, CRC24A, CRC24B = 0, 1, 2

vp = C.c_void_p
u32 = C.c_uint32


class LibraryMissing(RuntimeError):
    pass


@lru_cache(maxsize=None)
def lib() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} not built: run `python -m srslte_b200.build` (or __graft_entry__.build()). "
            "srslte_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.srsran_b200_device_count.restype = C.c_int
    L.srsran_b200_kernel_launches.restype = C.c_uint64
    L.srsran_b200_tdec_init.argtypes = [C.POINTER(vp), C.c_int, u32]
    L.srsran_b200_tdec_free.argtypes = [vp]
    L.srsran_b200_tdec_free.restype = None
    L.srsran_b200_tdec_run.argtypes = [vp, vp, u32, u32, u32, C.c_int, C.c_int, vp, vp, vp, u32, vp]
    L.srsran_b200_tdec_run_mixed.argtypes = [vp, vp, u32, vp, vp, u32, C.c_int, C.c_int, vp, vp, vp, u32, vp]
    L.srsran_b200_tdec_profile_get_ex.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.srsran_b200_tdec_profile_spans.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]
    L.srsran_b200_tdec_profile_reset.argtypes = [vp, C.c_int]
    L.srsran_b200_tdec_profile_reset.restype = None
    L.srsran_b200_tdec_profile_get.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    u64 = C.c_uint64
    L.srsran_b200_sch_init.argtypes = [C.POINTER(vp), C.c_int]
    L.srsran_b200_sch_free.argtypes = [vp]
    L.srsran_b200_sch_free.restype = None
    L.srsran_b200_sch_set_max_noi.argtypes = [vp, u32]
    L.srsran_b200_sch_set_max_noi.restype = None
    L.srsran_b200_sch_decode_after.argtypes = [vp, vp]
    L.srsran_b200_sch_decode_after.restype = None
    L.srsran_b200_sch_decode_after_event.argtypes = [vp, vp]
    L.srsran_b200_sch_decode_begin.argtypes = [vp, vp, u64, vp, u64, vp, u64, vp, u32, u32]
    L.srsran_b200_sch_decode_finish.argtypes = [vp]
    L.srsran_b200_sch_decode_after_event.restype = None
    L.srsran_b200_rm_turbo_rx_batch.argtypes = [vp, vp, u64, vp, u64, vp, u32, u32, vp]
    L.srsran_b200_sch_decode_batch.argtypes = [vp, vp, u64, vp, u64, vp, u64, vp, u32, u32]
    L.srsran_b200_use_standard_symbol_size.argtypes = [C.c_int]
    L.srsran_b200_use_standard_symbol_size.restype = None
    L.srsran_b200_symbol_sz.argtypes = [u32]
    L.srsran_b200_ofdm_rx_init.argtypes = [C.POINTER(vp), C.c_int, vp]
    L.srsran_b200_ofdm_rx_reconfigure.argtypes = [vp, vp]
    L.srsran_b200_ofdm_rx_free.argtypes = [vp]
    L.srsran_b200_ofdm_rx_free.restype = None
    L.srsran_b200_ofdm_rx_geometry.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.srsran_b200_ofdm_rx_sf_batch.argtypes = [vp, vp, vp, u32, u32, vp]
    L.srsran_b200_demod_soft_demodulate_s.argtypes = [C.c_int, C.c_int, vp, vp, u32, u32, u32, vp]
    L.srsran_b200_pusch_init.argtypes = [C.POINTER(vp), C.c_int, vp]
    L.srsran_b200_pusch_free.argtypes = [vp]
    L.srsran_b200_pusch_free.restype = None
    L.srsran_b200_pusch_geometry.argtypes = [vp, C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]
    L.srsran_b200_refsignal_dmrs_pusch_gen.argtypes = [vp, u32, u32, vp]
    L.srsran_b200_chest_ul_pusch_batch.argtypes = [vp, vp, u32, vp, vp, vp, vp, u32, vp]
    L.srsran_b200_pusch_equalize_deprecode_batch.argtypes = [vp, vp, vp, vp, vp, u32, u32, vp]
    L.srsran_b200_pusch_demod_descramble_batch.argtypes = [vp, vp, vp, u32, vp, vp, u32, vp]
    L.srsran_b200_pusch_rx_batch.argtypes = [vp, vp, vp, vp, u32, vp, vp, vp, u32, vp]
    L.srsran_b200_pusch_uci_geometry.argtypes = [vp, u32, vp, vp]
    L.srsran_b200_pusch_rx_uci_batch.argtypes = [vp, vp, vp, vp, u32, vp, vp, vp, vp, vp, u32, vp]
    L.srsran_b200_pusch_uci_collect.argtypes = [vp, vp, u32]
    L.srsran_b200_uci_decide.argtypes = [vp, u32, u32, u32, u32, vp, vp, vp, vp]
    L.srsran_b200_enb_ul_pusch_uci_batch.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32]
    L.srsran_b200_enb_ul_pusch_batch_begin.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32]
    L.srsran_b200_enb_ul_pusch_batch_finish.argtypes = [vp]
    L.srsran_b200_enb_ul_init.argtypes = [C.POINTER(vp), C.c_int, vp]
    L.srsran_b200_enb_ul_free.argtypes = [vp]
    L.srsran_b200_enb_ul_free.restype = None
    L.srsran_b200_enb_ul_geometry.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
    L.srsran_b200_enb_ul_pusch_batch.argtypes = [vp, vp, u32, vp, vp, vp, vp, vp, vp, vp, u32]
    return L


# every symbol include/srslte_b200.h declares; tests check the built library exports all of them
EXPORTED_SYMBOLS = [
    "srsran_b200_device_count",
    "srsran_b200_kernel_launches",
    "srsran_b200_tdec_init",
    "srsran_b200_tdec_free",
    "srsran_b200_tdec_run",
    "srsran_b200_tdec_run_mixed",
    "srsran_b200_tdec_profile_get_ex",
    "srsran_b200_tdec_profile_spans",
    "srsran_b200_tdec_profile_reset",
    "srsran_b200_tdec_profile_get",
    "srsran_b200_tdec_resident_tiles_per_sm",
    "srsran_b200_pusch_demap_batch",
    "srsran_b200_pusch_init",
    "srsran_b200_pusch_free",
    "srsran_b200_pusch_geometry",
    "srsran_b200_refsignal_dmrs_pusch_gen",
    "srsran_b200_chest_ul_pusch_batch",
    "srsran_b200_pusch_equalize_deprecode_batch",
    "srsran_b200_pusch_demod_descramble_batch",
    "srsran_b200_pusch_rx_batch",
    "srsran_b200_pusch_uci_geometry",
    "srsran_b200_pusch_rx_uci_batch",
    "srsran_b200_pusch_uci_collect",
    "srsran_b200_uci_decide",
    "srsran_b200_enb_ul_pusch_uci_batch",
    "srsran_b200_enb_ul_pusch_batch_begin",
    "srsran_b200_enb_ul_pusch_batch_finish",
    "srsran_b200_enb_ul_init",
    "srsran_b200_enb_ul_free",
    "srsran_b200_enb_ul_geometry",
    "srsran_b200_enb_ul_pusch_batch",
    "srsran_b200_sch_init",
    "srsran_b200_sch_free",
    "srsran_b200_sch_decode_after",
    "srsran_b200_sch_decode_after_event",
    "srsran_b200_sch_decode_begin",
    "srsran_b200_sch_decode_finish",
    "srsran_b200_sch_set_max_noi",
    "srsran_b200_rm_turbo_rx_batch",
    "srsran_b200_sch_decode_batch",
    "srsran_b200_use_standard_symbol_size",
    "srsran_b200_symbol_sz",
    "srsran_b200_ofdm_rx_init",
    "srsran_b200_ofdm_rx_reconfigure",
    "srsran_b200_ofdm_rx_free",
    "srsran_b200_ofdm_rx_geometry",
    "srsran_b200_ofdm_rx_sf_batch",
    "srsran_b200_demod_soft_demodulate_s",
]
