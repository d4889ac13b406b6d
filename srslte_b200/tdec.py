"""Python mirror of the batched turbo decoder entry (include/srslte_b200.h, srsran_b200_tdec_*).

Accepts numpy arrays (host path: the library copies host<->device itself) or torch CUDA tensors (device path:
pointers are passed through, the work is enqueued on torch's current stream and nothing synchronises).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

_CRC = {None: _lib.CRC_NONE, "none": _lib.CRC_NONE, "A": _lib.CRC24A, "B": _lib.CRC24B}


class TurboDecoderBatch:
    """One decoder object per host thread and device, like ``srsran_tdec_t`` (turbodecoder.h:63-95)."""

    def __init__(self, device: int = 0, max_cb_hint: int = 0):
        self._lib = _lib.lib()
        self._h = C.c_void_p()
        self.device = device
        rc = self._lib.srsran_b200_tdec_init(C.byref(self._h), device, max_cb_hint)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_init failed ({rc}): no usable CUDA device {device}?")

    def close(self):
        if self._h:
            self._lib.srsran_b200_tdec_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- kernel timing (CUDA events inside the library, on the launching stream) ----------------------------
    def profile_reset(self, enable: bool = True):
        self._lib.srsran_b200_tdec_profile_reset(self._h, int(enable))

    def profile_get(self) -> dict:
        ms = (C.c_double * 3)()
        n = (C.c_uint64 * 3)()
        rc = self._lib.srsran_b200_tdec_profile_get(self._h, ms, n)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_profile_get failed ({rc})")
        return {"load_ms": ms[0], "siso_ms": ms[1], "decide_ms": ms[2], "total_ms": ms[0] + ms[1] + ms[2],
                "load_launches": int(n[0]), "siso_launches": int(n[1]), "decide_launches": int(n[2])}

    # -- host buffers ------------------------------------------------------------------------------------
    def decode(self, llr: np.ndarray, K: int, max_passes: int = 8, crc: str | None = "B", early_stop: bool = True):
        """llr: (ncb, 3K+12) int16 numpy.  Returns (bytes (ncb,K/8) uint8, crc_ok (ncb,), npass (ncb,))."""
        llr = np.ascontiguousarray(llr, dtype=np.int16).reshape(-1, 3 * K + 12)
        ncb = llr.shape[0]
        out = np.zeros((ncb, K // 8), np.uint8)
        ok = np.zeros(ncb, np.uint8)
        npass = np.zeros(ncb, np.uint8)
        rc = self._lib.srsran_b200_tdec_run(self._h, llr.ctypes.data, ncb, K, max_passes, _CRC[crc], int(early_stop),
                                            out.ctypes.data, ok.ctypes.data, npass.ctypes.data, 0, None)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_run failed ({rc})")
        return out, ok, npass

    def decode_pinned(self, llr_ptr: int, ncb: int, K: int, out_ptr: int, ok_ptr: int, npass_ptr: int,
                      max_passes: int = 8, crc: str | None = "B", early_stop: bool = True, llr_int8: bool = False):
        """Host path on raw (pinned) host pointers, for the end-to-end benchmark.  llr_int8: the buffer holds int8 values
        (SRSRAN_B200_FLAG_LLR_INT8), decoded with the same int16 arithmetic after widening on the device."""
        rc = self._lib.srsran_b200_tdec_run(self._h, llr_ptr, ncb, K, max_passes, _CRC[crc], int(early_stop),
                                            out_ptr, ok_ptr, npass_ptr, _lib.FLAG_LLR_INT8 if llr_int8 else 0, None)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_run failed ({rc})")

    # -- device buffers ----------------------------------------------------------------------------------
    def decode_device(self, llr, K: int, out, ok, npass, max_passes: int = 8, crc: str | None = "B",
                      early_stop: bool = True, stream_ptr: int | None = None):
        """llr/out/ok/npass: torch CUDA tensors (int16 / uint8) on this object's device; asynchronous."""
        import torch

        ncb = llr.numel() // (3 * K + 12)
        if stream_ptr is None:
            stream_ptr = torch.cuda.current_stream(llr.device).cuda_stream
        rc = self._lib.srsran_b200_tdec_run(self._h, llr.data_ptr(), ncb, K, max_passes, _CRC[crc], int(early_stop),
                                            out.data_ptr(), ok.data_ptr() if ok is not None else None,
                                            npass.data_ptr() if npass is not None else None,
                                            _lib.FLAG_DEVICE_PTRS | (_lib.FLAG_LLR_INT8 if llr.dtype == torch.int8 else 0), stream_ptr)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_run failed ({rc})")


def synth_llr(device: int, ncb: int, K: int, sigma: float, scale: float = 16.0, clip: int = 31, seed: int = 0xB200,
              attach_crc: bool = True):
    """Synthetic AWGN workload generated on the GPU: returns (llr int16 (ncb,3K+12), truth uint8 (ncb,K/8)) torch CUDA."""
    import torch

    dev = torch.device("cuda", device)
    llr = torch.empty((ncb, 3 * K + 12), dtype=torch.int16, device=dev)
    truth = torch.empty((ncb, K // 8), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    rc = _lib.lib().srsran_b200_synth_llr(device, llr.data_ptr(), truth.data_ptr(), ncb, K, sigma, scale, clip, seed,
                                          int(attach_crc), st)
    if rc != _lib.SUCCESS:
        raise RuntimeError(f"srsran_b200_synth_llr failed ({rc})")
    return llr, truth
