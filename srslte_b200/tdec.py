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
        ms = (C.c_double * 4)()
        n = (C.c_uint64 * 4)()
        rc = self._lib.srsran_b200_tdec_profile_get_ex(self._h, ms, n, 4)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_profile_get_ex failed ({rc})")
        return {"load_ms": ms[0], "siso_ms": ms[1], "decide_ms": ms[2], "repack_ms": ms[3], "total_ms": ms[0] + ms[1] + ms[2] + ms[3],
                "load_launches": int(n[0]), "siso_launches": int(n[1]), "decide_launches": int(n[2]), "repack_launches": int(n[3])}

    def profile_spans(self, max_spans: int = 8192):
        """[(class, ms), ...] of every timed span since the last reset, in launch order (0 load, 1 SISO pass, 2 decide, 3 re-packing)."""
        ms = (C.c_float * max_spans)()
        cls = (C.c_int * max_spans)()
        n = self._lib.srsran_b200_tdec_profile_spans(self._h, ms, cls, max_spans)
        if n < 0:
            raise RuntimeError(f"srsran_b200_tdec_profile_spans failed ({n})")
        return [(int(cls[i]), float(ms[i])) for i in range(n)]

    # -- several code block lengths in one batch (BASELINE config 3) ---------------------------------------------
    def decode_mixed(self, llrs, Ks, max_passes: int = 8, crc: str | None = "B", early_stop: bool = True):
        """llrs: list of (ncb_g, 3K_g+12) int16 numpy arrays, Ks: their code block lengths.  Host buffers.
        Returns a list of (bytes (ncb_g,K_g/8), crc_ok (ncb_g,), npass (ncb_g,)) per group."""
        ncbs = [int(np.asarray(a).reshape(-1, 3 * K + 12).shape[0]) for a, K in zip(llrs, Ks)]
        flat = np.ascontiguousarray(np.concatenate([np.asarray(a, np.int16).ravel() for a in llrs]) if llrs else np.zeros(0, np.int16))
        tot = sum(ncbs)
        out = np.zeros(sum(n * (K // 8) for n, K in zip(ncbs, Ks)), np.uint8)
        ok = np.zeros(tot, np.uint8)
        npass = np.zeros(tot, np.uint8)
        Ka = np.array(Ks, np.uint32)
        na = np.array(ncbs, np.uint32)
        rc = self._lib.srsran_b200_tdec_run_mixed(self._h, flat.ctypes.data, len(Ks), Ka.ctypes.data, na.ctypes.data, max_passes,
                                                  _CRC[crc], int(early_stop), out.ctypes.data, ok.ctypes.data, npass.ctypes.data, 0, None)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_run_mixed failed ({rc})")
        res, o0, c0 = [], 0, 0
        for n, K in zip(ncbs, Ks):
            res.append((out[o0:o0 + n * (K // 8)].reshape(n, K // 8), ok[c0:c0 + n], npass[c0:c0 + n]))
            o0 += n * (K // 8)
            c0 += n
        return res

    def decode_mixed_device(self, llr, Ks, ncbs, out, ok, npass, max_passes: int = 8, crc: str | None = "B",
                            early_stop: bool = True, stream_ptr: int | None = None):
        """llr / out / ok / npass: flat torch CUDA tensors laid out group after group; asynchronous."""
        import torch

        if stream_ptr is None:
            stream_ptr = torch.cuda.current_stream(llr.device).cuda_stream
        Ka = np.array(Ks, np.uint32)
        na = np.array(ncbs, np.uint32)
        rc = self._lib.srsran_b200_tdec_run_mixed(self._h, llr.data_ptr(), len(Ks), Ka.ctypes.data, na.ctypes.data, max_passes,
                                                  _CRC[crc], int(early_stop), out.data_ptr(), ok.data_ptr() if ok is not None else None,
                                                  npass.data_ptr() if npass is not None else None, _lib.FLAG_DEVICE_PTRS, stream_ptr)
        if rc != _lib.SUCCESS:
            raise RuntimeError(f"srsran_b200_tdec_run_mixed failed ({rc})")

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
    """Synthetic AWGN workload generated on the GPU by the test/bench helper library tools/synth/libsrslte_b200_synth.so
    (not part of the product library): returns (llr int16 (ncb,3K+12), truth uint8 (ncb,K/8)) torch CUDA tensors."""
    import ctypes as C

    import torch

    from .build import SYNTH_LIB_PATH

    L = C.CDLL(SYNTH_LIB_PATH)
    L.b200_synth_llr.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_int,
                                 C.c_uint64, C.c_int, C.c_void_p]
    dev = torch.device("cuda", device)
    llr = torch.empty((ncb, 3 * K + 12), dtype=torch.int16, device=dev)
    truth = torch.empty((ncb, K // 8), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    rc = L.b200_synth_llr(device, llr.data_ptr(), truth.data_ptr(), ncb, K, sigma, scale, clip, seed, int(attach_crc), st)
    if rc != 0:
        raise RuntimeError(f"b200_synth_llr failed ({rc})")
    return llr, truth
