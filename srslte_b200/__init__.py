"""srslte_b200 — B200-native receive-side PHY hot path of srsLTE/srsRAN behind the reference's C API.

The product is the shared library ``libsrslte_b200.so`` (hand-written sm_100a CUDA + a C ABI, see ``include/``);
this package is only its Python mirror for tests and benchmarks.
"""
from . import _lib  # noqa: F401
from .sch import SchDecoder  # noqa: F401
from .tdec import TurboDecoderBatch  # noqa: F401

__all__ = ["TurboDecoderBatch", "SchDecoder"]
