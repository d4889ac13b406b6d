#!/usr/bin/env python3
"""Tiny driver for profiler captures: one warm-up decode + one decode of NCB blocks (fixed 8 passes, no early stop)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200 import TurboDecoderBatch  # noqa: E402
from srslte_b200.tdec import synth_llr  # noqa: E402

K = int(os.environ.get("K", 6144))
ncb = int(os.environ.get("NCB", 65536))
llr, truth = synth_llr(0, ncb, K, sigma=0.79, scale=16.0, clip=31, seed=1)
dec = TurboDecoderBatch(0, ncb)
out = torch.empty((ncb, K // 8), dtype=torch.uint8, device="cuda")
ok = torch.empty(ncb, dtype=torch.uint8, device="cuda")
npass = torch.empty(ncb, dtype=torch.uint8, device="cuda")
for _ in range(2):
    dec.decode_device(llr, K, out, ok, npass, 8, "B", False)
    torch.cuda.synchronize()
print("ok", ok.float().mean().item())
