"""One tile of K=6144 through the low-latency SISO kernel with its cycle stamps printed (SRSLTE_B200_TDEC_LL_DEBUG)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SRSLTE_B200_TDEC_LL"] = "1"
os.environ["SRSLTE_B200_TDEC_LL_DEBUG"] = "1"
from srslte_b200 import TurboDecoderBatch  # noqa: E402
from srslte_b200.tdec import synth_llr  # noqa: E402

K, n = 6144, 64
dec = TurboDecoderBatch(0, n)
llr, truth = synth_llr(0, n, K, sigma=0.8, scale=16.0, clip=31, seed=3)
o, k, p = torch.empty((n, K // 8), dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
dec.decode_device(llr, K, o, k, p, 3, "B", False)
torch.cuda.synchronize()
