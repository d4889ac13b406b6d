#!/usr/bin/env python3
"""Developer timing of the device-resident batched decode (not the contract bench): prints per-call ms and Gbit/s."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200 import TurboDecoderBatch  # noqa: E402
from srslte_b200.tdec import synth_llr  # noqa: E402


def main():
    K = int(os.environ.get("K", 6144))
    ncb = int(os.environ.get("NCB", 65536))
    sigma = float(os.environ.get("SIGMA", 0.79))
    t0 = time.time()
    llr, truth = synth_llr(0, ncb, K, sigma=sigma, scale=16.0, clip=31, seed=1)
    torch.cuda.synchronize()
    print(f"synth {ncb} x K={K}: {time.time()-t0:.2f} s")
    dec = TurboDecoderBatch(0, ncb)
    from srslte_b200 import _lib
    print("resident tiles per SM:", _lib.lib().srsran_b200_tdec_resident_tiles_per_sm())
    out = torch.empty((ncb, K // 8), dtype=torch.uint8, device="cuda")
    ok = torch.empty(ncb, dtype=torch.uint8, device="cuda")
    npass = torch.empty(ncb, dtype=torch.uint8, device="cuda")
    for early, mp in ((False, 8), (True, 8), (False, 1), (False, 2)):
        for it in range(3):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            dec.decode_device(llr, K, out, ok, npass, mp, "B", early)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        good = ok.bool()
        ber_ok = (out[good] == truth[good]).all().item()
        print(f"early={early} max_pass={mp}: {ms:.3f} ms  {ncb*K/ms/1e6:.2f} Gbit/s  crc_ok={good.float().mean().item():.4f} "
              f"mean_pass={npass.float().mean().item():.2f} ok_blocks_match_truth={ber_ok}")


if __name__ == "__main__":
    main()
