"""Early-stop decode of 65,536 K=6144 blocks: time and per-kernel-class breakdown over Eb/N0, with and without lane re-packing."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from srslte_b200 import TurboDecoderBatch  # noqa: E402
from srslte_b200.tdec import synth_llr  # noqa: E402

K, ncb = 6144, int(os.environ.get("NCB", "65536"))
dec = TurboDecoderBatch(0, ncb)
out = torch.empty((ncb, K // 8), dtype=torch.uint8, device="cuda")
ok = torch.empty(ncb, dtype=torch.uint8, device="cuda")
npass = torch.empty(ncb, dtype=torch.uint8, device="cuda")
for eb in (1.5, 2.5, 4.0):
    sigma = (3.0 / (2.0 * 10 ** (eb / 10.0))) ** 0.5
    llr, truth = synth_llr(0, ncb, K, sigma=sigma, scale=16.0, clip=31, seed=int(eb * 10))
    for mode in ("repack", "repack_noll", "norepack"):
        os.environ.pop("SRSLTE_B200_TDEC_NO_COMPACT", None)
        os.environ.pop("SRSLTE_B200_TDEC_LL", None)
        if mode == "norepack":
            os.environ["SRSLTE_B200_TDEC_NO_COMPACT"] = "1"
        if mode == "repack_noll":
            os.environ["SRSLTE_B200_TDEC_LL"] = "0"
        dec.decode_device(llr, K, out, ok, npass, 8, "B", True)
        torch.cuda.synchronize()
        dec.profile_reset(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dec.decode_device(llr, K, out, ok, npass, 8, "B", True)
        e1.record()
        torch.cuda.synchronize()
        p = dec.profile_get()
        spans = dec.profile_spans()
        dec.profile_reset(False)
        last = spans[-(len(spans) // 3):]
        print("   last decode: " + " ".join(f"{'LSDR'[c]}{t:.2f}" for c, t in last), flush=True)
        ms = e0.elapsed_time(e1) / 3
        hist = torch.bincount(npass.int(), minlength=9).tolist()
        print(f"Eb/N0 {eb} {mode}: {ms:.2f} ms = {ncb * K / ms / 1e6:.1f} Gbit/s; load {p['load_ms'] / 3:.2f} siso {p['siso_ms'] / 3:.2f} "
              f"repack {p['repack_ms'] / 3:.2f} decide {p['decide_ms'] / 3:.2f}; mean passes {npass.float().mean().item():.2f} hist {hist}", flush=True)
    del llr, truth

# two decoder objects on two streams, steps alternating between them: the light tail passes of one batch overlap the heavy
# first passes of the next
os.environ.pop("SRSLTE_B200_TDEC_NO_COMPACT", None)
dec2 = TurboDecoderBatch(0, ncb)
out2, ok2, np2 = torch.empty_like(out), torch.empty_like(ok), torch.empty_like(npass)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for eb in (1.5, 2.5):
    sigma = (3.0 / (2.0 * 10 ** (eb / 10.0))) ** 0.5
    llr, truth = synth_llr(0, ncb, K, sigma=sigma, scale=16.0, clip=31, seed=int(eb * 10))
    for mode in ("repack", "norepack"):
        if mode == "norepack":
            os.environ["SRSLTE_B200_TDEC_NO_COMPACT"] = "1"
        else:
            os.environ.pop("SRSLTE_B200_TDEC_NO_COMPACT", None)
        torch.cuda.synchronize()
        for rep in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(4):
                dec.decode_device(llr, K, out, ok, npass, 8, "B", True, stream_ptr=s1.cuda_stream)
                dec2.decode_device(llr, K, out2, ok2, np2, 8, "B", True, stream_ptr=s2.cuda_stream)
            torch.cuda.synchronize()
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 8
        print(f"two streams, Eb/N0 {eb} {mode}: {ms:.2f} ms per batch = {ncb * K / ms / 1e6:.1f} Gbit/s", flush=True)
    del llr, truth

# lone-tile latency: 64 / 832 / 9472 blocks (1, 13, 148 tiles), fixed 8 passes, throughput kernel vs low-latency kernel
os.environ.pop("SRSLTE_B200_TDEC_NO_COMPACT", None)
for n in (64, 832, 9472, 18944, 37888):
    llr, truth = synth_llr(0, n, K, sigma=0.8, scale=16.0, clip=31, seed=3)
    o, k, p_ = torch.empty((n, K // 8), dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    res = {}
    for mode in ("0", "1"):
        os.environ["SRSLTE_B200_TDEC_LL"] = mode
        dec.decode_device(llr, K, o, k, p_, 8, "B", False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            dec.decode_device(llr, K, o, k, p_, 8, "B", False)
        e1.record()
        torch.cuda.synchronize()
        res[mode] = (e0.elapsed_time(e1) / 5, o.clone())
    same = bool((res["0"][1] == res["1"][1]).all())
    print(f"{n} blocks ({(n + 63) // 64} tiles), 8 passes: throughput kernel {res['0'][0]:.3f} ms, low-latency kernel {res['1'][0]:.3f} ms, same bytes {same}", flush=True)
os.environ.pop("SRSLTE_B200_TDEC_LL", None)
