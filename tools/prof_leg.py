"""Runs ONE secondary leg of bench.py on its own (for ncu launch lists and timing breakdowns):  python tools/prof_leg.py mixed_k [steps]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import numpy as np
    import torch

    leg = sys.argv[1]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    args = argparse.Namespace(steps=steps, warmup=3, no_cpu_baseline=True, no_pusch=False, skip=set(), gpus=1)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ident = lambda x: x
    R = {"barrier": torch.cuda.synchronize, "max": ident, "min": ident, "sum": ident, "rank": 0, "world": 1, "local": 0, "dev": dev,
         "torch": torch, "np": np}
    out = getattr(bench, leg + "_leg")(args, R, bench.helper_lib(), bench.measured_peaks())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
