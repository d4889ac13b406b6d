"""How do per-cell receiver calls overlap when several host threads drive one GPU?  N cells x 1000 subframes, host (pinned int16 I/Q)
and device inputs, 1..8 threads."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from srslte_b200 import _lib  # noqa: E402
from srslte_b200 import synth_pusch as sp  # noqa: E402
from srslte_b200.pusch import EnbUl, PUSCH_RES_DTYPE, PuschChain  # noqa: E402

ncell, nsf, tbs, nd = int(os.environ.get("NCELL", "12")), 1000, 75376, 8
hlp = bench.helper_lib()
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream(dev).cuda_stream
objs = []
for c in range(ncell):
    cell_id = 1 + c
    chain = PuschChain(cell_id=cell_id, cell_nof_prb=100, L_prb=100, n_prb=0, mod=3, llr_shift=4, device=0)
    rnti8 = np.arange(nd, dtype=np.uint32) * 97 + 62 + c
    tti8 = (np.arange(nd, dtype=np.uint32) * 3 + c) % 10240
    clean, payload8, G, amp, sigma_t = sp.make_subframes_full(cell_id, 100, 2048, tbs, 6, 0, sp.qpp_interleaver(5824), nd, rnti8, tti8,
                                                              lambda sf: chain.dmrs(sf, 0), 23.0, seed=0x77, noise=False, return_gain=True)
    chain.close()
    scale = 16384.0 / float(np.abs(clean.view(np.float32)).max())
    base_d, amp_d = torch.from_numpy(clean).to(dev), torch.from_numpy(amp).to(dev)
    x16 = torch.empty((nsf, 15 * 2048, 2), dtype=torch.int16, device=dev)
    assert hlp.b200_synth_pusch_iq16(0, base_d.data_ptr(), amp_d.data_ptr(), nd, nsf, 15 * 2048, sigma_t, scale, 0xCE11 + c, x16.data_ptr(), st) == 0
    enb = EnbUl(cell_id, 100, tbs, 3, llr_shift=4, max_noi=8, device=0, symbol_sz=2048)
    h16 = torch.empty((nsf, 15 * 2048, 2), dtype=torch.int16).pin_memory()
    h16.copy_(x16)
    objs.append({"enb": enb, "x16": x16, "h16": h16, "rnti": np.tile(rnti8, nsf // nd), "tti": np.tile(tti8, nsf // nd),
                 "data": torch.empty((nsf, enb.tb_bytes), dtype=torch.uint8, device=dev),
                 "hdata": torch.empty((nsf, enb.tb_bytes), dtype=torch.uint8).pin_memory(), "res": np.zeros(nsf, PUSCH_RES_DTYPE)})
torch.cuda.synchronize()


def serve_dev(o):
    o["enb"].run_ptr(o["x16"].data_ptr(), nsf, o["rnti"], o["tti"], o["data"].data_ptr(), o["res"], flags=_lib.FLAG_DEVICE_PTRS | _lib.FLAG_IQ_INT16)


def serve_host(o):
    o["enb"].run_ptr(o["h16"].data_ptr(), nsf, o["rnti"], o["tti"], o["hdata"].data_ptr(), o["res"], flags=_lib.FLAG_IQ_INT16)


for name, fn in (("device", serve_dev), ("host", serve_host)):
    for nt in (1, 2, 4, 8):
        with ThreadPoolExecutor(max_workers=nt) as pool:
            list(pool.map(fn, objs))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                list(pool.map(fn, objs))
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3 / 3
        print(f"{name} inputs, {nt} threads: {ms:.1f} ms for {ncell} cells x {nsf} subframes = {ms / ncell:.2f} ms per cell, {ncell * nsf / ms:.0f} k subframes/s", flush=True)
