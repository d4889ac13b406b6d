#!/usr/bin/env python3
"""Benchmark of the B200-native receive-side PHY hot path (driver contract: one JSON line on stdout).

Headline workload (BASELINE.json configs[1]): 65,536 code blocks of K=6144 per GPU, int16 LLRs (scale 16, clip +-31),
synthetic BPSK/AWGN, max-log-MAP turbo decoding, 8 SISO passes per block ("8 iterations" of the reference API = 4 full
turbo iterations, turbodecoder_iter.h:104-140), CRC24B evaluated after every pass.

  value      decoded info Gbit/s (K bits per block, as turbodecoder_test.c:284 counts), LLRs resident in HBM, EXACTLY 8 passes
             for every block (early stop off: the work per step is fixed), decided bits + CRC flags + pass counts copied back
             to pinned host memory INSIDE the timed region (second stream, overlapped with the next step), CUDA-event timed.
  e2e        the same metric through the C ABI with HOST buffers (pinned): H2D of the LLRs and D2H of the results inside the
             timed region, chunk-pipelined by the library; with the achieved host->device GB/s of every rank.
  roofline   SURVEY 8(d): the decoder is bound by the packed-int16 issue rate -- the reference's literal 78.5 int16 operations
             per trellis step and pass (two code blocks per packed lane) against the VIADD.16x2 + VIADDMNMX.S16x2 issue rate
             measured in this process (tools/synth/synth.cu: b200_ubench_int16_issue).  roofline_hbm: the step's compulsory
             bytes (LLRs in, bits and flags out) against the measured copy bandwidth.
  early_stop / early_stop_sweep   the batch with CRC early stop (sch.c:425-454) at 1.5 dB and over Eb/N0, blocks that still run
             re-packed into fewer tiles between passes; `two_in_flight`: two batches on two streams, the light tail passes of
             one overlapping the heavy first passes of the next.
  mixed_k    BASELINE configs[2]: all 188 code block lengths x 256 blocks in ONE batch through rate de-matching + decoding
             (rv 0-3, E = 0.4 / 1 / 1.7 x (3K+12)), one launch per pass over all of them.
  pusch_full BASELINE configs[3]: 20 MHz PUSCH subframes through the complete receiver, 4096 per step.
  multi_cell BASELINE configs[4]: 64 cells x 1000 subframes, cell c on GPU c mod G (STRONG scaling), HARQ buffers resident per
             cell, every subframe with its own noise realisation generated on the device.
  --impl reference : the reference's own CPU decoder (AVX2 16-lane window, srsran_tdec AUTO) from oracle/_ref on all host
             cores, same config, bounded sample per step.

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); code blocks / cells are independent, so ranks take disjoint
shards and only exchange their timings (no collective on the data path).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 6144
NCB_PER_GPU = 65536
MAX_PASSES = 8
SCALE, CLIP = 16.0, 31
EBN0_DB = 1.5  # rate-1/3 BPSK: sigma^2 = 3 / (2 * 10^(EbN0/10)); near the waterfall, so early stop is non-trivial
METRIC = "decoded_info_gbit_per_s_k6144_8pass"
UNIT = "Gbit/s"


def sigma_of(ebn0_db: float) -> float:
    return (3.0 / (2.0 * 10 ** (ebn0_db / 10.0))) ** 0.5


def workload_name() -> str:
    return (f"configs[1]: batched turbo decode, {NCB_PER_GPU} code blocks/GPU K={K}, {MAX_PASSES} SISO passes, int16 LLR "
            f"(scale {SCALE:g}, clip +-{CLIP}), BPSK/AWGN Eb/N0={EBN0_DB} dB, CRC24B checked every pass")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def measured_peaks() -> dict:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "MEASURED_PEAKS.json (measured copy bandwidth)"}
    return {"hbm_gbs": 6650.0, "source": "fallback 6.65 TB/s (B200_PROFILING.md)"}


def helper_lib():
    """tools/synth/libsrslte_b200_synth.so: synthetic inputs and the int16 issue micro-benchmark (bench/test helper, not product)."""
    from srslte_b200.build import SYNTH_LIB_PATH

    L = C.CDLL(SYNTH_LIB_PATH)
    L.b200_synth_pusch_iq16.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float,
                                        C.c_uint64, C.c_void_p, C.c_void_p]
    L.b200_ubench_int16_issue.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    return L


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's CPU implementation of the same path, all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np

    from oracle import loader

    cores = os.cpu_count() or 1
    have_ref = loader.have_ref()
    api = loader.api("ref" if have_ref else "port")
    kind = "reference" if have_ref else "port"
    # bounded sample: per step enough blocks for ~2 s on this box (AVX2 window decoder ~ 120 us / block / core)
    ncb = max(cores * 16, min(NCB_PER_GPU, cores * (1024 if have_ref else 16)))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import coded_llrs

    port = loader.api("port")
    base, _ = coded_llrs(port, K, 64, sigma_of(EBN0_DB), SCALE, CLIP, seed=0xB200)
    llr = np.ascontiguousarray(np.tile(base, ((ncb + 63) // 64, 1))[:ncb])
    impl = loader.TDEC_AUTO if have_ref else loader.TDEC_GENERIC
    times = []
    for it in range(args.warmup + args.steps):
        _, ok, npass, sec = api.decode_batch(llr, K, MAX_PASSES, "B", 0, False, nthreads=cores, impl=impl)
        if it >= args.warmup:
            times.append(sec)
    t = sum(times) / len(times)
    val = ncb * K / t / 1e9
    sample = (f"bounded sample: {ncb} of the {NCB_PER_GPU} code blocks per step (64 distinct blocks of the same distribution, tiled), "
              f"{MAX_PASSES} passes each, no early stop, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": workload_name(), "reference_impl": "srsran_tdec AUTO (AVX2 16-sub-block window, turbodecoder_win.h) "
                   "via srsran_tdec_iteration + srsran_crc_checksum_byte per pass" if have_ref else "oracle port (scalar generic int16)",
                   "threads": cores, "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    from srslte_b200 import TurboDecoderBatch, _lib, shard
    from srslte_b200.tdec import synth_llr

    rank, world, local = shard.rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; srslte_b200 has no CPU fallback")
    # several ranks on one host: the pinned buffers of a rank come from the memory next to its GPU (one rank keeps all cores, which
    # the CPU baseline of the same run needs)
    numa = shard.bind_to_gpu_numa_node(local) if world > 1 else {"bound": False, "note": "single rank: not bound"}
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        # the contract is ONE JSON line on stdout: NCCL's own log lines (NCCL_DEBUG as the launcher set it) go to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return shard.reduce_scalars([x], "max", dist, dev)[0]

    def min_over_ranks(x: float) -> float:
        return -shard.reduce_scalars([-x], "max", dist, dev)[0]

    def sum_over_ranks(x: float) -> float:
        return shard.reduce_scalars([x], "sum", dist, dev)[0]

    R = {"barrier": barrier, "max": max_over_ranks, "min": min_over_ranks, "sum": sum_over_ranks, "rank": rank, "world": world,
         "local": local, "dev": dev, "torch": torch, "np": np}

    ncb = NCB_PER_GPU
    sigma = sigma_of(EBN0_DB)
    lib = _lib.lib()
    hlp = helper_lib()
    dec = TurboDecoderBatch(local, ncb)
    llr, truth = synth_llr(local, ncb, K, sigma=sigma, scale=SCALE, clip=CLIP, seed=shard.shard_seed(0xB200, rank))
    nb = K // 8
    # results: two sets on the device (step s+1 decodes into the other one while step s is copied out) and one pinned host copy
    outs = [torch.empty((ncb, nb), dtype=torch.uint8, device=dev) for _ in range(2)]
    oks = [torch.empty(ncb, dtype=torch.uint8, device=dev) for _ in range(2)]
    nps = [torch.empty(ncb, dtype=torch.uint8, device=dev) for _ in range(2)]
    h_out = torch.empty((ncb, nb), dtype=torch.uint8, pin_memory=True)
    h_ok = torch.empty(ncb, dtype=torch.uint8, pin_memory=True)
    h_np = torch.empty(ncb, dtype=torch.uint8, pin_memory=True)
    out, ok, npass = outs[0], oks[0], nps[0]
    copy_stream = torch.cuda.Stream(dev)
    done = [torch.cuda.Event(), torch.cuda.Event()]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    cur = torch.cuda.current_stream(dev)
    torch.cuda.synchronize()

    # ---- the packed-int16 issue peak of this GPU, measured in this process (roofline denominator) -------------------
    lanes = (C.c_double * 3)()
    ub_mhz = C.c_double(0)
    if hlp.b200_ubench_int16_issue(local, 32, lanes, C.byref(ub_mhz)) != 0:
        lanes[2] = 127.0
    torch.cuda.synchronize()

    # ---- device-resident: exactly 8 passes per block, results copied to the host inside the timed region -------------
    step_no = [0]

    def step_fixed():
        b = step_no[0] & 1
        cur.wait_event(copied[b])                     # the copy that last read this result set has finished
        dec.decode_device(llr, K, outs[b], oks[b], nps[b], MAX_PASSES, "B", False)
        done[b].record(cur)
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(done[b])
            h_out.copy_(outs[b], non_blocking=True)
            h_ok.copy_(oks[b], non_blocking=True)
            h_np.copy_(nps[b], non_blocking=True)
            copied[b].record(copy_stream)
        step_no[0] += 1

    for _ in range(args.warmup):
        step_fixed()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.srsran_b200_kernel_launches()
    dec.profile_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_fixed()
    cur.wait_stream(copy_stream)                      # the last step's results are on the host when the clock stops
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.srsran_b200_kernel_launches() - launches0
    prof = dec.profile_get()  # per-kernel-class CUDA-event time accumulated inside the timed region
    spans = dec.profile_spans()
    dec.profile_reset(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * ncb * K / (ms_step * 1e-3) / 1e9
    frac_ok_fixed = float(h_ok.float().mean().item())
    okb = h_ok.bool()
    ber_ok = bool((h_out[okb] == truth.cpu()[okb]).all().item())
    siso = [t for c, t in spans if c == 1]
    by_kind = {"first": siso[0::8], "dec2": [t for i, t in enumerate(siso) if i % 8 in (1, 3, 5, 7)],
               "dec1": [t for i, t in enumerate(siso) if i % 8 in (2, 4, 6)]}
    pass_ms = {k: (sum(v) / len(v) if v else None) for k, v in by_kind.items()}

    # ---- roofline of the dominant kernel (tdec_siso_pass_kernel) -----------------------------------------------------
    siso_ms = prof["siso_ms"] / max(1, prof["siso_launches"])
    peaks = measured_peaks()
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_hz = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
    # literal operation count of the reference's generic SISO (turbodecoder_gen.c: beta 2 + 12 + 8 + 7/4, alpha 2 + 12 + 16 + 14 +
    # 8 + 1 + 7/4) = 78.5 int16 operations per trellis step and pass; one packed lane-operation carries two code blocks
    lane_ops = ncb * K * 78.5 / 2.0
    alu_peak = float(lanes[2]) * sm_count * sm_hz
    alu_ach = lane_ops / (siso_ms * 1e-3) if siso_ms > 0 else 0.0
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    traffic = None
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("tdec_siso_pass_kernel_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "int16-issue", "kernel": "tdec_siso_pass_kernel", "achieved": alu_ach / 1e12, "peak": alu_peak / 1e12,
                "unit": "T lane-ops/s", "frac": alu_ach / alu_peak if alu_peak else None, "traffic": traffic,
                "algorithmic_lane_ops_per_launch": lane_ops, "avg_launch_ms": siso_ms, "launches_timed": prof["siso_launches"],
                "launch_ms_by_kind": pass_ms, "share_of_step": prof["siso_ms"] / max(1e-9, prof["total_ms"]),
                "peak_source": f"measured in this process: VIADD.16x2 {lanes[0]:.1f}, VIADDMNMX.S16x2 {lanes[1]:.1f}, interleaved 1:1 "
                               f"{lanes[2]:.1f} lanes/clk/SM (32 warps/SM, SM clock {ub_mhz.value:.0f} MHz during the probe) x {sm_count} SMs "
                               f"x the median SM clock of the timed region",
                "note": "78.5 int16 ops per trellis step and pass (the reference's literal count) / 2 code blocks per packed lane"}
    compulsory = ncb * ((3 * K + 12) * 2 + nb + 2)          # LLRs in, decided bits + CRC flag + pass count out
    hbm_ach = compulsory / (ms_step * 1e-3) / 1e9
    roofline_hbm = {"bound": "hbm", "achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                    "compulsory_bytes_per_step": compulsory, "peak_source": peaks["source"],
                    "note": "whole step: compulsory bytes only; the SISO kernel's own DRAM traffic is in roofline.traffic (bytes per launch, ncu)"}

    # ---- the same batch with CRC early stop (sch.c:425-454 loop) -----------------------------------------------------
    dec2 = TurboDecoderBatch(local, ncb)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def time_es(x, reps):
        """(ms per batch single stream, ms per batch with two batches in flight)"""
        dec.decode_device(x, K, out, ok, npass, MAX_PASSES, "B", True)
        barrier()
        e0.record()
        for _ in range(reps):
            dec.decode_device(x, K, out, ok, npass, MAX_PASSES, "B", True)
        e1.record()
        barrier()
        one = max_over_ranks(e0.elapsed_time(e1)) / reps
        for _ in range(2):
            barrier()
            e0.record()
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            for _ in range(reps):
                dec.decode_device(x, K, outs[0], oks[0], nps[0], MAX_PASSES, "B", True, stream_ptr=s1.cuda_stream)
                dec2.decode_device(x, K, outs[1], oks[1], nps[1], MAX_PASSES, "B", True, stream_ptr=s2.cuda_stream)
            cur.wait_stream(s1)
            cur.wait_stream(s2)
            e1.record()
            barrier()
        two = max_over_ranks(e0.elapsed_time(e1)) / (2 * reps)
        return one, two

    n_es = max(3, min(args.steps, 10))
    ms_es, ms_es2 = time_es(llr, n_es)
    mean_pass = sum_over_ranks(float(npass.float().mean().item())) / world
    frac_ok = sum_over_ranks(float(ok.float().mean().item())) / world
    gb = lambda ms: world * ncb * K / (ms * 1e-3) / 1e9
    early_stop = {"value": gb(ms_es), "unit": UNIT, "ms_per_step": ms_es, "mean_passes": mean_pass, "crc_ok_fraction": frac_ok,
                  "ebn0_db": EBN0_DB, "two_in_flight": {"value": gb(ms_es2), "unit": UNIT, "ms_per_batch": ms_es2}}
    es_sweep = []
    for eb in (0.5, 1.0, 2.0, 2.5, 4.0):
        llr_s, truth_s = synth_llr(local, ncb, K, sigma=sigma_of(eb), scale=SCALE, clip=CLIP, seed=shard.shard_seed(0xB200 + int(eb * 10), rank))
        ms_s, ms_s2 = time_es(llr_s, 3)
        okb = ok.bool()
        es_sweep.append({"ebn0_db": eb, "value": gb(ms_s), "unit": UNIT, "ms_per_step": ms_s,
                         "two_in_flight": {"value": gb(ms_s2), "unit": UNIT, "ms_per_batch": ms_s2},
                         "mean_passes": sum_over_ranks(float(npass.float().mean().item())) / world,
                         "crc_ok_fraction": sum_over_ranks(float(okb.float().mean().item())) / world,
                         "crc_ok_blocks_equal_transmitted_bits": bool((out[okb] == truth_s[okb]).all().item())})
        del llr_s, truth_s
    dec2.close()

    # ---- end to end through the C ABI with host (pinned) buffers -----------------------------------------------------
    h_llr = torch.empty((ncb, 3 * K + 12), dtype=torch.int16, pin_memory=True)
    h_llr.copy_(llr)
    torch.cuda.synchronize()
    # plain host->device copy rate of every rank, all ranks at once (what the end-to-end figures are bounded by)
    probe = torch.empty_like(llr)
    barrier()
    t0 = time.perf_counter()
    probe.copy_(h_llr, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = h_llr.numel() * 2 / (time.perf_counter() - t0) / 1e9
    del probe

    def step_e2e():
        dec.decode_pinned(h_llr.data_ptr(), ncb, K, h_out.data_ptr(), h_ok.data_ptr(), h_np.data_ptr(), MAX_PASSES, "B", False)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 8))
    for _ in range(n_e2e):
        step_e2e()
    barrier()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_e2e
    e2e_val = world * ncb * K / (ms_e2e * 1e-3) / 1e9
    e2e_same = bool((h_out == outs[(step_no[0] - 1) & 1].cpu()).all().item()) if False else None
    # the same call with the LLRs in the reference's 8-bit soft-bit container (SRSRAN_B200_FLAG_LLR_INT8): identical values
    # (|LLR| <= 31 here), identical int16 arithmetic after widening on the device, half the PCIe bytes.  Reported beside e2e.
    h_llr8 = torch.empty((ncb, 3 * K + 12), dtype=torch.int8, pin_memory=True)
    h_llr8.copy_(llr.to(torch.int8))
    h_out8 = torch.empty((ncb, nb), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

    def step_e2e8():
        dec.decode_pinned(h_llr8.data_ptr(), ncb, K, h_out8.data_ptr(), h_ok.data_ptr(), h_np.data_ptr(), MAX_PASSES, "B", False,
                          llr_int8=True)

    step_e2e8()
    same8 = bool((h_out8 == h_out).all().item())
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        step_e2e8()
    barrier()
    ms_e2e8 = max_over_ranks((time.perf_counter() - t0) * 1e3) / n_e2e
    del h_llr8, h_llr
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(ncb * (3 * K + 12) * 2), "d2h_bytes_per_step": int(ncb * (nb + 2)),
           "ms_per_step": ms_e2e, "steps": n_e2e,
           "h2d_gbs_per_rank": {"achieved_in_e2e": ncb * (3 * K + 12) * 2 / (ms_e2e * 1e-3) / 1e9,
                                "plain_copy_all_ranks_at_once": {"min": min_over_ranks(h2d_gbs), "max": max_over_ranks(h2d_gbs)}},
           "numa": numa,
           "note": "host pinned LLRs in, bits+crc+passes out, chunk-pipelined; bounded by PCIe (6 B per info bit)"}
    e2e8 = {"value": world * ncb * K / (ms_e2e8 * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e8,
            "h2d_bytes_per_step": int(ncb * (3 * K + 12)), "same_bytes_out_as_int16_call": same8,
            "note": "same LLR values handed over as int8 (the reference's 8-bit soft-bit container), widened on the device"}

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's AVX2 decoder on the host cores -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import loader

            cores = os.cpu_count() or 1
            have_ref = loader.have_ref()
            api = loader.api("ref" if have_ref else "port")
            # bounded sample: ~1.5 s of wall time on all cores (~20-25 core-seconds) for the AVX2 decoder at ~130 us/block
            n = min(ncb, cores * (8192 if have_ref else 24))
            sub = llr[:n].cpu().numpy()
            api.decode_batch(sub[:cores], K, MAX_PASSES, "B", 0, False, nthreads=cores,
                             impl=loader.TDEC_AUTO if have_ref else loader.TDEC_GENERIC)
            sec, reps = 0.0, 0
            while sec * cores < 16.0 and reps < 8:  # about 16-20 core-seconds of CPU work in total
                _, _, _, s_1 = api.decode_batch(sub, K, MAX_PASSES, "B", 0, False, nthreads=cores,
                                                impl=loader.TDEC_AUTO if have_ref else loader.TDEC_GENERIC)
                sec += s_1
                reps += 1
            cpu = {"value": reps * n * K / sec / 1e9, "unit": UNIT, "cores": cores, "kind": "reference" if have_ref else "port",
                   "sample": f"first {n} code blocks of the same batch x {reps} repetitions, {MAX_PASSES} passes each, no early stop, {sec:.1f} s wall on {cores} threads, "
                             + ("srsran_tdec AUTO (AVX2 window decoder)" if have_ref else "scalar generic port")}
        except Exception as ex:  # the baseline is informative; never fail the bench on it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    del llr, truth, outs, oks, nps, out, ok, npass
    dec.close()
    torch.cuda.empty_cache()

    legs = {}
    if not args.no_pusch:
        for name, fn in (("mixed_k", mixed_k_leg), ("pusch_full", pusch_full_leg), ("multi_cell", multi_cell_leg)):
            if name in args.skip:
                continue
            try:
                legs[name] = fn(args, R, hlp, peaks)
            except Exception as ex:  # secondary metrics: report the failure, keep the headline
                import traceback

                legs[name] = {"error": repr(ex), "trace": traceback.format_exc()[-1500:]}
            torch.cuda.empty_cache()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int16",
        "data": "synthetic",
        "config": {"workload": workload_name(), "blocks_per_gpu": ncb, "K": K, "passes": MAX_PASSES, "early_stop": False,
                   "l2": "inputs larger than L2 (2.4 GB LLRs + 3.7 GB decoder state per GPU vs 126 MB L2)",
                   "results": "decided bits, CRC flags and pass counts (50 MB) copied to pinned host memory inside the timed region, overlapped",
                   "sharding": f"{world} x {ncb} independent code blocks, no collective on the data path"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": e2e, "e2e_int8_container": e2e8,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "cpu_baseline": cpu,
        "early_stop": early_stop, "early_stop_sweep": es_sweep,
        "checks": {"crc_ok_fraction_fixed8": frac_ok_fixed, "crc_ok_blocks_equal_transmitted_bits": ber_ok},
        "kernel_ms": prof,
    }
    line.update(legs)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
MIXED_PER_K = 256


def mixed_k_leg(args, R, hlp, peaks):
    """BASELINE configs[2]: all 188 LTE code block lengths in ONE batch through rate de-matching and turbo decoding.  256 blocks
    per length, each a single-code-block transport block (TBS = K - 24, CRC24A) so that it runs through the real decode loop
    (srsran_b200_sch_decode_batch: E split, de-matching into the soft buffers, up to 8 passes with a CRC check after each, TB
    CRC); redundancy versions 0-3 and E = 0.4 / 1.0 / 1.7 x (3K+12) (punctured blocks with rv 0 only: the others cannot decode)."""
    import numpy as np

    torch, dev, rank, world = R["torch"], R["dev"], R["rank"], R["world"]
    from srslte_b200 import _lib
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.lte import CB_SIZES
    from srslte_b200.pusch import TB_DTYPE
    from srslte_b200.sch import SOFTBUFFER_SIZE, SchDecoder

    combos = [(0, 0.4)] + [(rv, 1.0) for rv in range(4)] + [(rv, 1.7) for rv in range(4)]
    rng = np.random.default_rng(0xC3 + rank)
    tb = np.zeros(len(CB_SIZES) * MIXED_PER_K, TB_DTYPE)
    tx_parts, sig_parts, payloads = [], [], []
    e_off = 0
    i = 0
    for Kc in CB_SIZES:
        tbs = Kc - 24
        payload = rng.integers(0, 2, (MIXED_PER_K, tbs)).astype(np.uint8)
        blk = np.concatenate([payload, sp.crc24(payload, sp.CRC24A)], axis=1)
        d = sp.turbo_encode(blk, sp.qpp_interleaver(Kc))
        payloads.append(np.packbits(blk, axis=1))
        for ci, (rv, ef) in enumerate(combos):
            rows = np.arange(ci, MIXED_PER_K, len(combos))
            E = int(ef * (3 * Kc + 12)) // 2 * 2
            tx_parts.append(sp.rate_match(d[rows], E, rv).ravel())
            sig_parts.append(np.full(rows.size * E, 0.45 if ef < 0.9 else 0.8, np.float32))
            for r_ in rows:
                t = tb[i + r_]
                t["tbs"], t["Qm"], t["rv"], t["nof_e_bits"], t["e_offset"] = tbs, 2, rv, E, e_off
                t["soft_offset"] = (i + r_) * SOFTBUFFER_SIZE
                t["new_data"] = 1
                e_off += E
        i += MIXED_PER_K
    ntb = tb.size
    stride = (6120 // 8 + 3 + 768 + 15) // 16 * 16
    tb["data_offset"] = np.arange(ntb, dtype=np.uint64) * stride
    tx = torch.from_numpy(np.concatenate(tx_parts)).to(dev)
    sg = torch.from_numpy(np.concatenate(sig_parts)).to(dev)
    g = torch.Generator(device=dev)
    g.manual_seed(0xC3 + rank)
    y = (2.0 * tx.float() - 1.0) + torch.randn(tx.numel(), device=dev, generator=g) * sg
    e_bits = torch.clamp(torch.round(SCALE * y), -CLIP, CLIP).to(torch.int16)
    del tx, sg, y
    soft = torch.zeros(ntb * SOFTBUFFER_SIZE, dtype=torch.int16, device=dev)
    data = torch.zeros(ntb * stride + 1024, dtype=torch.uint8, device=dev)
    sch = SchDecoder(R["local"], MAX_PASSES)
    lib = _lib.lib()

    def step():
        t = tb.copy()
        rc = lib.srsran_b200_sch_decode_batch(sch._h, e_bits.data_ptr(), e_bits.numel(), soft.data_ptr(), soft.numel(), data.data_ptr(),
                                              data.numel(), t.ctypes.data, ntb, _lib.FLAG_DEVICE_PTRS)
        if rc != 0:
            raise RuntimeError(f"srsran_b200_sch_decode_batch failed ({rc})")
        return t

    launches0 = lib.srsran_b200_kernel_launches()
    res = step()
    launches = lib.srsran_b200_kernel_launches() - launches0
    step()
    steps = max(3, min(args.steps, 10))
    R["barrier"]()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = step()
    torch.cuda.synchronize()
    ms = R["max"]((time.perf_counter() - t0) * 1e3) / steps
    # two batches in flight from one thread (srsran_b200_sch_decode_begin / _finish on two decode objects): the host side of a
    # 48,128-entry list -- comparing it with the cached plan, ~50 launches, reading the verdicts back -- overlaps the other batch's kernels
    two = None
    bits = float(sum(CB_SIZES)) * MIXED_PER_K
    try:
        sch2 = SchDecoder(R["local"], MAX_PASSES)
        soft2, data2 = torch.zeros_like(soft), torch.zeros_like(data)
        sets = [(sch, soft, data, tb.copy()), (sch2, soft2, data2, tb.copy())]

        def begin(k):
            q, so, da, t = sets[k]
            t[:] = tb
            rc = lib.srsran_b200_sch_decode_begin(q._h, e_bits.data_ptr(), e_bits.numel(), so.data_ptr(), so.numel(), da.data_ptr(), da.numel(),
                                                  t.ctypes.data, ntb, _lib.FLAG_DEVICE_PTRS)
            if rc != 0:
                raise RuntimeError(f"srsran_b200_sch_decode_begin failed ({rc})")

        def finish(k):
            if lib.srsran_b200_sch_decode_finish(sets[k][0]._h) != 0:
                raise RuntimeError("srsran_b200_sch_decode_finish failed")

        begin(0)
        for i in range(4 + steps):
            if i == 4:
                R["barrier"]()
                t0 = time.perf_counter()
            begin((i + 1) % 2)
            finish(i % 2)
        ms2 = R["max"]((time.perf_counter() - t0) * 1e3) / steps
        finish((4 + steps) % 2)
        torch.cuda.synchronize()
        same = bool(torch.equal(data, data2)) and bool((sets[0][3]["result"] == sets[1][3]["result"]).all()) and bool((sets[0][3]["result"] == res["result"]).all())
        two = {"value": world * bits / (ms2 * 1e-3) / 1e9, "unit": UNIT, "ms_per_batch": ms2, "same_results_from_both_objects": same}
        sch2.close()
        del soft2, data2
    except Exception as ex:  # noqa: BLE001
        two = {"error": repr(ex)}
    okm = res["result"] == 0
    dh = data.cpu().numpy()
    good = True
    i = 0
    for ki, Kc in enumerate(CB_SIZES):
        nbytes = Kc // 8
        for r_ in range(0, MIXED_PER_K, 37):
            if okm[i + r_]:
                o = int(tb["data_offset"][i + r_])
                good = good and bool((dh[o:o + nbytes] == payloads[ki][r_]).all())
        i += MIXED_PER_K
    bits = float(sum(CB_SIZES)) * MIXED_PER_K
    sch.close()
    return {"metric": "mixed_k_info_gbit_per_s_188_sizes_dematch_decode", "value": world * bits / (ms * 1e-3) / 1e9, "unit": UNIT,
            "ms_per_step": ms, "code_blocks_per_step": ntb, "code_blocks_per_s": world * ntb / (ms * 1e-3),
            "mean_passes": float(res["avg_iterations"].mean()), "tb_ok_fraction": float(okm.mean()),
            "decoded_payloads_equal_transmitted": good, "kernel_launches_per_step": int(launches), "two_in_flight": two,
            "config": f"configs[2]: all 188 code block lengths K=40..6144 x {MIXED_PER_K} blocks per GPU in ONE srsran_b200_sch_decode_batch call: "
                      "rate de-matching (rv 0-3, E = 0.4 / 1.0 / 1.7 x (3K+12)) + turbo decoding with CRC24A early stop, max 8 passes; tiles "
                      "ordered by length, ONE launch per pass over all 752 tiles; info bits counted as K per block"}


# ---------------------------------------------------------------------------------------------------------------------
PUSCH_SF_PER_GPU = 4096   # (cell, subframe) pairs per GPU and step: 53,248 code blocks K=5824 = 832 tiles, one wave
PUSCH_SNR_DB = 23.0


def pusch_e2e(torch, dev, steps, h_iq, x, run_fn, result_dev, h_data, barrier, max_over_ranks):
    """End-to-end PUSCH steps: every step copies its own input from pinned host memory and reads its result back; the copy of
    step s+1 (second stream, second device buffer) overlaps the processing of step s, as a receiver fed by a radio would run."""
    x2 = torch.empty_like(x)
    bufs = (x, x2)
    cs = torch.cuda.Stream(dev)
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    cur = torch.cuda.current_stream(dev)

    def copy_in(b):
        with torch.cuda.stream(cs):
            bufs[b].copy_(h_iq, non_blocking=True)
            ev[b].record(cs)

    def loop(n):
        copy_in(0)
        for s_ in range(n):
            b = s_ & 1
            cur.wait_event(ev[b])
            if s_ + 1 < n:
                copy_in(1 - b)
            run_fn(bufs[b])                                   # returns when the transport blocks are decoded
            h_data.copy_(result_dev, non_blocking=True)
        torch.cuda.synchronize()

    loop(2)
    barrier()
    t0 = time.perf_counter()
    loop(steps)
    return max_over_ranks((time.perf_counter() - t0) * 1e3) / steps


def ofdm_cpu_substitute(iq, cores: int, return_grid: bool = False):
    """the reference's srsran_ofdm_rx_sf needs FFTW, which this image does not have, so its speed cannot be measured;
    this is the SUBSTITUTE the survey asks for instead -- the same arithmetic (half-subcarrier shift over the subframe, the 14
    windows advanced by half a CP, 2048-point transforms, the 1200 occupied bins in fftshift order, window-offset ramp) with
    scipy.fft (pocketfft) on all host cores.  iq: (n, 30720) complex64.  Returns subframes/s."""
    import numpy as np
    import scipy.fft

    n, N, R = iq.shape[0], 2048, 1200
    cp = [160] + [144] * 6
    noff = 72
    starts, rel, t0 = [], np.zeros(15 * N), 0
    for slot in range(2):
        for l in range(7):
            rel[t0:t0 + cp[l] + N] = np.arange(-cp[l], N)   # sample index relative to the end of the symbol's cyclic prefix (ofdm.c:347-355)
            t0 += cp[l]
            starts.append(t0 - noff)
            t0 += N
    shift = np.exp(-2j * np.pi * 0.5 * rel / N).astype(np.complex64)
    ramp = np.exp(2j * np.pi * noff * (np.arange(R) - R // 2) / N).astype(np.complex64)
    idx = (np.array(starts)[:, None] + np.arange(N)[None, :]).reshape(-1)
    t_0 = time.perf_counter()
    x = iq * shift[None, :]
    w = x[:, idx].reshape(n, 14, N)
    X = scipy.fft.fft(w, axis=2, workers=cores)
    out = np.concatenate([X[:, :, N - R // 2:], X[:, :, :R // 2]], axis=2) * ramp[None, None, :]
    sec = time.perf_counter() - t_0
    assert out.shape == (n, 14, R)
    return out if return_grid else n / sec


def pusch_full_leg(args, R, hlp, peaks):
    """BASELINE configs[3] batched: the 20 MHz subframe through the COMPLETE receive chain: transmit side with channel interleaver,
    scrambling, transform precoding and DMRS; per-subframe flat fading + timing offset + AWGN (every one of the 4096 subframes of a
    step with its own noise realisation, generated on the device); receive side OFDM rx -> channel estimation -> MMSE equaliser +
    transform de-precoding -> soft demap + descrambling + UL-SCH de-interleave -> rate de-matching -> turbo decoding with CRC
    early stop."""
    import numpy as np

    torch, dev, rank, world, local = R["torch"], R["dev"], R["rank"], R["world"], R["local"]
    barrier, max_over_ranks, sum_over_ranks = R["barrier"], R["max"], R["sum"]
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import PuschRxFull

    nsf, tbs, nd, cell_id = PUSCH_SF_PER_GPU, 75376, 16, 1 + rank
    rx = PuschRxFull(cell_id, 100, tbs, 3, llr_shift=4, max_noi=MAX_PASSES, device=local, symbol_sz=2048)
    rnti8 = np.arange(nd, dtype=np.uint32) * 97 + 62
    tti8 = np.arange(nd, dtype=np.uint32) * 3 + rank
    clean, payload8, G, amp, sigma_t = sp.make_subframes_full(cell_id, 100, 2048, tbs, 6, 0, sp.qpp_interleaver(5824), nd, rnti8, tti8,
                                                              lambda sf: rx.chain.dmrs(sf, 0), PUSCH_SNR_DB, seed=0x77 + rank, noise=False,
                                                              return_gain=True)
    rnti, tti = np.tile(rnti8, nsf // nd), np.tile(tti8, nsf // nd)
    # int16 I/Q (the radio's wire format), AGC-like scaling to half of full scale; every subframe gets its own noise
    peak = float(np.abs(clean.view(np.float32)).max())
    scale = 16384.0 / peak
    base_d = torch.from_numpy(clean).to(dev)
    amp_d = torch.from_numpy(amp).to(dev)
    x16 = torch.empty((nsf, 15 * 2048, 2), dtype=torch.int16, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    assert hlp.b200_synth_pusch_iq16(local, base_d.data_ptr(), amp_d.data_ptr(), nd, nsf, 15 * 2048, sigma_t, scale, 0x5EED + rank,
                                     x16.data_ptr(), st) == 0
    x = torch.view_as_complex((x16.float() / scale).contiguous())   # the same samples as float I/Q
    nbytes = tbs // 8 + 3
    steps, warm = max(3, min(args.steps, 10)), 2
    ok, its = None, None
    for _ in range(warm):
        ok, its = rx.run(x, nsf, rnti, tti)
    want = np.tile(payload8, (nsf // nd, 1))
    got = rx.data[:nsf, :nbytes].cpu().numpy()
    # with a noise realisation per subframe a few transport blocks fail at this SNR (the reference's receiver loses them too):
    # every block whose CRC passed must carry the transmitted bytes
    good = bool((got[ok] == want[ok]).all())
    tb_ok_fraction, tb_bytes_equal_fraction = float(ok.mean()), float((got == want).all(axis=1).mean())
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        rx.run(x, nsf, rnti, tti)
    torch.cuda.synchronize()
    ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
    hist = np.bincount(np.ceil(its - 1e-6).astype(np.int64), minlength=9).tolist()   # subframes by their mean passes per code block, rounded up
    # Two batches in flight from ONE caller thread (srsran_b200_sch_decode_begin / _finish): two receiver objects on two
    # streams, the next batch is queued before the previous one is waited for.  A step's last passes belong to the handful of
    # blocks that fail their CRC and run all 8 passes on an almost empty GPU; they now overlap the bulk of the other batch.
    two = None
    try:
        rx2 = PuschRxFull(cell_id, 100, tbs, 3, llr_shift=4, max_noi=MAX_PASSES, device=local, symbol_sz=2048)
        objs, streams = [rx, rx2], [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        for k in (0, 1):
            streams[k].wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(streams[k]):
                objs[k].run(x, nsf, rnti, tti)
        same = bool(torch.equal(rx.data[:nsf, :nbytes], rx2.data[:nsf, :nbytes]))
        with torch.cuda.stream(streams[0]):
            objs[0].run_begin(x, nsf, rnti, tti)
        stamps, warm2 = [], 4   # untimed pipelined iterations first: the second decoder workspace is allocated when two batches first overlap
        for i in range(warm2 + steps):
            if i == warm2:
                barrier()
                t0 = time.perf_counter()
            k = (i + 1) % 2
            with torch.cuda.stream(streams[k]):
                objs[k].run_begin(x, nsf, rnti, tti)
            ok2, _ = objs[1 - k].run_finish()
            if i >= warm2:
                stamps.append(time.perf_counter())
        ms2 = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
        objs[(warm2 + steps) % 2].run_finish()
        torch.cuda.synchronize()
        two = {"value": world * nsf / (ms2 * 1e-3), "unit": "subframes/s", "ms_per_step": ms2, "tb_ok_fraction": float(ok2.mean()),
               "same_bytes_from_both_objects": same,
               "ms_per_iteration": [round((b - a) * 1e3, 2) for a, b in zip([t0] + stamps[:-1], stamps)],
               "note": "two receiver objects on two streams driven by one thread through srsran_b200_sch_decode_begin / _finish"}
        rx2.close()
        del rx2, objs
    except Exception as ex:  # noqa: BLE001
        two = {"error": repr(ex)}
    # the front-end kernels alone (CUDA events on the launching stream)
    ch = rx.chain
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    t_ofdm = t_chest = t_eq = t_demod = 0.0
    grid = rx.grid[:nsf]
    ce, meas = ch.chest(grid, tti)          # outputs allocated once; the timed calls below only launch kernels
    d = ch.equalize_deprecode(grid, ce, meas)
    ch.demod_descramble(d, rnti, tti, out=rx.llr)
    torch.cuda.synchronize()
    for _ in range(steps):
        # each stage between its own pair of events, with the device idle before it, so that host-side argument marshalling
        # (the per-subframe rnti/tti arrays are copied to the device inside the calls) is not counted as kernel time
        e[3].record()
        rx.ofdm.rx_sf_device(x, rx.grid, nsf, st)
        e[4].record()
        ch.chest(grid, tti, out=(ce, meas))
        torch.cuda.synchronize()
        t_ofdm += e[3].elapsed_time(e[4]) / steps
        e[0].record()
        ch._lib.srsran_b200_chest_ul_pusch_batch(ch._h, grid.data_ptr(), nsf, None, None, ce.data_ptr(), meas.data_ptr(), 1, st)
        e[1].record()
        ch.equalize_deprecode(grid, ce, meas, out=d)
        e[2].record()
        torch.cuda.synchronize()
        t_chest += e[0].elapsed_time(e[1]) / steps
        t_eq += e[1].elapsed_time(e[2]) / steps
        ch.demod_descramble(d, rnti, tti, out=rx.llr)
        torch.cuda.synchronize()
        e[0].record()
        ch._lib.srsran_b200_pusch_demod_descramble_batch(ch._h, d.data_ptr(), rx.llr.data_ptr(), nsf, None, None, 1, st)
        e[3].record()
        torch.cuda.synchronize()
        t_demod += e[0].elapsed_time(e[3]) / steps
    mean_its = sum_over_ranks(float(its.mean())) / world
    snr_est = float(rx.meas[:nd, 1].log10().mean().item() * 10.0)
    grids8 = rx.grid[:nd].cpu().numpy()
    rx.close()
    del x, rx
    torch.cuda.empty_cache()

    # end to end through the native one-call entry (srsran_b200_enb_ul_pusch_batch): pinned host samples in, transport-block
    # bytes out, nothing but the C ABI in between (chunked copies on a second stream inside the call)
    from srslte_b200.pusch import EnbUl, PUSCH_RES_DTYPE

    enb = EnbUl(cell_id, 100, tbs, 3, llr_shift=4, max_noi=MAX_PASSES, device=local, symbol_sz=2048)
    h_out = torch.empty((nsf, enb.tb_bytes), dtype=torch.uint8).pin_memory()
    res_np = np.zeros(nsf, PUSCH_RES_DTYPE)
    h_iq16 = torch.empty((nsf, 15 * 2048, 2), dtype=torch.int16).pin_memory()
    h_iq16.copy_(x16)
    h_iq = torch.empty((nsf, 15 * 2048, 2), dtype=torch.float32).pin_memory()
    h_iq.copy_(x16.float() / scale)
    del x16
    torch.cuda.synchronize()
    native = {}
    for name, hbuf, fl in (("float_iq", h_iq, 0), ("int16_iq", h_iq16, 8)):
        enb.run_ptr(hbuf.data_ptr(), nsf, rnti, tti, h_out.data_ptr(), res_np, flags=fl)
        okm = res_np["crc_ok"] != 0
        okn = bool((h_out.numpy()[okm] == want[okm]).all()) and float(okm.mean()) > 0.99
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            enb.run_ptr(hbuf.data_ptr(), nsf, rnti, tti, h_out.data_ptr(), res_np, flags=fl)
        msn = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
        native[name] = {"value": world * nsf / (msn * 1e-3), "unit": "subframes/s", "ms_per_step": msn, "tb_ok_fraction": float(okm.mean()),
                        "crc_ok_blocks_equal_transmitted_bytes": okn}
    enb.close()
    iq16_sample = h_iq16[:256].numpy().copy()   # for the CPU substitute of the OFDM stage below
    del h_iq, h_iq16
    # CPU baseline: the reference's own receiver after the OFDM demodulator (FFTW is not available to build its srsran_ofdm)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import loader

            if loader.have_ref():
                Rf = loader.api("ref")
                cores = os.cpu_count() or 1
                n_cpu = 16 * cores
                lk = loader.pusch_link(cell_id=cell_id, rnti=int(rnti8[0]), tti=int(tti8[0]), tbs=tbs, max_iter=MAX_PASSES)
                okc, sec = Rf.pusch_rx_bench(lk, np.ascontiguousarray(np.tile(grids8[:1], (n_cpu, 1, 1))), cores)
                try:
                    sub_iq = np.ascontiguousarray(iq16_sample.astype(np.float32) / np.float32(scale)).view(np.complex64).reshape(256, -1)
                    ofdm_sub = {"value": ofdm_cpu_substitute(sub_iq, cores), "unit": "subframes/s", "cores": cores,
                                "label": "FFTW not available; substitute: numpy + scipy.fft (pocketfft) doing what srsran_ofdm_rx_sf does, 256 subframes"}
                except Exception as ex2:  # noqa: BLE001
                    ofdm_sub = {"error": repr(ex2)}
                cpu = {"value": n_cpu / sec, "unit": "subframes/s", "cores": cores, "kind": "reference", "ofdm_substitute": ofdm_sub,
                       "sample": f"{n_cpu} copies of one subframe's resource grid (demodulated on the GPU: the reference's OFDM needs FFTW, "
                                 f"absent here): srsran_chest_ul_estimate_pusch + srsran_pusch_decode per subframe, one object set per thread, "
                                 f"all crc ok = {bool(okc.all())}"}
        except Exception as ex:  # noqa: BLE001
            cpu = {"error": repr(ex)}
    M = 1200
    ofdm_bytes = nsf * (15 * 2048 * 8 + 14 * 1200 * 8)
    chest_bytes = nsf * (2 * M * 8 * 2 + 2 * M * 8)          # two DMRS symbols + known sequence in, two slot estimates out
    eq_bytes = nsf * (12 * M * 8 + 2 * M * 8 + 12 * M * 8)   # data symbols + estimates in, de-precoded symbols out
    demod_bytes = nsf * 12 * M * (8 + 12 + 6 / 8.0 * 2)      # symbols in, soft bits out, scrambling bits written + read
    gbs = lambda b, t: b / (t * 1e-3) / 1e9
    fr = lambda b, t: gbs(b, t) / peaks["hbm_gbs"]
    return {"metric": "pusch_full_chain_subframes_per_s_20mhz_64qam_tbs75376", "value": world * nsf / (ms * 1e-3), "unit": "subframes/s",
            "ms_per_step": ms, "subframes_per_gpu_per_step": nsf, "info_gbit_per_s": world * nsf * tbs / (ms * 1e-3) / 1e9,
            "mean_passes": mean_its, "passes_histogram_per_tb_mean": hist, "snr_db": PUSCH_SNR_DB, "estimated_snr_db": snr_est,
            "tb_ok_fraction": tb_ok_fraction, "tb_bytes_equal_fraction": tb_bytes_equal_fraction, "crc_ok_blocks_equal_transmitted_bytes": good,
            "two_in_flight": two,
            # end to end = ONE C-ABI call per step (srsran_b200_enb_ul_pusch_batch) with pinned host samples in and host bytes out
            "e2e": dict(native["int16_iq"], h2d_bytes_per_step=int(nsf * 15 * 2048 * 4), d2h_bytes_per_step=int(nsf * (tbs // 8 + 3)),
                        note="srsran_b200_enb_ul_pusch_batch with the samples as int16 I/Q pairs (the radio's wire format), converted in the "
                             "first FFT pass; copies chunked on a second stream inside the call"),
            "e2e_float_iq": dict(native["float_iq"], h2d_bytes_per_step=int(nsf * 15 * 2048 * 8), d2h_bytes_per_step=int(nsf * (tbs // 8 + 3))),
            "front_end": {"ofdm_ms": t_ofdm, "ofdm_frac_of_hbm_peak": fr(ofdm_bytes, t_ofdm), "chest_ms": t_chest,
                          "chest_frac_of_hbm_peak": fr(chest_bytes, t_chest), "equalize_deprecode_ms": t_eq,
                          "equalize_deprecode_frac_of_hbm_peak": fr(eq_bytes, t_eq), "demod_descramble_deinterleave_ms": t_demod,
                          "demod_descramble_deinterleave_frac_of_hbm_peak": fr(demod_bytes, t_demod)},
            "cpu_baseline": cpu,
            "config": "configs[3] batched: 100 PRB, N=2048, normal CP, f=-0.5, window offset 0.5, 64QAM, TBS 75376 -> 13 x K=5824, rv 0, "
                      "DMRS + channel interleaver + scrambling + transform precoding, flat fading with timing offset per subframe "
                      f"({nd} distinct), {nsf} subframes per step each with its own AWGN realisation (generated on the device), soft bits >> 4"}


# ---------------------------------------------------------------------------------------------------------------------
NCELLS, SF_PER_CELL = 64, 1000


def multi_cell_leg(args, R, hlp, peaks):
    """BASELINE configs[4]: 64 cells x 1000 subframes of 20 MHz PUSCH, cell c on GPU c mod G (STRONG scaling: the 64,000 subframes
    are the whole job at every G).  One receiver object per cell (its own cell id: scrambling, DMRS), whose HARQ soft buffers stay
    resident on its GPU; every subframe carries its own noise realisation, generated on the device before the clock starts; the
    cells of a GPU are served by a few host threads so that several cells' batches are in flight at once."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor

    torch, dev, rank, world, local = R["torch"], R["dev"], R["rank"], R["world"], R["local"]
    from srslte_b200 import _lib
    from srslte_b200 import shard
    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import EnbUl, PUSCH_RES_DTYPE, PuschChain

    tbs, nd, nsf = 75376, 8, SF_PER_CELL
    cells = [c for c in range(NCELLS) if shard.cell_to_rank(c, world) == rank]
    objs = []
    st = torch.cuda.current_stream(dev).cuda_stream
    for c in cells:
        cell_id = 1 + c
        chain = PuschChain(cell_id=cell_id, cell_nof_prb=100, L_prb=100, n_prb=0, mod=3, llr_shift=4, device=local)
        rnti8 = np.arange(nd, dtype=np.uint32) * 97 + 62 + c
        tti8 = (np.arange(nd, dtype=np.uint32) * 3 + c) % 10240
        clean, payload8, G, amp, sigma_t = sp.make_subframes_full(cell_id, 100, 2048, tbs, 6, 0, sp.qpp_interleaver(5824), nd, rnti8, tti8,
                                                                  lambda sf: chain.dmrs(sf, 0), PUSCH_SNR_DB, seed=0x77, noise=False,
                                                                  return_gain=True)
        chain.close()
        scale = 16384.0 / float(np.abs(clean.view(np.float32)).max())
        base_d = torch.from_numpy(clean).to(dev)
        amp_d = torch.from_numpy(amp).to(dev)
        x16 = torch.empty((nsf, 15 * 2048, 2), dtype=torch.int16, device=dev)
        assert hlp.b200_synth_pusch_iq16(local, base_d.data_ptr(), amp_d.data_ptr(), nd, nsf, 15 * 2048, sigma_t, scale, 0xCE11 + 977 * c,
                                         x16.data_ptr(), st) == 0
        enb = EnbUl(cell_id, 100, tbs, 3, llr_shift=4, max_noi=MAX_PASSES, device=local, symbol_sz=2048)
        objs.append({"enb": enb, "x16": x16, "rnti": np.tile(rnti8, nsf // nd), "tti": np.tile(tti8, nsf // nd),
                     "want": np.tile(payload8, (nsf // nd, 1)), "data": torch.empty((nsf, enb.tb_bytes), dtype=torch.uint8, device=dev),
                     "res": np.zeros(nsf, PUSCH_RES_DTYPE)})
    torch.cuda.synchronize()
    flags = _lib.FLAG_DEVICE_PTRS | _lib.FLAG_IQ_INT16
    # device-resident samples: 8 cells in flight keep the SMs busiest; PCIe-fed: beyond 6 the threads only queue behind the copy engine
    workers = max(1, min(int(os.environ.get("SRSLTE_B200_BENCH_CELL_THREADS", "8")), len(objs)))
    workers_host = max(1, min(int(os.environ.get("SRSLTE_B200_BENCH_CELL_THREADS_HOST", "6")), len(objs)))

    def serve(o):
        torch.cuda.set_device(local)
        o["enb"].run_ptr(o["x16"].data_ptr(), nsf, o["rnti"], o["tti"], o["data"].data_ptr(), o["res"], flags=flags)

    with ThreadPoolExecutor(max_workers=workers) as pool:
        def step():
            list(pool.map(serve, objs))

        step()
        def check(o, got):
            okm = o["res"]["crc_ok"] != 0
            return bool((got[okm] == o["want"][okm]).all()), float(okm.mean())

        chk = [check(o, o["data"].cpu().numpy()) for o in objs]
        good, ok_frac = all(c[0] for c in chk), float(np.mean([c[1] for c in chk])) if chk else 1.0
        mean_its = float(np.mean([o["res"]["avg_iterations"].mean() for o in objs])) if objs else 0.0
        step()
        steps = max(2, min(args.steps, 5))
        R["barrier"]()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        ms = R["max"]((time.perf_counter() - t0) * 1e3) / steps
        # the same job with the samples in pinned HOST memory (int16 I/Q): the end-to-end figure
        for o in objs:
            o["h16"] = torch.empty((nsf, 15 * 2048, 2), dtype=torch.int16).pin_memory()
            o["h16"].copy_(o["x16"])
            o["hdata"] = torch.empty((nsf, o["enb"].tb_bytes), dtype=torch.uint8).pin_memory()
        torch.cuda.synchronize()

        def serve_host(o):
            torch.cuda.set_device(local)
            o["enb"].run_ptr(o["h16"].data_ptr(), nsf, o["rnti"], o["tti"], o["hdata"].data_ptr(), o["res"], flags=_lib.FLAG_IQ_INT16)

        with ThreadPoolExecutor(max_workers=workers_host) as pool_h:
            list(pool_h.map(serve_host, objs))
            good_h = all(check(o, o["hdata"].numpy())[0] for o in objs)
            R["barrier"]()
            t0 = time.perf_counter()
            for _ in range(steps):
                list(pool_h.map(serve_host, objs))
            ms_h = R["max"]((time.perf_counter() - t0) * 1e3) / steps
    # the same two jobs from ONE thread through srsran_b200_enb_ul_pusch_batch_begin / _finish: a rolling window of cells in flight
    one = None
    try:
        def rolling(window, begin):
            inflight = []
            for o in objs:
                begin(o)
                inflight.append(o)
                if len(inflight) >= window:
                    inflight.pop(0)["enb"].finish()
            for o in inflight:
                o["enb"].finish()

        def begin_dev(o):
            o["enb"].begin_ptr(o["x16"].data_ptr(), nsf, o["rnti"], o["tti"], o["data"].data_ptr(), o["res"], flags=flags)

        def begin_host(o):
            o["enb"].begin_ptr(o["h16"].data_ptr(), nsf, o["rnti"], o["tti"], o["hdata"].data_ptr(), o["res"], flags=_lib.FLAG_IQ_INT16)

        one = {}
        for name, fn, window in (("device", begin_dev, workers), ("e2e", begin_host, workers_host)):
            rolling(window, fn)
            ok1 = all(check(o, (o["data"].cpu().numpy() if name == "device" else o["hdata"].numpy()))[0] for o in objs)
            R["barrier"]()
            t0 = time.perf_counter()
            for _ in range(steps):
                rolling(window, fn)
            torch.cuda.synchronize()
            ms1 = R["max"]((time.perf_counter() - t0) * 1e3) / steps
            one[name] = {"value": NCELLS * SF_PER_CELL / (ms1 * 1e-3), "unit": "subframes/s", "ms_per_step": ms1, "cells_in_flight": window,
                         "crc_ok_blocks_equal_transmitted_bytes": bool(R["min"](1.0 if ok1 else 0.0) > 0.5)}
    except Exception as ex:  # noqa: BLE001
        one = {"error": repr(ex)}
    for o in objs:
        o["enb"].close()
    total = NCELLS * SF_PER_CELL
    return {"metric": "multi_cell_pusch_subframes_per_s_64cells_x_1000sf", "value": total / (ms * 1e-3), "unit": "subframes/s", "scaling": "strong",
            "ms_per_step": ms, "cells_on_this_gpu": len(cells), "subframes_per_step_whole_job": total, "host_threads_per_gpu": workers, "host_threads_per_gpu_e2e": workers_host, "one_thread_begin_finish": one,
            "info_gbit_per_s": total * tbs / (ms * 1e-3) / 1e9, "mean_passes": R["sum"](mean_its) / world,
            "tb_ok_fraction": R["sum"](ok_frac) / world, "crc_ok_blocks_equal_transmitted_bytes": bool(R["min"](1.0 if good else 0.0) > 0.5),
            "e2e": {"value": total / (ms_h * 1e-3), "unit": "subframes/s", "ms_per_step": ms_h, "h2d_bytes_per_step": int(total * 15 * 2048 * 4),
                    "d2h_bytes_per_step": int(total * (tbs // 8 + 3)), "crc_ok_blocks_equal_transmitted_bytes": bool(R["min"](1.0 if good_h else 0.0) > 0.5),
                    "note": "int16 I/Q samples in pinned host memory in, transport-block bytes in pinned host memory out, one "
                            "srsran_b200_enb_ul_pusch_batch call per cell and step"},
            "config": f"configs[4]: {NCELLS} cells x {SF_PER_CELL} subframes (100 PRB, 64QAM, TBS 75376, complete PUSCH chain), cell c -> GPU c mod G, one "
                      "srsran_b200_enb_ul object per cell (HARQ soft buffers resident, first transmissions), 64,000 distinct noise realisations "
                      f"generated on the device ({nd} distinct payload/fading sets per cell), int16 I/Q samples resident in HBM"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pusch", action="store_true", help="skip the secondary legs (mixed_k, pusch_full, multi_cell)")
    ap.add_argument("--skip", default="", help="comma-separated secondary legs to skip: mixed_k,pusch_full,multi_cell")
    args = ap.parse_args()
    args.skip = set(x for x in args.skip.split(",") if x)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
