#!/usr/bin/env python3
"""Benchmark of the B200-native receive-side PHY hot path (driver contract: one JSON line on stdout).

Workload (BASELINE.json configs[1]): 65,536 code blocks of K=6144, int16 LLRs (scale 16, clip +-31), synthetic
BPSK/AWGN, max-log-MAP turbo decoding, 8 SISO passes per block ("8 iterations" of the reference API = 4 full turbo
iterations, turbodecoder_iter.h:104-140), CRC24B evaluated after every pass.

  value : decoded info Gbit/s (K bits per block, as turbodecoder_test.c:284 counts) with the LLRs already resident in
          HBM, EXACTLY 8 passes for every block (early stop off, so the work per step is fixed), CUDA-event timed.
          Extra keys report the same batch with CRC early stop on (the reference's decode_tb_cb loop, sch.c:425-454).
  e2e   : the same metric through the C ABI with HOST buffers (pinned): H2D of the LLRs and D2H of bits/flags inside
          the timed region, chunk-pipelined by the library.
  e2e_int8_container : e2e with the same LLR values in the 8-bit soft-bit container (half the PCIe bytes)
  roofline / roofline_int16_alu : SISO launch time against the measured HBM copy peak (algorithmic bytes) and against the
          measured packed-int16 issue peak (the reference's literal operation count)
  early_stop / early_stop_sweep : the batch (and fresh batches at 0.5 / 1 / 2.5 / 4 dB) with CRC early stop on
  pusch      : BASELINE's second metric, 20 MHz PUSCH subframes/s through OFDM rx -> demap -> de-match -> decode (identity channel)
  pusch_full : the same subframe through the complete receiver (channel estimation, equaliser, transform de-precoding,
          descrambling, de-interleave); its e2e is ONE native call per step (srsran_b200_enb_ul_pusch_batch) with host samples
  --impl reference : the reference's own CPU decoder (AVX2 16-lane window, srsran_tdec AUTO) from oracle/_ref on all
          host cores, same config, bounded sample per step.

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); code blocks are independent, so ranks shard the batch
(weak scaling: 65,536 blocks per GPU) and only exchange their timings.
"""
from __future__ import annotations

import argparse
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
    sample = f"{ncb} of {NCB_PER_GPU} code blocks per step (64 distinct, tiled), {MAX_PASSES} passes each, no early stop"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16", "data": "synthetic",
        "config": {"workload": workload_name(), "reference_impl": "srsran_tdec AUTO (AVX2 16-sub-block window, turbodecoder_win.h) "
                   "via srsran_tdec_iteration + srsran_crc_checksum_byte per pass" if have_ref else "oracle port (scalar generic int16)",
                   "threads": cores},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    from srslte_b200 import TurboDecoderBatch, _lib
    from srslte_b200.tdec import synth_llr

    from srslte_b200 import shard

    rank, world, local = shard.rank_info()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; srslte_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: the contract is ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        return shard.reduce_scalars([x], "max", dist, dev)[0]

    def sum_over_ranks(x: float) -> float:
        return shard.reduce_scalars([x], "sum", dist, dev)[0]

    ncb = NCB_PER_GPU
    sigma = sigma_of(EBN0_DB)
    lib = _lib.lib()
    dec = TurboDecoderBatch(local, ncb)
    llr, truth = synth_llr(local, ncb, K, sigma=sigma, scale=SCALE, clip=CLIP, seed=shard.shard_seed(0xB200, rank))
    out = torch.empty((ncb, K // 8), dtype=torch.uint8, device=dev)
    ok = torch.empty(ncb, dtype=torch.uint8, device=dev)
    npass = torch.empty(ncb, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    # ---- device-resident: exactly 8 passes per block -----------------------------------------------------------
    def step_fixed():
        dec.decode_device(llr, K, out, ok, npass, MAX_PASSES, "B", False)

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
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = lib.srsran_b200_kernel_launches() - launches0
    prof = dec.profile_get()  # per-kernel-class CUDA-event time accumulated inside the timed region
    dec.profile_reset(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * ncb * K / (ms_step * 1e-3) / 1e9
    frac_ok_fixed = float(ok.float().mean().item())
    ber_ok = bool((out[ok.bool()] == truth[ok.bool()]).all().item())

    # ---- the same batch with CRC early stop (sch.c:425-454 loop) -----------------------------------------------
    def step_es():
        dec.decode_device(llr, K, out, ok, npass, MAX_PASSES, "B", True)

    for _ in range(max(1, args.warmup // 2)):
        step_es()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_es()
    e1.record()
    barrier()
    ms_es = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    mean_pass = sum_over_ranks(float(npass.float().mean().item())) / world
    frac_ok = sum_over_ranks(float(ok.float().mean().item())) / world
    # ... and at the other operating points SURVEY 8d lists for config 2 (fresh batches, same quantisation)
    es_sweep = []
    for eb in (0.5, 1.0, 2.5, 4.0):
        llr_s, truth_s = synth_llr(local, ncb, K, sigma=sigma_of(eb), scale=SCALE, clip=CLIP, seed=shard.shard_seed(0xB200 + int(eb * 10), rank))
        dec.decode_device(llr_s, K, out, ok, npass, MAX_PASSES, "B", True)
        barrier()
        e0.record()
        for _ in range(3):
            dec.decode_device(llr_s, K, out, ok, npass, MAX_PASSES, "B", True)
        e1.record()
        barrier()
        ms_s = max_over_ranks(e0.elapsed_time(e1)) / 3
        okb = ok.bool()
        es_sweep.append({"ebn0_db": eb, "value": world * ncb * K / (ms_s * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_s,
                         "mean_passes": sum_over_ranks(float(npass.float().mean().item())) / world,
                         "crc_ok_fraction": sum_over_ranks(float(okb.float().mean().item())) / world,
                         "crc_ok_blocks_equal_transmitted_bits": bool((out[okb] == truth_s[okb]).all().item())})
        del llr_s, truth_s

    # ---- end to end through the C ABI with host (pinned) buffers -----------------------------------------------
    h_llr = torch.empty((ncb, 3 * K + 12), dtype=torch.int16, pin_memory=True)
    h_llr.copy_(llr)
    h_out = torch.empty((ncb, K // 8), dtype=torch.uint8, pin_memory=True)
    h_ok = torch.empty(ncb, dtype=torch.uint8, pin_memory=True)
    h_np = torch.empty(ncb, dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()

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
    # the same call with the LLRs in the reference's 8-bit soft-bit container (SRSRAN_B200_FLAG_LLR_INT8): identical values
    # (|LLR| <= 31 here), identical int16 arithmetic after widening on the device, half the PCIe bytes.  Reported beside e2e.
    h_llr8 = torch.empty((ncb, 3 * K + 12), dtype=torch.int8, pin_memory=True)
    h_llr8.copy_(llr.to(torch.int8))
    h_out8 = torch.empty((ncb, K // 8), dtype=torch.uint8, pin_memory=True)
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
    del h_llr8
    e2e_match = bool((h_out.to(dev) == out).all().item()) if False else None  # (out now holds the early-stop decode)

    # ---- roofline of the dominant kernel (tdec_siso_pass_kernel) -----------------------------------------------
    # Algorithmic bytes per SISO launch and code block: one read of the pass's inputs and one
    # write of the extrinsics, int16:  DEC1 (even pass): S, E, P0 in + E out = 8K bytes (6K on pass 0, no a-priori yet);
    # DEC2 (odd pass): E, P1 in + E out = 6K bytes.  Averaged over the 8 launches of a step.  (The kernel's actual DRAM
    # traffic -- "traffic", from ncu -- is higher by construction: checkpoint + recompute reads the inputs twice.)
    alg_per_cb = (6 * K + 4 * 6 * K + 3 * 8 * K) / 8.0
    siso_ms = prof["siso_ms"] / max(1, prof["siso_launches"])
    peaks = measured_peaks()
    achieved = ncb * alg_per_cb / (siso_ms * 1e-3) / 1e9 if siso_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "tdec_siso_pass_kernel", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": None, "peak_source": peaks["source"],
                "avg_launch_ms": siso_ms, "launches_timed": prof["siso_launches"],
                "algorithmic_bytes_per_launch": ncb * alg_per_cb,
                "share_of_step": prof["siso_ms"] / max(1e-9, prof["total_ms"])}
    # Second ceiling: the decoder's integer work against the packed-int16 issue peak.  Literal operation count of the
    # reference's generic SISO (turbodecoder_gen.c: beta 2+12+8+7/4, alpha 2+12+16+14+8+1+7/4) = 78.5 int16 ops per trellis step
    # and pass; one packed instruction lane carries two code blocks.  Peak = 127 lanes/clk/SM measured for VIADD.16x2 +
    # VIADDMNMX.S16x2 issued together (tools/ubench_int16.cu, profiles/r01_int16_issue.txt) x SMs x the SM clock under load.
    sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_hz = (clocks or {}).get("sm_mhz") or 1965.0
    alu_peak = 127.0 * sm_count * sm_hz * 1e6
    alu_ach = ncb * K * 78.5 / 2.0 / (siso_ms * 1e-3) if siso_ms > 0 else 0.0
    roofline_alu = {"bound": "packed-int16 issue", "kernel": "tdec_siso_pass_kernel", "achieved": alu_ach / 1e12, "peak": alu_peak / 1e12,
                    "unit": "T lane-ops/s", "frac": alu_ach / alu_peak,
                    "note": "78.5 int16 ops per trellis step and pass (reference's literal count) / 2 blocks per lane; peak 127 lanes/clk/SM measured"}
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get("tdec_siso_pass_kernel_bytes_per_launch")
        except Exception:
            pass

    # ---- CPU baseline beside it (rank 0, N=1 only): the reference's AVX2 decoder on the host cores ------------------
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
                _, _, _, s1 = api.decode_batch(sub, K, MAX_PASSES, "B", 0, False, nthreads=cores,
                                               impl=loader.TDEC_AUTO if have_ref else loader.TDEC_GENERIC)
                sec += s1
                reps += 1
            cpu = {"value": reps * n * K / sec / 1e9, "unit": UNIT, "cores": cores, "kind": "reference" if have_ref else "port",
                   "sample": f"first {n} code blocks of the same batch x {reps} repetitions, {MAX_PASSES} passes each, no early stop, {sec:.1f} s wall on {cores} threads, "
                             + ("srsran_tdec AUTO (AVX2 window decoder)" if have_ref else "scalar generic port")}
        except Exception as ex:  # the baseline is informative; never fail the bench on it
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex}"}

    pusch = None
    if not args.no_pusch:
        try:
            pusch = pusch_leg(args, torch, dev, local, rank, world, barrier, max_over_ranks, sum_over_ranks, peaks)
        except Exception as ex:  # secondary metric: report the failure, keep the headline
            pusch = {"error": repr(ex)}
    pusch_full = None
    if not args.no_pusch:
        try:
            pusch_full = pusch_full_leg(args, torch, dev, local, rank, world, barrier, max_over_ranks, sum_over_ranks, peaks)
        except Exception as ex:  # noqa: BLE001
            pusch_full = {"error": repr(ex)}

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
                   "sharding": f"{world} x {ncb} independent code blocks, no collective on the data path"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(ncb * (3 * K + 12) * 2),
                "d2h_bytes_per_step": int(ncb * (K // 8 + 2)), "ms_per_step": ms_e2e, "steps": n_e2e,
                "note": "host pinned LLRs in, bits+crc+passes out, chunk-pipelined; bounded by PCIe (6 B per info bit)"},
        "e2e_int8_container": {"value": world * ncb * K / (ms_e2e8 * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_e2e8,
                               "h2d_bytes_per_step": int(ncb * (3 * K + 12)), "same_bytes_out_as_int16_call": same8,
                               "note": "same LLR values handed over as int8 (the reference's 8-bit soft-bit container), widened on the device"},
        "roofline": roofline, "roofline_int16_alu": roofline_alu, "cpu_baseline": cpu,
        "early_stop": {"value": world * ncb * K / (ms_es * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms_es, "mean_passes": mean_pass,
                       "crc_ok_fraction": frac_ok, "ebn0_db": EBN0_DB},
        "early_stop_sweep": es_sweep,
        "checks": {"crc_ok_fraction_fixed8": frac_ok_fixed, "crc_ok_blocks_equal_transmitted_bits": ber_ok},
        "kernel_ms": prof,
        "pusch": pusch,
        "pusch_full": pusch_full,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
PUSCH_SF_PER_GPU = 4096   # (cell, subframe) pairs per GPU and step: 53,248 code blocks K=5824 = 832 tiles, one wave
PUSCH_SNR_DB = 23.0


def pusch_leg(args, torch, dev, local, rank, world, barrier, max_over_ranks, sum_over_ranks, peaks):
    """BASELINE's second metric: PUSCH subframes/s for the config-4 pipeline (20 MHz, 100 PRB, 2048-point OFDM rx of 14
    symbols, 64QAM soft demap, rate de-matching, turbo decoding of the 13 code blocks of TBS 75,376 with CRC early stop),
    batched as in config 5 (cells x subframes sharded across the GPUs, no collective)."""
    import numpy as np

    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import PuschRx

    nsf, tbs, nd = PUSCH_SF_PER_GPU, 75376, 8
    iq8, payload8, G = sp.make_subframes(100, 2048, tbs, 6, 0, sp.qpp_interleaver(5824), nd, PUSCH_SNR_DB, seed=0x5F + rank)
    rx = PuschRx(100, tbs, 3, llr_shift=4, max_noi=MAX_PASSES, device=local, symbol_sz=2048)
    h_iq = torch.from_numpy(np.ascontiguousarray(np.tile(iq8, (nsf // nd, 1)))).pin_memory()
    x = h_iq.to(dev)
    nbytes = tbs // 8 + 3
    h_data = torch.empty((nsf, rx.data_stride), dtype=torch.uint8).pin_memory()
    steps, warm = max(3, min(args.steps, 10)), 2

    ok, its = None, None
    for _ in range(warm):
        ok, its = rx.run(x, nsf)
    good = bool(ok.all()) and bool((rx.data[:nd, :nbytes].cpu().numpy() == payload8).all())
    rx_grid_check = rx.grid[:nd].cpu().numpy()
    # device-resident: IQ already in HBM; the decode entry is synchronous, so wall clock between two device syncs
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        rx.run(x, nsf)
    torch.cuda.synchronize()
    ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
    # front-end kernels alone (CUDA events on the launching stream)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fe_ofdm = fe_demap = 0.0
    st = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(steps):
        e[0].record()
        rx.ofdm.rx_sf_device(x, rx.grid, nsf, st)
        e[1].record()
        rx._lib.srsran_b200_pusch_demap_batch(local, 3, rx.grid.data_ptr(), rx.llr.data_ptr(), nsf, 14, rx.nof_re, 0x3BF7, 4, 1, st)
        e[2].record()
        torch.cuda.synchronize()
        fe_ofdm += e[0].elapsed_time(e[1]) / steps
        fe_demap += e[1].elapsed_time(e[2]) / steps
    # end to end: pinned host IQ in, transport block bytes out, every step
    ms_e2e = pusch_e2e(torch, dev, steps, h_iq, x, lambda xb: rx.run(xb, nsf), rx.data[:nsf], h_data, barrier, max_over_ranks)
    mean_its = sum_over_ranks(float(its.mean())) / world
    rx.close()
    # CPU figure for the OFDM stage alone.  The reference's srsran_ofdm_rx_sf sits on FFTW, which is not available here,
    # so this is a SUBSTITUTE: the same windows, half-subcarrier shift, 2048-point transforms and bin selection with scipy.fft
    # (pocketfft, complex64) on all host cores -- reported, not a target.
    ofdm_cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            import scipy.fft as sfft

            cores = os.cpu_count() or 1
            n_cpu = 64 * cores
            xs = np.ascontiguousarray(np.tile(iq8, (n_cpu // nd, 1)))
            N, cp1, cp2, noff = 2048, 160, 144, 72
            nn = np.arange(N)
            shift = np.exp(-1j * np.pi * (nn - noff) / N).astype(np.complex64)
            ramp = np.exp(2j * np.pi * noff * np.concatenate([np.arange(N - 600, N), np.arange(0, 600)]) / N).astype(np.complex64)
            t0 = time.perf_counter()
            grid_cpu = np.empty((n_cpu, 14, 1200), np.complex64)
            for l in range(14):
                slot, ls = divmod(l, 7)
                start = slot * (15 * N // 2) + cp1 + ls * (N + cp2) - noff
                X = sfft.fft(xs[:, start:start + N] * shift, axis=1, workers=cores)
                grid_cpu[:, l, :600] = X[:, N - 600:]
                grid_cpu[:, l, 600:] = X[:, :600]
                grid_cpu[:, l, :] *= ramp
            sec = time.perf_counter() - t0
            ref_grid = rx_grid_check
            err = float(np.linalg.norm(grid_cpu[:nd] - ref_grid) / np.linalg.norm(ref_grid))
            ofdm_cpu = {"value": n_cpu / sec, "unit": "subframes/s", "cores": cores, "kind": "substitute",
                        "sample": f"{n_cpu} subframes, scipy.fft complex64 with {cores} workers; FFTW (the reference's DFT) is not available; "
                                  f"relative L2 distance to the GPU grid {err:.2e}"}
        except Exception as ex:  # noqa: BLE001
            ofdm_cpu = {"error": repr(ex)}
    ofdm_bytes = nsf * (15 * 2048 * 8 + 14 * 1200 * 8)
    demap_bytes = nsf * 12 * 1200 * (8 + 12)
    return {"metric": "pusch_subframes_per_s_20mhz_64qam_tbs75376", "value": world * nsf / (ms * 1e-3), "unit": "subframes/s",
            "ms_per_step": ms, "subframes_per_gpu_per_step": nsf, "info_gbit_per_s": world * nsf * tbs / (ms * 1e-3) / 1e9,
            "mean_passes": mean_its, "snr_db": PUSCH_SNR_DB, "all_tb_crc_ok_and_bytes_equal_payload": good,
            "e2e": {"value": world * nsf / (ms_e2e * 1e-3), "unit": "subframes/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": int(nsf * 15 * 2048 * 8), "d2h_bytes_per_step": int(nsf * rx.data_stride)},
            "front_end": {"ofdm_ms": fe_ofdm, "ofdm_gbs": ofdm_bytes / (fe_ofdm * 1e-3) / 1e9, "ofdm_frac_of_hbm_peak":
                          ofdm_bytes / (fe_ofdm * 1e-3) / 1e9 / peaks["hbm_gbs"], "demap_ms": fe_demap,
                          "demap_gbs": demap_bytes / (fe_demap * 1e-3) / 1e9,
                          "demap_frac_of_hbm_peak": demap_bytes / (fe_demap * 1e-3) / 1e9 / peaks["hbm_gbs"],
                          "ofdm_cpu_substitute": ofdm_cpu},
            "config": "configs[3] pipeline batched as configs[4]: 100 PRB, N=2048, normal CP, f=-0.5, window offset 0.5, 64QAM, "
                      "TBS 75376 -> 13 x K=5824, rv 0, identity channel, 8 distinct subframes tiled, soft bits >> 4"}


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


def pusch_full_leg(args, torch, dev, local, rank, world, barrier, max_over_ranks, sum_over_ranks, peaks):
    """The same 20 MHz subframe through the COMPLETE receive chain (SURVEY 8f ranks 1-3 added to config 4): transmit side with
    channel interleaver, scrambling, transform precoding and DMRS; per-subframe flat fading + timing offset + AWGN; receive side
    OFDM rx -> channel estimation -> MMSE equaliser + transform de-precoding -> soft demap + descrambling + UL-SCH de-interleave
    -> rate de-matching -> turbo decoding with CRC early stop."""
    import numpy as np

    from srslte_b200 import synth_pusch as sp
    from srslte_b200.pusch import PuschRxFull

    nsf, tbs, nd, cell_id = PUSCH_SF_PER_GPU, 75376, 8, 1 + rank
    rx = PuschRxFull(cell_id, 100, tbs, 3, llr_shift=4, max_noi=MAX_PASSES, device=local, symbol_sz=2048)
    rnti8 = np.arange(nd, dtype=np.uint32) * 97 + 62
    tti8 = np.arange(nd, dtype=np.uint32) * 3 + rank
    iq8, payload8, G = sp.make_subframes_full(cell_id, 100, 2048, tbs, 6, 0, sp.qpp_interleaver(5824), nd, rnti8, tti8,
                                              lambda sf: rx.chain.dmrs(sf, 0), PUSCH_SNR_DB, seed=0x77 + rank)
    rnti, tti = np.tile(rnti8, nsf // nd), np.tile(tti8, nsf // nd)
    h_iq = torch.from_numpy(np.ascontiguousarray(np.tile(iq8, (nsf // nd, 1)))).pin_memory()
    x = h_iq.to(dev)
    nbytes = tbs // 8 + 3
    h_data = torch.empty((nsf, rx.data_stride), dtype=torch.uint8).pin_memory()
    steps, warm = max(3, min(args.steps, 10)), 2
    ok, its = None, None
    for _ in range(warm):
        ok, its = rx.run(x, nsf, rnti, tti)
    good = bool(ok.all()) and bool((rx.data[:nd, :nbytes].cpu().numpy() == payload8).all())
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        rx.run(x, nsf, rnti, tti)
    torch.cuda.synchronize()
    ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
    # the new front-end kernels alone (CUDA events on the launching stream)
    ch = rx.chain
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_chest = t_eq = t_demod = 0.0
    st = torch.cuda.current_stream(dev).cuda_stream
    grid = rx.grid[:nsf]
    ce, meas = ch.chest(grid, tti)          # outputs allocated once; the timed calls below only launch kernels
    d = ch.equalize_deprecode(grid, ce, meas)
    ch.demod_descramble(d, rnti, tti, out=rx.llr)
    torch.cuda.synchronize()
    for _ in range(steps):
        # each stage between its own pair of events, with the device idle before it, so that host-side argument marshalling
        # (the per-subframe rnti/tti arrays are copied to the device inside the calls) is not counted as kernel time
        ch.chest(grid, tti, out=(ce, meas))
        torch.cuda.synchronize()
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
    ch.chest(grid, tti, out=(ce, meas))     # leave the buffers as the real parameters produce them
    ch.demod_descramble(ch.equalize_deprecode(grid, ce, meas, out=d), rnti, tti, out=rx.llr)

    ms_e2e = pusch_e2e(torch, dev, steps, h_iq, x, lambda xb: rx.run(xb, nsf, rnti, tti), rx.data[:nsf], h_data, barrier, max_over_ranks)
    # the same with the samples in the radio's int16 I/Q wire format (SRSRAN_B200_FLAG_IQ_INT16): the RF front ends deliver sc16
    # and srsRAN converts to float on the host; here the first FFT pass converts, and half the bytes cross PCIe.  AGC-like scaling
    # to +-0.5 full scale before quantisation; the transport blocks must still decode to the same bytes.
    peak = float(np.abs(iq8.view(np.float32)).max())
    q8 = np.round(iq8.view(np.float32).reshape(nd, -1, 2) * (16384.0 / peak)).astype(np.int16)
    h_iq16 = torch.from_numpy(np.ascontiguousarray(np.tile(q8, (nsf // nd, 1, 1)))).pin_memory()
    x16 = h_iq16.to(dev)
    ok16, _ = rx.run(x16, nsf, rnti, tti)
    good16 = bool(ok16.all()) and bool((rx.data[:nd, :nbytes].cpu().numpy() == payload8).all())
    ms_e2e16 = pusch_e2e(torch, dev, steps, h_iq16, x16, lambda xb: rx.run(xb, nsf, rnti, tti), rx.data[:nsf], h_data, barrier,
                         max_over_ranks)
    del x16
    # the same step through the native one-call entry (srsran_b200_enb_ul_pusch_batch): pinned host samples in, transport-block
    # bytes out, nothing but the C ABI in between (chunked copies on a second stream inside the call; no overlap across calls)
    from srslte_b200.pusch import EnbUl, PUSCH_RES_DTYPE

    enb = EnbUl(cell_id, 100, tbs, 3, llr_shift=4, max_noi=MAX_PASSES, device=local, symbol_sz=2048)
    h_out = torch.empty((nsf, enb.tb_bytes), dtype=torch.uint8).pin_memory()
    res_np = np.zeros(nsf, PUSCH_RES_DTYPE)
    native = {}
    for name, hbuf, fl in (("float_iq", h_iq, 0), ("int16_iq", h_iq16, 8)):
        enb.run_ptr(hbuf.data_ptr(), nsf, rnti, tti, h_out.data_ptr(), res_np, flags=fl)
        okn = bool(res_np["crc_ok"].all()) and bool((h_out[:nd].numpy() == payload8).all())
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            enb.run_ptr(hbuf.data_ptr(), nsf, rnti, tti, h_out.data_ptr(), res_np, flags=fl)
        msn = max_over_ranks((time.perf_counter() - t0) * 1e3) / steps
        native[name] = {"value": world * nsf / (msn * 1e-3), "unit": "subframes/s", "ms_per_step": msn,
                        "all_tb_crc_ok_and_bytes_equal_payload": okn}
    enb.close()
    mean_its = sum_over_ranks(float(its.mean())) / world
    snr_est = float(rx.meas[:nd, 1].log10().mean().item() * 10.0)
    # CPU baseline: the reference's own receiver after the OFDM demodulator (FFTW is not available to build its srsran_ofdm)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import loader

            if loader.have_ref():
                R = loader.api("ref")
                cores = os.cpu_count() or 1
                grids8 = rx.grid[:nd].cpu().numpy()
                n_cpu = 16 * cores
                lk = loader.pusch_link(cell_id=cell_id, rnti=int(rnti8[0]), tti=int(tti8[0]), tbs=tbs, max_iter=MAX_PASSES)
                okc, sec = R.pusch_rx_bench(lk, np.ascontiguousarray(np.tile(grids8[:1], (n_cpu, 1, 1))), cores)
                cpu = {"value": n_cpu / sec, "unit": "subframes/s", "cores": cores, "kind": "reference",
                       "sample": f"{n_cpu} copies of one subframe's resource grid (demodulated on the GPU: the reference's OFDM needs FFTW, "
                                 f"absent here): srsran_chest_ul_estimate_pusch + srsran_pusch_decode per subframe, one object set per thread, "
                                 f"all crc ok = {bool(okc.all())}"}
        except Exception as ex:  # noqa: BLE001
            cpu = {"error": repr(ex)}
    rx.close()
    M = 1200
    chest_bytes = nsf * (2 * M * 8 * 2 + 2 * M * 8)          # two DMRS symbols + known sequence in, two slot estimates out
    eq_bytes = nsf * (12 * M * 8 + 2 * M * 8 + 12 * M * 8)   # data symbols + estimates in, de-precoded symbols out
    demod_bytes = nsf * 12 * M * (8 + 12 + 6 / 8.0 * 2)      # symbols in, soft bits out, scrambling bits written + read
    gbs = lambda b, t: b / (t * 1e-3) / 1e9
    return {"metric": "pusch_full_chain_subframes_per_s_20mhz_64qam_tbs75376", "value": world * nsf / (ms * 1e-3), "unit": "subframes/s",
            "ms_per_step": ms, "subframes_per_gpu_per_step": nsf, "info_gbit_per_s": world * nsf * tbs / (ms * 1e-3) / 1e9,
            "mean_passes": mean_its, "snr_db": PUSCH_SNR_DB, "estimated_snr_db": snr_est, "all_tb_crc_ok_and_bytes_equal_payload": good,
            # end to end = ONE C-ABI call per step (srsran_b200_enb_ul_pusch_batch) with pinned host samples in and host bytes out
            "e2e": dict(native["float_iq"], h2d_bytes_per_step=int(nsf * 15 * 2048 * 8), d2h_bytes_per_step=int(nsf * (tbs // 8 + 3)),
                        note="srsran_b200_enb_ul_pusch_batch, float I/Q; copies chunked on a second stream inside the call"),
            "e2e_int16_iq": dict(native["int16_iq"], h2d_bytes_per_step=int(nsf * 15 * 2048 * 4), d2h_bytes_per_step=int(nsf * (tbs // 8 + 3)),
                                 note="same call with the samples as int16 I/Q pairs (radio wire format), converted in the first FFT pass"),
            # the stage entries driven from Python with the copy of step s+1 overlapping the processing of step s
            "e2e_pipelined_steps": {"float_iq": {"value": world * nsf / (ms_e2e * 1e-3), "unit": "subframes/s", "ms_per_step": ms_e2e},
                                    "int16_iq": {"value": world * nsf / (ms_e2e16 * 1e-3), "unit": "subframes/s", "ms_per_step": ms_e2e16,
                                                 "all_tb_crc_ok_and_bytes_equal_payload": good16}},
            "front_end": {"chest_ms": t_chest, "chest_gbs": gbs(chest_bytes, t_chest), "chest_frac_of_hbm_peak": gbs(chest_bytes, t_chest) / peaks["hbm_gbs"],
                          "equalize_deprecode_ms": t_eq, "equalize_deprecode_gbs": gbs(eq_bytes, t_eq),
                          "equalize_deprecode_frac_of_hbm_peak": gbs(eq_bytes, t_eq) / peaks["hbm_gbs"],
                          "demod_descramble_deinterleave_ms": t_demod, "demod_descramble_deinterleave_gbs": gbs(demod_bytes, t_demod),
                          "demod_descramble_deinterleave_frac_of_hbm_peak": gbs(demod_bytes, t_demod) / peaks["hbm_gbs"]},
            "cpu_baseline": cpu,
            "config": "configs[3]/[4] with the complete chain: 100 PRB, N=2048, normal CP, f=-0.5, window offset 0.5, 64QAM, TBS 75376 -> "
                      "13 x K=5824, rv 0, DMRS + channel interleaver + scrambling + transform precoding, flat fading with timing offset per "
                      "subframe, 8 distinct subframes tiled, soft bits >> 4"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pusch", action="store_true", help="skip the secondary PUSCH subframes/s leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
