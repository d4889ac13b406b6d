#!/usr/bin/env python3
"""Generates tests/golden/*.npz by running the REFERENCE's own code (oracle/_ref/libsrsref.so, built in place from
/root/reference by oracle/Makefile).  Run in the authoring container; the fixtures are committed so that the oracle
port and the CUDA path can be checked on machines where /root/reference does not exist.

usage: python tools/gen_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import coded_llrs  # noqa: E402
from oracle import loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    loader.build()
    R = loader.api("ref")
    os.makedirs(OUT, exist_ok=True)

    # ---- tables: sizes, QPP, rate matching, segmentation, CRC ---------------------------------------------------
    Ks = R.cb_sizes()
    qpp_sha = []
    for K in Ks:
        f, r = R.interleaver(int(K))
        qpp_sha.append(sha(f) + sha(r))
    rm_sha = [[sha(R.rm_table(i, rv)) for rv in range(4)] for i in range(len(Ks))]
    tbs_list = np.array([16, 40, 256, 1000, 2216, 6120, 6121, 6200, 12960, 30576, 36696, 51024, 75376, 97896], np.uint32)
    segm = np.array([[R.cbsegm(int(t))[k] for k in ("F", "C", "K1", "K2", "K1_idx", "K2_idx", "C1", "C2", "tbs")] for t in tbs_list], np.uint32)
    rng = np.random.default_rng(2024)
    crc_msg = rng.integers(0, 256, 2048).astype(np.uint8)
    crc_lens = np.array([8, 24, 40, 1000, 6144, 16384], np.int32)
    crc_vals = np.array([[R.crc24(k, crc_msg, int(n)) for n in crc_lens] for k in "AB"], np.uint32)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), Ks=Ks, qpp_sha=np.array(qpp_sha), rm_sha=np.array(rm_sha),
                        rm_table_k40_rv0=R.rm_table(0, 0), rm_table_k6144_rv2=R.rm_table(187, 2),
                        tbs_list=tbs_list, segm=segm, crc_msg=crc_msg, crc_lens=crc_lens, crc_vals=crc_vals)

    # ---- encoder + rate matching ------------------------------------------------------------------------------------
    enc = {}
    for K in (40, 512, 6144):
        bits = rng.integers(0, 2, K).astype(np.uint8)
        cw = R.tcod_encode(bits)
        enc[f"bits_{K}"] = bits
        enc[f"coded_{K}"] = cw
        for rv in range(4):
            E = int(1.3 * (3 * K + 12))
            enc[f"tx_{K}_rv{rv}"] = R.rm_tx(cw, K, E, rv)
            e = rng.integers(-60, 60, E).astype(np.int16)
            soft = rng.integers(-300, 300, 3 * K + 12).astype(np.int16)
            enc[f"rx_in_{K}_rv{rv}"] = e
            enc[f"rx_soft_{K}_rv{rv}"] = soft.copy()
            buf = np.concatenate([soft, np.zeros(64, np.int16)])
            R.rm_rx(e, buf, R.cbindex(K), rv)
            enc[f"rx_out_{K}_rv{rv}"] = buf[:3 * K + 12].copy()
    np.savez_compressed(os.path.join(OUT, "coding.npz"), **enc)

    # ---- decoder: generic int16 implementation, decisions after every pass + the decode_tb_cb-style loop --------
    dec = {}
    cases = [(40, 6, 0.8, 16.0, 31), (512, 4, 0.95, 16.0, 31), (1008, 3, 1.0, 16.0, 31), (6144, 2, 0.93, 16.0, 31),
             (2048, 2, 0.9, 500.0, 2000), (6144, 1, 0.8, 8000.0, 30000)]
    for ci, (K, ncb, sigma, scale, clip) in enumerate(cases):
        llr, bits = coded_llrs(R, K, ncb, sigma, scale, clip, seed=100 + ci)
        dec[f"c{ci}_meta"] = np.array([K, ncb, clip], np.int64)
        dec[f"c{ci}_llr"] = llr
        dec[f"c{ci}_bits"] = np.packbits(bits, axis=1)
        dec[f"c{ci}_per_pass"] = np.stack([R.tdec_passes(llr[c], K, 8, loader.TDEC_GENERIC) for c in range(ncb)])
        for es in (0, 1):
            out, ok, npass, _ = R.decode_batch(llr, K, 8, "B", 0, bool(es), 1, loader.TDEC_GENERIC)
            dec[f"c{ci}_loop{es}_out"] = out
            dec[f"c{ci}_loop{es}_ok"] = ok
            dec[f"c{ci}_loop{es}_npass"] = npass
    np.savez_compressed(os.path.join(OUT, "tdec.npz"), **dec)

    # ---- OFDM rx + 64QAM/16QAM demap --------------------------------------------------------------------------------
    of = {}
    cfgs = [(6, 0, 0, 0.0, 0.0, 0, 0), (6, 0, 0, -0.5, 0.5, 0, 0), (15, 0, 1, 0.0, 0.0, 1, 0), (25, 0, 0, 0.5, 0.25, 0, 1),
            (25, 512, 0, -0.5, 0.5, 0, 0), (100, 2048, 0, -0.5, 0.5, 0, 0)]
    for i, (prb, N, cp, fs, wo, nm, kd) in enumerate(cfgs):
        n = N or R.symbol_sz(prb)
        x = (rng.normal(size=15 * n) + 1j * rng.normal(size=15 * n)).astype(np.complex64)
        y, _ = R.ofdm_rx(x, prb, bool(cp), N, fs, wo, bool(nm), bool(kd))
        of[f"cfg{i}"] = np.array([prb, N, cp, fs, wo, nm, kd], np.float64)
        of[f"in{i}"] = x
        of[f"out{i}"] = y
    for mod, name in ((1, "qpsk"), (2, "qam16"), (3, "qam64")):
        s = ((rng.normal(size=1003) + 1j * rng.normal(size=1003)) * 0.7).astype(np.complex64)
        s[:4] = [50 + 3j, -60 - 1j, 46.81 - 46.82j, 0.5 * 1j]
        of[f"{name}_sym"] = s
        of[f"{name}_llr"] = R.demod_s(mod, s)
    np.savez_compressed(os.path.join(OUT, "ofdm_demod.npz"), **of)

    # ---- PUSCH chain between OFDM and de-matching: DMRS, chest, equaliser, transform de-precoding, descrambling, de-interleave
    pc = {}
    links = [loader.pusch_link(cell_id=301, nof_prb=25, L_prb=12, n_prb=7, mod=2, tbs=4584, tti=7, n_dmrs=3, cyclic_shift=5, delta_ss=11, rnti=4660),
             loader.pusch_link(cell_id=9, nof_prb=6, L_prb=3, n_prb=2, mod=1, tbs=392, tti=12, rnti=62),
             loader.pusch_link(cell_id=77, nof_prb=15, L_prb=10, n_prb=0, mod=3, tbs=5160, tti=5, group_hopping=1, rnti=65535),
             # 1 and 2 PRB: the phi(n) base sequences of TS 36.211 5.5.1.2 (appended: the vectors above do not change)
             loader.pusch_link(cell_id=12, nof_prb=6, L_prb=1, n_prb=4, mod=1, tbs=136, tti=3, rnti=101),
             loader.pusch_link(cell_id=250, nof_prb=25, L_prb=2, n_prb=11, mod=2, tbs=328, tti=8, n_dmrs=5, group_hopping=1, delta_ss=17, rnti=9)]
    for i, lk in enumerate(links):
        data = rng.integers(0, 256, int(lk[12]) // 8, dtype=np.uint8)
        tx = R.pusch_encode(lk, data)
        noise = (rng.normal(size=tx.shape) + 1j * rng.normal(size=tx.shape)).astype(np.complex64) * np.float32(0.02)
        rx = (tx * np.complex64(0.8 * np.exp(0.7j)) + noise).astype(np.complex64)
        res = R.pusch_decode(lk, rx)
        assert res["crc"] and (res["data"] == data).all()
        pc[f"link{i}"] = lk
        pc[f"data{i}"] = data
        pc[f"rx{i}"] = rx
        pc[f"dmrs{i}"] = R.dmrs_pusch_gen(lk)
        for k in ("d", "q", "g", "ce"):
            pc[f"{k}{i}"] = res[k]
        pc[f"meas{i}"] = np.array([res["noise"], res["snr"], res["cfo_hz"]], np.float32)
    np.savez_compressed(os.path.join(OUT, "pusch_chain.npz"), **pc)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
