"""Drop-in proof with the reference's OWN code: its unit-test programs (turbodecoder_test.c, rm_turbo_test.c, ofdm_test.c,
pusch_test.c) are compiled from /root/reference where they lie (oracle/Makefile target `dropin`; nothing is copied) and linked
against libsrslte_b200.so IN FRONT OF the reference's remaining sources, so the turbo decoder, rate de-matcher, DFT plans,
OFDM receiver, CRC, segmentation and interleaver tables resolve to this repo's library -- for the programs and for the
reference's own sch.c / pusch.c / ofdm.c / turbocoder.c that call them.  The programs run with the argument sets of the
reference's CMakeLists (ctest); exit code 0 = pass, exactly as ctest judges them.

The programs are built here (where the reference tree exists) into oracle/_ref/dropin/ and travel to the GPU box as built
files; the -m gpu tests only execute them."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "dropin")
PROGS = ["turbodecoder_test", "rm_turbo_test", "ofdm_test", "pusch_test", "turbodecoder_test_ref"]
HAVE_REF = os.path.exists("/root/reference/lib/include/srsran/config.h")


def build():
    from srslte_b200 import build as b

    b.build_library()
    from oracle import loader

    loader.build()
    subprocess.check_call(["make", "-s", "dropin"], cwd=os.path.join(ROOT, "oracle"))


def have_programs():
    return all(os.path.exists(os.path.join(DROPIN, p)) for p in PROGS)


@pytest.mark.skipif(not HAVE_REF, reason="reference tree absent: the prebuilt programs are used as they are")
def test_reference_programs_build_and_bind_to_this_library():
    """Every replaced entry point the programs (and the reference's own callers inside libsrsref.so) use must bind to
    libsrslte_b200.so; the reference's transmit side stays the reference's."""
    build()
    assert have_programs()
    env = dict(os.environ, LD_DEBUG="bindings")
    r = subprocess.run([os.path.join(DROPIN, "turbodecoder_test"), "-n", "1", "-l", "40", "-e", "2"], capture_output=True, text=True, env=env)
    bind = {}
    for m in re.finditer(r"binding file (\S+) \[0\] to (\S+) \[0\]: normal symbol `(srsran_\w+)'", r.stderr):
        bind.setdefault(m.group(3), set()).add((os.path.basename(m.group(1)), os.path.basename(m.group(2))))
    assert ("turbodecoder_test", "libsrslte_b200.so") in bind["srsran_tdec_init_manual"]
    assert ("turbodecoder_test", "libsrslte_b200.so") in bind["srsran_tdec_run_all"]
    # the reference's turbo ENCODER (libsrsref.so) takes its interleaver table from this library
    assert ("libsrsref.so", "libsrslte_b200.so") in bind["srsran_tc_interl_LTE_gen"]
    assert ("turbodecoder_test", "libsrsref.so") in bind["srsran_tcod_encode"]


def run_prog(name, *args, timeout=300):
    exe = os.path.join(DROPIN, name)
    if not os.path.exists(exe):
        if HAVE_REF:
            build()
        else:
            pytest.skip(f"{exe} not prebuilt and no reference tree to build it from")
    r = subprocess.run([exe, *[str(a) for a in args]], capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, f"{name} {' '.join(map(str, args))}: exit {r.returncode}\n{r.stdout[-1500:]}\n{r.stderr[-1500:]}"
    return r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("args", [("-n", 100, "-s", 1, "-l", 504, "-e", 1.0, "-t"), ("-n", 100, "-s", 1, "-l", 504, "-e", 2.0, "-t"),
                                  ("-n", 100, "-s", 1, "-l", 6144, "-e", 1.5, "-t"), ("-n", 1, "-s", 1, "-k", "-e", 0.5)])
def test_reference_turbodecoder_test(args):
    """lib/src/phy/fec/turbo/test/CMakeLists.txt: turbodecoder_test_504_1, _504_2, _6114_1_5, _known."""
    out = run_prog("turbodecoder_test", *args)
    assert "Done" in out
    # the program only reports its bit errors (ctest judges the exit code).  Stronger: the same program on the reference
    # alone, with the reference's GENERIC int16 decoder selected (-d 1 = SRSRAN_TDEC_GENERIC), must report the same
    # number of bit errors for the same seed -- the decoder this library replaces it with is bit-exact with that one.
    ref = run_prog("turbodecoder_test_ref", *args, "-d", 1)
    errs = lambda txt: (re.findall(r"(\d+) Errors", txt) or ["0"])[-1]
    assert errs(out) == errs(ref), (errs(out), errs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("args", [("-e", 1920), ("-e", 8192), ("-c", 0, "-e", 200), ("-c", 100, "-e", 9000, "-i", 2), ("-c", 187, "-e", 20000, "-i", 3)])
def test_reference_rm_turbo_test(args):
    """rm_turbo_test_1 / _2 of the reference's CMakeLists plus three points of its -c / -i (code block size, rv) sweep: the
    program rate-matches with the reference's transmitter and compares srsran_rm_turbo_rx_lut_ (this library) with the
    reference's float de-matcher."""
    run_prog("rm_turbo_test", *args)


@pytest.mark.gpu
@pytest.mark.parametrize("args", [("-r", 1), ("-e", "-r", 1), ("-s", 0.5, "-r", 1), ("-o", 0.5, "-r", 1), ("-N", 4096, "-r", 1),
                                  ("-e", "-o", 0.5, "-s", 0.5, "-N", 4096, "-r", 1)])
def test_reference_ofdm_test(args):
    """lib/src/phy/dft/test/CMakeLists.txt: ofdm_normal, _extended, _shifted, _offset, _force, _extended_shifted_offset_force.
    The reference's OFDM TRANSMITTER (its ofdm.c, on this library's DFT plans) feeds this library's receiver; the program
    fails when the round-trip error reaches 1e-4."""
    out = run_prog("ofdm_test", *args)
    assert "MSE too large" not in out


def pusch_cases():
    cases = []
    for cell, mcs_list in ((50, (0, 7, 14, 21, 28)), (100, (7, 28)), (75, (14,))):
        for mcs in mcs_list:
            for ack in (0, 1, 2, 10):
                for cqi in ("none", "wideband"):
                    a = ["-n", cell, "-L", 50]
                    m = mcs
                    if ack:
                        a += ["-p", "uci_ack", ack]
                        m = 27 if m == 28 else m
                    if cqi != "none":
                        a += ["-p", "cqi", cqi]
                        m = 27 if m == 28 else m
                    if m > 24:
                        a += ["-p", "enable_64qam"]
                    a += ["-m", m]
                    cases.append(tuple(a))
    return cases


@pytest.mark.gpu
def test_reference_pusch_test():
    """lib/src/phy/phch/test/CMakeLists.txt pusch_test sweep (cells of 50 / 75 / 100 PRB, L = 50 PRB, MCS 0..28, HARQ-ACK and
    CQI multiplexed or not): the reference's srsran_pusch_encode -> srsran_pusch_decode, whose sch.c / pusch.c call this
    library for transform (de)precoding, rate de-matching, turbo decoding, CRC and segmentation."""
    for a in pusch_cases():
        run_prog("pusch_test", *a)
