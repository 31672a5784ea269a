"""The run-table representation of rubato's f64 phase recurrence (streamkit_b200/csrc/phase_runs.h, the same
header the CUDA kernels compile) must reproduce the sequential `idx += t` chain BIT-EXACTLY: every element,
the output count and the carried last_index, over long streams and adversarial ratios."""
import ctypes as C
import math
import os
import random

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(HERE), "streamkit_b200", "csrc", "libsk_phase_host.so")


@pytest.fixture(scope="module")
def check():
    lib = C.CDLL(SO)
    f = lib.skp_check_stream
    f.restype = C.c_uint64
    f.argtypes = [C.c_double, C.c_uint32, C.c_double, C.c_uint32] + [C.c_void_p] * 4

    def run(ratio, chunk, last_index, calls):
        mr, ov, lo, tot = C.c_uint32(), C.c_uint32(), C.c_double(), C.c_uint64()
        bad = f(ratio, chunk, last_index, calls, C.byref(mr), C.byref(ov), C.byref(lo), C.byref(tot))
        return bad, mr.value, ov.value, lo.value, tot.value

    return run


COMMON = [(44100, 48000, 882), (48000, 16000, 960), (48000, 24000, 480), (16000, 48000, 320), (8000, 48000, 160),
          (48000, 44100, 960), (44100, 16000, 882), (22050, 48000, 441), (48000, 8000, 960), (96000, 48000, 1920),
          (11025, 48000, 221), (48000, 48001, 960), (32000, 48000, 640), (48000, 16000, 20), (32000, 48000, 7)]


@pytest.mark.parametrize("i,o,chunk", COMMON)
def test_common_ratios_long_streams(check, i, o, chunk):
    bad, max_runs, ovf, last, total = check(o / i, chunk, -4.0, 5000)
    assert bad == 0 and ovf == 0
    assert max_runs <= 64
    assert total > 0 or chunk < 16


def test_known_orbits(check):
    # 48k -> 16k: t = 3.0 exactly, steady last_index = -10, 318 then 320 frames
    bad, _, _, last, total = check(16000 / 48000, 960, -4.0, 1)
    assert (bad, last, total) == (0, -10.0, 318)
    bad, _, _, last, total = check(16000 / 48000, 960, -10.0, 1)
    assert (bad, last, total) == (0, -10.0, 320)
    bad, _, _, last, total = check(24000 / 48000, 480, -4.0, 1)
    assert (bad, total) == (0, 237)


def test_random_ratios_and_phases(check):
    rnd = random.Random(12345)
    n = 0
    worst = 0
    while n < 1500:
        i, o = rnd.randint(4000, 192000), rnd.randint(4000, 192000)
        if not (1 / 16 < o / i < 16):
            continue
        chunk = rnd.randint(1, 3000)
        t = i / o
        L = -4.0 if rnd.random() < 0.5 else -(9 + math.ceil(t)) + rnd.random() * t
        bad, mr, ovf, _, _ = check(o / i, chunk, L, 60)
        assert bad == 0 and ovf == 0, (i, o, chunk, L)
        worst = max(worst, mr)
        n += 1
    assert worst <= 96


def test_tie_prone_ratios(check):
    """t with few mantissa bits makes every addition in some binade an exact tie or exact: the run rule
    (anchor only after three same-binade elements) must still hold."""
    for num, den in [(3, 2), (5, 4), (9, 8), (17, 16), (3, 4), (5, 8), (7, 8), (1, 3), (2, 3), (1, 7), (255, 256), (257, 256)]:
        for chunk in (64, 441, 960, 2880):
            bad, _, ovf, _, _ = check(den / num, chunk, -4.0, 300)
            assert bad == 0 and ovf == 0, (num, den, chunk)
    # t = 1 + 2^-52 * k: increments that are odd multiples of half an ulp in the upper binades
    for k in (1, 3, 5, 1023, 4097):
        t = 1.0 + k * 2.0 ** -52
        bad, _, ovf, _, _ = check(1.0 / t, 960, -4.0, 300)
        assert bad == 0 and ovf == 0, k
