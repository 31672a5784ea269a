"""Frame programs of the fused chain (streamkit_b200/csrc/chain_prog.h, the same header k_phase_chain and k_chain
compile). The host shim simulates a stream tick by tick exactly as k_phase_chain does (generator -> tail of the
pending packet -> part 1 of the next), then executes every emitted packet's program the way k_chain's consumers
do (block map -> segments -> lanes, including the integer floor / fraction split of FAST run segments) and
compares each frame's (buffer offset, f32 fraction) BIT FOR BIT with rubato's plain sequential recurrence
(`idx += t`; floor; `(idx - floor) as f32`, resampler.rs:404-407 -> rubato FastFixedIn::process_into_buffer).
Every packet frame must be produced by exactly one (segment, lane)."""
import ctypes as C
import os
import random

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(os.path.dirname(HERE), "streamkit_b200", "csrc", "libsk_phase_host.so")


@pytest.fixture(scope="module")
def check():
    lib = C.CDLL(SO)
    f = lib.skc_check_stream
    f.restype = C.c_uint64
    f.argtypes = [C.c_double] + [C.c_uint32] * 6 + [C.c_void_p] * 4

    def run(ratio, chunk, F, channels, calls, cap_seg=64, cap_exp=1024):
        out = [C.c_uint32() for _ in range(4)]
        bad = f(ratio, chunk, F, channels, calls, cap_seg, cap_exp, *[C.byref(x) for x in out])
        packets, max_seg, max_exp, status = (x.value for x in out)
        return bad, packets, max_seg, max_exp, status

    return run


# (in_rate, out_rate, chunk_frames, output_frame_size): chunk * out / in == F, the fused chain's eligibility rule
ELIGIBLE = [(44100, 48000, 882, 960), (32000, 48000, 640, 960), (8000, 48000, 160, 960), (16000, 48000, 320, 960),
            (24000, 48000, 480, 960), (22050, 48000, 441, 960), (11025, 48000, 441, 1920), (44100, 48000, 2646, 2880),
            (44100, 48000, 441, 480), (47000, 48000, 940, 960), (96000, 48000, 1920, 960), (48000, 48000, 960, 960),
            (48000, 24000, 480, 240), (48000, 16000, 360, 120), (4000, 48000, 80, 960)]


@pytest.mark.parametrize("i,o,chunk,F", ELIGIBLE)
@pytest.mark.parametrize("channels", [1, 2])
def test_programs_reproduce_the_recurrence(check, i, o, chunk, F, channels):
    bad, packets, max_seg, max_exp, status = check(o / i, chunk, F, channels, 400)
    assert status == 0, status
    assert bad == 0
    assert packets >= 398           # a packet every tick once the stream has started
    assert max_seg <= 24 and max_exp <= 200


def test_steady_state_44k1_program_is_small(check):
    """BASELINE config #5: the record the mixing kernel stages per stream-tick stays below 1 KB."""
    bad, packets, max_seg, max_exp, status = check(48000 / 44100, 882, 960, 2, 2000)
    assert (bad, status) == (0, 0) and packets == 1999
    assert max_seg <= 10 and max_exp <= 32      # runs shorter than SKC_MIN_RUN (16) frames are explicit
    assert 64 + 32 * (max_seg + 4) + 8 * (max_exp + 8) <= 1024


def test_random_eligible_ratios(check):
    rnd = random.Random(4242)
    done = 0
    while done < 150:
        F = rnd.choice([120, 240, 480, 960, 1920, 2880])
        chunk = rnd.randint(max(16, F // 12), min(3 * F, 4000))
        # out/in chosen so that chunk frames yield exactly F output frames
        bad, packets, max_seg, max_exp, status = check(F / chunk, chunk, F, rnd.choice([1, 2]), 120)
        if status & 2:      # table overflow: such ratios are rejected at plan time (skgpu_plan_add_chain)
            continue
        assert bad == 0, (F, chunk)
        assert status == 0 or status == 4, (F, chunk, status)   # 4: tail needs more head frames than the kernel stages
        done += 1


def test_capacity_overflow_is_reported_not_silent(check):
    bad, packets, max_seg, max_exp, status = check(48000 / 8000, 160, 960, 2, 50, cap_seg=4, cap_exp=8)
    assert status & 2
