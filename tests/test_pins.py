"""CPU: the oracle's resampler against pins that do not depend on the restated rubato source (tests/pins.py)."""
import numpy as np
import pytest

from oracle import np_oracle, sko
from tests import pins


def _oracle_resampler(in_rate, out_rate, chunk, channels, impl=sko.FastFixedIn):
    r = impl(in_rate, out_rate, chunk, channels)
    return lambda chunks: [r.process(c) for c in chunks]


@pytest.mark.parametrize("impl", [sko.FastFixedIn, np_oracle.FastFixedIn])
@pytest.mark.parametrize("in_rate,out_rate,chunk,channels", [(48000, 16000, 960, 2), (48000, 24000, 480, 2), (48000, 8000, 960, 1),
                                                             (48000, 24000, 960, 1), (32000, 16000, 640, 2)])
def test_integer_ratio_outputs_are_input_samples(impl, in_rate, out_rate, chunk, channels):
    pins.check_integer_ratio_identity(_oracle_resampler(in_rate, out_rate, chunk, channels, impl), in_rate, out_rate, chunk, channels)


@pytest.mark.parametrize("in_rate,out_rate,chunk,channels", [(44100, 48000, 882, 2), (48000, 44100, 960, 1), (16000, 48000, 320, 2),
                                                             (22050, 48000, 441, 1), (48000, 16000, 960, 2), (8000, 48000, 160, 2)])
def test_ramp_in_ramp_out(in_rate, out_rate, chunk, channels):
    n = pins.check_ramp(_oracle_resampler(in_rate, out_rate, chunk, channels), in_rate, out_rate, chunk, channels)
    assert n > 0


@pytest.mark.parametrize("in_rate,out_rate,chunk", [(44100, 48000, 882), (48000, 44100, 960), (8000, 44100, 160), (48000, 16000, 960),
                                                    (22050, 48000, 441), (11025, 48000, 441)])
def test_output_counts_match_exact_rational_arithmetic_over_10k_chunks(in_rate, out_rate, chunk):
    def counts(n):
        r = sko.FastFixedIn(in_rate, out_rate, chunk, 1)
        z = np.zeros(chunk, np.float32)
        return [r.process(z).size for _ in range(n)]
    pins.check_counts(counts, in_rate, out_rate, chunk, 10000)


def test_reference_length_asserts():
    # resampler.rs:826-837: one 960-sample stereo 48 kHz packet -> 24 kHz; "approximately" 480 samples: |n - 480| < 10
    n = sko.ResamplerNode(24000, 960, 0)
    n.push(48000, 2, np.full(960, 0.5, np.float32))
    n.finish()
    total = sum(p["samples"].size for p in n.out)
    assert abs(total - 480) < 10 and total == 2 * pins.exact_total_after(48000, 24000, 480, 1)
    # resampler.rs:886-906: three 480-sample packets buffer up; the first emitted packet is non-empty
    n = sko.ResamplerNode(24000, 960, 0)
    for _ in range(3):
        n.push(48000, 2, np.full(480, 0.5, np.float32))
    n.finish()
    assert n.out and n.out[0]["samples"].size > 0
