"""CPU tests of the oracle itself: the C restatement against (a) the reference's OWN C gain plugin compiled
from /root/reference (oracle/_ref), (b) every constant the reference's unit tests pin for this path,
(c) an independent numpy restatement, (d) committed golden vectors (tests/golden)."""
import os

import numpy as np
import pytest

from oracle import np_oracle, sko
from streamkit_b200 import synth
from tests import plugin_host as ph

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


# ---------------------------------------------------------------- gain

def test_gain_oracle_matches_reference_c_plugin_bit_exact():
    """oracle/_ref/libgain_plugin_c.so is examples/plugins/gain-native-c/gain_plugin.c compiled unmodified."""
    if not os.path.exists(sko.REF_GAIN_PLUGIN):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    p = ph.NativePlugin(sko.REF_GAIN_PLUGIN)
    assert p.kind == "gain_c" and p.inputs == ["in"] and p.outputs == ["out"]
    inst = p.create('{"gain": 1.0}')
    x = synth.uniform_pcm(1, 48000)
    for g in [0.0, 0.25, 0.5, 1.0, 1.7, 2.0, 3.999, 4.0]:
        ok, _ = inst.update_params('{"gain": %r}' % g)
        assert ok
        out = inst.process_audio(48000, 2, x)
        assert len(out) == 1 and out[0][0] == "out"
        assert np.array_equal(bits(out[0][1]["samples"]), bits(sko.gain(x, np.float32(g))))
    inst.destroy()


def test_gain_validation_table():
    # gain.rs:465-509
    for g in (0.0, 1.0, 2.0, 4.0, 0.5, 3.5):
        assert sko.gain_validate(g)[0] == 0
    for g in (4.1, -0.1, 100.0, -10.0):
        rc, msg = sko.gain_validate(g)
        assert rc == 2 and "must be between" in msg
    for g in (float("nan"), float("inf"), float("-inf")):
        rc, msg = sko.gain_validate(g)
        assert rc == 1 and "finite number" in msg


def test_gain_reference_constants():
    # gain.rs:283-286 (0.5*2 -> 1.0), :323-331, :402-404 (mute is exact 0.0), :433-435 (0.5*4 -> 2.0)
    assert np.all(np.abs(sko.gain(np.full(100, 0.5, np.float32), 2.0) - 1.0) < 1e-3)
    for v in (0.2, 0.4, 0.6):
        assert np.all(np.abs(sko.gain(np.full(20, v, np.float32), 0.5) - v * 0.5) < 1e-3)
    assert np.all(sko.gain(np.ones(20, np.float32), 0.0) == 0.0)
    assert np.all(np.abs(sko.gain(np.full(20, 0.5, np.float32), 4.0) - 2.0) < 1e-3)
    x = synth.uniform_pcm(3, 10000)
    assert np.array_equal(bits(sko.gain(x, 1.7)), bits(np_oracle.gain(x, 1.7)))


# ---------------------------------------------------------------- s16

def test_s16_definition_known_answers():
    x = np.array([0.0, 1.0, -1.0, 0.5, -0.5, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768, -0.5 / 32768, -1.5 / 32768,
                  32767 / 32768, 32767.5 / 32768, 2.0, -2.0, np.nan, np.inf, -np.inf, 1e-30], dtype=np.float32)
    want = np.array([0, 32767, -32768, 16384, -16384, 0, 2, 2, 0, -2, 32767, 32767, 32767, -32768, 0, 32767, -32768, 0], np.int16)
    assert np.array_equal(sko.f32_to_s16(x), want)
    assert np.array_equal(np_oracle.f32_to_s16(x), want)
    allv = np.arange(-32768, 32768, dtype=np.int16)
    f = sko.s16_to_f32(allv)
    assert np.array_equal(bits(f), bits(np_oracle.s16_to_f32(allv)))
    assert np.array_equal(sko.f32_to_s16(f), allv)
    r = synth.uniform_pcm(8, 200000, over_range_frac=0.05)
    assert np.array_equal(sko.f32_to_s16(r), np_oracle.f32_to_s16(r))


# ---------------------------------------------------------------- mixer

def F(v, ch, n=10, uniq=True):
    return (np.full(n * ch, v, np.float32), ch, uniq)


def test_mixer_reference_scenarios():
    o, oc = sko.mix_sync([F(0.5, 2), F(0.3, 2)])                         # mixer.rs:1698-1701
    assert oc == 2 and o.size == 20 and np.all(np.abs(o - 0.8) < 1e-3)
    o, _ = sko.mix_sync([F(0.1, 2), F(0.2, 2), F(0.3, 2)])               # :1807-1809
    assert np.all(np.abs(o - 0.6) < 1e-3)
    o, _ = sko.mix_sync([F(0.5, 2), F(-0.3, 2)])                         # :1941-1943
    assert np.all(np.abs(o - 0.2) < 1e-3)
    o, oc = sko.mix_sync([F(0.75, 2)])                                   # single input pass-through :1981-1983
    assert oc == 2 and np.all(o == np.float32(0.75))
    o, oc = sko.mix_sync([F(0.5, 2), F(0.3, 1)])                         # upmix :1738
    assert oc == 2 and abs(o[0] - 0.8) < 1e-3
    o, oc = sko.mix_sync([F(0.25, 1)], max_channels_seen=2)              # sticky stereo :1752-1754
    assert oc == 2 and o.size == 20 and abs(o[0] - 0.25) < 1e-3 and abs(o[1] - 0.25) < 1e-3
    o = sko.mix_clocked([F(0.5, 2), F(0.3, 2)], 2, 10)                   # clocked :2047-2052
    assert o.size == 20 and np.all(np.abs(o - 0.8) < 1e-3)
    o = sko.mix_clocked([F(0.75, 2)], 2, 10)                             # missing input = silence :2097-2102
    assert np.all(np.abs(o - 0.75) < 1e-3)
    o, oc = sko.mix_sync([])
    assert o.size == 0


def test_mixer_base_selection_and_swap_remove_order():
    # all 4 frames have the output shape and are unique: base = last (idx 3); order after swap_remove = [0,1,2]
    fr = [F(0.1 * (i + 1), 2) for i in range(4)]
    assert sko.mix_plan(fr, 2, 20) == ([3, 0, 1, 2], True)
    # frame 1 is the only unique one -> base = 1, the last frame takes its slot: [0, 3, 2]
    fr = [F(0.1, 2, uniq=False), F(0.2, 2, uniq=True), F(0.3, 2, uniq=False), F(0.4, 2, uniq=False)]
    assert sko.mix_plan(fr, 2, 20) == ([1, 0, 3, 2], True)
    # no frame of output shape (all mono, stereo output): zero-initialised accumulator, pin order
    fr = [F(0.1, 1), F(0.2, 1)]
    assert sko.mix_plan(fr, 2, 20) == ([0, 1], False)
    # numpy restatement agrees
    for frames, oc, osz in [([F(0.1, 2), F(0.2, 2, uniq=False), F(0.3, 1)], 2, 20)]:
        assert sko.mix_plan(frames, oc, osz) == np_oracle.mix_order(frames, oc, osz)


def test_mixer_order_matters_at_full_scale_and_c_equals_numpy():
    rng = np.random.default_rng(0)
    g = [((rng.random(1920, dtype=np.float32) * 2 - 1), 2, True) for _ in range(64)]
    a = sko.mix_clocked(g, 2, 960)
    b = np_oracle.mix(g, 2, 1920)
    assert np.array_equal(bits(a), bits(b))
    rev = sko.mix_clocked(g[::-1], 2, 960)
    assert not np.array_equal(bits(a), bits(rev))     # the order is part of the contract
    # negative zero: base frame is NOT added to +0.0
    z = [(np.array([-0.0, -0.0], np.float32), 2, True)]
    assert np.signbit(sko.mix_clocked(z, 2, 1)).all()
    z2 = [(np.array([-0.0], np.float32), 1, True)]     # mono into stereo: no base -> 0.0 + -0.0 = +0.0
    assert not np.signbit(sko.mix_clocked(z2, 2, 1)).any()


def test_mixer_random_cross_check():
    rng = np.random.default_rng(2)
    for _ in range(200):
        n = int(rng.integers(1, 7))
        frames = []
        for _ in range(n):
            ch = int(rng.choice([1, 2, 3]))
            fr = int(rng.integers(1, 40))
            frames.append((rng.standard_normal(fr * ch).astype(np.float32), ch, bool(rng.random() < 0.7)))
        seen = int(rng.choice([0, 1, 2]))
        a, oc = sko.mix_sync(frames, seen)
        b, oc2 = np_oracle.mix_sync(frames, seen)
        assert oc == oc2 and np.array_equal(bits(a), bits(b))


# ---------------------------------------------------------------- resampler

def test_resampler_known_answers():
    # SURVEY Appendix B worked checks (control flow of rubato FastFixedIn)
    r = sko.FastFixedIn(48000, 16000, 960, 2)
    x = synth.tone_streams(1, 0, 1, 960, 2, 48000)[0]
    o1 = r.process(x)
    assert o1.size // 2 == 318 and r.last_index == -10.0
    assert o1[0] == 0.0 and o1[1] == 0.0           # first output reads zero history with frac 0
    assert r.process(x).size // 2 == 320 and r.last_index == -10.0
    # reference test resampler.rs:816-835: one 960-sample stereo packet, chunk_frames 960 -> remainder path with a
    # fresh FastFixedIn(chunk = 480) -> 237 frames = 474 samples, inside the asserted 480 +- 10
    n = sko.ResamplerNode(24000, 960, 0)
    n.push(48000, 2, np.full(960, 0.5, np.float32))
    assert n.out == []
    n.finish()
    assert len(n.out) == 1 and n.out[0]["samples"].size == 474 and n.out[0]["sample_rate"] == 24000 and n.out[0]["channels"] == 2
    assert abs(n.out[0]["samples"].size - 480) < 10
    # resampler.rs:886-906 buffering: 3 x 480-sample packets -> first output non-empty
    n = sko.ResamplerNode(24000, 960, 0)
    for _ in range(3):
        n.push(48000, 2, np.full(480, 0.5, np.float32))
    n.finish()
    assert len(n.out) >= 1 and n.out[0]["samples"].size > 0


def test_resampler_config_validation():
    with pytest.raises(ValueError, match="target_sample_rate must be greater than 0"):
        sko.ResamplerNode(0)                                                  # resampler.rs:917-922
    with pytest.raises(ValueError, match="chunk_frames must be greater than 0"):
        sko.ResamplerNode(48000, 0)
    with pytest.raises(ValueError, match="valid Opus frame size"):
        sko.ResamplerNode(48000, 960, 1000)
    for ofs in (0, 120, 240, 480, 960, 1920, 2880):
        sko.ResamplerNode(48000, 960, ofs)


@pytest.mark.parametrize("in_rate,out_rate,chunk,ch", [(44100, 48000, 882, 2), (48000, 16000, 960, 1), (16000, 48000, 320, 2),
                                                       (48000, 44100, 960, 2), (22050, 48000, 441, 3)])
def test_resampler_c_equals_numpy(in_rate, out_rate, chunk, ch):
    a = sko.FastFixedIn(in_rate, out_rate, chunk, ch)
    b = np_oracle.FastFixedIn(in_rate, out_rate, chunk, ch)
    for c in range(8):
        x = synth.tone_streams(4, c, 1, chunk, ch, in_rate)[0]
        ya, yb = a.process(x), b.process(x)
        assert ya.size == yb.size and np.array_equal(bits(ya), bits(yb))
        assert a.last_index == b.last_index


def test_resampler_node_reframing_timestamps_and_flush():
    # 44.1k mono -> 48k, default chunk 960 / output_frame_size 960, packets of 882 frames, first timestamp 1000 us
    n = sko.ResamplerNode(48000, 960, 960)
    total_in = 0
    for c in range(25):
        n.push(44100, 1, synth.tone_streams(6, c, 1, 882, 1, 44100)[0], timestamp_us=1000 if c == 0 else 5)
        total_in += 882
    n.finish()
    sizes = [p["samples"].size for p in n.out]
    assert all(s == 960 for s in sizes[:-1]) and 0 < sizes[-1] <= 960
    assert [p["sequence"] for p in n.out[:-1]] == list(range(len(n.out) - 1))
    assert n.out[0]["timestamp_us"] == 1000
    assert n.out[1]["timestamp_us"] == 1000 + sko.duration_us_for_frames(48000, 960)   # 20000 us
    assert all(p["duration_us"] == 20000 for p in n.out[:-1])
    assert n.out[-1]["duration_us"] == sko.duration_us_for_frames(48000, sizes[-1])
    # format change mid-stream is fatal (resampler.rs:253-279)
    n2 = sko.ResamplerNode(48000)
    n2.push(44100, 2, np.zeros(882 * 2, np.float32))
    with pytest.raises(RuntimeError, match="Audio format changed mid-stream: expected 44100Hz/2ch, got 48000Hz/2ch"):
        n2.push(48000, 2, np.zeros(960 * 2, np.float32))
    # equal rates: passthrough untouched (ofs 0) or re-framed only (resampler.rs:299-373)
    n3 = sko.ResamplerNode(48000, 960, 0)
    x = synth.uniform_pcm(1, 700)
    n3.push(48000, 1, x)
    assert len(n3.out) == 1 and np.array_equal(bits(n3.out[0]["samples"]), bits(x))
    n4 = sko.ResamplerNode(48000, 960, 480)
    n4.push(48000, 1, x)
    assert [p["samples"].size for p in n4.out] == [480]
    n4.finish()
    assert [p["samples"].size for p in n4.out] == [480, 220]
    assert np.array_equal(bits(np.concatenate([p["samples"] for p in n4.out])), bits(x))


# ---------------------------------------------------------------- golden vectors

def test_golden_vectors():
    """tests/golden/hotpath_v1.npz is produced by tests/golden/make_golden.py (C oracle, cross-checked with the numpy
    restatement at generation time). It pins the oracle against silent drift; the GPU suite re-uses it."""
    g = np.load(os.path.join(GOLDEN, "hotpath_v1.npz"))
    x = g["pcm"]
    assert np.array_equal(bits(sko.gain(x, g["gain"][0])), g["gain_out_bits"])
    assert np.array_equal(sko.gain_f32_to_s16(x, g["gain"][0]), g["gain_s16"])
    assert np.array_equal(sko.f32_to_s16(g["edge"]), g["edge_s16"])
    frames = [(g["mix_in"][i], 2, True) for i in range(g["mix_in"].shape[0])]
    assert np.array_equal(bits(sko.mix_clocked(frames, 2, 960)), g["mix_out_bits"])
    r = sko.FastFixedIn(44100, 48000, 882, 2)
    outs = np.concatenate([r.process(g["rs_in"][c]) for c in range(g["rs_in"].shape[0])])
    assert np.array_equal(bits(outs), g["rs_out_bits"]) and r.last_index == float(g["rs_last_index"][0])
    r = sko.FastFixedIn(48000, 16000, 960, 1)
    outs = np.concatenate([r.process(g["rs2_in"][c]) for c in range(g["rs2_in"].shape[0])])
    assert np.array_equal(bits(outs), g["rs2_out_bits"])


def test_reference_wav_fixture_chain_functional():
    """BASELINE config #1 restated: s16 stereo 48 kHz fixture -> f32 -> per-channel mono streams ->
    resample 48k->16k -> gain 2.0 -> s16. The expected bytes are committed (tests/golden/config1_*.npy); the WAV
    fixture itself is the reference's crates/nodes/testdata/audio/sample.wav decoded at generation time."""
    pcm = np.load(os.path.join(GOLDEN, "config1_input_s16.npy"))
    want = np.load(os.path.join(GOLDEN, "config1_output_s16.npy"))
    assert pcm.shape == (4800, 2)
    f = sko.s16_to_f32(pcm.reshape(-1)).reshape(4800, 2)
    outs = []
    for ch in range(2):
        n = sko.ResamplerNode(16000, 960, 960)
        for c in range(5):
            n.push(48000, 1, f[c * 960:(c + 1) * 960, ch])
        n.finish()
        y = np.concatenate([p["samples"] for p in n.out])
        outs.append(sko.gain_f32_to_s16(y, 2.0))
    assert np.array_equal(np.stack(outs), want)
