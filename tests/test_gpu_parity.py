"""GPU parity tests: every CUDA kernel, called THROUGH the C ABI (ctypes -> libskgpu.so), against the
oracle (oracle/libsk_oracle.so) on identical seeded inputs.

Bars (BASELINE.json north_star): s16 conversion / clipping bit-exact; f32 gain / mix bit-exact (they
are single IEEE operations in a defined order); f32 resample within 2e-6 max-abs -- in practice the
kernels reproduce the oracle bit for bit, the tests assert the tolerance AND report exactness.
"""
import numpy as np
import pytest

from oracle import np_oracle, sko
from streamkit_b200 import chain, lib as L, synth
from tests import chain_ref

pytestmark = pytest.mark.gpu

RESAMPLE_TOL = 2e-6  # north_star: f32 resample within 2e-6 max-abs


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ctx():
    c = L.Context(device=0, max_streams=4096, max_channels=8, fifo_frames=2048)
    yield c
    c.close()


def _al(x, a=256):
    return (x + a - 1) // a * a


# ------------------------------------------------------------------ K1 / K2: gain + format conversion

EDGE = np.array([0.0, -0.0, 1.0, -1.0, 0.999969482421875, 1.0000001, -1.0000001, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768,
                 -0.5 / 32768, -1.5 / 32768, 32766.5 / 32768, 32767.5 / 32768, -32768.5 / 32768, 3.9, -3.9, 1e-40, -1e-40,
                 np.inf, -np.inf, np.nan, 1e30, -1e30, 1.17549435e-38], dtype=np.float32)


def _run_convert(ctx, mode, frames, gains, offsets_in=None):
    """frames: list of 1-D arrays (f32 or s16); returns list of outputs"""
    in_b = 2 if mode == L.CVT_S16_TO_F32 else 4
    out_b = 2 if mode == L.CVT_F32_TO_S16 else 4
    segs = np.zeros(len(frames), dtype=L.SEG_DT)
    off = 0
    for i, f in enumerate(frames):
        pad = 0 if offsets_in is None else offsets_in[i]
        segs[i]["in_off"] = off + pad
        segs[i]["n_samples"] = f.size
        segs[i]["gain_idx"] = i if gains is not None else L.SKGPU_NO_GAIN
        off = _al(off + pad + f.size * in_b)
    in_bytes = off
    for i, f in enumerate(frames):
        pad = 0 if offsets_in is None else (offsets_in[i] // in_b) * out_b
        segs[i]["out_off"] = off + pad
        off = _al(off + pad + f.size * out_b)
    plan = L.Plan(ctx, max(off, 256))
    try:
        if gains is not None:
            plan.set_gains(gains)
        plan.add_convert(mode, segs)
        plan.set_io(0, in_bytes, in_bytes, off - in_bytes)
        plan.finalize()
        host_in = np.zeros(in_bytes, dtype=np.uint8)
        for s, f in zip(segs, frames):
            b = np.ascontiguousarray(f).view(np.uint8)
            host_in[int(s["in_off"]): int(s["in_off"]) + b.size] = b
        host_out = np.zeros(off - in_bytes, dtype=np.uint8)
        plan.submit(host_in, host_out)
        plan.wait()
        outs = []
        for s, f in zip(segs, frames):
            o = int(s["out_off"]) - in_bytes
            outs.append(host_out[o: o + f.size * out_b].view(np.int16 if out_b == 2 else np.float32).copy())
        return outs
    finally:
        plan.destroy()


@pytest.mark.parametrize("n", [0, 1, 3, 4, 7, 8, 100, 1919, 1920, 1921, 2048, 2049, 5000, 7680])
def test_gain_f32_bit_exact(ctx, n):
    x = synth.uniform_pcm(n + 11, n)
    g = np.float32(1.7)
    (y,) = _run_convert(ctx, L.CVT_F32_TO_F32, [x], np.array([g], np.float32))
    assert np.array_equal(bits(y), bits(sko.gain(x, g)))


def test_gain_reference_unit_test_constants(ctx):
    # gain.rs:233-435: 0.5 * 2.0 -> 1.0 ; 0.5x on 3 packets ; zero gain -> exact 0.0 ; 4.0 * 0.5 -> 2.0
    x = np.full(100, 0.5, np.float32)
    (y,) = _run_convert(ctx, L.CVT_F32_TO_F32, [x], np.array([2.0], np.float32))
    assert y.size == 100 and np.all(np.abs(y - 1.0) < 1e-3)
    frames = [np.full(20, v, np.float32) for v in (0.2, 0.4, 0.6)]
    outs = _run_convert(ctx, L.CVT_F32_TO_F32, frames, np.array([0.5, 0.5, 0.5], np.float32))
    for o, v in zip(outs, (0.2, 0.4, 0.6)):
        assert np.all(np.abs(o - v * 0.5) < 1e-3)
    (z,) = _run_convert(ctx, L.CVT_F32_TO_F32, [np.ones(20, np.float32)], np.array([0.0], np.float32))
    assert np.all(z == 0.0)
    (m,) = _run_convert(ctx, L.CVT_F32_TO_F32, [np.full(20, 0.5, np.float32)], np.array([4.0], np.float32))
    assert np.all(np.abs(m - 2.0) < 1e-3)


def test_gain_many_sessions_ragged(ctx):
    rng = np.random.default_rng(3)
    frames = [synth.uniform_pcm(100 + i, int(n)) for i, n in enumerate(rng.integers(1, 4000, size=97))]
    gains = synth.gains(5, len(frames))
    outs = _run_convert(ctx, L.CVT_F32_TO_F32, frames, gains)
    for x, g, y in zip(frames, gains, outs):
        assert np.array_equal(bits(y), bits(sko.gain(x, g)))


def test_gain_unaligned_offsets(ctx):
    frames = [synth.uniform_pcm(200 + i, 501 + i) for i in range(8)]
    gains = synth.gains(6, len(frames))
    outs = _run_convert(ctx, L.CVT_F32_TO_F32, frames, gains, offsets_in=[4 * (i % 4) for i in range(8)])
    for x, g, y in zip(frames, gains, outs):
        assert np.array_equal(bits(y), bits(sko.gain(x, g)))


def test_f32_to_s16_edge_vectors_bit_exact(ctx):
    x = np.concatenate([EDGE, np.arange(-40000, 40000, 1, dtype=np.float32) / np.float32(32768.0) + np.float32(0.5 / 32768)])
    (s,) = _run_convert(ctx, L.CVT_F32_TO_S16, [x], None)
    assert np.array_equal(s, sko.f32_to_s16(x))
    assert np.array_equal(s, np_oracle.f32_to_s16(x))


@pytest.mark.parametrize("n", [1, 7, 8, 9, 1920, 4097])
def test_gain_clip_s16_bit_exact(ctx, n):
    x = synth.uniform_pcm(n, n, over_range_frac=0.05)
    g = np.float32(2.3)
    (s,) = _run_convert(ctx, L.CVT_F32_TO_S16, [x], np.array([g], np.float32))
    assert np.array_equal(s, sko.gain_f32_to_s16(x, g))


def test_s16_to_f32_all_values_and_roundtrip(ctx):
    allv = np.arange(-32768, 32768, dtype=np.int16)
    (f,) = _run_convert(ctx, L.CVT_S16_TO_F32, [allv], None)
    assert np.array_equal(bits(f), bits(sko.s16_to_f32(allv)))
    (back,) = _run_convert(ctx, L.CVT_F32_TO_S16, [f], None)
    assert np.array_equal(back, allv)  # s16 -> f32 -> s16 is the identity


def test_config2_shape_4096_sessions_property(ctx):
    """BASELINE config #2 at full size: 4096 sessions x 1920 samples, fused gain -> s16.
    Checked by oracle on a sample of sessions and by a checksum of all of them."""
    S, N = 4096, 1920
    x = synth.uniform_pcm(42, S * N).reshape(S, N)
    gains = synth.gains(43, S)
    outs = _run_convert(ctx, L.CVT_F32_TO_S16, [x[i] for i in range(S)], gains)
    got = np.stack(outs)
    want = np.empty_like(got)
    for i in range(S):
        want[i] = sko.gain_f32_to_s16(x[i], gains[i])
    assert np.array_equal(got, want)


# ------------------------------------------------------------------ K3: mixer

def _run_mix(ctx, groups_frames, out_shapes, in_gains=None, master=None, s16=False, present=None):
    """groups_frames: list of groups, each a list of (samples, channels, unique); out_shapes: [(oc, out_frames)]"""
    n_in = sum(len(g) for g in groups_frames)
    inputs = np.zeros(n_in, dtype=L.MIX_INPUT_DT)
    groups = np.zeros(len(groups_frames), dtype=L.MIX_GROUP_DT)
    off = 0
    blobs = []
    k = 0
    for gi, g in enumerate(groups_frames):
        groups[gi]["first_input"] = k
        groups[gi]["n_inputs"] = len(g)
        for (smp, ch, uniq) in g:
            smp = np.ascontiguousarray(smp, dtype=np.float32)
            inputs[k]["in_off"] = off
            inputs[k]["n_frames"] = smp.size // ch
            inputs[k]["channels"] = ch
            inputs[k]["flags"] = L.MIX_IN_UNIQUE if uniq else 0
            inputs[k]["gain_idx"] = k if in_gains is not None else L.SKGPU_NO_GAIN
            blobs.append((off, smp))
            off = _al(off + smp.size * 4, 16)
            k += 1
    in_bytes = _al(off)
    off = in_bytes
    ob = 2 if s16 else 4
    for gi, (oc, of) in enumerate(out_shapes):
        groups[gi]["out_off"] = off
        groups[gi]["out_frames"] = of
        groups[gi]["out_channels"] = oc
        groups[gi]["flags"] = L.MIX_OUT_S16 if s16 else 0
        groups[gi]["gain_idx"] = (n_in + gi) if master is not None else L.SKGPU_NO_GAIN
        off = _al(off + oc * of * ob, 16)
    plan = L.Plan(ctx, max(_al(off), 256))
    try:
        gt = []
        if in_gains is not None:
            gt = list(in_gains)
        elif master is not None:
            gt = [1.0] * n_in
        if master is not None:
            gt += list(master)
        if gt:
            plan.set_gains(np.array(gt, np.float32))
        op = plan.add_mix(groups, inputs)
        if present is not None:
            plan.set_present(op, present)
        plan.set_io(0, in_bytes, in_bytes, max(off - in_bytes, 0))
        plan.finalize()
        host_in = np.zeros(in_bytes, dtype=np.uint8)
        for o, smp in blobs:
            host_in[o:o + smp.size * 4] = smp.view(np.uint8)
        host_out = np.zeros(max(off - in_bytes, 1), dtype=np.uint8)
        plan.submit(host_in, host_out)
        plan.wait()
        res = []
        for gi, (oc, of) in enumerate(out_shapes):
            o = int(groups[gi]["out_off"]) - in_bytes
            res.append(host_out[o:o + oc * of * ob].view(np.int16 if s16 else np.float32).copy())
        return res
    finally:
        plan.destroy()


def test_mixer_reference_scenarios(ctx):
    f = lambda v, ch, n=10: (np.full(n * ch, v, np.float32), ch, True)
    # mixer.rs:1698-1701  0.5 + 0.3 ~ 0.8
    (o,) = _run_mix(ctx, [[f(0.5, 2), f(0.3, 2)]], [(2, 10)])
    assert o.size == 20 and np.all(np.abs(o - 0.8) < 1e-3)
    # three inputs 0.1 + 0.2 + 0.3 ~ 0.6 ; negative 0.5 + (-0.3) ~ 0.2
    (o,) = _run_mix(ctx, [[f(0.1, 2), f(0.2, 2), f(0.3, 2)]], [(2, 10)])
    assert np.all(np.abs(o - 0.6) < 1e-3)
    (o,) = _run_mix(ctx, [[f(0.5, 2), f(-0.3, 2)]], [(2, 10)])
    assert np.all(np.abs(o - 0.2) < 1e-3)
    # single input is a pass-through (mixer.rs:1947-1984)
    (o,) = _run_mix(ctx, [[f(0.75, 2)]], [(2, 10)])
    assert np.all(o == np.float32(0.75))
    # stereo + mono -> stereo upmix (mixer.rs:1728-1738), then sticky stereo with mono only (:1741-1754)
    (o,) = _run_mix(ctx, [[f(0.5, 2), f(0.3, 1)]], [(2, 10)])
    assert abs(o[0] - 0.8) < 1e-3 and abs(o[1] - 0.8) < 1e-3
    (o,) = _run_mix(ctx, [[f(0.25, 1)]], [(2, 10)])
    assert o.size == 20 and np.all(np.abs(o - 0.25) < 1e-3)
    # clocked: missing input is silence (mixer.rs:2097-2102)
    (o,) = _run_mix(ctx, [[f(0.75, 2), f(0.1, 2)]], [(2, 10)], present=[1, 0])
    assert np.all(np.abs(o - 0.75) < 1e-3)


def _rand_group(rng, n_inputs, oc, frames, full_scale, ragged=False, chans=(1, 2)):
    g = []
    for _ in range(n_inputs):
        ch = int(rng.choice(chans))
        fr = frames if not ragged else int(rng.integers(max(1, frames - 50), frames + 50))
        amp = 1.0 if full_scale else 0.125
        smp = ((rng.random(fr * ch, dtype=np.float32) * 2 - 1) * np.float32(amp)).astype(np.float32)
        if rng.random() < 0.1:
            smp[rng.integers(0, smp.size)] = -0.0
        g.append((smp, ch, bool(rng.random() < 0.8)))
    return g


@pytest.mark.parametrize("full_scale", [False, True])
def test_mixer_64_inputs_bit_exact_order(ctx, full_scale):
    """config #3 shape: 64 stereo inputs per group. With full-scale inputs the partial sums reach |64| where one
    ulp is 7.6e-6 > 2e-6, so only the reference's summation ORDER gives the right bits."""
    rng = np.random.default_rng(11 + full_scale)
    groups = [_rand_group(rng, 64, 2, 960, full_scale, chans=(2,)) for _ in range(6)]
    outs = _run_mix(ctx, groups, [(2, 960)] * 6)
    for g, o in zip(groups, outs):
        want = sko.mix_clocked(g, 2, 960)
        assert np.array_equal(bits(o), bits(want))
        assert np.array_equal(bits(want), bits(np_oracle.mix(g, 2, 1920)))


def test_mixer_mixed_channels_ragged_sync_mode(ctx):
    rng = np.random.default_rng(5)
    groups, shapes, wants = [], [], []
    for gi in range(40):
        n = int(rng.integers(1, 9))
        g = _rand_group(rng, n, 2, int(rng.integers(5, 700)), False, ragged=True, chans=(1, 2, 3) if gi % 5 == 0 else (1, 2))
        want, oc = sko.mix_sync(g, max_channels_seen=int(rng.choice([0, 1, 2])))
        groups.append(g)
        shapes.append((oc, want.size // oc))
        wants.append(want)
    outs = _run_mix(ctx, groups, shapes)
    for o, w in zip(outs, wants):
        assert np.array_equal(bits(o), bits(w))


def test_mixer_stereo_to_mono_and_generic(ctx):
    rng = np.random.default_rng(9)
    g1 = _rand_group(rng, 5, 1, 333, False, chans=(1, 2))
    w1 = sko.mix_clocked(g1, 1, 333)
    g2 = _rand_group(rng, 4, 3, 200, False, chans=(1, 2, 3, 4))
    w2 = sko.mix_clocked(g2, 3, 200)
    o1, o2 = _run_mix(ctx, [g1, g2], [(1, 333), (3, 200)])
    assert np.array_equal(bits(o1), bits(w1))
    assert np.array_equal(bits(o2), bits(w2))


def test_mixer_gain_clip_s16_epilogue(ctx):
    rng = np.random.default_rng(21)
    groups = [_rand_group(rng, 8, 2, 960, True, chans=(2,)) for _ in range(5)]
    in_g = synth.gains(1, 40, 0.0, 2.0)
    master = synth.gains(2, 5, 0.1, 1.5)
    outs = _run_mix(ctx, groups, [(2, 960)] * 5, in_gains=in_g, master=master, s16=True)
    for gi, (g, o) in enumerate(zip(groups, outs)):
        scaled = [(sko.gain(s, in_g[gi * 8 + j]), ch, u) for j, (s, ch, u) in enumerate(g)]
        want = sko.gain_f32_to_s16(sko.mix_clocked(scaled, 2, 960), master[gi])
        assert np.array_equal(o, want)


def test_mixer_presence_changes_base_and_order(ctx):
    rng = np.random.default_rng(33)
    g = _rand_group(rng, 12, 2, 480, True, chans=(2,))
    for trial in range(6):
        present = (rng.random(12) < 0.6).astype(np.uint8)
        (o,) = _run_mix(ctx, [g], [(2, 480)], present=present)
        sel = [f for f, p in zip(g, present) if p]
        want = sko.mix_clocked(sel, 2, 480)
        assert np.array_equal(bits(o), bits(want))


# ------------------------------------------------------------------ K4: resampler

def _run_resample_stream(ctx, in_rate, out_rate, chunk, channels, n_chunks, seed, n_streams=3):
    """streams x chunks through the GPU with persistent state; returns per-stream list of per-chunk outputs"""
    slots = [ctx.stream_open(in_rate, out_rate, chunk, channels) for _ in range(n_streams)]
    cap = L.Context.max_out_frames(in_rate, out_rate, chunk, channels)
    in_stride = _al(chunk * channels * 4, 16)
    out_stride = _al(cap * channels * 4, 16)
    in_bytes = _al(n_streams * in_stride)
    res_off = in_bytes
    out_off = _al(res_off + 8 * n_streams)
    total = _al(out_off + n_streams * out_stride)
    plan = L.Plan(ctx, total)
    items = np.zeros(n_streams, dtype=L.RS_ITEM_DT)
    items["in_off"] = np.arange(n_streams) * in_stride
    items["out_off"] = out_off + np.arange(n_streams) * out_stride
    items["slot"] = slots
    items["out_cap_frames"] = cap
    plan.add_resample(items, res_off)
    plan.set_io(0, in_bytes, res_off, total - res_off)
    plan.finalize()
    outs = [[] for _ in range(n_streams)]
    try:
        for c in range(n_chunks):
            x = synth.tone_streams(seed, c, n_streams, chunk, channels, in_rate)
            host_in = np.zeros(in_bytes, np.uint8)
            for s in range(n_streams):
                host_in[s * in_stride: s * in_stride + chunk * channels * 4] = x[s].view(np.uint8)
            host_out = np.zeros(total - res_off, np.uint8)
            plan.submit(host_in, host_out, L.SUBMIT_GRAPH if c % 2 else 0)
            plan.wait()
            res = host_out[: 8 * n_streams].view(L.RS_RESULT_DT)
            for s in range(n_streams):
                assert res[s]["status"] == 0
                n = int(res[s]["out_frames"])
                o = (out_off - res_off) + s * out_stride
                outs[s].append(host_out[o: o + n * channels * 4].view(np.float32).copy())
        states = [ctx.stream_state(sl, channels) for sl in slots]
    finally:
        plan.destroy()
        for sl in slots:
            ctx.stream_close(sl)
    return outs, states


@pytest.mark.parametrize("in_rate,out_rate,chunk,channels", [
    (44100, 48000, 882, 2), (48000, 16000, 960, 2), (48000, 16000, 960, 1), (44100, 48000, 882, 1),
    (48000, 24000, 480, 2), (16000, 48000, 320, 2), (8000, 48000, 160, 1), (48000, 44100, 960, 2),
    (44100, 16000, 441, 1), (22050, 48000, 441, 3), (48000, 8000, 960, 2), (32000, 48000, 7, 2), (48000, 16000, 20, 2),
])
def test_resampler_streaming_parity(ctx, in_rate, out_rate, chunk, channels):
    n_chunks, n_streams = 12, 3
    outs, states = _run_resample_stream(ctx, in_rate, out_rate, chunk, channels, n_chunks, seed=chunk + channels, n_streams=n_streams)
    exact = True
    for s in range(n_streams):
        r = sko.FastFixedIn(in_rate, out_rate, chunk, channels)
        for c in range(n_chunks):
            x = synth.tone_streams(chunk + channels, c, n_streams, chunk, channels, in_rate)[s]
            want = r.process(x)
            got = outs[s][c]
            assert got.size == want.size, (s, c, got.size, want.size)  # output COUNT is exact, not +-1
            if got.size:
                assert np.max(np.abs(got - want)) <= RESAMPLE_TOL
                exact = exact and np.array_equal(bits(got), bits(want))
        li, hist, _, _ = states[s]
        assert li == r.last_index                      # f64 phase state bit-identical after 12 chunks
        assert np.array_equal(bits(hist), bits(r.history()))
    assert exact, "within tolerance but not bit-exact (unexpected: arithmetic is IEEE, uncontracted)"


def test_resampler_known_answer_lengths(ctx):
    # 48k->16k chunk 960: 318 frames first call, then 320
    outs, _ = _run_resample_stream(ctx, 48000, 16000, 960, 2, 3, seed=1, n_streams=1)
    assert [o.size // 2 for o in outs[0]] == [318, 320, 320]
    # reference test resampler.rs:816-835 (remainder path): 48k->24k, fresh resampler with chunk 480 -> 237 frames = 474 samples
    outs, _ = _run_resample_stream(ctx, 48000, 24000, 480, 2, 1, seed=2, n_streams=1)
    assert outs[0][0].size == 474 and abs(outs[0][0].size - 480) < 10


def test_resampler_long_run_count_exact(ctx):
    """500 chunks of 44.1k->48k: accumulated f64 phase must track the sequential recurrence exactly
    (a closed-form idx0 + k*t would drift and eventually flip an output count)."""
    outs, states = _run_resample_stream(ctx, 44100, 48000, 882, 1, 500, seed=77, n_streams=1)
    r = sko.FastFixedIn(44100, 48000, 882, 1)
    for c in range(500):
        want = r.process(synth.tone_streams(77, c, 1, 882, 1, 44100)[0])
        assert outs[0][c].size == want.size
        assert np.array_equal(bits(outs[0][c]), bits(want))
    assert states[0][0] == r.last_index


def test_resampler_reset_gives_fresh_state(ctx):
    slot = ctx.stream_open(48000, 16000, 960, 2)
    li, hist, _, _ = ctx.stream_state(slot, 2)
    assert li == -4.0 and np.all(hist == 0)
    ctx.stream_close(slot)


def test_resampler_many_streams_16384_property(ctx):
    """config #4 scale-down on shared ctx is limited to 4096 slots; full 16384 runs in a dedicated context.
    Property: identical streams produce identical outputs and every stream matches the oracle's count."""
    c2 = L.Context(device=0, max_streams=16384, max_channels=2, fifo_frames=0)
    try:
        S, N, C = 16384, 882, 2
        slots = c2.stream_open_many(44100, 48000, N, C, S)
        cap = L.Context.max_out_frames(44100, 48000, N, C)
        in_stride, out_stride = N * C * 4, _al(cap * C * 4, 16)
        in_bytes = _al(S * in_stride)
        res_off = in_bytes
        out_off = _al(res_off + 8 * S)
        total = _al(out_off + S * out_stride)
        plan = L.Plan(c2, total)
        items = np.zeros(S, dtype=L.RS_ITEM_DT)
        items["in_off"] = np.arange(S, dtype=np.uint64) * in_stride
        items["out_off"] = out_off + np.arange(S, dtype=np.uint64) * out_stride
        items["slot"] = slots
        items["out_cap_frames"] = cap
        plan.add_resample(items, res_off)
        plan.set_io(0, in_bytes, res_off, total - res_off)
        plan.finalize()
        base = synth.tone_streams(5, 0, 64, N, C, 44100)          # 64 distinct streams tiled 256x
        x = np.tile(base, (S // 64, 1))
        host_in = np.zeros(in_bytes, np.uint8)
        host_in[: S * in_stride] = x.reshape(-1).view(np.uint8)
        host_out = np.zeros(total - res_off, np.uint8)
        r = [sko.FastFixedIn(44100, 48000, N, C) for _ in range(64)]
        for tick in range(3):
            plan.submit(host_in, host_out)
            plan.wait()
            res = host_out[: 8 * S].view(L.RS_RESULT_DT)
            want = [rr.process(base[i]) for i, rr in enumerate(r)]
            assert np.all(res["status"] == 0)
            for i in range(64):
                assert np.all(res["out_frames"][i::64] == want[i].size // C)
            o0 = out_off - res_off
            outs = host_out[o0: o0 + S * out_stride].reshape(S, out_stride)
            for i in (0, 17, 63):
                n = want[i].size
                blk = outs[i::64, : n * 4].copy().view(np.float32)
                assert np.array_equal(blk.view(np.uint32), np.tile(bits(want[i]), (S // 64, 1)))
        plan.destroy()
    finally:
        c2.close()


# ------------------------------------------------------------------ full chain (config #5)

@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("k_inputs,channels,in_rate", [(1, 2, 44100), (2, 2, 44100), (3, 1, 44100), (2, 2, 32000), (6, 2, 44100)])
def test_full_chain_bit_exact(k_inputs, channels, in_rate, fused):
    """fused k_chain (lagged recompute from the double-banked input arena) and the unfused ops
    (k_resample -> device ring -> k_mix) must both reproduce the oracle's s16 bytes."""
    S, T = 12, 7
    got = chain.run_chain_gpu(S, k_inputs, T, seed=3, in_rate=in_rate, channels=channels, fused=fused)
    want = chain_ref.run_chain_oracle(S, k_inputs, T, seed=3, in_rate=in_rate, channels=channels)
    for t in range(T):
        assert np.array_equal(got[t], want[t]), f"tick {t}: {(got[t] != want[t]).sum()} s16 samples differ"
    assert any(np.any(g != 0) for g in got)


@pytest.mark.parametrize("in_rate,chunk,out_frames,channels,k_inputs", [
    (44100, 441, 480, 2, 2), (44100, 1764, 1920, 2, 2), (44100, 2646, 2880, 2, 3), (22050, 441, 960, 1, 2),
    (16000, 320, 960, 2, 2), (8000, 160, 960, 2, 1), (11025, 441, 1920, 2, 2), (44100, 2646, 2880, 1, 2), (96000, 1920, 960, 2, 2),
    (44100, 882, 960, 2, 33),
])
def test_fused_chain_other_packet_sizes_and_ratios(in_rate, chunk, out_frames, channels, k_inputs):
    """every valid output_frame_size with more than one kernel iteration per packet (F > 1024), up- and down-sampling
    ratios whose programs contain slow (negative-position) run segments, and a session with more inputs than fit one
    staging batch: s16 bytes identical to the oracle."""
    S, T = 5, 6
    got = chain.run_chain_gpu(S, k_inputs, T, seed=21, in_rate=in_rate, channels=channels, chunk_frames=chunk, out_frames=out_frames)
    want = chain_ref.run_chain_oracle(S, k_inputs, T, seed=21, in_rate=in_rate, channels=channels, chunk_frames=chunk, out_frames=out_frames)
    for t in range(T):
        assert np.array_equal(got[t], want[t]), f"tick {t}: {(got[t] != want[t]).sum()} s16 samples differ"
    assert any(np.any(g != 0) for g in got)


@pytest.mark.parametrize("channels,k_inputs", [(1, 1), (2, 1), (2, 2)])
def test_fused_chain_config0_48k_to_16k_with_960_frame_packets(channels, k_inputs):
    """BASELINE configs[0] / samples/pipelines/oneshot/speech_to_text.yml:14-18: audio::resampler 48k -> 16k, chunk_frames 960,
    output_frame_size 960 (one packet per THREE chunks). Fused, this runs at the packet cadence: one 60 ms tick carries the
    stream's three 20 ms chunks as one 2,880-frame chunk. t = 3.0 is exact, so every position of the f64 recurrence is an
    integer whichever way the input is cut: the s16 bytes equal the reference-shaped nodes that process 960-frame chunks."""
    S, T = 4, 8
    got = chain.run_chain_gpu(S, k_inputs, T, seed=41, in_rate=48000, channels=channels, chunk_frames=2880, out_frames=960, out_rate=16000)
    want = chain_ref.run_chain_oracle(S, k_inputs, T, seed=41, in_rate=48000, channels=channels, chunk_frames=2880, out_frames=960, out_rate=16000,
                                      node_chunk_frames=960)
    for t in range(T):
        assert np.array_equal(got[t], want[t]), f"tick {t}: {(got[t] != want[t]).sum()} s16 samples differ"
    assert any(np.any(g != 0) for g in got[1:])


def test_fused_chain_long_run_stays_bit_exact():
    """400 ticks (8 s of audio): the f64 phase state, the carry bookkeeping and the alternating side records of the
    fused path must track the oracle tick for tick (a drifting phase would flip a packet boundary sooner or later)."""
    S, K, T = 2, 2, 400
    got = chain.run_chain_gpu(S, K, T, seed=31)
    want = chain_ref.run_chain_oracle(S, K, T, seed=31)
    bad = [t for t in range(T) if not np.array_equal(got[t], want[t])]
    assert not bad, f"first differing ticks: {bad[:5]}"


def test_full_chain_graph_equals_stream_launch():
    a = chain.run_chain_gpu(8, 2, 6, seed=9, graph=False)
    b = chain.run_chain_gpu(8, 2, 6, seed=9, graph=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_fused_chain_state_and_results():
    """carry / emission bookkeeping of the fused kernel: 44.1k->48k, first chunk yields 954 < 960 frames (no packet),
    afterwards one packet per tick with 954 carried frames recomputed from the previous bank."""
    ct = chain.ChainTick(4, 2, seed=5)
    try:
        for t in range(4):
            ct.tick(synth.tone_streams(5, t, ct.n_streams, ct.chunk, 2, 44100))
            res = ct.results()
            assert np.all(res["status"] == 0)
            assert np.all(res["emitted"] == (0 if t == 0 else 1))
        assert ct.plan.tick_count() == 4
    finally:
        ct.close()


def test_fused_chain_absent_stream_protocol():
    """A stream that delivers nothing in a tick is marked absent and its previous chunk is repeated by the host so
    the other bank stays valid (include/skgpu_batch.h, chain protocol). The oracle session simply gets no packet."""
    S, K, T = 6, 2, 8
    ct = chain.ChainTick(S, K, seed=11)
    rng = np.random.default_rng(4)
    import collections
    nodes = [sko.ResamplerNode(48000, chunk_frames=ct.chunk, output_frame_size=960) for _ in range(S * K)]
    queues = [collections.deque() for _ in range(S * K)]
    last = np.zeros((S * K, ct.chunk * 2), np.float32)
    n_sent = np.zeros(S * K, dtype=int)
    try:
        for t in range(T):
            present = (rng.random(S * K) < 0.75).astype(np.uint8) if t >= 2 else np.ones(S * K, np.uint8)
            x = np.empty_like(last)
            for s in range(S * K):
                if present[s]:
                    x[s] = synth.tone_streams(11 + s, n_sent[s], 1, ct.chunk, 2, 44100)[0]
                    n_sent[s] += 1
                    last[s] = x[s]
                else:
                    x[s] = last[s]                      # host repeats the previous chunk bytes
            ct.plan.set_present(ct.op_chain, present)
            got = ct.tick(x)
            want = np.zeros_like(got)
            for s in range(S * K):
                if present[s]:
                    nodes[s].out.clear()
                    nodes[s].push(44100, 2, x[s])
                    for pkt in nodes[s].out:
                        queues[s].append(sko.gain(pkt["samples"], float(ct.in_gains[s])))
            for g in range(S):
                frames = []
                for i in range(K):
                    q = queues[g * K + i]
                    if q:
                        frames.append((q.popleft(), 2, True))
                want[g] = sko.gain_f32_to_s16(sko.mix_clocked(frames, 2, 960), float(ct.master_gains[g]))
            assert np.array_equal(got, want), f"tick {t}"
            assert np.all(ct.results()["status"] == 0)
    finally:
        ct.close()


def test_chain_rejects_ineligible_streams(ctx):
    """48k -> 16k with chunk 960 yields 320 frames per chunk: a 960-frame packet would span 3 chunks -> unfused ops only"""
    slot = ctx.stream_open(48000, 16000, 960, 2)
    plan = L.Plan(ctx, 1 << 20)
    try:
        plan.set_io(0, 960 * 8, 65536, 4096)
        plan.set_banks(32768)
        cin = np.zeros(1, dtype=L.CHAIN_INPUT_DT)
        cin["slot"] = slot
        cin["gain_idx"] = L.SKGPU_NO_GAIN
        cg = np.zeros(1, dtype=L.CHAIN_GROUP_DT)
        cg["out_off"] = 65536
        cg["n_inputs"] = 1
        cg["gain_idx"] = L.SKGPU_NO_GAIN
        cg["out_channels"] = 2
        with pytest.raises(L.SkgpuError) as e:
            plan.add_chain(cg, cin, 960, 131072)
        assert "use the unfused ops" in e.value.msg
    finally:
        plan.destroy()
        ctx.stream_close(slot)


def test_chain_op_of_plain_inputs_rejects_another_input_kind_on_update(ctx):
    """an op created from resampled f32 inputs of the output's channel count runs the instantiation without input-kind code;
    a later table update must not smuggle a bypass / s16 / other-channel stream into it"""
    a = ctx.stream_open(44100, 48000, 882, 2)
    a2 = ctx.stream_open(44100, 48000, 882, 2)
    b = ctx.stream_open(48000, 48000, 960, 2)          # rate-equal: bypass
    c = ctx.stream_open(44100, 48000, 882, 2, L.STREAM_S16)
    plan = L.Plan(ctx, 1 << 20)
    try:
        plan.set_io(0, 32768, 200000, 8192)
        plan.set_banks(65536)
        cin = np.zeros(2, dtype=L.CHAIN_INPUT_DT)
        cin["slot"] = [a, a2]
        cin["in_off"] = [0, 8192]
        cin["gain_idx"] = L.SKGPU_NO_GAIN
        cg = np.zeros(1, dtype=L.CHAIN_GROUP_DT)
        cg["out_off"], cg["n_inputs"], cg["gain_idx"], cg["out_channels"] = 200000, 2, L.SKGPU_NO_GAIN, 2
        op = plan.add_chain(cg, cin, 960, 150000)
        plan.finalize()
        for other in (b, c):
            cin["slot"] = [a, other]
            with pytest.raises(L.SkgpuError) as e:
                plan.update_chain(op, cg, cin)
            assert "another kind" in e.value.msg, e.value.msg
        cin["slot"] = [a2, a]                              # the same kind in another order is fine
        plan.update_chain(op, cg, cin)
    finally:
        plan.destroy()
        for s_ in (a, a2, b, c):
            ctx.stream_close(s_)


def test_smoke_entry():
    import __graft_entry__ as g

    g.smoke()


# ------------------------------------------------------------------ error behaviour at the boundary

def test_invalid_configs_are_rejected(ctx):
    with pytest.raises(L.SkgpuError) as e:
        ctx.stream_open(48000, 0, 960, 2)
    assert "target_sample_rate must be greater than 0" in e.value.msg  # resampler.rs:82-86
    with pytest.raises(L.SkgpuError) as e:
        ctx.stream_open(48000, 16000, 0, 2)
    assert "chunk_frames must be greater than 0" in e.value.msg        # resampler.rs:88-92
    with pytest.raises(L.SkgpuError):
        ctx.stream_open(48000, 16000, 960, 9)
    plan = L.Plan(ctx, 4096)
    segs = np.zeros(1, dtype=L.SEG_DT)
    segs[0]["in_off"] = 4000
    segs[0]["n_samples"] = 1000
    segs[0]["gain_idx"] = L.SKGPU_NO_GAIN
    with pytest.raises(L.SkgpuError) as e:
        plan.add_convert(L.CVT_F32_TO_F32, segs)
    assert "outside" in e.value.msg
    plan.destroy()


# ------------------------------------------------------------------ sliced ticks

@pytest.mark.parametrize("n_slices", [1, 3, 8, 64])
def test_sliced_tick_equals_whole_tick(n_slices):
    """SKGPU_SUBMIT_SLICED (upload / kernels / read-back of consecutive table slices overlapped on three streams) must give
    the same bytes as the plain submit, tick after tick, also when sliced and plain ticks alternate and ticks are pipelined."""
    S, K, T = 53, 2, 9
    a = chain.ChainTick(S, K, seed=13)
    b = chain.ChainTick(S, K, seed=13)
    try:
        b.plan.auto_slices(b.op_chain, n_slices)
        for t in range(T):
            x = synth.tone_streams(13, t, a.n_streams, a.chunk, 2, 44100)
            want = a.tick(x)
            sliced = (t % 4) != 3
            got = b.tick(x, L.SUBMIT_SLICED if sliced else L.SUBMIT_GRAPH)
            assert np.array_equal(got, want), f"tick {t}"
            if sliced:
                tm = b.plan.slice_timing(b.plan.tick_count())
                assert len(tm) == min(n_slices, S) and all(lat >= ker >= 0 for _, ker, lat in tm)
        res = b.results()
        assert np.all(res["status"] == 0) and np.all(res["emitted"] == 1)
    finally:
        a.close()
        b.close()


def test_sliced_ticks_pipelined_two_in_flight():
    """two sliced ticks in flight (submit n + 1 before collecting n, two host output buffers): the cross-tick ordering --
    uploads into the other bank wait for the kernels that still read it, kernels wait for the read-back of the rows they
    overwrite -- keeps every tick's bytes equal to the one-at-a-time run"""
    S, K, T = 300, 2, 10
    a = chain.ChainTick(S, K, seed=21)
    b = chain.ChainTick(S, K, seed=21)
    try:
        b.plan.auto_slices(b.op_chain, 7)
        xs = [synth.tone_streams(21, t, a.n_streams, a.chunk, 2, 44100) for t in range(T)]
        want = [a.tick(x) for x in xs]
        ins = [b.ctx.pinned(b.in_bytes, np.float32) for _ in range(2)]
        outs = [b.ctx.pinned(b.out_bytes, np.int16) for _ in range(2)]
        got = []
        for t in range(T):
            ins[t & 1].reshape(b.n_streams, b.in_stride // 4)[:, : xs[t].shape[1]] = xs[t]
            b.plan.submit(ins[t & 1], outs[t & 1], L.SUBMIT_SLICED)
            if t >= 1:
                b.plan.wait_for(t)                                  # tick number t = the previous submit
                got.append(outs[(t - 1) & 1].reshape(S, -1).copy())
        b.plan.wait()
        got.append(outs[(T - 1) & 1].reshape(S, -1).copy())
        for t in range(T):
            assert np.array_equal(got[t], want[t]), f"tick {t}"
    finally:
        a.close()
        b.close()


def test_slices_are_validated():
    ct = chain.ChainTick(8, 2, seed=1)
    try:
        sl = np.zeros(2, dtype=L.SLICE_DT)
        sl["group_end"] = [4, 8]
        sl["input_end"] = [8, 16]
        sl["h2d_end"] = [ct.in_bytes // 2, ct.in_bytes]
        ct.plan.set_slices(ct.op_chain, sl)
        bad = sl.copy()
        bad["h2d_end"][0] = 16                                       # slice 0's inputs are not uploaded by then
        with pytest.raises(L.SkgpuError) as e:
            ct.plan.set_slices(ct.op_chain, bad)
        assert "beyond the slice's h2d_end" in e.value.msg
        bad = sl.copy()
        bad["group_end"][1] = 7
        with pytest.raises(L.SkgpuError):
            ct.plan.set_slices(ct.op_chain, bad)
        ct.plan.set_slices(ct.op_chain, sl[:0])                      # clears
        with pytest.raises(L.SkgpuError) as e:
            ct.plan.submit(ct.host_in, ct.host_out, L.SUBMIT_SLICED)
        assert "no slices set" in e.value.msg
    finally:
        ct.close()


def test_pinned_alloc_placement_flags(ctx):
    a = ctx.pinned(1 << 20, np.uint8)                               # NUMA-local (default): bound to the GPU's node when the box has nodes
    assert ctx.last_pinned_node in (-1, ctx.numa_node())
    b = ctx.pinned(1 << 20, np.uint8, flags=L.PIN_WRITE_COMBINED)
    a[:] = 7
    b[:] = 9
    assert a[-1] == 7
    ctx.bind_thread()                                               # no-op on a single-node box


@pytest.mark.parametrize("graph", [False, True])
def test_resident_sliced_kernel_network_equals_whole_tick(graph):
    """inputs resident, nothing read back: the sliced tick is a fork / join of k_phase_chain (stream_p) and k_chain (stream_k)
    launches -- optionally one captured CUDA graph -- in which phase(i + 1) overlaps chain(i). Same bytes as whole-tick launches."""
    S, K, T = 97, 2, 7
    a = chain.ChainTick(S, K, seed=29)
    b = chain.ChainTick(S, K, seed=29)
    try:
        b.plan.auto_slices(b.op_chain, 5)
        x = synth.tone_streams(29, 0, a.n_streams, a.chunk, 2, 44100)
        hin = np.zeros(a.in_bytes // 4, np.float32)
        hin.reshape(a.n_streams, a.in_stride // 4)[:, : x.shape[1]] = x
        for ct in (a, b):
            ct.plan.upload(0, hin)
            ct.plan.upload(ct.bank_stride, hin)
        for t in range(T):
            a.plan.submit(None, None, L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H)
            b.plan.submit(None, None, L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_SLICED | (L.SUBMIT_GRAPH if graph else 0))
            a.plan.wait()
            b.plan.wait()
            want = a.plan.download(a.out_off, a.out_bytes, np.int16)
            got = b.plan.download(b.out_off, b.out_bytes, np.int16)
            assert np.array_equal(got, want), f"tick {t}"
        assert np.any(want != 0)
        assert np.array_equal(a.results(), b.results())
    finally:
        a.close()
        b.close()


# ------------------------------------------------------------------ input kinds of real pipelines

def _kind_inputs(rng, ct, t, seed):
    xs = []
    for i, (r, c, f) in enumerate(zip(ct.rates, ct.chunks, ct.fmts)):
        x = synth.tone_streams(seed + 17 * i, t, ct.S, c, ct.C, r)
        if t % 5 == 3:
            x[rng.integers(0, ct.S), rng.integers(0, x.shape[1])] = np.float32(-0.0)   # signed zero must survive a bypass input
        xs.append(np.clip(np.rint(x * 32767.0), -32768, 32767).astype(np.int16) if f else x)
    return xs


@pytest.mark.parametrize("rates,s16,channels", [
    ([48000], [False], 1),                         # one Opus-decoded participant: 48 kHz mono, untouched (opus.rs:103,122-131)
    ([48000, 48000, 48000], [False] * 3, 1),       # samples/pipelines/dynamic/moq_mixing.yml: decoders -> gain -> clocked mixer, no resampler
    ([48000, 48000], [False, False], 2),
    ([44100, 48000], [False, False], 2),           # a resampled and a bypass participant in one mix
    ([48000, 44100, 16000], [False, False, False], 1),
    ([44100, 44100], [True, True], 2),             # s16 ingest: half the PCIe bytes
    ([44100, 48000], [True, True], 2),
    ([48000, 44100], [True, False], 1),
    ([32000], [True], 2),
])
def test_chain_input_kinds_bypass_mono_s16(rates, s16, channels):
    S, T, seed = 9, 8, 41
    ct = chain.ChainTick(S, len(rates), in_rate=rates, channels=channels, seed=seed, s16=s16)
    ref = chain_ref.MixedChainOracle(S, rates, s16, channels, seed)
    rng = np.random.default_rng(3)
    try:
        for t in range(T):
            xs = _kind_inputs(rng, ct, t, seed)
            got = ct.tick(xs, L.SUBMIT_GRAPH if t % 2 else 0)
            want = ref.tick(xs)
            assert np.array_equal(got, want), f"tick {t}: {(got != want).sum()} s16 samples differ"
            res = ct.results()
            assert np.all(res["status"] == 0)
        assert np.any(got != 0)
    finally:
        ct.close()


def test_bypass_input_is_not_interpolated():
    """(1 - 0) * y0 + 0 * y1 is NOT y0 when y1 is not finite or y0 is -0.0; a bypass input must come through untouched"""
    S = 3
    ct = chain.ChainTick(S, 1, in_rate=[48000], channels=1, seed=2)
    try:
        ct.plan.set_gains(np.ones(S + S, np.float32))              # unit gains: the packet itself reaches the s16 stage
        x = np.zeros((S, 960), np.float32)
        x[0, 10] = np.inf                                           # 0 * inf would poison frame 9 under interpolation
        x[0, 9] = 0.25
        x[1, :] = -0.0
        x[2, 100] = 1.0
        got = ct.tick([x])
        assert got[0, 9] == 8192 and got[0, 10] == 32767 and got[0, 11] == 0
        assert not np.any(got[1]) and got[2, 100] == 32767 and np.count_nonzero(got[2]) == 1
    finally:
        ct.close()


def test_bypass_and_s16_absent_streams():
    S, T, seed = 7, 10, 43
    rates, s16 = [48000, 44100], [True, True]
    ct = chain.ChainTick(S, 2, in_rate=rates, channels=2, seed=seed, s16=s16)
    ref = chain_ref.MixedChainOracle(S, rates, s16, 2, seed)
    rng = np.random.default_rng(5)
    last = None
    try:
        for t in range(T):
            xs = _kind_inputs(rng, ct, t, seed)
            present = (rng.random((S, 2)) < 0.7) if t >= 2 else np.ones((S, 2), bool)
            if last is not None:                                     # the host repeats an absent stream's previous chunk bytes
                for i in range(2):
                    xs[i][~present[:, i]] = last[i][~present[:, i]]
            ct.plan.set_present(ct.op_chain, present.reshape(-1).astype(np.uint8))
            got = ct.tick(xs)
            want = ref.tick(xs, present)
            assert np.array_equal(got, want), f"tick {t}"
            last = xs
    finally:
        ct.close()


def test_resample_op_rejects_bypass_streams(ctx):
    slot = ctx.stream_open(48000, 48000, 960, 2)
    plan = L.Plan(ctx, 1 << 20)
    try:
        items = np.zeros(1, dtype=L.RS_ITEM_DT)
        items["slot"] = slot
        items["out_off"] = 65536
        items["out_cap_frames"] = 1000
        with pytest.raises(L.SkgpuError) as e:
            plan.add_resample(items, 32768)
        assert "bypasses the resampler" in e.value.msg
    finally:
        plan.destroy()
        ctx.stream_close(slot)
