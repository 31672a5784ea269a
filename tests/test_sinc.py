"""Windowed-sinc polyphase resampler mode (BASELINE.json north star; spec in include/skgpu_batch.h skgpu_ctx_set_sinc).
The reference has no such mode (it uses rubato's Linear, resampler.rs:232-238): parity here is against the build's own
two oracles -- C (oracle/sk_sinc.c, sequential f32 fma like the kernel) and numpy (f64 dot products) -- plus properties of
the specification itself (unit DC gain, a band-limited sine comes through to < -110 dBFS, exact output counts)."""
import numpy as np
import pytest

from oracle import np_oracle, sko
from streamkit_b200 import lib as L, synth

TOL = 2e-6    # north_star: f32 resample within 2e-6 max-abs


def test_tap_table_unit_dc_gain_and_symmetry():
    for Ln, O, fc in [(64, 256, 0.95), (32, 128, 0.95 / 3), (128, 64, 0.9)]:
        t = sko.sinc_taps(Ln, O, fc)
        assert t.shape == (O + 1, Ln)
        assert np.allclose(t.astype(np.float64).sum(axis=1), 1.0, atol=2e-6)          # every phase passes DC unchanged
        assert np.array_equal(t, np_oracle.sinc_taps(Ln, O, fc))                        # two independent statements of the table
        # phase p read backwards is phase O - p shifted by one tap (the kernel is even in tau)
        assert np.allclose(t[0, :-1], t[O, 1:], atol=1e-7)


@pytest.mark.parametrize("in_rate,out_rate,chunk,ch", [(44100, 48000, 882, 2), (48000, 16000, 960, 1), (16000, 48000, 320, 2), (48000, 44100, 960, 2)])
def test_sinc_c_oracle_equals_numpy_within_tolerance(in_rate, out_rate, chunk, ch):
    a = sko.SincFixedIn(in_rate, out_rate, chunk, ch)
    b = np_oracle.SincFixedIn(in_rate, out_rate, chunk, ch)
    for c in range(5):
        x = synth.tone_streams(4, c, 1, chunk, ch, in_rate)[0]
        ya, yb = a.process(x), b.process(x)
        assert ya.size == yb.size and a.last_index == b.last_index
        assert np.max(np.abs(ya - yb)) <= TOL


def test_sinc_reconstructs_a_band_limited_sine_below_minus_110_dbfs():
    for in_rate, out_rate, chunk in [(44100, 48000, 882), (48000, 16000, 960), (16000, 48000, 320)]:
        r = sko.SincFixedIn(in_rate, out_rate, chunk, 1)
        n = chunk * 8
        f = 1000.0
        x = (0.9 * np.sin(2 * np.pi * f * np.arange(n) / in_rate)).astype(np.float32)
        y = np.concatenate([r.process(x[i * chunk:(i + 1) * chunk]) for i in range(8)]).astype(np.float64)
        pos = -32 + (np.arange(y.size) + 1) * (in_rate / out_rate)                       # -(L/2) + (k + 1) t
        ideal = 0.9 * np.sin(2 * np.pi * f * pos / in_rate)
        m = pos > 72
        assert np.max(np.abs(y[m] - ideal[m])) < 10 ** (-110 / 20)                       # 3.2e-6


def test_sinc_output_counts_match_exact_arithmetic():
    # same recurrence as the linear mode with the sinc constants: positions -(L/2) + j t are emitted while < (m - 1) N + N - L/2 - 1 - ceil(t)
    from fractions import Fraction
    import math
    for in_rate, out_rate, chunk in [(44100, 48000, 882), (48000, 44100, 960), (8000, 44100, 160)]:
        r = sko.SincFixedIn(in_rate, out_rate, chunk, 1)
        z = np.zeros(chunk, np.float32)
        tot = np.cumsum([r.process(z).size for _ in range(2000)])
        t = Fraction(in_rate, out_rate)
        for m in (1, 2, 3, 500, 2000):
            bound = Fraction((m - 1) * chunk + chunk - 32 - 1 - math.ceil(t) + 32)
            q = bound / t
            assert q.denominator != 1
            assert int(tot[m - 1]) == math.ceil(q)


# ------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("in_rate,out_rate,chunk,ch,params", [
    (44100, 48000, 882, 2, (64, 256, 0.95)), (48000, 16000, 960, 1, (64, 256, 0.95)), (48000, 16000, 960, 2, (64, 256, 0.95)),
    (16000, 48000, 320, 2, (64, 256, 0.95)), (44100, 48000, 882, 1, (32, 128, 0.9)), (48000, 44100, 960, 2, (128, 512, 0.95)),
    (22050, 48000, 441, 2, (64, 256, 0.95)),
])
def test_sinc_kernel_matches_the_c_oracle(in_rate, out_rate, chunk, ch, params):
    from tests.gpu_helpers import GpuResampler, bits
    ctx = L.Context(device=0, max_streams=8, max_channels=2)
    try:
        ctx.set_sinc(*params)
        g = GpuResampler(ctx, in_rate, out_rate, chunk, ch, n_streams=3, flags=L.STREAM_SINC)
        refs = [sko.SincFixedIn(in_rate, out_rate, chunk, ch, *params) for _ in range(3)]
        exact = True
        for c in range(10):
            x = synth.tone_streams(chunk + ch, c, 3, chunk, ch, in_rate)
            outs = g.process(x)
            for s in range(3):
                want = refs[s].process(x[s])
                assert outs[s].size == want.size, (c, s, outs[s].size, want.size)        # counts exact
                assert np.max(np.abs(outs[s] - want)) <= TOL
                exact = exact and np.array_equal(bits(outs[s]), bits(want))
        assert g.state(0)[0] == refs[0].last_index
        assert exact, "within tolerance but not bit-exact (the kernel runs the oracle's sequential f32 fma chains)"
        g.close()
        with pytest.raises(L.SkgpuError):
            ctx.set_sinc(64, 256, 0.95)                                                  # fixed for the context's lifetime
    finally:
        ctx.close()


@pytest.mark.gpu
def test_sinc_and_linear_streams_do_not_share_an_op_and_chain_rejects_sinc():
    ctx = L.Context(device=0, max_streams=8, max_channels=2)
    try:
        with pytest.raises(L.SkgpuError) as e:
            ctx.stream_open(44100, 48000, 882, 2, L.STREAM_SINC)
        assert "skgpu_ctx_set_sinc first" in e.value.msg
        ctx.set_sinc()
        a = ctx.stream_open(44100, 48000, 882, 2, L.STREAM_SINC)
        b = ctx.stream_open(44100, 48000, 882, 2)
        plan = L.Plan(ctx, 1 << 20)
        items = np.zeros(2, dtype=L.RS_ITEM_DT)
        items["slot"] = [a, b]
        items["in_off"] = [0, 8192]
        items["out_off"] = [65536, 131072]
        items["out_cap_frames"] = 1000
        with pytest.raises(L.SkgpuError) as e:
            plan.add_resample(items, 32768)
        assert "cannot share" in e.value.msg
        plan.destroy()
    finally:
        ctx.close()


def _run_sinc_op(ctx, cfgs, ticks, params, update_at=None):
    """cfgs: list of (in_rate, out_rate, chunk, channels); one resample op over all of them, `ticks` chunks each, checked
    bit for bit against one C-oracle resampler per stream. Returns the number of compared frames."""
    from tests.gpu_helpers import al, bits
    n = len(cfgs)
    slots = [ctx.stream_open(i, o, c, ch, L.STREAM_SINC) for (i, o, c, ch) in cfgs]
    caps = [L.Context.max_out_frames(i, o, c, ch) for (i, o, c, ch) in cfgs]
    in_off, out_off, pos = [], [], 0
    for (i, o, c, ch) in cfgs:
        in_off.append(pos)
        pos += al(c * ch * 4, 16)
    in_bytes = al(pos)
    res_off = in_bytes
    pos = al(res_off + 8 * n)
    for (i, o, c, ch), cap in zip(cfgs, caps):
        out_off.append(pos)
        pos += al(cap * ch * 4, 16)
    total = al(pos)
    plan = L.Plan(ctx, total)
    items = np.zeros(n, dtype=L.RS_ITEM_DT)
    items["in_off"], items["out_off"], items["slot"], items["out_cap_frames"] = in_off, out_off, slots, caps
    plan.add_resample(items, res_off)
    plan.set_io(0, in_bytes, res_off, total - res_off)
    plan.finalize()
    host_in, host_out = np.zeros(in_bytes, np.uint8), np.zeros(total - res_off, np.uint8)
    refs = [sko.SincFixedIn(i, o, c, ch, *params) for (i, o, c, ch) in cfgs]
    frames = 0
    for t in range(ticks):
        xs = []
        for s, (i, o, c, ch) in enumerate(cfgs):
            x = synth.tone_streams(1000 + s, t, 1, c, ch, i)[0]
            xs.append(x)
            host_in[in_off[s]: in_off[s] + x.size * 4] = x.view(np.uint8)
        plan.submit(host_in, host_out, L.SUBMIT_GRAPH if t % 2 else 0)
        plan.wait()
        res = host_out[: 8 * n].view(L.RS_RESULT_DT)
        for s, (i, o, c, ch) in enumerate(cfgs):
            want = refs[s].process(xs[s])
            assert res[s]["status"] == 0 and int(res[s]["out_frames"]) * ch == want.size, (t, s)
            got = host_out[out_off[s] - res_off: out_off[s] - res_off + want.size * 4].view(np.float32)
            assert np.array_equal(bits(got), bits(want)), (t, s, cfgs[s])
            frames += want.size // ch
    plan.destroy()
    for sl in slots:
        ctx.stream_close(sl)
    return frames


@pytest.mark.gpu
def test_sinc_persistent_kernel_many_streams_more_passes_than_ctas():
    """700 stereo 44.1k->48k streams = more passes than SMs: every CTA of the persistent kernel refills both ring stages"""
    params = (64, 256, 0.95)
    ctx = L.Context(device=0, max_streams=1024, max_channels=2)
    try:
        ctx.set_sinc(*params)
        assert _run_sinc_op(ctx, [(44100, 48000, 882, 2)] * 700, 3, params) > 700 * 2700
    finally:
        ctx.close()


@pytest.mark.gpu
def test_sinc_one_op_with_mixed_up_sampling_ratios_channels_stay_per_op():
    """up-sampling streams share the tap table (cutoff 0.95): different periods, chunk sizes and item counts in one op"""
    params = (64, 256, 0.95)
    ctx = L.Context(device=0, max_streams=256, max_channels=2)
    try:
        ctx.set_sinc(*params)
        cfgs = [(44100, 48000, 882, 2), (16000, 48000, 320, 2), (22050, 48000, 441, 2), (8000, 48000, 160, 2), (32000, 48000, 640, 2),
                (11025, 48000, 441, 2), (44100, 48000, 441, 2), (24000, 48000, 480, 2)] * 9
        _run_sinc_op(ctx, cfgs, 4, params)
        # mono, with a chunk that is not a 16-byte multiple (441 x 4 B): the cooperative copy path
        _run_sinc_op(ctx, [(44100, 48000, 441, 1), (44100, 48000, 882, 1), (22050, 48000, 441, 1)] * 5, 4, params)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_sinc_one_op_with_several_tap_tables_takes_the_per_stream_kernel():
    """down-sampling scales the cutoff with the ratio: streams with different tables in one op still resample exactly"""
    params = (64, 256, 0.95)
    ctx = L.Context(device=0, max_streams=64, max_channels=2)
    try:
        ctx.set_sinc(*params)
        _run_sinc_op(ctx, [(44100, 48000, 882, 2), (48000, 16000, 960, 2), (48000, 44100, 960, 2), (48000, 8000, 960, 2)] * 3, 4, params)
        _run_sinc_op(ctx, [(48000, 16000, 960, 2)] * 40, 3, params)      # one down-sampling table: persistent kernel, every output one sub-phase
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_sinc_random_rate_pairs_and_filter_shapes(seed):
    """random common-rate pairs (up- and down-sampling mixed tables -> per-stream kernel; same-table groups -> persistent kernel),
    random filter shapes: every output frame equals the C oracle's bit for bit"""
    rng = np.random.default_rng(seed)
    rates = [8000, 11025, 12000, 16000, 22050, 24000, 32000, 44100, 48000, 88200, 96000]
    params = (int(rng.choice([16, 32, 64, 128])), int(rng.choice([32, 128, 256, 512])), float(rng.choice([0.8, 0.9, 0.95, 1.0])))
    ch = int(rng.choice([1, 2]))
    ctx = L.Context(device=0, max_streams=64, max_channels=2)
    try:
        ctx.set_sinc(*params)
        def chunk_of(r):
            c = max(r // 50, params[0] + 8)
            return c + (c * ch) % 2                      # keep stereo / mono chunks a whole number of 8-byte units
        ups, mixed = [], []
        for _ in range(6):
            i, o = (int(v) for v in rng.choice(rates, 2, replace=False))
            (ups if o > i else mixed).append((i, o, chunk_of(i), ch))
        mixed = mixed + ups[:2]
        if ups:
            _run_sinc_op(ctx, ups * 3, 3, params)        # one tap table (cutoff unscaled): the persistent kernel
        if mixed:
            _run_sinc_op(ctx, mixed, 3, params)          # several tables: one CTA per stream
    finally:
        ctx.close()
