"""GPU tests of the drop-in boundary: the native plugin ABI v2 libraries (libskgpu_plugin_*.so) driven by the same
host shim that drives the reference's own C plugin, and the C++ node mirror (libskgpu_nodes.so)."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import sko
from streamkit_b200 import synth
from tests import plugin_host as ph

pytestmark = pytest.mark.gpu
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "streamkit_b200", "csrc")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_gpu_gain_plugin_is_a_drop_in_for_the_reference_c_plugin():
    gpu = ph.NativePlugin(os.path.join(CSRC, "libskgpu_plugin_gain.so"))
    assert gpu.kind == "gpu_gain" and gpu.inputs == ["in"] and gpu.outputs == ["out"]
    assert gpu.input_formats == [(0, (0, 0, 0))]            # RawAudio{0, 0, F32}: same wildcard pin as gain.rs:89-101
    ref = ph.NativePlugin(sko.REF_GAIN_PLUGIN) if os.path.exists(sko.REF_GAIN_PLUGIN) else None
    gi = gpu.create('{"gain": 2.0}')
    ri = ref.create('{"gain": 2.0}') if ref else None
    # reference unit test gain.rs:233-286: 50 stereo frames of 0.5 at gain 2.0 -> 100 samples of ~1.0
    out = gi.process_audio(48000, 2, np.full(100, 0.5, np.float32))
    assert len(out) == 1 and out[0][0] == "out"
    pkt = out[0][1]
    assert pkt["sample_rate"] == 48000 and pkt["channels"] == 2 and pkt["samples"].size == 100
    assert np.all(np.abs(pkt["samples"] - 1.0) < 1e-3)
    x = synth.uniform_pcm(9, 1920)
    for g in (0.0, 0.37, 1.0, 3.999, 4.0):
        ok, _ = gi.update_params('{"gain": %r}' % g)
        assert ok
        y = gi.process_audio(48000, 2, x)[0][1]["samples"]
        assert np.array_equal(bits(y), bits(sko.gain(x, np.float32(g))))
        if ri:
            ri.update_params('{"gain": %r}' % g)
            assert np.array_equal(bits(y), bits(ri.process_audio(48000, 2, x)[0][1]["samples"]))
    # invalid live update: rejected, old gain stays (gain.rs:157-171)
    ok, msg = gi.update_params('{"gain": 4.5}')
    assert not ok and "must be between" in msg
    ok, msg = gi.update_params('{"gain": "loud"}')
    assert not ok
    y = gi.process_audio(48000, 2, x)[0][1]["samples"]
    assert np.array_equal(bits(y), bits(sko.gain(x, np.float32(4.0))))
    # variable packet sizes, empty packet
    for n in (1, 7, 960, 5000):
        xx = synth.uniform_pcm(n, n)
        assert np.array_equal(bits(gi.process_audio(16000, 1, xx)[0][1]["samples"]), bits(sko.gain(xx, np.float32(4.0))))
    assert gi.flush() == []
    gi.destroy()
    if ri:
        ri.destroy()
    # construction with an out-of-range gain fails (gain.rs:81-84 -> StreamKitError::Configuration)
    with pytest.raises(ph.PluginError):
        gpu.create('{"gain": 9.0}')
    # missing / empty params fall back to the default gain 1.0 (gain.rs:29,38-42)
    d = gpu.create(None)
    assert np.array_equal(bits(d.process_audio(48000, 2, x)[0][1]["samples"]), bits(x))
    d.destroy()


@pytest.mark.parametrize("in_rate,target,chunk,ofs,ch,pkt_frames", [
    (48000, 16000, 960, 960, 1, 960), (44100, 48000, 960, 960, 2, 882), (48000, 24000, 960, 0, 2, 480), (16000, 48000, 320, 480, 1, 320),
])
def test_gpu_resampler_plugin_matches_reference_node_semantics(in_rate, target, chunk, ofs, ch, pkt_frames):
    plug = ph.NativePlugin(os.path.join(CSRC, "libskgpu_plugin_resampler.so"))
    assert plug.kind == "gpu_resampler"
    inst = plug.create('{"target_sample_rate": %d, "chunk_frames": %d, "output_frame_size": %d}' % (target, chunk, ofs))
    node = sko.ResamplerNode(target, chunk, ofs)
    got = []
    for c in range(9):
        x = synth.tone_streams(17, c, 1, pkt_frames, ch, in_rate)[0]
        got += [p for _, p in inst.process_audio(in_rate, ch, x)]
        node.push(in_rate, ch, x)
    got += [p for _, p in inst.flush()]
    node.finish()
    assert len(got) == len(node.out) and len(got) > 0
    for a, b in zip(got, node.out):
        assert a["sample_rate"] == b["sample_rate"] == target and a["channels"] == b["channels"] == ch
        assert a["samples"].size == b["samples"].size
        assert np.array_equal(bits(a["samples"]), bits(b["samples"]))
    inst.destroy()


def test_gpu_resampler_plugin_errors():
    plug = ph.NativePlugin(os.path.join(CSRC, "libskgpu_plugin_resampler.so"))
    for bad in ('{"target_sample_rate": 0}', '{"target_sample_rate": 48000, "chunk_frames": 0}',
                '{"target_sample_rate": 48000, "output_frame_size": 1000}', '{"chunk_frames": 960}'):
        with pytest.raises(ph.PluginError):
            plug.create(bad)                                   # resampler.rs:81-102, parse_config_required
    inst = plug.create('{"target_sample_rate": 48000}')
    inst.process_audio(44100, 2, np.zeros(882 * 2, np.float32))
    with pytest.raises(ph.PluginError, match="Audio format changed mid-stream: expected 44100Hz/2ch, got 48000Hz/2ch"):
        inst.process_audio(48000, 2, np.zeros(960 * 2, np.float32))   # fatal for the node (resampler.rs:253-279)
    inst.destroy()
    # equal rates: pure re-framing, no DSP (resampler.rs:299-373)
    inst = plug.create('{"target_sample_rate": 48000, "output_frame_size": 480}')
    x = synth.uniform_pcm(1, 700)
    out = inst.process_audio(48000, 1, x) + inst.flush()
    assert [p["samples"].size for _, p in out] == [480, 220]
    assert np.array_equal(bits(np.concatenate([p["samples"] for _, p in out])), bits(x))
    inst.destroy()


def test_gpu_pcm16_plugin_bit_exact():
    plug = ph.NativePlugin(os.path.join(CSRC, "libskgpu_plugin_pcm16.so"))
    assert plug.kind == "gpu_pcm16"
    inst = plug.create('{"gain": 1.5}')
    x = synth.uniform_pcm(5, 1920, over_range_frac=0.05)
    out = inst.process_audio(48000, 2, x)
    assert out[0][1]["kind"] == "binary"
    s = np.frombuffer(out[0][1]["data"], dtype="<i2")
    assert np.array_equal(s, sko.gain_f32_to_s16(x, np.float32(1.5)))
    inst.destroy()


def _nodes():
    lib = C.CDLL(os.path.join(CSRC, "libskgpu_nodes.so"))
    lib.skn_mixer_create.restype = C.c_void_p
    lib.skn_mixer_create.argtypes = [C.c_char_p]
    lib.skn_mixer_destroy.argtypes = [C.c_void_p]
    lib.skn_mixer_num_input_pins.restype = C.c_uint32
    lib.skn_mixer_num_input_pins.argtypes = [C.c_void_p]
    lib.skn_mixer_pin_name.argtypes = [C.c_void_p, C.c_uint32, C.c_char_p, C.c_size_t]
    lib.skn_last_error.restype = C.c_char_p
    return lib


class SknFrame(C.Structure):
    _fields_ = [("samples", C.POINTER(C.c_float)), ("n_samples", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint16),
                ("unique", C.c_uint16)]


def _mix(lib, h, clocked, frames, rate=48000):
    keep = [np.ascontiguousarray(f[0], np.float32) for f in frames]
    arr = (SknFrame * max(len(frames), 1))()
    for i, (f, k) in enumerate(zip(frames, keep)):
        arr[i] = SknFrame(k.ctypes.data_as(C.POINTER(C.c_float)), k.size, rate, f[1], 1 if f[2] else 0)
    out = np.zeros(65536, np.float32)
    ol, oc, orate = C.c_size_t(), C.c_uint16(), C.c_uint32()
    rc = lib.skn_mixer_mix(C.c_void_p(h), clocked, arr, len(frames), out.ctypes.data_as(C.POINTER(C.c_float)), out.size, C.byref(ol), C.byref(oc), C.byref(orate))
    assert rc == 0, lib.skn_last_error()
    return out[: ol.value].copy(), oc.value


def test_mixer_node_mirror_reference_scenarios():
    lib = _nodes()
    h = lib.skn_mixer_create(b'{"sync_timeout_ms": 100, "num_inputs": 2}')
    assert lib.skn_mixer_num_input_pins(C.c_void_p(h)) == 2            # in_0, in_1 (mixer.rs:128-143)
    buf = C.create_string_buffer(16)
    lib.skn_mixer_pin_name(C.c_void_p(h), 1, buf, 16)
    assert buf.value == b"in_1"
    F = lambda v, ch, n=10: (np.full(n * ch, v, np.float32), ch, True)
    # test_mixer_continues_after_eof_with_sticky_channels (mixer.rs:1705-1760)
    o, oc = _mix(lib, h, 0, [F(0.5, 2), F(0.3, 1)])
    assert oc == 2 and abs(o[0] - 0.8) < 1e-3
    o, oc = _mix(lib, h, 0, [F(0.25, 1)])
    assert oc == 2 and o.size == 20 and abs(o[0] - 0.25) < 1e-3 and abs(o[1] - 0.25) < 1e-3
    lib.skn_mixer_destroy(C.c_void_p(h))
    # random cross-check against the oracle incl. sticky state
    rng = np.random.default_rng(8)
    h = lib.skn_mixer_create(None)
    seen = 0
    for _ in range(25):
        frames = []
        for _ in range(int(rng.integers(1, 6))):
            ch = int(rng.choice([1, 2]))
            frames.append((rng.standard_normal(int(rng.integers(1, 300)) * ch).astype(np.float32), ch, bool(rng.random() < 0.7)))
        want, oc_w = sko.mix_sync(frames, seen)
        seen = max(seen, oc_w)
        got, oc = _mix(lib, h, 0, frames)
        assert oc == oc_w and np.array_equal(bits(got), bits(want))
    lib.skn_mixer_destroy(C.c_void_p(h))
    # clocked (mixer.rs:2047-2052, :2097-2102)
    h = lib.skn_mixer_create(b'{"clocked": {"sample_rate": 48000, "frame_samples_per_channel": 10, "jitter_buffer_frames": 2, "generate_silence": false}}')
    o, oc = _mix(lib, h, 1, [F(0.5, 2), F(0.3, 2)])
    assert oc == 2 and o.size == 20 and np.all(np.abs(o - 0.8) < 1e-3)
    o, oc = _mix(lib, h, 1, [F(0.75, 2)])
    assert np.all(np.abs(o - 0.75) < 1e-3)
    lib.skn_mixer_destroy(C.c_void_p(h))


# ------------------------------------------------------------------ audio::resampler node: integer packet metadata

class _SknMeta(C.Structure):
    _fields_ = [("timestamp_us", C.c_uint64), ("duration_us", C.c_uint64), ("sequence", C.c_uint64), ("has_timestamp", C.c_uint8),
                ("has_duration", C.c_uint8), ("has_sequence", C.c_uint8), ("pad", C.c_uint8)]


_SKN_EMIT = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_uint16, C.POINTER(C.c_float), C.c_size_t, C.POINTER(_SknMeta))


class _GpuResamplerNode:
    def __init__(self, params: str):
        lib = C.CDLL(os.path.join(CSRC, "libskgpu_nodes.so"))
        lib.skn_resampler_create.restype = C.c_void_p
        lib.skn_resampler_create.argtypes = [C.c_char_p]
        lib.skn_resampler_destroy.argtypes = [C.c_void_p]
        lib.skn_resampler_push.argtypes = [C.c_void_p, C.c_uint32, C.c_uint16, C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, _SKN_EMIT, C.c_void_p]
        lib.skn_resampler_finish.argtypes = [C.c_void_p, _SKN_EMIT, C.c_void_p]
        lib.skn_last_error.restype = C.c_char_p
        self.lib = lib
        self.h = lib.skn_resampler_create(params.encode())
        assert self.h, lib.skn_last_error()
        self.out = []
        self._cb = _SKN_EMIT(self._emit)

    def _emit(self, ud, rate, ch, samples, n, meta):
        m = meta.contents
        self.out.append(dict(sample_rate=rate, channels=ch, samples=np.ctypeslib.as_array(samples, shape=(n,)).copy() if n else np.zeros(0, np.float32),
                             timestamp_us=m.timestamp_us if m.has_timestamp else None, duration_us=m.duration_us if m.has_duration else None,
                             sequence=m.sequence if m.has_sequence else None))

    def push(self, rate, ch, x, timestamp_us=None):
        x = np.ascontiguousarray(x, dtype=np.float32)
        rc = self.lib.skn_resampler_push(C.c_void_p(self.h), rate, ch, x.ctypes.data_as(C.c_void_p), x.size, 0 if timestamp_us is None else 1,
                                         0 if timestamp_us is None else timestamp_us, self._cb, None)
        if rc != 0:
            raise RuntimeError(self.lib.skn_last_error().decode())

    def finish(self):
        assert self.lib.skn_resampler_finish(C.c_void_p(self.h), self._cb, None) == 0

    def close(self):
        self.lib.skn_resampler_destroy(C.c_void_p(self.h))


@pytest.mark.parametrize("in_rate,target,ch,ofs,packet,first_ts", [
    (44100, 48000, 1, 960, 882, 1000), (48000, 16000, 2, 960, 960, 123456789), (44100, 48000, 2, 0, 1000, 7),
    (48000, 48000, 2, 480, 700, 42), (22050, 48000, 1, 1920, 441, None),
])
def test_resampler_node_timestamp_duration_sequence_match_the_oracle(in_rate, target, ch, ofs, packet, first_ts):
    """resampler.rs:286-297 / :108-116: the first input packet's timestamp seeds output_timestamp_us, every emitted packet
    advances it by frames * 1_000_000 / rate (INTEGER division), sequence counts 0, 1, 2, ...; the flushed partial packet at EOF
    carries the next sequence number WITHOUT incrementing it (:707-711). Integer metadata is a bit-exact requirement: samples
    AND metadata of every packet equal the oracle node's (oracle/sk_oracle.c sko_rsnode_*)."""
    node = _GpuResamplerNode('{"target_sample_rate": %d, "chunk_frames": 960, "output_frame_size": %d}' % (target, ofs))
    ref = sko.ResamplerNode(target, 960, ofs)
    try:
        for c in range(11):
            x = synth.tone_streams(9, c, 1, packet, ch, in_rate)[0]
            ts = first_ts if (c == 0 and first_ts is not None) else (5 if first_ts is not None else None)   # later timestamps are ignored (:211-214)
            node.push(in_rate, ch, x, timestamp_us=ts)
            ref.push(in_rate, ch, x, timestamp_us=ts)
        node.finish()
        ref.finish()
        assert len(node.out) == len(ref.out) and len(node.out) >= 2
        for a, b in zip(node.out, ref.out):
            assert (a["sample_rate"], a["channels"]) == (b["sample_rate"], b["channels"])
            assert np.array_equal(a["samples"].view(np.uint32), b["samples"].view(np.uint32))
            assert a["sequence"] == b["sequence"] and a["duration_us"] == b["duration_us"] and a["timestamp_us"] == b["timestamp_us"]
        if first_ts is not None and in_rate != target and ofs:
            assert node.out[0]["timestamp_us"] == first_ts
            assert node.out[1]["timestamp_us"] == first_ts + sko.duration_us_for_frames(target, ofs)
            assert [p["sequence"] for p in node.out[:-1]] == list(range(len(node.out) - 1))
    finally:
        node.close()


# ------------------------------------------------------------------ PROPOSED ABI v3: C host + GPU mixer plugin

def _lcg_inputs(n_in, rounds, s16):
    seed = np.uint32(12345)
    outs = []
    state = int(seed)
    for _ in range(rounds):
        per = []
        for _i in range(n_in):
            v = np.empty(1920, np.int32)
            for k in range(1920):
                state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
                v[k] = (state >> 16) - 32768
            if s16:
                per.append(sko.s16_to_f32(v.astype(np.int16)))
            else:
                per.append((v.astype(np.float32) * np.float32(1.0 / 32768.0) * np.float32(0.75)).astype(np.float32))
        outs.append(per)
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("n_in,in_fmt,out_fmt", [(3, "f32", "s16"), (2, "s16", "f32"), (1, "f32", "f32")])
def test_abi_v3_c_host_drives_the_gpu_mixer_plugin(n_in, in_fmt, out_fmt, tmp_path):
    """tests/host/host_v3.c (a C host, like crates/plugin-native loads plugins) against libskgpu_plugin_mixer_v3.so: dynamic pins,
    batched process_packets, s16 payloads in / out, metadata passthrough; output bytes checked against the oracle's mixer."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "host_v3")
    subprocess.check_call(["gcc", "-std=c11", "-O1", "-Wall", "-o", exe, os.path.join(root, "tests", "host", "host_v3.c"), "-ldl"])
    rounds, gain = 3, 1.5
    out_file = str(tmp_path / "out.bin")
    res = subprocess.run([exe, os.path.join(CSRC, "libskgpu_plugin_mixer_v3.so"), '{"gain": %r, "output_format": "%s"}' % (gain, out_fmt), str(n_in), in_fmt,
                          str(rounds), out_file], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lines = res.stdout.strip().splitlines()
    assert lines[0] == "kind gpu_mixer inputs 1 cardinality 1 prefix in accepts 2"
    pk = [ln.split() for ln in lines if ln.startswith("packet ")]
    assert len(pk) == rounds + 1 and lines[-1] == "packets %d" % (rounds + 1)
    assert "update_params(gain 9.0) success 0" in lines
    ins = _lcg_inputs(n_in, rounds, in_fmt == "s16")
    raw = open(out_file, "rb").read()
    bps = 2 if out_fmt == "s16" else 4
    off = 0
    for r in range(rounds + 1):
        if r < rounds:
            n_used = n_in - 1 if (r == rounds - 1 and n_in > 1) else n_in       # the removed pin's frame is not mixed
            frames = [(ins[r][i], 2, True) for i in range(n_used)]
            ts, seq = 1000000 + 20000 * r, 100 + r                               # the mix carries the FIRST frame's metadata (mixer.rs:994)
        else:
            first = (_lcg_inputs(n_in, rounds, False)[rounds - 1][0])           # the host re-sends the f32 buffer of input 0 (last round)
            frames = [(first, 2, True)]
            ts = seq = None
        mixed, oc = sko.mix_sync(frames)
        want = sko.gain_f32_to_s16(mixed, gain) if out_fmt == "s16" else sko.gain(mixed, gain)
        got = np.frombuffer(raw[off: off + 1920 * bps], dtype=np.int16 if out_fmt == "s16" else np.float32)
        off += 1920 * bps
        assert np.array_equal(got.view(np.uint16 if out_fmt == "s16" else np.uint32), want.view(np.uint16 if out_fmt == "s16" else np.uint32)), f"packet {r}"
        f = pk[r]
        assert int(f[1]) == 1920 * bps and f[3] == "48000" and f[5] == "2" and f[7] == ("1" if out_fmt == "s16" else "0")
        assert f[9] == (str(ts) if ts is not None else "-") and f[11] == (str(seq) if seq is not None else "-")
    assert off == len(raw)
