"""The frame-batching layer (include/skgpu_hub.h, libskgpu_hub.so): the C++ stand-in for the layer the north star adds
to crates/engine. CPU tests: the library loads, exports what the header declares and fails loudly without a GPU.
GPU tests: sessions driven through the hub -- with session churn, inputs that skip ticks, gain updates -- deliver the
oracle's s16 bytes (per input: audio::resampler -> audio::gain; audio::mixer clocked; audio::gain; f32 -> s16)."""
import collections
import os
import re
import subprocess

import numpy as np
import pytest

from streamkit_b200 import hub as H, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_present():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_hub_library_exports_every_declared_symbol():
    src = open(os.path.join(ROOT, "include", "skgpu_hub.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(skgpu_hub_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 20
    out = subprocess.check_output(["nm", "-D", "--defined-only", H.HUB_LIB_PATH], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert not [n for n in names if n not in exported]
    assert sorted(H.EXPORTS) == names


@pytest.mark.skipif(_gpu_present(), reason="only meaningful on a box without a GPU")
def test_hub_without_gpu_fails_loudly():
    with pytest.raises(H.HubError) as e:
        H.Hub(4, 8, [44100])
    assert e.value.rc == -5 and "no CPU fallback" in e.value.msg


# ------------------------------------------------------------------------------------------------ GPU

class _OracleSession:
    """the same session expressed with the CPU restatement of the reference nodes (oracle/sko.py)"""

    def __init__(self, in_rates, channels, out_frames):
        from oracle import sko
        self.sko = sko
        self.C, self.F = channels, out_frames
        self.nodes = [sko.ResamplerNode(48000, chunk_frames=r * out_frames // 48000, output_frame_size=out_frames) for r in in_rates]
        self.queues = [collections.deque() for _ in in_rates]
        self.in_gain = [1.0] * len(in_rates)
        self.master = 1.0
        self.rates = list(in_rates)

    def push(self, i, x):
        n = self.nodes[i]
        n.out.clear()
        n.push(self.rates[i], self.C, x)
        for pkt in n.out:
            self.queues[i].append(self.sko.gain(pkt["samples"], float(self.in_gain[i])))

    def tick(self):
        frames = []
        for q in self.queues:
            if q:
                frames.append((q.popleft(), self.C, True))
        mixed = self.sko.mix_clocked(frames, self.C, self.F)
        return self.sko.gain_f32_to_s16(mixed, float(self.master)), len(frames)


def _chunk(seed, tick, rate, frames, channels):
    return synth.tone_streams(seed, tick, 1, frames, channels, rate)[0]


@pytest.mark.gpu
def test_hub_sessions_match_oracle_with_churn_absences_and_gain_updates():
    rng = np.random.default_rng(7)
    hub = H.Hub(max_sessions=8, max_streams=24, in_rates=[44100, 32000, 16000], max_inputs_per_session=4)
    try:
        live = {}     # hub session id -> (oracle session, seed)

        def open_session(rates, seed):
            sid = hub.session_open(rates)
            live[sid] = (_OracleSession(rates, 2, 960), seed, [0] * len(rates))

        open_session([44100, 44100], 1)
        open_session([32000], 2)
        open_session([44100, 16000, 32000], 3)
        for t in range(24):
            if t == 5:
                open_session([16000, 44100], 4)                       # joins mid-run: fresh resampler state
            if t == 9:
                sid = sorted(live)[1]
                hub.session_close(sid)
                del live[sid]
            if t == 11:
                open_session([44100, 44100, 44100, 44100], 5)        # reuses freed slots
            if t == 7:
                sid = sorted(live)[0]
                hub.set_input_gain(sid, 1, 0.25)
                hub.set_master_gain(sid, 1.75)
                live[sid][0].in_gain[1] = 0.25
                live[sid][0].master = 1.75
                with pytest.raises(H.HubError):                      # gain.rs:50-66: rejected, the old gain stays
                    hub.set_input_gain(sid, 0, 4.5)
                with pytest.raises(H.HubError):
                    hub.set_master_gain(sid, float("nan"))
            want = {}
            for sid, (osess, seed, sent) in live.items():
                for i, r in enumerate(osess.rates):
                    if t >= 3 and rng.random() < 0.2:
                        continue                                      # this input skips the tick: silence, state kept
                    n = hub.chunk_frames(sid, i)
                    assert n == r * 960 // 48000
                    x = _chunk(seed * 10 + i, sent[i], r, n, 2)
                    sent[i] += 1
                    if (t + i) % 3 == 0:                              # zero-copy path: write straight into the pinned slot
                        dst = hub.acquire(sid, i)
                        assert dst.size == n * 2
                        dst[:] = x
                        hub.commit(sid, i)
                    else:
                        hub.push(sid, i, x)
                    osess.push(i, x)
                want[sid] = osess.tick()
            hub.tick()
            hub.wait()
            for sid, (w, n_frames) in want.items():
                got, n_mixed, status = hub.output(sid)
                assert status == 0
                assert n_mixed == n_frames, (t, sid, n_mixed, n_frames)
                assert got is not None and np.array_equal(got, w), f"tick {t} session {sid}: {(got != w).sum()} samples differ"
        assert hub.live_sessions == len(live)
        assert hub.live_streams == sum(len(o.rates) for o, _, _ in live.values())
    finally:
        hub.close()


@pytest.mark.gpu
def test_hub_pipelined_collection_matches_oracle():
    """submit tick n + 1 first, then collect tick n (skgpu_hub_wait_tick): same bytes as the oracle, one tick late"""
    hub = H.Hub(max_sessions=4, max_streams=8, in_rates=[44100, 32000], max_inputs_per_session=2)
    try:
        rates = [[44100, 32000], [44100], [32000, 32000]]
        sids = [hub.session_open(r) for r in rates]
        osess = [_OracleSession(r, 2, 960) for r in rates]
        pending = None   # (tick number, {sid: (want, n)})
        for t in range(12):
            want = {}
            for sid, o, r in zip(sids, osess, rates):
                for i, rate in enumerate(r):
                    x = _chunk(100 + sid * 4 + i, t, rate, rate * 960 // 48000, 2)
                    hub.push(sid, i, x)
                    o.push(i, x)
                want[sid] = o.tick()
            hub.tick()
            k = hub.ticks
            if pending is not None:
                hub.wait_tick(pending[0])          # tick n while tick n + 1 is in flight
                for sid, (w, n) in pending[1].items():
                    got, n_mixed, status = hub.output(sid)
                    assert status == 0 and n_mixed == n and np.array_equal(got, w), (t, sid)
            pending = (k, want)
        hub.wait_tick(pending[0])
        for sid, (w, n) in pending[1].items():
            got, n_mixed, status = hub.output(sid)
            assert status == 0 and n_mixed == n and np.array_equal(got, w)
        with pytest.raises(H.HubError):
            hub.wait_tick(hub.ticks - 2)           # only the two most recent ticks can be waited for
    finally:
        hub.close()


@pytest.mark.gpu
def test_hub_jitter_queue_matches_the_clocked_mixer_ring():
    """bursty arrival: an input may deliver 0..4 chunks between two ticks. With jitter_frames = 3 the hub queues up to three
    chunks per input, consumes one per tick and drops the OLDEST on overflow -- InputRingBuffer of the clocked mixer
    (mixer.rs:1185-1206, jitter_buffer_frames default 3). The oracle models exactly that with a deque(maxlen=3)."""
    J = 3
    rng = np.random.default_rng(21)
    rates = [[44100, 44100], [32000]]
    hub = H.Hub(max_sessions=2, max_streams=4, in_rates=[44100, 32000], max_inputs_per_session=2, jitter_frames=J)
    try:
        sids = [hub.session_open(r) for r in rates]
        osess = [_OracleSession(r, 2, 960) for r in rates]
        rings = [[collections.deque(maxlen=J) for _ in r] for r in rates]
        sent = [[0] * len(r) for r in rates]
        for t in range(30):
            for a, (sid, r) in enumerate(zip(sids, rates)):
                for i, rate in enumerate(r):
                    burst = int(rng.choice([0, 1, 1, 1, 2, 4])) if t >= 2 else 1
                    for _ in range(burst):
                        x = _chunk(300 + a * 4 + i, sent[a][i], rate, rate * 960 // 48000, 2)
                        sent[a][i] += 1
                        hub.push(sid, i, x)
                        rings[a][i].append(x)          # maxlen: the oldest chunk falls out
            want = []
            for a, o in enumerate(osess):
                for i in range(len(rates[a])):
                    if rings[a][i]:
                        o.push(i, rings[a][i].popleft())   # one chunk per input per tick reaches the resampler
                want.append(o.tick())
            hub.tick()
            hub.wait()
            for sid, (w, n) in zip(sids, want):
                got, n_mixed, status = hub.output(sid)
                assert status == 0 and n_mixed == n, (t, sid, n_mixed, n)
                assert np.array_equal(got, w), f"tick {t} session {sid}: {(got != w).sum()} samples differ"
    finally:
        hub.close()


@pytest.mark.gpu
def test_hub_rejects_what_the_fused_chain_cannot_do():
    with pytest.raises(H.HubError) as e:
        H.Hub(4, 8, [48000])                    # equal rates: the reference bypasses the resampler (resampler.rs:299-373)
    assert "bypass" in e.value.msg
    with pytest.raises(H.HubError):
        H.Hub(4, 8, [44101])                    # 20 ms of 44101 Hz is not a whole number of frames
    hub = H.Hub(2, 4, [44100], max_inputs_per_session=2)
    try:
        hub.tick()                              # a tick with no session at all is a no-op, not an error
        hub.wait()
        with pytest.raises(H.HubError):
            hub.session_open([22050])           # rate not declared at creation
        with pytest.raises(H.HubError):
            hub.session_open([44100] * 3)       # more inputs than max_inputs_per_session
        a = hub.session_open([44100, 44100])
        b = hub.session_open([44100, 44100])
        with pytest.raises(H.HubError):
            hub.session_open([44100])           # out of session and stream slots
        with pytest.raises(H.HubError):
            hub.push(a, 0, np.zeros(100 * 2, np.float32))   # wrong chunk length
        hub.session_close(b)
        c = hub.session_open([44100])
        hub.tick()                              # sessions that never pushed anything: silence, nothing mixed
        hub.wait()
        for sid in (a, c):
            got, n_mixed, status = hub.output(sid)
            assert n_mixed == 0 and status == 0 and got is not None and not np.any(got)
    finally:
        hub.close()
