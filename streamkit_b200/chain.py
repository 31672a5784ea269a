"""Host-side builder of the full-chain tick (BASELINE config #5) on top of the batch C ABI:

    per input i of a session:  audio::resampler{target 48000, chunk_frames = in frames/tick, output_frame_size F}
                               -> audio::gain{g_i}
    audio::mixer (clocked 48 kHz / F, inputs in pin order) -> audio::gain{master} -> f32 -> s16

Sessions are the unit of sharding: every session's streams live on one GPU, no collective.
This is what a frame-batching layer does each tick: gather all sessions' 20 ms frames into one pinned
arena, one submit, scatter the s16 results.

Input kinds per input index: `in_rate` may be a list (one rate per input of a session); a rate equal to
48000 is a BYPASS input (the reference's resampler forwards such packets untouched, resampler.rs:299-373); `s16` (bool or
list) makes an input arrive as interleaved s16 on PCIe (x = s / 32768).
"""
from __future__ import annotations

import numpy as np

from . import lib as L
from . import synth

OUT_RATE = 48000
OUT_FRAMES = 960  # output_frame_size / frame_samples_per_channel


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


class ChainTick:
    def __init__(self, n_sessions: int, k_inputs: int, in_rate=44100, channels: int = 2, device: int = 0,
                 seed: int = 0, chunk_frames: int | None = None, fused: bool = True, out_frames: int = OUT_FRAMES, alloc_host: bool = True,
                 s16=False, out_rate: int = OUT_RATE):
        """fused=True: one k_chain launch per tick (double-banked input, lagged recompute, no HBM intermediates);
        fused=False: the general unfused ops (k_resample -> device re-framing ring -> k_mix)."""
        self.fused = fused
        self.F = out_frames   # output_frame_size of the resampler nodes = frame_samples_per_channel of the clocked mixer
        self.S, self.K, self.C = n_sessions, k_inputs, channels
        self.out_rate = out_rate
        rates = list(in_rate) if isinstance(in_rate, (list, tuple)) else [in_rate] * k_inputs
        fmts = list(s16) if isinstance(s16, (list, tuple)) else [bool(s16)] * k_inputs
        assert len(rates) == k_inputs and len(fmts) == k_inputs
        self.rates, self.fmts = rates, fmts
        self.in_rate = rates[0]
        # frames per tick of every input index: a tick is F / out_rate seconds
        self.chunks = [chunk_frames if chunk_frames is not None else r * out_frames // out_rate for r in rates]
        self.chunk = self.chunks[0]
        self.n_streams = n_sessions * k_inputs
        self.ctx = L.Context(device=device, max_streams=self.n_streams, max_channels=channels, fifo_frames=0 if fused else 2048)
        # one arena row per stream; the row stride is the largest chunk of the session shape (+ 16 bytes of slack after s16 rows)
        self.row_bytes = [c * channels * (2 if f else 4) for c, f in zip(self.chunks, fmts)]
        self.in_stride = _align(max(b + (16 if f else 0) for b, f in zip(self.row_bytes, fmts)), 16)
        self.out_stride = self.F * channels * 2
        self.in_bytes = self.n_streams * self.in_stride
        self.bank_stride = _align(self.in_bytes) if fused else 0
        self.res_off = _align(self.in_bytes) + self.bank_stride
        self.res_bytes = self.n_streams * 8
        self.out_off = _align(self.res_off + self.res_bytes)
        self.out_bytes = n_sessions * self.out_stride
        arena = _align(self.out_off + self.out_bytes)
        self.plan = L.Plan(self.ctx, arena)
        # gain table: [per-input gains (S*K) | master gains (S)]
        self.in_gains = synth.gains(seed, self.n_streams, 0.25, 1.5)
        self.master_gains = synth.gains(seed + 1, n_sessions, 0.5, 2.0)
        self.plan.set_gains(np.concatenate([self.in_gains, self.master_gains]))
        slots = np.zeros(self.n_streams, dtype=np.uint32)
        for i in range(k_inputs):   # stream (session s, input i) = s * K + i
            slots[i::k_inputs] = self.ctx.stream_open_many(rates[i], out_rate, self.chunks[i], channels, n_sessions, L.STREAM_S16 if fmts[i] else 0)
        self.slots = slots
        if fused:
            self.plan.set_io(0, self.in_bytes, self.out_off, self.out_bytes)
            self.plan.set_banks(self.bank_stride)
            cin = np.zeros(self.n_streams, dtype=L.CHAIN_INPUT_DT)
            cin["in_off"] = np.arange(self.n_streams, dtype=np.uint64) * self.in_stride
            cin["slot"] = slots
            cin["gain_idx"] = np.arange(self.n_streams, dtype=np.uint32)
            cin["flags"] = L.MIX_IN_UNIQUE
            cg = np.zeros(n_sessions, dtype=L.CHAIN_GROUP_DT)
            cg["out_off"] = self.out_off + np.arange(n_sessions, dtype=np.uint64) * self.out_stride
            cg["first_input"] = np.arange(n_sessions, dtype=np.uint32) * k_inputs
            cg["n_inputs"] = k_inputs
            cg["gain_idx"] = self.n_streams + np.arange(n_sessions, dtype=np.uint32)
            cg["out_channels"] = channels
            cg["flags"] = L.MIX_OUT_S16
            self.op_chain = self.plan.add_chain(cg, cin, self.F, self.res_off)
            self.op_rs = self.op_mix = None
            self.plan.finalize()
            if alloc_host:
                self.host_in = self.ctx.pinned(self.in_bytes, np.float32)
                self.host_out = self.ctx.pinned(self.out_bytes, np.int16)
            return
        assert not any(fmts) and all(r != out_rate for r in rates), "the unfused ops take f32 chunks of resampled streams"
        items = np.zeros(self.n_streams, dtype=L.RS_ITEM_DT)
        items["in_off"] = np.arange(self.n_streams, dtype=np.uint64) * self.in_stride
        items["slot"] = slots
        items["flags"] = L.RS_TO_FIFO
        self.op_rs = self.plan.add_resample(items, self.res_off)
        inputs = np.zeros(self.n_streams, dtype=L.MIX_INPUT_DT)
        inputs["n_frames"] = self.F
        inputs["channels"] = channels
        inputs["flags"] = L.MIX_IN_UNIQUE | L.MIX_IN_FIFO
        inputs["gain_idx"] = np.arange(self.n_streams, dtype=np.uint32)
        inputs["slot"] = slots
        groups = np.zeros(n_sessions, dtype=L.MIX_GROUP_DT)
        groups["out_off"] = self.out_off + np.arange(n_sessions, dtype=np.uint64) * self.out_stride
        groups["first_input"] = np.arange(n_sessions, dtype=np.uint32) * k_inputs
        groups["n_inputs"] = k_inputs
        groups["out_frames"] = self.F
        groups["out_channels"] = channels
        groups["flags"] = L.MIX_OUT_S16
        groups["gain_idx"] = self.n_streams + np.arange(n_sessions, dtype=np.uint32)
        self.op_mix = self.plan.add_mix(groups, inputs)
        self.plan.set_io(0, self.in_bytes, self.out_off, self.out_bytes)
        self.plan.finalize()
        self.host_in = self.ctx.pinned(self.in_bytes, np.float32)
        self.host_out = self.ctx.pinned(self.out_bytes, np.int16)

    # algorithmic bytes: per session-tick K*(in + state r/w) + s16 out; a bypass input has no resampler state
    def algorithmic_bytes_per_tick(self) -> int:
        per_session = sum(b + (0 if r == self.out_rate else 2 * (8 + 16 * self.C * 4)) for b, r in zip(self.row_bytes, self.rates))
        return self.S * (per_session + self.out_stride)

    def fill_rows(self, host_in: np.ndarray, inputs):
        """inputs: per input index i an array [n_sessions, chunk_i * channels] (float32, or int16 for s16 inputs) -- or one
        float32 array [n_streams, chunk * channels] when every input has the same shape"""
        rows = host_in.view(np.uint8).reshape(self.n_streams, self.in_stride)
        if isinstance(inputs, np.ndarray):
            inputs = [inputs.reshape(self.S, self.K, -1)[:, i] for i in range(self.K)]
        for i, x in enumerate(inputs):
            x = np.ascontiguousarray(x, dtype=np.int16 if self.fmts[i] else np.float32).reshape(self.S, -1)
            rows[i::self.K, : self.row_bytes[i]] = x.view(np.uint8).reshape(self.S, -1)

    def tick(self, inputs, flags: int = 0) -> np.ndarray:
        """inputs: see fill_rows; returns int16 [n_sessions, F*channels]"""
        self.fill_rows(self.host_in, inputs)
        self.plan.submit(self.host_in, self.host_out, flags)
        self.plan.wait()
        return self.host_out.reshape(self.S, self.F * self.C).copy()

    def results(self) -> np.ndarray:
        return self.plan.download(self.res_off, self.res_bytes, L.CHAIN_RESULT_DT if self.fused else L.RS_RESULT_DT)

    def close(self):
        self.plan.destroy()
        self.ctx.close()


def run_chain_gpu(n_sessions: int, k_inputs: int, ticks: int, seed: int, in_rate: int = 44100, channels: int = 2,
                  graph: bool = False, fused: bool = True, chunk_frames: int | None = None, out_frames: int = OUT_FRAMES, out_rate: int = OUT_RATE):
    ct = ChainTick(n_sessions, k_inputs, in_rate=in_rate, channels=channels, seed=seed, fused=fused, chunk_frames=chunk_frames,
                   out_frames=out_frames, out_rate=out_rate)
    try:
        outs = []
        for t in range(ticks):
            x = synth.tone_streams(seed, t, ct.n_streams, ct.chunk, channels, in_rate)
            outs.append(ct.tick(x, L.SUBMIT_GRAPH if graph else 0))
        return outs
    finally:
        ct.close()
