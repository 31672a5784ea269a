"""Host-side builder of the full-chain tick (BASELINE config #5) on top of the batch C ABI:

    per input i of a session:  audio::resampler{target 48000, chunk_frames = in frames/tick, output_frame_size 960}
                               -> audio::gain{g_i}
    audio::mixer (clocked 48 kHz / 960, inputs in pin order) -> audio::gain{master} -> f32 -> s16

Sessions are the unit of sharding: every session's streams live on one GPU, no collective.
This is what a frame-batching layer does each tick: gather all sessions' 20 ms frames into one pinned
arena, one submit, scatter the s16 results.
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
    def __init__(self, n_sessions: int, k_inputs: int, in_rate: int = 44100, channels: int = 2, device: int = 0,
                 seed: int = 0, chunk_frames: int | None = None, fused: bool = True, out_frames: int = OUT_FRAMES, alloc_host: bool = True):
        """fused=True: one k_chain launch per tick (double-banked input, lagged recompute, no HBM intermediates);
        fused=False: the general unfused ops (k_resample -> device re-framing ring -> k_mix)."""
        self.fused = fused
        self.F = out_frames   # output_frame_size of the resampler nodes = frame_samples_per_channel of the clocked mixer
        self.S, self.K, self.C = n_sessions, k_inputs, channels
        self.in_rate = in_rate
        self.chunk = chunk_frames if chunk_frames is not None else in_rate // 50  # frames per 20 ms tick
        self.n_streams = n_sessions * k_inputs
        self.ctx = L.Context(device=device, max_streams=self.n_streams, max_channels=channels, fifo_frames=0 if fused else 2048)
        self.in_stride = _align(self.chunk * channels * 4, 16)
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
        slots = self.ctx.stream_open_many(in_rate, OUT_RATE, self.chunk, channels, self.n_streams)
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

    # algorithmic bytes: per session-tick K*(in + state r/w) + s16 out
    def algorithmic_bytes_per_tick(self) -> int:
        per_stream = self.in_stride + 2 * (8 + 16 * self.C * 4)
        return self.S * (self.K * per_stream + self.out_stride)

    def tick(self, inputs: np.ndarray, flags: int = 0) -> np.ndarray:
        """inputs: float32 [n_streams, chunk*channels]; returns int16 [n_sessions, 960*channels]"""
        x = np.ascontiguousarray(inputs, dtype=np.float32).reshape(self.n_streams, -1)
        self.host_in.reshape(self.n_streams, self.in_stride // 4)[:, : x.shape[1]] = x
        self.plan.submit(self.host_in, self.host_out, flags)
        self.plan.wait()
        return self.host_out.reshape(self.S, self.F * self.C).copy()

    def results(self) -> np.ndarray:
        return self.plan.download(self.res_off, self.res_bytes, L.CHAIN_RESULT_DT if self.fused else L.RS_RESULT_DT)

    def close(self):
        self.plan.destroy()
        self.ctx.close()


def run_chain_gpu(n_sessions: int, k_inputs: int, ticks: int, seed: int, in_rate: int = 44100, channels: int = 2,
                  graph: bool = False, fused: bool = True, chunk_frames: int | None = None, out_frames: int = OUT_FRAMES):
    ct = ChainTick(n_sessions, k_inputs, in_rate=in_rate, channels=channels, seed=seed, fused=fused, chunk_frames=chunk_frames,
                   out_frames=out_frames)
    try:
        outs = []
        for t in range(ticks):
            x = synth.tone_streams(seed, t, ct.n_streams, ct.chunk, channels, in_rate)
            outs.append(ct.tick(x, L.SUBMIT_GRAPH if graph else 0))
        return outs
    finally:
        ct.close()
