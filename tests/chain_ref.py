"""Oracle composition of the full chain (test infrastructure): the same pipeline streamkit_b200.chain
builds on the GPU, expressed with the CPU restatement of the reference nodes (oracle/sko.py)."""
from __future__ import annotations

import collections

import numpy as np

from oracle import sko
from streamkit_b200 import synth

OUT_RATE = 48000
OUT_FRAMES = 960


def run_chain_oracle(n_sessions: int, k_inputs: int, ticks: int, seed: int, in_rate: int = 44100, channels: int = 2,
                     inputs_fn=None, chunk_frames: int | None = None, out_frames: int = OUT_FRAMES, out_rate: int = OUT_RATE,
                     node_chunk_frames: int | None = None):
    """node_chunk_frames: the resampler nodes' own chunk_frames when it differs from the frames delivered per tick (a tick
    that carries several node chunks, e.g. 60 ms ticks over 20 ms chunks)"""
    chunk = chunk_frames if chunk_frames is not None else in_rate // 50
    OUT_FRAMES = out_frames
    OUT_RATE = out_rate
    n_streams = n_sessions * k_inputs
    in_gains = synth.gains(seed, n_streams, 0.25, 1.5)
    master = synth.gains(seed + 1, n_sessions, 0.5, 2.0)
    nodes = [sko.ResamplerNode(OUT_RATE, chunk_frames=node_chunk_frames or chunk, output_frame_size=OUT_FRAMES) for _ in range(n_streams)]
    queues = [collections.deque() for _ in range(n_streams)]
    outs = []
    for t in range(ticks):
        x = (inputs_fn or synth.tone_streams)(seed, t, n_streams, chunk, channels, in_rate) if inputs_fn is None else inputs_fn(t)
        out = np.zeros((n_sessions, OUT_FRAMES * channels), dtype=np.int16)
        for s in range(n_streams):
            nodes[s].out.clear()
            nodes[s].push(in_rate, channels, x[s])
            for pkt in nodes[s].out:                     # audio::gain on every resampled packet (gain.rs:187-189)
                queues[s].append(sko.gain(pkt["samples"], float(in_gains[s])))
        for g in range(n_sessions):
            frames = []
            for i in range(k_inputs):
                q = queues[g * k_inputs + i]
                if q:                                     # clocked mixer pops <= 1 frame per input per tick (mixer.rs:1323-1353)
                    frames.append((q.popleft(), channels, True))
            mixed = sko.mix_clocked(frames, channels, OUT_FRAMES)      # zeros when nothing is present
            out[g] = sko.gain_f32_to_s16(mixed, float(master[g]))     # master gain -> clip -> s16
        outs.append(out)
    return outs


class MixedChainOracle:
    """sessions whose inputs differ in kind: per input index a sample rate (== out_rate: the resampler node's
    bypass, resampler.rs:299-373) and a wire format (s16 inputs are converted with x = s / 32768 first). Driven tick by
    tick with the same per-input arrays streamkit_b200.chain.ChainTick.tick takes."""

    def __init__(self, n_sessions, rates, s16, channels, seed, out_frames=OUT_FRAMES, out_rate=OUT_RATE):
        self.S, self.K, self.C, self.F = n_sessions, len(rates), channels, out_frames
        self.rates, self.s16, self.out_rate = list(rates), list(s16), out_rate
        n_streams = self.S * self.K
        self.in_gains = synth.gains(seed, n_streams, 0.25, 1.5)
        self.master = synth.gains(seed + 1, self.S, 0.5, 2.0)
        self.nodes = [[sko.ResamplerNode(out_rate, chunk_frames=r * out_frames // out_rate, output_frame_size=out_frames) for r in rates]
                      for _ in range(self.S)]
        self.queues = [[collections.deque() for _ in rates] for _ in range(self.S)]

    def tick(self, inputs, present=None):
        """inputs[i]: [n_sessions, chunk_i * channels]; present[s][i] False = the input delivers nothing this tick"""
        out = np.zeros((self.S, self.F * self.C), dtype=np.int16)
        for s in range(self.S):
            frames = []
            for i in range(self.K):
                if present is None or present[s][i]:
                    x = inputs[i][s]
                    if self.s16[i]:
                        x = sko.s16_to_f32(np.ascontiguousarray(x, dtype=np.int16))
                    n = self.nodes[s][i]
                    n.out.clear()
                    n.push(self.rates[i], self.C, np.ascontiguousarray(x, dtype=np.float32))
                    for pkt in n.out:
                        self.queues[s][i].append(sko.gain(pkt["samples"], float(self.in_gains[s * self.K + i])))
                q = self.queues[s][i]
                if q:
                    frames.append((q.popleft(), self.C, True))
            out[s] = sko.gain_f32_to_s16(sko.mix_clocked(frames, self.C, self.F), float(self.master[s]))
        return out
