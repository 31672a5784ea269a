"""Host-side builders of the standalone-node workloads (BASELINE.json configs[1..3]) on top of the batch C ABI.
Shared by bench.py (--config 2|3|4) and tools/bench_kernels.py. Device working sets are larger than the 126 MB L2
(stated per workload) so the timings are HBM timings without an explicit flush.

  config #2  audio::gain + f32<->s16 over 4,096 sessions x 1,920 samples  -> k_convert  (x ROT rotated tick buffers)
  config #3  audio::mixer, 1,024 groups x 64 stereo inputs (+ clip/s16)    -> k_mix
  config #4  audio::resampler, 16,384 stereo streams, both ratios          -> k_phase_prog + k_resample_prog
"""
from __future__ import annotations

import numpy as np

from . import lib as L
from . import synth


def al(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


class Workload:
    """common surface: plan + dominant op, algorithmic bytes per launch, host buffers for the end-to-end leg"""
    name = ""
    unit = ""
    kernel = ""

    def close(self):
        self.plan.destroy()
        if getattr(self, "own_ctx", False):
            self.ctx.close()


class GainS16(Workload):
    """config #2: per session y = sat_s16(rint((x * g) * 32768)) (gain.rs:184-190 + the build-defined s16 packing), ROT
    ticks' worth of sessions per launch so that the working set (ROT x 47 MB) exceeds L2."""
    unit = "sessions"

    def __init__(self, ctx, mode=L.CVT_F32_TO_S16, sessions=4096, rot=8, seed=0):
        self.ctx, self.mode, self.S, self.rot, self.N = ctx, mode, sessions, rot, 1920
        n = sessions * rot
        self.in_b = 2 if mode == L.CVT_S16_TO_F32 else 4
        self.out_b = 2 if mode == L.CVT_F32_TO_S16 else 4
        self.bps = self.in_b + self.out_b
        self.kernel = {L.CVT_F32_TO_F32: "k_convert<f32->f32>", L.CVT_F32_TO_S16: "k_convert<f32->s16>", L.CVT_S16_TO_F32: "k_convert<s16->f32>"}[mode]
        self.in_bytes = n * self.N * self.in_b
        self.out_off = al(self.in_bytes)
        self.out_bytes = n * self.N * self.out_b
        self.plan = L.Plan(ctx, al(self.out_off + self.out_bytes))
        self.gains = synth.gains(seed + 1, n)
        self.plan.set_gains(self.gains)
        segs = np.zeros(n, dtype=L.SEG_DT)
        segs["in_off"] = np.arange(n, dtype=np.uint64) * (self.N * self.in_b)
        segs["out_off"] = self.out_off + np.arange(n, dtype=np.uint64) * (self.N * self.out_b)
        segs["n_samples"] = self.N
        segs["gain_idx"] = np.arange(n, dtype=np.uint32)
        self.op = self.plan.add_convert(mode, segs)
        self.plan.set_io(0, self.in_bytes, self.out_off, self.out_bytes)
        self.plan.finalize()
        self.units_per_launch = n
        self.algorithmic_bytes = n * self.N * self.bps
        self.name = "BASELINE configs[1]: gain + f32<->s16 over %d concurrent 48 kHz stereo sessions, 20 ms frames (%s); %d rotated tick buffers per launch (working set %.0f MB > L2)" % (
            sessions, self.kernel, rot, (self.in_bytes + self.out_bytes) / 1e6)
        self.ops = [(self.op, 0, self.kernel)]

    def fill_host(self, host_in: np.ndarray, seed=7):
        P = 2048
        n = self.S * self.rot
        if self.mode == L.CVT_S16_TO_F32:
            self.pool = np.random.default_rng(seed).integers(-32768, 32768, (P, self.N), dtype=np.int16)
            v = host_in.view(np.int16)[: n * self.N].reshape(n, self.N)
        else:
            self.pool = synth.uniform_pcm(seed, P * self.N, over_range_frac=0.02).reshape(P, self.N)
            v = host_in.view(np.float32)[: n * self.N].reshape(n, self.N)
        for b in range(0, n, P):
            m = min(P, n - b)
            v[b:b + m] = self.pool[:m]


class Mix64(Workload):
    """config #3: 1,024 mix groups x 64 stereo inputs of 960 frames, reference summation order, optional gain+clip+s16 epilogue"""
    unit = "mix groups"
    kernel = "k_mix"

    def __init__(self, ctx, s16=True, groups=1024, k=64, seed=0):
        self.ctx, self.G, self.K, self.N, self.s16 = ctx, groups, k, 1920, s16
        self.in_bytes = groups * k * self.N * 4
        self.out_off = al(self.in_bytes)
        ob = 2 if s16 else 4
        self.out_bytes = groups * self.N * ob
        self.plan = L.Plan(ctx, al(self.out_off + self.out_bytes))
        inputs = np.zeros(groups * k, dtype=L.MIX_INPUT_DT)
        inputs["in_off"] = np.arange(groups * k, dtype=np.uint64) * (self.N * 4)
        inputs["n_frames"] = self.N // 2
        inputs["channels"] = 2
        inputs["flags"] = L.MIX_IN_UNIQUE
        inputs["gain_idx"] = L.SKGPU_NO_GAIN
        grp = np.zeros(groups, dtype=L.MIX_GROUP_DT)
        grp["out_off"] = self.out_off + np.arange(groups, dtype=np.uint64) * (self.N * ob)
        grp["first_input"] = np.arange(groups, dtype=np.uint32) * k
        grp["n_inputs"] = k
        grp["out_frames"] = self.N // 2
        grp["out_channels"] = 2
        grp["flags"] = L.MIX_OUT_S16 if s16 else 0
        grp["gain_idx"] = np.arange(groups, dtype=np.uint32) if s16 else L.SKGPU_NO_GAIN
        self.master = synth.gains(seed + 2, groups, 0.01, 0.06)      # 64 full-scale inputs sum to +-64: keep some of the output unclipped
        if s16:
            self.plan.set_gains(self.master)
        self.op = self.plan.add_mix(grp, inputs)
        self.plan.set_io(0, self.in_bytes, self.out_off, self.out_bytes)
        self.plan.finalize()
        self.units_per_launch = groups
        self.algorithmic_bytes = groups * (k * self.N * 4 + self.N * ob)
        self.name = "BASELINE configs[2]: %d-input mixer (with clip%s) x %d mix groups per tick, stereo 960-frame packets (inputs %.0f MB > L2)" % (
            k, " + s16 pack" if s16 else "", groups, self.in_bytes / 1e6)
        self.ops = [(self.op, 0, self.kernel)]

    def fill_host(self, host_in: np.ndarray, seed=5):
        P = 2048
        pool = (np.random.default_rng(seed).random((P, self.N), dtype=np.float32) * 2 - 1).astype(np.float32)
        self.pool = pool
        self.idx = (np.arange(self.G)[:, None] * 37 + np.arange(self.K)[None, :] * 101) % P
        host_in.view(np.float32)[: self.G * self.K * self.N].reshape(self.G * self.K, self.N)[:] = pool[self.idx.reshape(-1)]


class Resample(Workload):
    """config #4: N stereo streams, one 20 ms chunk each per launch, output_frame_size = 0 (variable-length output)"""
    unit = "streams"
    kernel = "k_resample_prog<2>"

    def __init__(self, in_rate=44100, out_rate=48000, streams=16384, device=0, sinc=None):
        """sinc: None = rubato Linear (the reference's mode); (sinc_len, oversampling, f_cutoff) = the windowed-sinc polyphase mode"""
        self.own_ctx = True
        self.in_rate, self.out_rate, self.S, self.C = in_rate, out_rate, streams, 2
        self.chunk = in_rate // 50
        self.sinc = sinc
        self.ctx = L.Context(device=device, max_streams=streams, max_channels=2, fifo_frames=0)
        if sinc:
            self.ctx.set_sinc(*sinc)
            self.kernel = "k_resample_sinc_tiled<2>"
        slots = self.ctx.stream_open_many(in_rate, out_rate, self.chunk, self.C, streams, L.STREAM_SINC if sinc else 0)
        self.cap = L.Context.max_out_frames(in_rate, out_rate, self.chunk, self.C)
        self.in_stride, self.out_stride = al(self.chunk * self.C * 4, 16), al(self.cap * self.C * 4, 16)
        self.in_bytes = al(streams * self.in_stride)
        self.res_off = self.in_bytes
        self.out_off = al(self.res_off + 8 * streams)
        self.out_bytes = al(self.out_off + streams * self.out_stride) - self.res_off
        self.plan = L.Plan(self.ctx, self.res_off + self.out_bytes)
        items = np.zeros(streams, dtype=L.RS_ITEM_DT)
        items["in_off"] = np.arange(streams, dtype=np.uint64) * self.in_stride
        items["out_off"] = self.out_off + np.arange(streams, dtype=np.uint64) * self.out_stride
        items["slot"] = slots
        items["out_cap_frames"] = self.cap
        self.op = self.plan.add_resample(items, self.res_off)
        self.plan.set_io(0, self.in_bytes, self.res_off, self.out_bytes)
        self.plan.finalize()
        self.n_out = round(self.chunk * out_rate / in_rate)
        self.units_per_launch = streams
        # in + out + state read/write (last_index 8 B + 16-frame history 128 B, each way): BASELINE.md / SURVEY 8d
        self.algorithmic_bytes = streams * (self.chunk * self.C * 4 + self.n_out * self.C * 4 + 2 * (8 + 16 * self.C * 4))
        self.name = "BASELINE configs[3]: batched resampling %d->%d Hz over %d stereo streams, 20 ms chunks (in+out %.0f MB > L2)" % (
            in_rate, out_rate, streams, streams * (self.chunk + self.n_out) * self.C * 4 / 1e6)
        self.ops = [(self.op, 1, self.kernel), (self.op, 0, "k_phase" if sinc else "k_phase_prog")]
        if sinc:
            hist = (sinc[0] + 8) * self.C * 4
            self.algorithmic_bytes = streams * (self.chunk * self.C * 4 + self.n_out * self.C * 4 + 2 * (8 + hist))
            self.fma_per_launch = streams * self.n_out * self.C * 2 * sinc[0]       # two tap rows x L taps per channel and output frame
            self.name += " -- windowed-sinc polyphase mode (sinc_len %d, oversampling %d, cutoff %.2f, BlackmanHarris2, linear phase interpolation)" % sinc

    def fill_host(self, host_in: np.ndarray, seed=9):
        D = 256
        self.base = synth.tone_streams(seed, 0, D, self.chunk, self.C, self.in_rate)
        v = host_in.view(np.float32)[: self.S * self.in_stride // 4].reshape(self.S, self.in_stride // 4)
        for b in range(0, self.S, D):
            m = min(D, self.S - b)
            v[b:b + m, : self.base.shape[1]] = self.base[:m]
