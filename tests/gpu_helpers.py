"""Shared helpers of the -m gpu tests: thin drivers of the batch C ABI (ctypes -> libskgpu.so)."""
from __future__ import annotations

import numpy as np

from streamkit_b200 import lib as L


def al(x, a=256):
    return (x + a - 1) // a * a


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


class GpuResampler:
    """n_streams identical-configuration resampler streams driven chunk by chunk through a resample op"""

    def __init__(self, ctx, in_rate, out_rate, chunk, channels, n_streams=1, flags=0):
        self.ctx, self.chunk, self.channels, self.n = ctx, chunk, channels, n_streams
        self.slots = [ctx.stream_open(in_rate, out_rate, chunk, channels, flags) for _ in range(n_streams)]
        cap = L.Context.max_out_frames(in_rate, out_rate, chunk, channels)
        self.in_stride = al(chunk * channels * 4, 16)
        self.out_stride = al(cap * channels * 4, 16)
        self.in_bytes = al(n_streams * self.in_stride)
        self.res_off = self.in_bytes
        self.out_off = al(self.res_off + 8 * n_streams)
        self.total = al(self.out_off + n_streams * self.out_stride)
        self.plan = L.Plan(ctx, self.total)
        items = np.zeros(n_streams, dtype=L.RS_ITEM_DT)
        items["in_off"] = np.arange(n_streams) * self.in_stride
        items["out_off"] = self.out_off + np.arange(n_streams) * self.out_stride
        items["slot"] = self.slots
        items["out_cap_frames"] = cap
        self.plan.add_resample(items, self.res_off)
        self.plan.set_io(0, self.in_bytes, self.res_off, self.total - self.res_off)
        self.plan.finalize()
        self.host_in = np.zeros(self.in_bytes, np.uint8)
        self.host_out = np.zeros(self.total - self.res_off, np.uint8)
        self.ticks = 0

    def process(self, x):
        """x: [n_streams, chunk*channels] f32 (or 1-D for one stream); returns list of per-stream outputs"""
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(self.n, -1)
        for s in range(self.n):
            self.host_in[s * self.in_stride: s * self.in_stride + x.shape[1] * 4] = x[s].view(np.uint8)
        self.plan.submit(self.host_in, self.host_out, L.SUBMIT_GRAPH if self.ticks % 2 else 0)
        self.plan.wait()
        self.ticks += 1
        res = self.host_out[: 8 * self.n].view(L.RS_RESULT_DT)
        outs = []
        for s in range(self.n):
            assert res[s]["status"] == 0
            n = int(res[s]["out_frames"])
            o = (self.out_off - self.res_off) + s * self.out_stride
            outs.append(self.host_out[o: o + n * self.channels * 4].view(np.float32).copy())
        return outs

    def counts_only(self):
        """one more chunk of whatever is in host_in; returns out_frames per stream"""
        self.plan.submit(self.host_in, self.host_out, L.SUBMIT_GRAPH)
        self.plan.wait()
        self.ticks += 1
        return self.host_out[: 8 * self.n].view(L.RS_RESULT_DT)["out_frames"].copy()

    def state(self, s=0):
        return self.ctx.stream_state(self.slots[s], self.channels)

    def close(self):
        self.plan.destroy()
        for sl in self.slots:
            self.ctx.stream_close(sl)
