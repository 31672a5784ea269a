#!/usr/bin/env python
"""Per-kernel roofline numbers for BASELINE configs #2, #3, #4 (the standalone node kernels), device resident.
Working sets are scaled past the 126 MB L2 (stated per line) so the figures are HBM figures. Prints one JSON
line per kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak (MEASURED_PEAKS.json)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from streamkit_b200 import lib as L, synth  # noqa: E402

PEAK = 6547.2
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    PEAK = float(json.load(open(pp))["hbm_gbs"])
ITERS, WARM = 20, 3


def al(x, a=256):
    return (x + a - 1) // a * a


def timed(plan, ops):
    flags = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_TIME_OPS
    for _ in range(WARM):
        plan.submit(None, None, flags)
    plan.wait()
    plan.reset_op_times()
    for _ in range(ITERS):
        plan.submit(None, None, flags)
    plan.wait()
    return [plan.op_time(op, sub)[0] for op, sub in ops]


def report(name, cfg, algo_bytes, ms, extra=None):
    gbs = algo_bytes / (ms * 1e-3) / 1e9
    line = {"kernel": name, "config": cfg, "algorithmic_bytes": algo_bytes, "ms": ms, "achieved_gbs": gbs, "peak_gbs": PEAK,
            "frac": gbs / PEAK}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def bench_convert(ctx, mode, name, bps, sessions):
    N = 1920
    in_b = 2 if mode == L.CVT_S16_TO_F32 else 4
    out_b = 2 if mode == L.CVT_F32_TO_S16 else 4
    in_bytes = sessions * N * in_b
    out_off = al(in_bytes)
    plan = L.Plan(ctx, al(out_off + sessions * N * out_b))
    plan.set_gains(synth.gains(1, sessions))
    segs = np.zeros(sessions, dtype=L.SEG_DT)
    segs["in_off"] = np.arange(sessions, dtype=np.uint64) * (N * in_b)
    segs["out_off"] = out_off + np.arange(sessions, dtype=np.uint64) * (N * out_b)
    segs["n_samples"] = N
    segs["gain_idx"] = np.arange(sessions, dtype=np.uint32)
    op = plan.add_convert(mode, segs)
    plan.finalize()
    plan.fill(0, 0, in_bytes)
    (ms,) = timed(plan, [(op, 0)])
    report(name, "config #2 shape: %d sessions x 1920 samples (x%d of 4096, working set %.0f MB > L2)" % (
        sessions, sessions // 4096, (in_bytes + sessions * N * out_b) / 1e6), sessions * N * bps, ms,
        {"sessions_4096_equiv_us": ms * 1e3 * 4096 / sessions})
    plan.destroy()


def bench_mix(ctx, s16):
    G, K, N = 1024, 64, 1920
    in_bytes = G * K * N * 4
    out_off = al(in_bytes)
    ob = 2 if s16 else 4
    plan = L.Plan(ctx, al(out_off + G * N * ob))
    inputs = np.zeros(G * K, dtype=L.MIX_INPUT_DT)
    inputs["in_off"] = np.arange(G * K, dtype=np.uint64) * (N * 4)
    inputs["n_frames"] = N // 2
    inputs["channels"] = 2
    inputs["flags"] = L.MIX_IN_UNIQUE
    inputs["gain_idx"] = L.SKGPU_NO_GAIN
    groups = np.zeros(G, dtype=L.MIX_GROUP_DT)
    groups["out_off"] = out_off + np.arange(G, dtype=np.uint64) * (N * ob)
    groups["first_input"] = np.arange(G, dtype=np.uint32) * K
    groups["n_inputs"] = K
    groups["out_frames"] = N // 2
    groups["out_channels"] = 2
    groups["flags"] = L.MIX_OUT_S16 if s16 else 0
    groups["gain_idx"] = L.SKGPU_NO_GAIN
    op = plan.add_mix(groups, inputs)
    plan.finalize()
    plan.fill(0, 0, in_bytes)
    (ms,) = timed(plan, [(op, 0)])
    report("k_mix (%s out)" % ("s16" if s16 else "f32"), "config #3: 1024 groups x 64 stereo inputs x 960 frames (503 MB > L2)",
           G * (K * N * 4 + N * ob), ms)
    plan.destroy()


def bench_resample(in_rate, out_rate, chunk, streams):
    C = 2
    ctx = L.Context(device=0, max_streams=streams, max_channels=2, fifo_frames=0)
    slots = ctx.stream_open_many(in_rate, out_rate, chunk, C, streams)
    cap = L.Context.max_out_frames(in_rate, out_rate, chunk, C)
    in_stride, out_stride = al(chunk * C * 4, 16), al(cap * C * 4, 16)
    in_bytes = al(streams * in_stride)
    res_off = in_bytes
    out_off = al(res_off + 8 * streams)
    plan = L.Plan(ctx, al(out_off + streams * out_stride))
    items = np.zeros(streams, dtype=L.RS_ITEM_DT)
    items["in_off"] = np.arange(streams, dtype=np.uint64) * in_stride
    items["out_off"] = out_off + np.arange(streams, dtype=np.uint64) * out_stride
    items["slot"] = slots
    items["out_cap_frames"] = cap
    op = plan.add_resample(items, res_off)
    plan.finalize()
    plan.fill(0, 0, in_bytes)
    ph_ms, rs_ms = timed(plan, [(op, 0), (op, 1)])
    n_out = round(chunk * out_rate / in_rate)
    per_stream = chunk * C * 4 + n_out * C * 4 + 2 * (8 + 16 * C * 4)
    cfg = "config #4: %d stereo streams %d->%d Hz, chunk %d (in+out %.0f MB)" % (streams, in_rate, out_rate, chunk,
                                                                                 streams * (chunk + n_out) * C * 4 / 1e6)
    report("k_resample<2>", cfg, streams * per_stream, rs_ms, {"k_phase_ms": ph_ms})
    plan.destroy()
    ctx.close()


def main():
    ctx = L.Context(device=0, max_streams=16, max_channels=2)
    name, sms, *_ = ctx.device_info()
    print(json.dumps({"device": name, "sms": sms, "peak_gbs": PEAK, "iters": ITERS, "warmup": WARM}), flush=True)
    bench_convert(ctx, L.CVT_F32_TO_F32, "k_convert f32->f32 (audio::gain)", 8, 32768)
    bench_convert(ctx, L.CVT_F32_TO_S16, "k_convert f32->s16 (gain+clip+pack)", 6, 32768)
    bench_convert(ctx, L.CVT_S16_TO_F32, "k_convert s16->f32", 6, 32768)
    bench_mix(ctx, False)
    bench_mix(ctx, True)
    ctx.close()
    bench_resample(44100, 48000, 882, 16384)
    bench_resample(48000, 16000, 960, 16384)
    bench_resample(44100, 48000, 882, 65536)


if __name__ == "__main__":
    main()
