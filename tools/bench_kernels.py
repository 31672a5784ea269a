#!/usr/bin/env python
"""Per-kernel roofline numbers for BASELINE configs #2, #3, #4 (the standalone node kernels), device resident.
Working sets are larger than the 126 MB L2 (stated per line) so the figures are HBM figures. Prints one JSON line per
kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak (MEASURED_PEAKS.json). The driver-facing
measurement of the same workloads is `bench.py --config 2|3|4`; this tool adds the variants (f32->f32, s16->f32,
f32 mixer output, 48k->16k, 65,536 streams)."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from streamkit_b200 import lib as L, workloads as W  # noqa: E402

PEAK = 6547.2
pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pp):
    PEAK = float(json.load(open(pp))["hbm_gbs"])
ITERS, WARM = (1, 1) if os.environ.get("SK_PROFILE") else (20, 3)      # SK_PROFILE: two launches per workload, for an ncu capture


def run(w):
    flags = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_TIME_OPS
    # real signal in the arena (seeded noise / tones, uploaded once): all-zero inputs time a few percent faster on this part
    host_in = w.ctx.pinned(w.in_bytes, np.uint8)
    w.fill_host(host_in)
    w.plan.submit(host_in, None, L.SUBMIT_NO_D2H)
    w.plan.wait()
    for _ in range(WARM):
        w.plan.submit(None, None, flags)
    w.plan.wait()
    w.plan.reset_op_times()
    for _ in range(ITERS):
        w.plan.submit(None, None, flags)
    w.plan.wait()
    times = {name: w.plan.op_time(op, sub)[0] for op, sub, name in w.ops}
    ms = times[w.ops[0][2]]
    gbs = w.algorithmic_bytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": w.ops[0][2], "config": w.name, "algorithmic_bytes": w.algorithmic_bytes, "ms": ms, "achieved_gbs": gbs,
                      "peak_gbs": PEAK, "frac": gbs / PEAK, "kernels_ms": times}), flush=True)
    w.close()


def main():
    if os.environ.get("SK_ONLY") == "sinc":      # profiling helper: the sinc workloads alone
        run(W.Resample(44100, 48000, 16384, sinc=(64, 256, 0.95)))
        run(W.Resample(48000, 16000, 16384, sinc=(64, 256, 0.95)))
        return
    ctx = L.Context(device=0, max_streams=16, max_channels=2)
    if os.environ.get("SK_ONLY") == "mix":
        run(W.Mix64(ctx, s16=True))
        ctx.close()
        return
    name, sms, *_ = ctx.device_info()
    print(json.dumps({"device": name, "sms": sms, "peak_gbs": PEAK, "iters": ITERS, "warmup": WARM}), flush=True)
    for mode in (L.CVT_F32_TO_F32, L.CVT_F32_TO_S16, L.CVT_S16_TO_F32):
        run(W.GainS16(ctx, mode))
    run(W.Mix64(ctx, s16=False))
    run(W.Mix64(ctx, s16=True))
    ctx.close()
    run(W.Resample(44100, 48000, 16384))
    run(W.Resample(48000, 16000, 16384))
    run(W.Resample(44100, 48000, 65536))
    run(W.Resample(44100, 48000, 16384, sinc=(64, 256, 0.95)))
    run(W.Resample(48000, 16000, 16384, sinc=(64, 256, 0.95)))


if __name__ == "__main__":
    main()
