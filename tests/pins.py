"""Pins of the resampler that do NOT depend on the restated rubato source (VERDICT r1, next #1a).

rubato 0.16.2 is not under /root/reference, so `oracle/sk_oracle.c` restates FastFixedIn/Linear from its published
algorithm. The checks here are properties of *what that resampler is* -- a linear interpolator that reads the input
at positions -4 + (k+1) * in_rate/out_rate -- derived with exact rational arithmetic, not from the restatement:

  identity   integer ratios (t = 2, 3, 6): every fraction is exactly 0, so output k IS input sample -4 + (k+1) t
             (zeros before the stream starts), bit for bit;
  ramp       x[i] = a i + b in, a * pos_k + b out (linear interpolation reproduces a straight line); pos_k from
             Fractions; tolerance = a few f32 ulps of the value;
  counts     frames emitted after m chunks = #{j >= 0 : -4 + j t < m N + N - 9 - ceil(t)} (closed form over Fractions);
  lengths    the reference's own length asserts (resampler.rs:826-837, :901-906).

The same functions run against the C oracle (tests/test_pins.py, CPU) and against the CUDA kernels through the C ABI
(tests/test_gpu_pins.py): `resample(chunks) -> list of per-chunk interleaved outputs`.
"""
from __future__ import annotations

import math
from fractions import Fraction

import numpy as np


def exact_positions(in_rate: int, out_rate: int, n: int):
    """input position of output k (k = 0..n-1) relative to the first input sample, exact"""
    t = Fraction(in_rate, out_rate)
    return [Fraction(-4) + (k + 1) * t for k in range(n)]


def exact_total_after(in_rate: int, out_rate: int, chunk: int, m_chunks: int) -> int:
    """frames emitted by the first m_chunks process() calls (exact arithmetic; valid while no position lies exactly
    on a chunk's end_idx boundary, which the callers assert)"""
    t = Fraction(in_rate, out_rate)
    end_idx = chunk - 9 - math.ceil(t)
    bound = Fraction((m_chunks - 1) * chunk + end_idx + 4)   # emits j-th position (j >= 0, value -4 + j t) while it is < the bound
    q = bound / t
    assert q.denominator != 1, "tie: the exact count is ambiguous for this configuration"
    return math.ceil(q)


def check_integer_ratio_identity(resample, in_rate: int, out_rate: int, chunk: int, channels: int, n_chunks: int = 4):
    assert in_rate % out_rate == 0
    t = in_rate // out_rate
    rng = np.random.default_rng(in_rate + chunk)
    x = (rng.random((n_chunks * chunk, channels), dtype=np.float32) - np.float32(0.5)).astype(np.float32)
    x[x == 0] = np.float32(0.25)   # keep signed zeros out (y0 + 0*y1 of -0.0 is +0.0: equal, not identical)
    outs = resample([x[c * chunk:(c + 1) * chunk].reshape(-1) for c in range(n_chunks)])
    y = np.concatenate(outs).reshape(-1, channels)
    pos = -4 + (np.arange(y.shape[0]) + 1) * t
    want = np.where((pos >= 0)[:, None], x[np.clip(pos, 0, x.shape[0] - 1)], np.float32(0.0))
    assert y.shape[0] == exact_total_after(in_rate, out_rate, chunk, n_chunks)
    assert pos[-1] < x.shape[0]
    assert np.array_equal(y.view(np.uint32), want.astype(np.float32).view(np.uint32)), "integer ratio: outputs must BE input samples"


def check_ramp(resample, in_rate: int, out_rate: int, chunk: int, channels: int, n_chunks: int = 6, ulps: float = 3.0):
    a = [np.float32(1.0 / 1024.0) * (c + 1) for c in range(channels)]
    b = [np.float32(0.125) * c for c in range(channels)]
    i = np.arange(n_chunks * chunk, dtype=np.float64)
    x = np.stack([(float(a[c]) * i + float(b[c])) for c in range(channels)], axis=1).astype(np.float32)   # exact in f32 (small dyadic values)
    assert np.array_equal(x.astype(np.float64), np.stack([(float(a[c]) * i + float(b[c])) for c in range(channels)], axis=1))
    outs = resample([x[c * chunk:(c + 1) * chunk].reshape(-1) for c in range(n_chunks)])
    y = np.concatenate(outs).reshape(-1, channels)
    pos = exact_positions(in_rate, out_rate, y.shape[0])
    first = next(k for k, p in enumerate(pos) if p >= 0)   # before that the interpolator reads the zero history
    for c in range(channels):
        want = np.array([float(a[c]) * float(p) + float(b[c]) for p in pos[first:]])
        got = y[first:, c].astype(np.float64)
        tol = ulps * np.spacing(np.abs(want).astype(np.float32)).astype(np.float64) + 1e-12   # + the f64 phase's own rounding (|a| * 1e-13)
        bad = np.abs(got - want) > tol
        assert not bad.any(), f"ramp: {bad.sum()} outputs off the line (first at {first + int(np.argmax(bad))})"
    return y.shape[0]


def check_counts(resample_counts, in_rate: int, out_rate: int, chunk: int, n_chunks: int):
    """resample_counts(n_chunks) -> per-chunk output frame counts"""
    counts = np.asarray(resample_counts(n_chunks), dtype=np.int64)
    tot = np.cumsum(counts)
    for m in sorted({1, 2, 3, 7, n_chunks // 3, n_chunks // 2, n_chunks - 1, n_chunks}):
        if m >= 1:
            assert int(tot[m - 1]) == exact_total_after(in_rate, out_rate, chunk, m), f"total frames after {m} chunks"
    # every cumulative total, vectorised with integers: ceil(bound * out / in)
    t_num, t_den = Fraction(in_rate, out_rate).numerator, Fraction(in_rate, out_rate).denominator
    end_idx = chunk - 9 - math.ceil(Fraction(in_rate, out_rate))
    m = np.arange(1, n_chunks + 1, dtype=np.int64)
    num = ((m - 1) * chunk + end_idx + 4) * t_den
    assert not np.any(num % t_num == 0)
    assert np.array_equal(tot, -(-num // t_num))
