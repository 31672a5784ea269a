"""GPU parity at BASELINE.json's FULL sizes (VERDICT r1 next #1c): the persistent 592-CTA schedule of k_chain, its block
rotation and prefetch ring only run at full occupancy, so the small-S tests do not cover them.

  config #5  65,536 sessions x 2 stereo 44.1 kHz inputs, fused chain, 4 ticks: every s16 byte of the last tick against the
             multi-threaded CPU chain (oracle/sk_chain.c, out_last)
  config #3  1,024 mix groups x 64 full-scale stereo inputs, f32 and gain+clip+s16 epilogue, bit-exact vs the oracle
  config #4  16,384 resampler streams, 44.1k->48k and 48k->16k, 3 chunks: counts + every sample vs the oracle
"""
import os

import numpy as np
import pytest

from oracle import sko
from streamkit_b200 import chain, lib as L, synth
from tests.gpu_helpers import al, bits

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("S,K,T", [(65536, 2, 4), (50000, 1, 3)])
def test_config5_full_size_fused_chain_every_byte(S, K, T):
    P = 4096                                              # distinct input streams; stream s plays pool[s % P] every tick
    ct = chain.ChainTick(S, K, seed=17)
    try:
        pool = synth.noise_streams(99, 0, P, ct.chunk, 2, amp=0.9)
        hin = ct.host_in.reshape(ct.n_streams, ct.in_stride // 4)
        for b in range(0, ct.n_streams, P):
            n = min(P, ct.n_streams - b)
            hin[b:b + n, : pool.shape[1]] = pool[:n]
        for _ in range(T):
            ct.plan.submit(ct.host_in, ct.host_out, L.SUBMIT_GRAPH)
            ct.plan.wait()
        got = ct.host_out.reshape(S, 960 * 2).copy()
        res = ct.results()
        assert np.all(res["status"] == 0) and np.all(res["emitted"] == 1)
        threads = len(os.sched_getaffinity(0))
        _sec, _cs, want = sko.chain_bench(S, K, T, 44100, 2, pool, ct.in_gains, ct.master_gains, threads, want_last=True)
        diff = np.flatnonzero(np.any(got != want, axis=1))
        assert diff.size == 0, f"{diff.size} of {S} sessions differ from the CPU chain (first: {diff[:8]})"
        assert np.count_nonzero(got) > got.size // 2
    finally:
        ct.close()


@pytest.mark.parametrize("s16", [False, True])
def test_config3_full_size_1024_groups_x_64_inputs(s16):
    G, K, N = 1024, 64, 1920
    P = 2048
    rng = np.random.default_rng(5)
    pool = ((rng.random((P, N), dtype=np.float32) * 2 - 1)).astype(np.float32)        # full scale: only the reference ORDER gives the bits
    idx = (np.arange(G)[:, None] * 37 + np.arange(K)[None, :] * 101) % P             # input (g, i) plays pool[idx[g, i]]
    in_gains = synth.gains(3, G * K, 0.0, 2.0)
    master = synth.gains(4, G, 0.1, 1.5)
    ctx = L.Context(device=0, max_streams=1, max_channels=2, fifo_frames=0)
    try:
        in_stride = N * 4
        in_bytes = al(G * K * in_stride)
        ob = 2 if s16 else 4
        out_stride = N * ob
        plan = L.Plan(ctx, in_bytes + al(G * out_stride))
        inputs = np.zeros(G * K, dtype=L.MIX_INPUT_DT)
        inputs["in_off"] = np.arange(G * K, dtype=np.uint64) * in_stride
        inputs["n_frames"] = N // 2
        inputs["channels"] = 2
        inputs["flags"] = L.MIX_IN_UNIQUE
        inputs["gain_idx"] = np.arange(G * K) if s16 else L.SKGPU_NO_GAIN
        groups = np.zeros(G, dtype=L.MIX_GROUP_DT)
        groups["out_off"] = in_bytes + np.arange(G, dtype=np.uint64) * out_stride
        groups["first_input"] = np.arange(G) * K
        groups["n_inputs"] = K
        groups["out_frames"] = N // 2
        groups["out_channels"] = 2
        groups["flags"] = L.MIX_OUT_S16 if s16 else 0
        groups["gain_idx"] = (G * K + np.arange(G)) if s16 else L.SKGPU_NO_GAIN
        if s16:
            plan.set_gains(np.concatenate([in_gains, master]))
        plan.add_mix(groups, inputs)
        plan.set_io(0, in_bytes, in_bytes, al(G * out_stride))
        plan.finalize()
        hin = ctx.pinned(in_bytes, np.float32)
        hin[: G * K * N].reshape(G * K, N)[:] = pool[idx.reshape(-1)]
        hout = ctx.pinned(al(G * out_stride), np.uint8)
        plan.submit(hin, hout)
        plan.wait()
        got = hout[: G * out_stride].view(np.int16 if s16 else np.uint32).reshape(G, N)
        for g in range(G):
            if s16:
                frames = [(sko.gain(pool[idx[g, i]], in_gains[g * K + i]), 2, True) for i in range(K)]
                want = sko.gain_f32_to_s16(sko.mix_clocked(frames, 2, N // 2), master[g])
            else:
                want = bits(sko.mix_clocked([(pool[idx[g, i]], 2, True) for i in range(K)], 2, N // 2))
            assert np.array_equal(got[g], want), f"group {g}"
        plan.destroy()
    finally:
        ctx.close()


@pytest.mark.parametrize("in_rate,out_rate,chunk", [(44100, 48000, 882), (48000, 16000, 960)])
def test_config4_full_size_16384_streams(in_rate, out_rate, chunk):
    S, C, D = 16384, 2, 128            # D distinct streams tiled over the S slots
    ctx = L.Context(device=0, max_streams=S, max_channels=2, fifo_frames=0)
    try:
        slots = ctx.stream_open_many(in_rate, out_rate, chunk, C, S)
        cap = L.Context.max_out_frames(in_rate, out_rate, chunk, C)
        in_stride, out_stride = chunk * C * 4, al(cap * C * 4, 16)
        in_bytes = al(S * in_stride)
        res_off = in_bytes
        out_off = al(res_off + 8 * S)
        total = al(out_off + S * out_stride)
        plan = L.Plan(ctx, total)
        items = np.zeros(S, dtype=L.RS_ITEM_DT)
        items["in_off"] = np.arange(S, dtype=np.uint64) * in_stride
        items["out_off"] = out_off + np.arange(S, dtype=np.uint64) * out_stride
        items["slot"] = slots
        items["out_cap_frames"] = cap
        plan.add_resample(items, res_off)
        plan.set_io(0, in_bytes, res_off, total - res_off)
        plan.finalize()
        hin = ctx.pinned(in_bytes, np.float32)
        hout = ctx.pinned(total - res_off, np.uint8)
        refs = [sko.FastFixedIn(in_rate, out_rate, chunk, C) for _ in range(D)]
        for tick in range(3):
            base = synth.tone_streams(5 + tick, tick, D, chunk, C, in_rate)
            hin[: S * chunk * C].reshape(S // D, D, chunk * C)[:] = base[None]
            plan.submit(hin, hout, L.SUBMIT_GRAPH if tick else 0)
            plan.wait()
            res = hout[: 8 * S].view(L.RS_RESULT_DT)
            assert np.all(res["status"] == 0)
            outs = hout[out_off - res_off: out_off - res_off + S * out_stride].reshape(S // D, D, out_stride)
            for i in range(D):
                want = refs[i].process(base[i])
                assert np.all(res["out_frames"][i::D] == want.size // C)
                blk = np.ascontiguousarray(outs[:, i, : want.size * 4]).view(np.uint32)
                assert np.array_equal(blk, np.broadcast_to(bits(want), blk.shape)), f"tick {tick} stream class {i}"
        plan.destroy()
    finally:
        ctx.close()
