#!/usr/bin/env python
"""Device-resident kernel times of the fused chain for sessions that mix a resampled (44.1 kHz) and a rate-equal (48 kHz) input --
the CHAIN_F32 instantiation -- next to the all-resampled shape (CHAIN_PLAIN). 65,536 sessions x 2 stereo inputs."""
import sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from streamkit_b200 import chain, lib as L, synth
for rates in ([44100, 48000], [44100, 44100]):
    ct = chain.ChainTick(65536, 2, in_rate=rates, channels=2, seed=0)
    xs = [synth.noise_streams(7 + i, 0, ct.S, c, 2) for i, c in enumerate(ct.chunks)]
    ct.fill_rows(ct.host_in, xs)
    fl = L.SUBMIT_NO_D2H
    ct.plan.submit(ct.host_in, None, fl); ct.plan.submit(ct.host_in, None, fl); ct.plan.wait()
    dev = L.SUBMIT_NO_H2D | L.SUBMIT_NO_D2H | L.SUBMIT_TIME_OPS
    for _ in range(5): ct.plan.submit(None, None, dev)
    ct.plan.wait(); ct.plan.reset_op_times()
    for _ in range(30): ct.plan.submit(None, None, dev)
    ct.plan.wait()
    print(rates, 'phase', round(ct.plan.op_time(ct.op_chain, 0)[0], 4), 'chain', round(ct.plan.op_time(ct.op_chain, 1)[0], 4))
    ct.close()
