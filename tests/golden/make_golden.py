"""Generates the committed golden vectors. Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference only for the WAV fixture (crates/nodes/testdata/audio/sample.wav, decoded here with the
stdlib `wave` module); everything else comes from the oracle. At generation time the C oracle and the numpy
restatement must agree bit for bit, otherwise nothing is written."""
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import np_oracle, sko  # noqa: E402
from streamkit_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
bits = lambda a: np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)

pcm = synth.uniform_pcm(1234, 1920, over_range_frac=0.05)
g = np.array([1.7], np.float32)
edge = np.array([0.0, -0.0, 1.0, -1.0, 0.999969482421875, 1.0000001, -1.0000001, 0.5 / 32768, 1.5 / 32768, 2.5 / 32768,
                 -0.5 / 32768, -1.5 / 32768, 32766.5 / 32768, 32767.5 / 32768, -32768.5 / 32768, 3.9, -3.9, 1e-40, -1e-40,
                 np.inf, -np.inf, np.nan, 1e30, -1e30], dtype=np.float32)
rng = np.random.default_rng(99)
mix_in = (rng.random((16, 1920), dtype=np.float32) * 2 - 1).astype(np.float32)
frames = [(mix_in[i], 2, True) for i in range(16)]
mix_out = sko.mix_clocked(frames, 2, 960)
assert np.array_equal(bits(mix_out), bits(np_oracle.mix(frames, 2, 1920)))

rs_in = np.stack([synth.tone_streams(55, c, 1, 882, 2, 44100)[0] for c in range(4)])
a, b = sko.FastFixedIn(44100, 48000, 882, 2), np_oracle.FastFixedIn(44100, 48000, 882, 2)
ya = np.concatenate([a.process(x) for x in rs_in]); yb = np.concatenate([b.process(x) for x in rs_in])
assert np.array_equal(bits(ya), bits(yb)) and a.last_index == b.last_index
rs2_in = np.stack([synth.tone_streams(56, c, 1, 960, 1, 48000)[0] for c in range(3)])
a2, b2 = sko.FastFixedIn(48000, 16000, 960, 1), np_oracle.FastFixedIn(48000, 16000, 960, 1)
y2a = np.concatenate([a2.process(x) for x in rs2_in]); y2b = np.concatenate([b2.process(x) for x in rs2_in])
assert np.array_equal(bits(y2a), bits(y2b))
assert np.array_equal(sko.f32_to_s16(edge), np_oracle.f32_to_s16(edge))

np.savez_compressed(os.path.join(HERE, "hotpath_v1.npz"), pcm=pcm, gain=g, gain_out_bits=bits(sko.gain(pcm, g[0])),
                    gain_s16=sko.gain_f32_to_s16(pcm, g[0]), edge=edge, edge_s16=sko.f32_to_s16(edge), mix_in=mix_in,
                    mix_out_bits=bits(mix_out), rs_in=rs_in, rs_out_bits=bits(ya), rs_last_index=np.array([a.last_index]),
                    rs2_in=rs2_in, rs2_out_bits=bits(y2a))

wav = "/root/reference/crates/nodes/testdata/audio/sample.wav"
with wave.open(wav, "rb") as w:
    assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 2, 48000, 4800)
    s = np.frombuffer(w.readframes(4800), dtype="<i2").reshape(4800, 2).copy()
np.save(os.path.join(HERE, "config1_input_s16.npy"), s)
f = sko.s16_to_f32(s.reshape(-1)).reshape(4800, 2)
outs = []
for ch in range(2):
    n = sko.ResamplerNode(16000, 960, 960)
    for c in range(5):
        n.push(48000, 1, f[c * 960:(c + 1) * 960, ch])
    n.finish()
    outs.append(sko.gain_f32_to_s16(np.concatenate([p["samples"] for p in n.out]), 2.0))
np.save(os.path.join(HERE, "config1_output_s16.npy"), np.stack(outs))
print("golden vectors written:", os.listdir(HERE))
