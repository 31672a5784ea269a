"""streamkit_b200 -- B200-native (sm_100a) batched PCM DSP hot path of StreamKit:
audio::resampler -> audio::mixer -> audio::gain -> f32/s16 conversion.

The product is the C-ABI shared library built from streamkit_b200/csrc (see include/skgpu_batch.h);
`streamkit_b200.lib` is a thin ctypes declaration of it for tests and benchmarks.
"""
__version__ = "0.1.0"
