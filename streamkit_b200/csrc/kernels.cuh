// kernels.cuh -- hand-written sm_100a kernels for StreamKit's PCM DSP hot path (umbrella include).
//
//   k_convert        gain / f32->s16 / s16->f32 over frame segments        (gain.rs:184-190)
//   k_phase          per-stream rubato phase recurrence -> phase table      (rubato FastFixedIn, resampler.rs:404-407)
//   k_resample       linear interpolation from TMA-staged chunk + history   (rubato interp_lin; resampler.rs:397-417)
//   k_mix            ordered N-input sum, channel conversion, epilogue      (mixer.rs:944-1013, :1027-1078)
//   k_fifo_commit    advances device re-framing ring read cursors           (resampler.rs:420-458 re-framing, unfused path)
//   k_chain          all of the above fused per session, lagged recompute   (BASELINE config #5)
//   k_resample_sinc  windowed-sinc polyphase mode (north star; own spec)    (skgpu_ctx_set_sinc)
#pragma once
#include "common.cuh"
#include "k_convert.cuh"
#include "k_resample.cuh"
#include "k_mix.cuh"
#include "k_chain.cuh"
#include "k_resample_prog.cuh"
#include "k_sinc.cuh"
