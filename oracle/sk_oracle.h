/*
 * sk_oracle.h -- CPU restatement of StreamKit's PCM DSP hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle for streamkit_b200. It is NOT part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it. The product (libskgpu.so) never links or calls it.
 *
 * Every function cites the reference file:line (relative to /root/reference) it restates.
 *
 * Pin status (see DESIGN.md "Oracle"):
 *   gain      : pinned against the reference's own C plugin compiled from
 *               examples/plugins/gain-native-c/gain_plugin.c (oracle/_ref) and the constants of
 *               the reference unit tests (gain.rs:283-286,323-331,402-404,433-435,465-509).
 *   mixer     : pinned at 1e-3 by the reference's unit-test scenarios (mixer.rs:1698-2102);
 *               summation order is DEFINED here (see sko_mix_plan) because the reference's
 *               HashMap iteration order is not deterministic.
 *   resampler : PARITY UNPINNED for sample values. rubato 0.16.2 (crates.io, Cargo.lock:3708-3711)
 *               is not vendored under /root/reference; FastFixedIn<f32>/PolynomialDegree::Linear is
 *               restated from its published algorithm (src/asynchro_fast.rs). Output lengths are
 *               pinned by resampler.rs:826-837 (474 samples inside 480+-10).
 *   s16       : PARITY UNPINNED. The reference has no f32<->s16 code (types.rs:26-29 only declares
 *               the tag). Defined here as sat_s16(rint_half_even(x*32768)), NaN->0, and s/32768.
 */
#ifndef SK_ORACLE_H
#define SK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- gain (gain.rs) */

/* gain.rs:50-66 AudioGainConfig::validate. Returns 0 if ok; 1 = not finite, 2 = out of [0,4].
 * On error writes the reference's message text into err (if non-NULL). */
int sko_gain_validate(float gain, char *err, size_t err_len);

/* gain.rs:187-189: for sample in frame.make_samples_mut() { *sample *= gain }  (in place) */
void sko_gain_apply(float *samples, size_t n, float gain);

/* ---------------------------------------------------------------- s16 (build-defined) */

int16_t sko_f32_to_s16(float x);
void sko_f32_to_s16_buf(const float *in, int16_t *out, size_t n);
void sko_s16_to_f32_buf(const int16_t *in, float *out, size_t n);
/* fused gain -> clip -> s16 (two roundings: x*gain then *32768 exact, then rint) */
void sko_gain_f32_to_s16_buf(const float *in, int16_t *out, size_t n, float gain);

/* ---------------------------------------------------------------- mixer (mixer.rs) */

typedef struct sko_frame {
    const float *samples; /* interleaved */
    uint32_t n_samples;   /* total across channels (AudioFrame.samples.len()) */
    uint16_t channels;
    uint8_t unique;       /* AudioFrame::has_unique_samples() (types.rs) -- Arc not shared */
    uint8_t _pad;
} sko_frame;

/* mixer.rs:1027-1078 mix_frame_with_channel_conversion: output[..] += source with channel mapping. */
void sko_mix_frame_with_channel_conversion(float *output, size_t output_len, const sko_frame *source,
                                           uint16_t output_channels);

/* Base-frame selection + swap_remove order, mixer.rs:960-980 (sync) / :1447-1467 (clocked).
 * frames[] is in pin order (the order defined by this build).
 * order[] receives n indices: if *has_base, order[0] is the base frame (its samples initialise the
 * accumulator, it is NOT added to zeros), order[1..] the frames added to it in Vec order after
 * swap_remove; if !*has_base the accumulator starts at +0.0 and order[0..n] are added in Vec order. */
void sko_mix_plan(const sko_frame *frames, size_t n, uint16_t out_channels, size_t out_size, uint32_t *order,
                  int *has_base);

/* mixer.rs:944-1013 mix_and_send arithmetic (sync mode).
 * out_channels = max(max_channels_seen, max over frames, 1); out_len = max frames/ch * out_channels.
 * Returns 0, or -1 if out_cap is too small. n == 0 produces *out_len = 0 (no output packet). */
int sko_mix_sync(const sko_frame *frames, size_t n, uint16_t max_channels_seen, float *out, size_t out_cap,
                 uint16_t *out_channels, size_t *out_len);

/* mixer.rs:1436-1492 mix_clocked_frames: fixed output shape frame_samples_per_channel*out_channels. */
int sko_mix_clocked(const sko_frame *frames, size_t n, uint16_t out_channels, size_t frame_samples_per_channel,
                    float *out, size_t out_cap);

/* ---------------------------------------------------------------- resampler core (rubato 0.16.2) */

typedef struct sko_ffi sko_ffi; /* rubato::FastFixedIn<f32> with PolynomialDegree::Linear */

/* resampler.rs:232-238: FastFixedIn::<f32>::new(ratio, 1.0, Linear, chunk_frames, channels) */
sko_ffi *sko_ffi_new(double ratio, size_t chunk_frames, size_t channels);
void sko_ffi_free(sko_ffi *r);
/* upper bound on frames one process() call may emit (rubato output_frames_max-like, generous) */
size_t sko_ffi_out_max(const sko_ffi *r);
/* resampler.rs:397-417 deinterleave -> process(&planar, None) -> interleave, fused:
 * in = chunk_frames*channels interleaved samples; out receives n*channels; returns n (frames). */
size_t sko_ffi_process_interleaved(sko_ffi *r, const float *in, float *out, size_t out_cap_frames);
double sko_ffi_last_index(const sko_ffi *r);
/* copies the 16-frame history (interleaved, 16*channels floats) -- for state comparison with the GPU */
void sko_ffi_history(const sko_ffi *r, float *hist);

/* ---------------------------------------------------------------- resampler node (resampler.rs:197-740) */

typedef struct sko_packet_meta {
    uint64_t timestamp_us;
    uint8_t has_timestamp;
    uint64_t duration_us;
    uint64_t sequence;
} sko_packet_meta;

/* emit callback: one output packet ("out" pin) */
typedef void (*sko_emit_fn)(void *ud, uint32_t sample_rate, uint16_t channels, const float *samples,
                            size_t n_samples, const sko_packet_meta *meta);

typedef struct sko_rsnode sko_rsnode;

/* resampler.rs:81-102 factory validation. Returns NULL and writes err on invalid config. */
sko_rsnode *sko_rsnode_new(uint32_t target_sample_rate, size_t chunk_frames, size_t output_frame_size, char *err,
                           size_t err_len);
void sko_rsnode_free(sko_rsnode *n);
/* one input audio packet (resampler.rs:203-528). Returns 0, or -1 on "Audio format changed mid-stream". */
int sko_rsnode_push(sko_rsnode *n, uint32_t sample_rate, uint16_t channels, const float *samples, size_t n_samples,
                    int has_timestamp, uint64_t timestamp_us, sko_emit_fn emit, void *ud, char *err, size_t err_len);
/* input closed (resampler.rs:543-730): remainder via fresh FastFixedIn + flush of partial frame */
int sko_rsnode_finish(sko_rsnode *n, sko_emit_fn emit, void *ud);

/* resampler.rs:108-116 */
uint64_t sko_duration_us_for_frames(uint32_t sample_rate, size_t frames_per_channel);

#ifdef __cplusplus
}
#endif
#endif
