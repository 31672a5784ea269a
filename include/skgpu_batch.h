/*
 * skgpu_batch.h -- batch C ABI of streamkit_b200 (libskgpu.so), version 1.
 *
 * This is the per-tick boundary a frame-batching layer in StreamKit's crates/engine binds over FFI
 * (SURVEY.md 8b "Batch C ABI"; INTEGRATION.md shows the Rust `extern "C"` block). It does not exist in
 * the reference: there every node is one tokio task fed one packet at a time
 * (crates/engine/src/dynamic_actor.rs:393-495, crates/core/src/node.rs:191-226). Each entry point below
 * names the reference code whose per-packet work it replaces.
 *
 * Conventions (mirroring the native plugin ABI, sdks/plugin-sdk/native/src/types.rs):
 *   - plain C, POD structs, no exceptions cross the boundary;
 *   - every call returns skgpu_rc (0 = ok, <0 = error); the message for the last error on the calling
 *     thread is borrowed from skgpu_last_error() and stays valid until the next error on that thread
 *     (same ownership rule as CResult.error_message, types.rs:42-48, conversions.rs:441-461);
 *   - a context is thread-compatible: one submitting thread at a time per context; different contexts
 *     (one per GPU) are independent. Sessions shard across GPUs by mix-group id, no collective.
 *   - all audio is interleaved (crates/core/src/types.rs:210-211). Offsets are BYTE offsets into the
 *     plan's device arena and must be 4-byte aligned (2 for s16); 16-byte alignment enables the
 *     128-bit / TMA paths.
 *   - there is NO CPU fallback: without a CUDA device skgpu_ctx_create fails.
 */
#ifndef SKGPU_BATCH_H
#define SKGPU_BATCH_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKGPU_ABI_VERSION 1u

typedef int32_t skgpu_rc;
#define SKGPU_OK 0
#define SKGPU_ERR_INVALID (-1)   /* bad argument / configuration (StreamKitError::Configuration) */
#define SKGPU_ERR_CUDA (-2)      /* CUDA runtime error (StreamKitError::Runtime, node -> Failed) */
#define SKGPU_ERR_NOMEM (-3)
#define SKGPU_ERR_STATE (-4)     /* call not valid in the current state */
#define SKGPU_ERR_NODEVICE (-5)

typedef struct skgpu_ctx skgpu_ctx;
typedef struct skgpu_plan skgpu_plan;

uint32_t skgpu_abi_version(void);
const char *skgpu_last_error(void);

/* ------------------------------------------------------------------ context */

typedef struct skgpu_ctx_config {
    uint32_t max_streams;     /* resampler stream slots (per-stream state in HBM) */
    uint32_t max_channels;    /* 1..8; history is stored 16 frames x max_channels per slot */
    uint32_t fifo_frames;     /* per-slot device re-framing ring (frames, power of two), 0 = none */
    uint32_t flags;           /* reserved, 0 */
} skgpu_ctx_config;

skgpu_rc skgpu_ctx_create(int32_t device_ordinal, const skgpu_ctx_config *cfg, skgpu_ctx **out);
void skgpu_ctx_destroy(skgpu_ctx *ctx);
/* "NVIDIA B200" etc.; sm count; for logs */
skgpu_rc skgpu_ctx_device_info(skgpu_ctx *ctx, char *name, size_t name_len, int32_t *sm_count, int32_t *cc_major,
                               int32_t *cc_minor);

/* Pinned (page-locked) host memory for the tick arenas the batching layer gathers frames into.
 * Replaces AudioFramePool buckets (crates/core/src/frame_pool.rs:302-317) on the batched path. */
skgpu_rc skgpu_pinned_alloc(skgpu_ctx *ctx, size_t bytes, void **out);
skgpu_rc skgpu_pinned_free(skgpu_ctx *ctx, void *p);
/* Placement-aware variant ("one pinned arena pair per GPU, NUMA-pin each to the GPU's socket"). The pages are
 * mapped on the host NUMA node the GPU hangs off (sysfs numa_node of its PCI function; mbind MPOL_BIND before the pages
 * are touched), transparent huge pages are requested, then the range is registered with CUDA (cudaHostRegister).
 * SKGPU_PIN_WRITE_COMBINED (input arenas only: the CPU never reads them back) uses cudaHostAllocWriteCombined instead.
 * skgpu_pinned_alloc == skgpu_pinned_alloc_ex(flags = SKGPU_PIN_NUMA_LOCAL). *node_out (may be NULL) receives the node the
 * pages were bound to, -1 when the platform exposes none (single-node VM) and the default policy was kept. */
#define SKGPU_PIN_NUMA_LOCAL 1u
#define SKGPU_PIN_WRITE_COMBINED 2u
skgpu_rc skgpu_pinned_alloc_ex(skgpu_ctx *ctx, size_t bytes, uint32_t flags, void **out, int32_t *node_out);
/* host NUMA node of the context's GPU (-1 unknown) and the CPUs of that node as a bitmask string for logs ("0-15") */
int32_t skgpu_ctx_numa_node(skgpu_ctx *ctx);
/* pins the CALLING thread to the CPUs of the GPU's NUMA node (the submit / gather thread of this GPU); no-op when unknown */
skgpu_rc skgpu_ctx_bind_thread(skgpu_ctx *ctx);

/* ------------------------------------------------------------------ resampler stream slots
 * One slot = one rubato::FastFixedIn<f32>(Linear) instance (resampler.rs:232-238): last_index (f64) and
 * a 16-frame history live in HBM for the lifetime of the stream. */

/* stream flags */
#define SKGPU_STREAM_S16 1u  /* chain op only: the stream's chunks arrive as interleaved s16 (x = s / 32768, exact) -- half the
                              * PCIe bytes of f32 for natively 16-bit sources; expanded to f32 in shared memory */
#define SKGPU_STREAM_SINC 2u /* resample op only: windowed-sinc polyphase interpolation (skgpu_ctx_set_sinc) instead of rubato's Linear */
typedef struct skgpu_stream_cfg {
    uint32_t in_rate;           /* Hz, from the first packet (resampler.rs:206-209) */
    uint32_t out_rate;          /* AudioResamplerConfig.target_sample_rate (resampler.rs:22-27) */
    uint32_t chunk_frames;      /* AudioResamplerConfig.chunk_frames, input frames per process() */
    uint16_t channels;          /* 1..max_channels */
    uint16_t flags;             /* SKGPU_STREAM_* */
} skgpu_stream_cfg;
/* in_rate == out_rate opens a BYPASS stream: the reference's resampler node does no DSP for such inputs and forwards /
 * re-frames the packets untouched (resampler.rs:299-373). In a chain op a bypass stream delivers exactly one packet of
 * output_frame_size frames per tick (chunk_frames == output_frame_size) that enters the mix as it is -- the path of the
 * 48 kHz mono frames an Opus decoder emits (crates/nodes/src/audio/codecs/opus.rs:103,122-131; moq_mixing.yml has no
 * resampler at all). No arithmetic is applied to it (not a frac = 0 interpolation). Resample ops reject bypass streams. */

/* Windowed-sinc polyphase mode (BASELINE.json north star: "a batched polyphase windowed-sinc resampler that keeps filter state per
 * stream in HBM"). The reference resamples with rubato FastFixedIn / Linear (resampler.rs:232-238), so this mode has no reference
 * implementation; its specification, in rubato's SincInterpolationParameters vocabulary:
 *   sinc_len L (taps per phase, multiple of 8, <= 256), oversampling_factor O (<= 1024), f_cutoff (of the lower Nyquist),
 *   window BlackmanHarris2 (4-term Blackman-Harris squared), interpolation Linear between the two nearest phases;
 *   fc = f_cutoff * min(1, out_rate / in_rate);  g(tau) = fc sinc(fc tau) w((tau + L/2) / L);
 *   T[p][n] = g(L/2 - 1 - n + p/O) / sum_n g(..) (f64 -> f32), p = 0..O;
 *   per chunk of N frames: idx starts at last_index (-(L/2) for a new stream) and advances by 1/ratio (f64) while
 *   idx < N - L/2 - 1 - ceil(1/ratio); output = (1 - q) dot(window, T[p]) + q dot(window, T[p + 1]) with p = floor(frac O),
 *   q = f32(frac O - p), window = L input frames from floor(idx) - L/2 + 1 (zeros before the stream starts), every dot product a
 *   sequential f32 fma chain in ascending tap order. State per stream in HBM: last_index (f64) + the last L + 8 input frames.
 * The oracle restates this independently (oracle/sk_sinc.c, oracle/np_oracle.py). Call before opening SKGPU_STREAM_SINC streams;
 * the parameters are fixed for the context's lifetime.
 * Performance note: a resample op whose sinc streams share ONE tap table (all up-sampling streams do: fc = f_cutoff; down-sampling
 * streams of one ratio do) runs the persistent kernel that keeps the table in shared memory (k_resample_sinc_tiled); an op that
 * mixes tables, or whose table does not fit beside two staged chunks, runs one CTA per stream with the taps read through L1. Same
 * results either way. */
skgpu_rc skgpu_ctx_set_sinc(skgpu_ctx *ctx, uint32_t sinc_len, uint32_t oversampling_factor, double f_cutoff);

skgpu_rc skgpu_stream_open(skgpu_ctx *ctx, const skgpu_stream_cfg *cfg, uint32_t *slot_out);
/* bulk open of n identical streams (session ramp-up); slots_out[n] */
skgpu_rc skgpu_stream_open_many(skgpu_ctx *ctx, const skgpu_stream_cfg *cfg, uint32_t n, uint32_t *slots_out);
/* back to a fresh FastFixedIn: zero history, last_index = -4.0 */
skgpu_rc skgpu_stream_reset(skgpu_ctx *ctx, uint32_t slot);
skgpu_rc skgpu_stream_close(skgpu_ctx *ctx, uint32_t slot);
/* read back state (tests, checkpointing): hist receives 16*channels floats (interleaved) */
skgpu_rc skgpu_stream_get_state(skgpu_ctx *ctx, uint32_t slot, double *last_index, float *hist, uint64_t *fifo_written,
                                uint64_t *fifo_read);
/* upper bound of frames one chunk can produce for this configuration */
uint32_t skgpu_stream_max_out_frames(const skgpu_stream_cfg *cfg);

/* ------------------------------------------------------------------ descriptors */

#define SKGPU_NO_GAIN 0xFFFFFFFFu /* gain_idx sentinel: no multiply */

/* format-conversion / gain segments (one per frame). Replaces gain.rs:184-190 and the build-defined
 * f32<->s16 conversion (types.rs:26-29 only declares SampleFormat::S16Le). */
typedef enum skgpu_cvt_mode {
    SKGPU_CVT_F32_TO_F32 = 0, /* y = x * g                                  (audio::gain)               */
    SKGPU_CVT_F32_TO_S16 = 1, /* s = sat_s16(rint_even((x * g) * 32768))    (gain -> clip -> s16 pack)  */
    SKGPU_CVT_S16_TO_F32 = 2  /* y = (s / 32768) * g                         (s16 ingest)                */
} skgpu_cvt_mode;

typedef struct skgpu_seg {
    uint64_t in_off;
    uint64_t out_off;    /* may equal in_off for F32_TO_F32 (in place, like make_samples_mut) */
    uint32_t n_samples;  /* total samples (all channels) */
    uint32_t gain_idx;   /* index into the plan's gain table, or SKGPU_NO_GAIN */
} skgpu_seg;

/* resampler work item: one chunk of one stream. Replaces resampler.rs:377-417 (+ rubato process). */
#define SKGPU_RS_TO_FIFO 1u /* append output to the slot's device ring instead of out_off */
typedef struct skgpu_rs_item {
    uint64_t in_off;          /* chunk_frames * channels f32, interleaved */
    uint64_t out_off;         /* receives out_frames * channels f32 (ignored with SKGPU_RS_TO_FIFO) */
    uint32_t slot;
    uint32_t out_cap_frames;  /* capacity at out_off, frames */
    uint32_t flags;
    uint32_t reserved;
} skgpu_rs_item;

typedef struct skgpu_rs_result { /* written per item at the op's results offset */
    uint32_t out_frames;
    uint32_t status;          /* 0 ok; 1 = output truncated (out_cap_frames too small); 2 = run table overflow */
} skgpu_rs_result;

/* mixer. Replaces mixer.rs:922-1019 / :1436-1492 (+ :1027-1078). Inputs are listed in PIN ORDER
 * (in_0, in_1, ...); base-frame selection and swap_remove ordering (mixer.rs:960-980) run on the device
 * every tick from the inputs that are present. */
#define SKGPU_MIX_IN_UNIQUE 1u   /* AudioFrame::has_unique_samples() */
#define SKGPU_MIX_IN_FIFO 2u     /* read one output_frame_size packet from slot's device ring */
typedef struct skgpu_mix_input {
    uint64_t in_off;
    uint32_t n_frames;    /* frames per channel in this input frame */
    uint16_t channels;
    uint16_t flags;
    uint32_t gain_idx;    /* per-input audio::gain applied before the sum (rounded separately), or NO_GAIN */
    uint32_t slot;        /* for SKGPU_MIX_IN_FIFO */
} skgpu_mix_input;

#define SKGPU_MIX_OUT_S16 1u     /* epilogue: clip + pack s16 instead of storing f32 */
typedef struct skgpu_mix_group {
    uint64_t out_off;
    uint32_t first_input;  /* index of the group's first skgpu_mix_input */
    uint32_t n_inputs;
    uint32_t out_frames;   /* frames per channel of the output (sync: longest input; clocked: fixed) */
    uint16_t out_channels; /* sticky max channel count (mixer.rs:947-948), decided by the host */
    uint16_t flags;
    uint32_t gain_idx;     /* master audio::gain after the mix, or NO_GAIN */
    uint32_t reserved;
} skgpu_mix_group;

/* fused chain (BASELINE config #5): per session K x [audio::resampler{output_frame_size F} -> audio::gain] ->
 * audio::mixer -> audio::gain -> (s16). One kernel, no f32 intermediate in HBM. Replaces, per tick and session,
 * resampler.rs:377-470, gain.rs:187-189 (K + 1 times), mixer.rs:960-980 + :1027-1078.
 *
 * Protocol (what the frame-batching layer must do):
 *  - the H2D range is DOUBLE-BANKED (skgpu_plan_set_banks): tick n uploads into bank n & 1; a stream keeps the
 *    same in_off while it is alive, so its previous chunk is found at the same offset in the other bank;
 *  - a stream that delivers no chunk in a tick is marked absent (skgpu_plan_set_present) and the host repeats
 *    the stream's previous chunk bytes at in_off, so the other bank stays valid for the next tick;
 *  - eligible streams: channels 1 or 2, chunk_frames >= 16, chunk_frames * out_rate / in_rate == F (a packet
 *    never spans more than two chunks). Everything else uses the unfused resample + mix ops. */
typedef struct skgpu_chain_input {
    uint64_t in_off;      /* byte offset of the stream's chunk inside bank 0 (bank 1 = + bank_stride), 16-byte aligned; the chunk
                           * is chunk_frames x channels f32 (s16 with SKGPU_STREAM_S16: then chunk_frames x channels must be even
                           * and at most 4096, and 16 bytes past the chunk's end must still lie inside the arena) */
    uint32_t slot;        /* resampler stream slot (state in HBM) */
    uint32_t gain_idx;    /* per-input audio::gain, or SKGPU_NO_GAIN */
    uint32_t flags;       /* SKGPU_MIX_IN_UNIQUE */
    uint32_t reserved;
} skgpu_chain_input;

typedef struct skgpu_chain_group {
    uint64_t out_off;      /* F * out_channels samples (s16 with SKGPU_MIX_OUT_S16, else f32), 16-byte aligned */
    uint32_t first_input;
    uint32_t n_inputs;     /* <= 64 */
    uint32_t gain_idx;     /* master audio::gain, or SKGPU_NO_GAIN */
    uint16_t out_channels; /* 1 or 2 (sticky max, decided by the host) */
    uint16_t flags;        /* SKGPU_MIX_OUT_S16 */
} skgpu_chain_group;

typedef struct skgpu_chain_result { /* per input, written at the op's results offset every tick */
    uint32_t emitted;      /* 1 = this input contributed an F-frame packet to the mix this tick */
    uint32_t status;       /* bit0 backlog (second packet pending), bit1 frame-program overflow (record incomplete),
                            * bit2 unsupported packet (carry spans two chunks, or its tail needs more than the 32 staged
                            * frames of the current chunk); with bit1 or bit2 set nothing is emitted */
} skgpu_chain_result;

/* ------------------------------------------------------------------ plan = one compiled tick
 * A plan is the steady-state shape of a tick: an ordered list of ops whose descriptor tables live on the
 * device, one H2D range and one D2H range. Built once, replayed every 20 ms; tables can be replaced
 * when sessions come and go (same capacity), gains / presence change per tick. */

skgpu_rc skgpu_plan_create(skgpu_ctx *ctx, size_t arena_bytes, skgpu_plan **out);
void skgpu_plan_destroy(skgpu_plan *plan);

/* each add_* returns the op index in *op_out (may be NULL) */
skgpu_rc skgpu_plan_add_convert(skgpu_plan *plan, skgpu_cvt_mode mode, const skgpu_seg *segs, uint32_t n,
                                uint32_t *op_out);
skgpu_rc skgpu_plan_add_resample(skgpu_plan *plan, const skgpu_rs_item *items, uint32_t n, uint64_t results_off,
                                 uint32_t *op_out);
skgpu_rc skgpu_plan_add_mix(skgpu_plan *plan, const skgpu_mix_group *groups, uint32_t n_groups,
                            const skgpu_mix_input *inputs, uint32_t n_inputs, uint32_t *op_out);
/* requires skgpu_plan_set_banks first; output_frame_size = AudioResamplerConfig.output_frame_size (resampler.rs:33-38) */
skgpu_rc skgpu_plan_add_chain(skgpu_plan *plan, const skgpu_chain_group *groups, uint32_t n_groups,
                              const skgpu_chain_input *inputs, uint32_t n_inputs, uint32_t output_frame_size,
                              uint64_t results_off, uint32_t *op_out);

/* same, with explicit table capacities (sessions come and go: skgpu_plan_update_chain may grow the tables up to these)
 * and the largest n_inputs any later group will have (0 = the initial tables' maximum). The initial inputs must cover
 * every stream configuration (rate pair, chunk size, channels) later updates will use: staging is sized from them.
 * They also fix which kernel the op runs: when every initial input is of ONE kind -- f32 resampled, f32 rate-equal, s16 resampled,
 * s16 rate-equal, each with the group's channel count -- or an f32 mixture of resampled and rate-equal inputs of that shape, the op
 * gets the instantiation specialised for it (faster), and a later update must stay within that kind
 * (SKGPU_ERR_INVALID otherwise). Initial tables that already mix channel counts or formats get the general kernel. */
skgpu_rc skgpu_plan_add_chain_cap(skgpu_plan *plan, const skgpu_chain_group *groups, uint32_t n_groups,
                                  const skgpu_chain_input *inputs, uint32_t n_inputs, uint32_t cap_groups, uint32_t cap_inputs,
                                  uint32_t max_inputs_per_group, uint32_t output_frame_size, uint64_t results_off, uint32_t *op_out);

/* replace an op's descriptor table in place (n <= capacity given at add time) */
skgpu_rc skgpu_plan_update_convert(skgpu_plan *plan, uint32_t op, const skgpu_seg *segs, uint32_t n);
skgpu_rc skgpu_plan_update_resample(skgpu_plan *plan, uint32_t op, const skgpu_rs_item *items, uint32_t n);
skgpu_rc skgpu_plan_update_mix(skgpu_plan *plan, uint32_t op, const skgpu_mix_group *groups, uint32_t n_groups,
                               const skgpu_mix_input *inputs, uint32_t n_inputs);
skgpu_rc skgpu_plan_update_chain(skgpu_plan *plan, uint32_t op, const skgpu_chain_group *groups, uint32_t n_groups,
                                 const skgpu_chain_input *inputs, uint32_t n_inputs);

/* tick I/O ranges: host_in[0..bytes) -> arena[h2d_off..), arena[d2h_off..) -> host_out[0..bytes) */
skgpu_rc skgpu_plan_set_io(skgpu_plan *plan, uint64_t h2d_off, uint64_t h2d_bytes, uint64_t d2h_off, uint64_t d2h_bytes);
/* double-bank the H2D range: tick n uploads host_in to h2d_off + (n & 1) * bank_stride (bank_stride >= h2d_bytes,
 * 16-byte multiple). Offsets of chain inputs are relative to bank 0. Call before skgpu_plan_add_chain. */
skgpu_rc skgpu_plan_set_banks(skgpu_plan *plan, uint64_t bank_stride);
/* number of ticks submitted so far (bank of the NEXT tick = result & 1) */
uint64_t skgpu_plan_tick_count(const skgpu_plan *plan);

/* ---- sliced ticks (chain op). A tick of tens of thousands of sessions is ~1 GB up and ~250 MB down: submitted as one
 * upload -> kernels -> read-back sequence, the first result leaves the device only after the LAST input byte arrived
 * (17 ms + 0.4 ms + 4.4 ms at 65,536 sessions). With slices the tables are cut into n consecutive pieces; slice i's kernels
 * start as soon as ITS inputs are uploaded and its results are read back while slice i + 1 still uploads (three streams,
 * PCIe is full duplex). Per slice the added device latency -- upload-done -> results-in-host-memory -- is
 * (kernels + read-back) / n. Requirements: the plan's only op is the chain op; slices are consecutive, cover the tables,
 * and a slice's inputs lie below its h2d_end inside the H2D range (groups sorted by input offset do that). */
typedef struct skgpu_slice {
    uint32_t group_end;   /* this slice runs groups [previous group_end, group_end) */
    uint32_t input_end;   /* ... which own inputs [previous input_end, input_end) */
    uint64_t h2d_end;     /* its kernels wait for host_in[0 .. h2d_end) (byte offset inside the H2D range, non-decreasing) */
    uint64_t d2h_off[2];  /* up to two byte ranges of the D2H range that are final after this slice (results rows, output rows) */
    uint64_t d2h_bytes[2];
} skgpu_slice;
skgpu_rc skgpu_plan_set_slices(skgpu_plan *plan, uint32_t chain_op, const skgpu_slice *slices, uint32_t n);
/* even split of the CURRENT tables into n slices (tables in input-offset order, outputs in group order, results rows first
 * then outputs inside the D2H range -- the layout streamkit_b200's own hosts use); re-run after skgpu_plan_update_chain */
skgpu_rc skgpu_plan_auto_slices(skgpu_plan *plan, uint32_t chain_op, uint32_t n);

typedef struct skgpu_slice_timing { /* CUDA events of one slice of a finished sliced tick */
    float upload_done_ms;   /* since the tick's first upload started */
    float kernels_ms;       /* upload-done -> kernels-done (includes waiting for the previous slice's kernels) */
    float latency_ms;       /* upload-done -> read-back-done: the slice's added device latency */
} skgpu_slice_timing;
/* timings of tick number `tick` (one of the two most recent, finished); returns the number of slices in *n_out */
skgpu_rc skgpu_tick_slice_timing(skgpu_plan *plan, uint64_t tick, skgpu_slice_timing *out, uint32_t cap, uint32_t *n_out);

/* per-tick dynamic parameters, snapshotted at submit (gain.rs:151: control messages are drained before
 * each packet, so a new gain applies from the next frame on). present[i] != 0 <=> input i of mix or chain op
 * `mix_op` delivered a frame this tick; absent inputs are silence (mixer.rs:999-1009, :1354-1367). */
skgpu_rc skgpu_plan_set_gains(skgpu_plan *plan, const float *gains, uint32_t n);
skgpu_rc skgpu_plan_set_present(skgpu_plan *plan, uint32_t mix_op, const uint8_t *present, uint32_t n);

/* validates, uploads tables, optionally captures the CUDA graph */
skgpu_rc skgpu_plan_finalize(skgpu_plan *plan);

#define SKGPU_SUBMIT_NO_H2D 1u   /* inputs already resident in the device arena */
#define SKGPU_SUBMIT_NO_D2H 2u
#define SKGPU_SUBMIT_GRAPH 4u    /* replay the captured CUDA graph instead of individual launches */
#define SKGPU_SUBMIT_TIME_OPS 8u /* record a CUDA event pair around every op (stream mode only) */
#define SKGPU_SUBMIT_SLICED 32u   /* run the tick slice by slice (skgpu_plan_set_slices): upload, kernels and read-back of
                                   * consecutive slices overlap on three streams; implies the overlapped read-back */
#define SKGPU_SUBMIT_OVERLAP_D2H 16u /* read results back on a second stream so the copy overlaps the NEXT tick's upload
                                      * (host_out must stay untouched until the next skgpu_tick_wait) */

/* asynchronous: enqueues H2D, kernels, D2H on the context stream and returns */
skgpu_rc skgpu_tick_submit(skgpu_plan *plan, const void *host_in, void *host_out, uint32_t flags);

typedef struct skgpu_tick_timing { /* of the most recent submitted tick, CUDA events on the ctx stream */
    float h2d_ms;
    float kernels_ms;
    float d2h_ms;
    float total_ms;
} skgpu_tick_timing;
/* blocks until everything submitted so far has finished; timing may be NULL */
skgpu_rc skgpu_tick_wait(skgpu_plan *plan, skgpu_tick_timing *timing);

/* blocks until tick number `tick` (1-based: the value of skgpu_plan_tick_count right after its submit) has finished,
 * read-back included, WITHOUT waiting for a later tick that is already submitted: with SKGPU_SUBMIT_OVERLAP_D2H and two
 * host_out buffers the caller collects tick n while tick n + 1 uploads. Only the two most recent ticks can be waited for. */
skgpu_rc skgpu_tick_wait_for(skgpu_plan *plan, uint64_t tick);

/* averaged device time of one op over the ticks submitted with SKGPU_SUBMIT_TIME_OPS since the last reset.
 * For a resample op, sub = 0 is the phase-table kernel, sub = 1 the interpolation kernel. */
skgpu_rc skgpu_plan_op_time(skgpu_plan *plan, uint32_t op, uint32_t sub, float *avg_ms, uint32_t *n_samples);
skgpu_rc skgpu_plan_reset_op_times(skgpu_plan *plan);
/* number of kernel launches one tick of this plan performs (for bench.py's gpu_launches) */
uint32_t skgpu_plan_launches_per_tick(const skgpu_plan *plan);

/* direct arena access (tests, device-resident benchmarking) */
skgpu_rc skgpu_arena_upload(skgpu_plan *plan, uint64_t off, const void *host, size_t bytes);
skgpu_rc skgpu_arena_download(skgpu_plan *plan, uint64_t off, void *host, size_t bytes);
skgpu_rc skgpu_arena_fill(skgpu_plan *plan, uint64_t off, int byte_value, size_t bytes);

/* CUDA-event stopwatch on the context stream (device time of a whole timed region) */
skgpu_rc skgpu_timer_start(skgpu_ctx *ctx);
skgpu_rc skgpu_timer_stop(skgpu_ctx *ctx);
skgpu_rc skgpu_timer_elapsed_ms(skgpu_ctx *ctx, float *ms); /* synchronises on the stop event */
skgpu_rc skgpu_ctx_sync(skgpu_ctx *ctx);
/* writes a buffer larger than L2 (flush between timed iterations of L2-sized workloads) */
skgpu_rc skgpu_ctx_flush_l2(skgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
