/*
 * skgpu_hub.h -- the frame-batching layer (libskgpu_hub.so), C ABI.
 *
 * BASELINE.json's north star puts a frame-batching layer into StreamKit's crates/engine: "gathers each tick's
 * 10-20 ms frames from all live sessions into pinned host buffers and sends them to the device through a thin
 * C-ABI FFI". In the reference nothing is batched across sessions: every node of every session is its own tokio
 * task with its own channels (crates/engine/src/dynamic_actor.rs:393-495, crates/engine/src/graph_builder.rs:378-422),
 * one DynamicEngine actor per session (apps/skit/src/session.rs:173-200). The reference is Rust and this image has
 * no Rust toolchain, so the layer is written in C++ above include/skgpu_batch.h; a Rust engine binds THIS header
 *  and keeps only the node shims (forward a frame, await the result).
 *
 * One hub = one GPU. A SESSION is one instance of the chain
 *     n x [ audio::resampler{target 48 kHz, chunk = 20 ms, output_frame_size F} -> audio::gain ]
 *       -> audio::mixer{clocked, F frames} -> audio::gain -> f32->s16
 * (samples/pipelines/dynamic/moq_mixing.yml:63-74 is this shape with n = 2 and 48 kHz decoder outputs: bypass inputs, no
 * resampler work at all). Per 20 ms tick the engine pushes at most
 * one chunk per input, calls skgpu_hub_tick (asynchronous: upload, one fused kernel pass, read-back) and collects
 * every session's mixed packet after skgpu_hub_wait.
 *
 * Semantics kept from the reference nodes:
 *   - an input that delivers nothing in a tick contributes silence to that tick's mix and keeps its resampler state
 *     (clocked mixer: mixer.rs:1354-1367; the hub marks the stream absent and repeats its previous chunk bytes as the
 *     fused kernel's protocol requires);
 *   - gains are validated like AudioGainConfig (finite, 0.0 ..= 4.0, gain.rs:50-66); an invalid update is rejected and
 *     the old gain stays (gain.rs:153-173); an accepted update applies from the next tick (gain.rs:151);
 *   - a new stream starts from a fresh FastFixedIn (zero history, last_index = -4, resampler.rs:232-238) and emits its
 *     first packet once F frames are available (resampler.rs:425-428): the first tick of a 44.1 kHz input mixes silence.
 *
 * Threading: skgpu_hub_push may be called concurrently from any number of threads for DISTINCT (session, input)
 * pairs (the engine's node tasks); every other call belongs to the hub's tick thread.
 */
#ifndef SKGPU_HUB_H
#define SKGPU_HUB_H

#include "skgpu_batch.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skgpu_hub skgpu_hub;

#define SKGPU_HUB_OUT_S16 1u /* sessions deliver s16 (clip + pack); without it f32 */
#define SKGPU_HUB_IN_S16 2u  /* inputs arrive as interleaved s16 (x = s / 32768, exact): half the PCIe bytes for natively 16-bit
                              * sources; push / acquire then take and hand out int16_t samples */

typedef struct skgpu_hub_config {
    uint32_t max_sessions;           /* capacity: concurrently open sessions */
    uint32_t max_streams;            /* capacity: resampled inputs over all sessions */
    uint32_t max_inputs_per_session; /* 1..64 */
    uint32_t out_rate;               /* mixer / resampler target rate, e.g. 48000 */
    uint32_t out_frames;             /* F: output_frame_size = frame_samples_per_channel, e.g. 960 */
    uint16_t channels;               /* 1 | 2, inputs and output */
    uint16_t flags;                  /* SKGPU_HUB_OUT_S16 */
    const uint32_t *in_rates;        /* every input sample rate sessions may use; a chunk is in_rate * F / out_rate frames. A rate
                                      * equal to out_rate is a BYPASS input: its F-frame packets enter the mix untouched, as the
                                      * reference's resampler node forwards them (resampler.rs:299-373) -- the 48 kHz frames an
                                      * Opus decoder emits (opus.rs:103,122-131) need no resampler in front of a 48 kHz mixer */
    uint32_t n_in_rates;
    uint32_t jitter_frames;          /* chunks an input may queue ahead (ClockedMixerConfig.jitter_buffer_frames, default 3 in
                                      * the reference, mixer.rs:46-55); 0 = 1. Costs jitter_frames + 2 pinned input arenas. */
    uint32_t slices;                 /* a tick runs as this many slices (upload / kernels / read-back overlapped, skgpu_batch.h
                                      * "sliced ticks"): the first sessions' packets are back in host memory while the last ones
                                      * still upload. 0 = 16; 1 = whole-tick submit */
} skgpu_hub_config;

/* message of the last error on the calling thread (borrowed, like skgpu_last_error) */
const char *skgpu_hub_last_error(void);

skgpu_rc skgpu_hub_create(int32_t device_ordinal, const skgpu_hub_config *cfg, skgpu_hub **out);
void skgpu_hub_destroy(skgpu_hub *hub);

/* in_rates[n_inputs]: sample rate of every input (each must be one of cfg.in_rates). Gains start at 1.0. */
skgpu_rc skgpu_hub_session_open(skgpu_hub *hub, uint32_t n_inputs, const uint32_t *in_rates, uint32_t *session_out);

/* ---- audio::mixer SYNC mode on the batch path (mixer.rs:554-918; the default above is the clocked mode, :1242-1434).
 * The state machine that decides WHICH frames enter a mix stays on the host and is evaluated once per hub tick:
 *   cold start   nothing is mixed until every active input has delivered at least once (:728-731)
 *   ready        every input that is not marked slow holds a frame -> mix now; slow inputs that delivered again are
 *                "recovered" and the session returns to Running (:746-762)
 *   timeout      a frame has been waiting for sync_timeout_ms (AudioMixerConfig.sync_timeout_ms, default 100, :639-709,
 *                :782-838) -> the inputs still missing are marked SLOW (stats.discarded += their number), the session becomes
 *                Degraded{"slow_input_timeout"} and the mix goes out with silence in their place; later mixes no longer wait
 *                for slow inputs
 *   hold         otherwise: nothing is consumed, nothing is emitted (n_mixed = 0); the waiting chunks stay queued
 *   EOF          skgpu_hub_input_eof removes an input from the active set (:848-898); buffered frames of the others are
 *                mixed at the next tick; a session whose inputs are all closed is Stopped{"all_inputs_closed"}
 * Time is the hub's tick clock: a timeout of T ms expires after ceil(T / tick_ms) ticks (tick_ms = 1000 F / out_rate).
 * A waiting input keeps its LATEST chunk (jitter_frames = 1: a second push replaces the first, like slot.frame = Some(frame));
 * unlike the reference, where the upstream resampler node has already consumed the replaced packet, the replaced chunk never
 * reaches the stream's resampler state. */
#define SKGPU_SESSION_SYNC 1u
skgpu_rc skgpu_hub_session_open_ex(skgpu_hub *hub, uint32_t n_inputs, const uint32_t *in_rates, uint32_t mode, uint32_t sync_timeout_ms,
                                   uint32_t *session_out);
skgpu_rc skgpu_hub_input_eof(skgpu_hub *hub, uint32_t session, uint32_t input);
#define SKGPU_SESSION_RUNNING 1u
#define SKGPU_SESSION_DEGRADED 2u   /* NodeState::Degraded{reason: "slow_input_timeout"} (crates/core/src/state.rs:122-186) */
#define SKGPU_SESSION_STOPPED 4u    /* all inputs closed */
typedef struct skgpu_session_state {
    uint32_t state;          /* SKGPU_SESSION_* */
    uint32_t mixed;          /* 1 = the last tick mixed (consumed the waiting frames); 0 = it held or had nothing */
    uint64_t slow_mask;      /* bit i: input i is marked slow */
    uint64_t eof_mask;       /* bit i: input i reached EOF */
    uint64_t newly_slow;     /* inputs marked slow by the last tick (the "newly_slow_pins" of the Degraded details) */
    uint64_t recovered;      /* inputs that left the slow state at the last tick */
} skgpu_session_state;
skgpu_rc skgpu_hub_session_state(skgpu_hub *hub, uint32_t session, skgpu_session_state *out);
skgpu_rc skgpu_hub_session_close(skgpu_hub *hub, uint32_t session);
skgpu_rc skgpu_hub_set_input_gain(skgpu_hub *hub, uint32_t session, uint32_t input, float gain);
skgpu_rc skgpu_hub_set_master_gain(skgpu_hub *hub, uint32_t session, float gain);
/* frames a chunk of this input must have (in_rate * F / out_rate) */
skgpu_rc skgpu_hub_chunk_frames(skgpu_hub *hub, uint32_t session, uint32_t input, uint32_t *frames_out);

/* one chunk (interleaved f32 -- int16_t with SKGPU_HUB_IN_S16 --, n_frames == chunk frames of the input), copied into the pinned arena of the first tick
 * that has no chunk of this input yet: an input may queue up to jitter_frames chunks, every tick consumes one; a push
 * into a full queue drops the oldest chunk (the clocked mixer's InputRingBuffer, mixer.rs:1185-1206).
 * Threading: push / push_batch / acquire / commit may be called from any thread, concurrently with each other (distinct
 * streams) AND with skgpu_hub_tick: the hub serialises them against the tick's cut, so a racing chunk lands either in this
 * tick or in the next one, never in an arena that is being uploaded. Everything else is the tick thread's. */
skgpu_rc skgpu_hub_push(skgpu_hub *hub, uint32_t session, uint32_t input, const void *samples, uint32_t n_frames);

/* zero-copy variant: *dst_out is the stream's slot in the pinned arena of the NEXT tick (chunk frames x channels f32);
 * the producer (e.g. a decoder) writes its samples there and calls skgpu_hub_commit BEFORE the next skgpu_hub_tick
 * (a commit after a tick intervened fails with SKGPU_ERR_STATE and the chunk is dropped). This is how the pinned arenas
 * replace AudioFramePool buffers (crates/core/src/frame_pool.rs:302-317) on the batched path: no gather copy at all.
 * The pointer is valid until the next skgpu_hub_tick. */
skgpu_rc skgpu_hub_acquire(skgpu_hub *hub, uint32_t session, uint32_t input, void **dst_out, uint32_t *n_frames_out);
skgpu_rc skgpu_hub_commit(skgpu_hub *hub, uint32_t session, uint32_t input);
/* every live stream delivered its chunk in place (producers that always write their slot) */
skgpu_rc skgpu_hub_commit_all(skgpu_hub *hub);

/* many chunks at once, copied by n_threads worker threads (the gather of a whole tick is ~1 GB at 65 k sessions:
 * one thread cannot keep up with PCIe). frames[i] must name distinct (session, input) pairs. */
typedef struct skgpu_hub_frame {
    const void *samples;   /* f32, or int16_t with SKGPU_HUB_IN_S16 */
    uint32_t session, input, n_frames, reserved;
} skgpu_hub_frame;
skgpu_rc skgpu_hub_push_batch(skgpu_hub *hub, const skgpu_hub_frame *frames, uint32_t n, uint32_t n_threads);

/* asynchronous: table updates after session churn, presence + gains, upload, kernels, read-back */
skgpu_rc skgpu_hub_tick(skgpu_hub *hub);
/* blocks until the last tick has finished; timing may be NULL */
skgpu_rc skgpu_hub_wait(skgpu_hub *hub, skgpu_tick_timing *timing);
/* pipelined collection: blocks until tick number `tick` (1-based, = skgpu_hub_ticks() right after its skgpu_hub_tick)
 * has finished without waiting for the next tick if that one is already submitted; skgpu_hub_session_output then
 * reports tick `tick`. Lets the engine submit tick n + 1 first and collect tick n while n + 1 uploads (the read-back of
 * n overlaps the upload of n + 1). Only the two most recent ticks can be waited for. */
skgpu_rc skgpu_hub_wait_tick(skgpu_hub *hub, uint64_t tick);
/* the session's packet of the last waited tick: F * channels samples (s16 or f32) inside the hub's pinned output arena,
 * valid until the next skgpu_hub_wait. *n_mixed = inputs that contributed a packet (0 = silence), *status = OR of the
 * inputs' chain status bits (skgpu_chain_result.status). Sessions opened after that tick was submitted report 0 / NULL. */
skgpu_rc skgpu_hub_session_output(skgpu_hub *hub, uint32_t session, const void **samples, uint32_t *n_mixed, uint32_t *status);

/* NodeStatsTracker counters (crates/core/src/stats.rs:131-152) of the batched nodes, summed over the hub:
 *   received   chunks consumed by ticks (a resampler node's stats.received(), resampler.rs:283)
 *   sent       mixed packets produced (one per live session per tick, mixer.rs:1011 stats.sent())
 *   discarded  chunks dropped by overwrite-oldest when an input's jitter queue was full (mixer.rs:1195-1201)
 *   errored    rejected gain updates (gain.rs:157-171 stats.errored()), rejected pushes, ticks that failed on the device */
typedef struct skgpu_hub_stats {
    uint64_t received, sent, discarded, errored;
} skgpu_hub_stats;
skgpu_rc skgpu_hub_get_stats(skgpu_hub *hub, skgpu_hub_stats *out);

/* node state of the hub as a whole (crates/core/src/state.rs:122-186): 1 = Running, 2 = Degraded (a sync-mode session
 * timed out on a slow input this tick), 3 = Failed (a CUDA error surfaced; every later call returns SKGPU_ERR_STATE and
 * the engine must emit NodeState::Failed{reason} for the sessions of this hub). */
#define SKGPU_HUB_RUNNING 1u
#define SKGPU_HUB_DEGRADED 2u
#define SKGPU_HUB_FAILED 3u
uint32_t skgpu_hub_state(const skgpu_hub *hub, const char **reason_out);

/* pins the CALLING thread to the CPUs of the hub GPU's NUMA node (the thread that ticks / gathers for this hub); returns the
 * node, or -1 when the platform exposes none */
int32_t skgpu_hub_bind_thread(skgpu_hub *hub);

/* counters for logs / tests */
uint32_t skgpu_hub_live_sessions(const skgpu_hub *hub);
uint32_t skgpu_hub_live_streams(const skgpu_hub *hub);
uint64_t skgpu_hub_ticks(const skgpu_hub *hub);

#ifdef __cplusplus
}
#endif
#endif /* SKGPU_HUB_H */
