// hub.cpp -- the frame-batching layer (include/skgpu_hub.h) above the batch C ABI (include/skgpu_batch.h).
//
// C++ stand-in for the layer BASELINE.json's north star adds to StreamKit's crates/engine (the reference is Rust and this
// image has no Rust toolchain): it owns the pinned tick arenas, assigns every stream a fixed offset in the double-banked
// input range, keeps the fused chain's descriptor tables in step with session churn (skgpu_plan_update_chain), gathers
// the frames node tasks push, and runs one asynchronous tick per 20 ms. Reference behaviour it preserves is cited in
// the header.
#include "../../../include/skgpu_hub.h"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;

skgpu_rc hub_fail(skgpu_rc rc, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return rc;
}
// propagate an error of the batch layer (its message lives in libskgpu's thread-local slot)
skgpu_rc hub_pass(skgpu_rc rc) {
    const char *m = skgpu_last_error();
    g_err = m ? m : "batch layer error";
    return rc;
}
#define PASS(call)                          \
    do {                                    \
        skgpu_rc rc__ = (call);             \
        if (rc__ != SKGPU_OK) return hub_pass(rc__); \
    } while (0)
// same, inside tick / wait: a CUDA error is fatal for every session of the hub -- the node state becomes Failed{reason}
// (crates/core/src/state.rs:122-186; wrapper.rs:468-483 does the same for a plugin whose process_packet fails)
#define PASS_FATAL(h, call)                 \
    do {                                    \
        skgpu_rc rc__ = (call);             \
        if (rc__ != SKGPU_OK) {             \
            hub_pass(rc__);                 \
            if (rc__ == SKGPU_ERR_CUDA) { (h)->fail_reason = g_err; (h)->stats.errored += 1; } \
            return rc__;                    \
        }                                   \
    } while (0)

uint64_t align_up(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

// Copy of one chunk into the pinned arena with NON-TEMPORAL stores: the destination is read next by the PCIe DMA engine,
// not by a CPU, so allocating it in the caches (and reading the lines first) only costs host memory bandwidth -- the
// resource the gather of ~1 GB per tick competes for with the upload itself. dst is 16-byte aligned (arena offsets are).
void stream_copy(uint8_t *dst, const uint8_t *src, size_t bytes) {
#if defined(__SSE2__)
    size_t i = 0;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (; i + 64 <= bytes; i += 64) {
            const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i));
            const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 16));
            const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 32));
            const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + i + 48));
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i), a);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 16), b);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 32), c);
            _mm_stream_si128(reinterpret_cast<__m128i *>(dst + i + 48), d);
        }
    }
    if (i < bytes) memcpy(dst + i, src + i, bytes - i);
#else
    memcpy(dst, src, bytes);
#endif
}
void stream_fence() {
#if defined(__SSE2__)
    _mm_sfence();
#endif
}

struct Stream {
    uint32_t slot = 0;         // resampler slot of the batch context
    uint32_t chunk = 0;        // frames per pushed chunk
    uint32_t in_rate = 0;
    bool live = false;
    bool ever_pushed = false;
    uint64_t acq_tick = ~0ull;   // value of `ticks` when the zero-copy slot was handed out (skgpu_hub_acquire)
};

struct Session {
    bool live = false;
    std::vector<uint32_t> streams;   // hub stream ids, pin order
    int64_t tab_first = -1;          // index of its first input in the submitted tables (-1: not in the tables yet)
    // ---- audio::mixer sync mode (mixer.rs:554-918), evaluated once per tick
    bool sync = false;
    uint32_t timeout_ticks = 0;      // 0 = no timeout (sync_timeout_ms: None)
    uint64_t has_sent = 0, slow = 0, eof = 0, have_frame = 0;   // per-input bit masks (<= 64 inputs)
    int64_t waiting_since = -1;      // tick number at which the first frame of the pending mix arrived
    bool force_mix = false;          // an input reached EOF while frames were buffered
    bool degraded_reported = false;
    uint64_t consumed = 0;           // inputs whose frame the current tick's mix takes
    skgpu_session_state st{SKGPU_SESSION_RUNNING, 0, 0, 0, 0, 0};
};

}  // namespace

struct skgpu_hub {
    skgpu_hub_config cfg{};
    std::vector<uint32_t> rates;
    skgpu_ctx *ctx = nullptr;
    skgpu_plan *plan = nullptr;
    uint32_t op = 0;
    uint32_t C = 2, F = 960, ob = 2, ib = 4;   // ib: bytes per input sample (2 with SKGPU_HUB_IN_S16)
    uint32_t n_slices = 16;
    bool slices_dirty = true;
    uint64_t in_stride = 0, in_bytes = 0, bank_stride = 0, res_off = 0, res_bytes = 0, out_off = 0, out_stride = 0, d2h_bytes = 0;
    static constexpr uint32_t MAX_RING = 10;    // jitter_frames (<= 8) + 2
    uint8_t *host_in[MAX_RING] = {};            // ring of pinned input arenas: tick k uploads host_in[k % R]
    uint32_t J = 1, R = 3;                      // queue depth per input (mixer.rs:1185-1206 InputRingBuffer) and ring size J + 2
    uint8_t *host_out[2] = {nullptr, nullptr};
    std::vector<Stream> streams;
    std::vector<Session> sessions;
    std::vector<uint32_t> free_streams, free_sessions;
    std::vector<std::atomic<uint8_t>> pushed;   // per stream: chunks queued for the next ticks (0..J); chunk i of the queue
                                                // sits in the input arena of tick (ticks + 1 + i)
    std::vector<float> gains;                   // [stream gains | master gains]
    bool gains_dirty = true, tables_dirty = true;
    std::vector<skgpu_chain_group> groups;
    std::vector<skgpu_chain_input> inputs;
    std::vector<uint32_t> tab_stream;           // table input index -> hub stream id
    std::vector<uint32_t> tab_session, tab_input;   // table input index -> session index / input index inside the session
    uint32_t n_sync_sessions = 0;
    std::vector<uint8_t> present;
    uint32_t cur = 0;                           // OUTPUT arena the next tick reads back into (ping-pong)
    int last = -1;                              // arena of the last submitted tick
    bool in_flight = false;
    bool out_ready = false;                     // host_out[last] holds a finished tick
    uint64_t epoch = 1;                         // bumped by every table rebuild
    uint64_t arena_epoch[2] = {0, 0};           // table epoch the tick in each output arena was submitted with
    std::vector<int64_t> first_by_arena[2];     // per session: first table input of that tick (-1: not part of it)
    std::vector<uint32_t> count_by_arena[2];
    std::atomic<uint64_t> ticks{0};             // ticks submitted; pushers read it (under the shared lock) to find their arena
    uint64_t waited = 0;                        // highest tick number known to have finished (read-back included)
    // THE CUT between "this tick" and "the next one": pushers (push / acquire / commit / push_batch, any thread) hold the lock
    // shared, skgpu_hub_tick holds it exclusively while it consumes the queue counters, copies absent streams and advances
    // `ticks`. A push that races with a tick therefore lands entirely before the cut (mixed by this tick) or entirely after
    // it (queued in the next tick's arena) -- never in the arena that is being uploaded.
    std::shared_mutex cut;
    std::atomic<uint64_t> n_discarded{0}, n_errored{0};   // bumped from pusher threads
    std::string fail_reason;                    // non-empty: the hub is Failed (state.rs:122-186)
    bool degraded = false;
    skgpu_hub_stats stats{};                    // NodeStatsTracker counters of the batched nodes (crates/core/src/stats.rs:131-152)
    uint32_t n_live_sessions = 0, n_live_streams = 0;
    explicit skgpu_hub(uint32_t n_streams) : pushed(n_streams) {}
};

// input arena of the tick that is `ahead` ticks after the next one to be submitted
static inline uint8_t *in_arena(skgpu_hub *h, uint32_t ahead) { return h->host_in[(h->ticks.load(std::memory_order_relaxed) + ahead) % h->R]; }

// queue one chunk of stream sid (skgpu_hub_push and friends). Thread-safe for distinct streams.
// copy == nullptr: only reserve the slot (zero-copy acquire). Returns the slot to write.
static uint8_t *enqueue_slot(skgpu_hub *h, uint32_t sid) {
    const uint64_t off = (uint64_t)sid * h->in_stride;
    const size_t bytes = (size_t)h->streams[sid].chunk * h->C * h->ib;
    uint32_t q = h->pushed[sid].load(std::memory_order_relaxed);
    if (q >= h->J) {
        // queue full: the oldest chunk is dropped, the rest move up (overwrite-oldest, mixer.rs:1195-1201); rare
        for (uint32_t i = 0; i + 1 < h->J; ++i) memcpy(in_arena(h, i) + off, in_arena(h, i + 1) + off, bytes);
        q = h->J - 1;
        h->pushed[sid].store((uint8_t)q, std::memory_order_relaxed);
        h->n_discarded.fetch_add(1, std::memory_order_relaxed);
    }
    return in_arena(h, q) + off;
}

static bool valid_gain(float g) { return std::isfinite(g) && g >= 0.0f && g <= 4.0f; }   // gain.rs:50-66

static skgpu_rc rebuild_tables(skgpu_hub *h) {
    h->groups.clear();
    h->inputs.clear();
    h->tab_stream.clear();
    h->tab_session.clear();
    h->tab_input.clear();
    for (uint32_t si = 0; si < h->sessions.size(); ++si) {
        Session &s = h->sessions[si];
        if (!s.live) continue;
        skgpu_chain_group g{};
        g.out_off = h->out_off + (uint64_t)si * h->out_stride;
        g.first_input = (uint32_t)h->inputs.size();
        g.n_inputs = (uint32_t)s.streams.size();
        g.gain_idx = h->cfg.max_streams + si;
        g.out_channels = (uint16_t)h->C;
        g.flags = (h->cfg.flags & SKGPU_HUB_OUT_S16) ? 1u : 0u;   // SKGPU_MIX_OUT_S16
        s.tab_first = (int64_t)h->inputs.size();
        for (uint32_t sid : s.streams) {
            skgpu_chain_input in{};
            in.in_off = (uint64_t)sid * h->in_stride;
            in.slot = h->streams[sid].slot;
            in.gain_idx = sid;
            in.flags = 1u;   // SKGPU_MIX_IN_UNIQUE: every stream owns its samples
            h->inputs.push_back(in);
            h->tab_stream.push_back(sid);
            h->tab_session.push_back(si);
            h->tab_input.push_back((uint32_t)(h->inputs.size() - 1u - (size_t)s.tab_first));
        }
        h->groups.push_back(g);
    }
    PASS(skgpu_plan_update_chain(h->plan, h->op, h->groups.data(), (uint32_t)h->groups.size(), h->inputs.data(), (uint32_t)h->inputs.size()));
    h->tables_dirty = false;
    h->slices_dirty = true;
    h->epoch += 1;
    return SKGPU_OK;
}

// One evaluation of the sync-mode state machine of audio::mixer (mixer.rs:554-918) for tick number T. Decides whether the
// session mixes at this tick (s.st.mixed, s.consumed) and maintains has_sent / slow / waiting_since like the reference's
// InputSlot flags; arrivals between two ticks are observed at the later one.
static void sync_evaluate(skgpu_hub *h, Session &s, int64_t T) {
    const uint32_t n = (uint32_t)s.streams.size();
    const uint64_t all = n >= 64 ? ~0ull : ((1ull << n) - 1ull);
    const uint64_t active = all & ~s.eof;
    s.st.newly_slow = s.st.recovered = 0;
    s.st.mixed = 0;
    s.consumed = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t bit = 1ull << i;
        if (!(active & bit)) continue;
        if (h->pushed[s.streams[i]].load(std::memory_order_acquire) != 0 && !(s.have_frame & bit)) {
            // RecvResult::Audio (:717-745): the first frame into an empty mix starts the timeout clock when more than one input is expected
            if ((s.have_frame & active) == 0 && __builtin_popcountll(active & ~s.slow) > 1) s.waiting_since = T;
            s.have_frame |= bit;
            s.has_sent |= bit;
        }
    }
    const uint64_t frames = s.have_frame & active;
    bool mix = false;
    if (active != 0 && frames != 0) {
        const bool cold_start_complete = (s.has_sent & active) == active;                     // :728-731
        if (s.force_mix) {
            mix = true;                                                                       // PinEof with frames buffered (:872-886)
        } else if (cold_start_complete) {
            if (((s.slow | s.have_frame) & active) == active) {                               // ready_to_mix (:746)
                s.st.recovered = s.slow & frames;                                             // :748-762
                s.slow &= ~s.st.recovered;
                s.waiting_since = -1;
                mix = true;
            } else if (s.timeout_ticks && s.waiting_since >= 0 && T - s.waiting_since >= (int64_t)s.timeout_ticks) {
                const uint64_t missing = active & ~s.slow & ~s.have_frame;                    // :782-838 / :639-709
                if (missing) {
                    s.slow |= missing;
                    s.st.newly_slow = missing;
                    h->stats.discarded += (uint64_t)__builtin_popcountll(missing);
                    s.waiting_since = -1;
                    mix = true;
                }
            }
        }
    }
    s.force_mix = false;
    if (mix) {
        s.st.mixed = 1;
        s.consumed = frames;
        s.have_frame &= ~frames;
    }
    s.st.slow_mask = s.slow & active;
    s.st.eof_mask = s.eof;
    s.st.state = active == 0 ? SKGPU_SESSION_STOPPED : (s.st.slow_mask ? SKGPU_SESSION_DEGRADED : SKGPU_SESSION_RUNNING);
}

extern "C" const char *skgpu_hub_last_error(void) { return g_err.c_str(); }

extern "C" skgpu_rc skgpu_hub_create(int32_t device, const skgpu_hub_config *cfg, skgpu_hub **out) {
    if (!cfg || !out) return hub_fail(SKGPU_ERR_INVALID, "null argument");
    if (!cfg->max_sessions || !cfg->max_streams || !cfg->n_in_rates || !cfg->in_rates) return hub_fail(SKGPU_ERR_INVALID, "empty capacity or rate list");
    if (cfg->channels != 1 && cfg->channels != 2) return hub_fail(SKGPU_ERR_INVALID, "channels must be 1 or 2");
    if (cfg->max_inputs_per_session < 1 || cfg->max_inputs_per_session > 64) return hub_fail(SKGPU_ERR_INVALID, "max_inputs_per_session must be 1..64");
    if (cfg->n_in_rates > 64) return hub_fail(SKGPU_ERR_INVALID, "at most 64 distinct input rates");
    if (cfg->jitter_frames > 8) return hub_fail(SKGPU_ERR_INVALID, "jitter_frames must be 0 (= 1) .. 8");
    uint32_t max_chunk = 0;
    for (uint32_t i = 0; i < cfg->n_in_rates; ++i) {
        const uint64_t num = (uint64_t)cfg->in_rates[i] * cfg->out_frames;
        if (!cfg->in_rates[i] || num % cfg->out_rate) return hub_fail(SKGPU_ERR_INVALID, "input rate %u does not give whole chunks of %u output frames at %u Hz", cfg->in_rates[i], cfg->out_frames, cfg->out_rate);
        max_chunk = std::max<uint32_t>(max_chunk, (uint32_t)(num / cfg->out_rate));
    }
    skgpu_hub *h = new skgpu_hub(cfg->max_streams);
    h->cfg = *cfg;
    h->rates.assign(cfg->in_rates, cfg->in_rates + cfg->n_in_rates);
    h->cfg.in_rates = h->rates.data();
    h->C = cfg->channels;
    h->F = cfg->out_frames;
    h->ob = (cfg->flags & SKGPU_HUB_OUT_S16) ? 2u : 4u;
    h->ib = (cfg->flags & SKGPU_HUB_IN_S16) ? 2u : 4u;
    h->n_slices = cfg->slices ? cfg->slices : 16u;
    if (h->ib == 2u && (((uint64_t)max_chunk * h->C) % 2u || (uint64_t)(max_chunk + 32u) * h->C > 4096u))
        { delete h; return hub_fail(SKGPU_ERR_INVALID, "s16 inputs need chunks of an even number of samples, at most 4096 with the 32-frame head"); }
    h->in_stride = align_up((uint64_t)max_chunk * h->C * h->ib + (h->ib == 2u ? 16u : 0u), 16);   // s16 rows are copied in whole 16-byte units
    h->in_bytes = (uint64_t)cfg->max_streams * h->in_stride;
    h->bank_stride = align_up(h->in_bytes, 256);
    h->res_off = 2 * h->bank_stride;
    h->res_bytes = align_up((uint64_t)cfg->max_streams * sizeof(skgpu_chain_result), 256);
    h->out_off = h->res_off + h->res_bytes;
    h->out_stride = align_up((uint64_t)h->F * h->C * h->ob, 16);
    h->d2h_bytes = h->res_bytes + (uint64_t)cfg->max_sessions * h->out_stride;
    const uint64_t arena = align_up(h->res_off + h->d2h_bytes, 256);
    auto bail = [&](skgpu_rc rc) { skgpu_hub_destroy(h); return rc; };

    skgpu_ctx_config cc{};
    cc.max_streams = cfg->max_streams + cfg->n_in_rates;   // + the probe streams that size the op
    cc.max_channels = h->C;
    cc.fifo_frames = 0;
    skgpu_rc rc0 = skgpu_ctx_create(device, &cc, &h->ctx);
    if (rc0 != SKGPU_OK) return bail(hub_pass(rc0));   // no CUDA device: SKGPU_ERR_NODEVICE, there is no CPU fallback
    rc0 = skgpu_plan_create(h->ctx, arena, &h->plan);
    if (rc0 != SKGPU_OK) return bail(hub_pass(rc0));
    h->J = cfg->jitter_frames ? cfg->jitter_frames : 1u;
    h->R = h->J + 2u;   // J queued ticks + up to two ticks in flight (pipelined collection)
    for (uint32_t b = 0; b < h->R; ++b) {
        void *p = nullptr;
        if (skgpu_pinned_alloc(h->ctx, h->in_bytes, &p) != SKGPU_OK) return bail(hub_pass(SKGPU_ERR_NOMEM));
        h->host_in[b] = (uint8_t *)p;
        memset(p, 0, h->in_bytes);
    }
    for (int b = 0; b < 2; ++b) {
        void *p = nullptr;
        if (skgpu_pinned_alloc(h->ctx, h->d2h_bytes, &p) != SKGPU_OK) return bail(hub_pass(SKGPU_ERR_NOMEM));
        h->host_out[b] = (uint8_t *)p;
        memset(p, 0, h->d2h_bytes);
    }
    h->streams.resize(cfg->max_streams);
    h->sessions.resize(cfg->max_sessions);
    for (uint32_t i = cfg->max_streams; i-- > 0;) h->free_streams.push_back(i);
    for (uint32_t i = cfg->max_sessions; i-- > 0;) h->free_sessions.push_back(i);
    h->gains.assign((size_t)cfg->max_streams + cfg->max_sessions, 1.0f);
    h->present.assign(cfg->max_streams, 0);

    // ---- the op is sized from one probe stream per allowed rate (staging buffers, frame-program capacities), with the
    // table capacities of the whole hub; afterwards the tables are emptied and the probe streams closed
    std::vector<uint32_t> probe_slots(cfg->n_in_rates);
    std::vector<skgpu_chain_input> pin(cfg->n_in_rates);
    for (uint32_t i = 0; i < cfg->n_in_rates; ++i) {
        skgpu_stream_cfg sc{};
        sc.in_rate = cfg->in_rates[i];
        sc.out_rate = cfg->out_rate;
        sc.chunk_frames = (uint32_t)((uint64_t)cfg->in_rates[i] * cfg->out_frames / cfg->out_rate);
        sc.channels = (uint16_t)h->C;
        sc.flags = h->ib == 2u ? SKGPU_STREAM_S16 : 0u;
        if (skgpu_stream_open(h->ctx, &sc, &probe_slots[i]) != SKGPU_OK) return bail(hub_pass(SKGPU_ERR_INVALID));
        pin[i] = skgpu_chain_input{};
        pin[i].in_off = 0;
        pin[i].slot = probe_slots[i];
        pin[i].gain_idx = SKGPU_NO_GAIN;
        pin[i].flags = 1u;
    }
    skgpu_chain_group pg{};
    pg.out_off = h->out_off;
    pg.first_input = 0;
    pg.n_inputs = cfg->n_in_rates;
    pg.gain_idx = SKGPU_NO_GAIN;
    pg.out_channels = (uint16_t)h->C;
    pg.flags = (cfg->flags & SKGPU_HUB_OUT_S16) ? 1u : 0u;
    skgpu_rc rc = skgpu_plan_set_io(h->plan, 0, h->in_bytes, h->res_off, h->d2h_bytes);
    if (rc == SKGPU_OK) rc = skgpu_plan_set_banks(h->plan, h->bank_stride);
    if (rc == SKGPU_OK) rc = skgpu_plan_set_gains(h->plan, h->gains.data(), (uint32_t)h->gains.size());
    if (rc == SKGPU_OK)
        rc = skgpu_plan_add_chain_cap(h->plan, &pg, 1, pin.data(), cfg->n_in_rates, cfg->max_sessions, cfg->max_streams, cfg->max_inputs_per_session,
                                      cfg->out_frames, h->res_off, &h->op);
    if (rc == SKGPU_OK) rc = skgpu_plan_finalize(h->plan);
    if (rc == SKGPU_OK) rc = skgpu_plan_update_chain(h->plan, h->op, nullptr, 0, nullptr, 0);
    if (rc != SKGPU_OK) return bail(hub_pass(rc));
    for (uint32_t s : probe_slots) skgpu_stream_close(h->ctx, s);
    h->tables_dirty = false;
    *out = h;
    return SKGPU_OK;
}

extern "C" void skgpu_hub_destroy(skgpu_hub *h) {
    if (!h) return;
    if (h->plan && h->in_flight) skgpu_tick_wait(h->plan, nullptr);
    if (h->ctx) {
        for (uint32_t b = 0; b < skgpu_hub::MAX_RING; ++b)
            if (h->host_in[b]) skgpu_pinned_free(h->ctx, h->host_in[b]);
        for (int b = 0; b < 2; ++b)
            if (h->host_out[b]) skgpu_pinned_free(h->ctx, h->host_out[b]);
    }
    if (h->plan) skgpu_plan_destroy(h->plan);
    if (h->ctx) skgpu_ctx_destroy(h->ctx);
    delete h;
}

extern "C" skgpu_rc skgpu_hub_session_open(skgpu_hub *h, uint32_t n_inputs, const uint32_t *in_rates, uint32_t *session_out) {
    return skgpu_hub_session_open_ex(h, n_inputs, in_rates, 0u, 0u, session_out);
}

extern "C" skgpu_rc skgpu_hub_session_open_ex(skgpu_hub *h, uint32_t n_inputs, const uint32_t *in_rates, uint32_t mode, uint32_t sync_timeout_ms,
                                              uint32_t *session_out) {
    if (!h || !in_rates || !session_out) return hub_fail(SKGPU_ERR_INVALID, "null argument");
    if (mode & ~SKGPU_SESSION_SYNC) return hub_fail(SKGPU_ERR_INVALID, "unknown session mode 0x%x", mode);
    if ((mode & SKGPU_SESSION_SYNC) && h->J != 1u) return hub_fail(SKGPU_ERR_INVALID, "sync-mode sessions keep the latest frame per input: the hub must have jitter_frames = 1");
    if (n_inputs < 1 || n_inputs > h->cfg.max_inputs_per_session) return hub_fail(SKGPU_ERR_INVALID, "n_inputs %u outside 1..%u", n_inputs, h->cfg.max_inputs_per_session);
    if (h->free_sessions.empty()) return hub_fail(SKGPU_ERR_NOMEM, "out of session slots (%u)", h->cfg.max_sessions);
    if (h->free_streams.size() < n_inputs) return hub_fail(SKGPU_ERR_NOMEM, "out of stream slots (%u requested, %zu free)", n_inputs, h->free_streams.size());
    for (uint32_t i = 0; i < n_inputs; ++i) {
        bool ok = false;
        for (uint32_t r : h->rates) ok |= (r == in_rates[i]);
        if (!ok) return hub_fail(SKGPU_ERR_INVALID, "input %u: sample rate %u is not in the hub's rate list", i, in_rates[i]);
    }
    const uint32_t si = h->free_sessions.back();
    Session s;
    s.live = true;
    s.sync = (mode & SKGPU_SESSION_SYNC) != 0;
    if (s.sync && sync_timeout_ms) {
        const double tick_ms = 1000.0 * (double)h->F / (double)h->cfg.out_rate;
        s.timeout_ticks = (uint32_t)std::ceil((double)sync_timeout_ms / tick_ms);
    }
    for (uint32_t i = 0; i < n_inputs; ++i) {
        skgpu_stream_cfg sc{};
        sc.in_rate = in_rates[i];
        sc.out_rate = h->cfg.out_rate;
        sc.chunk_frames = (uint32_t)((uint64_t)in_rates[i] * h->F / h->cfg.out_rate);
        sc.channels = (uint16_t)h->C;
        sc.flags = h->ib == 2u ? SKGPU_STREAM_S16 : 0u;
        uint32_t slot = 0;
        const skgpu_rc rc = skgpu_stream_open(h->ctx, &sc, &slot);
        if (rc != SKGPU_OK) {
            for (uint32_t sid : s.streams) { skgpu_stream_close(h->ctx, h->streams[sid].slot); h->streams[sid] = Stream{}; h->free_streams.push_back(sid); }
            return hub_pass(rc);
        }
        const uint32_t sid = h->free_streams.back();
        h->free_streams.pop_back();
        Stream st;
        st.slot = slot; st.chunk = sc.chunk_frames; st.in_rate = in_rates[i]; st.live = true;
        h->streams[sid] = st;
        h->pushed[sid].store(0, std::memory_order_relaxed);
        h->gains[sid] = 1.0f;
        s.streams.push_back(sid);
    }
    h->free_sessions.pop_back();
    h->gains[h->cfg.max_streams + si] = 1.0f;
    h->sessions[si] = std::move(s);
    h->n_live_sessions += 1;
    h->n_live_streams += n_inputs;
    if (h->sessions[si].sync) h->n_sync_sessions += 1;
    h->gains_dirty = h->tables_dirty = true;
    *session_out = si;
    return SKGPU_OK;
}

static Session *live_session(skgpu_hub *h, uint32_t si) { return (h && si < h->sessions.size() && h->sessions[si].live) ? &h->sessions[si] : nullptr; }

extern "C" skgpu_rc skgpu_hub_session_close(skgpu_hub *h, uint32_t si) {
    Session *s = live_session(h, si);
    if (!s) return hub_fail(SKGPU_ERR_INVALID, "session %u is not open", si);
    if (h->in_flight) { PASS(skgpu_tick_wait(h->plan, nullptr)); h->in_flight = false; }   // its slots may be in use by the tick in flight
    for (uint32_t sid : s->streams) {
        skgpu_stream_close(h->ctx, h->streams[sid].slot);
        h->streams[sid] = Stream{};
        h->pushed[sid].store(0, std::memory_order_relaxed);
        h->free_streams.push_back(sid);
    }
    h->n_live_streams -= (uint32_t)s->streams.size();
    h->n_live_sessions -= 1;
    if (s->sync) h->n_sync_sessions -= 1;
    *s = Session{};
    for (int a = 0; a < 2; ++a)   // a session index that is reused must not see the closed session's packets
        if (si < h->first_by_arena[a].size()) h->first_by_arena[a][si] = -1;
    h->free_sessions.push_back(si);
    h->tables_dirty = true;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_set_input_gain(skgpu_hub *h, uint32_t si, uint32_t input, float gain) {
    Session *s = live_session(h, si);
    if (!s || input >= s->streams.size()) return hub_fail(SKGPU_ERR_INVALID, "no such session / input");
    if (!valid_gain(gain)) { h->n_errored.fetch_add(1, std::memory_order_relaxed); return hub_fail(SKGPU_ERR_INVALID, "gain must be a finite number between 0.0 and 4.0"); }   // gain.rs:50-66, :157-171; the old gain stays
    h->gains[s->streams[input]] = gain;
    h->gains_dirty = true;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_set_master_gain(skgpu_hub *h, uint32_t si, float gain) {
    if (!live_session(h, si)) return hub_fail(SKGPU_ERR_INVALID, "session %u is not open", si);
    if (!valid_gain(gain)) { h->n_errored.fetch_add(1, std::memory_order_relaxed); return hub_fail(SKGPU_ERR_INVALID, "gain must be a finite number between 0.0 and 4.0"); }
    h->gains[h->cfg.max_streams + si] = gain;
    h->gains_dirty = true;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_chunk_frames(skgpu_hub *h, uint32_t si, uint32_t input, uint32_t *frames_out) {
    Session *s = live_session(h, si);
    if (!s || input >= s->streams.size() || !frames_out) return hub_fail(SKGPU_ERR_INVALID, "no such session / input");
    *frames_out = h->streams[s->streams[input]].chunk;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_push(skgpu_hub *h, uint32_t si, uint32_t input, const void *samples, uint32_t n_frames) {
    Session *s = live_session(h, si);
    if (!s || input >= s->streams.size() || !samples) return hub_fail(SKGPU_ERR_INVALID, "no such session / input");
    const uint32_t sid = s->streams[input];
    const Stream &st = h->streams[sid];
    if (n_frames != st.chunk) return hub_fail(SKGPU_ERR_INVALID, "chunk of %u frames, the stream delivers %u per tick", n_frames, st.chunk);
    std::shared_lock<std::shared_mutex> lk(h->cut);
    stream_copy(enqueue_slot(h, sid), reinterpret_cast<const uint8_t *>(samples), (size_t)n_frames * h->C * h->ib);
    stream_fence();
    h->pushed[sid].fetch_add(1, std::memory_order_release);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_acquire(skgpu_hub *h, uint32_t si, uint32_t input, void **dst_out, uint32_t *n_frames_out) {
    Session *s = live_session(h, si);
    if (!s || input >= s->streams.size() || !dst_out) return hub_fail(SKGPU_ERR_INVALID, "no such session / input");
    const uint32_t sid = s->streams[input];
    std::shared_lock<std::shared_mutex> lk(h->cut);
    *dst_out = enqueue_slot(h, sid);
    h->streams[sid].acq_tick = h->ticks.load(std::memory_order_relaxed);
    if (n_frames_out) *n_frames_out = h->streams[sid].chunk;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_commit(skgpu_hub *h, uint32_t si, uint32_t input) {
    Session *s = live_session(h, si);
    if (!s || input >= s->streams.size()) return hub_fail(SKGPU_ERR_INVALID, "no such session / input");
    const uint32_t sid = s->streams[input];
    std::shared_lock<std::shared_mutex> lk(h->cut);
    if (h->streams[sid].acq_tick != ~0ull && h->streams[sid].acq_tick != h->ticks.load(std::memory_order_relaxed)) {
        h->streams[sid].acq_tick = ~0ull;
        h->n_errored.fetch_add(1, std::memory_order_relaxed);
        return hub_fail(SKGPU_ERR_STATE, "commit after a tick: the slot handed out by skgpu_hub_acquire belonged to an earlier tick; chunk dropped");
    }
    h->streams[sid].acq_tick = ~0ull;
    if (h->pushed[sid].load(std::memory_order_relaxed) < h->J) h->pushed[sid].fetch_add(1, std::memory_order_release);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_commit_all(skgpu_hub *h) {
    if (!h) return hub_fail(SKGPU_ERR_INVALID, "null hub");
    std::shared_lock<std::shared_mutex> lk(h->cut);
    for (uint32_t sid = 0; sid < h->streams.size(); ++sid)
        if (h->streams[sid].live && h->pushed[sid].load(std::memory_order_relaxed) == 0) h->pushed[sid].store(1, std::memory_order_relaxed);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_push_batch(skgpu_hub *h, const skgpu_hub_frame *frames, uint32_t n, uint32_t n_threads) {
    if (!h || (!frames && n)) return hub_fail(SKGPU_ERR_INVALID, "null argument");
    // validate on the calling thread (errors are thread-local), copy on the workers
    for (uint32_t i = 0; i < n; ++i) {
        Session *s = live_session(h, frames[i].session);
        if (!s || frames[i].input >= s->streams.size() || !frames[i].samples) return hub_fail(SKGPU_ERR_INVALID, "frame %u: no such session / input", i);
        if (frames[i].n_frames != h->streams[s->streams[frames[i].input]].chunk)
            return hub_fail(SKGPU_ERR_INVALID, "frame %u: chunk of %u frames, the stream delivers %u per tick", i, frames[i].n_frames, h->streams[s->streams[frames[i].input]].chunk);
    }
    auto work = [&](uint32_t lo, uint32_t hi) {
        std::shared_lock<std::shared_mutex> lk(h->cut);
        for (uint32_t i = lo; i < hi; ++i) {
            const uint32_t sid = h->sessions[frames[i].session].streams[frames[i].input];
            stream_copy(enqueue_slot(h, sid), reinterpret_cast<const uint8_t *>(frames[i].samples), (size_t)frames[i].n_frames * h->C * h->ib);
            h->pushed[sid].fetch_add(1, std::memory_order_relaxed);
        }
        stream_fence();
    };
    const uint32_t T = std::max(1u, std::min(n_threads, n / 256u + 1u));
    if (T == 1) { work(0, n); return SKGPU_OK; }
    std::vector<std::thread> pool;
    const uint32_t per = (n + T - 1) / T;
    for (uint32_t t = 0; t < T; ++t) pool.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
    for (auto &th : pool) th.join();
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_tick(skgpu_hub *h) {
    if (!h) return hub_fail(SKGPU_ERR_INVALID, "null hub");
    if (!h->fail_reason.empty()) return hub_fail(SKGPU_ERR_STATE, "hub failed: %s", h->fail_reason.c_str());
    const uint64_t t0 = h->ticks.load(std::memory_order_relaxed);
    // at most two unfinished ticks: the input ring (J + 2 arenas) and the two output arenas are sized for exactly that
    // (a third submit would let pushes overwrite an arena whose upload is pending and reuse host_out[cur])
    if (t0 >= 2 && h->waited + 2 <= t0) {
        PASS(skgpu_tick_wait_for(h->plan, t0 - 1));
        h->waited = t0 - 1;
    }
    if (h->tables_dirty) {
        if (h->in_flight) { PASS(skgpu_tick_wait(h->plan, nullptr)); h->in_flight = false; h->waited = t0; }
        skgpu_rc rc = rebuild_tables(h);
        if (rc != SKGPU_OK) return rc;
    }
    // ---- the cut: from here to the advance of `ticks` no push can interleave
    std::unique_lock<std::shared_mutex> cut(h->cut);
    const uint32_t n_in = (uint32_t)h->tab_stream.size();
    uint8_t *in_cur = in_arena(h, 0), *in_prev = h->host_in[(t0 + h->R - 1u) % h->R];   // this tick's arena, the previous tick's
    uint32_t n_sent = 0;
    if (h->n_sync_sessions) {
        for (Session &s : h->sessions)
            if (s.live && s.sync) sync_evaluate(h, s, (int64_t)t0 + 1);
    }
    h->degraded = false;
    for (const Session &s : h->sessions) {
        if (!s.live) continue;
        if (!s.sync) n_sent += 1;
        else { n_sent += s.st.mixed; h->degraded |= (s.st.state & SKGPU_SESSION_DEGRADED) != 0; }
    }
    uint8_t *in_next = in_arena(h, 1);
    for (uint32_t i = 0; i < n_in; ++i) {
        const uint32_t sid = h->tab_stream[i];
        const Stream &st = h->streams[sid];
        bool got = h->pushed[sid].load(std::memory_order_acquire) != 0;   // the head of the stream's queue sits in this tick's arena
        const Session &ss = h->sessions[h->tab_session[i]];
        if (ss.sync && got && !(ss.st.mixed && ((ss.consumed >> h->tab_input[i]) & 1ull))) {
            // sync mode holds this frame: it stays queued for a later mix (its chunk moves on to the next tick's arena); for the
            // kernel the stream is absent
            const uint64_t off = (uint64_t)sid * h->in_stride;
            memcpy(in_next + off, in_cur + off, (size_t)st.chunk * h->C * h->ib);
            got = false;
        }
        h->present[i] = got ? 1 : 0;
        if (!got && st.ever_pushed) {
            // absent this tick: the other input bank must keep holding the stream's previous chunk (skgpu_batch.h, chain protocol)
            const uint64_t off = (uint64_t)sid * h->in_stride;
            memcpy(in_cur + off, in_prev + off, (size_t)st.chunk * h->C * h->ib);
        }
    }
    PASS(skgpu_plan_set_present(h->plan, h->op, h->present.data(), n_in));
    if (h->gains_dirty) {
        PASS(skgpu_plan_set_gains(h->plan, h->gains.data(), (uint32_t)h->gains.size()));
        h->gains_dirty = false;
    }
    const bool sliced = h->n_slices > 1u && !h->groups.empty();
    if (sliced && h->slices_dirty) {
        PASS(skgpu_plan_auto_slices(h->plan, h->op, h->n_slices));
        h->slices_dirty = false;
    }
    PASS_FATAL(h, skgpu_tick_submit(h->plan, in_cur, h->host_out[h->cur], sliced ? SKGPU_SUBMIT_SLICED : (SKGPU_SUBMIT_GRAPH | SKGPU_SUBMIT_OVERLAP_D2H)));
    // only a tick that was really submitted consumes the queues (a failed submit keeps every queued chunk)
    for (uint32_t i = 0; i < n_in; ++i) {
        if (!h->present[i]) continue;
        const uint32_t sid = h->tab_stream[i];
        h->pushed[sid].fetch_sub(1, std::memory_order_relaxed);
        h->streams[sid].ever_pushed = true;
        h->stats.received += 1;
    }
    h->stats.sent += n_sent;
    if (h->last == (int)h->cur) h->out_ready = false;   // this tick reuses the arena the collected results lived in
    if (h->arena_epoch[h->cur] != h->epoch) {           // remember which table rows this tick's results belong to
        auto &fa = h->first_by_arena[h->cur];
        auto &ca = h->count_by_arena[h->cur];
        fa.assign(h->sessions.size(), -1);
        ca.assign(h->sessions.size(), 0);
        for (size_t si = 0; si < h->sessions.size(); ++si)
            if (h->sessions[si].live) { fa[si] = h->sessions[si].tab_first; ca[si] = (uint32_t)h->sessions[si].streams.size(); }
        h->arena_epoch[h->cur] = h->epoch;
    }
    h->cur ^= 1u;
    h->in_flight = true;
    h->ticks.store(t0 + 1, std::memory_order_release);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_wait(skgpu_hub *h, skgpu_tick_timing *timing) {
    if (!h) return hub_fail(SKGPU_ERR_INVALID, "null hub");
    if (!h->fail_reason.empty()) return hub_fail(SKGPU_ERR_STATE, "hub failed: %s", h->fail_reason.c_str());
    PASS_FATAL(h, skgpu_tick_wait(h->plan, timing));
    h->in_flight = false;
    h->waited = h->ticks;
    h->out_ready = h->ticks > 0;
    if (h->ticks > 0) h->last = (int)((h->ticks - 1) & 1u);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_wait_tick(skgpu_hub *h, uint64_t tick) {
    if (!h) return hub_fail(SKGPU_ERR_INVALID, "null hub");
    if (tick == 0 || tick > h->ticks || tick + 1 < h->ticks) return hub_fail(SKGPU_ERR_INVALID, "tick %llu is not one of the two most recent ticks", (unsigned long long)tick);
    PASS_FATAL(h, skgpu_tick_wait_for(h->plan, tick));
    h->waited = std::max(h->waited, tick);
    h->last = (int)((tick - 1) & 1u);          // tick k uploaded from / read back into arena (k - 1) & 1
    h->in_flight = tick < h->ticks;           // a later tick may still be running
    h->out_ready = true;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_session_output(skgpu_hub *h, uint32_t si, const void **samples, uint32_t *n_mixed, uint32_t *status) {
    Session *s = live_session(h, si);
    if (!s) return hub_fail(SKGPU_ERR_INVALID, "session %u is not open", si);
    if (!h->out_ready) return hub_fail(SKGPU_ERR_STATE, "no finished tick: call skgpu_hub_wait / skgpu_hub_wait_tick first");
    if (samples) *samples = nullptr;
    if (n_mixed) *n_mixed = 0;
    if (status) *status = 0;
    if (h->last < 0 || si >= h->first_by_arena[h->last].size() || h->first_by_arena[h->last][si] < 0) return SKGPU_OK;   // not part of that tick
    const uint8_t *o = h->host_out[h->last];
    const skgpu_chain_result *res = reinterpret_cast<const skgpu_chain_result *>(o) + h->first_by_arena[h->last][si];
    uint32_t nm = 0, stt = 0;
    for (uint32_t i = 0; i < h->count_by_arena[h->last][si]; ++i) { nm += res[i].emitted; stt |= res[i].status; }
    if (samples) *samples = o + (h->out_off - h->res_off) + (uint64_t)si * h->out_stride;
    if (n_mixed) *n_mixed = nm;
    if (status) *status = stt;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_input_eof(skgpu_hub *h, uint32_t si, uint32_t input) {
    Session *s = live_session(h, si);
    if (!s || input >= s->streams.size()) return hub_fail(SKGPU_ERR_INVALID, "no such session / input");
    if (!s->sync) return hub_fail(SKGPU_ERR_STATE, "session %u is a clocked-mode session: an input that stops delivering is simply silent there", si);
    std::unique_lock<std::shared_mutex> lk(h->cut);
    const uint64_t bit = 1ull << input;
    if (s->eof & bit) return SKGPU_OK;
    s->eof |= bit;                                   // slots.remove(slot_idx) (mixer.rs:856)
    s->have_frame &= ~bit;
    h->pushed[s->streams[input]].store(0, std::memory_order_relaxed);
    s->waiting_since = -1;                           // :866
    const uint32_t n = (uint32_t)s->streams.size();
    const uint64_t active = (n >= 64 ? ~0ull : ((1ull << n) - 1ull)) & ~s->eof;
    if (active && (s->have_frame & active)) s->force_mix = true;   // :872-886
    s->st.eof_mask = s->eof;
    s->st.slow_mask = s->slow & active;
    s->st.state = active == 0 ? SKGPU_SESSION_STOPPED : (s->st.slow_mask ? SKGPU_SESSION_DEGRADED : SKGPU_SESSION_RUNNING);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_session_state(skgpu_hub *h, uint32_t si, skgpu_session_state *out) {
    Session *s = live_session(h, si);
    if (!s || !out) return hub_fail(SKGPU_ERR_INVALID, "no such session");
    *out = s->st;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_hub_get_stats(skgpu_hub *h, skgpu_hub_stats *out) {
    if (!h || !out) return hub_fail(SKGPU_ERR_INVALID, "null argument");
    *out = h->stats;
    out->discarded += h->n_discarded.load(std::memory_order_relaxed);
    out->errored += h->n_errored.load(std::memory_order_relaxed);
    return SKGPU_OK;
}

extern "C" uint32_t skgpu_hub_state(const skgpu_hub *h, const char **reason_out) {
    if (reason_out) *reason_out = (h && !h->fail_reason.empty()) ? h->fail_reason.c_str() : nullptr;
    if (!h) return SKGPU_HUB_FAILED;
    if (!h->fail_reason.empty()) return SKGPU_HUB_FAILED;
    return h->degraded ? SKGPU_HUB_DEGRADED : SKGPU_HUB_RUNNING;
}

extern "C" int32_t skgpu_hub_bind_thread(skgpu_hub *h) {
    if (!h || !h->ctx) return -1;
    if (skgpu_ctx_bind_thread(h->ctx) != SKGPU_OK) return -1;
    return skgpu_ctx_numa_node(h->ctx);
}

extern "C" uint32_t skgpu_hub_live_sessions(const skgpu_hub *h) { return h ? h->n_live_sessions : 0; }
extern "C" uint32_t skgpu_hub_live_streams(const skgpu_hub *h) { return h ? h->n_live_streams : 0; }
extern "C" uint64_t skgpu_hub_ticks(const skgpu_hub *h) { return h ? h->ticks.load(std::memory_order_relaxed) : 0; }
