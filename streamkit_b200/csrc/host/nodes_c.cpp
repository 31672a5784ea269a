// nodes_c.cpp -- C shim over the node mirror for the pieces the single-input plugin ABI cannot carry
// (audio::mixer: wrapper.rs:224,410 supports exactly one input pin). Used by the pytest suite; a Rust host would
// bind the batch ABI directly instead.
#include <cstring>
#include <string>
#include <vector>

#include "nodes.hpp"

using namespace skhost;

static thread_local std::string g_err;

extern "C" {

const char *skn_last_error(void) { return g_err.c_str(); }

void *skn_mixer_create(const char *params_json) {
    StreamKitError err{StreamKitError::Configuration, ""};
    auto n = AudioMixerNode::create(params_json, &err);
    if (!n) { g_err = err.message; return nullptr; }
    return n.release();
}
void skn_mixer_destroy(void *h) { delete static_cast<AudioMixerNode *>(h); }
uint32_t skn_mixer_num_input_pins(void *h) { return (uint32_t)static_cast<AudioMixerNode *>(h)->input_pins().size(); }
int skn_mixer_pin_name(void *h, uint32_t i, char *buf, size_t n) {
    auto pins = static_cast<AudioMixerNode *>(h)->input_pins();
    if (i >= pins.size()) return -1;
    std::strncpy(buf, pins[i].name.c_str(), n);
    return 0;
}

struct skn_frame { const float *samples; uint32_t n_samples; uint32_t sample_rate; uint16_t channels; uint16_t unique; };

// clocked = 0: mix_and_send arithmetic (sync mode, sticky channels); clocked = 1: mix_clocked_frames.
// out must hold out_cap floats; returns 0 and fills *out_len / *out_channels / *out_rate, or -1 (see skn_last_error)
int skn_mixer_mix(void *h, int clocked, const skn_frame *frames, uint32_t n, float *out, size_t out_cap, size_t *out_len,
                  uint16_t *out_channels, uint32_t *out_rate) {
    auto *node = static_cast<AudioMixerNode *>(h);
    std::vector<AudioFrame> fr(n);
    for (uint32_t i = 0; i < n; ++i) {
        fr[i].sample_rate = frames[i].sample_rate;
        fr[i].channels = frames[i].channels;
        fr[i].samples.assign(frames[i].samples, frames[i].samples + frames[i].n_samples);
        fr[i].unique = frames[i].unique != 0;
    }
    AudioFrame o;
    StreamKitError err{StreamKitError::Runtime, ""};
    try {
        const bool ok = clocked ? node->mix_clocked(fr, o, &err) : node->mix(fr, o, &err);
        if (!ok) { g_err = err.message; return -1; }
    } catch (const StreamKitError &e) {
        g_err = e.message;
        return -1;
    }
    if (o.samples.size() > out_cap) { g_err = "output buffer too small"; return -1; }
    std::memcpy(out, o.samples.data(), o.samples.size() * sizeof(float));
    *out_len = o.samples.size();
    *out_channels = o.channels;
    *out_rate = o.sample_rate;
    return 0;
}

// ---- audio::resampler with packet metadata (the native plugin ABI v2 drops metadata, conversions.rs:342-346, so the
// timestamp / duration / sequence stamping of resampler.rs:286-297 and the no-increment-after-flush rule of :707-711 are
// reachable only here)
struct skn_packet_meta { uint64_t timestamp_us, duration_us, sequence; uint8_t has_timestamp, has_duration, has_sequence, pad; };
typedef void (*skn_emit_fn)(void *ud, uint32_t sample_rate, uint16_t channels, const float *samples, size_t n_samples, const skn_packet_meta *meta);

void *skn_resampler_create(const char *params_json) {
    StreamKitError err{StreamKitError::Configuration, ""};
    try {
        auto n = AudioResamplerNode::create(params_json, &err);
        if (!n) { g_err = err.message; return nullptr; }
        return n.release();
    } catch (const StreamKitError &e) {
        g_err = e.message;
        return nullptr;
    }
}
void skn_resampler_destroy(void *h) { delete static_cast<AudioResamplerNode *>(h); }

static void emit_all(const std::vector<AudioFrame> &out, skn_emit_fn emit, void *ud) {
    for (const AudioFrame &f : out) {
        skn_packet_meta m{};
        if (f.metadata) {
            if (f.metadata->timestamp_us) { m.timestamp_us = *f.metadata->timestamp_us; m.has_timestamp = 1; }
            if (f.metadata->duration_us) { m.duration_us = *f.metadata->duration_us; m.has_duration = 1; }
            if (f.metadata->sequence) { m.sequence = *f.metadata->sequence; m.has_sequence = 1; }
        }
        emit(ud, f.sample_rate, f.channels, f.samples.data(), f.samples.size(), &m);
    }
}

// one input packet; has_timestamp = 0 -> the packet carries no metadata. Returns 0, or -1 (fatal for the node: skn_last_error)
int skn_resampler_push(void *h, uint32_t sample_rate, uint16_t channels, const float *samples, size_t n_samples, int has_timestamp,
                       uint64_t timestamp_us, skn_emit_fn emit, void *ud) {
    auto *node = static_cast<AudioResamplerNode *>(h);
    AudioFrame in;
    in.sample_rate = sample_rate;
    in.channels = channels;
    in.samples.assign(samples, samples + n_samples);
    if (has_timestamp) { PacketMetadata md; md.timestamp_us = timestamp_us; in.metadata = md; }
    std::vector<AudioFrame> out;
    StreamKitError err{StreamKitError::Runtime, ""};
    try {
        if (!node->process(in, out, &err)) { g_err = err.message; return -1; }
    } catch (const StreamKitError &e) {
        g_err = e.message;
        return -1;
    }
    emit_all(out, emit, ud);
    return 0;
}
int skn_resampler_finish(void *h, skn_emit_fn emit, void *ud) {
    auto *node = static_cast<AudioResamplerNode *>(h);
    std::vector<AudioFrame> out;
    StreamKitError err{StreamKitError::Runtime, ""};
    try {
        if (!node->finish(out, &err)) { g_err = err.message; return -1; }
    } catch (const StreamKitError &e) {
        g_err = e.message;
        return -1;
    }
    emit_all(out, emit, ud);
    return 0;
}
}
