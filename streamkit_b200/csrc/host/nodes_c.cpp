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
}
