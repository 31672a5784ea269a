// plugin_v3.cpp -- libskgpu_plugin_mixer_v3.so: the GPU audio::mixer behind the PROPOSED native plugin ABI v3
// (include/streamkit_native_abi_v3.h). What v2 cannot express and this plugin uses: a dynamic family of input pins
// ("in_0", "in_1", ...: mixer.rs:122-173), typed s16 audio payloads in and out, packet metadata carried through
// (the mix inherits the first frame's metadata, mixer.rs:994), and a batched process_packets call that hands over
// the frames of all pins at once -- one FFI hop and one GPU submit per mix instead of one per input frame.
//
// params: {"num_inputs": n (pre-created pins, mixer.rs:128-143), "gain": g (audio::gain after the mix, [0, 4]),
//          "output_format": "f32" | "s16"}
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/streamkit_native_abi_v3.h"
#include "mini_json.hpp"
#include "nodes.hpp"

using namespace skhost;

namespace {

thread_local std::string g_err;
sk_result ok() { return sk_result{true, nullptr}; }
sk_result fail(const std::string &m) {
    g_err = m;
    return sk_result{false, g_err.c_str()};
}

struct Instance {
    sk_log_callback log_cb = nullptr;
    void *log_ud = nullptr;
    std::unique_ptr<AudioMixerNode> mixer;
    std::unique_ptr<AudioGainNode> gain;        // f32 output: audio::gain after the mix
    std::unique_ptr<AudioPcm16Node> pcm16;      // s16 output: gain -> clip -> s16 in one pass
    bool out_s16 = false;
    std::vector<std::string> pins;              // active input pins, creation order = mix order
    std::map<std::string, AudioFrame> pending;  // process_packet path: latest frame per pin (slot.frame = Some(frame), mixer.rs:741)
};

const sk_audio_format kAnyF32 = {0, 0, SK_SAMPLE_F32};
const sk_audio_format kAnyS16 = {0, 0, SK_SAMPLE_S16LE};
const sk_packet_type_info kInTypes[] = {{SK_PACKET_RAW_AUDIO, &kAnyF32, nullptr}, {SK_PACKET_RAW_AUDIO, &kAnyS16, nullptr}};
const sk_input_pin_v3 kInputs[] = {{"in", kInTypes, 2, SK_PIN_DYNAMIC}};
const sk_output_pin kOutputs[] = {{"out", {SK_PACKET_RAW_AUDIO, &kAnyF32, nullptr}}};
const char *const kCategories[] = {"audio", "filters", "gpu"};
const char *kSchema =
    "{\"type\":\"object\",\"properties\":{\"num_inputs\":{\"type\":\"integer\",\"minimum\":0},\"gain\":{\"type\":\"number\",\"default\":1.0,"
    "\"minimum\":0.0,\"maximum\":4.0,\"tunable\":true},\"output_format\":{\"type\":\"string\",\"enum\":[\"f32\",\"s16\"],\"default\":\"f32\"}}}";
const sk_node_metadata_v3 kMeta = {"gpu_mixer", "audio::mixer on the GPU (streamkit_b200): ordered N-input sum, channel conversion, gain, optional s16 output",
                                   kInputs, 1, kOutputs, 1, kSchema, kCategories, 3};

const sk_node_metadata_v3 *get_metadata() { return &kMeta; }

sk_plugin_handle create_instance(const char *params_json, sk_log_callback log_cb, void *log_ud) {
    try {
        std::unique_ptr<Instance> inst(new Instance());
        inst->log_cb = log_cb;
        inst->log_ud = log_ud;
        StreamKitError err{StreamKitError::Configuration, ""};
        inst->mixer = AudioMixerNode::create(params_json, &err);
        if (!inst->mixer) return nullptr;
        std::string gain_json = "{}";
        if (params_json && *params_json) {
            JsonValue v;
            std::string perr;
            if (JsonParser(params_json).parse(v, perr) && v.kind == JsonValue::Object) {
                if (const JsonValue *f = v.get("output_format")) inst->out_s16 = f->kind == JsonValue::String && f->str == "s16";
                if (const JsonValue *g = v.get("gain"))
                    if (g->kind == JsonValue::Number) gain_json = "{\"gain\": " + std::to_string(g->num) + "}";
            }
        }
        if (inst->out_s16) inst->pcm16 = AudioPcm16Node::create(gain_json.c_str(), &err);
        else inst->gain = AudioGainNode::create(gain_json.c_str(), &err);
        if (!inst->pcm16 && !inst->gain) {
            if (log_cb) log_cb(SK_LOG_ERROR, "streamkit_b200", err.message.c_str(), log_ud);
            return nullptr;   // gain outside [0, 4] etc. (gain.rs:50-66)
        }
        for (const PinSpec &p : inst->mixer->input_pins()) inst->pins.push_back(p.name);
        GpuRuntime::get();    // fail at creation when there is no GPU: there is no CPU fallback
        return inst.release();
    } catch (...) {
        return nullptr;
    }
}

bool to_frame(const sk_packet_v3 *pkt, AudioFrame &f, std::string &why) {
    if (!pkt || pkt->packet_type != SK_PACKET_RAW_AUDIO || !pkt->data || pkt->len != sizeof(sk_audio_frame_v3)) { why = "not a v3 RawAudio packet"; return false; }
    const auto *af = static_cast<const sk_audio_frame_v3 *>(pkt->data);
    if (!af->samples && af->sample_count) { why = "Invalid audio frame"; return false; }
    if (af->channels == 0 || af->sample_count % af->channels) { why = "sample_count is not a multiple of channels"; return false; }
    f.sample_rate = af->sample_rate;
    f.channels = af->channels;
    f.samples.resize(af->sample_count);
    const size_t frames = af->sample_count / af->channels;
    for (size_t i = 0; i < af->sample_count; ++i) {
        // planar -> interleaved (resampler.rs:413-417 does the same for rubato's planar output)
        const size_t src = af->layout == SK_LAYOUT_PLANAR ? (i % af->channels) * frames + i / af->channels : i;
        f.samples[i] = af->sample_format == SK_SAMPLE_S16LE ? (float)static_cast<const int16_t *>(af->samples)[src] * (1.0f / 32768.0f)   // exact
                                                              : static_cast<const float *>(af->samples)[src];
    }
    if (pkt->metadata) {
        PacketMetadata md;
        if (pkt->metadata->has_timestamp_us) md.timestamp_us = pkt->metadata->timestamp_us;
        if (pkt->metadata->has_duration_us) md.duration_us = pkt->metadata->duration_us;
        if (pkt->metadata->has_sequence) md.sequence = pkt->metadata->sequence;
        f.metadata = md;
    }
    return true;
}

sk_result mix_and_emit(Instance *inst, std::vector<AudioFrame> &frames, sk_output_callback_v3 cb, void *ud) {
    if (frames.empty()) return ok();
    AudioFrame mixed;
    StreamKitError err{StreamKitError::Runtime, ""};
    if (!inst->mixer->mix(frames, mixed, &err)) return fail(err.message);
    sk_packet_metadata md{};
    const sk_packet_metadata *mdp = nullptr;
    if (mixed.metadata) {   // mixer.rs:994: the mix carries the first frame's metadata
        if (mixed.metadata->timestamp_us) { md.timestamp_us = *mixed.metadata->timestamp_us; md.has_timestamp_us = true; }
        if (mixed.metadata->duration_us) { md.duration_us = *mixed.metadata->duration_us; md.has_duration_us = true; }
        if (mixed.metadata->sequence) { md.sequence = *mixed.metadata->sequence; md.has_sequence = true; }
        mdp = &md;
    }
    if (inst->out_s16) {
        std::vector<int16_t> s16;
        if (!inst->pcm16->process(mixed, s16, &err)) return fail(err.message);
        sk_audio_frame_v3 af{mixed.sample_rate, mixed.channels, SK_SAMPLE_S16LE, SK_LAYOUT_INTERLEAVED, s16.data(), s16.size()};
        sk_packet_v3 pkt{SK_PACKET_RAW_AUDIO, &af, sizeof af, mdp};
        return cb("out", &pkt, ud);
    }
    AudioFrame out;
    if (!inst->gain->process(mixed, out, &err)) return fail(err.message);
    sk_audio_frame_v3 af{out.sample_rate, out.channels, SK_SAMPLE_F32, SK_LAYOUT_INTERLEAVED, out.samples.data(), out.samples.size()};
    sk_packet_v3 pkt{SK_PACKET_RAW_AUDIO, &af, sizeof af, mdp};
    return cb("out", &pkt, ud);
}

// frames of `by_pin` in pin order (the order the pins were created in: the build's definition of the mix order)
std::vector<AudioFrame> ordered(Instance *inst, std::map<std::string, AudioFrame> &by_pin) {
    std::vector<AudioFrame> frames;
    for (const std::string &p : inst->pins) {
        auto it = by_pin.find(p);
        if (it != by_pin.end()) frames.push_back(std::move(it->second));
    }
    by_pin.clear();
    return frames;
}

sk_result process_packets(sk_plugin_handle h, const sk_pin_packet_v3 *items, size_t n, sk_output_callback_v3 cb, void *ud, sk_telemetry_callback, void *) {
    if (!h) return fail("Null handle");
    if (!items && n) return fail("Null batch");
    auto *inst = static_cast<Instance *>(h);
    try {
        std::map<std::string, AudioFrame> by_pin;
        for (size_t i = 0; i < n; ++i) {
            if (!items[i].input_pin || !items[i].packet) return fail("Null batch item");
            if (items[i].packet->packet_type != SK_PACKET_RAW_AUDIO) continue;               // non-audio packets are ignored (mixer.rs:899-901)
            const std::string pin = items[i].input_pin;
            if (std::find(inst->pins.begin(), inst->pins.end(), pin) == inst->pins.end()) return fail("Unknown input pin '" + pin + "'");
            AudioFrame f;
            std::string why;
            if (!to_frame(items[i].packet, f, why)) return fail(why);
            by_pin[pin] = std::move(f);                                                      // the latest frame per pin wins
        }
        std::vector<AudioFrame> frames = ordered(inst, by_pin);
        return mix_and_emit(inst, frames, cb, ud);
    } catch (const StreamKitError &e) {
        return fail(e.message);
    } catch (const std::exception &e) {
        return fail(e.what());
    } catch (...) {
        return fail("unknown exception");
    }
}

sk_result process_packet(sk_plugin_handle h, const char *pin, const sk_packet_v3 *pkt, sk_output_callback_v3 cb, void *ud, sk_telemetry_callback tcb, void *tud) {
    if (!h) return fail("Null handle");
    if (!pin || !pkt) return fail("Null packet");
    auto *inst = static_cast<Instance *>(h);
    try {
        if (pkt->packet_type != SK_PACKET_RAW_AUDIO) return ok();
        if (std::find(inst->pins.begin(), inst->pins.end(), std::string(pin)) == inst->pins.end()) return fail(std::string("Unknown input pin '") + pin + "'");
        AudioFrame f;
        std::string why;
        if (!to_frame(pkt, f, why)) return fail(why);
        inst->pending[pin] = std::move(f);
        if (inst->pending.size() < inst->pins.size()) return ok();       // ready_to_mix: every pin holds a frame (mixer.rs:746)
        std::vector<AudioFrame> frames = ordered(inst, inst->pending);
        (void)tcb; (void)tud;
        return mix_and_emit(inst, frames, cb, ud);
    } catch (const StreamKitError &e) {
        return fail(e.message);
    } catch (const std::exception &e) {
        return fail(e.what());
    } catch (...) {
        return fail("unknown exception");
    }
}

sk_result update_params(sk_plugin_handle h, const char *params_json) {
    if (!h) return fail("Null handle");
    auto *inst = static_cast<Instance *>(h);
    try {
        auto bad = inst->out_s16 ? inst->pcm16->update_params(params_json) : inst->gain->update_params(params_json);
        if (bad) return fail(*bad);       // the old gain stays (gain.rs:153-173)
    } catch (...) {
        return fail("unknown exception");
    }
    return ok();
}

sk_result flush(sk_plugin_handle h, sk_output_callback_v3 cb, void *ud, sk_telemetry_callback, void *) {
    if (!h) return fail("Null handle");
    auto *inst = static_cast<Instance *>(h);
    try {
        std::vector<AudioFrame> frames = ordered(inst, inst->pending);   // frames still buffered when the inputs close are mixed (mixer.rs:872-886)
        return mix_and_emit(inst, frames, cb, ud);
    } catch (...) {
        return fail("unknown exception");
    }
}

void destroy_instance(sk_plugin_handle h) {
    try { delete static_cast<Instance *>(h); } catch (...) {}
}

sk_result input_pin_added(sk_plugin_handle h, const char *pin) {
    if (!h || !pin) return fail("Null argument");
    auto *inst = static_cast<Instance *>(h);
    if (std::strncmp(pin, "in_", 3) != 0) return fail("input pins of audio::mixer are named in_<n>");
    if (std::find(inst->pins.begin(), inst->pins.end(), std::string(pin)) == inst->pins.end()) inst->pins.push_back(pin);
    return ok();
}
sk_result input_pin_removed(sk_plugin_handle h, const char *pin) {
    if (!h || !pin) return fail("Null argument");
    auto *inst = static_cast<Instance *>(h);
    inst->pins.erase(std::remove(inst->pins.begin(), inst->pins.end(), std::string(pin)), inst->pins.end());
    inst->pending.erase(pin);
    return ok();
}

const sk_native_plugin_api_v3 kApi = {SK_NATIVE_PLUGIN_API_VERSION_3, get_metadata, create_instance, process_packet, process_packets, update_params,
                                      flush, destroy_instance, input_pin_added, input_pin_removed};

}  // namespace

extern "C" __attribute__((visibility("default"))) const sk_native_plugin_api_v3 *streamkit_native_plugin_api(void) { return &kApi; }
