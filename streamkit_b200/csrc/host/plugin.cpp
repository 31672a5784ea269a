// plugin.cpp -- StreamKit native plugin ABI v2 (include/streamkit_native_abi.h, mirroring
// sdks/plugin-sdk/native/src/types.rs:205-264) over the GPU node mirror. Compiled three times:
//   -DSK_PLUGIN_KIND=1  libskgpu_plugin_gain.so       kind "gpu_gain"       (drop-in for examples/plugins/gain-native-c)
//   -DSK_PLUGIN_KIND=2  libskgpu_plugin_resampler.so  kind "gpu_resampler"  (audio::resampler semantics incl. flush)
//   -DSK_PLUGIN_KIND=3  libskgpu_plugin_pcm16.so      kind "gpu_pcm16"      (f32 -> s16le, emitted as a Binary packet)
// The host loads them as plugin::native::gpu_* (crates/plugin-native/src/lib.rs:307-333) and calls them one packet
// at a time on pin "in" (wrapper.rs:398-457). Output packets are handed to the callback and may be freed as soon as
// it returns: the host copies inside the callback (conversions.rs:340-346).
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../../include/streamkit_native_abi.h"
#include "nodes.hpp"

using namespace skhost;

#ifndef SK_PLUGIN_KIND
#error "define SK_PLUGIN_KIND (1 gain, 2 resampler, 3 pcm16)"
#endif

namespace {

thread_local std::string g_err;   // borrowed error text, valid until the next error on this thread (types.rs:42-48)
sk_result ok() { return sk_result{true, nullptr}; }
sk_result fail(const std::string &m) {
    g_err = m;
    return sk_result{false, g_err.c_str()};
}

struct Instance {
    sk_log_callback log_cb = nullptr;
    void *log_ud = nullptr;
#if SK_PLUGIN_KIND == 1
    std::unique_ptr<AudioGainNode> node;
#elif SK_PLUGIN_KIND == 2
    std::unique_ptr<AudioResamplerNode> node;
#else
    std::unique_ptr<AudioPcm16Node> node;
#endif
    void log(sk_log_level lvl, const std::string &m) const {
        if (log_cb) log_cb(lvl, "streamkit_b200", m.c_str(), log_ud);
    }
};

const sk_audio_format kAnyF32 = {0, 0, SK_SAMPLE_F32};   // wildcard rate / channels (gain.rs:93-97)
const sk_packet_type_info kInTypes[] = {{SK_PACKET_RAW_AUDIO, &kAnyF32, nullptr}};
const sk_input_pin kInputs[] = {{"in", kInTypes, 1}};
#if SK_PLUGIN_KIND == 3
const sk_output_pin kOutputs[] = {{"out", {SK_PACKET_BINARY, nullptr, nullptr}}};
#else
const sk_output_pin kOutputs[] = {{"out", {SK_PACKET_RAW_AUDIO, &kAnyF32, nullptr}}};
#endif
const char *const kCategories[] = {"audio", "filters", "gpu"};

#if SK_PLUGIN_KIND == 1
const char *kKind = "gpu_gain";
const char *kDesc = "audio::gain on the GPU (streamkit_b200): y = x * gain, gain in [0, 4], live-tunable";
const char *kSchema =
    "{\"type\":\"object\",\"properties\":{\"gain\":{\"type\":\"number\",\"default\":1.0,\"minimum\":0.0,\"maximum\":4.0,\"tunable\":true,"
    "\"description\":\"Linear gain multiplier. 0.0 = mute, 1.0 = unity (no change), 2.0 = +6dB, 4.0 = +12dB. Range: 0.0 to 4.0\"}}}";
#elif SK_PLUGIN_KIND == 2
const char *kKind = "gpu_resampler";
const char *kDesc = "audio::resampler on the GPU (streamkit_b200): rubato FastFixedIn linear interpolation, per-stream state in HBM";
const char *kSchema =
    "{\"type\":\"object\",\"required\":[\"target_sample_rate\"],\"properties\":{\"target_sample_rate\":{\"type\":\"integer\",\"minimum\":1},"
    "\"chunk_frames\":{\"type\":\"integer\",\"minimum\":1,\"default\":960},\"output_frame_size\":{\"type\":\"integer\",\"default\":960,"
    "\"description\":\"0 (variable) or a valid Opus frame size: 120, 240, 480, 960, 1920, 2880\"}}}";
#else
const char *kKind = "gpu_pcm16";
const char *kDesc = "gain -> clip -> s16le packing on the GPU (streamkit_b200); emits Binary packets of little-endian PCM16";
const char *kSchema =
    "{\"type\":\"object\",\"properties\":{\"gain\":{\"type\":\"number\",\"default\":1.0,\"minimum\":0.0,\"maximum\":4.0,\"tunable\":true}}}";
#endif

const sk_node_metadata kMetadata = {nullptr, nullptr, kInputs, 1, kOutputs, 1, nullptr, kCategories, 3};
sk_node_metadata g_md;

const sk_node_metadata *get_metadata() {
    g_md = kMetadata;
    g_md.kind = kKind;
    g_md.description = kDesc;
    g_md.param_schema = kSchema;
    return &g_md;
}

sk_plugin_handle create_instance(const char *params_json, sk_log_callback log_cb, void *log_ud) {
    try {
        // owned until every step has succeeded: a throw below (no GPU, out of memory) must not leak the instance
        std::unique_ptr<Instance> inst(new Instance());
        inst->log_cb = log_cb;
        inst->log_ud = log_ud;
        StreamKitError err{StreamKitError::Configuration, ""};
#if SK_PLUGIN_KIND == 1
        inst->node = AudioGainNode::create(params_json, &err);
#elif SK_PLUGIN_KIND == 2
        inst->node = AudioResamplerNode::create(params_json, &err);
#else
        inst->node = AudioPcm16Node::create(params_json, &err);
#endif
        if (!inst->node) {
            inst->log(SK_LOG_ERROR, err.message);
            return nullptr;   // host: StreamKitError::Configuration("Plugin failed to create instance") (wrapper.rs:184-188)
        }
        GpuRuntime::get();    // fail at creation, not at the first packet, when there is no GPU
        inst->log(SK_LOG_INFO, std::string("created ") + kKind + " instance");
        return inst.release();
    } catch (const StreamKitError &e) {
        if (log_cb) log_cb(SK_LOG_ERROR, "streamkit_b200", e.message.c_str(), log_ud);
        return nullptr;
    } catch (...) {
        return nullptr;
    }
}

[[maybe_unused]] sk_result emit_audio(const AudioFrame &f, sk_output_callback cb, void *ud) {
    sk_audio_frame af{f.sample_rate, f.channels, f.samples.data(), f.samples.size()};
    sk_packet pkt{SK_PACKET_RAW_AUDIO, &af, sizeof(sk_audio_frame)};
    return cb("out", &pkt, ud);
}

sk_result process_packet(sk_plugin_handle h, const char *pin, const sk_packet *pkt, sk_output_callback cb, void *ud,
                         sk_telemetry_callback, void *) {
    (void)pin;
    if (!h) return fail("Null handle");
    if (!pkt || !pkt->data) return fail("Null packet");
    auto *inst = static_cast<Instance *>(h);
    if (pkt->packet_type != SK_PACKET_RAW_AUDIO) {
        // the built-in nodes forward non-audio packets unchanged (gain.rs:184-196, resampler.rs:529-538)
        return cb("out", pkt, ud);
    }
    const auto *af = static_cast<const sk_audio_frame *>(pkt->data);
    if (!af->samples && af->sample_count) return fail("Invalid audio frame");
    try {
        AudioFrame in;
        in.sample_rate = af->sample_rate;
        in.channels = af->channels;
        in.samples.assign(af->samples, af->samples + af->sample_count);
        StreamKitError err{StreamKitError::Runtime, ""};
#if SK_PLUGIN_KIND == 1
        AudioFrame out;
        if (!inst->node->process(in, out, &err)) return fail(err.message);
        return emit_audio(out, cb, ud);
#elif SK_PLUGIN_KIND == 2
        std::vector<AudioFrame> outs;
        if (!inst->node->process(in, outs, &err)) return fail(err.message);
        for (const auto &f : outs) {
            sk_result r = emit_audio(f, cb, ud);
            if (!r.success) return r;
        }
        return ok();
#else
        std::vector<int16_t> s16;
        if (!inst->node->process(in, s16, &err)) return fail(err.message);
        sk_packet out{SK_PACKET_BINARY, s16.data(), s16.size() * sizeof(int16_t)};
        static const int16_t kEmpty = 0;
        if (s16.empty()) out.data = &kEmpty;
        return cb("out", &out, ud);
#endif
    } catch (const StreamKitError &e) {
        return fail(e.message);
    } catch (const std::exception &e) {
        return fail(e.what());
    } catch (...) {
        return fail("unknown exception");   // nothing may unwind across the C ABI into the Rust host
    }
}

sk_result update_params(sk_plugin_handle h, const char *params_json) {
    if (!h) return fail("Null handle");
#if SK_PLUGIN_KIND == 2
    (void)params_json;
    return ok();   // audio::resampler has no tunable parameters
#else
    auto *inst = static_cast<Instance *>(h);
    try {
        if (auto bad = inst->node->update_params(params_json)) {
            inst->log(SK_LOG_WARN, *bad);
            return fail(*bad);   // host logs a warning and keeps running (wrapper.rs:297-299); the old gain stays in effect
        }
    } catch (const std::exception &e) {
        return fail(e.what());
    } catch (...) {
        return fail("unknown exception");
    }
    return ok();
#endif
}

sk_result flush(sk_plugin_handle h, sk_output_callback cb, void *ud, sk_telemetry_callback, void *) {
    if (!h) return fail("Null handle");
#if SK_PLUGIN_KIND == 2
    auto *inst = static_cast<Instance *>(h);
    try {
        std::vector<AudioFrame> outs;
        StreamKitError err{StreamKitError::Runtime, ""};
        if (!inst->node->finish(outs, &err)) return fail(err.message);
        for (const auto &f : outs) {
            sk_result r = emit_audio(f, cb, ud);
            if (!r.success) return r;
        }
    } catch (const StreamKitError &e) {
        return fail(e.message);
    } catch (const std::exception &e) {
        return fail(e.what());
    } catch (...) {
        return fail("unknown exception");
    }
#else
    (void)cb; (void)ud;
#endif
    return ok();
}

void destroy_instance(sk_plugin_handle h) {
    try { delete static_cast<Instance *>(h); } catch (...) {}
}

const sk_native_plugin_api kApi = {SK_NATIVE_PLUGIN_API_VERSION, get_metadata, create_instance, process_packet, update_params, flush, destroy_instance};

}  // namespace

extern "C" __attribute__((visibility("default"))) const sk_native_plugin_api *streamkit_native_plugin_api(void) { return &kApi; }
