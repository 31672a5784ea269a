// nodes.cpp -- see nodes.hpp. Every arithmetic step goes through libskgpu.so (no CPU audio math here:
// the host side only parses parameters, keeps the reference's buffering / re-framing state and moves bytes).
#include "nodes.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "mini_json.hpp"

namespace skhost {

static StreamKitError rt_err(const char *what) {
    return StreamKitError{StreamKitError::Runtime, std::string(what) + ": " + skgpu_last_error()};
}

// ------------------------------------------------------------------ runtime

GpuRuntime::GpuRuntime() {
    int dev = 0;
    if (const char *e = std::getenv("SKGPU_DEVICE")) dev = std::atoi(e);
    skgpu_ctx_config cfg{};
    cfg.max_streams = 4096;
    if (const char *e = std::getenv("SKGPU_MAX_STREAMS")) cfg.max_streams = (uint32_t)std::atoi(e);
    cfg.max_channels = 8;
    cfg.fifo_frames = 0;
    if (skgpu_ctx_create(dev, &cfg, &ctx_) != SKGPU_OK) throw rt_err("skgpu_ctx_create");
}
GpuRuntime::~GpuRuntime() {
    if (ctx_) skgpu_ctx_destroy(ctx_);
}
GpuRuntime &GpuRuntime::get() {
    // deliberately leaked: the context must not be torn down from a static destructor at process exit, when the CUDA runtime
    // may already be unloading; the OS reclaims it
    static GpuRuntime *rt = new GpuRuntime();
    return *rt;
}

void PlanHolder::reset() {
    if (plan) {
        std::lock_guard<std::recursive_mutex> lk(GpuRuntime::get().mutex());
        skgpu_plan_destroy(plan);
        plan = nullptr;
    }
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

uint64_t duration_us_for_frames(uint32_t sample_rate, size_t frames_per_channel) {
    if (sample_rate == 0) return 0;
    return ((uint64_t)frames_per_channel * 1000000ull) / (uint64_t)sample_rate;
}

// ------------------------------------------------------------------ gain

std::optional<std::string> AudioGainConfig::validate() const {
    char buf[128];
    if (!std::isfinite(gain)) {
        std::snprintf(buf, sizeof(buf), "Gain must be a finite number, got: %g", (double)gain);
        return std::string(buf);
    }
    if (gain < 0.0f || gain > 4.0f) {
        std::snprintf(buf, sizeof(buf), "Gain must be between 0 and 4, got: %g", (double)gain);
        return std::string(buf);
    }
    return std::nullopt;
}

// serde: #[serde(default)] struct { gain: f32 } -- missing field -> default 1.0; wrong type -> error
static bool parse_gain_config(const char *params_json, AudioGainConfig &cfg, std::string &err) {
    cfg = AudioGainConfig{};
    if (!params_json || !*params_json) return true;
    JsonValue v;
    if (!JsonParser(params_json).parse(v, err)) return false;
    if (v.kind == JsonValue::Null) return true;
    if (v.kind != JsonValue::Object) { err = "invalid type: expected struct AudioGainConfig"; return false; }
    if (const JsonValue *g = v.get("gain")) {
        if (g->kind != JsonValue::Number) { err = "invalid type for `gain`: expected f32"; return false; }
        cfg.gain = (float)g->num;
    }
    return true;
}

static bool build_convert_plan(PlanHolder &ph, skgpu_cvt_mode mode, size_t n_samples, StreamKitError *err) {
    GpuRuntime &rt = GpuRuntime::get();
    std::lock_guard<std::recursive_mutex> lk(rt.mutex());
    if (ph.plan) { skgpu_plan_destroy(ph.plan); ph.plan = nullptr; }
    const size_t in_b = (mode == SKGPU_CVT_S16_TO_F32) ? 2 : 4, out_b = (mode == SKGPU_CVT_F32_TO_S16) ? 2 : 4;
    ph.in_bytes = n_samples * in_b;
    ph.out_off = align_up(std::max<size_t>(ph.in_bytes, 16), 256);
    ph.out_bytes = n_samples * out_b;
    const size_t arena = align_up(ph.out_off + std::max<size_t>(ph.out_bytes, 16), 256);
    if (skgpu_plan_create(rt.ctx(), arena, &ph.plan) != SKGPU_OK) { if (err) *err = rt_err("skgpu_plan_create"); return false; }
    float one = 1.0f;
    skgpu_seg seg{0, ph.out_off, (uint32_t)n_samples, 0};
    if (skgpu_plan_set_gains(ph.plan, &one, 1) != SKGPU_OK || skgpu_plan_add_convert(ph.plan, mode, &seg, 1, nullptr) != SKGPU_OK ||
        skgpu_plan_set_io(ph.plan, 0, ph.in_bytes, ph.out_off, ph.out_bytes) != SKGPU_OK || skgpu_plan_finalize(ph.plan) != SKGPU_OK) {
        if (err) *err = rt_err("building the convert plan");
        skgpu_plan_destroy(ph.plan);
        ph.plan = nullptr;
        return false;
    }
    return true;
}

static bool run_convert(PlanHolder &ph, float gain, const void *in, void *out, StreamKitError *err) {
    GpuRuntime &rt = GpuRuntime::get();
    std::lock_guard<std::recursive_mutex> lk(rt.mutex());
    // a CUDA error in a tick surfaces as a StreamKitError::Runtime (node -> Failed), never as an abort
    if (skgpu_plan_set_gains(ph.plan, &gain, 1) != SKGPU_OK || skgpu_tick_submit(ph.plan, in, out, SKGPU_SUBMIT_GRAPH) != SKGPU_OK ||
        skgpu_tick_wait(ph.plan, nullptr) != SKGPU_OK) {
        if (err) *err = rt_err("gpu tick");
        return false;
    }
    return true;
}

std::unique_ptr<AudioGainNode> AudioGainNode::create(const char *params_json, StreamKitError *err) {
    auto n = std::unique_ptr<AudioGainNode>(new AudioGainNode());
    std::string perr;
    if (!parse_gain_config(params_json, n->cfg_, perr)) n->cfg_ = AudioGainConfig{};  // parse_config_optional: defaults (helpers.rs:26-32)
    if (auto bad = n->cfg_.validate()) {
        if (err) *err = StreamKitError{StreamKitError::Configuration, *bad};           // filters/mod.rs:131-133
        return nullptr;
    }
    return n;
}

std::vector<PinSpec> AudioGainNode::input_pins() const { return {PinSpec{"in", 0, 0, true, "One"}}; }
std::vector<PinSpec> AudioGainNode::output_pins() const { return {PinSpec{"out", 0, 0, true, "Broadcast"}}; }

std::optional<std::string> AudioGainNode::update_params(const char *params_json) {
    AudioGainConfig nc;
    std::string perr;
    if (!parse_gain_config(params_json, nc, perr)) return "Failed to deserialize params for volume_adjust: " + perr;  // gain.rs:164-167
    if (auto bad = nc.validate()) return "Rejected invalid gain parameter: " + *bad;                                  // gain.rs:157-162
    cfg_ = nc;
    return std::nullopt;
}

bool AudioGainNode::process(const AudioFrame &in, AudioFrame &out, StreamKitError *err) {
    out.sample_rate = in.sample_rate;
    out.channels = in.channels;
    out.metadata = in.metadata;
    out.samples.resize(in.samples.size());
    if (in.samples.empty()) return true;
    if (!ph_.plan || plan_samples_ != in.samples.size()) {
        if (!build_convert_plan(ph_, SKGPU_CVT_F32_TO_F32, in.samples.size(), err)) return false;
        plan_samples_ = in.samples.size();
    }
    return run_convert(ph_, cfg_.gain, in.samples.data(), out.samples.data(), err);
}

std::unique_ptr<AudioPcm16Node> AudioPcm16Node::create(const char *params_json, StreamKitError *err) {
    auto n = std::unique_ptr<AudioPcm16Node>(new AudioPcm16Node());
    std::string perr;
    if (!parse_gain_config(params_json, n->cfg_, perr)) n->cfg_ = AudioGainConfig{};
    if (auto bad = n->cfg_.validate()) {
        if (err) *err = StreamKitError{StreamKitError::Configuration, *bad};
        return nullptr;
    }
    return n;
}
std::optional<std::string> AudioPcm16Node::update_params(const char *params_json) {
    AudioGainConfig nc;
    std::string perr;
    if (!parse_gain_config(params_json, nc, perr)) return "Failed to deserialize params: " + perr;
    if (auto bad = nc.validate()) return "Rejected invalid gain parameter: " + *bad;
    cfg_ = nc;
    return std::nullopt;
}
bool AudioPcm16Node::process(const AudioFrame &in, std::vector<int16_t> &out, StreamKitError *err) {
    out.resize(in.samples.size());
    if (in.samples.empty()) return true;
    if (!ph_.plan || plan_samples_ != in.samples.size()) {
        if (!build_convert_plan(ph_, SKGPU_CVT_F32_TO_S16, in.samples.size(), err)) return false;
        plan_samples_ = in.samples.size();
    }
    return run_convert(ph_, cfg_.gain, in.samples.data(), out.data(), err);
}

// ------------------------------------------------------------------ resampler

std::unique_ptr<AudioResamplerNode> AudioResamplerNode::create(const char *params_json, StreamKitError *err) {
    auto fail = [&](const std::string &m) -> std::unique_ptr<AudioResamplerNode> {
        if (err) *err = StreamKitError{StreamKitError::Configuration, m};
        return nullptr;
    };
    AudioResamplerConfig cfg;
    if (!params_json || !*params_json) {
        cfg.target_sample_rate = 48000;  // factory(None): defaults used for schema generation (resampler.rs:72-77)
    } else {
        JsonValue v;
        std::string perr;
        if (!JsonParser(params_json).parse(v, perr) || v.kind != JsonValue::Object) return fail("Failed to parse configuration: " + perr);
        const JsonValue *t = v.get("target_sample_rate");
        if (!t || t->kind != JsonValue::Number || t->num < 0) return fail("Failed to parse configuration: missing field `target_sample_rate`");
        cfg.target_sample_rate = (uint32_t)t->num;
        if (const JsonValue *c = v.get("chunk_frames")) {
            if (c->kind != JsonValue::Number || c->num < 0) return fail("Failed to parse configuration: invalid `chunk_frames`");
            cfg.chunk_frames = (size_t)c->num;
        }
        if (const JsonValue *o = v.get("output_frame_size")) {
            if (o->kind != JsonValue::Number || o->num < 0) return fail("Failed to parse configuration: invalid `output_frame_size`");
            cfg.output_frame_size = (size_t)o->num;
        }
    }
    if (cfg.target_sample_rate == 0) return fail("target_sample_rate must be greater than 0");   // resampler.rs:82-86
    if (cfg.chunk_frames == 0) return fail("chunk_frames must be greater than 0");               // resampler.rs:88-92
    if (cfg.output_frame_size != 0) {                                                           // resampler.rs:95-102
        static const size_t valid[] = {120, 240, 480, 960, 1920, 2880};
        if (std::find(std::begin(valid), std::end(valid), cfg.output_frame_size) == std::end(valid))
            return fail("output_frame_size must be 0 (disabled) or a valid Opus frame size: [120, 240, 480, 960, 1920, 2880]");
    }
    auto n = std::unique_ptr<AudioResamplerNode>(new AudioResamplerNode());
    n->cfg_ = cfg;
    return n;
}

AudioResamplerNode::~AudioResamplerNode() {
    ph_.reset();
    if (slot_ >= 0) {
        GpuRuntime &rt = GpuRuntime::get();
        std::lock_guard<std::recursive_mutex> lk(rt.mutex());
        skgpu_stream_close(rt.ctx(), (uint32_t)slot_);
    }
}

std::vector<PinSpec> AudioResamplerNode::input_pins() const { return {PinSpec{"in", 0, 0, true, "One"}}; }
std::vector<PinSpec> AudioResamplerNode::output_pins() const { return {PinSpec{"out", cfg_.target_sample_rate, 0, true, "Broadcast"}}; }

std::optional<PacketMetadata> AudioResamplerNode::next_metadata(uint64_t duration_us) {   // resampler.rs:286-297
    PacketMetadata m;
    m.timestamp_us = output_timestamp_us_;
    m.duration_us = duration_us;
    m.sequence = output_sequence_;
    output_sequence_ += 1;
    if (output_timestamp_us_) *output_timestamp_us_ += duration_us;
    return m;
}

void AudioResamplerNode::drain_output(std::vector<AudioFrame> &out) {   // resampler.rs:425-470
    const size_t ofs = cfg_.output_frame_size * channels_;
    while (output_buffer_.size() - output_off_ >= ofs) {
        AudioFrame f;
        f.sample_rate = cfg_.target_sample_rate;
        f.channels = channels_;
        f.samples.assign(output_buffer_.begin() + (long)output_off_, output_buffer_.begin() + (long)(output_off_ + ofs));
        output_off_ += ofs;
        f.metadata = next_metadata(duration_us_for_frames(cfg_.target_sample_rate, cfg_.output_frame_size));
        out.push_back(std::move(f));
    }
    if (output_off_ == output_buffer_.size()) {
        output_buffer_.clear();
        output_off_ = 0;
    } else if (output_off_ > 0 && (output_off_ >= ofs * 8 || output_off_ * 2 >= output_buffer_.size())) {
        output_buffer_.erase(output_buffer_.begin(), output_buffer_.begin() + (long)output_off_);
        output_off_ = 0;
    }
}

// one rubato process() call on the GPU: chunk -> resampled frames (variable count)
bool AudioResamplerNode::run_chunk(const float *chunk, uint32_t slot, size_t chunk_frames, std::vector<float> &resampled, StreamKitError *err) {
    GpuRuntime &rt = GpuRuntime::get();
    std::lock_guard<std::recursive_mutex> lk(rt.mutex());
    const size_t C = channels_;
    PlanHolder local;
    PlanHolder *ph = (slot == (uint32_t)slot_) ? &ph_ : &local;  // the remainder pass uses a throw-away plan
    if (!ph->plan) {
        skgpu_stream_cfg sc{sample_rate_, cfg_.target_sample_rate, (uint32_t)chunk_frames, (uint16_t)C, 0};
        const uint32_t cap = skgpu_stream_max_out_frames(&sc);
        ph->in_bytes = chunk_frames * C * 4;
        const size_t res_off = align_up(ph->in_bytes, 256);
        ph->out_off = res_off;
        const size_t data_off = res_off + 256;
        ph->out_bytes = 256 + (size_t)cap * C * 4;
        const size_t arena = align_up(data_off + (size_t)cap * C * 4, 256);
        if (skgpu_plan_create(rt.ctx(), arena, &ph->plan) != SKGPU_OK) { if (err) *err = rt_err("skgpu_plan_create"); return false; }
        skgpu_rs_item it{0, data_off, slot, cap, 0, 0};
        if (skgpu_plan_add_resample(ph->plan, &it, 1, res_off, nullptr) != SKGPU_OK ||
            skgpu_plan_set_io(ph->plan, 0, ph->in_bytes, ph->out_off, ph->out_bytes) != SKGPU_OK || skgpu_plan_finalize(ph->plan) != SKGPU_OK) {
            if (err) *err = rt_err("building the resample plan");
            skgpu_plan_destroy(ph->plan);
            ph->plan = nullptr;
            return false;
        }
        if (ph == &ph_) out_cap_ = cap;
    }
    std::vector<uint8_t> host_out(ph->out_bytes);
    if (skgpu_tick_submit(ph->plan, chunk, host_out.data(), SKGPU_SUBMIT_GRAPH) != SKGPU_OK || skgpu_tick_wait(ph->plan, nullptr) != SKGPU_OK) {
        if (err) *err = rt_err("Resampling failed");   // resampler.rs:404-407
        return false;
    }
    skgpu_rs_result res;
    std::memcpy(&res, host_out.data(), sizeof(res));
    if (res.status != 0) {
        if (err) *err = StreamKitError{StreamKitError::Runtime, "Resampling failed: device status " + std::to_string(res.status)};
        return false;
    }
    const float *o = reinterpret_cast<const float *>(host_out.data() + 256);
    resampled.assign(o, o + (size_t)res.out_frames * C);
    if (ph == &local) {
        skgpu_plan_destroy(local.plan);
        local.plan = nullptr;
    }
    return true;
}

bool AudioResamplerNode::process(const AudioFrame &in, std::vector<AudioFrame> &out, StreamKitError *err) {
    if (!initialised_) {                                                // resampler.rs:206-249
        initialised_ = true;
        needs_resample_ = in.sample_rate != cfg_.target_sample_rate;
        sample_rate_ = in.sample_rate;
        channels_ = in.channels;
        if (in.metadata && in.metadata->timestamp_us) output_timestamp_us_ = in.metadata->timestamp_us;
        if (needs_resample_) {
            GpuRuntime &rt = GpuRuntime::get();
            std::lock_guard<std::recursive_mutex> lk(rt.mutex());
            skgpu_stream_cfg sc{sample_rate_, cfg_.target_sample_rate, (uint32_t)cfg_.chunk_frames, channels_, 0};
            uint32_t slot = 0;
            if (skgpu_stream_open(rt.ctx(), &sc, &slot) != SKGPU_OK) {
                if (err) *err = StreamKitError{StreamKitError::Runtime, std::string("Failed to create resampler: ") + skgpu_last_error()};
                return false;
            }
            slot_ = slot;
        }
    }
    if (in.sample_rate != sample_rate_ || in.channels != channels_) {   // resampler.rs:253-279: fatal
        char buf[160];
        std::snprintf(buf, sizeof(buf), "Audio format changed mid-stream: expected %uHz/%uch, got %uHz/%uch", sample_rate_, (unsigned)channels_,
                      in.sample_rate, (unsigned)in.channels);
        if (err) *err = StreamKitError{StreamKitError::Runtime, buf};
        return false;
    }
    if (!needs_resample_) {                                             // resampler.rs:299-373
        if (cfg_.output_frame_size == 0) {
            out.push_back(in);
            return true;
        }
        output_buffer_.insert(output_buffer_.end(), in.samples.begin(), in.samples.end());
        drain_output(out);
        return true;
    }
    sample_buffer_.insert(sample_buffer_.end(), in.samples.begin(), in.samples.end());   // resampler.rs:377
    const size_t chunk_samples = cfg_.chunk_frames * channels_;
    std::vector<float> resampled;
    while (sample_buffer_.size() - sample_off_ >= chunk_samples) {
        if (!run_chunk(sample_buffer_.data() + sample_off_, (uint32_t)slot_, cfg_.chunk_frames, resampled, err)) return false;
        if (cfg_.output_frame_size > 0) {
            output_buffer_.insert(output_buffer_.end(), resampled.begin(), resampled.end());
            drain_output(out);
        } else {                                                        // resampler.rs:471-512
            AudioFrame f;
            f.sample_rate = cfg_.target_sample_rate;
            f.channels = channels_;
            f.samples = resampled;
            f.metadata = next_metadata(duration_us_for_frames(cfg_.target_sample_rate, resampled.size() / channels_));
            out.push_back(std::move(f));
        }
        sample_off_ += chunk_samples;
    }
    if (sample_off_ == sample_buffer_.size()) {                          // resampler.rs:518-526
        sample_buffer_.clear();
        sample_off_ = 0;
    } else if (sample_off_ > 0 && (sample_off_ >= chunk_samples * 4 || sample_off_ * 2 >= sample_buffer_.size())) {
        sample_buffer_.erase(sample_buffer_.begin(), sample_buffer_.begin() + (long)sample_off_);
        sample_off_ = 0;
    }
    return true;
}

bool AudioResamplerNode::finish(std::vector<AudioFrame> &out, StreamKitError *err) {
    if (sample_buffer_.size() > sample_off_ && initialised_ && needs_resample_) {   // resampler.rs:543-689
        const size_t remaining_frames = (sample_buffer_.size() - sample_off_) / channels_;
        if (remaining_frames > 0) {
            // "create a new resampler with the exact size": a FRESH FastFixedIn(chunk = remaining_frames)
            GpuRuntime &rt = GpuRuntime::get();
            uint32_t tmp_slot = 0;
            {
                std::lock_guard<std::recursive_mutex> lk(rt.mutex());
                skgpu_stream_cfg sc{sample_rate_, cfg_.target_sample_rate, (uint32_t)remaining_frames, channels_, 0};
                if (skgpu_stream_open(rt.ctx(), &sc, &tmp_slot) != SKGPU_OK) {
                    if (err) *err = StreamKitError{StreamKitError::Runtime, std::string("Failed to create remainder resampler: ") + skgpu_last_error()};
                    return false;
                }
            }
            std::vector<float> resampled;
            const bool ok = run_chunk(sample_buffer_.data() + sample_off_, tmp_slot, remaining_frames, resampled, err);
            {
                std::lock_guard<std::recursive_mutex> lk(rt.mutex());
                skgpu_stream_close(rt.ctx(), tmp_slot);
            }
            if (!ok) return false;
            if (cfg_.output_frame_size > 0) {
                output_buffer_.insert(output_buffer_.end(), resampled.begin(), resampled.end());
                drain_output(out);
            } else {
                AudioFrame f;
                f.sample_rate = cfg_.target_sample_rate;
                f.channels = channels_;
                f.samples = resampled;
                f.metadata = next_metadata(duration_us_for_frames(cfg_.target_sample_rate, resampled.size() / channels_));
                out.push_back(std::move(f));
            }
        }
        sample_buffer_.clear();
        sample_off_ = 0;
    }
    if (output_buffer_.size() > output_off_ && cfg_.output_frame_size > 0) {       // resampler.rs:691-730
        AudioFrame f;
        f.sample_rate = cfg_.target_sample_rate;
        f.channels = channels_;
        f.samples.assign(output_buffer_.begin() + (long)output_off_, output_buffer_.end());
        PacketMetadata m;
        m.timestamp_us = output_timestamp_us_;
        m.duration_us = duration_us_for_frames(cfg_.target_sample_rate, f.samples.size() / channels_);
        m.sequence = output_sequence_;                                               // not incremented afterwards (:707-711)
        if (output_timestamp_us_) *output_timestamp_us_ += *m.duration_us;
        f.metadata = m;
        out.push_back(std::move(f));
        output_buffer_.clear();
        output_off_ = 0;
    }
    return true;
}

// ------------------------------------------------------------------ mixer

std::unique_ptr<AudioMixerNode> AudioMixerNode::create(const char *params_json, StreamKitError *err) {
    auto n = std::unique_ptr<AudioMixerNode>(new AudioMixerNode());
    if (params_json && *params_json) {
        JsonValue v;
        std::string perr;
        if (JsonParser(params_json).parse(v, perr) && v.kind == JsonValue::Object) {
            if (const JsonValue *t = v.get("sync_timeout_ms")) {
                if (t->kind == JsonValue::Null) n->cfg_.sync_timeout_ms = std::nullopt;
                else if (t->kind == JsonValue::Number) n->cfg_.sync_timeout_ms = (uint64_t)t->num;
            }
            if (const JsonValue *t = v.get("num_inputs"))
                if (t->kind == JsonValue::Number) n->cfg_.num_inputs = (size_t)t->num;
            if (const JsonValue *c = v.get("clocked")) {
                if (c->kind == JsonValue::Object) {
                    ClockedMixerConfig cc;
                    if (const JsonValue *x = c->get("sample_rate")) if (x->kind == JsonValue::Number) cc.sample_rate = (uint32_t)x->num;
                    if (const JsonValue *x = c->get("frame_samples_per_channel")) if (x->kind == JsonValue::Number) cc.frame_samples_per_channel = (size_t)x->num;
                    if (const JsonValue *x = c->get("jitter_buffer_frames")) if (x->kind == JsonValue::Number) cc.jitter_buffer_frames = (size_t)x->num;
                    if (const JsonValue *x = c->get("generate_silence")) if (x->kind == JsonValue::Bool) cc.generate_silence = x->b;
                    n->cfg_.clocked = cc;
                }
            }
        }  // parse failure -> defaults (parse_config_optional)
    }
    (void)err;
    return n;
}

std::vector<PinSpec> AudioMixerNode::input_pins() const {
    std::vector<PinSpec> pins;
    if (cfg_.num_inputs)
        for (size_t i = 0; i < *cfg_.num_inputs; ++i) pins.push_back(PinSpec{"in_" + std::to_string(i), 0, 0, true, "One"});
    return pins;
}

bool AudioMixerNode::run(const std::vector<AudioFrame> &frames, uint16_t oc, size_t out_frames, uint32_t rate, AudioFrame &out, StreamKitError *err) {
    GpuRuntime &rt = GpuRuntime::get();
    std::lock_guard<std::recursive_mutex> lk(rt.mutex());
    std::vector<skgpu_mix_input> ins(frames.size());
    size_t off = 0;
    for (size_t i = 0; i < frames.size(); ++i) {
        const AudioFrame &f = frames[i];
        if (f.channels == 0) { if (err) *err = StreamKitError{StreamKitError::Runtime, "mixer input with 0 channels"}; return false; }
        ins[i] = skgpu_mix_input{off, (uint32_t)(f.samples.size() / f.channels), f.channels, (uint16_t)(f.unique ? SKGPU_MIX_IN_UNIQUE : 0), SKGPU_NO_GAIN, 0};
        off = align_up(off + f.samples.size() * 4, 16);
    }
    const size_t in_bytes = align_up(std::max<size_t>(off, 16), 256);
    const size_t out_bytes = out_frames * oc * 4;
    // shape of this mix: reuse the compiled plan while it stays the same (the steady state of a pipeline)
    std::vector<uint64_t> shape{(uint64_t)oc, (uint64_t)out_frames};
    for (const AudioFrame &f : frames) shape.push_back(((uint64_t)f.samples.size() << 32) | ((uint64_t)f.channels << 8) | (f.unique ? 1u : 0u));
    if (!ph_.plan || shape != shape_) {
        ph_.reset();
        skgpu_mix_group g{in_bytes, 0, (uint32_t)frames.size(), (uint32_t)out_frames, oc, 0, SKGPU_NO_GAIN, 0};
        if (skgpu_plan_create(rt.ctx(), align_up(in_bytes + std::max<size_t>(out_bytes, 16), 256), &ph_.plan) != SKGPU_OK) { if (err) *err = rt_err("skgpu_plan_create"); return false; }
        const bool built = skgpu_plan_add_mix(ph_.plan, &g, 1, ins.data(), (uint32_t)ins.size(), nullptr) == SKGPU_OK &&
                           skgpu_plan_set_io(ph_.plan, 0, in_bytes, in_bytes, out_bytes) == SKGPU_OK && skgpu_plan_finalize(ph_.plan) == SKGPU_OK;
        if (!built) { if (err) *err = rt_err("gpu mix plan"); ph_.reset(); return false; }
        shape_ = shape;
        host_in_.assign(in_bytes, 0);
    }
    for (size_t i = 0; i < frames.size(); ++i) std::memcpy(host_in_.data() + ins[i].in_off, frames[i].samples.data(), frames[i].samples.size() * 4);
    out.sample_rate = rate;
    out.channels = oc;
    out.samples.assign(out_frames * oc, 0.0f);
    const bool ok = skgpu_tick_submit(ph_.plan, host_in_.data(), out.samples.data(), SKGPU_SUBMIT_GRAPH) == SKGPU_OK && skgpu_tick_wait(ph_.plan, nullptr) == SKGPU_OK;
    if (!ok && err) *err = rt_err("gpu mix");
    return ok;
}

bool AudioMixerNode::mix(const std::vector<AudioFrame> &frames, AudioFrame &out, StreamKitError *err) {
    if (frames.empty()) { out = AudioFrame{}; return true; }                         // mixer.rs:940-942
    uint16_t current_max = 0;
    size_t max_spc = 0;
    for (const auto &f : frames) {
        current_max = std::max(current_max, f.channels);
        if (f.channels) max_spc = std::max(max_spc, f.samples.size() / f.channels);
    }
    max_channels_seen_ = std::max(max_channels_seen_, current_max);                   // sticky (mixer.rs:722)
    const uint16_t oc = std::max<uint16_t>(std::max(max_channels_seen_, current_max), 1);   // mixer.rs:947-948
    if (!run(frames, oc, max_spc, frames.front().sample_rate, out, err)) return false;
    out.metadata = frames.front().metadata;                                           // mixer.rs:994
    return true;
}

bool AudioMixerNode::mix_clocked(const std::vector<AudioFrame> &frames, AudioFrame &out, StreamKitError *err) {
    const ClockedMixerConfig cc = cfg_.clocked.value_or(ClockedMixerConfig{});
    for (const auto &f : frames) {
        if (f.sample_rate != cc.sample_rate) {                                        // mixer.rs:1326-1341: fatal
            if (err) *err = StreamKitError{StreamKitError::Runtime, "Clocked mixer input sample_rate mismatch: got " + std::to_string(f.sample_rate) +
                                                                          ", expected " + std::to_string(cc.sample_rate)};
            return false;
        }
        max_channels_seen_ = std::max(max_channels_seen_, f.channels);                // mixer.rs:1343
    }
    if (max_channels_seen_ == 0 && frames.empty()) { out = AudioFrame{}; return true; }   // mixer.rs:1404-1407
    const uint16_t oc = std::max<uint16_t>(max_channels_seen_, 1);
    if (!run(frames, oc, cc.frame_samples_per_channel, cc.sample_rate, out, err)) return false;
    PacketMetadata m;
    if (!frames.empty() && frames.front().metadata) m = *frames.front().metadata;
    else m.duration_us = duration_us_for_frames(cc.sample_rate, cc.frame_samples_per_channel);   // mixer.rs:1413-1418
    out.metadata = m;
    return true;
}

}  // namespace skhost
