// nodes.hpp -- C++ host-side mirror of the reference's audio filter nodes, running their arithmetic on the GPU
// through the batch C ABI (libskgpu.so). Same names, parameters, validation rules and error behaviour as
// crates/nodes/src/audio/filters/{gain,resampler,mixer}.rs so a maintainer can line the two up:
//
//   AudioGainNode       gain.rs:30-67 (config + validate), :153-173 (UpdateParams), :184-190 (process)
//   AudioResamplerNode  resampler.rs:22-38,:81-102 (config), :206-279 (format latch), :377-470 (chunk + re-frame),
//                       :286-297 (metadata), :543-730 (EOF remainder + flush)
//   AudioMixerNode      mixer.rs:60-79 (config), :944-1013 (mix_and_send arithmetic), :1436-1492 (clocked)
//   AudioPcm16Node      build-defined f32 -> s16le packing (the reference only declares SampleFormat::S16Le)
//
// The reference is Rust; no Rust toolchain exists in this image, so the host side above the C ABI is C++
// (INTEGRATION.md shows the Rust binding a maintainer adds). One packet per call is the drop-in shape, not the
// fast one: the batch ABI (skgpu_batch.h) is what reaches the throughput numbers.
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <optional>
#include <string>
#include <vector>

#include "../../../include/skgpu_batch.h"

namespace skhost {

struct StreamKitError {                 // crates/core/src/error.rs: Configuration | Runtime
    enum Kind { Configuration, Runtime } kind;
    std::string message;
};

struct PacketMetadata {                 // crates/core/src/types.rs:40-53
    std::optional<uint64_t> timestamp_us, duration_us, sequence;
};

struct AudioFrame {                     // crates/core/src/types.rs:207-216 (interleaved f32)
    uint32_t sample_rate = 0;
    uint16_t channels = 0;
    std::vector<float> samples;
    std::optional<PacketMetadata> metadata;
    bool unique = true;                 // AudioFrame::has_unique_samples()
};

struct PinSpec {                        // crates/core/src/pins.rs:32-106, reduced to what the filters declare
    std::string name;
    uint32_t sample_rate;               // 0 = wildcard
    uint16_t channels;                  // 0 = wildcard
    bool f32;
    const char *cardinality;            // "One" | "Broadcast" | "Dynamic"
};

// ---- process-wide GPU runtime shared by all node instances (instances run concurrently on spawn_blocking
// threads, wrapper.rs:398-457; a context is thread-compatible, so calls are serialised by a mutex)
class GpuRuntime {
  public:
    static GpuRuntime &get();           // throws StreamKitError{Runtime} if no CUDA device: there is no CPU fallback
    skgpu_ctx *ctx() { return ctx_; }
    std::recursive_mutex &mutex() { return mu_; }
    ~GpuRuntime();

  private:
    GpuRuntime();
    skgpu_ctx *ctx_ = nullptr;
    std::recursive_mutex mu_;
};

// one-op plan cache: a node instance keeps the compiled tick for its current packet shape
struct PlanHolder {
    skgpu_plan *plan = nullptr;
    size_t in_bytes = 0, out_bytes = 0, out_off = 0;
    void reset();
    ~PlanHolder() { reset(); }
};

// ------------------------------------------------------------------ audio::gain
struct AudioGainConfig {
    float gain = 1.0f;                                          // gain.rs:38-42
    std::optional<std::string> validate() const;                // gain.rs:50-66 (returns the reference's message)
};

class AudioGainNode {
  public:
    static constexpr const char *kKind = "audio::gain";
    // factory (filters/mod.rs:127-143): parse_config_optional -> defaults on parse failure, then validate
    static std::unique_ptr<AudioGainNode> create(const char *params_json, StreamKitError *err);
    std::vector<PinSpec> input_pins() const;                    // gain.rs:89-101
    std::vector<PinSpec> output_pins() const;                   // gain.rs:103-114
    // NodeControlMessage::UpdateParams (gain.rs:153-173): invalid -> old gain kept, error text returned
    std::optional<std::string> update_params(const char *params_json);
    bool process(const AudioFrame &in, AudioFrame &out, StreamKitError *err);   // gain.rs:184-190
    float gain() const { return cfg_.gain; }

  private:
    AudioGainConfig cfg_;
    PlanHolder ph_;
    size_t plan_samples_ = 0;
};

// ------------------------------------------------------------------ f32 -> s16le (+ optional gain)
class AudioPcm16Node {
  public:
    static constexpr const char *kKind = "audio::pcm16";
    static std::unique_ptr<AudioPcm16Node> create(const char *params_json, StreamKitError *err);
    std::optional<std::string> update_params(const char *params_json);
    bool process(const AudioFrame &in, std::vector<int16_t> &out, StreamKitError *err);

  private:
    AudioGainConfig cfg_;
    PlanHolder ph_;
    size_t plan_samples_ = 0;
};

// ------------------------------------------------------------------ audio::resampler
struct AudioResamplerConfig {
    uint32_t target_sample_rate = 0;    // required (resampler.rs:22-27)
    size_t chunk_frames = 960;          // resampler.rs:40-42
    size_t output_frame_size = 960;     // resampler.rs:44-46
};

class AudioResamplerNode {
  public:
    static constexpr const char *kKind = "audio::resampler";
    static std::unique_ptr<AudioResamplerNode> create(const char *params_json, StreamKitError *err);   // resampler.rs:68-105
    ~AudioResamplerNode();
    std::vector<PinSpec> input_pins() const;
    std::vector<PinSpec> output_pins() const;                   // RawAudio{target_rate, 0, F32} (resampler.rs:134-145)
    // one input packet (resampler.rs:203-528); emitted packets are appended to `out`
    bool process(const AudioFrame &in, std::vector<AudioFrame> &out, StreamKitError *err);
    // input closed (resampler.rs:543-730)
    bool finish(std::vector<AudioFrame> &out, StreamKitError *err);

  private:
    bool run_chunk(const float *chunk, uint32_t slot, size_t chunk_frames, std::vector<float> &resampled, StreamKitError *err);
    void drain_output(std::vector<AudioFrame> &out);
    std::optional<PacketMetadata> next_metadata(uint64_t duration_us);
    AudioResamplerConfig cfg_;
    bool initialised_ = false, needs_resample_ = false;
    uint32_t sample_rate_ = 0;
    uint16_t channels_ = 0;
    uint64_t output_sequence_ = 0;
    std::optional<uint64_t> output_timestamp_us_;
    std::vector<float> sample_buffer_, output_buffer_;
    size_t sample_off_ = 0, output_off_ = 0;
    int64_t slot_ = -1;
    PlanHolder ph_;
    uint32_t out_cap_ = 0;
};

// ------------------------------------------------------------------ audio::mixer (arithmetic only; the sync / jitter
// state machines of mixer.rs:554-918 and :1242-1434 decide WHICH frames are mixed and stay with the caller)
struct ClockedMixerConfig {             // mixer.rs:23-55
    uint32_t sample_rate = 48000;
    size_t frame_samples_per_channel = 960;
    size_t jitter_buffer_frames = 3;
    bool generate_silence = true;
};
struct AudioMixerConfig {               // mixer.rs:60-88
    std::optional<uint64_t> sync_timeout_ms = 100;
    std::optional<size_t> num_inputs;
    std::optional<ClockedMixerConfig> clocked;
};

class AudioMixerNode {
  public:
    static constexpr const char *kKind = "audio::mixer";
    static std::unique_ptr<AudioMixerNode> create(const char *params_json, StreamKitError *err);
    std::vector<PinSpec> input_pins() const;                    // in_0..in_{n-1} when num_inputs is set (mixer.rs:128-143)
    const AudioMixerConfig &config() const { return cfg_; }
    // mix_and_send arithmetic (mixer.rs:944-1013): frames in pin order; sticky output channels are tracked here
    bool mix(const std::vector<AudioFrame> &frames, AudioFrame &out, StreamKitError *err);
    // mix_clocked_frames (mixer.rs:1436-1492)
    bool mix_clocked(const std::vector<AudioFrame> &frames, AudioFrame &out, StreamKitError *err);

  private:
    bool run(const std::vector<AudioFrame> &frames, uint16_t oc, size_t out_frames, uint32_t rate, AudioFrame &out, StreamKitError *err);
    AudioMixerConfig cfg_;
    uint16_t max_channels_seen_ = 0;
    // the compiled tick of the current packet shape: a steady stream of equally shaped mixes is one submit each, not a
    // plan build + graph capture + teardown per mix
    PlanHolder ph_;
    std::vector<uint64_t> shape_;
    std::vector<uint8_t> host_in_;
};

uint64_t duration_us_for_frames(uint32_t sample_rate, size_t frames_per_channel);   // resampler.rs:108-116

}  // namespace skhost
