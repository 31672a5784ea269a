/*
 * streamkit_native_abi_v3.h -- PROPOSAL: version 3 of StreamKit's native plugin ABI.
 *
 * Version 2 (include/streamkit_native_abi.h; sdks/plugin-sdk/native/src/types.rs:13-264) cannot carry the hot path as a
 * drop-in .so:
 *   - one input pin named "in" only (crates/plugin-native/src/wrapper.rs:224,410; inputs always PinCardinality::One,
 *     crates/plugin-native/src/lib.rs:162)                                              -> no audio::mixer
 *   - audio payloads are interleaved f32 only (types.rs:137-143); an s16 result has to travel as PACKET_TYPE_BINARY and loses
 *     its content type (conversions.rs:389-395)                                         -> no f32 <-> s16 node with typed pins
 *   - packet metadata is dropped on the way out (conversions.rs:342-346)                 -> timestamps / sequence numbers end
 *     at the first plugin of a pipeline
 *   - one packet per call (wrapper.rs:398-457)                                           -> one FFI hop (and, for a GPU node,
 *     one launch) per 20 ms frame per session
 * Version 3 keeps every v2 type that does not need to change (sk_result, sk_audio_format, sk_packet_type, sk_packet_metadata,
 * the log / telemetry callbacks) and adds, marked NEW below: pin cardinality + dynamic pins, a typed audio frame with sample
 * format and layout, per-packet metadata in both directions, and a batched entry point. A v3 host can still load v2 plugins
 * (the version field tells them apart); a v3 plugin exports the same symbol.
 *
 * This header is implemented by libskgpu_plugin_mixer_v3.so (streamkit_b200/csrc/host/plugin_v3.cpp: the GPU audio::mixer with
 * per-input gain, master gain and an optional s16 output) and exercised by a C host (tests/host/host_v3.c).
 */
#ifndef STREAMKIT_NATIVE_ABI_V3_H
#define STREAMKIT_NATIVE_ABI_V3_H

#include "streamkit_native_abi.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SK_NATIVE_PLUGIN_API_VERSION_3 3u

/* NEW: crates/core/src/pins.rs PinCardinality for INPUT pins (v2 hard-codes One) */
typedef enum sk_pin_cardinality {
    SK_PIN_ONE = 0,      /* exactly one upstream connection */
    SK_PIN_DYNAMIC = 1   /* a family of pins created on demand: name = prefix + "_" + index (mixer.rs:122-173: "in_0", "in_1", ...) */
} sk_pin_cardinality;

typedef enum sk_sample_layout { SK_LAYOUT_INTERLEAVED = 0, SK_LAYOUT_PLANAR = 1 } sk_sample_layout; /* NEW */

typedef struct sk_input_pin_v3 { /* NEW: v2's sk_input_pin + cardinality */
    const char *name;              /* One: the pin name; Dynamic: the prefix ("in") */
    const sk_packet_type_info *accepts_types;
    size_t accepts_types_count;
    sk_pin_cardinality cardinality;
} sk_input_pin_v3;

typedef struct sk_node_metadata_v3 { /* v2's sk_node_metadata with v3 input pins */
    const char *kind;
    const char *description;
    const sk_input_pin_v3 *inputs;
    size_t inputs_count;
    const sk_output_pin *outputs;
    size_t outputs_count;
    const char *param_schema;
    const char *const *categories;
    size_t categories_count;
} sk_node_metadata_v3;

/* NEW: audio payload with format + layout. samples: interleaved (or planar: channel after channel) f32 or s16le; sample_count
 * counts samples of all channels. Borrowed for the duration of the call, like v2 (conversions.rs:219). */
typedef struct sk_audio_frame_v3 {
    uint32_t sample_rate;
    uint16_t channels;
    sk_sample_format sample_format;
    sk_sample_layout layout;
    const void *samples;
    size_t sample_count;
} sk_audio_frame_v3;

/* NEW: every packet may carry its metadata (crates/core/src/types.rs:40-53); NULL = none */
typedef struct sk_packet_v3 {
    sk_packet_type packet_type;
    const void *data;                    /* RawAudio: -> sk_audio_frame_v3, len = sizeof(sk_audio_frame_v3) */
    size_t len;
    const sk_packet_metadata *metadata;
} sk_packet_v3;

typedef struct sk_pin_packet_v3 { /* NEW: one element of a batch */
    const char *input_pin;
    const sk_packet_v3 *packet;
} sk_pin_packet_v3;

/* (pin_name, packet, user_data): the host copies the packet AND its metadata before returning */
typedef sk_result (*sk_output_callback_v3)(const char *, const sk_packet_v3 *, void *);

typedef struct sk_native_plugin_api_v3 {
    uint32_t version; /* 3 */
    const sk_node_metadata_v3 *(*get_metadata)(void);
    sk_plugin_handle (*create_instance)(const char *params_json, sk_log_callback log_cb, void *log_user_data);
    /* one packet on one pin, as in v2 but typed / with metadata */
    sk_result (*process_packet)(sk_plugin_handle handle, const char *input_pin, const sk_packet_v3 *packet, sk_output_callback_v3 output_cb,
                                void *output_user_data, sk_telemetry_callback telemetry_cb, void *telemetry_user_data);
    /* NEW: everything the host has for this instance right now, in one call (the packets of all of a mixer's pins for a tick; a
     * burst of frames on one pin). The plugin may answer with any number of output packets. Same threading contract as
     * process_packet: strictly sequential per instance (wrapper.rs:241-486). */
    sk_result (*process_packets)(sk_plugin_handle handle, const sk_pin_packet_v3 *items, size_t n_items, sk_output_callback_v3 output_cb,
                                 void *output_user_data, sk_telemetry_callback telemetry_cb, void *telemetry_user_data);
    sk_result (*update_params)(sk_plugin_handle handle, const char *params_json);
    sk_result (*flush)(sk_plugin_handle handle, sk_output_callback_v3 output_cb, void *output_user_data, sk_telemetry_callback telemetry_cb,
                       void *telemetry_user_data);
    void (*destroy_instance)(sk_plugin_handle handle);
    /* NEW: dynamic pin management (PinManagementMessage::AddedInputPin / RemoveInputPin, mixer.rs:341-389) */
    sk_result (*input_pin_added)(sk_plugin_handle handle, const char *pin_name);
    sk_result (*input_pin_removed)(sk_plugin_handle handle, const char *pin_name);
} sk_native_plugin_api_v3;

typedef const sk_native_plugin_api_v3 *(*sk_plugin_entry_v3_fn)(void);

#ifdef __cplusplus
}
#endif
#endif
