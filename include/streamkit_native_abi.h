/*
 * streamkit_native_abi.h -- C view of StreamKit's native plugin ABI, version 2, as the Rust host
 * actually calls it. Written from sdks/plugin-sdk/native/src/types.rs (the authoritative #[repr(C)]
 * definitions); each declaration cites the lines it mirrors.
 *
 * Why not reuse examples/plugins/gain-native-c/streamkit_plugin.h: that header is stale. It declares
 * process_packet with 5 parameters and flush with 3 (its lines 264-266, 285-286) while the host
 * passes 7 and 5 (types.rs:229-237, :250-256) -- a telemetry callback and its user data were added
 * in v2. The old header only works because the SysV x86-64 convention ignores surplus register
 * arguments. This header declares the real signatures.
 *
 * Names are prefixed sk_ to keep them apart from the reference's C names; layouts are identical.
 */
#ifndef STREAMKIT_NATIVE_ABI_H
#define STREAMKIT_NATIVE_ABI_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SK_NATIVE_PLUGIN_API_VERSION 2u /* types.rs:13; host rejects anything else (plugin-native/src/lib.rs:90-95) */
#define SK_PLUGIN_API_SYMBOL "streamkit_native_plugin_api" /* types.rs:264 */

typedef void *sk_plugin_handle; /* types.rs:16 */

typedef enum sk_log_level { SK_LOG_TRACE = 0, SK_LOG_DEBUG = 1, SK_LOG_INFO = 2, SK_LOG_WARN = 3, SK_LOG_ERROR = 4 } sk_log_level; /* :19-27 */

/* (level, target, message, user_data)  types.rs:35 */
typedef void (*sk_log_callback)(sk_log_level, const char *, const char *, void *);

/* types.rs:38-49. error_message is BORROWED: the callee keeps it alive until its next error on the same
 * thread; the caller copies it and never frees it. */
typedef struct sk_result {
    bool success;
    const char *error_message;
} sk_result;

typedef enum sk_sample_format { SK_SAMPLE_F32 = 0, SK_SAMPLE_S16LE = 1 } sk_sample_format; /* :62-67 */

typedef struct sk_audio_format { /* :70-76 */
    uint32_t sample_rate; /* 0 = wildcard */
    uint16_t channels;    /* 0 = wildcard */
    sk_sample_format sample_format;
} sk_audio_format;

typedef enum sk_packet_type { /* :79-90 */
    SK_PACKET_RAW_AUDIO = 0,
    SK_PACKET_OPUS_AUDIO = 1,
    SK_PACKET_TEXT = 2,
    SK_PACKET_TRANSCRIPTION = 3,
    SK_PACKET_CUSTOM = 4,
    SK_PACKET_BINARY = 5,
    SK_PACKET_ANY = 6,
    SK_PACKET_PASSTHROUGH = 7
} sk_packet_type;

typedef struct sk_packet_metadata { /* :100-109 */
    uint64_t timestamp_us;
    bool has_timestamp_us;
    uint64_t duration_us;
    bool has_duration_us;
    uint64_t sequence;
    bool has_sequence;
} sk_packet_metadata;

typedef struct sk_packet_type_info { /* :126-134 */
    sk_packet_type type_discriminant;
    const sk_audio_format *audio_format; /* RawAudio only */
    const char *custom_type_id;          /* Custom only */
} sk_packet_type_info;

/* :137-143. samples is interleaved f32 (crates/core/src/types.rs:210-211), borrowed for the call. */
typedef struct sk_audio_frame {
    uint32_t sample_rate;
    uint16_t channels;
    const float *samples;
    size_t sample_count; /* all channels */
} sk_audio_frame;

/* :147-152. RawAudio: data -> sk_audio_frame, len = sizeof(sk_audio_frame) (conversions.rs:222-226);
 * Binary: data -> bytes, len = byte count (conversions.rs:296-303). */
typedef struct sk_packet {
    sk_packet_type packet_type;
    const void *data;
    size_t len;
} sk_packet;

typedef struct sk_input_pin { /* :155-161 */
    const char *name;
    const sk_packet_type_info *accepts_types;
    size_t accepts_types_count;
} sk_input_pin;

typedef struct sk_output_pin { /* :164-168 */
    const char *name;
    sk_packet_type_info produces_type;
} sk_output_pin;

typedef struct sk_node_metadata { /* :171-185; must stay valid for the library's lifetime */
    const char *kind;
    const char *description; /* may be NULL */
    const sk_input_pin *inputs;
    size_t inputs_count;
    const sk_output_pin *outputs;
    size_t outputs_count;
    const char *param_schema; /* JSON schema text */
    const char *const *categories;
    size_t categories_count;
} sk_node_metadata;

/* (pin_name, packet, user_data); the host copies the packet before returning (conversions.rs:340-346)  :189 */
typedef sk_result (*sk_output_callback)(const char *, const sk_packet *, void *);
/* (event_type, data_json, data_len, metadata, user_data); may be NULL  :199-201 */
typedef sk_result (*sk_telemetry_callback)(const char *, const uint8_t *, size_t, const sk_packet_metadata *, void *);

typedef struct sk_native_plugin_api { /* :205-261 */
    uint32_t version;
    const sk_node_metadata *(*get_metadata)(void);
    sk_plugin_handle (*create_instance)(const char *params_json, sk_log_callback log_cb, void *log_user_data);
    sk_result (*process_packet)(sk_plugin_handle handle, const char *input_pin, const sk_packet *packet,
                                sk_output_callback output_cb, void *output_user_data,
                                sk_telemetry_callback telemetry_cb, void *telemetry_user_data);
    sk_result (*update_params)(sk_plugin_handle handle, const char *params_json);
    sk_result (*flush)(sk_plugin_handle handle, sk_output_callback output_cb, void *output_user_data,
                       sk_telemetry_callback telemetry_cb, void *telemetry_user_data);
    void (*destroy_instance)(sk_plugin_handle handle);
} sk_native_plugin_api;

/* every plugin library exports exactly this symbol */
typedef const sk_native_plugin_api *(*sk_plugin_entry_fn)(void);

#ifdef __cplusplus
}
#endif
#endif
