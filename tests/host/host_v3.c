/* host_v3.c -- a C host for the PROPOSED native plugin ABI v3 (include/streamkit_native_abi_v3.h): test infrastructure.
 *
 * Loads a v3 plugin .so like crates/plugin-native/src/lib.rs loads v2 plugins (dlopen, the streamkit_native_plugin_api
 * symbol, version check), creates one instance and drives the parts v2 does not have:
 *   - dynamic input pins (input_pin_added / input_pin_removed)
 *   - process_packets: the frames of all pins in ONE call, f32 or s16 payloads, metadata attached
 *   - typed s16 output, metadata copied inside the output callback (what the Rust host does in conversions.rs:340-346)
 * Inputs come from a fixed LCG so the Python test can regenerate them and check the output bytes against the oracle.
 *
 * usage: host_v3 <plugin.so> <params_json> <n_inputs> <in_fmt f32|s16> <rounds> <out_file>
 * prints one line per emitted packet: "packet <bytes> rate <r> ch <c> fmt <f> ts <t|-> seq <s|->" and appends the payload
 * bytes to out_file. Exit code 0 on success. */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/streamkit_native_abi_v3.h"

static FILE *g_out;
static int g_packets;

static sk_result on_output(const char *pin, const sk_packet_v3 *pkt, void *ud) {
    (void)ud;
    sk_result ok = {true, NULL}, bad = {false, "host: unsupported packet"};
    if (!pin || strcmp(pin, "out") != 0 || !pkt || pkt->packet_type != SK_PACKET_RAW_AUDIO || pkt->len != sizeof(sk_audio_frame_v3)) return bad;
    const sk_audio_frame_v3 *f = (const sk_audio_frame_v3 *)pkt->data;
    const size_t bytes = f->sample_count * (f->sample_format == SK_SAMPLE_S16LE ? 2 : 4);
    fwrite(f->samples, 1, bytes, g_out);   /* the host copies payload and metadata before returning */
    char ts[32] = "-", seq[32] = "-";
    if (pkt->metadata && pkt->metadata->has_timestamp_us) snprintf(ts, sizeof ts, "%llu", (unsigned long long)pkt->metadata->timestamp_us);
    if (pkt->metadata && pkt->metadata->has_sequence) snprintf(seq, sizeof seq, "%llu", (unsigned long long)pkt->metadata->sequence);
    printf("packet %zu rate %u ch %u fmt %d ts %s seq %s\n", bytes, f->sample_rate, (unsigned)f->channels, (int)f->sample_format, ts, seq);
    g_packets++;
    return ok;
}

static void on_log(sk_log_level lvl, const char *target, const char *msg, void *ud) {
    (void)ud;
    fprintf(stderr, "[plugin %d %s] %s\n", (int)lvl, target, msg);
}

static uint32_t lcg(uint32_t *s) { *s = *s * 1664525u + 1013904223u; return *s; }

int main(int argc, char **argv) {
    if (argc != 7) { fprintf(stderr, "usage: host_v3 plugin.so params n_inputs f32|s16 rounds out_file\n"); return 2; }
    const int n_in = atoi(argv[3]), s16 = strcmp(argv[4], "s16") == 0, rounds = atoi(argv[5]);
    void *lib = dlopen(argv[1], RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 3; }
    sk_plugin_entry_v3_fn entry = (sk_plugin_entry_v3_fn)dlsym(lib, SK_PLUGIN_API_SYMBOL);
    if (!entry) { fprintf(stderr, "no %s symbol\n", SK_PLUGIN_API_SYMBOL); return 3; }
    const sk_native_plugin_api_v3 *api = entry();
    if (api->version != SK_NATIVE_PLUGIN_API_VERSION_3) { fprintf(stderr, "plugin speaks ABI version %u, this host needs 3\n", api->version); return 4; }
    const sk_node_metadata_v3 *md = api->get_metadata();
    printf("kind %s inputs %zu cardinality %d prefix %s accepts %zu\n", md->kind, md->inputs_count, (int)md->inputs[0].cardinality, md->inputs[0].name,
           md->inputs[0].accepts_types_count);
    sk_plugin_handle h = api->create_instance(argv[2], on_log, NULL);
    if (!h) { fprintf(stderr, "create_instance failed\n"); return 5; }
    g_out = fopen(argv[6], "wb");
    char pin[64][16];
    for (int i = 0; i < n_in; i++) {
        snprintf(pin[i], sizeof pin[i], "in_%d", i);
        sk_result r = api->input_pin_added(h, pin[i]);
        if (!r.success) { fprintf(stderr, "input_pin_added: %s\n", r.error_message); return 6; }
    }
    const size_t N = 960 * 2;   /* one 20 ms stereo frame at 48 kHz */
    float *f32 = (float *)malloc((size_t)n_in * N * sizeof(float));
    int16_t *i16 = (int16_t *)malloc((size_t)n_in * N * sizeof(int16_t));
    uint32_t seed = 12345u;
    for (int r = 0; r < rounds; r++) {
        sk_audio_frame_v3 fr[64];
        sk_packet_v3 pk[64];
        sk_packet_metadata mdv[64];
        sk_pin_packet_v3 items[64];
        for (int i = 0; i < n_in; i++) {
            for (size_t k = 0; k < N; k++) {
                const int32_t v = (int32_t)(lcg(&seed) >> 16) - 32768;          /* -32768 .. 32767 */
                i16[(size_t)i * N + k] = (int16_t)v;
                f32[(size_t)i * N + k] = (float)v * (1.0f / 32768.0f) * 0.75f;
            }
            fr[i].sample_rate = 48000; fr[i].channels = 2; fr[i].layout = SK_LAYOUT_INTERLEAVED; fr[i].sample_count = N;
            fr[i].sample_format = s16 ? SK_SAMPLE_S16LE : SK_SAMPLE_F32;
            fr[i].samples = s16 ? (const void *)(i16 + (size_t)i * N) : (const void *)(f32 + (size_t)i * N);
            memset(&mdv[i], 0, sizeof mdv[i]);
            mdv[i].timestamp_us = 1000000ull * (uint64_t)(i + 1) + 20000ull * (uint64_t)r; mdv[i].has_timestamp_us = true;
            mdv[i].sequence = (uint64_t)(100 * (i + 1) + r); mdv[i].has_sequence = true;
            pk[i].packet_type = SK_PACKET_RAW_AUDIO; pk[i].data = &fr[i]; pk[i].len = sizeof fr[i]; pk[i].metadata = &mdv[i];
            items[i].input_pin = pin[i]; items[i].packet = &pk[i];
        }
        if (r == rounds - 1 && n_in > 1) {
            /* last round: the last pin is removed first, its frame must not be mixed */
            sk_result rr = api->input_pin_removed(h, pin[n_in - 1]);
            if (!rr.success) { fprintf(stderr, "input_pin_removed: %s\n", rr.error_message); return 6; }
        }
        const size_t n_items = (r == rounds - 1 && n_in > 1) ? (size_t)n_in - 1 : (size_t)n_in;
        sk_result res = api->process_packets(h, items, n_items, on_output, NULL, NULL, NULL);
        if (!res.success) { fprintf(stderr, "process_packets: %s\n", res.error_message); return 7; }
    }
    /* one frame through the v2-style per-packet entry point, then flush: the buffered frame is mixed alone */
    {
        sk_audio_frame_v3 fr = {48000, 2, SK_SAMPLE_F32, SK_LAYOUT_INTERLEAVED, f32, N};
        sk_packet_v3 pk = {SK_PACKET_RAW_AUDIO, &fr, sizeof fr, NULL};
        sk_result res = api->process_packet(h, pin[0], &pk, on_output, NULL, NULL, NULL);
        if (!res.success) { fprintf(stderr, "process_packet: %s\n", res.error_message); return 7; }
        res = api->flush(h, on_output, NULL, NULL, NULL);
        if (!res.success) { fprintf(stderr, "flush: %s\n", res.error_message); return 7; }
    }
    sk_result bad = api->update_params(h, "{\"gain\": 9.0}");   /* rejected, the old gain stays (gain.rs:153-173) */
    printf("update_params(gain 9.0) success %d\n", (int)bad.success);
    api->destroy_instance(h);
    fclose(g_out);
    printf("packets %d\n", g_packets);
    free(f32); free(i16);
    return 0;
}
