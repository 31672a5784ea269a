/* host_shim.c -- the C half of the test host for the native plugin ABI (test infrastructure).
 * ctypes cannot return structs from callbacks, so the output / telemetry callbacks live here. They do what
 * the Rust host does inside output_callback_shim: COPY the packet (conversions.rs:340-346) and queue it
 * (wrapper.rs:554). */
#include <stdlib.h>
#include <string.h>

#include "../../include/streamkit_native_abi.h"

typedef struct skh_out {
    int packet_type;
    uint32_t sample_rate;
    uint16_t channels;
    size_t n;      /* samples (audio) or bytes (binary) */
    void *data;    /* owned copy */
    char pin[32];
} skh_out;

typedef struct skh_collector {
    skh_out *outs;
    size_t n, cap;
    size_t telemetry_events;
} skh_collector;

skh_collector *skh_collector_new(void) { return (skh_collector *)calloc(1, sizeof(skh_collector)); }

void skh_collector_clear(skh_collector *c) {
    for (size_t i = 0; i < c->n; i++) free(c->outs[i].data);
    c->n = 0;
}
void skh_collector_free(skh_collector *c) {
    if (!c) return;
    skh_collector_clear(c);
    free(c->outs);
    free(c);
}
size_t skh_collector_count(const skh_collector *c) { return c->n; }
const skh_out *skh_collector_get(const skh_collector *c, size_t i) { return &c->outs[i]; }

sk_result skh_output_cb(const char *pin, const sk_packet *pkt, void *ud) {
    skh_collector *c = (skh_collector *)ud;
    sk_result ok = {true, NULL};
    sk_result bad = {false, "host: unsupported packet"};
    if (!pkt || !pkt->data) return bad;
    if (c->n == c->cap) {
        c->cap = c->cap ? c->cap * 2 : 8;
        c->outs = (skh_out *)realloc(c->outs, c->cap * sizeof(skh_out));
    }
    skh_out *o = &c->outs[c->n];
    memset(o, 0, sizeof(*o));
    strncpy(o->pin, pin ? pin : "", sizeof(o->pin) - 1);
    o->packet_type = (int)pkt->packet_type;
    if (pkt->packet_type == SK_PACKET_RAW_AUDIO) {
        const sk_audio_frame *f = (const sk_audio_frame *)pkt->data;
        if (!f->samples && f->sample_count) return bad;
        o->sample_rate = f->sample_rate;
        o->channels = f->channels;
        o->n = f->sample_count;
        o->data = malloc(f->sample_count * sizeof(float) + 1);
        memcpy(o->data, f->samples, f->sample_count * sizeof(float));
    } else if (pkt->packet_type == SK_PACKET_BINARY) {
        o->n = pkt->len;
        o->data = malloc(pkt->len + 1);
        memcpy(o->data, pkt->data, pkt->len);
    } else {
        return bad;
    }
    c->n++;
    return ok;
}

sk_result skh_telemetry_cb(const char *event_type, const uint8_t *json, size_t len, const sk_packet_metadata *md, void *ud) {
    (void)event_type; (void)json; (void)len; (void)md;
    if (ud) ((skh_collector *)ud)->telemetry_events++;
    sk_result ok = {true, NULL};
    return ok;
}
