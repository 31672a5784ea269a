/*
 * sk_oracle.c -- CPU restatement of StreamKit's PCM DSP hot path. TEST INFRASTRUCTURE ONLY
 * (see sk_oracle.h header for scope and pin status). Build: gcc -O2 -ffp-contract=off (no
 * -ffast-math): Rust never contracts a*b+c and never reassociates f32 arithmetic.
 */
#include "sk_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ================================================================ gain */

/* gain.rs:50-66 */
int sko_gain_validate(float gain, char *err, size_t err_len) {
    const float MIN_GAIN = 0.0f, MAX_GAIN = 4.0f;
    if (!isfinite(gain)) {
        if (err) snprintf(err, err_len, "Gain must be a finite number, got: %g", (double)gain);
        return 1;
    }
    if (gain < MIN_GAIN || gain > MAX_GAIN) {
        if (err) snprintf(err, err_len, "Gain must be between 0 and 4, got: %g", (double)gain);
        return 2;
    }
    return 0;
}

/* gain.rs:187-189 */
void sko_gain_apply(float *samples, size_t n, float gain) {
    for (size_t i = 0; i < n; i++) samples[i] *= gain;
}

/* ================================================================ s16 (build-defined) */

int16_t sko_f32_to_s16(float x) {
    float y = x * 32768.0f; /* exact power-of-two scaling (may overflow to +-inf: saturates below) */
    if (isnan(y)) return 0;
    if (y >= 32767.0f) return 32767;
    if (y <= -32768.0f) return -32768;
    return (int16_t)lrintf(y); /* default FP environment: round-half-to-even */
}

void sko_f32_to_s16_buf(const float *in, int16_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = sko_f32_to_s16(in[i]);
}

void sko_s16_to_f32_buf(const int16_t *in, float *out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (float)in[i] * (1.0f / 32768.0f);
}

void sko_gain_f32_to_s16_buf(const float *in, int16_t *out, size_t n, float gain) {
    for (size_t i = 0; i < n; i++) {
        float g = in[i] * gain; /* gain.rs:188, rounded to f32 first */
        out[i] = sko_f32_to_s16(g);
    }
}

/* ================================================================ mixer */

/* mixer.rs:1027-1078 */
void sko_mix_frame_with_channel_conversion(float *output, size_t output_len, const sko_frame *source,
                                           uint16_t output_channels) {
    uint16_t source_channels = source->channels;
    size_t samples_per_channel = source->n_samples / source_channels;
    size_t output_samples_per_channel = output_len / output_channels;
    size_t mix_spc = samples_per_channel < output_samples_per_channel ? samples_per_channel : output_samples_per_channel;

    if (source_channels == output_channels) {
        /* :1039-1046  zip(output, source).take(mix_len) */
        size_t mix_len = mix_spc * output_channels;
        if (mix_len > output_len) mix_len = output_len;
        if (mix_len > source->n_samples) mix_len = source->n_samples;
        for (size_t i = 0; i < mix_len; i++) output[i] += source->samples[i];
    } else if (source_channels == 1 && output_channels == 2) {
        /* :1047-1054 */
        for (size_t i = 0; i < mix_spc; i++) {
            float m = source->samples[i];
            output[i * 2] += m;
            output[i * 2 + 1] += m;
        }
    } else if (source_channels == 2 && output_channels == 1) {
        /* :1055-1061  add-then-halve: two roundings */
        for (size_t i = 0; i < mix_spc; i++) {
            float left = source->samples[i * 2];
            float right = source->samples[i * 2 + 1];
            output[i] += (left + right) * 0.5f;
        }
    } else {
        /* :1062-1076 generic cyclic mapping */
        for (size_t i = 0; i < mix_spc; i++) {
            for (size_t ch = 0; ch < output_channels; ch++) {
                size_t sch = ch % source_channels;
                output[i * output_channels + ch] += source->samples[i * source_channels + sch];
            }
        }
    }
}

/* mixer.rs:960-980: base = max_by_key((has_unique_samples, idx)) among frames already of output shape,
 * removed with Vec::swap_remove (the last frame takes its slot), the rest are added in Vec order. */
void sko_mix_plan(const sko_frame *frames, size_t n, uint16_t out_channels, size_t out_size, uint32_t *order,
                  int *has_base) {
    long base = -1;
    int base_unique = -1;
    for (size_t i = 0; i < n; i++) {
        if (frames[i].channels == out_channels && frames[i].n_samples == out_size) {
            int u = frames[i].unique ? 1 : 0;
            /* max_by_key returns the LAST maximum; key = (unique, idx) is strictly increasing in idx
             * for equal unique, so ">=" on unique with ascending idx reproduces it. */
            if (u > base_unique || (u == base_unique)) {
                base = (long)i;
                base_unique = u;
            }
        }
    }
    if (base < 0) {
        *has_base = 0;
        for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
        return;
    }
    *has_base = 1;
    /* Vec after swap_remove(base): element n-1 moves into slot `base` (unless base == n-1). */
    uint32_t *vec = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    for (size_t i = 0; i < n; i++) vec[i] = (uint32_t)i;
    vec[base] = vec[n - 1];
    size_t m = n - 1;
    order[0] = (uint32_t)base;
    for (size_t i = 0; i < m; i++) order[1 + i] = vec[i];
    free(vec);
}

static void mix_with_plan(const sko_frame *frames, size_t n, uint16_t out_channels, size_t out_size, float *out) {
    uint32_t *order = (uint32_t *)malloc(sizeof(uint32_t) * (n ? n : 1));
    int has_base = 0;
    sko_mix_plan(frames, n, out_channels, out_size, order, &has_base);
    size_t first = 0;
    if (has_base) {
        /* base.make_samples_mut(): the base frame's own samples ARE the accumulator (mixer.rs:969-972) */
        memcpy(out, frames[order[0]].samples, out_size * sizeof(float));
        first = 1;
    } else {
        for (size_t i = 0; i < out_size; i++) out[i] = 0.0f; /* mixer.rs:983 vec![0.0f32; output_size] */
    }
    for (size_t i = first; i < n; i++) sko_mix_frame_with_channel_conversion(out, out_size, &frames[order[i]], out_channels);
    free(order);
}

/* mixer.rs:944-1013 */
int sko_mix_sync(const sko_frame *frames, size_t n, uint16_t max_channels_seen, float *out, size_t out_cap,
                 uint16_t *out_channels, size_t *out_len) {
    if (n == 0) { /* :940-942 */
        *out_len = 0;
        *out_channels = max_channels_seen ? max_channels_seen : 1;
        return 0;
    }
    uint16_t current_max = 0; /* :947 (n>0 so unwrap_or(1) never triggers) */
    for (size_t i = 0; i < n; i++)
        if (frames[i].channels > current_max) current_max = frames[i].channels;
    uint16_t oc = max_channels_seen > current_max ? max_channels_seen : current_max;
    if (oc < 1) oc = 1; /* :948 */
    size_t max_spc = 0; /* :952-953 */
    for (size_t i = 0; i < n; i++) {
        size_t spc = frames[i].n_samples / frames[i].channels;
        if (spc > max_spc) max_spc = spc;
    }
    size_t out_size = max_spc * oc; /* :954 */
    if (out_size > out_cap) return -1;
    mix_with_plan(frames, n, oc, out_size, out);
    *out_channels = oc;
    *out_len = out_size;
    return 0;
}

/* mixer.rs:1436-1492 */
int sko_mix_clocked(const sko_frame *frames, size_t n, uint16_t out_channels, size_t frame_samples_per_channel,
                    float *out, size_t out_cap) {
    size_t out_size = frame_samples_per_channel * out_channels; /* :1444 */
    if (out_size > out_cap) return -1;
    mix_with_plan(frames, n, out_channels, out_size, out);
    return 0;
}

/* ================================================================ rubato 0.16.2 FastFixedIn<f32>, Linear */

#define POLYNOMIAL_LEN 8 /* rubato asynchro_fast.rs POLYNOMIAL_LEN_U / _I */

struct sko_ffi {
    size_t nbr_channels;
    size_t chunk_size;
    double last_index;
    double resample_ratio;
    float *buffer; /* planar [channels][chunk + 16] */
};

sko_ffi *sko_ffi_new(double ratio, size_t chunk_frames, size_t channels) {
    if (!(ratio > 0.0) || chunk_frames == 0 || channels == 0) return NULL; /* validate_ratios */
    sko_ffi *r = (sko_ffi *)calloc(1, sizeof(*r));
    r->nbr_channels = channels;
    r->chunk_size = chunk_frames;
    r->last_index = -(double)(POLYNOMIAL_LEN / 2); /* -4.0 */
    r->resample_ratio = ratio;
    r->buffer = (float *)calloc(channels * (chunk_frames + 2 * POLYNOMIAL_LEN), sizeof(float));
    return r;
}

void sko_ffi_free(sko_ffi *r) {
    if (!r) return;
    free(r->buffer);
    free(r);
}

size_t sko_ffi_out_max(const sko_ffi *r) { return (size_t)((double)r->chunk_size * r->resample_ratio + 10.0) + 8; }

double sko_ffi_last_index(const sko_ffi *r) { return r->last_index; }

void sko_ffi_history(const sko_ffi *r, float *hist) {
    size_t stride = r->chunk_size + 2 * POLYNOMIAL_LEN;
    for (size_t j = 0; j < 2 * POLYNOMIAL_LEN; j++)
        for (size_t ch = 0; ch < r->nbr_channels; ch++)
            hist[j * r->nbr_channels + ch] = r->buffer[ch * stride + r->chunk_size + j];
}

/* rubato FastFixedIn::process_into_buffer, PolynomialDegree::Linear branch, fixed ratio
 * (target_ratio == resample_ratio so t_ratio_increment == 0.0 and `t_ratio += 0.0` is the identity).
 * The interleave/deinterleave passes of resampler.rs:397-401 / :413-417 are folded into the indexing. */
size_t sko_ffi_process_interleaved(sko_ffi *r, const float *in, float *out, size_t out_cap_frames) {
    const size_t C = r->nbr_channels, N = r->chunk_size;
    const size_t stride = N + 2 * POLYNOMIAL_LEN;
    /* buf.copy_within(chunk..chunk+16, 0) */
    for (size_t ch = 0; ch < C; ch++) memmove(r->buffer + ch * stride, r->buffer + ch * stride + N, 2 * POLYNOMIAL_LEN * sizeof(float));
    /* buffer[ch][16..16+N] = wave_in[ch][..N] */
    for (size_t i = 0; i < N; i++)
        for (size_t ch = 0; ch < C; ch++) r->buffer[ch * stride + 2 * POLYNOMIAL_LEN + i] = in[i * C + ch];

    double t_ratio = 1.0 / r->resample_ratio;
    long end_idx = (long)N - (POLYNOMIAL_LEN + 1) - (long)ceil(t_ratio);
    double idx = r->last_index;
    size_t n = 0;
    while (idx < (double)end_idx) {
        idx += t_ratio;
        double fl = floor(idx);
        long start_idx = (long)fl;
        float frac = (float)(idx - fl); /* T::coerce(frac) : f64 -> f32 */
        if (n < out_cap_frames) {
            for (size_t ch = 0; ch < C; ch++) {
                const float *bp = r->buffer + ch * stride + (size_t)(start_idx + 2 * POLYNOMIAL_LEN);
                /* interp_lin: (1 - x) * y0 + x * y1, f32, no contraction */
                float a = (1.0f - frac) * bp[0];
                float b = frac * bp[1];
                out[n * C + ch] = a + b;
            }
        }
        n++;
    }
    r->last_index = idx - (double)N;
    return n;
}

/* ================================================================ resampler node (resampler.rs) */

uint64_t sko_duration_us_for_frames(uint32_t sample_rate, size_t frames_per_channel) {
    if (sample_rate == 0) return 0; /* :109-111 */
    return ((uint64_t)frames_per_channel * 1000000ull) / (uint64_t)sample_rate; /* :115 */
}

typedef struct fvec {
    float *p;
    size_t len, cap;
} fvec;

static void fvec_extend(fvec *v, const float *src, size_t n) {
    if (v->len + n > v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 4096;
        while (nc < v->len + n) nc *= 2;
        v->p = (float *)realloc(v->p, nc * sizeof(float));
        v->cap = nc;
    }
    if (n) memcpy(v->p + v->len, src, n * sizeof(float));
    v->len += n;
}
static void fvec_drain_front(fvec *v, size_t n) {
    memmove(v->p, v->p + n, (v->len - n) * sizeof(float));
    v->len -= n;
}

struct sko_rsnode {
    uint32_t target_sample_rate;
    size_t chunk_frames, output_frame_size;
    /* stream state, latched on first audio packet (:206-249) */
    int initialised, needs_resample;
    uint32_t sample_rate;
    uint16_t channels;
    sko_ffi *resampler;
    uint64_t output_sequence;
    int has_ts;
    uint64_t output_timestamp_us;
    fvec sample_buffer;
    size_t sample_buffer_offset;
    fvec output_buffer;
    size_t output_buffer_offset;
    float *scratch;
    size_t scratch_frames;
};

sko_rsnode *sko_rsnode_new(uint32_t target_sample_rate, size_t chunk_frames, size_t output_frame_size, char *err,
                           size_t err_len) {
    if (target_sample_rate == 0) { /* :82-86 */
        if (err) snprintf(err, err_len, "target_sample_rate must be greater than 0");
        return NULL;
    }
    if (chunk_frames == 0) { /* :88-92 */
        if (err) snprintf(err, err_len, "chunk_frames must be greater than 0");
        return NULL;
    }
    if (output_frame_size != 0) { /* :95-102 */
        static const size_t valid[] = {120, 240, 480, 960, 1920, 2880};
        int ok = 0;
        for (size_t i = 0; i < 6; i++) ok |= (valid[i] == output_frame_size);
        if (!ok) {
            if (err)
                snprintf(err, err_len,
                         "output_frame_size must be 0 (disabled) or a valid Opus frame size: [120, 240, 480, 960, 1920, 2880]");
            return NULL;
        }
    }
    sko_rsnode *n = (sko_rsnode *)calloc(1, sizeof(*n));
    n->target_sample_rate = target_sample_rate;
    n->chunk_frames = chunk_frames;
    n->output_frame_size = output_frame_size;
    return n;
}

void sko_rsnode_free(sko_rsnode *n) {
    if (!n) return;
    sko_ffi_free(n->resampler);
    free(n->sample_buffer.p);
    free(n->output_buffer.p);
    free(n->scratch);
    free(n);
}

/* resampler.rs:286-297 next_metadata */
static sko_packet_meta next_metadata(sko_rsnode *n, uint64_t duration_us) {
    sko_packet_meta m;
    m.timestamp_us = n->output_timestamp_us;
    m.has_timestamp = (uint8_t)n->has_ts;
    m.duration_us = duration_us;
    m.sequence = n->output_sequence;
    n->output_sequence += 1;
    if (n->has_ts) n->output_timestamp_us += duration_us;
    return m;
}

/* resampler.rs:323-372 / :425-470: emit exact-size packets from output_buffer, then lazy compaction */
static void drain_output_frames(sko_rsnode *n, sko_emit_fn emit, void *ud) {
    size_t ofs = n->output_frame_size * n->channels;
    while (n->output_buffer.len - n->output_buffer_offset >= ofs) {
        size_t start = n->output_buffer_offset;
        n->output_buffer_offset = start + ofs;
        uint64_t dur = sko_duration_us_for_frames(n->target_sample_rate, n->output_frame_size);
        sko_packet_meta m = next_metadata(n, dur);
        emit(ud, n->target_sample_rate, n->channels, n->output_buffer.p + start, ofs, &m);
    }
    if (n->output_buffer_offset == n->output_buffer.len) {
        n->output_buffer.len = 0;
        n->output_buffer_offset = 0;
    } else if (n->output_buffer_offset > 0 &&
               (n->output_buffer_offset >= ofs * 8 || n->output_buffer_offset * 2 >= n->output_buffer.len)) {
        fvec_drain_front(&n->output_buffer, n->output_buffer_offset); /* perf-only compaction */
        n->output_buffer_offset = 0;
    }
}

int sko_rsnode_push(sko_rsnode *n, uint32_t sample_rate, uint16_t channels, const float *samples, size_t n_samples,
                    int has_timestamp, uint64_t timestamp_us, sko_emit_fn emit, void *ud, char *err, size_t err_len) {
    if (!n->initialised) { /* :206-249 */
        n->initialised = 1;
        n->needs_resample = (sample_rate != n->target_sample_rate);
        n->sample_rate = sample_rate;
        n->channels = channels;
        n->has_ts = has_timestamp ? 1 : 0;
        n->output_timestamp_us = has_timestamp ? timestamp_us : 0;
        if (n->needs_resample) {
            n->resampler = sko_ffi_new((double)n->target_sample_rate / (double)sample_rate, n->chunk_frames, channels);
            n->scratch_frames = sko_ffi_out_max(n->resampler);
            n->scratch = (float *)malloc(n->scratch_frames * channels * sizeof(float));
        }
    }
    if (sample_rate != n->sample_rate || channels != n->channels) { /* :253-279 */
        if (err)
            snprintf(err, err_len, "Audio format changed mid-stream: expected %uHz/%uch, got %uHz/%uch", n->sample_rate,
                     (unsigned)n->channels, sample_rate, (unsigned)channels);
        return -1;
    }
    const size_t C = n->channels;
    if (!n->needs_resample) { /* :299-373 */
        if (n->output_frame_size == 0) {
            /* forwarded untouched, original metadata: we report it as-is (no restamp) */
            sko_packet_meta m;
            m.timestamp_us = timestamp_us;
            m.has_timestamp = (uint8_t)(has_timestamp ? 1 : 0);
            m.duration_us = 0;
            m.sequence = 0;
            emit(ud, sample_rate, channels, samples, n_samples, &m);
            return 0;
        }
        fvec_extend(&n->output_buffer, samples, n_samples);
        drain_output_frames(n, emit, ud);
        return 0;
    }
    /* resampling path :375-527 */
    fvec_extend(&n->sample_buffer, samples, n_samples);
    size_t chunk_samples = n->chunk_frames * C;
    while (n->sample_buffer.len - n->sample_buffer_offset >= chunk_samples) {
        const float *chunk = n->sample_buffer.p + n->sample_buffer_offset;
        size_t out_frames = sko_ffi_process_interleaved(n->resampler, chunk, n->scratch, n->scratch_frames);
        if (n->output_frame_size > 0) {
            fvec_extend(&n->output_buffer, n->scratch, out_frames * C);
            drain_output_frames(n, emit, ud);
        } else { /* :471-512 variable-size packet */
            uint64_t dur = sko_duration_us_for_frames(n->target_sample_rate, out_frames);
            sko_packet_meta m = next_metadata(n, dur);
            emit(ud, n->target_sample_rate, n->channels, n->scratch, out_frames * C, &m);
        }
        n->sample_buffer_offset += chunk_samples;
    }
    if (n->sample_buffer_offset == n->sample_buffer.len) { /* :518-526 */
        n->sample_buffer.len = 0;
        n->sample_buffer_offset = 0;
    } else if (n->sample_buffer_offset > 0 && (n->sample_buffer_offset >= chunk_samples * 4 ||
                                               n->sample_buffer_offset * 2 >= n->sample_buffer.len)) {
        fvec_drain_front(&n->sample_buffer, n->sample_buffer_offset);
        n->sample_buffer_offset = 0;
    }
    return 0;
}

int sko_rsnode_finish(sko_rsnode *n, sko_emit_fn emit, void *ud) {
    /* :543-689 remainder through a FRESH FastFixedIn(chunk = remaining_frames) */
    if (n->sample_buffer.len > n->sample_buffer_offset && n->initialised && n->needs_resample) {
        const size_t C = n->channels;
        size_t remaining_samples = n->sample_buffer.len - n->sample_buffer_offset;
        size_t remaining_frames = remaining_samples / C;
        if (remaining_frames > 0) {
            sko_ffi *rr = sko_ffi_new((double)n->target_sample_rate / (double)n->sample_rate, remaining_frames, C);
            size_t cap = sko_ffi_out_max(rr);
            float *tmp = (float *)malloc((cap ? cap : 1) * C * sizeof(float));
            size_t out_frames = sko_ffi_process_interleaved(rr, n->sample_buffer.p + n->sample_buffer_offset, tmp, cap);
            if (n->output_frame_size > 0) {
                fvec_extend(&n->output_buffer, tmp, out_frames * C);
                drain_output_frames(n, emit, ud);
            } else { /* :655-686 */
                uint64_t dur = sko_duration_us_for_frames(n->target_sample_rate, out_frames);
                sko_packet_meta m = next_metadata(n, dur);
                emit(ud, n->target_sample_rate, n->channels, tmp, out_frames * C, &m);
            }
            free(tmp);
            sko_ffi_free(rr);
        }
    }
    /* :691-730 flush the partial output frame (sequence NOT incremented afterwards, :707-711) */
    if (n->output_buffer.len > n->output_buffer_offset && n->output_frame_size > 0) {
        if (n->output_buffer_offset > 0) {
            fvec_drain_front(&n->output_buffer, n->output_buffer_offset);
            n->output_buffer_offset = 0;
        }
        size_t fpc = n->output_buffer.len / n->channels;
        sko_packet_meta m;
        m.timestamp_us = n->output_timestamp_us;
        m.has_timestamp = (uint8_t)n->has_ts;
        m.duration_us = sko_duration_us_for_frames(n->target_sample_rate, fpc);
        m.sequence = n->output_sequence;
        if (n->has_ts) n->output_timestamp_us += m.duration_us;
        emit(ud, n->target_sample_rate, n->channels, n->output_buffer.p, n->output_buffer.len, &m);
        n->output_buffer.len = 0;
    }
    return 0;
}
