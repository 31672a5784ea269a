/*
 * sk_sinc.c -- CPU oracle of the windowed-sinc polyphase resampler mode (TEST INFRASTRUCTURE ONLY).
 *
 * BASELINE.json's north star names "a batched polyphase windowed-sinc resampler that keeps filter state per stream in HBM".
 * The reference itself resamples with rubato FastFixedIn / PolynomialDegree::Linear (resampler.rs:232-238), so this mode has
 * NO reference implementation: it is specified here (last paragraph: "restate rubato SincFixedIn parameters
 * explicitly in the build's own spec ... and pin with the build's own oracle"). PARITY: pinned to this spec only.
 *
 * Spec `sinc` (parameters as in rubato's SincInterpolationParameters):
 *   sinc_len L (taps per phase, multiple of 8), oversampling_factor O (phases), f_cutoff (relative to the lower Nyquist),
 *   window = BlackmanHarris2 (4-term Blackman-Harris, squared), interpolation = Linear (between the two nearest phases).
 *   fc = f_cutoff * min(1, out_rate / in_rate).
 *   Continuous kernel  g(tau) = fc * sinc(fc * tau) * w((tau + L/2) / L),  sinc(z) = sin(pi z) / (pi z),
 *                      w(u) = (0.35875 - 0.48829 cos(2 pi u) + 0.14128 cos(4 pi u) - 0.01168 cos(6 pi u))^2.
 *   Tap table (computed in f64, stored f32): T[p][n] = g(L/2 - 1 - n + p/O) / sum_n g(...),  p = 0..O, n = 0..L-1 (unit DC gain per phase).
 *   State per stream: last_index (f64, starts at -(L/2)) and the last H = L + 8 input frames (zeros at the start).
 *   process(chunk of N frames): buffer = [H history | chunk]; t = 1 / ratio; end_idx = N - L/2 - 1 - ceil(t);
 *       idx = last_index; while idx < end_idx { idx += t; fl = floor(idx); frac = idx - fl; fo = frac * O; p = min(floor(fo), O - 1);
 *       q = f32(fo - p); base = fl - L/2 + 1 + H; for each channel: y0 = sum_n fma(buf[base + n], T[p][n], .) (n ascending, f32 fma),
 *       y1 likewise with T[p + 1]; out = (1 - q) * y0 + q * y1 (three f32 roundings, as rubato's interp_lin) }
 *       last_index = idx - N; the history becomes the last H frames of the buffer.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct sko_sinc {
    double ratio, last_index;
    size_t chunk, channels, L, O, H;
    float *taps;   /* [(O + 1)][L] */
    float *buf;    /* [(H + chunk)][channels] interleaved */
} sko_sinc;

static double bh2(double u) {
    const double two_pi = 6.283185307179586476925286766559;
    const double w = 0.35875 - 0.48829 * cos(two_pi * u) + 0.14128 * cos(2.0 * two_pi * u) - 0.01168 * cos(3.0 * two_pi * u);
    return w * w;
}

/* the tap table of the spec; out[(O + 1) * L] */
void sko_sinc_taps(size_t L, size_t O, double fc, float *out) {
    const double pi = 3.14159265358979323846264338327950288;
    double *g = (double *)malloc(L * sizeof(double));
    for (size_t p = 0; p <= O; p++) {
        double sum = 0.0;
        for (size_t n = 0; n < L; n++) {
            const double tau = (double)L / 2.0 - 1.0 - (double)n + (double)p / (double)O;
            const double z = fc * tau;
            const double s = z == 0.0 ? 1.0 : sin(pi * z) / (pi * z);
            double u = (tau + (double)L / 2.0) / (double)L;
            double w = (u <= 0.0 || u >= 1.0) ? 0.0 : bh2(u);
            g[n] = fc * s * w;
            sum += g[n];
        }
        for (size_t n = 0; n < L; n++) out[p * L + n] = (float)(g[n] / sum);
    }
    free(g);
}

sko_sinc *sko_sinc_new(double ratio, size_t chunk_frames, size_t channels, size_t sinc_len, size_t oversampling, double f_cutoff) {
    if (!(ratio > 0.0) || chunk_frames == 0 || channels == 0 || sinc_len < 8 || sinc_len % 8 || oversampling < 1) return NULL;
    sko_sinc *r = (sko_sinc *)calloc(1, sizeof(*r));
    r->ratio = ratio; r->chunk = chunk_frames; r->channels = channels; r->L = sinc_len; r->O = oversampling; r->H = sinc_len + 8;
    r->last_index = -(double)(sinc_len / 2);
    r->taps = (float *)malloc((oversampling + 1) * sinc_len * sizeof(float));
    sko_sinc_taps(sinc_len, oversampling, f_cutoff * (ratio < 1.0 ? ratio : 1.0), r->taps);
    r->buf = (float *)calloc((r->H + chunk_frames) * channels, sizeof(float));
    return r;
}
void sko_sinc_free(sko_sinc *r) {
    if (!r) return;
    free(r->taps); free(r->buf); free(r);
}
size_t sko_sinc_out_max(const sko_sinc *r) { return (size_t)((double)r->chunk * r->ratio + 10.0) + 8; }
double sko_sinc_last_index(const sko_sinc *r) { return r->last_index; }

size_t sko_sinc_process_interleaved(sko_sinc *r, const float *in, float *out, size_t out_cap_frames) {
    const size_t C = r->channels, N = r->chunk, L = r->L, H = r->H;
    memmove(r->buf, r->buf + N * C, H * C * sizeof(float));              /* history := last H frames of [history | chunk] */
    memcpy(r->buf + H * C, in, N * C * sizeof(float));
    const double t = 1.0 / r->ratio;
    const long end_idx = (long)N - (long)(L / 2) - 1 - (long)ceil(t);
    double idx = r->last_index;
    size_t n_out = 0;
    while (idx < (double)end_idx) {
        idx += t;
        const double fl = floor(idx);
        const double fo = (idx - fl) * (double)r->O;
        long p = (long)floor(fo);
        if (p > (long)r->O - 1) p = (long)r->O - 1;
        const float q = (float)(fo - (double)p);
        const long base = (long)fl - (long)(L / 2) + 1 + (long)H;
        if (n_out < out_cap_frames) {
            const float *t0 = r->taps + (size_t)p * L, *t1 = t0 + L;
            for (size_t ch = 0; ch < C; ch++) {
                float y0 = 0.0f, y1 = 0.0f;
                for (size_t n = 0; n < L; n++) {
                    const float x = r->buf[((size_t)base + n) * C + ch];
                    y0 = fmaf(x, t0[n], y0);
                    y1 = fmaf(x, t1[n], y1);
                }
                const float a = (1.0f - q) * y0;
                const float b = q * y1;
                out[n_out * C + ch] = a + b;
            }
        }
        n_out++;
    }
    r->last_index = idx - (double)N;
    return n_out;
}
