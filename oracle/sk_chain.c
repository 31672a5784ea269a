/*
 * sk_chain.c -- CPU baseline runner for the full chain (TEST / BENCH INFRASTRUCTURE ONLY).
 *
 * "Reference-shaped" (BASELINE.md section 2 variant ii): one object per node per session,
 * one 20 ms packet per call, a freshly allocated frame per emitted packet -- the per-packet shape of
 * the real nodes -- composed exactly like streamkit_b200.chain builds the GPU tick:
 *     K x [audio::resampler(chunk = in frames/tick, output_frame_size 960) -> audio::gain]
 *       -> audio::mixer (clocked, 960) -> audio::gain -> f32->s16
 * using the oracle's node restatements (sk_oracle.c). Sessions are partitioned over pthreads.
 * tokio scheduling / channel hops of the real reference are NOT included, so this baseline is
 * optimistic for the reference.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sk_oracle.h"

#define OUT_FRAMES 960
#define QCAP 4

typedef struct pkt_queue {
    float *pk[QCAP];
    size_t n[QCAP];
    int head, count;
    float gain;
} pkt_queue;

static void on_emit(void *ud, uint32_t rate, uint16_t ch, const float *samples, size_t n, const sko_packet_meta *m) {
    (void)rate; (void)ch; (void)m;
    pkt_queue *q = (pkt_queue *)ud;
    float *copy = (float *)malloc(n * sizeof(float)); /* the emitted AudioFrame (pool get + copy, resampler.rs:184-194) */
    memcpy(copy, samples, n * sizeof(float));
    sko_gain_apply(copy, n, q->gain);                 /* audio::gain node on the packet (gain.rs:187-189) */
    if (q->count == QCAP) {                            /* InputRingBuffer overwrite-oldest (mixer.rs:1195-1201) */
        free(q->pk[q->head]);
        q->head = (q->head + 1) % QCAP;
        q->count--;
    }
    int tail = (q->head + q->count) % QCAP;
    q->pk[tail] = copy;
    q->n[tail] = n;
    q->count++;
}

typedef struct chain_job {
    uint32_t s_begin, s_end, k_inputs, ticks, in_rate, chunk, pool_streams;
    uint16_t channels;
    const float *input;  /* [pool_streams][chunk*channels] */
    const float *in_gains, *master_gains;
    int16_t *out_last;   /* [n_sessions][960*channels] or NULL */
    uint64_t checksum;
    pthread_barrier_t *start, *stop;
} chain_job;

static void *chain_worker(void *arg) {
    chain_job *j = (chain_job *)arg;
    const uint32_t K = j->k_inputs, C = j->channels;
    const size_t in_len = (size_t)j->chunk * C, out_len = (size_t)OUT_FRAMES * C;
    const uint32_t ns = j->s_end - j->s_begin;
    uint64_t sum = 0;
    float *mixed = (float *)malloc(out_len * sizeof(float));
    int16_t *s16 = (int16_t *)malloc(out_len * sizeof(int16_t));
    /* session set-up (node construction) is NOT timed: a server creates nodes once per session */
    sko_rsnode **nodes = (sko_rsnode **)calloc((size_t)ns * K + 1, sizeof(*nodes));
    pkt_queue *qs = (pkt_queue *)calloc((size_t)ns * K + 1, sizeof(*qs));
    for (uint32_t s = 0; s < ns; s++)
        for (uint32_t i = 0; i < K; i++) {
            nodes[(size_t)s * K + i] = sko_rsnode_new(48000, j->chunk, OUT_FRAMES, NULL, 0);
            qs[(size_t)s * K + i].gain = j->in_gains[(size_t)(j->s_begin + s) * K + i];
        }
    pthread_barrier_wait(j->start);
    /* tick-major, like a server: every 20 ms each live session processes one frame per input */
    for (uint32_t t = 0; t < j->ticks; t++) {
        for (uint32_t s = 0; s < ns; s++) {
            sko_frame frames[64];
            float *popped[64];
            size_t nf = 0;
            for (uint32_t i = 0; i < K; i++) {
                const size_t stream = (size_t)(j->s_begin + s) * K + i;
                const float *x = j->input + (stream % j->pool_streams) * in_len;
                pkt_queue *q = &qs[(size_t)s * K + i];
                sko_rsnode_push(nodes[(size_t)s * K + i], j->in_rate, (uint16_t)C, x, in_len, 0, 0, on_emit, q, NULL, 0);
                if (q->count > 0 && nf < 64) { /* clocked mixer pops <= 1 frame per input per tick */
                    frames[nf].samples = q->pk[q->head];
                    frames[nf].n_samples = (uint32_t)q->n[q->head];
                    frames[nf].channels = (uint16_t)C;
                    frames[nf].unique = 1;
                    popped[nf] = q->pk[q->head];
                    q->head = (q->head + 1) % QCAP;
                    q->count--;
                    nf++;
                }
            }
            sko_mix_clocked(frames, nf, (uint16_t)C, OUT_FRAMES, mixed, out_len);
            sko_gain_f32_to_s16_buf(mixed, s16, out_len, j->master_gains[j->s_begin + s]);
            for (size_t f = 0; f < nf; f++) free(popped[f]);
            for (size_t e = 0; e < out_len; e += 97) sum += (uint16_t)s16[e];
            if (j->out_last && t + 1 == j->ticks) memcpy(j->out_last + (size_t)(j->s_begin + s) * out_len, s16, out_len * sizeof(int16_t));
        }
    }
    pthread_barrier_wait(j->stop);
    for (size_t i = 0; i < (size_t)ns * K; i++) {
        sko_rsnode_free(nodes[i]);
        while (qs[i].count > 0) { free(qs[i].pk[qs[i].head]); qs[i].head = (qs[i].head + 1) % QCAP; qs[i].count--; }
    }
    free(nodes);
    free(qs);
    free(mixed);
    free(s16);
    j->checksum = sum;
    return NULL;
}

/* Runs `ticks` ticks of `n_sessions` sessions on `threads` pthreads. Returns the wall seconds (CLOCK_MONOTONIC) of the
 * tick loop only (node construction / teardown excluded). */
double sko_chain_bench(uint32_t n_sessions, uint32_t k_inputs, uint32_t ticks, uint32_t in_rate, uint16_t channels,
                       const float *input, uint32_t pool_streams, const float *in_gains, const float *master_gains,
                       int threads, int16_t *out_last, uint64_t *checksum) {
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > n_sessions) threads = (int)(n_sessions ? n_sessions : 1);
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(*th));
    chain_job *jobs = (chain_job *)calloc((size_t)threads, sizeof(*jobs));
    pthread_barrier_t start, stop;
    pthread_barrier_init(&start, NULL, (unsigned)threads + 1);
    pthread_barrier_init(&stop, NULL, (unsigned)threads + 1);
    struct timespec t0, t1;
    for (int i = 0; i < threads; i++) {
        jobs[i].s_begin = (uint32_t)((uint64_t)n_sessions * i / threads);
        jobs[i].s_end = (uint32_t)((uint64_t)n_sessions * (i + 1) / threads);
        jobs[i].k_inputs = k_inputs; jobs[i].ticks = ticks; jobs[i].in_rate = in_rate; jobs[i].chunk = in_rate / 50;
        jobs[i].pool_streams = pool_streams; jobs[i].channels = channels; jobs[i].input = input;
        jobs[i].in_gains = in_gains; jobs[i].master_gains = master_gains; jobs[i].out_last = out_last;
        jobs[i].start = &start; jobs[i].stop = &stop;
        pthread_create(&th[i], NULL, chain_worker, &jobs[i]);
    }
    pthread_barrier_wait(&start);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&stop);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    uint64_t sum = 0;
    for (int i = 0; i < threads; i++) { pthread_join(th[i], NULL); sum += jobs[i].checksum; }
    if (checksum) *checksum = sum;
    pthread_barrier_destroy(&start);
    pthread_barrier_destroy(&stop);
    free(th);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------------------------------------------------------
 * CPU runners of the standalone node workloads (BASELINE configs 2-4), reference-shaped: one node object per
 * stream / mixer, one 20 ms packet per call, output in a freshly written buffer. `unit` u plays input
 * pool[(u * stride) % pool_units]. Returns the wall seconds of `iters` passes over `n_units` units on `threads` pthreads.
 * kind: 0 = audio::gain -> f32->s16 (1920 samples), 1 = audio::mixer 64 stereo inputs (clocked, 960) -> gain -> s16,
 *       2 = audio::resampler chunk = in_rate/50, variable-length output. */
typedef struct node_job {
    int kind;
    uint32_t u_begin, u_end, iters, in_rate, out_rate, k_inputs, pool_units;
    const float *pool;
    const float *gains;
    int16_t *out_s16;   /* kind 0/1: [n_units][1920] of the last pass, or NULL */
    float *out_f32;     /* kind 2: [n_units][cap*2] of the last pass, or NULL */
    uint32_t *out_n;    /* kind 2: frames per unit of the last pass */
    uint32_t cap;
    pthread_barrier_t *start, *stop;
} node_job;

static void *node_worker(void *arg) {
    node_job *j = (node_job *)arg;
    const size_t N = 1920;
    const uint32_t nu = j->u_end - j->u_begin;
    int16_t *s16 = (int16_t *)malloc(N * sizeof(int16_t));
    float *mixed = (float *)malloc(N * sizeof(float));
    sko_ffi **rs = NULL;
    float *rs_out = NULL;
    const uint32_t chunk = j->in_rate / 50;
    if (j->kind == 2) {
        rs = (sko_ffi **)calloc(nu + 1, sizeof(*rs));
        for (uint32_t u = 0; u < nu; u++) rs[u] = sko_ffi_new((double)j->out_rate / (double)j->in_rate, chunk, 2);
        rs_out = (float *)malloc((size_t)j->cap * 2 * sizeof(float));
    }
    pthread_barrier_wait(j->start);
    for (uint32_t it = 0; it < j->iters; it++) {
        for (uint32_t u = 0; u < nu; u++) {
            const uint32_t gu = j->u_begin + u;
            if (j->kind == 0) {
                const float *x = j->pool + (size_t)(gu % j->pool_units) * N;
                sko_gain_f32_to_s16_buf(x, s16, N, j->gains[gu]);
                if (j->out_s16 && it + 1 == j->iters) memcpy(j->out_s16 + (size_t)gu * N, s16, N * sizeof(int16_t));
            } else if (j->kind == 1) {
                sko_frame fr[64];
                for (uint32_t i = 0; i < j->k_inputs; i++) {
                    fr[i].samples = j->pool + (size_t)(((size_t)gu * 37 + (size_t)i * 101) % j->pool_units) * N;
                    fr[i].n_samples = (uint32_t)N; fr[i].channels = 2; fr[i].unique = 1;
                }
                sko_mix_clocked(fr, j->k_inputs, 2, N / 2, mixed, N);
                sko_gain_f32_to_s16_buf(mixed, s16, N, j->gains[gu]);
                if (j->out_s16 && it + 1 == j->iters) memcpy(j->out_s16 + (size_t)gu * N, s16, N * sizeof(int16_t));
            } else {
                const float *x = j->pool + (size_t)(gu % j->pool_units) * chunk * 2;
                const size_t n = sko_ffi_process_interleaved(rs[u], x, rs_out, j->cap);
                if (j->out_f32 && it + 1 == j->iters) {
                    memcpy(j->out_f32 + (size_t)gu * j->cap * 2, rs_out, n * 2 * sizeof(float));
                    j->out_n[gu] = (uint32_t)n;
                }
            }
        }
    }
    pthread_barrier_wait(j->stop);
    if (rs) { for (uint32_t u = 0; u < nu; u++) sko_ffi_free(rs[u]); free(rs); }
    free(rs_out); free(s16); free(mixed);
    return NULL;
}

double sko_node_bench(int kind, uint32_t n_units, uint32_t iters, uint32_t in_rate, uint32_t out_rate, uint32_t k_inputs,
                      const float *pool, uint32_t pool_units, const float *gains, int threads, int16_t *out_s16, float *out_f32,
                      uint32_t *out_n, uint32_t cap) {
    if (threads < 1) threads = 1;
    if ((uint32_t)threads > n_units) threads = (int)(n_units ? n_units : 1);
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(*th));
    node_job *jobs = (node_job *)calloc((size_t)threads, sizeof(*jobs));
    pthread_barrier_t start, stop;
    pthread_barrier_init(&start, NULL, (unsigned)threads + 1);
    pthread_barrier_init(&stop, NULL, (unsigned)threads + 1);
    struct timespec t0, t1;
    for (int i = 0; i < threads; i++) {
        jobs[i].kind = kind;
        jobs[i].u_begin = (uint32_t)((uint64_t)n_units * i / threads);
        jobs[i].u_end = (uint32_t)((uint64_t)n_units * (i + 1) / threads);
        jobs[i].iters = iters; jobs[i].in_rate = in_rate; jobs[i].out_rate = out_rate; jobs[i].k_inputs = k_inputs;
        jobs[i].pool = pool; jobs[i].pool_units = pool_units; jobs[i].gains = gains;
        jobs[i].out_s16 = out_s16; jobs[i].out_f32 = out_f32; jobs[i].out_n = out_n; jobs[i].cap = cap;
        jobs[i].start = &start; jobs[i].stop = &stop;
        pthread_create(&th[i], NULL, node_worker, &jobs[i]);
    }
    pthread_barrier_wait(&start);
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&stop);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int i = 0; i < threads; i++) pthread_join(th[i], NULL);
    pthread_barrier_destroy(&start);
    pthread_barrier_destroy(&stop);
    free(th);
    free(jobs);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
