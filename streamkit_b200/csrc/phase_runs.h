// phase_runs.h -- exact, parallel-friendly representation of rubato's sequential phase recurrence.
//
// rubato 0.16.2 FastFixedIn::process_into_buffer (called from the reference at
// crates/nodes/src/audio/filters/resampler.rs:404-407) advances the read position by REPEATED f64
// addition:   while idx < end_idx { idx += t; emit(floor(idx), frac(idx)) }
// A parallel kernel cannot use idx_k = idx_0 + k*t: the roundings differ, and with them (rarely) the
// output count per chunk. This header turns the recurrence into a short table of "runs" that thousands
// of threads can evaluate independently and BIT-EXACTLY:
//
//     for k in [k_a, next run's k_a):   idx_k == fma((double)(k - k_a), delta, x_a)      (exact)
//
// Why it is exact: inside one binade B = +-[2^e, 2^(e+1)) with unit u = 2^(e-52), every chain element y
// is a multiple of u, so fl(y + t) = y + RN_u(t) as long as the result stays strictly inside B; the
// increment is a constant multiple of u and the fma result is representable, hence unrounded. Ties
// (t an odd multiple of u/2) settle after one step because round-half-even leaves an even mantissa.
// A run is therefore opened only after three consecutive TRUE chain elements a, b, c in the same
// binade (b, c not on the binade's lower boundary): anchor b, delta = c - b. Every other element is
// stored as a singleton (delta = 0). Membership of later elements is decided on the TRUE chain value
// (same sign/exponent word, mantissa != 0), which implies equality with the fma prediction.
// tests/test_phase_runs.py checks this against the plain recurrence for millions of (t, last_index).
//
// The generator is the sequential chain itself (one thread per stream, data independent: it needs only
// last_index, t and end_idx), so correctness never depends on the closed form being clever.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SK_HD __host__ __device__ __forceinline__
#else
#define SK_HD static inline
#endif

#define SK_RUNS_MAX 96u

struct SkRun {
    double x_a;    // chain value of the anchor element
    double delta;  // constant increment inside the run (0 for a singleton)
    uint32_t k_a;  // output index (within the chunk) of the anchor
    uint32_t _pad;
};

SK_HD uint64_t sk_d2bits(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    union { double d; uint64_t u; } v;
    v.d = x;
    return v.u;
#endif
}

SK_HD double sk_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

// Generates the run table for one process() call.
//   last_index : rubato's self.last_index on entry
//   t          : 1.0 / resample_ratio (f64)
//   end_idx    : chunk - 9 - ceil(t)
// Returns the number of output frames n; *n_runs receives the table length (clamped to rmax and
// *overflow set when more runs were needed); *idx_end receives the final idx (last_index' = idx_end - chunk).
SK_HD uint32_t sk_phase_runs(double last_index, double t, int32_t end_idx, SkRun *runs, uint32_t rmax,
                             uint32_t *n_runs, double *idx_end, int *overflow) {
    const double end = (double)end_idx;
    double x = last_index;
    uint32_t k = 0, nr = 0;
    bool have_run = false;
    uint32_t run_key = 0;
    uint32_t prev_key = 0xFFFFFFFFu, cnt = 0;
    bool prev_strict = false;
    int ovf = 0;
    while (x < end) {
        x = sk_dadd(x, t);
        const uint64_t bits = sk_d2bits(x);
        const uint32_t key = (uint32_t)(bits >> 52);                 // sign + exponent
        const bool strict = (bits & 0x000FFFFFFFFFFFFFull) != 0ull;  // not on the binade's lower edge
        if (have_run) {
            if (key == run_key && strict) {
                ++k;
                continue;
            }
            have_run = false;
            prev_key = 0xFFFFFFFFu;
            cnt = 0;
        }
        cnt = (key == prev_key) ? cnt + 1 : 1;
        prev_key = key;
        if (cnt >= 3 && strict && prev_strict && nr > 0 && nr <= rmax) {
            // a = element k-2, b = element k-1 (entry nr-1), c = this element
            runs[nr - 1].delta = x - runs[nr - 1].x_a;  // exact: same binade
            have_run = true;
            run_key = key;
        } else {
            if (nr < rmax) {
                runs[nr].x_a = x;
                runs[nr].delta = 0.0;
                runs[nr].k_a = k;
                runs[nr]._pad = 0;
            } else {
                ovf = 1;
            }
            ++nr;
        }
        prev_strict = strict;
        ++k;
    }
    *n_runs = nr < rmax ? nr : rmax;
    *idx_end = x;
    *overflow = ovf;
    return k;
}

// Evaluate element k from its run (the consumer side; exact by construction).
SK_HD double sk_phase_eval(const SkRun &r, uint32_t k) {
#if defined(__CUDA_ARCH__)
    return __fma_rn((double)(k - r.k_a), r.delta, r.x_a);
#else
    return __builtin_fma((double)(k - r.k_a), r.delta, r.x_a);
#endif
}
