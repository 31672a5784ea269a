// phase_runs.h -- exact, parallel-friendly representation of rubato's sequential phase recurrence.
//
// rubato 0.16.2 FastFixedIn::process_into_buffer (called from the reference at
// crates/nodes/src/audio/filters/resampler.rs:404-407) advances the read position by REPEATED f64
// addition:   while idx < end_idx { idx += t; emit(floor(idx), frac(idx)) }
// A parallel kernel cannot use idx_k = idx_0 + k*t: the roundings differ, and with them (rarely) the
// output count per chunk. This header turns the recurrence into a compact PHASE TABLE that thousands of
// threads evaluate independently and BIT-EXACTLY:
//
//   k <  n_prefix             : idx_k = prefix[k]                      (the first few elements, stored)
//   k in [k_a, k_e) of run r  : idx_k = fma((double)(k - k_a), delta, x_a)        (exact, see below)
//   k == k_e of run r (gap)   : idx_k = idx_{k-1} + t                  (one true addition)
//
// Why a run is exact: inside one binade B = +-[2^e, 2^(e+1)) with unit u = 2^(e-52) every chain element y
// is a multiple of u, so fl(y + t) = y + RN_u(t) as long as the result stays strictly inside B: the
// increment is a constant multiple of u and the fma result is representable, hence unrounded. Ties
// (t an odd multiple of u/2) settle after one step because round-half-even leaves an even mantissa.
// A run is therefore opened only after three consecutive TRUE chain elements a, b, c in the same binade
// (b, c not on the binade's lower edge): anchor b, delta = c - b. Because the increment is constant, the
// generator does not walk a long run element by element: it JUMPS to the last member below
// min(end_idx, binade edge) with one division (+ exact fix-up), so one chunk costs O(#binades) ~ 50
// dependent steps instead of ~960. Everything that is not provably inside a run is produced by the true
// sequential addition. tests/test_phase_runs.py checks every element against the plain recurrence for
// hundreds of millions of (t, last_index) pairs, including tie-prone ratios.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SK_HD __host__ __device__ __forceinline__
#define SK_HD_MEMBER __host__ __device__ __forceinline__
#else
#define SK_HD static inline
#define SK_HD_MEMBER inline
#endif

#define SK_PREFIX_MAX 40u
#define SK_RUNS_MAX 40u

struct SkRun {
    double x_a;    // chain value of the anchor element
    double delta;  // constant increment inside the run (0 for a degenerate single-element entry)
    uint32_t k_a;  // first output index covered
    uint32_t k_e;  // one past the last output index covered; index k_e itself may be an uncovered "gap" element
};

struct SkPhaseTable {       // one process() call of one stream
    uint32_t n_out;         // frames this call produces
    uint32_t n_prefix;      // elements [0, n_prefix) are stored verbatim
    uint32_t n_runs;
    uint32_t overflow;      // 1 = table capacity exceeded (never seen for ratios in [1/256, 256])
    double prefix[SK_PREFIX_MAX];
    SkRun runs[SK_RUNS_MAX];
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
SK_HD double sk_bits2d(uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    union { double d; uint64_t u; } v;
    v.u = b;
    return v.d;
#endif
}
SK_HD double sk_dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
SK_HD double sk_dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// Generates the phase table of one process() call.
//   last_index : rubato's self.last_index on entry;  t : 1.0 / resample_ratio;  end_idx : chunk - 9 - ceil(t)
// Returns n_out; *idx_end receives the final idx (the caller stores last_index' = idx_end - chunk).
// Pointer-based form: prefix[] (capacity cap_np) and runs[] (capacity cap_nr) may live anywhere (a compact per-op layout
// in HBM for the fused chain). Sizes and the overflow flag are returned through out-parameters.
// Streaming form: every prefix element goes to sink.prefix(k, x) and every COMPLETED table entry to sink.run(index, entry),
// both in increasing k, so a consumer (chain_prog.h) can build on the fly without a stored table.
template <class Sink>
SK_HD uint32_t sk_phase_stream(double last_index, double t, int32_t end_idx, uint32_t cap_np, uint32_t cap_nr, Sink &sink,
                               uint32_t *n_prefix_out, uint32_t *n_runs_out, uint32_t *overflow_out, double *idx_end) {
    const double end = (double)end_idx;
    double x = last_index;
    uint32_t k = 0, np = 0, nr = 0;
    bool in_prefix = true, have_run = false;
    uint32_t run_key = 0, prev_key = 0xFFFFFFFFu, cnt = 0;
    bool prev_strict = false;
    double xb = 0.0;              // value of the previous element
    SkRun cur;                    // the open (last) entry, kept in registers; flushed to T->runs[nr-1]
    cur.x_a = 0.0; cur.delta = 0.0; cur.k_a = 0; cur.k_e = 0;
    bool cur_valid = false;
    uint32_t ovf = 0;

#define SK_FLUSH_CUR()                                              \
    do {                                                            \
        if (cur_valid) {                                            \
            if (nr <= cap_nr && nr > 0) sink.run(nr - 1, cur);      \
        }                                                           \
    } while (0)

    while (x < end) {
        x = sk_dadd(x, t);
        const uint64_t bits = sk_d2bits(x);
        const uint32_t key = (uint32_t)(bits >> 52);                 // sign + exponent
        const bool strict = (bits & 0x000FFFFFFFFFFFFFull) != 0ull;  // not on the binade's lower edge
        if (have_run) {
            if (key == run_key && strict) {  // provably fma(k - k_a, delta, x_a): extend the open run
                ++k;
                cur.k_e = k;
                xb = x;
                continue;
            }
            have_run = false;
            prev_key = 0xFFFFFFFFu;
            cnt = 0;
        }
        cnt = (key == prev_key) ? cnt + 1 : 1;
        prev_key = key;
        const bool can_open = cnt >= 3 && strict && prev_strict && k >= 1;
        if (can_open && !(in_prefix && !(xb > 0.0) && np < cap_np)) {
            // a = element k-2, b = element k-1 (value xb), c = this element: open a run anchored at b
            if (in_prefix) {  // b leaves the prefix and becomes the first table entry
                in_prefix = false;
                np = k - 1;
            }
            const bool b_has_entry = cur_valid && cur.k_a == k - 1 && cur.k_e == k;  // degenerate entry for b
            if (!b_has_entry) {
                SK_FLUSH_CUR();
                ++nr;
                if (nr > cap_nr) ovf = 1;
            }
            cur.x_a = xb;
            cur.delta = x - xb;  // exact: same binade
            cur.k_a = k - 1;
            cur.k_e = k + 1;
            cur_valid = true;
            have_run = true;
            run_key = key;
            // ---- jump to the last member strictly below min(end, binade edge)
            const uint32_t E = key & 0x7FFu;
            if (E > 0u && E < 0x7FEu && cur.delta > 0.0) {
                const bool neg = (key & 0x800u) != 0u;
                const double lim = neg ? sk_bits2d((uint64_t)key << 52)            // -2^e (values below it are inside)
                                       : sk_bits2d((uint64_t)(E + 1u) << 52);      // +2^(e+1)
                const double M = lim < end ? lim : end;
                const double q = (M - cur.x_a) / cur.delta;
                if (q > 2.0 && q < 1048576.0) {
                    uint32_t j = (uint32_t)q;
                    while (j > 1u && !(sk_dfma((double)j, cur.delta, cur.x_a) < M)) --j;
                    while (sk_dfma((double)(j + 1u), cur.delta, cur.x_a) < M) ++j;
                    if (j > 1u) {
                        x = sk_dfma((double)j, cur.delta, cur.x_a);
                        k = cur.k_a + j;
                        cur.k_e = k + 1;
                    }
                }
            }
            xb = x;
            prev_strict = true;
            ++k;
            continue;
        }
        // ---- element outside any run
        if (in_prefix) {
            if (np < cap_np) {
                sink.prefix(np++, x);
                prev_strict = strict;
                xb = x;
                ++k;
                continue;
            }
            in_prefix = false;  // prefix full: continue with table entries
        }
        // allowed uncovered: exactly one element right after the end of the last entry (consumer adds t once)
        const bool gap_ok = cur_valid && cur.k_e == k && cur.delta != 0.0;
        if (!gap_ok) {
            SK_FLUSH_CUR();
            ++nr;
            if (nr > cap_nr) ovf = 1;
            cur.x_a = x;
            cur.delta = 0.0;
            cur.k_a = k;
            cur.k_e = k + 1;
            cur_valid = true;
        }
        prev_strict = strict;
        xb = x;
        ++k;
    }
    SK_FLUSH_CUR();
#undef SK_FLUSH_CUR
    *n_prefix_out = np < k ? np : k;
    *n_runs_out = nr < cap_nr ? nr : cap_nr;
    *overflow_out = ovf;
    *idx_end = x;
    return k;
}

struct SkArraySink {   // stores the table (prefix[] / runs[] may live anywhere)
    double *prefix_out;
    SkRun *runs_out;
    SK_HD_MEMBER void prefix(uint32_t i, double x) { prefix_out[i] = x; }
    SK_HD_MEMBER void run(uint32_t i, const SkRun &r) { runs_out[i] = r; }
};
SK_HD uint32_t sk_phase_table_ex(double last_index, double t, int32_t end_idx, double *prefix_out, uint32_t cap_np, SkRun *runs_out,
                                 uint32_t cap_nr, uint32_t *n_prefix_out, uint32_t *n_runs_out, uint32_t *overflow_out, double *idx_end) {
    SkArraySink sink;
    sink.prefix_out = prefix_out;
    sink.runs_out = runs_out;
    return sk_phase_stream(last_index, t, end_idx, cap_np, cap_nr, sink, n_prefix_out, n_runs_out, overflow_out, idx_end);
}

// Generates the phase table of one process() call into a SkPhaseTable.
//   last_index : rubato's self.last_index on entry;  t : 1.0 / resample_ratio;  end_idx : chunk - 9 - ceil(t)
// Returns n_out; *idx_end receives the final idx (the caller stores last_index' = idx_end - chunk).
SK_HD uint32_t sk_phase_table(double last_index, double t, int32_t end_idx, SkPhaseTable *T, double *idx_end) {
    uint32_t np, nr, ovf;
    const uint32_t n = sk_phase_table_ex(last_index, t, end_idx, T->prefix, SK_PREFIX_MAX, T->runs, SK_RUNS_MAX, &np, &nr, &ovf, idx_end);
    T->n_out = n;
    T->n_prefix = np;
    T->n_runs = nr;
    T->overflow = ovf;
    return n;
}

// Consumer side: idx of output k (k < n_out). `r` is a cursor the caller may carry between calls with
// non-decreasing k (start it at 0).
SK_HD double sk_phase_eval(const double *prefix, uint32_t n_prefix, const SkRun *runs, uint32_t n_runs, double t, uint32_t k,
                           uint32_t *r_io) {
    if (k < n_prefix) return prefix[k];
    uint32_t r = *r_io;
    while (r + 1u < n_runs && runs[r + 1u].k_a <= k) ++r;
    *r_io = r;
    const SkRun rn = runs[r];
    if (k < rn.k_e) return sk_dfma((double)(k - rn.k_a), rn.delta, rn.x_a);
    // gap element: one true addition after the run's last member
    return sk_dadd(sk_dfma((double)(rn.k_e - 1u - rn.k_a), rn.delta, rn.x_a), t);
}
