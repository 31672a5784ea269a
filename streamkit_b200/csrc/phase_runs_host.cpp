// Host-side test shim for phase_runs.h (the SAME source the CUDA kernels compile): lets the CPU test
// suite verify the phase-table representation against the plain sequential recurrence.
#include "phase_runs.h"
#include <cmath>
#include <vector>

extern "C" {

uint32_t skp_table(double last_index, double t, int32_t end_idx, SkPhaseTable *T, double *idx_end) {
    return sk_phase_table(last_index, t, end_idx, T, idx_end);
}

// Runs `calls` consecutive process() calls of a stream (chunk frames each) starting from last_index and
// compares every element of the table reconstruction with the sequential chain.
// Returns the number of mismatching elements (0 = exact); writes stats.
uint64_t skp_check_stream(double ratio, uint32_t chunk, double last_index, uint32_t calls, uint32_t *max_runs,
                          uint32_t *overflows, double *last_index_out, uint64_t *total_out) {
    const double t = 1.0 / ratio;
    const int32_t end_idx = (int32_t)chunk - 9 - (int32_t)std::ceil(t);
    SkPhaseTable T;
    uint64_t bad = 0, total = 0;
    uint32_t mr = 0, ov = 0;
    double L = last_index;
    for (uint32_t c = 0; c < calls; ++c) {
        double idx_end = 0;
        uint32_t n = sk_phase_table(L, t, end_idx, &T, &idx_end);
        if (T.n_runs > mr) mr = T.n_runs;
        ov += T.overflow;
        // sequential reference
        double idx = L;
        uint32_t k = 0, r = 0, r2;
        while (idx < (double)end_idx) {
            idx += t;
            if (!T.overflow && k < n) {
                double pred = sk_phase_eval(T.prefix, T.n_prefix, T.runs, T.n_runs, t, k, &r);
                if (!(pred == idx)) ++bad;
                if ((k & 63u) == 0u) {  // random access with a cold cursor must agree too
                    r2 = 0;
                    if (!(sk_phase_eval(T.prefix, T.n_prefix, T.runs, T.n_runs, t, k, &r2) == idx)) ++bad;
                }
            }
            ++k;
        }
        if (k != n) ++bad;
        if (idx != idx_end) ++bad;
        total += n;
        L = idx_end - (double)chunk;
    }
    *max_runs = mr;
    *overflows = ov;
    *last_index_out = L;
    *total_out = total;
    return bad;
}

// number of sequential generator iterations is not observable from outside; expose table sizes for the docs
void skp_table_sizes(uint32_t *table_bytes, uint32_t *prefix_max, uint32_t *runs_max) {
    *table_bytes = (uint32_t)sizeof(SkPhaseTable);
    *prefix_max = SK_PREFIX_MAX;
    *runs_max = SK_RUNS_MAX;
}
}
