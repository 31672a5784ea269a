// Host-side test shim for phase_runs.h (the SAME source the CUDA kernels compile): lets the CPU test
// suite verify the run-table representation against the plain sequential recurrence.
#include "phase_runs.h"
#include <cmath>
#include <vector>

extern "C" {

uint32_t skp_runs(double last_index, double t, int32_t end_idx, SkRun *runs, uint32_t rmax, uint32_t *n_runs,
                  double *idx_end, int *overflow) {
    return sk_phase_runs(last_index, t, end_idx, runs, rmax, n_runs, idx_end, overflow);
}

// Runs `calls` consecutive process() calls of a stream (chunk frames each) starting from last_index and
// compares every element of the run-table reconstruction with the sequential chain.
// Returns the number of mismatching elements (0 = exact); writes stats.
uint64_t skp_check_stream(double ratio, uint32_t chunk, double last_index, uint32_t calls, uint32_t *max_runs,
                          uint32_t *overflows, double *last_index_out, uint64_t *total_out) {
    const double t = 1.0 / ratio;
    const int32_t end_idx = (int32_t)chunk - 9 - (int32_t)std::ceil(t);
    std::vector<SkRun> runs(SK_RUNS_MAX);
    uint64_t bad = 0, total = 0;
    uint32_t mr = 0, ov = 0;
    double L = last_index;
    for (uint32_t c = 0; c < calls; ++c) {
        uint32_t nr = 0;
        double idx_end = 0;
        int ovf = 0;
        uint32_t n = sk_phase_runs(L, t, end_idx, runs.data(), SK_RUNS_MAX, &nr, &idx_end, &ovf);
        if (nr > mr) mr = nr;
        ov += (uint32_t)ovf;
        // sequential reference
        double idx = L;
        uint32_t k = 0, r = 0;
        while (idx < (double)end_idx) {
            idx += t;
            if (!ovf) {
                while (r + 1 < nr && runs[r + 1].k_a <= k) ++r;
                double pred = sk_phase_eval(runs[r], k);
                if (!(pred == idx) || runs[r].k_a > k) ++bad;
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
}
