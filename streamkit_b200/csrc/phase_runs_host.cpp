// Host-side test shim for phase_runs.h (the SAME source the CUDA kernels compile): lets the CPU test
// suite verify the phase-table representation against the plain sequential recurrence.
#include "phase_runs.h"
#include <algorithm>
#include <cmath>
#include <cstring>
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

// ---------------------------------------------------------------------------------------------------------------
// Frame programs of the fused chain (chain_prog.h): simulate `calls` ticks of one stream exactly as k_phase_chain does
// (generator -> tail of the pending packet -> part 1 of the next), execute every emitted packet's program the way
// k_chain's consumers do (block map -> segments -> lanes, including the integer floor/fraction split of fast run
// segments) and compare each frame's (buffer offset, fraction) with the plain sequential recurrence.
#include "chain_prog.h"

extern "C" uint64_t skc_check_stream(double ratio, uint32_t chunk, uint32_t F, uint32_t channels, uint32_t calls, uint32_t cap_seg,
                                     uint32_t cap_exp, uint32_t *packets_out, uint32_t *max_seg, uint32_t *max_exp, uint32_t *status_or) {
    const double t = 1.0 / ratio;
    const int32_t end_idx = (int32_t)chunk - 9 - (int32_t)std::ceil(t);
    const uint32_t fb = channels * 4u, head = std::min(32u, chunk);
    ChainProgDims d;
    d.nblk = (F + 31u) / 32u;
    d.map_bytes = skc_map_bytes(d.nblk);
    d.cap_seg = cap_seg;
    d.cap_exp = cap_exp;
    std::vector<uint8_t> rec[2] = {std::vector<uint8_t>(skc_prog_cap(d) + 64), std::vector<uint8_t>(skc_prog_cap(d) + 64)};
    uint32_t n_exp_rec[2] = {0, 0}, n_out_rec[2] = {0, 0};
    bool rec_bad[2] = {false, false};
    std::vector<double> seq_prev, seq_cur;   // the true chain of the previous / current chunk
    uint64_t bad = 0;
    uint32_t packets = 0, ms = 0, me = 0, st_or = 0;
    double L = -4.0;
    uint32_t carry = 0;
    for (uint32_t c = 0; c < calls; ++c) {
        const uint32_t par_new = c & 1u, par_old = par_new ^ 1u;
        // ---- exactly what k_phase_chain does: stream the generator into the tail of the old record + part 1 of the new one
        double idx_end = 0;
        uint32_t np, nr, ovf, ns = 0, ne = 0;
        const bool pending = c >= 1;
        uint32_t kd = pending ? F - std::min(carry, F) : 0u;
        const uint32_t ne_old = n_exp_rec[par_old];
        ChainExp *tailp = reinterpret_cast<ChainExp *>(rec[par_old].data() + skc_exp_off(d)) + ne_old;
        SkcStream sb;
        sb.begin(rec[par_new].data(), d, F, fb, kd, tailp, ne_old, kd, d.cap_exp - std::min(ne_old, d.cap_exp), chunk, head, t);
        uint32_t n_cur = sk_phase_stream(L, t, end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
        bool completes = pending;
        if (n_cur < kd) {
            completes = false;
            sb.begin(rec[par_new].data(), d, F, fb, 0u, tailp, ne_old, 0u, 0u, chunk, head, t);
            n_cur = sk_phase_stream(L, t, end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
        }
        const uint32_t tail_status = completes ? sb.tail_status : 0u;
        const uint32_t stb = sb.finish(n_cur, &ns, &ne);
        if (ovf) st_or |= SKC_ST_OVERFLOW;
        seq_cur.clear();
        for (double idx = L; idx < (double)end_idx;) { idx += t; seq_cur.push_back(idx); }
        if (seq_cur.size() != n_cur) ++bad;
        const uint32_t avail = carry + n_cur;
        const uint32_t new_carry = avail >= F ? avail - F : avail;
        if (completes && carry <= n_out_rec[par_old] && !rec_bad[par_old]) {
            uint8_t *ro = rec[par_old].data();
            const uint32_t st = tail_status;
            st_or |= st;
            if (!st) {
                ++packets;
                me = std::max(me, ne_old + (F - std::min(carry, F)));
                // ---- execute the program like the consumers
                const uint16_t *map = reinterpret_cast<const uint16_t *>(ro);
                const ChainSeg *segs = reinterpret_cast<const ChainSeg *>(ro + skc_seg_off(d));
                std::vector<int> hits(F, 0);
                const uint32_t kd = n_out_rec[par_old] - carry;
                for (uint32_t b = 0; b < d.nblk; ++b) {
                    const uint32_t ent = map[b];
                    for (uint32_t s = ent & 0xFFu; s <= (ent >> 8); ++s) {
                        const ChainSeg sg = segs[s];
                        const uint32_t j0 = sg.jj & 0xFFFFu, len = (sg.jj >> 16) - j0;
                        for (uint32_t lane = 0; lane < 32; ++lane) {
                            if (b * 32u + lane >= F) continue;
                            const uint32_t j = b * 32u + lane, rel = j - j0;
                            if (!(rel < len)) continue;
                            uint32_t off;
                            float frac;
                            if (sg.himask == SKC_KIND_E) {
                                const ChainExp e = *reinterpret_cast<const ChainExp *>(ro + sg.aux + rel * 8u);
                                off = e.aoff; frac = e.frac;
                            } else {
                                const double x = __builtin_fma((double)rel, sg.delta, sg.x0);
                                if (sg.himask != SKC_KIND_SLOW) {
                                    const uint32_t flh = (uint32_t)(sk_d2bits(x) >> 32) & sg.himask;
                                    frac = (float)(x - sk_bits2d((uint64_t)flh << 32));
                                    off = 16u * fb + (flh >> sg.sh) - sg.aux;
                                } else {
                                    int32_t fl;
                                    skc_split(x, &fl, &frac);
                                    off = (uint32_t)(16 + fl) * fb;
                                }
                            }
                            // expectation from the true chain
                            const bool from_cur = j >= carry;
                            const double xt = from_cur ? seq_cur[j - carry] : seq_prev[kd + j];
                            const double fl_t = std::floor(xt);
                            const uint32_t off_t = (uint32_t)((from_cur ? (int)chunk : 0) + 16 + (int)fl_t) * fb;
                            const float frac_t = (float)(xt - fl_t);
                            if (off != off_t || memcmp(&frac, &frac_t, 4) != 0) ++bad;
                            ++hits[j];
                        }
                    }
                }
                for (uint32_t j = 0; j < F; ++j) if (hits[j] != 1) ++bad;
            }
        }
        st_or |= stb;
        rec_bad[par_new] = stb != 0;
        ms = std::max(ms, ns);
        n_exp_rec[par_new] = ne;
        n_out_rec[par_new] = n_cur;
        carry = new_carry;
        seq_prev.swap(seq_cur);
        L = idx_end - (double)chunk;
    }
    *packets_out = packets;
    *max_seg = ms;
    *max_exp = me;
    *status_or = st_or;
    return bad;
}
