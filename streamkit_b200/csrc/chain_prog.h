// chain_prog.h -- the per-packet FRAME PROGRAM of the fused chain (host + device, like phase_runs.h).
//
// The resampler node emits packets of exactly F frames (resampler.rs:420-458). A packet of the fused chain consists of
// `carry` frames produced by the stream's PREVIOUS chunk followed by F - carry frames of the current one. The phase
// kernel (one thread per stream) turns rubato's phase recurrence (phase_runs.h) into a list of SEGMENTS over the
// packet's frame index j in [0, F) that the mixing kernel can execute without searching anything:
//
//   run segment       x_j = fma((double)(j - j0), delta, x0)         exact (all values in one binade, see phase_runs.h);
//                     FAST: the binade is [2^e, 2^(e+1)) with 0 <= e <= 17, so floor(x) and x - floor(x) come from
//                     integer operations on the high word of the double (mask, shift and offset precomputed)
//   explicit segment  (buffer offset, f32 fraction) stored per frame: prefix / gap / tiny-run elements and the packet's
//                     tail, whose frames are produced by the CURRENT chunk (appended one tick later)
//   block map         for every 32-frame block of the packet: first and last segment touching it
//
// Record layout (one per stream and chunk parity, in HBM; copied to shared memory by ONE bulk copy of the used bytes):
//   [ map: nblk x u16, padded to 16 B | segs: cap_seg x ChainSeg (32 B) | exps: ChainExp (8 B) ... ]
#pragma once
#include <stdint.h>

#include "phase_runs.h"

#define SKC_TAB_PREFIX 24u        // capacities of the thread-private phase table the program is built from (host and
#define SKC_TAB_RUNS 24u          // device must use the same ones: they shape the table, hence the program)
#define SKC_MIN_RUN 8u            // shorter run pieces are stored explicitly (one pass of the consumer costs ~25 instructions)
#define SKC_ST_OVERFLOW 2u        // status bit1: a table of the record overflowed
#define SKC_ST_UNSUPPORTED 4u     // status bit2: the packet needs frames the kernel does not stage
#define SKC_KIND_E 0u             // ChainSeg.himask: explicit segment
#define SKC_KIND_SLOW 1u          // ChainSeg.himask: run segment outside [1, 2^18): floor / fraction by real conversions
                                  // any other value: FAST run segment, the mask that clears the fraction bits of the high word

// One segment = two 16-byte shared-memory loads in the consumer. FAST run segments lie in one binade [2^e, 2^(e+1)),
// 0 <= e <= 17, so with hi = high word of x:  floor(x) as a double = {hi & himask, 0}  and the byte offset of buffer
// frame floor(x) is ((hi & himask) >> sh) - cs   (sh = 20 - e - log2(frame_bytes), cs = (1022 + e) << (20 - sh)).
struct ChainSeg {
    double x0, delta;   // run: x_j = fma((double)(j - j0), delta, x0)
    uint32_t jj;        // j0 | j1 << 16
    uint32_t himask;    // SKC_KIND_E, SKC_KIND_SLOW, or the FAST mask
    uint32_t aux;       // E: byte offset of its first ChainExp from the record start; FAST: cs
    uint32_t sh;        // FAST: shift
};
struct ChainExp { uint32_t aoff; float frac; };  // byte offset of frame y0 from the start of the 16-frame history

struct ChainProgDims {
    uint32_t nblk;       // ceil(F / 32)
    uint32_t map_bytes;  // nblk * 2 rounded up to 16
    uint32_t cap_seg;
    uint32_t cap_exp;    // even
};
SK_HD uint32_t skc_seg_off(const ChainProgDims &d) { return d.map_bytes; }
SK_HD uint32_t skc_exp_off(const ChainProgDims &d) { return d.map_bytes + d.cap_seg * 32u; }
SK_HD uint32_t skc_prog_cap(const ChainProgDims &d) { return d.map_bytes + d.cap_seg * 32u + d.cap_exp * 8u; }

// T::coerce(idx - idx.floor()) and floor(idx) of rubato's interpolation loop
SK_HD void skc_split(double x, int32_t *fl, float *frac) {
#if defined(__CUDA_ARCH__)
    const int f = __double2int_rd(x);
    *fl = f;
    *frac = __double2float_rn(__dsub_rn(x, (double)f));
#else
    const double f = __builtin_floor(x);
    *fl = (int32_t)f;
    *frac = (float)(x - f);
#endif
}

struct SkcBuilder {   // appends segments in increasing j and completes the block map on the fly
    uint16_t *map;
    ChainSeg *segs;
    ChainExp *exps;
    ChainProgDims d;
    uint32_t F, frame_bytes;
    uint32_t n_seg, n_exp, bcur, first_cur, status;
    uint64_t map_acc;       // four map entries are collected and stored as one 8-byte word
    // the open explicit segment (consecutive explicit frames share one segment)
    uint32_t e_j0, e_first;
    bool e_open;
};

SK_HD void skc_append(SkcBuilder &b, uint32_t j0, uint32_t j1, double x0, double delta, uint32_t himask, uint32_t aux, uint32_t sh) {
    if (b.n_seg >= b.d.cap_seg || b.n_seg >= 255u) { b.status |= SKC_ST_OVERFLOW; return; }
    const uint32_t s = b.n_seg++;
    ChainSeg sg;
    sg.x0 = x0; sg.delta = delta; sg.jj = j0 | (j1 << 16); sg.himask = himask; sg.aux = aux; sg.sh = sh;
    b.segs[s] = sg;
    while (b.bcur < b.d.nblk) {
        const uint32_t bs = 32u * b.bcur;
        const uint32_t be = (bs + 31u < b.F - 1u) ? bs + 31u : b.F - 1u;
        if (bs >= j0 && bs < j1) b.first_cur = s;
        if (be >= j1) break;
        b.map_acc |= (uint64_t)(b.first_cur | (s << 8)) << (16u * (b.bcur & 3u));
        if ((b.bcur & 3u) == 3u) {
            reinterpret_cast<uint64_t *>(b.map)[b.bcur >> 2] = b.map_acc;
            b.map_acc = 0;
        }
        ++b.bcur;
    }
}
SK_HD void skc_flush_map(SkcBuilder &b) {
    if (b.bcur & 3u) reinterpret_cast<uint64_t *>(b.map)[b.bcur >> 2] = b.map_acc;
}
SK_HD void skc_close_exp(SkcBuilder &b, uint32_t j_end) {
    if (!b.e_open) return;
    b.e_open = false;
    skc_append(b, b.e_j0, j_end, 0.0, 0.0, SKC_KIND_E, skc_exp_off(b.d) + b.e_first * 8u, 0u);
}
// one explicit frame at packet index j reading buffer frame a_idx (index into [16 history | previous chunk | head of current])
SK_HD void skc_push_exp(SkcBuilder &b, uint32_t j, uint32_t a_idx, float frac) {
    if (b.n_exp >= b.d.cap_exp) { b.status |= SKC_ST_OVERFLOW; return; }
    if (!b.e_open) { b.e_open = true; b.e_j0 = j; b.e_first = b.n_exp; }
    ChainExp e; e.aoff = a_idx * b.frame_bytes; e.frac = frac;
    b.exps[b.n_exp++] = e;
}

// PART 1 (written when chunk n is processed, consumed one tick later): the `carry` frames chunk n contributes to the
// NEXT packet, i.e. chunk outputs k in [n_out - carry, n_out) -> packet frames j = k - (n_out - carry); then the
// (still empty) explicit tail segment [carry, F) whose entries skc_fill_tail() appends when chunk n + 1 is known.
// Returns status bits; *n_seg_out / *n_exp_out receive the table sizes (n_exp WITHOUT the tail).
SK_HD uint32_t skc_build(const double *prefix, uint32_t np, const SkRun *runs, uint32_t nr, double t, uint32_t n_out, uint32_t carry,
                         uint32_t F, uint32_t frame_bytes, const ChainProgDims &d, uint8_t *rec, uint32_t *n_seg_out, uint32_t *n_exp_out) {
    SkcBuilder b;
    b.map = reinterpret_cast<uint16_t *>(rec);
    b.segs = reinterpret_cast<ChainSeg *>(rec + skc_seg_off(d));
    b.exps = reinterpret_cast<ChainExp *>(rec + skc_exp_off(d));
    b.d = d; b.F = F; b.frame_bytes = frame_bytes;
    b.n_seg = 0; b.n_exp = 0; b.bcur = 0; b.first_cur = 0; b.status = 0; b.map_acc = 0;
    b.e_j0 = 0; b.e_first = 0; b.e_open = false;
    const uint32_t c = carry < F ? carry : F;          // frames beyond F belong to a later packet (backlog, reported by the caller)
    const uint32_t kd = n_out - (carry < n_out ? carry : n_out);
    uint32_t k = kd, r = 0;
    const uint32_t k_end = kd + c;
    while (k < k_end) {
        const uint32_t j = k - kd;
        int32_t fl;
        float frac;
        if (k < np || nr == 0u) {
            skc_split(k < np ? prefix[k] : 0.0, &fl, &frac);
            skc_push_exp(b, j, (uint32_t)(fl + 16), frac);
            ++k;
            continue;
        }
        while (r + 1u < nr && runs[r + 1u].k_a <= k) ++r;
        const SkRun rn = runs[r];
        double x0, delta = rn.delta;
        uint32_t ke;
        if (k < rn.k_e) {
            x0 = sk_dfma((double)(k - rn.k_a), rn.delta, rn.x_a);
            ke = rn.k_e;
        } else {
            // gap element: one true addition after the run's last member. It usually lies on the NEXT run's lattice
            // (x_a' - delta' exactly, same binade): then it becomes the first member of that run's segment.
            x0 = sk_dadd(sk_dfma((double)(rn.k_e - 1u - rn.k_a), rn.delta, rn.x_a), t);
            ke = k + 1u;
            delta = 0.0;
            if (r + 1u < nr) {
                const SkRun nx = runs[r + 1u];
                if (nx.k_a == k + 1u && nx.delta > 0.0 && (sk_d2bits(x0) >> 52) == (sk_d2bits(nx.x_a) >> 52) &&
                    sk_dfma(-1.0, nx.delta, nx.x_a) == x0) {
                    delta = nx.delta;
                    ke = nx.k_e;
                }
            }
        }
        if (ke > k_end) ke = k_end;
        if (delta > 0.0 && ke - k >= SKC_MIN_RUN) {
            skc_close_exp(b, j);
            const uint32_t hi = (uint32_t)(sk_d2bits(x0) >> 32);
            uint32_t himask = SKC_KIND_SLOW, cs = 0, sh = 0;
            if ((hi - 0x3FF00000u) < (18u << 20)) {   // positive, exponent e in 0..17, shared by the whole run
                const uint32_t e = (hi >> 20) - 1023u;
                const uint32_t lfb = frame_bytes == 8u ? 3u : 2u;
                himask = 0xFFFFFFFFu << (20u - e);
                sh = 20u - e - lfb;
                cs = (1022u + e) << (e + lfb);
            }
            skc_append(b, j, ke - kd, x0, delta, himask, cs, sh);
            k = ke;
        } else {
            const uint32_t k0 = k;   // short piece: explicit frames (x0 + i * delta is exact inside a run; delta == 0 for a lone element)
            for (; k < ke; ++k) {
                skc_split(sk_dfma((double)(k - k0), delta, x0), &fl, &frac);
                skc_push_exp(b, k - kd, (uint32_t)(fl + 16), frac);
            }
        }
    }
    *n_exp_out = b.n_exp;
    if (c < F) {   // the tail: frames produced by the next chunk; entries follow the explicit entries of part 1
        if (b.e_open) {
            // the tail simply extends an open explicit segment
            b.e_open = false;
            skc_append(b, b.e_j0, F, 0.0, 0.0, SKC_KIND_E, skc_exp_off(d) + b.e_first * 8u, 0u);
        } else {
            skc_append(b, c, F, 0.0, 0.0, SKC_KIND_E, skc_exp_off(d) + b.n_exp * 8u, 0u);
        }
        if (b.n_exp + (F - c) > d.cap_exp) b.status |= SKC_ST_OVERFLOW;
    } else {
        skc_close_exp(b, c);
    }
    skc_flush_map(b);
    *n_seg_out = b.n_seg;
    return b.status;
}

// PART 2 (written when chunk n + 1 is processed): the packet's tail, frames j in [carry, F) = outputs 0 .. F - carry - 1 of
// the current chunk, whose buffer is [.. tail of the previous chunk | head_frames of the current chunk]. `n_frames_prev`
// is the previous chunk's length N: buffer frame of position p (= floor + 16) is N + p.
SK_HD uint32_t skc_fill_tail(const double *prefix, uint32_t np, const SkRun *runs, uint32_t nr, double t, uint32_t n_cur, uint32_t carry,
                             uint32_t F, uint32_t n_frames_prev, uint32_t head_frames, uint32_t frame_bytes, ChainExp *tail, uint32_t cap) {
    uint32_t r = 0, status = 0;
    if (carry >= F) return 0;
    const uint32_t n = F - carry;
    if (n > n_cur || n > cap) return SKC_ST_UNSUPPORTED;
    for (uint32_t k = 0; k < n; ++k) {
        const double x = sk_phase_eval(prefix, np, runs, nr, t, k, &r);
        int32_t fl;
        float frac;
        skc_split(x, &fl, &frac);
        const uint32_t p = (uint32_t)(fl + 16);
        if (p + 1u >= 16u + head_frames) { status |= SKC_ST_UNSUPPORTED; break; }
        ChainExp e; e.aoff = (n_frames_prev + p) * frame_bytes; e.frac = frac;
        tail[k] = e;
    }
    return status;
}
