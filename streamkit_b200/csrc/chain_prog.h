// chain_prog.h -- the per-packet FRAME PROGRAM of the fused chain (host + device, like phase_runs.h).
//
// The resampler node emits packets of exactly F frames (resampler.rs:420-458). A packet of the fused chain consists of
// `carry` frames produced by the stream's PREVIOUS chunk followed by F - carry frames of the current one. The phase
// kernel (one thread per stream) turns rubato's phase recurrence (phase_runs.h) into a list of SEGMENTS over the
// packet's frame index j in [0, F) that the mixing kernel can execute without searching anything:
//
//   run segment       x_j = fma((double)(j - j0), delta, x0)         exact (all values in one binade, see phase_runs.h);
//                     FAST: the binade is [2^e, 2^(e+1)) with 0 <= e <= 16, so floor(x) and x - floor(x) come from
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

#ifndef SKC_TAB_PREFIX
#define SKC_TAB_PREFIX 24u        // prefix capacity handed to the phase generator (host and device must use the same
#endif                            // value: it decides where the generator switches from prefix elements to run entries)
#define SKC_MIN_RUN 16u           // shorter run pieces are stored explicitly (one pass of the consumer costs ~25 instructions)
#define SKC_ST_OVERFLOW 2u        // status bit1: a table of the record overflowed
#define SKC_ST_UNSUPPORTED 4u     // status bit2: the packet needs frames the kernel does not stage
#define SKC_KIND_E 0u             // ChainSeg.himask: explicit segment
#define SKC_KIND_SLOW 1u          // ChainSeg.himask: run segment outside [1, 2^17): floor / fraction by real conversions
                                  // any other value: FAST run segment, the mask that clears the fraction bits of the high word

// One segment = two 16-byte shared-memory loads in the consumer. FAST run segments lie in one binade [2^e, 2^(e+1)),
// 0 <= e <= 16, so with hi = high word of x:  floor(x) as a double = {hi & himask, 0}  and the byte offset of buffer
// frame floor(x) is ((hi & himask) >> sh) - cs   (sh = 20 - e - log2(frame_bytes), cs = (1022 + e) << (20 - sh)).
struct alignas(32) ChainSeg {
    double x0, delta;   // run: x_j = fma((double)(j - j0), delta, x0)
    uint32_t jj;        // j0 | j1 << 16
    uint32_t himask;    // SKC_KIND_E, SKC_KIND_SLOW, or the FAST mask
    uint32_t aux;       // E: byte offset of its first ChainExp from the record start; FAST: cs
    uint32_t sh;        // FAST: shift
};
struct alignas(8) ChainExp { uint32_t aoff; float frac; };  // byte offset of frame y0 from the start of the 16-frame history
struct alignas(16) ChainExp2 { ChainExp a, b; };  // the builder stores explicit entries in aligned pairs: one thread per stream
                                                   // writes its own record, so every store is a scattered request and wide
                                                   // stores halve their number (k_phase_chain is bound by them)

SK_HD uint32_t skc_map_bytes(uint32_t nblk) { return (nblk * 2u + 31u) & ~31u; }   // segments stay 32-byte aligned
struct ChainProgDims {
    uint32_t nblk;       // ceil(F / 32)
    uint32_t map_bytes;  // skc_map_bytes(nblk)
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

SK_HD void skc_store32(uint8_t *dst, const uint64_t (&w)[4]) {   // one 256-bit store on the device (STG.256, sm_100)
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "l"(w[0]), "l"(w[1]), "l"(w[2]), "l"(w[3]) : "memory");
#else
    uint64_t *d = reinterpret_cast<uint64_t *>(dst);
    d[0] = w[0]; d[1] = w[1]; d[2] = w[2]; d[3] = w[3];
#endif
}

struct SkcBuilder {   // appends segments in increasing j and completes the block map on the fly
    uint16_t *map;
    ChainSeg *segs;
    ChainExp *exps;
    ChainProgDims d;
    uint32_t F, frame_bytes;
    uint32_t n_seg, n_exp, bcur, first_cur, status;
    uint64_t map_w[4];       // sixteen map entries are collected and stored as one 32-byte word
    ChainExp exp_pend;       // first entry of an aligned pair of explicit entries (stored with the second one)
    // the open explicit segment (consecutive explicit frames share one segment)
    uint32_t e_j0, e_first;
    bool e_open;
};

SK_HD void skc_map_put(SkcBuilder &b, uint32_t s) {   // entry of block b.bcur: (first_cur, s); advances bcur
    const uint64_t ent = (uint64_t)(b.first_cur | (s << 8)) << (16u * (b.bcur & 3u));
    const uint32_t w = (b.bcur >> 2) & 3u;
    if (w == 0u) b.map_w[0] |= ent; else if (w == 1u) b.map_w[1] |= ent; else if (w == 2u) b.map_w[2] |= ent; else b.map_w[3] |= ent;
    if ((b.bcur & 15u) == 15u) {
        skc_store32(reinterpret_cast<uint8_t *>(b.map) + (b.bcur >> 4) * 32u, b.map_w);
        b.map_w[0] = b.map_w[1] = b.map_w[2] = b.map_w[3] = 0;
    }
    ++b.bcur;
}
SK_HD void skc_append(SkcBuilder &b, uint32_t j0, uint32_t j1, double x0, double delta, uint32_t himask, uint32_t aux, uint32_t sh) {
    if (b.n_seg >= b.d.cap_seg || b.n_seg >= 255u) { b.status |= SKC_ST_OVERFLOW; return; }
    const uint32_t s = b.n_seg++;
#if defined(__CUDA_ARCH__)
    // one 256-bit store (STG.256, sm_100): k_phase_chain is bound by the number of scattered store requests
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(b.segs + s), "r"(__double2loint(x0)), "r"(__double2hiint(x0)),
                 "r"(__double2loint(delta)), "r"(__double2hiint(delta)), "r"(j0 | (j1 << 16)), "r"(himask), "r"(aux), "r"(sh) : "memory");
#else
    ChainSeg sg;
    sg.x0 = x0; sg.delta = delta; sg.jj = j0 | (j1 << 16); sg.himask = himask; sg.aux = aux; sg.sh = sh;
    b.segs[s] = sg;
#endif
    while (b.bcur < b.d.nblk) {
        const uint32_t bs = 32u * b.bcur;
        const uint32_t be = (bs + 31u < b.F - 1u) ? bs + 31u : b.F - 1u;
        if (bs >= j0 && bs < j1) b.first_cur = s;
        if (be >= j1) break;
        skc_map_put(b, s);
    }
}
SK_HD void skc_flush_map(SkcBuilder &b) {
    if (b.bcur & 15u) skc_store32(reinterpret_cast<uint8_t *>(b.map) + (b.bcur >> 4) * 32u, b.map_w);
    if (b.n_exp & 1u) b.exps[b.n_exp - 1u] = b.exp_pend;   // an unpaired last explicit entry
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
    if (b.n_exp & 1u) {
        ChainExp2 pr; pr.a = b.exp_pend; pr.b = e;
        *reinterpret_cast<ChainExp2 *>(b.exps + (b.n_exp - 1u)) = pr;
    } else {
        b.exp_pend = e;
    }
    ++b.n_exp;
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming builder: a sink of sk_phase_stream (phase_runs.h). While the generator walks chunk n it receives the chunk's
// outputs k = 0, 1, .. in order (prefix elements one by one, run entries as soon as they are complete) and writes
//   PART 2 of the pending packet  its tail, frames j in [carry, F) = outputs k < n_tail of this chunk, explicit entries
//                                 appended to the OLD record (buffer = [.. previous chunk | head_frames of this chunk]:
//                                 position p = floor + 16 of this chunk is buffer frame n_frames_prev + p);
//   PART 1 of the next packet     outputs k >= kd of this chunk -> packet frames j = k - kd of the NEW record, then the
//                                 (still empty) explicit tail segment that the next chunk will fill.
// Both boundaries are F - carry when this chunk completes the pending packet (kd == n_tail); the caller re-runs the
// generator with kd = n_tail = 0 in the rare case that it does not (right after a stream starts).
struct SkcStream {
    SkcBuilder b;
    ChainExp *tail;          // old record: where the tail's explicit entries go
    uint32_t n_tail, tail_cap, n_frames_prev, head_frames;
    uint32_t tail_status;    // problems of the PENDING packet's tail (the builder's own status concerns the new record)
    uint32_t tail_odd;       // parity of the absolute index of tail[0]: entries are stored in 16-byte aligned pairs
    ChainExp tail_pend;
    double t;
    uint32_t kd;
    uint32_t k_next;         // outputs below k_next have been consumed
    bool have_last;
    SkRun last;              // the previous run entry (a gap element may follow it)

    SK_HD_MEMBER void begin(uint8_t *rec_new, const ChainProgDims &d, uint32_t F, uint32_t frame_bytes, uint32_t kd_, ChainExp *tail_, uint32_t tail_index0,
                            uint32_t n_tail_, uint32_t tail_cap_, uint32_t n_frames_prev_, uint32_t head_frames_, double t_) {
        tail_odd = tail_index0 & 1u;
        tail_pend.aoff = 0; tail_pend.frac = 0.0f;
        b.map = reinterpret_cast<uint16_t *>(rec_new);
        b.segs = reinterpret_cast<ChainSeg *>(rec_new + skc_seg_off(d));
        b.exps = reinterpret_cast<ChainExp *>(rec_new + skc_exp_off(d));
        b.d = d; b.F = F; b.frame_bytes = frame_bytes;
        b.n_seg = 0; b.n_exp = 0; b.bcur = 0; b.first_cur = 0; b.status = 0; b.map_w[0] = b.map_w[1] = b.map_w[2] = b.map_w[3] = 0;
        b.exp_pend.aoff = 0; b.exp_pend.frac = 0.0f;
        b.e_j0 = 0; b.e_first = 0; b.e_open = false;
        tail = tail_; n_tail = n_tail_; tail_cap = tail_cap_; n_frames_prev = n_frames_prev_; head_frames = head_frames_;
        t = t_; kd = kd_; k_next = 0; have_last = false;
        last.x_a = 0.0; last.delta = 0.0; last.k_a = 0; last.k_e = 0;
        tail_status = n_tail > tail_cap ? SKC_ST_UNSUPPORTED : 0u;
    }
    // one output whose value is known explicitly
    SK_HD_MEMBER void element(uint32_t k, double x) {
        int32_t fl;
        float frac;
        skc_split(x, &fl, &frac);
        const uint32_t p = (uint32_t)(fl + 16);
        if (k < n_tail && k < tail_cap) {
            if (p + 1u >= 16u + head_frames) tail_status |= SKC_ST_UNSUPPORTED;   // needs frames of this chunk the kernel does not stage
            ChainExp e; e.aoff = (n_frames_prev + p) * b.frame_bytes; e.frac = frac;
            if (((k + tail_odd) & 1u) && k > 0u) {          // second entry of an aligned pair
                ChainExp2 pr; pr.a = tail_pend; pr.b = e;
                *reinterpret_cast<ChainExp2 *>(tail + (k - 1u)) = pr;
            } else if (k + 1u == n_tail || ((k + tail_odd) & 1u)) {
                tail[k] = e;                                // last entry, or an unaligned first one
            } else {
                tail_pend = e;
            }
        }
        if (k >= kd && k - kd < b.F) skc_push_exp(b, k - kd, p, frac);
    }
    // members kfirst .. ke-1 of a run anchored at k0: x = x0 + (k - k0) * delta (exact)
    SK_HD_MEMBER void members(uint32_t kfirst, uint32_t ke, uint32_t k0, double x0, double delta) {
        uint32_t k = kfirst;
        const uint32_t lo_end = ke < kd ? ke : kd;   // part below kd: only the tail may want it
        for (; k < lo_end; ++k) {
            if (k >= n_tail) { k = lo_end; break; }
            element(k, sk_dfma((double)(k - k0), delta, x0));
        }
        if (k >= ke) return;
        uint32_t ke_c = ke;                           // clip to the packet
        if (ke_c - kd > b.F) ke_c = kd + b.F;
        if (k >= ke_c) return;
        if (delta > 0.0 && ke_c - k >= SKC_MIN_RUN && k >= n_tail) {
            skc_close_exp(b, k - kd);
            const double xs = sk_dfma((double)(k - k0), delta, x0);
            const uint32_t hi = (uint32_t)(sk_d2bits(xs) >> 32);
            uint32_t himask = SKC_KIND_SLOW, cs = 0, sh = 0;
            if ((hi - 0x3FF00000u) < (17u << 20)) {   // positive, exponent e in 0..16 (shift >= 1), shared by the whole run
                const uint32_t e = (hi >> 20) - 1023u;
                const uint32_t lfb = b.frame_bytes == 8u ? 3u : 2u;
                himask = 0xFFFFFFFFu << (20u - e);
                sh = 20u - e - lfb;
                cs = (1022u + e) << (e + lfb);
            }
            skc_append(b, k - kd, ke_c - kd, xs, delta, himask, cs, sh);
        } else {
            for (; k < ke_c; ++k) element(k, sk_dfma((double)(k - k0), delta, x0));   // short piece: explicit frames
        }
    }
    SK_HD_MEMBER double gap_value() const { return sk_dadd(sk_dfma((double)(last.k_e - 1u - last.k_a), last.delta, last.x_a), t); }
    SK_HD_MEMBER void prefix(uint32_t k, double x) {
        element(k, x);
        k_next = k + 1u;
    }
    SK_HD_MEMBER void run(uint32_t, const SkRun &rn) {
        uint32_t k0 = rn.k_a;
        double x0 = rn.x_a;
        if (have_last && k_next < rn.k_a) {
            // the one uncovered element after the previous run: a true addition. It usually lies on THIS run's lattice
            // (x_a - delta exactly, same binade): then it simply becomes the run's first member.
            const double xg = gap_value();
            if (rn.k_a == k_next + 1u && rn.delta > 0.0 && (sk_d2bits(xg) >> 52) == (sk_d2bits(rn.x_a) >> 52) && sk_dfma(-1.0, rn.delta, rn.x_a) == xg) {
                k0 = k_next;
                x0 = xg;
            } else {
                element(k_next, xg);
            }
        }
        // the generator may re-issue the last prefix element as the anchor of the first run: it has been consumed already
        members(k0 < k_next ? k_next : k0, rn.k_e, k0, x0, rn.delta);
        last = rn;
        have_last = true;
        k_next = rn.k_e;
    }
    // after the generator: n_out outputs in total. Returns status bits; sizes of the new record in *n_seg_out / *n_exp_out
    // (n_exp WITHOUT the tail the next chunk will append).
    SK_HD_MEMBER uint32_t finish(uint32_t n_out, uint32_t *n_seg_out, uint32_t *n_exp_out) {
        if (have_last && k_next < n_out) element(k_next, gap_value());   // trailing gap element
        const uint32_t c = n_out > kd ? (n_out - kd < b.F ? n_out - kd : b.F) : 0u;   // frames carried into the next packet
        *n_exp_out = b.n_exp;
        if (c < b.F) {   // the tail: frames produced by the next chunk; its entries follow the explicit entries of part 1
            if (b.e_open) {
                b.e_open = false;   // the tail simply extends an open explicit segment
                skc_append(b, b.e_j0, b.F, 0.0, 0.0, SKC_KIND_E, skc_exp_off(b.d) + b.e_first * 8u, 0u);
            } else {
                skc_append(b, c, b.F, 0.0, 0.0, SKC_KIND_E, skc_exp_off(b.d) + b.n_exp * 8u, 0u);
            }
            if (b.n_exp + (b.F - c) > b.d.cap_exp) b.status |= SKC_ST_OVERFLOW;
        } else {
            skc_close_exp(b, c);
        }
        skc_flush_map(b);
        *n_seg_out = b.n_seg;
        return b.status;
    }
    // variant for a record WITHOUT a tail (plain resample op: the record covers exactly the n_out outputs of one chunk,
    // begin() was called with kd = n_tail = 0 and F = the record's frame capacity)
    SK_HD_MEMBER uint32_t finish_open(uint32_t n_out, uint32_t *n_seg_out, uint32_t *n_exp_out) {
        if (have_last && k_next < n_out) element(k_next, gap_value());
        const uint32_t c = n_out < b.F ? n_out : b.F;
        if (n_out > b.F) b.status |= SKC_ST_OVERFLOW;
        skc_close_exp(b, c);
        if (b.n_seg > 0u) {   // the block the last segment left open ends with the record
            const uint32_t s = b.n_seg - 1u;
            while (b.bcur < b.d.nblk && b.bcur * 32u < c) {
                skc_map_put(b, s);
                b.first_cur = s;
            }
        }
        skc_flush_map(b);
        *n_seg_out = b.n_seg;
        *n_exp_out = b.n_exp;
        return b.status;
    }
};
