// k_chain.cuh -- the fused hot path: resample -> per-input gain -> ordered mix -> master gain -> clip -> s16
// in ONE pass, no f32 intermediate ever written to HBM (BASELINE config #5, SURVEY 2.3 "K5 chain_fused").
//
// It fuses, per session (mix group):
//   K x audio::resampler{output_frame_size F}   resampler.rs:377-470 (+ rubato process)      -> rubato interp_lin
//   K x audio::gain                             gain.rs:187-189                              -> one f32 multiply
//   audio::mixer                                mixer.rs:960-980, :1027-1078                 -> ordered f32 sum
//   audio::gain (master) + f32 -> s16           gain.rs:187-189, SURVEY A5                   -> multiply, cvt.rni.sat
//
// The resampler node re-frames its output into packets of exactly F frames (resampler.rs:420-458); a packet
// therefore straddles two input chunks: the `carry` frames the previous chunk produced but that did not fill
// a packet, plus the first F - carry frames of the current chunk (steady state 44.1k->48k: 954 + 6). Instead
// of parking those carry frames in an HBM ring (write + read of 7.6 KB per stream-tick), the kernel RECOMPUTES
// them from the previous tick's input chunk, which is still resident in the other half of the double-banked
// input arena. HBM traffic per stream-tick is one pass over one input chunk plus a ~0.7 KB frame program
// (chain_prog.h) and a 16-frame history.
//
// Two kernels per tick:
//   k_phase_chain  one THREAD per input stream (massively parallel, data independent): rubato's f64 phase recurrence
//                  (phase_runs.h) -> the packet's frame program (chain_prog.h): run segments, explicit frames and a
//                  per-32-frame block map; emission / carry / status bookkeeping (resampler.rs:425-428 re-framing);
//                  a ready-made 64-byte ChainRec per input for k_chain's producer.
//   k_chain        PERSISTENT, WARP-SPECIALISED CTAs (looping over sessions):
//       warp 0  (producer)   prefetches ChainRecs with cp.async, derives the summation order (base selection +
//                            swap_remove, mixer.rs:960-980) from warp ballots, and issues 3 TMA bulk copies per
//                            input ([frame program | history], previous chunk, head of the current chunk) into a
//                            2-stage shared-memory ring guarded by full/empty mbarriers;
//       warps 1-4 (consumers) execute the frame programs: a warp owns EIGHT consecutive 32-frame blocks in which lane l
//                            owns frame (block start + l) -- neighbouring lanes read neighbouring input frames, so the
//                            two shared-memory loads of a frame are conflict free. A block inside the FAST run the warp
//                            is already in (RunCache) is straight-line code; blocks holding a run boundary or explicit
//                            frames walk the segments per lane. Inputs are added SEQUENTIALLY in the reference's order
//                            (f32 addition is not associative); master gain, clip + s16 pack, one coalesced
//                            store per block.
// Input kinds: resampled f32 (the reference's path), rate-equal BYPASS (resampler.rs:299-373: the staged
// chunk is the packet), s16 ingest (expanded to f32 in shared memory); the kernel is instantiated per mixture of kinds
// (CHAIN_PLAIN ... CHAIN_ANY below) because the consumer loop is sensitive to its instruction footprint.
#pragma once
#include "chain_prog.h"
#include "common.cuh"

namespace skgpu {

constexpr int CH_CWARPS = 4;                      // consumer warps
constexpr int CH_NB = 8;                          // consecutive 32-frame blocks a consumer warp owns per iteration (256 frames)
constexpr int CH_CONSUMERS = CH_CWARPS * 32;
constexpr int CH_THREADS = CH_CONSUMERS + 32;     // + producer warp (warp 0)
constexpr int CH_MAX_STAGES = 4;                  // pipeline depth is a launch parameter (2..4)
constexpr int CH_HEAD = 32;                       // frames of the CURRENT chunk staged behind the previous one
constexpr int CH_MAX_INPUTS = 64;                 // inputs per session
constexpr int CH_MAX_KB = 4;                      // inputs staged per batch
constexpr int CH_PF = 8;                          // inputs per session whose records the producer prefetches (more: general path)
constexpr uint32_t CH_PROG_MAX = SK_SIDE_STRIDE - SK_SIDE_HIST;   // side record = [frame program (prog_cap bytes) | 128-byte history field]

constexpr uint32_t CK_BYPASS = 0x100u, CK_S16 = 0x200u;   // ChainCons.sc kind bits
struct __align__(16) ChainCons {   // what a consumer needs per input
    float gain;             // 1.0 when the input has no audio::gain in front of the mixer (x * 1.0 == x)
    uint32_t sc;            // source channels (low byte) | CK_BYPASS: the staged chunk IS the packet (resampler.rs:299-373)
                            //                            | CK_S16: the staged chunk is s16, expand it in place before use
    uint32_t n_prev;        // s16: samples of the first staged piece (previous chunk; bypass: the packet)
    uint32_t n_head;        // s16: samples of the second staged piece (head of the current chunk)
};

// per-input record written by k_phase_chain every tick, consumed by k_chain's producer (64 bytes, one per chain input):
// every address the producer needs is final, so staging an input is three bulk copies and two small stores
constexpr uint32_t CR_EMIT = 1u, CR_UNIQUE = 2u, CR_UNALIGNED = 4u;
struct __align__(16) ChainRec {
    ChainCons cons;
    const uint8_t *prog_src;  // side record of the packet: [frame program (prog_cap bytes) | 128-byte history field]
    const float *prev_g;      // previous chunk (other input bank)
    const float *cur_g;       // current chunk
    float *hist_dst;          // where the last 16 frames of the previous chunk go (history field of the current chunk's record)
    uint32_t chunk_bytes;     // bytes of the first staged piece (s16 streams: padded to whole 16-byte units)
    uint32_t head_bytes;      // staged frames of the current chunk, bytes
    uint32_t flags;           // CR_*
    uint32_t tail_off;        // (N - 16) * channels: float offset of those 16 frames inside the chunk
};
static_assert(sizeof(ChainRec) == 64, "ChainRec is one 64-byte record");

struct __align__(16) ChainTail {   // per staged input: what the end-of-batch history update needs
    float *hist_dst;
    uint32_t tail_off, n;   // n = 16 * channels floats
};

struct __align__(16) ChainStage {   // header of one pipeline stage (shared memory)
    uint32_t nb;            // inputs in this batch           } one 16-byte load
    uint32_t first, last;   // first / last batch of the session
    uint32_t has_base;      // bit0: the first input of the first batch is the base frame (mixer.rs:960-972); bit1: every staged input is stereo;
                            // bits 8-10: block-ownership rotation of this session (load balance, see the consumers)
    uint64_t out_off;
    float master_gain;      // 1.0 when the session has no master audio::gain
    uint32_t flags;
    uint32_t stop;          // no more work
    uint32_t pad[3];
    ChainCons cons[CH_MAX_KB];
    ChainTail tail[CH_MAX_KB];
};

struct ChainDims {          // launch-time geometry of the staging ring (host: chain_size_smem)
    uint32_t kb;            // inputs per batch
    uint32_t chunk_cap;     // bytes reserved per input for [previous chunk | CH_HEAD frames of the current one], 16-aligned
    ChainProgDims prog;     // frame-program capacities
    uint32_t nstages;
    uint32_t max_k;         // largest n_inputs of any session (sizes the producer scratch)
    uint32_t reserved;
    float one;              // 1.0f, deliberately a run-time value (see add2)
};
// shared-memory slot of one staged input:  [frame program | history field 128 B | previous chunk | head of current]
__host__ __device__ __forceinline__ uint32_t chain_in_bytes(const ChainDims &dm) { return skc_prog_cap(dm.prog) + SK_SIDE_HIST + dm.chunk_cap; }

// packed f32x2 arithmetic (Blackwell FMUL2 / FFMA2): two IEEE-rounded results per instruction
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, float b) {
    unsigned long long d, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(bb));
    return d;
}
__device__ __forceinline__ unsigned long long mul2v(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b));
    return d;
}
// packed add. ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (one rounding less than the reference) even
// though both carry .rn, so the add is written as fma(a, 1.0, b) with the 1.0 taken from a kernel parameter: ptxas
// cannot fold a multiplier it does not know, and fl(a * 1 + b) == fl(a + b) for every a, b.
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b, unsigned long long one2) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(one2), "l"(b));
    return d;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

// one frame of one input: rubato interp_lin between the buffer frames at `addr` and the next one, the input's
// audio::gain (gain.rs:187-189), the mixer's channel mapping (mixer.rs:1027-1078); returns the value to ADD
template <int OC, int SC>
__device__ __forceinline__ unsigned long long chain_frame(uint32_t addr, float frac, float gain, unsigned long long one2) {
    if (SC == 2) {
        unsigned long long y0, y1;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(y0) : "r"(addr));
        asm volatile("ld.shared.b64 %0, [%1+8];" : "=l"(y1) : "r"(addr));
        const unsigned long long r = mul2(add2(mul2(y0, __fsub_rn(1.0f, frac)), mul2(y1, frac), one2), gain);
        if (OC == 2) return r;
        float a, b;
        unpack2(r, a, b);
        return pack2(__fmul_rn(__fadd_rn(a, b), 0.5f), 0.0f);   // stereo -> mono: average
    } else {
        const float v = __fmul_rn(interp_lin(frac, lds_f32(addr), lds_f32(addr + 4u)), gain);
        return pack2(v, OC == 2 ? v : 0.0f);                    // mono -> stereo: duplicate
    }
}

// rare run segments outside [1, 2^18) (negative positions right after a stream starts, very long chunks): real conversions.
// Out of line so that the hot loop carries none of its instructions; returns (byte offset from buffer position 16, fraction).
__device__ __forceinline__ unsigned long long chain_split_slow(double x, uint32_t frame_bytes) {
    int32_t fl;
    float frac;
    skc_split(x, &fl, &frac);
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"((uint32_t)(fl * (int32_t)frame_bytes)), "f"(frac));
    return r;
}

// acc += v (packed f32x2, in place; see add2 for the fma-with-one form)
__device__ __forceinline__ void acc_add(unsigned long long &acc, unsigned long long v, unsigned long long one2) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(v), "l"(one2));
}

// one frame of a BYPASS input (input rate == mixer rate): the frame enters the mix as it is -- the input's audio::gain and
// the mixer's channel mapping only; no interpolation arithmetic (resampler.rs:299-373 forwards such packets untouched)
template <int OC, int SC>
__device__ __forceinline__ unsigned long long chain_pass(uint32_t addr, float gain) {
    if (SC == 2) {
        unsigned long long y;
        asm volatile("ld.shared.b64 %0, [%1];" : "=l"(y) : "r"(addr));
        const unsigned long long r = mul2(y, gain);
        if (OC == 2) return r;
        float a, b;
        unpack2(r, a, b);
        return pack2(__fmul_rn(__fadd_rn(a, b), 0.5f), 0.0f);   // stereo -> mono: average
    } else {
        const float v = __fmul_rn(lds_f32(addr), gain);
        return pack2(v, OC == 2 ? v : 0.0f);                    // mono -> stereo: duplicate
    }
}
template <int OC, int SC, int ITERS>
__device__ __forceinline__ void chain_consume_pass(unsigned long long (&acc)[ITERS][CH_NB], uint32_t a_chunk, uint32_t F, uint32_t cw, uint32_t lane,
                                                   float gain, unsigned long long one2) {
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int f = 0; f < CH_NB; ++f) {
            const uint32_t jb = (((uint32_t)it * CH_CWARPS + cw) * CH_NB + (uint32_t)f) * 32u;
            if (jb < F) acc_add(acc[it][f], chain_pass<OC, SC>(a_chunk + min(jb + lane, F - 1u) * (SC * 4u), gain), one2);
        }
    }
}

// s16 ingest (SKGPU_STREAM_S16): the staged pieces are s16 (x = s / 32768, exact); the four consumer warps expand them to f32 in
// place before the input is consumed. All reads of a round happen before its writes (named barrier), rounds run from the top
// down, so the growing f32 image never overwrites s16 samples that are still unread.
__device__ __forceinline__ void chain_expand_s16(uint32_t chunk_sm, uint32_t n_prev, uint32_t n_head, uint32_t ct) {
    const uint32_t off_head = (n_prev * 2u + 15u) & ~15u;      // the head piece starts on the next 16-byte unit
    const uint32_t T = n_prev + n_head;                         // both even
    for (int r = (int)((T + 2047u) / 2048u) - 1; r >= 0; --r) {
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t idx = (((uint32_t)r * CH_CONSUMERS + ct) * 8u + (uint32_t)k) * 2u;
            w[k] = 0u;
            if (idx < T) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[k]) : "r"(chunk_sm + (idx < n_prev ? idx * 2u : off_head + (idx - n_prev) * 2u)));
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CH_CONSUMERS) : "memory");
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t idx = (((uint32_t)r * CH_CONSUMERS + ct) * 8u + (uint32_t)k) * 2u;
            if (idx < T) {
                const float lo = s16_to_f32((int)(short)(w[k] & 0xFFFFu)), hi = s16_to_f32((int)w[k] >> 16);
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(chunk_sm + idx * 4u), "f"(lo), "f"(hi) : "memory");
            }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CH_CONSUMERS) : "memory");
    }
}

// run cache of a consumer warp (registers, warp-uniform except xl): the FAST run segment the warp is inside of
struct RunCache {
    uint32_t j0, j1;      // packet frames the run covers (j1 = 0: nothing cached)
    uint32_t himask;      // clears the fraction bits of the high word
    uint32_t sh;          // byte offset of buffer frame floor(x) = (flh >> sh) - cs
    uint32_t kc;          // shared-memory address of buffer frame 0 of the run's binade offset: a_chunk - cs
    double xl;            // per lane: phase of this lane's frame in the block processed last
    double dl32;          // 32 * delta (exact: a power-of-two multiple)
};

// floor / fraction of a phase inside a FAST run (one binade [2^e, 2^(e+1)), 0 <= e <= 16): integer operations on the high word
__device__ __forceinline__ void fast_split(double x, uint32_t himask, uint32_t &flh, float &frac) {
    flh = (uint32_t)__double2hiint(x) & himask;                                   // floor(x) as a double = {flh, 0}
    frac = __double2float_rn(__dsub_rn(x, __hiloint2double((int)flh, 0)));        // T::coerce(idx - idx.floor())
}

// Executes the frame program of one staged input for the CH_NB consecutive blocks (256 frames) this warp owns per
// iteration. Lane l owns frame (block start + l). A block that lies inside the cached FAST run takes the straight-line
// path: the lane's phase advances by 32 * delta (exact, phase_runs.h) -- no segment decode, no search, no predicate.
// Any other block (a binade boundary inside it, explicit frames, the packet's head and tail: ~7 of 30) takes the general
// path: every lane finds ITS segment (short divergent walk from the block map's first entry) and evaluates it.
//   prog    shared-memory address of the program record; a_hist: of the 16-frame history (buffer position 0)
// N consecutive blocks that lie inside the cached run, as ONE basic block: the N dependency chains (DADD -> split -> LDS ->
// 5 packed operations) interleave, which is what hides their fixed latencies with only ~6 warps per scheduler
template <int OC, int SC, int N>
__device__ __forceinline__ void chain_pure(unsigned long long *acc, RunCache &rc, float gain, unsigned long long one2) {
    double x[N];
    uint32_t flh[N];
    float frac[N];
#pragma unroll
    for (int n = 0; n < N; ++n) x[n] = __dadd_rn(n ? x[n - 1] : rc.xl, rc.dl32);
    rc.xl = x[N - 1];
#pragma unroll
    for (int n = 0; n < N; ++n) fast_split(x[n], rc.himask, flh[n], frac[n]);
#pragma unroll
    for (int n = 0; n < N; ++n) acc_add(acc[n], chain_frame<OC, SC>((flh[n] >> rc.sh) + rc.kc, frac[n], gain, one2), one2);   // a_chunk + floor(x) * frame bytes
}

// one block, any shape: straight line when it lies inside the cached run (possibly after entering the run that covers it),
// else the per-lane segment walk; `refresh`: the next block belongs to this warp too, so try to cache the block's last run
template <int OC, int SC>
__device__ __forceinline__ unsigned long long chain_block(RunCache &rc, uint32_t blk, bool refresh, uint32_t prog, uint32_t segs, uint32_t a_hist,
                                                          uint32_t a_chunk, uint32_t F, uint32_t lane, float gain, unsigned long long one2) {
    const uint32_t jb = blk * 32u;
    bool pure = jb >= rc.j0 && jb + 32u <= rc.j1;
    uint32_t e = 0;
    if (!pure) {
        if (jb >= F) return 0x8000000080000000ull;             // warp-uniform: past the packet, add -0.0 (the identity)
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(prog + blk * 2u));
    }
    if (pure) {
        // ---- straight line: the whole block lies inside the cached run
        rc.xl = __dadd_rn(rc.xl, rc.dl32);
        uint32_t flh;
        float frac;
        fast_split(rc.xl, rc.himask, flh, frac);
        return chain_frame<OC, SC>((flh >> rc.sh) + rc.kc, frac, gain, one2);   // a_chunk + floor(x) * frame bytes
    } else {
        uint32_t addr;
        float frac;
        // ---- general: per-lane segment walk
        const uint32_t j = min(jb + lane, F - 1u);               // lanes past the packet's end recompute its last frame
        uint32_t sa = segs + (e & 0xFFu) * 32u;
        const uint32_t sa_last = segs + (e >> 8) * 32u;
        uint32_t jj, himask, aux, sh;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(jj), "=r"(himask), "=r"(aux), "=r"(sh) : "r"(sa));
        while ((jj >> 16) <= j && sa < sa_last) {
            sa += 32u;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(jj), "=r"(himask), "=r"(aux), "=r"(sh) : "r"(sa));
        }
        const uint32_t rel = j - (jj & 0xFFFFu);
        if (himask == SKC_KIND_E) {
            uint32_t aoff;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(aoff), "=f"(frac) : "r"(prog + aux + rel * 8u));
            addr = a_hist + aoff;
        } else {
            double x0, dl;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(dl) : "r"(sa));
            const double x = __fma_rn((double)rel, dl, x0);     // exact inside a run (phase_runs.h)
            if (himask != SKC_KIND_SLOW) {
                uint32_t flh;
                fast_split(x, himask, flh, frac);
                addr = (flh >> sh) + a_chunk - aux;
            } else {
                const unsigned long long pf = chain_split_slow(x, SC * 4u);
                uint32_t off;
                asm("mov.b64 {%0, %1}, %2;" : "=r"(off), "=f"(frac) : "l"(pf));
                addr = a_chunk + off;
            }
        }
        const unsigned long long v = chain_frame<OC, SC>(addr, frac, gain, one2);
        // ---- refresh the run cache from the block's last segment when that is a FAST run reaching past the block
        rc.j1 = 0;
        if (refresh) {
            uint32_t ljj, lhimask, laux, lsh;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(ljj), "=r"(lhimask), "=r"(laux), "=r"(lsh) : "r"(sa_last));
            if (lhimask > SKC_KIND_SLOW && (ljj >> 16) >= jb + 64u) {   // warp-uniform
                double x0, dl;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(dl) : "r"(sa_last));
                rc.j0 = ljj & 0xFFFFu;
                rc.j1 = ljj >> 16;
                rc.himask = lhimask;
                rc.sh = lsh;
                rc.kc = a_chunk - laux;
                rc.dl32 = __dmul_rn(dl, 32.0);
                // this lane's frame of THIS block on the run's lattice (lanes before the run's start extrapolate below
                // the binade: still a multiple of the binade's unit, so every later + 32 delta step is exact)
                rc.xl = __fma_rn((double)((int)(jb + lane) - (int)rc.j0), dl, x0);
            }
        }
        return v;
    }
}

template <int OC, int SC, int ITERS>
__device__ __forceinline__ void chain_consume(unsigned long long (&acc)[ITERS][CH_NB], uint32_t prog, uint32_t a_hist, const ChainProgDims &pd,
                                              uint32_t F, uint32_t cw, uint32_t lane, float gain, unsigned long long one2) {
    static_assert(CH_NB % 4 == 0, "blocks are processed in aligned groups of four");
    const uint32_t segs = prog + skc_seg_off(pd);
    const uint32_t a_chunk = a_hist + 16u * SC * 4u;   // buffer position 16: floor(idx) == 0
    RunCache rc;
    rc.j0 = 0; rc.j1 = 0; rc.himask = 0; rc.sh = 0; rc.kc = 0; rc.xl = 0.0; rc.dl32 = 0.0;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        rc.j1 = 0;   // nothing cached: the warp's first block of this input / iteration is not adjacent to its last one
        const uint32_t blk0 = ((uint32_t)it * CH_CWARPS + cw) * CH_NB;
        if (blk0 * 32u < F) {
            // enter the run that covers the warp's first block, if one FAST run does (else the block takes the general path)
            uint32_t e;
            asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(prog + blk0 * 2u));
            const uint32_t sa = segs + (e & 0xFFu) * 32u;
            uint32_t ljj, lhimask, laux, lsh;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(ljj), "=r"(lhimask), "=r"(laux), "=r"(lsh) : "r"(sa));
            if ((e & 0xFFu) == (e >> 8) && lhimask > SKC_KIND_SLOW && (ljj >> 16) >= blk0 * 32u + 32u) {
                double x0, dl;
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(dl) : "r"(sa));
                rc.j0 = ljj & 0xFFFFu;
                rc.j1 = ljj >> 16;
                rc.himask = lhimask;
                rc.sh = lsh;
                rc.kc = a_chunk - laux;
                rc.dl32 = __dmul_rn(dl, 32.0);
                // the lane's frame one block EARLIER on the run's lattice (a multiple of the binade's unit below 2^(e+1), hence
                // exact, also when it extrapolates below the run's start); the straight-line paths add 32 delta per block
                rc.xl = __fma_rn((double)((int)(blk0 * 32u + lane) - (int)rc.j0 - 32), dl, x0);
            }
        }
#pragma unroll
        for (int f4 = 0; f4 < CH_NB; f4 += 4) {
            const uint32_t jb4 = (blk0 + (uint32_t)f4) * 32u;
            if (jb4 >= rc.j0 && jb4 + 128u <= rc.j1) {
                chain_pure<OC, SC, 4>(&acc[it][f4], rc, gain, one2);
            } else {
#pragma unroll
                for (int f2 = f4; f2 < f4 + 4; f2 += 2) {
                    const uint32_t jb2 = (blk0 + (uint32_t)f2) * 32u;
                    if (jb2 >= rc.j0 && jb2 + 64u <= rc.j1) {
                        chain_pure<OC, SC, 2>(&acc[it][f2], rc, gain, one2);
                    } else {
                        acc_add(acc[it][f2], chain_block<OC, SC>(rc, blk0 + (uint32_t)f2, true, prog, segs, a_hist, a_chunk, F, lane, gain, one2), one2);
                        acc_add(acc[it][f2 + 1], chain_block<OC, SC>(rc, blk0 + (uint32_t)f2 + 1u, f2 + 2 < CH_NB, prog, segs, a_hist, a_chunk, F, lane, gain, one2), one2);
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------ k_phase_chain
constexpr int PHASE_CHAIN_THREADS = 64;

__global__ void __launch_bounds__(PHASE_CHAIN_THREADS, 14) k_phase_chain(const OpHeader *__restrict__ hdr, const skgpu_chain_input *__restrict__ inputs,
                                                                     const uint8_t *__restrict__ present, const float *__restrict__ gains, SlotTables st,
                                                                     uint8_t *__restrict__ arena, const uint32_t *__restrict__ tick, uint64_t bank_stride,
                                                                     uint32_t F, uint64_t results_off, ChainDims dm, ChainRec *__restrict__ recs) {
    const uint32_t il = blockIdx.x * PHASE_CHAIN_THREADS + threadIdx.x;
    if (il >= hdr->count2) return;
    const uint32_t i = il + hdr->first2;   // a sliced tick launches this kernel once per slice of the tables
    const skgpu_chain_input in = inputs[i];
    const uint32_t slot = in.slot;
    SlotRec *recp = st.rec + slot;
    SlotRec rec = *recp;
    const bool pres = present ? (present[i] != 0) : true;
    const uint32_t ch = rec.channels, N = rec.chunk, fb = ch * 4u;
    const bool s16 = (rec.flags & SLOT_S16) != 0;
    const uint32_t sbytes = s16 ? ch * 2u : fb;          // bytes of one frame in the input arena
    const uint32_t parity = tick[0] & 1u;
    const uint8_t *cur_b = arena + in.in_off + (uint64_t)parity * bank_stride;
    const uint8_t *prev_b = arena + in.in_off + (uint64_t)(1u - parity) * bank_stride;
    const float gain = in.gain_idx != SKGPU_NO_GAIN ? gains[in.gain_idx] : 1.0f;
    uint32_t *dst = reinterpret_cast<uint32_t *>(recs + i);
    if (rec.flags & SLOT_BYPASS) {
        // ---- input rate == mixer rate: no resampler state, no program; this tick's chunk is the packet (resampler.rs:299-373)
        skgpu_chain_result res;
        res.emitted = pres ? 1u : 0u;
        res.status = 0;
        reinterpret_cast<skgpu_chain_result *>(arena + results_off)[i] = res;
        const uint64_t a_prog = (uint64_t)(uintptr_t)slot_side(st, slot, 0), a_cur = (uint64_t)(uintptr_t)cur_b;
        const uint32_t bytes = (N * sbytes + 15u) & ~15u;
        const uint32_t flags = (pres ? CR_EMIT : 0u) | ((in.flags & SKGPU_MIX_IN_UNIQUE) ? CR_UNIQUE : 0u) | (((N * sbytes) & 15u) && !s16 ? CR_UNALIGNED : 0u);
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(__float_as_uint(gain)), "r"(ch | CK_BYPASS | (s16 ? CK_S16 : 0u)),
                     "r"(N * ch), "r"(0u), "r"((uint32_t)a_prog), "r"((uint32_t)(a_prog >> 32)), "r"((uint32_t)a_cur), "r"((uint32_t)(a_cur >> 32)) : "memory");
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8), "r"((uint32_t)a_cur), "r"((uint32_t)(a_cur >> 32)), "r"((uint32_t)a_prog),
                     "r"((uint32_t)(a_prog >> 32)), "r"(s16 ? bytes : N * sbytes), "r"(0u), "r"(flags), "r"(0u) : "memory");
        return;
    }
    const uint32_t head = min((uint32_t)CH_HEAD, N);
    // slot record fields reused by the chain op: n_prefix[par] = explicit entries of the record's part 1, n_runs[par] = segments
    const uint32_t count0 = rec.chunk_count;
    const uint32_t par_new = count0 & 1u, par_old = par_new ^ 1u;   // record of the chunk processed now / of the previous chunk
    const uint32_t carry = rec.carry;
    uint32_t status = 0, emit = 0;
    if (pres) {
        // ---- rubato's phase recurrence for this chunk, streamed into the frame programs (chain_prog.h): the tail of the
        // pending packet (old record) and part 1 of the next packet (new record). A packet is emitted when
        // carry + n_cur >= F (resampler.rs:425-428), i.e. when this chunk has at least kd = F - carry outputs.
        // (static field accesses only: the record stays in registers and is written back with two 256-bit stores)
        const uint32_t n_prev = count0 >= 1u ? (par_old ? rec.n_out[1] : rec.n_out[0]) : 0u;
        if (carry > n_prev) status |= SKC_ST_UNSUPPORTED;                 // carried frames span more than one chunk
        if ((rec.overflow >> par_old) & 1u) status |= SKC_ST_OVERFLOW;    // the record the packet would execute is incomplete
        const bool pending = count0 >= 1u;
        uint32_t kd = pending ? F - min(carry, F) : 0u;
        const uint32_t ne_old = par_old ? rec.n_prefix[1] : rec.n_prefix[0];
        ChainExp *tail = reinterpret_cast<ChainExp *>(slot_side(st, slot, par_old) + skc_exp_off(dm.prog)) + ne_old;
        uint8_t *rec_new = slot_side(st, slot, par_new);
        SkcStream sb;
        uint32_t np, nr, ovf;
        double idx_end;
        sb.begin(rec_new, dm.prog, F, fb, kd, tail, ne_old, kd, dm.prog.cap_exp - min(ne_old, dm.prog.cap_exp), N, head, rec.t_ratio);
        uint32_t n_cur = sk_phase_stream(rec.last_index, rec.t_ratio, rec.end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
        if (n_cur < kd) {
            // the chunk does not complete the packet (only right after a stream starts): everything is carried
            kd = 0u;
            sb.begin(rec_new, dm.prog, F, fb, 0u, tail, ne_old, 0u, 0u, N, head, rec.t_ratio);
            n_cur = sk_phase_stream(rec.last_index, rec.t_ratio, rec.end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
        } else if (pending) {
            status |= sb.tail_status;
            emit = (status & (SKC_ST_OVERFLOW | SKC_ST_UNSUPPORTED)) ? 0u : 1u;
        }
        uint32_t n_seg = 0, n_exp = 0;
        const uint32_t st_new = sb.finish(n_cur, &n_seg, &n_exp);
        const uint32_t avail = carry + n_cur;
        const uint32_t new_carry = (avail >= F) ? avail - F : avail;
        if (new_carry > n_cur) status |= 1u;                              // backlog: a second packet is pending
        const double li = __dsub_rn(idx_end, (double)rec.chunk);           // self.last_index = idx - chunk_size as f64
        const uint32_t n_out0 = par_new ? rec.n_out[0] : n_cur, n_out1 = par_new ? n_cur : rec.n_out[1];
        const uint32_t np01 = par_new ? ((uint32_t)rec.n_prefix[0] | (n_exp << 16)) : ((n_exp & 0xFFFFu) | ((uint32_t)rec.n_prefix[1] << 16));
        const uint32_t nr01 = par_new ? ((uint32_t)rec.n_runs[0] | (n_seg << 16)) : ((n_seg & 0xFFFFu) | ((uint32_t)rec.n_runs[1] << 16));
        const uint32_t ovfl = (rec.overflow & ~(1u << par_new)) | ((st_new ? 1u : 0u) << par_new);
        uint32_t *rw = reinterpret_cast<uint32_t *>(recp);
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(rw), "r"(__double2loint(rec.t_ratio)), "r"(__double2hiint(rec.t_ratio)),
                     "r"(__double2loint(li)), "r"(__double2hiint(li)), "r"(rec.chunk), "r"(rec.channels), "r"(count0 + 1u), "r"(new_carry) : "memory");
        asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(rw + 8), "r"(n_out0), "r"(n_out1), "r"(np01), "r"(nr01),
                     "r"((uint32_t)rec.end_idx), "r"(ovfl), "r"(rec.flags), "r"(0u) : "memory");
    }
    skgpu_chain_result res;
    res.emitted = emit;
    res.status = status;
    reinterpret_cast<skgpu_chain_result *>(arena + results_off)[i] = res;

    const float *cur_g = reinterpret_cast<const float *>(cur_b);
    const float *prev_g = reinterpret_cast<const float *>(prev_b);
    // a present input that emits nothing still retires its previous chunk: the history before the current chunk
    // (= last 16 frames of the previous one) goes into the current chunk's side record. Emitting inputs get it from
    // k_chain's consumers, which have the previous chunk in shared memory anyway.
    const uint32_t prog_cap = skc_prog_cap(dm.prog);
    float *hist_dst = reinterpret_cast<float *>(slot_side(st, slot, par_new) + prog_cap + SK_SIDE_HIST) - 16u * ch;
    if (pres && !emit && count0 >= 1u) {
        if (s16) {
            const short *ps = reinterpret_cast<const short *>(prev_b);
            for (uint32_t e = 0; e < 16u * ch; ++e) hist_dst[e] = s16_to_f32((int)ps[(size_t)(N - 16u) * ch + e]);
        } else {
            for (uint32_t e = 0; e < 16u * ch; ++e) hist_dst[e] = prev_g[(size_t)(N - 16u) * ch + e];
        }
    }
    // the 64-byte ChainRec, written with two 256-bit stores (the kernel is bound by scattered store requests)
    const uint64_t a_prog = (uint64_t)(uintptr_t)slot_side(st, slot, par_old), a_prev = (uint64_t)(uintptr_t)prev_g,
                   a_cur = (uint64_t)(uintptr_t)cur_g, a_hist = (uint64_t)(uintptr_t)hist_dst;
    // staged pieces: previous chunk, head of the current one (s16 streams: copied in whole 16-byte units)
    const uint32_t cb = s16 ? ((N * sbytes + 15u) & ~15u) : N * fb, hb = s16 ? ((head * sbytes + 15u) & ~15u) : head * fb;
    const uint32_t flags = (emit ? CR_EMIT : 0u) | ((in.flags & SKGPU_MIX_IN_UNIQUE) ? CR_UNIQUE : 0u) | (((cb | hb) & 15u) ? CR_UNALIGNED : 0u);
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(__float_as_uint(gain)), "r"(ch | (s16 ? CK_S16 : 0u)), "r"(N * ch), "r"(head * ch),
                 "r"((uint32_t)a_prog), "r"((uint32_t)(a_prog >> 32)), "r"((uint32_t)a_prev), "r"((uint32_t)(a_prev >> 32)) : "memory");
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst + 8), "r"((uint32_t)a_cur), "r"((uint32_t)(a_cur >> 32)), "r"((uint32_t)a_hist),
                 "r"((uint32_t)(a_hist >> 32)), "r"(cb), "r"(hb), "r"(flags), "r"((N - 16u) * ch) : "memory");
}

// OC output channels (1 | 2); ITERS = ceil(F / 1024); KIND: what the op's inputs are (host: validate_chain) --
//   CHAIN_PLAIN   every input is a resampled f32 stream with OC channels: ONE copy of the consumer code, no input-kind code
//                 (the instruction footprint matters);
//   CHAIN_BYPASS  every input is a rate-equal f32 stream with OC channels (Opus decoders into a 48 kHz mix,
//                 samples/pipelines/dynamic/moq_mixing.yml): no frame programs at all, the staged chunk is the packet;
//   CHAIN_ANY     anything the chain admits (mono and stereo, bypass, s16 ingest), in any mixture
//   CHAIN_PLAIN_S16 / CHAIN_BYPASS_S16   the same two with every input arriving as s16 (expanded in shared memory first)
//   CHAIN_F32      f32 inputs of the output's channel count, resampled and rate-equal ones mixed (a 48 kHz mix with 44.1 kHz
//                  and 48 kHz participants): the plain code plus the small pass-through loop, nothing else
constexpr int CHAIN_ANY = 0, CHAIN_PLAIN = 1, CHAIN_BYPASS = 2, CHAIN_PLAIN_S16 = 3, CHAIN_BYPASS_S16 = 4, CHAIN_F32 = 5;
template <int OC, int ITERS, int KIND>
__global__ void __launch_bounds__(CH_THREADS, 5) k_chain(const OpHeader *__restrict__ hdr, const skgpu_chain_group *__restrict__ groups,
                                                         const ChainRec *__restrict__ recs, const float *__restrict__ gains, SlotTables st,
                                                         uint8_t *__restrict__ arena, uint32_t F, ChainDims dm, uint32_t *tick_advance) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[CH_MAX_STAGES], bar_empty[CH_MAX_STAGES];
    __shared__ __align__(16) ChainStage s_stage[CH_MAX_STAGES];
    __shared__ uint8_t s_order[CH_MAX_INPUTS];
    __shared__ float s_one;   // 1.0f in a place ptxas cannot see through (add2): a register value, not a constant-bank operand

    // dynamic smem: nstages x kb input slots, then the producer's scratch for sessions that need several batches
    const uint32_t kb = dm.kb, nstages = dm.nstages;
    const uint32_t prog_cap = skc_prog_cap(dm.prog);
    const uint32_t in_bytes = chain_in_bytes(dm);
    const uint32_t stage_bytes = kb * in_bytes;
    ChainRec *s_res = reinterpret_cast<ChainRec *>(smem_raw + (size_t)stage_bytes * nstages);
    const uint32_t n_groups = hdr->count;
    groups += hdr->first;                  // a sliced tick launches this kernel once per slice of the tables
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    if (threadIdx.x == 0) {
        s_one = dm.one;
        for (uint32_t s = 0; s < nstages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], CH_CONSUMERS / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == 0) {
        // =============================================================== producer warp
        __shared__ __align__(16) skgpu_chain_group pf_grp[4];
        __shared__ __align__(16) ChainRec pf_rec[2][CH_PF];
        const uint32_t gstep = gridDim.x;
        uint32_t stage = 0, ephase = 1;  // waiting on parity 1 of a fresh mbarrier returns immediately
        auto sess = [&](uint32_t n) -> uint32_t { return blockIdx.x + n * gstep; };
        auto issue_grp = [&](uint32_t n) {
            if (lane < 3u && sess(n) < n_groups) cp_async8(reinterpret_cast<uint8_t *>(&pf_grp[n & 3u]) + lane * 8u, reinterpret_cast<const uint8_t *>(&groups[sess(n)]) + lane * 8u);
        };
        auto issue_rec = [&](uint32_t n) {
            if (sess(n) >= n_groups) return;
            const skgpu_chain_group &g = pf_grp[n & 3u];
            if (lane < min(g.n_inputs, (uint32_t)CH_PF)) {
                const uint8_t *src = reinterpret_cast<const uint8_t *>(recs + g.first_input + lane);
                uint8_t *dst = reinterpret_cast<uint8_t *>(&pf_rec[n & 1u][lane]);
#pragma unroll
                for (int w = 0; w < 4; ++w) cp_async16(dst + w * 16, src + w * 16);
            }
        };
        // one input (general path): copy its consumer record into the stage and issue its bulk copies; returns the bytes the barrier must expect
        auto stage_input = [&](const ChainRec &r, uint32_t q, ChainStage *S, uint8_t *sm, uint64_t *bar) -> uint32_t {
            S->cons[q] = r.cons;
            ChainTail tl;
            tl.hist_dst = r.hist_dst; tl.tail_off = r.tail_off; tl.n = (r.cons.sc & CK_BYPASS) ? 0u : 16u * (r.cons.sc & 0xFFu);
            S->tail[q] = tl;
            uint8_t *slot_sm = sm + (size_t)q * in_bytes;
            uint8_t *chunk_sm = slot_sm + prog_cap + SK_SIDE_HIST;
            const uint32_t cb = r.chunk_bytes, hb = r.head_bytes;
            uint32_t bytes = 0;
            if (KIND != CHAIN_BYPASS && KIND != CHAIN_BYPASS_S16) {   // a bypass input has neither a program nor a history
                bytes = prog_cap + SK_SIDE_HIST;
                tma_bulk_g2s(slot_sm, r.prog_src, prog_cap + SK_SIDE_HIST, bar);   // frame program + the 16 frames before the previous chunk
            }
            if (!(r.flags & CR_UNALIGNED)) {
                tma_bulk_g2s(chunk_sm, r.prev_g, cb, bar);
                if (hb) tma_bulk_g2s(chunk_sm + cb, r.cur_g, hb, bar);
                bytes += cb + hb;
            } else {   // chunk size not a multiple of 16 bytes (e.g. mono 882 frames): no bulk copy, this lane copies
                float *dst = reinterpret_cast<float *>(chunk_sm);
                for (uint32_t e = 0; e < cb / 4u; ++e) dst[e] = r.prev_g[e];
                for (uint32_t e = 0; e < hb / 4u; ++e) dst[cb / 4u + e] = r.cur_g[e];
            }
            return bytes;
        };
        auto stage_header = [&](ChainStage *S, const skgpu_chain_group &grp, uint32_t nb, bool first, bool last, uint32_t has_base, uint32_t rot) {
            S->nb = nb;
            S->first = first;
            S->last = last;
            S->has_base = has_base | (rot << 8);
            S->out_off = grp.out_off;
            S->master_gain = grp.gain_idx != SKGPU_NO_GAIN ? gains[grp.gain_idx] : 1.0f;
            S->flags = grp.flags;
            S->stop = 0;
        };
        issue_grp(0); issue_grp(1);
        cp_async_wait_all(); __syncwarp();
        issue_rec(0);
        cp_async_wait_all(); __syncwarp();

        for (uint32_t n = 0; sess(n) < n_groups; ++n) {
            issue_grp(n + 2u);   // prefetch (completes in the background while this session is staged)
            issue_rec(n + 1u);
            const skgpu_chain_group grp = pf_grp[n & 3u];
            const uint32_t K = min(grp.n_inputs, (uint32_t)CH_MAX_INPUTS);
            ChainRec r;
            r.flags = 0;
            r.cons.sc = 0;
            if (lane < min(K, (uint32_t)CH_PF)) r = pf_rec[n & 1u][lane];
            else if (lane < min(K, 32u)) r = recs[grp.first_input + lane];   // big sessions: straight from HBM
            const bool emit = (r.flags & CR_EMIT) != 0;
            const bool elig = emit && (r.cons.sc & 0xFFu) == (uint32_t)OC;   // packet already has the output shape
            // ---- summation order over the inputs that deliver a packet: base selection + swap_remove (mixer.rs:960-980),
            // computed by every lane from three ballots
            const uint32_t emit_mask = __ballot_sync(0xffffffffu, emit);
            const uint32_t elig_mask = __ballot_sync(0xffffffffu, elig);
            const uint32_t uniq_mask = __ballot_sync(0xffffffffu, elig && (r.flags & CR_UNIQUE));
            const uint32_t unal_mask = __ballot_sync(0xffffffffu, emit && (r.flags & CR_UNALIGNED));
            const uint32_t mono_mask = __ballot_sync(0xffffffffu, emit && (r.cons.sc & 0xFFu) != 2u);
            uint32_t m = __popc(emit_mask);
            if (K <= (uint32_t)CH_PF && m <= kb && unal_mask == 0u) {
                // ---- common path. max_by_key((unique, idx)): the last unique full-shape frame, else the last full-shape frame
                const int base_lane = uniq_mask ? 31 - __clz(uniq_mask) : (elig_mask ? 31 - __clz(elig_mask) : -1);
                const uint32_t rank = __popc(emit_mask & ((1u << lane) - 1u));
                uint32_t pos = rank;
                if (base_lane >= 0) {
                    const uint32_t base_rank = __popc(emit_mask & ((1u << base_lane) - 1u));
                    if ((int)lane == base_lane) pos = 0;
                    else pos = 1u + ((rank == m - 1u) ? base_rank : rank);   // Vec::swap_remove: the last element takes the base's slot
                }
                mbar_wait(&bar_empty[stage], ephase);
                ChainStage *S = &s_stage[stage];
                if (emit) {   // the lanes fill the consumer-side records of their inputs in parallel
                    S->cons[pos] = r.cons;
                    ChainTail tl;
                    tl.hist_dst = r.hist_dst; tl.tail_off = r.tail_off; tl.n = (r.cons.sc & CK_BYPASS) ? 0u : 16u * (r.cons.sc & 0xFFu);
                    S->tail[pos] = tl;
                    s_order[pos] = (uint8_t)lane;
                }
                __syncwarp();
                if (lane == 0) {
                    stage_header(S, grp, m, true, true, (base_lane >= 0 ? 1u : 0u) | (mono_mask == 0u ? 2u : 0u), (n * 3u) & (CH_CWARPS - 1u));
                    // one thread issues the 3 bulk copies of every input, in summation order, from the prefetched records
                    uint8_t *dst = smem_raw + (size_t)stage * stage_bytes;
                    uint32_t bytes = 0;
                    for (uint32_t q = 0; q < m; ++q, dst += in_bytes) {
                        const ChainRec *rr = &pf_rec[n & 1u][s_order[q]];
                        const uint32_t cb = rr->chunk_bytes, hb = rr->head_bytes;
                        if (KIND != CHAIN_BYPASS && KIND != CHAIN_BYPASS_S16) {   // a bypass input has neither a program nor a history
                            tma_bulk_g2s(dst, rr->prog_src, prog_cap + SK_SIDE_HIST, &bar_full[stage]);
                            bytes += prog_cap + SK_SIDE_HIST;
                        }
                        tma_bulk_g2s(dst + prog_cap + SK_SIDE_HIST, rr->prev_g, cb, &bar_full[stage]);
                        if (hb) tma_bulk_g2s(dst + prog_cap + SK_SIDE_HIST + cb, rr->cur_g, hb, &bar_full[stage]);
                        bytes += cb + hb;
                    }
                    mbar_expect_tx(&bar_full[stage], bytes);   // arrive: the smem writes above are ordered before it
                }
                if (++stage == nstages) { stage = 0; ephase ^= 1u; }
            } else {
                // ---- general path: more emitting inputs than fit one batch, or more than 32 inputs: serial order, several batches
                if (lane < min(K, 32u)) s_res[lane] = r;
                for (uint32_t j = 32u + lane; j < K; j += 32u) s_res[j] = recs[grp.first_input + j];
                __syncwarp();
                uint32_t has_base = 0;
                m = 0;
                if (lane == 0) {
                    int base = -1, base_unique = -1;
                    for (uint32_t j = 0; j < K; ++j) {
                        if (!(s_res[j].flags & CR_EMIT)) continue;
                        if ((s_res[j].cons.sc & 0xFFu) == (uint32_t)OC) {
                            const int u = (s_res[j].flags & CR_UNIQUE) ? 1 : 0;
                            if (u >= base_unique) { base = (int)m; base_unique = u; }
                        }
                        s_order[m++] = (uint8_t)j;
                    }
                    if (base >= 0 && m > 0) {
                        const uint8_t b = s_order[base];
                        s_order[base] = s_order[m - 1];
                        for (uint32_t q = m - 1; q > 0; --q) s_order[q] = s_order[q - 1];
                        s_order[0] = b;
                    }
                    has_base = (base >= 0) ? 1u : 0u;
                }
                m = __shfl_sync(0xffffffffu, m, 0);
                has_base = __shfl_sync(0xffffffffu, has_base, 0);
                __syncwarp();
                const uint32_t n_batches = (m + kb - 1u) / kb + (m == 0u ? 1u : 0u);
                for (uint32_t b = 0; b < n_batches; ++b) {
                    const uint32_t b0 = b * kb;
                    const uint32_t nb = min(kb, m - min(m, b0));
                    mbar_wait(&bar_empty[stage], ephase);
                    ChainStage *S = &s_stage[stage];
                    uint8_t *sm = smem_raw + (size_t)stage * stage_bytes;
                    uint32_t bytes = 0;
                    if (lane < nb) {
                        const ChainRec rr = s_res[s_order[b0 + lane]];
                        bytes = stage_input(rr, lane, S, sm, &bar_full[stage]);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
                    if (lane == 0) stage_header(S, grp, nb, b == 0u, b + 1u == n_batches, has_base, (n * 3u) & (CH_CWARPS - 1u));
                    __syncwarp();
                    if (lane == 0) mbar_expect_tx(&bar_full[stage], bytes);
                    if (++stage == nstages) { stage = 0; ephase ^= 1u; }
                }
            }
            cp_async_wait_all();
            __syncwarp();
        }
        // ---- tell the consumers there is no more work
        mbar_wait(&bar_empty[stage], ephase);
        if (lane == 0) {
            s_stage[stage].stop = 1;
            mbar_expect_tx(&bar_full[stage], 0);
            // bank parity of the next tick (k_phase_chain of THIS tick has finished; nothing in this kernel reads the counter)
            if (tick_advance && blockIdx.x == 0) tick_advance[0] += 1u;
        }
        return;
    }

    // =================================================================== consumer warps
    // Block ownership ROTATES from session to session: the first 128 frames of a packet hold most of the program's
    // segments (the binades below 128) and the last ones its explicit tail, so a fixed assignment would make the same
    // warp the slowest of every session while a stage is only released when all eight are done.
    const uint32_t ct = threadIdx.x - 32u;
    unsigned long long acc[ITERS][CH_NB];   // (left, right) packed f32x2 per owned frame (mono: low half)
    float one_r;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(one_r) : "r"(smem_u32(&s_one)));
    const unsigned long long one2 = pack2(one_r, one_r);   // 1.0f the compiler cannot see (add2)
    uint32_t stage = 0, fphase = 0;
    for (;;) {
        mbar_wait(&bar_full[stage], fphase);
        const ChainStage *S = &s_stage[stage];
        const uint4 hd = *reinterpret_cast<const uint4 *>(&S->nb);   // nb, first, last, has_base (one broadcast load)
        if (S->stop) break;
        const uint32_t sm = smem_u32(smem_raw) + stage * stage_bytes;
        const uint32_t nb = hd.x;
        const uint32_t cw = (warp - 1u + (hd.w >> 8)) & (CH_CWARPS - 1u);   // this session's block group of the warp
        if (hd.y != 0) {
            // the base frame IS the accumulator (mixer.rs:969-972): start from -0.0, the additive identity of every f32
            // (-0.0 + v == v bit for bit, also for v == -0.0); without a base frame the mix starts from vec![0.0; n]
            const unsigned long long init = (hd.w & 1u) != 0 ? 0x8000000080000000ull : 0ull;
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
#pragma unroll
                for (int f = 0; f < CH_NB; ++f) acc[it][f] = init;
        }
        if (KIND == CHAIN_BYPASS || KIND == CHAIN_BYPASS_S16) {
            for (uint32_t q = 0; q < nb; ++q) {
                const uint32_t a_chunk = sm + q * in_bytes + prog_cap + SK_SIDE_HIST;
                // (expanding two inputs behind one pair of barriers was measured slower: the per-input barriers let the warps drift)
                if (KIND == CHAIN_BYPASS_S16) chain_expand_s16(a_chunk, S->cons[q].n_prev, S->cons[q].n_head, ct);
                chain_consume_pass<OC, OC, ITERS>(acc, a_chunk, F, cw, lane, S->cons[q].gain, one2);
            }
        } else if (KIND == CHAIN_PLAIN_S16) {
            for (uint32_t q = 0; q < nb; ++q) {
                const uint32_t prog = sm + q * in_bytes, a_chunk = prog + prog_cap + SK_SIDE_HIST;
                chain_expand_s16(a_chunk, S->cons[q].n_prev, S->cons[q].n_head, ct);
                chain_consume<OC, OC, ITERS>(acc, prog, a_chunk - 16u * OC * 4u, dm.prog, F, cw, lane, S->cons[q].gain, one2);
            }
        } else if (KIND == CHAIN_F32) {
            for (uint32_t q = 0; q < nb; ++q) {
                const ChainCons c = S->cons[q];
                const uint32_t prog = sm + q * in_bytes, a_chunk = prog + prog_cap + SK_SIDE_HIST;
                if (c.sc & CK_BYPASS) chain_consume_pass<OC, OC, ITERS>(acc, a_chunk, F, cw, lane, c.gain, one2);
                else chain_consume<OC, OC, ITERS>(acc, prog, a_chunk - 16u * OC * 4u, dm.prog, F, cw, lane, c.gain, one2);
            }
        } else if (KIND == CHAIN_PLAIN) {
            // every input of the op is a resampled f32 stream with the output's channel count (host: validate_chain): this
            // instantiation carries no input-kind code at all -- the consumer loop is sensitive to its instruction footprint
            for (uint32_t q = 0; q < nb; ++q) {
                const uint32_t prog = sm + q * in_bytes;
                chain_consume<OC, OC, ITERS>(acc, prog, prog + prog_cap + (OC == 2 ? 0u : SK_SIDE_HIST - 64u), dm.prog, F, cw, lane, S->cons[q].gain, one2);
            }
        } else {
            for (uint32_t q = 0; q < nb; ++q) {
                const ChainCons c = S->cons[q];
                const uint32_t prog = sm + q * in_bytes, a_chunk = prog + prog_cap + SK_SIDE_HIST;   // program record; staged chunk pieces
                if (c.sc & CK_S16) chain_expand_s16(a_chunk, c.n_prev, c.n_head, ct);
                if ((c.sc & 0xFFu) == 2u) {
                    if (c.sc & CK_BYPASS) chain_consume_pass<OC, 2, ITERS>(acc, a_chunk, F, cw, lane, c.gain, one2);
                    else chain_consume<OC, 2, ITERS>(acc, prog, a_chunk - 128u, dm.prog, F, cw, lane, c.gain, one2);
                } else {
                    if (c.sc & CK_BYPASS) chain_consume_pass<OC, 1, ITERS>(acc, a_chunk, F, cw, lane, c.gain, one2);
                    else chain_consume<OC, 1, ITERS>(acc, prog, a_chunk - 64u, dm.prog, F, cw, lane, c.gain, one2);
                }
            }
        }
        if (hd.z != 0) {
            // ---- epilogue: master gain, then clip + s16 pack (or f32); a warp stores 32 consecutive frames per instruction
            const float mg = S->master_gain;   // audio::gain after the mixer (1.0 when there is none: x * 1.0 == x)
            const uint32_t flags = S->flags;
            uint8_t *out_base = arena + S->out_off;
            const uint32_t jw = cw * (CH_NB * 32u) + lane;
            if (flags & SKGPU_MIX_OUT_S16) {
                // s16 = cvt.rni.sat(fl(a * g) * 32768). The power-of-two scale commutes with the rounding of a * g, so one
                // multiply by g * 32768 gives the same integer (a subnormal a * g rounds to 0 either way, overflow saturates).
                const float mgs = __fmul_rn(mg, 32768.0f);
                const unsigned long long mgs2 = pack2(mgs, mgs);
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                    for (int f = 0; f < CH_NB; ++f) {
                        const uint32_t j = jw + (uint32_t)it * (CH_CWARPS * CH_NB * 32u) + (uint32_t)f * 32u;
                        if (j >= F) continue;
                        float a0, a1;
                        unpack2(mul2v(acc[it][f], mgs2), a0, a1);
                        int r0, r1;
                        asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r0) : "f"(a0));
                        asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r1) : "f"(a1));
                        if (OC == 2) stg_stream_u32(reinterpret_cast<uint32_t *>(out_base) + j, __byte_perm((uint32_t)r0, (uint32_t)r1, 0x5410));
                        else reinterpret_cast<uint16_t *>(out_base)[j] = (uint16_t)r0;
                    }
                }
            } else {
                const unsigned long long mg2 = pack2(mg, mg);
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
#pragma unroll
                    for (int f = 0; f < CH_NB; ++f) {
                        const uint32_t j = jw + (uint32_t)it * (CH_CWARPS * CH_NB * 32u) + (uint32_t)f * 32u;
                        if (j >= F) continue;
                        float a0, a1;
                        unpack2(mul2v(acc[it][f], mg2), a0, a1);
                        if (OC == 2) stg_stream_f2(reinterpret_cast<float2 *>(out_base) + j, make_float2(a0, a1));
                        else reinterpret_cast<float *>(out_base)[j] = a0;
                    }
                }
            }
        }
        // ---- history before the current chunk := last 16 frames of the previous chunk (now retired), written into the
        // current chunk's side record (next tick it is the "previous" one)
        if (ct < 16u * 2u) {
            for (uint32_t q = 0; q < hd.x; ++q) {
                const ChainTail tl = S->tail[q];
                if (ct < tl.n) {
                    const float *chunk_f = reinterpret_cast<const float *>(smem_raw + (size_t)stage * stage_bytes + (size_t)q * in_bytes + prog_cap + SK_SIDE_HIST);
                    tl.hist_dst[ct] = chunk_f[tl.tail_off + ct];
                }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_empty[stage])) : "memory");
        if (++stage == nstages) { stage = 0; fphase ^= 1u; }
    }
}

// advances the device-side tick counter (bank parity) once per tick; part of the captured graph
__global__ void k_tick_advance(uint32_t *tick) { tick[0] += 1u; }

}  // namespace skgpu
