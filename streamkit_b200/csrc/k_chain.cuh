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
// input arena. HBM traffic per stream-tick is one pass over one input chunk plus a ~0.5 KB "side record"
// (phase table + 16-frame history), i.e. the "fully fused" algorithmic bytes of SURVEY 8(d).
//
// Two kernels per tick:
//   k_phase_chain  one THREAD per input stream (massively parallel, data independent): rubato's f64 phase recurrence
//                  -> compact phase table in the stream's side record; emission / carry / status bookkeeping
//                  (resampler.rs:425-428 re-framing); writes a ready-made 64-byte ChainRec per input.
//   k_chain        PERSISTENT, WARP-SPECIALISED CTAs (4 per SM, looping over sessions):
//       warp 0  (producer)   prefetches ChainRecs with cp.async, derives the summation order (base selection +
//                            swap_remove, mixer.rs:960-980) from warp ballots, and issues 3 TMA bulk copies per
//                            input (side record of the previous chunk, the previous chunk, table of the current
//                            chunk) into a 2-stage shared-memory ring guarded by full/empty mbarriers;
//       warps 1-8 (consumers) interpolate 4 consecutive output frames per thread from shared memory, add the inputs
//                            SEQUENTIALLY in the reference's order (f32 addition is not associative),
//                            apply the master gain, clip + pack s16 and store 16 bytes per thread.
#pragma once
#include "common.cuh"

namespace skgpu {

constexpr int CH_CONSUMERS = 256;                 // 8 consumer warps
constexpr int CH_THREADS = CH_CONSUMERS + 32;     // + producer warp (warp 0)
constexpr int CH_MAX_STAGES = 4;                  // pipeline depth is a launch parameter (2..4)
constexpr int CH_HEAD = 32;                       // frames of the CURRENT chunk staged behind the previous one (rarely needed)
constexpr int CH_MAX_INPUTS = 64;                 // inputs per session
constexpr int CH_FPT = 4;                         // output frames per consumer thread per iteration
constexpr int CH_MAX_KB = 4;                      // inputs staged per batch

struct __align__(16) ChainCons {   // what a consumer thread needs per input: 32 bytes = two broadcast 16-byte loads
    double t;               // 1 / resample_ratio
    float gain;
    uint32_t carry;
    uint32_t kdelta;        // n_prev - carry: chunk-output index of packet frame j (j < carry) is kdelta + j
    uint32_t n_cur;         // >= 1
    uint32_t np_nr;         // np_prev | nr_prev << 8 | np_cur << 16 | nr_cur << 24
    uint32_t na_flags;      // NA (24 bits) | channels << 24 | has_gain << 26 | needs_global << 27
};

// per-input record written by k_phase_chain every tick, consumed by k_chain's producer (64 bytes, one per chain input)
constexpr uint32_t CR_EMIT = 1u, CR_UNIQUE = 2u, CR_NEEDS_HEAD = 4u, CR_PAR_PREV = 8u, CR_HAS_PREV = 16u;
struct __align__(16) ChainRec {
    ChainCons cons;
    const float *prev_g;    // previous chunk (other input bank)
    const float *cur_g;     // current chunk
    uint32_t slot;
    uint32_t chunk_bytes;   // N * channels * 4
    uint32_t flags;         // CR_*
    uint32_t N;
};
static_assert(sizeof(ChainRec) == 64, "ChainRec is one 64-byte record");

struct ChainTail {          // per staged input: what the end-of-batch history update and the rare HBM path need
    uint8_t *side_cur;      // side record of the current chunk (its history field is written here)
    const float *cur_g;
    uint32_t N, ch, has_prev, pad;
};

struct __align__(16) ChainStage {   // header of one pipeline stage (shared memory)
    uint32_t nb;            // inputs in this batch           } one 16-byte load
    uint32_t first, last;   // first / last batch of the session
    uint32_t has_base;      // the first input of the first batch is the base frame (mixer.rs:960-972)
    uint64_t out_off;
    float master_gain;
    uint32_t has_master;
    uint32_t flags;
    uint32_t stop;          // no more work
    uint32_t pad[2];
    ChainCons cons[CH_MAX_KB];
    ChainTail tail[CH_MAX_KB];
};

struct ChainDims {          // launch-time geometry of the staging ring (host: chain_size_smem)
    uint32_t kb;            // inputs per batch
    uint32_t chunk_cap;     // bytes reserved per input for [previous chunk | CH_HEAD frames of the current one], 16-aligned
    uint32_t cap_np, cap_nr;   // phase-table capacity: prefix doubles / runs (both even) -> tab_bytes = cap_np*8 + cap_nr*24
    uint32_t nstages;
    uint32_t max_k;         // largest n_inputs of any session (sizes the producer scratch)
    uint32_t debug;         // profiling only: bit0 = consumers skip the arithmetic (isolates the load pipeline)
};
// shared-memory slot of one staged input:  [table prev | history field 128 B | previous chunk | head of current] [table cur]
__host__ __device__ __forceinline__ uint32_t chain_tab_bytes(const ChainDims &dm) { return dm.cap_np * 8u + dm.cap_nr * (uint32_t)sizeof(SkRun); }
__host__ __device__ __forceinline__ uint32_t chain_in_bytes(const ChainDims &dm) { return 2u * chain_tab_bytes(dm) + SK_SIDE_HIST + dm.chunk_cap; }

// packed f32x2 multiply (Blackwell FMUL2): two IEEE-rounded products per instruction. Additions stay scalar
// FADDs on purpose: ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 (one rounding less than the reference).
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, float b) {
    unsigned long long d, bb;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(bb));
    return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long pack2(float a, float b) {
    unsigned long long d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b));
    return d;
}

struct ChainInRegs {        // what the inner loop needs, in registers
    const float *cur_g;     // only set on the rare HBM path
    double t;
    float gain;
    uint32_t has_gain, carry, kdelta, n_cur, NA;
};

// one (input frame -> output frame) accumulate step for the non-stereo/stereo combinations (mixer.rs:1027-1078)
template <int OC, int SC, bool IS_BASE>
__device__ __forceinline__ void chain_accumulate(float *acc, const float *y) {
    float v[2];
    if (SC == OC) { v[0] = y[0]; v[1] = y[1]; }
    else if (SC == 1 && OC == 2) { v[0] = y[0]; v[1] = y[0]; }
    else { v[0] = __fmul_rn(__fadd_rn(y[0], y[1]), 0.5f); v[1] = 0.0f; }
#pragma unroll
    for (int c = 0; c < OC; ++c) acc[c] = IS_BASE ? v[c] : __fadd_rn(acc[c], v[c]);
}

// interpolate + gain + accumulate one frame at phase x; `idx_base` selects history++previous (0) or the current chunk (NA)
template <int OC, int SC, bool IS_BASE>
__device__ __forceinline__ void chain_frame(const ChainInRegs &in, const float *A, double x, uint32_t idx_base, float *acc) {
    uint32_t p;
    float frac;
    phase_split(x, p, frac);
    const uint32_t idx = idx_base + p;   // frame index into A
    if (SC == 2 && OC == 2) {
        const unsigned long long y0 = *reinterpret_cast<const unsigned long long *>(A + 2u * idx);
        const unsigned long long y1 = *reinterpret_cast<const unsigned long long *>(A + 2u * idx + 2u);
        float a0, a1, b0, b1;
        unpack2(mul2(y0, __fsub_rn(1.0f, frac)), a0, a1);   // rubato interp_lin: (1 - frac) * y0 + frac * y1
        unpack2(mul2(y1, frac), b0, b1);
        float r0 = __fadd_rn(a0, b0), r1 = __fadd_rn(a1, b1);
        if (in.has_gain) unpack2(mul2(pack2(r0, r1), in.gain), r0, r1);   // the input's audio::gain (gain.rs:187-189)
        acc[0] = IS_BASE ? r0 : __fadd_rn(acc[0], r0);
        acc[1] = IS_BASE ? r1 : __fadd_rn(acc[1], r1);
    } else {
        float y[2] = {0.0f, 0.0f};
#pragma unroll
        for (int c = 0; c < SC; ++c) {
            y[c] = interp_lin(frac, A[idx * SC + c], A[(idx + 1u) * SC + c]);
            if (in.has_gain) y[c] = __fmul_rn(y[c], in.gain);
        }
        chain_accumulate<OC, SC, IS_BASE>(acc, y);
    }
}

// rare path (right after a stream starts, or odd ratios): frames of the current chunk beyond the staged head are read from HBM
template <int OC, int SC>
__device__ __noinline__ void chain_input_global(const ChainInRegs &in, const float *A, const PhaseView &Tp, const PhaseView &Tc, uint32_t F,
                                                uint32_t j0, float *acc, bool is_base) {
    for (int f = 0; f < CH_FPT; ++f) {
        const uint32_t j = j0 + f;
        if (j >= F) break;
        uint32_t p;
        float frac;
        float y[2] = {0.0f, 0.0f};
        if (j < in.carry) {
            phase_split(pv_eval(Tp, in.t, in.kdelta + j), p, frac);
            for (int c = 0; c < SC; ++c) y[c] = interp_lin(frac, A[p * SC + c], A[(p + 1u) * SC + c]);
        } else {
            phase_split(pv_eval(Tc, in.t, min(j - in.carry, in.n_cur - 1u)), p, frac);
            for (int c = 0; c < SC; ++c) {
                const float y0 = (p < 16u) ? A[(in.NA + p) * SC + c] : in.cur_g[(size_t)(p - 16u) * SC + c];
                const float y1 = (p + 1u < 16u) ? A[(in.NA + p + 1u) * SC + c] : in.cur_g[(size_t)(p + 1u - 16u) * SC + c];
                y[c] = interp_lin(frac, y0, y1);
            }
        }
        for (int c = 0; c < SC; ++c)
            if (in.has_gain) y[c] = __fmul_rn(y[c], in.gain);
        if (is_base) chain_accumulate<OC, SC, true>(acc + f * OC, y);
        else chain_accumulate<OC, SC, false>(acc + f * OC, y);
    }
}

// one input of one session, 4 consecutive output frames j0..j0+3 of this thread (frames >= F are computed on clamped
// indices and never stored). A = [16 history | previous chunk (NA frames) | CH_HEAD frames of the current chunk].
template <int OC, int SC, bool IS_BASE>
__device__ __forceinline__ void chain_input(const ChainInRegs &in, const float *A, const PhaseView &Tp, const PhaseView &Tc, uint32_t j0,
                                            float *acc) {
    if (j0 + CH_FPT <= in.carry) {
        // FAST PATH (954 of 960 frames in steady state, 7 of 8 warps entirely): all four frames were produced by the
        // PREVIOUS chunk; one run lookup, then x, x+d, x+2d, x+3d (exact inside a run)
        const uint32_t k0 = in.kdelta + j0;
        double x[4];
        bool in_run = false;
        if (k0 >= Tp.n_prefix) {
            uint32_t r = Tp.n_runs - 1u;
            while (r > 0u && Tp.runs[r].k_a > k0) --r;
            const SkRun rn = Tp.runs[r];
            if (k0 + 3u < rn.k_e) {
                x[0] = __fma_rn((double)(k0 - rn.k_a), rn.delta, rn.x_a);
                x[1] = __dadd_rn(x[0], rn.delta);
                x[2] = __dadd_rn(x[1], rn.delta);
                x[3] = __dadd_rn(x[2], rn.delta);
                in_run = true;
            }
        }
        if (!in_run) {
#pragma unroll
            for (int f = 0; f < CH_FPT; ++f) x[f] = pv_eval(Tp, in.t, k0 + (uint32_t)f);
        }
#pragma unroll
        for (int f = 0; f < CH_FPT; ++f) chain_frame<OC, SC, IS_BASE>(in, A, x[f], 0u, acc + f * OC);
        return;
    }
    // the thread that straddles the prev/cur boundary and the few behind it (one warp per session): frame by frame.
    // Frames produced by the CURRENT chunk use its own table; their history is the tail of the previous chunk, which
    // precedes the staged head of the current chunk in A.
#pragma unroll
    for (int f = 0; f < CH_FPT; ++f) {
        const uint32_t j = j0 + f;
        const bool from_cur = j >= in.carry;
        const double x = from_cur ? pv_eval(Tc, in.t, min(j - in.carry, in.n_cur - 1u)) : pv_eval(Tp, in.t, in.kdelta + j);
        chain_frame<OC, SC, IS_BASE>(in, A, x, from_cur ? in.NA : 0u, acc + f * OC);
    }
}


// ------------------------------------------------------------------ k_phase_chain
constexpr int PHASE_CHAIN_THREADS = 64;

__global__ void __launch_bounds__(PHASE_CHAIN_THREADS) k_phase_chain(const OpHeader *__restrict__ hdr, const skgpu_chain_input *__restrict__ inputs,
                                                                     const uint8_t *__restrict__ present, const float *__restrict__ gains, SlotTables st,
                                                                     uint8_t *__restrict__ arena, const uint32_t *__restrict__ tick, uint64_t bank_stride,
                                                                     uint32_t F, uint64_t results_off, ChainDims dm, ChainRec *__restrict__ recs) {
    const uint32_t i = blockIdx.x * PHASE_CHAIN_THREADS + threadIdx.x;
    if (i >= hdr->count2) return;
    const skgpu_chain_input in = inputs[i];
    const uint32_t slot = in.slot;
    SlotRec *recp = st.rec + slot;
    SlotRec rec = *recp;
    const bool pres = present ? (present[i] != 0) : true;
    const uint32_t tab_bytes = chain_tab_bytes(dm);
    uint32_t status = 0;
    if (pres) {
        // ---- rubato's phase recurrence for this chunk -> compact table in the side record of (chunk number & 1)
        const uint32_t par = rec.chunk_count & 1u;
        uint8_t *side = slot_side(st, slot, par);
        uint32_t np, nr, ovf;
        double idx_end;
        const uint32_t n = sk_phase_table_ex(rec.last_index, rec.t_ratio, rec.end_idx, reinterpret_cast<double *>(side), dm.cap_np,
                                             reinterpret_cast<SkRun *>(side + dm.cap_np * 8u), dm.cap_nr, &np, &nr, &ovf, &idx_end);
        rec.last_index = __dsub_rn(idx_end, (double)rec.chunk);   // self.last_index = idx - chunk_size as f64
        rec.chunk_count += 1u;
        rec.n_out[par] = n;
        rec.n_prefix[par] = (uint16_t)np;
        rec.n_runs[par] = (uint16_t)nr;
        rec.overflow = (rec.overflow & ~(1u << par)) | ((ovf ? 1u : 0u) << par);
        if (ovf) status |= 2u;
    }
    // ---- re-framing bookkeeping (resampler.rs:425-428): a packet is emitted when carry + n_cur >= F
    const uint32_t count = rec.chunk_count;
    const uint32_t par_cur = (count - 1u) & 1u, par_prev = count & 1u;
    const uint32_t n_cur = (pres && count >= 1u) ? rec.n_out[par_cur] : 0u;
    const uint32_t n_prev = (count >= 2u) ? rec.n_out[par_prev] : 0u;
    const uint32_t carry = rec.carry;
    uint32_t emit = 0;
    if (pres) {
        if (carry > n_prev) status |= 4u;                     // carried frames span more than one chunk: unsupported
        if ((rec.overflow >> par_prev) & 1u) status |= 2u;
        const uint32_t avail = carry + n_cur;
        emit = (avail >= F && !(status & 6u)) ? 1u : 0u;
        const uint32_t new_carry = (avail >= F) ? avail - F : avail;
        if (new_carry > n_cur) status |= 1u;                  // backlog: a second packet is pending
        rec.carry = new_carry;
        *recp = rec;
    }
    skgpu_chain_result res;
    res.emitted = emit;
    res.status = status;
    reinterpret_cast<skgpu_chain_result *>(arena + results_off)[i] = res;

    const uint32_t parity = tick[0] & 1u;
    const uint32_t ch = rec.channels, N = rec.chunk;
    const float *cur_g = reinterpret_cast<const float *>(arena + in.in_off + (uint64_t)parity * bank_stride);
    const float *prev_g = reinterpret_cast<const float *>(arena + in.in_off + (uint64_t)(1u - parity) * bank_stride);
    // a present input that emits nothing still retires its previous chunk: the history before the current chunk
    // (= last 16 frames of the previous one) goes into the current chunk's side record. Emitting inputs get it from
    // k_chain's consumers, which have the previous chunk in shared memory anyway.
    if (pres && !emit && count >= 2u) {
        float *h = reinterpret_cast<float *>(slot_side(st, slot, par_cur) + tab_bytes + SK_SIDE_HIST) - 16u * ch;
        for (uint32_t e = 0; e < 16u * ch; ++e) h[e] = prev_g[(size_t)(N - 16u) * ch + e];
    }
    ChainRec r;
    const bool has_gain = in.gain_idx != SKGPU_NO_GAIN;
    r.cons.t = rec.t_ratio;
    r.cons.gain = has_gain ? gains[in.gain_idx] : 1.0f;
    r.cons.carry = carry;
    r.cons.kdelta = n_prev - min(carry, n_prev);
    r.cons.n_cur = max(n_cur, 1u);
    r.cons.np_nr = (count >= 2u ? ((uint32_t)rec.n_prefix[par_prev] | ((uint32_t)rec.n_runs[par_prev] << 8)) : 0u) |
                   ((uint32_t)rec.n_prefix[par_cur] << 16) | ((uint32_t)rec.n_runs[par_cur] << 24);
    // positions the current-chunk frames can reach: idx < -9 + (F - carry) * t  (last_index < -9 after the first chunk)
    const float reach = (emit && carry < F) ? (float)(F - carry) * (float)rec.t_ratio : 0.0f;
    r.cons.na_flags = ((count >= 2u) ? N : 0u) | (ch << 24) | ((has_gain ? 1u : 0u) << 26) | ((reach > (float)(CH_HEAD + 6) ? 1u : 0u) << 27);
    r.prev_g = prev_g;
    r.cur_g = cur_g;
    r.slot = slot;
    r.chunk_bytes = N * ch * 4u;
    r.flags = (emit ? CR_EMIT : 0u) | ((in.flags & SKGPU_MIX_IN_UNIQUE) ? CR_UNIQUE : 0u) | (reach >= 7.0f ? CR_NEEDS_HEAD : 0u) |
              (par_prev ? CR_PAR_PREV : 0u) | (count >= 2u ? CR_HAS_PREV : 0u);
    r.N = N;
    recs[i] = r;
}

template <int OC, int ITERS>  // output channels (1 | 2); ITERS = ceil(F / 1024)
__global__ void __launch_bounds__(CH_THREADS, 4) k_chain(const OpHeader *__restrict__ hdr, const skgpu_chain_group *__restrict__ groups,
                                                         const ChainRec *__restrict__ recs, const float *__restrict__ gains, SlotTables st,
                                                         uint8_t *__restrict__ arena, uint32_t F, ChainDims dm) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[CH_MAX_STAGES], bar_empty[CH_MAX_STAGES];
    __shared__ __align__(16) ChainStage s_stage[CH_MAX_STAGES];
    __shared__ uint8_t s_order[CH_MAX_INPUTS];

    // dynamic smem: nstages x kb input slots, then the producer's scratch for sessions that need several batches
    const uint32_t kb = dm.kb, nstages = dm.nstages;
    const uint32_t tab_bytes = chain_tab_bytes(dm);
    const uint32_t in_bytes = chain_in_bytes(dm);
    const uint32_t stage_bytes = kb * in_bytes;
    ChainRec *s_res = reinterpret_cast<ChainRec *>(smem_raw + (size_t)stage_bytes * nstages);
    const uint32_t n_groups = hdr->count;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;

    if (threadIdx.x == 0) {
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
        __shared__ __align__(16) ChainRec pf_rec[2][32];
        const uint32_t gstep = gridDim.x;
        uint32_t stage = 0, ephase = 1;  // waiting on parity 1 of a fresh mbarrier returns immediately
        auto sess = [&](uint32_t n) -> uint32_t { return blockIdx.x + n * gstep; };
        auto issue_grp = [&](uint32_t n) {
            if (lane < 3u && sess(n) < n_groups) cp_async8(reinterpret_cast<uint8_t *>(&pf_grp[n & 3u]) + lane * 8u, reinterpret_cast<const uint8_t *>(&groups[sess(n)]) + lane * 8u);
        };
        auto issue_rec = [&](uint32_t n) {
            if (sess(n) >= n_groups) return;
            const skgpu_chain_group &g = pf_grp[n & 3u];
            if (lane < min(g.n_inputs, 32u)) {
                const uint8_t *src = reinterpret_cast<const uint8_t *>(recs + g.first_input + lane);
                uint8_t *dst = reinterpret_cast<uint8_t *>(&pf_rec[n & 1u][lane]);
#pragma unroll
                for (int w = 0; w < 4; ++w) cp_async16(dst + w * 16, src + w * 16);
            }
        };
        // one input: copy its consumer record into the stage and issue its bulk copies; returns the bytes the barrier must expect
        auto stage_input = [&](const ChainRec &r, uint32_t q, ChainStage *S, uint8_t *sm, uint64_t *bar) -> uint32_t {
            S->cons[q] = r.cons;
            const uint32_t ch = (r.cons.na_flags >> 24) & 3u;
            const uint32_t par_prev = (r.flags & CR_PAR_PREV) ? 1u : 0u;
            ChainTail tl;
            tl.side_cur = slot_side(st, r.slot, par_prev ^ 1u);
            tl.cur_g = r.cur_g;
            tl.N = r.N; tl.ch = ch; tl.has_prev = (r.flags & CR_HAS_PREV) ? 1u : 0u; tl.pad = 0;
            S->tail[q] = tl;
            uint8_t *slot_sm = sm + (size_t)q * in_bytes;
            uint8_t *chunk_sm = slot_sm + tab_bytes + SK_SIDE_HIST;
            const uint32_t cb = (r.flags & CR_HAS_PREV) ? r.chunk_bytes : 0u;
            const uint32_t hb = (r.flags & CR_NEEDS_HEAD) ? min((uint32_t)CH_HEAD, r.N) * ch * 4u : 0u;
            uint32_t bytes = 2u * tab_bytes + SK_SIDE_HIST;
            tma_bulk_g2s(slot_sm, slot_side(st, r.slot, par_prev), tab_bytes + SK_SIDE_HIST, bar);              // table + history of the previous chunk
            tma_bulk_g2s(chunk_sm + dm.chunk_cap, slot_side(st, r.slot, par_prev ^ 1u), tab_bytes, bar);       // table of the current chunk
            if (!(cb & 15u)) {
                if (cb) { tma_bulk_g2s(chunk_sm, r.prev_g, cb, bar); bytes += cb; }
                if (hb) { tma_bulk_g2s(chunk_sm + cb, r.cur_g, hb, bar); bytes += hb; }
            } else {   // chunk size not a multiple of 16 bytes (e.g. mono 882 frames): no bulk copy, this lane copies
                float *dst = reinterpret_cast<float *>(chunk_sm);
                for (uint32_t e = 0; e < cb / 4u; ++e) dst[e] = r.prev_g[e];
                for (uint32_t e = 0; e < hb / 4u; ++e) dst[cb / 4u + e] = r.cur_g[e];
            }
            return bytes;
        };
        auto stage_header = [&](ChainStage *S, const skgpu_chain_group &grp, uint32_t nb, bool first, bool last, uint32_t has_base) {
            S->nb = nb;
            S->first = first;
            S->last = last;
            S->has_base = has_base;
            S->out_off = grp.out_off;
            S->has_master = grp.gain_idx != SKGPU_NO_GAIN;
            S->master_gain = S->has_master ? gains[grp.gain_idx] : 1.0f;
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
            r.cons.na_flags = 0;
            if (lane < min(K, 32u)) r = pf_rec[n & 1u][lane];
            const bool emit = (r.flags & CR_EMIT) != 0;
            const bool elig = emit && ((r.cons.na_flags >> 24) & 3u) == (uint32_t)OC;   // packet already has the output shape
            // ---- summation order over the inputs that deliver a packet: base selection + swap_remove (mixer.rs:960-980),
            // computed by every lane from three ballots
            const uint32_t emit_mask = __ballot_sync(0xffffffffu, emit);
            const uint32_t elig_mask = __ballot_sync(0xffffffffu, elig);
            const uint32_t uniq_mask = __ballot_sync(0xffffffffu, elig && (r.flags & CR_UNIQUE));
            uint32_t m = __popc(emit_mask);
            if (K <= 32u && m <= kb) {
                // max_by_key((unique, idx)): the last unique full-shape frame, else the last full-shape frame
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
                uint8_t *sm = smem_raw + (size_t)stage * stage_bytes;
                uint32_t bytes = emit ? stage_input(r, pos, S, sm, &bar_full[stage]) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
                if (lane == 0) stage_header(S, grp, m, true, true, base_lane >= 0 ? 1u : 0u);
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&bar_full[stage], bytes);   // arrive: the lanes' smem writes are ordered before it
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
                        if (((s_res[j].cons.na_flags >> 24) & 3u) == (uint32_t)OC) {
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
                    if (lane == 0) stage_header(S, grp, nb, b == 0u, b + 1u == n_batches, has_base);
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
        }
        return;
    }

    // =================================================================== consumer warps
    const uint32_t ct = threadIdx.x - 32u;  // 0..255
    float acc[ITERS][CH_FPT * OC];
    uint32_t stage = 0, fphase = 0;
    for (;;) {
        mbar_wait(&bar_full[stage], fphase);
        const ChainStage *S = &s_stage[stage];
        const uint4 hd = *reinterpret_cast<const uint4 *>(&S->nb);   // nb, first, last, has_base (one broadcast load)
        if (S->stop) break;
        const uint8_t *sm = smem_raw + (size_t)stage * stage_bytes;
        const uint32_t nb = (dm.debug & 1u) ? 0u : hd.x;
        const bool first = hd.y != 0;
        if (first) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
#pragma unroll
                for (int e = 0; e < CH_FPT * OC; ++e) acc[it][e] = 0.0f;  // vec![0.0f32; output_size] when there is no base frame
        }
        for (uint32_t q = 0; q < nb; ++q) {
            // everything the inner loop needs about this input: two broadcast 16-byte loads
            const uint4 c0 = *reinterpret_cast<const uint4 *>(&S->cons[q]);
            const uint4 c1 = *(reinterpret_cast<const uint4 *>(&S->cons[q]) + 1);
            ChainInRegs in;
            in.t = __hiloint2double((int)c0.y, (int)c0.x);
            in.gain = __uint_as_float(c0.z);
            in.carry = c0.w;
            in.kdelta = c1.x;
            in.n_cur = c1.y;
            in.NA = c1.w & 0xFFFFFFu;
            in.has_gain = (c1.w >> 26) & 1u;
            in.cur_g = nullptr;
            const uint32_t sc = (c1.w >> 24) & 3u;
            const bool needs_global = ((c1.w >> 27) & 1u) != 0;
            const uint8_t *slot_sm = sm + (size_t)q * in_bytes;
            const uint8_t *chunk_sm = slot_sm + tab_bytes + SK_SIDE_HIST;
            // A = [16 history frames | previous chunk | head of the current chunk]; the history field ends where the chunk begins
            const float *A = reinterpret_cast<const float *>(chunk_sm) - 16u * sc;
            const uint8_t *tc = chunk_sm + dm.chunk_cap;
            PhaseView Tp, Tc;
            Tp.prefix = reinterpret_cast<const double *>(slot_sm);
            Tp.runs = reinterpret_cast<const SkRun *>(slot_sm + dm.cap_np * 8u);
            Tp.n_prefix = c1.z & 0xFFu; Tp.n_runs = (c1.z >> 8) & 0xFFu; Tp.n_out = max(in.kdelta + in.carry, 1u);
            Tc.prefix = reinterpret_cast<const double *>(tc);
            Tc.runs = reinterpret_cast<const SkRun *>(tc + dm.cap_np * 8u);
            Tc.n_prefix = (c1.z >> 16) & 0xFFu; Tc.n_runs = (c1.z >> 24) & 0xFFu; Tc.n_out = in.n_cur;
            const bool is_base = first && hd.w != 0 && q == 0u;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t j0 = (ct + it * CH_CONSUMERS) * CH_FPT;
                const uint32_t j0_warp_first = (ct - lane + it * CH_CONSUMERS) * CH_FPT;   // warp-uniform
                if (j0_warp_first >= F) continue;                                            // whole warp past the packet
                if (needs_global) {
                    if (j0 < F) {
                        // rare path, out of line: it works on a copy so that `acc` itself never has its address taken
                        // (an escaping pointer would push the accumulators into local memory for the hot path too)
                        float tmp[CH_FPT * OC];
#pragma unroll
                        for (int e = 0; e < CH_FPT * OC; ++e) tmp[e] = acc[it][e];
                        in.cur_g = S->tail[q].cur_g;
                        if (sc == 2u) chain_input_global<OC, 2>(in, A, Tp, Tc, F, j0, tmp, is_base);
                        else chain_input_global<OC, 1>(in, A, Tp, Tc, F, j0, tmp, is_base);
#pragma unroll
                        for (int e = 0; e < CH_FPT * OC; ++e) acc[it][e] = tmp[e];
                    }
                } else if (is_base) {   // uniform: the base frame IS the accumulator (no add), mixer.rs:969-972
                    if (sc == 2u) chain_input<OC, 2, true>(in, A, Tp, Tc, j0, acc[it]);
                    else chain_input<OC, 1, true>(in, A, Tp, Tc, j0, acc[it]);
                } else {
                    if (sc == 2u) chain_input<OC, 2, false>(in, A, Tp, Tc, j0, acc[it]);
                    else chain_input<OC, 1, false>(in, A, Tp, Tc, j0, acc[it]);
                }
            }
        }
        if (hd.z != 0) {
            // ---- epilogue: master gain, then clip + s16 pack (or f32)
            const bool has_master = S->has_master != 0;
            const float mg = S->master_gain;
            const uint32_t flags = S->flags;
            uint8_t *out_base = arena + S->out_off;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t j0 = (ct + it * CH_CONSUMERS) * CH_FPT;
                if (j0 >= F) continue;
                float *a = acc[it];
                if (has_master) {
#pragma unroll
                    for (int e = 0; e < CH_FPT * OC; ++e) a[e] = __fmul_rn(a[e], mg);
                }
                const uint32_t nfr = min((uint32_t)CH_FPT, F - j0);
                if (flags & SKGPU_MIX_OUT_S16) {
                    uint16_t *o = reinterpret_cast<uint16_t *>(out_base) + (size_t)j0 * OC;
                    if (nfr == CH_FPT && OC == 2) {
                        stg_stream_u4(reinterpret_cast<uint4 *>(o), make_uint4(pack_s16x2(a[0], a[1]), pack_s16x2(a[2], a[3]),
                                                                              pack_s16x2(a[4 % (CH_FPT * OC)], a[5 % (CH_FPT * OC)]),
                                                                              pack_s16x2(a[6 % (CH_FPT * OC)], a[7 % (CH_FPT * OC)])));
                    } else if (nfr == CH_FPT && OC == 1) {
                        stg_stream_u2(reinterpret_cast<uint2 *>(o), make_uint2(pack_s16x2(a[0], a[1]), pack_s16x2(a[2], a[3])));
                    } else {
                        for (uint32_t e = 0; e < nfr * OC; ++e) o[e] = (uint16_t)f32_to_s16_bits(a[e]);
                    }
                } else {
                    float *o = reinterpret_cast<float *>(out_base) + (size_t)j0 * OC;
                    if (nfr == CH_FPT) {
                        stg_stream_f4(reinterpret_cast<float4 *>(o), make_float4(a[0], a[1], a[2], a[3]));
                        if (OC == 2) stg_stream_f4(reinterpret_cast<float4 *>(o) + 1, make_float4(a[4 % (CH_FPT * OC)], a[5 % (CH_FPT * OC)], a[6 % (CH_FPT * OC)], a[7 % (CH_FPT * OC)]));
                    } else {
                        for (uint32_t e = 0; e < nfr * OC; ++e) o[e] = a[e];
                    }
                }
            }
        }
        // ---- history before the current chunk := last 16 frames of the previous chunk (now retired), written into the
        // current chunk's side record (next tick it is the "previous" one and travels with its table in one bulk copy)
        if (ct < 16u * 2u) {
            for (uint32_t q = 0; q < hd.x; ++q) {
                const ChainTail tl = S->tail[q];
                if (tl.has_prev && ct < 16u * tl.ch) {
                    const float *chunk_f = reinterpret_cast<const float *>(sm + (size_t)q * in_bytes + tab_bytes + SK_SIDE_HIST);
                    float *h = reinterpret_cast<float *>(tl.side_cur + tab_bytes + SK_SIDE_HIST) - 16u * tl.ch;
                    h[ct] = chunk_f[(size_t)(tl.N - 16u) * tl.ch + ct];
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
