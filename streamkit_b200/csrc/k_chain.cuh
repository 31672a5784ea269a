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
// input arena. HBM traffic per stream-tick is one pass over one input chunk (+ 2 x 128 B history state and a
// ~0.3 KB phase table), i.e. the "fully fused" algorithmic bytes of SURVEY 8(d).
//
// Shape: PERSISTENT, WARP-SPECIALISED CTAs (a few per SM, looping over sessions):
//   warp 0  (producer)  resolves the next session (descriptors -> 64-byte slot records -> emission / carry
//                       bookkeeping -> summation order) and issues TMA bulk copies (history, previous chunk,
//                       phase tables) into a 2-stage shared-memory ring, signalling an mbarrier per stage;
//   warps 1-8 (consumers) wait on the stage, interpolate 4 consecutive output frames per thread from shared
//                       memory, add the inputs SEQUENTIALLY in the reference's order (f32 addition is not
//                       associative), and store 16 bytes of s16 per thread.
// The producer runs ahead, so the chain of dependent global loads that precedes every session's TMA is
// hidden behind the previous session's arithmetic.
#pragma once
#include "common.cuh"

namespace skgpu {

constexpr int CH_CONSUMERS = 256;                 // 8 consumer warps
constexpr int CH_THREADS = CH_CONSUMERS + 32;     // + producer warp (warp 0)
constexpr int CH_MAX_STAGES = 4;                // pipeline depth is a launch parameter (2..4)
constexpr int CH_HEAD = 32;                     // frames of the CURRENT chunk staged behind the previous one
constexpr int CH_MAX_INPUTS = 64;                 // inputs per session
constexpr int CH_FPT = 4;                         // output frames per consumer thread per iteration
constexpr int CH_MAX_KB = 4;                      // inputs staged per batch

struct ChainIn {            // per-tick view of one input of the session
    const float *prev_g;    // previous chunk (other bank)
    const float *cur_g;     // current chunk
    float *hist_g;          // st.hist of the slot: 16 frames before the previous chunk
    double t;
    float gain;
    uint32_t has_gain;
    uint32_t slot, N, ch;
    uint32_t carry, n_prev, n_cur, count;
    uint32_t emit, unique, par_prev, par_cur;
    uint16_t np_prev, nr_prev, np_cur, nr_cur;   // phase-table sizes (from the slot record)
};

struct ChainStage {         // header of one pipeline stage (shared memory)
    uint64_t out_off;
    float master_gain;
    uint32_t has_master;
    uint32_t flags;
    uint32_t nb;            // inputs in this batch
    uint32_t first, last;   // first / last batch of the session
    uint32_t has_base;      // the first input of the first batch is the base frame (mixer.rs:960-972)
    uint32_t stop;          // no more work
    ChainIn in[CH_MAX_KB];
};

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

struct ChainInRegs {        // the fields of ChainIn the inner loop needs, in registers
    const float *cur_g;
    double t;
    float gain;
    uint32_t has_gain, sc, carry, n_prev, n_cur, NA;
};

// one input of one session, 4 consecutive output frames j0..j0+3 of this thread (frames >= F are computed on clamped
// indices and never stored). A = [16 history | previous chunk (NA frames) | CH_HEAD frames of the current chunk].
template <int OC, int SC>
__device__ __forceinline__ void chain_input(const ChainInRegs &in, const float *A, const SmemPhase *Tp, const SmemPhase *Tc, uint32_t F,
                                            uint32_t j0, float *acc, bool is_base) {
    double x[4];
    if (j0 + CH_FPT <= in.carry) {
        // FAST PATH (954 of 960 frames in steady state): all four frames were produced by the PREVIOUS chunk;
        // recompute them from (history ++ previous chunk), everything in shared memory, no selects.
        phase_eval4(Tp, in.t, in.n_prev - in.carry + j0, 4u, x);
        if (SC == 2 && OC == 2) {
#pragma unroll
            for (int f = 0; f < CH_FPT; ++f) {
                uint32_t p;
                float frac;
                phase_split(x[f], p, frac);
                const unsigned long long y0 = *reinterpret_cast<const unsigned long long *>(A + 2u * p);
                const unsigned long long y1 = *reinterpret_cast<const unsigned long long *>(A + 2u * p + 2u);
                float a0, a1, b0, b1;
                unpack2(mul2(y0, __fsub_rn(1.0f, frac)), a0, a1);   // rubato interp_lin: (1 - frac) * y0 + frac * y1
                unpack2(mul2(y1, frac), b0, b1);
                float r0 = __fadd_rn(a0, b0), r1 = __fadd_rn(a1, b1);
                if (in.has_gain) unpack2(mul2(pack2(r0, r1), in.gain), r0, r1);   // the input's audio::gain
                acc[f * 2] = is_base ? r0 : __fadd_rn(acc[f * 2], r0);
                acc[f * 2 + 1] = is_base ? r1 : __fadd_rn(acc[f * 2 + 1], r1);
            }
            return;
        }
    } else if (j0 >= in.carry) {
        // all four from the CURRENT chunk (its history is the tail of the previous chunk, which precedes it in A)
        const uint32_t k0 = j0 - in.carry;
        phase_eval4(Tc, in.t, min(k0, in.n_cur - 1u), min(4u, in.n_cur - min(k0, in.n_cur - 1u)), x);
    } else {
#pragma unroll
        for (int f = 0; f < CH_FPT; ++f) {
            const uint32_t j = j0 + f;
            x[f] = (j < in.carry) ? phase_eval_smem(Tp, in.t, in.n_prev - in.carry + j)
                                  : phase_eval_smem(Tc, in.t, min(j - in.carry, in.n_cur - 1u));
        }
    }
#pragma unroll
    for (int f = 0; f < CH_FPT; ++f) {
        uint32_t p;
        float frac;
        phase_split(x[f], p, frac);
        const bool from_cur = (j0 + f) >= in.carry;
        const uint32_t idx = (from_cur ? in.NA : 0u) + p;   // frame index into A
        if (SC == 2 && OC == 2) {
            unsigned long long y0, y1;
            if (from_cur && p + 1u >= 16u + CH_HEAD) {       // beyond the staged head (only right after a stream starts)
                y0 = *reinterpret_cast<const unsigned long long *>(in.cur_g + 2u * (size_t)(p - 16u));
                y1 = *reinterpret_cast<const unsigned long long *>(in.cur_g + 2u * (size_t)(p - 16u) + 2u);
            } else {
                y0 = *reinterpret_cast<const unsigned long long *>(A + 2u * idx);
                y1 = *reinterpret_cast<const unsigned long long *>(A + 2u * idx + 2u);
            }
            // rubato interp_lin per channel: (1 - frac) * y0 + frac * y1, then the input's audio::gain, then the sum
            float a0, a1, b0, b1;
            unpack2(mul2(y0, __fsub_rn(1.0f, frac)), a0, a1);
            unpack2(mul2(y1, frac), b0, b1);
            float r0 = __fadd_rn(a0, b0), r1 = __fadd_rn(a1, b1);
            if (in.has_gain) unpack2(mul2(pack2(r0, r1), in.gain), r0, r1);
            acc[f * 2] = is_base ? r0 : __fadd_rn(acc[f * 2], r0);
            acc[f * 2 + 1] = is_base ? r1 : __fadd_rn(acc[f * 2 + 1], r1);
        } else {
            float y[2] = {0.0f, 0.0f};
#pragma unroll
            for (int c = 0; c < SC; ++c) {
                float y0, y1;
                if (from_cur && p + 1u >= 16u + CH_HEAD) {
                    y0 = in.cur_g[(size_t)(p - 16u) * SC + c];
                    y1 = in.cur_g[(size_t)(p + 1u - 16u) * SC + c];
                } else {
                    y0 = A[idx * SC + c];
                    y1 = A[(idx + 1u) * SC + c];
                }
                y[c] = interp_lin(frac, y0, y1);
                if (in.has_gain) y[c] = __fmul_rn(y[c], in.gain);   // upstream audio::gain, rounded separately (gain.rs:187-189)
            }
            // channel conversion as mixer.rs:1027-1078
            float v[2];
            if (SC == OC) { v[0] = y[0]; v[1] = y[1]; }
            else if (SC == 1 && OC == 2) { v[0] = y[0]; v[1] = y[0]; }
            else { v[0] = __fmul_rn(__fadd_rn(y[0], y[1]), 0.5f); v[1] = 0.0f; }
#pragma unroll
            for (int c = 0; c < OC; ++c) acc[f * OC + c] = is_base ? v[c] : __fadd_rn(acc[f * OC + c], v[c]);
        }
    }
}

template <int OC, int ITERS>  // output channels (1 | 2); ITERS = ceil(F / 1024)
__global__ void __launch_bounds__(CH_THREADS, 4) k_chain(const OpHeader *__restrict__ hdr, const skgpu_chain_group *__restrict__ groups,
                                                      const skgpu_chain_input *__restrict__ inputs, const uint8_t *__restrict__ present,
                                                      const float *__restrict__ gains, SlotTables st, uint8_t *__restrict__ arena,
                                                      const uint32_t *__restrict__ tick, uint64_t bank_stride, uint32_t F,
                                                      uint64_t results_off, uint32_t kb, uint32_t buf_floats, uint32_t nstages) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[CH_MAX_STAGES], bar_empty[CH_MAX_STAGES];
    __shared__ __align__(16) ChainStage s_stage[CH_MAX_STAGES];
    __shared__ ChainIn s_res[CH_MAX_INPUTS];     // producer scratch: resolved inputs of the session being prepared
    __shared__ uint8_t s_order[CH_MAX_INPUTS];

    // dynamic smem per stage: kb staging buffers [(16 + N) * ch floats], then kb x 2 phase tables
    const size_t buf_bytes = (((size_t)kb * buf_floats * 4u) + 15u) & ~(size_t)15u;
    const size_t stage_bytes = buf_bytes + (size_t)kb * 2u * sizeof(SmemPhase);
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
        const uint32_t parity = tick[0] & 1u;
        uint32_t stage = 0, ephase = 1;  // waiting on parity 1 of a fresh mbarrier returns immediately
        // The descriptors of a session hang off each other (group -> inputs -> slot records / gains / presence): three
        // dependent global round trips. They are software-pipelined three sessions deep in registers, so by the time a
        // session is processed everything it needs was requested a full iteration earlier.
        struct InPre { skgpu_chain_input in; SlotRec rec; float gain; uint32_t pres; };
        const skgpu_chain_group grp_none = {0, 0, 0, SKGPU_NO_GAIN, (uint16_t)OC, 0};
        auto load_grp = [&](uint32_t g) -> skgpu_chain_group { return (g < n_groups) ? groups[g] : grp_none; };
        auto load_in = [&](const skgpu_chain_group &g, uint32_t j) -> skgpu_chain_input {
            skgpu_chain_input z = {0, 0, SKGPU_NO_GAIN, 0, 0};
            return (j < min(g.n_inputs, (uint32_t)CH_MAX_INPUTS)) ? inputs[g.first_input + j] : z;
        };
        auto load_pre = [&](const skgpu_chain_group &g, const skgpu_chain_input &in, uint32_t j, InPre &o) {
            o.in = in;
            if (j < min(g.n_inputs, (uint32_t)CH_MAX_INPUTS)) {
                o.rec = st.rec[in.slot];
                o.gain = (in.gain_idx != SKGPU_NO_GAIN) ? gains[in.gain_idx] : 1.0f;
                o.pres = present ? (uint32_t)(present[g.first_input + j] != 0) : 1u;
            }
        };
        const uint32_t gstep = gridDim.x;
        skgpu_chain_group grp0 = load_grp(blockIdx.x), grp1 = load_grp(blockIdx.x + gstep), grp2 = load_grp(blockIdx.x + 2u * gstep);
        skgpu_chain_input in0 = load_in(grp0, lane), in1 = load_in(grp1, lane);
        InPre pre0;
        load_pre(grp0, in0, lane, pre0);
        for (uint32_t g_i = blockIdx.x; g_i < n_groups; g_i += gstep) {
            const skgpu_chain_group grp = grp0;
            const InPre pre = pre0;
            // prefetch for the next three sessions
            const skgpu_chain_group grp3 = load_grp(g_i + 3u * gstep);
            const skgpu_chain_input in2 = load_in(grp2, lane);
            InPre pre1;
            load_pre(grp1, in1, lane, pre1);
            const uint32_t K = min(grp.n_inputs, (uint32_t)CH_MAX_INPUTS);
            // ---- resolve inputs (one lane per input): slot record, emission decision, carry bookkeeping, results
            for (uint32_t j = lane; j < K; j += 32u) {
                const uint32_t gi = grp.first_input + j;
                InPre cur = pre;
                if (j >= 32u) {  // sessions with more than 32 inputs: the rest is fetched on demand
                    cur.in = inputs[gi];
                    load_pre(grp, cur.in, j, cur);
                }
                const skgpu_chain_input in = cur.in;
                SlotRec *recp = st.rec + in.slot;
                const SlotRec rec = cur.rec;
                ChainIn r;
                r.slot = in.slot;
                r.N = rec.chunk;
                r.ch = rec.channels;
                r.t = rec.t_ratio;
                r.count = rec.chunk_count;   // k_phase already counted the current chunk
                r.carry = rec.carry;
                const uint32_t pres = cur.pres;
                r.has_gain = in.gain_idx != SKGPU_NO_GAIN;
                r.gain = cur.gain;
                r.unique = (in.flags & SKGPU_MIX_IN_UNIQUE) ? 1u : 0u;
                r.cur_g = reinterpret_cast<const float *>(arena + in.in_off + (uint64_t)parity * bank_stride);
                r.prev_g = reinterpret_cast<const float *>(arena + in.in_off + (uint64_t)(1u - parity) * bank_stride);
                r.hist_g = st.hist + (size_t)in.slot * 16u * st.max_channels;
                r.par_cur = (r.count - 1u) & 1u;
                r.par_prev = r.count & 1u;   // == (count - 2) & 1
                r.n_cur = (pres && r.count >= 1u) ? rec.n_out[r.par_cur] : 0u;
                r.n_prev = (r.count >= 2u) ? rec.n_out[r.par_prev] : 0u;
                r.np_prev = (r.count >= 2u) ? rec.n_prefix[r.par_prev] : (uint16_t)0;
                r.nr_prev = (r.count >= 2u) ? rec.n_runs[r.par_prev] : (uint16_t)0;
                r.np_cur = rec.n_prefix[r.par_cur];
                r.nr_cur = rec.n_runs[r.par_cur];
                uint32_t status = 0;
                r.emit = 0;
                uint32_t new_carry = r.carry;
                if (pres) {
                    if (r.carry > r.n_prev) status |= 4u;                   // carried frames span more than one chunk: unsupported
                    const uint32_t avail = r.carry + r.n_cur;
                    r.emit = (avail >= F && !(status & 4u)) ? 1u : 0u;      // a whole F-frame packet is ready (resampler.rs:425-428)
                    new_carry = r.emit ? avail - F : avail;
                    if (new_carry > r.n_cur) status |= 1u;                  // backlog: a second packet is pending
                    if (r.count >= 1u && ((rec.overflow >> r.par_cur) & 1u)) status |= 2u;
                    recp->carry = new_carry;
                }
                s_res[j] = r;
                skgpu_chain_result res;
                res.emitted = r.emit;
                res.status = status;
                reinterpret_cast<skgpu_chain_result *>(arena + results_off)[gi] = res;
                // a present input that emits nothing still retires its previous chunk: advance the history here
                // (emitting inputs are handled by the consumers, after the TMA read of the old history)
                if (pres && !r.emit && r.count >= 2u) {
                    for (uint32_t e = 0; e < 16u * r.ch; ++e) r.hist_g[e] = r.prev_g[(size_t)(r.N - 16u) * r.ch + e];
                }
            }
            grp0 = grp1; grp1 = grp2; grp2 = grp3;
            in0 = in1; in1 = in2;
            pre0 = pre1;
            __syncwarp();
            // ---- summation order over the inputs that deliver a packet: base selection + swap_remove (mixer.rs:960-980)
            uint32_t m = 0, has_base = 0;
            if (lane == 0) {
                int base = -1, base_unique = -1;
                for (uint32_t j = 0; j < K; ++j) {
                    if (!s_res[j].emit) continue;
                    if (s_res[j].ch == (uint32_t)OC) {  // packet already has the output shape (F frames x OC channels)
                        const int u = (int)s_res[j].unique;
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
            // ---- one pipeline item per batch of <= kb inputs (an empty session still produces one item: silence)
            const uint32_t n_batches = (m + kb - 1u) / kb + (m == 0u ? 1u : 0u);
            for (uint32_t b = 0; b < n_batches; ++b) {
                const uint32_t b0 = b * kb;
                const uint32_t nb = min(kb, m - min(m, b0));
                mbar_wait(&bar_empty[stage], ephase);
                ChainStage *S = &s_stage[stage];
                uint8_t *sm = smem_raw + (size_t)stage * stage_bytes;
                float *s_buf = reinterpret_cast<float *>(sm);
                SmemPhase *s_tab = reinterpret_cast<SmemPhase *>(sm + buf_bytes);
                // chunks whose byte size is not a multiple of 16 (e.g. mono 882 frames) cannot use the bulk copy:
                // the producer warp copies them itself
                for (uint32_t q = 0; q < nb; ++q) {
                    const ChainIn &in = s_res[s_order[b0 + q]];
                    const uint32_t cb = (in.count >= 2u) ? in.N * in.ch * 4u : 0u;
                    if (cb & 15u) {
                        float *dst = s_buf + (size_t)q * buf_floats + 16u * in.ch;
                        for (uint32_t e = lane; e < in.N * in.ch; e += 32u) dst[e] = in.prev_g[e];
                        const uint32_t head = min((uint32_t)CH_HEAD, in.N) * in.ch;
                        for (uint32_t e = lane; e < head; e += 32u) dst[in.N * in.ch + e] = in.cur_g[e];
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    S->out_off = grp.out_off;
                    S->has_master = grp.gain_idx != SKGPU_NO_GAIN;
                    S->master_gain = S->has_master ? gains[grp.gain_idx] : 1.0f;
                    S->flags = grp.flags;
                    S->nb = nb;
                    S->first = (b == 0u);
                    S->last = (b + 1u == n_batches);
                    S->has_base = has_base;
                    S->stop = 0;
                    uint32_t total = 0;
                    for (uint32_t q = 0; q < nb; ++q) {
                        const ChainIn in = s_res[s_order[b0 + q]];
                        S->in[q] = in;
                        SmemPhase *Tp = &s_tab[q * 2u], *Tc = &s_tab[q * 2u + 1u];
                        const uint32_t npp = in.np_prev, nrp = in.nr_prev, npc = in.np_cur, nrc = in.nr_cur;
                        Tp->n_out = in.n_prev; Tp->n_prefix = npp; Tp->n_runs = nrp; Tp->overflow = 0;
                        Tc->n_out = in.n_cur; Tc->n_prefix = npc; Tc->n_runs = nrc; Tc->overflow = 0;
                        const uint32_t cb = (in.count >= 2u) ? in.N * in.ch * 4u : 0u;
                        const uint32_t b_hist = 16u * in.ch * 4u;
                        const uint32_t b_chunk = (cb & 15u) ? 0u : cb + min((uint32_t)CH_HEAD, in.N) * in.ch * 4u;  // + head of the current chunk
                        const uint32_t b_pp = (npp * 8u + 15u) & ~15u, b_pr = (nrp * 24u + 15u) & ~15u;
                        const uint32_t b_cp = (npc * 8u + 15u) & ~15u, b_cr = (nrc * 24u + 15u) & ~15u;
                        total += b_hist + b_chunk + b_pp + b_pr + b_cp + b_cr;
                    }
                    mbar_expect_tx(&bar_full[stage], total);
                    for (uint32_t q = 0; q < nb; ++q) {
                        const ChainIn &in = S->in[q];
                        float *dst = s_buf + (size_t)q * buf_floats;
                        const SkPhaseTable *tab = st.tab + (size_t)in.slot * 2u;
                        SmemPhase *Tp = &s_tab[q * 2u], *Tc = &s_tab[q * 2u + 1u];
                        const uint32_t cb = (in.count >= 2u) ? in.N * in.ch * 4u : 0u;
                        tma_bulk_g2s(dst, in.hist_g, 16u * in.ch * 4u, &bar_full[stage]);
                        if (cb && !(cb & 15u)) tma_bulk_g2s(dst + 16u * in.ch, in.prev_g, cb, &bar_full[stage]);
                        if (!(cb & 15u)) tma_bulk_g2s(dst + (16u + (cb ? in.N : 0u)) * in.ch, in.cur_g, min((uint32_t)CH_HEAD, in.N) * in.ch * 4u, &bar_full[stage]);
                        const uint32_t b_pp = (Tp->n_prefix * 8u + 15u) & ~15u, b_pr = (Tp->n_runs * 24u + 15u) & ~15u;
                        const uint32_t b_cp = (Tc->n_prefix * 8u + 15u) & ~15u, b_cr = (Tc->n_runs * 24u + 15u) & ~15u;
                        if (b_pp) tma_bulk_g2s(Tp->prefix, tab[in.par_prev].prefix, b_pp, &bar_full[stage]);
                        if (b_pr) tma_bulk_g2s(Tp->runs, tab[in.par_prev].runs, b_pr, &bar_full[stage]);
                        if (b_cp) tma_bulk_g2s(Tc->prefix, tab[in.par_cur].prefix, b_cp, &bar_full[stage]);
                        if (b_cr) tma_bulk_g2s(Tc->runs, tab[in.par_cur].runs, b_cr, &bar_full[stage]);
                    }
                }
                __syncwarp();
                if (++stage == nstages) { stage = 0; ephase ^= 1u; }
            }
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
        if (S->stop) break;
        const uint8_t *sm = smem_raw + (size_t)stage * stage_bytes;
        const float *s_buf = reinterpret_cast<const float *>(sm);
        const SmemPhase *s_tab = reinterpret_cast<const SmemPhase *>(sm + buf_bytes);
        if (S->first) {
#pragma unroll
            for (int it = 0; it < ITERS; ++it)
#pragma unroll
                for (int e = 0; e < CH_FPT * OC; ++e) acc[it][e] = 0.0f;  // vec![0.0f32; output_size] when there is no base frame
        }
        const uint32_t nb = S->nb;
        const bool first_base = S->has_base && S->first;
        for (uint32_t q = 0; q < nb; ++q) {
            const ChainIn *ip = &S->in[q];
            ChainInRegs in;
            in.cur_g = ip->cur_g;
            in.t = ip->t;
            in.gain = ip->gain;
            in.has_gain = ip->has_gain;
            in.sc = ip->ch;
            in.carry = ip->carry;
            in.n_prev = ip->n_prev;
            in.n_cur = max(ip->n_cur, 1u);
            in.NA = (ip->count >= 2u) ? ip->N : 0u;
            const float *A = s_buf + (size_t)q * buf_floats;
            const SmemPhase *Tp = &s_tab[q * 2u], *Tc = &s_tab[q * 2u + 1u];
            const bool is_base = first_base && q == 0u;
#pragma unroll
            for (int it = 0; it < ITERS; ++it) {
                const uint32_t j0 = (ct + it * CH_CONSUMERS) * CH_FPT;
                if (j0 < F) {
                    if (in.sc == 2u) chain_input<OC, 2>(in, A, Tp, Tc, F, j0, acc[it], is_base);
                    else chain_input<OC, 1>(in, A, Tp, Tc, F, j0, acc[it], is_base);
                }
            }
        }
        if (S->last) {
            // ---- epilogue: master gain, then clip + s16 pack (or f32)
            const bool has_master = S->has_master != 0;
            const float mg = S->master_gain;
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
                if (S->flags & SKGPU_MIX_OUT_S16) {
                    uint16_t *o = reinterpret_cast<uint16_t *>(arena + S->out_off) + (size_t)j0 * OC;
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
                    float *o = reinterpret_cast<float *>(arena + S->out_off) + (size_t)j0 * OC;
                    if (nfr == CH_FPT) {
                        stg_stream_f4(reinterpret_cast<float4 *>(o), make_float4(a[0], a[1], a[2], a[3]));
                        if (OC == 2) stg_stream_f4(reinterpret_cast<float4 *>(o) + 1, make_float4(a[4 % (CH_FPT * OC)], a[5 % (CH_FPT * OC)], a[6 % (CH_FPT * OC)], a[7 % (CH_FPT * OC)]));
                    } else {
                        for (uint32_t e = 0; e < nfr * OC; ++e) o[e] = a[e];
                    }
                }
            }
        }
        // ---- history of every input of the batch := last 16 frames of its previous chunk (now retired); the bulk
        // read of the old history has completed (the full barrier flipped), so overwriting it is safe
        for (uint32_t q = 0; q < nb; ++q) {
            const ChainIn &in = S->in[q];
            if (in.count >= 2u && ct < 16u * in.ch) in.hist_g[ct] = s_buf[(size_t)q * buf_floats + (size_t)in.N * in.ch + ct];
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar_empty[stage])) : "memory");
        if (++stage == nstages) { stage = 0; fphase ^= 1u; }
    }
}

// advances the device-side tick counter (bank parity) once per tick; part of the captured graph
__global__ void k_tick_advance(uint32_t *tick) { tick[0] += 1u; }

}  // namespace skgpu
