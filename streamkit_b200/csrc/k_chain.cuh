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
// One CTA per session. Inputs are staged (history ++ previous chunk) by TMA bulk copies, up to `kb` inputs per
// batch; each thread owns 4 consecutive output frames and adds the inputs sequentially in the reference's
// summation order (f32 addition is not associative).
#pragma once
#include "common.cuh"

namespace skgpu {

constexpr int CH_THREADS = 256;
constexpr int CH_MAX_INPUTS = 64;   // inputs per session
constexpr int CH_FPT = 4;           // output frames per thread per iteration
constexpr int CH_MAX_ITERS = 3;     // F <= 3072 (largest valid output_frame_size is 2880)

struct ChainIn {            // per-tick view of one input of the session (shared memory)
    const float *prev_g;    // previous chunk (other bank)
    const float *cur_g;     // current chunk
    float *hist_g;          // st.hist of the slot: 16 frames before the previous chunk
    double t;
    float gain;
    uint32_t has_gain;
    uint32_t slot, N, ch;
    uint32_t carry, n_prev, n_cur, count;
    uint32_t present, emit, unique, status, new_carry;
};

template <int OC>
__device__ __forceinline__ void chain_accumulate(float *acc, const float *y, uint32_t sc, float gain, bool has_gain, bool is_base) {
    // y: one input frame (sc channels) -> one output frame (OC channels); channel conversion as mixer.rs:1027-1078,
    // upstream audio::gain applied per sample first (rounded separately, gain.rs:187-189)
    float v[2];
    if (sc == (uint32_t)OC) {
#pragma unroll
        for (int c = 0; c < OC; ++c) v[c] = has_gain ? __fmul_rn(y[c], gain) : y[c];
    } else if (sc == 1 && OC == 2) {
        const float m = has_gain ? __fmul_rn(y[0], gain) : y[0];
        v[0] = m; v[1] = m;
    } else {  // sc == 2 && OC == 1
        const float l = has_gain ? __fmul_rn(y[0], gain) : y[0];
        const float r = has_gain ? __fmul_rn(y[1], gain) : y[1];
        v[0] = __fmul_rn(__fadd_rn(l, r), 0.5f);
    }
#pragma unroll
    for (int c = 0; c < OC; ++c) acc[c] = is_base ? v[c] : __fadd_rn(acc[c], v[c]);
}

template <int OC>  // output channels: 1 or 2
__global__ void __launch_bounds__(CH_THREADS) k_chain(const OpHeader *__restrict__ hdr, const skgpu_chain_group *__restrict__ groups,
                                                      const skgpu_chain_input *__restrict__ inputs, const uint8_t *__restrict__ present,
                                                      const float *__restrict__ gains, SlotTables st, uint8_t *__restrict__ arena,
                                                      const uint32_t *__restrict__ tick, uint64_t bank_stride, uint32_t F,
                                                      uint64_t results_off, uint32_t kb, uint32_t buf_floats) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ ChainIn s_in[CH_MAX_INPUTS];
    __shared__ uint8_t s_order[CH_MAX_INPUTS];
    __shared__ uint32_t s_m, s_has_base;

    const uint32_t g_i = blockIdx.x;
    if (g_i >= hdr->count) return;
    const skgpu_chain_group grp = groups[g_i];
    const uint32_t K = min(grp.n_inputs, (uint32_t)CH_MAX_INPUTS);
    const uint32_t parity = tick[0] & 1u;

    // dynamic smem: kb staging buffers [(16 + N) * ch floats] then kb x 2 phase tables
    float *s_buf = reinterpret_cast<float *>(smem_raw);
    SmemPhase *s_tab = reinterpret_cast<SmemPhase *>(smem_raw + (((size_t)kb * buf_floats * 4u + 15u) & ~(size_t)15u));

    // ---- resolve the session's inputs (one thread per input): slot state, emission decision, carry bookkeeping
    if (threadIdx.x < K) {
        const uint32_t gi = grp.first_input + threadIdx.x;
        const skgpu_chain_input in = inputs[gi];
        ChainIn r;
        r.slot = in.slot;
        r.N = st.chunk[in.slot];
        r.ch = st.channels[in.slot];
        r.t = st.t_ratio[in.slot];
        r.count = st.chunk_count[in.slot];   // k_phase already counted the current chunk
        r.carry = st.carry[in.slot];
        r.present = present ? (present[gi] != 0) : 1u;
        r.has_gain = in.gain_idx != SKGPU_NO_GAIN;
        r.gain = r.has_gain ? gains[in.gain_idx] : 1.0f;
        r.unique = (in.flags & SKGPU_MIX_IN_UNIQUE) ? 1u : 0u;
        r.cur_g = reinterpret_cast<const float *>(arena + in.in_off + (uint64_t)parity * bank_stride);
        r.prev_g = reinterpret_cast<const float *>(arena + in.in_off + (uint64_t)(1u - parity) * bank_stride);
        r.hist_g = st.hist + (size_t)in.slot * 16u * st.max_channels;
        const SkPhaseTable *tab = st.tab + (size_t)in.slot * 2u;
        r.n_cur = (r.present && r.count >= 1u) ? tab[(r.count - 1u) & 1u].n_out : 0u;
        r.n_prev = (r.count >= 2u) ? tab[(r.count - 2u) & 1u].n_out : 0u;
        r.status = 0;
        r.emit = 0;
        r.new_carry = r.carry;
        if (r.present) {
            if (r.carry > r.n_prev) r.status |= 4u;                 // carried frames span more than one chunk: unsupported
            const uint32_t avail = r.carry + r.n_cur;
            r.emit = (avail >= F && !(r.status & 4u)) ? 1u : 0u;    // a whole F-frame packet is ready (resampler.rs:425-428)
            r.new_carry = r.emit ? avail - F : avail;
            if (r.new_carry > r.n_cur) r.status |= 1u;               // backlog: a second packet is pending / carry spans two chunks
        }
        s_in[threadIdx.x] = r;
    }
    __syncthreads();
    // ---- summation order over the inputs that deliver a packet: base selection + swap_remove (mixer.rs:960-980)
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        int base = -1, base_unique = -1;
        for (uint32_t j = 0; j < K; ++j) {
            if (!s_in[j].emit) continue;
            if (s_in[j].ch == (uint32_t)OC) {  // packet already has the output shape (F frames x OC channels)
                const int u = (int)s_in[j].unique;
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
        s_m = m;
        s_has_base = (base >= 0) ? 1u : 0u;
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const uint32_t m = s_m;
    const bool has_base = s_has_base != 0;

    float acc[CH_MAX_ITERS][CH_FPT * OC];
#pragma unroll
    for (int it = 0; it < CH_MAX_ITERS; ++it)
#pragma unroll
        for (int e = 0; e < CH_FPT * OC; ++e) acc[it][e] = 0.0f;  // vec![0.0f32; output_size] when there is no base frame

    uint32_t phase_bit = 0;
    for (uint32_t b0 = 0; b0 < m; b0 += kb) {
        const uint32_t nb = min(kb, m - b0);
        if (b0 > 0) {
            __syncthreads();  // previous batch fully consumed before its buffers are overwritten
            if (threadIdx.x == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic reads -> async writes
        }
        // ---- stage (history ++ previous chunk) of every input of the batch with TMA bulk copies
        // (chunks whose byte size is not a multiple of 16, e.g. mono 882 frames, are copied cooperatively instead)
        if (threadIdx.x == 0) {
            uint32_t total = 0;
            for (uint32_t q = 0; q < nb; ++q) {
                const ChainIn &in = s_in[s_order[b0 + q]];
                const uint32_t cb = (in.count >= 2u) ? in.N * in.ch * 4u : 0u;
                total += 16u * in.ch * 4u + ((cb & 15u) ? 0u : cb);
            }
            mbar_expect_tx(&bar, total);
            for (uint32_t q = 0; q < nb; ++q) {
                const ChainIn &in = s_in[s_order[b0 + q]];
                float *dst = s_buf + (size_t)q * buf_floats;
                const uint32_t cb = (in.count >= 2u) ? in.N * in.ch * 4u : 0u;
                tma_bulk_g2s(dst, in.hist_g, 16u * in.ch * 4u, &bar);
                if (cb && !(cb & 15u)) tma_bulk_g2s(dst + 16u * in.ch, in.prev_g, cb, &bar);
            }
        }
        for (uint32_t q = 0; q < nb; ++q) {
            const ChainIn &in = s_in[s_order[b0 + q]];
            const uint32_t cb = (in.count >= 2u) ? in.N * in.ch * 4u : 0u;
            if (cb & 15u) {
                float *dst = s_buf + (size_t)q * buf_floats + 16u * in.ch;
                for (uint32_t e = threadIdx.x; e < in.N * in.ch; e += CH_THREADS) dst[e] = in.prev_g[e];
            }
        }
        // ---- phase tables (previous and current chunk) of the batch -> smem, overlapping the bulk copies
        for (uint32_t q = 0; q < nb; ++q) {
            const ChainIn &in = s_in[s_order[b0 + q]];
            const SkPhaseTable *tab = st.tab + (size_t)in.slot * 2u;
            if (in.count >= 2u) load_phase_table(&s_tab[q * 2u], tab + ((in.count - 2u) & 1u), threadIdx.x, CH_THREADS);
            load_phase_table(&s_tab[q * 2u + 1u], tab + ((in.count - 1u) & 1u), threadIdx.x, CH_THREADS);
        }
        __syncthreads();
        mbar_wait(&bar, phase_bit);
        phase_bit ^= 1u;

        // ---- interpolate + gain + ordered accumulate
        for (uint32_t q = 0; q < nb; ++q) {
            const ChainIn in = s_in[s_order[b0 + q]];
            const float *A = s_buf + (size_t)q * buf_floats;     // [16 history | previous chunk (N frames, if any)]
            const uint32_t NA = (in.count >= 2u) ? in.N : 0u;    // frames of the previous chunk present in A
            const SmemPhase *Tp = &s_tab[q * 2u], *Tc = &s_tab[q * 2u + 1u];
            const bool is_base = has_base && (b0 + q == 0u);
            const uint32_t sc = in.ch;
#pragma unroll
            for (int it = 0; it < CH_MAX_ITERS; ++it) {
                const uint32_t j0 = (threadIdx.x + it * CH_THREADS) * CH_FPT;
                if (j0 >= F) break;
#pragma unroll
                for (int f = 0; f < CH_FPT; ++f) {
                    const uint32_t j = j0 + f;
                    if (j >= F) break;
                    float y[2];
                    uint32_t p;
                    float frac;
                    if (j < in.carry) {
                        // frame produced by the PREVIOUS chunk: recompute it from (history ++ previous chunk)
                        const uint32_t k = in.n_prev - in.carry + j;
                        phase_split(phase_eval_smem(Tp, in.t, k), p, frac);
                        for (uint32_t c = 0; c < sc; ++c) y[c] = interp_lin(frac, A[p * sc + c], A[(p + 1u) * sc + c]);
                    } else {
                        // frame produced by the CURRENT chunk: its history is the tail of the previous chunk (in A),
                        // positions beyond it are read from the current chunk in HBM (a handful in steady state)
                        const uint32_t k = j - in.carry;
                        phase_split(phase_eval_smem(Tc, in.t, k), p, frac);
                        for (uint32_t c = 0; c < sc; ++c) {
                            const float y0 = (p < 16u) ? A[(NA + p) * sc + c] : in.cur_g[(size_t)(p - 16u) * sc + c];
                            const float y1 = (p + 1u < 16u) ? A[(NA + p + 1u) * sc + c] : in.cur_g[(size_t)(p + 1u - 16u) * sc + c];
                            y[c] = interp_lin(frac, y0, y1);
                        }
                    }
                    chain_accumulate<OC>(&acc[it][f * OC], y, sc, in.gain, in.has_gain != 0, is_base);
                }
            }
        }
    }

    // ---- epilogue: master gain, then clip + s16 pack (or f32)
    const bool has_master = grp.gain_idx != SKGPU_NO_GAIN;
    const float mg = has_master ? gains[grp.gain_idx] : 1.0f;
#pragma unroll
    for (int it = 0; it < CH_MAX_ITERS; ++it) {
        const uint32_t j0 = (threadIdx.x + it * CH_THREADS) * CH_FPT;
        if (j0 >= F) break;
        float *a = acc[it];
        if (has_master) {
#pragma unroll
            for (int e = 0; e < CH_FPT * OC; ++e) a[e] = __fmul_rn(a[e], mg);
        }
        const uint32_t nfr = min((uint32_t)CH_FPT, F - j0);
        if (grp.flags & SKGPU_MIX_OUT_S16) {
            uint16_t *o = reinterpret_cast<uint16_t *>(arena + grp.out_off) + (size_t)j0 * OC;
            if (nfr == CH_FPT && OC == 2 && ((((uintptr_t)o) & 15u) == 0)) {
                stg_stream_u4(reinterpret_cast<uint4 *>(o), make_uint4(pack_s16x2(a[0], a[1]), pack_s16x2(a[2], a[3]),
                                                                      pack_s16x2(a[4 % (CH_FPT * OC)], a[5 % (CH_FPT * OC)]),
                                                                      pack_s16x2(a[6 % (CH_FPT * OC)], a[7 % (CH_FPT * OC)])));
            } else if (nfr == CH_FPT && OC == 1 && ((((uintptr_t)o) & 7u) == 0)) {
                stg_stream_u2(reinterpret_cast<uint2 *>(o), make_uint2(pack_s16x2(a[0], a[1]), pack_s16x2(a[2], a[3])));
            } else {
                for (uint32_t e = 0; e < nfr * OC; ++e) o[e] = (uint16_t)f32_to_s16_bits(a[e]);
            }
        } else {
            float *o = reinterpret_cast<float *>(arena + grp.out_off) + (size_t)j0 * OC;
            for (uint32_t e = 0; e < nfr * OC; ++e) o[e] = a[e];
        }
    }

    // ---- state: carry, history (16 frames before the chunk that now becomes "previous"), per-input results
    __syncthreads();  // all TMA reads of st.hist have completed (every thread passed the last mbar_wait)
    for (uint32_t j = 0; j < K; ++j) {
        const ChainIn &in = s_in[j];
        if (in.present && in.count >= 2u) {
            // new history = last 16 frames of the previous chunk (N >= 16 is validated by the host)
            const uint32_t n = 16u * in.ch;
            if (threadIdx.x < n) in.hist_g[threadIdx.x] = in.prev_g[(size_t)(in.N - 16u) * in.ch + threadIdx.x];
        }
    }
    if (threadIdx.x < K) {
        const ChainIn &in = s_in[threadIdx.x];
        if (in.present) st.carry[in.slot] = in.new_carry;
        uint32_t status = in.status;
        const SkPhaseTable *tab = st.tab + (size_t)in.slot * 2u;
        if (in.present && in.count >= 1u && tab[(in.count - 1u) & 1u].overflow) status |= 2u;
        skgpu_chain_result res;
        res.emitted = in.emit;
        res.status = status;
        reinterpret_cast<skgpu_chain_result *>(arena + results_off)[grp.first_input + threadIdx.x] = res;
    }
}

// advances the device-side tick counter (bank parity) once per tick; part of the captured graph
__global__ void k_tick_advance(uint32_t *tick) { tick[0] += 1u; }

}  // namespace skgpu
