// k_sinc.cuh -- windowed-sinc polyphase resampler mode (interp = sinc), the resampler BASELINE.json's north star names.
//
// The reference resamples with rubato FastFixedIn / Linear (resampler.rs:232-238); that stays the parity default
// (k_resample.cuh, k_chain.cuh). This mode has no reference implementation: its specification is restated in
// include/skgpu_batch.h (skgpu_ctx_set_sinc) and implemented independently by the test suite's CPU checkers (C and numpy).
//
//   y(x) = (1 - q) * sum_n in[fl - L/2 + 1 + n] * T[p][n]  +  q * sum_n in[...] * T[p + 1][n]
//          fl = floor(x), p = floor((x - fl) * O), q = f32((x - fl) * O - p)
//
// The positions x follow the same f64 recurrence idx += 1/ratio as the linear mode, so k_phase (phase_runs.h) produces them
// exactly and data-independently. Per stream the state in HBM is last_index (f64) and the last L + 8 input frames.
//
// Two kernels. k_resample_sinc_tiled (further down; persistent, tap table in shared memory, same-phase output tiles) runs ops whose
// streams share one tap table; k_resample_sinc is the simple form and the fallback:
// k_resample_sinc: one CTA per stream-chunk. The input window every output needs -- [history | chunk], 7.6 KB for a 20 ms
// stereo chunk -- is staged in shared memory by two TMA bulk copies; each thread then owns whole output frames and walks the two
// tap rows of its sub-phase with 128-bit loads (rows are L + 4 floats apart, 16-byte aligned; the (O + 1) x (L + 4) table is 70 KB and lives in
// L1 / L2: every CTA of the launch reads the same table). The dot products are sequential f32 fma chains in ascending tap order,
// one chain per channel and row: bit-identical to a plain C fmaf loop. 256 fma per stereo frame at L = 64: the kernel is bound by
// the FMA / LSU pipes, not by HBM (16 fma per byte moved).
#pragma once
#include "common.cuh"
#include "k_resample.cuh"
#include "k_chain.cuh"

namespace skgpu {

constexpr int SINC_THREADS = 128;
constexpr uint32_t SINC_ROW_PAD = 4u;     // tap rows are L + 4 floats apart (HBM and shared memory): rows p and p + 8 fall on different banks

template <int C>
__global__ void __launch_bounds__(SINC_THREADS) k_resample_sinc(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items, SlotTables st,
                                                                uint8_t *__restrict__ arena) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) SmemPhase s_tab;

    const uint32_t i = blockIdx.x;
    if (i >= hdr->count) return;
    const skgpu_rs_item it = items[i];
    const uint32_t slot = it.slot;
    const SlotRec rec = st.rec[slot];
    const uint32_t N = rec.chunk, L = st.sinc_L, O = st.sinc_O, H = st.sinc_H;
    const double t = rec.t_ratio;
    const float *__restrict__ taps = st.sinc_tabs[(rec.flags >> 8) & 0xFFu];

    float *buf = reinterpret_cast<float *>(smem_raw);                         // [(H + N) * C]: history then chunk, interleaved
    float *hist_g = st.sinc_hist + (size_t)slot * H * st.max_channels;
    const float *in_g = reinterpret_cast<const float *>(arena + it.in_off);
    const uint32_t hist_bytes = H * C * 4u, in_bytes = N * C * 4u;
    const bool tma_ok = ((in_bytes & 15u) == 0) && ((((uintptr_t)in_g) & 15u) == 0) && ((hist_bytes & 15u) == 0);
    if (tma_ok) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&bar, hist_bytes + in_bytes);
            tma_bulk_g2s(buf, hist_g, hist_bytes, &bar);
            tma_bulk_g2s(buf + H * C, in_g, in_bytes, &bar);
        }
    } else {
        for (uint32_t s = threadIdx.x; s < H * C; s += SINC_THREADS) buf[s] = hist_g[s];
        for (uint32_t s = threadIdx.x; s < N * C; s += SINC_THREADS) buf[H * C + s] = in_g[s];
    }
    const uint32_t par = (rec.chunk_count - 1u) & 1u;                         // the chunk k_phase just processed
    load_phase_table(&s_tab, slot_tab(st, slot, par), rec.n_out[par], rec.n_prefix[par], rec.n_runs[par], threadIdx.x, SINC_THREADS);
    __syncthreads();
    if (tma_ok) mbar_wait(&bar, 0);

    const uint32_t n_out = min(s_tab.n_out, it.out_cap_frames);
    float *out_g = reinterpret_cast<float *>(arena + it.out_off);
    const double dO = (double)O;
    for (uint32_t k = threadIdx.x; k < n_out; k += SINC_THREADS) {
        const double x = phase_eval_smem(&s_tab, t, k);
        const int fl = __double2int_rd(x);
        const double fo = __dmul_rn(__dsub_rn(x, (double)fl), dO);
        int p = __double2int_rd(fo);
        p = min(p, (int)O - 1);
        const float q = __double2float_rn(__dsub_rn(fo, (double)p));
        const float *w = buf + (size_t)(fl - (int)(L / 2u) + 1 + (int)H) * C;  // first input frame under the taps
        const float4 *t0 = reinterpret_cast<const float4 *>(taps + (size_t)p * (L + SINC_ROW_PAD));
        const float4 *t1 = t0 + (L + SINC_ROW_PAD) / 4u;
        float y0[C], y1[C];
#pragma unroll
        for (int c = 0; c < C; ++c) y0[c] = y1[c] = 0.0f;
        for (uint32_t n4 = 0; n4 < L / 4u; ++n4) {
            const float4 a = __ldg(t0 + n4), b = __ldg(t1 + n4);
            const float ta[4] = {a.x, a.y, a.z, a.w}, tb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float xin = w[(n4 * 4u + (uint32_t)m) * C + c];
                    y0[c] = __fmaf_rn(xin, ta[m], y0[c]);
                    y1[c] = __fmaf_rn(xin, tb[m], y1[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < C; ++c) out_g[(size_t)k * C + c] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, q), y0[c]), __fmul_rn(q, y1[c]));
    }
    // new history = the last H frames of [history | chunk]
    __syncthreads();
    for (uint32_t s = threadIdx.x; s < H * C; s += SINC_THREADS) hist_g[s] = buf[(size_t)N * C + s];
}

// ------------------------------------------------------------------------------------------------------------------
// k_resample_sinc_tiled: the same arithmetic, organised for the FMA pipe instead of for simplicity.
//
// 256 fma per stereo output frame make this mode compute / shared-memory bound (16 fma per HBM byte), so the kernel
//   * is PERSISTENT (one 512-thread CTA per SM) and stages the op's whole tap table -- (O + 1) rows of L + 4 floats, 70 KB
//     at L = 64, O = 256 -- in shared memory ONCE with one bulk copy (cp.async.bulk, completion on an mbarrier);
//   * walks the op's streams in passes of G streams; a pass's inputs -- per stream the phase table k_phase just wrote, the
//     L + 8 frame history and the chunk -- arrive by bulk copies into a 2-stage ring (the next pass loads while this one
//     computes), issued by G lanes of warp 0 (one stream each) against one mbarrier per stage;
//   * gives a thread up to SINC_RA outputs that share ONE pair of tap rows: for a rational ratio the sub-phase repeats
//     every `pe` outputs (out_rate / gcd, rounded up to a multiple >= 32; host: slot_configure), so outputs k, k + pe,
//     k + 2 pe ... read the same rows p, p + 1 of the table and the tap quads are loaded once for all of them; consecutive
//     lanes own consecutive k, so their input windows are consecutive frames (conflict-free 64-bit loads) and their
//     stores coalesce. Every output still derives its own (floor, p, q) from ITS phase: where the f64 recurrence puts two
//     "same-phase" outputs into different rows (a position within rounding distance of a row boundary) the thread falls
//     back to one-output-at-a-time evaluation. Stereo frames are multiplied as packed f32x2 (FFMA2): two independent
//     fma.rn, bit-identical to the scalar chain.
// Shared-memory bandwidth bounds it: per output and tap 8 B of input (not shared) + 8 B of taps / SINC_RA.
constexpr int SINCT_THREADS = 512;
constexpr int SINC_RA = 6;
constexpr int SINC_GMAX = 8;
constexpr uint32_t SINC_SLOW_CAP = 1024;

struct __align__(16) SincStream {   // one stream of a pass: written by its loading lane, read by everyone after the stage's barrier
    float *out_g;
    float *hist_g;
    const float *in_g;
    double t;
    uint32_t n_out, pe, n_items, N;
    uint32_t coop, pad[3];          // coop: the chunk is not 16-byte copyable, threads copy it
};

struct SincDims {                   // host: rs_sinc_dims
    uint32_t G;                     // streams per pass
    uint32_t stream_bytes;          // [SkPhaseTable | history | chunk] per stream, 16-byte multiple
    uint32_t tab_bytes;             // (O + 1) * (L + 4) * 4
};

__device__ __forceinline__ double sinc_phase_eval(const SkPhaseTable *T, double t, uint32_t k) {
    if (k < T->n_prefix) return T->prefix[k];
    uint32_t r = T->n_runs - 1u;
    while (r > 0u && T->runs[r].k_a > k) --r;
    const SkRun rn = T->runs[r];
    if (k < rn.k_e) return __fma_rn((double)(k - rn.k_a), rn.delta, rn.x_a);
    return __dadd_rn(__fma_rn((double)(rn.k_e - 1u - rn.k_a), rn.delta, rn.x_a), t);  // gap element
}

__device__ __forceinline__ unsigned long long sinc_lds64(uint32_t a) {
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ float4 sinc_lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned long long sinc_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float sinc_mix(float q, float y0, float y1) { return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, q), y0), __fmul_rn(q, y1)); }

// one output frame on its own tap rows (shared-memory addresses)
template <int C>
__device__ __forceinline__ void sinc_one(uint32_t w, uint32_t row, uint32_t LS4, uint32_t L, float q, float *out) {
    if (C == 2) {
        unsigned long long y0 = 0ull, y1 = 0ull;
#pragma unroll 2
        for (uint32_t n4 = 0; n4 < L / 4u; ++n4) {
            const float4 a = sinc_lds128(row + n4 * 16u), b = sinc_lds128(row + LS4 + n4 * 16u);
            const unsigned long long A[4] = {pack2(a.x, a.x), pack2(a.y, a.y), pack2(a.z, a.z), pack2(a.w, a.w)};
            const unsigned long long B[4] = {pack2(b.x, b.x), pack2(b.y, b.y), pack2(b.z, b.z), pack2(b.w, b.w)};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const unsigned long long xf = sinc_lds64(w + (n4 * 4u + (uint32_t)m) * 8u);
                y0 = sinc_fma2(xf, A[m], y0);
                y1 = sinc_fma2(xf, B[m], y1);
            }
        }
        float l0, r0, l1, r1;
        unpack2(y0, l0, r0);
        unpack2(y1, l1, r1);
        out[0] = sinc_mix(q, l0, l1);
        out[C - 1] = sinc_mix(q, r0, r1);
    } else {
        float y0 = 0.0f, y1 = 0.0f;
#pragma unroll 2
        for (uint32_t n4 = 0; n4 < L / 4u; ++n4) {
            const float4 a = sinc_lds128(row + n4 * 16u), b = sinc_lds128(row + LS4 + n4 * 16u);
            const float ta[4] = {a.x, a.y, a.z, a.w}, tb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
                const float xin = lds_f32(w + (n4 * 4u + (uint32_t)m) * 4u);
                y0 = __fmaf_rn(xin, ta[m], y0);
                y1 = __fmaf_rn(xin, tb[m], y1);
            }
        }
        out[0] = sinc_mix(q, y0, y1);
    }
}

// output k of a staged stream, evaluated on its own (phase -> rows -> two dot products -> store)
template <int C>
__device__ __forceinline__ void sinc_single(const SincStream &m, const SkPhaseTable *T, uint32_t a_buf, uint32_t a_tab, uint32_t k, uint32_t L, uint32_t O,
                                            uint32_t H, uint32_t LS4) {
    if (k >= m.n_out) return;
    const double x = sinc_phase_eval(T, m.t, k);
    const int fl = __double2int_rd(x);
    const double fo = __dmul_rn(__dsub_rn(x, (double)fl), (double)O);
    int p = __double2int_rd(fo);
    p = min(p, (int)O - 1);
    const float q = __double2float_rn(__dsub_rn(fo, (double)p));
    float o[C];
    sinc_one<C>(a_buf + (uint32_t)(fl - (int)(L / 2u) + 1 + (int)H) * (C * 4u), a_tab + (uint32_t)p * LS4, LS4, L, q, o);
    if (C == 2) stg_stream_f2(reinterpret_cast<float2 *>(m.out_g) + k, make_float2(o[0], o[C - 1]));
    else m.out_g[k] = o[0];
}

template <int C>
__global__ void __launch_bounds__(SINCT_THREADS, 1) k_resample_sinc_tiled(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items, SlotTables st,
                                                                          uint8_t *__restrict__ arena, const float *__restrict__ taps, SincDims dm) {
    extern __shared__ __align__(16) uint8_t smem_raw[];   // [tap table | stage 0: G streams | stage 1: G streams]
    __shared__ __align__(8) uint64_t bar_tab, bar_full[2];
    __shared__ SincStream meta[2][SINC_GMAX];
    __shared__ uint32_t slow_n, slow_q[SINC_SLOW_CAP];   // work items whose outputs do not share their tap rows (see below)

    const uint32_t count = hdr->count, G = dm.G;
    const uint32_t n_pass = (count + G - 1u) / G;
    if (blockIdx.x >= n_pass) return;
    const uint32_t L = st.sinc_L, O = st.sinc_O, H = st.sinc_H, LS4 = (L + SINC_ROW_PAD) * 4u;
    const uint32_t tid = threadIdx.x;
    const uint32_t a_tab = smem_u32(smem_raw), a_stage = a_tab + dm.tab_bytes;
    uint8_t *stage_p = smem_raw + dm.tab_bytes;
    constexpr uint32_t PT_BYTES = (uint32_t)((sizeof(SkPhaseTable) + 15u) & ~15u);

    if (tid == 0) {
        mbar_init(&bar_tab, 1);
        mbar_init(&bar_full[0], G);
        mbar_init(&bar_full[1], G);
        mbar_fence_init();
        mbar_expect_tx(&bar_tab, dm.tab_bytes);
        tma_bulk_g2s(smem_raw, taps, dm.tab_bytes, &bar_tab);
    }
    __syncthreads();

    // loading lane g of warp 0: stream (q * G + g) of pass q into stage s
    auto load_stream = [&](uint32_t q, uint32_t s, uint32_t g) {
        const uint32_t i = q * G + g;
        SincStream m;
        m.out_g = nullptr; m.hist_g = nullptr; m.in_g = nullptr; m.t = 0.0; m.n_out = 0; m.pe = 1; m.n_items = 0; m.N = 0; m.coop = 0;
        uint32_t bytes = 0;
        uint8_t *dst = stage_p + (size_t)(s * G + g) * dm.stream_bytes;
        if (q < n_pass && i < count) {
            const skgpu_rs_item it = items[i];
            const SlotRec rec = st.rec[it.slot];
            const uint32_t par = (rec.chunk_count - 1u) & 1u;   // the chunk k_phase just processed
            m.N = rec.chunk;
            m.t = rec.t_ratio;
            m.n_out = min(par ? rec.n_out[1] : rec.n_out[0], it.out_cap_frames);
            m.pe = max(rec.carry, 1u);
            const uint32_t per_class = (m.n_out + m.pe - 1u) / m.pe;
            m.n_items = m.pe * ((per_class + SINC_RA - 1u) / SINC_RA);
            m.out_g = reinterpret_cast<float *>(arena + it.out_off);
            m.hist_g = st.sinc_hist + (size_t)it.slot * H * st.max_channels;
            m.in_g = reinterpret_cast<const float *>(arena + it.in_off);
            const uint32_t hist_bytes = H * C * 4u, in_bytes = m.N * C * 4u;
            m.coop = (((in_bytes & 15u) == 0) && ((((uintptr_t)m.in_g) & 15u) == 0)) ? 0u : 1u;
            bytes = PT_BYTES + hist_bytes + (m.coop ? 0u : in_bytes);
            meta[s][g] = m;
            mbar_expect_tx(&bar_full[s], bytes);
            tma_bulk_g2s(dst, slot_tab(st, it.slot, par), PT_BYTES, &bar_full[s]);
            tma_bulk_g2s(dst + PT_BYTES, m.hist_g, hist_bytes, &bar_full[s]);
            if (!m.coop) tma_bulk_g2s(dst + PT_BYTES + hist_bytes, m.in_g, in_bytes, &bar_full[s]);
        } else {
            meta[s][g] = m;
            mbar_expect_tx(&bar_full[s], 0u);
        }
    };

    if (tid < G) load_stream(blockIdx.x, 0u, tid);
    mbar_wait(&bar_tab, 0);

    uint32_t it_n = 0;
    for (uint32_t q = blockIdx.x; q < n_pass; q += gridDim.x, ++it_n) {
        const uint32_t s = it_n & 1u;
        if (tid < G && q + gridDim.x < n_pass) load_stream(q + gridDim.x, s ^ 1u, tid);   // stage s ^ 1 was released by the barrier that ended the previous pass
        if (tid == 0) slow_n = 0u;       // ordered before its first use by the barrier below (and after its last use by the pass-ending one)
        mbar_wait(&bar_full[s], (it_n >> 1) & 1u);
        __syncthreads();

        // chunks that could not be bulk-copied (unaligned / odd sizes)
        bool any_coop = false;
        for (uint32_t g = 0; g < G; ++g) {
            const SincStream &m = meta[s][g];
            if (m.coop) {
                any_coop = true;
                float *d = reinterpret_cast<float *>(stage_p + (size_t)(s * G + g) * dm.stream_bytes + PT_BYTES) + H * C;
                for (uint32_t e = tid; e < m.N * C; e += SINCT_THREADS) d[e] = m.in_g[e];
            }
        }
        if (any_coop) __syncthreads();   // block-uniform

        uint32_t total = 0;
        for (uint32_t g = 0; g < G; ++g) total += meta[s][g].n_items;
        for (uint32_t w = tid; w < total; w += SINCT_THREADS) {
            uint32_t g = 0, local = w;
            while (local >= meta[s][g].n_items) { local -= meta[s][g].n_items; ++g; }
            const SincStream &m = meta[s][g];
            const uint32_t pe = m.pe, n_out = m.n_out;
            const uint32_t cls = local % pe, grp = local / pe;
            const uint32_t k0 = cls + pe * grp * SINC_RA;
            if (k0 >= n_out) continue;
            const uint32_t a_str = a_stage + (s * G + g) * dm.stream_bytes;
            const SkPhaseTable *T = reinterpret_cast<const SkPhaseTable *>(stage_p + (size_t)(s * G + g) * dm.stream_bytes);
            const uint32_t a_buf = a_str + PT_BYTES;                  // [history | chunk], interleaved
            uint32_t wa[SINC_RA], p0 = 0;
            float qa[SINC_RA];
            bool same = true;
#pragma unroll
            for (int a = 0; a < SINC_RA; ++a) {
                const uint32_t k = k0 + (uint32_t)a * pe;
                const double x = sinc_phase_eval(T, m.t, k < n_out ? k : k0);
                const int fl = __double2int_rd(x);
                const double fo = __dmul_rn(__dsub_rn(x, (double)fl), (double)O);
                int p = __double2int_rd(fo);
                p = min(p, (int)O - 1);
                qa[a] = __double2float_rn(__dsub_rn(fo, (double)p));
                if (a == 0) p0 = (uint32_t)p;
                wa[a] = a_buf + (uint32_t)(fl - (int)(L / 2u) + 1 + (int)H) * (C * 4u);   // first input frame under the taps
                same = same && ((uint32_t)p == p0);
            }
            float *out_g = m.out_g;
            if (same) {
                const uint32_t row = a_tab + p0 * LS4;
                if (C == 2) {
                    unsigned long long y0[SINC_RA], y1[SINC_RA];
#pragma unroll
                    for (int a = 0; a < SINC_RA; ++a) y0[a] = y1[a] = 0ull;
#pragma unroll 2
                    for (uint32_t n4 = 0; n4 < L / 4u; ++n4) {
                        const float4 ta = sinc_lds128(row + n4 * 16u), tb = sinc_lds128(row + LS4 + n4 * 16u);
                        const unsigned long long A[4] = {pack2(ta.x, ta.x), pack2(ta.y, ta.y), pack2(ta.z, ta.z), pack2(ta.w, ta.w)};
                        const unsigned long long B[4] = {pack2(tb.x, tb.x), pack2(tb.y, tb.y), pack2(tb.z, tb.z), pack2(tb.w, tb.w)};
#pragma unroll
                        for (int a = 0; a < SINC_RA; ++a) {
#pragma unroll
                            for (int mm = 0; mm < 4; ++mm) {
                                const unsigned long long xf = sinc_lds64(wa[a] + (n4 * 4u + (uint32_t)mm) * 8u);
                                y0[a] = sinc_fma2(xf, A[mm], y0[a]);
                                y1[a] = sinc_fma2(xf, B[mm], y1[a]);
                            }
                        }
                    }
#pragma unroll
                    for (int a = 0; a < SINC_RA; ++a) {
                        const uint32_t k = k0 + (uint32_t)a * pe;
                        if (k < n_out) {
                            float l0, r0, l1, r1;
                            unpack2(y0[a], l0, r0);
                            unpack2(y1[a], l1, r1);
                            stg_stream_f2(reinterpret_cast<float2 *>(out_g) + k, make_float2(sinc_mix(qa[a], l0, l1), sinc_mix(qa[a], r0, r1)));
                        }
                    }
                } else {
                    float y0[SINC_RA], y1[SINC_RA];
#pragma unroll
                    for (int a = 0; a < SINC_RA; ++a) y0[a] = y1[a] = 0.0f;
#pragma unroll 2
                    for (uint32_t n4 = 0; n4 < L / 4u; ++n4) {
                        const float4 ta = sinc_lds128(row + n4 * 16u), tb = sinc_lds128(row + LS4 + n4 * 16u);
                        const float A[4] = {ta.x, ta.y, ta.z, ta.w}, B[4] = {tb.x, tb.y, tb.z, tb.w};
#pragma unroll
                        for (int a = 0; a < SINC_RA; ++a) {
#pragma unroll
                            for (int mm = 0; mm < 4; ++mm) {
                                const float xin = lds_f32(wa[a] + (n4 * 4u + (uint32_t)mm) * 4u);
                                y0[a] = __fmaf_rn(xin, A[mm], y0[a]);
                                y1[a] = __fmaf_rn(xin, B[mm], y1[a]);
                            }
                        }
                    }
#pragma unroll
                    for (int a = 0; a < SINC_RA; ++a) {
                        const uint32_t k = k0 + (uint32_t)a * pe;
                        if (k < n_out) out_g[k] = sinc_mix(qa[a], y0[a], y1[a]);
                    }
                }
            } else {
                // The outputs of this item sit within rounding distance of a tap-row boundary ((x - floor x) * O is an integer in
                // exact arithmetic: 1 class in 5 at 160/147), so the f64 recurrence puts them on either side of it. A lane that
                // evaluated them one by one here would stall its whole warp: queue the item, all threads share the queue below.
                const uint32_t at = atomicAdd(&slow_n, 1u);
                if (at < SINC_SLOW_CAP) {
                    slow_q[at] = (g << 24) | local;
                } else {
#pragma unroll 1
                    for (uint32_t a = 0; a < (uint32_t)SINC_RA; ++a) sinc_single<C>(m, T, a_buf, a_tab, k0 + a * pe, L, O, H, LS4);
                }
            }
        }
        __syncthreads();
        {   // queued items, one OUTPUT per thread
            const uint32_t ns = min(slow_n, SINC_SLOW_CAP);
            for (uint32_t u = tid; u < ns * SINC_RA; u += SINCT_THREADS) {
                const uint32_t e = slow_q[u % ns], a = u / ns, g = e >> 24, local = e & 0xFFFFFFu;
                const SincStream &m = meta[s][g];
                const uint32_t k = (local % m.pe) + m.pe * ((local / m.pe) * SINC_RA + a);
                const SkPhaseTable *T = reinterpret_cast<const SkPhaseTable *>(stage_p + (size_t)(s * G + g) * dm.stream_bytes);
                sinc_single<C>(m, T, a_stage + (s * G + g) * dm.stream_bytes + PT_BYTES, a_tab, k, L, O, H, LS4);
            }
        }
        // new history = the last H frames of [history | chunk]
        for (uint32_t g = 0; g < G; ++g) {
            const SincStream &m = meta[s][g];
            if (!m.hist_g) continue;
            const float *b = reinterpret_cast<const float *>(stage_p + (size_t)(s * G + g) * dm.stream_bytes + PT_BYTES);
            for (uint32_t e = tid; e < H * C; e += SINCT_THREADS) m.hist_g[e] = b[(size_t)m.N * C + e];
        }
        __syncthreads();   // the stage (and its meta) may be refilled
    }
}

}  // namespace skgpu
