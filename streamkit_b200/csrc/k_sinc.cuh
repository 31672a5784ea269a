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
// k_resample_sinc: one CTA per stream-chunk. The input window every output needs -- [history | chunk], 7.6 KB for a 20 ms
// stereo chunk -- is staged in shared memory by two TMA bulk copies; each thread then owns whole output frames and walks the two
// tap rows of its sub-phase with 128-bit loads (rows are L floats, 16-byte aligned; the (O + 1) x L table is 66 KB and lives in
// L1 / L2: every CTA of the launch reads the same table). The dot products are sequential f32 fma chains in ascending tap order,
// one chain per channel and row: bit-identical to a plain C fmaf loop. 256 fma per stereo frame at L = 64: the kernel is bound by
// the FMA / LSU pipes, not by HBM (16 fma per byte moved).
#pragma once
#include "common.cuh"
#include "k_resample.cuh"

namespace skgpu {

constexpr int SINC_THREADS = 128;

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
        const float4 *t0 = reinterpret_cast<const float4 *>(taps + (size_t)p * L);
        const float4 *t1 = t0 + L / 4u;
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

}  // namespace skgpu
