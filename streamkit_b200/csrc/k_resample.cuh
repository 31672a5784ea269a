// k_resample.cuh -- batched rubato FastFixedIn<f32>/Linear (resampler.rs:397-417 + rubato process_into_buffer)
//
//   k_phase     one THREAD per stream-chunk: the f64 phase recurrence -> compact phase table (phase_runs.h).
//               32 streams share a warp, so the inherently sequential chain is SIMD across streams; thanks to
//               the per-binade jump it is ~50 dependent steps per chunk. Data independent (reads slot state only).
//   k_resample  one CTA per stream-chunk: history (HBM state) + chunk staged into shared memory by TMA bulk
//               copies, phase table staged next to it, then every thread interpolates 16 bytes worth of
//               consecutive output frames per iteration and stores them with one 128-bit store.
#pragma once
#include "chain_prog.h"
#include "common.cuh"

namespace skgpu {

constexpr int PHASE_THREADS = 64;

__global__ void __launch_bounds__(PHASE_THREADS) k_phase(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items, SlotTables st,
                                                         uint8_t *__restrict__ arena, uint64_t results_off) {
    const uint32_t i = blockIdx.x * PHASE_THREADS + threadIdx.x;
    if (i >= hdr->count) return;
    const uint32_t slot = items[i].slot;
    SlotRec *rec = st.rec + slot;
    const uint32_t c = rec->chunk_count;
    const uint32_t par = c & 1u;
    SkPhaseTable *T = slot_tab(st, slot, par);
    double idx_end;
    const uint32_t n = sk_phase_table(rec->last_index, rec->t_ratio, rec->end_idx, T, &idx_end);
    rec->last_index = __dsub_rn(idx_end, (double)rec->chunk);  // self.last_index = idx - chunk_size as f64
    rec->chunk_count = c + 1u;
    rec->n_out[par] = n;
    rec->n_prefix[par] = (uint16_t)T->n_prefix;
    rec->n_runs[par] = (uint16_t)T->n_runs;
    rec->overflow = (rec->overflow & ~(1u << par)) | ((T->overflow ? 1u : 0u) << par);
    const skgpu_rs_item *it = items + i;
    skgpu_rs_result res;
    res.out_frames = n;
    res.status = 0;
    if (!(it->flags & SKGPU_RS_TO_FIFO) && n > it->out_cap_frames) {
        res.out_frames = it->out_cap_frames;
        res.status = 1;
    }
    if (T->overflow) res.status = 2;
    reinterpret_cast<skgpu_rs_result *>(arena + results_off)[i] = res;
}

// ------------------------------------------------------------------ interpolation

constexpr int RS_THREADS = 128;

// one output frame (C channels) from a staged buffer `buf` = [16 history frames | chunk], interleaved
template <int C>
__device__ __forceinline__ void interp_frame(const float *buf, uint32_t p, float frac, float *o) {
    if (C == 2) {
        const float2 y0 = *reinterpret_cast<const float2 *>(buf + 2u * p);
        const float2 y1 = *reinterpret_cast<const float2 *>(buf + 2u * p + 2u);
        o[0] = interp_lin(frac, y0.x, y1.x);
        o[1] = interp_lin(frac, y0.y, y1.y);
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = interp_lin(frac, buf[p * C + c], buf[(p + 1u) * C + c]);
    }
}

template <int C>  // C = 1, 2 specialised; 0 = runtime channel count
__global__ void __launch_bounds__(RS_THREADS) k_resample(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items,
                                                         SlotTables st, uint8_t *__restrict__ arena, uint32_t smem_frames) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(16) SmemPhase s_tab;

    const uint32_t i = blockIdx.x;
    if (i >= hdr->count) return;
    const skgpu_rs_item it = items[i];
    const uint32_t slot = it.slot;
    const SlotRec rec = st.rec[slot];
    const uint32_t ch = (C > 0) ? (uint32_t)C : rec.channels;
    const uint32_t N = rec.chunk;
    const double t = rec.t_ratio;
    const bool to_fifo = (it.flags & SKGPU_RS_TO_FIFO) != 0;

    float *buf = reinterpret_cast<float *>(smem_raw);  // [(16 + N) * ch]: history then chunk, interleaved
    float *hist_g = st.hist + (size_t)slot * 16u * st.max_channels;
    const float *in_g = reinterpret_cast<const float *>(arena + it.in_off);
    const uint32_t hist_bytes = 16u * ch * 4u;
    const uint32_t in_bytes = N * ch * 4u;
    const bool staged = (N + 16u) <= smem_frames;  // host sizes smem for the op's largest chunk
    const bool tma_ok = staged && ((in_bytes & 15u) == 0) && ((((uintptr_t)in_g) & 15u) == 0) && ((hist_bytes & 15u) == 0);

    if (staged) {
        if (tma_ok) {
            if (threadIdx.x == 0) {
                mbar_init(&bar, 1);
                mbar_fence_init();
                mbar_expect_tx(&bar, hist_bytes + in_bytes);
                tma_bulk_g2s(buf, hist_g, hist_bytes, &bar);
                tma_bulk_g2s(buf + 16u * ch, in_g, in_bytes, &bar);
            }
        } else {
            for (uint32_t s = threadIdx.x; s < 16u * ch; s += RS_THREADS) buf[s] = hist_g[s];
            for (uint32_t s = threadIdx.x; s < N * ch; s += RS_THREADS) buf[16u * ch + s] = in_g[s];
        }
    }
    // phase table of the chunk k_phase just processed (chunk_count was already advanced) -> smem, overlapping the bulk copy
    const uint32_t par = (rec.chunk_count - 1u) & 1u;
    load_phase_table(&s_tab, slot_tab(st, slot, par), rec.n_out[par], rec.n_prefix[par], rec.n_runs[par], threadIdx.x, RS_THREADS);
    __syncthreads();
    if (staged && tma_ok) mbar_wait(&bar, 0);

    const uint32_t n_total = s_tab.n_out;
    const uint32_t n_out = to_fifo ? n_total : min(n_total, it.out_cap_frames);

    float *out_g;
    unsigned long long fifo_w = 0;
    uint32_t fifo_mask = 0;
    if (to_fifo) {
        out_g = st.fifo + (size_t)slot * st.fifo_frames * st.max_channels;
        fifo_w = st.fifo_w[slot];
        fifo_mask = st.fifo_frames - 1u;
    } else {
        out_g = reinterpret_cast<float *>(arena + it.out_off);
    }

    if (C == 1 || C == 2) {
        constexpr int CC = (C == 0) ? 1 : C;
        constexpr int FPT = 4;  // consecutive frames per thread per iteration: one phase-run lookup, 16*CC bytes of output
        for (uint32_t k0 = threadIdx.x * FPT; k0 < n_out; k0 += RS_THREADS * FPT) {
            const uint32_t nfr = min((uint32_t)FPT, n_out - k0);
            double x[4];
            phase_eval4(&s_tab, t, k0, nfr, x);
            float o[4 * CC];
#pragma unroll
            for (int f = 0; f < FPT; ++f) {
                uint32_t p;
                float frac;
                phase_split(x[f], p, frac);
                if (staged) {
                    interp_frame<CC>(buf, p, frac, o + f * CC);
                } else {
#pragma unroll
                    for (int c = 0; c < CC; ++c) {
                        const float y0 = (p < 16u) ? hist_g[p * CC + c] : in_g[(p - 16u) * CC + c];
                        const float y1 = (p + 1u < 16u) ? hist_g[(p + 1u) * CC + c] : in_g[(p + 1u - 16u) * CC + c];
                        o[f * CC + c] = interp_lin(frac, y0, y1);
                    }
                }
            }
            const uint32_t of0 = to_fifo ? (uint32_t)((fifo_w + k0) & fifo_mask) : k0;
            float *dst = out_g + (size_t)of0 * CC;
            const bool contiguous = (nfr == FPT) && (!to_fifo || of0 + FPT <= st.fifo_frames);
            if (contiguous && ((((uintptr_t)dst) & 15u) == 0)) {
                stg_stream_f4(reinterpret_cast<float4 *>(dst), make_float4(o[0], o[1], o[2], o[3]));
                if (CC == 2) stg_stream_f4(reinterpret_cast<float4 *>(dst) + 1, make_float4(o[4 % (4 * CC)], o[5 % (4 * CC)], o[6 % (4 * CC)], o[7 % (4 * CC)]));
            } else {
#pragma unroll
                for (int f = 0; f < FPT; ++f) {
                    if ((uint32_t)f < nfr) {
                        const uint32_t of = to_fifo ? (uint32_t)((fifo_w + k0 + f) & fifo_mask) : (k0 + f);
#pragma unroll
                        for (int c = 0; c < CC; ++c) out_g[(size_t)of * CC + c] = o[f * CC + c];
                    }
                }
            }
        }
    } else {
        for (uint32_t k = threadIdx.x; k < n_out; k += RS_THREADS) {
            uint32_t p;
            float frac;
            phase_split(phase_eval_smem(&s_tab, t, k), p, frac);
            const uint32_t of = to_fifo ? (uint32_t)((fifo_w + k) & fifo_mask) : k;
            for (uint32_t c = 0; c < ch; ++c) {
                float y0, y1;
                if (staged) {
                    y0 = buf[p * ch + c];
                    y1 = buf[(p + 1u) * ch + c];
                } else {
                    y0 = (p < 16u) ? hist_g[p * ch + c] : in_g[(p - 16u) * ch + c];
                    y1 = (p + 1u < 16u) ? hist_g[(p + 1u) * ch + c] : in_g[(p + 1u - 16u) * ch + c];
                }
                out_g[(size_t)of * ch + c] = interp_lin(frac, y0, y1);
            }
        }
    }
    // new history = buffer frames [N, N+16): the last 16 frames of (history ++ chunk). 16*ch <= 128 threads.
    float hv = 0.0f;
    const bool hw = threadIdx.x < 16u * ch;
    if (hw) {
        const uint32_t f = N + threadIdx.x / ch, c = threadIdx.x % ch;  // frame index into history++chunk
        if (staged) hv = buf[f * ch + c];
        else hv = (f < 16u) ? hist_g[f * ch + c] : in_g[(f - 16u) * ch + c];
    }
    __syncthreads();  // everyone is done reading the old history (the non-staged path reads it from HBM)
    if (hw) hist_g[threadIdx.x] = hv;
    if (to_fifo && threadIdx.x == 0) st.fifo_w[slot] = fifo_w + n_total;
}

// ------------------------------------------------------------------ program-driven variant (the fast path)
// Same idea as the fused chain (k_chain.cuh): the phase kernel turns the chunk's phase recurrence into a FRAME PROGRAM
// (chain_prog.h: run segments, explicit frames, a per-32-frame block map; here without a tail: the record covers exactly
// the chunk's n_out outputs) so that the interpolation kernel searches nothing: a warp owns 32-frame blocks, lane l owns
// output frame (block start + l) -- neighbouring lanes read neighbouring input frames (conflict-free 8-byte shared-memory
// loads) and a warp stores 32 consecutive output frames per instruction.
// Used when the op's streams are mono / stereo, fit the staging buffer and their programs fit a side record.

constexpr int RSP_THREADS = 64;   // small CTAs: more stream-chunks in flight per SM (the kernel is bound by bytes in flight)

__global__ void __launch_bounds__(PHASE_THREADS, 14) k_phase_prog(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items,
                                                                  SlotTables st, uint8_t *__restrict__ arena, uint64_t results_off, ChainProgDims pd) {
    const uint32_t i = blockIdx.x * PHASE_THREADS + threadIdx.x;
    if (i >= hdr->count) return;
    const skgpu_rs_item it = items[i];
    const uint32_t slot = it.slot;
    SlotRec *recp = st.rec + slot;
    SlotRec rec = *recp;
    const uint32_t c = rec.chunk_count, par = c & 1u;
    SkcStream sb;
    uint32_t np, nr, ovf, n_seg = 0, n_exp = 0;
    double idx_end;
    sb.begin(slot_side(st, slot, par), pd, pd.nblk * 32u, rec.channels * 4u, 0u, nullptr, 0u, 0u, 0u, rec.chunk, 0u, rec.t_ratio);
    const uint32_t n = sk_phase_stream(rec.last_index, rec.t_ratio, rec.end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
    const uint32_t st_prog = sb.finish_open(n, &n_seg, &n_exp);
    rec.last_index = __dsub_rn(idx_end, (double)rec.chunk);  // self.last_index = idx - chunk_size as f64
    rec.chunk_count = c + 1u;
    if (par) { rec.n_out[1] = n; rec.n_prefix[1] = (uint16_t)n_exp; rec.n_runs[1] = (uint16_t)n_seg; }
    else { rec.n_out[0] = n; rec.n_prefix[0] = (uint16_t)n_exp; rec.n_runs[0] = (uint16_t)n_seg; }
    rec.overflow = (rec.overflow & ~(1u << par)) | ((st_prog ? 1u : 0u) << par);
    *recp = rec;
    skgpu_rs_result res;
    res.out_frames = n;
    res.status = 0;
    if (!(it.flags & SKGPU_RS_TO_FIFO) && n > it.out_cap_frames) {
        res.out_frames = it.out_cap_frames;
        res.status = 1;
    }
    if (st_prog) res.status = 2;
    reinterpret_cast<skgpu_rs_result *>(arena + results_off)[i] = res;
}

}  // namespace skgpu
