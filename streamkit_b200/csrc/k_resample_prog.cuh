// k_resample_prog.cuh -- the standalone audio::resampler kernel (BASELINE config #4): one CTA per stream-chunk, the chunk,
// its 16-frame history and the frame program k_phase_prog built (chain_prog.h) staged by three TMA bulk copies; the
// blocks are evaluated by the same straight-line / segment-walk code the fused chain uses (k_chain.cuh chain_block).
// `one` is 1.0f passed as a launch parameter: see add2 in k_chain.cuh.
#pragma once
#include "k_chain.cuh"

namespace skgpu {

template <int C>  // 1 | 2
__global__ void __launch_bounds__(RSP_THREADS) k_resample_prog(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items,
                                                              SlotTables st, uint8_t *__restrict__ arena, uint32_t smem_frames, ChainProgDims pd, float one) {
    extern __shared__ __align__(16) uint8_t smem_raw[];   // [program (prog_cap) | 16 history frames | chunk]
    __shared__ __align__(8) uint64_t bar;

    const uint32_t i = blockIdx.x;
    if (i >= hdr->count) return;
    const skgpu_rs_item it = items[i];
    const uint32_t slot = it.slot;
    const SlotRec rec = st.rec[slot];
    const uint32_t N = rec.chunk;
    const bool to_fifo = (it.flags & SKGPU_RS_TO_FIFO) != 0;
    const uint32_t par = (rec.chunk_count - 1u) & 1u;   // the chunk k_phase_prog just processed
    const uint32_t n_total = par ? rec.n_out[1] : rec.n_out[0];
    const uint32_t n_exp = par ? rec.n_prefix[1] : rec.n_prefix[0];
    const bool prog_ok = ((rec.overflow >> par) & 1u) == 0u;

    const uint32_t prog_cap = skc_prog_cap(pd);
    float *buf = reinterpret_cast<float *>(smem_raw + prog_cap);  // [(16 + N) * C]: history then chunk, interleaved
    float *hist_g = st.hist + (size_t)slot * 16u * st.max_channels;
    const float *in_g = reinterpret_cast<const float *>(arena + it.in_off);
    const uint32_t hist_bytes = 16u * C * 4u;
    const uint32_t in_bytes = N * C * 4u;
    const uint32_t prog_bytes = (skc_exp_off(pd) + n_exp * 8u + 15u) & ~15u;
    const bool tma_ok = ((in_bytes & 15u) == 0) && ((((uintptr_t)in_g) & 15u) == 0);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&bar, prog_bytes + hist_bytes + (tma_ok ? in_bytes : 0u));
        tma_bulk_g2s(smem_raw, slot_side(st, slot, par), prog_bytes, &bar);
        tma_bulk_g2s(buf, hist_g, hist_bytes, &bar);
        if (tma_ok) tma_bulk_g2s(buf + 16u * C, in_g, in_bytes, &bar);
    }
    if (!tma_ok)
        for (uint32_t s = threadIdx.x; s < N * C; s += RSP_THREADS) buf[16u * C + s] = in_g[s];
    __syncthreads();
    mbar_wait(&bar, 0);

    const uint32_t n_out = prog_ok ? (to_fifo ? n_total : min(n_total, it.out_cap_frames)) : 0u;
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

    const uint32_t prog = smem_u32(smem_raw), segs = prog + skc_seg_off(pd);
    const uint32_t a_hist = smem_u32(buf), a_chunk = a_hist + 16u * C * 4u;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    // each warp owns a run of CONSECUTIVE 32-frame blocks: a block inside the FAST run the warp is already in costs one
    // DADD + the split + the interpolation (chain_block's straight-line path, k_chain.cuh); only blocks that hold a run
    // boundary, explicit frames or the chunk's tail take the per-lane segment walk
    const uint32_t nblk = (n_out + 31u) >> 5, per = (nblk + (RSP_THREADS / 32) - 1u) / (RSP_THREADS / 32);
    const uint32_t b0 = warp * per, b1 = min(nblk, b0 + per);
    const unsigned long long one2 = pack2(one, one);
    RunCache rc;
    rc.j0 = 0; rc.j1 = 0; rc.himask = 0; rc.sh = 0; rc.kc = 0; rc.xl = 0.0; rc.dl32 = 0.0;
    if (b0 < b1) {   // enter the run that covers the warp's first block, if one FAST run does
        uint32_t e;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e) : "r"(prog + b0 * 2u));
        const uint32_t sa = segs + (e & 0xFFu) * 32u;
        uint32_t ljj, lhimask, laux, lsh;
        asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+16];" : "=r"(ljj), "=r"(lhimask), "=r"(laux), "=r"(lsh) : "r"(sa));
        if ((e & 0xFFu) == (e >> 8) && lhimask > SKC_KIND_SLOW && (ljj >> 16) >= b0 * 32u + 32u) {
            double x0, dl;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x0), "=d"(dl) : "r"(sa));
            rc.j0 = ljj & 0xFFFFu; rc.j1 = ljj >> 16; rc.himask = lhimask; rc.sh = lsh; rc.kc = a_chunk - laux;
            rc.dl32 = __dmul_rn(dl, 32.0);
            rc.xl = __fma_rn((double)((int)(b0 * 32u + lane) - (int)rc.j0 - 32), dl, x0);   // one block earlier on the run's lattice (exact)
        }
    }
#pragma unroll 2
    for (uint32_t b = b0; b < b1; ++b) {
        const unsigned long long v = chain_block<C, C>(rc, b, b + 1u < b1, prog, segs, a_hist, a_chunk, n_out, lane, 1.0f, one2);
        const uint32_t j = b * 32u + lane;
        if (j < n_out) {
            const uint32_t of = to_fifo ? (uint32_t)((fifo_w + j) & fifo_mask) : j;
            float o0, o1;
            unpack2(v, o0, o1);
            if (C == 2) stg_stream_f2(reinterpret_cast<float2 *>(out_g) + of, make_float2(o0, o1));
            else out_g[of] = o0;
        }
    }
    // new history = buffer frames [N, N+16): the last 16 frames of (history ++ chunk)
    float hv = 0.0f;
    const bool hw = threadIdx.x < 16u * C;
    if (hw) hv = buf[N * C + threadIdx.x];
    __syncthreads();  // everyone is done with the staged buffer; the old history in HBM was only read by the bulk copy
    if (hw) hist_g[threadIdx.x] = hv;
    if (to_fifo && threadIdx.x == 0) st.fifo_w[slot] = fifo_w + n_total;
}

}  // namespace skgpu
