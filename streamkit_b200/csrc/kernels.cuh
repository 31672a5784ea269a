// kernels.cuh -- hand-written sm_100a kernels for StreamKit's PCM DSP hot path.
//
// All kernels are HBM-bound streaming kernels (<= 3 flop/byte): no tensor cores. What matters is 128-bit
// coalesced access, enough bytes in flight per SM, shared-memory/TMA staging where the access pattern is
// data dependent (resampler), and exact IEEE arithmetic: the TU is compiled with -fmad=false and the
// parity-critical expressions additionally use __fmul_rn/__fadd_rn so nothing is ever contracted
// (Rust, the reference's language, never contracts a*b+c).
//
//   k_convert        gain / f32->s16 / s16->f32 over frame segments      (gain.rs:184-190)
//   k_phase          per-stream rubato phase recurrence -> run table      (rubato FastFixedIn, resampler.rs:404-407)
//   k_resample       linear interpolation from TMA-staged chunk + history (rubato interp_lin; resampler.rs:397-417)
//   k_mix            ordered N-input sum, channel conversion, epilogue    (mixer.rs:944-1013, :1027-1078)
//   k_fifo_commit    advances device re-framing ring read cursors         (resampler.rs:420-458 re-framing)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skgpu_batch.h"
#include "phase_runs.h"

namespace skgpu {

// ------------------------------------------------------------------ device-side tables

struct OpHeader {       // lives in device memory so a captured graph sees table-size updates
    uint32_t count;     // live entries (<= capacity the grid was sized for)
    uint32_t count2;    // second table (mix inputs)
    uint32_t pad[2];
};

struct SlotTables {     // SoA per-stream state + configuration, all device pointers
    double *last_index;     // rubato self.last_index
    double *t_ratio;        // 1.0 / resample_ratio
    int32_t *end_idx;       // chunk - 9 - ceil(t)
    uint32_t *chunk;        // chunk_frames
    uint32_t *channels;
    float *hist;            // [slot][16 * max_channels], frames interleaved with the SLOT's channel count
    SkRun *runs;            // [slot][SK_RUNS_MAX]
    uint32_t *n_runs;
    uint32_t *n_out;        // frames the current chunk produces
    float *fifo;            // [slot][fifo_frames * max_channels] (may be null)
    unsigned long long *fifo_w;  // total frames ever written
    unsigned long long *fifo_r;  // total frames ever consumed
    uint32_t max_channels;
    uint32_t fifo_frames;   // power of two
};

// ------------------------------------------------------------------ small helpers

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// streaming 128-bit accesses: inputs are read once, outputs written once -> keep them out of L1
__device__ __forceinline__ float4 ldg_stream_f4(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float2 ldg_stream_f2(const float2 *p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_u4(uint4 *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_u2(uint2 *p, uint2 v) {
    asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

// f32 -> s16: sat_s16(rint_half_even(x * 32768)), NaN -> 0  (SURVEY A5). cvt.rni.sat.s16.f32 is exactly
// this: round-to-nearest-even, saturating, NaN converts to 0.
__device__ __forceinline__ uint32_t f32_to_s16_bits(float x) {
    float y = __fmul_rn(x, 32768.0f);
    int r;
    asm("cvt.rni.sat.s16.f32 %0, %1;" : "=r"(r) : "f"(y));  // 16-bit result sign-extended in a b32 reg
    return (uint32_t)r & 0xFFFFu;
}
__device__ __forceinline__ uint32_t pack_s16x2(float a, float b) { return f32_to_s16_bits(a) | (f32_to_s16_bits(b) << 16); }
__device__ __forceinline__ float s16_to_f32(int s) { return __fmul_rn((float)s, 1.0f / 32768.0f); }

// ------------------------------------------------------------------ K1/K2: convert / gain
// One CTA per (segment, tile). TILE = 2048 samples: a 20 ms 48 kHz stereo frame (1920 samples) is one tile.
// Every thread issues all of its 128-bit loads before the first use (4 x 16 B in flight per thread).

constexpr int CVT_THREADS = 128;
constexpr int CVT_TILE = 2048;

template <int MODE>
__global__ void __launch_bounds__(CVT_THREADS) k_convert(const OpHeader *__restrict__ hdr, const skgpu_seg *__restrict__ segs,
                                                         const float *__restrict__ gains, uint8_t *__restrict__ arena,
                                                         uint32_t tiles_per_seg) {
    const uint32_t seg_i = blockIdx.x / tiles_per_seg;
    const uint32_t tile = blockIdx.x - seg_i * tiles_per_seg;
    if (seg_i >= hdr->count) return;
    const skgpu_seg sg = segs[seg_i];
    const uint32_t n = sg.n_samples;
    const uint32_t s_begin = tile * CVT_TILE;
    if (s_begin >= n) return;
    const uint32_t s_end = min(n, s_begin + CVT_TILE);
    const bool has_gain = sg.gain_idx != SKGPU_NO_GAIN;
    const float g = has_gain ? gains[sg.gain_idx] : 1.0f;

    constexpr int IN_B = (MODE == SKGPU_CVT_S16_TO_F32) ? 2 : 4;
    constexpr int OUT_B = (MODE == SKGPU_CVT_F32_TO_S16) ? 2 : 4;
    const uint8_t *in_p = arena + sg.in_off;
    uint8_t *out_p = arena + sg.out_off;
    const bool aligned = (((uintptr_t)in_p | (uintptr_t)out_p) & 15u) == 0;

    if (MODE == SKGPU_CVT_F32_TO_F32) {
        if (aligned) {
            const float4 *in4 = reinterpret_cast<const float4 *>(in_p);
            float4 *out4 = reinterpret_cast<float4 *>(out_p);
            const uint32_t v_begin = s_begin / 4, v_end = s_end / 4;  // whole vectors
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t vi = v_begin + threadIdx.x + j * CVT_THREADS;
                if (vi < v_end) v[j] = ldg_stream_f4(in4 + vi);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t vi = v_begin + threadIdx.x + j * CVT_THREADS;
                if (vi < v_end) {
                    float4 o = v[j];
                    if (has_gain) {
                        o.x = __fmul_rn(o.x, g); o.y = __fmul_rn(o.y, g);
                        o.z = __fmul_rn(o.z, g); o.w = __fmul_rn(o.w, g);
                    }
                    stg_stream_f4(out4 + vi, o);
                }
            }
            // tail (< 4 samples) of the segment
            const uint32_t t0 = v_end * 4;
            if (t0 + threadIdx.x < s_end && threadIdx.x < 4) {
                const float *in1 = reinterpret_cast<const float *>(in_p);
                float *out1 = reinterpret_cast<float *>(out_p);
                float x = in1[t0 + threadIdx.x];
                out1[t0 + threadIdx.x] = has_gain ? __fmul_rn(x, g) : x;
            }
            return;
        }
    } else if (MODE == SKGPU_CVT_F32_TO_S16) {
        if (aligned) {
            const float4 *in4 = reinterpret_cast<const float4 *>(in_p);
            uint4 *out8 = reinterpret_cast<uint4 *>(out_p);  // 8 x s16
            const uint32_t v_begin = s_begin / 8, v_end = s_end / 8;
            float4 a[2], b[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t vi = v_begin + threadIdx.x + j * CVT_THREADS;
                if (vi < v_end) {
                    a[j] = ldg_stream_f4(in4 + 2 * vi);
                    b[j] = ldg_stream_f4(in4 + 2 * vi + 1);
                }
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t vi = v_begin + threadIdx.x + j * CVT_THREADS;
                if (vi < v_end) {
                    float4 x = a[j], y = b[j];
                    if (has_gain) {
                        x.x = __fmul_rn(x.x, g); x.y = __fmul_rn(x.y, g); x.z = __fmul_rn(x.z, g); x.w = __fmul_rn(x.w, g);
                        y.x = __fmul_rn(y.x, g); y.y = __fmul_rn(y.y, g); y.z = __fmul_rn(y.z, g); y.w = __fmul_rn(y.w, g);
                    }
                    uint4 o;
                    o.x = pack_s16x2(x.x, x.y); o.y = pack_s16x2(x.z, x.w);
                    o.z = pack_s16x2(y.x, y.y); o.w = pack_s16x2(y.z, y.w);
                    stg_stream_u4(out8 + vi, o);
                }
            }
            const uint32_t t0 = v_end * 8;
            if (t0 + threadIdx.x < s_end && threadIdx.x < 8) {
                const float *in1 = reinterpret_cast<const float *>(in_p);
                uint16_t *out1 = reinterpret_cast<uint16_t *>(out_p);
                float x = in1[t0 + threadIdx.x];
                if (has_gain) x = __fmul_rn(x, g);
                out1[t0 + threadIdx.x] = (uint16_t)f32_to_s16_bits(x);
            }
            return;
        }
    } else {  // S16 -> F32
        if (aligned) {
            const uint4 *in8 = reinterpret_cast<const uint4 *>(in_p);
            float4 *out4 = reinterpret_cast<float4 *>(out_p);
            const uint32_t v_begin = s_begin / 8, v_end = s_end / 8;
            uint4 a[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t vi = v_begin + threadIdx.x + j * CVT_THREADS;
                if (vi < v_end) a[j] = ldg_stream_u4(in8 + vi);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t vi = v_begin + threadIdx.x + j * CVT_THREADS;
                if (vi < v_end) {
                    const uint32_t w[4] = {a[j].x, a[j].y, a[j].z, a[j].w};
                    float f[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        f[2 * q] = s16_to_f32((int)(int16_t)(w[q] & 0xFFFFu));
                        f[2 * q + 1] = s16_to_f32((int)(int16_t)(w[q] >> 16));
                    }
                    if (has_gain) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) f[q] = __fmul_rn(f[q], g);
                    }
                    stg_stream_f4(out4 + 2 * vi, make_float4(f[0], f[1], f[2], f[3]));
                    stg_stream_f4(out4 + 2 * vi + 1, make_float4(f[4], f[5], f[6], f[7]));
                }
            }
            const uint32_t t0 = v_end * 8;
            if (t0 + threadIdx.x < s_end && threadIdx.x < 8) {
                const int16_t *in1 = reinterpret_cast<const int16_t *>(in_p);
                float *out1 = reinterpret_cast<float *>(out_p);
                float x = s16_to_f32((int)in1[t0 + threadIdx.x]);
                out1[t0 + threadIdx.x] = has_gain ? __fmul_rn(x, g) : x;
            }
            return;
        }
    }
    // unaligned segment: scalar path (ragged offsets; correctness only)
    for (uint32_t s = s_begin + threadIdx.x; s < s_end; s += CVT_THREADS) {
        float x;
        if (IN_B == 2) x = s16_to_f32((int)reinterpret_cast<const int16_t *>(in_p)[s]);
        else x = reinterpret_cast<const float *>(in_p)[s];
        if (has_gain) x = __fmul_rn(x, g);
        if (OUT_B == 2) reinterpret_cast<uint16_t *>(out_p)[s] = (uint16_t)f32_to_s16_bits(x);
        else reinterpret_cast<float *>(out_p)[s] = x;
    }
}

// ------------------------------------------------------------------ K4a: phase recurrence -> run table
// One THREAD per stream-chunk: 32 streams share a warp, so the inherently sequential f64 chain
// (~chunk*ratio dependent DADDs) is SIMD across streams. Data independent: reads only slot state.

constexpr int PHASE_THREADS = 128;

__global__ void __launch_bounds__(PHASE_THREADS) k_phase(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items,
                                                         SlotTables st, uint8_t *__restrict__ arena, uint64_t results_off) {
    const uint32_t i = blockIdx.x * PHASE_THREADS + threadIdx.x;
    if (i >= hdr->count) return;
    const skgpu_rs_item it = items[i];
    const uint32_t slot = it.slot;
    const double L = st.last_index[slot];
    const double t = st.t_ratio[slot];
    const int32_t end_idx = st.end_idx[slot];
    uint32_t nr;
    double idx_end;
    int ovf;
    const uint32_t n = sk_phase_runs(L, t, end_idx, st.runs + (size_t)slot * SK_RUNS_MAX, SK_RUNS_MAX, &nr, &idx_end, &ovf);
    st.n_runs[slot] = nr;
    st.n_out[slot] = n;
    st.last_index[slot] = __dsub_rn(idx_end, (double)st.chunk[slot]);  // self.last_index = idx - chunk_size as f64
    skgpu_rs_result res;
    res.out_frames = n;
    res.status = 0;
    if (!(it.flags & SKGPU_RS_TO_FIFO) && n > it.out_cap_frames) {
        res.out_frames = it.out_cap_frames;
        res.status = 1;
    }
    if (ovf) res.status = 2;
    reinterpret_cast<skgpu_rs_result *>(arena + results_off)[i] = res;
}

// ------------------------------------------------------------------ K4b: interpolation
// One CTA per stream-chunk. The chunk (and the 16-frame history from HBM state) is staged into shared
// memory with ONE TMA bulk copy each (cp.async.bulk + mbarrier) -- the "tap window" staging: reads of
// y[p], y[p+1] are data dependent (p = floor(idx)) so they are served from smem, while HBM only sees a
// single fully coalesced pass over the input. While the copy is in flight every thread evaluates the
// f64 phase of its first output from the run table.

constexpr int RS_THREADS = 128;

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// rubato interp_lin: (1 - x) * y0 + x * y1, every operation rounded to f32
__device__ __forceinline__ float interp_lin(float frac, float y0, float y1) {
    return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, frac), y0), __fmul_rn(frac, y1));
}

template <int C>  // C = 1, 2 specialised; 0 = runtime channel count
__global__ void __launch_bounds__(RS_THREADS) k_resample(const OpHeader *__restrict__ hdr, const skgpu_rs_item *__restrict__ items,
                                                         SlotTables st, uint8_t *__restrict__ arena, uint32_t smem_frames) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ SkRun s_runs[SK_RUNS_MAX];

    const uint32_t i = blockIdx.x;
    if (i >= hdr->count) return;
    const skgpu_rs_item it = items[i];
    const uint32_t slot = it.slot;
    const uint32_t ch = (C > 0) ? (uint32_t)C : st.channels[slot];
    const uint32_t N = st.chunk[slot];
    const uint32_t nr = st.n_runs[slot];
    uint32_t n_out = st.n_out[slot];
    const bool to_fifo = (it.flags & SKGPU_RS_TO_FIFO) != 0;
    if (!to_fifo) n_out = min(n_out, it.out_cap_frames);

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
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                mbar_expect_tx(&bar, hist_bytes + in_bytes);
                tma_bulk_g2s(buf, hist_g, hist_bytes, &bar);
                tma_bulk_g2s(buf + 16u * ch, in_g, in_bytes, &bar);
            }
        } else {
            for (uint32_t s = threadIdx.x; s < 16u * ch; s += RS_THREADS) buf[s] = hist_g[s];
            for (uint32_t s = threadIdx.x; s < N * ch; s += RS_THREADS) buf[16u * ch + s] = in_g[s];
        }
    }
    // run table -> smem (overlaps the bulk copy)
    {
        const uint32_t words = nr * (uint32_t)(sizeof(SkRun) / 4);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(st.runs + (size_t)slot * SK_RUNS_MAX);
        uint32_t *dst = reinterpret_cast<uint32_t *>(s_runs);
        for (uint32_t w = threadIdx.x; w < words; w += RS_THREADS) dst[w] = src[w];
    }
    __syncthreads();
    if (staged && tma_ok) mbar_wait(&bar, 0);

    // output destination
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

    uint32_t r = 0;
    for (uint32_t k = threadIdx.x; k < n_out; k += RS_THREADS) {
        while (r + 1 < nr && s_runs[r + 1].k_a <= k) ++r;
        const double x = sk_phase_eval(s_runs[r], k);
        const double fl = floor(x);
        const float frac = (float)(x - fl);            // T::coerce(idx - floor(idx))
        const uint32_t p = (uint32_t)((int)fl + 16);    // start_idx + 2 * POLYNOMIAL_LEN
        const uint32_t of = to_fifo ? (uint32_t)((fifo_w + k) & fifo_mask) : k;
        if (C == 2) {
            float2 y0, y1;
            if (staged) {
                y0 = *reinterpret_cast<const float2 *>(buf + 2u * p);
                y1 = *reinterpret_cast<const float2 *>(buf + 2u * p + 2u);
            } else {
                const float *a = (p < 16u) ? hist_g + 2u * p : in_g + 2u * (p - 16u);
                const float *b = (p + 1u < 16u) ? hist_g + 2u * (p + 1u) : in_g + 2u * (p + 1u - 16u);
                y0 = make_float2(a[0], a[1]);
                y1 = make_float2(b[0], b[1]);
            }
            float2 o = make_float2(interp_lin(frac, y0.x, y1.x), interp_lin(frac, y0.y, y1.y));
            *reinterpret_cast<float2 *>(out_g + 2u * of) = o;
        } else {
            for (uint32_t c = 0; c < ch; ++c) {
                float y0, y1;
                if (staged) {
                    y0 = buf[p * ch + c];
                    y1 = buf[(p + 1u) * ch + c];
                } else {
                    y0 = (p < 16u) ? hist_g[p * ch + c] : in_g[(p - 16u) * ch + c];
                    y1 = (p + 1u < 16u) ? hist_g[(p + 1u) * ch + c] : in_g[(p + 1u - 16u) * ch + c];
                }
                out_g[of * ch + c] = interp_lin(frac, y0, y1);
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
    if (to_fifo && threadIdx.x == 0) st.fifo_w[slot] = fifo_w + st.n_out[slot];
}

// ------------------------------------------------------------------ K3: ordered mixer + epilogue
// One CTA per (group, tile of 512 output samples); one thread owns 4 consecutive output samples and adds the
// inputs SEQUENTIALLY in the reference's order (f32 addition is not associative: a warp-shuffle tree over
// inputs would not be bit-exact). Coalescing comes from adjacent threads owning adjacent samples;
// memory-level parallelism from the (unrolled) independent loads of successive inputs.

constexpr int MIX_THREADS = 128;
constexpr int MIX_TILE = MIX_THREADS * 4;
constexpr int MIX_MAX_INPUTS = 1024;

struct MixIn {           // resolved per-tick view of one present input, in summation order
    const float *ptr;    // frame base (arena) or ring base (fifo)
    uint32_t n_frames;   // frames available from this input (<= out_frames is NOT implied)
    uint16_t channels;
    uint16_t fifo;       // 1 = ring addressing
    float gain;
    uint32_t has_gain;
    uint32_t ring_start; // first ring frame of the packet
    uint32_t ring_mask;
};

__device__ __forceinline__ float mix_fetch(const MixIn &in, uint32_t frame, uint32_t c) {
    const uint32_t f = in.fifo ? ((in.ring_start + frame) & in.ring_mask) : frame;
    return in.ptr[(size_t)f * in.channels + c];
}

// epilogue: master audio::gain (gain.rs:187-189), then f32 store or clip + s16 pack
__device__ __forceinline__ void mix_epilogue(const skgpu_mix_group &grp, const float *__restrict__ gains, uint8_t *__restrict__ arena,
                                             uint32_t s0, uint32_t nvalid, float *acc) {
    if (grp.gain_idx != SKGPU_NO_GAIN) {
        const float g = gains[grp.gain_idx];
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = __fmul_rn(acc[e], g);
    }
    if (grp.flags & SKGPU_MIX_OUT_S16) {
        uint16_t *o = reinterpret_cast<uint16_t *>(arena + grp.out_off) + s0;
        if (nvalid == 4 && ((((uintptr_t)o) & 7u) == 0)) {
            stg_stream_u2(reinterpret_cast<uint2 *>(o), make_uint2(pack_s16x2(acc[0], acc[1]), pack_s16x2(acc[2], acc[3])));
        } else {
            for (uint32_t e = 0; e < nvalid; ++e) o[e] = (uint16_t)f32_to_s16_bits(acc[e]);
        }
    } else {
        float *o = reinterpret_cast<float *>(arena + grp.out_off) + s0;
        if (nvalid == 4 && ((((uintptr_t)o) & 15u) == 0)) {
            stg_stream_f4(reinterpret_cast<float4 *>(o), make_float4(acc[0], acc[1], acc[2], acc[3]));
        } else {
            for (uint32_t e = 0; e < nvalid; ++e) o[e] = acc[e];
        }
    }
}

__global__ void __launch_bounds__(MIX_THREADS) k_mix(const OpHeader *__restrict__ hdr, const skgpu_mix_group *__restrict__ groups,
                                                     const skgpu_mix_input *__restrict__ inputs, const uint8_t *__restrict__ present,
                                                     const float *__restrict__ gains, SlotTables st, uint8_t *__restrict__ arena,
                                                     uint32_t tiles_per_group) {
    __shared__ MixIn s_in[MIX_MAX_INPUTS > 64 ? 64 : MIX_MAX_INPUTS];  // first 64 inputs cached in smem
    __shared__ uint16_t s_order[MIX_MAX_INPUTS];
    __shared__ uint8_t s_flag[MIX_MAX_INPUTS];
    __shared__ uint32_t s_n, s_has_base;

    const uint32_t g_i = blockIdx.x / tiles_per_group;
    const uint32_t tile = blockIdx.x - g_i * tiles_per_group;
    if (g_i >= hdr->count) return;
    const skgpu_mix_group grp = groups[g_i];
    const uint32_t oc = grp.out_channels;
    const uint32_t out_size = grp.out_frames * oc;
    const uint32_t s0 = tile * MIX_TILE + threadIdx.x * 4u;
    if (tile * MIX_TILE >= out_size) return;
    const uint32_t n_in = min(grp.n_inputs, (uint32_t)MIX_MAX_INPUTS);

    // ---- prologue: which inputs are present, base-frame selection, swap_remove order (mixer.rs:960-980)
    // flags are computed by all threads in parallel (one global round trip), the tiny ordered compaction
    // runs on thread 0 out of shared memory.
    for (uint32_t j = threadIdx.x; j < n_in; j += MIX_THREADS) {
        const uint32_t gi = grp.first_input + j;
        const skgpu_mix_input in = inputs[gi];
        bool pres = present ? (present[gi] != 0) : true;
        if (pres && (in.flags & SKGPU_MIX_IN_FIFO)) {
            const unsigned long long avail = st.fifo_w[in.slot] - st.fifo_r[in.slot];
            pres = avail >= (unsigned long long)in.n_frames;  // a whole re-framed packet is ready
        }
        // frame.channels == output_channels && frame.samples.len() == output_size
        const bool elig = in.channels == oc && in.n_frames * in.channels == out_size;
        s_flag[j] = (uint8_t)((pres ? 1u : 0u) | (elig ? 2u : 0u) | ((in.flags & SKGPU_MIX_IN_UNIQUE) ? 4u : 0u));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t m = 0;
        int base = -1, base_unique = -1;
        for (uint32_t j = 0; j < n_in; ++j) {
            const uint32_t fl = s_flag[j];
            if (!(fl & 1u)) continue;
            if (fl & 2u) {
                const int u = (fl & 4u) ? 1 : 0;
                if (u >= base_unique) { base = (int)m; base_unique = u; }  // max_by_key((unique, idx)): last max
            }
            s_order[m++] = (uint16_t)j;
        }
        if (base >= 0 && m > 0) {
            // Vec::swap_remove(base): the last element takes the slot; the base goes first in our order and
            // order[1..] is vec[0..m-1] after the swap_remove.
            const uint16_t b = s_order[base];
            s_order[base] = s_order[m - 1];
            for (uint32_t q = m - 1; q > 0; --q) s_order[q] = s_order[q - 1];
            s_order[0] = b;
        }
        s_n = m;
        s_has_base = (base >= 0) ? 1u : 0u;
    }
    __syncthreads();
    const uint32_t m = s_n;
    // resolve inputs (in summation order) into smem
    bool simple = true;
    for (uint32_t q = threadIdx.x; q < m && q < 64u; q += MIX_THREADS) {
        const skgpu_mix_input in = inputs[grp.first_input + s_order[q]];
        MixIn r;
        r.n_frames = in.n_frames;
        r.channels = in.channels;
        r.has_gain = in.gain_idx != SKGPU_NO_GAIN;
        r.gain = r.has_gain ? gains[in.gain_idx] : 1.0f;
        if (in.flags & SKGPU_MIX_IN_FIFO) {
            r.fifo = 1;
            r.ptr = st.fifo + (size_t)in.slot * st.fifo_frames * st.max_channels;
            r.ring_mask = st.fifo_frames - 1u;
            r.ring_start = (uint32_t)(st.fifo_r[in.slot] & r.ring_mask);
        } else {
            r.fifo = 0;
            r.ptr = reinterpret_cast<const float *>(arena + in.in_off);
            r.ring_mask = 0; r.ring_start = 0;
        }
        simple = simple && !r.fifo && r.channels == oc && r.n_frames >= grp.out_frames && ((((uintptr_t)r.ptr) & 15u) == 0);
        s_in[q] = r;
    }
    const bool all_simple = __syncthreads_and(simple ? 1 : 0) && m <= 64u && (out_size % 4u == 0);
    if (s0 >= out_size) return;
    const bool has_base = s_has_base != 0;

    if (all_simple) {
        // fast path: every present input has the output's shape -> pure 128-bit streaming, 8 loads in flight,
        // still one sequential chain of f32 additions per output sample in the reference order.
        float4 a = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        uint32_t q = 0;
        if (has_base) {
            const MixIn in = s_in[0];
            float4 t = ldg_stream_f4(reinterpret_cast<const float4 *>(in.ptr + s0));
            if (in.has_gain) { t.x = __fmul_rn(t.x, in.gain); t.y = __fmul_rn(t.y, in.gain); t.z = __fmul_rn(t.z, in.gain); t.w = __fmul_rn(t.w, in.gain); }
            a = t;
            q = 1;
        }
        for (; q + 8 <= m; q += 8) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) t[u] = ldg_stream_f4(reinterpret_cast<const float4 *>(s_in[q + u].ptr + s0));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float4 x = t[u];
                if (s_in[q + u].has_gain) {
                    const float g = s_in[q + u].gain;
                    x.x = __fmul_rn(x.x, g); x.y = __fmul_rn(x.y, g); x.z = __fmul_rn(x.z, g); x.w = __fmul_rn(x.w, g);
                }
                a.x = __fadd_rn(a.x, x.x); a.y = __fadd_rn(a.y, x.y); a.z = __fadd_rn(a.z, x.z); a.w = __fadd_rn(a.w, x.w);
            }
        }
        for (; q < m; ++q) {
            float4 x = ldg_stream_f4(reinterpret_cast<const float4 *>(s_in[q].ptr + s0));
            if (s_in[q].has_gain) {
                const float g = s_in[q].gain;
                x.x = __fmul_rn(x.x, g); x.y = __fmul_rn(x.y, g); x.z = __fmul_rn(x.z, g); x.w = __fmul_rn(x.w, g);
            }
            a.x = __fadd_rn(a.x, x.x); a.y = __fadd_rn(a.y, x.y); a.z = __fadd_rn(a.z, x.z); a.w = __fadd_rn(a.w, x.w);
        }
        float accf[4] = {a.x, a.y, a.z, a.w};
        mix_epilogue(grp, gains, arena, s0, 4u, accf);
        return;
    }

    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};  // vec![0.0f32; output_size] when there is no base frame
    const uint32_t nvalid = min(4u, out_size - s0);

    for (uint32_t q = 0; q < m; ++q) {
        MixIn in;
        if (q < 64u) in = s_in[q];
        else {  // rare: > 64 present inputs, resolve on the fly
            const skgpu_mix_input gi = inputs[grp.first_input + s_order[q]];
            in.n_frames = gi.n_frames; in.channels = gi.channels;
            in.has_gain = gi.gain_idx != SKGPU_NO_GAIN; in.gain = in.has_gain ? gains[gi.gain_idx] : 1.0f;
            if (gi.flags & SKGPU_MIX_IN_FIFO) {
                in.fifo = 1; in.ptr = st.fifo + (size_t)gi.slot * st.fifo_frames * st.max_channels;
                in.ring_mask = st.fifo_frames - 1u; in.ring_start = (uint32_t)(st.fifo_r[gi.slot] & in.ring_mask);
            } else { in.fifo = 0; in.ptr = reinterpret_cast<const float *>(arena + gi.in_off); in.ring_mask = 0; in.ring_start = 0; }
        }
        const uint32_t sc = in.channels;
        // mix_samples_per_channel = min(source frames, output frames)  (mixer.rs:1034-1037)
        const uint32_t mix_frames = min(in.n_frames, grp.out_frames);
        const bool is_base = has_base && q == 0;
        float v[4];
        bool ok[4];
        if (sc == oc) {
            const uint32_t mix_len = mix_frames * oc;
            const bool vec = !in.fifo && (s0 + 4u <= mix_len) && ((((uintptr_t)(in.ptr + s0)) & 15u) == 0);
            if (vec) {
                const float4 t = ldg_stream_f4(reinterpret_cast<const float4 *>(in.ptr + s0));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
                ok[0] = ok[1] = ok[2] = ok[3] = true;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t s = s0 + e;
                    ok[e] = s < mix_len;
                    v[e] = ok[e] ? mix_fetch(in, s / oc, s % oc) : 0.0f;
                }
            }
        } else if (sc == 1 && oc == 2) {  // mono -> stereo: duplicate (mixer.rs:1047-1054)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t f = (s0 + e) >> 1;
                ok[e] = f < mix_frames;
                v[e] = ok[e] ? mix_fetch(in, f, 0) : 0.0f;
            }
        } else if (sc == 2 && oc == 1) {  // stereo -> mono: (L + R) * 0.5 (mixer.rs:1055-1061)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t f = s0 + e;
                ok[e] = f < mix_frames;
                if (ok[e]) {
                    float l = mix_fetch(in, f, 0), rr = mix_fetch(in, f, 1);
                    if (in.has_gain) { l = __fmul_rn(l, in.gain); rr = __fmul_rn(rr, in.gain); }
                    v[e] = __fmul_rn(__fadd_rn(l, rr), 0.5f);
                } else v[e] = 0.0f;
            }
        } else {  // generic cyclic mapping (mixer.rs:1062-1076)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t s = s0 + e;
                const uint32_t f = s / oc, c = s % oc;
                ok[e] = f < mix_frames;
                v[e] = ok[e] ? mix_fetch(in, f, c % sc) : 0.0f;
            }
        }
        const bool gain_pending = in.has_gain && !(sc == 2 && oc == 1);  // stereo->mono applied it per channel above
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!ok[e]) continue;
            float x = gain_pending ? __fmul_rn(v[e], in.gain) : v[e];  // upstream audio::gain, rounded separately
            acc[e] = is_base ? x : __fadd_rn(acc[e], x);               // base frame IS the accumulator
        }
    }

    mix_epilogue(grp, gains, arena, s0, nvalid, acc);
}

// advances ring read cursors of FIFO-sourced mix inputs that delivered a packet this tick
__global__ void k_fifo_commit(const OpHeader *__restrict__ hdr, const skgpu_mix_input *__restrict__ inputs,
                              const uint8_t *__restrict__ present, SlotTables st) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hdr->count2) return;
    const skgpu_mix_input in = inputs[i];
    if (!(in.flags & SKGPU_MIX_IN_FIFO)) return;
    if (present && !present[i]) return;
    const unsigned long long w = st.fifo_w[in.slot], r = st.fifo_r[in.slot];
    if (w - r >= (unsigned long long)in.n_frames) st.fifo_r[in.slot] = r + in.n_frames;
}

// (re)initialises stream slots: fresh FastFixedIn = zero history, last_index = -4.0 (rubato new())
__global__ void k_reset_slots(const uint32_t *__restrict__ slots, uint32_t n, SlotTables st) {
    const uint32_t i = blockIdx.x;
    if (i >= n) return;
    const uint32_t slot = slots[i];
    for (uint32_t s = threadIdx.x; s < 16u * st.max_channels; s += blockDim.x) st.hist[(size_t)slot * 16u * st.max_channels + s] = 0.0f;
    if (threadIdx.x == 0) {
        st.last_index[slot] = -4.0;
        st.n_runs[slot] = 0;
        st.n_out[slot] = 0;
        if (st.fifo_w) { st.fifo_w[slot] = 0ull; st.fifo_r[slot] = 0ull; }
    }
}

__global__ void k_l2_flush(uint4 *buf, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_uint4((uint32_t)i, 0u, 0u, 0u);
}

}  // namespace skgpu
