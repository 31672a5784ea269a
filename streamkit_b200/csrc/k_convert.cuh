// k_convert.cuh -- gain / f32<->s16 conversion kernels (gain.rs:184-190)
#pragma once
#include "common.cuh"

namespace skgpu {
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
                    // 8 floats = one 32-byte sector per thread: a single 256-bit store (STG.256, sm_100) instead of two
                    // half-sector stores
                    if ((((uintptr_t)(out4 + 2 * vi)) & 31u) == 0) {
                        asm volatile("st.global.L1::no_allocate.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(out4 + 2 * vi), "f"(f[0]), "f"(f[1]),
                                     "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
                    } else {
                        stg_stream_f4(out4 + 2 * vi, make_float4(f[0], f[1], f[2], f[3]));
                        stg_stream_f4(out4 + 2 * vi + 1, make_float4(f[4], f[5], f[6], f[7]));
                    }
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

}  // namespace skgpu
