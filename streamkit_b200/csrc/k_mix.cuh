// k_mix.cuh -- ordered N-input mixer with channel conversion and gain/clip/s16 epilogue (mixer.rs:944-1078)
#pragma once
#include "common.cuh"

namespace skgpu {
// ------------------------------------------------------------------ K3: ordered mixer + epilogue
// One CTA per (group, tile of 512 output samples); one thread owns 4 consecutive output samples and adds the
// inputs SEQUENTIALLY in the reference's order (f32 addition is not associative: a warp-shuffle tree over
// inputs would not be bit-exact). Coalescing comes from adjacent threads owning adjacent samples;
// memory-level parallelism from the (unrolled) independent loads of successive inputs.

constexpr int MIX_THREADS = 96;    // tile = 384 samples: a 20 ms 48 kHz stereo frame (1920 samples) is exactly five full tiles
constexpr int MIX_TILE = MIX_THREADS * 4;
constexpr int MIX_MLP = 8;   // 128-bit loads a thread keeps in flight in the same-shape fast path (12: 117 us, 16: 101 us, 8: 89 us at config #3)
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
                                                     uint32_t tiles_per_group, uint32_t tiles_per_cta) {
    __shared__ MixIn s_in[MIX_MAX_INPUTS > 64 ? 64 : MIX_MAX_INPUTS];  // first 64 inputs cached in smem
    __shared__ uint16_t s_order[MIX_MAX_INPUTS];
    __shared__ uint8_t s_flag[MIX_MAX_INPUTS];
    __shared__ uint32_t s_n, s_has_base;

    // a CTA owns tiles_per_cta consecutive tiles of one group: small groups (few inputs) are handled whole by one CTA so
    // that the prologue is paid once per group, large ones one tile per CTA for parallelism
    const uint32_t ctas_per_group = (tiles_per_group + tiles_per_cta - 1u) / tiles_per_cta;
    const uint32_t g_i = blockIdx.x / ctas_per_group;
    const uint32_t tile0 = (blockIdx.x - g_i * ctas_per_group) * tiles_per_cta;
    if (g_i >= hdr->count) return;
    const skgpu_mix_group grp = groups[g_i];
    const uint32_t oc = grp.out_channels;
    const uint32_t out_size = grp.out_frames * oc;
    if (tile0 * MIX_TILE >= out_size) return;
    const uint32_t n_in = min(grp.n_inputs, (uint32_t)MIX_MAX_INPUTS);

    // ---- prologue: which inputs are present, base-frame selection, swap_remove order (mixer.rs:960-980)
    bool simple = true;
    uint32_t m;
    auto resolve = [&](const skgpu_mix_input &in) -> MixIn {
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
        return r;
    };
    auto is_simple = [&](const MixIn &r) { return !r.fifo && r.channels == oc && r.n_frames >= grp.out_frames && ((((uintptr_t)r.ptr) & 15u) == 0); };
    if (n_in <= 32u) {
        // small groups (the common 2..8-input mixers): warp 0 does everything with three ballots -- lane j owns input j,
        // its position in the summation order follows from popcounts, and it resolves its own input straight into
        // shared memory: one global round trip, one barrier, no serial loop.
        if (threadIdx.x < 32u) {
            const uint32_t lane = threadIdx.x;
            skgpu_mix_input in{};
            bool pres = false, elig = false;
            if (lane < n_in) {
                const uint32_t gi = grp.first_input + lane;
                in = inputs[gi];
                pres = present ? (present[gi] != 0) : true;
                if (pres && (in.flags & SKGPU_MIX_IN_FIFO)) {
                    const unsigned long long avail = st.fifo_w[in.slot] - st.fifo_r[in.slot];
                    pres = avail >= (unsigned long long)in.n_frames;  // a whole re-framed packet is ready
                }
                elig = pres && in.channels == oc && in.n_frames * in.channels == out_size;   // already has the output's shape
            }
            const uint32_t pres_mask = __ballot_sync(0xffffffffu, pres);
            const uint32_t elig_mask = __ballot_sync(0xffffffffu, elig);
            const uint32_t uniq_mask = __ballot_sync(0xffffffffu, elig && (in.flags & SKGPU_MIX_IN_UNIQUE));
            const uint32_t mm = __popc(pres_mask);
            // max_by_key((unique, idx)): the last unique full-shape frame, else the last full-shape frame
            const int base_lane = uniq_mask ? 31 - __clz(uniq_mask) : (elig_mask ? 31 - __clz(elig_mask) : -1);
            if (pres) {
                const uint32_t rank = __popc(pres_mask & ((1u << lane) - 1u));
                uint32_t pos = rank;
                if (base_lane >= 0) {
                    const uint32_t base_rank = __popc(pres_mask & ((1u << base_lane) - 1u));
                    pos = ((int)lane == base_lane) ? 0u : 1u + ((rank == mm - 1u) ? base_rank : rank);   // Vec::swap_remove
                }
                const MixIn r = resolve(in);
                s_in[pos] = r;
                s_order[pos] = (uint16_t)lane;
                simple = is_simple(r);
            }
            if (lane == 0) { s_n = mm; s_has_base = base_lane >= 0 ? 1u : 0u; }
        }
    } else {
        // flags are computed by all threads in parallel (one global round trip), the ordered compaction
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
            uint32_t mm = 0;
            int base = -1, base_unique = -1;
            for (uint32_t j = 0; j < n_in; ++j) {
                const uint32_t fl = s_flag[j];
                if (!(fl & 1u)) continue;
                if (fl & 2u) {
                    const int u = (fl & 4u) ? 1 : 0;
                    if (u >= base_unique) { base = (int)mm; base_unique = u; }  // max_by_key((unique, idx)): last max
                }
                s_order[mm++] = (uint16_t)j;
            }
            if (base >= 0 && mm > 0) {
                // Vec::swap_remove(base): the last element takes the slot; the base goes first in our order and
                // order[1..] is vec[0..m-1] after the swap_remove.
                const uint16_t b = s_order[base];
                s_order[base] = s_order[mm - 1];
                for (uint32_t q = mm - 1; q > 0; --q) s_order[q] = s_order[q - 1];
                s_order[0] = b;
            }
            s_n = mm;
            s_has_base = (base >= 0) ? 1u : 0u;
        }
        __syncthreads();
        // resolve inputs (in summation order) into smem
        for (uint32_t q = threadIdx.x; q < s_n && q < 64u; q += MIX_THREADS) {
            const MixIn r = resolve(inputs[grp.first_input + s_order[q]]);
            simple = simple && is_simple(r);
            s_in[q] = r;
        }
    }
    const bool all_simple_v = __syncthreads_and(simple ? 1 : 0) != 0;
    m = s_n;
    const bool all_simple = all_simple_v && m <= 64u && (out_size % 4u == 0);
    const bool has_base = s_has_base != 0;
    for (uint32_t tile = tile0; tile < tile0 + tiles_per_cta; ++tile) {
    const uint32_t s0 = tile * MIX_TILE + threadIdx.x * 4u;
    if (s0 >= out_size) break;

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
        // batches of MIX_MLP loads in flight; the last batch is partial (guarded) rather than a tail of one-at-a-time loads, each of
        // which would expose a full memory latency. The adds stay in input order.
        for (; q < m; q += MIX_MLP) {
            float4 t[MIX_MLP];
#pragma unroll
            for (int u = 0; u < MIX_MLP; ++u)
                if (q + u < m) t[u] = ldg_stream_f4(reinterpret_cast<const float4 *>(s_in[q + u].ptr + s0));
#pragma unroll
            for (int u = 0; u < MIX_MLP; ++u) {
                if (q + u < m) {
                    float4 x = t[u];
                    if (s_in[q + u].has_gain) {
                        const float g = s_in[q + u].gain;
                        x.x = __fmul_rn(x.x, g); x.y = __fmul_rn(x.y, g); x.z = __fmul_rn(x.z, g); x.w = __fmul_rn(x.w, g);
                    }
                    a.x = __fadd_rn(a.x, x.x); a.y = __fadd_rn(a.y, x.y); a.z = __fadd_rn(a.z, x.z); a.w = __fadd_rn(a.w, x.w);
                }
            }
        }
        float accf[4] = {a.x, a.y, a.z, a.w};
        mix_epilogue(grp, gains, arena, s0, 4u, accf);
        continue;
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
            // four consecutive samples of the packet: contiguous in a plain frame; in a ring when they do not wrap
            // (ring sizes and packet sizes are multiples of 4 samples, so an aligned quad never straddles the end)
            const float *src = in.ptr + s0;
            if (in.fifo) {
                const uint32_t ring_samples = (in.ring_mask + 1u) * oc;
                const uint32_t pos = (in.ring_start * oc + s0) & (ring_samples - 1u);   // ring_mask + 1 is a power of two; oc is 1 or 2 here
                src = in.ptr + pos;
            }
            const bool pow2_oc = (oc & (oc - 1u)) == 0u;
            const bool vec = (!in.fifo || pow2_oc) && (s0 + 4u <= mix_len) && ((((uintptr_t)src) & 15u) == 0);
            if (vec) {
                const float4 t = ldg_stream_f4(reinterpret_cast<const float4 *>(src));
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
    }   // tiles of this CTA
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

}  // namespace skgpu
