// common.cuh -- device-side tables and small helpers shared by all streamkit_b200 kernels.
//
// All kernels are HBM-bound streaming kernels (<= 3 flop/byte): no tensor cores. What matters is 128-bit
// coalesced access, enough bytes in flight per SM, shared-memory/TMA staging where the access pattern is
// data dependent (resampler), and exact IEEE arithmetic: the TU is compiled with -fmad=false and the
// parity-critical expressions additionally use __fmul_rn/__fadd_rn so nothing is ever contracted
// (Rust, the reference's language, never contracts a*b+c).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/skgpu_batch.h"
#include "phase_runs.h"

namespace skgpu {

// ------------------------------------------------------------------ device-side tables

struct OpHeader {       // lives in device memory so a captured graph sees table-size updates
    uint32_t count;     // live entries (<= capacity the grid was sized for)
    uint32_t count2;    // second table (mix / chain inputs)
    uint32_t first;     // chain op, sliced ticks: this launch covers groups [first, first + count) ...
    uint32_t first2;    // ... and inputs [first2, first2 + count2) of the op's tables (0 for a whole-tick launch)
};

struct __align__(16) SlotRec {  // everything a kernel needs to know about one resampler stream: ONE 64-byte load
    double t_ratio;        // 1.0 / resample_ratio                                  (host config)
    double last_index;     // rubato self.last_index for the NEXT chunk             (k_phase)
    uint32_t chunk;        // chunk_frames                                          (host config)
    uint32_t channels;     //                                                       (host config)
    uint32_t chunk_count;  // chunks processed so far                               (k_phase)
    uint32_t carry;        // chain op: frames produced but not yet emitted in a packet (k_chain)
    uint32_t n_out[2];     // frames produced by the two most recent chunks, by (chunk number & 1)
    uint16_t n_prefix[2];  // phase-table sizes of those chunks
    uint16_t n_runs[2];
    int32_t end_idx;       // chunk - 9 - ceil(t)                                   (host config)
    uint32_t overflow;     // bit (chunk number & 1): phase table overflowed
    uint32_t flags;        // SLOT_*                                                (host config)
};
constexpr uint32_t SLOT_BYPASS = 1u, SLOT_S16 = 2u, SLOT_SINC = 4u;   // bits 8-15: index of the stream's sinc tap table
static_assert(sizeof(SlotRec) == 64, "SlotRec must be one 64-byte record");

struct SlotCfgUpload {     // host -> device (re)configuration of one slot (k_config_slots)
    double t_ratio;
    uint32_t slot, chunk, channels;
    int32_t end_idx;
    uint32_t flags, aux;   // aux: sinc streams: the sub-phase period in outputs (kept in SlotRec.carry, which only chain streams use)
    double last_index0;    // rubato's initial last_index: -4.0 (FastFixedIn, POLYNOMIAL_LEN / 2), -(sinc_len / 2) in sinc mode
};

constexpr uint32_t SK_SIDE_STRIDE = 2048u;   // >= sizeof(SkPhaseTable); fused chain: frame program (<= 1920 B) + 128 B history
constexpr uint32_t SK_SIDE_HIST = 128u;      // 16 frames x 2 channels x 4 bytes

struct SlotTables {     // per-stream state + configuration, all device pointers
    SlotRec *rec;           // [slot]
    float *hist;            // [slot][16 * max_channels]: the 16 frames before the oldest chunk that still has
                            // unconsumed output (plain resample op: before the next chunk; chain op: before the previous one)
    uint8_t *side;          // [slot][2][SK_SIDE_STRIDE]: per (slot, chunk number & 1) record. Plain resample op: a SkPhaseTable.
                            // Fused chain: [frame program of the next packet (chain_prog.h) ... | 128 B history before that chunk]
    float *fifo;            // [slot][fifo_frames * max_channels] (may be null): unfused re-framing ring
    unsigned long long *fifo_w;  // total frames ever written
    unsigned long long *fifo_r;  // total frames ever consumed
    uint32_t max_channels;
    uint32_t fifo_frames;   // power of two
    // windowed-sinc mode (k_sinc.cuh); null / 0 until skgpu_ctx_set_sinc
    float *sinc_hist;                  // [slot][sinc_H * max_channels]: the last sinc_H input frames
    const float *const *sinc_tabs;     // tap tables [(sinc_O + 1)][sinc_L + 4], one per distinct cutoff
    uint32_t sinc_L, sinc_O, sinc_H, sinc_pad;
};

// ------------------------------------------------------------------ small helpers

__device__ __forceinline__ uint8_t *slot_side(const SlotTables &st, uint32_t slot, uint32_t par) {
    return st.side + ((size_t)slot * 2u + par) * SK_SIDE_STRIDE;
}
__device__ __forceinline__ SkPhaseTable *slot_tab(const SlotTables &st, uint32_t slot, uint32_t par) {
    return reinterpret_cast<SkPhaseTable *>(slot_side(st, slot, par));
}

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
__device__ __forceinline__ void stg_stream_f4(float4 *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_f2(float2 *p, float2 v) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream_u4(uint4 *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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
// s / 32768 without an int->float conversion (I2F runs on the quarter-rate XU pipe): 0x4B400000 is 1.5 * 2^23, where one ulp is
// 1.0, so adding s to the bit pattern yields exactly 12582912 + s; the subtraction and the power-of-two scaling are exact too --
// bit-identical to (float)s * (1.0f / 32768.0f) for every 16-bit s (test_s16_to_f32_all_values_and_roundtrip)
__device__ __forceinline__ float s16_to_f32(int s) {
    return __fmul_rn(__fsub_rn(__int_as_float(0x4B400000 + s), 12582912.0f), 1.0f / 32768.0f);
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP): one instruction moves a whole chunk into smem
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Bounded wait: try_wait suspends the thread up to the hint (no issue slots while waiting). A barrier that never completes -- a
// bulk copy that was lost, a byte count that does not match -- must not hang the context forever: after
// ~2^21 expired hints (tens of seconds; a healthy wait is microseconds) the thread traps, the launch fails with a CUDA error,
// and the host surfaces it as SKGPU_ERR_CUDA / NodeState::Failed instead of a silent hang.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
        if (!done && spins > (1u << 21)) asm volatile("trap;");
    }
}

// ---- cp.async (LDGSTS): asynchronous global -> shared copies that need no destination registers
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// rubato interp_lin: (1 - x) * y0 + x * y1, every operation rounded to f32
__device__ __forceinline__ float interp_lin(float frac, float y0, float y1) {
    return __fadd_rn(__fmul_rn(__fsub_rn(1.0f, frac), y0), __fmul_rn(frac, y1));
}

// ---- phase table staged in shared memory (only the used part is copied)
struct SmemPhase {
    double prefix[SK_PREFIX_MAX];
    SkRun runs[SK_RUNS_MAX];
    uint32_t n_out, n_prefix, n_runs, overflow;
};

// cooperative copy global -> smem by `nthreads` threads (tid in [0, nthreads)); sizes come from the slot record,
// so no dependent header load is needed. Caller syncs afterwards.
__device__ __forceinline__ void load_phase_table(SmemPhase *dst, const SkPhaseTable *src, uint32_t n_out, uint32_t n_prefix,
                                                 uint32_t n_runs, uint32_t tid, uint32_t nthreads) {
    const uint32_t np = min(n_prefix, (uint32_t)SK_PREFIX_MAX), nr = min(n_runs, (uint32_t)SK_RUNS_MAX);
    if (tid == 0) { dst->n_out = n_out; dst->n_prefix = np; dst->n_runs = nr; dst->overflow = 0; }
    const uint2 *sp = reinterpret_cast<const uint2 *>(src->prefix);
    uint2 *dp = reinterpret_cast<uint2 *>(dst->prefix);
    for (uint32_t i = tid; i < np; i += nthreads) dp[i] = sp[i];
    const uint2 *sr = reinterpret_cast<const uint2 *>(src->runs);
    uint2 *dr = reinterpret_cast<uint2 *>(dst->runs);
    for (uint32_t i = tid; i < nr * 3u; i += nthreads) dr[i] = sr[i];
}

// idx of output k from a staged table. Searches the run from the END: the upper binades hold most outputs
// ([512,1024) alone half of them), so the expected number of steps is ~1.
__device__ __forceinline__ double phase_eval_smem(const SmemPhase *T, double t, uint32_t k) {
    if (k < T->n_prefix) return T->prefix[k];
    uint32_t r = T->n_runs - 1u;
    while (r > 0u && T->runs[r].k_a > k) --r;
    const SkRun rn = T->runs[r];
    if (k < rn.k_e) return __fma_rn((double)(k - rn.k_a), rn.delta, rn.x_a);
    return __dadd_rn(__fma_rn((double)(rn.k_e - 1u - rn.k_a), rn.delta, rn.x_a), t);  // gap element
}

// split idx into buffer position (floor(idx) + 16 = start_idx + 2*POLYNOMIAL_LEN) and f32 fraction
__device__ __forceinline__ void phase_split(double x, uint32_t &p, float &frac) {
    const int fl = __double2int_rd(x);
    frac = __double2float_rn(__dsub_rn(x, (double)fl));  // T::coerce(idx - floor(idx))
    p = (uint32_t)(fl + 16);
}

// idx of 4 consecutive outputs k0..k0+3 (the caller uses the first n). Inside one run consecutive members differ by
// exactly delta (all values are representable multiples of the binade's unit), so 1 fma + 3 adds; otherwise 4 lookups.
__device__ __forceinline__ void phase_eval4(const SmemPhase *T, double t, uint32_t k0, uint32_t n, double *x) {
    if (k0 >= T->n_prefix && T->n_runs > 0u) {
        uint32_t r = T->n_runs - 1u;
        while (r > 0u && T->runs[r].k_a > k0) --r;
        const SkRun rn = T->runs[r];
        if (k0 + 3u < rn.k_e) {
            x[0] = __fma_rn((double)(k0 - rn.k_a), rn.delta, rn.x_a);
            x[1] = __dadd_rn(x[0], rn.delta);
            x[2] = __dadd_rn(x[1], rn.delta);
            x[3] = __dadd_rn(x[2], rn.delta);
            return;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = phase_eval_smem(T, t, k0 + min((uint32_t)i, n - 1u));
}

// (re)configures stream slots from the host: fresh FastFixedIn = zero history, last_index = -4.0 (rubato new())
__global__ void k_config_slots(const SlotCfgUpload *__restrict__ cfgs, uint32_t n, SlotTables st) {
    const uint32_t i = blockIdx.x;
    if (i >= n) return;
    const SlotCfgUpload c = cfgs[i];
    const uint32_t slot = c.slot;
    for (uint32_t s = threadIdx.x; s < 16u * st.max_channels; s += blockDim.x) st.hist[(size_t)slot * 16u * st.max_channels + s] = 0.0f;
    if ((c.flags & SLOT_SINC) && st.sinc_hist)
        for (uint32_t s = threadIdx.x; s < st.sinc_H * st.max_channels; s += blockDim.x) st.sinc_hist[(size_t)slot * st.sinc_H * st.max_channels + s] = 0.0f;
    {   // both side records (phase tables / chain history) start zeroed
        uint4 *sd = reinterpret_cast<uint4 *>(slot_side(st, slot, 0));
        for (uint32_t s = threadIdx.x; s < 2u * SK_SIDE_STRIDE / 16u; s += blockDim.x) sd[s] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x == 0) {
        SlotRec r;
        r.t_ratio = c.t_ratio;
        r.last_index = c.last_index0;
        r.chunk = c.chunk;
        r.channels = c.channels;
        r.chunk_count = 0;
        r.carry = (c.flags & SLOT_SINC) ? c.aux : 0u;
        r.n_out[0] = r.n_out[1] = 0;
        r.n_prefix[0] = r.n_prefix[1] = 0;
        r.n_runs[0] = r.n_runs[1] = 0;
        r.end_idx = c.end_idx;
        r.overflow = 0;
        r.flags = c.flags;
        st.rec[slot] = r;
        if (st.fifo_w) { st.fifo_w[slot] = 0ull; st.fifo_r[slot] = 0ull; }
    }
}

__global__ void k_l2_flush(uint4 *buf, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_uint4((uint32_t)i, 0u, 0u, 0u);
}

}  // namespace skgpu
