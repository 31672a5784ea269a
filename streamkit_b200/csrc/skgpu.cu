// skgpu.cu -- implementation of the batch C ABI (include/skgpu_batch.h) over the kernels in kernels.cuh.
//
// Shape of the runtime: one context per GPU (one CUDA stream, per-stream resampler state in HBM as SoA
// tables), plans = compiled ticks (device-resident descriptor tables + one device arena + an optional
// captured CUDA graph), and an asynchronous submit that enqueues H2D -> kernels -> D2H on the context
// stream. Nothing here ever computes audio on the CPU: if CUDA is unavailable every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include "kernels.cuh"

using namespace skgpu;

// ------------------------------------------------------------------ errors (borrowed TLS string, like
// sdks/plugin-sdk/native/src/conversions.rs:441-461 error_to_c)

static thread_local std::string g_err;

static skgpu_rc fail(skgpu_rc rc, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return rc;
}

#define CU(expr)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) return fail(SKGPU_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

extern "C" uint32_t skgpu_abi_version(void) { return SKGPU_ABI_VERSION; }
extern "C" const char *skgpu_last_error(void) { return g_err.c_str(); }

// ------------------------------------------------------------------ context

static constexpr int DYN_RING = 4;  // staging ring depth for per-tick dynamic tables

struct skgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream_d2h = nullptr;   // result read-back runs here so it overlaps the next tick's upload (PCIe is full duplex)
    cudaStream_t stream_k = nullptr;     // sliced ticks: k_chain runs here, uploads stay on `stream`
    cudaStream_t stream_p = nullptr;     // sliced ticks: k_phase_chain (data independent) runs here, ahead of / next to k_chain
    int numa_node = -1;                  // host NUMA node of the GPU's PCI function (-1: unknown / single node)
    std::vector<int> node_cpus;          // CPUs of that node
    struct PinnedBlock { size_t bytes; int kind; };   // kind 0: cudaHostAlloc, 1: mmap + cudaHostRegister
    std::map<void *, PinnedBlock> pinned;
    std::mutex pinned_mu;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;
    skgpu_ctx_config cfg{};
    SlotTables st{};
    // host mirrors of slot configuration
    std::vector<double> h_t;
    std::vector<int32_t> h_end;
    std::vector<uint32_t> h_chunk, h_ch, h_flags;   // h_flags: SLOT_*
    std::vector<uint32_t> h_pe;                     // sinc streams: outputs after which the sub-phase repeats (k_resample_sinc_tiled)
    std::vector<double> h_li0;                      // initial last_index of the slot's resampler
    // windowed-sinc mode
    uint32_t sinc_L = 0, sinc_O = 0;
    double sinc_cutoff = 0.0;
    std::vector<double> sinc_fc;                    // cutoff of tap table i
    std::vector<float *> sinc_tab_dev;
    const float **d_sinc_tabs = nullptr;            // device array of the table pointers (capacity 256)
    std::vector<uint8_t> used;
    std::vector<uint32_t> free_list;
    uint32_t next_fresh = 0;
    std::vector<uint32_t> reset_list;   // slots to (re)configure + reset on the device before the next launch
    SlotCfgUpload *d_reset = nullptr;
    uint32_t d_reset_cap = 0;
    uint4 *l2buf = nullptr;
    size_t l2n = 0;
    int sm_count = 0;
};

template <typename T>
static cudaError_t dalloc(T **p, size_t n) {
    return cudaMalloc((void **)p, n * sizeof(T));
}

static skgpu_rc ctx_flush(skgpu_ctx *c) {
    // (re)configure freshly opened / reset slots on the device: one 24-byte record per slot, one small kernel
    if (!c->reset_list.empty()) {
        const uint32_t n = (uint32_t)c->reset_list.size();
        std::vector<SlotCfgUpload> up(n);
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t slot = c->reset_list[i];
            up[i].t_ratio = c->h_t[slot];
            up[i].slot = slot;
            up[i].chunk = c->h_chunk[slot];
            up[i].channels = c->h_ch[slot];
            up[i].end_idx = c->h_end[slot];
            up[i].flags = c->h_flags[slot];
            up[i].aux = c->h_pe[slot];
            up[i].last_index0 = c->h_li0[slot];
        }
        if (n > c->d_reset_cap) {
            if (c->d_reset) cudaFree(c->d_reset);
            c->d_reset_cap = std::max<uint32_t>(n, 1024u);
            CU(dalloc(&c->d_reset, c->d_reset_cap));
        }
        CU(cudaMemcpyAsync(c->d_reset, up.data(), n * sizeof(SlotCfgUpload), cudaMemcpyHostToDevice, c->stream));
        k_config_slots<<<n, 128, 0, c->stream>>>(c->d_reset, n, c->st);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));  // `up` is pageable host memory
        c->reset_list.clear();
    }
    return SKGPU_OK;
}

static void detect_numa(skgpu_ctx *c);

extern "C" skgpu_rc skgpu_ctx_create(int32_t device_ordinal, const skgpu_ctx_config *cfg, skgpu_ctx **out) {
    if (!cfg || !out) return fail(SKGPU_ERR_INVALID, "skgpu_ctx_create: null argument");
    if (cfg->max_streams == 0) return fail(SKGPU_ERR_INVALID, "max_streams must be > 0");
    if (cfg->max_channels == 0 || cfg->max_channels > 8) return fail(SKGPU_ERR_INVALID, "max_channels must be in 1..8");
    if (cfg->fifo_frames != 0 && (cfg->fifo_frames & (cfg->fifo_frames - 1)) != 0)
        return fail(SKGPU_ERR_INVALID, "fifo_frames must be a power of two");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(SKGPU_ERR_NODEVICE, "no CUDA device available (%s); streamkit_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device_ordinal < 0 || device_ordinal >= ndev) return fail(SKGPU_ERR_INVALID, "device ordinal %d out of range (have %d)", device_ordinal, ndev);
    CU(cudaSetDevice(device_ordinal));
    skgpu_ctx *c = new (std::nothrow) skgpu_ctx();
    if (!c) return fail(SKGPU_ERR_NOMEM, "out of host memory");
    c->device = device_ordinal;
    c->cfg = *cfg;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device_ordinal));
    c->sm_count = prop.multiProcessorCount;
    {   // sliced ticks: a slice's kernels and read-back are on the latency path, the bulk upload of the NEXT slices is not --
        // give the device's scheduler that order of preference
        int lo = 0, hi = 0;   // numerically lower = higher priority
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, lo));
        CU(cudaStreamCreateWithPriority(&c->stream_d2h, cudaStreamNonBlocking, hi));
        CU(cudaStreamCreateWithPriority(&c->stream_k, cudaStreamNonBlocking, hi));
        CU(cudaStreamCreateWithPriority(&c->stream_p, cudaStreamNonBlocking, hi));
    }
    detect_numa(c);
    CU(cudaEventCreate(&c->tm0));
    CU(cudaEventCreate(&c->tm1));
    const size_t S = cfg->max_streams;
    SlotTables &st = c->st;
    st.max_channels = cfg->max_channels;
    st.fifo_frames = cfg->fifo_frames;
    CU(dalloc(&st.rec, S));
    CU(dalloc(&st.hist, S * 16 * cfg->max_channels));
    CU(dalloc(&st.side, S * 2 * SK_SIDE_STRIDE));
    CU(cudaMemset(st.rec, 0, S * sizeof(SlotRec)));
    CU(cudaMemset(st.hist, 0, S * 16 * cfg->max_channels * sizeof(float)));
    CU(cudaMemset(st.side, 0, S * 2 * SK_SIDE_STRIDE));
    if (cfg->fifo_frames) {
        CU(dalloc(&st.fifo, S * cfg->fifo_frames * cfg->max_channels));
        CU(dalloc(&st.fifo_w, S));
        CU(dalloc(&st.fifo_r, S));
        CU(cudaMemset(st.fifo_w, 0, S * sizeof(unsigned long long)));
        CU(cudaMemset(st.fifo_r, 0, S * sizeof(unsigned long long)));
    }
    c->h_t.assign(S, 1.0);
    c->h_end.assign(S, 0);
    c->h_chunk.assign(S, 0);
    c->h_ch.assign(S, 0);
    c->h_flags.assign(S, 0);
    c->h_li0.assign(S, -4.0);
    c->h_pe.assign(S, 0u);
    c->used.assign(S, 0);
    *out = c;
    return SKGPU_OK;
}

extern "C" void skgpu_ctx_destroy(skgpu_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    SlotTables &st = c->st;
#define TEARDOWN(call)                                                                                                  \
    do {                                                                                                                \
        const cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) { fprintf(stderr, "streamkit_b200: %s failed during teardown: %s\n", #call, cudaGetErrorString(e__)); cudaGetLastError(); } \
    } while (0)
    TEARDOWN(cudaFree(st.rec)); TEARDOWN(cudaFree(st.hist)); TEARDOWN(cudaFree(st.side));
    if (st.fifo) { TEARDOWN(cudaFree(st.fifo)); TEARDOWN(cudaFree(st.fifo_w)); TEARDOWN(cudaFree(st.fifo_r)); }
    if (c->d_reset) TEARDOWN(cudaFree(c->d_reset));
    if (st.sinc_hist) TEARDOWN(cudaFree(st.sinc_hist));
    for (float *t : c->sinc_tab_dev) TEARDOWN(cudaFree(t));
    if (c->d_sinc_tabs) TEARDOWN(cudaFree((void *)c->d_sinc_tabs));
    if (c->l2buf) TEARDOWN(cudaFree(c->l2buf));
    TEARDOWN(cudaEventDestroy(c->tm0)); TEARDOWN(cudaEventDestroy(c->tm1));
    for (auto &kv : c->pinned) {
        if (kv.second.kind == 1) { TEARDOWN(cudaHostUnregister(kv.first)); munmap(kv.first, kv.second.bytes); }
        else TEARDOWN(cudaFreeHost(kv.first));
    }
    cudaStreamDestroy(c->stream_k);
    cudaStreamDestroy(c->stream_p);
    cudaStreamDestroy(c->stream_d2h);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" skgpu_rc skgpu_ctx_device_info(skgpu_ctx *c, char *name, size_t name_len, int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, c->device));
    if (name && name_len) snprintf(name, name_len, "%s", prop.name);
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return SKGPU_OK;
}

// ---- NUMA placement of the pinned tick arenas. No libnuma in the image: sysfs + the mbind syscall.
static void detect_numa(skgpu_ctx *c) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, c->device) != cudaSuccess) { cudaGetLastError(); return; }
    for (char *p = bus; *p; ++p) *p = (char)tolower(*p);
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    if (node < 0) return;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f) return;
    char buf[4096] = {0};
    if (!fgets(buf, sizeof buf, f)) buf[0] = 0;
    fclose(f);
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    sched_getaffinity(0, sizeof allowed, &allowed);
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int k = a; k <= b; ++k) if (k < CPU_SETSIZE && CPU_ISSET(k, &allowed)) c->node_cpus.push_back(k); }
        else if (sscanf(tok, "%d", &a) == 1) { if (a < CPU_SETSIZE && CPU_ISSET(a, &allowed)) c->node_cpus.push_back(a); }
    }
    c->numa_node = node;
}

extern "C" int32_t skgpu_ctx_numa_node(skgpu_ctx *c) { return c ? c->numa_node : -1; }

extern "C" skgpu_rc skgpu_ctx_bind_thread(skgpu_ctx *c) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    if (c->node_cpus.empty()) return SKGPU_OK;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int k : c->node_cpus) CPU_SET(k, &set);
    if (sched_setaffinity(0, sizeof set, &set) != 0) return fail(SKGPU_ERR_STATE, "sched_setaffinity to NUMA node %d failed", c->numa_node);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_pinned_alloc_ex(skgpu_ctx *c, size_t bytes, uint32_t flags, void **out, int32_t *node_out) {
    if (!c || !out) return fail(SKGPU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->device));
    if (node_out) *node_out = -1;
    if (bytes == 0) bytes = 16;
    if ((flags & SKGPU_PIN_NUMA_LOCAL) && c->numa_node >= 0 && !(flags & SKGPU_PIN_WRITE_COMBINED)) {
        // map, bind to the GPU's node BEFORE the first touch, ask for huge pages, then pin: cudaHostRegister faults the
        // pages in under the MPOL_BIND policy, so every page of the arena is local to the GPU's PCIe root
        const size_t len = (bytes + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1);
        void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask[16] = {0};
            mask[c->numa_node / (8 * sizeof(unsigned long))] |= 1ul << (c->numa_node % (8 * sizeof(unsigned long)));
            const long rc = syscall(SYS_mbind, p, len, 2 /* MPOL_BIND */, mask, sizeof(mask) * 8, 0);
            madvise(p, len, MADV_HUGEPAGE);
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                std::lock_guard<std::mutex> lk(c->pinned_mu);
                c->pinned[p] = {len, 1};
                *out = p;
                if (node_out) *node_out = rc == 0 ? c->numa_node : -1;
                return SKGPU_OK;
            }
            cudaGetLastError();
            munmap(p, len);   // fall through to the plain allocation
        }
    }
    // plain path: first touch on a thread bound to the GPU's node keeps the pages local under the default policy
    cpu_set_t saved;
    bool rebound = false;
    if ((flags & SKGPU_PIN_NUMA_LOCAL) && !c->node_cpus.empty() && sched_getaffinity(0, sizeof saved, &saved) == 0) {
        cpu_set_t set;
        CPU_ZERO(&set);
        for (int k : c->node_cpus) CPU_SET(k, &set);
        rebound = sched_setaffinity(0, sizeof set, &set) == 0;
    }
    const cudaError_t e = cudaHostAlloc(out, bytes, (flags & SKGPU_PIN_WRITE_COMBINED) ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (rebound) sched_setaffinity(0, sizeof saved, &saved);
    if (e != cudaSuccess) return fail(SKGPU_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    {
        std::lock_guard<std::mutex> lk(c->pinned_mu);
        c->pinned[*out] = {bytes, 0};
    }
    if (node_out && rebound) *node_out = c->numa_node;
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_pinned_alloc(skgpu_ctx *c, size_t bytes, void **out) { return skgpu_pinned_alloc_ex(c, bytes, SKGPU_PIN_NUMA_LOCAL, out, nullptr); }
extern "C" skgpu_rc skgpu_pinned_free(skgpu_ctx *c, void *p) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    if (!p) return SKGPU_OK;
    skgpu_ctx::PinnedBlock blk{0, 0};
    {
        std::lock_guard<std::mutex> lk(c->pinned_mu);
        auto it = c->pinned.find(p);
        if (it == c->pinned.end()) return fail(SKGPU_ERR_INVALID, "pointer was not allocated by skgpu_pinned_alloc on this context");
        blk = it->second;
        c->pinned.erase(it);
    }
    if (blk.kind == 1) { CU(cudaHostUnregister(p)); munmap(p, blk.bytes); }
    else CU(cudaFreeHost(p));
    return SKGPU_OK;
}

// ------------------------------------------------------------------ windowed-sinc mode: parameters + tap tables

// 4-term Blackman-Harris window, squared (rubato WindowFunction::BlackmanHarris2), on u in (0, 1)
static double sinc_window(double u) {
    const double two_pi = 6.283185307179586476925286766559;
    const double w = 0.35875 - 0.48829 * std::cos(two_pi * u) + 0.14128 * std::cos(2.0 * two_pi * u) - 0.01168 * std::cos(3.0 * two_pi * u);
    return w * w;
}

// index of the tap table for cutoff fc (built and uploaded on first use): T[p][n] = g(L/2 - 1 - n + p/O) / sum_n g, p = 0..O
static int sinc_table_for(skgpu_ctx *c, double fc) {
    for (size_t i = 0; i < c->sinc_fc.size(); ++i)
        if (c->sinc_fc[i] == fc) return (int)i;
    const uint32_t L = c->sinc_L, O = c->sinc_O;
    const double pi = 3.14159265358979323846264338327950288;
    const uint32_t LS = L + SINC_ROW_PAD;                       // padded row stride (k_sinc.cuh)
    std::vector<float> tab((size_t)(O + 1u) * LS, 0.0f);
    std::vector<double> g(L);
    for (uint32_t p = 0; p <= O; ++p) {
        double sum = 0.0;
        for (uint32_t n = 0; n < L; ++n) {
            const double tau = (double)L / 2.0 - 1.0 - (double)n + (double)p / (double)O;
            const double z = fc * tau;
            const double sc = z == 0.0 ? 1.0 : std::sin(pi * z) / (pi * z);
            const double u = (tau + (double)L / 2.0) / (double)L;
            const double w = (u <= 0.0 || u >= 1.0) ? 0.0 : sinc_window(u);
            g[n] = fc * sc * w;
            sum += g[n];
        }
        for (uint32_t n = 0; n < L; ++n) tab[(size_t)p * LS + n] = (float)(g[n] / sum);
    }
    float *d = nullptr;
    if (cudaMalloc((void **)&d, tab.size() * sizeof(float)) != cudaSuccess) return 0;
    cudaMemcpy(d, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice);
    c->sinc_fc.push_back(fc);
    c->sinc_tab_dev.push_back(d);
    const float *dp = d;
    cudaMemcpy(c->d_sinc_tabs + (c->sinc_fc.size() - 1), &dp, sizeof(dp), cudaMemcpyHostToDevice);
    return (int)c->sinc_fc.size() - 1;
}

extern "C" skgpu_rc skgpu_ctx_set_sinc(skgpu_ctx *c, uint32_t sinc_len, uint32_t oversampling, double f_cutoff) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    if (c->sinc_L) return fail(SKGPU_ERR_STATE, "sinc parameters are fixed for the context's lifetime");
    if (sinc_len < 8 || sinc_len > 256 || sinc_len % 8) return fail(SKGPU_ERR_INVALID, "sinc_len must be a multiple of 8 in 8..256");
    if (oversampling < 1 || oversampling > 1024) return fail(SKGPU_ERR_INVALID, "oversampling_factor must be in 1..1024");
    if (!(f_cutoff > 0.0 && f_cutoff <= 1.0)) return fail(SKGPU_ERR_INVALID, "f_cutoff must be in (0, 1]");
    CU(cudaSetDevice(c->device));
    SlotTables &st = c->st;
    const uint32_t H = sinc_len + 8u;
    const size_t n = (size_t)c->cfg.max_streams * H * c->cfg.max_channels;
    CU(dalloc(&st.sinc_hist, n));
    CU(cudaMemset(st.sinc_hist, 0, n * sizeof(float)));
    CU(cudaMalloc((void **)&c->d_sinc_tabs, 256 * sizeof(float *)));
    CU(cudaMemset(c->d_sinc_tabs, 0, 256 * sizeof(float *)));
    st.sinc_tabs = c->d_sinc_tabs;
    st.sinc_L = sinc_len; st.sinc_O = oversampling; st.sinc_H = H;
    c->sinc_L = sinc_len; c->sinc_O = oversampling; c->sinc_cutoff = f_cutoff;
    return SKGPU_OK;
}

// ------------------------------------------------------------------ stream slots

static skgpu_rc validate_stream_cfg(const skgpu_ctx *c, const skgpu_stream_cfg *s) {
    if (!s) return fail(SKGPU_ERR_INVALID, "null stream config");
    if (s->out_rate == 0) return fail(SKGPU_ERR_INVALID, "target_sample_rate must be greater than 0");  // resampler.rs:82-86
    if (s->in_rate == 0) return fail(SKGPU_ERR_INVALID, "input sample rate must be greater than 0");
    if (s->chunk_frames == 0) return fail(SKGPU_ERR_INVALID, "chunk_frames must be greater than 0");     // resampler.rs:88-92
    if (s->chunk_frames > (1u << 20)) return fail(SKGPU_ERR_INVALID, "chunk_frames too large");
    if (s->channels == 0 || s->channels > c->cfg.max_channels)
        return fail(SKGPU_ERR_INVALID, "channels %u outside 1..%u", (unsigned)s->channels, c->cfg.max_channels);
    if (s->flags & ~(uint32_t)(SKGPU_STREAM_S16 | SKGPU_STREAM_SINC)) return fail(SKGPU_ERR_INVALID, "unknown stream flags 0x%x", (unsigned)s->flags);
    if (s->flags & SKGPU_STREAM_SINC) {
        if (!c->sinc_L) return fail(SKGPU_ERR_STATE, "sinc streams need skgpu_ctx_set_sinc first");
        if (s->flags & SKGPU_STREAM_S16) return fail(SKGPU_ERR_INVALID, "sinc streams take f32 chunks");
        if (s->in_rate == s->out_rate) return fail(SKGPU_ERR_INVALID, "input rate equals the target rate: nothing to resample");
        if (s->chunk_frames < c->sinc_L + 4u) return fail(SKGPU_ERR_INVALID, "chunk_frames %u too short for sinc_len %u", s->chunk_frames, c->sinc_L);
        if (c->sinc_fc.size() >= 255) return fail(SKGPU_ERR_NOMEM, "too many distinct sinc cutoffs");
    }
    if ((s->flags & SKGPU_STREAM_S16) && (((uint64_t)s->chunk_frames * s->channels) % 2u || (uint64_t)(s->chunk_frames + 32u) * s->channels > 4096u || s->channels > 2))
        return fail(SKGPU_ERR_INVALID, "s16 streams need mono / stereo chunks of an even number of samples, at most 4096 samples with the 32-frame head");
    const double ratio = (double)s->out_rate / (double)s->in_rate;
    if (!(ratio >= 1.0 / 256.0 && ratio <= 256.0)) return fail(SKGPU_ERR_INVALID, "resample ratio %g outside [1/256, 256]", ratio);
    return SKGPU_OK;
}

static void slot_configure(skgpu_ctx *c, uint32_t slot, const skgpu_stream_cfg *s) {
    const double ratio = (double)s->out_rate / (double)s->in_rate;  // resampler.rs:233 f64::from(out) / f64::from(in)
    const double t = 1.0 / ratio;                                    // rubato: t_ratio = 1.0 / resample_ratio
    c->h_t[slot] = t;
    c->h_end[slot] = (int32_t)s->chunk_frames - 9 - (int32_t)std::ceil(t);  // chunk - (POLYNOMIAL_LEN + 1) - ceil(t)
    c->h_chunk[slot] = s->chunk_frames;
    c->h_ch[slot] = s->channels;
    c->h_flags[slot] = (s->in_rate == s->out_rate ? SLOT_BYPASS : 0u) | ((s->flags & SKGPU_STREAM_S16) ? SLOT_S16 : 0u);
    c->h_li0[slot] = -4.0;                                            // rubato FastFixedIn: -(POLYNOMIAL_LEN / 2)
    if (s->flags & SKGPU_STREAM_SINC) {
        const int tab = sinc_table_for(c, c->sinc_cutoff * std::min(1.0, ratio));   // validated by the caller
        c->h_flags[slot] |= SLOT_SINC | ((uint32_t)tab << 8);
        c->h_end[slot] = (int32_t)s->chunk_frames - (int32_t)(c->sinc_L / 2u) - 1 - (int32_t)std::ceil(t);
        c->h_li0[slot] = -(double)(c->sinc_L / 2u);
        // out_rate / gcd outputs later the sub-phase is the same again (up to the rounding of the f64 recurrence, which the
        // kernel checks per output); a multiple >= 32 of it keeps a warp's lanes on consecutive outputs
        uint32_t a = s->in_rate, b = s->out_rate;
        while (b) { const uint32_t r = a % b; a = b; b = r; }
        const uint32_t P = s->out_rate / a;
        c->h_pe[slot] = P >= 32u ? P : P * ((32u + P - 1u) / P);
    } else {
        c->h_pe[slot] = 0u;
    }
    c->used[slot] = 1;
    c->reset_list.push_back(slot);
}

extern "C" uint32_t skgpu_stream_max_out_frames(const skgpu_stream_cfg *s) {
    if (!s || !s->in_rate) return 0;
    const double ratio = (double)s->out_rate / (double)s->in_rate;
    return (uint32_t)((double)s->chunk_frames * ratio + 10.0) + 8u;
}

extern "C" skgpu_rc skgpu_stream_open_many(skgpu_ctx *c, const skgpu_stream_cfg *s, uint32_t n, uint32_t *slots_out) {
    if (!c || !slots_out) return fail(SKGPU_ERR_INVALID, "null argument");
    skgpu_rc rc = validate_stream_cfg(c, s);
    if (rc) return rc;
    const size_t avail = c->free_list.size() + (c->cfg.max_streams - c->next_fresh);
    if (n > avail) return fail(SKGPU_ERR_NOMEM, "out of stream slots (%u requested, %zu free of %u)", n, avail, c->cfg.max_streams);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t slot;
        if (c->next_fresh < c->cfg.max_streams) slot = c->next_fresh++;  // fresh slots first: contiguous ranges
        else { slot = c->free_list.back(); c->free_list.pop_back(); }
        slot_configure(c, slot, s);
        slots_out[i] = slot;
    }
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_stream_open(skgpu_ctx *c, const skgpu_stream_cfg *s, uint32_t *slot_out) {
    return skgpu_stream_open_many(c, s, 1, slot_out);
}
extern "C" skgpu_rc skgpu_stream_reset(skgpu_ctx *c, uint32_t slot) {
    if (!c || slot >= c->cfg.max_streams || !c->used[slot]) return fail(SKGPU_ERR_INVALID, "invalid slot %u", slot);
    c->reset_list.push_back(slot);
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_stream_close(skgpu_ctx *c, uint32_t slot) {
    if (!c || slot >= c->cfg.max_streams || !c->used[slot]) return fail(SKGPU_ERR_INVALID, "invalid slot %u", slot);
    c->used[slot] = 0;
    c->free_list.push_back(slot);
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_stream_get_state(skgpu_ctx *c, uint32_t slot, double *last_index, float *hist, uint64_t *fifo_written, uint64_t *fifo_read) {
    if (!c || slot >= c->cfg.max_streams || !c->used[slot]) return fail(SKGPU_ERR_INVALID, "invalid slot %u", slot);
    CU(cudaSetDevice(c->device));
    skgpu_rc rc = ctx_flush(c);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    if (last_index) {
        SlotRec rec;
        CU(cudaMemcpy(&rec, c->st.rec + slot, sizeof(SlotRec), cudaMemcpyDeviceToHost));
        *last_index = rec.last_index;
    }
    if (hist) CU(cudaMemcpy(hist, c->st.hist + (size_t)slot * 16 * c->cfg.max_channels, 16 * c->h_ch[slot] * sizeof(float), cudaMemcpyDeviceToHost));
    if (fifo_written) { *fifo_written = 0; if (c->st.fifo_w) CU(cudaMemcpy(fifo_written, c->st.fifo_w + slot, 8, cudaMemcpyDeviceToHost)); }
    if (fifo_read) { *fifo_read = 0; if (c->st.fifo_r) CU(cudaMemcpy(fifo_read, c->st.fifo_r + slot, 8, cudaMemcpyDeviceToHost)); }
    return SKGPU_OK;
}

// ------------------------------------------------------------------ plan

enum OpKind { OP_CONVERT = 0, OP_RESAMPLE = 1, OP_MIX = 2, OP_CHAIN = 3 };

struct DynTable {  // small per-tick table with a ring of pinned staging buffers
    void *dev = nullptr;
    void *host[DYN_RING] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done[DYN_RING] = {nullptr, nullptr, nullptr, nullptr};
    bool used[DYN_RING] = {false, false, false, false};
    size_t bytes = 0;
    int cur = 0;
    bool dirty = false;
    bool valid = false;
};

struct Op {
    OpKind kind;
    int mode = 0;
    uint32_t cap = 0, cap2 = 0;   // table capacities
    uint32_t n = 0, n2 = 0;       // live entries
    void *d_tab = nullptr, *d_tab2 = nullptr;
    void *h_tab = nullptr, *h_tab2 = nullptr;  // pinned staging
    OpHeader *d_hdr = nullptr;
    OpHeader *h_hdr = nullptr;                 // pinned
    bool dirty = true;
    uint32_t tiles = 1;           // convert: tiles per segment; mix: tiles per group
    uint32_t max_unit = 0;        // convert: max n_samples; mix: max out samples; resample: max chunk frames
    uint32_t smem_frames = 0;     // resample: frames (history + chunk) the staged path can hold
    uint32_t smem_bytes = 0;
    int rs_channels = 0;          // resample: 1 / 2 specialisation, 0 = generic
    uint32_t mix_tpc = 1;         // mix: tiles one CTA handles (whole group for small groups)
    bool rs_prog = false;         // resample: program-driven kernels (k_phase_prog + k_resample_prog)
    bool rs_sinc = false;         // resample: windowed-sinc streams (k_phase + k_resample_sinc)
    bool rs_sinc_tiled = false;   // ... through the persistent tap-table-in-shared-memory kernel (one tap table, fits)
    SincDims sinc_dm{};
    const float *sinc_taps = nullptr;
    ChainProgDims rs_pd{};        // resample: frame-program capacities of the op
    uint64_t results_off = 0;
    bool has_fifo_inputs = false;
    uint32_t chain_F = 0;         // chain: output_frame_size
    uint32_t chain_kb = 0;        // chain: inputs staged per batch
    uint32_t chain_buf_floats = 0;  // chain: floats per staging buffer
    int chain_oc = 2;             // chain: output channels specialisation
    int chain_iters = 1;          // chain: ceil(F / 1024)
    int chain_kind = CHAIN_ANY;   // chain: CHAIN_PLAIN / _BYPASS / _PLAIN_S16 / _BYPASS_S16 when every input is of that one kind (specialised kernels), else CHAIN_ANY
    ChainDims chain_dm{};         // chain: staging-ring geometry passed to the kernel
    ChainRec *d_rec = nullptr;    // chain: per-input records written by k_phase_chain every tick
    uint32_t chain_grid = 0;      // chain: persistent grid size (CTAs per SM x SMs)
    DynTable present;             // mix / chain: per-input presence
    // sliced ticks (chain op)
    std::vector<skgpu_slice> slices;
    std::vector<uint32_t> last_reader;   // per slice i: the last slice whose kernels read bytes of upload piece i
    OpHeader *d_hdr_sl = nullptr, *h_hdr_sl = nullptr;   // one launch header per slice
    uint32_t hdr_sl_cap = 0;
    bool slices_dirty = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[2];
    uint32_t ev_used[2] = {0, 0};
};

struct skgpu_plan {
    skgpu_ctx *ctx = nullptr;
    uint8_t *arena = nullptr;
    size_t arena_bytes = 0;
    std::vector<Op> ops;
    DynTable gains;
    uint32_t n_gains = 0;
    uint64_t h2d_off = 0, h2d_bytes = 0, d2h_off = 0, d2h_bytes = 0;
    uint64_t bank_stride = 0;     // 0 = single-banked H2D
    uint64_t tick = 0;            // ticks submitted (host mirror of d_tick)
    uint32_t *d_tick = nullptr;
    bool finalized = false;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr, e3 = nullptr;
    cudaEvent_t ev_kernels_done = nullptr, ev_d2h_done = nullptr;  // cross-stream ordering for overlapped read-back
    cudaEvent_t ev_tick_done[2] = {nullptr, nullptr};              // completion of tick k (everything incl. read-back), by k & 1
    bool d2h_pending = false;
    bool timing_valid = false;
    // sliced ticks: events per (tick parity, slice); *_n = slices of the tick submitted with that parity
    std::vector<cudaEvent_t> ev_up[2], ev_k[2], ev_done[2];
    cudaEvent_t ev_t0[2] = {nullptr, nullptr}, ev_tables = nullptr, ev_k_all = nullptr, ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_p;   // phase(i) done
    cudaGraph_t sl_graph = nullptr;           // kernels of a sliced tick without transfers
    cudaGraphExec_t sl_graph_exec = nullptr;
    uint32_t sl_n[2] = {0, 0};
    bool sliced_pending = false;   // the previous submit was sliced (its kernels ran on stream_k)
};

static skgpu_rc dyn_alloc(DynTable &t, size_t bytes) {
    t.bytes = bytes;
    CU(cudaMalloc(&t.dev, bytes ? bytes : 16));
    for (int i = 0; i < DYN_RING; ++i) {
        CU(cudaHostAlloc(&t.host[i], bytes ? bytes : 16, cudaHostAllocDefault));
        CU(cudaEventCreateWithFlags(&t.done[i], cudaEventDisableTiming));
    }
    t.valid = true;
    return SKGPU_OK;
}
static void dyn_free(DynTable &t) {
    if (!t.valid) return;
    cudaFree(t.dev);
    for (int i = 0; i < DYN_RING; ++i) { cudaFreeHost(t.host[i]); cudaEventDestroy(t.done[i]); }
    t.valid = false;
}
// host writes the next staging buffer (waiting for its previous upload if still in flight)
static skgpu_rc dyn_write(DynTable &t, const void *src, size_t bytes) {
    const int nxt = (t.cur + 1) % DYN_RING;
    if (t.used[nxt]) CU(cudaEventSynchronize(t.done[nxt]));
    memcpy(t.host[nxt], src, bytes);
    t.cur = nxt;
    t.dirty = true;
    return SKGPU_OK;
}
static skgpu_rc dyn_upload(DynTable &t, cudaStream_t s) {
    if (!t.valid || !t.dirty) return SKGPU_OK;
    CU(cudaMemcpyAsync(t.dev, t.host[t.cur], t.bytes, cudaMemcpyHostToDevice, s));
    CU(cudaEventRecord(t.done[t.cur], s));
    t.used[t.cur] = true;
    t.dirty = false;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_create(skgpu_ctx *c, size_t arena_bytes, skgpu_plan **out) {
    if (!c || !out) return fail(SKGPU_ERR_INVALID, "null argument");
    if (arena_bytes == 0) return fail(SKGPU_ERR_INVALID, "arena_bytes must be > 0");
    CU(cudaSetDevice(c->device));
    skgpu_plan *p = new (std::nothrow) skgpu_plan();
    if (!p) return fail(SKGPU_ERR_NOMEM, "out of host memory");
    p->ctx = c;
    p->arena_bytes = arena_bytes;
    cudaError_t e = cudaMalloc((void **)&p->arena, arena_bytes);
    if (e != cudaSuccess) { delete p; return fail(SKGPU_ERR_NOMEM, "cudaMalloc(%zu) for the tick arena failed: %s", arena_bytes, cudaGetErrorString(e)); }
    CU(cudaMemset(p->arena, 0, arena_bytes));   // padding between frames and never-written capacity read back as zeros, not garbage
    CU(cudaEventCreate(&p->e0)); CU(cudaEventCreate(&p->e1)); CU(cudaEventCreate(&p->e2)); CU(cudaEventCreate(&p->e3));
    CU(cudaEventCreateWithFlags(&p->ev_kernels_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p->ev_d2h_done, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p->ev_tick_done[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&p->ev_tick_done[1], cudaEventDisableTiming));
    CU(cudaMalloc((void **)&p->d_tick, 16));
    CU(cudaMemset(p->d_tick, 0, 16));
    *out = p;
    return SKGPU_OK;
}

static void op_free(Op &op) {
    cudaFree(op.d_tab); cudaFree(op.d_tab2); cudaFree(op.d_hdr); cudaFree(op.d_rec); cudaFree(op.d_hdr_sl);
    if (op.h_hdr_sl) cudaFreeHost(op.h_hdr_sl);
    if (op.h_tab) cudaFreeHost(op.h_tab);
    if (op.h_tab2) cudaFreeHost(op.h_tab2);
    if (op.h_hdr) cudaFreeHost(op.h_hdr);
    dyn_free(op.present);
    for (int s = 0; s < 2; ++s)
        for (auto &pr : op.ev[s]) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
}

extern "C" void skgpu_plan_destroy(skgpu_plan *p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    cudaStreamSynchronize(p->ctx->stream_k);
    cudaStreamSynchronize(p->ctx->stream_d2h);
    for (auto &op : p->ops) op_free(op);
    dyn_free(p->gains);
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    if (p->graph) cudaGraphDestroy(p->graph);
    { cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) fprintf(stderr, "skgpu_plan_destroy: stale CUDA error: %s\n", cudaGetErrorString(e)); }
    cudaFree(p->arena);
    cudaFree(p->d_tick);
    cudaEventDestroy(p->ev_kernels_done); cudaEventDestroy(p->ev_d2h_done);
    cudaEventDestroy(p->ev_tick_done[0]); cudaEventDestroy(p->ev_tick_done[1]);
    cudaEventDestroy(p->e0); cudaEventDestroy(p->e1); cudaEventDestroy(p->e2); cudaEventDestroy(p->e3);
    for (int a = 0; a < 2; ++a) {
        for (auto e : p->ev_up[a]) cudaEventDestroy(e);
        for (auto e : p->ev_k[a]) cudaEventDestroy(e);
        for (auto e : p->ev_done[a]) cudaEventDestroy(e);
        if (p->ev_t0[a]) cudaEventDestroy(p->ev_t0[a]);
    }
    if (p->ev_tables) cudaEventDestroy(p->ev_tables);
    if (p->ev_k_all) cudaEventDestroy(p->ev_k_all);
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    for (auto e : p->ev_p) cudaEventDestroy(e);
    if (p->sl_graph_exec) cudaGraphExecDestroy(p->sl_graph_exec);
    if (p->sl_graph) cudaGraphDestroy(p->sl_graph);
    delete p;
}

static skgpu_rc check_range(const skgpu_plan *p, uint64_t off, uint64_t bytes, const char *what) {
    if (off > p->arena_bytes || bytes > p->arena_bytes - off)
        return fail(SKGPU_ERR_INVALID, "%s: range [%llu, +%llu) outside the %zu-byte arena", what, (unsigned long long)off, (unsigned long long)bytes, p->arena_bytes);
    return SKGPU_OK;
}

static skgpu_rc op_alloc_tables(Op &op, size_t entry, uint32_t cap, size_t entry2, uint32_t cap2) {
    op.cap = cap;
    op.cap2 = cap2;
    CU(cudaMalloc(&op.d_tab, std::max<size_t>(entry * cap, 16)));
    CU(cudaHostAlloc(&op.h_tab, std::max<size_t>(entry * cap, 16), cudaHostAllocDefault));
    if (entry2) {
        CU(cudaMalloc(&op.d_tab2, std::max<size_t>(entry2 * cap2, 16)));
        CU(cudaHostAlloc(&op.h_tab2, std::max<size_t>(entry2 * cap2, 16), cudaHostAllocDefault));
    }
    CU(cudaMalloc((void **)&op.d_hdr, sizeof(OpHeader)));
    CU(cudaHostAlloc((void **)&op.h_hdr, sizeof(OpHeader), cudaHostAllocDefault));
    memset(op.h_hdr, 0, sizeof(OpHeader));
    return SKGPU_OK;
}

// ---- convert

static skgpu_rc validate_segs(const skgpu_plan *p, int mode, const skgpu_seg *segs, uint32_t n, uint32_t *max_n) {
    const uint32_t in_b = mode == SKGPU_CVT_S16_TO_F32 ? 2 : 4, out_b = mode == SKGPU_CVT_F32_TO_S16 ? 2 : 4;
    uint32_t mx = 0;
    for (uint32_t i = 0; i < n; ++i) {
        skgpu_rc rc = check_range(p, segs[i].in_off, (uint64_t)segs[i].n_samples * in_b, "convert input");
        if (rc) return rc;
        rc = check_range(p, segs[i].out_off, (uint64_t)segs[i].n_samples * out_b, "convert output");
        if (rc) return rc;
        if ((segs[i].in_off % in_b) || (segs[i].out_off % out_b)) return fail(SKGPU_ERR_INVALID, "convert segment %u: misaligned offset", i);
        if (segs[i].gain_idx != SKGPU_NO_GAIN && segs[i].gain_idx >= p->n_gains)
            return fail(SKGPU_ERR_INVALID, "convert segment %u: gain_idx %u >= gain table size %u (call skgpu_plan_set_gains first)", i, segs[i].gain_idx, p->n_gains);
        mx = std::max(mx, segs[i].n_samples);
    }
    *max_n = mx;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_add_convert(skgpu_plan *p, skgpu_cvt_mode mode, const skgpu_seg *segs, uint32_t n, uint32_t *op_out) {
    if (!p || (!segs && n)) return fail(SKGPU_ERR_INVALID, "null argument");
    if (p->finalized) return fail(SKGPU_ERR_STATE, "plan already finalized");
    if ((int)mode < 0 || (int)mode > 2) return fail(SKGPU_ERR_INVALID, "unknown convert mode %d", (int)mode);
    CU(cudaSetDevice(p->ctx->device));
    uint32_t mx = 0;
    skgpu_rc rc = validate_segs(p, mode, segs, n, &mx);
    if (rc) return rc;
    Op op;
    op.kind = OP_CONVERT;
    op.mode = mode;
    rc = op_alloc_tables(op, sizeof(skgpu_seg), std::max(n, 1u), 0, 0);
    if (rc) return rc;
    if (n) memcpy(op.h_tab, segs, n * sizeof(skgpu_seg));
    op.n = n;
    op.max_unit = std::max(mx, 1u);
    op.tiles = (op.max_unit + CVT_TILE - 1) / CVT_TILE;
    if (op_out) *op_out = (uint32_t)p->ops.size();
    p->ops.push_back(op);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_update_convert(skgpu_plan *p, uint32_t opi, const skgpu_seg *segs, uint32_t n) {
    if (!p || opi >= p->ops.size() || p->ops[opi].kind != OP_CONVERT) return fail(SKGPU_ERR_INVALID, "not a convert op");
    Op &op = p->ops[opi];
    if (n > op.cap) return fail(SKGPU_ERR_INVALID, "update exceeds capacity (%u > %u)", n, op.cap);
    uint32_t mx = 0;
    skgpu_rc rc = validate_segs(p, op.mode, segs, n, &mx);
    if (rc) return rc;
    if (mx > op.tiles * (uint32_t)CVT_TILE) return fail(SKGPU_ERR_INVALID, "segment longer (%u) than the op was sized for (%u)", mx, op.tiles * CVT_TILE);
    CU(cudaStreamSynchronize(p->ctx->stream));
    if (n) memcpy(op.h_tab, segs, n * sizeof(skgpu_seg));
    op.n = n;
    op.dirty = true;
    return SKGPU_OK;
}

// ---- resample

static skgpu_rc validate_rs(const skgpu_plan *p, const skgpu_rs_item *items, uint32_t n, uint32_t *max_chunk, int *chan, bool *any_fifo) {
    const skgpu_ctx *c = p->ctx;
    uint32_t mx = 0;
    int ch = -1;
    bool fifo = false;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t slot = items[i].slot;
        if (slot >= c->cfg.max_streams || !c->used[slot]) return fail(SKGPU_ERR_INVALID, "resample item %u: slot %u is not open", i, slot);
        const uint32_t N = c->h_chunk[slot], C = c->h_ch[slot];
        if (c->h_flags[slot] & SLOT_BYPASS) return fail(SKGPU_ERR_INVALID, "resample item %u: the stream's input rate equals its target rate: the reference bypasses the resampler for such inputs (resampler.rs:299-373); mix them directly or use the chain op", i);
        if (c->h_flags[slot] & SLOT_S16) return fail(SKGPU_ERR_INVALID, "resample item %u: s16 streams are a chain-op feature (convert with SKGPU_CVT_S16_TO_F32 first)", i);
        if (((c->h_flags[slot] & SLOT_SINC) != 0) != ((c->h_flags[items[0].slot] & SLOT_SINC) != 0)) return fail(SKGPU_ERR_INVALID, "resample item %u: linear and sinc streams cannot share one resample op", i);
        if ((c->h_flags[slot] & SLOT_SINC) && (items[i].flags & SKGPU_RS_TO_FIFO)) return fail(SKGPU_ERR_INVALID, "resample item %u: sinc streams write to out_off, not to the device ring", i);
        skgpu_rc rc = check_range(p, items[i].in_off, (uint64_t)N * C * 4, "resample input");
        if (rc) return rc;
        if (items[i].in_off % 4) return fail(SKGPU_ERR_INVALID, "resample item %u: misaligned input", i);
        if (items[i].flags & SKGPU_RS_TO_FIFO) {
            if (!c->cfg.fifo_frames) return fail(SKGPU_ERR_INVALID, "resample item %u: context has no device FIFO (fifo_frames = 0)", i);
            fifo = true;
        } else {
            rc = check_range(p, items[i].out_off, (uint64_t)items[i].out_cap_frames * C * 4, "resample output");
            if (rc) return rc;
            if (items[i].out_off % (C == 2 ? 8 : 4)) return fail(SKGPU_ERR_INVALID, "resample item %u: misaligned output", i);
        }
        mx = std::max(mx, N);
        if (ch == -1) ch = (int)C;
        else if (ch != (int)C) ch = 0;
    }
    *max_chunk = mx;
    *chan = (ch == 1 || ch == 2) ? ch : 0;
    *any_fifo = fifo;
    return SKGPU_OK;
}

// Geometry of the persistent sinc kernel for this op, or rs_sinc_tiled = false (the one-CTA-per-stream kernel then runs):
// every stream must use the same tap table, and the table plus a 2-stage ring of at least one stream must fit.
static void rs_sinc_dims(const skgpu_ctx *c, Op &op, const skgpu_rs_item *items, uint32_t n, int ch) {
    op.rs_sinc_tiled = false;
    if (!n) return;
    const uint32_t tab = (c->h_flags[items[0].slot] >> 8) & 0xFFu;
    uint32_t max_items = 1;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t slot = items[i].slot;
        if (((c->h_flags[slot] >> 8) & 0xFFu) != tab) return;
        const uint32_t pe = std::max(c->h_pe[slot], 1u);
        // the NOMINAL output count sizes the pass (a chunk that yields one frame more costs its stream a second round of items)
        const uint32_t nominal = std::max(1u, (uint32_t)((double)c->h_chunk[slot] / c->h_t[slot] + 0.5));
        const uint32_t per_class = (nominal + pe - 1u) / pe;
        max_items = std::max(max_items, pe * ((per_class + SINC_RA - 1u) / SINC_RA));
    }
    SincDims d;
    d.tab_bytes = (c->sinc_O + 1u) * (c->sinc_L + SINC_ROW_PAD) * 4u;
    const uint32_t pt = ((uint32_t)sizeof(SkPhaseTable) + 15u) & ~15u;
    d.stream_bytes = pt + c->st.sinc_H * (uint32_t)ch * 4u + (((op.max_unit * (uint32_t)ch * 4u) + 15u) & ~15u);
    const uint32_t budget = 220u * 1024u;
    if (d.tab_bytes >= (1u << 20) || d.tab_bytes + 2u * d.stream_bytes > budget) return;
    uint32_t G = std::max(1u, (uint32_t)SINCT_THREADS / max_items);
    G = std::min(G, (uint32_t)SINC_GMAX);
    G = std::min(G, (budget - d.tab_bytes) / (2u * d.stream_bytes));
    d.G = G;
    op.sinc_dm = d;
    op.sinc_taps = c->sinc_tab_dev[tab];
    op.rs_sinc_tiled = true;
}

// Frame-program capacities of a resample op (k_resample_prog): the same generator + builder the phase kernel runs,
// over the first chunks of every distinct stream configuration and a spread of steady-state phases. Returns false when
// a program cannot be represented (the op then uses the table-driven kernels).
static bool rs_prog_dims(const skgpu_ctx *c, const skgpu_rs_item *items, uint32_t n, ChainProgDims *out) {
    uint32_t max_frames = 0, need_seg = 0, need_exp = 0;
    double last_t = -1.0;
    int32_t last_end = 0;
    uint32_t last_N = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t slot = items[i].slot;
        const double t = c->h_t[slot];
        const int32_t end_idx = c->h_end[slot];
        const uint32_t N = c->h_chunk[slot];
        if (t == last_t && end_idx == last_end && N == last_N) continue;
        last_t = t; last_end = end_idx; last_N = N;
        const uint32_t fcap = (uint32_t)((double)N / t + 10.0) + 8u;
        if (fcap > 60000u) return false;   // segment bounds are 16-bit
        ChainProgDims big{};
        big.nblk = (fcap + 31u) / 32u;
        big.map_bytes = skc_map_bytes(big.nblk);
        big.cap_seg = 254u;
        big.cap_exp = 8192u;
        std::vector<uint8_t> scratch(skc_prog_cap(big) + 16u);
        auto one_chunk = [&](double last_index, double *next_index) -> bool {
            SkcStream sb;
            uint32_t np, nr, ovf, ns = 0, ne = 0;
            double idx_end;
            sb.begin(scratch.data(), big, big.nblk * 32u, 8u, 0u, nullptr, 0u, 0u, 0u, N, 0u, t);
            const uint32_t n_out = sk_phase_stream(last_index, t, end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
            const uint32_t st = sb.finish_open(n_out, &ns, &ne);
            need_seg = std::max(need_seg, ns);
            need_exp = std::max(need_exp, ne);
            *next_index = idx_end - (double)N;
            return st == 0 && !ovf;
        };
        double li = -4.0, nli;
        for (int k = 0; k < 6; ++k) { if (!one_chunk(li, &nli)) return false; li = nli; }
        const double lo = -(9.0 + std::ceil(t));
        for (int k = 0; k < 64; ++k) if (!one_chunk(lo + t * (k + 0.37) / 64.0, &nli)) return false;
        max_frames = std::max(max_frames, fcap);
    }
    ChainProgDims d{};
    d.nblk = (std::max(max_frames, 32u) + 31u) / 32u;
    d.map_bytes = skc_map_bytes(d.nblk);
    d.cap_seg = (need_seg + 5u) & ~1u;
    d.cap_exp = (need_exp + 9u) & ~1u;
    if (d.cap_seg > 254u || skc_prog_cap(d) > SK_SIDE_STRIDE) return false;
    *out = d;
    return true;
}

static void rs_size_smem(const skgpu_ctx *c, Op &op) {
    // staged path: (16 + chunk) frames x channels x 4 B of dynamic shared memory, capped at 96 KB so that
    // at least two CTAs stay resident per SM; longer chunks take the direct-from-HBM path.
    const uint32_t chmax = op.rs_channels ? (uint32_t)op.rs_channels : c->cfg.max_channels;
    uint64_t need = (uint64_t)(op.max_unit + 16u) * chmax * 4u;
    if (need <= 96u * 1024u) { op.smem_frames = op.max_unit + 16u; op.smem_bytes = (uint32_t)((need + 15u) & ~15ull); }
    else { op.smem_frames = 0; op.smem_bytes = 0; }
    // the program-driven kernels need the staged path and mono / stereo streams
    if (op.rs_prog && (op.smem_frames == 0 || (op.rs_channels != 1 && op.rs_channels != 2))) op.rs_prog = false;
    if (op.rs_prog) op.smem_bytes += skc_prog_cap(op.rs_pd);
#ifdef SKGPU_TUNING_KNOBS
    if (std::getenv("SKGPU_RS_TABLE")) op.rs_prog = false;   // profiling knob: force the table-driven kernels
#endif
}

extern "C" skgpu_rc skgpu_plan_add_resample(skgpu_plan *p, const skgpu_rs_item *items, uint32_t n, uint64_t results_off, uint32_t *op_out) {
    if (!p || (!items && n)) return fail(SKGPU_ERR_INVALID, "null argument");
    if (p->finalized) return fail(SKGPU_ERR_STATE, "plan already finalized");
    CU(cudaSetDevice(p->ctx->device));
    uint32_t mx = 0;
    int ch = 0;
    bool fifo = false;
    skgpu_rc rc = validate_rs(p, items, n, &mx, &ch, &fifo);
    if (rc) return rc;
    rc = check_range(p, results_off, (uint64_t)std::max(n, 1u) * sizeof(skgpu_rs_result), "resample results");
    if (rc) return rc;
    if (results_off % 8) return fail(SKGPU_ERR_INVALID, "results_off must be 8-byte aligned");
    Op op;
    op.kind = OP_RESAMPLE;
    rc = op_alloc_tables(op, sizeof(skgpu_rs_item), std::max(n, 1u), 0, 0);
    if (rc) return rc;
    if (n) memcpy(op.h_tab, items, n * sizeof(skgpu_rs_item));
    op.n = n;
    op.max_unit = std::max(mx, 1u);
    op.rs_channels = ch;
    op.results_off = results_off;
    op.rs_sinc = n > 0 && (p->ctx->h_flags[items[0].slot] & SLOT_SINC) != 0;
    if (op.rs_sinc) {
        if (ch != 1 && ch != 2) return fail(SKGPU_ERR_INVALID, "sinc resample op: mono or stereo streams of one channel count");
        op.smem_bytes = (uint32_t)((((uint64_t)(op.max_unit + p->ctx->st.sinc_H) * (uint32_t)ch * 4u) + 15u) & ~15ull);
        if (op.smem_bytes > 200u * 1024u) return fail(SKGPU_ERR_INVALID, "sinc resample op: chunk too large for shared-memory staging");
        rs_sinc_dims(p->ctx, op, items, n, ch);
    } else {
        op.rs_prog = n > 0 && rs_prog_dims(p->ctx, items, n, &op.rs_pd);
        rs_size_smem(p->ctx, op);
    }
    if (op_out) *op_out = (uint32_t)p->ops.size();
    p->ops.push_back(op);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_update_resample(skgpu_plan *p, uint32_t opi, const skgpu_rs_item *items, uint32_t n) {
    if (!p || opi >= p->ops.size() || p->ops[opi].kind != OP_RESAMPLE) return fail(SKGPU_ERR_INVALID, "not a resample op");
    Op &op = p->ops[opi];
    if (n > op.cap) return fail(SKGPU_ERR_INVALID, "update exceeds capacity (%u > %u)", n, op.cap);
    uint32_t mx = 0;
    int ch = 0;
    bool fifo = false;
    skgpu_rc rc = validate_rs(p, items, n, &mx, &ch, &fifo);
    if (rc) return rc;
    if (n && ch != op.rs_channels && op.rs_channels != 0) return fail(SKGPU_ERR_INVALID, "update changes the op's channel specialisation (%d -> %d)", op.rs_channels, ch);
    if (n && op.rs_sinc != ((p->ctx->h_flags[items[0].slot] & SLOT_SINC) != 0)) return fail(SKGPU_ERR_INVALID, "update changes the op's interpolation mode");
    if (op.rs_sinc && mx > op.max_unit) return fail(SKGPU_ERR_INVALID, "update has a longer chunk (%u) than the op was sized for (%u)", mx, op.max_unit);
    if (op.rs_sinc && op.rs_sinc_tiled && n) {
        Op probe = op;
        rs_sinc_dims(p->ctx, probe, items, n, op.rs_channels);
        if (!probe.rs_sinc_tiled || probe.sinc_taps != op.sinc_taps) return fail(SKGPU_ERR_INVALID, "update changes the op's tap table (streams of one sinc op share one effective cutoff)");
        if (probe.sinc_dm.G < op.sinc_dm.G) op.sinc_dm.G = probe.sinc_dm.G;   // more work items per stream: fewer streams per pass (same shared memory)
    }
    if (op.smem_frames && mx + 16u > op.smem_frames) return fail(SKGPU_ERR_INVALID, "update has a longer chunk (%u) than the op was sized for (%u)", mx, op.smem_frames - 16u);
    if (op.rs_prog && n) {
        ChainProgDims d{};
        if (!rs_prog_dims(p->ctx, items, n, &d) || d.nblk > op.rs_pd.nblk || d.cap_seg > op.rs_pd.cap_seg || d.cap_exp > op.rs_pd.cap_exp)
            return fail(SKGPU_ERR_INVALID, "update adds a stream configuration whose frame program exceeds what the op was sized for");
    }
    CU(cudaStreamSynchronize(p->ctx->stream));
    if (n) memcpy(op.h_tab, items, n * sizeof(skgpu_rs_item));
    op.n = n;
    op.dirty = true;
    return SKGPU_OK;
}

// ---- mix

static skgpu_rc validate_mix(const skgpu_plan *p, const skgpu_mix_group *g, uint32_t ng, const skgpu_mix_input *in, uint32_t ni, uint32_t *max_out, bool *any_fifo) {
    const skgpu_ctx *c = p->ctx;
    uint32_t mx = 0;
    bool fifo = false;
    for (uint32_t i = 0; i < ni; ++i) {
        if (in[i].channels == 0) return fail(SKGPU_ERR_INVALID, "mix input %u: channels must be >= 1", i);
        if (in[i].gain_idx != SKGPU_NO_GAIN && in[i].gain_idx >= p->n_gains) return fail(SKGPU_ERR_INVALID, "mix input %u: gain_idx out of range", i);
        if (in[i].flags & SKGPU_MIX_IN_FIFO) {
            if (!c->cfg.fifo_frames) return fail(SKGPU_ERR_INVALID, "mix input %u: context has no device FIFO", i);
            if (in[i].slot >= c->cfg.max_streams || !c->used[in[i].slot]) return fail(SKGPU_ERR_INVALID, "mix input %u: slot %u is not open", i, in[i].slot);
            if (in[i].channels != c->h_ch[in[i].slot]) return fail(SKGPU_ERR_INVALID, "mix input %u: channel count differs from its stream slot", i);
            if (in[i].n_frames > c->cfg.fifo_frames) return fail(SKGPU_ERR_INVALID, "mix input %u: packet larger than the device FIFO", i);
            fifo = true;
        } else {
            skgpu_rc rc = check_range(p, in[i].in_off, (uint64_t)in[i].n_frames * in[i].channels * 4, "mix input");
            if (rc) return rc;
            if (in[i].in_off % 4) return fail(SKGPU_ERR_INVALID, "mix input %u: misaligned offset", i);
        }
    }
    for (uint32_t i = 0; i < ng; ++i) {
        if (g[i].out_channels == 0) return fail(SKGPU_ERR_INVALID, "mix group %u: out_channels must be >= 1", i);
        if ((uint64_t)g[i].first_input + g[i].n_inputs > ni) return fail(SKGPU_ERR_INVALID, "mix group %u: inputs [%u, +%u) outside the input table (%u)", i, g[i].first_input, g[i].n_inputs, ni);
        if (g[i].n_inputs > (uint32_t)MIX_MAX_INPUTS) return fail(SKGPU_ERR_INVALID, "mix group %u: more than %d inputs", i, MIX_MAX_INPUTS);
        if (g[i].gain_idx != SKGPU_NO_GAIN && g[i].gain_idx >= p->n_gains) return fail(SKGPU_ERR_INVALID, "mix group %u: gain_idx out of range", i);
        const uint64_t osz = (uint64_t)g[i].out_frames * g[i].out_channels;
        const uint32_t ob = (g[i].flags & SKGPU_MIX_OUT_S16) ? 2 : 4;
        skgpu_rc rc = check_range(p, g[i].out_off, osz * ob, "mix output");
        if (rc) return rc;
        if (g[i].out_off % ob) return fail(SKGPU_ERR_INVALID, "mix group %u: misaligned output", i);
        if (osz > 0xFFFFFFFFull) return fail(SKGPU_ERR_INVALID, "mix group %u: output too large", i);
        mx = std::max<uint32_t>(mx, (uint32_t)osz);
    }
    *max_out = mx;
    *any_fifo = fifo;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_add_mix(skgpu_plan *p, const skgpu_mix_group *groups, uint32_t ng, const skgpu_mix_input *inputs, uint32_t ni, uint32_t *op_out) {
    if (!p || (!groups && ng) || (!inputs && ni)) return fail(SKGPU_ERR_INVALID, "null argument");
    if (p->finalized) return fail(SKGPU_ERR_STATE, "plan already finalized");
    CU(cudaSetDevice(p->ctx->device));
    uint32_t mx = 0;
    bool fifo = false;
    skgpu_rc rc = validate_mix(p, groups, ng, inputs, ni, &mx, &fifo);
    if (rc) return rc;
    Op op;
    op.kind = OP_MIX;
    rc = op_alloc_tables(op, sizeof(skgpu_mix_group), std::max(ng, 1u), sizeof(skgpu_mix_input), std::max(ni, 1u));
    if (rc) return rc;
    if (ng) memcpy(op.h_tab, groups, ng * sizeof(skgpu_mix_group));
    if (ni) memcpy(op.h_tab2, inputs, ni * sizeof(skgpu_mix_input));
    op.n = ng;
    op.n2 = ni;
    op.max_unit = std::max(mx, 1u);
    op.tiles = (op.max_unit + MIX_TILE - 1) / MIX_TILE;
    {   // groups with few inputs: one CTA per group (the prologue costs more than the tile's loads); fixed at add time
        uint32_t max_k = 0;
        for (uint32_t i = 0; i < ng; ++i) max_k = std::max(max_k, groups[i].n_inputs);
        // ... and with several groups per SM a CTA takes a whole group whatever its size: the prologue (descriptor loads, summation
        // order) is paid once per group instead of once per tile (config #3, 1,024 groups x 64 inputs: 94.9 -> 89.0 us)
        op.mix_tpc = (ng >= 4096u && max_k <= 8u) ? std::min(op.tiles, 8u) : (ng >= 4u * (uint32_t)p->ctx->sm_count ? std::min(op.tiles, 8u) : 1u);
#ifdef SKGPU_TUNING_KNOBS
        if (const char *e = std::getenv("SKGPU_MIX_TPC")) op.mix_tpc = (uint32_t)std::max(1, std::atoi(e));
#endif
    }
    op.has_fifo_inputs = fifo;
    {   // presence table: every input present until skgpu_plan_set_present says otherwise
        rc = dyn_alloc(op.present, op.cap2);
        if (rc) return rc;
        std::vector<uint8_t> ones(op.cap2, 1);
        rc = dyn_write(op.present, ones.data(), op.cap2);
        if (rc) return rc;
    }
    if (op_out) *op_out = (uint32_t)p->ops.size();
    p->ops.push_back(op);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_update_mix(skgpu_plan *p, uint32_t opi, const skgpu_mix_group *groups, uint32_t ng, const skgpu_mix_input *inputs, uint32_t ni) {
    if (!p || opi >= p->ops.size() || p->ops[opi].kind != OP_MIX) return fail(SKGPU_ERR_INVALID, "not a mix op");
    Op &op = p->ops[opi];
    if (ng > op.cap || ni > op.cap2) return fail(SKGPU_ERR_INVALID, "update exceeds capacity");
    uint32_t mx = 0;
    bool fifo = false;
    skgpu_rc rc = validate_mix(p, groups, ng, inputs, ni, &mx, &fifo);
    if (rc) return rc;
    if (mx > op.tiles * (uint32_t)MIX_TILE) return fail(SKGPU_ERR_INVALID, "group output larger (%u) than the op was sized for (%u)", mx, op.tiles * MIX_TILE);
    if (fifo && !op.has_fifo_inputs && p->finalized) return fail(SKGPU_ERR_INVALID, "update introduces FIFO inputs into an op finalized without them");
    CU(cudaStreamSynchronize(p->ctx->stream));
    if (ng) memcpy(op.h_tab, groups, ng * sizeof(skgpu_mix_group));
    if (ni) memcpy(op.h_tab2, inputs, ni * sizeof(skgpu_mix_input));
    op.n = ng;
    op.n2 = ni;
    op.has_fifo_inputs = op.has_fifo_inputs || fifo;
    op.dirty = true;
    return SKGPU_OK;
}

// ---- io / dynamic tables

extern "C" skgpu_rc skgpu_plan_set_io(skgpu_plan *p, uint64_t h2d_off, uint64_t h2d_bytes, uint64_t d2h_off, uint64_t d2h_bytes) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    skgpu_rc rc = check_range(p, h2d_off, h2d_bytes, "h2d range");
    if (rc) return rc;
    rc = check_range(p, d2h_off, d2h_bytes, "d2h range");
    if (rc) return rc;
    p->h2d_off = h2d_off; p->h2d_bytes = h2d_bytes; p->d2h_off = d2h_off; p->d2h_bytes = d2h_bytes;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_set_gains(skgpu_plan *p, const float *gains, uint32_t n) {
    if (!p || (!gains && n)) return fail(SKGPU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(p->ctx->device));
    if (!p->gains.valid) {
        if (p->finalized) return fail(SKGPU_ERR_STATE, "gain table must be created before finalize");
        skgpu_rc rc = dyn_alloc(p->gains, std::max<size_t>(n, 1) * sizeof(float));
        if (rc) return rc;
        p->n_gains = n;
    } else if (n != p->n_gains) {
        return fail(SKGPU_ERR_INVALID, "gain table size is fixed at %u (got %u)", p->n_gains, n);
    }
    return dyn_write(p->gains, gains, n * sizeof(float));
}

extern "C" skgpu_rc skgpu_plan_set_present(skgpu_plan *p, uint32_t opi, const uint8_t *present, uint32_t n) {
    if (!p || opi >= p->ops.size() || (p->ops[opi].kind != OP_MIX && p->ops[opi].kind != OP_CHAIN)) return fail(SKGPU_ERR_INVALID, "not a mix or chain op");
    Op &op = p->ops[opi];
    CU(cudaSetDevice(p->ctx->device));
    if (n > op.cap2) return fail(SKGPU_ERR_INVALID, "presence table larger than the input table capacity");
    std::vector<uint8_t> full(op.cap2, 1);
    if (n) memcpy(full.data(), present, n);
    return dyn_write(op.present, full.data(), op.cap2);
}

extern "C" skgpu_rc skgpu_plan_set_banks(skgpu_plan *p, uint64_t bank_stride) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    if (p->finalized) return fail(SKGPU_ERR_STATE, "plan already finalized");
    if (bank_stride % 16) return fail(SKGPU_ERR_INVALID, "bank_stride must be a multiple of 16 bytes");
    if (bank_stride < p->h2d_bytes) return fail(SKGPU_ERR_INVALID, "bank_stride smaller than the H2D range (call skgpu_plan_set_io first)");
    skgpu_rc rc = check_range(p, p->h2d_off + bank_stride, p->h2d_bytes, "second input bank");
    if (rc) return rc;
    p->bank_stride = bank_stride;
    return SKGPU_OK;
}
extern "C" uint64_t skgpu_plan_tick_count(const skgpu_plan *p) { return p ? p->tick : 0; }

// ---- chain

// Upper bounds of the frame-program sizes (chain_prog.h) a stream configuration can produce: run the host-compiled
// generator + builder over the first chunks of a stream and over a spread of steady-state phases. The kernel
// re-checks at run time (status bit1).
static bool prog_bounds(double t, int32_t end_idx, uint32_t N, uint32_t F, uint32_t *need_seg, uint32_t *need_exp) {
    ChainProgDims big{};
    big.nblk = (F + 31u) / 32u;
    big.map_bytes = skc_map_bytes(big.nblk);
    big.cap_seg = 254u;
    big.cap_exp = 8192u;
    std::vector<uint8_t> scratch(skc_prog_cap(big));
    uint32_t ms = 0, me = 0;
    bool overflow = false;
    auto one_chunk = [&](double last_index, uint32_t carry, bool pending, double *next_index, uint32_t *next_carry) {
        // the same sequence k_phase_chain runs (k_chain.cuh)
        SkcStream sb;
        uint32_t np, nr, ovf, ns = 0, ne = 0;
        double idx_end;
        uint32_t kd = pending ? F - std::min(carry, F) : 0u;
        std::vector<ChainExp2> tail2(F / 2u + 2u);   // 16-byte aligned: the builder stores entries in pairs
        ChainExp *tail_p = reinterpret_cast<ChainExp *>(tail2.data());
        sb.begin(scratch.data(), big, F, 8u, kd, tail_p, 0u, kd, F, N, 32u, t);
        uint32_t n_cur = sk_phase_stream(last_index, t, end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
        if (n_cur < kd) {
            sb.begin(scratch.data(), big, F, 8u, 0u, tail_p, 0u, 0u, 0u, N, 32u, t);
            n_cur = sk_phase_stream(last_index, t, end_idx, SKC_TAB_PREFIX, 255u, sb, &np, &nr, &ovf, &idx_end);
        }
        overflow |= sb.finish(n_cur, &ns, &ne) != 0;
        const uint32_t avail = carry + n_cur;
        const uint32_t nc = avail >= F ? avail - F : avail;
        ms = std::max(ms, ns);
        me = std::max(me, ne + (F - std::min(nc, F)));
        *next_index = idx_end - (double)N;
        *next_carry = nc;
    };
    double li = -4.0, nli;
    uint32_t carry = 0, ncarry, steady = 0;
    for (int c = 0; c < 8; ++c) { one_chunk(li, carry, c > 0, &nli, &ncarry); li = nli; carry = ncarry; steady = carry; }
    const double lo = -(9.0 + std::ceil(t));
    for (int i = 0; i < 64; ++i) one_chunk(lo + t * (i + 0.37) / 64.0, steady, true, &nli, &ncarry);
    *need_seg = ms; *need_exp = me;
    return !overflow;
}

static skgpu_rc validate_chain(const skgpu_plan *p, const skgpu_chain_group *g, uint32_t ng, const skgpu_chain_input *in, uint32_t ni,
                               uint32_t F, uint32_t *max_k, uint32_t *max_buf_floats, int *oc_out, uint32_t *cap_seg, uint32_t *cap_exp, int *kind_out) {
    uint32_t need_seg = 0, need_exp = 0;
    double last_t = -1.0;
    int32_t last_end = 0;
    uint32_t last_N = 0;
    const skgpu_ctx *c = p->ctx;
    uint32_t mk = 0, mb = 0;
    int oc = -1;
    for (uint32_t i = 0; i < ni; ++i) {
        const uint32_t slot = in[i].slot;
        if (slot >= c->cfg.max_streams || !c->used[slot]) return fail(SKGPU_ERR_INVALID, "chain input %u: slot %u is not open", i, slot);
        const uint32_t N = c->h_chunk[slot], C = c->h_ch[slot];
        if (C != 1 && C != 2) return fail(SKGPU_ERR_INVALID, "chain input %u: %u channels (the fused chain handles mono and stereo)", i, C);
        if (N < 16) return fail(SKGPU_ERR_INVALID, "chain input %u: chunk_frames %u < 16", i, N);
        const bool bypass = (c->h_flags[slot] & SLOT_BYPASS) != 0, s16 = (c->h_flags[slot] & SLOT_S16) != 0;
        if (c->h_flags[slot] & SLOT_SINC) return fail(SKGPU_ERR_INVALID, "chain input %u: sinc streams belong to a resample op (the fused chain reproduces the reference's linear interpolation)", i);
        // nominal frames per chunk must equal the packet size: a packet then never spans more than two chunks
        const double nominal = (double)N / c->h_t[slot];
        if (bypass ? N != F : std::fabs(nominal - (double)F) > 0.5)
            return fail(SKGPU_ERR_INVALID, "chain input %u: chunk of %u frames yields %.2f output frames, not output_frame_size %u (use the unfused ops)", i, N, nominal, F);
        if (in[i].in_off % 16) return fail(SKGPU_ERR_INVALID, "chain input %u: in_off must be 16-byte aligned", i);
        // s16 chunks are copied in whole 16-byte units: up to 16 bytes past the chunk's end are read (and ignored)
        skgpu_rc rc = check_range(p, in[i].in_off + p->bank_stride, s16 ? (((uint64_t)N * C * 2 + 15u) & ~15ull) : (uint64_t)N * C * 4, "chain input (bank 1)");
        if (rc) return rc;
        if (in[i].gain_idx != SKGPU_NO_GAIN && in[i].gain_idx >= p->n_gains) return fail(SKGPU_ERR_INVALID, "chain input %u: gain_idx out of range", i);
        mb = std::max(mb, (N + (uint32_t)CH_HEAD) * C * 4u + 32u);   // bytes: previous chunk + staged head of the current one (+ s16 copy slack)
        if (bypass) continue;                                        // no phase, no program
        if (c->h_t[slot] != last_t || c->h_end[slot] != last_end || N != last_N) {   // streams of one op usually share a handful of configurations
            uint32_t a = 0, b = 0;
            if (!prog_bounds(c->h_t[slot], c->h_end[slot], N, F, &a, &b))
                return fail(SKGPU_ERR_INVALID, "chain input %u: the phase table of ratio %g does not fit the fused kernel (use the unfused ops)", i, 1.0 / c->h_t[slot]);
            need_seg = std::max(need_seg, a); need_exp = std::max(need_exp, b);
            last_t = c->h_t[slot]; last_end = c->h_end[slot]; last_N = N;
        }
    }
    for (uint32_t i = 0; i < ng; ++i) {
        if (g[i].out_channels != 1 && g[i].out_channels != 2) return fail(SKGPU_ERR_INVALID, "chain group %u: out_channels must be 1 or 2", i);
        if (oc == -1) oc = g[i].out_channels;
        else if (oc != g[i].out_channels) return fail(SKGPU_ERR_INVALID, "chain op: all groups must share out_channels (split into two ops)");
        if ((uint64_t)g[i].first_input + g[i].n_inputs > ni) return fail(SKGPU_ERR_INVALID, "chain group %u: inputs outside the input table", i);
        if (g[i].n_inputs > (uint32_t)CH_MAX_INPUTS) return fail(SKGPU_ERR_INVALID, "chain group %u: more than %d inputs", i, CH_MAX_INPUTS);
        if (g[i].gain_idx != SKGPU_NO_GAIN && g[i].gain_idx >= p->n_gains) return fail(SKGPU_ERR_INVALID, "chain group %u: gain_idx out of range", i);
        const uint32_t ob = (g[i].flags & SKGPU_MIX_OUT_S16) ? 2 : 4;
        skgpu_rc rc = check_range(p, g[i].out_off, (uint64_t)F * g[i].out_channels * ob, "chain output");
        if (rc) return rc;
        if (g[i].out_off % 16) return fail(SKGPU_ERR_INVALID, "chain group %u: out_off must be 16-byte aligned", i);
        for (uint32_t j = 0; j < g[i].n_inputs; ++j) {
            const uint32_t C = c->h_ch[in[g[i].first_input + j].slot];
            if (C > g[i].out_channels && !(C == 2 && g[i].out_channels == 1)) return fail(SKGPU_ERR_INVALID, "chain group %u: unsupported channel mapping", i);
        }
        mk = std::max(mk, g[i].n_inputs);
    }
    *max_k = mk;
    *max_buf_floats = (mb + 15u) & ~15u;
    *oc_out = oc < 0 ? 2 : oc;
    {
        bool plain = ni > 0, bypass = ni > 0, f32 = true, s16 = true, same_shape = ni > 0;   // one kind only: streams with the output's channel count, all resampled or all rate-equal, all f32 or all s16
        for (uint32_t i = 0; i < ni; ++i) {
            const bool shape = (int)c->h_ch[in[i].slot] == *oc_out;
            same_shape = same_shape && shape;
            plain = plain && shape && !(c->h_flags[in[i].slot] & SLOT_BYPASS);
            bypass = bypass && shape && (c->h_flags[in[i].slot] & SLOT_BYPASS);
            f32 = f32 && !(c->h_flags[in[i].slot] & SLOT_S16);
            s16 = s16 && (c->h_flags[in[i].slot] & SLOT_S16);
        }
        *kind_out = (plain && f32) ? CHAIN_PLAIN : (bypass && f32) ? CHAIN_BYPASS : (plain && s16) ? CHAIN_PLAIN_S16 : (bypass && s16) ? CHAIN_BYPASS_S16 : (same_shape && f32) ? CHAIN_F32 : CHAIN_ANY;
    }
    // margin for phases the sampling did not hit; even counts keep every staged array a multiple of 16 bytes
    *cap_seg = (need_seg + 5u) & ~1u;
    *cap_exp = (need_exp + 9u) & ~1u;
    {
        ChainProgDims d{};
        d.nblk = (F + 31u) / 32u;
        d.map_bytes = skc_map_bytes(d.nblk);
        d.cap_seg = *cap_seg;
        d.cap_exp = *cap_exp;
        if (*cap_seg > 254u || skc_prog_cap(d) > CH_PROG_MAX)
            return fail(SKGPU_ERR_INVALID, "chain op: frame program of %u segments + %u explicit frames does not fit a stream's side record (use the unfused ops)", *cap_seg, *cap_exp);
    }
    return SKGPU_OK;
}

static void chain_size_smem(Op &op, uint32_t max_k, uint32_t chunk_cap, uint32_t cap_seg, uint32_t cap_exp) {
    // nstages pipeline stages, each staging up to kb inputs: [frame program | history 128 B | previous chunk + head].
    // 4 CTAs per SM is the register limit (56 regs x 288 threads); pick the deepest ring that keeps 4 CTAs in 227 KB.
    ChainDims dm{};
    dm.chunk_cap = chunk_cap;
    dm.prog.nblk = (op.chain_F + 31u) / 32u;
    dm.prog.map_bytes = skc_map_bytes(dm.prog.nblk);
    dm.prog.cap_seg = cap_seg;
    dm.prog.cap_exp = cap_exp;
    dm.max_k = std::max(max_k, 1u);
    const uint64_t in_bytes = chain_in_bytes(dm);
    uint32_t kb = std::max(1u, std::min<uint32_t>(max_k, CH_MAX_KB));
    while (kb > 1 && 2ull * kb * in_bytes > 100u * 1024u) --kb;
#ifdef SKGPU_TUNING_KNOBS
    if (const char *e = std::getenv("SKGPU_CHAIN_KB")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= CH_MAX_KB) kb = (uint32_t)v;
    }
#endif
    const uint64_t stage = (uint64_t)kb * in_bytes;
    const uint64_t scratch = (uint64_t)dm.max_k * sizeof(ChainRec);
    const uint64_t static_smem = sizeof(ChainStage) * CH_MAX_STAGES + 5u * 1024u, sm_smem = 227u * 1024u, cta_overhead = 1024u;   // stage headers + prefetch slots
    uint32_t best_ns = 2;
    for (uint32_t ns = 2; ns <= 2u; ++ns) {   // measured on B200: deeper rings do not help (the kernel is issue bound), keep 2
        const uint64_t per_cta = stage * ns + scratch + static_smem + cta_overhead;
        if (per_cta * 4u <= sm_smem) best_ns = ns;
    }
#ifdef SKGPU_TUNING_KNOBS   // never in the product build: environment variables must not change what the library computes or how
    if (const char *e = std::getenv("SKGPU_CHAIN_STAGES")) {
        const int v = std::atoi(e);
        if (v >= 2 && v <= CH_MAX_STAGES) best_ns = (uint32_t)v;
    }
#endif
    dm.kb = kb;
    dm.one = 1.0f;
    dm.nstages = best_ns;
    dm.reserved = 0;
    op.chain_kb = kb;
    op.chain_buf_floats = chunk_cap;
    op.chain_dm = dm;
    op.smem_bytes = (uint32_t)(stage * best_ns + scratch);
}

extern "C" skgpu_rc skgpu_plan_add_chain(skgpu_plan *p, const skgpu_chain_group *groups, uint32_t ng, const skgpu_chain_input *inputs, uint32_t ni,
                                         uint32_t output_frame_size, uint64_t results_off, uint32_t *op_out) {
    return skgpu_plan_add_chain_cap(p, groups, ng, inputs, ni, ng, ni, 0, output_frame_size, results_off, op_out);
}

extern "C" skgpu_rc skgpu_plan_add_chain_cap(skgpu_plan *p, const skgpu_chain_group *groups, uint32_t ng, const skgpu_chain_input *inputs, uint32_t ni,
                                             uint32_t cap_groups, uint32_t cap_inputs, uint32_t max_inputs_per_group, uint32_t output_frame_size,
                                             uint64_t results_off, uint32_t *op_out) {
    if (cap_groups < ng || cap_inputs < ni) return fail(SKGPU_ERR_INVALID, "table capacities smaller than the initial tables");
    if (max_inputs_per_group > (uint32_t)CH_MAX_INPUTS) return fail(SKGPU_ERR_INVALID, "more than %d inputs per group", CH_MAX_INPUTS);
    if (!p || (!groups && ng) || (!inputs && ni)) return fail(SKGPU_ERR_INVALID, "null argument");
    if (p->finalized) return fail(SKGPU_ERR_STATE, "plan already finalized");
    if (!p->bank_stride) return fail(SKGPU_ERR_STATE, "the fused chain needs a double-banked input range: call skgpu_plan_set_banks first");
    static const uint32_t valid[] = {120, 240, 480, 960, 1920, 2880};  // resampler.rs:95-102
    bool okF = false;
    for (uint32_t v : valid) okF |= (v == output_frame_size);
    if (!okF) return fail(SKGPU_ERR_INVALID, "output_frame_size must be a valid Opus frame size: [120, 240, 480, 960, 1920, 2880]");
    CU(cudaSetDevice(p->ctx->device));
    uint32_t mk = 0, mb = 0, cnp = 0, cnr = 0;
    int oc = 2;
    int kind = CHAIN_ANY;
    skgpu_rc rc = validate_chain(p, groups, ng, inputs, ni, output_frame_size, &mk, &mb, &oc, &cnp, &cnr, &kind);
    if (rc) return rc;
    mk = std::max(mk, max_inputs_per_group);
    rc = check_range(p, results_off, (uint64_t)std::max(cap_inputs, 1u) * sizeof(skgpu_chain_result), "chain results");
    if (rc) return rc;
    if (results_off % 8) return fail(SKGPU_ERR_INVALID, "results_off must be 8-byte aligned");
    Op op;
    op.kind = OP_CHAIN;
    rc = op_alloc_tables(op, sizeof(skgpu_chain_group), std::max(cap_groups, 1u), sizeof(skgpu_chain_input), std::max(cap_inputs, 1u));
    if (rc) return rc;
    if (ng) memcpy(op.h_tab, groups, ng * sizeof(skgpu_chain_group));
    if (ni) memcpy(op.h_tab2, inputs, ni * sizeof(skgpu_chain_input));
    op.n = ng;
    op.n2 = ni;
    op.chain_F = output_frame_size;
    op.chain_oc = oc;
    op.chain_kind = kind;
    op.chain_iters = (int)((output_frame_size + 1023u) / 1024u);
    op.results_off = results_off;
    chain_size_smem(op, mk, std::max(mb, 64u), cnp, cnr);
    CU(cudaMalloc((void **)&op.d_rec, (size_t)op.cap2 * sizeof(ChainRec)));
    CU(cudaMemset(op.d_rec, 0, (size_t)op.cap2 * sizeof(ChainRec)));
    if (op.smem_bytes > 200u * 1024u) return fail(SKGPU_ERR_INVALID, "chain op: chunk too large for shared-memory staging (%u bytes)", op.smem_bytes);
    {
        rc = dyn_alloc(op.present, op.cap2);
        if (rc) return rc;
        std::vector<uint8_t> ones(op.cap2, 1);
        rc = dyn_write(op.present, ones.data(), op.cap2);
        if (rc) return rc;
    }
    if (op_out) *op_out = (uint32_t)p->ops.size();
    p->ops.push_back(op);
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_update_chain(skgpu_plan *p, uint32_t opi, const skgpu_chain_group *groups, uint32_t ng, const skgpu_chain_input *inputs, uint32_t ni) {
    if (!p || opi >= p->ops.size() || p->ops[opi].kind != OP_CHAIN) return fail(SKGPU_ERR_INVALID, "not a chain op");
    Op &op = p->ops[opi];
    if (ng > op.cap || ni > op.cap2) return fail(SKGPU_ERR_INVALID, "update exceeds capacity");
    uint32_t mk = 0, mb = 0, cnp = 0, cnr = 0;
    int oc = 2;
    int kind = CHAIN_ANY;
    skgpu_rc rc = validate_chain(p, groups, ng, inputs, ni, op.chain_F, &mk, &mb, &oc, &cnp, &cnr, &kind);
    if (rc) return rc;
    if (ng && oc != op.chain_oc) return fail(SKGPU_ERR_INVALID, "update changes the op's output channel count");
    const bool kind_ok = op.chain_kind == CHAIN_ANY || !ni || kind == op.chain_kind ||
                         (op.chain_kind == CHAIN_F32 && (kind == CHAIN_PLAIN || kind == CHAIN_BYPASS));   // the mixed-f32 kernel runs either
    if (!kind_ok) return fail(SKGPU_ERR_INVALID, "update adds an input of another kind (channel count differing from the output's, resampled vs rate-equal bypass, s16) to an op created from inputs of one kind only (include such a stream in the initial tables)");
    if (mb > op.chain_buf_floats) return fail(SKGPU_ERR_INVALID, "update has a longer chunk than the op was sized for");
    if (mk > op.chain_dm.max_k) return fail(SKGPU_ERR_INVALID, "update has a session with more inputs (%u) than the op was sized for (%u)", mk, op.chain_dm.max_k);
    if (cnp > op.chain_dm.prog.cap_seg || cnr > op.chain_dm.prog.cap_exp) return fail(SKGPU_ERR_INVALID, "update adds a resampling ratio whose phase tables exceed the op's staging capacity");
    CU(cudaStreamSynchronize(p->ctx->stream));
    if (ng) memcpy(op.h_tab, groups, ng * sizeof(skgpu_chain_group));
    if (ni) memcpy(op.h_tab2, inputs, ni * sizeof(skgpu_chain_input));
    op.n = ng;
    op.n2 = ni;
    op.dirty = true;
    op.slices.clear();   // the cut points referred to the old tables: set them again (skgpu_plan_set_slices / _auto_slices)
    return SKGPU_OK;
}

// ---- slices

static uint64_t chunk_bytes_of(const skgpu_ctx *c, uint32_t slot) {
    return (uint64_t)c->h_chunk[slot] * c->h_ch[slot] * ((c->h_flags[slot] & SLOT_S16) ? 2u : 4u);
}

extern "C" skgpu_rc skgpu_plan_set_slices(skgpu_plan *p, uint32_t opi, const skgpu_slice *sl, uint32_t n) {
    if (!p || opi >= p->ops.size() || p->ops[opi].kind != OP_CHAIN) return fail(SKGPU_ERR_INVALID, "not a chain op");
    if (p->ops.size() != 1) return fail(SKGPU_ERR_INVALID, "sliced ticks need a plan whose only op is the chain op");
    Op &op = p->ops[opi];
    if (n == 0) { op.slices.clear(); return SKGPU_OK; }
    if (!sl || n > 4096) return fail(SKGPU_ERR_INVALID, "invalid slice table");
    const skgpu_ctx *c = p->ctx;
    const skgpu_chain_group *g = (const skgpu_chain_group *)op.h_tab;
    const skgpu_chain_input *in = (const skgpu_chain_input *)op.h_tab2;
    std::vector<uint32_t> last_reader(n);
    uint32_t g0 = 0, i0 = 0;
    uint64_t up0 = 0;
    for (uint32_t k = 0; k < n; ++k) {
        if (sl[k].group_end < g0 || sl[k].group_end > op.n || sl[k].input_end < i0 || sl[k].input_end > op.n2) return fail(SKGPU_ERR_INVALID, "slice %u: table ranges must be consecutive and inside the tables", k);
        if (sl[k].h2d_end < up0 || sl[k].h2d_end > p->h2d_bytes) return fail(SKGPU_ERR_INVALID, "slice %u: h2d_end must be non-decreasing and inside the H2D range", k);
        for (int r = 0; r < 2; ++r)
            if (sl[k].d2h_off[r] > p->d2h_bytes || sl[k].d2h_bytes[r] > p->d2h_bytes - sl[k].d2h_off[r]) return fail(SKGPU_ERR_INVALID, "slice %u: read-back range outside the D2H range", k);
        for (uint32_t q = g0; q < sl[k].group_end; ++q)
            if (g[q].first_input < i0 || (uint64_t)g[q].first_input + g[q].n_inputs > sl[k].input_end) return fail(SKGPU_ERR_INVALID, "slice %u: group %u owns inputs outside the slice", k, q);
        last_reader[k] = k;
        g0 = sl[k].group_end; i0 = sl[k].input_end; up0 = sl[k].h2d_end;
    }
    if (g0 != op.n || i0 != op.n2) return fail(SKGPU_ERR_INVALID, "slices do not cover the tables (%u of %u groups, %u of %u inputs)", g0, op.n, i0, op.n2);
    if (sl[n - 1].h2d_end != p->h2d_bytes) return fail(SKGPU_ERR_INVALID, "the last slice must end the H2D range");
    // every input must be uploaded by its own slice's h2d_end; remember which slice reads each upload piece last (the next
    // tick's upload into the other bank overwrites what THIS tick's kernels read as the previous chunk)
    i0 = 0;
    for (uint32_t k = 0; k < n; ++k) {
        for (uint32_t q = i0; q < sl[k].input_end; ++q) {
            if (in[q].in_off < p->h2d_off) return fail(SKGPU_ERR_INVALID, "input %u lies below the H2D range", q);
            const uint64_t b0 = in[q].in_off - p->h2d_off;
            const uint64_t b1 = b0 + chunk_bytes_of(c, in[q].slot);
            if (b1 > sl[k].h2d_end) return fail(SKGPU_ERR_INVALID, "slice %u: input %u ends at byte %llu of the H2D range, beyond the slice's h2d_end %llu (sort the tables by input offset)", k, q, (unsigned long long)b1, (unsigned long long)sl[k].h2d_end);
            for (uint32_t u = 0; u <= k; ++u) {   // upload pieces [h2d_end[u-1], h2d_end[u]) the input overlaps
                const uint64_t lo = u ? sl[u - 1].h2d_end : 0, hi = sl[u].h2d_end;
                if (b0 < hi && b1 > lo) last_reader[u] = std::max(last_reader[u], k);
                if (hi >= b1) break;
            }
        }
        i0 = sl[k].input_end;
    }
    CU(cudaSetDevice(p->ctx->device));
    if (n > op.hdr_sl_cap) {
        CU(cudaStreamSynchronize(p->ctx->stream));
        CU(cudaStreamSynchronize(p->ctx->stream_k));
        if (op.d_hdr_sl) cudaFree(op.d_hdr_sl);
        if (op.h_hdr_sl) cudaFreeHost(op.h_hdr_sl);
        op.hdr_sl_cap = std::max(n, 64u);
        CU(cudaMalloc((void **)&op.d_hdr_sl, op.hdr_sl_cap * sizeof(OpHeader)));
        CU(cudaHostAlloc((void **)&op.h_hdr_sl, op.hdr_sl_cap * sizeof(OpHeader), cudaHostAllocDefault));
    }
    op.slices.assign(sl, sl + n);
    op.last_reader = last_reader;
    op.slices_dirty = true;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_auto_slices(skgpu_plan *p, uint32_t opi, uint32_t n) {
    if (!p || opi >= p->ops.size() || p->ops[opi].kind != OP_CHAIN) return fail(SKGPU_ERR_INVALID, "not a chain op");
    Op &op = p->ops[opi];
    if (n == 0 || op.n == 0) return skgpu_plan_set_slices(p, opi, nullptr, 0);
    n = std::min(n, op.n);
    const skgpu_ctx *c = p->ctx;
    const skgpu_chain_group *g = (const skgpu_chain_group *)op.h_tab;
    const skgpu_chain_input *in = (const skgpu_chain_input *)op.h_tab2;
    const uint64_t out_bytes = (uint64_t)op.chain_F * (uint32_t)op.chain_oc;
    std::vector<skgpu_slice> sl(n);
    uint32_t g0 = 0, i0 = 0;
    uint64_t up = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const uint32_t g1 = (uint32_t)((uint64_t)op.n * (k + 1) / n);
        uint32_t i1 = i0;
        uint64_t lo = ~0ull, hi = 0;
        for (uint32_t q = g0; q < g1; ++q) {
            i1 = std::max(i1, g[q].first_input + g[q].n_inputs);
            const uint64_t ob = out_bytes * ((g[q].flags & SKGPU_MIX_OUT_S16) ? 2u : 4u);
            lo = std::min(lo, g[q].out_off); hi = std::max(hi, g[q].out_off + ob);
        }
        for (uint32_t q = i0; q < i1; ++q) {
            if (in[q].in_off < p->h2d_off) return fail(SKGPU_ERR_INVALID, "input %u lies below the H2D range", q);
            up = std::max(up, in[q].in_off - p->h2d_off + chunk_bytes_of(c, in[q].slot));
        }
        if (k + 1 == n) up = p->h2d_bytes;
        sl[k].group_end = g1; sl[k].input_end = i1; sl[k].h2d_end = std::min<uint64_t>(up, p->h2d_bytes);
        // read-back: the slice's result rows, then its output rows (skipped when they lie outside the D2H range)
        sl[k].d2h_off[0] = sl[k].d2h_bytes[0] = sl[k].d2h_off[1] = sl[k].d2h_bytes[1] = 0;
        const uint64_t r0 = op.results_off + (uint64_t)i0 * sizeof(skgpu_chain_result), r1 = op.results_off + (uint64_t)i1 * sizeof(skgpu_chain_result);
        if (r0 >= p->d2h_off && r1 <= p->d2h_off + p->d2h_bytes && r1 > r0) { sl[k].d2h_off[0] = r0 - p->d2h_off; sl[k].d2h_bytes[0] = r1 - r0; }
        if (lo != ~0ull && lo >= p->d2h_off && hi <= p->d2h_off + p->d2h_bytes) { sl[k].d2h_off[1] = lo - p->d2h_off; sl[k].d2h_bytes[1] = hi - lo; }
        g0 = g1; i0 = i1;
    }
    return skgpu_plan_set_slices(p, opi, sl.data(), n);
}

// ---- launch

static skgpu_rc upload_dirty(skgpu_plan *p) {
    cudaStream_t s = p->ctx->stream;
    for (auto &op : p->ops) {
        if (op.dirty) {
            const size_t esz = op.kind == OP_CONVERT ? sizeof(skgpu_seg) : op.kind == OP_RESAMPLE ? sizeof(skgpu_rs_item)
                               : op.kind == OP_MIX ? sizeof(skgpu_mix_group) : sizeof(skgpu_chain_group);
            const size_t esz2 = op.kind == OP_MIX ? sizeof(skgpu_mix_input) : sizeof(skgpu_chain_input);
            if (op.n) CU(cudaMemcpyAsync(op.d_tab, op.h_tab, esz * op.n, cudaMemcpyHostToDevice, s));
            if ((op.kind == OP_MIX || op.kind == OP_CHAIN) && op.n2) CU(cudaMemcpyAsync(op.d_tab2, op.h_tab2, esz2 * op.n2, cudaMemcpyHostToDevice, s));
            op.h_hdr->count = op.n;
            op.h_hdr->count2 = op.n2;
            CU(cudaMemcpyAsync(op.d_hdr, op.h_hdr, sizeof(OpHeader), cudaMemcpyHostToDevice, s));
            op.dirty = false;
        }
        if (op.kind == OP_MIX || op.kind == OP_CHAIN) { skgpu_rc rc = dyn_upload(op.present, s); if (rc) return rc; }
    }
    return dyn_upload(p->gains, s);
}

static skgpu_rc op_event(Op &op, int sub, bool second, cudaStream_t s) {
    if (!second) {
        if (op.ev_used[sub] >= op.ev[sub].size()) {
            if (op.ev[sub].size() >= 8192) return SKGPU_OK;  // stop sampling, keep running
            cudaEvent_t a, b;
            CU(cudaEventCreate(&a));
            CU(cudaEventCreate(&b));
            op.ev[sub].push_back({a, b});
        }
        CU(cudaEventRecord(op.ev[sub][op.ev_used[sub]].first, s));
    } else {
        if (op.ev_used[sub] >= op.ev[sub].size()) return SKGPU_OK;
        CU(cudaEventRecord(op.ev[sub][op.ev_used[sub]].second, s));
        op.ev_used[sub]++;
    }
    return SKGPU_OK;
}

typedef void (*chain_kernel_t)(const OpHeader *, const skgpu_chain_group *, const ChainRec *, const float *, SlotTables, uint8_t *, uint32_t, ChainDims, uint32_t *);
static chain_kernel_t chain_kernel(int oc, int iters, int kind) {
#define SK_CHAIN_PICK(K)                                                                                                  \
    if (oc == 2) return iters == 1 ? k_chain<2, 1, K> : iters == 2 ? k_chain<2, 2, K> : k_chain<2, 3, K>;                \
    return iters == 1 ? k_chain<1, 1, K> : iters == 2 ? k_chain<1, 2, K> : k_chain<1, 3, K>;
    if (kind == CHAIN_PLAIN) { SK_CHAIN_PICK(CHAIN_PLAIN) }
    if (kind == CHAIN_BYPASS) { SK_CHAIN_PICK(CHAIN_BYPASS) }
    if (kind == CHAIN_PLAIN_S16) { SK_CHAIN_PICK(CHAIN_PLAIN_S16) }
    if (kind == CHAIN_BYPASS_S16) { SK_CHAIN_PICK(CHAIN_BYPASS_S16) }
    if (kind == CHAIN_F32) { SK_CHAIN_PICK(CHAIN_F32) }
    SK_CHAIN_PICK(CHAIN_ANY)
#undef SK_CHAIN_PICK
}

// the two kernels of the chain op over the table range a launch header describes (the whole tables, or one slice)
static skgpu_rc launch_chain(skgpu_plan *p, Op &op, const OpHeader *d_hdr, uint32_t n_groups, uint32_t n_inputs, cudaStream_t s, bool time_ops) {
    skgpu_ctx *c = p->ctx;
    const float *gains = (const float *)p->gains.dev;
    const uint8_t *present = op.present.valid ? (const uint8_t *)op.present.dev : nullptr;
    const skgpu_chain_input *cin = (const skgpu_chain_input *)op.d_tab2;
    if (time_ops) { skgpu_rc rc = op_event(op, 0, false, s); if (rc) return rc; }
    k_phase_chain<<<(std::max(n_inputs, 1u) + PHASE_CHAIN_THREADS - 1) / PHASE_CHAIN_THREADS, PHASE_CHAIN_THREADS, 0, s>>>(
        d_hdr, cin, present, gains, c->st, p->arena, p->d_tick, p->bank_stride, op.chain_F, op.results_off, op.chain_dm, op.d_rec);
    CU(cudaGetLastError());
    if (time_ops) { skgpu_rc rc = op_event(op, 0, true, s); if (rc) return rc; rc = op_event(op, 1, false, s); if (rc) return rc; }
    const uint32_t grid = std::max(1u, std::min<uint32_t>(n_groups, op.chain_grid));
    auto kfn = chain_kernel(op.chain_oc, op.chain_iters, op.chain_kind);
    kfn<<<grid, CH_THREADS, op.smem_bytes, s>>>(d_hdr, (const skgpu_chain_group *)op.d_tab, op.d_rec, gains, c->st, p->arena, op.chain_F, op.chain_dm,
                                                 p->ops.size() == 1 ? p->d_tick : nullptr);   // a chain-only plan: the kernel advances the bank parity itself
    CU(cudaGetLastError());
    if (time_ops) { skgpu_rc rc = op_event(op, 1, true, s); if (rc) return rc; }
    return SKGPU_OK;
}

static skgpu_rc launch_ops(skgpu_plan *p, bool time_ops) {
    skgpu_ctx *c = p->ctx;
    cudaStream_t s = c->stream;
    const float *gains = (const float *)p->gains.dev;
    for (auto &op : p->ops) {
        if (op.kind == OP_CONVERT) {
            const uint32_t grid = op.cap * op.tiles;
            if (time_ops) { skgpu_rc rc = op_event(op, 0, false, s); if (rc) return rc; }
            switch (op.mode) {
                case SKGPU_CVT_F32_TO_F32: k_convert<SKGPU_CVT_F32_TO_F32><<<grid, CVT_THREADS, 0, s>>>(op.d_hdr, (const skgpu_seg *)op.d_tab, gains, p->arena, op.tiles); break;
                case SKGPU_CVT_F32_TO_S16: k_convert<SKGPU_CVT_F32_TO_S16><<<grid, CVT_THREADS, 0, s>>>(op.d_hdr, (const skgpu_seg *)op.d_tab, gains, p->arena, op.tiles); break;
                default: k_convert<SKGPU_CVT_S16_TO_F32><<<grid, CVT_THREADS, 0, s>>>(op.d_hdr, (const skgpu_seg *)op.d_tab, gains, p->arena, op.tiles); break;
            }
            CU(cudaGetLastError());
            if (time_ops) { skgpu_rc rc = op_event(op, 0, true, s); if (rc) return rc; }
        } else if (op.kind == OP_RESAMPLE) {
            const skgpu_rs_item *items = (const skgpu_rs_item *)op.d_tab;
            if (time_ops) { skgpu_rc rc = op_event(op, 0, false, s); if (rc) return rc; }
            if (op.rs_sinc) k_phase<<<(op.cap + PHASE_THREADS - 1) / PHASE_THREADS, PHASE_THREADS, 0, s>>>(op.d_hdr, items, c->st, p->arena, op.results_off);
            else if (op.rs_prog) k_phase_prog<<<(op.cap + PHASE_THREADS - 1) / PHASE_THREADS, PHASE_THREADS, 0, s>>>(op.d_hdr, items, c->st, p->arena, op.results_off, op.rs_pd);
            else k_phase<<<(op.cap + PHASE_THREADS - 1) / PHASE_THREADS, PHASE_THREADS, 0, s>>>(op.d_hdr, items, c->st, p->arena, op.results_off);
            CU(cudaGetLastError());
            if (time_ops) { skgpu_rc rc = op_event(op, 0, true, s); if (rc) return rc; rc = op_event(op, 1, false, s); if (rc) return rc; }
            if (op.rs_sinc && op.rs_sinc_tiled) {
                const uint32_t grid = std::min<uint32_t>((op.cap + op.sinc_dm.G - 1u) / op.sinc_dm.G, (uint32_t)c->sm_count);
                const uint32_t sm = op.sinc_dm.tab_bytes + 2u * op.sinc_dm.G * op.sinc_dm.stream_bytes;
                if (op.rs_channels == 2) k_resample_sinc_tiled<2><<<grid, SINCT_THREADS, sm, s>>>(op.d_hdr, items, c->st, p->arena, op.sinc_taps, op.sinc_dm);
                else k_resample_sinc_tiled<1><<<grid, SINCT_THREADS, sm, s>>>(op.d_hdr, items, c->st, p->arena, op.sinc_taps, op.sinc_dm);
            } else if (op.rs_sinc && op.rs_channels == 2) k_resample_sinc<2><<<op.cap, SINC_THREADS, op.smem_bytes, s>>>(op.d_hdr, items, c->st, p->arena);
            else if (op.rs_sinc) k_resample_sinc<1><<<op.cap, SINC_THREADS, op.smem_bytes, s>>>(op.d_hdr, items, c->st, p->arena);
            else if (op.rs_prog && op.rs_channels == 2) k_resample_prog<2><<<op.cap, RSP_THREADS, op.smem_bytes, s>>>(op.d_hdr, items, c->st, p->arena, op.smem_frames, op.rs_pd, 1.0f);
            else if (op.rs_prog) k_resample_prog<1><<<op.cap, RSP_THREADS, op.smem_bytes, s>>>(op.d_hdr, items, c->st, p->arena, op.smem_frames, op.rs_pd, 1.0f);
            else if (op.rs_channels == 2) k_resample<2><<<op.cap, RS_THREADS, op.smem_bytes, s>>>(op.d_hdr, items, c->st, p->arena, op.smem_frames);
            else if (op.rs_channels == 1) k_resample<1><<<op.cap, RS_THREADS, op.smem_bytes, s>>>(op.d_hdr, items, c->st, p->arena, op.smem_frames);
            else k_resample<0><<<op.cap, RS_THREADS, op.smem_bytes, s>>>(op.d_hdr, items, c->st, p->arena, op.smem_frames);
            CU(cudaGetLastError());
            if (time_ops) { skgpu_rc rc = op_event(op, 1, true, s); if (rc) return rc; }
        } else if (op.kind == OP_CHAIN) {
            skgpu_rc rc = launch_chain(p, op, op.d_hdr, op.cap, op.cap2, s, time_ops);
            if (rc) return rc;
        } else {
            const uint8_t *present = op.present.valid ? (const uint8_t *)op.present.dev : nullptr;
            if (time_ops) { skgpu_rc rc = op_event(op, 0, false, s); if (rc) return rc; }
            k_mix<<<op.cap * ((op.tiles + op.mix_tpc - 1) / op.mix_tpc), MIX_THREADS, 0, s>>>(op.d_hdr, (const skgpu_mix_group *)op.d_tab, (const skgpu_mix_input *)op.d_tab2, present, gains, c->st, p->arena, op.tiles, op.mix_tpc);
            CU(cudaGetLastError());
            if (op.has_fifo_inputs) {
                k_fifo_commit<<<(op.cap2 + 127) / 128, 128, 0, s>>>(op.d_hdr, (const skgpu_mix_input *)op.d_tab2, present, c->st);
                CU(cudaGetLastError());
            }
            if (time_ops) { skgpu_rc rc = op_event(op, 0, true, s); if (rc) return rc; }
        }
    }
    if (p->bank_stride && p->ops.size() != 1) {  // bank parity of the next tick (a chain-only plan advances it inside k_chain)
        k_tick_advance<<<1, 1, 0, s>>>(p->d_tick);
        CU(cudaGetLastError());
    }
    return SKGPU_OK;
}

extern "C" uint32_t skgpu_plan_launches_per_tick(const skgpu_plan *p) {
    if (!p) return 0;
    uint32_t n = 0;
    for (auto &op : p->ops) n += (op.kind == OP_RESAMPLE || op.kind == OP_CHAIN) ? 2 : (op.kind == OP_MIX && op.has_fifo_inputs) ? 2 : 1;
    if (p->bank_stride && p->ops.size() != 1) n += 1;
    return n;
}

extern "C" skgpu_rc skgpu_plan_finalize(skgpu_plan *p) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    if (p->finalized) return fail(SKGPU_ERR_STATE, "plan already finalized");
    skgpu_ctx *c = p->ctx;
    CU(cudaSetDevice(c->device));
    if (!p->gains.valid) {  // an empty gain table keeps kernel arguments valid
        float one = 1.0f;
        skgpu_rc rc = skgpu_plan_set_gains(p, &one, 1);
        if (rc) return rc;
    }
    for (auto &op : p->ops) {
        if (op.kind == OP_RESAMPLE && op.rs_sinc && op.rs_sinc_tiled) {
            const int sm_fit = 220 * 1024;   // a cap, not a reservation (per function, shared by every plan of the process)
            CU(cudaFuncSetAttribute(k_resample_sinc_tiled<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_fit));
            CU(cudaFuncSetAttribute(k_resample_sinc_tiled<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm_fit));
        }
        if (op.kind == OP_RESAMPLE && op.smem_bytes > 48u * 1024u) {
            if (op.rs_sinc && op.rs_channels == 2) CU(cudaFuncSetAttribute(k_resample_sinc<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
            else if (op.rs_sinc) CU(cudaFuncSetAttribute(k_resample_sinc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
            else if (op.rs_prog && op.rs_channels == 2) CU(cudaFuncSetAttribute(k_resample_prog<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
            else if (op.rs_prog) CU(cudaFuncSetAttribute(k_resample_prog<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
            else if (op.rs_channels == 2) CU(cudaFuncSetAttribute(k_resample<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
            else if (op.rs_channels == 1) CU(cudaFuncSetAttribute(k_resample<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
            else CU(cudaFuncSetAttribute(k_resample<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
        }
    }
    for (auto &op : p->ops) {
        if (op.kind == OP_CHAIN) {
            auto kfn = chain_kernel(op.chain_oc, op.chain_iters, op.chain_kind);
            CU(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)op.smem_bytes));
            int per_sm = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, CH_THREADS, op.smem_bytes));
            if (per_sm < 1) return fail(SKGPU_ERR_INVALID, "chain op: kernel does not fit on an SM (%u bytes of shared memory)", op.smem_bytes);
            op.chain_grid = (uint32_t)per_sm * (uint32_t)c->sm_count;  // persistent CTAs: every SM fully occupied, each loops over sessions
        }
    }
    skgpu_rc rc = ctx_flush(c);
    if (rc) return rc;
    rc = upload_dirty(p);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    // capture the kernel sequence once: a tick becomes a single graph launch (hides per-kernel launch cost)
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    rc = launch_ops(p, false);
    cudaError_t e = cudaStreamEndCapture(c->stream, &p->graph);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(SKGPU_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    if (!p->ops.empty()) CU(cudaGraphInstantiate(&p->graph_exec, p->graph, 0));
    p->finalized = true;
    return SKGPU_OK;
}

// the two kernels of the chain op, separately (sliced ticks run them on different streams)
static skgpu_rc launch_chain_phase(skgpu_plan *p, Op &op, const OpHeader *d_hdr, uint32_t n_inputs, cudaStream_t s) {
    skgpu_ctx *c = p->ctx;
    const uint8_t *present = op.present.valid ? (const uint8_t *)op.present.dev : nullptr;
    k_phase_chain<<<(std::max(n_inputs, 1u) + PHASE_CHAIN_THREADS - 1) / PHASE_CHAIN_THREADS, PHASE_CHAIN_THREADS, 0, s>>>(
        d_hdr, (const skgpu_chain_input *)op.d_tab2, present, (const float *)p->gains.dev, c->st, p->arena, p->d_tick, p->bank_stride, op.chain_F,
        op.results_off, op.chain_dm, op.d_rec);
    CU(cudaGetLastError());
    return SKGPU_OK;
}
static skgpu_rc launch_chain_main(skgpu_plan *p, Op &op, const OpHeader *d_hdr, uint32_t n_groups, cudaStream_t s, bool advance) {
    skgpu_ctx *c = p->ctx;
    const uint32_t grid = std::max(1u, std::min<uint32_t>(n_groups, op.chain_grid));
    auto kfn = chain_kernel(op.chain_oc, op.chain_iters, op.chain_kind);
    kfn<<<grid, CH_THREADS, op.smem_bytes, s>>>(d_hdr, (const skgpu_chain_group *)op.d_tab, op.d_rec, (const float *)p->gains.dev, c->st, p->arena,
                                                 op.chain_F, op.chain_dm, advance ? p->d_tick : nullptr);
    CU(cudaGetLastError());
    return SKGPU_OK;
}

// One tick, slice by slice, on four streams: uploads on the context stream, k_phase_chain on stream_p, k_chain on stream_k,
// read-backs on stream_d2h.
//   phase(i)   is data independent (stream state + presence only): all slices' phase kernels are issued at the start of the
//              tick and run while the first slice still uploads; with inputs resident they overlap the previous slice's k_chain
//   upload(i)  follows the previous tick's kernels (they read the other bank as "previous chunk", and the per-tick tables)
//   chain(i)   waits for phase(i), upload(i) and the previous tick's read-back of the rows it overwrites
//   read-back(i) waits for chain(i)
// so slice i's results are in host memory while slice i + 1 .. n still upload (SURVEY 8d latency: upload-done -> read-back-done).
static skgpu_rc sliced_prepare(skgpu_plan *p, Op &op, uint32_t n) {
    skgpu_ctx *c = p->ctx;
    if (!p->ev_tables) {
        CU(cudaEventCreateWithFlags(&p->ev_tables, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&p->ev_k_all, cudaEventDisableTiming));
        CU(cudaEventCreate(&p->ev_t0[0]));
        CU(cudaEventCreate(&p->ev_t0[1]));
    }
    for (int a = 0; a < 2; ++a)
        while (p->ev_up[a].size() < n) {
            cudaEvent_t e1, e2, e3;
            CU(cudaEventCreate(&e1)); CU(cudaEventCreate(&e2)); CU(cudaEventCreate(&e3));
            p->ev_up[a].push_back(e1); p->ev_k[a].push_back(e2); p->ev_done[a].push_back(e3);
        }
    while (p->ev_p.size() < n) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        p->ev_p.push_back(e);
    }
    if (op.slices_dirty) {
        uint32_t g0 = 0, i0 = 0;
        // the staging buffer may still be read by an earlier upload: slices change only with the tables (rare), so just wait
        CU(cudaStreamSynchronize(c->stream));
        CU(cudaStreamSynchronize(c->stream_p));
        CU(cudaStreamSynchronize(c->stream_k));
        for (uint32_t k = 0; k < n; ++k) {
            op.h_hdr_sl[k].count = op.slices[k].group_end - g0;
            op.h_hdr_sl[k].count2 = op.slices[k].input_end - i0;
            op.h_hdr_sl[k].first = g0;
            op.h_hdr_sl[k].first2 = i0;
            g0 = op.slices[k].group_end; i0 = op.slices[k].input_end;
        }
        CU(cudaMemcpyAsync(op.d_hdr_sl, op.h_hdr_sl, n * sizeof(OpHeader), cudaMemcpyHostToDevice, c->stream));
        op.slices_dirty = false;
        if (p->sl_graph_exec) { cudaGraphExecDestroy(p->sl_graph_exec); p->sl_graph_exec = nullptr; }
        if (p->sl_graph) { cudaGraphDestroy(p->sl_graph); p->sl_graph = nullptr; }
    }
    return SKGPU_OK;
}

// kernels of one sliced tick as a fork / join between stream_k and stream_p (used directly, and captured as a CUDA graph for
// ticks without transfers): phase(i + 1) runs next to chain(i)
static skgpu_rc sliced_kernels(skgpu_plan *p, Op &op, uint32_t n, cudaEvent_t fork) {
    skgpu_ctx *c = p->ctx;
    cudaStream_t sk = c->stream_k, sp = c->stream_p;
    CU(cudaEventRecord(fork, sk));
    CU(cudaStreamWaitEvent(sp, fork, 0));
    uint32_t g0 = 0, i0 = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const skgpu_slice &sl = op.slices[k];
        skgpu_rc rc = launch_chain_phase(p, op, op.d_hdr_sl + k, sl.input_end - i0, sp);
        if (rc) return rc;
        CU(cudaEventRecord(p->ev_p[k], sp));
        CU(cudaStreamWaitEvent(sk, p->ev_p[k], 0));
        rc = launch_chain_main(p, op, op.d_hdr_sl + k, sl.group_end - g0, sk, k + 1 == n);   // the last slice advances the bank parity
        if (rc) return rc;
        g0 = sl.group_end; i0 = sl.input_end;
    }
    return SKGPU_OK;
}

static skgpu_rc submit_sliced(skgpu_plan *p, const void *host_in, void *host_out, bool do_h2d, bool do_d2h, uint32_t flags) {
    skgpu_ctx *c = p->ctx;
    cudaStream_t s = c->stream, sk = c->stream_k, sp = c->stream_p, sd = c->stream_d2h;
    if (p->ops.size() != 1 || p->ops[0].kind != OP_CHAIN) return fail(SKGPU_ERR_STATE, "sliced ticks need a plan whose only op is the chain op");
    Op &op = p->ops[0];
    if (op.slices.empty()) return fail(SKGPU_ERR_STATE, "no slices set (skgpu_plan_set_slices / skgpu_plan_auto_slices; table updates clear them)");
    const uint32_t n = (uint32_t)op.slices.size();
    skgpu_rc rc = sliced_prepare(p, op, n);
    if (rc) return rc;
    const uint32_t par = (uint32_t)((p->tick + 1) & 1ull), opar = par ^ 1u;   // parity of THIS tick's number / of the previous tick
    const bool prev_sliced = p->sliced_pending && p->sl_n[opar] > 0;
    // the previous tick's kernels read the per-tick tables and, as "previous chunk", the bank this tick uploads into
    if (p->sliced_pending) CU(cudaStreamWaitEvent(s, p->ev_k_all, 0));
    rc = upload_dirty(p);
    if (rc) return rc;
    CU(cudaEventRecord(p->ev_tables, s));             // tables, gains, presence (and any unsliced tick before) are on s
    CU(cudaStreamWaitEvent(sk, p->ev_tables, 0));
    CU(cudaEventRecord(p->ev_t0[par], s));
    if (!do_h2d && !do_d2h) {
        // ---- inputs resident, nothing read back: the kernel network alone, replayed as one CUDA graph
        if ((flags & SKGPU_SUBMIT_GRAPH) != 0) {
            if (!p->sl_graph_exec) {
                if (!p->ev_fork) CU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
                CU(cudaStreamSynchronize(sk));
                CU(cudaStreamSynchronize(sp));
                CU(cudaStreamBeginCapture(sk, cudaStreamCaptureModeThreadLocal));
                rc = sliced_kernels(p, op, n, p->ev_fork);
                const cudaError_t e = cudaStreamEndCapture(sk, &p->sl_graph);
                if (rc) return rc;
                if (e != cudaSuccess) return fail(SKGPU_ERR_CUDA, "capture of the sliced tick failed: %s", cudaGetErrorString(e));
                CU(cudaGraphInstantiate(&p->sl_graph_exec, p->sl_graph, 0));
            }
            CU(cudaGraphLaunch(p->sl_graph_exec, sk));
        } else {
            if (!p->ev_fork) CU(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
            rc = sliced_kernels(p, op, n, p->ev_fork);
            if (rc) return rc;
        }
        p->tick++;
        CU(cudaEventRecord(p->ev_k_all, sk));
        CU(cudaStreamWaitEvent(sd, p->ev_k_all, 0));
        for (uint32_t k = 0; k < n; ++k) {   // keep the per-slice events defined (timing of such a tick is not meaningful)
            CU(cudaEventRecord(p->ev_up[par][k], s));
            CU(cudaEventRecord(p->ev_k[par][k], sk));
            CU(cudaEventRecord(p->ev_done[par][k], sd));
        }
        CU(cudaEventRecord(p->ev_d2h_done, sd));
        CU(cudaEventRecord(p->ev_tick_done[p->tick & 1ull], sd));
        p->sl_n[par] = n;
        p->sliced_pending = true;
        p->d2h_pending = false;
        p->timing_valid = false;
        return SKGPU_OK;
    }
    // ---- all phase kernels first (stream_p): they need no input byte of this tick (but they rewrite the result rows the
    // previous tick's read-back may still be copying)
    CU(cudaStreamWaitEvent(sp, p->ev_tables, 0));
    if (p->d2h_pending) CU(cudaStreamWaitEvent(sp, p->ev_d2h_done, 0));
    {
        uint32_t i0 = 0;
        for (uint32_t k = 0; k < n; ++k) {
            rc = launch_chain_phase(p, op, op.d_hdr_sl + k, op.slices[k].input_end - i0, sp);
            if (rc) return rc;
            CU(cudaEventRecord(p->ev_p[k], sp));
            i0 = op.slices[k].input_end;
        }
    }
    uint8_t *bank = p->arena + p->h2d_off + (p->tick & 1ull) * p->bank_stride;
    uint64_t up0 = 0;
    uint32_t g0 = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const skgpu_slice &sl = op.slices[k];
        if (do_h2d && sl.h2d_end > up0) CU(cudaMemcpyAsync(bank + up0, (const uint8_t *)host_in + up0, sl.h2d_end - up0, cudaMemcpyHostToDevice, s));
        up0 = std::max(up0, sl.h2d_end);
        CU(cudaEventRecord(p->ev_up[par][k], s));
        CU(cudaStreamWaitEvent(sk, p->ev_up[par][k], 0));
        CU(cudaStreamWaitEvent(sk, p->ev_p[k], 0));
        if (prev_sliced && k < p->sl_n[opar]) CU(cudaStreamWaitEvent(sk, p->ev_done[opar][k], 0));
        else if (p->d2h_pending && k == 0) CU(cudaStreamWaitEvent(sk, p->ev_d2h_done, 0));
        rc = launch_chain_main(p, op, op.d_hdr_sl + k, sl.group_end - g0, sk, k + 1 == n);   // the last slice advances the bank parity
        if (rc) return rc;
        CU(cudaEventRecord(p->ev_k[par][k], sk));
        CU(cudaStreamWaitEvent(sd, p->ev_k[par][k], 0));
        if (do_d2h)
            for (int r = 0; r < 2; ++r)
                if (sl.d2h_bytes[r]) CU(cudaMemcpyAsync((uint8_t *)host_out + sl.d2h_off[r], p->arena + p->d2h_off + sl.d2h_off[r], sl.d2h_bytes[r], cudaMemcpyDeviceToHost, sd));
        CU(cudaEventRecord(p->ev_done[par][k], sd));
        g0 = sl.group_end;
    }
    p->tick++;
    CU(cudaEventRecord(p->ev_k_all, sk));
    CU(cudaEventRecord(p->ev_d2h_done, sd));
    CU(cudaEventRecord(p->ev_tick_done[p->tick & 1ull], sd));
    p->sl_n[par] = n;
    p->sliced_pending = true;
    p->d2h_pending = do_d2h;
    p->timing_valid = false;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_tick_slice_timing(skgpu_plan *p, uint64_t tick, skgpu_slice_timing *out, uint32_t cap, uint32_t *n_out) {
    if (!p || !n_out) return fail(SKGPU_ERR_INVALID, "null argument");
    if (tick == 0 || tick > p->tick || tick + 1 < p->tick) return fail(SKGPU_ERR_INVALID, "tick %llu is not one of the two most recent ticks", (unsigned long long)tick);
    const uint32_t par = (uint32_t)(tick & 1ull), n = p->sl_n[par];
    if (n == 0) return fail(SKGPU_ERR_STATE, "tick %llu was not a sliced tick", (unsigned long long)tick);
    CU(cudaSetDevice(p->ctx->device));
    CU(cudaEventSynchronize(p->ev_done[par][n - 1]));
    *n_out = n;
    for (uint32_t k = 0; k < n && k < cap; ++k) {
        CU(cudaEventElapsedTime(&out[k].upload_done_ms, p->ev_t0[par], p->ev_up[par][k]));
        CU(cudaEventElapsedTime(&out[k].kernels_ms, p->ev_up[par][k], p->ev_k[par][k]));
        CU(cudaEventElapsedTime(&out[k].latency_ms, p->ev_up[par][k], p->ev_done[par][k]));
    }
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_tick_submit(skgpu_plan *p, const void *host_in, void *host_out, uint32_t flags) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    if (!p->finalized) return fail(SKGPU_ERR_STATE, "plan not finalized");
    skgpu_ctx *c = p->ctx;
    cudaStream_t s = c->stream;
    CU(cudaSetDevice(c->device));
    const bool do_h2d = !(flags & SKGPU_SUBMIT_NO_H2D) && p->h2d_bytes;
    const bool do_d2h = !(flags & SKGPU_SUBMIT_NO_D2H) && p->d2h_bytes;
    if (do_h2d && !host_in) return fail(SKGPU_ERR_INVALID, "host_in is null");
    if (do_d2h && !host_out) return fail(SKGPU_ERR_INVALID, "host_out is null");
    skgpu_rc rc = ctx_flush(c);
    if (rc) return rc;
    if (flags & SKGPU_SUBMIT_SLICED) return submit_sliced(p, host_in, host_out, do_h2d, do_d2h, flags);
    if (p->sliced_pending) {   // the previous tick's kernels ran on the kernel stream: this tick's upload / kernels follow them
        CU(cudaStreamWaitEvent(s, p->ev_k_all, 0));
        p->sliced_pending = false;
    }
    rc = upload_dirty(p);
    if (rc) return rc;
    const bool overlap = (flags & SKGPU_SUBMIT_OVERLAP_D2H) != 0 && do_d2h;
    CU(cudaEventRecord(p->e0, s));
    if (do_h2d) CU(cudaMemcpyAsync(p->arena + p->h2d_off + (p->tick & 1ull) * p->bank_stride, host_in, p->h2d_bytes, cudaMemcpyHostToDevice, s));
    p->tick++;
    CU(cudaEventRecord(p->e1, s));
    // the kernels overwrite the output range: they must not start before the previous tick's read-back has finished
    if (p->d2h_pending) CU(cudaStreamWaitEvent(s, p->ev_d2h_done, 0));
    if ((flags & SKGPU_SUBMIT_GRAPH) && p->graph_exec) {
        CU(cudaGraphLaunch(p->graph_exec, s));
    } else {
        rc = launch_ops(p, (flags & SKGPU_SUBMIT_TIME_OPS) != 0);
        if (rc) return rc;
    }
    CU(cudaEventRecord(p->e2, s));
    if (overlap) {
        // read-back on the second stream: overlaps the NEXT tick's upload (H2D and D2H use opposite PCIe directions)
        CU(cudaEventRecord(p->ev_kernels_done, s));
        CU(cudaStreamWaitEvent(c->stream_d2h, p->ev_kernels_done, 0));
        CU(cudaMemcpyAsync(host_out, p->arena + p->d2h_off, p->d2h_bytes, cudaMemcpyDeviceToHost, c->stream_d2h));
        CU(cudaEventRecord(p->ev_d2h_done, c->stream_d2h));
        CU(cudaEventRecord(p->e3, c->stream_d2h));
        CU(cudaEventRecord(p->ev_tick_done[p->tick & 1ull], c->stream_d2h));
        p->d2h_pending = true;
    } else {
        if (do_d2h) CU(cudaMemcpyAsync(host_out, p->arena + p->d2h_off, p->d2h_bytes, cudaMemcpyDeviceToHost, s));
        CU(cudaEventRecord(p->e3, s));
        CU(cudaEventRecord(p->ev_tick_done[p->tick & 1ull], s));
    }
    p->timing_valid = true;
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_tick_wait_for(skgpu_plan *p, uint64_t tick) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    if (tick == 0 || tick > p->tick || tick + 1 < p->tick) return fail(SKGPU_ERR_INVALID, "tick %llu is not one of the two most recent ticks (%llu submitted)", (unsigned long long)tick, (unsigned long long)p->tick);
    CU(cudaSetDevice(p->ctx->device));
    CU(cudaEventSynchronize(p->ev_tick_done[tick & 1ull]));
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_tick_wait(skgpu_plan *p, skgpu_tick_timing *t) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    CU(cudaSetDevice(p->ctx->device));
    CU(cudaStreamSynchronize(p->ctx->stream));
    CU(cudaStreamSynchronize(p->ctx->stream_p));
    CU(cudaStreamSynchronize(p->ctx->stream_k));
    CU(cudaStreamSynchronize(p->ctx->stream_d2h));
    p->d2h_pending = false;
    p->sliced_pending = false;
    if (t) {
        memset(t, 0, sizeof(*t));
        if (p->timing_valid) {
            CU(cudaEventElapsedTime(&t->h2d_ms, p->e0, p->e1));
            CU(cudaEventElapsedTime(&t->kernels_ms, p->e1, p->e2));
            CU(cudaEventElapsedTime(&t->d2h_ms, p->e2, p->e3));
            CU(cudaEventElapsedTime(&t->total_ms, p->e0, p->e3));
        }
    }
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_plan_op_time(skgpu_plan *p, uint32_t opi, uint32_t sub, float *avg_ms, uint32_t *n_samples) {
    if (!p || opi >= p->ops.size() || sub > 1) return fail(SKGPU_ERR_INVALID, "invalid op");
    CU(cudaSetDevice(p->ctx->device));
    CU(cudaStreamSynchronize(p->ctx->stream));
    Op &op = p->ops[opi];
    double sum = 0;
    for (uint32_t i = 0; i < op.ev_used[sub]; ++i) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, op.ev[sub][i].first, op.ev[sub][i].second));
        sum += ms;
    }
    if (avg_ms) *avg_ms = op.ev_used[sub] ? (float)(sum / op.ev_used[sub]) : 0.0f;
    if (n_samples) *n_samples = op.ev_used[sub];
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_plan_reset_op_times(skgpu_plan *p) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    CU(cudaStreamSynchronize(p->ctx->stream));
    for (auto &op : p->ops) op.ev_used[0] = op.ev_used[1] = 0;
    return SKGPU_OK;
}

// ---- arena access / stopwatch

extern "C" skgpu_rc skgpu_arena_upload(skgpu_plan *p, uint64_t off, const void *host, size_t bytes) {
    if (!p || !host) return fail(SKGPU_ERR_INVALID, "null argument");
    skgpu_rc rc = check_range(p, off, bytes, "arena upload");
    if (rc) return rc;
    CU(cudaSetDevice(p->ctx->device));
    CU(cudaMemcpyAsync(p->arena + off, host, bytes, cudaMemcpyHostToDevice, p->ctx->stream));
    CU(cudaStreamSynchronize(p->ctx->stream));
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_arena_download(skgpu_plan *p, uint64_t off, void *host, size_t bytes) {
    if (!p || !host) return fail(SKGPU_ERR_INVALID, "null argument");
    skgpu_rc rc = check_range(p, off, bytes, "arena download");
    if (rc) return rc;
    CU(cudaSetDevice(p->ctx->device));
    CU(cudaMemcpyAsync(host, p->arena + off, bytes, cudaMemcpyDeviceToHost, p->ctx->stream));
    CU(cudaStreamSynchronize(p->ctx->stream));
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_arena_fill(skgpu_plan *p, uint64_t off, int byte_value, size_t bytes) {
    if (!p) return fail(SKGPU_ERR_INVALID, "null plan");
    skgpu_rc rc = check_range(p, off, bytes, "arena fill");
    if (rc) return rc;
    CU(cudaSetDevice(p->ctx->device));
    CU(cudaMemsetAsync(p->arena + off, byte_value, bytes, p->ctx->stream));
    return SKGPU_OK;
}

extern "C" skgpu_rc skgpu_timer_start(skgpu_ctx *c) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaEventRecord(c->tm0, c->stream));
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_timer_stop(skgpu_ctx *c) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    // the stop mark covers the read-back stream too
    CU(cudaEventRecord(c->tm1, c->stream_d2h));
    CU(cudaStreamWaitEvent(c->stream, c->tm1, 0));
    CU(cudaEventRecord(c->tm1, c->stream));
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_timer_elapsed_ms(skgpu_ctx *c, float *ms) {
    if (!c || !ms) return fail(SKGPU_ERR_INVALID, "null argument");
    CU(cudaSetDevice(c->device));
    CU(cudaEventSynchronize(c->tm1));
    CU(cudaEventElapsedTime(ms, c->tm0, c->tm1));
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_ctx_sync(skgpu_ctx *c) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->stream_p));
    CU(cudaStreamSynchronize(c->stream_k));
    CU(cudaStreamSynchronize(c->stream_d2h));
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_ctx_flush_l2(skgpu_ctx *c) {
    if (!c) return fail(SKGPU_ERR_INVALID, "null context");
    CU(cudaSetDevice(c->device));
    if (!c->l2buf) {
        c->l2n = (size_t)256 * 1024 * 1024 / sizeof(uint4);  // 256 MB > 126 MB L2
        CU(dalloc(&c->l2buf, c->l2n));
    }
    k_l2_flush<<<c->sm_count * 8, 256, 0, c->stream>>>(c->l2buf, c->l2n);
    CU(cudaGetLastError());
    return SKGPU_OK;
}
