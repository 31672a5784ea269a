// router.cpp -- single-process multi-GPU session router (include/skgpu_router.h): fnv1a64(session id) % n_gpus, one hub and
// one NUMA-pinned tick thread per GPU. No collective: sessions never leave their GPU.
#include "../../../include/skgpu_router.h"

#include <sched.h>

#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_err;
skgpu_rc rfail(skgpu_rc rc, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return rc;
}

enum Cmd { CMD_NONE = 0, CMD_CREATE, CMD_TICK, CMD_WAIT, CMD_RUN, CMD_EXIT };

struct Gpu {
    int32_t device = 0;
    skgpu_hub *hub = nullptr;
    int32_t numa = -1;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    Cmd cmd = CMD_NONE;
    bool done = true;
    skgpu_rc rc = SKGPU_OK;
    std::string err;
    uint32_t run_n = 0;
    double run_ms = 0.0;
};

}  // namespace

struct skgpu_router {
    std::vector<Gpu *> gpus;
    skgpu_hub_config cfg{};
    std::vector<uint32_t> rates;
};

extern "C" uint64_t skgpu_fnv1a64(const void *data, size_t len) {
    const uint8_t *p = static_cast<const uint8_t *>(data);
    uint64_t h = 0xcbf29ce484222325ull;          // FNV_OFFSET_BASIS (session.rs:36)
    for (size_t i = 0; i < len; ++i) {
        h ^= p[i];
        h *= 0x100000001b3ull;                   // FNV_PRIME (session.rs:37)
    }
    return h;
}
extern "C" uint32_t skgpu_router_gpu_for(const void *id, size_t len, uint32_t n) { return n ? (uint32_t)(skgpu_fnv1a64(id, len) % n) : 0u; }

static void gpu_thread(skgpu_router *r, Gpu *g) {
    for (;;) {
        Cmd cmd;
        {
            std::unique_lock<std::mutex> lk(g->mu);
            g->cv.wait(lk, [&] { return g->cmd != CMD_NONE; });
            cmd = g->cmd;
        }
        skgpu_rc rc = SKGPU_OK;
        if (cmd == CMD_CREATE) {
            // the hub is created HERE so that its context, streams and pinned arenas belong to this thread; the arenas are bound
            // to the GPU's NUMA node by skgpu_pinned_alloc, and the thread pins itself to that node's CPUs right after
            rc = skgpu_hub_create(g->device, &r->cfg, &g->hub);
            if (rc == SKGPU_OK) {
                g->numa = skgpu_hub_bind_thread(g->hub);
            } else {
                g->err = skgpu_hub_last_error();
            }
        } else if (cmd == CMD_TICK) {
            rc = skgpu_hub_tick(g->hub);
            if (rc != SKGPU_OK) g->err = skgpu_hub_last_error();
        } else if (cmd == CMD_WAIT) {
            rc = skgpu_hub_wait(g->hub, nullptr);
            if (rc != SKGPU_OK) g->err = skgpu_hub_last_error();
        } else if (cmd == CMD_RUN) {
            // steady-state zero-copy loop: commit, submit tick n + 1, collect tick n (its read-back overlaps the next upload)
            const auto t0 = std::chrono::steady_clock::now();
            uint64_t prev = 0;
            for (uint32_t i = 0; i < g->run_n && rc == SKGPU_OK; ++i) {
                rc = skgpu_hub_commit_all(g->hub);
                if (rc == SKGPU_OK) rc = skgpu_hub_tick(g->hub);
                if (rc == SKGPU_OK && prev) rc = skgpu_hub_wait_tick(g->hub, prev);
                prev = skgpu_hub_ticks(g->hub);
            }
            if (rc == SKGPU_OK) rc = skgpu_hub_wait(g->hub, nullptr);
            g->run_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / (g->run_n ? g->run_n : 1);
            if (rc != SKGPU_OK) g->err = skgpu_hub_last_error();
        }
        {
            std::lock_guard<std::mutex> lk(g->mu);
            g->rc = rc;
            g->cmd = CMD_NONE;
            g->done = true;
        }
        g->cv.notify_all();
        if (cmd == CMD_EXIT) return;
    }
}

static void post(Gpu *g, Cmd c) {
    {
        std::lock_guard<std::mutex> lk(g->mu);
        g->cmd = c;
        g->done = false;
    }
    g->cv.notify_all();
}
static skgpu_rc collect(skgpu_router *r, const char *what) {
    skgpu_rc rc = SKGPU_OK;
    for (size_t i = 0; i < r->gpus.size(); ++i) {
        Gpu *g = r->gpus[i];
        std::unique_lock<std::mutex> lk(g->mu);
        g->cv.wait(lk, [&] { return g->done; });
        if (g->rc != SKGPU_OK && rc == SKGPU_OK) rc = rfail(g->rc, "%s on GPU %d: %s", what, g->device, g->err.c_str());
    }
    return rc;
}
static skgpu_rc broadcast(skgpu_router *r, Cmd c, const char *what) {
    for (Gpu *g : r->gpus) post(g, c);
    return collect(r, what);
}

extern "C" skgpu_rc skgpu_router_create(const int32_t *devices, uint32_t n, const skgpu_hub_config *cfg, skgpu_router **out) {
    if (!devices || !n || !cfg || !out) return rfail(SKGPU_ERR_INVALID, "null argument");
    if (n > 64) return rfail(SKGPU_ERR_INVALID, "at most 64 GPUs");
    skgpu_router *r = new skgpu_router();
    r->cfg = *cfg;
    r->rates.assign(cfg->in_rates, cfg->in_rates + cfg->n_in_rates);
    r->cfg.in_rates = r->rates.data();
    for (uint32_t i = 0; i < n; ++i) {
        Gpu *g = new Gpu();
        g->device = devices[i];
        r->gpus.push_back(g);
        g->th = std::thread(gpu_thread, r, g);
    }
    const skgpu_rc rc = broadcast(r, CMD_CREATE, "hub creation");
    if (rc != SKGPU_OK) {
        const std::string keep = g_err;
        skgpu_router_destroy(r);
        g_err = keep;
        return rc;
    }
    *out = r;
    return SKGPU_OK;
}

extern "C" void skgpu_router_destroy(skgpu_router *r) {
    if (!r) return;
    for (Gpu *g : r->gpus) {
        post(g, CMD_EXIT);
        if (g->th.joinable()) g->th.join();
        if (g->hub) skgpu_hub_destroy(g->hub);
        delete g;
    }
    delete r;
}

extern "C" uint32_t skgpu_router_gpus(const skgpu_router *r) { return r ? (uint32_t)r->gpus.size() : 0u; }
extern "C" skgpu_hub *skgpu_router_hub(skgpu_router *r, uint32_t g) { return (r && g < r->gpus.size()) ? r->gpus[g]->hub : nullptr; }
extern "C" int32_t skgpu_router_numa_node(skgpu_router *r, uint32_t g) { return (r && g < r->gpus.size()) ? r->gpus[g]->numa : -1; }

static Gpu *gpu_of(skgpu_router *r, skgpu_session_handle h) {
    const uint32_t g = (uint32_t)(h >> 32);
    return (r && g < r->gpus.size()) ? r->gpus[g] : nullptr;
}
#define HUBCALL(g, call)                                                     \
    do {                                                                     \
        if (!(g)) return rfail(SKGPU_ERR_INVALID, "invalid session handle"); \
        const skgpu_rc rc__ = (call);                                        \
        if (rc__ != SKGPU_OK) return rfail(rc__, "%s", skgpu_hub_last_error()); \
        return SKGPU_OK;                                                     \
    } while (0)

extern "C" skgpu_rc skgpu_router_session_open(skgpu_router *r, const void *id, size_t id_len, uint32_t n_inputs, const uint32_t *in_rates,
                                              skgpu_session_handle *handle_out) {
    if (!r || !id || !handle_out) return rfail(SKGPU_ERR_INVALID, "null argument");
    const uint32_t gi = skgpu_router_gpu_for(id, id_len, (uint32_t)r->gpus.size());
    uint32_t s = 0;
    const skgpu_rc rc = skgpu_hub_session_open(r->gpus[gi]->hub, n_inputs, in_rates, &s);
    if (rc != SKGPU_OK) return rfail(rc, "GPU %d: %s", r->gpus[gi]->device, skgpu_hub_last_error());
    *handle_out = ((uint64_t)gi << 32) | s;
    return SKGPU_OK;
}
extern "C" skgpu_rc skgpu_router_session_close(skgpu_router *r, skgpu_session_handle h) { Gpu *g = gpu_of(r, h); HUBCALL(g, skgpu_hub_session_close(g->hub, (uint32_t)h)); }
extern "C" skgpu_rc skgpu_router_push(skgpu_router *r, skgpu_session_handle h, uint32_t input, const void *samples, uint32_t n_frames) {
    Gpu *g = gpu_of(r, h);
    HUBCALL(g, skgpu_hub_push(g->hub, (uint32_t)h, input, samples, n_frames));
}
extern "C" skgpu_rc skgpu_router_set_input_gain(skgpu_router *r, skgpu_session_handle h, uint32_t input, float gain) {
    Gpu *g = gpu_of(r, h);
    HUBCALL(g, skgpu_hub_set_input_gain(g->hub, (uint32_t)h, input, gain));
}
extern "C" skgpu_rc skgpu_router_set_master_gain(skgpu_router *r, skgpu_session_handle h, float gain) {
    Gpu *g = gpu_of(r, h);
    HUBCALL(g, skgpu_hub_set_master_gain(g->hub, (uint32_t)h, gain));
}
extern "C" skgpu_rc skgpu_router_session_output(skgpu_router *r, skgpu_session_handle h, const void **samples, uint32_t *n_mixed, uint32_t *status) {
    Gpu *g = gpu_of(r, h);
    HUBCALL(g, skgpu_hub_session_output(g->hub, (uint32_t)h, samples, n_mixed, status));
}

extern "C" skgpu_rc skgpu_router_tick(skgpu_router *r) {
    if (!r) return rfail(SKGPU_ERR_INVALID, "null router");
    return broadcast(r, CMD_TICK, "tick");
}
extern "C" skgpu_rc skgpu_router_wait(skgpu_router *r) {
    if (!r) return rfail(SKGPU_ERR_INVALID, "null router");
    return broadcast(r, CMD_WAIT, "wait");
}
extern "C" skgpu_rc skgpu_router_run_ticks(skgpu_router *r, uint32_t n, double *ms_out) {
    if (!r) return rfail(SKGPU_ERR_INVALID, "null router");
    for (Gpu *g : r->gpus) g->run_n = n;
    const skgpu_rc rc = broadcast(r, CMD_RUN, "tick loop");
    double worst = 0.0;
    for (Gpu *g : r->gpus) worst = g->run_ms > worst ? g->run_ms : worst;
    if (ms_out) *ms_out = worst;
    return rc;
}

extern "C" const char *skgpu_router_last_error(void) { return g_err.c_str(); }
