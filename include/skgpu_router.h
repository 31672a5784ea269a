/*
 * skgpu_router.h -- single-process, multi-GPU session router above the frame-batching layer (libskgpu_hub.so), C ABI.
 *
 * sessions are the unit of sharding -- every input stream of a session's mixer lives on ONE GPU for the
 * session's lifetime, so the data path has no exchange step and no collective (NCCL is not used). The router is what a
 * StreamKit engine process on an 8-GPU box owns instead of eight processes:
 *   - gpu = fnv1a64(session id) % n_gpus, the hash the reference already applies to session ids
 *     (apps/skit/src/session.rs:35-45; ids are UUIDv4 strings, :180);
 *   - one hub (skgpu_hub.h) per GPU, created on a worker thread that is pinned to the CPUs of that GPU's NUMA node, so
 *     the hub's pinned arenas (NUMA-bound by skgpu_pinned_alloc) and every gather copy stay on the GPU's socket;
 *   - one tick thread per GPU: skgpu_router_tick wakes all of them, each submits its hub's tick (asynchronous, sliced) --
 *     the per-tick host work (presence tables, absent-stream copies, launches) runs in parallel across GPUs.
 * Pushers call skgpu_router_push from any thread; it forwards to the owning hub, whose cut against the tick applies.
 */
#ifndef SKGPU_ROUTER_H
#define SKGPU_ROUTER_H

#include "skgpu_hub.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct skgpu_router skgpu_router;
typedef uint64_t skgpu_session_handle;   /* gpu index << 32 | hub session index */

/* message of the last error on the calling thread (borrowed) */
const char *skgpu_router_last_error(void);

/* FNV-1a 64 of the session id bytes (apps/skit/src/session.rs:35-45) */
uint64_t skgpu_fnv1a64(const void *data, size_t len);
/* the GPU a session id maps to among n_gpus (pure function: usable without any GPU) */
uint32_t skgpu_router_gpu_for(const void *session_id, size_t len, uint32_t n_gpus);

/* devices[n_gpus]: CUDA ordinals; cfg: the per-GPU hub configuration (capacities are per GPU) */
skgpu_rc skgpu_router_create(const int32_t *devices, uint32_t n_gpus, const skgpu_hub_config *cfg, skgpu_router **out);
void skgpu_router_destroy(skgpu_router *r);
uint32_t skgpu_router_gpus(const skgpu_router *r);
/* the hub of GPU index g (for calls the router does not wrap); NULL if out of range */
skgpu_hub *skgpu_router_hub(skgpu_router *r, uint32_t g);
/* host NUMA node of GPU index g (-1 unknown) */
int32_t skgpu_router_numa_node(skgpu_router *r, uint32_t g);

skgpu_rc skgpu_router_session_open(skgpu_router *r, const void *session_id, size_t id_len, uint32_t n_inputs, const uint32_t *in_rates,
                                   skgpu_session_handle *handle_out);
skgpu_rc skgpu_router_session_close(skgpu_router *r, skgpu_session_handle h);
skgpu_rc skgpu_router_push(skgpu_router *r, skgpu_session_handle h, uint32_t input, const void *samples, uint32_t n_frames);
skgpu_rc skgpu_router_set_input_gain(skgpu_router *r, skgpu_session_handle h, uint32_t input, float gain);
skgpu_rc skgpu_router_set_master_gain(skgpu_router *r, skgpu_session_handle h, float gain);

/* one tick on every GPU: wakes the per-GPU tick threads and returns when all of them have SUBMITTED (asynchronous device work) */
skgpu_rc skgpu_router_tick(skgpu_router *r);
/* blocks until every GPU's last tick has finished */
skgpu_rc skgpu_router_wait(skgpu_router *r);
/* tick + wait on every GPU, each on its own thread, `n` times back to back with `commit_all` before every tick (producers
 * that write their pinned slots in place): the steady-state loop of a zero-copy engine, used by bench.py. ms_per_tick_out
 * receives the wall time per tick of the slowest GPU. */
skgpu_rc skgpu_router_run_ticks(skgpu_router *r, uint32_t n, double *ms_per_tick_out);
skgpu_rc skgpu_router_session_output(skgpu_router *r, skgpu_session_handle h, const void **samples, uint32_t *n_mixed, uint32_t *status);

#ifdef __cplusplus
}
#endif
#endif /* SKGPU_ROUTER_H */
