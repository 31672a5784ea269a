//! `extern "C"` view of streamkit_b200's C ABI. Field-for-field restatement of `include/skgpu_hub.h` and
//! `include/skgpu_router.h` (and the pieces of `include/skgpu_batch.h` they mention). UNCOMPILED (no Rust toolchain in the
//! build image) -- see rust/README.md.
//!
//! Conventions are those of the native plugin SDK (sdks/plugin-sdk/native/src/types.rs): `#[repr(C)]` PODs, no unwinding across
//! the boundary, every call returns `skgpu_rc` (0 = ok), the error text of the calling thread is BORROWED from
//! `skgpu_*_last_error()` until that thread's next error (types.rs:42-48 has the same rule for `CResult.error_message`).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_void};

pub type skgpu_rc = i32;
pub const SKGPU_OK: skgpu_rc = 0;
pub const SKGPU_ERR_INVALID: skgpu_rc = -1; // StreamKitError::Configuration
pub const SKGPU_ERR_CUDA: skgpu_rc = -2; // StreamKitError::Runtime, node -> Failed
pub const SKGPU_ERR_NOMEM: skgpu_rc = -3;
pub const SKGPU_ERR_STATE: skgpu_rc = -4;
pub const SKGPU_ERR_NODEVICE: skgpu_rc = -5;

pub const SKGPU_HUB_OUT_S16: u16 = 1;
pub const SKGPU_HUB_IN_S16: u16 = 2;
pub const SKGPU_SESSION_SYNC: u32 = 1;
pub const SKGPU_SESSION_RUNNING: u32 = 1;
pub const SKGPU_SESSION_DEGRADED: u32 = 2;
pub const SKGPU_SESSION_STOPPED: u32 = 4;
pub const SKGPU_HUB_RUNNING: u32 = 1;
pub const SKGPU_HUB_DEGRADED: u32 = 2;
pub const SKGPU_HUB_FAILED: u32 = 3;

#[repr(C)]
pub struct skgpu_hub {
    _private: [u8; 0],
}
#[repr(C)]
pub struct skgpu_router {
    _private: [u8; 0],
}
pub type skgpu_session_handle = u64; // gpu index << 32 | hub session index

#[repr(C)]
#[derive(Clone, Copy)]
pub struct skgpu_hub_config {
    pub max_sessions: u32,
    pub max_streams: u32,
    pub max_inputs_per_session: u32,
    pub out_rate: u32,
    pub out_frames: u32,
    pub channels: u16,
    pub flags: u16,
    pub in_rates: *const u32,
    pub n_in_rates: u32,
    pub jitter_frames: u32,
    pub slices: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct skgpu_tick_timing {
    pub h2d_ms: f32,
    pub kernels_ms: f32,
    pub d2h_ms: f32,
    pub total_ms: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct skgpu_hub_stats {
    pub received: u64,
    pub sent: u64,
    pub discarded: u64,
    pub errored: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct skgpu_session_state {
    pub state: u32,
    pub mixed: u32,
    pub slow_mask: u64,
    pub eof_mask: u64,
    pub newly_slow: u64,
    pub recovered: u64,
}

#[link(name = "skgpu_hub")]
extern "C" {
    pub fn skgpu_hub_last_error() -> *const c_char;
    pub fn skgpu_hub_create(device_ordinal: i32, cfg: *const skgpu_hub_config, out: *mut *mut skgpu_hub) -> skgpu_rc;
    pub fn skgpu_hub_destroy(hub: *mut skgpu_hub);
    pub fn skgpu_hub_session_open(hub: *mut skgpu_hub, n_inputs: u32, in_rates: *const u32, session_out: *mut u32) -> skgpu_rc;
    pub fn skgpu_hub_session_open_ex(
        hub: *mut skgpu_hub, n_inputs: u32, in_rates: *const u32, mode: u32, sync_timeout_ms: u32, session_out: *mut u32,
    ) -> skgpu_rc;
    pub fn skgpu_hub_session_close(hub: *mut skgpu_hub, session: u32) -> skgpu_rc;
    pub fn skgpu_hub_input_eof(hub: *mut skgpu_hub, session: u32, input: u32) -> skgpu_rc;
    pub fn skgpu_hub_session_state(hub: *mut skgpu_hub, session: u32, out: *mut skgpu_session_state) -> skgpu_rc;
    pub fn skgpu_hub_set_input_gain(hub: *mut skgpu_hub, session: u32, input: u32, gain: f32) -> skgpu_rc;
    pub fn skgpu_hub_set_master_gain(hub: *mut skgpu_hub, session: u32, gain: f32) -> skgpu_rc;
    pub fn skgpu_hub_chunk_frames(hub: *mut skgpu_hub, session: u32, input: u32, frames_out: *mut u32) -> skgpu_rc;
    pub fn skgpu_hub_push(hub: *mut skgpu_hub, session: u32, input: u32, samples: *const c_void, n_frames: u32) -> skgpu_rc;
    pub fn skgpu_hub_acquire(hub: *mut skgpu_hub, session: u32, input: u32, dst_out: *mut *mut c_void, n_frames_out: *mut u32) -> skgpu_rc;
    pub fn skgpu_hub_commit(hub: *mut skgpu_hub, session: u32, input: u32) -> skgpu_rc;
    pub fn skgpu_hub_tick(hub: *mut skgpu_hub) -> skgpu_rc;
    pub fn skgpu_hub_wait(hub: *mut skgpu_hub, timing: *mut skgpu_tick_timing) -> skgpu_rc;
    pub fn skgpu_hub_wait_tick(hub: *mut skgpu_hub, tick: u64) -> skgpu_rc;
    pub fn skgpu_hub_session_output(
        hub: *mut skgpu_hub, session: u32, samples: *mut *const c_void, n_mixed: *mut u32, status: *mut u32,
    ) -> skgpu_rc;
    pub fn skgpu_hub_get_stats(hub: *mut skgpu_hub, out: *mut skgpu_hub_stats) -> skgpu_rc;
    pub fn skgpu_hub_state(hub: *const skgpu_hub, reason_out: *mut *const c_char) -> u32;
    pub fn skgpu_hub_ticks(hub: *const skgpu_hub) -> u64;
}

#[link(name = "skgpu_router")]
extern "C" {
    pub fn skgpu_router_last_error() -> *const c_char;
    pub fn skgpu_fnv1a64(data: *const c_void, len: usize) -> u64;
    pub fn skgpu_router_gpu_for(session_id: *const c_void, len: usize, n_gpus: u32) -> u32;
    pub fn skgpu_router_create(devices: *const i32, n_gpus: u32, cfg: *const skgpu_hub_config, out: *mut *mut skgpu_router) -> skgpu_rc;
    pub fn skgpu_router_destroy(r: *mut skgpu_router);
    pub fn skgpu_router_gpus(r: *const skgpu_router) -> u32;
    pub fn skgpu_router_hub(r: *mut skgpu_router, g: u32) -> *mut skgpu_hub;
    pub fn skgpu_router_session_open(
        r: *mut skgpu_router, session_id: *const c_void, id_len: usize, n_inputs: u32, in_rates: *const u32, handle_out: *mut skgpu_session_handle,
    ) -> skgpu_rc;
    pub fn skgpu_router_session_close(r: *mut skgpu_router, h: skgpu_session_handle) -> skgpu_rc;
    pub fn skgpu_router_push(r: *mut skgpu_router, h: skgpu_session_handle, input: u32, samples: *const c_void, n_frames: u32) -> skgpu_rc;
    pub fn skgpu_router_set_input_gain(r: *mut skgpu_router, h: skgpu_session_handle, input: u32, gain: f32) -> skgpu_rc;
    pub fn skgpu_router_set_master_gain(r: *mut skgpu_router, h: skgpu_session_handle, gain: f32) -> skgpu_rc;
    pub fn skgpu_router_tick(r: *mut skgpu_router) -> skgpu_rc;
    pub fn skgpu_router_wait(r: *mut skgpu_router) -> skgpu_rc;
    pub fn skgpu_router_session_output(
        r: *mut skgpu_router, h: skgpu_session_handle, samples: *mut *const c_void, n_mixed: *mut u32, status: *mut u32,
    ) -> skgpu_rc;
}

/// The borrowed error text of the calling thread as an owned `String` (copy it before the next call, like the host does with
/// `CResult.error_message`, crates/plugin-native/src/wrapper.rs:468-483).
pub fn last_router_error() -> String {
    unsafe {
        let p = skgpu_router_last_error();
        if p.is_null() { String::new() } else { std::ffi::CStr::from_ptr(p).to_string_lossy().into_owned() }
    }
}
