//! The engine half of the frame-batching layer (BASELINE.json north star: "crates/engine gains a frame-batching layer that
//! gathers each tick's 10-20 ms frames from all live sessions into pinned host buffers and sends them to the device through a
//! thin C-ABI FFI"). UNCOMPILED -- see rust/README.md.
//!
//! In the reference every node of every session is its own tokio task with its own channels
//! (crates/engine/src/dynamic_actor.rs:393-495), one DynamicEngine actor per session (apps/skit/src/session.rs:173-200);
//! nothing is shared across sessions. `GpuBatcher` is the one process-wide object the GPU nodes of ALL sessions talk to:
//!   * it owns the multi-GPU router (`skgpu_router`, one hub + one NUMA-pinned tick thread per GPU; sessions are routed by
//!     fnv1a64(session id) % n_gpus, the reference's own session hash, session.rs:35-45);
//!   * node tasks `push` their frames (thread-safe, copied straight into the pinned arena of the next tick);
//!   * one tick task fires every 20 ms, submits all GPUs' ticks, and hands every session's mixed packet to the node task that
//!     registered for it (a `tokio::sync::mpsc` per session, like any other pin).
use std::collections::HashMap;
use std::ffi::c_void;
use std::sync::{Arc, Mutex};
use std::time::Duration;

use streamkit_core::types::{AudioFrame, Packet, PacketMetadata};
use streamkit_core::StreamKitError;
use streamkit_gpu_sys as sys;
use tokio::sync::mpsc;

/// What `audio::gpu_chain` (and its single-purpose twins) ask the batcher for: one mixer session.
pub struct SessionSpec {
    pub session_id: String,        // NodeContext.session_id (node.rs:215): decides the GPU
    pub input_rates: Vec<u32>,     // one per input pin; == out_rate -> bypass input (resampler.rs:299-373)
    pub sync_timeout_ms: Option<u64>, // Some(..) -> audio::mixer sync mode (mixer.rs:554-918); None + clocked -> clocked mode
    pub clocked: bool,
}

pub struct SessionHandle {
    handle: sys::skgpu_session_handle,
    batcher: Arc<GpuBatcher>,
    /// the session's mixed packets, one per tick in which something was mixed
    pub output: mpsc::Receiver<Packet>,
}

struct Registered {
    out_tx: mpsc::Sender<Packet>,
    sequence: u64,
}

pub struct GpuBatcher {
    router: *mut sys::skgpu_router,
    out_rate: u32,
    out_frames: u32,
    channels: u16,
    sessions: Mutex<HashMap<sys::skgpu_session_handle, Registered>>,
}

// The C layer is internally synchronised for the calls used from node tasks (push / gains: skgpu_hub.h "Threading").
unsafe impl Send for GpuBatcher {}
unsafe impl Sync for GpuBatcher {}

fn rt_err(what: &str) -> StreamKitError {
    StreamKitError::Runtime(format!("{what}: {}", sys::last_router_error()))
}

impl GpuBatcher {
    /// the rate every session is mixed at (the hubs' `out_rate`)
    pub fn mixer_rate(&self) -> u32 {
        self.out_rate
    }
    /// frames per packet and tick (the hubs' `out_frames`)
    pub fn packet_frames(&self) -> u32 {
        self.out_frames
    }

    /// `devices`: CUDA ordinals of the box; capacities are per GPU. Fails (Configuration) when there is no GPU: there is no CPU
    /// fallback -- a deployment without GPUs keeps the built-in `audio::*` nodes.
    pub fn new(devices: &[i32], max_sessions_per_gpu: u32, max_inputs_per_session: u32, in_rates: &[u32]) -> Result<Arc<Self>, StreamKitError> {
        let cfg = sys::skgpu_hub_config {
            max_sessions: max_sessions_per_gpu,
            max_streams: max_sessions_per_gpu * max_inputs_per_session,
            max_inputs_per_session,
            out_rate: 48_000,
            out_frames: 960, // ClockedMixerConfig defaults (mixer.rs:46-55): 48 kHz, 960 frames = 20 ms
            channels: 2,
            flags: sys::SKGPU_HUB_OUT_S16,
            in_rates: in_rates.as_ptr(),
            n_in_rates: in_rates.len() as u32,
            jitter_frames: 3, // jitter_buffer_frames default (mixer.rs:52)
            slices: 16,
        };
        let mut router = std::ptr::null_mut();
        let rc = unsafe { sys::skgpu_router_create(devices.as_ptr(), devices.len() as u32, &cfg, &mut router) };
        if rc != sys::SKGPU_OK {
            return Err(StreamKitError::Configuration(format!("GPU batcher: {}", sys::last_router_error())));
        }
        Ok(Arc::new(Self { router, out_rate: 48_000, out_frames: 960, channels: 2, sessions: Mutex::new(HashMap::new()) }))
    }

    pub fn open_session(self: &Arc<Self>, spec: &SessionSpec) -> Result<SessionHandle, StreamKitError> {
        let mut handle = 0u64;
        let id = spec.session_id.as_bytes();
        let rc = unsafe {
            sys::skgpu_router_session_open(self.router, id.as_ptr() as *const c_void, id.len(), spec.input_rates.len() as u32, spec.input_rates.as_ptr(), &mut handle)
        };
        if rc != sys::SKGPU_OK {
            return Err(StreamKitError::Configuration(format!("GPU session: {}", sys::last_router_error())));
        }
        let (out_tx, output) = mpsc::channel(8);
        self.sessions.lock().unwrap().insert(handle, Registered { out_tx, sequence: 0 });
        Ok(SessionHandle { handle, batcher: Arc::clone(self), output })
    }

    /// One 20 ms tick on every GPU, then delivery. Spawned once by the engine next to its session actors:
    /// `tokio::spawn(batcher.clone().run_ticks())`.
    pub async fn run_ticks(self: Arc<Self>) {
        let mut interval = tokio::time::interval(Duration::from_millis(20));
        interval.set_missed_tick_behavior(tokio::time::MissedTickBehavior::Skip); // like the clocked mixer's ticker (mixer.rs:1290)
        loop {
            interval.tick().await;
            let this = Arc::clone(&self);
            // the FFI calls block for ~1 ms of submit work: keep them off the async workers, like the plugin wrapper does
            // (crates/plugin-native/src/wrapper.rs:398-457 uses spawn_blocking for every process_packet)
            let delivered = tokio::task::spawn_blocking(move || this.tick_blocking()).await;
            if let Ok(Err(e)) = delivered {
                tracing::error!("GPU batcher tick failed: {e}"); // a CUDA error: every GPU session's node reports Failed
                return;
            }
        }
    }

    fn tick_blocking(&self) -> Result<(), StreamKitError> {
        unsafe {
            if sys::skgpu_router_tick(self.router) != sys::SKGPU_OK {
                return Err(rt_err("skgpu_router_tick"));
            }
            if sys::skgpu_router_wait(self.router) != sys::SKGPU_OK {
                return Err(rt_err("skgpu_router_wait"));
            }
        }
        let n = (self.out_frames as usize) * (self.channels as usize);
        let mut sessions = self.sessions.lock().unwrap();
        for (handle, reg) in sessions.iter_mut() {
            let (mut ptr, mut n_mixed, mut status) = (std::ptr::null::<c_void>(), 0u32, 0u32);
            let rc = unsafe { sys::skgpu_router_session_output(self.router, *handle, &mut ptr, &mut n_mixed, &mut status) };
            if rc != sys::SKGPU_OK || ptr.is_null() || n_mixed == 0 {
                continue; // nothing mixed this tick (sync mode holding, or a session opened after the tick was submitted)
            }
            // s16 -> the engine's f32 AudioFrame for downstream CPU nodes (an Opus encoder takes i16 directly: see INTEGRATION.md)
            let s16 = unsafe { std::slice::from_raw_parts(ptr as *const i16, n) };
            let samples: Vec<f32> = s16.iter().map(|s| f32::from(*s) * (1.0 / 32768.0)).collect();
            let metadata = PacketMetadata {
                timestamp_us: None,
                duration_us: Some(u64::from(self.out_frames) * 1_000_000 / u64::from(self.out_rate)), // mixer.rs:1413-1418
                sequence: Some(reg.sequence),
            };
            reg.sequence += 1;
            let frame = AudioFrame::with_metadata(self.out_rate, self.channels, samples, Some(metadata));
            let _ = reg.out_tx.try_send(Packet::Audio(frame)); // a slow consumer drops mixes, like OutputMailbox (mixer.rs:1150-1183)
        }
        Ok(())
    }
}

impl SessionHandle {
    /// One input frame of pin `input` (interleaved f32, exactly the pin's 20 ms chunk). Callable from the node task.
    pub fn push(&self, input: u32, frame: &AudioFrame) -> Result<(), StreamKitError> {
        let n_frames = (frame.samples.len() / frame.channels as usize) as u32;
        let rc = unsafe { sys::skgpu_router_push(self.batcher.router, self.handle, input, frame.samples.as_ptr() as *const c_void, n_frames) };
        if rc == sys::SKGPU_OK { Ok(()) } else { Err(rt_err("skgpu_router_push")) }
    }
    pub fn set_input_gain(&self, input: u32, gain: f32) -> Result<(), String> {
        let rc = unsafe { sys::skgpu_router_set_input_gain(self.batcher.router, self.handle, input, gain) };
        if rc == sys::SKGPU_OK { Ok(()) } else { Err(sys::last_router_error()) } // rejected like gain.rs:157-171: the old gain stays
    }
    pub fn set_master_gain(&self, gain: f32) -> Result<(), String> {
        let rc = unsafe { sys::skgpu_router_set_master_gain(self.batcher.router, self.handle, gain) };
        if rc == sys::SKGPU_OK { Ok(()) } else { Err(sys::last_router_error()) }
    }
}

impl Drop for SessionHandle {
    fn drop(&mut self) {
        self.batcher.sessions.lock().unwrap().remove(&self.handle);
        unsafe { sys::skgpu_router_session_close(self.batcher.router, self.handle) };
    }
}

impl Drop for GpuBatcher {
    fn drop(&mut self) {
        unsafe { sys::skgpu_router_destroy(self.router) };
    }
}
