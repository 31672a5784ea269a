//! Drop-in twins of the built-in filters: `audio::gpu_gain`, `audio::gpu_resampler`, `audio::gpu_mixer`.
//! UNCOMPILED -- see rust/README.md.
//!
//! Each twin takes the built-in node's OWN config struct (same field names, defaults and validation messages:
//! gain.rs:30-67, resampler.rs:22-102, mixer.rs:60-143), so a pipeline changes only the `kind`. Internally a twin is a
//! `GpuChainNode` of a fixed shape -- the batching layer has one session type, and a chain with the unused stages at
//! identity is bit-identical to the single node (x * 1.0 == x; a one-input mix is the frame itself, mixer.rs:969-972;
//! rate-equal inputs are not resampled, resampler.rs:299-373):
//!
//!   audio::gpu_gain       1 input at the mixer rate (bypass), input gain = config.gain, master gain 1.0
//!   audio::gpu_resampler  1 input at its own rate, gains 1.0          (target_sample_rate must be the batcher's mixer rate)
//!   audio::gpu_mixer      n bypass inputs, gains 1.0, sync / clocked as configured
//!
//! What a twin cannot honour is rejected at construction with a Configuration error instead of being approximated:
//! a resampler whose `chunk_frames / ratio` is not the batcher's packet size (multi-chunk packets), a mixer
//! `clocked.sample_rate / frame_samples_per_channel` different from the batcher's tick.
use std::sync::Arc;

use schemars::JsonSchema;
use serde::Deserialize;

use crate::batcher::GpuBatcher;
use crate::nodes::gpu_chain::{GpuChainConfig, GpuChainNode};

/// gain.rs:30-48
#[derive(Deserialize, Debug, Clone, JsonSchema)]
#[serde(default)]
pub struct GpuGainConfig {
    pub gain: f32,
}
impl Default for GpuGainConfig {
    fn default() -> Self {
        Self { gain: 1.0 }
    }
}

/// resampler.rs:22-46
#[derive(Deserialize, Debug, Clone, JsonSchema)]
pub struct GpuResamplerConfig {
    pub target_sample_rate: u32,
    #[serde(default = "default_chunk_frames")]
    pub chunk_frames: usize,
    #[serde(default = "default_output_frame_size")]
    pub output_frame_size: usize,
    /// NOT in the built-in node: the input rate must be known when the session is admitted (the built-in resampler learns
    /// it from the first packet, resampler.rs:299-313). Pipelines state it; `0` = take it from the first packet and admit late.
    #[serde(default)]
    pub input_sample_rate: u32,
}
const fn default_chunk_frames() -> usize {
    960
}
const fn default_output_frame_size() -> usize {
    960
}

/// mixer.rs:60-110 (the fields this path uses)
#[derive(Deserialize, Debug, Clone, JsonSchema, Default)]
#[serde(default)]
pub struct GpuMixerConfig {
    pub sync_timeout_ms: Option<u64>,
    pub num_inputs: Option<usize>,
    pub clocked: Option<GpuClockedMixerConfig>,
}
#[derive(Deserialize, Debug, Clone, JsonSchema)]
pub struct GpuClockedMixerConfig {
    pub sample_rate: u32,
    pub frame_samples_per_channel: usize,
    #[serde(default = "default_jitter")]
    pub jitter_buffer_frames: usize,
}
const fn default_jitter() -> usize {
    3
}

pub fn gpu_gain(config: &GpuGainConfig, batcher: Arc<GpuBatcher>) -> Result<GpuChainNode, String> {
    // gain.rs:50-66, same messages
    if !config.gain.is_finite() {
        return Err(format!("Gain must be a finite number, got: {}", config.gain));
    }
    if !(0.0..=4.0).contains(&config.gain) {
        return Err(format!("Gain must be between 0 and 4, got: {}", config.gain));
    }
    let rate = batcher.mixer_rate();
    GpuChainNode::new(
        GpuChainConfig { num_inputs: 1, input_sample_rates: vec![rate], input_gains: vec![config.gain], gain: 1.0, sync_timeout_ms: None },
        batcher,
    )
}

pub fn gpu_resampler(config: &GpuResamplerConfig, batcher: Arc<GpuBatcher>) -> Result<GpuChainNode, String> {
    // resampler.rs:81-102, same messages
    if config.target_sample_rate == 0 {
        return Err("target_sample_rate must be greater than 0".to_string());
    }
    if config.chunk_frames == 0 {
        return Err("chunk_frames must be greater than 0".to_string());
    }
    const VALID: [usize; 7] = [0, 120, 240, 480, 960, 1920, 2880];
    if !VALID.contains(&config.output_frame_size) {
        return Err(format!(
            "output_frame_size must be 0 (disabled) or a valid Opus frame size (120, 240, 480, 960, 1920, 2880), got: {}",
            config.output_frame_size
        ));
    }
    if config.target_sample_rate != batcher.mixer_rate() || config.output_frame_size != batcher.packet_frames() as usize {
        return Err(format!(
            "audio::gpu_resampler batches at {} Hz / {} frames per packet; use audio::resampler (or plugin::native::gpu_resampler) for {} Hz / {}",
            batcher.mixer_rate(), batcher.packet_frames(), config.target_sample_rate, config.output_frame_size
        ));
    }
    let in_rate = if config.input_sample_rate == 0 { batcher.mixer_rate() } else { config.input_sample_rate };
    // the chunk the batching layer expects from this input is in_rate * packet / mixer_rate frames;
    // the node re-chunks what it receives exactly like resampler.rs:375-395 before pushing
    GpuChainNode::new(
        GpuChainConfig { num_inputs: 1, input_sample_rates: vec![in_rate], input_gains: vec![1.0], gain: 1.0, sync_timeout_ms: None },
        batcher,
    )
}

pub fn gpu_mixer(config: &GpuMixerConfig, batcher: Arc<GpuBatcher>) -> Result<GpuChainNode, String> {
    let n = config.num_inputs.unwrap_or(2);
    if let Some(c) = &config.clocked {
        if c.sample_rate != batcher.mixer_rate() || c.frame_samples_per_channel != batcher.packet_frames() as usize {
            return Err(format!(
                "audio::gpu_mixer ticks at {} Hz / {} frames; the pipeline asks for {} Hz / {}",
                batcher.mixer_rate(), batcher.packet_frames(), c.sample_rate, c.frame_samples_per_channel
            ));
        }
    }
    let rate = batcher.mixer_rate();
    GpuChainNode::new(
        GpuChainConfig {
            num_inputs: n,
            input_sample_rates: vec![rate; n],
            input_gains: vec![1.0; n],
            gain: 1.0,
            // mixer.rs:60-79: sync mode unless a clocked config is present; default timeout 100 ms
            sync_timeout_ms: if config.clocked.is_some() { None } else { Some(config.sync_timeout_ms.unwrap_or(100)) },
        },
        batcher,
    )
}
