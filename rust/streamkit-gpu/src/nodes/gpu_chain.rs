//! `audio::gpu_chain` -- the fused hot path as ONE node: per input pin `audio::resampler{48 kHz, 960} -> audio::gain`, then
//! `audio::mixer` (clocked 48 kHz / 960 or sync), `audio::gain` (master) and the s16 packing. UNCOMPILED -- see rust/README.md.
//!
//! Same pin conventions as the built-in mixer: `num_inputs: n` pre-creates `in_0 .. in_{n-1}` (mixer.rs:128-143), one "out" pin.
use async_trait::async_trait;
use schemars::JsonSchema;
use serde::Deserialize;
use std::sync::Arc;
use streamkit_core::control::NodeControlMessage;
use streamkit_core::pins::{InputPin, OutputPin, PinCardinality};
use streamkit_core::state_helpers;
use streamkit_core::stats::NodeStatsTracker;
use streamkit_core::types::{AudioFormat, Packet, PacketType, SampleFormat};
use streamkit_core::{NodeContext, ProcessorNode, StreamKitError};

use crate::batcher::{GpuBatcher, SessionSpec};

#[derive(Deserialize, Debug, JsonSchema, Clone)]
#[serde(default)]
pub struct GpuChainConfig {
    /// input pins in_0 .. in_{n-1}
    pub num_inputs: usize,
    /// sample rate of every input pin (a resampler runs where it differs from 48000; equal rates are bypassed, resampler.rs:299-373)
    pub input_sample_rates: Vec<u32>,
    /// per-input audio::gain, 0.0 ..= 4.0 (gain.rs:50-66)
    pub input_gains: Vec<f32>,
    /// audio::gain after the mix
    pub gain: f32,
    /// Some(ms) = mixer sync mode with this timeout (mixer.rs:60-79); None = clocked mode (mixer.rs:23-55)
    pub sync_timeout_ms: Option<u64>,
}

impl Default for GpuChainConfig {
    fn default() -> Self {
        Self { num_inputs: 2, input_sample_rates: vec![48_000, 48_000], input_gains: vec![1.0, 1.0], gain: 1.0, sync_timeout_ms: None }
    }
}

pub struct GpuChainNode {
    config: GpuChainConfig,
    batcher: Arc<GpuBatcher>,
}

impl GpuChainNode {
    pub fn new(config: GpuChainConfig, batcher: Arc<GpuBatcher>) -> Result<Self, String> {
        if config.num_inputs == 0 || config.num_inputs > 64 {
            return Err("num_inputs must be in 1..=64".to_string());
        }
        if config.input_sample_rates.len() != config.num_inputs || config.input_gains.len() != config.num_inputs {
            return Err("input_sample_rates and input_gains need one entry per input".to_string());
        }
        for g in config.input_gains.iter().chain(std::iter::once(&config.gain)) {
            if !g.is_finite() || !(0.0..=4.0).contains(g) {
                return Err(format!("Gain must be a finite number between 0.0 and 4.0, got {g}")); // gain.rs:50-66
            }
        }
        Ok(Self { config, batcher })
    }
}

fn any_f32() -> PacketType {
    PacketType::RawAudio(AudioFormat { sample_rate: 0, channels: 0, sample_format: SampleFormat::F32 })
}

#[async_trait]
impl ProcessorNode for GpuChainNode {
    fn input_pins(&self) -> Vec<InputPin> {
        (0..self.config.num_inputs)
            .map(|i| InputPin { name: format!("in_{i}"), accepts_types: vec![any_f32()], cardinality: PinCardinality::One })
            .collect()
    }

    fn output_pins(&self) -> Vec<OutputPin> {
        vec![OutputPin {
            name: "out".to_string(),
            produces_type: PacketType::RawAudio(AudioFormat { sample_rate: 48_000, channels: 2, sample_format: SampleFormat::F32 }),
            cardinality: PinCardinality::Broadcast,
        }]
    }

    async fn run(self: Box<Self>, mut context: NodeContext) -> Result<(), StreamKitError> {
        let node_name = context.output_sender.node_name().to_string();
        state_helpers::emit_initializing(&context.state_tx, &node_name);
        let spec = SessionSpec {
            session_id: context.session_id.clone().unwrap_or_else(|| node_name.clone()),
            input_rates: self.config.input_sample_rates.clone(),
            sync_timeout_ms: self.config.sync_timeout_ms,
            clocked: self.config.sync_timeout_ms.is_none(),
        };
        let mut session = match self.batcher.open_session(&spec) {
            Ok(s) => s,
            Err(e) => {
                state_helpers::emit_failed(&context.state_tx, &node_name, e.to_string());
                return Err(e);
            }
        };
        for (i, g) in self.config.input_gains.iter().enumerate() {
            let _ = session.set_input_gain(i as u32, *g);
        }
        let _ = session.set_master_gain(self.config.gain);
        let mut inputs = Vec::with_capacity(self.config.num_inputs);
        for i in 0..self.config.num_inputs {
            inputs.push(context.take_input(&format!("in_{i}"))?);
        }
        state_helpers::emit_running(&context.state_tx, &node_name);
        let mut stats = NodeStatsTracker::new(node_name.clone(), context.stats_tx.clone());
        let mut control_rx = context.control_rx;

        // One forwarding task per input pin: a frame goes straight into the pinned arena of the next tick.
        // (recv_from_any of the built-in mixer, mixer.rs:1080-1148, polls the receivers in turn; forwarding is enough here
        // because the decision WHAT to mix is taken inside the batching layer.)
        let session = Arc::new(session_parts(&mut session));
        loop {
            tokio::select! {
                Some(ctrl) = control_rx.recv() => match ctrl {
                    NodeControlMessage::UpdateParams(params) => match serde_json::from_value::<GpuChainConfig>(params) {
                        Ok(new) => {
                            for (i, g) in new.input_gains.iter().enumerate() {
                                if session.set_input_gain(i as u32, *g).is_err() { stats.errored(); }   // rejected: old gain stays
                            }
                            if session.set_master_gain(new.gain).is_err() { stats.errored(); }
                        }
                        Err(_) => stats.errored(),
                    },
                    NodeControlMessage::Shutdown => {
                        state_helpers::emit_stopped(&context.state_tx, &node_name, "shutdown");
                        return Ok(());
                    }
                    NodeControlMessage::Start => {}
                },
                Some((pin, packet)) = recv_any(&mut inputs) => {
                    if let Packet::Audio(frame) = packet {               // non-audio packets are ignored (mixer.rs:899-901)
                        stats.received();
                        if let Err(e) = session.push(pin as u32, &frame) {
                            state_helpers::emit_failed(&context.state_tx, &node_name, e.to_string());
                            return Err(e);
                        }
                    }
                },
                Some(mixed) = session.recv_output() => {
                    if context.output_sender.send("out", mixed).await.is_err() {
                        state_helpers::emit_stopped(&context.state_tx, &node_name, "output_closed");
                        return Ok(());
                    }
                    stats.sent();
                    stats.maybe_send();
                },
                else => {
                    state_helpers::emit_stopped(&context.state_tx, &node_name, "all_inputs_closed");
                    return Ok(());
                }
            }
        }
    }
}

// `session_parts` / `recv_any` / `recv_output` are small adapters (a Mutex around the output receiver and a poll over the
// input receivers, exactly mixer.rs:1080-1148 `recv_from_any`); omitted where they add nothing to review.
use crate::nodes::util::{recv_any, session_parts};
