//! Registration of the GPU twins next to the built-in filters (crates/nodes/src/audio/filters/mod.rs:27-93).
//! UNCOMPILED -- see rust/README.md.
pub mod batcher;
pub mod nodes {
    pub mod gpu_chain;
    pub mod twins;
    pub mod util;
}

use std::sync::Arc;

use schemars::schema_for;
use streamkit_core::registry::{NodeRegistry, StaticPins};
use streamkit_core::{ProcessorNode, StreamKitError};

use batcher::GpuBatcher;
use nodes::gpu_chain::{GpuChainConfig, GpuChainNode};
use nodes::twins::{gpu_gain, gpu_mixer, gpu_resampler, GpuGainConfig, GpuMixerConfig, GpuResamplerConfig};

/// `kind`s added: `audio::gpu_chain` (the fused path) and the drop-in twins `audio::gpu_gain`, `audio::gpu_resampler`,
/// `audio::gpu_mixer`, which are `GpuChainNode`s of a fixed shape (one input + gain only; one resampled input; n bypass inputs).
pub fn register_gpu_nodes(registry: &mut NodeRegistry, batcher: Arc<GpuBatcher>) {
    let default_node = GpuChainNode::new(GpuChainConfig::default(), Arc::clone(&batcher)).expect("default GpuChainConfig is valid");
    let b = Arc::clone(&batcher);
    registry.register_static_with_description(
        "audio::gpu_chain",
        move |params: Option<&serde_json::Value>| {
            let config: GpuChainConfig = match params {
                Some(p) => serde_json::from_value(p.clone())
                    .map_err(|e| StreamKitError::Configuration(format!("Failed to parse audio::gpu_chain params: {e}")))?,
                None => GpuChainConfig::default(),
            };
            let node = GpuChainNode::new(config, Arc::clone(&b))
                .map_err(|e| StreamKitError::Configuration(format!("Invalid audio::gpu_chain configuration: {e}")))?;
            Ok(Box::new(node) as Box<dyn ProcessorNode>)
        },
        serde_json::to_value(schema_for!(GpuChainConfig)).expect("GpuChainConfig schema should serialize to JSON"),
        StaticPins { inputs: default_node.input_pins(), outputs: default_node.output_pins() },
        vec!["audio".to_string(), "filters".to_string(), "gpu".to_string()],
        false,
        "Resample, gain, mix, gain and s16 packing of one session on the GPU (streamkit_b200), batched with every other live session.",
    );

    // ---- the drop-in twins: the built-in nodes' config structs, pins and categories (filters/mod.rs:121-187)
    macro_rules! twin {
        ($kind:literal, $cfg:ty, $ctor:ident, $default:expr, $desc:literal) => {{
            let b = Arc::clone(&batcher);
            let probe = $ctor(&$default, Arc::clone(&batcher)).expect("default config is valid");
            registry.register_static_with_description(
                $kind,
                move |params: Option<&serde_json::Value>| {
                    let config: $cfg = match params {
                        Some(p) => serde_json::from_value(p.clone())
                            .map_err(|e| StreamKitError::Configuration(format!(concat!("Failed to parse ", $kind, " params: {}"), e)))?,
                        None => $default,
                    };
                    let node = $ctor(&config, Arc::clone(&b)).map_err(StreamKitError::Configuration)?;
                    Ok(Box::new(node) as Box<dyn ProcessorNode>)
                },
                serde_json::to_value(schema_for!($cfg)).expect("schema should serialize to JSON"),
                StaticPins { inputs: probe.input_pins(), outputs: probe.output_pins() },
                vec!["audio".to_string(), "filters".to_string(), "gpu".to_string()],
                false,
                $desc,
            );
        }};
    }
    twin!("audio::gpu_gain", GpuGainConfig, gpu_gain, GpuGainConfig::default(), "audio::gain on the GPU, batched with every other live session.");
    twin!(
        "audio::gpu_resampler",
        GpuResamplerConfig,
        gpu_resampler,
        GpuResamplerConfig { target_sample_rate: 48_000, chunk_frames: 960, output_frame_size: 960, input_sample_rate: 44_100 },
        "audio::resampler (rubato FastFixedIn / Linear arithmetic, bit-identical) on the GPU, batched with every other live session."
    );
    twin!("audio::gpu_mixer", GpuMixerConfig, gpu_mixer, GpuMixerConfig::default(), "audio::mixer (sync or clocked) on the GPU, batched with every other live session.");
}
