//! Small adapters used by gpu_chain.rs. UNCOMPILED -- see rust/README.md.
use streamkit_core::types::Packet;
use tokio::sync::mpsc;

use crate::batcher::SessionHandle;

/// `Self::recv_from_any` of the built-in mixer (mixer.rs:1080-1148): the first input that has a packet, round robin.
pub async fn recv_any(inputs: &mut [mpsc::Receiver<Packet>]) -> Option<(usize, Packet)> {
    use std::future::poll_fn;
    use std::task::Poll;
    poll_fn(|cx| {
        let mut open = false;
        for (i, rx) in inputs.iter_mut().enumerate() {
            match rx.poll_recv(cx) {
                Poll::Ready(Some(p)) => return Poll::Ready(Some((i, p))),
                Poll::Ready(None) => {}
                Poll::Pending => open = true,
            }
        }
        if open { Poll::Pending } else { Poll::Ready(None) }
    })
    .await
}

pub struct SessionParts<'a> {
    inner: &'a mut SessionHandle,
}
pub fn session_parts(s: &mut SessionHandle) -> SessionParts<'_> {
    SessionParts { inner: s }
}
impl SessionParts<'_> {
    pub fn push(&self, input: u32, frame: &streamkit_core::types::AudioFrame) -> Result<(), streamkit_core::StreamKitError> {
        self.inner.push(input, frame)
    }
    pub fn set_input_gain(&self, input: u32, gain: f32) -> Result<(), String> {
        self.inner.set_input_gain(input, gain)
    }
    pub fn set_master_gain(&self, gain: f32) -> Result<(), String> {
        self.inner.set_master_gain(gain)
    }
    pub async fn recv_output(&mut self) -> Option<Packet> {
        self.inner.output.recv().await
    }
}
