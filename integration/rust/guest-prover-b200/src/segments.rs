//! Segment-parallel proving: the reference's hot loop #0 (risc0-zkvm `ProverImpl::prove_session`: `for segment in session.segments
//! { prove_segment(..) }`; reached from /root/reference/crates/guest-prover-r0/src/prover.rs:90) spread over the GPUs of one box.
//! Segments are independent, so there is no collective: a shared queue, `inflight` worker threads per device, one B200Hal
//! (ctx + stream + memory pool) per worker -- a Hal and its buffers are !Send, so each is created on the thread that uses it.
use std::sync::{atomic::{AtomicUsize, Ordering}, Mutex};

use anyhow::{anyhow, Result};
use risc0_circuit_rv32im::{prove::{SegmentProver, SegmentProverImpl}, CircuitImpl};
use risc0_zkp::hal::CircuitHal;
use risc0_zkvm::{Segment, SegmentReceipt, Session};

use crate::{blob::circuit_blob, hal::{B200CircuitHal, B200Hal}};

/// One worker: risc0's own segment prover, generic over the Hal pair (risc0-circuit-rv32im `prove::SegmentProverImpl<H, C>`).
struct Worker { prover: SegmentProverImpl<B200Hal, B200CircuitHal> }

impl Worker {
    fn new(device: i32, blob: &[u32]) -> Self {
        let hal = std::rc::Rc::new(B200Hal::new(device));
        let circuit_hal = std::rc::Rc::new(B200CircuitHal::new(&hal, blob.to_vec()));
        Self { prover: SegmentProverImpl::new(hal, circuit_hal) }
    }
    /// `ProverImpl::prove_segment`: seal from the circuit prover, claim decoded from the seal's io words, integrity check.
    fn prove(&self, segment: &Segment) -> Result<SegmentReceipt> {
        let seal = self.prover.prove_segment(&segment.inner)?;
        let receipt = risc0_zkvm::receipt::segment::segment_receipt_from_seal(segment, seal, "poseidon2")?;
        receipt.verify_integrity_with_context(&risc0_zkvm::VerifierContext::default())?;
        Ok(receipt)
    }
}

/// Proves every segment of `session`; returns the receipts in segment order.
pub fn prove_all(session: &Session, devices: &[i32], inflight: usize) -> Result<Vec<SegmentReceipt>> {
    let blob = circuit_blob(&CircuitImpl::new());           // TapSet + PolyExtStep program of rv32im, serialised once (blob.rs)
    let n = session.segments.len();
    let next = AtomicUsize::new(0);
    let out: Vec<Mutex<Option<SegmentReceipt>>> = (0..n).map(|_| Mutex::new(None)).collect();
    let failed: Mutex<Option<anyhow::Error>> = Mutex::new(None);
    std::thread::scope(|scope| {
        for &device in devices {
            for _ in 0..inflight.max(1) {
                let (blob, next, out, failed) = (&blob, &next, &out, &failed);
                scope.spawn(move || {
                    let worker = Worker::new(device, blob);
                    loop {
                        let i = next.fetch_add(1, Ordering::Relaxed);          // dynamic queue: the tail is one segment, not one round
                        if i >= n || failed.lock().unwrap().is_some() { break; }
                        let result = session.segments[i].resolve().and_then(|segment| worker.prove(&segment));
                        match result {
                            Ok(r) => *out[i].lock().unwrap() = Some(r),
                            Err(e) => { *failed.lock().unwrap() = Some(e); break; }
                        }
                    }
                });
            }
        }
    });
    if let Some(e) = failed.into_inner().unwrap() { return Err(e); }
    out.into_iter().enumerate().map(|(i, m)| m.into_inner().unwrap().ok_or_else(|| anyhow!("segment {i} was not proven"))).collect()
}

// keep the trait in scope for readers: B200CircuitHal is the CircuitHal<B200Hal> SegmentProverImpl is instantiated with
#[allow(unused)] fn _assert_circuit_hal<C: CircuitHal<B200Hal>>(_: &C) {}

/// Device ordinals 0..N-1 for the N GPUs CUDA exposes to this process (CUDA_VISIBLE_DEVICES already applied by the driver).
pub fn visible_devices() -> Vec<i32> {
    let mut n = 0i32;
    crate::hal::ok(unsafe { zkb200_sys::zkb_device_count(&mut n) });
    (0..n.max(1)).collect()
}
