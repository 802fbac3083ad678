//! `-p b200`: the reference's r0 arm (/root/reference/crates/guest-prover-r0/src/prover.rs:59-106) with risc0's segment prover
//! running on B200Hal / B200CircuitHal over libzkb200.
//!
//! SOURCE ONLY: the image this repository is built in has no Rust toolchain and the risc0 crates are not vendored, so none of
//! this is compiled here.  It is kept self-consistent (every `crate::` path resolves to a module below; tests/test_abi_symbols.py
//! checks that, and that zkb200-sys is regenerated from include/zkb200.h).  risc0 API names follow risc0-zkvm / risc0-zkp /
//! risc0-circuit-rv32im 1.2.5 as pinned in /root/reference/Cargo.lock:4961-5128; they are recalled, not compiled against.
pub mod blob;
pub mod hal;
pub mod recursion;
pub mod segments;

use std::{future::Future, panic};

use anyhow::Result;
use risc0_zkvm::{ExecutorEnv, ExecutorImpl, ProverOpts, VerifierContext};
use zktls_core::ZkProver;
use zktls_program_core::GuestInput;

/// Same shape as `Risc0GuestProver` (/root/reference/crates/guest-prover-r0/src/prover.rs:32-57); there is no mock / network
/// mode: this backend only exists to prove locally on B200s.
#[derive(Default)]
pub struct B200GuestProver {
    /// CUDA device ordinals; empty = device 0.  Segment i is proven on `devices[i % devices.len()]`.
    pub devices: Vec<i32>,
    /// segments in flight per device (one `B200Hal` = one ctx / stream / memory pool each); 0 = 3
    pub inflight: usize,
}

impl B200GuestProver {
    pub fn devices(mut self, devices: Vec<i32>) -> Self { self.devices = devices; self }
}

impl ZkProver for B200GuestProver {
    fn prove(&mut self, input: GuestInput, guest_program: &[u8]) -> impl Future<Output = Result<(Vec<u8>, Vec<u8>)>> + Send {
        let devices = if self.devices.is_empty() { vec![0] } else { self.devices.clone() };
        let inflight = if self.inflight == 0 { 3 } else { self.inflight };
        let elf = guest_program.to_vec();
        // the trait returns anyhow errors and traps panics, like the r0 arm (prover.rs:70-76); B200Hal panics on a C-ABI error string
        async move { panic::catch_unwind(move || prove_blocking(input, &elf, &devices, inflight)).map_err(|e| anyhow::anyhow!("{:?}", e))? }
    }
}

fn prove_blocking(input: GuestInput, elf: &[u8], devices: &[i32], inflight: usize) -> Result<(Vec<u8>, Vec<u8>)> {
    let mut input_bytes = Vec::new();
    ciborium::into_writer(&input, &mut input_bytes)?;                       // prover.rs:81-82
    let env = ExecutorEnv::builder().write_slice(&input_bytes).build()?;    // prover.rs:86
    // 1. execute: continuation segments (host, unchanged)
    let session = ExecutorImpl::from_elf(env, elf)?.run()?;
    // 2. prove every segment: segment i -> devices[i % G], `inflight` worker threads per device, each with its own B200Hal;
    //    no collective on this path (the Python mirror is zktls_b200/shard.py + bench.py's SegmentQueue)
    let receipts = crate::segments::prove_all(&session, devices, inflight)?;
    // 3. lift / join / identity_p254 / groth16 exactly as risc0's ProverImpl::compress does (recursion circuit: unchanged)
    let receipt = crate::recursion::compress(&session, receipts, &ProverOpts::groth16(), &VerifierContext::default())?;
    let journal = receipt.journal.bytes.clone();                            // prover.rs:95
    let mut seal = receipt.inner.groth16()?.seal.clone();                   // prover.rs:96
    if seal.len() <= 4 { seal = Vec::new(); }                               // prover.rs:101-103
    Ok((journal, seal))
}
