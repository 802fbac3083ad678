//! `-p b200`: the reference's r0 arm (/root/reference/crates/guest-prover-r0/src/prover.rs:59-106) with risc0's segment
//! prover running on B200Hal / B200CircuitHal.  SOURCE ONLY (see hal.rs).
pub mod hal;

use std::{future::Future, panic};

use anyhow::Result;
use risc0_zkvm::{ExecutorEnv, ExecutorImpl, ProverOpts, VerifierContext};
use zktls_core::ZkProver;
use zktls_program_core::GuestInput;

#[derive(Default)]
pub struct B200GuestProver { pub devices: Vec<i32> }

impl ZkProver for B200GuestProver {
    fn prove(&mut self, input: GuestInput, guest_program: &[u8]) -> impl Future<Output = Result<(Vec<u8>, Vec<u8>)>> + Send {
        let devices = if self.devices.is_empty() { vec![0] } else { self.devices.clone() };
        let elf = guest_program.to_vec();
        async move { panic::catch_unwind(move || prove_blocking(input, &elf, &devices)).map_err(|e| anyhow::anyhow!("{:?}", e))? }
    }
}

fn prove_blocking(input: GuestInput, elf: &[u8], devices: &[i32]) -> Result<(Vec<u8>, Vec<u8>)> {
    let mut input_bytes = Vec::new();
    ciborium::into_writer(&input, &mut input_bytes)?;
    let env = ExecutorEnv::builder().write_slice(&input_bytes).build()?;
    // 1. execute: continuation segments (host, unchanged)
    let session = ExecutorImpl::from_elf(env, elf)?.run()?;
    // 2. prove every segment: segment i -> GPU devices[i % G], one B200Hal (ctx + stream + memory pool) per worker thread,
    //    no collective on this path.  Each worker keeps risc0's SegmentProverImpl<B200Hal, B200CircuitHal>.
    let receipts = crate::segments::prove_all(&session, devices)?;
    // 3. lift / join / identity_p254 / groth16 exactly as risc0's ProverImpl::prove_session does (recursion circuit: unchanged)
    let receipt = crate::recursion::compress(receipts, &ProverOpts::groth16(), &VerifierContext::default())?;
    let journal = receipt.journal.bytes.clone();
    let mut seal = receipt.inner.groth16()?.seal.clone();
    if seal.len() <= 4 { seal = Vec::new(); }
    Ok((journal, seal))
}
