//! lift / join / identity_p254 / Groth16: the stage after the segment proofs.  It runs risc0's recursion prover unchanged
//! (`ProverServer::compress`, risc0-zkvm 1.2.5 `host/server/prove/mod.rs`): the recursion circuit calls the same Hal operators, so
//! it can be moved onto B200Hal the same way once its `poly_ext` is exported as a blob (needs the recursion
//! circuit's control program `recursion_zkr.zip` and a poseidon_254 hash kernel for identity_p254; neither is in this repository).
use anyhow::Result;
use risc0_zkvm::{get_prover_server, CompositeReceipt, InnerReceipt, ProverOpts, Receipt, SegmentReceipt, Session, VerifierContext};

/// Composite receipt from the segment receipts (+ the session's assumptions, none for the TLS guest), then compress to `opts.receipt_kind`.
pub fn compress(session: &Session, segments: Vec<SegmentReceipt>, opts: &ProverOpts, ctx: &VerifierContext) -> Result<Receipt> {
    let composite = CompositeReceipt { segments, assumption_receipts: Vec::new(), verifier_parameters: ctx.composite_verifier_parameters().map(|p| p.digest()).unwrap_or_default() };
    let journal = session.journal.clone().unwrap_or_default().bytes;
    let receipt = Receipt::new(InnerReceipt::Composite(composite), journal);
    receipt.verify_integrity_with_context(ctx)?;
    // succinct: lift every segment, join pairwise (a binary tree over the same work queue, still no collective); groth16 adds
    // identity_p254 + the SNARK wrapper (Docker, as in the reference arm)
    get_prover_server(opts)?.compress(opts, &receipt)
}
