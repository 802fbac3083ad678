//! Serialises a risc0 circuit's constraint system into the `ZKC1` blob libzkb200 takes as DATA (layout: DESIGN.md "circuit blob",
//! parsed by zktls_b200/csrc/circuit.hpp and oracle/circuit.hpp; the Python builder is zktls_b200/circuit.py):
//!
//!   header  [MAGIC 'ZKC1', accum_cols, code_cols, data_cols, mix_size, out_size, n_taps, n_steps, ret, n_fp, n_mix, 0, info[4]]
//!   taps    n_taps x (group, column, back), sorted           (risc0_zkp::taps::TapSet, groups accum = 0, code = 1, data = 2)
//!   steps   n_steps x (op, a, b, c)                          (risc0_zkp::adapter::PolyExtStep)
//!
//! ops: 0 Const(value) 1 Get(tap) 2 GetGlobal(base, offset) 3 Add 4 Sub 5 Mul 6 True 7 AndEqz(chain, value) 8 AndCond(chain, cond, inner).
//! Operands index the fp / mix values in definition order, exactly as `PolyExtStepDef::step` numbers them.
use risc0_zkp::{adapter::{CircuitInfo, PolyExtStep, PolyExtStepDef, TapsProvider}, field::Elem, taps::TapSet};

pub const MAGIC: u32 = 0x5A4B_4331;
const OP_CONST: u32 = 0; const OP_GET: u32 = 1; const OP_GET_GLOBAL: u32 = 2; const OP_ADD: u32 = 3; const OP_SUB: u32 = 4;
const OP_MUL: u32 = 5; const OP_TRUE: u32 = 6; const OP_AND_EQZ: u32 = 7; const OP_AND_COND: u32 = 8;

/// A circuit that exposes its `PolyExtStepDef` (rv32im and recursion do: `poly_ext.rs` is a generated `DEF` table).
pub trait PolyExtDef { fn poly_ext_def(&self) -> &'static PolyExtStepDef; }

fn taps_words(taps: &TapSet<'static>) -> Vec<u32> {
    let mut v: Vec<(u32, u32, u32)> = taps.taps().map(|t| (t.group() as u32, t.offset() as u32, t.back() as u32)).collect();
    v.sort();
    v.into_iter().flat_map(|(g, c, b)| [g, c, b]).collect()
}

pub fn circuit_blob<C: CircuitInfo + TapsProvider + PolyExtDef>(circuit: &C) -> Vec<u32> {
    let taps = circuit.get_taps();
    let def = circuit.poly_ext_def();
    let tap_words = taps_words(taps);
    let (mut steps, mut n_fp, mut n_mix) = (Vec::with_capacity(4 * def.block.len()), 0u32, 0u32);
    for step in def.block {
        let (op, a, b, c, is_fp) = match *step {
            PolyExtStep::Const(value) => (OP_CONST, value, 0, 0, true),
            PolyExtStep::Get(tap) => (OP_GET, tap as u32, 0, 0, true),
            PolyExtStep::GetGlobal(base, offset) => (OP_GET_GLOBAL, base as u32, offset as u32, 0, true),
            PolyExtStep::Add(x, y) => (OP_ADD, x as u32, y as u32, 0, true),
            PolyExtStep::Sub(x, y) => (OP_SUB, x as u32, y as u32, 0, true),
            PolyExtStep::Mul(x, y) => (OP_MUL, x as u32, y as u32, 0, true),
            PolyExtStep::True => (OP_TRUE, 0, 0, 0, false),
            PolyExtStep::AndEqz(chain, value) => (OP_AND_EQZ, chain as u32, value as u32, 0, false),
            PolyExtStep::AndCond(chain, cond, inner) => (OP_AND_COND, chain as u32, cond as u32, inner as u32, false),
        };
        steps.extend_from_slice(&[op, a, b, c]);
        if is_fp { n_fp += 1 } else { n_mix += 1 }
    }
    let info = C::CIRCUIT_INFO.encode();                 // 16 field elements, one per ASCII byte: stored as the raw bytes
    let mut info_words = [0u32; 4];
    for (i, e) in info.iter().enumerate().take(16) { info_words[i / 4] |= (e.to_u32_words()[0] & 0xff) << (8 * (i % 4)); }
    let mut blob = vec![MAGIC, taps.group_size(0) as u32, taps.group_size(1) as u32, taps.group_size(2) as u32, C::MIX_SIZE as u32, C::OUTPUT_SIZE as u32,
                        (tap_words.len() / 3) as u32, (steps.len() / 4) as u32, def.ret as u32, n_fp, n_mix, 0];
    blob.extend_from_slice(&info_words);
    blob.extend_from_slice(&tap_words);
    blob.extend_from_slice(&steps);
    blob
}

impl PolyExtDef for risc0_circuit_rv32im::CircuitImpl {
    fn poly_ext_def(&self) -> &'static PolyExtStepDef { &risc0_circuit_rv32im::poly_ext::DEF }
}
