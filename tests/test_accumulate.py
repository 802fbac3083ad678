"""CircuitHal::accumulate on the device (zkb_accumulate, csrc/k_accum.cu) against the host witness generator
(zktls_b200/synth.py trace_b_accum, the numpy statement of the SYN family's accumulation step), and as the middle step of a
segment proof that never brings the accum trace to the host."""
import numpy as np
import pytest

from zktls_b200 import circuit, synth

pytestmark = pytest.mark.gpu

SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)
MID = dict(accum_cols=7, code_cols=5, data_cols=33, mix_size=20, out_size=32)
ONE = dict(accum_cols=1, code_cols=2, data_cols=1, mix_size=1, out_size=1)


@pytest.fixture(scope="module")
def hal():
    from zktls_b200.hal import B200Hal
    h = B200Hal(0)
    yield h
    h.close()


@pytest.mark.parametrize("shape,po2", [(ONE, 7), (SMALL, 6), (SMALL, 11), (MID, 12), (circuit.SYN280, 10)])
def test_accumulate_matches_host_witness_generator(hal, shape, po2):
    blob = circuit.syn_circuit(**shape).blob()
    io, code, data = synth.trace_b_code_data(shape, po2, seed=3 + po2)
    mix = synth.encode(synth.splitmix_fp(99 + po2, shape["mix_size"])).astype(np.uint32)      # any Montgomery words: the op does not care where mix came from
    want = synth.to_mont(synth.trace_b_accum(shape, po2, 3 + po2, code, data, io, mix))
    # the device op overwrites live rows only: start from the same noise the host generator starts from
    n = 1 << po2
    noise = synth.to_mont(synth.splitmix_fp((3 + po2) * 4 + 3, shape["accum_cols"] * n).reshape(shape["accum_cols"], n))
    accum = hal.copy_from_elem(noise)
    hal.accumulate(blob, accum, hal.copy_from_elem(synth.to_mont(code)), hal.copy_from_elem(synth.to_mont(data)), mix, io, po2)
    got = accum.to_numpy()
    assert np.array_equal(got, want)
    live = code[0].astype(bool)
    assert np.array_equal(got.reshape(-1, n)[:, ~live], noise.reshape(-1, n)[:, ~live]), "rows outside the selector must be left alone"


def _grand_product_circuit():
    """a witness program the SYN family does not have: an unconditional Set from taps at back 0 / 2, a barrier, a second phase
    that reads the first phase's column at back 1, and a PrefixProduct over four accum columns (a grand-product column)"""
    from zktls_b200.circuit import CircuitBuilder, GROUP_ACCUM, GROUP_CODE, GROUP_DATA, GLOBAL_MIX, GLOBAL_OUT
    b = CircuitBuilder(6, 2, 3, 2, 2, info=b"GRANDPROD:v1____")
    for g, n in ((GROUP_ACCUM, 6), (GROUP_CODE, 2), (GROUP_DATA, 3)):
        for c in range(n):
            b.add_tap(g, c, 0)
    b.finish_taps()
    b.ret = b.and_eqz(b.true(), b.sub(b.get(GROUP_ACCUM, 0, 0), b.get(GROUP_ACCUM, 0, 0)))
    m0, m1 = b.w_get_global(GLOBAL_MIX, 0), b.w_get_global(GLOBAL_MIX, 1)
    for k in range(4):                                   # columns 0..3 = (m0 + d0[i] * d1[i-2] + k, ...): the factors of the grand product
        v = b.w_add(b.w_add(m0, b.w_mul(b.w_get(GROUP_DATA, 0, 0), b.w_get(GROUP_DATA, 1, 2))), b.w_const(k + 1))
        b.w_set(k, b.w_mul(v, m1) if k == 3 else v)
    b.w_barrier()
    prev = b.w_get(GROUP_ACCUM, 1, 1)                    # phase 2 reads what phase 1 wrote, one row back
    b.w_set(4, b.w_sub(prev, b.w_get_global(GLOBAL_OUT, 1)), b.w_get(GROUP_CODE, 0, 0))
    b.w_prefix_product(0)                                # columns 0..3 become the running product over the rows
    b.w_set(5, b.w_add(b.w_get(GROUP_ACCUM, 0, 0), b.w_get(GROUP_ACCUM, 3, 1)))          # phase 3 reads the scanned columns
    return b


@pytest.mark.parametrize("po2", [5, 9, 13])
def test_accumulate_program_with_prefix_product_matches_oracle(hal, oracle, po2):
    b = _grand_product_circuit(); blob = b.blob()
    rng = np.random.default_rng(po2)
    n = 1 << po2
    accum, code, data = (oracle.random_fp(rng, c * n) for c in b.group_size)
    code[:n] = oracle.encode(rng.integers(0, 2, size=n))      # the condition column: zeros and ones
    mix, io = oracle.random_fp(rng, 2), oracle.random_fp(rng, 2)
    want = oracle.accumulate(blob, accum, code, data, mix, io, po2)
    d_accum = hal.copy_from_elem(accum)
    hal.accumulate(blob, d_accum, hal.copy_from_elem(code), hal.copy_from_elem(data), mix, io, po2)
    assert np.array_equal(d_accum.to_numpy(), want)


def test_accumulate_is_rejected_for_a_circuit_without_a_witness_program(hal):
    from zktls_b200._lib import ZkbError
    b = circuit.syn_circuit(**SMALL)
    b.wsteps = []                       # a blob without a witness program (e.g. exported from a circuit whose accumulate stays on the host)
    blob = b.blob()
    buf = hal.alloc_elem(4 << 6)
    with pytest.raises(ZkbError, match="no witness program"):
        hal.accumulate(blob, buf, buf, buf, np.zeros(5, np.uint32), np.zeros(4, np.uint32), 6)


@pytest.mark.parametrize("shape,po2", [(SMALL, 9), (MID, 12)])
def test_segment_proof_with_device_side_accumulate(hal, oracle, shape, po2):
    """begin() -> accumulate on the device -> finish(device buffer): the seal equals the oracle's (host accum) and verifies."""
    from zktls_b200.prover import SegmentProver, verify_segment, control_id
    blob = circuit.syn_circuit(**shape).blob()
    io, code, data = synth.trace_b_code_data(shape, po2, seed=21)
    code_m, data_m = synth.to_mont(code), synth.to_mont(data)
    d_code, d_data = hal.copy_from_elem(code_m), hal.copy_from_elem(data_m)
    gp = SegmentProver(hal, blob)
    mix = gp.begin(po2, io, d_code, d_data)
    n = 1 << po2
    d_accum = hal.copy_from_elem(synth.to_mont(synth.splitmix_fp(21 * 4 + 3, shape["accum_cols"] * n).reshape(shape["accum_cols"], n)))
    hal.accumulate(blob, d_accum, d_code, d_data, mix, io, po2)
    seal = gp.finish(d_accum)
    op = oracle.Prover(blob)
    assert np.array_equal(op.begin(po2, io, code_m, data_m), mix)
    seal_o = op.finish(synth.to_mont(synth.trace_b_accum(shape, po2, 21, code, data, io, mix)))
    assert np.array_equal(seal, seal_o)
    verify_segment(blob, seal, control_id(po2, op.roots()[0]))
    gp.close()


@pytest.mark.parametrize("shape,po2", [(SMALL, 9), (MID, 11)])
def test_staged_proof_computes_the_accum_group_on_the_device(hal, oracle, shape, po2):
    """zkb_prover_stage_traces with a NULL accum trace: zkb_prove_staged runs the witness program between the data commit and the accum
    commit (prove_segment's order) over a zeroed group; seal = oracle prover fed with the oracle's accumulate of the same inputs;
    two segments through the two staging slots give the same seal; a circuit without a witness program is refused."""
    from zktls_b200.prover import SegmentProver, verify_segment, control_id
    from zktls_b200._lib import ZkbError
    blob = circuit.syn_circuit(**shape).blob()
    io, code, data = synth.trace_b_code_data(shape, po2, seed=33)
    code_m, data_m = synth.to_mont(code), synth.to_mont(data)
    gp = SegmentProver(hal, blob)
    gp.stage(po2, code_m, data_m); gp.stage(po2, code_m, data_m)
    seal1 = gp.prove_staged(io); seal2 = gp.prove_staged(io)
    op = oracle.Prover(blob)
    mix = op.begin(po2, io, code_m, data_m)
    n = 1 << po2
    accum_o = oracle.accumulate(blob, np.zeros(shape["accum_cols"] * n, np.uint32), code_m.reshape(-1), data_m.reshape(-1), mix, io, po2)
    seal_o = op.finish(accum_o)
    assert np.array_equal(seal1, seal_o) and np.array_equal(seal2, seal_o)
    verify_segment(blob, seal1, control_id(po2, op.roots()[0]))
    gp.close()
    b = circuit.syn_circuit(**shape); b.wsteps = []
    gq = SegmentProver(hal, b.blob())
    with pytest.raises(ZkbError, match="no witness program"):
        gq.stage(po2, code_m, data_m)
    gq.close()
