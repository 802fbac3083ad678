"""zkb_verify_segment (risc0-zkp verify/* restated): seals of valid traces VERIFY, anything else is rejected.

CPU leg: the verifier against seals made by the CPU oracle prover (independent code: the oracle shares nothing with
libzkb200).  GPU leg: seals made by the B200 prover verify too."""
import numpy as np
import pytest

from zktls_b200 import circuit, synth, ZkbError
from zktls_b200.prover import verify_segment, control_id, seal_code_root

SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)
MID = dict(accum_cols=7, code_cols=5, data_cols=33, mix_size=20, out_size=32)


def oracle_seal(oracle, shape, po2, seed, valid=True):
    blob = circuit.syn_circuit(**shape).blob()
    pr = oracle.Prover(blob)
    if valid:
        io, code, data = synth.trace_b_code_data(shape, po2, seed)
        code_m, data_m = synth.to_mont(code), synth.to_mont(data)
        mix = pr.begin(po2, io, code_m, data_m)
        accum_m = synth.to_mont(synth.trace_b_accum(shape, po2, seed, code, data, io, mix))
    else:
        io, code_m, data_m, accum_m = synth.trace_a(shape, po2, seed)
        pr.begin(po2, io, code_m, data_m)
    seal = pr.finish(accum_m)
    return blob, seal, control_id(po2, pr.roots()[0])


@pytest.mark.parametrize("shape,po2", [(SMALL, 8), (SMALL, 9), (MID, 10), (SMALL, 13)])
def test_valid_oracle_seal_verifies(oracle, shape, po2):
    blob, seal, cid = oracle_seal(oracle, shape, po2, seed=3 + po2)
    got_po2, got_root = verify_segment(blob, seal, cid)
    assert got_po2 == po2 and np.array_equal(got_root, cid[1:])


def test_unsatisfied_constraints_are_rejected(oracle):
    blob, seal, cid = oracle_seal(oracle, SMALL, 8, seed=4, valid=False)
    with pytest.raises(ZkbError, match="constraint polynomial"):
        verify_segment(blob, seal, cid)


def test_every_region_of_the_seal_is_bound(oracle):
    blob, seal, cid = oracle_seal(oracle, SMALL, 8, seed=5)
    verify_segment(blob, seal, cid)
    rng = np.random.default_rng(1)
    # flip one word at positions spread over the whole seal: header, top layers, coeff_u, FRI commitments, final
    # coefficients and query answers -- every single-word change must be caught
    positions = sorted(set([0, 3, 4, 5, 40, seal.size - 1, seal.size - 9] + list(rng.integers(0, seal.size, 60))))
    for pos in positions:
        bad = seal.copy()
        bad[pos] = (int(bad[pos]) + 1) % 2013265921
        with pytest.raises(ZkbError, match="invalid proof|out of range"):
            verify_segment(blob, bad, cid)
    with pytest.raises(ZkbError, match="too short"):
        verify_segment(blob, seal[:-1], cid)
    with pytest.raises(ZkbError, match="trailing"):
        verify_segment(blob, np.concatenate([seal, np.zeros(1, np.uint32)]), cid)
    # a seal for one circuit does not verify under another circuit's description
    other = circuit.syn_circuit(**dict(SMALL, data_cols=7)).blob()
    with pytest.raises(ZkbError):
        verify_segment(other, seal, cid)


def test_golden_seal_verifies():
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
    blob = circuit.syn_circuit(**SMALL).blob()
    cid = control_id(int(G["seg_valid_po2"][0]), G["seg_valid_roots"][:8])       # roots in commit order: code first
    verify_segment(blob, G["seg_valid_seal"], cid)
    with pytest.raises(ZkbError):
        verify_segment(blob, G["seg_random_seal"], control_id(int(G["seg_random_po2"][0]), G["seg_random_roots"][:8]))


def test_code_root_is_bound_to_the_control_id(oracle):
    """the code group is the PROGRAM.  A prover that picks its own code trace -- here an all-zero selector column,
    which makes every SYN constraint vanish -- can produce an internally consistent seal for ANY io; it must not verify against
    the genuine program's control ID, and a verifier call that neither checks nor returns the code root is refused."""
    shape, po2, seed = SMALL, 8, 6
    blob, seal, cid = oracle_seal(oracle, shape, po2, seed)
    # the forgery: sel = 0 everywhere, arbitrary data / accum / io
    io, code, data = synth.trace_b_code_data(shape, po2, seed)
    code[0][:] = 0
    forged_io = synth.encode(np.arange(7, 7 + shape["out_size"], dtype=np.uint64)).astype(np.uint32)
    pr = oracle.Prover(blob)
    pr.begin(po2, forged_io, synth.to_mont(code), synth.to_mont(data))
    _, _, _, junk = synth.trace_a(shape, po2, 99)
    forged = pr.finish(junk)
    forged_cid = control_id(po2, pr.roots()[0])
    assert not np.array_equal(forged_cid, cid)
    verify_segment(blob, forged, forged_cid)                 # consistent with ITS OWN code root: the constraints do vanish ...
    with pytest.raises(ZkbError, match="control ID"):
        verify_segment(blob, forged, cid)                    # ... but it is not the program the verifier asked about
    with pytest.raises(ZkbError, match="control ID"):
        verify_segment(blob, seal, control_id(po2 + 1, cid[1:]))     # right root, wrong po2
    verify_segment(blob, seal, np.concatenate([forged_cid, control_id(po2 + 1, cid[1:]), cid]))   # table with several entries
    assert seal_code_root(blob, seal)[0] == po2 and np.array_equal(seal_code_root(blob, seal)[1], cid[1:])
    # the C entry point refuses to run with neither a table nor an output buffer
    import ctypes as C
    from zktls_b200._lib import lib, check
    from zktls_b200.hal import _hp, _sz
    b = np.ascontiguousarray(blob, np.uint32); s_ = np.ascontiguousarray(seal, np.uint32)
    with pytest.raises(ZkbError, match="control IDs"):
        check(lib().zkb_verify_segment(_hp(b), _sz(b.size), _hp(s_), _sz(s_.size), None, _sz(0), None))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,po2", [(SMALL, 10), (MID, 12), (MID, 15)])
def test_gpu_seal_verifies(shape, po2):
    from zktls_b200.hal import B200Hal
    from zktls_b200.prover import SegmentProver
    hal = B200Hal(0)
    blob = circuit.syn_circuit(**shape).blob()
    gp = SegmentProver(hal, blob)
    io, code, data = synth.trace_b_code_data(shape, po2, 77)
    code_m, data_m = synth.to_mont(code), synth.to_mont(data)
    mix = gp.begin(po2, io, code_m, data_m)
    seal = gp.finish(synth.to_mont(synth.trace_b_accum(shape, po2, 77, code, data, io, mix)))
    cid = control_id(po2, gp.roots()[0])
    verify_segment(blob, seal, cid)
    bad = seal.copy(); bad[seal.size // 2] ^= 1
    with pytest.raises(ZkbError):
        verify_segment(blob, bad, cid)
    gp.close(); hal.close()


def _fib_circuit():
    """A satisfiable circuit OUTSIDE the SYN family: registers of three taps (back 0, 1, 2) and a gated linear recurrence
         sel[i] * (d_j[i] - d_j[i-1] - k[i] * d_j[i-2]) == 0,   acc[i] = mix_0 * d_0[i] + d_1[i-2] on live rows
    so that the verifier's per-register interpolation (three points), the three divisions of the {0,1,2} combo and the DEEP relation
    are exercised by a seal that must VERIFY, not merely equal another prover's."""
    from zktls_b200.circuit import CircuitBuilder, GROUP_ACCUM, GROUP_CODE, GROUP_DATA, GLOBAL_MIX
    b = CircuitBuilder(1, 2, 3, 1, 1, info=b"FIB3:v1_________")
    b.add_tap(GROUP_ACCUM, 0, 0)
    for c in range(2):
        b.add_tap(GROUP_CODE, c, 0)
    for c in range(3):
        for k in (0, 1, 2):
            b.add_tap(GROUP_DATA, c, k)
    b.finish_taps()
    sel, kk = b.get(GROUP_CODE, 0, 0), b.get(GROUP_CODE, 1, 0)
    top = b.and_eqz(b.true(), b.mul(sel, b.sub(sel, b.const(1))))
    inner = b.true()
    for j in range(3):
        d0, d1, d2 = (b.get(GROUP_DATA, j, k) for k in (0, 1, 2))
        inner = b.and_eqz(inner, b.sub(b.sub(d0, d1), b.mul(kk, d2)))
    m0 = b.get_global(GLOBAL_MIX, 0)
    inner = b.and_eqz(inner, b.sub(b.sub(b.get(GROUP_ACCUM, 0, 0), b.mul(m0, b.get(GROUP_DATA, 0, 0))), b.get(GROUP_DATA, 1, 2)))
    b.ret = b.and_cond(top, sel, inner)
    # the witness program of the accum column (device-side accumulate)
    v = b.w_add(b.w_mul(b.w_get_global(GLOBAL_MIX, 0), b.w_get(GROUP_DATA, 0, 0)), b.w_get(GROUP_DATA, 1, 2))
    b.w_set(0, v, b.w_get(GROUP_CODE, 0, 0))
    return b


def _fib_trace(po2, seed):
    from zktls_b200.circuit import ZK_ROWS
    P = 2013265921
    n = 1 << po2
    rng = np.random.default_rng(seed)
    code = rng.integers(0, P, size=(2, n), dtype=np.uint64)
    sel = np.zeros(n, np.uint64); sel[2: n - ZK_ROWS] = 1
    code[0] = sel
    data = rng.integers(0, P, size=(3, n), dtype=np.uint64)
    for i in range(2, n - ZK_ROWS):
        data[:, i] = (data[:, i - 1] + code[1, i] * data[:, i - 2] % P) % P
    io = synth.encode(rng.integers(0, P, size=1, dtype=np.uint64)).astype(np.uint32)
    return io, code, data


def _fib_accum(po2, seed, code, data, mix_mont):
    P = 2013265921
    n = 1 << po2
    m0 = int(synth.decode(mix_mont)[0])
    acc = np.random.default_rng(seed + 1).integers(0, P, size=(1, n), dtype=np.uint64)
    live = code[0].astype(bool)
    v = (m0 * data[0] % P + np.roll(data[1], 2)) % P
    acc[0][live] = v[live]
    return acc


@pytest.mark.parametrize("po2", [8, 10])
def test_three_tap_registers_prove_and_verify_on_the_oracle(oracle, po2):
    b = _fib_circuit(); blob = b.blob()
    io, code, data = _fib_trace(po2, 40 + po2)
    code_m, data_m = synth.to_mont(code), synth.to_mont(data)
    pr = oracle.Prover(blob)
    mix = pr.begin(po2, io, code_m, data_m)
    acc = _fib_accum(po2, 40 + po2, code, data, mix)
    assert np.array_equal(oracle.accumulate(blob, synth.to_mont(np.random.default_rng(41 + po2).integers(0, 2013265921, size=(1, 1 << po2), dtype=np.uint64)),
                                            code_m, data_m, mix, io, po2), synth.to_mont(acc))          # the witness program computes the same column
    seal = pr.finish(synth.to_mont(acc))
    cid = control_id(po2, pr.roots()[0])
    verify_segment(blob, seal, cid)
    # one broken recurrence row -> the DEEP relation fails
    bad = data.copy(); bad[1, 77] = (bad[1, 77] + 1) % 2013265921
    pr2 = oracle.Prover(blob)
    pr2.begin(po2, io, code_m, synth.to_mont(bad))
    with pytest.raises(ZkbError, match="constraint polynomial"):
        verify_segment(blob, pr2.finish(synth.to_mont(acc)), cid)


@pytest.mark.gpu
def test_three_tap_registers_prove_and_verify_on_the_gpu(oracle):
    from zktls_b200.hal import B200Hal
    from zktls_b200.prover import SegmentProver
    hal = B200Hal(0)
    b = _fib_circuit(); blob = b.blob(); po2 = 11
    io, code, data = _fib_trace(po2, 7)
    code_m, data_m = synth.to_mont(code), synth.to_mont(data)
    gp = SegmentProver(hal, blob)
    d_code, d_data = hal.copy_from_elem(code_m), hal.copy_from_elem(data_m)
    mix = gp.begin(po2, io, d_code, d_data)
    d_acc = hal.copy_from_elem(synth.to_mont(np.random.default_rng(8).integers(0, 2013265921, size=(1, 1 << po2), dtype=np.uint64)))
    hal.accumulate(blob, d_acc, d_code, d_data, mix, io, po2)           # accum never visits the host
    assert np.array_equal(d_acc.to_numpy(), synth.to_mont(_fib_accum(po2, 7, code, data, mix)))
    seal = gp.finish(d_acc)
    verify_segment(blob, seal, control_id(po2, gp.roots()[0]))
    op = oracle.Prover(blob)
    op.begin(po2, io, code_m, data_m)
    assert np.array_equal(seal, op.finish(d_acc.to_numpy()))
    gp.close(); hal.close()
