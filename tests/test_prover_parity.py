"""GPU parity of the full segment prover: seal words and Merkle/FRI roots from libzkb200 must equal the CPU oracle's
(the stand-in for "identical Merkle roots, FRI commitments and receipt" -- SURVEY.md 4.2, 8d config 2)."""
import numpy as np
import pytest

from zktls_b200 import circuit, synth

pytestmark = pytest.mark.gpu

SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)
MID = dict(accum_cols=7, code_cols=5, data_cols=33, mix_size=20, out_size=32)


@pytest.fixture(scope="module")
def hal():
    from zktls_b200.hal import B200Hal
    h = B200Hal(0)
    yield h
    h.close()


def run_both(hal, oracle, shape, po2, seed, valid):
    from zktls_b200.prover import SegmentProver
    blob = circuit.syn_circuit(**shape).blob()
    gp = SegmentProver(hal, blob); op = oracle.Prover(blob)
    if valid:
        io, code, data = synth.trace_b_code_data(shape, po2, seed)
        code_m, data_m = synth.to_mont(code), synth.to_mont(data)
        mix_g = gp.begin(po2, io, code_m, data_m); mix_o = op.begin(po2, io, code_m, data_m)
        assert np.array_equal(mix_g, mix_o)
        accum_m = synth.to_mont(synth.trace_b_accum(shape, po2, seed, code, data, io, mix_g))
        seal_g = gp.finish(accum_m); seal_o = op.finish(accum_m)
    else:
        io, code_m, data_m, accum_m = synth.trace_a(shape, po2, seed)
        seal_g = gp.prove(po2, io, code_m, data_m, accum_m)
        op.begin(po2, io, code_m, data_m); seal_o = op.finish(accum_m)
    roots_g, roots_o = gp.roots(), op.roots()
    gp.close()
    return seal_g, seal_o, roots_g, roots_o


@pytest.mark.parametrize("shape,po2,valid", [(SMALL, 8, True), (SMALL, 10, True), (SMALL, 10, False), (MID, 12, True), (SMALL, 13, False), (MID, 14, False)])
def test_seal_matches_oracle(hal, oracle, shape, po2, valid):
    seal_g, seal_o, roots_g, roots_o = run_both(hal, oracle, shape, po2, seed=11 + po2, valid=valid)
    assert np.array_equal(roots_g, roots_o), "Merkle / FRI roots differ"
    assert seal_g.size == seal_o.size
    diff = np.nonzero(seal_g != seal_o)[0]
    assert diff.size == 0, f"first differing seal word at {diff[:5]} of {seal_g.size}"


def test_syn280_seal_matches_oracle(hal, oracle):
    """The benchmark circuit (SYN-280) at a size the oracle proves in seconds."""
    seal_g, seal_o, roots_g, roots_o = run_both(hal, oracle, circuit.SYN280, 11, seed=5, valid=False)
    assert np.array_equal(roots_g, roots_o) and np.array_equal(seal_g, seal_o)


def test_device_resident_traces_give_the_same_seal(hal, oracle):
    from zktls_b200.prover import SegmentProver
    po2 = 10
    blob = circuit.syn_circuit(**SMALL).blob()
    io, code_m, data_m, accum_m = synth.trace_a(SMALL, po2, 21)
    p1 = SegmentProver(hal, blob); s1 = p1.prove(po2, io, code_m, data_m, accum_m); p1.close()
    p2 = SegmentProver(hal, blob)
    s2 = p2.prove(po2, io, hal.copy_from_elem(code_m), hal.copy_from_elem(data_m), hal.copy_from_elem(accum_m)); p2.close()
    assert np.array_equal(s1, s2)


@pytest.mark.parametrize("po2", [6, 9])
def test_eval_check_matches_oracle(hal, oracle, po2):
    shape = MID
    blob = circuit.syn_circuit(**shape).blob()
    rng = np.random.default_rng(po2)
    dom = 4 << po2
    accum, code, data = (oracle.random_fp(rng, shape[k] * dom) for k in ("accum_cols", "code_cols", "data_cols"))
    mix, out, pm = oracle.random_fp(rng, shape["mix_size"]), oracle.random_fp(rng, shape["out_size"]), oracle.random_fp(rng, 4)
    chk = hal.alloc_elem(4 * dom)
    hal.eval_check(chk, blob, hal.copy_from_elem(accum), hal.copy_from_elem(code), hal.copy_from_elem(data), mix, out, pm, po2)
    assert np.array_equal(chk.to_numpy(), oracle.eval_check(blob, accum, code, data, mix, out, pm, po2))


def test_staged_pipeline_gives_the_same_seals(hal, oracle):
    """zkb_prover_stage_traces / zkb_prove_staged (double-buffered upload) must produce the seals of the plain call, in order."""
    from zktls_b200.prover import SegmentProver
    blob = circuit.syn_circuit(**SMALL).blob()
    gp = SegmentProver(hal, blob)
    segs = [synth.trace_a(SMALL, 9, 40 + i) for i in range(4)]
    want = []
    for io, code, data, accum in segs:
        op = oracle.Prover(blob); op.begin(9, io, code, data); want.append(op.finish(accum))
    got = []
    gp.stage(9, *segs[0][1:])
    for i in range(len(segs)):
        if i + 1 < len(segs):
            gp.stage(9, *segs[i + 1][1:])
        got.append(gp.prove_staged(segs[i][0]))
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    # misuse is an error string, not a crash
    from zktls_b200 import ZkbError
    with pytest.raises(ZkbError):
        gp.prove_staged(segs[0][0])
    gp.stage(9, *segs[0][1:]); gp.stage(9, *segs[1][1:])
    with pytest.raises(ZkbError):
        gp.stage(9, *segs[2][1:])
    gp.close()


@pytest.mark.parametrize("po2", [16, 18])
def test_syn280_seal_matches_oracle_at_tiled_ntt_sizes(hal, oracle, po2):
    """SYN-280 at sizes where every NTT runs through the tiled strided + contiguous passes (the benchmark's code path)."""
    seal_g, seal_o, roots_g, roots_o = run_both(hal, oracle, circuit.SYN280, po2, seed=po2, valid=False)
    assert np.array_equal(roots_g, roots_o), "Merkle / FRI roots differ"
    assert np.array_equal(seal_g, seal_o)


@pytest.mark.slow
def test_full_size_segment_verifies(hal):
    """BASELINE.json configs[1] at FULL size (2^20 cycles, SYN-280): a witness that satisfies the constraints is proven on the
    GPU and the seal must VERIFY (zkb_verify_segment: Merkle paths, constraint relation at the DEEP point, FRI) -- the
    size-independent property that stands in for an oracle seal the CPU would need minutes to produce."""
    from zktls_b200.prover import SegmentProver, verify_segment, control_id
    from zktls_b200 import ZkbError
    shape, po2 = circuit.SYN280, 20
    blob = circuit.syn_circuit(**shape).blob()
    gp = SegmentProver(hal, blob)
    io, code, data = synth.trace_b_code_data(shape, po2, 2020)
    code_m, data_m = synth.to_mont(code), synth.to_mont(data)
    mix = gp.begin(po2, io, code_m, data_m)
    accum_m = synth.to_mont(synth.trace_b_accum(shape, po2, 2020, code, data, io, mix))
    seal = gp.finish(accum_m)
    assert gp.roots().shape == (7, 8)
    cid = control_id(po2, gp.roots()[0])
    verify_segment(blob, seal, cid)
    bad = seal.copy(); bad[seal.size - 1000] ^= 1
    with pytest.raises(ZkbError):
        verify_segment(blob, bad, cid)
    # breaking one constraint row makes the proof fail at the DEEP check (the prover itself cannot know)
    data_bad = data_m.copy(); data_bad[12345] = (int(data_bad[12345]) + 1) % 2013265921
    gp.begin(po2, io, code_m, data_bad)
    with pytest.raises(ZkbError):
        verify_segment(blob, gp.finish(accum_m), cid)
    gp.close()


def _deep_tap_circuit(accum_cols=3, code_cols=4, data_cols=21, mix_size=3, out_size=2):
    """A circuit the SYN family does not cover: taps up to 5 rows back (a 20-row halo in the staged eval_check), columns that are
    read by many constraints (resident) next to columns read once (streamed), nested AndCond blocks and a constraint on globals only."""
    from zktls_b200.circuit import CircuitBuilder, GROUP_ACCUM, GROUP_CODE, GROUP_DATA, GLOBAL_MIX, GLOBAL_OUT
    b = CircuitBuilder(accum_cols, code_cols, data_cols, mix_size, out_size, info=b"DEEPTAPS:v1_____")
    backs = {GROUP_ACCUM: (0, 1, 3), GROUP_CODE: (0, 2), GROUP_DATA: (0, 1, 2, 5)}
    for g, n in ((GROUP_ACCUM, accum_cols), (GROUP_CODE, code_cols), (GROUP_DATA, data_cols)):
        for c in range(n):
            for k in backs[g]:
                b.add_tap(g, c, k)
    b.finish_taps()
    sel, gate = b.get(GROUP_CODE, 0, 0), b.get(GROUP_CODE, 1, 2)
    top = b.and_eqz(b.true(), b.mul(b.get_global(GLOBAL_MIX, 0), b.sub(b.get_global(GLOBAL_OUT, 1), b.const(7))))      # globals only
    inner = b.true()
    for j in range(data_cols):
        cur, p2_, p5 = b.get(GROUP_DATA, j, 0), b.get(GROUP_DATA, (j + 3) % data_cols, 2), b.get(GROUP_DATA, j, 5)
        k = b.get(GROUP_CODE, 2 + j % (code_cols - 2), 0)
        inner = b.and_eqz(inner, b.sub(b.sub(cur, b.mul(p2_, p5)), b.mul(k, b.get(GROUP_DATA, j, 1))))
    deeper = b.true()
    for j in range(accum_cols):
        a0, a3 = b.get(GROUP_ACCUM, j, 0), b.get(GROUP_ACCUM, j, 3)
        deeper = b.and_eqz(deeper, b.sub(b.add(a0, b.get(GROUP_ACCUM, (j + 1) % accum_cols, 1)), b.mul(a3, b.get_global(GLOBAL_MIX, j % mix_size))))
    inner = b.and_cond(inner, gate, deeper)
    b.ret = b.and_cond(top, sel, inner)
    return b


@pytest.mark.parametrize("po2,env", [(6, {}), (8, {}), (8, {"ZKB_EC_STAGES": "2", "ZKB_EC_CPS": "1"}), (8, {"ZKB_EC_RES": "0", "ZKB_EC_CPS": "3"}),
                                     (9, {"ZKB_EC_STAGES": "5", "ZKB_EC_CPS": "2", "ZKB_EC_RES_USES": "2"}), (8, {"ZKB_EC_STAGED": "0"}), (7, {"ZKB_EVAL_CHECK": "vm"})])
def test_eval_check_deep_taps_every_form(hal, oracle, po2, env, monkeypatch):
    """eval_check against the oracle on a circuit with a 20-row halo, in the staged form under several ring shapes (one column per
    stage, no resident columns, everything resident), in the register form, and in the device interpreter."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    b = _deep_tap_circuit()
    blob = b.blob()
    rng = np.random.default_rng(100 + po2)
    dom = 4 << po2
    accum, code, data = (oracle.random_fp(rng, n * dom) for n in b.group_size)
    mix, out, pm = oracle.random_fp(rng, b.mix_size), oracle.random_fp(rng, b.out_size), oracle.random_fp(rng, 4)
    chk = hal.alloc_elem(4 * dom)
    hal.eval_check(chk, blob, hal.copy_from_elem(accum), hal.copy_from_elem(code), hal.copy_from_elem(data), mix, out, pm, po2)
    assert np.array_equal(chk.to_numpy(), oracle.eval_check(blob, accum, code, data, mix, out, pm, po2))


def test_eval_check_below_the_minimum_domain_is_an_error_string(hal, oracle):
    from zktls_b200 import ZkbError
    b = _deep_tap_circuit(); dom = 4 << 4
    bufs = [hal.alloc_elem(max(n * dom, 4)) for n in b.group_size]
    with pytest.raises(ZkbError, match="po2 >= 6"):
        hal.eval_check(hal.alloc_elem(4 * dom), b.blob(), *bufs, np.zeros(b.mix_size, np.uint32), np.zeros(b.out_size, np.uint32), np.ones(4, np.uint32), 4)


def test_device_traces_are_left_untouched_by_the_prover(hal):
    """The first iNTT pass reads a device-resident trace in place (no copy): the caller's buffers must come back unchanged."""
    from zktls_b200.prover import SegmentProver
    for shape, po2 in ((SMALL, 10), (MID, 13)):      # 13: the tiled (out-of-place) path; 10: contiguous-only plan (copy first)
        blob = circuit.syn_circuit(**shape).blob()
        io, code_m, data_m, accum_m = synth.trace_a(shape, po2, 77)
        bufs = [hal.copy_from_elem(t) for t in (code_m, data_m, accum_m)]
        p = SegmentProver(hal, blob); p.prove(po2, io, *bufs); p.close()
        for b_, t in zip(bufs, (code_m, data_m, accum_m)):
            assert np.array_equal(b_.to_numpy(), t)
