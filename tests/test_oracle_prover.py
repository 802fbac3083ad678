"""CPU-only checks of the oracle prover, the circuit blob format and the synthetic traces (no GPU needed)."""
import numpy as np
import pytest

from zktls_b200 import circuit, synth

SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)


def lde(oracle, trace_mont, cols, po2):
    co = oracle.zk_shift(oracle.batch_interpolate_ntt(trace_mont, cols, po2), cols, po2)
    return oracle.batch_expand_into_evaluate_ntt(co, cols, po2, 2)


def test_trace_b_satisfies_constraints_check_poly_is_low_degree(oracle):
    """For a valid witness the check polynomial Q = C / Z has degree < 3n, so its top n coefficients vanish;
    for a random trace (Trace A) they do not.  Exercises eval_check + the circuit + the trace generator together."""
    po2 = 7; n = 1 << po2
    blob = circuit.syn_circuit(**SMALL).blob()
    rng = np.random.default_rng(5)
    mix = oracle.random_fp(rng, SMALL["mix_size"]); poly_mix = oracle.random_fp(rng, 4)
    io, code, data = synth.trace_b_code_data(SMALL, po2, seed=3)
    accum = synth.trace_b_accum(SMALL, po2, 3, code, data, io, mix)
    ev = [lde(oracle, synth.to_mont(m), m.shape[0], po2) for m in (accum, code, data)]
    check = oracle.eval_check(blob, ev[0], ev[1], ev[2], mix, io, poly_mix, po2)
    co = oracle.batch_bit_reverse(oracle.batch_interpolate_ntt(check, 4, po2 + 2), 4, po2 + 2).reshape(4, 4 * n)
    assert not co[:, 3 * n:].any() and co[:, : 3 * n].any()
    # Trace A: constraints violated -> full-degree quotient
    io, c2, d2, a2 = synth.trace_a(SMALL, po2, seed=4)
    ev = [lde(oracle, m, cols, po2) for m, cols in ((a2, 4), (c2, 3), (d2, 6))]
    check = oracle.eval_check(blob, ev[0], ev[1], ev[2], mix, io, poly_mix, po2)
    co = oracle.batch_bit_reverse(oracle.batch_interpolate_ntt(check, 4, po2 + 2), 4, po2 + 2).reshape(4, 4 * n)
    assert co[:, 3 * n:].any()


@pytest.mark.parametrize("po2,n_roots", [(8, 4), (10, 5)])
def test_oracle_prover_runs_and_is_deterministic(oracle, po2, n_roots):
    blob = circuit.syn_circuit(**SMALL).blob()
    seals = []
    for _ in range(2):
        io, code, data = synth.trace_b_code_data(SMALL, po2, seed=1)
        pr = oracle.Prover(blob)
        mix = pr.begin(po2, io, synth.to_mont(code), synth.to_mont(data))
        accum = synth.trace_b_accum(SMALL, po2, 1, code, data, io, mix)
        seals.append(pr.finish(synth.to_mont(accum)))
        assert pr.roots().shape == (n_roots, 8)
    assert np.array_equal(seals[0], seals[1]) and seals[0].size > 1000
    assert (seals[0] < 2013265921).all()        # every seal word is a canonical field element / digest word


def test_circuit_blob_rejects_malformed(oracle):
    blob = circuit.syn_circuit(**SMALL).blob()
    bad = blob.copy(); bad[0] ^= 1
    with pytest.raises(RuntimeError):
        oracle.Prover(bad)
    with pytest.raises(RuntimeError):
        oracle.Prover(blob[:-1])
    bad = blob.copy(); bad[16 + 3] = 0; bad[16 + 4] = 0; bad[16 + 5] = 0      # duplicate tap -> not strictly sorted
    with pytest.raises(RuntimeError):
        oracle.Prover(bad)


def test_splitmix_is_uniform_and_deterministic():
    a = synth.splitmix_fp(7, 10000); b = synth.splitmix_fp(7, 10000)
    assert np.array_equal(a, b) and a.max() < synth.P and abs(a.mean() / synth.P - 0.5) < 0.02


def test_oracle_accumulate_runs_the_witness_program(oracle):
    """CircuitHal::accumulate as data: the oracle's interpreter of the blob's witness program against the numpy statement of the
    SYN family's accumulation step (zktls_b200/synth.py), and the parser's phase rule."""
    import numpy as np
    from zktls_b200 import circuit, synth
    shape = dict(accum_cols=5, code_cols=3, data_cols=6, mix_size=5, out_size=4)
    blob = circuit.syn_circuit(**shape).blob()
    po2, seed = 8, 3
    n = 1 << po2
    io, code, data = synth.trace_b_code_data(shape, po2, seed)
    mix = synth.encode(synth.splitmix_fp(7, shape["mix_size"])).astype(np.uint32)
    noise = synth.to_mont(synth.splitmix_fp(seed * 4 + 3, shape["accum_cols"] * n).reshape(shape["accum_cols"], n))
    want = synth.to_mont(synth.trace_b_accum(shape, po2, seed, code, data, io, mix))
    got = oracle.accumulate(blob, noise, synth.to_mont(code), synth.to_mont(data), mix, io, po2)
    assert np.array_equal(got, want)
    # the C-ABI parser refuses a phase that reads an accum column it writes (rows run in parallel on the device)
    import ctypes as C
    from zktls_b200 import lib
    from zktls_b200._lib import check, ZkbError
    import pytest
    b = circuit.syn_circuit(**shape)
    b.wsteps = [s for s in b.wsteps if s[0] != circuit.W_BARRIER]
    bad = b.blob()
    need = C.c_size_t()
    with pytest.raises(ZkbError, match="reads an accum column it writes|writes an accum column it reads"):
        check(lib().zkb_eval_check_source(bad.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(bad.size), None, C.c_size_t(0), C.byref(need)))
