"""zkb_verify_segment (risc0-zkp verify/* restated): seals of valid traces VERIFY, anything else is rejected.

CPU leg: the verifier against seals made by the CPU oracle prover (independent code: the oracle shares nothing with
libzkb200).  GPU leg: seals made by the B200 prover verify too."""
import numpy as np
import pytest

from zktls_b200 import circuit, synth, ZkbError
from zktls_b200.prover import verify_segment

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
    return blob, pr.finish(accum_m)


@pytest.mark.parametrize("shape,po2", [(SMALL, 8), (SMALL, 9), (MID, 10), (SMALL, 13)])
def test_valid_oracle_seal_verifies(oracle, shape, po2):
    blob, seal = oracle_seal(oracle, shape, po2, seed=3 + po2)
    verify_segment(blob, seal)


def test_unsatisfied_constraints_are_rejected(oracle):
    blob, seal = oracle_seal(oracle, SMALL, 8, seed=4, valid=False)
    with pytest.raises(ZkbError, match="constraint polynomial"):
        verify_segment(blob, seal)


def test_every_region_of_the_seal_is_bound(oracle):
    blob, seal = oracle_seal(oracle, SMALL, 8, seed=5)
    verify_segment(blob, seal)
    rng = np.random.default_rng(1)
    # flip one word at positions spread over the whole seal: header, top layers, coeff_u, FRI commitments, final
    # coefficients and query answers -- every single-word change must be caught
    positions = sorted(set([0, 3, 4, 5, 40, seal.size - 1, seal.size - 9] + list(rng.integers(0, seal.size, 60))))
    for pos in positions:
        bad = seal.copy()
        bad[pos] = (int(bad[pos]) + 1) % 2013265921
        with pytest.raises(ZkbError, match="invalid proof|out of range"):
            verify_segment(blob, bad)
    with pytest.raises(ZkbError, match="too short"):
        verify_segment(blob, seal[:-1])
    with pytest.raises(ZkbError, match="trailing"):
        verify_segment(blob, np.concatenate([seal, np.zeros(1, np.uint32)]))
    # a seal for one circuit does not verify under another circuit's description
    other = circuit.syn_circuit(**dict(SMALL, data_cols=7)).blob()
    with pytest.raises(ZkbError):
        verify_segment(other, seal)


def test_golden_seal_verifies():
    import os
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
    blob = circuit.syn_circuit(**SMALL).blob()
    verify_segment(blob, G["seg_valid_seal"])
    with pytest.raises(ZkbError):
        verify_segment(blob, G["seg_random_seal"])


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
    verify_segment(blob, seal)
    bad = seal.copy(); bad[seal.size // 2] ^= 1
    with pytest.raises(ZkbError):
        verify_segment(blob, bad)
    gp.close(); hal.close()
