"""Golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py from the CPU oracle).

CPU leg: the oracle still reproduces them (guards the checker itself).  GPU leg: libzkb200 reproduces them through the
C-ABI with NO oracle in the loop -- bit-exact, integer arithmetic only."""
import hashlib
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)
HR = ["hr_a", "hr_b", "hr_c", "hr_d"]


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint32).ravel(), np.asarray(b, dtype=np.uint32).ravel())


# ---- CPU: oracle vs golden ------------------------------------------------------------------------------------------
def test_oracle_reproduces_operator_goldens(oracle):
    O = oracle
    assert eq(O.poseidon2_mix(G["mix_in"]), G["mix_out"])
    for k in HR:
        rows, cols = (int(v) for v in G[k + "_shape"])
        assert eq(O.hash_rows(G[k + "_in"], rows, cols), G[k + "_out"]), k
    assert eq(O.merkle_build(G["merkle_in"], 64), G["merkle_out"])
    assert eq(O.batch_interpolate_ntt(G["intt_in"], 3, 6), G["intt_out"])
    assert eq(O.zk_shift(G["intt_out"], 3, 6), G["intt_shift_out"])
    assert eq(O.batch_expand_into_evaluate_ntt(G["lde_in"], 2, 5, 2), G["lde_out"])
    assert eq(O.batch_bit_reverse(G["brev_in"], 2, 7), G["brev_out"])
    assert eq(O.batch_evaluate_any(G["any_coeffs"], 3, 8, G["any_which"], G["any_xs"]), G["any_out"])
    assert eq(O.mix_poly_coeffs(G["mixc_out0"], G["mixc_start"], G["mixc_mix"], G["mixc_in"], G["mixc_combos"], 5, 100), G["mixc_out"])
    q, rem = O.poly_divide(G["div_in"], G["div_z"])
    assert eq(q, G["div_out"]) and eq(rem, G["div_rem"])
    assert eq(O.eltwise_sum_extelem(G["sum_in"], 50, 3), G["sum_out"])
    assert eq(O.fri_fold(G["fold_in"], G["fold_mix"], 8), G["fold_out"])
    assert eq(O.prefix_products(G["pp_in"]), G["pp_out"])
    assert eq(O.eval_check(G["ec_blob"], G["ec_accum"], G["ec_code"], G["ec_data"], G["ec_mix"], G["ec_out_g"], G["ec_poly_mix"], 6), G["ec_check"])


@pytest.mark.parametrize("name", ["seg_valid", "seg_random"])
def test_oracle_reproduces_segment_goldens(oracle, name):
    from zktls_b200 import circuit
    pr = oracle.Prover(circuit.syn_circuit(**SMALL).blob())
    po2 = int(G[name + "_po2"][0])
    pr.begin(po2, G[name + "_io"], G[name + "_code"], G[name + "_data"])
    seal = pr.finish(G[name + "_accum"])
    assert eq(seal, G[name + "_seal"])
    assert eq(pr.roots(), G[name + "_roots"])
    assert hashlib.sha256(seal.tobytes()).digest() == G[name + "_seal_sha256"].tobytes()


# ---- GPU: libzkb200 vs golden (no oracle) -----------------------------------------------------------------------------
@pytest.fixture(scope="module")
def hal():
    from zktls_b200.hal import B200Hal
    h = B200Hal(0)
    yield h
    h.close()


@pytest.mark.gpu
def test_gpu_reproduces_operator_goldens(hal):
    for k in HR:
        rows, cols = (int(v) for v in G[k + "_shape"])
        out = hal.alloc_digest(rows)
        m = hal.copy_from_elem(G[k + "_in"]) if cols else hal.alloc_elem(1)
        import ctypes as C
        from zktls_b200._lib import lib, check
        check(lib().zkb_poseidon2_hash_rows(hal.ctx, C.c_void_p(out.ptr), C.c_void_p(m.ptr), C.c_size_t(rows), C.c_size_t(cols)))
        assert eq(out.to_numpy(), G[k + "_out"]), k
    nodes = hal.copy_from_digest(G["merkle_in"]); hal.merkle_build(nodes, 64)
    assert eq(nodes.to_numpy(), G["merkle_out"])
    b = hal.copy_from_elem(G["intt_in"]); hal.batch_interpolate_ntt(b, 3); assert eq(b.to_numpy(), G["intt_out"])
    hal.zk_shift(b, 3); assert eq(b.to_numpy(), G["intt_shift_out"])
    b = hal.copy_from_elem(G["intt_in"]); hal.batch_interpolate_ntt_zk_shift(b, 3); assert eq(b.to_numpy(), G["intt_shift_out"])
    o = hal.alloc_elem(G["lde_out"].size); hal.batch_expand_into_evaluate_ntt(o, hal.copy_from_elem(G["lde_in"]), 2, 2)
    assert eq(o.to_numpy(), G["lde_out"])
    b = hal.copy_from_elem(G["brev_in"]); hal.batch_bit_reverse(b, 2); assert eq(b.to_numpy(), G["brev_out"])
    o = hal.alloc_extelem(4)
    hal.batch_evaluate_any(hal.copy_from_elem(G["any_coeffs"]), 3, hal.copy_from_u32(G["any_which"]), hal.copy_from_extelem(G["any_xs"]), o)
    assert eq(o.to_numpy(), G["any_out"])
    o = hal.copy_from_extelem(G["mixc_out0"])
    hal.mix_poly_coeffs(o, G["mixc_start"], G["mixc_mix"], hal.copy_from_elem(G["mixc_in"]), hal.copy_from_u32(G["mixc_combos"]), 5, 100)
    assert eq(o.to_numpy(), G["mixc_out"])
    p = hal.copy_from_extelem(G["div_in"]); rem = hal.poly_divide(p, G["div_z"])
    assert eq(p.to_numpy(), G["div_out"]) and eq(rem, G["div_rem"])
    o = hal.alloc_elem(200); hal.eltwise_sum_extelem(o, hal.copy_from_extelem(G["sum_in"])); assert eq(o.to_numpy(), G["sum_out"])
    o = hal.alloc_elem(32); hal.fri_fold(o, hal.copy_from_elem(G["fold_in"]), G["fold_mix"]); assert eq(o.to_numpy(), G["fold_out"])
    b = hal.copy_from_extelem(G["pp_in"]); hal.prefix_products(b); assert eq(b.to_numpy(), G["pp_out"])
    chk = hal.alloc_elem(16 << 6)
    hal.eval_check(chk, G["ec_blob"], hal.copy_from_elem(G["ec_accum"]), hal.copy_from_elem(G["ec_code"]), hal.copy_from_elem(G["ec_data"]),
                   G["ec_mix"], G["ec_out_g"], G["ec_poly_mix"], 6)
    assert eq(chk.to_numpy(), G["ec_check"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["seg_valid", "seg_random"])
def test_gpu_reproduces_segment_goldens(hal, name):
    from zktls_b200 import circuit
    from zktls_b200.prover import SegmentProver
    pr = SegmentProver(hal, circuit.syn_circuit(**SMALL).blob())
    po2 = int(G[name + "_po2"][0])
    pr.begin(po2, G[name + "_io"], G[name + "_code"], G[name + "_data"])
    seal = pr.finish(G[name + "_accum"])
    assert eq(seal, G[name + "_seal"])
    assert eq(pr.roots(), G[name + "_roots"])
    pr.close()
