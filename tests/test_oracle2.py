"""The second, independent oracle (oracle/oracle2.py: pure numpy, canonical-value arithmetic, iterative DFTs, zero-pad LDE,
own Grain LFSR, transcript from SURVEY App. D) against (a) the spec's known answers, (b) the committed golden vectors and
(c) the C++ oracle on inputs the goldens do not hold -- so every item SURVEY 8(c) lists as "resting on recall" (sponge padding,
rng draw rule, Merkle top_size rule, fri_fold index formula, transcript order) is derived twice, by two code paths that share
nothing.  CPU only."""
import os

import numpy as np
import pytest

from oracle import oracle2 as O2
from zktls_b200 import circuit, synth

P = 2013265921
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))
SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)
MID = dict(accum_cols=7, code_cols=5, data_cols=33, mix_size=20, out_size=32)


def eq(a, b):
    return np.array_equal(np.asarray(a, dtype=np.uint32).ravel(), np.asarray(b, dtype=np.uint32).ravel())


def hexs(words):
    return " ".join(f"{int(v):08x}" for v in words)


def test_spec_known_answers():
    """permutation KAT, sponge padding cases, hash_pair, Merkle, rng draw rule, NTT conventions, fri_fold."""
    assert hexs(O2.poseidon2_mix(np.arange(24, dtype=np.uint64)[None, :])[0][[0, 1, 2, 23]]) == "2ed3e23d 12921fb0 0e659e79 57a99864"
    f2 = O2.hash_words(O2.enc(np.arange(40)))
    assert hexs(O2.dec(f2)) == "2759bb7a 774b8b4b 3a27f4c5 371d263a 18c62843 358cd9dc 236c1314 5905cec9"
    assert hexs(f2) == "440d5efd 2180f8fc 53ef06e6 026c7b41 33261083 05c27381 4c6ec63d 1a15be55"
    assert hexs(O2.dec(O2.hash_words(np.zeros(0, np.uint32)))) == "39fa4dee 386ee43e 45e695ae 24392948 560b05e1 009e435d 4ae29966 31a12e7b"
    assert hexs(O2.dec(O2.hash_words(O2.enc(np.arange(16))))) == "68445441 346b4f69 1dc5cbbc 6dd0663b 6f6c16ed 0442d646 0126152a 2929339c"
    f5 = O2.hash_pair_many(f2[None, :], f2[None, :])[0]
    assert hexs(O2.dec(f5)) == "40b4e587 5a8815c0 73bd632a 445486f7 032d913b 7339f5b5 3124512c 1c95a7c7"
    # F.6 Merkle
    m = O2.enc(np.arange(1, 25))
    nodes = O2.merkle_build(O2.hash_rows(m, 8, 3), 8)
    assert hexs(O2.dec(nodes[8])) == "0d3e41e5 410c0143 255f1f15 7022cd38 348f443e 1047ee89 656400e5 45a3691e"
    assert hexs(O2.dec(nodes[1])) == "32c79a75 0b85a1e5 161ac2ce 46fca593 3079ff96 6489d2ef 5affcfae 74816cc9"
    # F.7 rng
    r = O2.Rng(); r.mix(f2)
    assert [r.random_elem() for _ in range(3)] == [0x07BC76C7, 0x37AD3F28, 0x1EF5831B]
    assert r.random_bits(22) == 2793831
    assert hexs(r.random_ext_elem()) == "05e1cbe6 380294f1 6c99657f 525ad26f"
    r.mix(f5)
    assert r.random_elem() == 0x34FDEAE0
    # F.8 NTT
    ev = O2.enc(np.arange(1, 9))
    c = O2.batch_interpolate_ntt(ev, 1, 3)
    assert list(O2.dec(c)) == [1006632965, 1006632960, 1149063664, 864202256, 247018960, 37842447, 1975423473, 1766246960]
    cs = O2.zk_shift(c, 1, 3)
    assert list(O2.dec(cs)) == [1006632965, 1006632920, 275243371, 1864477272, 741056880, 1142650937, 991519825, 1338065042]
    lde = O2.dec(O2.batch_expand_into_evaluate_ntt(cs, 1, 3, 2))
    assert list(lde[:4]) == [313215528, 1279427784, 1093256194, 742674195] and int(lde[31]) == 510579356
    # F.9 / F.10
    assert list(O2.f4mul(O2.f4(1, 2, 3, 4), O2.f4(5, 6, 7, 8))) == [2013265255, 2013265365, 2013265603, 60]
    assert list(O2.f4inv(O2.f4(1, 2, 3, 4))) == [913204995, 645615856, 471318424, 1520759288]
    out = O2.dec(O2.fri_fold(O2.enc(np.arange(128)), O2.enc(np.array([2, 3, 5, 7])), 2))
    assert list(out) == [1462926330, 956898116, 745969380, 811640646, 524950586, 604349898, 1263579424, 54795267]


def test_merkle_top_size_rule():
    # D.2: top_layer = max{ i in 1..layers : 2^i <= 50 }
    assert O2.merkle_params(1 << 22) == (22, 5, 32)
    assert O2.merkle_params(1 << 6) == (6, 5, 32)
    assert O2.merkle_params(1 << 5) == (5, 4, 16)
    assert O2.merkle_params(2) == (1, 0, 1)


def test_oracle2_reproduces_operator_goldens():
    assert eq(O2.poseidon2_mix_words(G["mix_in"]), G["mix_out"])
    for k in ("hr_a", "hr_b", "hr_c", "hr_d"):
        rows, cols = (int(v) for v in G[k + "_shape"])
        assert eq(O2.hash_rows(G[k + "_in"], rows, cols), G[k + "_out"]), k
    assert eq(O2.merkle_build_words(G["merkle_in"], 64), G["merkle_out"])
    assert eq(O2.batch_interpolate_ntt(G["intt_in"], 3, 6), G["intt_out"])
    assert eq(O2.zk_shift(G["intt_out"], 3, 6), G["intt_shift_out"])
    assert eq(O2.batch_expand_into_evaluate_ntt(G["lde_in"], 2, 5, 2), G["lde_out"])
    assert eq(O2.batch_bit_reverse(G["brev_in"], 2, 7), G["brev_out"])
    assert eq(O2.batch_evaluate_any(G["any_coeffs"], 3, 8, G["any_which"], G["any_xs"]), G["any_out"])
    assert eq(O2.mix_poly_coeffs(G["mixc_out0"], G["mixc_start"], G["mixc_mix"], G["mixc_in"], G["mixc_combos"], 5, 100), G["mixc_out"])
    q, rem = O2.poly_divide_words(G["div_in"], G["div_z"])
    assert eq(q, G["div_out"]) and eq(rem, G["div_rem"])
    assert eq(O2.eltwise_sum_extelem(G["sum_in"], 50, 3), G["sum_out"])
    assert eq(O2.fri_fold(G["fold_in"], G["fold_mix"], 8), G["fold_out"])
    assert eq(O2.prefix_products(G["pp_in"]), G["pp_out"])
    assert eq(O2.eval_check(G["ec_blob"], G["ec_accum"], G["ec_code"], G["ec_data"], G["ec_mix"], G["ec_out_g"], G["ec_poly_mix"], 6), G["ec_check"])


@pytest.mark.parametrize("name", ["seg_valid", "seg_random"])
def test_oracle2_reproduces_segment_goldens(name):
    pr = O2.Prover(circuit.syn_circuit(**SMALL).blob())
    po2 = int(G[name + "_po2"][0])
    pr.begin(po2, G[name + "_io"], G[name + "_code"], G[name + "_data"])
    assert eq(pr.finish(G[name + "_accum"]), G[name + "_seal"])
    assert eq(pr.roots(), G[name + "_roots"])


def _deep_tap_circuit():
    """taps up to 5 rows back, three combos per group, nested AndCond, a constraint on globals only (as tests/test_prover_parity.py)"""
    from zktls_b200.circuit import CircuitBuilder, GROUP_ACCUM, GROUP_CODE, GROUP_DATA, GLOBAL_MIX, GLOBAL_OUT
    accum_cols, code_cols, data_cols, mix_size, out_size = 3, 4, 9, 3, 2
    b = CircuitBuilder(accum_cols, code_cols, data_cols, mix_size, out_size, info=b"DEEPTAPS:v1_____")
    backs = {GROUP_ACCUM: (0, 1, 3), GROUP_CODE: (0, 2), GROUP_DATA: (0, 1, 2, 5)}
    for g, n in ((GROUP_ACCUM, accum_cols), (GROUP_CODE, code_cols), (GROUP_DATA, data_cols)):
        for c in range(n):
            for k in backs[g]:
                b.add_tap(g, c, k)
    b.finish_taps()
    sel, gate = b.get(GROUP_CODE, 0, 0), b.get(GROUP_CODE, 1, 2)
    top = b.and_eqz(b.true(), b.mul(b.get_global(GLOBAL_MIX, 0), b.sub(b.get_global(GLOBAL_OUT, 1), b.const(7))))
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


def test_both_oracles_agree_beyond_the_goldens(oracle):
    """operators at other shapes, eval_check with nested AndCond / deep taps, and whole seals of circuits with 3-4 backs per register
    (poly_interpolate of size > 2, three divisions per combo) and of the MID shape."""
    O = oracle
    rng = np.random.default_rng(77)
    fp = lambda n: rng.integers(0, P, size=n, dtype=np.uint32)
    for po2, count, eb in ((1, 2, 2), (4, 3, 1), (9, 2, 2), (10, 1, 0)):
        x = fp(count << po2)
        assert eq(O.batch_interpolate_ntt(x, count, po2), O2.batch_interpolate_ntt(x, count, po2))
        assert eq(O.zk_shift(x, count, po2), O2.zk_shift(x, count, po2))
        assert eq(O.batch_expand_into_evaluate_ntt(x, count, po2, eb), O2.batch_expand_into_evaluate_ntt(x, count, po2, eb))
    for rows, cols in ((3, 15), (16, 17), (33, 48), (7, 224)):
        m = fp(rows * cols)
        assert eq(O.hash_rows(m, rows, cols), O2.hash_rows(m, rows, cols))
    b = _deep_tap_circuit(); blob = b.blob()
    for po2 in (6, 7):
        dom = 4 << po2
        accum, code, data = (fp(n * dom) for n in b.group_size)
        mix, out, pm = fp(b.mix_size), fp(b.out_size), fp(4)
        assert eq(O.eval_check(blob, accum, code, data, mix, out, pm, po2), O2.eval_check(blob, accum, code, data, mix, out, pm, po2))
    # whole seals (random traces: the transcript does not care whether the constraints hold)
    for blob_, shape, po2, seed in ((blob, dict(accum_cols=3, code_cols=4, data_cols=9, out_size=2), 8, 31), (circuit.syn_circuit(**MID).blob(), MID, 9, 32)):
        io, code_m, data_m, accum_m = synth.trace_a(shape, po2, seed)
        p1, p2 = O.Prover(blob_), O2.Prover(blob_)
        assert eq(p1.begin(po2, io, code_m, data_m), p2.begin(po2, io, code_m, data_m))
        s1, s2 = p1.finish(accum_m), p2.finish(accum_m)
        assert s1.size == s2.size and eq(s1, s2)
        assert eq(p1.roots(), p2.roots())


def test_both_oracles_agree_on_a_heavy_circuit(oracle):
    """nested AndCond four deep, five combos (registers of 1, 2, 3 and 5 taps): eval_check and the whole seal"""
    red = dict(accum_cols=6, code_cols=6, data_cols=12, mix_size=5, out_size=4, majors=3, fanout=(2, 2, 2), leaf_constraints=12)
    b = circuit.syn_heavy_circuit(**red); blob = b.blob()
    rng = np.random.default_rng(5)
    fp = lambda n: rng.integers(0, P, size=n, dtype=np.uint32)
    po2 = 6; dom = 4 << po2
    accum, code, data = (fp(n * dom) for n in b.group_size)
    mix, out, pm = fp(5), fp(4), fp(4)
    assert eq(oracle.eval_check(blob, accum, code, data, mix, out, pm, po2), O2.eval_check(blob, accum, code, data, mix, out, pm, po2))
    shape = dict(accum_cols=6, code_cols=6, data_cols=12, out_size=4)
    io, code_m, data_m, accum_m = synth.trace_a(shape, 8, 33)
    p1, p2 = oracle.Prover(blob), O2.Prover(blob)
    assert eq(p1.begin(8, io, code_m, data_m), p2.begin(8, io, code_m, data_m))
    assert eq(p1.finish(accum_m), p2.finish(accum_m)) and eq(p1.roots(), p2.roots())


def test_product_never_imports_oracle2():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for dirpath, _, files in os.walk(os.path.join(root, "zktls_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not re.search(r"oracle2", open(os.path.join(dirpath, f)).read()), f
    assert "oracle2" not in open(os.path.join(root, "bench.py")).read()
