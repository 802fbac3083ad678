"""Generates tests/golden/golden_v1.npz: small input/output vectors for every operator on the hot path and for the
whole segment prover, produced by the CPU oracle (oracle/, the restated CpuHal -- the reference itself is un-vendored
Rust and cannot be run).  Run once, commit the .npz:

    python tests/golden/make_golden.py

EVERY array is computed twice -- by the C++ oracle (oracle/*.hpp) and by the independent pure-numpy oracle
(oracle/oracle2.py: canonical-value arithmetic, iterative DFTs, zero-pad LDE, its own Grain LFSR, transcript written from
SURVEY App. D) -- and the script REFUSES to write the file unless the two agree word for word, both seals included.

tests/test_golden.py then checks (a) the oracle still reproduces these vectors (CPU) and (b) libzkb200 reproduces them
on the GPU without the oracle in the loop.  Inputs are stored too, so nothing depends on numpy's RNG stream.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O            # noqa: E402
from oracle import oracle2 as O2          # noqa: E402
from zktls_b200 import circuit, synth     # noqa: E402

P = 2013265921
SMALL = dict(accum_cols=4, code_cols=3, data_cols=6, mix_size=5, out_size=4)


def both(name, a, b):
    """the two oracles' results for one golden array; identical or no golden file"""
    a = np.asarray(a, dtype=np.uint32).ravel(); b = np.asarray(b, dtype=np.uint32).ravel()
    if a.shape != b.shape or not np.array_equal(a, b):
        raise SystemExit(f"make_golden: the C++ oracle and oracle2.py DISAGREE on {name}: not writing a golden file")
    return a


def seal_case(shape, po2, seed, valid):
    blob = circuit.syn_circuit(**shape).blob()
    pr = O.Prover(blob); pr2 = O2.Prover(blob)
    if valid:
        io, code, data = synth.trace_b_code_data(shape, po2, seed)
        code_m, data_m = synth.to_mont(code), synth.to_mont(data)
        mix = both("mix globals", pr.begin(po2, io, code_m, data_m), pr2.begin(po2, io, code_m, data_m))
        accum_m = synth.to_mont(synth.trace_b_accum(shape, po2, seed, code, data, io, mix))
    else:
        io, code_m, data_m, accum_m = synth.trace_a(shape, po2, seed)
        pr.begin(po2, io, code_m, data_m); pr2.begin(po2, io, code_m, data_m)
    seal = both("seal", pr.finish(accum_m), pr2.finish(accum_m))
    both("roots", pr.roots(), pr2.roots())
    return dict(io=io, code=code_m, data=data_m, accum=accum_m, seal=seal, roots=pr.roots().ravel(),
                seal_sha256=np.frombuffer(hashlib.sha256(seal.tobytes()).digest(), dtype=np.uint8))


def main():
    rng = np.random.default_rng(0xB200)
    g = {}
    fp = lambda *shape: rng.integers(0, P, size=shape, dtype=np.uint32)
    # Poseidon2
    g["mix_in"] = fp(24); g["mix_out"] = both("mix", O.poseidon2_mix(g["mix_in"]), O2.poseidon2_mix_words(g["mix_in"]))
    for name, (rows, cols) in {"hr_a": (37, 33), "hr_b": (64, 16), "hr_c": (5, 1), "hr_d": (9, 0)}.items():
        m = fp(rows * cols); g[name + "_in"] = m; g[name + "_shape"] = np.array([rows, cols]); g[name + "_out"] = both(name, O.hash_rows(m, rows, cols), O2.hash_rows(m, rows, cols))
    nodes = np.zeros(2 * 64 * 8, np.uint32); nodes[64 * 8:] = fp(64 * 8)
    g["merkle_in"] = nodes; g["merkle_out"] = both("merkle", O.merkle_build(nodes, 64), O2.merkle_build_words(nodes, 64))
    # NTT family
    x = fp(3 << 6); g["intt_in"] = x; g["intt_out"] = both("intt", O.batch_interpolate_ntt(x, 3, 6), O2.batch_interpolate_ntt(x, 3, 6))
    g["intt_shift_out"] = both("zk_shift", O.zk_shift(g["intt_out"], 3, 6), O2.zk_shift(g["intt_out"], 3, 6))
    x = fp(2 << 5); g["lde_in"] = x; g["lde_out"] = both("lde", O.batch_expand_into_evaluate_ntt(x, 2, 5, 2), O2.batch_expand_into_evaluate_ntt(x, 2, 5, 2))
    x = fp(2 << 7); g["brev_in"] = x; g["brev_out"] = both("brev", O.batch_bit_reverse(x, 2, 7), O2.batch_bit_reverse(x, 2, 7))
    # DEEP / mix / divide / sum / fold
    co = fp(3 << 8); which = np.array([0, 2, 2, 1], np.uint32); xs = fp(16)
    g["any_coeffs"] = co; g["any_which"] = which; g["any_xs"] = xs; g["any_out"] = both("any", O.batch_evaluate_any(co, 3, 8, which, xs), O2.batch_evaluate_any(co, 3, 8, which, xs))
    inp = fp(5 * 100); combos = np.array([0, 1, 1, 0, 2], np.uint32); out0 = fp(3 * 100 * 4); ms, mx = fp(4), fp(4)
    g["mixc_in"] = inp; g["mixc_combos"] = combos; g["mixc_out0"] = out0; g["mixc_start"] = ms; g["mixc_mix"] = mx
    g["mixc_out"] = both("mix_poly_coeffs", O.mix_poly_coeffs(out0, ms, mx, inp, combos, 5, 100), O2.mix_poly_coeffs(out0, ms, mx, inp, combos, 5, 100))
    p = fp(300 * 4); z = fp(4); q, rem = O.poly_divide(p, z); q2, rem2 = O2.poly_divide_words(p, z); both("div q", q, q2); both("div rem", rem, rem2)
    g["div_in"] = p; g["div_z"] = z; g["div_out"] = q; g["div_rem"] = rem
    s = fp(3 * 50 * 4); g["sum_in"] = s; g["sum_out"] = both("sum", O.eltwise_sum_extelem(s, 50, 3), O2.eltwise_sum_extelem(s, 50, 3))
    f = fp(64 * 8); fm = fp(4); g["fold_in"] = f; g["fold_mix"] = fm; g["fold_out"] = both("fri_fold", O.fri_fold(f, fm, 8), O2.fri_fold(f, fm, 8))
    pp = fp(77 * 4); g["pp_in"] = pp; g["pp_out"] = both("prefix_products", O.prefix_products(pp), O2.prefix_products(pp))
    # eval_check on a small circuit
    po2 = 6; n = 1 << po2; blob = circuit.syn_circuit(**SMALL).blob()
    acc, code, data = fp(4 * 4 * n), fp(3 * 4 * n), fp(6 * 4 * n); mg, og, pm = fp(5), fp(4), fp(4)
    g["ec_blob"] = blob; g["ec_accum"] = acc; g["ec_code"] = code; g["ec_data"] = data; g["ec_mix"] = mg; g["ec_out_g"] = og; g["ec_poly_mix"] = pm
    g["ec_check"] = both("eval_check", O.eval_check(blob, acc, code, data, mg, og, pm, po2), O2.eval_check(blob, acc, code, data, mg, og, pm, po2))
    # whole segment: a valid trace (seal verifies) and a random one
    for name, (po2, seed, valid) in {"seg_valid": (8, 5, True), "seg_random": (7, 9, False)}.items():
        for k, v in seal_case(SMALL, po2, seed, valid).items():
            g[f"{name}_{k}"] = np.asarray(v)
        g[f"{name}_po2"] = np.array([po2])
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(out, **g)
    print(out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
