"""Pins the CPU oracle against every known-answer vector available for this path.

The reference tree's own tests pin nothing at the Hal boundary: the vectors here
are (i) the upstream risc0 KATs recalled in SURVEY.md App. A/B (ROU tables, Poseidon2 constants pin,
`poseidon2_test_vectors`) and (ii) the spec-derived App. F values, which were computed by an
independent Python evaluation of App. A-C during the survey.
"""
import hashlib
import numpy as np
import pytest

P = 2013265921


def hx(s):
    return np.array([int(x, 16) for x in s.split()], dtype=np.uint32)


ROU_FWD = [1, 2013265920, 284861408, 1801542727, 567209306, 740045640, 918899846, 1881002012, 1453957774, 65325759,
           1538055801, 515192888, 483885487, 157393079, 1695124103, 2005211659, 1540072241, 88064245, 1542985445,
           1269900459, 1461624142, 825701067, 682402162, 1311873874, 1164520853, 352275361, 18769, 137]
ROU_REV = [1, 2013265920, 1728404513, 1592366214, 196396260, 1253260071, 72041623, 1091445674, 145223211, 1446820157,
           1030796471, 2010749425, 1827366325, 1239938613, 246299276, 596347512, 1893145354, 246074437, 1525739923,
           1194341128, 1463599021, 704606912, 95395244, 15672543, 647517488, 584175179, 137728885, 749463956]

P2_KAT = hx("2ed3e23d 12921fb0 0e659e79 61d81dc9 32bae33b 62486ae3 1e681b60 24b91325 2a2ef5b9 50e8593e 5bc818ec 10691997 "
            "35a14520 2ba6a3c5 279d47ec 55014e81 5953a67f 2f403111 6b8828ff 1801301f 2749207a 3dc9cf21 3c985ba2 57a99864")
F2 = hx("2759bb7a 774b8b4b 3a27f4c5 371d263a 18c62843 358cd9dc 236c1314 5905cec9")
F2_MONT = hx("440d5efd 2180f8fc 53ef06e6 026c7b41 33261083 05c27381 4c6ec63d 1a15be55")
F3 = hx("39fa4dee 386ee43e 45e695ae 24392948 560b05e1 009e435d 4ae29966 31a12e7b")
F4 = hx("68445441 346b4f69 1dc5cbbc 6dd0663b 6f6c16ed 0442d646 0126152a 2929339c")
F5 = hx("40b4e587 5a8815c0 73bd632a 445486f7 032d913b 7339f5b5 3124512c 1c95a7c7")
F6_LEAF0 = hx("0d3e41e5 410c0143 255f1f15 7022cd38 348f443e 1047ee89 656400e5 45a3691e")
F6_ROOT = hx("32c79a75 0b85a1e5 161ac2ce 46fca593 3079ff96 6489d2ef 5affcfae 74816cc9")


def test_montgomery_encodings(oracle):
    assert list(oracle.encode([0, 1, 2, 3, P - 1])) == list(hx("00000000 0ffffffe 1ffffffc 2ffffffa 68000003"))
    L = oracle.lib()
    for x in [0, 1, 2, 3, P - 1, 12345678]:
        assert L.orc_fp_encode(x) == int(oracle.encode(x))
        assert L.orc_fp_decode(L.orc_fp_encode(x)) == x


def test_field_axioms(oracle):
    L = oracle.lib()
    rng = np.random.default_rng(1)
    for a, b in rng.integers(0, P, size=(500, 2)):
        a, b = int(a), int(b)
        ea, eb = L.orc_fp_encode(a), L.orc_fp_encode(b)
        assert L.orc_fp_decode(L.orc_fp_mul(ea, eb)) == a * b % P
        assert L.orc_fp_decode(L.orc_fp_add(ea, eb)) == (a + b) % P
        assert L.orc_fp_decode(L.orc_fp_sub(ea, eb)) == (a - b) % P
        if a:
            assert L.orc_fp_decode(L.orc_fp_mul(ea, L.orc_fp_inv(ea))) == 1
    assert L.orc_fp_inv(0) == 0


def test_rou_tables(oracle):
    f, r = oracle.rou_tables()
    assert list(f) == ROU_FWD and list(r) == ROU_REV
    for i in range(28):
        assert pow(ROU_FWD[i], 1 << i, P) == 1 and (i == 0 or pow(ROU_FWD[i], 1 << (i - 1), P) != 1)
    assert pow(3, 1 << 27, P) != 1


def test_fp4(oracle):
    got = oracle.decode(oracle.fp4_mul(oracle.encode([1, 2, 3, 4]), oracle.encode([5, 6, 7, 8])))
    assert list(got) == [2013265255, 2013265365, 2013265603, 60]
    inv = oracle.decode(oracle.fp4_inv(oracle.encode([1, 2, 3, 4])))
    assert list(inv) == [913204995, 645615856, 471318424, 1520759288]
    rng = np.random.default_rng(2)
    one = oracle.encode([1, 0, 0, 0])
    for _ in range(200):
        a = oracle.random_fp(rng, 4)
        assert list(oracle.fp4_mul(a, oracle.fp4_inv(a))) == list(one)


def test_poseidon2_constants(oracle):
    rc = oracle.poseidon2_round_constants()
    assert [int(x) for x in rc[:4]] == [0x0fa20c37, 0x0795bb97, 0x12c60b9c, 0x0eabd88e]
    assert int(rc[96]) == 0x1da78ec2 and int(rc[116]) == 0x6beb839d and int(rc[117]) == 0x032959ad and int(rc[212]) == 0x5244e9d4
    assert hashlib.sha256(rc.astype("<u4").tobytes()).hexdigest() == "9f7aa102258e5f0e2fbfcb1a50b141bcb1914a08385a4c5640e1ae4dde1da983"


def test_poseidon2_permutation_kat(oracle):
    out = oracle.decode(oracle.poseidon2_mix(oracle.encode(np.arange(24))))
    assert list(out) == list(P2_KAT)


def test_sponge_kats(oracle):
    d = oracle.hash_elem_slice(oracle.encode(np.arange(40)))
    assert list(d) == list(F2_MONT) and list(oracle.decode(d)) == list(F2)
    assert list(oracle.decode(oracle.hash_elem_slice(np.zeros(0, np.uint32)))) == list(F3)
    assert list(oracle.decode(oracle.hash_elem_slice(oracle.encode(np.arange(16))))) == list(F4)
    assert list(oracle.decode(oracle.hash_pair(F2_MONT, F2_MONT))) == list(F5)


def test_merkle_kat(oracle):
    rows, cols = 8, 3
    m = oracle.encode(np.arange(rows * cols) + 1)
    nodes = np.zeros(2 * rows * 8, np.uint32)
    nodes[rows * 8:] = oracle.hash_rows(m, rows, cols)
    assert list(oracle.decode(nodes[rows * 8: rows * 8 + 8])) == list(F6_LEAF0)
    nodes = oracle.merkle_build(nodes, rows)
    assert list(oracle.decode(nodes[8:16])) == list(F6_ROOT)
    # per-level hash_fold == merkle_build
    n2 = np.zeros(2 * rows * 8, np.uint32); n2[rows * 8:] = nodes[rows * 8:]
    for l in (2, 1, 0):
        n2 = oracle.hash_fold(n2, 2 << l, 1 << l)
    assert np.array_equal(n2[8:], nodes[8:])


def test_rng_kat(oracle):
    r = oracle.Rng()
    r.mix(F2_MONT)
    dec = lambda w: int(oracle.decode(w))
    assert [dec(r.random_elem()) for _ in range(3)] == [0x07bc76c7, 0x37ad3f28, 0x1ef5831b]
    assert r.random_bits(22) == 2793831
    assert [dec(r.random_elem()) for _ in range(4)] == [0x05e1cbe6, 0x380294f1, 0x6c99657f, 0x525ad26f]
    r.mix(oracle.encode(F5))
    assert dec(r.random_elem()) == 0x34fdeae0


def test_ntt_kat(oracle):
    ev = oracle.encode(np.arange(1, 9))
    co = oracle.batch_interpolate_ntt(ev, 1, 3)
    assert list(oracle.decode(co)) == [1006632965, 1006632960, 1149063664, 864202256, 247018960, 37842447, 1975423473, 1766246960]
    sh = oracle.zk_shift(co, 1, 3)
    assert list(oracle.decode(sh)) == [1006632965, 1006632920, 275243371, 1864477272, 741056880, 1142650937, 991519825, 1338065042]
    lde = oracle.batch_expand_into_evaluate_ntt(sh, 1, 3, 2)
    assert list(oracle.decode(lde)) == [
        313215528, 1279427784, 1093256194, 742674195, 738823641, 855455262, 1074788723, 738184687, 1658149204, 1163040150,
        1198233669, 435928816, 1952958124, 847195501, 968122175, 1149975769, 1952959765, 297292304, 1515036166, 1767882799,
        1021531450, 1606666157, 502269526, 1362503268, 102207201, 1360408393, 377573918, 1345334830, 313218807, 643578169,
        1323783349, 510579356]
    # split API == fused API
    split = oracle.batch_evaluate_ntt(oracle.batch_expand(sh, 1, 3, 2), 1, 5, 2)
    assert np.array_equal(split, lde)


def test_ntt_matches_direct_evaluation(oracle):
    """LDE output i == f(3 * w_4n^i) for the interpolant f of the trace column (App. C.3)."""
    rng = np.random.default_rng(3)
    k = 5; n = 1 << k
    vals = rng.integers(0, P, size=n)
    co = oracle.decode(oracle.batch_bit_reverse(oracle.batch_interpolate_ntt(oracle.encode(vals), 1, k), 1, k))
    w = ROU_FWD[k]
    for i in (0, 1, 7, n - 1):     # interpolant reproduces the trace on <w_n>
        assert sum(int(c) * pow(w, i * j, P) for j, c in enumerate(co)) % P == int(vals[i])
    lde = oracle.decode(oracle.batch_expand_into_evaluate_ntt(oracle.zk_shift(oracle.batch_interpolate_ntt(oracle.encode(vals), 1, k), 1, k), 1, k, 2))
    w4 = ROU_FWD[k + 2]
    for i in (0, 1, 2, 3, 50, 4 * n - 1):
        x = 3 * pow(w4, i, P) % P
        assert sum(int(c) * pow(x, j, P) for j, c in enumerate(co)) % P == int(lde[i])


def test_fri_fold_kat(oracle):
    out = oracle.fri_fold(oracle.encode(np.arange(128)), oracle.encode([2, 3, 5, 7]), 2)
    assert list(oracle.decode(out)) == [1462926330, 956898116, 745969380, 811640646, 524950586, 604349898, 1263579424, 54795267]


def test_poly_divide_and_eval(oracle):
    rng = np.random.default_rng(4)
    n = 33
    p = oracle.random_fp(rng, 4 * n); z = oracle.random_fp(rng, 4)
    q, rem = oracle.poly_divide(p, z)
    # p(x) == q(x) (x - z) + rem, checked at a random point
    x = oracle.random_fp(rng, 4)
    # direct Horner in python ints via fp4_mul
    def horner(poly):
        acc = np.zeros(4, np.uint32)
        for i in reversed(range(len(poly) // 4)):
            acc = oracle.fp4_mul(acc, x)
            acc = np.array([oracle.lib().orc_fp_add(int(a), int(b)) for a, b in zip(acc, poly[4 * i: 4 * i + 4])], np.uint32)
        return acc
    lhs = horner(p)
    xm = np.array([oracle.lib().orc_fp_sub(int(a), int(b)) for a, b in zip(x, z)], np.uint32)
    rhs = oracle.fp4_mul(horner(q), xm)
    rhs = np.array([oracle.lib().orc_fp_add(int(a), int(b)) for a, b in zip(rhs, rem)], np.uint32)
    assert np.array_equal(lhs, rhs)
