"""GPU parity: every Hal operator of libzkb200 (through the C-ABI) against the CPU oracle, bit-exact.

Mirrors risc0-zkp's `hal::testutil` differential tests: random field elements from fixed seeds,
a handful of sizes crossing tile boundaries, whole-buffer equality.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 2013265921


@pytest.fixture(scope="module")
def hal():
    from zktls_b200.hal import B200Hal
    h = B200Hal(0)
    yield h
    h.close()


def rnd(seed, n):
    return np.random.default_rng(seed).integers(0, P, size=n, dtype=np.uint32)


NTT_SHAPES = [(0, 3), (1, 1), (3, 1), (3, 5), (5, 2), (8, 3), (10, 7), (11, 2), (12, 3), (13, 2), (14, 5), (16, 3), (18, 2), (20, 1)]


@pytest.mark.parametrize("po2,count", NTT_SHAPES)
def test_batch_interpolate_ntt(hal, oracle, po2, count):
    x = rnd(100 + po2, count << po2)
    b = hal.copy_from_elem(x)
    hal.batch_interpolate_ntt(b, count)
    assert np.array_equal(b.to_numpy(), oracle.batch_interpolate_ntt(x, count, po2))


@pytest.mark.parametrize("po2,count", NTT_SHAPES)
def test_interpolate_zk_shift(hal, oracle, po2, count):
    x = rnd(200 + po2, count << po2)
    want = oracle.zk_shift(oracle.batch_interpolate_ntt(x, count, po2), count, po2)
    b = hal.copy_from_elem(x)
    hal.batch_interpolate_ntt_zk_shift(b, count)
    assert np.array_equal(b.to_numpy(), want)
    b2 = hal.copy_from_elem(x)
    hal.batch_interpolate_ntt(b2, count)
    hal.zk_shift(b2, count)
    assert np.array_equal(b2.to_numpy(), want)


@pytest.mark.parametrize("po2,count", [(0, 2), (1, 3), (3, 1), (6, 4), (9, 3), (10, 2), (11, 3), (12, 1), (14, 3), (16, 2), (18, 1)])
@pytest.mark.parametrize("eb", [2, 0, 1])
def test_batch_expand_into_evaluate_ntt(hal, oracle, po2, count, eb):
    x = rnd(300 + po2, count << po2)
    out = hal.alloc_elem(count << (po2 + eb))
    hal.batch_expand_into_evaluate_ntt(out, hal.copy_from_elem(x), count, eb)
    assert np.array_equal(out.to_numpy(), oracle.batch_expand_into_evaluate_ntt(x, count, po2, eb))


def test_split_expand_then_evaluate(hal, oracle):
    po2, count, eb = 9, 3, 2
    x = rnd(7, count << po2)
    out = hal.alloc_elem(count << (po2 + eb))
    hal.batch_expand(out, hal.copy_from_elem(x), count)
    assert np.array_equal(out.to_numpy(), oracle.batch_expand(x, count, po2, eb))
    hal.batch_evaluate_ntt(out, count, eb)
    assert np.array_equal(out.to_numpy(), oracle.batch_expand_into_evaluate_ntt(x, count, po2, eb))


def test_ntt_roundtrip_large(hal):
    """size-independent property at a size the oracle would take long on: evaluate(interpolate(x)) == x."""
    po2, count = 22, 4
    x = rnd(8, count << po2)
    b = hal.copy_from_elem(x)
    hal.batch_interpolate_ntt(b, count)
    out = hal.alloc_elem(count << po2)
    hal.batch_expand_into_evaluate_ntt(out, b, count, 0)
    assert np.array_equal(out.to_numpy(), x)


@pytest.mark.parametrize("po2,count", [(0, 1), (1, 2), (4, 3), (10, 5), (15, 2)])
def test_batch_bit_reverse(hal, oracle, po2, count):
    x = rnd(400 + po2, count << po2)
    b = hal.copy_from_elem(x)
    hal.batch_bit_reverse(b, count)
    assert np.array_equal(b.to_numpy(), oracle.batch_bit_reverse(x, count, po2))


@pytest.mark.parametrize("rows,cols", [(1, 1), (8, 3), (32, 16), (100, 17), (256, 0), (1000, 40), (4096, 64), (1 << 14, 224), (777, 33), (1 << 16, 16)])
def test_hash_rows(hal, oracle, rows, cols):
    m = rnd(500 + rows + cols, rows * cols)
    out = hal.alloc_digest(rows)
    hal.hash_rows(out, hal.copy_from_elem(m) if cols else hal.alloc_elem(0))
    assert np.array_equal(out.to_numpy(), oracle.hash_rows(m, rows, cols) if cols else np.tile(oracle.hash_elem_slice(np.zeros(0, np.uint32)), rows))


@pytest.mark.parametrize("rows", [1, 2, 8, 64, 2048, 4096, 1 << 15])
def test_merkle_build_and_fold(hal, oracle, rows):
    nodes = np.zeros(2 * rows * 8, np.uint32)
    nodes[rows * 8:] = rnd(600 + rows, rows * 8)
    want = oracle.merkle_build(nodes, rows)
    b = hal.copy_from_digest(nodes)
    hal.merkle_build(b, rows)
    assert np.array_equal(b.to_numpy()[8:], want[8:])
    # level-at-a-time hash_fold gives the same tree
    b2 = hal.copy_from_digest(nodes)
    size = rows
    while size > 1:
        hal.hash_fold(b2, size, size // 2)
        size //= 2
    assert np.array_equal(b2.to_numpy()[8:], want[8:])


def test_merkle_kat(hal, oracle):
    """SURVEY App. F.6 through the GPU path."""
    rows, cols = 8, 3
    nodes = hal.alloc_digest(2 * rows)
    hal.hash_rows(nodes.slice(rows, rows), hal.copy_from_elem(oracle.encode(np.arange(rows * cols) + 1)))
    hal.merkle_build(nodes, rows)
    root = oracle.decode(nodes.to_numpy()[8:16])
    assert [f"{int(x):08x}" for x in root] == "32c79a75 0b85a1e5 161ac2ce 46fca593 3079ff96 6489d2ef 5affcfae 74816cc9".split()


@pytest.mark.parametrize("po2,polys,n_eval", [(0, 2, 3), (3, 2, 4), (8, 5, 9), (14, 3, 5), (15, 2, 3), (17, 3, 4)])
def test_batch_evaluate_any(hal, oracle, po2, polys, n_eval):
    rng = np.random.default_rng(700 + po2)
    coeffs = rnd(701 + po2, polys << po2)
    which = rng.integers(0, polys, size=n_eval, dtype=np.uint32)
    xs = rnd(702 + po2, 4 * n_eval)
    out = hal.alloc_extelem(n_eval)
    hal.batch_evaluate_any(hal.copy_from_elem(coeffs), polys, hal.copy_from_u32(which), hal.copy_from_extelem(xs), out)
    assert np.array_equal(out.to_numpy(), oracle.batch_evaluate_any(coeffs, polys, po2, which, xs))


@pytest.mark.parametrize("count,input_size,n_combos", [(1, 1, 1), (16, 5, 2), (1000, 40, 3), (1 << 14, 300, 3), (4097, 17, 4)])
def test_mix_poly_coeffs(hal, oracle, count, input_size, n_combos):
    rng = np.random.default_rng(800 + count)
    inp = rnd(801 + count, input_size * count)
    combos = rng.integers(0, n_combos, size=input_size, dtype=np.uint32)
    out0 = rnd(802 + count, 4 * n_combos * count)
    ms, mx = rnd(803, 4), rnd(804, 4)
    out = hal.copy_from_extelem(out0)
    hal.mix_poly_coeffs(out, ms, mx, hal.copy_from_elem(inp), hal.copy_from_u32(combos), input_size, count)
    assert np.array_equal(out.to_numpy(), oracle.mix_poly_coeffs(out0, ms, mx, inp, combos, input_size, count))


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 256, 257, 4096, 5000, 1 << 16, (1 << 16) + 3])
def test_poly_divide(hal, oracle, n):
    p = rnd(900 + n, 4 * n); z = rnd(901 + n, 4)
    b = hal.copy_from_extelem(p)
    rem = hal.poly_divide(b, z)
    q, r = oracle.poly_divide(p, z)
    assert np.array_equal(b.to_numpy(), q) and np.array_equal(rem, r)


@pytest.mark.parametrize("n", [1, 700, 1 << 13])
def test_combos_divide(hal, oracle, n):
    """the division step of Prover::finalize in one call: a combo is divided once per tap offset it holds, in order"""
    combos = rnd(950 + n, 3 * 4 * n).reshape(3, 4 * n)
    order, pts = [0, 2, 2, 1, 2], rnd(951 + n, 5 * 4).reshape(5, 4)
    b = hal.copy_from_extelem(combos.reshape(-1))
    rems = hal.combos_divide(b, n, order, pts)
    want = [c.copy() for c in combos]
    for k, c in enumerate(order):
        want[c], r = oracle.poly_divide(want[c], pts[k])
        assert np.array_equal(rems[k], r)
    assert np.array_equal(b.to_numpy(), np.concatenate(want))
    with pytest.raises(Exception, match="out of range"):
        hal.combos_divide(b, n, [3], pts[:1])


@pytest.mark.parametrize("count,to_add", [(1, 1), (100, 3), (1 << 12, 4)])
def test_eltwise_sum_extelem(hal, oracle, count, to_add):
    inp = rnd(1000 + count, 4 * count * to_add)
    out = hal.alloc_elem(4 * count)
    hal.eltwise_sum_extelem(out, hal.copy_from_extelem(inp))
    assert np.array_equal(out.to_numpy(), oracle.eltwise_sum_extelem(inp, count, to_add))


@pytest.mark.parametrize("m", [1, 2, 16, 1000, 1 << 14])
def test_fri_fold(hal, oracle, m):
    inp = rnd(1100 + m, 64 * m); mix = rnd(1101, 4)
    out = hal.alloc_elem(4 * m)
    hal.fri_fold(out, hal.copy_from_elem(inp), mix)
    assert np.array_equal(out.to_numpy(), oracle.fri_fold(inp, mix, m))


def test_fri_fold_kat(hal, oracle):
    out = hal.alloc_elem(8)
    hal.fri_fold(out, hal.copy_from_elem(oracle.encode(np.arange(128))), oracle.encode([2, 3, 5, 7]))
    assert list(oracle.decode(out.to_numpy())) == [1462926330, 956898116, 745969380, 811640646, 524950586, 604349898, 1263579424, 54795267]


def test_eltwise_misc(hal, oracle):
    n = 10007
    a, b = rnd(1200, n), rnd(1201, n)
    out = hal.alloc_elem(n)
    hal.eltwise_add_elem(out, hal.copy_from_elem(a), hal.copy_from_elem(b))
    assert np.array_equal(out.to_numpy(), oracle.eltwise_add_elem(a, b))
    hal.eltwise_copy_elem(out, hal.copy_from_elem(a))
    assert np.array_equal(out.to_numpy(), a)
    z = a.copy(); z[::7] = 0xffffffff
    bz = hal.copy_from_elem(z)
    hal.eltwise_zeroize_elem(bz)
    assert np.array_equal(bz.to_numpy(), oracle.eltwise_zeroize_elem(z))
    g = hal.alloc_elem(50)
    hal.gather_sample(g, hal.copy_from_elem(a), 13, 50, 199)
    assert np.array_equal(g.to_numpy(), oracle.gather_sample(a, 13, 50, 199))
    # batched form: one launch for all queries of a tree
    idx = np.array([0, 13, 198, 57, 13], dtype=np.uint32)
    gr = hal.alloc_elem(idx.size * 50)
    hal.gather_rows(gr, hal.copy_from_elem(a), idx, 50, 199)
    assert np.array_equal(gr.to_numpy(), np.concatenate([oracle.gather_sample(a, int(i), 50, 199) for i in idx]))
    with pytest.raises(Exception, match="outside the source"):
        hal.gather_rows(gr, hal.copy_from_elem(a), np.array([a.size - 10], dtype=np.uint32), 50, 199)


@pytest.mark.parametrize("n", [1, 5, 64, 257, 5000, 1 << 15])
def test_prefix_products(hal, oracle, n):
    x = rnd(1300 + n, 4 * n)
    b = hal.copy_from_extelem(x)
    hal.prefix_products(b)
    assert np.array_equal(b.to_numpy(), oracle.prefix_products(x))


def test_errors_are_strings_not_crashes(hal):
    from zktls_b200.hal import ZkbError
    b = hal.alloc_elem(16)
    with pytest.raises(ZkbError):
        hal.hash_fold(hal.alloc_digest(8), 5, 2)
    with pytest.raises(ZkbError):
        hal.batch_interpolate_ntt(hal.alloc_elem(24), 2)      # 12 is not a power of two
    from zktls_b200._lib import lib, check
    import ctypes as C
    with pytest.raises(ZkbError):
        check(lib().zkb_batch_interpolate_ntt(hal.ctx, C.c_void_p(b.ptr), C.c_size_t(1), C.c_int(40)))


@pytest.mark.parametrize("po2,count,shift", [(21, 2, True), (22, 3, False), (23, 2, True), (24, 1, True), (25, 1, False)])
def test_tiled_inverse_ntt_matches_level_path_at_large_sizes(hal, po2, count, shift, monkeypatch):
    """Beyond what the CPU oracle finishes in seconds: the tiled kernels against the level-at-a-time device path (which
    the oracle pins at small sizes), including the three-pass plans (po2 > 22)."""
    x = rnd(1400 + po2, count << po2)
    a = hal.copy_from_elem(x)
    (hal.batch_interpolate_ntt_zk_shift if shift else hal.batch_interpolate_ntt)(a, count)
    monkeypatch.setenv("ZKB_NTT_FORCE_LEVELS", "1")
    b = hal.copy_from_elem(x)
    (hal.batch_interpolate_ntt_zk_shift if shift else hal.batch_interpolate_ntt)(b, count)
    monkeypatch.delenv("ZKB_NTT_FORCE_LEVELS")
    assert np.array_equal(a.to_numpy(), b.to_numpy())


@pytest.mark.parametrize("po2,count,eb", [(19, 3, 2), (20, 2, 2), (21, 1, 2), (22, 1, 2), (23, 1, 0), (24, 1, 2)])
def test_tiled_forward_ntt_matches_level_path_at_large_sizes(hal, po2, count, eb, monkeypatch):
    x = rnd(1500 + po2, count << po2)
    src = hal.copy_from_elem(x)
    a = hal.alloc_elem(count << (po2 + eb))
    hal.batch_expand_into_evaluate_ntt(a, src, count, eb)
    monkeypatch.setenv("ZKB_NTT_FORCE_LEVELS", "1")
    b = hal.alloc_elem(count << (po2 + eb))
    hal.batch_expand_into_evaluate_ntt(b, src, count, eb)
    monkeypatch.delenv("ZKB_NTT_FORCE_LEVELS")
    assert np.array_equal(a.to_numpy(), b.to_numpy())


def test_ntt_many_columns_crosses_l2_batches(hal, oracle):
    po2, count = 14, 300       # > one L2 batch at the default budget? no -- exercise the batching loop with a tiny budget instead
    x = rnd(1600, count << po2)
    b = hal.copy_from_elem(x)
    hal.batch_interpolate_ntt_zk_shift(b, count)
    want = oracle.zk_shift(oracle.batch_interpolate_ntt(x, count, po2), count, po2)
    assert np.array_equal(b.to_numpy(), want)


@pytest.mark.parametrize("rows,per_row", [(1, 1), (7, 3), (1000, 5)])
def test_scatter(hal, oracle, rows, per_row):
    rng = np.random.default_rng(rows)
    counts = rng.integers(0, per_row + 1, size=rows)
    index = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
    total = int(index[-1]); size = max(2 * total, 8)
    offsets = rng.permutation(size)[:total].astype(np.uint32)        # distinct targets: order-independent result
    values = oracle.random_fp(rng, total)
    into0 = oracle.random_fp(rng, size)
    buf = hal.copy_from_elem(into0)
    hal.scatter(buf, index, offsets, values)
    assert np.array_equal(buf.to_numpy(), oracle.scatter(into0, index, offsets, values))
    if total:
        from zktls_b200 import ZkbError
        bad = offsets.copy(); bad[0] = size
        with pytest.raises(ZkbError, match="outside"):
            hal.scatter(buf, index, bad, values)
